import json, sys, signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value","ms_per_step","gpu_launches_per_step","clocks")})
print("e2e", d["e2e"]); print("cpu", d.get("cpu_baseline"))
print("roofline", d["roofline"])
for k,v in d["kernels"].items(): print(f"{k:32s} calls {v['calls_per_step']:5.1f} ms/step {v['ms_per_step']:7.3f} share {v['share_of_step']:.3f} GB/s {v['achieved_GBps']:8.1f} frac {v['frac_of_hbm_peak']:.3f}")
