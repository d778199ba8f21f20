"""Where a cfg3 training step (PTv1-Seg50, ScanNet-shaped scenes, fwd + bwd + SGD, f32) spends its time on one GPU:
torch.profiler kernel table + CPU-side launch time, for the whole 8-scene batch and for one scene (an 8-GPU rank's share).
    python tools/train_profile.py [scenes]"""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from pointcloudpdf_b200 import synthetic as S
from pointcloudpdf_b200.ptv1 import PointTransformerSeg50
import pointcloudpdf_b200.pointops as pointops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(2027)
sizes = [int(x) for x in torch.randint(90000, 100001, (8,), generator=g)]
torch.manual_seed(2024)
net = PointTransformerSeg50(in_channels=9, num_classes=20).to(dev).train()
opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)

for ns in ([int(a) for a in sys.argv[1:]] or [8, 1]):
    b = S.scannet_batch(sizes[:ns], seed=2027)
    d = {k: b[k].to(dev) for k in ("coord", "feat", "offset")}
    label = torch.randint(0, 20, (d["coord"].shape[0],), device=dev)
    off_host = b["offset"].tolist()

    def step():
        pointops.clear_caches()
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(net(d, off_host), label).backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts, cs = [], []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); step(); t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); cs.append((t1 - t0) * 1e3)
    print(f"== {ns} scene(s), {d['coord'].shape[0]} points: step {statistics.median(ts):.1f} ms on the device, host issue time {statistics.median(cs):.1f} ms")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step(); torch.cuda.synchronize()
    ev = prof.key_averages()
    tot = sum(e.self_device_time_total for e in ev)
    print(f"   device kernel time {tot / 1e3:.1f} ms in {sum(e.count for e in ev if e.self_device_time_total > 0)} launches")
    rows = sorted(ev, key=lambda e: -e.self_device_time_total)[:28]
    for e in rows:
        print(f"   {e.self_device_time_total / 1e3:8.2f} ms {100 * e.self_device_time_total / tot:5.1f}% x{e.count:4d}  {e.key[:110]}")
