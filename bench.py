#!/usr/bin/env python
"""bench.py -- PTv1 openseg inference throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one synthetic S3DIS-Area-5-shaped room of 80 000
points per GPU: Point Transformer v1 (Seg50, 13 classes, random-init, eval, f32) + the fused MSP
open-set score -- BASELINE.json configs[1], "openseg-pt-v1-0-msp inference".  Rooms are
independent, so N GPUs run N rooms with no data-path collective (weak scaling).

Printed JSON (one line, rank 0):
  value      whole-job points/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        the same through the host-buffer API OpenSegPTv1.infer(): pinned-host -> device copy
             of coord/feat/offset and device -> host read of score + prediction inside the timing
  roofline   the kernel of ours with the largest share of the step, timed live with CUDA events
             on the launching stream inside the timed region (achieved = algorithmic bytes / time)
  kernels    the same accounting for every C-ABI entry point the step calls
  cpu_baseline  the reference op sequence on the host cores (oracle port), bounded sample
--impl reference times that CPU port alone (the reference has no CPU implementation of this path
and its CUDA kernels are not a CPU baseline).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "ptv1_openseg_inference_points_per_sec"
UNIT = "points/s"
N_POINTS = 80_000
NUM_CLASSES = 13
IN_CHANNELS = 6
L2_FLUSH_BYTES = 256 << 20  # > 126 MB L2


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout: route everything libraries print there (NCCL's version banner, ...)
    to stderr and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--cpu-sample-points", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--literal", action="store_true", help="reference op sequence (kNN per block, einsum)")
    ap.add_argument("--linear", default="auto", choices=["auto", "pob", "cublas"],
                    help="frozen linears: pob_linear_forward (fused epilogue), the cuBLAS route, or per shape (auto)")
    ap.add_argument("--depth", type=int, default=12,
                    help="rooms whose H2D copy + coordinate-only work run ahead of the feature path (1 = serial)")
    return ap.parse_args()


# ----------------------------------------------------------------------- clocks sampling --

class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.skip = index, None, [], 0

    def __enter__(self):
        if self.index is None:   # only rank 0 polls NVML: N concurrent nvidia-smi loops contend on the driver lock
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.time()   # NVML start-up (attach, first query) happens before the caller starts its clock
            while not self.lines and time.time() - t0 < 3.0 and self.proc.poll() is None:
                time.sleep(0.01)
            self.skip = len(self.lines)
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[min(self.skip, max(len(self.lines) - 1, 0)):]:   # samples taken under load
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


class NvmlSampler:
    """The same samples taken in-process (pynvml, a daemon thread, every 100 ms): NVML is initialised BEFORE the
    timed region and a poll is two cheap driver queries.  Preferred over the nvidia-smi loop, whose process start-up
    and per-iteration full queries were seen to stall the launching thread for ~200 ms inside a 0.5 s timed region
    (profiles/r01d_experiments.md)."""
    NAMES = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
             ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        handle = None
        try:   # CUDA ordinal -> NVML handle by UUID (CUDA_VISIBLE_DEVICES may renumber)
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            handle = None
        self.handle = handle if handle is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        self._poll()                      # fails here, not in the thread, if a query is unsupported
        self.sm, self.reasons, self.stop = [], set(), threading.Event()

    def _poll(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        return mhz, mask

    def _loop(self):
        while not self.stop.is_set():
            try:
                mhz, mask = self._poll()
                self.sm.append(mhz)
                for name, const in self.NAMES:
                    if mask & int(getattr(self.nv, const)):
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.thread.join(timeout=2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "pynvml"}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "pynvml"}


def make_sampler(index):
    if index is None:
        return ClockSampler(None)
    try:
        return NvmlSampler(index)
    except Exception:
        return ClockSampler(index)      # nvidia-smi loop (B200_PROFILING.md recipe)


# ------------------------------------------------------------------------------ CPU port --

def cpu_port_points_per_sec(n_points: int, steps: int, warmup: int, threads: int, budget_s: float = 150.0):
    """The reference's op sequence for the same workload on host cores: PTv1 Seg50 (literal path:
    kNN in every block, gather k and v, einsum) over the oracle's brute-force operators + MSP.
    This is the one place bench.py executes oracle/ (cpu_baseline / --impl reference)."""
    from oracle import pointops_oracle as O
    from pointcloudpdf_b200 import ptv1, synthetic as S
    import types

    torch.set_num_threads(threads)
    shim = types.SimpleNamespace(
        knn_query=lambda k, xyz, off, nx=None, noff=None: O.knn_query(k, xyz, off, nx, noff),
        farthest_point_sampling=O.farthest_point_sampling, grouping=O.grouping, aggregation=O.aggregation,
        knn_query_and_group=O.knn_query_and_group, interpolation=O.interpolation)
    saved = ptv1.pointops
    ptv1.pointops = shim
    try:
        torch.manual_seed(2024)
        net = ptv1.PointTransformerSeg50(in_channels=IN_CHANNELS, num_classes=NUM_CLASSES).eval().set_fused(False)
        batch = S.s3dis_batch([n_points], seed=2026)
        times = []
        with torch.no_grad():
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                logits = net(dict(coord=batch["coord"], feat=batch["feat"], offset=batch["offset"]),
                             batch["offset"].tolist())
                O.msp_score(logits)
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
                    if sum(times) > budget_s:   # bounded: a slow host must not turn K steps into an hour
                        break
    finally:
        ptv1.pointops = saved
    return n_points / statistics.median(times), statistics.median(times), len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.cpu_sample_points
    pps, sec, timed = cpu_port_points_per_sec(n, max(1, args.steps), max(0, min(args.warmup, 1)), cores)
    sample = f"one S3DIS-shaped room of {n} points per step (bounded sample of the 80000-point workload; " \
             f"brute-force kNN/FPS are O(n^2), so points/s at 80000 would be lower); median of {timed} timed steps"
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "openseg-pt-v1-0-msp inference, PTv1-Seg50, S3DIS-shaped room", "points_per_step": n,
                       "classes": NUM_CLASSES},
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------ operator-level lines (cfg1) --

def ops_cfg1(dev, hbm_peak):
    """The second half of BASELINE.json's metric -- "kNN + group + aggregate GB/s vs HBM peak" -- on
    configs[0]'s shape: one 24 000-point S3DIS-shaped cloud, k = 16, C = 32, share_planes = 8.  Each
    operator is launched 24 times back to back (one CUDA graph, so no host gaps) over 8 rotating argument
    sets (~60 MB each, > L2 in total) between two CUDA events; GB/s = SURVEY.md 8(d) algorithmic bytes /
    mean time per launch."""
    from pointcloudpdf_b200 import synthetic as S
    import pointcloudpdf_b200.pointops as pointops
    n, k, c, wc = 24000, 16, 32, 4
    b = S.s3dis_batch([n], seed=2025)
    xyz, off = b["coord"].to(dev), b["offset"].to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    idx, _ = pointops.knn_query(k, xyz, off)
    sets = [dict(xyz=xyz.clone(), feat=torch.randn(n, c, device=dev, generator=g),
                 pos=torch.randn(n, k, c, device=dev, generator=g), w=torch.randn(n, k, wc, device=dev, generator=g))
            for _ in range(8)]

    def timed(fn, reps=24):
        for a in sets[:2]:
            fn(a)
        torch.cuda.synchronize()
        st = torch.cuda.Stream(device=dev)
        graph = torch.cuda.CUDAGraph()   # back-to-back launches without host gaps
        with torch.cuda.graph(graph, stream=st, capture_error_mode="thread_local"):
            for r in range(reps):
                fn(sets[r % len(sets)])
        times = []
        with torch.cuda.stream(st):
            graph.replay()
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                graph.replay()
                e1.record(st)
                e1.synchronize()
                times.append(e0.elapsed_time(e1))
        return statistics.median(times) / reps * 1e-3

    def knn(a):
        pointops.clear_caches()
        pointops.knn_query(k, a["xyz"], off)

    out = {}
    with torch.no_grad():
        for name, fn, nbytes, flops in (
                ("knn_query k=16 (grid build + query)", knn, 12 * n + 12 * n + 8 * k * n, 8 * n * n),
                ("grouping with_xyz (knn_query_and_group's gather)", lambda a: pointops.grouping(idx, a["feat"], xyz, xyz, with_xyz=True),
                 4 * (n * c + 3 * n + 3 * n + n * k + n * k * (3 + c)), 0),
                ("grouping2 (feature gather)", lambda a: pointops.grouping2(a["feat"], idx), 4 * (n * c + n * k + n * k * c), 0),
                ("aggregation forward", lambda a: pointops.aggregation(a["feat"], a["pos"], a["w"], idx),
                 4 * (n * c + n * k * c + n * k * wc + n * k + n * c), 0)):
            sec = timed(fn)
            out[name] = {"us": sec * 1e6, "alg_MB": nbytes / 1e6, "GBps": nbytes / sec / 1e9,
                         "frac_of_hbm_peak": nbytes / sec / 1e9 / hbm_peak}
            if flops:
                out[name]["bruteforce_equivalent_TFLOPs"] = flops / sec / 1e12
    pointops.clear_caches()
    return {"shape": "N=24000, k=16, C=32, w_c=4 (BASELINE configs[0])",
            "timing": "a CUDA graph of 24 back-to-back launches over 8 rotating argument sets, CUDA events around a replay, median of 5", "ops": out}


# ------------------------------------------------------------------------------- B200 arm --

def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    from pointcloudpdf_b200 import _lib, synthetic as S
    from pointcloudpdf_b200 import ptv1 as _ptv1
    from pointcloudpdf_b200.ptv1 import OpenSegPTv1
    import pointcloudpdf_b200.pointops as pointops
    _ptv1.set_linear_backend(args.linear)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    K, W = args.steps, max(args.warmup, 3)
    torch.backends.cuda.matmul.allow_tf32 = False  # f32 means f32
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(2024)
    net = OpenSegPTv1(in_channels=IN_CHANNELS, num_classes=NUM_CLASSES, method="msp").to(dev).eval()
    net.backbone.set_fused(not args.literal)

    # one distinct room per rank and per step slot (weak scaling: every GPU does a full room)
    n_rooms = min(K, 4)
    rooms = [S.s3dis_batch([args.points], seed=2026 + 101 * rank + i) for i in range(n_rooms)]
    host = [dict(coord=r["coord"].pin_memory(), feat=r["feat"].pin_memory(), offset=r["offset"].pin_memory()) for r in rooms]
    resident = [dict(coord=r["coord"].to(dev), feat=r["feat"].to(dev), offset=r["offset"].to(dev)) for r in rooms]
    off_host = [r["offset"].tolist() for r in rooms]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    depth = max(1, args.depth) if not args.literal else 1

    def step_resident(i):
        flush.zero_()                       # evict L2 between timed iterations
        pointops.clear_caches()             # nothing computed in a previous step may be reused
        with torch.no_grad():
            return net(resident[i % n_rooms], off_host[i % n_rooms])["score"]

    def step_e2e(i):
        flush.zero_()
        pointops.clear_caches()
        h = host[i % n_rooms]
        return net.infer(h["coord"], h["feat"], h["offset"], device=dev)

    class Flushed:
        """Room sequence for infer_stream that evicts L2 before each room is handed out."""
        def __init__(self, src, n):
            self.src, self.n = src, n
        def __iter__(self):
            for i in range(self.n):
                r = self.src[i % n_rooms]
                yield (r["coord"], r["feat"], r["offset"])

    def run_stream(src, n, graphs="auto"):
        rooms_seq = list(Flushed(src, n))
        last = None
        for k, (score, pred) in enumerate(net.infer_stream(rooms_seq, depth=depth, device=dev, graphs=graphs)):
            flush.zero_()                   # L2 eviction on the main stream between rooms
            last = (score, pred)
        return last

    for i in range(W):
        step_resident(i)
        step_e2e(i)
    if depth > 1:   # warm the side streams' allocator pools, capture the room graphs, warm the schedule itself
        run_stream(resident, max(W, depth + 2), graphs=False)
        run_stream(resident, max(W, depth + 2))
        run_stream(host, max(W, depth + 2))
    barrier()

    # ---- value: device-resident inputs, K steps, CUDA events on the main stream ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with make_sampler(local if rank == 0 else None) as clocks:
        # the sampler's start-up (NVML attach, up to a second) leaves the GPU idle and its clocks parked: a short
        # untimed burst of the same schedule brings them back before the clock starts (W warm-up steps were done above)
        if depth > 1:
            run_stream(resident, depth + 2)
        else:
            step_resident(0)
        barrier()
        launches0 = _lib.launch_count()
        e0.record()
        if depth > 1:
            run_stream(resident, K)
        else:
            for i in range(K):
                step_resident(i)
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0

    # ---- the same K steps once more with a CUDA-event bracket around every C-ABI call (on the
    #      stream it launches on): attributes the step to kernels; its own wall time is reported
    #      separately because ~250 event records per room are not free ----
    prof = _lib.OpProfile()
    _lib.PROFILE = prof
    barrier()
    p0_, p1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0_.record()
    if depth > 1:
        run_stream(resident, K, graphs=False)   # eager launches: the C-ABI calls are what the events bracket
    else:
        for i in range(K):
            step_resident(i)
    p1_.record()
    barrier()
    _lib.PROFILE = None
    ms_profiled = p0_.elapsed_time(p1_)
    ops = prof.summary()

    # ---- e2e: host buffers in, host score out ----
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    if depth > 1:
        score, pred = run_stream(host, K)
    else:
        for i in range(K):
            score, pred = step_e2e(i)
    e3.record()
    barrier()
    e2e_ms_total = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)

    t = torch.tensor([ms_total, e2e_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total = float(t[0]), float(t[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        step_ms = ms_total / K
        kernels = {}
        for name, d in sorted(ops.items(), key=lambda kv: -kv[1]["ms"]):
            per_call_ms = d["ms"] / d["calls"]
            gbs = d["alg_bytes"] / d["calls"] / (per_call_ms * 1e-3) / 1e9 if per_call_ms > 0 else 0.0
            kernels[name] = {"calls_per_step": d["calls"] / K, "ms_per_step": d["ms"] / K,
                             "share_of_step": d["ms"] / ms_profiled, "alg_MB_per_call": d["alg_bytes"] / d["calls"] / 1e6,
                             "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / hbm_peak,
                             "alg_GFLOP_per_call": d["alg_flops"] / d["calls"] / 1e9,
                             "achieved_TFLOPs": d["alg_flops"] / d["calls"] / (per_call_ms * 1e-3) / 1e12 if per_call_ms > 0 else 0.0}
        # FPS is the largest kernel by time but it is a serial dependent chain on 16 SMs (latency-bound: no
        # byte or flop roofline describes it; its accounting is reported as `dominant_kernel`).  `roofline`
        # is the largest bandwidth-class kernel of the step: the fused group + aggregate layer.
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        except (OSError, ValueError):
            pass
        # FP32-issue / latency class kernels (their byte roofline says nothing): FPS, kNN, the linears
        bw_class = [k for k in kernels if k not in ("pob_farthest_point_sampling", "pob_knn_grid_query", "pob_knn_grid_build",
                                                    "pob_linear_forward")]
        top = bw_class[0] if bw_class else None
        roof = None
        if top:
            kd = kernels[top]
            tr = traffic.get(top, {})
            roof = {"kernel": top, "bound": "hbm", "achieved": kd["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kd["frac_of_hbm_peak"], "traffic": tr.get("dram_bytes_per_launch"),
                    "traffic_source": tr.get("source"), "peak_source": peak_src,
                    "share_of_step": kd["share_of_step"], "avg_launch_ms": kd["ms_per_step"] / kd["calls_per_step"],
                    "alg_bytes_per_launch": kd["alg_MB_per_call"] * 1e6,
                    "unfused_operator_equivalent": {
                        "bytes_per_launch": ops[top]["unfused_bytes"] / ops[top]["calls"],
                        "GBps": ops[top]["unfused_bytes"] / ops[top]["calls"] / (kd["ms_per_step"] / kd["calls_per_step"] * 1e-3) / 1e9,
                        "what": "SURVEY.md 8(d) algorithmic bytes of the reference operators one launch replaces (knn_query_and_group "
                                "gather of k with xyz + grouping of v + aggregation forward) / the same launch time: the traffic "
                                "fusion removed, for context -- NOT bytes this kernel moves"} if ops[top].get("unfused_bytes") else None,
                    "note": "achieved = compulsory bytes of the FUSED op (q, k, v, out rows, indices, coordinates once) / "
                            "mean launch time by CUDA events in the instrumented pass; the (n, ns, C) tensors the unfused "
                            "reference ops would move are never materialised, so the kernel is bound by L2 gathers and FP32 "
                            "issue, not HBM (profiles/)"}
        dominant = None
        if kernels:
            name = next(iter(kernels))
            kd = kernels[name]
            clk = (clocks.summary().get("sm_mhz") or 1965.0) * 1e6
            fp32_peak = 148 * 128 * 2 * clk / 1e12
            tf = kd["alg_GFLOP_per_call"] / (kd["ms_per_step"] / kd["calls_per_step"]) if kd["ms_per_step"] > 0 else 0.0
            dominant = {"kernel": name, "bound": "latency (serial chain; FP32 CUDA cores)", "share_of_step": kd["share_of_step"],
                        "avg_launch_ms": kd["ms_per_step"] / kd["calls_per_step"],
                        "bruteforce_equivalent_TFLOPs": tf, "fp32_peak_TFLOPs": fp32_peak, "frac_of_fp32_peak": tf / fp32_peak,
                        "note": "exact pruning skips ~95 % of the distance updates the flop count assumes; the kernel runs on "
                                "one 16-CTA cluster and overlaps other rooms' kernels (share_of_step can exceed 1)"}
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        d2h = score.numel() * score.element_size() + pred.numel() * pred.element_size()
        line = {"metric": METRIC, "value": world * args.points * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "openseg-pt-v1-0-msp inference (BASELINE configs[1]): PTv1-Seg50 + fused MSP score, "
                                       "one S3DIS-Area-5-shaped room per GPU per step",
                           "points_per_step_per_gpu": args.points, "classes": NUM_CLASSES, "in_channels": IN_CHANNELS,
                           "op_sequence": "literal (kNN per block, einsum)" if args.literal else
                                          "one kNN per stage + fused aggregation kernel",
                           "linears": {"pob": "pob_linear_forward (FP32 FFMA tiles, bias / skip / ReLU on the accumulators)",
                                       "cublas": "cuBLAS through torch (SIMT sgemm + cuBLASLt bias pass)",
                                       "auto": "per shape: pob_linear_forward for linears with a bias / skip / ReLU epilogue and the "
                                               "80000-row layers, cuBLAS for the plain q/k/v GEMMs of the deeper stages"}[args.linear],
                           "l2": "256 MiB memset between timed iterations (inside the timed region)",
                           "schedule": (f"rooms served in order by OpenSegPTv1.infer_stream, {depth} in flight: each room is one "
                                        f"CUDA-graph replay (coordinate branch: FPS + kNN, forked; feature branch; joined) on "
                                        f"its own stream, so the serial FPS chains of the next rooms run under the feature "
                                        f"path of the current one; every room is computed in full inside the timed region") if depth > 1
                                       else "one room at a time (geometry side stream within the room)",
                           "parallelism": f"scene-sharded x{world}, no data-path collective"},
                "e2e": {"value": world * args.points * K / (e2e_ms_total * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_total / K},
                "gpu_launches": launches, "gpu_launches_per_step": launches / K,
                "kernel_attribution": {"how": "second pass of the same K steps, launched eagerly (no graph replay), with CUDA events around every C-ABI call; "
                                              "with rooms in flight concurrently, kernel times overlap and shares can sum past 1",
                                       "ms_per_step_instrumented": ms_profiled / K},
                "roofline": roof, "dominant_kernel": dominant, "kernels": kernels, "clocks": clocks.summary()}
        line["ops_cfg1"] = ops_cfg1(dev, hbm_peak)
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n = args.cpu_sample_points
            pps, sec, _ = cpu_port_points_per_sec(n, 3, 1, cores)
            line["cpu_baseline"] = {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"one {n}-point S3DIS-shaped room, reference op sequence over the "
                                              f"brute-force oracle operators, {sec:.2f} s per room (O(n^2): an "
                                              f"80000-point room would be slower per point)"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
