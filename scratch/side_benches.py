"""Side measurements for the BASELINE configs that are parity cases rather than the bench line
(cfg3 training step, cfg4 PDF scoring, cfg5 large-scene sweep).  Writes one JSON document.
    python scratch/side_benches.py [out.json]            (1 GPU)
    torchrun --nproc-per-node N scratch/side_benches.py   (cfg3 + sharded kNN at N GPUs)"""
import json, os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pointcloudpdf_b200 import synthetic as S, sharding, _lib
from pointcloudpdf_b200.ptv1 import OpenSegPTv1
import pointcloudpdf_b200.pointops as pointops

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = False
out = {"world": world}

def ev_time(fn, warm=2, reps=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        if world > 1: dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = torch.tensor([statistics.median(ts)], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])

# ---- cfg3: PTv1 ScanNet20-shaped training step, 8 scenes x ~95k points, scene-sharded ----
g = torch.Generator().manual_seed(2027)
sizes = [int(x) for x in torch.randint(90000, 100001, (8,), generator=g)]
mine = sharding.shard_scenes(8, rank, world)
batch = S.scannet_batch([sizes[i] for i in mine], seed=2027 + rank)
d = {k: batch[k].to(dev) for k in ("coord", "feat", "offset")}
label = torch.randint(0, 20, (d["coord"].shape[0],), device=dev)
torch.manual_seed(2024)
net = OpenSegPTv1(in_channels=9, num_classes=20, method="msp").to(dev).train()
opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)
off_host = batch["offset"].tolist()
def train_step():
    pointops.clear_caches()
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(net.backbone(d, off_host), label)
    loss.backward()
    sharding.allreduce_gradients(list(net.backbone.parameters()))
    opt.step()
ms = ev_time(train_step, warm=2, reps=4)
out["cfg3_training_step"] = {"scenes_total": 8, "points_total": sum(sizes), "scenes_per_gpu": len(mine), "ms_per_step": ms,
                             "points_per_sec": sum(sizes) / ms * 1e3, "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30,
                             "what": "fwd + bwd + gradient all-reduce + SGD step, f32, autograd through every pointops kernel"}
del net, opt, d, label; torch.cuda.empty_cache()

if rank == 0:
    # ---- cfg4: PDF U-decoder + fused score on ScanNet-shaped scenes ----
    torch.manual_seed(2024)
    net = OpenSegPTv1(in_channels=9, num_classes=20, method="pdf").to(dev).eval()
    rooms = [S.scannet_batch([150000], seed=3000 + i) for i in range(3)]
    seq = [(r["coord"].pin_memory(), r["feat"].pin_memory(), r["offset"]) for r in rooms]
    run = lambda n: [None for _ in net.infer_stream([seq[i % 3] for i in range(n)], depth=6)]
    run(8); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(24); torch.cuda.synchronize(); sec = time.perf_counter() - t0
    out["cfg4_pdf_inference"] = {"points_per_room": 150000, "classes": 20, "rooms": 24, "ms_per_room": sec / 24 * 1e3,
                                 "points_per_sec": 24 * 150000 / sec, "what": "backbone + PDF U-decoder + fused softmax score, host buffers in / out, graph replay depth 6"}
    from pointcloudpdf_b200.scoring import pseudo_label_prefix
    lg, conf, _unk, lab = S.openset_logits(150000, 20)
    lg_d = lg.to(dev); off = torch.tensor([150000], dtype=torch.int32, device=dev)
    from pointcloudpdf_b200.scoring import fused_scores
    ms = ev_time(lambda: fused_scores(lg_d, offset=off, beta=1.5, want=("msp_prob", "ml_norm")), warm=3, reps=9)
    out["cfg4_pseudo_label_scoring_pass"] = {"n": 150000, "K": 20, "ms": ms, "GBps": 4 * 150000 * (20 + 2) / ms / 1e6}
    del net; torch.cuda.empty_cache()

# ---- cfg5: large single scene: kNN sweep (queries sharded when world > 1), FPS sweep (rank 0) ----
out["cfg5_knn"], out["cfg5_fps"] = [], []
for n in (100000, 250000, 500000, 1000000, 2000000):
    b = S.s3dis_batch([n], seed=2029)
    xyz, off = b["coord"].to(dev), b["offset"].to(dev)
    for k in (16, 32):
        def run_knn():
            pointops.clear_caches()
            return sharding.sharded_knn_query(k, xyz, off, [n])
        ms = ev_time(run_knn, warm=1, reps=3)
        out["cfg5_knn"].append({"n": n, "k": k, "ms": ms, "queries_per_sec": n / ms * 1e3, "gpus": world})
    if rank == 0 and world == 1 and n <= 500000:
        m = n // 4
        noff = torch.tensor([m], dtype=torch.int32, device=dev)
        def run_fps():
            pointops.clear_caches()
            return pointops.farthest_point_sampling(xyz, off, noff)
        t0 = time.perf_counter(); run_fps(); torch.cuda.synchronize(); first = time.perf_counter() - t0
        ms = first * 1e3 if first > 2.0 else ev_time(run_fps, warm=0, reps=2)
        out["cfg5_fps"].append({"n": n, "m": m, "ms": ms, "ns_per_sample": ms * 1e6 / m,
                                "kernel": "register-resident cluster" if n <= 131072 else "streamed (fps_stream_kernel)"})
    del xyz, off; torch.cuda.empty_cache()
if rank == 0:
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/side_benches.json"
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
