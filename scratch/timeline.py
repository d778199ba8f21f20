"""Timeline of OpenSegPTv1.infer_stream: host timestamps and CUDA-event times of each room's
geometry (side stream) and feature path (main stream).  python scratch/timeline.py [depth] [rooms]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200 import synthetic as S
from pointcloudpdf_b200.ptv1 import OpenSegPTv1

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_rooms = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(2024)
net = OpenSegPTv1(in_channels=6, num_classes=13, method="msp").to(dev).eval()
rooms = [S.s3dis_batch([80000], seed=2026 + i) for i in range(4)]
host = [(r["coord"].pin_memory(), r["feat"].pin_memory(), r["offset"].pin_memory()) for r in rooms]
seq = [host[i % 4] for i in range(n_rooms)]

log = []
geo_orig = net.backbone.geometry
fwd_orig = net.forward
t00 = [0.0]
ev0 = torch.cuda.Event(enable_timing=True)


def geometry(p0, o0, offset_host, stream=None):
    h0 = time.perf_counter()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    stream.wait_stream(torch.cuda.current_stream())
    s.record(stream)
    out = geo_orig(p0, o0, offset_host, stream)
    e.record(stream)
    log.append(("geo", h0 - t00[0], time.perf_counter() - t00[0], s, e))
    return out


def forward(d, offset_host=None, geometry=None):
    h0 = time.perf_counter()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    out = fwd_orig(d, offset_host, geometry)
    e.record()
    log.append(("fwd", h0 - t00[0], time.perf_counter() - t00[0], s, e))
    return out


for _ in net.infer_stream(seq[:6], depth=depth, device=dev):
    pass
torch.cuda.synchronize()
net.backbone.geometry = geometry
net.forward = forward
t00[0] = time.perf_counter()
ev0.record()
k = 0
for _ in net.infer_stream(seq, depth=depth, device=dev):
    log.append(("yield", time.perf_counter() - t00[0], 0, None, None))
torch.cuda.synchronize()
total = time.perf_counter() - t00[0]
print(f"depth {depth}: {n_rooms} rooms in {total*1e3:.1f} ms = {total*1e3/n_rooms:.2f} ms/room")
print("kind   host_start host_end | dev_start dev_end (ms)")
for kind, h0, h1, s, e in log:
    if s is None:
        print(f"{kind:5s} {h0*1e3:9.2f}")
    else:
        print(f"{kind:5s} {h0*1e3:9.2f} {h1*1e3:8.2f} | {ev0.elapsed_time(s):8.2f} {ev0.elapsed_time(e):8.2f}")
