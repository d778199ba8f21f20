import sys, torch
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S, _lib
from pointcloudpdf_b200.ptv1 import OpenSegPTv1, _RoomGraph
import pointops
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(2024)
net = OpenSegPTv1(in_channels=6, num_classes=13, method="msp").to(dev).eval()
r = S.s3dis_batch([5000, 1200], seed=40)
coord, feat, off = r["coord"].pin_memory(), r["feat"].pin_memory(), r["offset"]
def eager():
    pointops.clear_caches()
    d = dict(coord=coord.to(dev), feat=feat.to(dev), offset=off.to(dev))
    out = net.forward(d, off.tolist())
    return out["seg_logits"].clone(), out["score"].clone()
l0, s0 = eager(); l1, s1 = eager()
print("eager vs eager: logits", (l0 - l1).abs().max().item(), "score", (s0 - s1).abs().max().item())
g = [_RoomGraph(net, off.tolist(), 6, dev) for _ in range(2)]
res = []
for k in (0, 0, 1, 1, 0):
    ev, s, p = g[k].run(coord, feat); ev.synchronize(); res.append(s.clone())
for i, k in enumerate((0, 0, 1, 1, 0)):
    print("slot", k, "vs eager", (res[i] - s0.cpu()).abs().max().item(), "vs first replay", (res[i] - res[0]).abs().max().item())
# literal (unfused) model for reference
net.backbone.set_fused(False)
l2, s2 = eager()
print("unfrozen literal vs frozen eager: score", (s2 - s0).abs().max().item())
