import sys, torch
sys.path.insert(0, '.')
from pointcloudpdf_b200.ptv1 import OpenSegPTv1
gold = torch.load('tests/golden/ptv1_small.pt')
cuda = torch.device('cuda:0')
d = dict(coord=gold["coord"].to(cuda), feat=gold["feat"].to(cuda), offset=gold["offset"].to(cuda))
label = torch.randint(0, 13, (3400,), device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
grads = []
for fused in (True, False, False):
    torch.manual_seed(2024)
    net = OpenSegPTv1(in_channels=6, num_classes=13, method="msp").to(cuda)
    net.backbone.set_fused(fused)
    net.eval()
    loss = torch.nn.functional.cross_entropy(net.backbone(d), label)
    loss.backward()
    print("loss", float(loss))
    grads.append({k: p.grad.clone() for k, p in net.backbone.named_parameters()})
scale = max(float(g.abs().max()) for g in grads[1].values())
for a, b, tag in ((0, 1, "fused vs literal"), (1, 2, "literal vs literal (atomic noise)")):
    errs = sorted(((float((grads[a][k] - grads[b][k]).abs().max()) / max(float(grads[b][k].abs().max()), 1e-3 * scale), k) for k in grads[a]), reverse=True)
    print(tag, errs[:8])
print("---- per layer, backward order")
order = ["cls", "dec1.1", "dec1.0", "dec2.1", "dec2.0", "dec3.1", "dec3.0", "dec4.1", "dec4.0", "dec5.1", "dec5.0", "enc5.2", "enc5.1", "enc5.0", "enc4.5", "enc4.1", "enc3.1", "enc2.1", "enc1.1", "enc1.0"]
for pre in order:
    ks = [k for k in grads[0] if k.startswith(pre + ".")]
    e = max((float((grads[0][k] - grads[1][k]).abs().max()) / max(float(grads[1][k].abs().max()), 1e-3 * scale), k) for k in ks)
    e2 = max((float((grads[1][k] - grads[2][k]).abs().max()) / max(float(grads[1][k].abs().max()), 1e-3 * scale), k) for k in ks)
    print(f"{pre:8s} fused-vs-literal {e[0]:.2e} ({e[1]})   literal noise {e2[0]:.2e}")
