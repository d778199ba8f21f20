"""GPU parity of the inference-only fused kernels (csrc/ptlayer.cu) against plain PyTorch f32
restatements of the reference caller's arithmetic (point_transformer_seg.py:48-81, 106-119,
168-170, 188-195), and of the frozen (folded-BatchNorm) model against the unfolded one.

Tolerance: 1e-5 relative to the output's magnitude (f32, different summation order), as
north_star states for aggregation-type ops; end-to-end frozen vs unfrozen logits 1e-4 absolute."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def randomise_bn(module, gen):
    for m in module.modules():
        if isinstance(m, nn.BatchNorm1d):
            c = m.num_features
            m.running_mean.copy_(torch.randn(c, generator=gen) * 0.2)
            m.running_var.copy_(torch.rand(c, generator=gen) + 0.5)
            m.weight.data.copy_(torch.rand(c, generator=gen) + 0.5)
            m.bias.data.copy_(torch.randn(c, generator=gen) * 0.2)


def layer_reference(layer, bn2, x_q, x_k, x_v, xyz, idx):
    """PointTransformerLayer.forward after the q/k/v linears, in plain torch, for a given idx
    (placeholders -1 group to zero rows, functions/grouping.py:41-57), then relu(bn2(.))."""
    n, ns = idx.shape
    mask = (idx >= 0)
    j = idx.clamp(min=0).long()
    p_r = (xyz[j] - xyz[:, None, :]) * mask[..., None]
    x_kg = x_k[j] * mask[..., None]
    x_vg = x_v[j] * mask[..., None]
    p_r = layer.linear_p(p_r)
    r_qk = x_kg - x_q[:, None, :] + p_r
    w = torch.softmax(layer.linear_w(r_qk), dim=1)
    s = layer.share_planes
    c = layer.out_planes
    out = torch.einsum("ntsi,nti->nsi", (x_vg + p_r).view(n, ns, s, c // s), w).reshape(n, c)
    return torch.relu(bn2(out)) if bn2 is not None else out


@pytest.mark.parametrize("c,ns", [(32, 8), (64, 16), (128, 16), (256, 16), (512, 16), (32, 16), (128, 8)])
@pytest.mark.parametrize("tail", [True, False])
@pytest.mark.parametrize("split", [0, -1, 1, 4, 16])
def test_pt_layer_forward_matches_torch(cuda, c, ns, tail, split):
    """split: 0 = the CTA-tiled kernel (default), -1 / 1 / 4 / 16 = warp-per-point variants: every kernel
    (one warp per point, 4 warps, one warp per neighbour) must agree with the torch restatement."""
    from pointcloudpdf_b200 import _lib, ptv1, synthetic as S
    from pointcloudpdf_b200.pointops import fused as FZ
    import pointops
    gen = torch.Generator().manual_seed(c * 100 + ns)
    torch.manual_seed(c + ns)
    block = ptv1.Bottleneck(c, c, 8, ns)
    randomise_bn(block, gen)
    block = block.to(cuda).eval()
    n = 3000 if c <= 128 else 700
    batch = S.s3dis_batch([n - 40, 40], seed=7)
    xyz, off = batch["coord"].to(cuda), batch["offset"].to(cuda)
    idx, _ = pointops.knn_query(ns, xyz, off)
    idx = idx.clone()
    idx[5, ns - 3:] = -1          # placeholders, as a scene with fewer than ns points would give
    idx[n - 1, 1:] = -1
    qkv = torch.randn(n, 3 * c, generator=gen).to(cuda)
    q, k, v = qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:]
    with torch.no_grad():
        ref = layer_reference(block.transformer, block.bn2 if tail else None, q, k, v, xyz, idx)
        f = block.frozen()
        out = FZ.pt_layer_forward(q, k, v, xyz, idx, f["params"], out_affine=tail, split=split)   # per-call option
    err = (out - ref).abs().max().item()
    assert err <= 1e-5 * max(ref.abs().max().item(), 1.0), err


@pytest.mark.parametrize("cin,cout,ns", [(32, 64, 16), (64, 128, 16), (256, 512, 16), (32, 64, 8)])
def test_transition_down_pool_matches_torch(cuda, cin, cout, ns):
    from pointcloudpdf_b200 import ptv1, synthetic as S
    import pointops
    gen = torch.Generator().manual_seed(cin + cout)
    torch.manual_seed(cin)
    td = ptv1.TransitionDown(cin, cout, 4, ns)
    randomise_bn(td, gen)
    td = td.to(cuda).eval()
    batch = S.s3dis_batch([4000, 2000], seed=11)
    xyz, off = batch["coord"].to(cuda), batch["offset"].to(cuda)
    x = torch.randn(6000, cin, generator=gen).to(cuda)
    cloud = ptv1.Cloud(xyz, x, off, batch["offset"].tolist())
    with torch.no_grad():
        td.use_frozen = False
        ref = td(cloud)
        td.use_frozen = True
        out = td(cloud)
    assert torch.equal(ref.p, out.p) and ref.o_host == out.o_host
    err = (out.x - ref.x).abs().max().item()
    assert err <= 1e-5 * max(ref.x.abs().max().item(), 1.0), err


def test_interpolation_add_and_affine_act(cuda):
    from pointcloudpdf_b200.pointops import fused as FZ
    import pointops
    from pointcloudpdf_b200 import synthetic as S
    gen = torch.Generator().manual_seed(3)
    batch = S.s3dis_batch([5000], seed=5)
    fine, off = batch["coord"].to(cuda), batch["offset"].to(cuda)
    coarse = fine[::4].contiguous()
    coff = torch.tensor([coarse.shape[0]], dtype=torch.int32, device=cuda)
    feat = torch.randn(coarse.shape[0], 64, generator=gen).to(cuda)
    base = torch.randn(fine.shape[0], 64, generator=gen).to(cuda)
    ref = base + pointops.interpolation(coarse, fine, feat, coff, off)
    from pointcloudpdf_b200.pointops.interpolation import _neighbours_and_weights
    idx, w = _neighbours_and_weights(coarse, fine, coff, off, 3)
    out = FZ.interpolation_add(feat, idx, w, base=base)
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    out2 = FZ.interpolation_add(feat, idx, w, base=None)
    assert (out2 - (ref - base)).abs().max().item() <= 2e-5 * ref.abs().max().item()

    x = torch.randn(1000, 32, generator=gen).to(cuda)
    sc, sh = torch.rand(32, generator=gen).to(cuda), torch.randn(32, generator=gen).to(cuda)
    res = torch.randn(1000, 32, generator=gen).to(cuda)
    assert torch.allclose(FZ.affine_act(x, sc, sh, res, relu=True), torch.relu(x * sc + sh + res), atol=1e-6)
    assert torch.allclose(FZ.affine_act(x, None, sh, None, relu=False), x + sh, atol=1e-6)
    y = x.clone()
    FZ.affine_act(y, None, sh, res, relu=True, inplace=True)
    assert torch.allclose(y, torch.relu(x + sh + res), atol=1e-6)


@pytest.mark.parametrize("method", ["msp", "pdf"])
@pytest.mark.parametrize("sizes", [[9000, 5000], [9000, 600]])
def test_frozen_model_matches_unfrozen(cuda, method, sizes):
    """The folded / fused inference form against the same model run module by module (eval).
    [9000, 600]: the small scene has fewer than nsample points from the third level on, so those levels
    see placeholder neighbours and must keep the q / k / v biases in the GEMM (ptv1.Bottleneck._freeze)."""
    from pointcloudpdf_b200 import ptv1, synthetic as S
    torch.manual_seed(2024)
    net = ptv1.OpenSegPTv1(in_channels=6, num_classes=13, method=method)
    randomise_bn(net, torch.Generator().manual_seed(1))
    net = net.to(cuda).eval()
    batch = S.s3dis_batch(sizes, seed=3)
    d = dict(coord=batch["coord"].to(cuda), feat=batch["feat"].to(cuda), offset=batch["offset"].to(cuda))
    outs = {}
    for frozen in (True, False):
        for m in net.modules():
            if isinstance(m, ptv1._Freezable):
                m.use_frozen = frozen
        with torch.no_grad():
            outs[frozen] = net(d, batch["offset"].tolist())
    a, b = outs[True], outs[False]
    scale = max(b["seg_logits"].abs().max().item(), 1.0)
    assert (a["seg_logits"] - b["seg_logits"]).abs().max().item() <= 1e-4 * scale
    assert (a["score"] - b["score"]).abs().max().item() <= 1e-4
    assert (a["pred"].long() != b["seg_logits"].argmax(-1)).float().mean().item() < 1e-3


def test_frozen_cache_is_dropped_when_weights_may_change(cuda):
    from pointcloudpdf_b200 import ptv1
    torch.manual_seed(0)
    block = ptv1.Bottleneck(32, 32, 8, 8).to(cuda).eval()
    f1 = block.frozen()
    assert block.frozen() is f1
    block.train()
    assert block._frozen is None
    block.eval()
    block.frozen()
    block.load_state_dict(block.state_dict())
    assert block._frozen is None
    block.frozen()
    block.float()
    assert block._frozen is None
    # training mode or grad mode never takes the frozen path
    x = torch.randn(8, 32, device=cuda)
    assert not block._can_freeze(x)            # grad enabled
    with torch.no_grad():
        assert block._can_freeze(x)
        block.train()
        assert not block._can_freeze(x)
