"""GPU parity of pob_linear_forward (csrc/linear.cu) -- the FP32 linear with fused bias / skip / ReLU
epilogue that the frozen PTv1 form runs instead of cuBLAS GEMM + eager BN / ReLU / add
(point_transformer_seg.py:87-95,128-147,178-195) -- against a float64 torch restatement.

Tolerance: 1e-5 relative to the output's magnitude (f32 FFMA accumulation in a different order than
the reference's cuBLAS GEMM), the bar north_star sets for the floating-point operators."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# the shapes of PTv1-Seg50 on an 80 000-point room (rows shrunk where the tile choice does not depend on them)
SHAPES = [
    (20000, 6, 32),      # enc1 input layer: K not a multiple of 4 -> scalar A loads
    (20000, 32, 13),     # classifier: N not a multiple of 4 -> scalar W loads / stores
    (17001, 32, 96),     # stage 1 q/k/v, ragged last row tile, tile 128x32
    (5000, 128, 128),    # tile 64x32, split-K 2
    (5000, 128, 384),
    (1250, 256, 768),    # tile 32x32, split-K 4
    (312, 512, 512),     # tile 16x32, split-K 8
    (312, 512, 1536),
    (1, 512, 512),       # the scene-mean row of the decoder head
    (37, 20, 7),         # nothing aligned
    (4099, 72, 40),      # K not a multiple of the k-tile, N not a multiple of the column tile
]


def reference(x, wt, bias, residual, relu):
    y = x.double() @ wt.double()
    if bias is not None:
        y = y + bias.double()
    if residual is not None:
        y = y + residual.double()
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("m,k,n", SHAPES)
@pytest.mark.parametrize("epilogue", ["plain", "bias_relu", "bias_residual_relu", "residual"])
def test_linear_matches_float64(cuda, m, k, n, epilogue):
    from pointcloudpdf_b200.pointops import fused as FZ
    g = torch.Generator(device=cuda).manual_seed(m * 31 + k * 7 + n)
    x = torch.randn(m, k, device=cuda, generator=g)
    wt = torch.randn(k, n, device=cuda, generator=g) / k ** 0.5
    bias = torch.randn(n, device=cuda, generator=g) if "bias" in epilogue else None
    res = torch.randn(m, n, device=cuda, generator=g) if "residual" in epilogue else None
    relu = "relu" in epilogue
    out = FZ.linear(x, wt, bias, res, relu)
    ref = reference(x, wt, bias, res, relu)
    assert out.shape == (m, n) and out.dtype == torch.float32
    scale = float(ref.abs().max())
    assert float((out.double() - ref).abs().max()) <= 1e-5 * scale


@pytest.mark.parametrize("config", list(range(1, 22)))
@pytest.mark.parametrize("m,k,n", [(700, 64, 96), (130, 8, 12), (33, 512, 136)])
def test_every_tile_configuration_agrees(cuda, config, m, k, n):
    """Configurations 1-16 are the FFMA tiles, 17-20 the tensor-core (3xTF32) tiles, 21 the FFMA tile the shape would pick.
    The tile is normally picked from the shape; forced here so that each instantiation (4x4 / 8x4 / 8x8
    register tiles, split-K 1..8 with its reduction tree) sees ragged rows, ragged columns and short K."""
    from pointcloudpdf_b200 import _lib
    from pointcloudpdf_b200.pointops import fused as FZ
    g = torch.Generator(device=cuda).manual_seed(config * 1000 + m)
    x = torch.randn(m, k, device=cuda, generator=g)
    wt = torch.randn(k, n, device=cuda, generator=g) / k ** 0.5
    bias = torch.randn(n, device=cuda, generator=g)
    res = torch.randn(m, n, device=cuda, generator=g)
    out = FZ.linear(x, wt, bias, res, True, config=config)   # per-call option: the library keeps no tuning state
    ref = reference(x, wt, bias, res, True)
    assert float((out.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_linear_on_column_block_and_rejects_bad_input(cuda):
    from pointcloudpdf_b200.pointops import fused as FZ
    g = torch.Generator(device=cuda).manual_seed(5)
    wide = torch.randn(3000, 96, device=cuda, generator=g)
    wt = torch.randn(32, 64, device=cuda, generator=g)
    x = wide[:, 32:64]                                   # row stride 96, unit column stride
    out = FZ.linear(x, wt, None, None, False)
    ref = reference(x, wt, None, None, False)
    assert float((out.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    with pytest.raises(ValueError):
        FZ.linear(x.cpu(), wt)                           # no CPU path
    with pytest.raises(ValueError):
        FZ.linear(wide, wt)                              # K mismatch
    with pytest.raises(ValueError):
        FZ.linear(torch.randn(32, 3000, device=cuda).t(), wt)   # column stride != 1


def test_frozen_model_same_logits_on_both_linear_backends(cuda):
    """OpenSegPTv1 (frozen form) with pob_linear_forward vs the cuBLAS route: same logits to f32
    summation-order noise, same predictions up to near-ties."""
    from pointcloudpdf_b200 import ptv1, synthetic as S
    torch.manual_seed(2024)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net = ptv1.PointTransformerSeg50(in_channels=6, num_classes=13).to(cuda).eval()
        b = S.s3dis_batch([9000, 3000], seed=11)
        data = {k: v.to(cuda) for k, v in b.items() if k in ("coord", "feat", "offset")}
        outs = {}
        for backend in ("cublas", "pob", "auto"):
            ptv1.set_linear_backend(backend)
            with torch.no_grad():
                outs[backend] = net(data, b["offset"].tolist()).clone()
    finally:
        ptv1.set_linear_backend("auto")
        torch.backends.cuda.matmul.allow_tf32 = prev
    c = outs["cublas"]
    for name in ("pob", "auto"):
        assert float((outs[name] - c).abs().max()) <= 1e-4 * max(1.0, float(c.abs().max())), name


@pytest.mark.parametrize("m,k,n", [(80000, 32, 96), (20000, 64, 64), (5000, 128, 384), (1250, 256, 768), (312, 512, 1536), (312, 512, 256)])
def test_tensor_core_form_is_f32_accurate(cuda, m, k, n):
    """3xTF32: error against float64 within 1e-5 of the largest output (the bar of every linear test), and within a few f32
    ulps of the FFMA form -- NOT the 1e-3 of plain TF32 (which the parity bars of this path exclude)."""
    from pointcloudpdf_b200.pointops import fused as FZ
    g = torch.Generator(device=cuda).manual_seed(m + k + n)
    x = torch.randn(m, k, device=cuda, generator=g) * 3
    wt = torch.randn(k, n, device=cuda, generator=g) / k ** 0.5
    bias = torch.randn(n, device=cuda, generator=g)
    res = torch.randn(m, n, device=cuda, generator=g)
    ref = torch.relu(x.double() @ wt.double() + bias.double() + res.double())
    scale = float(ref.abs().max())
    mma = FZ.linear(x, wt, bias, res, True, config=0)
    ffma = FZ.linear(x, wt, bias, res, True, config=21)
    err_mma = float((mma.double() - ref).abs().max()) / scale
    err_ffma = float((ffma.double() - ref).abs().max()) / scale
    assert err_mma <= 5e-6 and err_ffma <= 2e-6, (err_mma, err_ffma)
    tf32 = torch.relu((x.double().float().to(torch.float32)) @ wt + bias + res)   # what one TF32 pass would cost, for the record
    assert float((mma - ffma).abs().max()) <= 6e-6 * scale
