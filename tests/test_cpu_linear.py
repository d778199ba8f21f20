"""CPU-side checks around pob_linear_forward (csrc/linear.cu): no GPU here, so
  * the kernel's index arithmetic (k-major shared-memory staging, 4x4 / 8x4 / 8x8 register tiles, the intra-CTA
    split-K tree, ragged row / column tiles) is replayed thread by thread in numpy for every instantiated tile
    configuration (a transliteration of the kernel) and compared with float64;
  * the per-shape backend policy of the frozen PTv1 form and the bench's one-JSON-line contract are exercised."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emulator():
    spec = importlib.util.spec_from_file_location("linear_emulate", os.path.join(ROOT, "tools", "linear_emulate.py"))
    src = open(spec.origin).read().split("\ncfgs = [")[0]          # the functions only, not the sweep at the bottom
    mod = {}
    exec(compile(src, spec.origin, "exec"), mod)
    return mod["run"]


# (BM, BN, BK, KS, TM, TN) exactly as instantiated in pob_linear_forward's switch
CONFIGS = [(128, 32, 16, 1, 4, 4), (64, 32, 16, 2, 4, 4), (32, 32, 32, 4, 4, 4), (16, 32, 32, 8, 4, 4), (64, 64, 16, 1, 4, 4),
           (32, 64, 32, 2, 4, 4), (16, 64, 32, 4, 4, 4), (128, 32, 16, 1, 8, 4), (256, 32, 16, 1, 8, 4), (128, 64, 16, 1, 8, 8),
           (64, 64, 16, 2, 8, 8), (64, 64, 32, 4, 8, 8), (32, 64, 32, 8, 8, 8), (128, 64, 16, 2, 8, 8), (64, 128, 16, 2, 8, 8),
           (32, 128, 32, 4, 8, 8)]


def test_configs_match_the_kernel_source():
    src = open(os.path.join(ROOT, "pointcloudpdf_b200", "csrc", "linear.cu")).read()
    import re
    found = [tuple(int(v) for v in m.groups()[1:]) for m in
             re.finditer(r"POB_LINEAR_CASE\((\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+)\)", src)]
    assert found == CONFIGS


@pytest.mark.parametrize("cfg", CONFIGS)
def test_tile_index_arithmetic(cfg):
    run = _emulator()
    for (m, k, n) in [(37, 40, 36), (5, 8, 136)]:        # ragged rows and columns, K not a multiple of the k-tile
        assert run(*cfg, m, k, n, lda=k + 4) < 1e-5


def test_backend_policy():
    from pointcloudpdf_b200 import ptv1
    prev = ptv1._LINEAR_BACKEND
    try:
        ptv1.set_linear_backend("auto")
        assert ptv1._use_pob_linear(80000, 6, 32, True)          # not 16-byte friendly -> only pob handles it in one launch
        assert ptv1._use_pob_linear(80000, 32, 96, False)        # 80 000-row layers
        assert ptv1._use_pob_linear(20000, 64, 192, False)       # tensor-core (3xTF32) tiles beat cuBLAS's f32 GEMM there
        assert ptv1._use_pob_linear(5000, 128, 128, True)        # epilogue: one launch instead of two
        assert not ptv1._use_pob_linear(5000, 128, 384, False)   # plain q/k/v of the deeper stages: cuBLAS
        assert not ptv1._use_pob_linear(312, 512, 512, True)
        ptv1.set_linear_backend("pob")
        assert ptv1._use_pob_linear(312, 512, 1536, False)
        ptv1.set_linear_backend("cublas")
        assert not ptv1._use_pob_linear(80000, 32, 32, True)
        with pytest.raises(ValueError):
            ptv1.set_linear_backend("triton")
    finally:
        ptv1.set_linear_backend(prev)


def test_reference_arm_prints_one_json_line():
    """bench.py --impl reference runs on host cores only (no driver needed) and owns stdout."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-points", "2048"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
