"""The merged-list FPS protocol at the algorithm level, on the CPU: tools/fps_merge_sim.py restates what csrc/fps_merge.cuh
does per round (local continuation lists per warp, CTA-level top-KC + terminal, merged order, pairwise conflict test, KEEP
policy, one local step per warp and round) and checks the emitted sequence against plain one-at-a-time FPS
(sampling_cuda_kernel.cu:15-129 semantics: d2 in f32 order is irrelevant here, ties broken by the lower index).  This pins
the PROTOCOL; the kernel itself is pinned bit-for-bit against the oracle in tests/test_gpu_parity.py."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("order,d0,dmax,kc", [("hilbert", 1, 2, 4), ("row", 2, 2, 4), ("hilbert", 1, 2, 0)])
def test_merged_list_protocol_emits_the_plain_fps_sequence(order, d0, dmax, kc):
    env = dict(os.environ, ORDER=order, KEEP="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fps_merge_sim.py"), "2400", "32", str(d0), str(dmax), str(kc), "0", "32"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = out.stdout.strip().splitlines()[-1]
    assert "identical=True" in line, line
    chain = float(re.search(r"mean chain=([0-9.]+)", line).group(1))
    assert chain > 3.0, line          # the point of the protocol: several exact samples per exchange
