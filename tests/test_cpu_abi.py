"""CPU tests of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports
every symbol include/pointops_b200.h declares; the ctypes table matches the header; the Python
package exposes the reference's names; product code never touches the oracle."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pointops_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|int64_t|size_t|long long|const char\*)\s+(pob_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(2)] = n
    return decls


@pytest.fixture(scope="module")
def lib():
    from pointcloudpdf_b200 import build, _lib
    build.build()           # nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def test_header_declares_the_hot_path():
    d = declared_functions()
    for name in ("pob_knn_query", "pob_farthest_point_sampling", "pob_grouping_forward", "pob_grouping_backward",
                 "pob_subtraction_forward", "pob_subtraction_backward", "pob_aggregation_forward",
                 "pob_aggregation_backward", "pob_interpolation_forward", "pob_interpolation_backward",
                 "pob_group_xyz_forward", "pob_group_xyz_backward", "pob_score_fused"):
        assert name in d


def test_library_exports_every_declared_symbol(lib):
    from pointcloudpdf_b200 import _lib
    d = declared_functions()
    assert len(d) >= 20
    for name, nargs in d.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} missing from the ctypes table"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: ctypes table has the wrong arity"
    assert set(_lib.SIGNATURES) == set(d)


def test_no_compute_needed_entry_points(lib):
    assert lib.pob_version() == 1
    assert b"success" in lib.pob_error_string(0)
    assert b"workspace" in lib.pob_error_string(10002)
    assert lib.pob_knn_grid_workspace_bytes(80000, 1, 2.0) > 80000 * 20
    assert lib.pob_knn_grid_workspace_bytes(-1, 1, 2.0) == 0
    assert lib.pob_score_workspace_bytes(4) == 64 * (4 + 1)   # b scene slots + the finished-blocks counter


def test_python_surface_matches_reference_names():
    import pointops
    names = ["knn_query", "ball_query", "random_ball_query", "farthest_point_sampling", "grouping", "grouping2",
             "interpolation", "interpolation2", "subtraction", "aggregation", "attention_relation_step",
             "attention_fusion_step", "query_and_group", "knn_query_and_group", "ball_query_and_group",
             "batch2offset", "offset2batch", "furthestsampling", "knnquery", "queryandgroup"]
    for n in names:
        assert callable(getattr(pointops, n)), n


def test_cpu_tensors_are_rejected_not_silently_computed():
    import pointops
    xyz = torch.rand(10, 3)
    off = torch.tensor([10], dtype=torch.int32)
    with pytest.raises(ValueError, match="CUDA"):
        pointops.knn_query(3, xyz, off)
    with pytest.raises(ValueError, match="CUDA"):
        pointops.farthest_point_sampling(xyz, off, off)
    with pytest.raises(ValueError, match="CUDA"):
        pointops.grouping(torch.zeros(10, 3, dtype=torch.int32), torch.rand(10, 4), xyz)
    from pointcloudpdf_b200.scoring import fused_scores
    with pytest.raises(ValueError, match="CUDA"):
        fused_scores(torch.rand(5, 13))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pointcloudpdf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PointopsB200Error, match="no CPU or eager fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    bad = []
    for base in ("pointcloudpdf_b200", "pointops"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|oracle_c|liboracle|_ref/", txt, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_offset_helpers():
    from pointcloudpdf_b200.pointops import _common as C
    assert C.scene_sizes([5, 12, 12, 20]) == [5, 7, 0, 8]


def test_fps_variant_constants_match_the_header():
    """`variant` is a per-call argument of pob_farthest_point_sampling: the python names must be the header's numbers."""
    import re
    from pointcloudpdf_b200.pointops import sampling
    text = open(os.path.join(ROOT, "include", "pointops_b200.h")).read()
    header = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"#define\s+POB_FPS_(\w+)\s+(\d+)", text)}
    assert header == sampling.VARIANTS, (header, sampling.VARIANTS)
