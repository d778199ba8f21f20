"""Host-side checks of the drop-in loader (no GPU): the reference's unmodified caller files import against this
repo's ``pointops`` package, and the staged copy under baseline/_ref is byte-identical to the mounted reference."""
import hashlib
import json
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def glue():
    from oracle import ref_glue
    if not ref_glue.available():
        pytest.skip("reference python files neither mounted nor staged")
    return ref_glue


def test_reference_callers_import_over_the_product_package(glue):
    with glue.reference_modules("product") as R:
        assert os.path.dirname(os.path.abspath(R.pointops.__file__)) == os.path.join(ROOT, "pointops")
        for name in ("knn_query", "farthest_point_sampling", "grouping", "knn_query_and_group", "interpolation",
                     "aggregation", "subtraction", "grouping2", "interpolation2", "query_and_group", "offset2batch", "batch2offset"):
            assert hasattr(R.pointops, name), name
        torch.manual_seed(2024)
        model = R.ptseg.PointTransformerSeg50(in_channels=6, num_classes=13)
        # the module under test is the reference's file, not the mirror
        assert type(model).__module__ == "pointcept.models.point_transformer.point_transformer_seg"
        from pointcloudpdf_b200.ptv1 import PointTransformerSeg50 as Mirror
        torch.manual_seed(2024)
        mirror = Mirror(in_channels=6, num_classes=13)
        a, b = model.state_dict(), mirror.state_dict()
        assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
        # no CPU path: the unmodified caller on CPU tensors must fail loudly in the drop-in, not fall back
        x = torch.rand(64, 3)
        with pytest.raises((ValueError, RuntimeError)):
            R.pointops.knn_query(4, x, torch.tensor([64], dtype=torch.int32))


def test_staged_copy_is_byte_identical_to_the_reference(glue):
    staged = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir("/root/reference") or not os.path.exists(os.path.join(staged, "MANIFEST.json")):
        pytest.skip("needs both the mounted reference and the staged copy")
    manifest = json.load(open(os.path.join(staged, "MANIFEST.json")))
    assert len(manifest) >= 13
    for rel, digest in manifest.items():
        assert hashlib.sha256(open(os.path.join("/root/reference", rel), "rb").read()).hexdigest() == digest, rel
        assert hashlib.sha256(open(os.path.join(staged, rel), "rb").read()).hexdigest() == digest, rel
