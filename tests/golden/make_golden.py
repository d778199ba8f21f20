"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own Python code.

Run here (needs /root/reference; never on the GPU box):   python tests/golden/make_golden.py

The reference's ``libs/pointops/functions`` wrappers, ``PointTransformerSeg`` and recognizers are
imported unmodified (oracle/ref_glue.py) over a ``pointops._C`` stub that calls the literal C
restatement of the CUDA kernels (oracle/oracle_c.c), which is itself pinned bit-for-bit against
the compiled reference kernels on the GPU box (tests/test_gpu_reference_ext.py).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_glue  # noqa: E402
from pointcloudpdf_b200 import synthetic as S  # noqa: E402


def ops_small(R):
    g = torch.Generator().manual_seed(2024)
    sizes = [300, 5, 120]
    xyz = torch.rand(sum(sizes), 3, generator=g) * torch.tensor([3.0, 2.0, 1.5])
    offset = torch.tensor([300, 305, 425], dtype=torch.int32)
    new_offset = torch.tensor([75, 76, 106], dtype=torch.int32)
    feat = torch.randn(425, 8, generator=g)
    feat2 = torch.randn(425, 8, generator=g)
    pos = torch.randn(425, 16, 8, generator=g)
    w = torch.randn(425, 16, 2, generator=g)
    po = R.pointops
    G = dict(xyz=xyz, offset=offset, new_offset=new_offset, feat=feat, feat2=feat2, pos=pos, w=w)
    G["knn_idx"], G["knn_dist"] = po.knn_query(16, xyz, offset)
    G["fps_idx"] = po.farthest_point_sampling(xyz, offset, new_offset)
    G["grouped_xyz"] = po.grouping(G["knn_idx"], feat, xyz, xyz, with_xyz=True)
    n_p = xyz[G["fps_idx"].long()].contiguous()
    G["cross_grouped"], G["cross_idx"] = po.knn_query_and_group(feat, xyz, offset, n_p, new_offset, nsample=16,
                                                                with_xyz=True)
    G["interp"] = po.interpolation(n_p, xyz, feat[G["fps_idx"].long()].contiguous(), new_offset, offset)
    safe = G["knn_idx"].clamp(min=0)
    # aggregation / subtraction / grouping2 read out of bounds on -1 in the reference: use idx >= 0
    G["aggregation"] = po.aggregation(feat, pos, w, safe)
    G["subtraction"] = po.subtraction(feat, feat2, safe)
    G["grouping2"] = po.grouping2(feat, safe)
    G["qag"], G["qag_idx"] = po.query_and_group(8, xyz, xyz, feat, None, offset, offset, dilation=1)
    torch.save(G, os.path.join(HERE, "ops_small.pt"))
    return G


def scores_small(R):
    logits, conf, unknown, label = S.openset_logits(2000, 13, seed=2024)
    rec = R.msp.MaxProbability(method="msp")
    rec.model_hooks = {"backbone": {"forward_output": logits}}
    msp = rec({})["score"]
    rec2 = R.msp.MaxProbability(method="max_logits")
    rec2.model_hooks = {"backbone": {"forward_output": logits}}
    ml = rec2({})["score"]
    torch.save(dict(logits=logits, conf=conf, msp=msp, ml=ml), os.path.join(HERE, "scores_small.pt"))


def ptv1_small(R):
    """Unmodified reference PointTransformerSeg50 + MaxProbability + PTRecognizer + the PDF score
    line, eval mode, default init under fixed seeds (weights are NOT stored: the mirror in
    pointcloudpdf_b200/ptv1.py reproduces them from the same seeds, see tests/test_cpu_ptv1.py)."""
    batch = S.s3dis_batch([2500, 900], seed=2024)
    torch.manual_seed(2024)
    model = R.ptseg.PointTransformerSeg50(in_channels=6, num_classes=13).eval()
    torch.manual_seed(2025)
    rec_model = R.pt_rec.PTRecognizer().eval()
    hooks = {}

    def tap(name, mod):
        mod.register_forward_hook(lambda m, i, o: hooks.__setitem__(name, {"forward_output": o}))

    for i in range(1, 6):
        tap(f"backbone.enc{i}", getattr(model, f"enc{i}"))
        tap(f"backbone.dec{i}.1", getattr(model, f"dec{i}")[1])
    with torch.no_grad():
        logits = model(dict(coord=batch["coord"], feat=batch["feat"], offset=batch["offset"]))
        hooks["backbone"] = {"forward_output": logits}
        rec = R.msp.MaxProbability(method="msp")
        rec.model_hooks = hooks
        msp = rec({})["score"]
        conf = rec_model(hooks)
        pdf = torch.cat([logits, conf], -1).softmax(-1)[:, -1]  # pointpdf_v1m1_base.py:110-113
    torch.save(dict(coord=batch["coord"], feat=batch["feat"], offset=batch["offset"], logits=logits, msp=msp,
                    conf=conf, pdf=pdf, seeds=(2024, 2025)), os.path.join(HERE, "ptv1_small.pt"))
    return model, rec_model


def pseudo_small():
    """Region growth of PointPdfV1.pseudo_labeling, produced by the REFERENCE's own function (oracle/pseudo_oracle.py
    ::reference_growth runs the unmodified staticmethod and captures the region at the end of its growth loop).  Scenes
    carry a planted low-confidence blob so that the region really grows (hundreds of points, dozens of iterations)."""
    from oracle import pseudo_oracle as PO
    cases = []
    for n, seed, radius, sw in ((3000, 31, 0.06, False), (3000, 31, 0.06, True), (4000, 32, 0.05, False), (2500, 33, 0.07, True)):
        b = S.scannet_batch([n], seed=seed)
        coord = b["coord"]
        g = torch.Generator().manual_seed(seed)
        label = torch.randint(0, 20, (n,), generator=g)
        logits = torch.randn(n, 20, generator=g) * 0.7
        logits[torch.arange(n), label] += 5.0
        centre = coord[torch.randint(0, n, (1,), generator=g)]
        blob = (coord - centre).norm(dim=-1) < 0.3 * float((coord.max(0)[0] - coord.min(0)[0]).max())
        logits[blob] = torch.randn(int(blob.sum()), 20, generator=g) * 0.3     # flat logits: low msp, low max logit
        batch = torch.zeros(n, dtype=torch.long)
        nbrs, _ = PO.ball_query_partial_dense(radius, 16, coord, coord, batch, batch)
        region = PO.reference_growth(coord, logits, nbrs, "msp", 1.0, "ml", 0.15, 30, sw, seed=1000 + seed)
        # neighbours are not stored (the oracle recomputes them from coord / radius in a second)
        cases.append(dict(coord=coord, logits=logits.half().float() if False else logits, radius=radius, max_neighbor=16, condition_from="msp",
                          beta=1.0, seed_from="ml", seed_range=0.15, num_seed=30, slide_window=sw, torch_seed=1000 + seed,
                          region=region, blob=int(blob.sum())))
        print("pseudo_small", n, sw, "blob", int(blob.sum()), "region", region.numel())
    torch.save(cases, os.path.join(HERE, "pseudo_small.pt"))


if __name__ == "__main__":
    with ref_glue.reference_modules() as R:
        ops_small(R)
        scores_small(R)
        ptv1_small(R)
    pseudo_small()
    print("golden fixtures written to", HERE)
