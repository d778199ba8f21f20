"""pointcloudpdf_b200 -- B200-native (sm_100a) implementation of PointCloudPDF's PTv1
point-operator hot path and open-set scoring, behind the reference's ``pointops`` API.

    import pointcloudpdf_b200.pointops as pointops      # or simply ``import pointops`` from the repo root
    from pointcloudpdf_b200.scoring import MaxProbability, fused_scores

The native library (include/pointops_b200.h) is built by ``python -m pointcloudpdf_b200.build``.
"""
__version__ = "0.1.0"
