"""Stand-in for the one torch-points-kernels call the PDF recognizer makes:

    import torch_points_kernels as tp                      # pointcept/recognizers/ours/pointpdf_v1m1_base.py:6
    tp.ball_query(radius, max_neighbor, coord, coord, mode="partial_dense", batch_x=batch, batch_y=batch)[0]   # :121-129

torch-points-kernels is neither vendored nor pinned by the reference (and absent from this image).  Registering this
module under that name (``sys.modules["torch_points_kernels"] = pointcloudpdf_b200.tp_compat``) serves the call from the
kNN search grid of this library; only the mode the reference uses is implemented, anything else raises."""
from .pseudo import ball_query_partial_dense


def ball_query(radius, nsample, x, y, mode="dense", batch_x=None, batch_y=None, sort=False):
    if mode.lower() != "partial_dense":
        raise NotImplementedError("tp_compat.ball_query: only mode='partial_dense' (the PDF recognizer's call) is provided")
    if sort:
        raise NotImplementedError("tp_compat.ball_query: sort=True is not used by the reference and not provided")
    return ball_query_partial_dense(radius, nsample, x, y, batch_x, batch_y)
