"""ctypes binding of libpointops_b200.so (the C ABI of include/pointops_b200.h).

There is no CPU or eager-torch fallback behind this module: if the shared library is missing
or cannot be loaded, every operator raises.  Build it with ``python -m pointcloudpdf_b200.build``
(nvcc, sm_100a) -- ``__graft_entry__.build()`` does that.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_float, c_int, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libpointops_b200.so")

_lock = threading.Lock()
_lib = None

P, I, L, F, Z = c_void_p, c_int, c_int64, c_float, c_size_t

# name -> (restype, argtypes); mirrors include/pointops_b200.h line by line
SIGNATURES = {
    "pob_version": (I, []),
    "pob_kernel_launch_count": (ctypes.c_longlong, []),
    "pob_error_string": (ctypes.c_char_p, [I]),
    "pob_knn_grid_workspace_bytes": (Z, [L, I, F]),
    "pob_knn_grid_build": (I, [L, I, P, P, F, P, Z, P]),
    "pob_knn_grid_query": (I, [L, I, L, I, P, P, P, F, P, P, P, P, I, P, P]),
    "pob_knn_grid_query_scatter": (I, [L, I, L, I, P, P, P, F, P, L, I, P, P, I, P]),
    "pob_knn_query": (I, [L, I, L, I, P, P, P, P, P, P, I, P, Z, P]),
    "pob_knn_query_bruteforce": (I, [L, I, I, P, P, P, P, P, P, I, P, Z, P]),
    "pob_ball_query": (I, [L, I, F, F, L, I, P, P, P, F, P, P, P, P, P]),
    "pob_random_ball_query": (I, [L, I, F, F, L, I, P, P, P, P, F, P, P, P, P, P]),
    "pob_farthest_point_sampling": (I, [I, L, P, P, P, P, P, I, P, L, F, I, P, P]),
    "pob_grouping_forward": (I, [L, I, I, P, P, P, P]),
    "pob_grouping_backward": (I, [L, I, I, P, P, P, P]),
    "pob_subtraction_forward": (I, [L, I, I, P, P, P, P, P]),
    "pob_subtraction_backward": (I, [L, I, I, P, P, P, P, P]),
    "pob_aggregation_forward": (I, [L, I, I, I, P, P, P, P, P, P]),
    "pob_aggregation_backward": (I, [L, I, I, I, P, P, P, P, P, P, P, P, P]),
    "pob_interpolation_forward": (I, [L, I, I, P, P, P, P, P]),
    "pob_interpolation_backward": (I, [L, I, I, P, P, P, P, P]),
    "pob_group_xyz_forward": (I, [L, I, I, I, P, I, P, P, P, P, P]),
    "pob_group_xyz_backward": (I, [L, I, I, I, P, P, P, P]),
    "pob_group_relxyz_forward": (I, [L, I, P, P, P, P, P]),
    "pob_pt_layer_param_floats": (L, [I, I]),
    "pob_pt_layer_forward": (I, [L, I, I, I, P, L, P, L, P, L, P, P, P, I, P, L, I, P]),
    "pob_affine_act": (I, [L, I, P, P, P, P, I, P, P]),
    "pob_transition_down_pool": (I, [L, I, I, P, P, P, P, P, P, P, P, P]),
    "pob_interpolation_add_forward": (I, [L, I, I, P, P, P, P, P, P]),
    "pob_linear_forward": (I, [L, I, I, P, L, P, P, P, L, I, P, L, I, P]),
    "pob_attention_relation_step_forward": (I, [L, I, I, P, P, P, P, P, P, P]),
    "pob_attention_relation_step_backward": (I, [L, I, I, P, P, P, P, P, P, P, P, P, P]),
    "pob_attention_fusion_step_forward": (I, [L, I, I, P, P, P, P, P, P]),
    "pob_attention_fusion_step_backward": (I, [L, I, I, P, P, P, P, P, P, P, P]),
    "pob_grid_hash": (I, [L, P, ctypes.c_double, ctypes.c_double, ctypes.c_double, P, P, P, P]),
    "pob_scatter_mean": (I, [L, I, P, P, L, P, P, P]),
    "pob_score_workspace_bytes": (Z, [I]),
    "pob_score_fused": (I, [L, I, I, P, P, P, F, P, P, P, P, P, P, P, P, P, Z, P]),
}


class PointopsB200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the library once; raise loudly if it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise PointopsB200Error(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                "`python -m pointcloudpdf_b200.build` (needs nvcc; targets sm_100a). "
                "pointcloudpdf_b200 has no CPU or eager fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def error_string(code: int) -> str:
    return load().pob_error_string(int(code)).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise PointopsB200Error(f"{what} failed with code {code}: {error_string(code)}")


def ptr(t) -> c_void_p:
    """Device pointer of a tensor (None -> NULL)."""
    return c_void_p(None) if t is None else c_void_p(t.data_ptr())


_torch = None


def _t():
    global _torch
    if _torch is None:
        import torch
        _torch = torch
    return _torch


def raw_stream(device) -> int:
    """cudaStream_t of torch's current stream on `device`, without building a Stream object
    (torch.cuda.current_stream costs ~9 us of python per call; this is ~0.3 us)."""
    C = _t()._C
    idx = device.index
    return C._cuda_getCurrentRawStream(C._cuda_getDevice() if idx is None else idx)


def current_stream(device) -> c_void_p:
    return c_void_p(raw_stream(device))


_STREAM_OBJS = {}


def current_stream_obj(device=None):
    """torch.cuda.current_stream(device), memoised on the raw handle (pool streams are never destroyed)."""
    torch = _t()
    idx = torch._C._cuda_getDevice() if device is None or device.index is None else device.index
    key = (idx, torch._C._cuda_getCurrentRawStream(idx))
    s = _STREAM_OBJS.get(key)
    if s is None:
        s = _STREAM_OBJS[key] = torch.cuda.current_stream(idx)
    return s


class device_guard:
    """`with torch.cuda.device(dev)` for the common case that dev is already current: two C calls
    instead of ~5 us of python."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index

    def __enter__(self):
        C = _t()._C
        prev = C._cuda_getDevice()
        if self.idx is not None and prev != self.idx:
            C._cuda_setDevice(self.idx)
            self.prev = prev
        else:
            self.prev = -1
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            _t()._C._cuda_setDevice(self.prev)
        return False


# ---------------------------------------------------------------- op-level timing (bench.py) --
class OpProfile:
    """CUDA-event bracket around every C-ABI call, on the stream the kernels are launched on.
    Off by default; bench.py switches it on for the timed region to attribute step time to
    kernels and to compute achieved GB/s from the algorithmic bytes the wrappers report."""

    def __init__(self):
        self.records = []  # (name, start_event, end_event, alg_bytes, alg_flops, unfused_bytes)

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, s, e, nbytes, flops, unfused in self.records:
            d = out.setdefault(name, dict(calls=0, ms=0.0, alg_bytes=0, alg_flops=0, unfused_bytes=0))
            d["unfused_bytes"] += unfused
            d["calls"] += 1
            d["ms"] += s.elapsed_time(e)
            d["alg_bytes"] += nbytes
            d["alg_flops"] += flops
        return out


PROFILE = None  # set to an OpProfile() to record


def run(name: str, *args, alg_bytes: int = 0, alg_flops: int = 0, unfused_bytes: int = 0) -> None:
    """Call entry point `name`, raise on a non-zero status.  The last positional argument is the
    stream; when profiling, events are recorded on torch's current stream (the same one)."""
    fn = getattr(_lib if _lib is not None else load(), name)
    prof = PROFILE
    if prof is None:
        rc = fn(*args)
    else:
        import torch
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        prof.records.append((name, s, e, int(alg_bytes), int(alg_flops), int(unfused_bytes)))
    check(rc, name)


_REPLAYED = 0  # kernels of this library launched through CUDA-graph replays (counted at capture time)


def add_replayed_launches(n: int) -> None:
    global _REPLAYED
    _REPLAYED += int(n)


def launch_count() -> int:
    """Kernels of this library launched so far: direct launches (counted in C) + graph replays."""
    return int(load().pob_kernel_launch_count()) + _REPLAYED
