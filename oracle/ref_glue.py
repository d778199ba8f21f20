"""Run the REFERENCE's own Python code on CPU, on top of the C restatement.

TEST INFRASTRUCTURE ONLY (see pointops_oracle.py).  Works only where
/root/reference exists (this container; never on the GPU box).

The reference package ``libs/pointops/functions`` is imported *unmodified* from
where it lies, under the name ``pointops``; the CUDA extension it binds
(``pointops._C``, src/pointops_api.cpp:15-32) is replaced by a stub whose
``*_cuda`` entry points have the reference's signatures and call oracle_c.c
(literal restatements: heap kNN, block-reduction FPS).  ``torch.cuda.IntTensor``
/ ``FloatTensor`` (used by the wrappers to allocate outputs) are pointed at CPU
constructors for the duration.  The same trick loads
``pointcept/models/point_transformer`` and the recognizers with a 3-line
registry stub, so fixtures under tests/golden/ are outputs of reference code.
"""
from __future__ import annotations

import contextlib
import ctypes
import importlib
import importlib.util
import os
import sys
import types

import torch

from . import pointops_oracle as O

REF_ROOT = os.environ.get("POINTCLOUDPDF_REFERENCE", "/root/reference")
I64, I32 = ctypes.c_int64, ctypes.c_int


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "libs", "pointops", "functions"))


def _p(t):
    assert t.device.type == "cpu" and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _make_C_stub() -> types.ModuleType:
    """pointops._C with the signatures of src/*/*_cuda.cpp (sizes as python ints,
    tensors in the same positions, outputs written in place)."""
    L = O.lib()
    C = types.ModuleType("pointops._C")

    def knn_query_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
        L.oracle_knn_ref_heap(I64(m), I32(nsample), _p(xyz), _p(new_xyz), _p(offset), _p(new_offset),
                              _p(idx), _p(dist2))

    def farthest_point_sampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
        L.oracle_fps_ref_block(I32(b), I32(O._pow2_block(int(n_max))), _p(xyz), _p(offset), _p(new_offset), _p(idx))

    def grouping_forward_cuda(m, nsample, c, input, idx, output):
        L.oracle_grouping_fwd(I64(m), I32(nsample), I32(c), _p(input), _p(idx), _p(output))

    def grouping_backward_cuda(m, nsample, c, grad_output, idx, grad_input):
        L.oracle_grouping_bwd(I64(m), I32(nsample), I32(c), _p(grad_output.contiguous()), _p(idx), _p(grad_input))

    def subtraction_forward_cuda(n, nsample, c, input1, input2, idx, output):
        L.oracle_subtraction_fwd(I64(n), I32(nsample), I32(c), _p(input1), _p(input2), _p(idx), _p(output))

    def subtraction_backward_cuda(n, nsample, c, idx, grad_output, grad_input1, grad_input2):
        L.oracle_subtraction_bwd(I64(n), I32(nsample), I32(c), _p(idx), _p(grad_output.contiguous()),
                                 _p(grad_input1), _p(grad_input2))

    def aggregation_forward_cuda(n, nsample, c, w_c, input, position, weight, idx, output):
        L.oracle_aggregation_fwd(I64(n), I32(nsample), I32(c), I32(w_c), _p(input), _p(position), _p(weight),
                                 _p(idx), _p(output))

    def aggregation_backward_cuda(n, nsample, c, w_c, input, position, weight, idx, grad_output, grad_input,
                                  grad_position, grad_weight):
        L.oracle_aggregation_bwd(I64(n), I32(nsample), I32(c), I32(w_c), _p(input), _p(position), _p(weight),
                                 _p(idx), _p(grad_output.contiguous()), _p(grad_input), _p(grad_position),
                                 _p(grad_weight))

    def interpolation_forward_cuda(n, c, k, input, idx, weight, output):
        L.oracle_interpolation_fwd(I64(n), I32(c), I32(k), _p(input), _p(idx), _p(weight), _p(output))

    def interpolation_backward_cuda(n, c, k, grad_output, idx, weight, grad_input):
        L.oracle_interpolation_bwd(I64(n), I32(c), I32(k), _p(grad_output.contiguous()), _p(idx), _p(weight),
                                   _p(grad_input))

    def _unsupported(*a, **k):
        raise NotImplementedError("not on the PTv1 hot path (SURVEY.md section 8 f-4)")

    for name, fn in list(locals().items()):
        if name.endswith("_cuda"):
            setattr(C, name, fn)
    for name in ("ball_query_cuda", "random_ball_query_cuda", "attention_relation_step_forward_cuda",
                 "attention_relation_step_backward_cuda", "attention_fusion_step_forward_cuda",
                 "attention_fusion_step_backward_cuda"):
        setattr(C, name, _unsupported)
    return C


class _CpuTensorCtor:
    """torch.cuda.IntTensor(...) / FloatTensor(...) stand-in: same call forms, CPU result."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __call__(self, *args):
        if len(args) == 1 and isinstance(args[0], (list, tuple)):
            return torch.tensor(args[0], dtype=self.dtype)
        if len(args) == 1 and isinstance(args[0], torch.Tensor):
            return torch.empty(int(args[0]), dtype=self.dtype)
        return torch.empty(*[int(a) for a in args], dtype=self.dtype)


@contextlib.contextmanager
def reference_modules():
    """Context in which ``import pointops`` is the reference's functions package over the
    stub, and ``pointcept.models.point_transformer`` / recognizers are importable.
    Yields a namespace with .pointops, .ptseg (point_transformer_seg), .msp, .pt_rec."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    saved_modules = {k: v for k, v in sys.modules.items() if k == "pointops" or k.startswith("pointops.")
                     or k == "pointcept" or k.startswith("pointcept.")}
    for k in saved_modules:
        del sys.modules[k]
    saved_ctor = (torch.cuda.IntTensor, torch.cuda.FloatTensor)
    saved_sqrt = torch.sqrt
    torch.cuda.IntTensor = _CpuTensorCtor(torch.int32)
    torch.cuda.FloatTensor = _CpuTensorCtor(torch.float32)
    # The reference only ever calls torch.sqrt on CUDA tensors (IEEE-correct); torch's CPU float
    # sqrt is an inexact SIMD routine, so give the wrappers the correctly rounded one here.
    torch.sqrt = lambda x, *a, **k: O.sqrt_f32(x) if (x.dtype == torch.float32 and not a and not k) else saved_sqrt(x, *a, **k)
    try:
        fdir = os.path.join(REF_ROOT, "libs", "pointops", "functions")
        spec = importlib.util.spec_from_file_location("pointops", os.path.join(fdir, "__init__.py"),
                                                      submodule_search_locations=[fdir])
        pointops = importlib.util.module_from_spec(spec)
        sys.modules["pointops"] = pointops
        sys.modules["pointops._C"] = _make_C_stub()
        spec.loader.exec_module(pointops)

        # registry stubs for pointcept.models.builder / pointcept.recognizers.builder
        class _Reg:
            def register_module(self, *a, **k):
                return lambda cls: cls

        def _pkg(name, path=None):
            mod = types.ModuleType(name)
            mod.__path__ = [path] if path else []
            sys.modules[name] = mod
            return mod

        pc = os.path.join(REF_ROOT, "pointcept")
        _pkg("pointcept", pc)
        _pkg("pointcept.models", os.path.join(pc, "models"))
        b = _pkg("pointcept.models.builder")
        b.MODELS = _Reg()
        _pkg("pointcept.recognizers", os.path.join(pc, "recognizers"))
        rb = _pkg("pointcept.recognizers.builder")
        rb.RECOGNIZER = _Reg()

        def _load(name, relpath, is_pkg=False):
            path = os.path.join(pc, relpath)
            kw = dict(submodule_search_locations=[os.path.dirname(path)]) if is_pkg else {}
            sp = importlib.util.spec_from_file_location(name, path, **kw)
            mod = importlib.util.module_from_spec(sp)
            sys.modules[name] = mod
            sp.loader.exec_module(mod)
            return mod

        pt_pkg = _pkg("pointcept.models.point_transformer", os.path.join(pc, "models", "point_transformer"))
        _load("pointcept.models.point_transformer.utils", "models/point_transformer/utils.py")
        ptseg = _load("pointcept.models.point_transformer.point_transformer_seg",
                      "models/point_transformer/point_transformer_seg.py")
        pt_pkg.TransitionUp, pt_pkg.Bottleneck = ptseg.TransitionUp, ptseg.Bottleneck
        _pkg("pointcept.recognizers.max_probability", os.path.join(pc, "recognizers", "max_probability"))
        msp = _load("pointcept.recognizers.max_probability.max_probability_v1m1_base",
                    "recognizers/max_probability/max_probability_v1m1_base.py")
        _pkg("pointcept.recognizers.recognizer_model", os.path.join(pc, "recognizers", "recognizer_model"))
        pt_rec = _load("pointcept.recognizers.recognizer_model.pt_v1", "recognizers/recognizer_model/pt_v1.py")
        ns = types.SimpleNamespace(pointops=pointops, ptseg=ptseg, msp=msp, pt_rec=pt_rec)
        yield ns
    finally:
        torch.cuda.IntTensor, torch.cuda.FloatTensor = saved_ctor
        torch.sqrt = saved_sqrt
        for k in [k for k in sys.modules if k == "pointops" or k.startswith("pointops.")
                  or k == "pointcept" or k.startswith("pointcept.")]:
            del sys.modules[k]
        sys.modules.update(saved_modules)
