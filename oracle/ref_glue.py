"""Run the REFERENCE's own Python code -- on CPU on top of the C restatement, or on the GPU on top of
either the reference's own compiled kernels or this repo's drop-in.

TEST INFRASTRUCTURE ONLY (see pointops_oracle.py).  The reference files are read from /root/reference
where it exists (this container) and otherwise from baseline/_ref/, the git-ignored staging copy that
``python -m oracle.stage_reference`` makes of exactly the files listed there (it travels to the GPU box
with gpurun like a built .so; BASELINE.md 3.2 reserves the directory for the reference install).

backend (reference_modules):
  "oracle"   ``pointops`` = the reference's functions package over a ``pointops._C`` stub backed by
             oracle_c.c, on CPU tensors (golden fixtures, CPU baseline);
  "refgpu"   the same package over a ``pointops._C`` stub that calls the reference's OWN kernels
             (oracle/_ref/libpointops_ref.so, compiled unmodified) on CUDA tensors: the GPU "before";
  "product"  ``pointops`` = this repo's drop-in package: the unmodified callers
             (point_transformer_seg.py, pt_v1.py, max_probability_v1m1_base.py) on the new kernels.
  "shim"     the reference's functions package over a COMPILED ``pointops._C`` replacement
             (integration/pointops_C_shim.cpp: pybind shims with the reference's signatures over the C ABI
             of include/pointops_b200.h) -- the binding a maintainer would add to keep ``pointops._C``.

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

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(REPO_ROOT, "baseline", "_ref")
REF_SO = os.path.join(REPO_ROOT, "oracle", "_ref", "libpointops_ref.so")
I64, I32 = ctypes.c_int64, ctypes.c_int


def _has_tree(root: str) -> bool:
    return os.path.isdir(os.path.join(root, "libs", "pointops", "functions")) and \
        os.path.isfile(os.path.join(root, "pointcept", "models", "point_transformer", "point_transformer_seg.py"))


def ref_root():
    """Where the reference's python files are read from: the mounted tree, else the staged copy."""
    for cand in (os.environ.get("POINTCLOUDPDF_REFERENCE"), "/root/reference", STAGED_ROOT):
        if cand and _has_tree(cand):
            return cand
    return None


REF_ROOT = ref_root() or "/root/reference"


def available() -> bool:
    return ref_root() is not None


def _p(t):
    assert t.device.type == "cpu" and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _make_C_stub_refgpu() -> types.ModuleType:
    """pointops._C over the reference's own kernels (oracle/_ref/libpointops_ref.so): the pybind shims of
    src/*/*_cuda.cpp only unwrap data pointers and call the extern "C" launchers (src/pointops_api.cpp:15-32);
    this does the same through ctypes.  The launchers use the legacy default stream, so must the caller."""
    if not os.path.exists(REF_SO):
        raise RuntimeError("oracle/_ref/libpointops_ref.so not built (make -C oracle ref needs /root/reference)")
    L = ctypes.CDLL(REF_SO)
    C = types.ModuleType("pointops._C")
    I = ctypes.c_int

    def g(t):
        assert t.is_cuda and t.is_contiguous(), "the reference launchers take contiguous CUDA tensors"
        return ctypes.c_void_p(t.data_ptr())

    def knn_query_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
        L.knn_query_cuda_launcher(I(m), I(nsample), g(xyz), g(new_xyz), g(offset), g(new_offset), g(idx), g(dist2))

    def farthest_point_sampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
        L.farthest_point_sampling_cuda_launcher(I(b), I(int(n_max)), g(xyz), g(offset), g(new_offset), g(tmp), g(idx))

    def grouping_forward_cuda(m, nsample, c, input, idx, output):
        L.grouping_forward_cuda_launcher(I(m), I(nsample), I(c), g(input), g(idx), g(output))

    def grouping_backward_cuda(m, nsample, c, grad_output, idx, grad_input):
        L.grouping_backward_cuda_launcher(I(m), I(nsample), I(c), g(grad_output.contiguous()), g(idx), g(grad_input))

    def subtraction_forward_cuda(n, nsample, c, input1, input2, idx, output):
        L.subtraction_forward_cuda_launcher(I(n), I(nsample), I(c), g(input1), g(input2), g(idx), g(output))

    def subtraction_backward_cuda(n, nsample, c, idx, grad_output, grad_input1, grad_input2):
        L.subtraction_backward_cuda_launcher(I(n), I(nsample), I(c), g(idx), g(grad_output.contiguous()), g(grad_input1), g(grad_input2))

    def aggregation_forward_cuda(n, nsample, c, w_c, input, position, weight, idx, output):
        L.aggregation_forward_cuda_launcher(I(n), I(nsample), I(c), I(w_c), g(input), g(position), g(weight), g(idx), g(output))

    def aggregation_backward_cuda(n, nsample, c, w_c, input, position, weight, idx, grad_output, grad_input,
                                  grad_position, grad_weight):
        L.aggregation_backward_cuda_launcher(I(n), I(nsample), I(c), I(w_c), g(input), g(position), g(weight), g(idx),
                                             g(grad_output.contiguous()), g(grad_input), g(grad_position), g(grad_weight))

    def interpolation_forward_cuda(n, c, k, input, idx, weight, output):
        L.interpolation_forward_cuda_launcher(I(n), I(c), I(k), g(input), g(idx), g(weight), g(output))

    def interpolation_backward_cuda(n, c, k, grad_output, idx, weight, grad_input):
        L.interpolation_backward_cuda_launcher(I(n), I(c), I(k), g(grad_output.contiguous()), g(idx), g(weight), g(grad_input))

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
def reference_modules(backend: str = "oracle"):
    """Context in which ``import pointops`` resolves to the chosen backend (module docstring) and
    ``pointcept.models.point_transformer`` / recognizers are importable from the unmodified files.
    Yields a namespace with .pointops, .ptseg (point_transformer_seg), .msp, .pt_rec."""
    REF_ROOT = ref_root()
    if REF_ROOT is None:
        raise RuntimeError("reference python files found neither at /root/reference nor under baseline/_ref "
                           "(python -m oracle.stage_reference stages them where the reference is mounted)")
    if backend not in ("oracle", "refgpu", "product", "shim"):
        raise ValueError(backend)
    saved_modules = {k: v for k, v in sys.modules.items() if k == "pointops" or k.startswith("pointops.")
                     or k == "pointcept" or k.startswith("pointcept.")}
    for k in saved_modules:
        del sys.modules[k]
    saved_ctor = (torch.cuda.IntTensor, torch.cuda.FloatTensor)
    saved_sqrt = torch.sqrt
    if backend == "oracle":
        torch.cuda.IntTensor = _CpuTensorCtor(torch.int32)
        torch.cuda.FloatTensor = _CpuTensorCtor(torch.float32)
        # The reference only ever calls torch.sqrt on CUDA tensors (IEEE-correct); torch's CPU float
        # sqrt is an inexact SIMD routine, so give the wrappers the correctly rounded one here.
        torch.sqrt = lambda x, *a, **k: O.sqrt_f32(x) if (x.dtype == torch.float32 and not a and not k) else saved_sqrt(x, *a, **k)
    try:
        if backend == "product":
            # this repo's drop-in: the root-level `pointops` package (re-export of pointcloudpdf_b200.pointops)
            if REPO_ROOT not in sys.path:
                sys.path.insert(0, REPO_ROOT)
            pointops = importlib.import_module("pointops")
            assert os.path.dirname(os.path.abspath(pointops.__file__)) == os.path.join(REPO_ROOT, "pointops"), pointops.__file__
        else:
            fdir = os.path.join(REF_ROOT, "libs", "pointops", "functions")
            spec = importlib.util.spec_from_file_location("pointops", os.path.join(fdir, "__init__.py"),
                                                          submodule_search_locations=[fdir])
            pointops = importlib.util.module_from_spec(spec)
            sys.modules["pointops"] = pointops
            if backend == "shim":
                if REPO_ROOT not in sys.path:
                    sys.path.insert(0, REPO_ROOT)
                from integration import build_shim
                sys.modules["pointops._C"] = build_shim.load()
            else:
                sys.modules["pointops._C"] = _make_C_stub() if backend == "oracle" else _make_C_stub_refgpu()
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
