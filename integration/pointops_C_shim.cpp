// pointops._C over libpointops_b200.so -- the pybind module the reference's libs/pointops/functions/*.py import
// (`from pointops._C import knn_query_cuda, ...`), for a maintainer who wants to keep `pointops._C` and the
// reference's own python wrappers and swap only the native library underneath.
//
// Each function has the signature of the reference's shim (libs/pointops/src/*/ *_cuda.cpp, registered in
// src/pointops_api.cpp:15-32): sizes as ints, tensors in the same positions, outputs written in place.  The
// body unwraps data pointers and calls the C ABI of include/pointops_b200.h on torch's current stream.
// Built by integration/build_shim.py (torch.utils.cpp_extension, links libpointops_b200.so); exercised by
// tests/test_cpu_shim.py (compiles, exports every name) and tests/test_gpu_dropin.py (the reference's unmodified
// functions package over it, on the GPU).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "pointops_b200.h"

namespace {

inline cudaStream_t stream_of(const at::Tensor& t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
inline void check(int rc, const char* what) { TORCH_CHECK(rc == 0, what, ": ", pob_error_string(rc)); }

// knn_query_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2)   src/knn_query/knn_query_cuda.cpp:7-16
void knn_query_cuda(int m, int nsample, at::Tensor xyz, at::Tensor new_xyz, at::Tensor offset, at::Tensor new_offset,
                    at::Tensor idx, at::Tensor dist2) {
    const c10::cuda::CUDAGuard guard(xyz.device());
    const int64_t n = xyz.size(0);
    const int b = (int)offset.size(0);
    auto ws = at::empty({(int64_t)pob_knn_grid_workspace_bytes(n, b, 2.0f)}, xyz.options().dtype(at::kByte));
    check(pob_knn_query(m, nsample, n, b, xyz.data_ptr<float>(), new_xyz.data_ptr<float>(), offset.data_ptr<int>(),
                        new_offset.data_ptr<int>(), idx.data_ptr<int>(), dist2.data_ptr<float>(), /*take_sqrt=*/0,
                        ws.data_ptr(), (size_t)ws.numel(), stream_of(xyz)), "knn_query_cuda");
}

// farthest_point_sampling_cuda(b, n, xyz, offset, new_offset, tmp, idx)      src/sampling/sampling_cuda.cpp:7-15
void farthest_point_sampling_cuda(int b, int n, at::Tensor xyz, at::Tensor offset, at::Tensor new_offset, at::Tensor tmp,
                                  at::Tensor idx) {
    const c10::cuda::CUDAGuard guard(xyz.device());
    check(pob_farthest_point_sampling(b, n, xyz.data_ptr<float>(), offset.data_ptr<int>(), new_offset.data_ptr<int>(),
                                      tmp.data_ptr<float>(), idx.data_ptr<int>(), /*cluster_hint=*/0, /*grid=*/nullptr, 0, 0.f,
                                      POB_FPS_AUTO, /*stats=*/nullptr, stream_of(xyz)), "farthest_point_sampling_cuda");
}

// src/grouping/grouping_cuda.cpp:7-24
void grouping_forward_cuda(int m, int nsample, int c, at::Tensor input, at::Tensor idx, at::Tensor output) {
    const c10::cuda::CUDAGuard guard(input.device());
    check(pob_grouping_forward(m, nsample, c, input.data_ptr<float>(), idx.data_ptr<int>(), output.data_ptr<float>(),
                               stream_of(input)), "grouping_forward_cuda");
}
void grouping_backward_cuda(int m, int nsample, int c, at::Tensor grad_output, at::Tensor idx, at::Tensor grad_input) {
    const c10::cuda::CUDAGuard guard(idx.device());
    auto g = grad_output.contiguous();
    check(pob_grouping_backward(m, nsample, c, g.data_ptr<float>(), idx.data_ptr<int>(), grad_input.data_ptr<float>(),
                                stream_of(idx)), "grouping_backward_cuda");
}

// src/subtraction/subtraction_cuda.cpp:7-24
void subtraction_forward_cuda(int n, int nsample, int c, at::Tensor input1, at::Tensor input2, at::Tensor idx, at::Tensor output) {
    const c10::cuda::CUDAGuard guard(input1.device());
    check(pob_subtraction_forward(n, nsample, c, input1.data_ptr<float>(), input2.data_ptr<float>(), idx.data_ptr<int>(),
                                  output.data_ptr<float>(), stream_of(input1)), "subtraction_forward_cuda");
}
void subtraction_backward_cuda(int n, int nsample, int c, at::Tensor idx, at::Tensor grad_output, at::Tensor grad_input1,
                               at::Tensor grad_input2) {
    const c10::cuda::CUDAGuard guard(idx.device());
    auto g = grad_output.contiguous();
    check(pob_subtraction_backward(n, nsample, c, idx.data_ptr<int>(), g.data_ptr<float>(), grad_input1.data_ptr<float>(),
                                   grad_input2.data_ptr<float>(), stream_of(idx)), "subtraction_backward_cuda");
}

// src/aggregation/aggregation_cuda.cpp:7-28
void aggregation_forward_cuda(int n, int nsample, int c, int w_c, at::Tensor input, at::Tensor position, at::Tensor weight,
                              at::Tensor idx, at::Tensor output) {
    const c10::cuda::CUDAGuard guard(input.device());
    check(pob_aggregation_forward(n, nsample, c, w_c, input.data_ptr<float>(), position.data_ptr<float>(),
                                  weight.data_ptr<float>(), idx.data_ptr<int>(), output.data_ptr<float>(), stream_of(input)),
          "aggregation_forward_cuda");
}
void aggregation_backward_cuda(int n, int nsample, int c, int w_c, at::Tensor input, at::Tensor position, at::Tensor weight,
                               at::Tensor idx, at::Tensor grad_output, at::Tensor grad_input, at::Tensor grad_position,
                               at::Tensor grad_weight) {
    const c10::cuda::CUDAGuard guard(input.device());
    auto g = grad_output.contiguous();
    check(pob_aggregation_backward(n, nsample, c, w_c, input.data_ptr<float>(), position.data_ptr<float>(),
                                   weight.data_ptr<float>(), idx.data_ptr<int>(), g.data_ptr<float>(),
                                   grad_input.data_ptr<float>(), grad_position.data_ptr<float>(), grad_weight.data_ptr<float>(),
                                   stream_of(input)), "aggregation_backward_cuda");
}

// src/interpolation/interpolation_cuda.cpp:7-24
void interpolation_forward_cuda(int n, int c, int k, at::Tensor input, at::Tensor idx, at::Tensor weight, at::Tensor output) {
    const c10::cuda::CUDAGuard guard(input.device());
    check(pob_interpolation_forward(n, c, k, input.data_ptr<float>(), idx.data_ptr<int>(), weight.data_ptr<float>(),
                                    output.data_ptr<float>(), stream_of(input)), "interpolation_forward_cuda");
}
void interpolation_backward_cuda(int n, int c, int k, at::Tensor grad_output, at::Tensor idx, at::Tensor weight,
                                 at::Tensor grad_input) {
    const c10::cuda::CUDAGuard guard(idx.device());
    auto g = grad_output.contiguous();
    check(pob_interpolation_backward(n, c, k, g.data_ptr<float>(), idx.data_ptr<int>(), weight.data_ptr<float>(),
                                     grad_input.data_ptr<float>(), stream_of(idx)), "interpolation_backward_cuda");
}

void not_on_this_path(py::args, py::kwargs) {
    TORCH_CHECK(false, "pointops._C shim: ball_query / random_ball_query / attention_* are served by the python package "
                       "(pointcloudpdf_b200.pointops), not by this shim");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {   // names of src/pointops_api.cpp:15-32
    m.def("knn_query_cuda", &knn_query_cuda, "knn_query_cuda");
    m.def("farthest_point_sampling_cuda", &farthest_point_sampling_cuda, "farthest_point_sampling_cuda");
    m.def("grouping_forward_cuda", &grouping_forward_cuda, "grouping_forward_cuda");
    m.def("grouping_backward_cuda", &grouping_backward_cuda, "grouping_backward_cuda");
    m.def("subtraction_forward_cuda", &subtraction_forward_cuda, "subtraction_forward_cuda");
    m.def("subtraction_backward_cuda", &subtraction_backward_cuda, "subtraction_backward_cuda");
    m.def("aggregation_forward_cuda", &aggregation_forward_cuda, "aggregation_forward_cuda");
    m.def("aggregation_backward_cuda", &aggregation_backward_cuda, "aggregation_backward_cuda");
    m.def("interpolation_forward_cuda", &interpolation_forward_cuda, "interpolation_forward_cuda");
    m.def("interpolation_backward_cuda", &interpolation_backward_cuda, "interpolation_backward_cuda");
    for (const char* name : {"ball_query_cuda", "random_ball_query_cuda", "attention_relation_step_forward_cuda",
                             "attention_relation_step_backward_cuda", "attention_fusion_step_forward_cuda",
                             "attention_fusion_step_backward_cuda"})
        m.def(name, &not_on_this_path);
}
