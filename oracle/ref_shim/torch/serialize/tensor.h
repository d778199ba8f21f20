// Build shim (test infrastructure, not product): lets the reference's
// *_cuda_kernel.cu files compile with plain nvcc, without the torch header
// tree. Their headers only *declare* at::Tensor-taking wrappers; the kernels
// and extern "C" launchers we call never touch the type.
#pragma once
namespace at { class Tensor {}; }
