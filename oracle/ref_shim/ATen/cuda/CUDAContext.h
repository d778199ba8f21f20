// Build shim, see ../../torch/serialize/tensor.h
#pragma once
#include <cuda_runtime.h>
