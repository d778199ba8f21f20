// api.cu -- version and error text for the C ABI declared in include/pointops_b200.h.
#include "common.cuh"

#include <atomic>

static std::atomic<long long> g_launches{0};

extern "C" void pob_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

POB_API int pob_version(void) { return 1; }

POB_API long long pob_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

POB_API const char* pob_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case POB_ERR_BAD_ARG: return "pointops_b200: bad argument (null pointer, negative size or unsupported value)";
        case POB_ERR_WORKSPACE: return "pointops_b200: workspace too small";
        case POB_ERR_UNSUPPORTED: return "pointops_b200: unsupported configuration";
        default: return cudaGetErrorString((cudaError_t)code);
    }
}
