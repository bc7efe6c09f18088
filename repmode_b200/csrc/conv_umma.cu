// tcgen05 implicit-GEMM conv kernels (placeholder until the hardware probe settles the descriptor layout).
#include "common.cuh"
namespace mode {
bool conv3d_umma_supported(int, int, int, int, int) { return false; }
int conv3d_umma(const __half*, const __half*, const int32_t*, float*, int, int, int, int, int, int, float, const float*, double*,
                cudaStream_t) { MODE_FAIL("conv3d_umma: not built"); }
bool wgrad_umma_supported(int, int, int, int, int) { return false; }
int64_t wgrad_umma_workspace_bytes(int, int, int, int, int, int) { return 0; }
int wgrad_umma(const __half*, const __half*, float*, int, int, int, int, int, int, float, const float*, void*, cudaStream_t) {
    MODE_FAIL("wgrad_umma: not built");
}
}  // namespace mode
