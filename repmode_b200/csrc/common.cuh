// Shared host-side helpers for the C-ABI translation units: thread-local error message, CUDA error
// checking that never throws across the extern "C" boundary, small device utilities.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/repmode_b200.h"
#include "peer.cuh"

namespace mode {

char* err_buf();                       // thread-local, 512 bytes (defined in mode_abi.cu)
constexpr int kErrLen = 512;

#define MODE_FAIL(...)                                         \
    do {                                                       \
        snprintf(mode::err_buf(), mode::kErrLen, __VA_ARGS__); \
        return -1;                                             \
    } while (0)

#define MODE_CUDA(x)                                                                                        \
    do {                                                                                                    \
        cudaError_t e_ = (x);                                                                               \
        if (e_ != cudaSuccess) MODE_FAIL("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
    } while (0)

void count_launch();                   // bumps the process-wide kernel-launch counter (mode_launch_count)
#define MODE_LAUNCH_CHECK()           \
    do {                              \
        mode::count_launch();         \
        MODE_CUDA(cudaGetLastError()); \
    } while (0)

__host__ __device__ constexpr int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Cached per-device SM count (read-only after first use; benign race).
int sm_count();

// mode_conv_opts_t resolved by the dispatcher (Dx / Dy16 / y16_scale defaults filled in): haloed input + fused epilogue.
struct ConvExt {
    int Dx, x_off;                 // input planes, input plane of output plane 0's centre tap
    const float* ep_scale;         // per-channel affine after out_scale (eval-mode BatchNorm) or null
    const float* ep_shift;
    int relu;
    __half* y16;                   // optional fp16 copy of the result [N, Dy16, H, W, Nout], output plane q -> plane q + y16_off
    int Dy16, y16_off;
    float y16_scale;
    void* splitk_ws;               // split-K scratch (conv_umma.cu) or null
    long long splitk_ws_bytes;
    PeerPush push;                 // D-sharded slabs: the last CTA broadcasts bn_sums (2 * Nout doubles) to every rank
    __host__ __device__ int p_lo() const { return -x_off; }              // valid input planes in OUTPUT plane coordinates
    __host__ __device__ int p_hi() const { return Dx - x_off - 1; }
};

// Floats of one K4 (wgrad_deep.cu) work-unit partial: [60 entries][32 co][32 ci].  K1b reads these partials directly when a
// layer's wgrad runs one slab per unit (reparam.cu), so the layout is shared.
constexpr int K4_DEEP_PARTIAL_FLOATS = 60 * 32 * 32;

// Streaming 16-byte load for tensors that are read once per kernel: read-only path, no L1 allocation.  Measured with
// tools/stream_probe.cu on this pool's B200 (profiles/r2_stream_probe.txt): two 64 MB input streams read at 5.1 TB/s with
// these loads against 4.5 TB/s with plain ld.global (L1 allocation of data that is never reused).
__device__ __forceinline__ float4 ld_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// fp32 -> fp16 with saturation (never inf): operands of the tcgen05 kernels
__device__ __forceinline__ __half2 sat_half2(float a, float b) {
    return __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
}

}  // namespace mode
