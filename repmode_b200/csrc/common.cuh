// Shared host-side helpers for the C-ABI translation units: thread-local error message, CUDA error
// checking that never throws across the extern "C" boundary, small device utilities.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/repmode_b200.h"

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

}  // namespace mode
