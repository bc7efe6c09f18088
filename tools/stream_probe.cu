// Hardware probe (test infrastructure): what does a streaming READ kernel need to reach HBM speed on this B200?
// Variants: number of input streams (1 / 2), 16-byte loads in flight per thread (1 / 2 / 4 / 8), blocks per SM, load flavour
// (plain / ld.global.cs / ld.global.nc.L1::no_allocate).  Prints GB/s per variant.   nvcc -arch=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <int FL>
__device__ __forceinline__ float4 ld(const float4* p) {
    float4 v;
    if (FL == 0) v = *p;
    else if (FL == 1) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int U, int NS, int FL>
__global__ void __launch_bounds__(256) rd(const float4* __restrict__ a, const float4* __restrict__ b, long long nvec, float* out) {
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += U * stride) {
        float4 va[U], vb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long k = i + u * stride;
            if (k < nvec) { va[u] = ld<FL>(a + k); if (NS == 2) vb[u] = ld<FL>(b + k); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long k = i + u * stride;
            if (k < nvec) { acc += va[u].x + va[u].y + va[u].z + va[u].w; if (NS == 2) acc += vb[u].x * vb[u].w; }
        }
    }
    if (acc == 123.456f) *out = acc;
}

// block-contiguous variant: each block owns a contiguous chunk (no grid stride)
template <int U, int NS>
__global__ void __launch_bounds__(256) rd_chunk(const float4* __restrict__ a, const float4* __restrict__ b, long long nvec, float* out) {
    const long long per = (nvec + gridDim.x - 1) / gridDim.x;
    const long long lo = per * blockIdx.x, hi = min(nvec, lo + per);
    float acc = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += U * 256) {
        float4 va[U], vb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long k = i + u * 256;
            if (k < hi) { va[u] = a[k]; if (NS == 2) vb[u] = b[k]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long k = i + u * 256;
            if (k < hi) { acc += va[u].x + va[u].y + va[u].z + va[u].w; if (NS == 2) acc += vb[u].x * vb[u].w; }
        }
    }
    if (acc == 123.456f) *out = acc;
}

__global__ void cp(const float4* __restrict__ a, float4* __restrict__ b, long long nvec) {
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * stride < nvec) v[u] = a[i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * stride < nvec) b[i + u * stride] = v[u];
    }
}

template <typename F>
static float timeit(F f, int it = 20) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < it; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / it;
}

int main(int argc, char** argv) {
    const long long bytes = 64ll << 20;      // per stream: the headline activation tensor (67 MB)
    const long long nvec = bytes / 16;
    float4 *a, *b, *c, *big;
    float* out;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&c, bytes); cudaMalloc(&out, 4);
    cudaMalloc(&big, 512ll << 20);
    cudaMemset(a, 1, bytes); cudaMemset(b, 1, bytes); cudaMemset(c, 0, bytes); cudaMemset(big, 0, 512ll << 20);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto flush = [&]() { cudaMemsetAsync(big, 0, 512ll << 20); };      // L2 flush between timed launches
    // baseline for the flush cost
    const float t_flush = timeit([&]() { flush(); });
#define RUN(name, NSV, ...)                                                                                \
    {                                                                                                       \
        const float t = timeit([&]() { flush(); __VA_ARGS__; }) - t_flush;                                  \
        printf("%-44s %7.2f us  %7.1f GB/s\n", name, t * 1e3, NSV * bytes / (t * 1e-3) / 1e9);             \
    }
    printf("SMs %d, flush %.1f us\n", sms, t_flush * 1e3);
    RUN("copy 64MB->64MB (r+w counted), 8 blk/SM u4", 2, (cp<<<sms * 8, 256>>>(a, c, nvec)));
    for (int bps : {2, 4, 8, 16, 32}) {
        char nm[96];
        snprintf(nm, 96, "read 1 stream  u1 plain  %2d blk/SM", bps); RUN(nm, 1, (rd<1, 1, 0><<<sms * bps, 256>>>(a, b, nvec, out)));
        snprintf(nm, 96, "read 1 stream  u4 plain  %2d blk/SM", bps); RUN(nm, 1, (rd<4, 1, 0><<<sms * bps, 256>>>(a, b, nvec, out)));
        snprintf(nm, 96, "read 1 stream  u8 plain  %2d blk/SM", bps); RUN(nm, 1, (rd<8, 1, 0><<<sms * bps, 256>>>(a, b, nvec, out)));
        snprintf(nm, 96, "read 2 streams u2 plain  %2d blk/SM", bps); RUN(nm, 2, (rd<2, 2, 0><<<sms * bps, 256>>>(a, b, nvec, out)));
        snprintf(nm, 96, "read 2 streams u4 plain  %2d blk/SM", bps); RUN(nm, 2, (rd<4, 2, 0><<<sms * bps, 256>>>(a, b, nvec, out)));
        snprintf(nm, 96, "read 2 streams u4 .cs    %2d blk/SM", bps); RUN(nm, 2, (rd<4, 2, 1><<<sms * bps, 256>>>(a, b, nvec, out)));
        snprintf(nm, 96, "read 2 streams u4 .nc    %2d blk/SM", bps); RUN(nm, 2, (rd<4, 2, 2><<<sms * bps, 256>>>(a, b, nvec, out)));
        snprintf(nm, 96, "read 2 streams u4 chunk  %2d blk/SM", bps); RUN(nm, 2, (rd_chunk<4, 2><<<sms * bps, 256>>>(a, b, nvec, out)));
    }
    // many small blocks (torch-style): one u4 pass per block, no loop
    {
        const long long blocks = (nvec + 256 * 4 - 1) / (256 * 4);
        RUN("read 2 streams u4 plain, one pass per block", 2, (rd<4, 2, 0><<<(unsigned)blocks, 256>>>(a, b, nvec, out)));
        RUN("read 1 stream  u4 plain, one pass per block", 1, (rd<4, 1, 0><<<(unsigned)blocks, 256>>>(a, b, nvec, out)));
        const long long blocks8 = (nvec + 256 * 8 - 1) / (256 * 8);
        RUN("read 2 streams u8 plain, one pass per block", 2, (rd<8, 2, 0><<<(unsigned)blocks8, 256>>>(a, b, nvec, out)));
    }
    // no flush: back-to-back (what a warm loop sees)
    {
        const float t = timeit([&]() { rd<4, 2, 0><<<sms * 8, 256>>>(a, b, nvec, out); });
        printf("%-44s %7.2f us  %7.1f GB/s\n", "read 2 streams u4, no L2 flush between", t * 1e3, 2 * bytes / (t * 1e-3) / 1e9);
    }
    return 0;
}
