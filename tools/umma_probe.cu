// Hardware probe (test infrastructure, not product code): runs single tcgen05.mma / TMA operations on
// host-supplied shared-memory images and descriptors so that tools/probe_umma.py can check, against
// numpy models, exactly which shared-memory layouts the sm_100a UMMA and TMA units implement
// (shifted start addresses inside a swizzle atom, non-atom-multiple SBO, overlapping MN-major blocks,
// TMA out-of-bounds zero fill) and how many cycles one MMA of a given shape costs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../repmode_b200/csrc/ptx_sm100.cuh"

using namespace sm100;

static thread_local char g_err[512] = "";
extern "C" const char* probe_last_error(void) { return g_err; }
#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            snprintf(g_err, sizeof(g_err), "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)

// ------------------------------------------------------------------------------------------------
// probe_mma: img -> smem, issue n_k * repeat MMAs, dump the 128 x ncols fp32 accumulator.
__global__ void __launch_bounds__(128, 1)
probe_mma_kernel(const uint8_t* __restrict__ img, uint32_t img_bytes, uint64_t adesc, uint64_t bdesc,
                 uint32_t idesc, uint32_t kind, uint32_t n_k, uint32_t a_step16, uint32_t b_step16,
                 uint32_t repeat, uint32_t ncols, float* __restrict__ out, long long* __restrict__ cycles,
                 int* __restrict__ status) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_slot;

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);

    for (uint32_t i = threadIdx.x * 16; i < img_bytes; i += blockDim.x * 16) {
        *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(img + i);
    }
    fence_proxy_async();

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_slot));
    if (threadIdx.x == 32) {
        mbar_init(smem_u32(&mbar), 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;

    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        uint64_t a0 = adesc + (base >> 4), b0 = bdesc + (base >> 4);
        t0 = clock64();
        uint32_t acc = 0;
        for (uint32_t r = 0; r < repeat; ++r) {
            for (uint32_t k = 0; k < n_k; ++k) {
                uint64_t a = a0 + (uint64_t)(k * a_step16), b = b0 + (uint64_t)(k * b_step16);
                if (kind == 0) mma_f16_ss(tmem, a, b, idesc, acc);
                else mma_tf32_ss(tmem, a, b, idesc, acc);
                acc = 1;
            }
        }
        mma_commit(smem_u32(&mbar));
    }
    bool ok = mbar_wait(smem_u32(&mbar), 0, 1u << 24);
    if (threadIdx.x == 0) {
        t1 = clock64();
        cycles[0] = t1 - t0;
    }
    tc_fence_after();
    if (!ok) {
        if (threadIdx.x == 0) status[0] = 1;
    } else {
        for (uint32_t c = 0; c < ncols; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem + ((warp * 32u) << 16) + c, v);
            tmem_ld_wait();
            const uint32_t row = warp * 32 + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c + j < ncols) out[(size_t)row * ncols + c + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

extern "C" int probe_mma(const void* img_dev, uint32_t img_bytes, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                         uint32_t kind, uint32_t n_k, uint32_t a_step16, uint32_t b_step16, uint32_t repeat,
                         uint32_t ncols, float* out_dev, long long* cycles_dev, int* status_dev) {
    const int smem = 200 * 1024;
    if (img_bytes + 1024 > (uint32_t)smem || (img_bytes & 15)) {
        snprintf(g_err, sizeof(g_err), "probe_mma: bad image size %u", img_bytes);
        return -1;
    }
    CK(cudaFuncSetAttribute(probe_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemset(status_dev, 0, sizeof(int)));
    probe_mma_kernel<<<1, 128, smem>>>((const uint8_t*)img_dev, img_bytes, adesc, bdesc, idesc, kind, n_k, a_step16,
                                       b_step16, repeat, ncols, out_dev, cycles_dev, status_dev);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}


// ------------------------------------------------------------------------------------------------
// probe_mma_rate: issue-rate probe. Unrolled groups of 8 MMAs round-robin over NACC accumulators
// (column stride acc_cols), operand descriptor offsets from small tables (16-byte units).
struct RateOffs { uint32_t a[8]; uint32_t b[8]; };
__device__ int g_commit_every = 0;     // probe knob: a tcgen05.commit to a scratch mbarrier after every k-th MMA
extern "C" int probe_set_commit_every(int k) { CK(cudaMemcpyToSymbol(g_commit_every, &k, sizeof(int))); return 0; }

template <int NACC>
__global__ void __launch_bounds__(128, 1)
probe_rate_kernel(const uint8_t* __restrict__ img, uint32_t img_bytes, uint64_t adesc, uint64_t bdesc,
                  uint32_t idesc, uint32_t kind, RateOffs offs, uint32_t acc_cols, uint32_t repeat,
                  long long* __restrict__ cycles, int* __restrict__ status) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar, scratch_bar;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    for (uint32_t i = threadIdx.x * 16; i < img_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(img + i);
    fence_proxy_async();
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_slot));
    if (threadIdx.x == 32) { mbar_init(smem_u32(&mbar), 1); mbar_init(smem_u32(&scratch_bar), 1u << 20); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    if (threadIdx.x == 0) {
        const uint64_t a0 = adesc + (base >> 4), b0 = bdesc + (base >> 4);
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ad[j] = a0 + offs.a[j]; bd[j] = b0 + offs.b[j]; }
        // first group initialises the accumulators
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t d = tmem + (j % NACC) * acc_cols;
            if (kind == 0) mma_f16_ss(d, ad[j], bd[j], idesc, j >= NACC ? 1u : 0u);
            else mma_tf32_ss(d, ad[j], bd[j], idesc, j >= NACC ? 1u : 0u);
        }
        const long long t0 = clock64();
        for (uint32_t r = 0; r < repeat; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t d = tmem + (j % NACC) * acc_cols;
                if (kind == 0) mma_f16_ss(d, ad[j], bd[j], idesc, 1u);
                else mma_tf32_ss(d, ad[j], bd[j], idesc, 1u);
            }
        }
        const long long t_issue = clock64();
        mma_commit(smem_u32(&mbar));
        cycles[1] = t_issue - t0;
        cycles[2] = t0;
    }
    bool ok = mbar_wait(smem_u32(&mbar), 0, 1u << 26);
    if (threadIdx.x == 0) {
        cycles[0] = clock64() - cycles[2];
        if (!ok) status[0] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

extern "C" int probe_mma_rate(const void* img_dev, uint32_t img_bytes, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                              uint32_t kind, const uint32_t* a_offs16, const uint32_t* b_offs16, int nacc,
                              uint32_t acc_cols, uint32_t repeat, long long* cycles_dev, int* status_dev) {
    const int smem = 200 * 1024;
    RateOffs offs;
    for (int j = 0; j < 8; ++j) { offs.a[j] = a_offs16[j]; offs.b[j] = b_offs16[j]; }
    CK(cudaMemset(status_dev, 0, sizeof(int)));
#define LAUNCH_RATE(NA)                                                                                         \
    do {                                                                                                        \
        CK(cudaFuncSetAttribute(probe_rate_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));      \
        probe_rate_kernel<NA><<<1, 128, smem>>>((const uint8_t*)img_dev, img_bytes, adesc, bdesc, idesc, kind,   \
                                                offs, acc_cols, repeat, cycles_dev, status_dev);                 \
    } while (0)
    if (nacc == 1) LAUNCH_RATE(1);
    else if (nacc == 2) LAUNCH_RATE(2);
    else if (nacc == 4) LAUNCH_RATE(4);
    else if (nacc == 8) LAUNCH_RATE(8);
    else { snprintf(g_err, sizeof(g_err), "nacc must be 1,2,4,8"); return -1; }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// probe_mma_rate_pair: the same issue-rate probe with a CTA pair (cta_group::2, M = 256): each CTA holds its own
// 128 A rows and N/2 of the B rows; the leader issues, the commit is multicast to both CTAs.
template <int NACC, int CE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_rate_pair_kernel(const uint8_t* __restrict__ img, uint32_t img_bytes, uint64_t adesc, uint64_t bdesc,
                       uint32_t idesc, RateOffs offs, uint32_t acc_cols, uint32_t repeat,
                       long long* __restrict__ cycles, int* __restrict__ status) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar, scratch_bar;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    for (uint32_t i = threadIdx.x * 16; i < img_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(img + i);
    fence_proxy_async();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    if (warp == 0) tmem_alloc_pair<512>(smem_u32(&tmem_base_slot));
    if (threadIdx.x == 32) { mbar_init(smem_u32(&mbar), 1); mbar_init(smem_u32(&scratch_bar), 1u << 20); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    if (threadIdx.x == 0 && rank == 0) {
        const uint64_t a0 = adesc + (base >> 4), b0 = bdesc + (base >> 4);
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ad[j] = a0 + offs.a[j]; bd[j] = b0 + offs.b[j]; }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            mma_f16_ss_pair(tmem + (j % NACC) * acc_cols, ad[j], bd[j], idesc, j >= NACC ? 1u : 0u);
        const long long t0 = clock64();
        for (uint32_t r = 0; r < repeat; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                mma_f16_ss_pair(tmem + (j % NACC) * acc_cols, ad[j], bd[j], idesc, 1u);
                if (CE && ((j + 1) % (CE ? CE : 1)) == 0) mma_commit_pair(smem_u32(&scratch_bar), 3u);
            }
        }
        const long long t_issue = clock64();
        mma_commit_pair(smem_u32(&mbar), 3u);
        cycles[1] = t_issue - t0;
        cycles[2] = t0;
    }
    bool ok = mbar_wait(smem_u32(&mbar), 0, 1u << 26);
    if (threadIdx.x == 0 && rank == 0) cycles[0] = clock64() - cycles[2];
    if (threadIdx.x == 0 && !ok) status[0] = 1 + (int)rank;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

extern "C" int probe_mma_rate_pair(const void* img_dev, uint32_t img_bytes, uint64_t adesc, uint64_t bdesc,
                                   uint32_t idesc, const uint32_t* a_offs16, const uint32_t* b_offs16, int nacc,
                                   uint32_t acc_cols, uint32_t repeat, long long* cycles_dev, int* status_dev) {
    const int smem = 200 * 1024;
    RateOffs offs;
    for (int j = 0; j < 8; ++j) { offs.a[j] = a_offs16[j]; offs.b[j] = b_offs16[j]; }
    CK(cudaMemset(status_dev, 0, sizeof(int)));
#define LAUNCH_RATE2(NA, CE)                                                                                      \
    do {                                                                                                          \
        CK(cudaFuncSetAttribute(probe_rate_pair_kernel<NA, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        probe_rate_pair_kernel<NA, CE><<<2, 128, smem>>>((const uint8_t*)img_dev, img_bytes, adesc, bdesc, idesc, offs, \
                                                     acc_cols, repeat, cycles_dev, status_dev);                   \
    } while (0)
    int ce = 0;
    CK(cudaMemcpyFromSymbol(&ce, g_commit_every, sizeof(int)));
    if (nacc == 1 && ce == 0) LAUNCH_RATE2(1, 0);
    else if (nacc == 2 && ce == 0) LAUNCH_RATE2(2, 0);
    else if (nacc == 2 && ce == 1) LAUNCH_RATE2(2, 1);
    else if (nacc == 2 && ce == 2) LAUNCH_RATE2(2, 2);
    else if (nacc == 2 && ce == 4) LAUNCH_RATE2(2, 4);
    else if (nacc == 2 && ce == 8) LAUNCH_RATE2(2, 8);
    else { snprintf(g_err, sizeof(g_err), "unsupported nacc / commit_every"); return -1; }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// probe_tma: one tiled TMA load of a (<=5-d) box into shared memory, then dump the bytes.
__global__ void __launch_bounds__(128, 1)
probe_tma_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, int c4, uint32_t dst_off,
                 uint32_t expect_bytes, uint8_t* __restrict__ dump, uint32_t dump_bytes, int* __restrict__ status) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    for (uint32_t i = threadIdx.x * 16; i < dump_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0xEEEEEEEEu, 0xEEEEEEEEu, 0xEEEEEEEEu, 0xEEEEEEEEu);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&mbar), 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&mbar), expect_bytes);
        tma_load_5d(base + dst_off, &map, smem_u32(&mbar), c0, c1, c2, c3, c4);
    }
    bool ok = mbar_wait(smem_u32(&mbar), 0, 1u << 24);
    if (!ok && threadIdx.x == 0) status[0] = 1;
    __syncthreads();
    for (uint32_t i = threadIdx.x * 16; i < dump_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(dump + i) = *reinterpret_cast<uint4*>(smem + i);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// dims/box innermost-first (rank 5); strides_bytes has rank-1 entries (dims 1..4). elem: 2 = fp16, 4 = fp32.
// swizzle: 0 none, 1 32B, 2 64B, 3 128B (CUtensorMapSwizzle values).
extern "C" int probe_tma(void* gptr, int elem_bytes, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle, const int* coords, uint32_t dst_off, uint32_t expect_bytes,
                         uint8_t* dump_dev, uint32_t dump_bytes, int* status_dev) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled entry point not found");
        return -1;
    }
    CUtensorMap map;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 5; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i < 4; ++i) gstr[i] = strides_bytes[i];
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = ((EncodeTiledFn)fn)(&map, dt, 5, gptr, gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed: %d", (int)r);
        return -1;
    }
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemset(status_dev, 0, sizeof(int)));
    probe_tma_kernel<<<1, 128, smem>>>(map, coords[0], coords[1], coords[2], coords[3], coords[4], dst_off,
                                       expect_bytes, dump_dev, dump_bytes, status_dev);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}
