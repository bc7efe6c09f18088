// K4: wgrad of the 5x5x5 conv on tcgen05 tensor cores (sm_100a).
//   d_weff[n][tap][o][i] = sum_p dy[n][p][o] * x[n][p + tap - 2][i]      (autograd of RepMode.py:207)
//
// GEMM view: K = voxels, M = output channels, N = input channels, one GEMM per tap.  With 32 channels a
// single tap would fill a quarter of the 128-row MMA, so taps are STACKED through overlapping operand views
// of the same shared-memory bricks (MN-major operands: a "row" of 64 B is one voxel's 32 channels):
//   A = dy brick [19 rows x 8 w], M = 4 blocks x 32 co, block bm = the brick shifted by bm rows  (LBO = 1 row)
//   B = x  brick [20 rows x 12 w], N = 5 blocks x 32 ci, block bn = the brick shifted by bn voxels (LBO = 64 B)
//   one MMA (M128 N160 K16 = 2 rows x 8 voxels) therefore accumulates 20 taps at once:
//     set A (columns 0..159):   x rows offset +1  -> kh = 3 - bm (0..3), kw = bn
//     set B (columns 160..319): x rows offset +5  -> kh = 7 - bm: only bm = 3 (kh = 4) is a real tap, the
//                               other three blocks are discarded (uniform code; 10 sets instead of the minimal 7)
//   kd is fixed per CTA (the x brick comes from plane d + kd - 2).
// Zero padding = TMA out-of-bounds fill on both bricks; the v-space tiling starts 3 rows above the volume so
// every dy row meets every block shift.
// Work split: unit = (sample, co chunk, ci chunk, kd, slab of tiles); each CTA accumulates its whole slab in
// TMEM and writes one [25 taps x 32 x 32] fp32 partial; wgrad_reduce_kernel sums the slabs (deterministic).
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace mode {

using namespace sm100;

namespace wg {
constexpr int TW = 8, TH = 16;
constexpr int DY_ROWS = TH + 3, X_ROWS = TH + 4, X_COLS = TW + 4;
constexpr int DY_BYTES = DY_ROWS * TW * 64;        // 9728
constexpr int DY_SLOT = 10240;
constexpr int X_BYTES = X_ROWS * X_COLS * 64;      // 15360
constexpr int STAGE = DY_SLOT + X_BYTES;           // 25600
constexpr int STAGES = 7;
constexpr int THREADS = 256;
constexpr int PARTIAL_FLOATS = 25 * 32 * 32;
}  // namespace wg

struct WgradParams {
    float* partial;                 // [units][25][32][32]
    int N, D, H, W, Ci, Co;
    int ncic, ncoc, S;              // chunk counts, slabs per (n, coc, cic, kd)
    int tiles_h, tiles_w;
    int trim;                       // 1: skip the all-zero K steps of the last tile row and cut slabs by cost
    int* error_flag;
};

__global__ void __launch_bounds__(wg::THREADS, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap dymap,
                  const WgradParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + wg::STAGES * wg::STAGE;
    const uint32_t full = bars, empty = bars + 8 * wg::STAGES, done = empty + 8 * wg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + wg::STAGES * wg::STAGE + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // decode the unit
    int u = blockIdx.x;
    const int s = u % P.S; u /= P.S;
    const int kd = u % 5; u /= 5;
    const int cic = u % P.ncic; u /= P.ncic;
    const int coc = u % P.ncoc; u /= P.ncoc;
    const int n = u;
    // v-space planes whose x plane d + kd - 2 lies inside the volume
    const int dlo = max(0, 2 - kd), dhi = min(P.D, P.D + 2 - kd);          // [dlo, dhi)
    const int nd = max(0, dhi - dlo);
    const int tiles = nd * P.tiles_h * P.tiles_w;
    // slab boundaries by COST (two-row K steps), not by tile count: tiles of the last tile row are cheaper (see nk below)
    const int nk_last = P.trim ? min(8, (P.H - ((P.tiles_h - 1) * wg::TH - 3) + 1) >> 1) : 8;
    auto cum_cost = [&](int t) -> int64_t {
        const int per_plane = P.tiles_h * P.tiles_w;
        const int td = t / per_plane, rem = t - td * per_plane;
        const int th = rem / P.tiles_w, tw = rem - th * P.tiles_w;
        const int64_t plane_cost = (int64_t)P.tiles_w * (8 * (P.tiles_h - 1) + nk_last);
        return td * plane_cost + (int64_t)th * 8 * P.tiles_w + (int64_t)tw * (th == P.tiles_h - 1 ? nk_last : 8);
    };
    auto cut = [&](int k) -> int {                       // smallest t with cum_cost(t) >= total * k / S
        if (k <= 0) return 0;
        if (k >= P.S) return tiles;
        const int64_t target = cum_cost(tiles) * k / P.S;
        int lo = 0, hi = tiles;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cum_cost(mid) >= target) hi = mid; else lo = mid + 1;
        }
        return lo;
    };
    const int t0 = cut(s), t1 = cut(s + 1);

    if (threadIdx.x == 0) {
        for (int i = 0; i < wg::STAGES; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&xmap); tma_prefetch_desc(&dymap); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t st = 0, use = 0;
            for (int t = t0; t < t1; ++t) {
                const int tw = t % P.tiles_w, th = (t / P.tiles_w) % P.tiles_h, td = t / (P.tiles_w * P.tiles_h);
                const int vw0 = tw * wg::TW, vh0 = th * wg::TH - 3, dv = dlo + td;
                if (!mbar_wait(empty + 8 * st, (use & 1) ^ 1)) { atomicExch(P.error_flag, 11); return; }
                mbar_expect_tx(full + 8 * st, wg::DY_BYTES + wg::X_BYTES);
                const uint32_t dst = base + st * wg::STAGE;
                tma_load_5d(dst, &dymap, full + 8 * st, coc * 32, vw0, vh0, dv, n);
                tma_load_5d(dst + wg::DY_SLOT, &xmap, full + 8 * st, cic * 32, vw0 - 2, vh0 + 1, dv + kd - 2, n);
                if (++st == wg::STAGES) { st = 0; ++use; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // MN-major, 64B swizzle: lo = addr>>4 | (LBO>>4)<<16 ; hi = SBO>>4 | version | swizzle
            const uint32_t hi_a = (512u >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);              // K-group = next row (8 voxels)
            const uint32_t hi_b = ((wg::X_COLS * 64u) >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);  // next brick row
            const uint32_t lbo_a = (512u >> 4) << 16;    // block bm = +1 row of the dy brick
            const uint32_t lbo_b = (64u >> 4) << 16;     // block bn = +1 voxel of the x brick
            const uint32_t idesc = make_idesc(FMT_F16, 128, 160, 1, 1);
            uint32_t st = 0, use = 0, acc = 0;
            // v-rows at or below the volume's last row pair with no dy row: tiles of the LAST tile row need only
            // nk_last = ceil((H - vh0) / 2) of their 8 two-row K steps (H = 128: 2 of 8, i.e. 66 instead of 72 steps per
            // column).  The issue rate of this single-thread loop is what bounds the kernel (r1e: an extra compare-and-branch
            // per MMA pair cost 30 us), so full tiles keep the branch-free fully unrolled body and only the last tile row
            // takes the counted loop; the tile-row index is a wrapped counter, not a division per tile.
            int tw_c = t0 % P.tiles_w, th_c = (t0 / P.tiles_w) % P.tiles_h;
            for (int t = t0; t < t1; ++t) {
                if (!mbar_wait(full + 8 * st, use & 1)) { atomicExch(P.error_flag, 12); return; }
                tc_fence_after();
                const uint32_t dyb = base + st * wg::STAGE, xb = dyb + wg::DY_SLOT;
                if (th_c != P.tiles_h - 1 || nk_last == 8) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t a_lo = ((dyb + (2 * kk) * 512) >> 4) | lbo_a;
                        const uint32_t bA_lo = ((xb + (2 * kk) * wg::X_COLS * 64) >> 4) | lbo_b;
                        const uint32_t bB_lo = ((xb + (2 * kk + 4) * wg::X_COLS * 64) >> 4) | lbo_b;
                        const uint64_t ad = ((uint64_t)hi_a << 32) | a_lo;
                        mma_f16_ss(tmem, ad, ((uint64_t)hi_b << 32) | bA_lo, idesc, acc);
                        mma_f16_ss(tmem + 160, ad, ((uint64_t)hi_b << 32) | bB_lo, idesc, acc);
                        acc = 1;
                    }
                } else {
#pragma unroll 1
                    for (int kk = 0; kk < nk_last; ++kk) {
                        const uint32_t a_lo = ((dyb + (2 * kk) * 512) >> 4) | lbo_a;
                        const uint32_t bA_lo = ((xb + (2 * kk) * wg::X_COLS * 64) >> 4) | lbo_b;
                        const uint32_t bB_lo = ((xb + (2 * kk + 4) * wg::X_COLS * 64) >> 4) | lbo_b;
                        const uint64_t ad = ((uint64_t)hi_a << 32) | a_lo;
                        mma_f16_ss(tmem, ad, ((uint64_t)hi_b << 32) | bA_lo, idesc, acc);
                        mma_f16_ss(tmem + 160, ad, ((uint64_t)hi_b << 32) | bB_lo, idesc, acc);
                        acc = 1;
                    }
                }
                mma_commit(empty + 8 * st);
                if (++st == wg::STAGES) { st = 0; ++use; }
                if (++tw_c == P.tiles_w) { tw_c = 0; if (++th_c == P.tiles_h) th_c = 0; }
            }
            mma_commit(done);
        }
    } else if (warp >= 4) {
        const int bm = warp - 4;                 // TMEM lanes 32*bm .. : M block bm, lane = output channel
        float* out = P.partial + (size_t)blockIdx.x * wg::PARTIAL_FLOATS;
        bool ok = true;
        if (t1 > t0) {
            ok = mbar_wait(done, 0);
            if (!ok && lane == 0) atomicExch(P.error_flag, 13);
            tc_fence_after();
        }
        const uint32_t lane_addr = (uint32_t)(bm * 32) << 16;
        for (int set = 0; set < 2; ++set) {
            const int kh = (set == 0) ? 3 - bm : 7 - bm;
            if (kh > 4) continue;                                        // set B: only bm == 3 is a real tap
            for (int bn = 0; bn < 5; ++bn) {
                float f[32];
                if (t1 > t0 && ok) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem + set * 160 + bn * 32 + lane_addr, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = 0.f;
                }
                float* dst = out + ((size_t)(kh * 5 + bn) * 32 + lane) * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// d_weff[n][kd*25+t][coc*32+o][cic*32+i] = scale * sum_s partial[unit(n,coc,cic,kd,s)][t][o][i]
// one thread = 4 consecutive i (float4): coalesced 16-byte loads from each slab partial, 32-bit index math.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                           int N, int Ci, int Co, int S, float out_scale,
                                                           const float* __restrict__ out_scale_dev) {
    const uint32_t ci4 = (uint32_t)Ci >> 2;
    const uint32_t total4 = (uint32_t)N * 125u * (uint32_t)Co * ci4;
    const uint32_t ncic = Ci / 32, ncoc = Co / 32;
    float scale = out_scale;
    if (out_scale_dev != nullptr) scale *= *out_scale_dev;
    for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < total4; idx += gridDim.x * 256u) {
        const uint32_t i = (idx % ci4) << 2;
        const uint32_t row = idx / ci4;                 // (n*125 + tap)*Co + o
        const uint32_t o = row % (uint32_t)Co;
        const uint32_t nt = row / (uint32_t)Co;
        const uint32_t tap = nt % 125u, n = nt / 125u;
        const uint32_t kd = tap / 25u, t = tap - kd * 25u;
        const size_t unit0 = ((((size_t)n * ncoc + (o >> 5)) * ncic + (i >> 5)) * 5 + kd) * (size_t)S;
        const float* src = partial + unit0 * wg::PARTIAL_FLOATS + ((size_t)t * 32 + (o & 31)) * 32 + (i & 31);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < S; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)s * wg::PARTIAL_FLOATS);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
        *reinterpret_cast<float4*>(dw + (size_t)idx * 4) = acc;
    }
}

// ------------------------------------------------------------------------------------------------ host
int* device_error_flag();   // mode_abi.cu
int make_act_map(CUtensorMap* map, const __half* x, int N, int D, int H, int W, int K, int box_w, int box_h, int box_d);  // conv_umma.cu

static int wgrad_slabs(int N, int Ci, int Co) {
    const int units_per_slab = N * (Ci / 32) * (Co / 32) * 5;
    return max(1, sm_count() / units_per_slab);
}

bool wgrad_umma_supported(int D, int H, int W, int Ci, int Co) {
    (void)D; (void)H;
    return Ci % 32 == 0 && Co % 32 == 0 && Ci >= 32 && Co >= 32 && W % wg::TW == 0;
}

int64_t wgrad_umma_workspace_bytes(int N, int D, int H, int W, int Ci, int Co) {
    (void)D; (void)H; (void)W;
    const int64_t units = (int64_t)N * (Ci / 32) * (Co / 32) * 5 * wgrad_slabs(N, Ci, Co);
    return units * wg::PARTIAL_FLOATS * (int64_t)sizeof(float);
}

int wgrad_umma(const __half* x, const __half* dy, float* dw, int N, int D, int H, int W, int Ci, int Co,
               float out_scale, const float* out_scale_dev, void* workspace, cudaStream_t st) {
    if (!workspace) MODE_FAIL("wgrad_umma: workspace is NULL");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15))
        MODE_FAIL("wgrad_umma: pointers must be 16-byte aligned");
    WgradParams P;
    P.partial = (float*)workspace;
    P.N = N; P.D = D; P.H = H; P.W = W; P.Ci = Ci; P.Co = Co;
    P.ncic = Ci / 32; P.ncoc = Co / 32;
    P.S = wgrad_slabs(N, Ci, Co);
    P.tiles_h = (int)ceil_div(H + 3, wg::TH);
    P.tiles_w = W / wg::TW;
    static const bool notrim = getenv("REPMODE_WGRAD_NOTRIM") != nullptr;     // A/B knob: the untrimmed r1b behaviour
    P.trim = notrim ? 0 : 1;
    P.error_flag = device_error_flag();
    if (!P.error_flag) MODE_FAIL("wgrad_umma: could not allocate the device error flag");
    const int64_t units = (int64_t)N * P.ncic * P.ncoc * 5 * P.S;
    if (units > 0x7fffffff) MODE_FAIL("wgrad_umma: too many work units");
    CUtensorMap xmap, dymap;
    if (make_act_map(&xmap, x, N, D, H, W, Ci, wg::X_COLS, wg::X_ROWS, 1) != 0) return -1;
    if (make_act_map(&dymap, dy, N, D, H, W, Co, wg::TW, wg::DY_ROWS, 1) != 0) return -1;
    const int smem_bytes = wg::STAGES * wg::STAGE + 512 + 1024;
    MODE_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    wgrad_umma_kernel<<<(unsigned)units, wg::THREADS, smem_bytes, st>>>(xmap, dymap, P);
    MODE_LAUNCH_CHECK();
    const int64_t total = (int64_t)N * 125 * Co * Ci / 4;
    if (total > 0x7fffffff) MODE_FAIL("wgrad_umma: d_weff too large for 32-bit indexing");
    const int grid = (int)max((int64_t)1, min(ceil_div(total, 256), (int64_t)sm_count() * 16));
    wgrad_reduce_kernel<<<grid, 256, 0, st>>>((const float*)workspace, dw, N, Ci, Co, P.S, out_scale, out_scale_dev);
    MODE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mode
