// K4, round-1 kernel (A/B arm: REPMODE_WGRAD_SPLIT=1 or impl = 5; the default is wgrad_deep.cu): wgrad of the 5x5x5 conv
// on tcgen05 with the tap set split so that (almost) every accumulator column is a real tap.
//   d_weff[n][tap][o][i] = sum_p dy[n][p][o] * x[n][p + tap - 2][i]      (autograd of RepMode.py:207)
//
// The first-generation kernel (removed) stacked taps through overlapping MN-major views: M = 4 dy row shifts x 32 co, N = 5 x voxel shifts x 32 ci,
// one M128 N160 K16 MMA = 20 taps.  Five kh rows do not fit four row shifts, so that kernel spends a second MMA per K
// step on kh = 4 alone (25 useful of 40 computed tap blocks).  Here the 125 taps are covered by two kinds of work unit:
//   K units ("rows"):   kh = 0..3 of one or two kd.  Per (d-plane, 8x16 patch): one dy brick (19 rows) and the x brick of
//                       each kd (x plane d + kd - 2); M blocks = dy row shifts (kh = 3 - bm), N blocks = kw.  One MMA per
//                       K step per kd, every block useful.  The kd are grouped (0,1) (2,3) (4) so a dy brick is shared.
//   L units ("leftover"): kh = 4 of ALL kd.  Per (pair of x planes p, p+1; patch): ONE 6-plane dy box (planes p-2..p+3 at
//                       a uniform 8 KB stride) and the two x bricks; M blocks = dy PLANE shifts (LBO = plane stride):
//                       MMA 1 starts at plane p-1 -> kd = 3 - bm (20 taps, all useful), MMA 2 at plane p-2 -> bm = 0 is
//                       kd = 4 (5 useful taps; the 15 others are discarded).
//   => 7 MMAs per K step and x plane instead of 10 (125 useful of 140 computed tap blocks).
// Zero padding = TMA out-of-bounds fill on every brick; x planes outside the volume are neither loaded nor multiplied.
// The last tile row only issues the K steps that can meet a dy row (see nk below).
// Work split: every unit accumulates its slab of tiles in TMEM and writes one fp32 partial; wgrad_split_reduce_kernel sums
// the slabs in a fixed order (deterministic).  Slabs are cut by COST (K steps), and the number of slabs per unit kind is
// chosen so that all CTAs of one wave carry about the same number of MMAs.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.
// Roofline: tensor pipe; algorithmic work 2*125*Ci*Co FLOP per voxel.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace mode {

using namespace sm100;

namespace ws {
constexpr int TW = 8, TH = 16, X_COLS = TW + 4;
constexpr int X_BYTES = TH * X_COLS * 64;            // 12288: x brick, 16 rows x 12 voxels x 32 ch fp16
constexpr int K_DY_ROWS = TH + 3;
constexpr int K_DY_BYTES = K_DY_ROWS * TW * 64;      // 9728
constexpr int K_DY_SLOT = 10240;
constexpr int K_STAGE = K_DY_SLOT + 2 * X_BYTES;     // 34816
constexpr int K_STAGES = 6;
constexpr int L_PLANE = TH * TW * 64;                // 8192
constexpr int L_PLANES = 6;
constexpr int L_DY_BYTES = L_PLANES * L_PLANE;       // 49152
constexpr int L_STAGE = L_DY_BYTES + 2 * X_BYTES;    // 73728
constexpr int L_STAGES = 3;
constexpr int OPERAND_BYTES = (K_STAGES * K_STAGE > L_STAGES * L_STAGE) ? K_STAGES * K_STAGE : L_STAGES * L_STAGE;
constexpr int MAX_STAGES = 6;
constexpr int THREADS = 256;
constexpr int ENTRIES = 40;                          // [32 co][32 ci] fp32 blocks per unit partial
constexpr int PARTIAL_FLOATS = ENTRIES * 32 * 32;
static_assert(K_STAGE % 1024 == 0 && L_STAGE % 1024 == 0 && K_DY_SLOT % 1024 == 0 && L_PLANE % 1024 == 0, "swizzle atoms");
static_assert(OPERAND_BYTES + 512 + 1024 <= 227 * 1024, "shared memory budget");
}  // namespace ws

struct SplitParams {
    float* partial;                 // [units][ENTRIES][32][32]
    int N, D, H, W, Ci, Co;
    int ncic, ncoc;
    int SL, SK2, SK1;               // slabs per L unit group, per kd-pair K unit group, per kd = 4 K unit group
    int nL, nK2;                    // unit counts: grid = [L units][K2 units][K1 units]
    int tiles_hK, tiles_hL, tiles_w;
    int* error_flag;
};

// cumulative cost (two-row K steps) of the first t tiles of a (plane, tile row, tile column) walk whose last tile
// row needs nk_last instead of 8 steps
__device__ __forceinline__ int64_t cum_steps(int t, int tiles_h, int tiles_w, int nk_last) {
    const int per_plane = tiles_h * tiles_w;
    const int td = t / per_plane, rem = t - td * per_plane;
    const int th = rem / tiles_w, tw = rem - th * tiles_w;
    const int64_t plane_cost = (int64_t)tiles_w * (8 * (tiles_h - 1) + nk_last);
    return td * plane_cost + (int64_t)th * 8 * tiles_w + (int64_t)tw * (th == tiles_h - 1 ? nk_last : 8);
}
// smallest t with cum_steps(t) >= total * k / S
__device__ __forceinline__ int slab_cut(int k, int S, int tiles, int tiles_h, int tiles_w, int nk_last) {
    if (k <= 0) return 0;
    if (k >= S) return tiles;
    const int64_t target = cum_steps(tiles, tiles_h, tiles_w, nk_last) * k / S;
    int lo = 0, hi = tiles;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum_steps(mid, tiles_h, tiles_w, nk_last) >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__global__ void __launch_bounds__(ws::THREADS, 1)
wgrad_split_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap dymapK,
                   const __grid_constant__ CUtensorMap dymapL, const SplitParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + ws::OPERAND_BYTES;
    const uint32_t full = bars, empty = bars + 8 * ws::MAX_STAGES, done = empty + 8 * ws::MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ws::OPERAND_BYTES + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- decode the unit: kind, channel chunks, kd group, slab ------------------------------------------------
    int u = blockIdx.x;
    const bool is_l = u < P.nL;
    int s, S, kd0 = 0, nkd = 0;
    if (is_l) {
        S = P.SL; s = u % S; u /= S;
    } else if (u < P.nL + P.nK2) {
        u -= P.nL;
        S = P.SK2; s = u % S; u /= S;
        kd0 = 2 * (u & 1); nkd = 2; u >>= 1;
    } else {
        u -= P.nL + P.nK2;
        S = P.SK1; s = u % S; u /= S;
        kd0 = 4; nkd = 1;
    }
    const int cic = u % P.ncic; u /= P.ncic;
    const int coc = u % P.ncoc; u /= P.ncoc;
    const int n = u;

    // ---- the unit's tile walk (plane-major, then tile row, then tile column) and this slab's share of it ----------
    int dlo, nd, tiles_h, nk_last;
    if (is_l) {
        dlo = 0; nd = (P.D + 1) >> 1;                                   // pairs of x planes
        tiles_h = P.tiles_hL;
        nk_last = min(8, (P.H - (tiles_h - 1) * ws::TH + 1) >> 1);
    } else {
        // dy planes d whose x plane d + kd - 2 lies inside the volume for at least one kd of the group
        dlo = max(0, 2 - (kd0 + nkd - 1));
        const int dhi = min(P.D, P.D + 2 - kd0);
        nd = max(0, dhi - dlo);
        tiles_h = P.tiles_hK;
        nk_last = min(8, (P.H - ((tiles_h - 1) * ws::TH - 3) + 1) >> 1);
    }
    const int tiles = nd * tiles_h * P.tiles_w;
    const int t0 = slab_cut(s, S, tiles, tiles_h, P.tiles_w, nk_last);
    const int t1 = slab_cut(s + 1, S, tiles, tiles_h, P.tiles_w, nk_last);
    const int stages = is_l ? ws::L_STAGES : ws::K_STAGES;
    const uint32_t stage_bytes = is_l ? ws::L_STAGE : ws::K_STAGE;

    if (threadIdx.x == 0) {
        for (int i = 0; i < ws::MAX_STAGES; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&xmap); tma_prefetch_desc(is_l ? &dymapL : &dymapK); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t st = 0, use = 0;
            for (int t = t0; t < t1; ++t) {
                const int tw = t % P.tiles_w, th = (t / P.tiles_w) % tiles_h, td = t / (P.tiles_w * tiles_h);
                const int vw0 = tw * ws::TW;
                if (!mbar_wait(empty + 8 * st, (use & 1) ^ 1)) { atomicExch(P.error_flag, 21); return; }
                const uint32_t dst = base + st * stage_bytes;
                if (is_l) {
                    const int p = 2 * td, vh0 = th * ws::TH;
                    const int nx = (p + 1 < P.D) ? 2 : 1;
                    mbar_expect_tx(full + 8 * st, ws::L_DY_BYTES + nx * ws::X_BYTES);
                    tma_load_5d(dst, &dymapL, full + 8 * st, coc * 32, vw0, vh0, p - 2, n);             // planes p-2 .. p+3
                    for (int xi = 0; xi < nx; ++xi)                                                       // x rows = dy rows + 2 (kh = 4)
                        tma_load_5d(dst + ws::L_DY_BYTES + xi * ws::X_BYTES, &xmap, full + 8 * st, cic * 32, vw0 - 2,
                                    vh0 + 2, p + xi, n);
                } else {
                    const int d = dlo + td, vh0 = th * ws::TH - 3;
                    int nvalid = 0;
                    for (int j = 0; j < nkd; ++j) nvalid += (d + kd0 + j - 2 >= 0 && d + kd0 + j - 2 < P.D) ? 1 : 0;
                    mbar_expect_tx(full + 8 * st, ws::K_DY_BYTES + nvalid * ws::X_BYTES);
                    tma_load_5d(dst, &dymapK, full + 8 * st, coc * 32, vw0, vh0, d, n);
                    for (int j = 0; j < nkd; ++j) {
                        const int xp = d + kd0 + j - 2;
                        if (xp >= 0 && xp < P.D)                                                          // x rows = v + 1: kh = 3 - bm
                            tma_load_5d(dst + ws::K_DY_SLOT + j * ws::X_BYTES, &xmap, full + 8 * st, cic * 32, vw0 - 2,
                                        vh0 + 1, xp, n);
                    }
                }
                if (++st == (uint32_t)stages) { st = 0; ++use; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The WHOLE warp runs this loop on warp-uniform values (kernel parameters, block index, loop counters, vote
        // results) and one elected lane is predicated inside the asm: under `if (lane == 0)` ptxas must assume divergence
        // and wraps every UTCHMMA in an ELECT / 5x R2UR.BROADCAST loop, and that issue cost (~100-117 cycles per MMA
        // measured) -- not the 81 cycles the tensor pipe needs -- is what bounded this kernel (r1d) and the stacked one.
        // Every full tile takes a branch-free, fully unrolled body chosen ONCE per tile; only the last tile row (fewer K
        // steps) and the volume's boundary planes take the counted loops.  Tile coordinates are wrapped counters.
        {
            const uint32_t sel = elect_one() ? 1u : 0u;
            const uint32_t tm = __reduce_max_sync(0xffffffffu, tmem);             // provably uniform copy of the TMEM base
            // MN-major operands, 64B swizzle: lo = addr>>4 | (LBO>>4)<<16 ; hi = SBO>>4 | version | swizzle.
            // SBO = stride between the two 8-voxel K groups of a K16 step = the next brick row.
            const uint32_t hi_a = (512u >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);
            const uint32_t hi_b = ((ws::X_COLS * 64u) >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);
            const uint32_t lbo_row = (512u >> 4) << 16;                 // M block bm = dy brick + bm rows
            const uint32_t lbo_plane = ((uint32_t)ws::L_PLANE >> 4) << 16;   // M block bm = dy plane + bm
            const uint32_t lbo_vox = (64u >> 4) << 16;                  // N block bn = x brick + bn voxels
            const uint32_t idesc = make_idesc(FMT_F16, 128, 160, 1, 1);
            uint32_t st = 0, use = 0, acc0 = 0, acc1 = 0;
            const int per_plane = tiles_h * P.tiles_w;
            int tw_c = t0 % P.tiles_w, th_c = (t0 / P.tiles_w) % tiles_h, td = t0 / per_plane;
            bool ok = true;
#define WS_B(xbase, kk) ((((xbase) + 2 * (kk) * ws::X_COLS * 64) >> 4) | lbo_vox)
#define WS_MMA(col, alo, blo, acc) mma_f16_ss_sel(tm + (col), (alo), hi_a, (blo), hi_b, idesc, (acc), sel)
            for (int t = t0; t < t1; ++t) {
                if (!mbar_wait_warp<false>(full + 8 * st, use & 1)) { ok = false; break; }
                tc_fence_after();
                const uint32_t sb = base + st * stage_bytes;
                const bool full_row = th_c != tiles_h - 1 || nk_last == 8;
                if (is_l) {
                    const uint32_t x0 = sb + ws::L_DY_BYTES, x1 = x0 + ws::X_BYTES;
                    const int nx = (2 * td + 1 < P.D) ? 2 : 1;
                    if (full_row && nx == 2) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            // x plane p: dy planes p-1.. (kd = 3 - bm) and p-2.. (bm = 0: kd = 4); x plane p+1: one plane on
                            const uint32_t b0 = WS_B(x0, kk), b1 = WS_B(x1, kk);
                            const uint32_t a0 = ((sb + 0 * ws::L_PLANE + kk * 1024) >> 4) | lbo_plane;
                            const uint32_t a1 = ((sb + 1 * ws::L_PLANE + kk * 1024) >> 4) | lbo_plane;
                            const uint32_t a2 = ((sb + 2 * ws::L_PLANE + kk * 1024) >> 4) | lbo_plane;
                            WS_MMA(0, a1, b0, acc0);
                            WS_MMA(160, a0, b0, acc0);
                            WS_MMA(0, a2, b1, 1u);
                            WS_MMA(160, a1, b1, 1u);
                            acc0 = 1;
                        }
                    } else {
                        const int nk = full_row ? 8 : nk_last;
#pragma unroll 1
                        for (int kk = 0; kk < nk; ++kk) {
#pragma unroll 1
                            for (int xi = 0; xi < nx; ++xi) {
                                const uint32_t bd = WS_B(x0 + xi * ws::X_BYTES, kk);
                                const uint32_t a1 = ((sb + (1 + xi) * ws::L_PLANE + kk * 1024) >> 4) | lbo_plane;   // planes p-1..p+2
                                const uint32_t a2 = ((sb + xi * ws::L_PLANE + kk * 1024) >> 4) | lbo_plane;         // planes p-2..p+1
                                WS_MMA(0, a1, bd, acc0);
                                WS_MMA(160, a2, bd, acc0);
                                acc0 = 1;
                            }
                        }
                    }
                } else {
                    const int d = dlo + td;
                    const bool v0 = d + kd0 - 2 >= 0 && d + kd0 - 2 < P.D;
                    const bool v1 = nkd > 1 && d + kd0 - 1 >= 0 && d + kd0 - 1 < P.D;
                    const uint32_t x0 = sb + ws::K_DY_SLOT, x1 = x0 + ws::X_BYTES;
                    if (full_row && v0 && v1) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint32_t ad = ((sb + kk * 1024) >> 4) | lbo_row;
                            WS_MMA(0, ad, WS_B(x0, kk), acc0);
                            WS_MMA(160, ad, WS_B(x1, kk), acc1);
                            acc0 = 1; acc1 = 1;
                        }
                    } else if (full_row && v0 && nkd == 1) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint32_t ad = ((sb + kk * 1024) >> 4) | lbo_row;
                            WS_MMA(0, ad, WS_B(x0, kk), acc0);
                            acc0 = 1;
                        }
                    } else {
                        const int nk = full_row ? 8 : nk_last;
#pragma unroll 1
                        for (int kk = 0; kk < nk; ++kk) {
                            const uint32_t ad = ((sb + kk * 1024) >> 4) | lbo_row;
                            if (v0) { WS_MMA(0, ad, WS_B(x0, kk), acc0); acc0 = 1; }
                            if (v1) { WS_MMA(160, ad, WS_B(x1, kk), acc1); acc1 = 1; }
                        }
                    }
                }
                mma_commit_sel(empty + 8 * st, sel);
                if (++st == (uint32_t)stages) { st = 0; ++use; }
                if (++tw_c == P.tiles_w) { tw_c = 0; if (++th_c == tiles_h) { th_c = 0; ++td; } }
            }
#undef WS_MMA
#undef WS_B
            if (ok) mma_commit_sel(done, sel);
            else if (lane == 0) atomicExch(P.error_flag, 22);
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> fp32 partial =====================
        const int bm = warp - 4;                 // TMEM lanes 32*bm .. : M block bm, lane = output channel
        float* out = P.partial + (size_t)blockIdx.x * ws::PARTIAL_FLOATS;
        bool ok = true;
        if (t1 > t0) {
            ok = mbar_wait(done, 0);
            if (!ok && lane == 0) atomicExch(P.error_flag, 23);
            tc_fence_after();
        }
        const uint32_t lane_addr = (uint32_t)(bm * 32) << 16;
        const int nsets = is_l ? 2 : nkd;
        // planes this slab walked (a K set whose x planes all fell outside the volume never received an MMA: its TMEM
        // columns are stale and the partial is zero)
        const int per_plane = tiles_h * P.tiles_w;
        const int d_first = dlo + t0 / per_plane, d_last = dlo + (max(t1, 1) - 1) / per_plane;
        for (int set = 0; set < nsets; ++set) {
            int entry0;
            if (is_l) {
                if (set == 1 && bm != 0) continue;                     // MMA 2: only plane block 0 (kd = 4) is a tap
                entry0 = (set == 0 ? 3 - bm : 4) * 5;                  // entry = kd*5 + kw
            } else {
                entry0 = set * 20 + (3 - bm) * 5;                      // entry = set*20 + kh*5 + kw
            }
            bool have = t1 > t0 && ok;
            if (!is_l) have = have && max(d_first, 2 - (kd0 + set)) <= min(d_last, P.D + 1 - (kd0 + set));
            for (int bn = 0; bn < 5; ++bn) {
                float f[32];
                if (have) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem + set * 160 + bn * 32 + lane_addr, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = 0.f;
                }
                float* dst = out + ((size_t)(entry0 + bn) * 32 + lane) * 32;
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

// d_weff[n][kd*25 + kh*5 + kw][coc*32+o][cic*32+i] = scale * sum over the slabs of the unit group that owns the tap.
// One thread = 4 consecutive i (float4): coalesced 16-byte loads from each slab partial.
__global__ void __launch_bounds__(256) wgrad_split_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                 int N, int Ci, int Co, int SL, int SK2, int SK1, int nL,
                                                                 int nK2, float out_scale,
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
        const uint32_t kd = tap / 25u, kh = (tap / 5u) % 5u, kw = tap % 5u;
        const uint32_t g = (n * ncoc + (o >> 5)) * ncic + (i >> 5);
        uint32_t unit0, entry;
        int S;
        if (kh == 4u) { unit0 = g * SL; S = SL; entry = kd * 5u + kw; }
        else if (kd < 4u) { unit0 = nL + (g * 2u + (kd >> 1)) * SK2; S = SK2; entry = (kd & 1u) * 20u + kh * 5u + kw; }
        else { unit0 = nL + nK2 + g * SK1; S = SK1; entry = kh * 5u + kw; }
        const float* src = partial + (size_t)unit0 * ws::PARTIAL_FLOATS + ((size_t)entry * 32 + (o & 31)) * 32 + (i & 31);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < S; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)s * ws::PARTIAL_FLOATS);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
        *reinterpret_cast<float4*>(dw + (size_t)idx * 4) = acc;
    }
}

// ------------------------------------------------------------------------------------------------ host
int* device_error_flag();   // mode_abi.cu
int make_act_map(CUtensorMap* map, const __half* x, int N, int D, int H, int W, int K, int box_w, int box_h, int box_d);  // conv_umma.cu

struct SplitPlan { int SL, SK2, SK1; };

// Slabs per unit kind: one wave of CTAs when the layer has few channel chunks, each CTA carrying about the same number
// of K steps (a K2 group is two kd of the row taps, K1 one kd, L the kh = 4 taps of all kd at two MMAs per step).
static SplitPlan split_plan(int N, int D, int H, int W, int Ci, int Co) {
    const int groups = N * (Ci / 32) * (Co / 32);
    const int slots = std::max(1, sm_count() / groups);
    const int tiles_w = W / ws::TW;
    const int tiles_hK = (int)ceil_div(H + 3, ws::TH), tiles_hL = (int)ceil_div(H, ws::TH);
    const double stepsK = tiles_w * (8.0 * (tiles_hK - 1) + std::min(8, (H - ((tiles_hK - 1) * ws::TH - 3) + 1) >> 1));
    const double stepsL = tiles_w * (8.0 * (tiles_hL - 1) + std::min(8, (H - (tiles_hL - 1) * ws::TH + 1) >> 1));
    auto nd = [&](int kd) { return (double)std::max(0, std::min(D, D + 2 - kd) - std::max(0, 2 - kd)); };
    const double cK2 = 0.5 * ((nd(0) + nd(1)) + (nd(2) + nd(3))) * stepsK;      // MMAs of one kd-pair group (average)
    const double cK1 = nd(4) * stepsK;
    // Slabs are weighted by MMA count alone.  Measured on the headline layer (ncu sm__cycles_active min / avg / max over the
    // SMs): this plan (43 L + 84 K2 + 21 K1 CTAs) 123k / 167k / 215k cycles, 134 us (r1g); weighting the L slabs 1.7x
    // (60 L + 70 K2 + 17 K1) 106k / 167k / 245k, 153 us (r1h) -- i.e. the L units are the FAST ones (~80 cycles per MMA,
    // the tensor pipe's own rate) and the K units the slow ones (~115-140 cycles per MMA although the issue loop is no
    // longer the limit).  Why the K units stall is the open question for this kernel.
    const double cL = 2.0 * D * stepsL;
    SplitPlan p{1, 1, 1};
    while (2 * p.SK2 + p.SK1 + p.SL < slots) {
        const double a = cK2 / p.SK2, b = cK1 / p.SK1, c = cL / p.SL;
        if (a >= b && a >= c && 2 * (p.SK2 + 1) + p.SK1 + p.SL <= slots) ++p.SK2;     // a K2 group takes two CTAs per slab
        else if (c >= b) ++p.SL;
        else ++p.SK1;
    }
    return p;
}

bool wgrad_split_supported(int D, int H, int W, int Ci, int Co) {
    (void)D; (void)H;
    return Ci % 32 == 0 && Co % 32 == 0 && Ci >= 32 && Co >= 32 && W % ws::TW == 0;
}

int64_t wgrad_split_workspace_bytes(int N, int D, int H, int W, int Ci, int Co) {
    const SplitPlan p = split_plan(N, D, H, W, Ci, Co);
    const int64_t groups = (int64_t)N * (Ci / 32) * (Co / 32);
    return groups * (p.SL + 2 * p.SK2 + p.SK1) * ws::PARTIAL_FLOATS * (int64_t)sizeof(float);
}

int wgrad_split(const __half* x, const __half* dy, float* dw, int N, int D, int H, int W, int Ci, int Co,
                float out_scale, const float* out_scale_dev, void* workspace, cudaStream_t st) {
    if (!workspace) MODE_FAIL("wgrad_split: workspace is NULL");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15))
        MODE_FAIL("wgrad_split: pointers must be 16-byte aligned");
    const SplitPlan plan = split_plan(N, D, H, W, Ci, Co);
    SplitParams P;
    P.partial = (float*)workspace;
    P.N = N; P.D = D; P.H = H; P.W = W; P.Ci = Ci; P.Co = Co;
    P.ncic = Ci / 32; P.ncoc = Co / 32;
    P.SL = plan.SL; P.SK2 = plan.SK2; P.SK1 = plan.SK1;
    const int64_t groups = (int64_t)N * P.ncic * P.ncoc;
    const int64_t units = groups * (P.SL + 2 * P.SK2 + P.SK1);
    if (units > 0x7fffffff / 2) MODE_FAIL("wgrad_split: too many work units");
    P.nL = (int)(groups * P.SL);
    P.nK2 = (int)(groups * 2 * P.SK2);
    P.tiles_hK = (int)ceil_div(H + 3, ws::TH);
    P.tiles_hL = (int)ceil_div(H, ws::TH);
    P.tiles_w = W / ws::TW;
    P.error_flag = device_error_flag();
    if (!P.error_flag) MODE_FAIL("wgrad_split: could not allocate the device error flag");
    CUtensorMap xmap, dymapK, dymapL;
    if (make_act_map(&xmap, x, N, D, H, W, Ci, ws::X_COLS, ws::TH, 1) != 0) return -1;
    if (make_act_map(&dymapK, dy, N, D, H, W, Co, ws::TW, ws::K_DY_ROWS, 1) != 0) return -1;
    if (make_act_map(&dymapL, dy, N, D, H, W, Co, ws::TW, ws::TH, ws::L_PLANES) != 0) return -1;
    const int smem_bytes = ws::OPERAND_BYTES + 512 + 1024;
    MODE_CUDA(cudaFuncSetAttribute(wgrad_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    wgrad_split_kernel<<<(unsigned)units, ws::THREADS, smem_bytes, st>>>(xmap, dymapK, dymapL, P);
    MODE_LAUNCH_CHECK();
    const int64_t total = (int64_t)N * 125 * Co * Ci / 4;
    if (total > 0x7fffffff) MODE_FAIL("wgrad_split: d_weff too large for 32-bit indexing");
    const int grid = (int)std::max((int64_t)1, std::min(ceil_div(total, 256), (int64_t)sm_count() * 16));
    wgrad_split_reduce_kernel<<<grid, 256, 0, st>>>((const float*)workspace, dw, N, Ci, Co, P.SL, P.SK2, P.SK1, P.nL,
                                                    P.nK2, out_scale, out_scale_dev);
    MODE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mode
