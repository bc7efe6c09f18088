// K4 (default): wgrad of the 5x5x5 conv on tcgen05, split-tap cover with DEEP tiles.
//   d_weff[n][tap][o][i] = sum_p dy[n][p][o] * x[n][p + tap - 2][i]      (autograd of RepMode.py:207)
// Same arithmetic and tap cover as wgrad_split.cu (the round-1 kernel, kept as the A/B arm REPMODE_WGRAD_SPLIT=1 / impl 5):
// taps are stacked through overlapping MN-major views of the same shared-memory bricks -- A = dy brick seen through 4 M
// blocks (row shifts -> kh, or plane shifts -> kd for the kh = 4 units), B = x brick seen through 5 N blocks (voxel shifts ->
// kw) -- so one M128 N160 K16 MMA accumulates 20 taps.
//
// Why deep tiles: the r1g capture of wgrad_split_kernel fits a per-SM TMA model of ~300 cycles per bulk-tensor request plus
// ~48 B/clk (profiles/README.md): a tile of 2-3 requests / 22-35 KB feeding only 8-16 MMAs (648-1296
// cycles of tensor work) is REQUEST-bound -- the K units ran at 115-140 cycles per MMA -- while the L units (3 requests,
// 74 KB, 32 MMAs) ran at the tensor pipe's own 80.  So every unit kind here gets >= 32 MMAs per 2 requests:
//   A units: kh = 0..3 of kd 0,1   -- tile = 2 dy planes (one 2-plane box) x the 3 x planes they meet (one box): 32 MMAs
//   B units: kh = 0..3 of kd 2,3,4 -- tile = 2 dy planes x 4 x planes: 48 MMAs, three accumulator sets (480 TMEM columns)
//   L units: kh = 4 of all kd      -- as wgrad_split.cu, the two x planes in ONE box: 32 MMAs
// Planes outside the volume are TMA zero fill and are multiplied like any other (~3 % extra MMAs) so that every tile of
// a kind runs the same branch-free, fully unrolled, warp-uniform issue body.
// Measured (r2a, headline layer 32 -> 32 @ 32x128x128): 109 us + 4 us reduce against 128 + 7 for wgrad_split.cu.
// Roofline: tensor pipe; algorithmic work 2*125*Ci*Co FLOP per voxel.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace mode {

using namespace sm100;

namespace wd {
constexpr int TW = 8, TH = 16, X_COLS = TW + 4;
constexpr int PT = 2;                                // dy planes per K tile
constexpr int X_PLANE = TH * X_COLS * 64;            // 12288
constexpr int K_DY_ROWS = TH + 3;
constexpr int K_DY_PLANE = K_DY_ROWS * TW * 64;      // 9728 (a multiple of the 512-byte swizzle atom)
constexpr int K_DY_SLOT = 20480;                     // PT planes, padded to 1 KB
constexpr int A_STAGE = K_DY_SLOT + 3 * X_PLANE;     // 57344
constexpr int B_STAGE = K_DY_SLOT + 4 * X_PLANE;     // 69632
constexpr int L_PLANE = TH * TW * 64;                // 8192
constexpr int L_PLANES = 6;
constexpr int L_DY_BYTES = L_PLANES * L_PLANE;       // 49152
constexpr int L_STAGE = L_DY_BYTES + 2 * X_PLANE;    // 73728
constexpr int STAGES = 3;
constexpr int OPERAND_BYTES = STAGES * L_STAGE;      // the largest kind
constexpr int THREADS = 256;
constexpr int ENTRIES = 60;                          // [32 co][32 ci] fp32 blocks per unit partial (B units: 3 kd x 20)
constexpr int PARTIAL_FLOATS = ENTRIES * 32 * 32;
static_assert(PARTIAL_FLOATS == K4_DEEP_PARTIAL_FLOATS, "K1b (reparam.cu) reads the partials through this layout");
static_assert(PT * K_DY_PLANE <= K_DY_SLOT && K_DY_PLANE % 512 == 0, "dy slot");
static_assert(A_STAGE % 1024 == 0 && B_STAGE % 1024 == 0 && L_STAGE % 1024 == 0, "swizzle atoms");
static_assert(STAGES * B_STAGE <= OPERAND_BYTES && STAGES * A_STAGE <= OPERAND_BYTES, "operand area");
static_assert(OPERAND_BYTES + 512 + 1024 <= 227 * 1024, "shared memory budget");
}  // namespace wd

struct DeepParams {
    float* partial;                 // [units][ENTRIES][32][32]
    int N, D, H, W, Ci, Co;
    int Dx, x_off;                  // haloed x: Dx planes, dy plane p is centred on x plane p + x_off (Dx = D, 0 without halo)
    int ncic, ncoc;
    int SL, SA, SB;                 // slabs per unit group of each kind
    int nL, nA;                     // unit counts: grid = [L units][A units][B units]
    int tiles_hK, tiles_hL, tiles_w;
    int* error_flag;
};

// cumulative cost (two-row K steps) of the first t tiles of a (plane group, tile row, tile column) walk whose last tile row
// needs nk_last instead of 8 steps -- same helper as wgrad_split.cu
__device__ __forceinline__ int64_t wd_cum_steps(int t, int tiles_h, int tiles_w, int nk_last) {
    const int per_plane = tiles_h * tiles_w;
    const int td = t / per_plane, rem = t - td * per_plane;
    const int th = rem / tiles_w, tw = rem - th * tiles_w;
    const int64_t plane_cost = (int64_t)tiles_w * (8 * (tiles_h - 1) + nk_last);
    return td * plane_cost + (int64_t)th * 8 * tiles_w + (int64_t)tw * (th == tiles_h - 1 ? nk_last : 8);
}
__device__ __forceinline__ int wd_slab_cut(int k, int S, int tiles, int tiles_h, int tiles_w, int nk_last) {
    if (k <= 0) return 0;
    if (k >= S) return tiles;
    const int64_t target = wd_cum_steps(tiles, tiles_h, tiles_w, nk_last) * k / S;
    int lo = 0, hi = tiles;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (wd_cum_steps(mid, tiles_h, tiles_w, nk_last) >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__global__ void __launch_bounds__(wd::THREADS, 1)
wgrad_deep_kernel(const __grid_constant__ CUtensorMap dymapK, const __grid_constant__ CUtensorMap xmapA,
                  const __grid_constant__ CUtensorMap xmapB, const __grid_constant__ CUtensorMap dymapL,
                  const __grid_constant__ CUtensorMap xmapL, const DeepParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + wd::OPERAND_BYTES;
    const uint32_t full = bars, empty = bars + 8 * wd::STAGES, done = empty + 8 * wd::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + wd::OPERAND_BYTES + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- decode the unit: kind (0 = L, 1 = A: kd 0,1, 2 = B: kd 2,3,4), channel chunks, slab ----------------------
    int u = blockIdx.x;
    int kind, s, S;
    if (u < P.nL) { kind = 0; S = P.SL; }
    else if (u < P.nL + P.nA) { kind = 1; u -= P.nL; S = P.SA; }
    else { kind = 2; u -= P.nL + P.nA; S = P.SB; }
    s = u % S; u /= S;
    const int cic = u % P.ncic; u /= P.ncic;
    const int coc = u % P.ncoc; u /= P.ncoc;
    const int n = u;
    const int kd0 = kind == 2 ? 2 : 0, nkd = kind == 2 ? 3 : 2;

    // ---- the unit's tile walk (plane group, then tile row, then tile column) and this slab's share of it ----------
    int dlo, ngroups, tiles_h, nk_last;
    if (kind == 0) {
        // x planes (dy coordinates) that meet a dy plane through a kh = 4 tap of some kd: [-2, D + 1] clipped to the tensor
        dlo = max(-P.x_off, -2);
        const int phi = min(P.Dx - P.x_off - 1, P.D + 1);
        ngroups = (max(0, phi - dlo + 1) + 1) >> 1;                     // pairs of x planes
        tiles_h = P.tiles_hL;
        nk_last = min(8, (P.H - (tiles_h - 1) * wd::TH + 1) >> 1);
    } else {
        // dy planes d whose x plane d + kd - 2 lies inside the volume for at least one kd of the group, in pairs
        dlo = max(0, 2 - (kd0 + nkd - 1) - P.x_off);
        const int dhi = min(P.D, P.Dx + 2 - P.x_off - kd0);
        ngroups = (max(0, dhi - dlo) + wd::PT - 1) / wd::PT;
        tiles_h = P.tiles_hK;
        nk_last = min(8, (P.H - ((tiles_h - 1) * wd::TH - 3) + 1) >> 1);
    }
    const int tiles = ngroups * tiles_h * P.tiles_w;
    const int t0 = wd_slab_cut(s, S, tiles, tiles_h, P.tiles_w, nk_last);
    const int t1 = wd_slab_cut(s + 1, S, tiles, tiles_h, P.tiles_w, nk_last);
    const uint32_t stage_bytes = kind == 0 ? wd::L_STAGE : (kind == 1 ? wd::A_STAGE : wd::B_STAGE);

    if (threadIdx.x == 0) {
        for (int i = 0; i < wd::STAGES; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
    if (warp == 0 && lane == 0) {
        if (kind == 0) { tma_prefetch_desc(&dymapL); tma_prefetch_desc(&xmapL); }
        else { tma_prefetch_desc(&dymapK); tma_prefetch_desc(kind == 1 ? &xmapA : &xmapB); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: two requests per tile =====================
        if (lane == 0) {
            uint32_t st = 0, use = 0;
            for (int t = t0; t < t1; ++t) {
                const int tw = t % P.tiles_w, th = (t / P.tiles_w) % tiles_h, tg = t / (P.tiles_w * tiles_h);
                const int vw0 = tw * wd::TW;
                if (!mbar_wait(empty + 8 * st, (use & 1) ^ 1)) { atomicExch(P.error_flag, 31); return; }
                const uint32_t dst = base + st * stage_bytes;
                if (kind == 0) {
                    const int p = dlo + 2 * tg, vh0 = th * wd::TH;
                    mbar_expect_tx(full + 8 * st, wd::L_DY_BYTES + 2 * wd::X_PLANE);
                    tma_load_5d(dst, &dymapL, full + 8 * st, coc * 32, vw0, vh0, p - 2, n);                    // dy planes p-2 .. p+3
                    tma_load_5d(dst + wd::L_DY_BYTES, &xmapL, full + 8 * st, cic * 32, vw0 - 2, vh0 + 2, p + P.x_off, n); // x planes p, p+1
                } else {
                    const int d = dlo + wd::PT * tg, vh0 = th * wd::TH - 3;
                    const int nxp = wd::PT + nkd - 1;
                    mbar_expect_tx(full + 8 * st, wd::PT * wd::K_DY_PLANE + nxp * wd::X_PLANE);
                    tma_load_5d(dst, &dymapK, full + 8 * st, coc * 32, vw0, vh0, d, n);                        // dy planes d, d+1
                    tma_load_5d(dst + wd::K_DY_SLOT, kind == 1 ? &xmapA : &xmapB, full + 8 * st, cic * 32, vw0 - 2,
                                vh0 + 1, d + kd0 - 2 + P.x_off, n);                                            // x planes d+kd0-2 ..
                }
                if (++st == wd::STAGES) { st = 0; ++use; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: warp-uniform loop, elected lane predicated in the asm =====================
        {
            const uint32_t sel = elect_one() ? 1u : 0u;
            const uint32_t tm = __reduce_max_sync(0xffffffffu, tmem);
            const uint32_t hi_a = (512u >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);
            const uint32_t hi_b = ((wd::X_COLS * 64u) >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);
            const uint32_t lbo_row = (512u >> 4) << 16;                      // M block bm = dy brick + bm rows
            const uint32_t lbo_plane = ((uint32_t)wd::L_PLANE >> 4) << 16;   // M block bm = dy plane + bm
            const uint32_t lbo_vox = (64u >> 4) << 16;                       // N block bn = x brick + bn voxels
            const uint32_t idesc = make_idesc(FMT_F16, 128, 160, 1, 1);
            uint32_t st = 0, use = 0, acc = 0;
            int th_c = (t0 / P.tiles_w) % tiles_h, tw_c = t0 % P.tiles_w;
            bool ok = true;
// Descriptor low words are (base >> 4 | LBO field) + compile-time offsets: ONE uniform add per operand and MMA
            // (the address field is 14 bits and every offset keeps it below 2^14, so the add never carries into the LBO
            // field).  r2a capture: with ((base + off) >> 4) | lbo spelled out per MMA the issue warp ran ~11 uniform
            // instructions per UTCHMMA and was itself the limiter (tensor pipe 80 % of active cycles, MMA warp waiting
            // for operands only 7 % of its samples).
#define WD_OFF(bytes) ((uint32_t)(bytes) >> 4)
#define WD_MMA(col, alo, blo, a) mma_f16_ss_sel(tm + (col), (alo), hi_a, (blo), hi_b, idesc, (a), sel)
            for (int t = t0; t < t1; ++t) {
                if (!mbar_wait_warp<false>(full + 8 * st, use & 1)) { ok = false; break; }
                tc_fence_after();
                const uint32_t sb = base + st * stage_bytes;
                const bool full_row = th_c != tiles_h - 1 || nk_last == 8;
                if (kind == 0) {
                    const uint32_t ap = (sb >> 4) | lbo_plane;                                   // dy plane 0 of the box
                    const uint32_t bx = ((sb + wd::L_DY_BYTES) >> 4) | lbo_vox;                  // x plane p
                    if (full_row) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            // x plane p: dy planes p-1.. (kd = 3 - bm) and p-2.. (bm = 0: kd = 4); x plane p+1: one plane on
                            const uint32_t b0 = bx + WD_OFF(2 * kk * wd::X_COLS * 64);
                            const uint32_t b1 = bx + WD_OFF(wd::X_PLANE + 2 * kk * wd::X_COLS * 64);
                            const uint32_t a0 = ap + WD_OFF(0 * wd::L_PLANE + kk * 1024);
                            const uint32_t a1 = ap + WD_OFF(1 * wd::L_PLANE + kk * 1024);
                            const uint32_t a2 = ap + WD_OFF(2 * wd::L_PLANE + kk * 1024);
                            WD_MMA(0, a1, b0, acc);
                            WD_MMA(160, a0, b0, acc);
                            WD_MMA(0, a2, b1, 1u);
                            WD_MMA(160, a1, b1, 1u);
                            acc = 1;
                        }
                    } else {
#pragma unroll 1
                        for (int kk = 0; kk < nk_last; ++kk) {
                            const uint32_t b0 = bx + WD_OFF(2 * kk * wd::X_COLS * 64);
                            const uint32_t b1 = b0 + WD_OFF(wd::X_PLANE);
                            const uint32_t a0 = ap + WD_OFF(kk * 1024);
                            const uint32_t a1 = a0 + WD_OFF(wd::L_PLANE);
                            const uint32_t a2 = a0 + WD_OFF(2 * wd::L_PLANE);
                            WD_MMA(0, a1, b0, acc);
                            WD_MMA(160, a0, b0, acc);
                            WD_MMA(0, a2, b1, 1u);
                            WD_MMA(160, a1, b1, 1u);
                            acc = 1;
                        }
                    }
                } else {
                    // dy plane pl (row-shifted M blocks: kh = 3 - bm) x x plane pl + j  ->  kd = kd0 + j, accumulator set j
                    const uint32_t ar = (sb >> 4) | lbo_row;
                    const uint32_t bx = ((sb + wd::K_DY_SLOT) >> 4) | lbo_vox;
                    if (full_row && kind == 1) {
#pragma unroll
                        for (int pl = 0; pl < wd::PT; ++pl) {
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk) {
                                const uint32_t ad = ar + WD_OFF(pl * wd::K_DY_PLANE + kk * 1024);
                                const uint32_t bd = bx + WD_OFF(pl * wd::X_PLANE + 2 * kk * wd::X_COLS * 64);
                                WD_MMA(0, ad, bd, (pl | kk) ? 1u : acc);
                                WD_MMA(160, ad, bd + WD_OFF(wd::X_PLANE), (pl | kk) ? 1u : acc);
                            }
                        }
                    } else if (full_row) {
#pragma unroll
                        for (int pl = 0; pl < wd::PT; ++pl) {
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk) {
                                const uint32_t ad = ar + WD_OFF(pl * wd::K_DY_PLANE + kk * 1024);
                                const uint32_t bd = bx + WD_OFF(pl * wd::X_PLANE + 2 * kk * wd::X_COLS * 64);
                                WD_MMA(0, ad, bd, (pl | kk) ? 1u : acc);
                                WD_MMA(160, ad, bd + WD_OFF(wd::X_PLANE), (pl | kk) ? 1u : acc);
                                WD_MMA(320, ad, bd + WD_OFF(2 * wd::X_PLANE), (pl | kk) ? 1u : acc);
                            }
                        }
                    } else {
#pragma unroll 1
                        for (int pl = 0; pl < wd::PT; ++pl) {
#pragma unroll 1
                            for (int kk = 0; kk < nk_last; ++kk) {
                                const uint32_t ad = ar + WD_OFF(pl * wd::K_DY_PLANE + kk * 1024);
                                const uint32_t bd = bx + WD_OFF(pl * wd::X_PLANE + 2 * kk * wd::X_COLS * 64);
                                const uint32_t a = (pl | kk) ? 1u : acc;
#pragma unroll 1
                                for (int j = 0; j < nkd; ++j) WD_MMA(160 * j, ad, bd + j * WD_OFF(wd::X_PLANE), a);
                            }
                        }
                    }
                    acc = 1;
                }
                mma_commit_sel(empty + 8 * st, sel);
                if (++st == wd::STAGES) { st = 0; ++use; }
                if (++tw_c == P.tiles_w) { tw_c = 0; if (++th_c == tiles_h) th_c = 0; }
            }
#undef WD_OFF
#undef WD_MMA
            if (ok) mma_commit_sel(done, sel);
            else if (lane == 0) atomicExch(P.error_flag, 32);
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> fp32 partial =====================
        const int bm = warp - 4;                 // TMEM lanes 32*bm .. : M block bm, lane = output channel
        float* out = P.partial + (size_t)blockIdx.x * wd::PARTIAL_FLOATS;
        bool ok = true;
        if (t1 > t0) {
            ok = mbar_wait(done, 0);
            if (!ok && lane == 0) atomicExch(P.error_flag, 33);
            tc_fence_after();
        }
        const uint32_t lane_addr = (uint32_t)(bm * 32) << 16;
        const int nsets = kind == 0 ? 2 : nkd;
        const bool have = t1 > t0 && ok;           // every set of a unit receives MMAs in its first tile
        for (int set = 0; set < nsets; ++set) {
            int entry0;
            if (kind == 0) {
                if (set == 1 && bm != 0) continue;                     // MMA 2: only plane block 0 (kd = 4) is a tap
                entry0 = (set == 0 ? 3 - bm : 4) * 5;                  // entry = kd*5 + kw
            } else {
                entry0 = set * 20 + (3 - bm) * 5;                      // entry = set*20 + kh*5 + kw
            }
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
__global__ void __launch_bounds__(256) wgrad_deep_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                int N, int Ci, int Co, int SL, int SA, int SB, int nL, int nA,
                                                                float out_scale, const float* __restrict__ out_scale_dev) {
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
        else if (kd < 2u) { unit0 = nL + g * SA; S = SA; entry = kd * 20u + kh * 5u + kw; }
        else { unit0 = nL + nA + g * SB; S = SB; entry = (kd - 2u) * 20u + kh * 5u + kw; }
        const float* src = partial + (size_t)unit0 * wd::PARTIAL_FLOATS + ((size_t)entry * 32 + (o & 31)) * 32 + (i & 31);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < S; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)s * wd::PARTIAL_FLOATS);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
        *reinterpret_cast<float4*>(dw + (size_t)idx * 4) = acc;
    }
}

// ------------------------------------------------------------------------------------------------ host
int* device_error_flag();   // mode_abi.cu
int make_act_map(CUtensorMap* map, const __half* x, int N, int D, int H, int W, int K, int box_w, int box_h, int box_d);  // conv_umma.cu

struct DeepPlan { int SL, SA, SB; };

// Slabs per unit kind: one wave of CTAs when the layer has few channel chunks, every CTA carrying about the same number
// of MMAs (A: 2 kd, B: 3 kd of the row taps per dy plane; L: 2 MMAs per K step and x plane).
static DeepPlan deep_plan(int N, int D, int H, int W, int Ci, int Co, int Dx, int x_off) {
    const int groups = N * (Ci / 32) * (Co / 32);
    const int slots = std::max(1, sm_count() / groups);
    const int tiles_w = W / wd::TW;
    const int tiles_hK = (int)ceil_div(H + 3, wd::TH), tiles_hL = (int)ceil_div(H, wd::TH);
    const double stepsK = tiles_w * (8.0 * (tiles_hK - 1) + std::min(8, (H - ((tiles_hK - 1) * wd::TH - 3) + 1) >> 1));
    const double stepsL = tiles_w * (8.0 * (tiles_hL - 1) + std::min(8, (H - (tiles_hL - 1) * wd::TH + 1) >> 1));
    auto planes = [&](int kd0, int nkd) {
        const int dlo = std::max(0, 2 - (kd0 + nkd - 1) - x_off), dhi = std::min(D, Dx + 2 - x_off - kd0);
        return (double)(((std::max(0, dhi - dlo) + wd::PT - 1) / wd::PT) * wd::PT);
    };
    const double cA = planes(0, 2) * 2 * stepsK, cB = planes(2, 3) * 3 * stepsK;
    const int lplanes = std::max(0, std::min(Dx - x_off - 1, D + 1) - std::max(-x_off, -2) + 1);
    const double cL = 2.0 * (2 * ((lplanes + 1) / 2)) * stepsL;
    DeepPlan p{1, 1, 1};
    while (p.SA + p.SB + p.SL < slots) {
        const double a = cA / p.SA, b = cB / p.SB, c = cL / p.SL;
        if (a >= b && a >= c) ++p.SA;
        else if (b >= c) ++p.SB;
        else ++p.SL;
    }
    return p;
}

bool wgrad_deep_supported(int D, int H, int W, int Ci, int Co) {
    (void)D; (void)H;
    return Ci % 32 == 0 && Co % 32 == 0 && Ci >= 32 && Co >= 32 && W % wd::TW == 0;
}

// {SL, SA, SB, nL, nA}: slabs per unit group of each kind and the first unit of the A / B groups -- what a reader of the
// partials needs (the reduce kernel below, or K1b when every S is 1)
void wgrad_deep_layout(int N, int D, int H, int W, int Ci, int Co, int Dx, int x_off, int32_t out[5]) {
    const DeepPlan p = deep_plan(N, D, H, W, Ci, Co, Dx, x_off);
    const int64_t groups = (int64_t)N * (Ci / 32) * (Co / 32);
    out[0] = p.SL; out[1] = p.SA; out[2] = p.SB;
    out[3] = (int32_t)(groups * p.SL); out[4] = (int32_t)(groups * p.SA);
}

int64_t wgrad_deep_workspace_bytes(int N, int D, int H, int W, int Ci, int Co, int Dx, int x_off) {
    const DeepPlan p = deep_plan(N, D, H, W, Ci, Co, Dx, x_off);
    const int64_t groups = (int64_t)N * (Ci / 32) * (Co / 32);
    return groups * (p.SL + p.SA + p.SB) * wd::PARTIAL_FLOATS * (int64_t)sizeof(float);
}

// phase: 0 = both kernels; 1 = the tensor-core kernel only (partials into the workspace); 2 = the reduce kernel only -- so
// that a caller can start K3 (which needs nothing from K4) between the two (functional.py)
int wgrad_deep(const __half* x, const __half* dy, float* dw, int N, int D, int H, int W, int Ci, int Co,
               float out_scale, const float* out_scale_dev, void* workspace, int Dx, int x_off, cudaStream_t st,
               int phase) {
    if (!workspace) MODE_FAIL("wgrad_deep: workspace is NULL");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15))
        MODE_FAIL("wgrad_deep: pointers must be 16-byte aligned");
    if (x_off < 0 || x_off + D > Dx) MODE_FAIL("wgrad_deep: need 0 <= x_off and x_off + D <= Dx (x_off=%d D=%d Dx=%d)", x_off, D, Dx);
    const DeepPlan plan = deep_plan(N, D, H, W, Ci, Co, Dx, x_off);
    DeepParams P;
    P.partial = (float*)workspace;
    P.N = N; P.D = D; P.H = H; P.W = W; P.Ci = Ci; P.Co = Co;
    P.Dx = Dx; P.x_off = x_off;
    P.ncic = Ci / 32; P.ncoc = Co / 32;
    P.SL = plan.SL; P.SA = plan.SA; P.SB = plan.SB;
    const int64_t groups = (int64_t)N * P.ncic * P.ncoc;
    const int64_t units = groups * (P.SL + P.SA + P.SB);
    if (units > 0x7fffffff / 2) MODE_FAIL("wgrad_deep: too many work units");
    P.nL = (int)(groups * P.SL);
    P.nA = (int)(groups * P.SA);
    P.tiles_hK = (int)ceil_div(H + 3, wd::TH);
    P.tiles_hL = (int)ceil_div(H, wd::TH);
    P.tiles_w = W / wd::TW;
    P.error_flag = device_error_flag();
    if (!P.error_flag) MODE_FAIL("wgrad_deep: could not allocate the device error flag");
    CUtensorMap dymapK, xmapA, xmapB, dymapL, xmapL;
    if (make_act_map(&dymapK, dy, N, D, H, W, Co, wd::TW, wd::K_DY_ROWS, wd::PT) != 0) return -1;
    if (make_act_map(&xmapA, x, N, Dx, H, W, Ci, wd::X_COLS, wd::TH, wd::PT + 1) != 0) return -1;
    if (make_act_map(&xmapB, x, N, Dx, H, W, Ci, wd::X_COLS, wd::TH, wd::PT + 2) != 0) return -1;
    if (make_act_map(&dymapL, dy, N, D, H, W, Co, wd::TW, wd::TH, wd::L_PLANES) != 0) return -1;
    if (make_act_map(&xmapL, x, N, Dx, H, W, Ci, wd::X_COLS, wd::TH, 2) != 0) return -1;
    const int smem_bytes = wd::OPERAND_BYTES + 512 + 1024;
    if (phase != 2) {
        MODE_CUDA(cudaFuncSetAttribute(wgrad_deep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        wgrad_deep_kernel<<<(unsigned)units, wd::THREADS, smem_bytes, st>>>(dymapK, xmapA, xmapB, dymapL, xmapL, P);
        MODE_LAUNCH_CHECK();
    }
    if (phase == 1) return 0;
    const int64_t total = (int64_t)N * 125 * Co * Ci / 4;
    if (total > 0x7fffffff) MODE_FAIL("wgrad_deep: d_weff too large for 32-bit indexing");
    const int grid = (int)std::max((int64_t)1, std::min(ceil_div(total, 256), (int64_t)sm_count() * 16));
    wgrad_deep_reduce_kernel<<<grid, 256, 0, st>>>((const float*)workspace, dw, N, Ci, Co, P.SL, P.SA, P.SB, P.nL, P.nA,
                                                   out_scale, out_scale_dev);
    MODE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mode
