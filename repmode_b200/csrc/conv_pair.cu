// K2 / K3, CTA-pair version: the 5x5x5 'same' convolution (and, with K1's dgrad pack, its input gradient) on
// tcgen05.mma.cta_group::2 -- two SMs of a cluster execute ONE M = 256 MMA, each bringing its own 128 A rows and
// half of the B rows.  Replaces F.conv3d(x[i:i+1], w[i], padding='same') (fnet/nn_modules/RepMode.py:204-210 in
// the reference tree) exactly like conv_umma.cu; this file is the fast path for large volumes.
//
// Why a pair: measured on B200 (tools/probe_pair.py, profiles/probe_r1.md) a single-CTA SS-mode MMA costs
// max(93, 42 + N/2) cycles -- the 4 KB A read is exposed -- while the pair instruction costs max(44, N/2):
// N = 160 runs at 80 cycles instead of 122, i.e. at the full tensor rate.
//
// Mapping
//   * the pair owns an 8(w) x 32(h) patch column: CTA rank r computes the 8 x 16 patch at h0 + 16 r (its 128 A rows).
//   * "march": the pair walks its d-range plane by plane.  Input plane p feeds output planes q = p-2 .. p+2 with the
//     kd taps in DESCENDING order (K1's stage-major pack), so EVERY plane issues the same N = 5 * 32 = 160 wide MMA
//     into five neighbouring 32-column accumulator slots of TMEM.  Output planes outside the run's [da, db) -- and
//     outside the volume -- are ordinary slots whose content is thrown away ("garbage slots"); no partial windows,
//     hence the B operand splits 80 + 80 rows over the pair at a fixed address.
//   * accumulator slots are circular with period 12 (+4 overflow slots so a window never wraps: a plane whose window
//     starts in slots 8..11 spills into slots 12..15, which the epilogue adds to slots 0..3 when it drains them).
//     Planes are processed in groups of 4; after a group the 4 output planes it completed are drained and re-zeroed
//     by the epilogue warps of BOTH CTAs while the leader already issues the next group (12 = 4 draining + 8 live).
//   * only planes inside the volume are ever loaded or multiplied; a run that starts or stops inside the volume pays
//     2 extra planes on that side (the only redundant work).
// Roles per CTA (256 threads): warp 0 activation-plane TMA producer, warp 3 weight producer (each CTA copies its
//   half of every weight stage), warp 2 TMEM allocator, warps 4-7 epilogue.  Warp 1: in the leader (cluster rank 0)
//   the single MMA-issuing thread; in the peer a relay that forwards "my plane / my weight half has landed" to the
//   leader's barriers.  Slot release (tcgen05.commit) is multicast to both CTAs.
// Roofline: tensor-pipe bound; algorithmic work 2*125*K*Nout FLOP per output voxel.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace mode {

using namespace sm100;

namespace cp {
constexpr int TW = 8, TH = 16;                 // output patch per CTA
constexpr int BW = TW + 4, BH = TH + 4;        // haloed brick
constexpr int ROWB = 64;                       // bytes per voxel row (32 fp16 channels)
constexpr int PLANE_BYTES = BH * BW * ROWB;    // 15360
constexpr int NT = 32;                         // output channels per pass (one accumulator slot = 32 columns)
constexpr int GROUP = 4;                       // input planes per accumulator hand-off group
constexpr int WST_BYTES = 5 * NT * ROWB / 2;   // this CTA's half of one (chunk, kh, kw) stage: 80 rows = 5120 B
constexpr int PERIOD = 12;                     // circular accumulator slots (+4 overflow)
constexpr int THREADS = 256;
constexpr int MAX_CLUSTERS = 80;
constexpr int MAX_RING = 12, MAX_WST = 25;
// Two shared-memory plans.  RESIDENT (K == 32): this CTA's half of ALL 25 (kh, kw) stages stays in shared memory
// (125 KB, loaded once per sample) and planes are consumed one at a time (25 taps x 2 K-halves each), so a 6-deep plane
// ring is enough.  STREAMING (K > 32): weight stages are re-streamed per group of 4 planes through a 6-stage ring and the
// plane ring is 12 deep (3 groups).
template <bool RES> struct Plan {
    static constexpr int RING = RES ? 6 : 12;
    static constexpr int WST = RES ? 25 : 6;
    static constexpr uint32_t PLANE_OFF = 0;
    static constexpr uint32_t W_OFF = RING * PLANE_BYTES;
    static constexpr uint32_t BAR_OFF = W_OFF + WST * WST_BYTES;
    static constexpr uint32_t BN_OFF = BAR_OFF + 1024;
    static constexpr uint32_t EP_OFF = BN_OFF + 2 * NT * sizeof(double);      // epilogue affine: scale[NT], shift[NT]
    static constexpr uint32_t TOTAL = EP_OFF + 2 * NT * sizeof(float);
};
}  // namespace cp

struct PairParams {
    const __half* w;             // stage-major fp16 pack, all Nout rows
    const int32_t* sample_u;
    float* y;                    // [N,D,H,W,Nout]
    double* bn_sums;             // [2*Nout] or null
    const float* out_scale_dev;
    float out_scale;
    int N, D, H, W, K, Nout;
    int stat_lo, stat_hi;
    int tiles_w, tiles_h2;       // pair patches: 8 wide, 32 high
    int32_t bounds[cp::MAX_CLUSTERS + 1];   // cluster c owns plane-units [bounds[c], bounds[c+1]) of (n, th2, tw, d)
    ConvExt ext;                 // haloed input (Dx, x_off) and fused epilogue (affine + ReLU, fp16 copy)
    int p_lo, p_hi;              // valid input planes in output-plane coordinates: [-x_off, Dx - x_off - 1]
    // K > 32 runs as one RESIDENT-plan launch per 32-channel chunk: this launch reads input channels [x_chan0, x_chan0 + 32)
    // and the weight chunk P.w points at, and (accum) adds its result to what the previous chunk's launch left in y
    int x_chan0, accum;
    long long w_u_stride;        // elements between the weight sets of two gate inputs (covers ALL chunks)
    int* error_flag;
    long long* prof;
    int flags;                   // debug: bit 1 = force the streaming plan even when K == 32
};

// A "run" is a contiguous d-range [da, db) of one patch column; every role of both CTAs walks the same runs.
struct RunWalker {
    int64_t u, uend;
    int D, tiles_w, tiles_h2;
    __device__ RunWalker(const PairParams& P, int cluster)
        : u(P.bounds[cluster]), uend(P.bounds[cluster + 1]), D(P.D), tiles_w(P.tiles_w), tiles_h2(P.tiles_h2) {}
    __device__ bool next(int& n, int& h0, int& w0, int& da, int& db) {
        if (u >= uend) return false;
        const int64_t col = u / D;
        da = (int)(u - col * D);
        db = (int)min((int64_t)D, da + (uend - u));
        u += db - da;
        w0 = (int)(col % tiles_w) * cp::TW;
        h0 = (int)((col / tiles_w) % tiles_h2) * (2 * cp::TH);
        n = (int)(col / ((int64_t)tiles_w * tiles_h2));
        return true;
    }
};

template <bool RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(cp::THREADS, 1)
conv3d_pair_kernel(const __grid_constant__ CUtensorMap xmap, const PairParams P) {
    using PL = cp::Plan<RES>;
    constexpr int RING = PL::RING, WST = PL::WST;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + PL::BAR_OFF;
    // barrier table (8 bytes each).  The leader's *_full barriers take two arrivals: its own producer's expect_tx and the
    // peer relay's remote arrive (the operands live in both CTAs); tmem_empty is only used in the leader.  In the RESIDENT
    // plan w_full[kh] covers the five stages of tap row kh (so the first plane can start before all 125 KB landed) and
    // only w_empty[0] is used (one release per weight reload).
    const uint32_t plane_full = bars, plane_empty = bars + 8 * cp::MAX_RING;
    const uint32_t w_full = bars + 16 * cp::MAX_RING, w_empty = w_full + 8 * cp::MAX_WST;
    const uint32_t tmem_full = w_full + 16 * cp::MAX_WST, tmem_empty = tmem_full + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + PL::BAR_OFF + 1000);
    double* s_bn = reinterpret_cast<double*>(smem + PL::BN_OFF);
    float* s_ep = reinterpret_cast<float*>(smem + PL::EP_OFF);
    static_assert(16 * cp::MAX_RING + 16 * cp::MAX_WST + 32 <= 1000, "barrier table overflows its 1 KB");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x >> 1;
    const int n0 = blockIdx.y * cp::NT;
    const int nchunk = P.K / 32;
    const uint64_t ts_entry = (P.prof != nullptr && threadIdx.x == 0) ? globaltimer_ns() : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) {
            mbar_init(plane_full + 8 * i, rank == 0 ? 2 : 1); mbar_init(plane_empty + 8 * i, 1);
        }
        for (int i = 0; i < (RES ? 5 : WST); ++i) {
            mbar_init(w_full + 8 * i, rank == 0 ? 2 : 1); mbar_init(w_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, 8); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_pair<512>(smem_u32(tmem_slot));
    if (warp == 0 && lane == 0) tma_prefetch_desc(&xmap);
    for (int i = threadIdx.x; i < 2 * cp::NT; i += cp::THREADS) s_bn[i] = 0.0;
    const bool has_ep = P.ext.ep_scale != nullptr || P.ext.ep_shift != nullptr;
    if (has_ep && threadIdx.x < cp::NT) {
        s_ep[threadIdx.x] = P.ext.ep_scale ? P.ext.ep_scale[n0 + threadIdx.x] : 1.f;
        s_ep[cp::NT + threadIdx.x] = P.ext.ep_shift ? P.ext.ep_shift[n0 + threadIdx.x] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();              // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== activation-plane producer (own 8x16 patch, both CTAs) =====================
        // plane order = consumption order: run, group of 4 planes, K chunk, plane
        if (lane == 0) {
            uint32_t slot = 0, use = 0;
            RunWalker rw(P, cluster);
            int n, h0, w0, da, db;
            bool ok = true;
            while (ok && rw.next(n, h0, w0, da, db)) {
                const int p0 = max(da - 2, P.p_lo), p1 = min(db + 1, P.p_hi);
                for (int gp = p0; ok && gp <= p1; gp += cp::GROUP) {
                    const int gn = min(cp::GROUP, p1 - gp + 1);
                    for (int c = 0; ok && c < nchunk; ++c) {
                        for (int i = 0; i < gn; ++i) {
                            if (!mbar_wait(plane_empty + 8 * slot, (use & 1) ^ 1)) { atomicExch(P.error_flag, 11); ok = false; break; }
                            mbar_expect_tx(plane_full + 8 * slot, cp::PLANE_BYTES);
                            tma_load_5d(base + PL::PLANE_OFF + slot * cp::PLANE_BYTES, &xmap, plane_full + 8 * slot, c * 32 + P.x_chan0,
                                        w0 - 2, h0 + (int)rank * cp::TH - 2, gp + i + P.ext.x_off, n);
                            if (++slot == RING) { slot = 0; ++use; }
                        }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== weight producer: this CTA's 80 rows of every (chunk, kh, kw) stage =====================
        // stage rows are [kd = 4..0][32 channels]; rank 0 holds rows 0..79 (kd 4, kd 3, first half of kd 2), rank 1 the rest
        if (lane == 0) {
            uint32_t st = 0, use = 0;
            RunWalker rw(P, cluster);
            int n, h0, w0, da, db;
            bool ok = true;
            const size_t blk = (size_t)P.Nout * 32;            // elements between kd blocks in the pack
            auto copy_stage = [&](uint32_t dst, const __half* src, uint32_t bar) {
                if (rank == 0) {
                    bulk_load(dst, src, 2048, bar);
                    bulk_load(dst + 2048, src + blk, 2048, bar);
                    bulk_load(dst + 4096, src + 2 * blk, 1024, bar);
                } else {
                    bulk_load(dst, src + 2 * blk + 16 * 32, 1024, bar);
                    bulk_load(dst + 1024, src + 3 * blk, 2048, bar);
                    bulk_load(dst + 3072, src + 4 * blk, 2048, bar);
                }
            };
            int cur_u = -1;
            uint32_t nreload = 0;
            while (ok && rw.next(n, h0, w0, da, db)) {
                const int u = P.sample_u ? P.sample_u[n] : 0;
                const __half* wu = P.w + (size_t)u * P.w_u_stride;
                if constexpr (RES) {
                    if (u != cur_u) {                          // (re)load all 25 stages; the previous sample's MMAs must be done
                        if (nreload > 0 && !mbar_wait(w_empty, (nreload - 1) & 1)) { atomicExch(P.error_flag, 12); break; }
                        for (int t = 0; t < 25; ++t) {
                            const uint32_t bar = w_full + 8 * (t / 5);
                            if (t % 5 == 0) mbar_expect_tx(bar, 5 * cp::WST_BYTES);
                            copy_stage(base + PL::W_OFF + t * cp::WST_BYTES, wu + ((size_t)t * 5 * P.Nout + n0) * 32, bar);
                        }
                        cur_u = u;
                        ++nreload;
                    }
                } else {
                    const int p0 = max(da - 2, P.p_lo), p1 = min(db + 1, P.p_hi);
                    for (int gp = p0; ok && gp <= p1; gp += cp::GROUP) {
                        for (int c = 0; ok && c < nchunk; ++c) {
                            for (int t = 0; t < 25; ++t) {
                                if (!mbar_wait(w_empty + 8 * st, (use & 1) ^ 1)) { atomicExch(P.error_flag, 12); ok = false; break; }
                                mbar_expect_tx(w_full + 8 * st, cp::WST_BYTES);
                                copy_stage(base + PL::W_OFF + st * cp::WST_BYTES, wu + ((size_t)(c * 25 + t) * 5 * P.Nout + n0) * 32,
                                           w_full + 8 * st);
                                if (++st == WST) { st = 0; ++use; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && rank != 0) {
        // ===================== peer relay: forward "landed" events to the leader in the order it consumes them ==========
        if (lane == 0) {
            uint32_t pslot = 0, puse = 0, wst = 0, wuse = 0, nreload = 0;
            int cur_u = -1;
            RunWalker rw(P, cluster);
            int n, h0, w0, da, db;
            bool ok = true;
            while (ok && rw.next(n, h0, w0, da, db)) {
                const int p0 = max(da - 2, P.p_lo), p1 = min(db + 1, P.p_hi);
                if constexpr (RES) {
                    const int u = P.sample_u ? P.sample_u[n] : 0;
                    if (u != cur_u) {
                        for (int kh = 0; ok && kh < 5; ++kh) {
                            if (!mbar_wait(w_full + 8 * kh, nreload & 1)) { atomicExch(P.error_flag, 13); ok = false; break; }
                            mbar_arrive_remote(w_full + 8 * kh, 0);
                        }
                        if (!ok) break;
                        cur_u = u;
                        ++nreload;
                    }
                    for (int p = p0; p <= p1; ++p) {
                        if (!mbar_wait(plane_full + 8 * pslot, puse & 1)) { atomicExch(P.error_flag, 14); ok = false; break; }
                        mbar_arrive_remote(plane_full + 8 * pslot, 0);
                        if (++pslot == RING) { pslot = 0; ++puse; }
                    }
                } else {
                    for (int gp = p0; ok && gp <= p1; gp += cp::GROUP) {
                        const int gn = min(cp::GROUP, p1 - gp + 1);
                        for (int c = 0; ok && c < nchunk; ++c) {
                            for (int t = 0; ok && t < 25; ++t) {
                                if (!mbar_wait(w_full + 8 * wst, wuse & 1)) { atomicExch(P.error_flag, 13); ok = false; break; }
                                mbar_arrive_remote(w_full + 8 * wst, 0);
                                if (++wst == WST) { wst = 0; ++wuse; }
                                if (t == 0) {
                                    for (int i = 0; i < gn; ++i) {
                                        if (!mbar_wait(plane_full + 8 * pslot, puse & 1)) { atomicExch(P.error_flag, 14); ok = false; break; }
                                        mbar_arrive_remote(plane_full + 8 * pslot, 0);
                                        if (++pslot == RING) { pslot = 0; ++puse; }
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader only) =====================
        // The WHOLE warp runs this loop on warp-uniform values (kernel parameters, loop counters, vote results), and one
        // elected lane is predicated inside the asm.  Under `if (lane == 0)` ptxas has to assume divergence and wraps
        // every UTCHMMA in an ELECT / 5x R2UR.BROADCAST loop (~115 cycles per MMA measured -- more than the 80 cycles the
        // pair MMA itself takes); with uniform control flow the descriptors live in uniform registers.
        {
            const uint32_t sel = elect_one() ? 1u : 0u;
            const uint32_t tm = __reduce_max_sync(0xffffffffu, tmem);             // uniform copy of the TMEM base
            const uint32_t hi_a = ((cp::BW * cp::ROWB) >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);
            const uint32_t hi_b = (512u >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);
            const uint32_t lbo_lo = 1u << 16;
            const uint32_t idesc = make_idesc(FMT_F16, 256, 5 * cp::NT, 0, 0);
            const uint32_t plane0 = ((base + PL::PLANE_OFF) >> 4) | lbo_lo, wst0 = ((base + PL::W_OFF) >> 4) | lbo_lo;
            const bool prof = P.prof != nullptr;
            uint32_t pslot = 0, puse = 0, wst = 0, wuse = 0, nreload = 0;
            uint32_t g = 0;                                   // global group counter (accumulator hand-off parity)
            long long c_tmem = 0, c_w = 0, c_plane = 0, c_issue = 0, c_total = clock64(), t0 = 0;
            int u = P.bounds[cluster];
            const int uend = P.bounds[cluster + 1];
            int cur_u = -1;
            int err = 0;
            bool fresh = false;
            while (err == 0 && u < uend) {
                const int col = u / P.D;
                const int da = u - col * P.D;
                const int db = min(P.D, da + (uend - u));
                u += db - da;
                const int p0 = max(da - 2, P.p_lo), p1 = min(db + 1, P.p_hi);
                if constexpr (RES) {
                    const int n = col / (P.tiles_w * P.tiles_h2);
                    const int su = P.sample_u ? P.sample_u[n] : 0;
                    if (su != cur_u) {
                        // weights of another sample: release the resident copy once everything issued so far retires
                        if (nreload > 0) mma_commit_pair_sel(w_empty, 3u, sel);
                        cur_u = su;
                        fresh = true;                         // the first plane waits for each tap row's weights as it gets there
                        ++nreload;
                    }
                }
                uint32_t s = 0;                               // accumulator slot of the group's first window
                bool first = true;
                for (int gp = p0; err == 0 && gp <= p1; gp += cp::GROUP, ++g) {
                    const int gn = min(cp::GROUP, p1 - gp + 1);
                    // slots this group touches were drained two groups ago; a new run restarts at slot 0, so it
                    // also needs the drain of the group just before it (the wait group g+1 would do anyway)
                    if (prof) t0 = clock64();
                    if (!mbar_wait_warp<false>(tmem_empty + 8 * (g & 1), (g >> 1) & 1)) { err = 15; break; }
                    if (first && g > 0 && !mbar_wait_warp<false>(tmem_empty + 8 * ((g + 1) & 1), ((g + 1) >> 1) & 1)) { err = 16; break; }
                    first = false;
                    if (prof) c_tmem += clock64() - t0;
                    tc_fence_after();
                    if constexpr (RES) {
                        uint32_t ss = s;
                        for (int i = 0; i < gn; ++i) {
                            if (prof) t0 = clock64();
                            if (!mbar_wait_warp<false>(plane_full + 8 * pslot, puse & 1)) { err = 18; break; }
                            if (prof) c_plane += clock64() - t0;
                            tc_fence_after();
                            const long long ti = prof ? clock64() : 0;
                            const uint32_t a0 = plane0 + pslot * (cp::PLANE_BYTES >> 4);
                            const uint32_t dcol = tm + ss * cp::NT;
#pragma unroll
                            for (int t = 0; t < 25; ++t) {
                                if (t % 5 == 0 && fresh) {
                                    if (prof) t0 = clock64();
                                    if (!mbar_wait_warp<false>(w_full + 8 * (t / 5), (nreload - 1) & 1)) { err = 17; break; }
                                    if (prof) c_w += clock64() - t0;
                                    tc_fence_after();
                                }
                                const uint32_t al = a0 + ((((t / 5) * cp::BW + (t % 5)) * cp::ROWB) >> 4);
                                const uint32_t bl = wst0 + t * (cp::WST_BYTES >> 4);
                                mma_f16_ss_pair_sel(dcol, al, hi_a, bl, hi_b, idesc, sel);
                                mma_f16_ss_pair_sel(dcol, al + 2, hi_a, bl + 2, hi_b, idesc, sel);
                            }
                            if (err) break;
                            fresh = false;
                            mma_commit_pair_sel(plane_empty + 8 * pslot, 3u, sel);
                            if (prof) c_issue += clock64() - ti;
                            if (++pslot == RING) { pslot = 0; ++puse; }
                            if (++ss == cp::PERIOD) ss = 0;
                        }
                    } else {
                        for (int c = 0; err == 0 && c < nchunk; ++c) {
                            for (int t = 0; t < 25; ++t) {
                                const int kh = t / 5, kw = t - kh * 5;
                                if (prof) t0 = clock64();
                                if (!mbar_wait_warp<false>(w_full + 8 * wst, wuse & 1)) { err = 17; break; }
                                if (prof) c_w += clock64() - t0;
                                tc_fence_after();
                                const long long ti = prof ? clock64() : 0;
                                const uint32_t b_lo = wst0 + wst * (cp::WST_BYTES >> 4);
                                const uint32_t a_tap = ((kh * cp::BW + kw) * cp::ROWB) >> 4;
                                uint32_t slot = pslot, use = puse, ss = s;
                                for (int i = 0; i < gn; ++i) {
                                    if (t == 0) {
                                        if (prof) t0 = clock64();
                                        if (!mbar_wait_warp<false>(plane_full + 8 * slot, use & 1)) { err = 18; break; }
                                        if (prof) c_plane += clock64() - t0;
                                        tc_fence_after();
                                    }
                                    const uint32_t al = plane0 + slot * (cp::PLANE_BYTES >> 4) + a_tap;
                                    const uint32_t dcol = tm + ss * cp::NT;
                                    mma_f16_ss_pair_sel(dcol, al, hi_a, b_lo, hi_b, idesc, sel);
                                    mma_f16_ss_pair_sel(dcol, al + 2, hi_a, b_lo + 2, hi_b, idesc, sel);
                                    if (t == 24) mma_commit_pair_sel(plane_empty + 8 * slot, 3u, sel);
                                    if (++slot == RING) { slot = 0; ++use; }
                                    if (++ss == cp::PERIOD) ss = 0;
                                }
                                if (err) break;
                                mma_commit_pair_sel(w_empty + 8 * wst, 3u, sel);
                                if (prof) c_issue += clock64() - ti;
                                if (++wst == WST) { wst = 0; ++wuse; }
                            }
                            pslot += gn;
                            if (pslot >= RING) { pslot -= RING; ++puse; }
                        }
                    }
                    if (err == 0) mma_commit_pair_sel(tmem_full + 8 * (g & 1), 3u, sel);
                    s += gn;
                    if (s >= cp::PERIOD) s -= cp::PERIOD;
                }
            }
            if (err != 0) atomicExch(P.error_flag, err);
            if (prof && lane == 0) {
                long long* o = P.prof + 8 * (size_t)(cluster % 160);
                o[0] = clock64() - c_total; o[1] = c_tmem; o[2] = c_w; o[3] = c_plane;
                o[5] = (long long)globaltimer_ns();
                o[7] = c_issue;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own 128 accumulator rows) =====================
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        const int th = row >> 3, tw = row & 7;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        float scale = P.out_scale;
        if (P.out_scale_dev) scale *= *P.out_scale_dev;
        uint32_t zeros[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) zeros[j] = 0u;
        for (int c = 0; c < 512; c += 32) tmem_st_32x32(tmem + c + lane_addr, zeros);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive_remote(tmem_empty, 0); mbar_arrive_remote(tmem_empty + 8, 0); }

        RunWalker rw(P, cluster);
        int n, h0, w0, da, db;
        uint32_t g = 0;
        bool ok = true;
        while (ok && rw.next(n, h0, w0, da, db)) {
            const int p0 = max(da - 2, P.p_lo), p1 = min(db + 1, P.p_hi);
            const int hh = h0 + (int)rank * cp::TH + th;
            const bool row_ok = hh < P.H;
            for (int gp = p0; gp <= p1; gp += cp::GROUP, ++g) {
                const int gl = min(gp + cp::GROUP - 1, p1);          // last plane of the group
                if (!mbar_wait(tmem_full + 8 * (g & 1), (g >> 1) & 1)) { atomicExch(P.error_flag, 19); ok = false; break; }
                tc_fence_after();
                // output planes completed by this group: q = gp-2 .. gl-2, plus the tail gl-1 .. gl+2 after the last group
                const int qa = gp - 2, qb = (gl == p1) ? gl + 2 : gl - 2;
                for (int q = qa; q <= qb; ++q) {
                    const int slot = (q - p0 + 2) % cp::PERIOD;
                    const uint32_t taddr = tmem + slot * cp::NT + lane_addr;
                    const bool spill = slot < 4;                       // windows starting in slots 8..11 spill into 12..15
                    if (q >= da && q < db) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr, v);
                        tmem_ld_wait();
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        if (spill) {
                            tmem_ld_32x32(taddr + cp::PERIOD * cp::NT, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] *= scale;
                        if (P.accum && row_ok) {                      // later K chunk: add to the earlier chunks' partial result
                            const float* old = P.y + ((((size_t)n * P.D + q) * P.H + hh) * P.W + w0 + tw) * P.Nout + n0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 o4 = *reinterpret_cast<const float4*>(old + j);
                                f[j] += o4.x; f[j + 1] += o4.y; f[j + 2] += o4.z; f[j + 3] += o4.w;
                            }
                        }
                        if (has_ep) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], s_ep[j], s_ep[cp::NT + j]);
                        }
                        if (P.ext.relu) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = row_ok ? f[j] : 0.f;
                        if (row_ok) {
                            const size_t vox = ((size_t)hh * P.W + w0 + tw);
                            if (P.y != nullptr) {
                                float* dst = P.y + ((((size_t)n * P.D + q) * P.H) * P.W + vox) * P.Nout + n0;
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                            }
                            if (P.ext.y16 != nullptr) {
                                __half* d16 = P.ext.y16 +
                                              ((((size_t)n * P.ext.Dy16 + q + P.ext.y16_off) * P.H) * P.W + vox) * P.Nout + n0;
                                const float s16 = P.ext.y16_scale;
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    __half2 h0 = sat_half2(f[j] * s16, f[j + 1] * s16), h1 = sat_half2(f[j + 2] * s16, f[j + 3] * s16);
                                    __half2 h2 = sat_half2(f[j + 4] * s16, f[j + 5] * s16), h3 = sat_half2(f[j + 6] * s16, f[j + 7] * s16);
                                    uint4 pk;
                                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                                    *reinterpret_cast<uint4*>(d16 + j) = pk;
                                }
                            }
                        }
                        if (P.bn_sums != nullptr && q >= P.stat_lo && q < P.stat_hi) {
                            float gq[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) gq[j] = f[j] * f[j];
                            // transposed butterfly: afterwards lane l holds the 32-voxel total of channel l
#pragma unroll
                            for (int sft = 16; sft >= 1; sft >>= 1) {
                                const bool up = (lane & sft) != 0;
#pragma unroll
                                for (int i = 0; i < sft; ++i) {
                                    const float send1 = up ? f[i] : f[i + sft], send2 = up ? gq[i] : gq[i + sft];
                                    const float r1 = __shfl_xor_sync(0xffffffffu, send1, sft);
                                    const float r2 = __shfl_xor_sync(0xffffffffu, send2, sft);
                                    f[i] = (up ? f[i + sft] : f[i]) + r1;
                                    gq[i] = (up ? gq[i + sft] : gq[i]) + r2;
                                }
                            }
                            atomicAdd(s_bn + lane, (double)f[0]);
                            atomicAdd(s_bn + cp::NT + lane, (double)gq[0]);
                        }
                    }
                    tmem_st_32x32(taddr, zeros);
                    if (spill) tmem_st_32x32(taddr + cp::PERIOD * cp::NT, zeros);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(tmem_empty + 8 * (g & 1), 0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();              // the peer's shared memory and barriers stay alive until both CTAs are done
    if (P.prof != nullptr && threadIdx.x == 0 && rank == 0) {
        long long* o = P.prof + 8 * (size_t)(cluster % 160);
        o[4] = (long long)ts_entry;
        o[6] = (long long)globaltimer_ns();
    }
    if (P.bn_sums != nullptr)
        for (int i = threadIdx.x; i < 2 * cp::NT; i += cp::THREADS) {
            const int which = i / cp::NT, ch = i % cp::NT;
            atomicAdd(P.bn_sums + which * P.Nout + n0 + ch, s_bn[i]);
        }
    // D-sharded slabs: the last CTA of the launch stores the slab's complete sums into every rank's slot and signals
    if (P.ext.push.n > 0) push_vector_from_last_block(P.bn_sums, 2 * P.Nout, P.ext.push);
    if (warp == 2) tmem_dealloc_pair<512>(tmem);
}

// ------------------------------------------------------------------------------------------------ host
int* device_error_flag();            // mode_abi.cu
long long* debug_profile_buffer();   // mode_abi.cu
int make_act_map(CUtensorMap* map, const __half* x, int N, int D, int H, int W, int K, int box_w, int box_h, int box_d);   // conv_umma.cu

// input planes a run [da, db) of a D-plane column has to process (+0.25 per group hand-off)
static double run_cost(int da, int db, int p_lo, int p_hi) {
    const int planes = std::min(db + 1, p_hi) - std::max(da - 2, p_lo) + 1;
    return planes + 0.25 * ((planes + cp::GROUP - 1) / cp::GROUP);
}

// take plane-units from u while the accumulated cost stays within the budget; returns the new position
static int64_t take_runs(int64_t u, int64_t units, int D, int p_lo, int p_hi, double budget) {
    double acc = 0;
    while (u < units) {
        const int da = (int)(u % D);
        const double whole = run_cost(da, D, p_lo, p_hi);
        if (acc + whole <= budget) { acc += whole; u += D - da; continue; }
        int lo = 0, hi = D - da;                           // largest k with cost(da, da+k) fitting (cost is monotone in k)
        while (hi - lo > 1) {
            const int mid = (lo + hi) / 2;
            if (acc + run_cost(da, da + mid, p_lo, p_hi) <= budget) lo = mid; else hi = mid;
        }
        u += lo;
        break;
    }
    return u;
}

static void partition_runs(int64_t units, int D, int p_lo, int p_hi, int G, int32_t* bounds) {
    struct Key {
        int64_t units; int D, p_lo, p_hi, G;
        bool operator<(const Key& o) const {
            return std::tie(units, D, p_lo, p_hi, G) < std::tie(o.units, o.D, o.p_lo, o.p_hi, o.G);
        }
    };
    static std::mutex mu;
    static std::map<Key, std::array<int32_t, cp::MAX_CLUSTERS + 1>> cache;
    const Key key{units, D, p_lo, p_hi, G};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
        double total = 0;
        for (int64_t u = 0; u < units; u += D) total += run_cost(0, D, p_lo, p_hi);
        auto feasible = [&](double budget) {
            int64_t u = 0;
            for (int c = 0; c < G && u < units; ++c) {
                const int64_t nu = take_runs(u, units, D, p_lo, p_hi, budget);
                if (nu == u) return false;
                u = nu;
            }
            return u >= units;
        };
        double h = total / G;
        while (!feasible(h)) h *= 1.05;
        double l = total / G;
        for (int i = 0; i < 24; ++i) {
            const double m = 0.5 * (l + h);
            if (feasible(m)) h = m; else l = m;
        }
        std::array<int32_t, cp::MAX_CLUSTERS + 1> b{};
        int64_t u = 0;
        for (int c = 0; c < G; ++c) {
            b[c] = (int32_t)u;
            u = std::min(units, take_runs(u, units, D, p_lo, p_hi, h));
        }
        for (int c = G; c <= cp::MAX_CLUSTERS; ++c) b[c] = (int32_t)units;
        it = cache.emplace(key, b).first;
    }
    std::copy(it->second.begin(), it->second.end(), bounds);
}

static int pair_clusters(int passes) { return std::max(1, std::min(cp::MAX_CLUSTERS, (sm_count() / 2) / passes)); }

// Auto-selected when every cluster gets a long enough march (otherwise the single-CTA kernel's finer tiles win).  The
// weights of ONE 32-channel chunk fit the RESIDENT plan; K > 32 runs as one resident launch per chunk, each later chunk
// adding to y in the epilogue (the streaming plan re-reads every weight stage per 4 planes from L2 and measured slower
// than the single-CTA kernel: 0.135 vs 0.086 ms on 64->64 @ 16x64x64).  REPMODE_PAIR_MAXK caps the chunked form (default 128).
bool conv3d_pair_supported(int N, int D, int H, int W, int K, int Nout) {
    static const int maxk = getenv("REPMODE_PAIR_MAXK") ? atoi(getenv("REPMODE_PAIR_MAXK")) : 128;
    if (!(K % 32 == 0 && K >= 32 && K <= std::max(32, maxk) && Nout % 32 == 0 && Nout >= 32 && W % cp::TW == 0)) return false;
    const int64_t units = (int64_t)N * ceil_div(H, 2 * cp::TH) * (W / cp::TW) * D;
    if (units > 0x7fffffff) return false;
    return units >= (int64_t)12 * pair_clusters(Nout / cp::NT);
}

int conv3d_pair(const __half* x, const __half* w, const int32_t* sample_u, float* y, int N, int D, int H, int W, int K,
                int Nout, float out_scale, const float* out_scale_dev, double* bn_sums, int stat_lo, int stat_hi,
                const ConvExt& ext, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) ||
        (reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(ext.y16) & 15))
        MODE_FAIL("conv3d_pair: pointers must be 16-byte aligned");
    if (!(K % 32 == 0 && Nout % 32 == 0 && W % cp::TW == 0)) MODE_FAIL("conv3d_pair: unsupported shape");
    PairParams P;
    P.w = w; P.sample_u = sample_u; P.y = y; P.bn_sums = bn_sums; P.out_scale_dev = out_scale_dev;
    P.out_scale = out_scale;
    P.N = N; P.D = D; P.H = H; P.W = W; P.K = K; P.Nout = Nout;
    P.stat_lo = stat_lo; P.stat_hi = stat_hi;
    P.tiles_w = W / cp::TW; P.tiles_h2 = (int)ceil_div(H, 2 * cp::TH);
    const int64_t units = (int64_t)N * P.tiles_h2 * P.tiles_w * D;
    if (units > 0x7fffffff) MODE_FAIL("conv3d_pair: volume too large for 32-bit unit indices");
    const int passes = Nout / cp::NT;
    int G = (int)std::min<int64_t>(pair_clusters(passes), std::max<int64_t>(1, units / 4));
    if (const char* e = getenv("REPMODE_PAIR_CLUSTERS"))          // test hook: force the number of clusters
        G = std::max(1, std::min(std::min(G, cp::MAX_CLUSTERS), atoi(e)));
    P.ext = ext;
    P.p_lo = ext.p_lo(); P.p_hi = ext.p_hi();
    partition_runs(units, D, P.p_lo, P.p_hi, G, P.bounds);
    P.error_flag = device_error_flag();
    if (!P.error_flag) MODE_FAIL("conv3d_pair: could not allocate the device error flag");
    P.prof = debug_profile_buffer();
    P.flags = 0;
    if (const char* e = getenv("REPMODE_PAIR_FLAGS")) P.flags = atoi(e);
    CUtensorMap xmap;
    if (make_act_map(&xmap, x, N, ext.Dx, H, W, K, cp::BW, cp::BH, 1) != 0) return -1;
    static_assert(cp::Plan<true>::TOTAL + 1024 <= 227 * 1024 && cp::Plan<false>::TOTAL + 1024 <= 227 * 1024,
                  "shared memory budget");
    P.x_chan0 = 0; P.accum = 0;
    P.w_u_stride = (long long)(K / 32) * 125 * Nout * 32;
    if (!(P.flags & 2)) {
        // RESIDENT plan, one launch per 32-channel chunk of K; statistics, affine epilogue, fp16 copy and the stats
        // broadcast belong to the LAST chunk's launch (the only one that sees the complete sum)
        const int nchunks = K / 32;
        if (nchunks > 1 && y == nullptr) MODE_FAIL("conv3d_pair: K > 32 needs the fp32 output buffer (chunk accumulation)");
        const int smem_bytes = (int)cp::Plan<true>::TOTAL + 1024;
        MODE_CUDA(cudaFuncSetAttribute(conv3d_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        PairParams Q = P;
        Q.K = 32;
        for (int c = 0; c < nchunks; ++c) {
            const bool last = c == nchunks - 1;
            Q.x_chan0 = c * 32;
            Q.w = w + (size_t)c * 125 * Nout * 32;
            Q.accum = c > 0;
            Q.bn_sums = last ? bn_sums : nullptr;
            Q.ext = ext;
            if (!last) {
                Q.ext.ep_scale = nullptr; Q.ext.ep_shift = nullptr; Q.ext.relu = 0; Q.ext.y16 = nullptr; Q.ext.push.n = 0;
                Q.out_scale = out_scale; Q.out_scale_dev = out_scale_dev;
            }
            conv3d_pair_kernel<true><<<dim3(2 * G, passes), cp::THREADS, smem_bytes, st>>>(xmap, Q);
            MODE_LAUNCH_CHECK();
        }
        return 0;
    } else {
        const int smem_bytes = (int)cp::Plan<false>::TOTAL + 1024;
        MODE_CUDA(cudaFuncSetAttribute(conv3d_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        conv3d_pair_kernel<false><<<dim3(2 * G, passes), cp::THREADS, smem_bytes, st>>>(xmap, P);
    }
    MODE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mode
