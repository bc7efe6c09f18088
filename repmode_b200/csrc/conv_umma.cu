// K2 / K3: 5x5x5 'same' convolution as an im2col-free implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces F.conv3d(x[i:i+1], w[i], padding='same') per sample (fnet/nn_modules/RepMode.py:204-210 in the
// reference tree) and -- called with K1's flipped/transposed weight pack and dy as input -- its dgrad.
//
// Mapping (one persistent CTA per SM, 256 threads, warp-specialised):
//   GEMM M = 128 output voxels = an 8(w) x 16(h) patch of one d-plane.  A CTA tile is TD consecutive
//   d-planes of that patch; its TD accumulators [128 x Nt] fp32 sit side by side in TMEM (TD*Nt <= 256
//   columns, double buffered in the 512-column TMEM so the epilogue of tile i overlaps tile i+1).
//   A operand: the haloed activation plane (20 x 12 voxels x 32 ch fp16 = 15 KB) is brought in ONCE per
//   (tile, 32-channel chunk) by a single 5-d TMA box load (out-of-range coordinates zero-fill = the conv's
//   zero padding) and stays resident for all 25 (kh,kw) taps: a tap is just a different descriptor start
//   address (rows of 64 B, SWIZZLE_64B, SBO = 12*64 B) -- no im2col copy.
//   B operand: for one (chunk, kh, kw) K1 stores the five kd taps as consecutive [Nout x 32] blocks in
//   DESCENDING kd order, already in the 64B-swizzled shared-memory image.  Input plane p feeds output planes
//   q = p-4 .. p with taps kd = p-q, i.e. a CONTIGUOUS run of B rows and a contiguous run of accumulator
//   columns: ONE tcgen05.mma with N = Nt * (#valid q) <= 256 updates up to five output planes at once.
//   (Measured on B200: an SS-mode MMA costs max(93, 42 + N/2) cycles -- the 4 KB A read is exposed -- so the
//   wide N is what amortises it; every column is useful work.)
// Roles: warp 0 = activation-plane TMA producer, warp 3 = weight producer (bulk copies), warp 1 = MMA issuer
//   (one thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> scale -> global store, fused
//   BatchNorm partial sums, then re-zero the accumulators with tcgen05.st so every MMA can accumulate).
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

namespace cu {
constexpr int TW = 8, TH = 16;                 // output patch
constexpr int BW = TW + 4, BH = TH + 4;        // haloed brick
constexpr int ROWB = 64;                       // bytes per voxel row (32 fp16 channels)
constexpr int PLANE_BYTES = BH * BW * ROWB;    // 15360
constexpr int MAX_RING = 12;
constexpr int MAX_WST = 4;
constexpr int THREADS = 256;
}  // namespace cu

struct ConvParams {
    const __half* w;             // stage-major fp16 pack (see header), all Nout rows
    const int32_t* sample_u;
    float* y;                    // [N,D,H,W,Nout]
    double* bn_sums;             // [2*Nout] or null
    const float* out_scale_dev;
    float out_scale;
    int N, D, H, W, K, Nout;     // Nout = total output channels (row stride of y / rows per weight block)
    int Nt;                      // CTA (x, y) computes output channels [y*Nt, (y+1)*Nt)
    int stat_lo, stat_hi;        // BatchNorm sums cover output planes [stat_lo, stat_hi) only (owned planes of a slab)
    int TD, ring, wstages;
    int tiles_w, tiles_h;
    ConvExt ext;                 // haloed input (Dx, x_off) and fused epilogue (affine + ReLU, fp16 copy)
    int p_lo, p_hi;              // valid input planes in output-plane coordinates: [-x_off, Dx - x_off - 1]
    int64_t units;               // N * tiles_h * tiles_w * D plane-patches
    int32_t bounds[160];         // CTA c owns units [bounds[c], bounds[c+1]) -- cost-balanced on the host
    // split-K (deep, small-volume layers): CTA (x, y, z) accumulates the 32-channel chunks [z*cps, (z+1)*cps) only and
    // stores its RAW fp32 accumulators into part[z] (same NDHWC indexing as y); conv_splitk_reduce_kernel sums the splits
    // in fixed order and applies the whole epilogue (scale, affine, ReLU, fp16 copy, BatchNorm sums)
    float* part;                 // null = no split
    long long part_stride;       // elements per split = N*D*H*W*Nout
    int cps;                     // chunks per split (nchunk when part == null)
    int* error_flag;
    long long* prof;             // optional [grid][8]: MMA-warp cycles total / wait tmem_empty / wait weights / wait
                                 // planes, then globaltimer ns at CTA entry / MMA loop end / all roles done
};

// Work distribution: a "unit" is one 8x16 patch of one output d-plane, ordered (n, th, tw, d).  CTA c owns the
// contiguous unit range [bounds[c], bounds[c+1]) and walks it in tiles of up to TD consecutive planes of one patch
// column.  The host chooses the bounds so that every CTA gets the same ESTIMATED CYCLES (short tiles cost more per
// plane: fewer output planes share each input plane), not the same number of planes.
struct TileWalker {
    int64_t u, uend;
    int D, TD, tiles_h, tiles_w;
    __device__ TileWalker(const ConvParams& P)
        : u(P.bounds[blockIdx.x]), uend(P.bounds[blockIdx.x + 1]), D(P.D), TD(P.TD),
          tiles_h(P.tiles_h), tiles_w(P.tiles_w) {}
    __device__ bool next(int& n, int& h0, int& w0, int& d0, int& td) {
        if (u >= uend) return false;
        const int64_t patch = u / D;
        d0 = (int)(u - patch * D);
        td = (int)min((int64_t)min(TD, D - d0), uend - u);
        w0 = (int)(patch % tiles_w) * cu::TW;
        h0 = (int)((patch / tiles_w) % tiles_h) * cu::TH;
        n = (int)(patch / ((int64_t)tiles_w * tiles_h));
        u += td;
        return true;
    }
};

struct SmemLayout {
    uint32_t plane_off, w_off, bar_off, bn_off, ep_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(int ring, int wstages, int Nt) {
    SmemLayout L;
    L.plane_off = 0;
    L.w_off = ring * cu::PLANE_BYTES;                         // multiples of 15 KB keep 512-byte alignment
    const uint32_t wst_bytes = 5u * Nt * cu::ROWB;
    L.bar_off = L.w_off + wstages * wst_bytes;
    L.bn_off = L.bar_off + 1024;       // 512 B of mbarriers + tmem slot, 512 B MMA-issuer plane table
    L.ep_off = L.bn_off + 2 * 256 * sizeof(double);     // epilogue affine: scale[Nt], shift[Nt]
    L.total = L.ep_off + 2 * 256 * sizeof(float);
    return L;
}

// valid (in-volume) input planes of a tile: p in [pmin, pmax], input plane (output coordinates) = d0 + p - 2 in [p_lo, p_hi]
__device__ __forceinline__ void plane_range(int d0, int TD, int p_lo, int p_hi, int& pmin, int& pmax) {
    pmin = max(0, p_lo + 2 - d0);
    pmax = min(TD + 3, p_hi + 2 - d0);
}

__global__ void __launch_bounds__(cu::THREADS, 1)
conv3d_umma_kernel(const __grid_constant__ CUtensorMap xmap, const ConvParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const SmemLayout L = smem_layout(P.ring, P.wstages, P.Nt);
    const uint32_t bars = base + L.bar_off;
    // barrier table (8 bytes each)
    const uint32_t plane_full = bars, plane_empty = bars + 8 * cu::MAX_RING;
    const uint32_t w_full = bars + 16 * cu::MAX_RING, w_empty = w_full + 8 * cu::MAX_WST;
    const uint32_t tmem_full = w_empty + 8 * cu::MAX_WST, tmem_empty = tmem_full + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 400);
    double* s_bn = reinterpret_cast<double*>(smem + L.bn_off);   // per-CTA BatchNorm partial sums [2][Nt]
    float* s_ep = reinterpret_cast<float*>(smem + L.ep_off);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * P.Nt;
    const uint64_t ts_entry = (P.prof != nullptr && threadIdx.x == 0) ? globaltimer_ns() : 0;
    const int nchunk = P.K / 32;
    const int c_lo = blockIdx.z * P.cps, c_hi = min(nchunk, c_lo + P.cps);      // this CTA's K slice
    const uint32_t blk_bytes = (uint32_t)P.Nt * cu::ROWB;          // one kd block of a weight stage
    const uint32_t wst_bytes = 5u * blk_bytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < P.ring; ++i) { mbar_init(plane_full + 8 * i, 1); mbar_init(plane_empty + 8 * i, 1); }
        for (int i = 0; i < P.wstages; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, 4); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
    if (warp == 0 && lane == 0) tma_prefetch_desc(&xmap);
    for (int i = threadIdx.x; i < 2 * P.Nt; i += cu::THREADS) s_bn[i] = 0.0;
    const bool has_ep = P.ext.ep_scale != nullptr || P.ext.ep_shift != nullptr;
    if (has_ep)
        for (int i = threadIdx.x; i < P.Nt; i += cu::THREADS) {
            s_ep[i] = P.ext.ep_scale ? P.ext.ep_scale[n0 + i] : 1.f;
            s_ep[P.Nt + i] = P.ext.ep_shift ? P.ext.ep_shift[n0 + i] : 0.f;
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== activation-plane producer =====================
        if (lane == 0) {
            uint32_t seq = 0;
            TileWalker tw_(P);
            int n, h0, w0, d0, td;
            while (tw_.next(n, h0, w0, d0, td)) {
                int pmin, pmax;
                plane_range(d0, td, P.p_lo, P.p_hi, pmin, pmax);
                for (int c = c_lo; c < c_hi; ++c) {
                    for (int p = pmin; p <= pmax; ++p, ++seq) {
                        const uint32_t slot = seq % P.ring, use = seq / P.ring;
                        if (!mbar_wait(plane_empty + 8 * slot, (use & 1) ^ 1)) { atomicExch(P.error_flag, 1); return; }
                        mbar_expect_tx(plane_full + 8 * slot, cu::PLANE_BYTES);
                        tma_load_5d(base + L.plane_off + slot * cu::PLANE_BYTES, &xmap, plane_full + 8 * slot, c * 32,
                                    w0 - 2, h0 - 2, d0 + p - 2 + P.ext.x_off, n);
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== weight producer =====================
        if (lane == 0) {
            uint32_t seq = 0;
            TileWalker tw_(P);
            int n, h0, w0, d0, td;
            while (tw_.next(n, h0, w0, d0, td)) {
                const int u = P.sample_u ? P.sample_u[n] : 0;
                const __half* wu = P.w + (size_t)u * nchunk * 125 * P.Nout * 32;
                for (int c = c_lo; c < c_hi; ++c) {
                    for (int t = 0; t < 25; ++t, ++seq) {
                        const uint32_t st = seq % P.wstages, use = seq / P.wstages;
                        if (!mbar_wait(w_empty + 8 * st, (use & 1) ^ 1)) { atomicExch(P.error_flag, 2); return; }
                        mbar_expect_tx(w_full + 8 * st, wst_bytes);
                        const uint32_t dst = base + L.w_off + st * wst_bytes;
                        const __half* src = wu + ((size_t)(c * 25 + t) * 5 * P.Nout + n0) * 32;
#pragma unroll
                        for (int b = 0; b < 5; ++b)
                            bulk_load(dst + b * blk_bytes, src + (size_t)b * P.Nout * 32, blk_bytes, w_full + 8 * st);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // One thread issues every tcgen05.mma of the CTA.  Everything that depends only on (tile, chunk, plane) --
        // ring slot address, accumulator column, instruction descriptor (N varies with the number of output planes
        // the input plane feeds), weight block offset -- is tabulated once per chunk in shared memory; the inner
        // loop is one 16-byte table load, two 32-bit adds and two MMAs per plane.  No divisions: ring/stage indices
        // are wrapped counters.  (Measured: a warp-uniform variant with the election inside the asm is 15 % slower.)
        if (lane == 0) {
            uint4* tab = reinterpret_cast<uint4*>(smem + L.bar_off + 512);   // [MAX_RING] entries
            const uint32_t hi_a = ((cu::BW * cu::ROWB) >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);
            const uint32_t hi_b = (512u >> 4) | (1u << 14) | ((uint32_t)SWZ_64B << 29);
            const uint32_t lbo_lo = 1u << 16;                                    // LBO field = 16 B (unused for K-major)
            uint32_t pslot = 0, puse = 0;        // ring slot / use count of the NEXT plane in sequence
            uint32_t wst = 0, wuse = 0;          // weight stage / use count
            long long c_tmem = 0, c_w = 0, c_plane = 0, c_total = clock64();     // wait-cycle profile (mode_debug_profile)
            TileWalker tw_(P);
            int n, h0, w0, d0, td;
            for (int it = 0; tw_.next(n, h0, w0, d0, td); ++it) {
                int pmin, pmax;
                plane_range(d0, td, P.p_lo, P.p_hi, pmin, pmax);
                const int nplanes = pmax - pmin + 1;
                const int buf = it & 1;
                const uint32_t acc_base = tmem + buf * 256;
                // accumulators of this buffer have been drained and re-zeroed by the epilogue
                long long t0 = clock64();
                if (!mbar_wait(tmem_empty + 8 * buf, (it >> 1) & 1)) { atomicExch(P.error_flag, 3); return; }
                c_tmem += clock64() - t0;
                tc_fence_after();
                for (int c = c_lo; c < c_hi; ++c) {
                    // ---- per-chunk plane table
                    uint32_t slot = pslot, use = puse;
                    for (int i = 0; i < nplanes; ++i) {
                        const int p = pmin + i;
                        const int qlo = max(0, p - 4), qhi = min(td - 1, p);
                        uint4 e;
                        e.x = ((base + L.plane_off + slot * cu::PLANE_BYTES) >> 4) | lbo_lo;          // A desc lo (tap 0)
                        e.y = ((uint32_t)(4 - (p - qlo)) * blk_bytes) >> 4;                            // B block offset
                        e.z = acc_base + qlo * P.Nt;                                                   // D column
                        e.w = make_idesc(FMT_F16, 128, (uint32_t)(qhi - qlo + 1) * P.Nt, 0, 0);
                        tab[i] = e;
                        if (++slot == (uint32_t)P.ring) { slot = 0; ++use; }
                    }
                    for (int t = 0; t < 25; ++t) {
                        const int kh = t / 5, kw = t - kh * 5;
                        t0 = clock64();
                        if (!mbar_wait(w_full + 8 * wst, wuse & 1)) { atomicExch(P.error_flag, 5); return; }
                        c_w += clock64() - t0;
                        tc_fence_after();
                        const uint32_t wb_lo = ((base + L.w_off + wst * wst_bytes) >> 4) | lbo_lo;
                        const uint32_t a_tap = ((kh * cu::BW + kw) * cu::ROWB) >> 4;
                        if (t > 0 && t < 24) {
#pragma unroll 4
                            for (int i = 0; i < nplanes; ++i) {
                                const uint4 e = tab[i];
                                const uint32_t a_lo = e.x + a_tap, b_lo = wb_lo + e.y;
                                mma_f16_ss(e.z, ((uint64_t)hi_a << 32) | a_lo, ((uint64_t)hi_b << 32) | b_lo, e.w, 1u);
                                mma_f16_ss(e.z, ((uint64_t)hi_a << 32) | (a_lo + 2), ((uint64_t)hi_b << 32) | (b_lo + 2), e.w, 1u);
                            }
                        } else {
                            // t == 0: first touch of each plane -> wait for its TMA (planes land in sequence order);
                            // t == 24: last touch -> release the ring slot as soon as its MMAs retire
                            slot = pslot; use = puse;
                            for (int i = 0; i < nplanes; ++i) {
                                if (t == 0) {
                                    t0 = clock64();
                                    if (!mbar_wait(plane_full + 8 * slot, use & 1)) { atomicExch(P.error_flag, 4); return; }
                                    c_plane += clock64() - t0;
                                    tc_fence_after();
                                }
                                const uint4 e = tab[i];
                                const uint32_t a_lo = e.x + a_tap, b_lo = wb_lo + e.y;
                                mma_f16_ss(e.z, ((uint64_t)hi_a << 32) | a_lo, ((uint64_t)hi_b << 32) | b_lo, e.w, 1u);
                                mma_f16_ss(e.z, ((uint64_t)hi_a << 32) | (a_lo + 2), ((uint64_t)hi_b << 32) | (b_lo + 2), e.w, 1u);
                                if (t == 24) mma_commit(plane_empty + 8 * slot);
                                if (++slot == (uint32_t)P.ring) { slot = 0; ++use; }
                            }
                        }
                        mma_commit(w_empty + 8 * wst);
                        if (++wst == (uint32_t)P.wstages) { wst = 0; ++wuse; }
                    }
                    pslot += nplanes;
                    if (pslot >= (uint32_t)P.ring) { pslot -= P.ring; ++puse; }
                }
                mma_commit(tmem_full + 8 * buf);
            }
            if (P.prof != nullptr) {
                long long* o = P.prof + 8 * (size_t)(blockIdx.x % 160);
                o[0] = clock64() - c_total; o[1] = c_tmem; o[2] = c_w; o[3] = c_plane;
                o[5] = (long long)globaltimer_ns();                       // MMA issue loop done
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        const int th = row >> 3, tw = row & 7;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        float scale = P.out_scale;
        if (P.out_scale_dev) scale *= *P.out_scale_dev;
        uint32_t zeros[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) zeros[j] = 0u;
        // zero both accumulator buffers once; afterwards each buffer is re-zeroed right after it is drained
        for (int c = 0; c < 512; c += 32) tmem_st_32x32(tmem + c + lane_addr, zeros);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(tmem_empty); mbar_arrive(tmem_empty + 8); }

        TileWalker tw_(P);
        int n, h0, w0, d0, td;
        for (int it = 0; tw_.next(n, h0, w0, d0, td); ++it) {
            const int buf = it & 1;
            if (!mbar_wait(tmem_full + 8 * buf, (it >> 1) & 1)) { atomicExch(P.error_flag, 6); break; }
            tc_fence_after();
            const int qn = td;
            const bool row_ok = h0 + th < P.H;          // H need not be a multiple of 16: rows past the volume are
            for (int q = 0; q < td; ++q) {              // computed (TMA zero fill) but neither stored nor counted
                const size_t vox = (size_t)(h0 + th) * P.W + w0 + tw;
                for (int cc = 0; cc < P.Nt; cc += 32) {
                    const uint32_t taddr = tmem + buf * 256 + q * P.Nt + cc + lane_addr;
                    if (q < qn) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr, v);
                        tmem_ld_wait();
                        if (P.part != nullptr) {              // split-K: raw partial sums, epilogue in the reduce kernel
                            if (row_ok) {
                                float* dst = P.part + (size_t)blockIdx.z * P.part_stride +
                                             ((((size_t)n * P.D + d0 + q) * P.H) * P.W + vox) * P.Nout + n0 + cc;
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<uint4*>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            }
                            tmem_st_32x32(taddr, zeros);
                            continue;
                        }
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * scale;
                        if (has_ep) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], s_ep[cc + j], s_ep[P.Nt + cc + j]);
                        }
                        if (P.ext.relu) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = row_ok ? f[j] : 0.f;
                        if (row_ok) {
                            if (P.y != nullptr) {
                                float* dst = P.y + ((((size_t)n * P.D + d0 + q) * P.H) * P.W + vox) * P.Nout + n0 + cc;
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                            }
                            if (P.ext.y16 != nullptr) {
                                __half* d16 = P.ext.y16 + ((((size_t)n * P.ext.Dy16 + d0 + q + P.ext.y16_off) * P.H) * P.W + vox) *
                                                              P.Nout + n0 + cc;
                                const float s16 = P.ext.y16_scale;
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    __half2 h0_ = sat_half2(f[j] * s16, f[j + 1] * s16), h1_ = sat_half2(f[j + 2] * s16, f[j + 3] * s16);
                                    __half2 h2_ = sat_half2(f[j + 4] * s16, f[j + 5] * s16), h3_ = sat_half2(f[j + 6] * s16, f[j + 7] * s16);
                                    uint4 pk;
                                    pk.x = *reinterpret_cast<uint32_t*>(&h0_); pk.y = *reinterpret_cast<uint32_t*>(&h1_);
                                    pk.z = *reinterpret_cast<uint32_t*>(&h2_); pk.w = *reinterpret_cast<uint32_t*>(&h3_);
                                    *reinterpret_cast<uint4*>(d16 + j) = pk;
                                }
                            }
                        }
                        if (P.bn_sums != nullptr && d0 + q >= P.stat_lo && d0 + q < P.stat_hi) {
                            float g[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) g[j] = f[j] * f[j];
                            // transposed butterfly: afterwards lane l holds the 32-voxel total of channel cc + l
#pragma unroll
                            for (int s = 16; s >= 1; s >>= 1) {
                                const bool up = (lane & s) != 0;
#pragma unroll
                                for (int i = 0; i < s; ++i) {
                                    const float send1 = up ? f[i] : f[i + s], send2 = up ? g[i] : g[i + s];
                                    const float r1 = __shfl_xor_sync(0xffffffffu, send1, s);
                                    const float r2 = __shfl_xor_sync(0xffffffffu, send2, s);
                                    f[i] = (up ? f[i + s] : f[i]) + r1;
                                    g[i] = (up ? g[i + s] : g[i]) + r2;
                                }
                            }
                            atomicAdd(s_bn + cc + lane, (double)f[0]);
                            atomicAdd(s_bn + P.Nt + cc + lane, (double)g[0]);
                        }
                    }
                    tmem_st_32x32(taddr, zeros);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (P.prof != nullptr && threadIdx.x == 0) {
        long long* o = P.prof + 8 * (size_t)(blockIdx.x % 160);
        o[4] = (long long)ts_entry;                                       // CTA entry
        o[6] = (long long)globaltimer_ns();                               // all roles done (before teardown)
    }
    if (P.bn_sums != nullptr && P.part == nullptr)
        for (int i = threadIdx.x; i < 2 * P.Nt; i += cu::THREADS) {
            const int which = i / P.Nt, ch = i % P.Nt;
            atomicAdd(P.bn_sums + which * P.Nout + n0 + ch, s_bn[i]);
        }
    if (P.ext.push.n > 0) push_vector_from_last_block(P.bn_sums, 2 * P.Nout, P.ext.push);
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// ---- split-K reduce + epilogue ----------------------------------------------------------------------------------
// Sums the ksplit partial results of conv3d_umma_kernel in fixed order (deterministic) and applies the epilogue the
// unsplit kernel applies in place: out_scale, per-channel affine + ReLU (eval BatchNorm), fp32 result and/or saturated fp16
// copy, BatchNorm sums over the planes [stat_lo, stat_hi).  HBM/L2-bound: (ksplit + 1) * 4 B per output element, float4
// accesses, one thread = 4 channels of one voxel.
struct SplitReduceParams {
    const float* part; int ksplit; long long part_stride;
    long long M; int Nout, D; long long HW;
    float out_scale; const float* out_scale_dev;
    const float* ep_scale; const float* ep_shift; int relu;
    float* y; __half* y16; int Dy16, y16_off; float y16_scale;
    double* bn_sums; int stat_lo, stat_hi;
};

__global__ void __launch_bounds__(256) conv_splitk_reduce_kernel(const SplitReduceParams R) {
    extern __shared__ double s_red[];                       // [2][Nout] when bn_sums
    const int cg = R.Nout >> 2, vpb = 256 / cg;
    const int tv = threadIdx.x / cg, tc = threadIdx.x - tv * cg;
    const bool active = tv < vpb;
    if (R.bn_sums != nullptr) {
        for (int i = threadIdx.x; i < 2 * R.Nout; i += 256) s_red[i] = 0.0;
        __syncthreads();
    }
    float scale = R.out_scale;
    if (R.out_scale_dev) scale *= *R.out_scale_dev;
    const bool has_ep = R.ep_scale != nullptr || R.ep_shift != nullptr;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active && R.ep_scale) sc = *reinterpret_cast<const float4*>(R.ep_scale + 4 * tc);
    if (active && R.ep_shift) sh = *reinterpret_cast<const float4*>(R.ep_shift + 4 * tc);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
        for (long long v = (long long)blockIdx.x * vpb + tv; v < R.M; v += (long long)gridDim.x * vpb) {
            const float* p = R.part + v * R.Nout + 4 * tc;
            float4 a = __ldcs(reinterpret_cast<const float4*>(p));
            for (int k = 1; k < R.ksplit; ++k) {
                const float4 b = __ldcs(reinterpret_cast<const float4*>(p + (long long)k * R.part_stride));
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            float f[4] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale};
            if (has_ep) {
                f[0] = fmaf(f[0], sc.x, sh.x); f[1] = fmaf(f[1], sc.y, sh.y);
                f[2] = fmaf(f[2], sc.z, sh.z); f[3] = fmaf(f[3], sc.w, sh.w);
            }
            if (R.relu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            const long long per_n = (long long)R.D * R.HW;
            const long long nn = v / per_n, rem = v - nn * per_n;
            const int dpl = (int)(rem / R.HW);
            if (R.y != nullptr) *reinterpret_cast<float4*>(R.y + v * R.Nout + 4 * tc) = make_float4(f[0], f[1], f[2], f[3]);
            if (R.y16 != nullptr) {
                const long long within = rem - (long long)dpl * R.HW;
                __half* d16 = R.y16 + ((nn * R.Dy16 + dpl + R.y16_off) * R.HW + within) * R.Nout + 4 * tc;
                const float s16 = R.y16_scale;
                __half2 h0 = sat_half2(f[0] * s16, f[1] * s16), h1 = sat_half2(f[2] * s16, f[3] * s16);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                *reinterpret_cast<uint2*>(d16) = pk;
            }
            if (R.bn_sums != nullptr && dpl >= R.stat_lo && dpl < R.stat_hi) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { s1[j] += f[j]; s2[j] = fmaf(f[j], f[j], s2[j]); }
            }
        }
    }
    if (R.bn_sums != nullptr) {
        if (active) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(s_red + 4 * tc + j, (double)s1[j]);
                atomicAdd(s_red + R.Nout + 4 * tc + j, (double)s2[j]);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * R.Nout; i += 256) atomicAdd(R.bn_sums + i, s_red[i]);
    }
}

// ------------------------------------------------------------------------------------------------ host
int* device_error_flag();   // mode_abi.cu
long long* debug_profile_buffer();   // mode_abi.cu

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_act_map(CUtensorMap* map, const __half* x, int N, int D, int H, int W, int K, int box_w, int box_h, int box_d) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) MODE_FAIL("cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[5] = {(cuuint64_t)K, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t str[4] = {(cuuint64_t)K * 2, (cuuint64_t)W * K * 2, (cuuint64_t)H * W * K * 2,
                         (cuuint64_t)D * H * W * K * 2};
    cuuint32_t box[5] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_d, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) MODE_FAIL("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}


// ---- cost-balanced unit partition (host) ---------------------------------------------------------------------
// Estimated MMA-issue cycles of a tile of td output planes (measured SS-mode cost max(92.7, 41.7 + N/2) per MMA,
// two K-steps x 25 taps per 32-channel chunk) plus a fixed per-tile overhead (pipeline refill, epilogue hand-off).
static double tile_cost(int td, int d0, int p_lo, int p_hi, int Nt, int nchunk) {
    const int pmin = std::max(0, p_lo + 2 - d0), pmax = std::min(td + 3, p_hi + 2 - d0);
    double c = 0;
    for (int p = pmin; p <= pmax; ++p) {
        const int nq = std::min(td - 1, p) - std::max(0, p - 4) + 1;
        c += std::max(92.7, 41.7 + 0.5 * nq * Nt);
    }
    return c * 50.0 * nchunk + 3000.0;
}

// Walk units from `u` taking whole tiles while the accumulated cost stays <= budget; the last tile may be shortened.
static int64_t take_units(int64_t u, int64_t units, int D, int p_lo, int p_hi, int TD, int Nt, int nchunk, double budget) {
    double acc = 0;
    while (u < units) {
        const int d0 = (int)(u % D);
        const int full = std::min(TD, D - d0);
        const double cf = tile_cost(full, d0, p_lo, p_hi, Nt, nchunk);
        if (acc + cf <= budget) { acc += cf; u += full; continue; }
        int best = 0;                                   // largest shortened tile that still fits
        for (int td = full - 1; td >= 1; --td)
            if (acc + tile_cost(td, d0, p_lo, p_hi, Nt, nchunk) <= budget) { best = td; break; }
        u += best;
        break;
    }
    return u;
}

static void partition_units_uncached(int64_t units, int D, int p_lo, int p_hi, int TD, int Nt, int nchunk, int G, int32_t* bounds) {
    double lo = 0, hi = 0;
    for (int64_t u = 0; u < units;) {                   // upper bound: everything on one CTA
        const int d0 = (int)(u % D), full = std::min(TD, D - d0);
        hi += tile_cost(full, d0, p_lo, p_hi, Nt, nchunk);
        u += full;
    }
    lo = hi / G;
    auto feasible = [&](double budget) {
        int64_t u = 0;
        for (int c = 0; c < G && u < units; ++c) {
            const int64_t nu = take_units(u, units, D, p_lo, p_hi, TD, Nt, nchunk, budget);
            if (nu == u) return false;
            u = nu;
        }
        return u >= units;
    };
    double h = lo;
    while (!feasible(h)) h *= 1.05;
    double l = lo;
    for (int i = 0; i < 24; ++i) {
        const double m = 0.5 * (l + h);
        if (feasible(m)) h = m; else l = m;
    }
    int64_t u = 0;
    for (int c = 0; c < G; ++c) {
        bounds[c] = (int32_t)u;
        u = std::min(units, take_units(u, units, D, p_lo, p_hi, TD, Nt, nchunk, h));
    }
    bounds[G] = (int32_t)units;
}

// The partition costs ~0.4 ms of host time; it depends on the layer shape only, so it is computed once per shape.
static void partition_units(int64_t units, int D, int p_lo, int p_hi, int TD, int Nt, int nchunk, int G, int32_t* bounds) {
    struct Key {
        int64_t units; int D, p_lo, p_hi, TD, Nt, nchunk, G;
        bool operator<(const Key& o) const {
            return std::tie(units, D, p_lo, p_hi, TD, Nt, nchunk, G) <
                   std::tie(o.units, o.D, o.p_lo, o.p_hi, o.TD, o.Nt, o.nchunk, o.G);
        }
    };
    static std::mutex mu;
    static std::map<Key, std::array<int32_t, 160>> cache;
    const Key key{units, D, p_lo, p_hi, TD, Nt, nchunk, G};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
        std::array<int32_t, 160> b{};
        partition_units_uncached(units, D, p_lo, p_hi, TD, Nt, nchunk, G, b.data());
        it = cache.emplace(key, b).first;
    }
    std::copy(it->second.begin(), it->second.end(), bounds);
}

// Launch plan: output channels per CTA (Nt), and the K split.  Nt as wide as TMEM allows (<= 128: wide MMAs amortise the
// exposed A read) when (tiles x channel passes) alone fills half the GPU; otherwise -- the deep, small-volume levels of the
// U-Net (8x32x32 and below: 4..128 tiles, K up to 512) -- the 32-channel K chunks are split over CTAs as well (blockIdx.z),
// so that e.g. the 512 -> 512 bottleneck layer on 2x8x8 voxels streams its 65 MB of weights through 64 SMs instead of 16.
struct UmmaPlan { int Nt, TD, ksplit, cps; int64_t units; };
static UmmaPlan umma_plan(int N, int D, int H, int W, int K, int Nout, bool no_split) {
    const char* ms = getenv("REPMODE_UMMA_SPLITK");          // A/B + test hook: cap on the K split (1 = never split)
    const int max_split = ms ? atoi(ms) : 64;
    UmmaPlan pl{0, 1, 1, K / 32, 0};
    pl.units = (int64_t)N * (int64_t)ceil_div(H, cu::TH) * (W / cu::TW) * D;
    if (Nout > 1024) no_split = true;                       // reduce kernel: one thread per 4 channels, <= 256 per voxel
    const int nchunk = K / 32, sms = sm_count();
    double best = 0;
    for (int nt = 128; nt >= 32; nt >>= 1) {
        if (Nout % nt != 0) continue;
        const int td = max(1, min(min(256 / nt, 8), D));
        const int64_t tiles = ceil_div(pl.units, td), gx = std::min<int64_t>(tiles, sms), passes = Nout / nt;
        for (int split = 0; split < 2; ++split) {
            int ks = 1, cps = nchunk;
            if (split) {
                const int64_t ctas = gx * passes;
                if (no_split || nchunk < 2 || max_split < 2 || ctas >= sms / 2) break;     // only when the GPU is under-filled
                const int want = (int)std::min<int64_t>(std::min(nchunk, max_split), std::max<int64_t>(1, sms / ctas));
                if (want < 2) break;
                cps = (int)ceil_div(nchunk, want);
                ks = (int)ceil_div(nchunk, cps);
            }
            // estimated MMA-issue cycles of the slowest SM: waves x tiles per CTA x cycles per tile (+ the reduce pass)
            const int64_t waves = ceil_div(gx * passes * ks, sms);
            const double t = (double)waves * (double)ceil_div(tiles, gx) * tile_cost(td, 0, 0, D - 1, nt, cps) + (ks > 1 ? 8000.0 : 0.0);
            if (pl.Nt == 0 || t < 0.97 * best) {            // ties go to the wider channel pass
                best = t;
                pl.Nt = nt; pl.TD = td; pl.ksplit = ks; pl.cps = cps;
            }
        }
    }
    return pl;
}

int64_t conv3d_umma_workspace_bytes(int N, int D, int H, int W, int K, int Nout) {
    if (K % 32 != 0 || Nout % 32 != 0 || W % cu::TW != 0) return 0;
    const UmmaPlan pl = umma_plan(N, D, H, W, K, Nout, false);
    return pl.ksplit > 1 ? (int64_t)pl.ksplit * N * D * H * W * Nout * 4 : 0;
}

bool conv3d_umma_supported(int D, int H, int W, int K, int Nout) {
    (void)D;
    (void)H;
    return K % 32 == 0 && K >= 32 && Nout % 32 == 0 && Nout >= 32 && W % cu::TW == 0 &&
           true;
}

int conv3d_umma(const __half* x, const __half* w, const int32_t* sample_u, float* y, int N, int D, int H, int W, int K,
                int Nout, float out_scale, const float* out_scale_dev, double* bn_sums, int stat_lo, int stat_hi,
                const ConvExt& ext, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) ||
        (reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(ext.y16) & 15))
        MODE_FAIL("conv3d_umma: pointers must be 16-byte aligned");
    ConvParams P;
    P.w = w; P.sample_u = sample_u; P.y = y; P.bn_sums = bn_sums; P.out_scale_dev = out_scale_dev;
    P.out_scale = out_scale;
    P.N = N; P.D = D; P.H = H; P.W = W; P.K = K; P.Nout = Nout;
    P.stat_lo = stat_lo; P.stat_hi = stat_hi;
    P.ext = ext;
    P.p_lo = ext.p_lo(); P.p_hi = ext.p_hi();
    const UmmaPlan plan = umma_plan(N, D, H, W, K, Nout, ext.push.n > 0);
    if (plan.Nt == 0) MODE_FAIL("conv3d_umma: Nout=%d is not a multiple of 32", Nout);
    P.units = plan.units;
    P.Nt = plan.Nt;
    P.cps = plan.cps;
    P.part = nullptr;
    P.part_stride = (long long)N * D * H * W * Nout;
    if (plan.ksplit > 1) {
        if (ext.splitk_ws == nullptr || ext.splitk_ws_bytes < (long long)plan.ksplit * P.part_stride * 4)
            MODE_FAIL("conv3d_umma: this shape runs split-K x%d and needs a %lld-byte workspace (mode_conv3d_workspace_bytes, "
                      "mode_conv_opts_t.splitk_ws)", plan.ksplit, (long long)plan.ksplit * P.part_stride * 4);
        if (reinterpret_cast<uintptr_t>(ext.splitk_ws) & 15) MODE_FAIL("conv3d_umma: split-K workspace must be 16-byte aligned");
        P.part = (float*)ext.splitk_ws;
    }
    P.TD = plan.TD;
    P.ring = P.TD + 4;
    const int wst_bytes = 5 * P.Nt * cu::ROWB;
    const int budget = 227 * 1024 - 1024 - P.ring * cu::PLANE_BYTES - 1024 - 4096 - 2048;    // barriers, BN sums, epilogue affine
    P.wstages = min(cu::MAX_WST, budget / wst_bytes);
    if (P.wstages < 2) MODE_FAIL("conv3d_umma: shared memory budget too small for Nt=%d", P.Nt);
    const SmemLayout L = smem_layout(P.ring, P.wstages, P.Nt);
    const int smem_bytes = (int)L.total + 1024;
    if (smem_bytes > 227 * 1024) MODE_FAIL("conv3d_umma: shared memory budget exceeded (%d B)", smem_bytes);
    P.tiles_w = W / cu::TW; P.tiles_h = (int)ceil_div(H, cu::TH);
    if (P.units > 0x7fffffff) MODE_FAIL("conv3d_umma: volume too large for 32-bit unit indices");
    const int64_t total = ceil_div(P.units, P.TD);          // upper bound on useful CTAs
    P.error_flag = device_error_flag();
    if (!P.error_flag) MODE_FAIL("conv3d_umma: could not allocate the device error flag");
    P.prof = debug_profile_buffer();

    CUtensorMap xmap;
    if (make_act_map(&xmap, x, N, ext.Dx, H, W, K, cu::BW, cu::BH, 1) != 0) return -1;
    MODE_CUDA(cudaFuncSetAttribute(conv3d_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int grid = total < (int64_t)sm_count() ? (int)total : std::min(sm_count(), 159);
    partition_units(P.units, D, P.p_lo, P.p_hi, P.TD, P.Nt, P.cps, grid, P.bounds);
    conv3d_umma_kernel<<<dim3(grid, Nout / P.Nt, plan.ksplit), cu::THREADS, smem_bytes, st>>>(xmap, P);
    MODE_LAUNCH_CHECK();
    if (plan.ksplit > 1) {
        SplitReduceParams R;
        R.part = P.part; R.ksplit = plan.ksplit; R.part_stride = P.part_stride;
        R.M = (long long)N * D * H * W; R.Nout = Nout; R.D = D; R.HW = (long long)H * W;
        R.out_scale = out_scale; R.out_scale_dev = out_scale_dev;
        R.ep_scale = ext.ep_scale; R.ep_shift = ext.ep_shift; R.relu = ext.relu;
        R.y = y; R.y16 = ext.y16; R.Dy16 = ext.Dy16; R.y16_off = ext.y16_off; R.y16_scale = ext.y16_scale;
        R.bn_sums = bn_sums; R.stat_lo = stat_lo; R.stat_hi = stat_hi;
        const int vpb = 256 / (Nout / 4);
        const int rgrid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(R.M, vpb), 4 * (int64_t)sm_count()));
        const size_t rsmem = bn_sums ? 2 * (size_t)Nout * sizeof(double) : 0;
        conv_splitk_reduce_kernel<<<rgrid, 256, rsmem, st>>>(R);
        MODE_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace mode
