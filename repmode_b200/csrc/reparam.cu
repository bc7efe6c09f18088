// K1 / K1b: gate softmax + expert re-parameterisation of RepMode's MoDEConv, forward and backward.
//
// Reference behaviour being replaced (paths relative to the reference tree):
//   gate Linear -> view(N,E,Co) -> Softmax(dim=1) ........ fnet/nn_modules/RepMode.py:198-200
//   trans_kernel (centre zero padding) ..................... fnet/nn_modules/RepMode.py:165-169
//   routing (pool constants, expert mixing) ................ fnet/nn_modules/RepMode.py:171-192
// The reference launches ~(9N+5) elementwise kernels per layer, each a full pass over [Co,Ci,125];
// here one launch reads the experts once (620*Co*Ci bytes) and writes the packed conv operand(s).
//
// HBM-bound kernels: every global access is a contiguous >=64-byte segment per warp, experts are read
// exactly once per distinct gate input, and the mixing arithmetic is done in fp32 in the reference's
// own association order so the fp32 output is bit-identical up to the softmax's exp().
#include <stdlib.h>

#include "common.cuh"

namespace mode {

int* device_error_flag();   // mode_abi.cu

__device__ __forceinline__ void store_w(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_w(__half* p, float v) {
    *p = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));      // saturate instead of producing inf
}

// Position of element `col` (0..31) of row `row` inside a packed [rows][32] block.
//   fp32 pack: plain.   fp16 pack: the 64-byte-swizzled shared-memory image the UMMA B operand reads
//   (16-byte chunk index XOR bits 7-8 of the byte address = (row >> 1) & 3), so that a linear bulk copy of
//   the block into 512-byte-aligned shared memory needs no further rearrangement.
template <typename T> __device__ __forceinline__ int pack_col(int row, int col);
template <> __device__ __forceinline__ int pack_col<float>(int, int col) { return col; }
template <> __device__ __forceinline__ int pack_col<__half>(int row, int col) {
    return ((((col >> 3) ^ (row >> 1)) & 3) << 3) | (col & 7);
}

// Linear index of element (u, tap, k-chunk, row, col) in the packed weight tensor.
//   fp32 ("tap-major", read by the SIMT kernels):  w[u][tap][chunk][row][32]
//   fp16 ("stage-major", read by the tcgen05 kernel): w[u][chunk][kh*5+kw][4-kd][row][32] -- the five kd taps of
//   one (chunk, kh, kw) are consecutive blocks in DESCENDING kd order, so the taps an input plane contributes to
//   consecutive output planes form one contiguous B operand (conv_umma.cu).
template <typename T>
__device__ __forceinline__ size_t pack_index(int u, int tap, int chunk, int nchunk, int row, int nrows, int col);
template <>
__device__ __forceinline__ size_t pack_index<float>(int u, int tap, int chunk, int nchunk, int row, int nrows, int col) {
    return ((((size_t)u * 125 + tap) * nchunk + chunk) * nrows + row) * MODE_KC + col;
}
template <>
__device__ __forceinline__ size_t pack_index<__half>(int u, int tap, int chunk, int nchunk, int row, int nrows, int col) {
    const int kd = tap / 25, t = tap - kd * 25;
    return (((((size_t)u * nchunk + chunk) * 25 + t) * 5 + (4 - kd)) * nrows + row) * MODE_KC + pack_col<__half>(row, col);
}

// Task ids come from the caller's data loader: an id outside [0, T) would index gate_w out of bounds (the reference raises
// IndexError in its host loop, RepMode.py:44-49).  Here: clamp, and raise the device error flag (code 50, mode_poll_error).
__device__ __forceinline__ int checked_task(const int32_t* task_ids, int u, int T, int* err) {
    int t = task_ids[u];
    if ((unsigned)t >= (unsigned)T) {
        if (err != nullptr) atomicExch(err, 50);
        t = min(max(t, 0), T - 1);
    }
    return t;
}

// grid (Co, ceil(Ci/32), U*5), block 256: one block = one (o, 32-ci block, gate input, kd slice of 25 taps).
// The kd split multiplies the number of resident loads (a 32x32 layer is otherwise only 32 blocks deep and
// the kernel is pure memory latency).
template <typename OutT>
__global__ void __launch_bounds__(256) reparam_fwd_kernel(mode_layer_t L, const int32_t* __restrict__ task_ids,
                                                          const float* __restrict__ t_dense,
                                                          float* __restrict__ g_out, OutT* __restrict__ w_fwd,
                                                          float w_scale, const float* __restrict__ w_scale_dev,
                                                          int rows_pad, int* err) {
    __shared__ float sw[32 * 25];    // [i][tap in slice]; stride 25 is odd -> column reads are conflict free
    __shared__ float slog[MODE_NUM_EXPERTS];
    __shared__ float sg[MODE_NUM_EXPERTS];
    const int o = blockIdx.x, ic = blockIdx.y, u = blockIdx.z / 5, kds = blockIdx.z % 5;
    const int Ci = L.ci, Co = L.co, T = L.num_tasks;
    const int tid = threadIdx.x;

    if (tid < MODE_NUM_EXPERTS) {
        const int row = tid * Co + o;                         // gate row e*Co+o (RepMode.py:199)
        float logit;
        if (task_ids != nullptr) {
            logit = __fadd_rn(L.gate_w[(size_t)row * T + checked_task(task_ids, u, T, err)], L.gate_b[row]);
        } else {
            float acc = 0.f;
            for (int t = 0; t < T; ++t) acc = fmaf(t_dense[(size_t)u * T + t], L.gate_w[(size_t)row * T + t], acc);
            logit = acc + L.gate_b[row];
        }
        slog[tid] = logit;
    }
    __syncthreads();
    if (tid == 0) {
        float m = slog[0];
        for (int e = 1; e < MODE_NUM_EXPERTS; ++e) m = fmaxf(m, slog[e]);
        float ex[MODE_NUM_EXPERTS], s = 0.f;
        for (int e = 0; e < MODE_NUM_EXPERTS; ++e) { ex[e] = expf(slog[e] - m); s += ex[e]; }
        for (int e = 0; e < MODE_NUM_EXPERTS; ++e) {
            sg[e] = ex[e] / s;
            if (ic == 0 && kds == 0 && g_out != nullptr) g_out[((size_t)u * MODE_NUM_EXPERTS + e) * Co + o] = sg[e];
        }
    }
    __syncthreads();
    const float g0 = sg[0], g1 = sg[1], g2 = sg[2], g3 = sg[3], g4 = sg[4];
    if (w_scale_dev != nullptr) w_scale *= *w_scale_dev;
    const float c3 = 1.0f / 27, c5 = 1.0f / 125;               // fp32-rounded pool constants (RepMode.py:161-163)

    for (int idx = tid; idx < 32 * 25; idx += 256) {
        const int i = idx / 25, tap = kds * 25 + (idx - i * 25);
        const int c = ic * 32 + i;
        float val = 0.f;
        if (c < Ci) {
            const size_t oc = (size_t)o * Ci + c;
            const int kd = tap / 25, kh = (tap / 5) % 5, kw = tap % 5;
            const bool inner = (kd >= 1 && kd <= 3 && kh >= 1 && kh <= 3 && kw >= 1 && kw <= 3);
            const float t0 = __fmul_rn(L.k5[oc * 125 + tap], g0);
            float t1 = 0.f, t2 = 0.f, t3 = 0.f;
            if (inner) {
                t1 = __fmul_rn(L.k3[oc * 27 + ((kd - 1) * 3 + (kh - 1)) * 3 + (kw - 1)], g1);
                t3 = __fmul_rn(__fmul_rn(L.a3[oc], c3), g3);
            }
            if (tap == 62) t2 = __fmul_rn(L.k1[oc], g2);
            const float t4 = __fmul_rn(__fmul_rn(L.a5[oc], c5), g4);
            // reference association order: ((((e5 + e3) + e1) + ea3) + ea5)   (RepMode.py:184-188)
            val = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3), t4);
        }
        sw[idx] = val;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int nci = gridDim.y;
    for (int tl = warp; tl < 25; tl += 8) {
        store_w(w_fwd + pack_index<OutT>(u, kds * 25 + tl, ic, nci, o, rows_pad, lane), sw[lane * 25 + tl] * w_scale);
    }
}

// Offset of tap `tap` inside the pack of ONE gate input for K chunk `chunk`, row 0, column 0 (the row / column part and
// the gate-input stride are added by the caller): pack_index = u * pack_u_stride + pack_tap_offset + row * 32 + pack_col.
template <typename T> __device__ __forceinline__ int pack_tap_offset(int tap, int chunk, int nchunk, int nrows);
template <> __device__ __forceinline__ int pack_tap_offset<float>(int tap, int chunk, int nchunk, int nrows) {
    return (tap * nchunk + chunk) * nrows * MODE_KC;
}
template <> __device__ __forceinline__ int pack_tap_offset<__half>(int tap, int chunk, int nchunk, int nrows) {
    const int kd = tap / 25, t = tap - kd * 25;
    return ((chunk * 25 + t) * 5 + (4 - kd)) * nrows * MODE_KC;
}

// K1, row-block form (the default whenever Ci % 32 == 0 and Co % K1R_ROWS == 0): one block owns K1R_ROWS output channels x
// one 32-channel chunk x ALL 125 taps and loops over the U gate inputs, so the experts are read from HBM exactly ONCE per
// step however many distinct tasks the batch holds.  Why this shape (r2a: reparam_fwd_kernel above reaches 1.14 TB/s = 17 %
// of HBM on 512 -> 512, the 4-row staging variant tried first 0.81 TB/s): the expert slab of a (row, chunk) is ONE contiguous
// 16 000-byte run of k5 plus 3 456 bytes of k3, so the load phase is ten independent 16-byte loads per thread issued back
// to back (40 KB in flight per block, 5 blocks per SM), a single barrier, then the mix straight out of shared memory into
// 128-byte (fp16) / 256-byte (fp32) contiguous pack stores: 2 rows x 32 columns per (chunk, tap).  Shared-memory rows have
// odd strides (125 / 27 floats) so the transposing reads are conflict free.  Same expressions and association order as
// reparam_fwd_kernel: bit-identical W_eff.
// grid (Co / K1R_ROWS, Ci / 32), block 256.
constexpr int K1R_ROWS = 2;
constexpr int K1R_MAXU = 32;          // gate inputs per launch the row-block kernel takes (more -> the per-slice kernel)
template <typename OutT>
__device__ __forceinline__ void reparam_rows_body(const mode_layer_t& L, const int32_t* __restrict__ task_ids,
                                                  const float* __restrict__ t_dense, int U, float* __restrict__ g_out,
                                                  OutT* __restrict__ w_fwd, float w_scale,
                                                  const float* __restrict__ w_scale_dev, int rows_pad, int* err,
                                                  const int row_block, const int chunk, const int nci) {
    __shared__ __align__(16) float s5[K1R_ROWS * 32 * 125];
    __shared__ __align__(16) float s3[K1R_ROWS * 32 * 27];
    __shared__ __align__(16) float s1[3 * K1R_ROWS * 32];     // k1, a3, a5
    __shared__ float sg[K1R_MAXU * K1R_ROWS * MODE_NUM_EXPERTS];
    __shared__ int s_tapoff[125];          // per tap: offset inside one gate input's pack (row 0, column 0)
    __shared__ int s_t3[125];              // per tap: index into the 3^3 expert, -1 outside the inner 3^3
    const int o0 = row_block * K1R_ROWS, ic = chunk;
    const int Ci = L.ci, Co = L.co, T = L.num_tasks;
    const int tid = threadIdx.x;
    // r2l ncu (512 -> 512): issue slots 76 % busy, DRAM 23 % -- the mix loop was INSTRUCTION bound (tap -> kd/kh/kw divisions
    // and 64-bit pack addressing per element).  Everything that depends on the tap alone is tabulated once per block.
    if (tid < 125) {
        const int kd = tid / 25, kh = (tid / 5) % 5, kw = tid % 5;
        const bool inner = (kd >= 1 && kd <= 3 && kh >= 1 && kh <= 3 && kw >= 1 && kw <= 3);
        s_t3[tid] = inner ? ((kd - 1) * 3 + (kh - 1)) * 3 + (kw - 1) : -1;
        s_tapoff[tid] = pack_tap_offset<OutT>(tid, ic, nci, rows_pad);
    }
    // ---- load phase: contiguous slabs, 16-byte loads, all issued before the first use; the gate softmax of every
    // (gate input, row) runs on the threads at the END of the block while those loads are in flight (its dependent global
    // loads -- task id, then gate column -- would otherwise add ~1.5 us of two-thread latency to every block)
    {
        constexpr int V5 = 32 * 125 / 4, V3 = 32 * 27 / 4;     // float4 per row slab: 1000, 216
        float4 r5[K1R_ROWS][4], r3[K1R_ROWS];
#pragma unroll
        for (int r = 0; r < K1R_ROWS; ++r) {
            const size_t oc0 = (size_t)(o0 + r) * Ci + (size_t)ic * 32;
            const float4* src5 = reinterpret_cast<const float4*>(L.k5 + oc0 * 125);
            const float4* src3 = reinterpret_cast<const float4*>(L.k3 + oc0 * 27);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (tid + 256 * k < V5) r5[r][k] = src5[tid + 256 * k];
            if (tid < V3) r3[r] = src3[tid];
        }
        float v1 = 0.f;
        if (tid < 3 * K1R_ROWS * 32) {
            const int which = tid / (K1R_ROWS * 32), r = (tid / 32) % K1R_ROWS, i = tid & 31;
            const float* src = which == 0 ? L.k1 : (which == 1 ? L.a3 : L.a5);
            v1 = src[(size_t)(o0 + r) * Ci + (size_t)ic * 32 + i];
        }
        // gate column + bias -> softmax over the 5 experts, one thread per (gate input, output channel) (RepMode.py:198-200)
        const int gt = 255 - tid;
        if (gt < U * K1R_ROWS) {
            const int u = gt / K1R_ROWS, r = gt % K1R_ROWS, o = o0 + r;
            float lg[MODE_NUM_EXPERTS];
            for (int e = 0; e < MODE_NUM_EXPERTS; ++e) {
                const int row = e * Co + o;
                if (task_ids != nullptr) {
                    lg[e] = __fadd_rn(L.gate_w[(size_t)row * T + checked_task(task_ids, u, T, err)], L.gate_b[row]);
                } else {
                    float acc = 0.f;
                    for (int t = 0; t < T; ++t) acc = fmaf(t_dense[(size_t)u * T + t], L.gate_w[(size_t)row * T + t], acc);
                    lg[e] = acc + L.gate_b[row];
                }
            }
            float m = lg[0];
            for (int e = 1; e < MODE_NUM_EXPERTS; ++e) m = fmaxf(m, lg[e]);
            float ex[MODE_NUM_EXPERTS], ssum = 0.f;
            for (int e = 0; e < MODE_NUM_EXPERTS; ++e) { ex[e] = expf(lg[e] - m); ssum += ex[e]; }
            for (int e = 0; e < MODE_NUM_EXPERTS; ++e) {
                const float gv = ex[e] / ssum;
                sg[(u * K1R_ROWS + r) * MODE_NUM_EXPERTS + e] = gv;
                if (ic == 0 && g_out != nullptr) g_out[((size_t)u * MODE_NUM_EXPERTS + e) * Co + o] = gv;
            }
        }
#pragma unroll
        for (int r = 0; r < K1R_ROWS; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (tid + 256 * k < V5) reinterpret_cast<float4*>(s5 + r * 32 * 125)[tid + 256 * k] = r5[r][k];
            if (tid < V3) reinterpret_cast<float4*>(s3 + r * 32 * 27)[tid] = r3[r];
        }
        if (tid < 3 * K1R_ROWS * 32) s1[tid] = v1;
    }
    if (w_scale_dev != nullptr) w_scale *= *w_scale_dev;
    const float c3 = 1.0f / 27, c5 = 1.0f / 125;               // fp32-rounded pool constants (RepMode.py:161-163)
    __syncthreads();                                           // slabs and gates landed
    for (int u = 0; u < U; ++u) {
        // mix + pack: a warp owns whole taps (tap = warp, warp + 8, ...), its lanes the 32 columns; the shared-memory reads
        // walk a column at an odd stride (conflict free) and each store instruction writes one whole 64-byte (fp16) /
        // 128-byte (fp32) pack row
        const int warp = tid >> 5, lane = tid & 31;
        OutT* wu = w_fwd + (size_t)u * 125 * nci * rows_pad * MODE_KC;
#pragma unroll
        for (int r = 0; r < K1R_ROWS; ++r) {
            const float* g = sg + (u * K1R_ROWS + r) * MODE_NUM_EXPERTS;
            const float g0 = g[0], g1 = g[1];
            // per-column constants of this row: 1^3 expert (centre tap only), avg3 (inner taps), avg5 (every tap)
            const float t2c = __fmul_rn(s1[(0 * K1R_ROWS + r) * 32 + lane], g[2]);
            const float t3c = __fmul_rn(__fmul_rn(s1[(1 * K1R_ROWS + r) * 32 + lane], c3), g[3]);
            const float t4c = __fmul_rn(__fmul_rn(s1[(2 * K1R_ROWS + r) * 32 + lane], c5), g[4]);
            const float* k5c = s5 + (r * 32 + lane) * 125;
            const float* k3c = s3 + (r * 32 + lane) * 27;
            OutT* wrow = wu + (o0 + r) * MODE_KC + pack_col<OutT>(o0 + r, lane);
#pragma unroll 4
            for (int tap = warp; tap < 125; tap += 8) {
                const int t3i = s_t3[tap];
                const float t0 = __fmul_rn(k5c[tap], g0);
                float t1 = 0.f, t2 = 0.f, t3 = 0.f;
                if (t3i >= 0) {
                    t1 = __fmul_rn(k3c[t3i], g1);
                    t3 = t3c;
                }
                if (tap == 62) t2 = t2c;
                // reference association order: ((((e5 + e3) + e1) + ea3) + ea5)   (RepMode.py:184-188)
                const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3), t4c);
                store_w(wrow + s_tapoff[tap], val * w_scale);
            }
        }
    }
}

template <typename OutT>
__global__ void __launch_bounds__(256, 4) reparam_fwd_rows_kernel(mode_layer_t L, const int32_t* __restrict__ task_ids,
                                                               const float* __restrict__ t_dense, int U,
                                                               float* __restrict__ g_out, OutT* __restrict__ w_fwd,
                                                               float w_scale, const float* __restrict__ w_scale_dev,
                                                               int rows_pad, int* err) {
    reparam_rows_body<OutT>(L, task_ids, t_dense, U, g_out, w_fwd, w_scale, w_scale_dev, rows_pad, err, (int)blockIdx.x,
                            (int)blockIdx.y, (int)gridDim.y);
}

// K1 for SEVERAL layers in one launch ("a single grouped launch, not 19x"): the per-layer descriptors
// travel as a __grid_constant__ kernel parameter; a block finds its layer from the cumulative block counts and runs the
// row-block body above on it.  All layers share the gate inputs of the step (Net.forward passes the same t to every
// MoDEConv, RepMode.py:51-71).  The 17 row-eligible layers of the U-Net are 4 000 - 70 000 blocks of identical shape: the
// 7 us latency floor of the narrow layers disappears into the wide ones.
struct K1GroupItem {
    mode_layer_t L;
    float* g_out;
    void* w_fwd;
    void* w_dgrad;
    int rows_pad;              // fwd pack rows (co padded to 32)
    int block_begin;           // first block of this layer in the grouped K1 grid
    long long tile_begin;      // first 32x32 tile of this layer in the grouped dgrad-pack grid
};
struct K1Group {
    int n;
    K1GroupItem it[MODE_REPARAM_GROUP_MAX];
};
template <typename OutT>
__global__ void __launch_bounds__(256, 4) reparam_fwd_rows_grouped_kernel(const __grid_constant__ K1Group G,
                                                                       const int32_t* __restrict__ task_ids,
                                                                       const float* __restrict__ t_dense, int U,
                                                                       float w_scale, int* err) {
    const int b = (int)blockIdx.x;
    int k = 0;
    while (k + 1 < G.n && b >= G.it[k + 1].block_begin) ++k;
    const K1GroupItem& it = G.it[k];
    const int local = b - it.block_begin;
    const int row_blocks = it.L.co / K1R_ROWS;
    reparam_rows_body<OutT>(it.L, task_ids, t_dense, U, it.g_out, (OutT*)it.w_fwd, w_scale, nullptr, it.rows_pad, err,
                            local % row_blocks, local / row_blocks, it.L.ci / 32);
}

// fwd pack (rows = co, k = ci)  ->  dgrad pack (tap -> 124-tap, rows = ci, k = co)
// grid (ceil(Ci/32), ceil(Co/32), U*125), block (32, 8)
template <typename T>
__global__ void __launch_bounds__(256) pack_dgrad_kernel(const T* __restrict__ src, T* __restrict__ dst, int Ci,
                                                         int Co, int src_rows_pad, int dst_rows_pad) {
    __shared__ T tile[32][33];
    const int ic = blockIdx.x, oc = blockIdx.y;
    const int u = blockIdx.z / 125, tap = blockIdx.z % 125;
    const int nci = gridDim.x, nco = gridDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += 8) {
        const int o = oc * 32 + r;
        T v = T(0);
        if (o < Co) v = src[pack_index<T>(u, tap, ic, nci, o, src_rows_pad, tx)];
        tile[r][tx] = v;                                     // (o = oc*32+r, i = ic*32+tx)
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = ic * 32 + r;
        if (i < Ci) dst[pack_index<T>(u, 124 - tap, oc, nco, i, dst_rows_pad, tx)] = tile[tx][r];
    }
}

// fwd pack -> dgrad pack for the fp16 (stage-major, 64-byte-swizzled) layout with 16-byte global accesses (r2u; the kernel
// above moves 2 bytes per thread and reached 1.5 TB/s: 1.05 ms of a 23 ms batch-4 train step).  A 32 x 32 tile of one
// (gate input, tap, ci chunk, co block) is a CONTIGUOUS 2 KB block of the source pack and a contiguous 2 KB block of the
// destination pack: 128 threads move one tile (one 16-byte chunk each way, transposed through shared memory), a block of
// 256 threads PD_TILES tiles with all of its loads in flight before the first use.  Pad rows / columns of the source are
// zero (the caller zero-initialises padded packs), so whole tiles are moved blindly.  grid ceil(n_tiles / PD_TILES).
constexpr int PD_TILES = 8;
struct PdLayer {                 // one layer's packs as the tile mover sees them
    const __half* src; __half* dst;
    int nci, nco, src_rows_pad, dst_rows_pad;
};
template <typename Locate>
__device__ __forceinline__ void pack_dgrad_h16_body(const Locate& locate, long long n_tiles) {
    __shared__ __align__(16) __half tile[PD_TILES][32][34];       // [ci][o], 68-byte rows (4-byte aligned, odd word stride)
    const int tid = threadIdx.x, grp = tid >> 7, j = tid & 127;
    const int r = j >> 2, pc = j & 3;
    const int lc = pc ^ ((r >> 1) & 3);                           // logical 8-column chunk behind physical chunk pc of row r
    const long long t0 = (long long)blockIdx.x * PD_TILES;
    int hint = locate.first_layer(t0);                            // block-uniform: the layer of the block's first tile
    uint4 v[PD_TILES / 2];
    __half* dptr[PD_TILES / 2];
#pragma unroll
    for (int k = 0; k < PD_TILES / 2; ++k) {
        const long long tg = t0 + 2 * k + grp;
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        dptr[k] = nullptr;
        if (tg < n_tiles) {
            PdLayer P;
            const long long tl = locate(tg, P, hint);             // tile index inside its layer (hint only moves forward)
            const int oc = (int)(tl % P.nco);
            long long q = tl / P.nco;
            const int ic = (int)(q % P.nci);
            q /= P.nci;
            const int tap = (int)(q % 125), u = (int)(q / 125);
            const int kd = tap / 25, t = tap - kd * 25;
            const size_t sbase = (((((size_t)u * P.nci + ic) * 25 + t) * 5 + (4 - kd)) * P.src_rows_pad + (size_t)oc * 32) * MODE_KC;
            // destination tap 124 - tap = (4 - kd, 24 - t): its block index inside the stage is 4 - (4 - kd) = kd
            dptr[k] = P.dst + (((((size_t)u * P.nco + oc) * 25 + (24 - t)) * 5 + kd) * P.dst_rows_pad + (size_t)ic * 32) * MODE_KC;
            v[k] = *reinterpret_cast<const uint4*>(P.src + sbase + r * 32 + pc * 8);
        }
    }
#pragma unroll
    for (int k = 0; k < PD_TILES / 2; ++k) {
        const __half* h = reinterpret_cast<const __half*>(&v[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) tile[2 * k + grp][lc * 8 + e][r] = h[e];     // source (row o = r, col ci) -> [ci][o]
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PD_TILES / 2; ++k) {
        if (dptr[k] == nullptr) continue;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&tile[2 * k + grp][r][lc * 8]);   // dest row ci = r, cols o
        *reinterpret_cast<uint4*>(dptr[k] + r * 32 + pc * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
struct PdSingleLocate {
    PdLayer P;
    __device__ __forceinline__ int first_layer(long long) const { return 0; }
    __device__ __forceinline__ long long operator()(long long tg, PdLayer& out, int&) const { out = P; return tg; }
};
__global__ void __launch_bounds__(256) pack_dgrad_h16_kernel(const __half* __restrict__ src, __half* __restrict__ dst,
                                                             int nci, int nco, int src_rows_pad, int dst_rows_pad,
                                                             long long n_tiles) {
    pack_dgrad_h16_body(PdSingleLocate{PdLayer{src, dst, nci, nco, src_rows_pad, dst_rows_pad}}, n_tiles);
}
// the same over the layers of a K1 group (one launch; a block's tiles may belong to two neighbouring layers)
struct PdGroupLocate {
    const K1Group* G;
    __device__ __forceinline__ int first_layer(long long t0) const {
        int k = 0;
        while (k + 1 < G->n && t0 >= G->it[k + 1].tile_begin) ++k;
        return k;
    }
    __device__ __forceinline__ long long operator()(long long tg, PdLayer& out, int& k) const {
        while (k + 1 < G->n && tg >= G->it[k + 1].tile_begin) ++k;      // a block's 8 tiles rarely cross a layer boundary
        const K1GroupItem& it = G->it[k];
        const int nci = it.L.ci / 32, nco = it.L.co / 32;
        out = PdLayer{(const __half*)it.w_fwd, (__half*)it.w_dgrad, nci, nco, nco * 32, nci * 32};
        return tg - it.tile_begin;
    }
};
__global__ void __launch_bounds__(256) pack_dgrad_h16_grouped_kernel(const __grid_constant__ K1Group G, long long n_tiles) {
    pack_dgrad_h16_body(PdGroupLocate{&G}, n_tiles);
}

// K1b main kernel. grid (ceil(Ci/32), Co), block 128 -- the ci chunk is the FAST grid index, so blocks that run together
// read neighbouring 128-byte segments of the same d_weff rows (DRAM page locality). Loops over the samples; every global
// access is a contiguous segment (dW rows of 32 ci, expert slabs of 32*125 floats).
__global__ void __launch_bounds__(128) reparam_bwd_kernel(mode_layer_t L, const int32_t* __restrict__ sample_u,
                                                          int n_samples, const float* __restrict__ g,
                                                          const float* __restrict__ d_weff,
                                                          float* __restrict__ dk5, float* __restrict__ dk3,
                                                          float* __restrict__ dk1, float* __restrict__ da3,
                                                          float* __restrict__ da5, float* __restrict__ dg_part) {
    __shared__ float sk5[32 * 125];
    __shared__ float sk3[32 * 27];
    __shared__ float sdk3[32 * 27];
    __shared__ float s_all[32], s_inner[32], s_center[32];
    __shared__ float s_dg[2];
    const int ic = blockIdx.x, o = blockIdx.y;
    const int Ci = L.ci, Co = L.co;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = ic * 32 + lane;
    const bool live = c < Ci;
    const int nlive = min(32, Ci - ic * 32);
    const float c3 = 1.0f / 27, c5 = 1.0f / 125;

    // stage this (o, ci-block)'s experts: contiguous slabs
    {
        const float* src5 = L.k5 + ((size_t)o * Ci + ic * 32) * 125;
        for (int idx = tid; idx < 32 * 125; idx += 128) sk5[idx] = idx < nlive * 125 ? src5[idx] : 0.f;
        const float* src3 = L.k3 + ((size_t)o * Ci + ic * 32) * 27;
        for (int idx = tid; idx < 32 * 27; idx += 128) {
            sk3[idx] = idx < nlive * 27 ? src3[idx] : 0.f;
            sdk3[idx] = 0.f;
        }
        if (tid < 32) { s_all[tid] = 0.f; s_inner[tid] = 0.f; s_center[tid] = 0.f; }
        if (tid < 2) s_dg[tid] = 0.f;
    }
    const float k1v = live ? L.k1[(size_t)o * Ci + c] : 0.f;
    const float a3v = live ? L.a3[(size_t)o * Ci + c] * c3 : 0.f;
    const float a5v = live ? L.a5[(size_t)o * Ci + c] * c5 : 0.f;
    __syncthreads();

    float acc5[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc5[j] = 0.f;
    float acc_k1 = 0.f, acc_a3 = 0.f, acc_a5 = 0.f;

    for (int n = 0; n < n_samples; ++n) {
        const int u = sample_u[n];
        const float* gu = g + (size_t)u * MODE_NUM_EXPERTS * Co + o;
        const float g0 = gu[0], g1 = gu[Co], g2 = gu[2 * Co], g3 = gu[3 * Co], g4 = gu[4 * Co];
        const float* dw = d_weff + ((size_t)n * 125 * Co + o) * Ci + c;
        float p_all = 0.f, p_inner = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int tap = warp + 4 * j;
            if (tap < 125) {
                const float v = live ? dw[(size_t)tap * Co * Ci] : 0.f;
                acc5[j] = fmaf(g0, v, acc5[j]);
                p_all += v;
                q0 = fmaf(sk5[lane * 125 + tap], v, q0);
                const int kd = tap / 25, kh = (tap / 5) % 5, kw = tap % 5;
                if (kd >= 1 && kd <= 3 && kh >= 1 && kh <= 3 && kw >= 1 && kw <= 3) {
                    const int t3 = ((kd - 1) * 3 + (kh - 1)) * 3 + (kw - 1);
                    p_inner += v;
                    q1 = fmaf(sk3[lane * 27 + t3], v, q1);
                    sdk3[lane * 27 + t3] = fmaf(g1, v, sdk3[lane * 27 + t3]);   // owned by this thread only
                    if (tap == 62) s_center[lane] = v;
                }
            }
        }
        atomicAdd(&s_all[lane], p_all);
        atomicAdd(&s_inner[lane], p_inner);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            q0 += __shfl_xor_sync(0xffffffffu, q0, s);
            q1 += __shfl_xor_sync(0xffffffffu, q1, s);
        }
        if (lane == 0) { atomicAdd(&s_dg[0], q0); atomicAdd(&s_dg[1], q1); }
        __syncthreads();
        if (warp == 0) {
            const float A = s_all[lane], I = s_inner[lane], Cn = s_center[lane];
            acc_k1 = fmaf(g2, Cn, acc_k1);
            acc_a3 = fmaf(g3, I * c3, acc_a3);
            acc_a5 = fmaf(g4, A * c5, acc_a5);
            float d2 = k1v * Cn, d3 = a3v * I, d4 = a5v * A;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                d2 += __shfl_xor_sync(0xffffffffu, d2, s);
                d3 += __shfl_xor_sync(0xffffffffu, d3, s);
                d4 += __shfl_xor_sync(0xffffffffu, d4, s);
            }
            if (lane == 0) {
                float* dst = dg_part + (((size_t)ic * n_samples + n) * MODE_NUM_EXPERTS) * Co + o;
                dst[0] = s_dg[0];
                dst[Co] = s_dg[1];
                dst[2 * Co] = d2;
                dst[3 * Co] = d3;
                dst[4 * Co] = d4;
                s_dg[0] = 0.f;
                s_dg[1] = 0.f;
            }
            s_all[lane] = 0.f;
            s_inner[lane] = 0.f;
            s_center[lane] = 0.f;
        }
        __syncthreads();
    }

    // write expert gradients through shared memory so the global stores are contiguous
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int tap = warp + 4 * j;
        if (tap < 125) sk5[lane * 125 + tap] = acc5[j];
    }
    __syncthreads();
    float* dst5 = dk5 + ((size_t)o * Ci + ic * 32) * 125;
    for (int idx = tid; idx < nlive * 125; idx += 128) dst5[idx] = sk5[idx];
    float* dst3 = dk3 + ((size_t)o * Ci + ic * 32) * 27;
    for (int idx = tid; idx < nlive * 27; idx += 128) dst3[idx] = sdk3[idx];
    if (warp == 0 && live) {
        dk1[(size_t)o * Ci + c] = acc_k1;
        da3[(size_t)o * Ci + c] = acc_a3;
        da5[(size_t)o * Ci + c] = acc_a5;
    }
}

// K1b, slab form (default): same arithmetic as reparam_bwd_kernel, restructured for occupancy and coalescing.  r2k: the
// kernel above needs 186 registers (32 loads in flight + 32 accumulators per thread) -> 2 blocks of 128 threads per SM and
// 1.1 TB/s = 17 % of HBM on 512 -> 512.  Here a block of 256 threads owns one (o, 32-channel chunk): the d_weff slab of a
// sample (125 segments of 128 bytes) goes through shared memory with four 16-byte loads per thread, the expert slabs with
// five; every thread then owns 16 fixed (ci, tap) entries of dk5 (a contiguous 16 000-byte run of the output) and 4 of dk3,
// so the gradients leave in coalesced runs and the per-thread state is ~20 accumulators: <= 85 registers, 3 blocks per SM.
// grid (ceil(Ci/32), Co), block 256.
constexpr int K1B_DW_STRIDE = 33;     // floats per tap row of the staged d_weff slab (odd: the transposing reads are conflict free)
__global__ void __launch_bounds__(256, 3) reparam_bwd_slab_kernel(mode_layer_t L, const int32_t* __restrict__ sample_u,
                                                                  int n_samples, const float* __restrict__ g,
                                                                  const float* __restrict__ d_weff,
                                                                  float* __restrict__ dk5, float* __restrict__ dk3,
                                                                  float* __restrict__ dk1, float* __restrict__ da3,
                                                                  float* __restrict__ da5, float* __restrict__ dg_part) {
    __shared__ float sk5[32 * 125];
    __shared__ float sk3[32 * 27];
    __shared__ float sdw[125 * K1B_DW_STRIDE];
    __shared__ float s_col[2][8][32];        // per-warp column sums: all taps / inner taps
    __shared__ float s_q[2][8];
    const int ic = blockIdx.x, o = blockIdx.y;
    const int Ci = L.ci, Co = L.co;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nlive = min(32, Ci - ic * 32);
    const float c3 = 1.0f / 27, c5 = 1.0f / 125;
    const size_t oc0 = (size_t)o * Ci + (size_t)ic * 32;
    // stage this (o, ci-block)'s experts: contiguous slabs (zeros beyond the live channels)
    for (int idx = tid; idx < 32 * 125; idx += 256) sk5[idx] = idx < nlive * 125 ? L.k5[oc0 * 125 + idx] : 0.f;
    for (int idx = tid; idx < 32 * 27; idx += 256) sk3[idx] = idx < nlive * 27 ? L.k3[oc0 * 27 + idx] : 0.f;
    const bool colthr = tid < 32;
    const bool live = colthr && lane < nlive;
    const float k1v = live ? L.k1[oc0 + lane] : 0.f;
    const float a3v = live ? L.a3[oc0 + lane] * c3 : 0.f;
    const float a5v = live ? L.a5[oc0 + lane] * c5 : 0.f;
    float acc5[16], acc3[4];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc5[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc3[k] = 0.f;
    float acc_k1 = 0.f, acc_a3 = 0.f, acc_a5 = 0.f;
    const bool vec_ok = (Ci & 3) == 0 && nlive == 32 && (reinterpret_cast<uintptr_t>(d_weff) & 15) == 0;

    for (int n = 0; n < n_samples; ++n) {
        const int u = sample_u[n];
        const float* gu = g + (size_t)u * MODE_NUM_EXPERTS * Co + o;
        const float g0 = gu[0], g1 = gu[Co], g2 = gu[2 * Co], g3 = gu[3 * Co], g4 = gu[4 * Co];
        const float* dw = d_weff + ((size_t)n * 125 * Co + o) * Ci + (size_t)ic * 32;
        __syncthreads();                                   // experts staged (n = 0) / previous slab consumed
        if (vec_ok) {
            float4 v[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int tap = (tid >> 3) + 32 * p;
                if (tap < 125) v[p] = *reinterpret_cast<const float4*>(dw + (size_t)tap * Co * Ci + (tid & 7) * 4);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int tap = (tid >> 3) + 32 * p;
                if (tap < 125) {
                    float* d = sdw + tap * K1B_DW_STRIDE + (tid & 7) * 4;
                    d[0] = v[p].x; d[1] = v[p].y; d[2] = v[p].z; d[3] = v[p].w;
                }
            }
        } else {
            for (int idx = tid; idx < 125 * 32; idx += 256) {
                const int tap = idx >> 5, ci = idx & 31;
                sdw[tap * K1B_DW_STRIDE + ci] = ci < nlive ? dw[(size_t)tap * Co * Ci + ci] : 0.f;
            }
        }
        __syncthreads();
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int idx = tid + 256 * k;
            if (idx < 32 * 125) {
                const int ci = idx / 125, tap = idx - ci * 125;
                const float v = sdw[tap * K1B_DW_STRIDE + ci];
                acc5[k] = fmaf(g0, v, acc5[k]);
                q0 = fmaf(sk5[idx], v, q0);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = tid + 256 * k;
            if (idx < 32 * 27) {
                const int ci = idx / 27, t3 = idx - ci * 27;
                const int tap = (t3 / 9 + 1) * 25 + ((t3 / 3) % 3 + 1) * 5 + (t3 % 3) + 1;
                const float v = sdw[tap * K1B_DW_STRIDE + ci];
                acc3[k] = fmaf(g1, v, acc3[k]);
                q1 = fmaf(sk3[idx], v, q1);
            }
        }
        // column sums over the taps (all / inner): warp w takes taps w, w + 8, ...; lane = ci
        float p_all = 0.f, p_inner = 0.f;
        for (int tap = warp; tap < 125; tap += 8) {
            const float v = sdw[tap * K1B_DW_STRIDE + lane];
            const int kd = tap / 25, kh = (tap / 5) % 5, kw = tap % 5;
            p_all += v;
            if (kd >= 1 && kd <= 3 && kh >= 1 && kh <= 3 && kw >= 1 && kw <= 3) p_inner += v;
        }
        s_col[0][warp][lane] = p_all;
        s_col[1][warp][lane] = p_inner;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            q0 += __shfl_xor_sync(0xffffffffu, q0, sft);
            q1 += __shfl_xor_sync(0xffffffffu, q1, sft);
        }
        if (lane == 0) { s_q[0][warp] = q0; s_q[1][warp] = q1; }
        __syncthreads();
        if (colthr) {
            float A = 0.f, I = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { A += s_col[0][w][lane]; I += s_col[1][w][lane]; }
            const float Cn = sdw[62 * K1B_DW_STRIDE + lane];
            acc_k1 = fmaf(g2, Cn, acc_k1);
            acc_a3 = fmaf(g3, I * c3, acc_a3);
            acc_a5 = fmaf(g4, A * c5, acc_a5);
            float d2 = k1v * Cn, d3 = a3v * I, d4 = a5v * A;
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                d2 += __shfl_xor_sync(0xffffffffu, d2, sft);
                d3 += __shfl_xor_sync(0xffffffffu, d3, sft);
                d4 += __shfl_xor_sync(0xffffffffu, d4, sft);
            }
            if (lane == 0) {
                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) { t0 += s_q[0][w]; t1 += s_q[1][w]; }
                float* dst = dg_part + (((size_t)ic * n_samples + n) * MODE_NUM_EXPERTS) * Co + o;
                dst[0] = t0;
                dst[Co] = t1;
                dst[2 * Co] = d2;
                dst[3 * Co] = d3;
                dst[4 * Co] = d4;
            }
        }
    }
    // expert gradients leave in contiguous runs
    float* dst5 = dk5 + oc0 * 125;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int idx = tid + 256 * k;
        if (idx < nlive * 125) dst5[idx] = acc5[k];
    }
    float* dst3 = dk3 + oc0 * 27;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = tid + 256 * k;
        if (idx < nlive * 27) dst3[idx] = acc3[k];
    }
    if (live) {
        dk1[oc0 + lane] = acc_k1;
        da3[oc0 + lane] = acc_a3;
        da5[oc0 + lane] = acc_a5;
    }
}

// K1b, register form (r2u, default whenever the 32-channel chunk is whole): same arithmetic as the slab kernel, without the
// shared-memory round trip of d_weff.  The slab kernel staged every sample's [125][32] slab through shared memory (two
// barriers per sample, no load of sample n + 1 in flight while sample n is consumed) and sat at 1.5 TB/s = 23 % of HBM on
// 512 -> 512 (r2q), 1.68 ms of a 23 ms batch-4 train step.  Here a thread KEEPS the 16 values it loaded -- taps tr, tr + 32,
// tr + 64, tr + 96 of one channel quad (four 16-byte loads) -- as its share of the work: the dk5 / dk3 accumulators for
// exactly those (tap, ci) entries, its part of the two expert dot products (the staged experts are read at [ci][tap], odd
// stride: conflict free) and of the per-channel column sums (reduced by shuffles inside the warp, across the 8 warps through
// a double-buffered table: ONE barrier per sample).  The next sample's loads are issued before the current one is consumed.
// dk5 leaves through the expert buffer (coalesced 16-byte stores), dk3 is accumulated in shared memory (every (ci, t3) entry
// has exactly one owner thread).  grid (Ci / 32, Co), block 256.
// d_weff read straight out of K4's work-unit partials (r2j): when a layer's wgrad runs ONE slab per unit group (every wide
// layer of the U-Net: N * Ci/32 * Co/32 >= 50), the reduce kernel of wgrad_deep.cu is a pure re-layout -- partial
// [unit][entry][o % 32][ci % 32] -> d_weff [n][tap][o][ci], times the dy scale -- that K1b can do in its own loads: a tap's
// 32-channel row is a contiguous 128-byte run in either layout.  Saves writing and re-reading d_weff (the largest tensor of
// the backward pass of those layers).  (0 + v) * scale is what the reduce kernel computes for one slab: bit-identical.
struct K4Partials {
    const float* partial;          // nullptr: read d_weff
    int nL, nA;                    // first unit of the A / B groups (SL = SA = SB = 1)
    int ncic, ncoc;
    float scale;
    const float* scale_dev;
};
template <bool PARTIAL>
__global__ void __launch_bounds__(256, 3) reparam_bwd_reg_kernel(mode_layer_t L, const int32_t* __restrict__ sample_u,
                                                                 int n_samples, const float* __restrict__ g,
                                                                 const float* __restrict__ d_weff,
                                                                 float* __restrict__ dk5, float* __restrict__ dk3,
                                                                 float* __restrict__ dk1, float* __restrict__ da3,
                                                                 float* __restrict__ da5, float* __restrict__ dg_part,
                                                                 K4Partials kp) {
    __shared__ __align__(16) float sk5[32 * 125];     // experts [ci][tap]; reused as the dk5 staging buffer at the end
    __shared__ __align__(16) float sk3[32 * 27];
    __shared__ __align__(16) float sdk3[32 * 27];
    __shared__ float s_col[2][2][8][32];              // [buffer][all / inner taps][warp][ci]
    __shared__ float s_q[2][2][8];                    // [buffer][k5 / k3 dot product][warp]
    __shared__ float s_cn[2][32];                     // centre-tap values [buffer][ci]
    const int ic = blockIdx.x, o = blockIdx.y;
    const int Ci = L.ci, Co = L.co;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cq = tid & 7, tr = tid >> 3;            // channel quad 4cq..4cq+3; taps tr + 32p
    const float c3 = 1.0f / 27, c5 = 1.0f / 125;
    const size_t oc0 = (size_t)o * Ci + (size_t)ic * 32;
    const size_t tap_stride = (size_t)Co * Ci;

    float4 cur[4], nxt[4];
    // partial mode: float offset of this thread's 16 bytes of each of its four taps inside sample 0's units; a sample further
    // on is `pstride` floats away (one unit per group and kind)
    uint32_t poff[4] = {0u, 0u, 0u, 0u};
    size_t pstride = 0;
    float pscale = 1.f;
    if (PARTIAL) {
        const uint32_t g0 = (uint32_t)((o >> 5) * kp.ncic + ic);
        pstride = (size_t)kp.ncoc * kp.ncic * K4_DEEP_PARTIAL_FLOATS;
        pscale = kp.scale * (kp.scale_dev != nullptr ? *kp.scale_dev : 1.f);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int tap = min(tr + 32 * p, 124);
            const int kd = tap / 25, kh = (tap / 5) % 5, kw = tap % 5;
            uint32_t unit, entry;
            if (kh == 4) { unit = g0; entry = kd * 5 + kw; }
            else if (kd < 2) { unit = kp.nL + g0; entry = kd * 20 + kh * 5 + kw; }
            else { unit = kp.nL + kp.nA + g0; entry = (kd - 2) * 20 + kh * 5 + kw; }
            poff[p] = unit * (uint32_t)K4_DEEP_PARTIAL_FLOATS + (entry * 32u + (uint32_t)(o & 31)) * 32u + cq * 4;
        }
    }
    auto load_slab = [&](int n, float4 (&v)[4]) {
        if (PARTIAL) {
            const float* base = kp.partial + (size_t)n * pstride;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tr + 32 * p < 125) {
                    w = ld_stream(base + poff[p]);
                    w.x = (0.f + w.x) * pscale; w.y = (0.f + w.y) * pscale;
                    w.z = (0.f + w.z) * pscale; w.w = (0.f + w.w) * pscale;
                }
                v[p] = w;
            }
            return;
        }
        const float* dw = d_weff + ((size_t)n * 125 * Co + o) * Ci + (size_t)ic * 32 + cq * 4;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int tap = tr + 32 * p;
            v[p] = tap < 125 ? ld_stream(dw + (size_t)tap * tap_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    load_slab(0, cur);
    // this thread's taps: index into the 3^3 expert (-1 outside the inner 3^3)
    int t3i[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int tap = tr + 32 * p;
        const int kd = tap / 25, kh = (tap / 5) % 5, kw = tap % 5;
        const bool inner = tap < 125 && kd >= 1 && kd <= 3 && kh >= 1 && kh <= 3 && kw >= 1 && kw <= 3;
        t3i[p] = inner ? ((kd - 1) * 3 + (kh - 1)) * 3 + (kw - 1) : -1;
    }
    {   // experts of this (o, chunk): contiguous slabs, 16-byte loads
        const float4* s5 = reinterpret_cast<const float4*>(L.k5 + oc0 * 125);
        const float4* s3 = reinterpret_cast<const float4*>(L.k3 + oc0 * 27);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (tid + 256 * k < 1000) reinterpret_cast<float4*>(sk5)[tid + 256 * k] = s5[tid + 256 * k];
        if (tid < 216) {
            reinterpret_cast<float4*>(sk3)[tid] = s3[tid];
            reinterpret_cast<float4*>(sdk3)[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const bool colthr = tid < 32;
    const float k1v = colthr ? L.k1[oc0 + lane] : 0.f;
    const float a3v = colthr ? L.a3[oc0 + lane] * c3 : 0.f;
    const float a5v = colthr ? L.a5[oc0 + lane] * c5 : 0.f;
    float acc5[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc5[p][q] = 0.f;
    float acc_k1 = 0.f, acc_a3 = 0.f, acc_a5 = 0.f;
    __syncthreads();                                  // experts staged

    for (int n = 0; n < n_samples; ++n) {
        const int buf = n & 1;
        if (n + 1 < n_samples) load_slab(n + 1, nxt);
        const int u = sample_u[n];
        const float* gu = g + (size_t)u * MODE_NUM_EXPERTS * Co + o;
        const float g0 = gu[0], g1 = gu[Co], g2 = gu[2 * Co], g3 = gu[3 * Co], g4 = gu[4 * Co];
        float q0 = 0.f, q1 = 0.f, pa[4] = {0.f, 0.f, 0.f, 0.f}, pi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int tap = tr + 32 * p;
            if (tap < 125) {
                const float vv[4] = {cur[p].x, cur[p].y, cur[p].z, cur[p].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ci = cq * 4 + q;
                    const float v = vv[q];
                    acc5[p][q] = fmaf(g0, v, acc5[p][q]);
                    q0 = fmaf(sk5[ci * 125 + tap], v, q0);
                    pa[q] += v;
                    if (t3i[p] >= 0) {
                        pi[q] += v;
                        q1 = fmaf(sk3[ci * 27 + t3i[p]], v, q1);
                        sdk3[ci * 27 + t3i[p]] = fmaf(g1, v, sdk3[ci * 27 + t3i[p]]);      // owned by this thread only
                    }
                    if (tap == 62) s_cn[buf][ci] = v;
                }
            }
        }
        // column sums: lanes with the same channel quad (lane & 7) hold different taps
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            pa[q] += __shfl_xor_sync(0xffffffffu, pa[q], 8);
            pa[q] += __shfl_xor_sync(0xffffffffu, pa[q], 16);
            pi[q] += __shfl_xor_sync(0xffffffffu, pi[q], 8);
            pi[q] += __shfl_xor_sync(0xffffffffu, pi[q], 16);
        }
        if (lane < 8) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                s_col[buf][0][warp][lane * 4 + q] = pa[q];
                s_col[buf][1][warp][lane * 4 + q] = pi[q];
            }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            q0 += __shfl_xor_sync(0xffffffffu, q0, sft);
            q1 += __shfl_xor_sync(0xffffffffu, q1, sft);
        }
        if (lane == 0) { s_q[buf][0][warp] = q0; s_q[buf][1][warp] = q1; }
        __syncthreads();          // the only barrier per sample: buffer `buf` is rewritten two samples later, i.e. after the
                                  // next barrier, which warp 0 only reaches once it has read it
        if (colthr) {
            float A = 0.f, I = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { A += s_col[buf][0][w][lane]; I += s_col[buf][1][w][lane]; }
            const float Cn = s_cn[buf][lane];
            acc_k1 = fmaf(g2, Cn, acc_k1);
            acc_a3 = fmaf(g3, I * c3, acc_a3);
            acc_a5 = fmaf(g4, A * c5, acc_a5);
            float d2 = k1v * Cn, d3 = a3v * I, d4 = a5v * A;
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                d2 += __shfl_xor_sync(0xffffffffu, d2, sft);
                d3 += __shfl_xor_sync(0xffffffffu, d3, sft);
                d4 += __shfl_xor_sync(0xffffffffu, d4, sft);
            }
            if (lane == 0) {
                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) { t0 += s_q[buf][0][w]; t1 += s_q[buf][1][w]; }
                float* dst = dg_part + (((size_t)ic * n_samples + n) * MODE_NUM_EXPERTS) * Co + o;
                dst[0] = t0;
                dst[Co] = t1;
                dst[2 * Co] = d2;
                dst[3 * Co] = d3;
                dst[4 * Co] = d4;
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) cur[p] = nxt[p];
    }
    __syncthreads();                                  // every thread is done with the staged experts
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int tap = tr + 32 * p;
        if (tap < 125) {
#pragma unroll
            for (int q = 0; q < 4; ++q) sk5[(cq * 4 + q) * 125 + tap] = acc5[p][q];
        }
    }
    __syncthreads();
    float4* dst5 = reinterpret_cast<float4*>(dk5 + oc0 * 125);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (tid + 256 * k < 1000) dst5[tid + 256 * k] = reinterpret_cast<const float4*>(sk5)[tid + 256 * k];
    if (tid < 216) reinterpret_cast<float4*>(dk3 + oc0 * 27)[tid] = reinterpret_cast<const float4*>(sdk3)[tid];
    if (colthr) {
        dk1[oc0 + lane] = acc_k1;
        da3[oc0 + lane] = acc_a3;
        da5[oc0 + lane] = acc_a5;
    }
}

// softmax + Linear backward, block form (r2v, default): the one-thread-per-channel kernel below is a single latency chain
// (5 * nci dependent-issue loads per sample on 1..4 warps of the whole GPU: 30 us per layer at batch 4, 0.58 ms of a 21.8 ms
// train step in r2u).  Here a block owns 32 output channels (lane = channel): warp w turns sample n0 + w of a chunk of eight
// into its logit gradients (all 5 * nci partial sums of a sample are independent, coalesced 128-byte loads), then the owner
// thread (expert e = warp, channel = lane) folds the chunk into db and into its [T] row of dgate_w in shared memory --
// samples in order, starting from 0: the same summation order as the serial kernel, bit-identical results.
// grid ceil(Co / 32), block 256, dynamic shared memory 5 * 32 * T floats.
__global__ void __launch_bounds__(256) gate_bwd_block_kernel(mode_layer_t L, const int32_t* __restrict__ task_ids,
                                                             const float* __restrict__ t_dense,
                                                             const int32_t* __restrict__ sample_u, int n_samples, int nci,
                                                             const float* __restrict__ g, const float* __restrict__ dg_part,
                                                             float* __restrict__ dgate_w, float* __restrict__ dgate_b,
                                                             int* err) {
    extern __shared__ float s_w[];                    // [expert][channel][T]
    __shared__ float s_dl[8][MODE_NUM_EXPERTS][32];
    __shared__ int s_col[8];
    const int Co = L.co, T = L.num_tasks;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int o = blockIdx.x * 32 + lane;
    const bool valid = o < Co;
    for (int i = tid; i < MODE_NUM_EXPERTS * 32 * T; i += 256) s_w[i] = 0.f;
    float db = 0.f;
    for (int n0 = 0; n0 < n_samples; n0 += 8) {
        const int n = n0 + warp;
        if (n < n_samples) {
            const int u = sample_u[n];
            if (lane == 0) s_col[warp] = task_ids != nullptr ? checked_task(task_ids, u, T, err) : u;
            float gv[MODE_NUM_EXPERTS], dg[MODE_NUM_EXPERTS], s = 0.f;
#pragma unroll
            for (int e = 0; e < MODE_NUM_EXPERTS; ++e) {
                gv[e] = valid ? g[((size_t)u * MODE_NUM_EXPERTS + e) * Co + o] : 0.f;
                dg[e] = 0.f;
            }
            if (valid) {
#pragma unroll 4
                for (int b = 0; b < nci; ++b) {
                    const float* src = dg_part + (((size_t)b * n_samples + n) * MODE_NUM_EXPERTS) * Co + o;
#pragma unroll
                    for (int e = 0; e < MODE_NUM_EXPERTS; ++e) dg[e] += src[(size_t)e * Co];
                }
            }
#pragma unroll
            for (int e = 0; e < MODE_NUM_EXPERTS; ++e) s = fmaf(gv[e], dg[e], s);
#pragma unroll
            for (int e = 0; e < MODE_NUM_EXPERTS; ++e) s_dl[warp][e][lane] = gv[e] * (dg[e] - s);
        }
        __syncthreads();
        if (warp < MODE_NUM_EXPERTS) {
            float* wrow = s_w + ((size_t)warp * 32 + lane) * T;
            const int cnt = min(8, n_samples - n0);
            for (int j = 0; j < cnt; ++j) db += s_dl[j][warp][lane];
            for (int t = 0; t < T; ++t) {              // per-chunk partial first, then added: the serial kernel's rounding
                float w = 0.f;
                if (task_ids != nullptr) {
                    for (int j = 0; j < cnt; ++j) w += s_col[j] == t ? s_dl[j][warp][lane] : 0.f;
                } else {
                    for (int j = 0; j < cnt; ++j) w = fmaf(s_dl[j][warp][lane], t_dense[(size_t)s_col[j] * T + t], w);
                }
                wrow[t] = n0 == 0 ? w : wrow[t] + w;
            }
        }
        __syncthreads();
    }
    // [E][Co][T]: the 32 channels of this block are one contiguous run of 32 * T floats per expert
    const int run = 32 * T;
    for (int i = tid; i < MODE_NUM_EXPERTS * run; i += 256) {
        const int e = i / run, r = i - e * run;
        if (blockIdx.x * 32 + r / T < Co) dgate_w[((size_t)e * Co + (size_t)blockIdx.x * 32) * T + r] = s_w[i];
    }
    if (warp < MODE_NUM_EXPERTS && valid) dgate_b[(size_t)warp * Co + o] = db;
}

// softmax + Linear backward; one thread per output channel o, samples in order -> deterministic.
__global__ void gate_bwd_kernel(mode_layer_t L, const int32_t* __restrict__ task_ids,
                                const float* __restrict__ t_dense, const int32_t* __restrict__ sample_u,
                                int n_samples, int nci, const float* __restrict__ g,
                                const float* __restrict__ dg_part, float* __restrict__ dgate_w,
                                float* __restrict__ dgate_b, int* err) {
    // r2u: samples are taken eight at a time with their logit gradients held in registers and dgate_w written once per
    // (expert, task) at the end of the chunk -- the first version zeroed dgate_w in global memory and then read-modified-
    // wrote it per sample (a chain of dependent global round trips: 10 us for ONE sample, 21 us at batch 4, r2q).  Same
    // summation order as before (samples in order, starting from 0): bit-identical results.
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int Co = L.co, T = L.num_tasks;
    if (o >= Co) return;
    float db[MODE_NUM_EXPERTS];
#pragma unroll
    for (int e = 0; e < MODE_NUM_EXPERTS; ++e) db[e] = 0.f;
    for (int n0 = 0; n0 < n_samples; n0 += 8) {
        float dlv[8][MODE_NUM_EXPERTS];
        int col[8];                                   // task column (id mode) / gate-input row (dense mode); -1 = no sample
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            col[j] = -1;
#pragma unroll
            for (int e = 0; e < MODE_NUM_EXPERTS; ++e) dlv[j][e] = 0.f;
            const int n = n0 + j;
            if (n < n_samples) {
                const int u = sample_u[n];
                col[j] = task_ids != nullptr ? checked_task(task_ids, u, T, err) : u;
                float gv[MODE_NUM_EXPERTS], dg[MODE_NUM_EXPERTS], s = 0.f;
#pragma unroll
                for (int e = 0; e < MODE_NUM_EXPERTS; ++e) {
                    gv[e] = g[((size_t)u * MODE_NUM_EXPERTS + e) * Co + o];
                    float a = 0.f;
                    for (int b = 0; b < nci; ++b) a += dg_part[(((size_t)b * n_samples + n) * MODE_NUM_EXPERTS + e) * Co + o];
                    dg[e] = a;
                    s = fmaf(gv[e], a, s);
                }
#pragma unroll
                for (int e = 0; e < MODE_NUM_EXPERTS; ++e) {
                    dlv[j][e] = gv[e] * (dg[e] - s);
                    db[e] += dlv[j][e];
                }
            }
        }
        for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int e = 0; e < MODE_NUM_EXPERTS; ++e) {
                float w = 0.f;
                if (task_ids != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) w += col[j] == t ? dlv[j][e] : 0.f;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (col[j] >= 0) w = fmaf(dlv[j][e], t_dense[(size_t)col[j] * T + t], w);
                }
                float* dst = dgate_w + ((size_t)e * Co + o) * T + t;
                *dst = n0 == 0 ? w : *dst + w;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < MODE_NUM_EXPERTS; ++e) dgate_b[(size_t)e * Co + o] = db[e];
}

}  // namespace mode

using namespace mode;

extern "C" int64_t mode_packed_weight_elems(int32_t k_channels, int32_t n_channels) {
    return (int64_t)MODE_TAPS * ceil_div(k_channels, MODE_KC) * n_channels * MODE_KC;
}
// fp16 pack: rows are padded to a multiple of 32 as well (the tcgen05 kernel's N granularity); pad rows are not
// written by K1 -- the caller zero-initialises the buffer when n_channels % 32 != 0.
extern "C" int64_t mode_packed_weight_elems_f16(int32_t k_channels, int32_t n_channels) {
    return (int64_t)MODE_TAPS * ceil_div(k_channels, MODE_KC) * ceil_div(n_channels, 32) * 32 * MODE_KC;
}

extern "C" int mode_reparam_fwd(const mode_layer_t* L, const int32_t* task_ids, const float* t_dense, int32_t U,
                                float* g_out, void* w_fwd, void* w_dgrad, mode_dtype_t w_dtype, float w_scale,
                                const float* w_scale_dev, void* stream) {
    if (!L || !w_fwd || U <= 0) MODE_FAIL("mode_reparam_fwd: null layer/output or U <= 0");
    if ((task_ids == nullptr) == (t_dense == nullptr)) MODE_FAIL("mode_reparam_fwd: pass exactly one of task_ids / t_dense");
    if (L->ci <= 0 || L->co <= 0 || L->num_tasks <= 0) MODE_FAIL("mode_reparam_fwd: bad layer dims");
    if ((int64_t)U * 125 > 65535) MODE_FAIL("mode_reparam_fwd: U too large (%d)", U);
    cudaStream_t st = (cudaStream_t)stream;
    int* err = device_error_flag();
    const int nci = (int)ceil_div(L->ci, 32), nco = (int)ceil_div(L->co, 32);
    dim3 grid(L->co, nci, U * 5);
    // row-block kernel (experts read once for all U gate inputs) whenever the chunk is whole; REPMODE_K1_ROWS=0 forces the
    // per-(row, kd slice) kernel (read per call so that a test can A/B both)
    const char* rows_env = getenv("REPMODE_K1_ROWS");
    const bool rows_on = !(rows_env != nullptr && rows_env[0] == '0');
    const bool rows = rows_on && U <= K1R_MAXU && L->ci % 32 == 0 && L->co % K1R_ROWS == 0 &&
                      ((reinterpret_cast<uintptr_t>(L->k5) | reinterpret_cast<uintptr_t>(L->k3)) & 15) == 0;
    const dim3 rgrid(L->co / K1R_ROWS, nci);
    if (w_dtype == MODE_F32) {
        if (rows)
            reparam_fwd_rows_kernel<float><<<rgrid, 256, 0, st>>>(*L, task_ids, t_dense, U, g_out, (float*)w_fwd, w_scale,
                                                                  w_scale_dev, L->co, err);
        else
            reparam_fwd_kernel<float><<<grid, 256, 0, st>>>(*L, task_ids, t_dense, g_out, (float*)w_fwd, w_scale,
                                                            w_scale_dev, L->co, err);
        MODE_LAUNCH_CHECK();
        if (w_dgrad) {
            pack_dgrad_kernel<float><<<dim3(nci, nco, U * 125), dim3(32, 8), 0, st>>>((const float*)w_fwd,
                                                                                   (float*)w_dgrad, L->ci, L->co, L->co,
                                                                                   L->ci);
            MODE_LAUNCH_CHECK();
        }
    } else if (w_dtype == MODE_F16) {
        if (rows)
            reparam_fwd_rows_kernel<__half><<<rgrid, 256, 0, st>>>(*L, task_ids, t_dense, U, g_out, (__half*)w_fwd, w_scale,
                                                                   w_scale_dev, nco * 32, err);
        else
            reparam_fwd_kernel<__half><<<grid, 256, 0, st>>>(*L, task_ids, t_dense, g_out, (__half*)w_fwd, w_scale,
                                                             w_scale_dev, nco * 32, err);
        MODE_LAUNCH_CHECK();
        if (w_dgrad) {
            const char* pd_env = getenv("REPMODE_PACK_DGRAD_V1");                // A/B arm: the 2-byte-per-thread kernel
            const bool pd_v1 = (pd_env != nullptr && pd_env[0] == '1') ||
                               ((reinterpret_cast<uintptr_t>(w_fwd) | reinterpret_cast<uintptr_t>(w_dgrad)) & 15) != 0;
            if (pd_v1) {
                pack_dgrad_kernel<__half><<<dim3(nci, nco, U * 125), dim3(32, 8), 0, st>>>((const __half*)w_fwd,
                                                                                        (__half*)w_dgrad, L->ci, L->co,
                                                                                        nco * 32, nci * 32);
            } else {
                const long long n_tiles = (long long)U * 125 * nci * nco;
                pack_dgrad_h16_kernel<<<(unsigned)ceil_div(n_tiles, PD_TILES), 256, 0, st>>>(
                    (const __half*)w_fwd, (__half*)w_dgrad, nci, nco, nco * 32, nci * 32, n_tiles);
            }
            MODE_LAUNCH_CHECK();
        }
    } else {
        MODE_FAIL("mode_reparam_fwd: unknown dtype %d", (int)w_dtype);
    }
    return 0;
}

extern "C" int mode_reparam_fwd_grouped(const mode_reparam_item_t* items, int32_t n_items, const int32_t* task_ids,
                                        const float* t_dense, int32_t U, mode_dtype_t w_dtype, float w_scale,
                                        void* stream) {
    if (!items || n_items <= 0 || n_items > MODE_REPARAM_GROUP_MAX || U <= 0 || U > K1R_MAXU)
        MODE_FAIL("mode_reparam_fwd_grouped: bad arguments (n_items=%d, U=%d)", n_items, U);
    if ((task_ids == nullptr) == (t_dense == nullptr))
        MODE_FAIL("mode_reparam_fwd_grouped: pass exactly one of task_ids / t_dense");
    if (w_dtype != MODE_F16) MODE_FAIL("mode_reparam_fwd_grouped: fp16 packs only (the tensor-core path)");
    K1Group G;
    G.n = n_items;
    long long blocks = 0, tiles = 0;
    bool any_dgrad = false;
    for (int k = 0; k < n_items; ++k) {
        const mode_reparam_item_t& s = items[k];
        const mode_layer_t& L = s.layer;
        if (L.ci <= 0 || L.co <= 0 || L.num_tasks <= 0 || L.ci % 32 != 0 || L.co % 32 != 0 || !s.w_fwd)
            MODE_FAIL("mode_reparam_fwd_grouped: item %d: needs Ci %% 32 == 0, Co %% 32 == 0 and a w_fwd buffer", k);
        if (((reinterpret_cast<uintptr_t>(L.k5) | reinterpret_cast<uintptr_t>(L.k3) | reinterpret_cast<uintptr_t>(s.w_fwd) |
              reinterpret_cast<uintptr_t>(s.w_dgrad)) & 15) != 0)
            MODE_FAIL("mode_reparam_fwd_grouped: item %d: experts and packs must be 16-byte aligned", k);
        K1GroupItem& it = G.it[k];
        it.L = L; it.g_out = s.g_out; it.w_fwd = s.w_fwd; it.w_dgrad = s.w_dgrad;
        it.rows_pad = L.co;                          // Co % 32 == 0: no padding rows
        it.block_begin = (int)blocks;
        it.tile_begin = tiles;
        blocks += (long long)(L.co / K1R_ROWS) * (L.ci / 32);
        if (s.w_dgrad) { tiles += (long long)U * 125 * (L.ci / 32) * (L.co / 32); any_dgrad = true; }
        if (blocks > 0x7fffffffLL) MODE_FAIL("mode_reparam_fwd_grouped: too many blocks");
    }
    // layers without a dgrad pack (no input gradient wanted) own an empty tile range: tile_begin of the next layer equals
    // theirs, and the locate loop (">= next begin") skips them
    cudaStream_t st = (cudaStream_t)stream;
    reparam_fwd_rows_grouped_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(G, task_ids, t_dense, U, w_scale,
                                                                             device_error_flag());
    MODE_LAUNCH_CHECK();
    if (any_dgrad) {
        pack_dgrad_h16_grouped_kernel<<<(unsigned)ceil_div(tiles, PD_TILES), 256, 0, st>>>(G, tiles);
        MODE_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int64_t mode_reparam_bwd_workspace_bytes(int32_t ci, int32_t co, int32_t n_samples) {
    return (int64_t)ceil_div(ci, 32) * n_samples * MODE_NUM_EXPERTS * co * sizeof(float);
}

static int reparam_bwd_launch(const mode_layer_t* L, const int32_t* task_ids, const float* t_dense, int32_t U,
                              const int32_t* sample_u, int32_t n_samples, const float* g, const float* d_weff,
                              const K4Partials& kp, float* dk5, float* dk3, float* dk1, float* da3, float* da5,
                              float* dgate_w, float* dgate_b, void* workspace, void* stream) {
    if (!L || !sample_u || !g || (!d_weff && !kp.partial) || !workspace) MODE_FAIL("mode_reparam_bwd: null argument");
    if ((task_ids == nullptr) == (t_dense == nullptr)) MODE_FAIL("mode_reparam_bwd: pass exactly one of task_ids / t_dense");
    if (!dk5 || !dk3 || !dk1 || !da3 || !da5 || !dgate_w || !dgate_b) MODE_FAIL("mode_reparam_bwd: null output");
    (void)U;
    cudaStream_t st = (cudaStream_t)stream;
    const int nci = (int)ceil_div(L->ci, 32);
    static const bool legacy = getenv("REPMODE_K1B_LEGACY") != nullptr;       // A/B arm: the round-1 kernel
    const bool reg_ok = L->ci % 32 == 0 &&
                        ((reinterpret_cast<uintptr_t>(d_weff) | reinterpret_cast<uintptr_t>(L->k5) |
                          reinterpret_cast<uintptr_t>(L->k3) | reinterpret_cast<uintptr_t>(dk5) |
                          reinterpret_cast<uintptr_t>(dk3) | reinterpret_cast<uintptr_t>(kp.partial)) & 15) == 0;
    if (kp.partial != nullptr) {
        if (!reg_ok || L->co % 32 != 0) MODE_FAIL("mode_reparam_bwd_partial: needs Ci %% 32 == 0, Co %% 32 == 0, 16-byte aligned buffers");
        reparam_bwd_reg_kernel<true><<<dim3(nci, L->co), 256, 0, st>>>(*L, sample_u, n_samples, g, nullptr, dk5, dk3, dk1,
                                                                       da3, da5, (float*)workspace, kp);
    } else if (legacy)
        reparam_bwd_kernel<<<dim3(nci, L->co), 128, 0, st>>>(*L, sample_u, n_samples, g, d_weff, dk5, dk3, dk1, da3, da5,
                                                            (float*)workspace);
    else if (reg_ok && getenv("REPMODE_K1B_SLAB") == nullptr)
        reparam_bwd_reg_kernel<false><<<dim3(nci, L->co), 256, 0, st>>>(*L, sample_u, n_samples, g, d_weff, dk5, dk3, dk1,
                                                                        da3, da5, (float*)workspace, K4Partials{});
    else
        reparam_bwd_slab_kernel<<<dim3(nci, L->co), 256, 0, st>>>(*L, sample_u, n_samples, g, d_weff, dk5, dk3, dk1, da3,
                                                                  da5, (float*)workspace);
    MODE_LAUNCH_CHECK();
    const size_t gate_smem = (size_t)MODE_NUM_EXPERTS * 32 * L->num_tasks * sizeof(float);
    const bool gate_serial = getenv("REPMODE_GATE_BWD_SERIAL") != nullptr;      // A/B arm: one thread per channel
    if (!gate_serial && gate_smem <= 40 * 1024)
        gate_bwd_block_kernel<<<(unsigned)ceil_div(L->co, 32), 256, gate_smem, st>>>(
            *L, task_ids, t_dense, sample_u, n_samples, nci, g, (const float*)workspace, dgate_w, dgate_b,
            device_error_flag());
    else
        gate_bwd_kernel<<<(unsigned)ceil_div(L->co, 128), 128, 0, st>>>(*L, task_ids, t_dense, sample_u, n_samples, nci, g,
                                                                      (const float*)workspace, dgate_w, dgate_b,
                                                                      device_error_flag());
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_reparam_bwd(const mode_layer_t* L, const int32_t* task_ids, const float* t_dense, int32_t U,
                                const int32_t* sample_u, int32_t n_samples, const float* g, const float* d_weff,
                                float* dk5, float* dk3, float* dk1, float* da3, float* da5, float* dgate_w,
                                float* dgate_b, void* workspace, void* stream) {
    return reparam_bwd_launch(L, task_ids, t_dense, U, sample_u, n_samples, g, d_weff, K4Partials{}, dk5, dk3, dk1, da3, da5,
                              dgate_w, dgate_b, workspace, stream);
}

extern "C" int mode_reparam_bwd_partial(const mode_layer_t* L, const int32_t* task_ids, const float* t_dense, int32_t U,
                                        const int32_t* sample_u, int32_t n_samples, const float* g,
                                        const float* k4_partials, const int32_t* layout5, float scale,
                                        const float* scale_dev, float* dk5, float* dk3, float* dk1, float* da3,
                                        float* da5, float* dgate_w, float* dgate_b, void* workspace, void* stream) {
    if (!L || !k4_partials || !layout5) MODE_FAIL("mode_reparam_bwd_partial: null argument");
    if (layout5[0] != 1 || layout5[1] != 1 || layout5[2] != 1)
        MODE_FAIL("mode_reparam_bwd_partial: the wgrad ran several slabs per unit (%d, %d, %d): reduce it first", layout5[0],
                  layout5[1], layout5[2]);
    const int64_t groups = (int64_t)n_samples * (L->ci / 32) * (L->co / 32);
    if (groups * 3 * K4_DEEP_PARTIAL_FLOATS > 0x7fffffffLL * 2)
        MODE_FAIL("mode_reparam_bwd_partial: partials too large for 32-bit offsets");
    K4Partials kp{k4_partials, layout5[3], layout5[4], L->ci / 32, L->co / 32, scale, scale_dev};
    return reparam_bwd_launch(L, task_ids, t_dense, U, sample_u, n_samples, g, nullptr, kp, dk5, dk3, dk1, da3, da5, dgate_w,
                              dgate_b, workspace, stream);
}
