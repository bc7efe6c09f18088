// BatchNorm3d (+ReLU) on NDHWC fp32, training and eval, forward and backward, plus the fp32->fp16
// operand staging kernels.  Replaces `subsequent_layer = Sequential(BatchNorm3d(Co), ReLU)`
// (fnet/nn_modules/RepMode.py:146-149, applied at :212) and its autograd.
//
// All kernels are HBM-bound streaming passes: float4 accesses, channel index = element index mod C
// (C is a multiple of 4 on every layer that has a BN; the scalar path covers the rest), grid sized to
// a multiple of the SM count, per-thread fp32 partials folded into fp64 block/global sums.
#include <stdlib.h>

#include <map>
#include <mutex>

#include "common.cuh"
#include "peer.cuh"

namespace mode {

int* device_error_flag();   // mode_abi.cu

constexpr int BN_THREADS = 256;
constexpr int BN_MAXC = 1024;


// D-sharded slab bookkeeping (mode_planes_t by value); kind(row): 0 = outside the global volume, 1 = halo copy,
// 2 = owned
struct Planes {
    long long rows_per_plane;
    int D, own_lo, own_hi, valid_lo, valid_hi;
    __device__ __forceinline__ int kind(long long row) const {
        // 32-bit division whenever the tensor allows it (always on this path: < 2^32 voxels per slab)
        const int d = (row >> 32) == 0 && (rows_per_plane >> 32) == 0
                          ? (int)(((unsigned)row / (unsigned)rows_per_plane) % (unsigned)D)
                          : (int)((row / rows_per_plane) % D);
        if (d < valid_lo || d >= valid_hi) return 0;
        return (d >= own_lo && d < own_hi) ? 2 : 1;
    }
};
// every plane owned and inside the volume: only m_global matters (a slab whose halos live elsewhere)
static bool planes_trivial(const mode_planes_t* p) {
    return p->own_lo <= 0 && p->own_hi >= p->D && p->valid_lo <= 0 && p->valid_hi >= p->D;
}
// Where row r of the tensor a kernel WALKS lives in a second tensor it reads or writes (mode_rowmap_t by value):
//   d2s: the walked tensor is the [voxels][8][C] output of the transposed stride-2 GEMM (RepMode.py:97-101 as a GEMM, rows in
//        (n, d, h, w, kd, kh, kw) order) and the mapped tensor is the NDHWC volume of twice the size per axis -- the
//        depth-to-space scatter / gather happens inside the BatchNorm kernels instead of as a permute copy;
//   pitch: floats between consecutive rows of the mapped tensor (a channel range of a wider tensor: the gradient of one
//        half of the decoder's concatenated input).
struct RowMap {
    int active, d2s;
    int D, H, W;               // the LOW-resolution grid (rows / 8 = N * D * H * W)
    long long pitch;
    __device__ __forceinline__ long long offset(long long row, int C) const {     // element offset of (row, channel 0)
        if (!active) return row * C;
        long long r = row;
        if (d2s) {
            const unsigned tap = (unsigned)(row & 7);
            unsigned m = (unsigned)(row >> 3);                                    // < 2^32 voxels on this path
            const unsigned w = m % (unsigned)W; m /= (unsigned)W;
            const unsigned h = m % (unsigned)H; m /= (unsigned)H;
            const unsigned d = m % (unsigned)D; const unsigned n = m / (unsigned)D;
            r = (((long long)n * (2 * D) + (2 * d + (tap >> 2))) * (2 * H) + (2 * h + ((tap >> 1) & 1))) * (2LL * W) +
                (2 * w + (tap & 1));
        }
        return r * pitch;
    }
};
static RowMap to_rowmap(const mode_rowmap_t* m, int C) {
    RowMap r{0, 0, 1, 1, 1, (long long)C};
    if (m != nullptr) {
        r.d2s = m->d2s; r.D = m->D; r.H = m->H; r.W = m->W;
        r.pitch = m->pitch > 0 ? m->pitch : C;
        r.active = (r.d2s != 0 || r.pitch != C) ? 1 : 0;
    }
    return r;
}
static Planes to_planes(const mode_planes_t* p) {
    Planes q;
    q.rows_per_plane = p->rows_per_plane; q.D = p->D; q.own_lo = p->own_lo; q.own_hi = p->own_hi;
    q.valid_lo = p->valid_lo; q.valid_hi = p->valid_hi;
    return q;
}


// Deterministic in-block reduction of per-thread 4-channel partials: thread t owns channels 4*(t % vpr) + j.
// Partials go to shared memory once; thread c < 2C then sums the 256/vpr threads that share its channel
// (no shared-memory atomics: with 256 threads on 32 addresses they serialise for longer than the streaming loop).
__device__ __forceinline__ void block_fold_and_flush(const double (&ds)[4], const double (&dq)[4], int C, int vpr,
                                                     double* __restrict__ gout, double* sh /* [2][4][256] */) {
    const int t = threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) { sh[j * BN_THREADS + t] = ds[j]; sh[(4 + j) * BN_THREADS + t] = dq[j]; }
    __syncthreads();
    for (int c = t; c < 2 * C; c += BN_THREADS) {
        const int which = c / C, ch = c % C, lv = ch >> 2, j = ch & 3;
        double a = 0.0;
        for (int k = lv; k < BN_THREADS; k += vpr) a += sh[(which * 4 + j) * BN_THREADS + k];
        atomicAdd(gout + c, a);
    }
}

// ---- statistics: sums[c] += sum y, sums[C+c] += sum y^2 -------------------------------------------------
// Each thread owns a fixed group of 4 channels (requires (blockDim*4) % C == 0 or C % (blockDim*4) == 0 ->
// we use rows-of-C iteration instead: thread t handles float4 #t of a row block).
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const float* __restrict__ y, int64_t M, int C,
                                                              double* __restrict__ sums) {
    // vector lanes per row
    const int vpr = C >> 2;                         // float4 per row (C % 4 == 0 path)
    const int rows_per_iter = BN_THREADS / vpr;     // host guarantees vpr <= BN_THREADS and divides it
    const int lane_v = threadIdx.x % vpr, lane_r = threadIdx.x / vpr;
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
    int cnt = 0;
    for (int64_t r = (int64_t)blockIdx.x * rows_per_iter + lane_r; r < M; r += (int64_t)gridDim.x * rows_per_iter) {
        const float4 v = *reinterpret_cast<const float4*>(y + r * C + lane_v * 4);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
        if (++cnt == 64) {   // fold fp32 partials into fp64 regularly
#pragma unroll
            for (int j = 0; j < 4; ++j) { ds[j] += s[j]; dq[j] += q[j]; s[j] = 0.f; q[j] = 0.f; }
            cnt = 0;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { ds[j] += s[j]; dq[j] += q[j]; }
    __shared__ double sh[8 * BN_THREADS];
    block_fold_and_flush(ds, dq, C, vpr, sums, sh);
}

__global__ void bn_stats_scalar_kernel(const float* __restrict__ y, int64_t M, int C, double* __restrict__ sums) {
    // generic C (used only for tiny test shapes): one block per channel
    const int c = blockIdx.x;
    double s = 0, q = 0;
    for (int64_t r = threadIdx.x; r < M; r += blockDim.x) {
        const double v = y[r * C + c];
        s += v; q += v * v;
    }
    __shared__ double sh[2][BN_THREADS];
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = q;
    __syncthreads();
    for (int st = BN_THREADS / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) { sh[0][threadIdx.x] += sh[0][threadIdx.x + st]; sh[1][threadIdx.x] += sh[1][threadIdx.x + st]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(sums + c, sh[0][0]); atomicAdd(sums + C + c, sh[1][0]); }
}

// With `g` (D-sharded slabs over peer memory): ONE block; thread 0 waits until every rank's {sum, sum of squares} vector has
// landed in the local slots, then the statistics are the rank-ordered sum of the slots (deterministic) instead of `sums`.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int64_t M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* mean,
                                   float* invstd, float* scale, float* shift, float* running_mean,
                                   float* running_var, PeerGather g, int* error_flag) {
    uint32_t target = 0;
    if (g.on()) {
        if (!gather_wait(g, target) && threadIdx.x == 0) atomicExch(error_flag, 43);
    }
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    double s1, s2;
    if (g.on()) {
        const double* sl = reinterpret_cast<const double*>(g.slots);
        s1 = 0.0; s2 = 0.0;
        for (int r = 0; r < g.world; ++r) { s1 += sl[(size_t)r * 2 * C + c]; s2 += sl[(size_t)r * 2 * C + C + c]; }
    } else {
        s1 = sums[c]; s2 = sums[C + c];
    }
    const double m = s1 / (double)M;
    double var = s2 / (double)M - m * m;      // biased variance (training-mode normalisation)
    if (var < 0) var = 0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    if (mean) mean[c] = (float)m;
    if (invstd) invstd[c] = is;
    const float sc = ga * is;
    scale[c] = sc;
    shift[c] = be - (float)m * sc;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    if (running_var) {
        const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;   // running_var tracks the unbiased estimate
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
    }
    if (g.on()) {
        __syncthreads();
        if (threadIdx.x == 0) *g.expect = target;
    }
}

// ---- apply: out = relu(y*scale + shift), optional fp16 copy ----------------------------------------------
// Four independent 16-byte loads in flight per thread (one per thread leaves the SM far below the bytes-in-flight HBM
// needs), and the tensor is walked BACKWARDS: K2 has just written y front to back, so its tail is what the L2 still holds.
// With `fin.sums` the kernel FINALIZES the statistics itself (r2w): every block derives scale / shift of all C channels from
// the fp64 sums (same arithmetic as bn_finalize_kernel, so the same bits), block 0 also writes mean / invstd / scale / shift
// and updates the running statistics -- one kernel boundary less on the critical path of every training forward.
struct BnFinalize {
    const double* sums; long long M; const float* gamma; const float* beta; float eps, momentum;
    float *mean, *invstd, *scale, *shift, *running_mean, *running_var;
};
template <bool PLANES>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const float* __restrict__ y, int64_t total, int C,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int relu,
                                                              float* __restrict__ out, __half* __restrict__ out16,
                                                              float f16_scale, Planes pl, BnFinalize fin, RowMap om) {
    __shared__ float ssc[BN_MAXC], ssh[BN_MAXC];
    if (fin.sums != nullptr) {
        for (int c = threadIdx.x; c < C; c += BN_THREADS) {
            const double s1 = fin.sums[c], s2 = fin.sums[C + c];
            const double m = s1 / (double)fin.M;
            double var = s2 / (double)fin.M - m * m;
            if (var < 0) var = 0;
            const float is = (float)(1.0 / sqrt(var + (double)fin.eps));
            const float ga = fin.gamma ? fin.gamma[c] : 1.f, be = fin.beta ? fin.beta[c] : 0.f;
            const float sc = ga * is;
            const float sh = be - (float)m * sc;
            ssc[c] = sc; ssh[c] = sh;
            if (blockIdx.x == 0) {
                if (fin.mean) fin.mean[c] = (float)m;
                if (fin.invstd) fin.invstd[c] = is;
                if (fin.scale) fin.scale[c] = sc;
                if (fin.shift) fin.shift[c] = sh;
                if (fin.running_mean) fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)m;
                if (fin.running_var) {
                    const double unb = fin.M > 1 ? var * (double)fin.M / (double)(fin.M - 1) : var;
                    fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unb;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < C; i += BN_THREADS) { ssc[i] = scale[i]; ssh[i] = shift[i]; }
    }
    __syncthreads();
    const int64_t nvec = total >> 2;
    const int64_t stride = (int64_t)gridDim.x * BN_THREADS;
    for (int64_t k = (int64_t)blockIdx.x * BN_THREADS + threadIdx.x; k < nvec; k += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t kk = k + u * stride;
            if (kk < nvec) v[u] = ld_stream(y + (nvec - 1 - kk) * 4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t kk = k + u * stride;
            if (kk >= nvec) continue;
            const int64_t i = nvec - 1 - kk;
            const int c = (int)((i * 4) % C);
            float4 w = v[u];
            w.x = fmaf(w.x, ssc[c], ssh[c]); w.y = fmaf(w.y, ssc[c + 1], ssh[c + 1]);
            w.z = fmaf(w.z, ssc[c + 2], ssh[c + 2]); w.w = fmaf(w.w, ssc[c + 3], ssh[c + 3]);
            if (relu) { w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f); }
            if (PLANES && pl.kind((i * 4) / C) == 0) w = make_float4(0.f, 0.f, 0.f, 0.f);   // beyond the global volume
            const int64_t o = om.active ? om.offset((i * 4) / C, C) + c : i * 4;            // (depth-to-space) destination
            if (out) *reinterpret_cast<float4*>(out + o) = w;
            if (out16) {
                __half2 a = sat_half2(w.x * f16_scale, w.y * f16_scale);
                __half2 b = sat_half2(w.z * f16_scale, w.w * f16_scale);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&a);
                pk.y = *reinterpret_cast<uint32_t*>(&b);
                *reinterpret_cast<uint2*>(out16 + o) = pk;
            }
        }
    }
}

__global__ void bn_apply_scalar_kernel(const float* __restrict__ y, int64_t total, int C,
                                       const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                       float* __restrict__ out, __half* __restrict__ out16, float f16_scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float v = fmaf(y[i], scale[c], shift[c]);
        if (relu) v = fmaxf(v, 0.f);
        if (out) out[i] = v;
        if (out16) out16[i] = __float2half_rn(v * f16_scale);
    }
}

// ---- backward -------------------------------------------------------------------------------------------
// pass 1: red[c] += sum dz, red[C+c] += sum dz*xhat, where z = xhat*gamma+beta, dz = dout*(z>0)
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const float* __restrict__ y,
                                                                   const float* __restrict__ dout, int64_t M, int C,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ invstd,
                                                                   double* __restrict__ red, long long* __restrict__ mx) {
    const int c = blockIdx.y;      // one channel per blockIdx.y (generic in C; strided reads hit L2 lines shared by
    const float mu = mean[c], is = invstd[c], ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;  // neighbours)
    double s = 0, q = 0;
    float fs = 0.f, fq = 0.f, mdz = 0.f, mxh = 0.f;
    int cnt = 0;
    for (int64_t r = (int64_t)blockIdx.x * BN_THREADS + threadIdx.x; r < M; r += (int64_t)gridDim.x * BN_THREADS) {
        const float xh = (y[r * C + c] - mu) * is;
        const float z = fmaf(xh, ga, be);
        const float dz = z > 0.f ? dout[r * C + c] : 0.f;
        mdz = fmaxf(mdz, fabsf(dz)); mxh = fmaxf(mxh, fabsf(xh));
        fs += dz; fq = fmaf(dz, xh, fq);
        if (++cnt == 64) { s += fs; q += fq; fs = 0.f; fq = 0.f; cnt = 0; }
    }
    s += fs; q += fq;
    __shared__ double sh[2][BN_THREADS];
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = q;
    __syncthreads();
    for (int st = BN_THREADS / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) { sh[0][threadIdx.x] += sh[0][threadIdx.x + st]; sh[1][threadIdx.x] += sh[1][threadIdx.x + st]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(red + c, sh[0][0]); atomicAdd(red + C + c, sh[1][0]); }
#pragma unroll
    for (int st = 16; st > 0; st >>= 1) {
        mdz = fmaxf(mdz, __shfl_xor_sync(0xffffffffu, mdz, st));
        mxh = fmaxf(mxh, __shfl_xor_sync(0xffffffffu, mxh, st));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(mx + c, __double_as_longlong((double)mdz));
        atomicMax(mx + C + c, __double_as_longlong((double)mxh));
    }
}

// vectorised pass 1 for C % 4 == 0: thread t owns the 4 channels of float4 lane t % (C/4).  Four rows (8 independent 16-byte
// loads) in flight per thread at <= 85 registers, i.e. three blocks per SM: ~96 KB in flight per SM.  A thread sums a few
// dozen values per channel, so fp32 partials are exact enough; the fold to fp64 happens once per block.
template <bool PLANES>
__global__ void __launch_bounds__(BN_THREADS, 3) bn_bwd_reduce_vec_kernel(const float* __restrict__ y,
                                                                          const float* __restrict__ dout, int64_t M,
                                                                          int C, const float* __restrict__ gamma,
                                                                          const float* __restrict__ beta,
                                                                          const float* __restrict__ mean,
                                                                          const float* __restrict__ invstd,
                                                                          double* __restrict__ red, long long* __restrict__ mx,
                                                                          Planes pl, PeerPush pp, RowMap dm) {
    const int vpr = C >> 2;
    const int rows_per_iter = BN_THREADS / vpr;
    const int lane_v = threadIdx.x % vpr, lane_r = threadIdx.x / vpr;
    float mu[4], is[4], ga[4], be[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = lane_v * 4 + j;
        mu[j] = mean[c]; is[j] = invstd[c];
        ga[j] = gamma ? gamma[c] : 1.f; be[j] = beta ? beta[c] : 0.f;
    }
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0}, mdz[4] = {0, 0, 0, 0}, mxh[4] = {0, 0, 0, 0};
    const int64_t stride = (int64_t)gridDim.x * rows_per_iter;
    for (int64_t r = (int64_t)blockIdx.x * rows_per_iter + lane_r; r < M; r += 4 * stride) {
        float4 yv[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t rr = r + u * stride;
            if (rr < M) {
                yv[u] = ld_stream(y + rr * C + lane_v * 4);
                dv[u] = ld_stream(dout + (dm.active ? dm.offset(rr, C) : rr * C) + lane_v * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (r + u * stride < M && (!PLANES || pl.kind(r + u * stride) != 0)) {
                const float ya[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w}, da[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xh = (ya[j] - mu[j]) * is[j];
                    const float dz = fmaf(xh, ga[j], be[j]) > 0.f ? da[j] : 0.f;
                    mdz[j] = fmaxf(mdz[j], fabsf(dz)); mxh[j] = fmaxf(mxh[j], fabsf(xh));
                    s[j] += dz; q[j] = fmaf(dz, xh, q[j]);
                }
            }
        }
    }
    double ds[4], dq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { ds[j] = s[j]; dq[j] = q[j]; }
    __shared__ double sh[8 * BN_THREADS];
    __shared__ int smx[2 * BN_MAXC];
    for (int i = threadIdx.x; i < 2 * C; i += BN_THREADS) smx[i] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        atomicMax(&smx[lane_v * 4 + j], __float_as_int(mdz[j]));
        atomicMax(&smx[C + lane_v * 4 + j], __float_as_int(mxh[j]));
    }
    block_fold_and_flush(ds, dq, C, vpr, red, sh);     // contains the __syncthreads that also orders smx
    for (int i = threadIdx.x; i < 2 * C; i += BN_THREADS)
        atomicMax(mx + i, __double_as_longlong((double)__int_as_float(smx[i])));
    // D-sharded slabs: the last block broadcasts {sum dz, sum dz*xhat, max|dz|, max|xhat|}[C] (red and mx are contiguous)
    if (pp.n > 0) push_vector_from_last_block(red, 4 * C, pp);
}

// (optional) gather of the D-sharded partial sums, then the power-of-two fp16 scale for dy from the per-channel bound
//   |dy_c| <= |gamma_c*invstd_c| * (max|dz|_c + |sum_dz_c|/M + max|xhat|_c * |sum_dzxhat_c|/M)
// With `g`: thread 0 waits for every rank's 4*C-double vector, the block sums the slots in rank order INTO the workspace
// (red, mx) -- the sum of the ranks' maxima bounds the global maximum, and every rank derives the same scale.
__global__ void bn_bwd_scale_kernel(double* __restrict__ red, double* __restrict__ mx, long long M, int C,
                                    const float* __restrict__ gamma, const float* __restrict__ invstd, float target,
                                    float* __restrict__ scale2, PeerGather g, int* error_flag) {
    if (g.on()) {
        uint32_t tgt = 0;
        if (!gather_wait(g, tgt) && threadIdx.x == 0) atomicExch(error_flag, 44);
        const double* sl = reinterpret_cast<const double*>(g.slots);
        for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
            double a = 0.0;
            for (int r = 0; r < g.world; ++r) a += sl[(size_t)r * 4 * C + i];
            red[i] = a;                                   // red[0..2C) then mx[0..2C): one contiguous vector
        }
        __syncthreads();
        if (threadIdx.x == 0) *g.expect = tgt;
    }
    if (scale2 == nullptr) return;
    float b = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float ga = gamma ? gamma[c] : 1.f;
        const float a = fabsf((float)(red[c] / (double)M)), bb = fabsf((float)(red[C + c] / (double)M));
        b = fmaxf(b, fabsf(ga * invstd[c]) * ((float)mx[c] + a + (float)mx[C + c] * bb));
    }
    __shared__ float sh[256];
    sh[threadIdx.x] = b;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) sh[threadIdx.x] = fmaxf(sh[threadIdx.x], sh[threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float sc = 1.f;
        const float bound = sh[0];
        if (bound > 0.f && isfinite(bound)) {
            int e = (int)floorf(log2f(target / bound));
            e = max(-60, min(60, e));
            sc = exp2f((float)e);
        }
        scale2[0] = sc;
        scale2[1] = 1.f / sc;
    }
}

// pass 2, vectorised (C % 4 == 0): float4 in, float4 and/or 4 x fp16 out.  Walks the tensors BACKWARDS: pass 1 has just
// streamed them front to back, so their tails are what the 126 MB L2 still holds.  Two float4 of each tensor (4 loads) in
// flight per thread.
template <bool PLANES>
__global__ void __launch_bounds__(BN_THREADS, 4) bn_bwd_apply_vec_kernel(const float* __restrict__ y,
                                                                         const float* __restrict__ dout, int64_t M, int C,
                                                                         const float* __restrict__ gamma,
                                                                         const float* __restrict__ beta,
                                                                         const float* __restrict__ mean,
                                                                         const float* __restrict__ invstd,
                                                                         const double* __restrict__ red,
                                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                         float* __restrict__ dy, __half* __restrict__ dy16,
                                                                         float* __restrict__ scale2, Planes pl,
                                                                         long long m_div, HaloPush hp,
                                                                         const double* __restrict__ mx_own, RowMap dm) {
    // mx_own != nullptr (r2w): the power-of-two fp16 scale of dy is derived HERE, by every block, from the per-channel
    // bound of bn_bwd_scale_kernel (same expression, same bits); block 0 publishes {scale, 1/scale} for K3 / K4.  Saves the
    // one-block scale kernel and its boundary on the critical path of every training backward.
    __shared__ float smu[BN_MAXC], sis[BN_MAXC], sga[BN_MAXC], sbe[BN_MAXC], sa[BN_MAXC], sb[BN_MAXC];
    __shared__ float s_bound[BN_THREADS];
    float bnd = 0.f;
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        smu[c] = mean[c]; sis[c] = invstd[c];
        sga[c] = gamma ? gamma[c] : 1.f; sbe[c] = beta ? beta[c] : 0.f;
        sa[c] = (float)(red[c] / (double)m_div);
        sb[c] = (float)(red[C + c] / (double)m_div);
        if (mx_own != nullptr)
            bnd = fmaxf(bnd, fabsf(sga[c] * sis[c]) * ((float)mx_own[c] + fabsf(sa[c]) + (float)mx_own[C + c] * fabsf(sb[c])));
        if (blockIdx.x == 0) {
            if (dbeta) dbeta[c] = (float)red[c];
            if (dgamma) dgamma[c] = (float)red[C + c];
        }
    }
    float f16_scale = 1.f;
    if (mx_own != nullptr) {
        s_bound[threadIdx.x] = bnd;
        __syncthreads();
        for (int st = BN_THREADS / 2; st > 0; st >>= 1) {
            if (threadIdx.x < st) s_bound[threadIdx.x] = fmaxf(s_bound[threadIdx.x], s_bound[threadIdx.x + st]);
            __syncthreads();
        }
        const float bound = s_bound[0];
        if (bound > 0.f && isfinite(bound)) {
            int e = (int)floorf(log2f(8192.f / bound));
            e = max(-60, min(60, e));
            f16_scale = exp2f((float)e);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { scale2[0] = f16_scale; scale2[1] = 1.f / f16_scale; }
    } else {
        if (dy16 != nullptr && scale2 != nullptr) f16_scale = scale2[0];
        __syncthreads();
    }
    const int64_t nvec = (M * C) >> 2;
    const int64_t stride = (int64_t)gridDim.x * BN_THREADS;
    for (int64_t k = (int64_t)blockIdx.x * BN_THREADS + threadIdx.x; k < nvec; k += 2 * stride) {
        float4 yv[2], dv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t kk = k + u * stride;
            if (kk < nvec) {
                const int64_t iv = nvec - 1 - kk;
                yv[u] = ld_stream(y + iv * 4);
                dv[u] = ld_stream(dout + (dm.active ? dm.offset((iv * 4) / C, C) + (iv * 4) % C : iv * 4));
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t kk = k + u * stride;
            if (kk >= nvec) continue;
            const int64_t i = nvec - 1 - kk;
            const int c = (int)((i * 4) % C);
            const float ya[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w}, da[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
            float v[4];
            const int kind = PLANES ? pl.kind((i * 4) / C) : 2;   // 2 owned: full formula; 1 halo copy: no mean terms; 0: zero
            const float own = kind == 2 ? 1.f : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xh = (ya[j] - smu[c + j]) * sis[c + j];
                const float dz = (kind != 0 && fmaf(xh, sga[c + j], sbe[c + j]) > 0.f) ? da[j] : 0.f;
                v[j] = sga[c + j] * sis[c + j] * (dz - own * (sa[c + j] + xh * sb[c + j]));
            }
            if (dy) {
                const float4 o4 = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(dy + i * 4) = o4;
                if (hp.on() && dy16 == nullptr) halo_store(hp, i * 16, nvec * 16, o4);
            }
            if (dy16) {
                __half2 a = sat_half2(v[0] * f16_scale, v[1] * f16_scale);
                __half2 b = sat_half2(v[2] * f16_scale, v[3] * f16_scale);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&a);
                pk.y = *reinterpret_cast<uint32_t*>(&b);
                *reinterpret_cast<uint2*>(dy16 + i * 4) = pk;
                if (hp.on()) halo_store(hp, i * 8, nvec * 8, pk);
            }
        }
    }
    if (hp.on()) halo_finish(hp);       // the neighbours' counters move once every block's boundary stores are out
}

// pass 2: dy = gamma*invstd*(dz - sum_dz/M - xhat*sum_dzxhat/M); also emits dgamma/dbeta (block 0)
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const float* __restrict__ y,
                                                                  const float* __restrict__ dout, int64_t M, int C,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd,
                                                                  const double* __restrict__ red,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                  float* __restrict__ dy, __half* __restrict__ dy16,
                                                                  const float* __restrict__ scale2) {
    const float f16_scale = (dy16 != nullptr && scale2 != nullptr) ? scale2[0] : 1.f;
    __shared__ float smu[BN_MAXC], sis[BN_MAXC], sga[BN_MAXC], sbe[BN_MAXC], sa[BN_MAXC], sb[BN_MAXC];
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        smu[c] = mean[c]; sis[c] = invstd[c];
        sga[c] = gamma ? gamma[c] : 1.f; sbe[c] = beta ? beta[c] : 0.f;
        sa[c] = (float)(red[c] / (double)M);
        sb[c] = (float)(red[C + c] / (double)M);
        if (blockIdx.x == 0) {
            if (dbeta) dbeta[c] = (float)red[c];
            if (dgamma) dgamma[c] = (float)red[C + c];
        }
    }
    __syncthreads();
    const int64_t total = M * C;
    for (int64_t i = (int64_t)blockIdx.x * BN_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * BN_THREADS) {
        const int c = (int)(i % C);
        const float xh = (y[i] - smu[c]) * sis[c];
        const float dz = fmaf(xh, sga[c], sbe[c]) > 0.f ? dout[i] : 0.f;
        const float v = sga[c] * sis[c] * (dz - sa[c] - xh * sb[c]);
        if (dy) dy[i] = v;
        if (dy16) dy16[i] = __float2half_rn(v * f16_scale);
    }
}

// ---- operand staging ---------------------------------------------------------------------------------------
// fp32 -> fp16, round to nearest even, SATURATED to +-65504 (never inf: an out-of-range activation clips instead of
// poisoning the accumulators).  Four 16-byte loads in flight per thread.
// With a halo push (D-sharded slabs) the traversal is ROTATED so that it starts at the upper boundary region and wraps
// into the lower one: both regions are written -- locally and into the neighbours' halo planes -- in the first pass of the
// grid-stride loop, and every block takes its ticket as soon as its share of them is done.  The neighbours' counters then
// move after ~a quarter of the kernel instead of at its end (r2v: the conv that follows marches from the lower halo plane
// upwards, so it needs the halo at once; the rest of this kernel now hides the NVLink flight time and the rank skew).
__global__ void __launch_bounds__(256) cast_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst,
                                                       int64_t n, float scale, const float* __restrict__ scale_dev,
                                                       HaloPush hp) {
    if (scale_dev != nullptr) scale *= *scale_dev;
    const int64_t nvec = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * 256;
    const bool push = hp.on();
    const int64_t bvec = push ? (hp.bytes >> 3) : 0;               // 8-byte output vectors per boundary region
    const int64_t shift = push ? nvec - bvec : 0;
    const int64_t early = 2 * bvec;                                // rotated indices below this touch a boundary region
    bool signalled = false;
    for (int64_t j0 = (int64_t)blockIdx.x * 256; j0 < nvec; j0 += 4 * stride) {          // block-uniform trip count
        if (push && !signalled && j0 >= early) {
            halo_finish(hp);
            signalled = true;
        }
        const int64_t j = j0 + threadIdx.x;
        float4 v[4];
        int64_t idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int64_t i = j + u * stride + shift;
            if (i >= nvec) i -= nvec;
            idx[u] = j + u * stride < nvec ? i : -1;
            if (idx[u] >= 0) v[u] = ld_stream(src + idx[u] * 4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (idx[u] < 0) continue;
            __half2 a = sat_half2(v[u].x * scale, v[u].y * scale), b = sat_half2(v[u].z * scale, v[u].w * scale);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&a);
            pk.y = *reinterpret_cast<uint32_t*>(&b);
            *reinterpret_cast<uint2*>(dst + idx[u] * 4) = pk;
            if (push) halo_store(hp, idx[u] * 8, nvec * 8, pk);                    // host guarantees n % 4 == 0 with a push
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (nvec << 2) + threadIdx.x;
        dst[i] = __float2half_rn(fminf(fmaxf(src[i] * scale, -65504.f), 65504.f));
    }
    if (push && !signalled) halo_finish(hp);
}

// fp32 [rows][c] -> fp16 [rows][c_pad] with zero channels appended (the U-Net stem: c = 1 -> 32), saturated like
// cast_f16_kernel.  One thread per 16-byte output chunk (8 channels): the stores of a warp are 512 contiguous bytes.
// Replaces cast + torch.zeros + strided copy (three passes, 283 us of a batch-4 train step in r2q) by one.
__global__ void __launch_bounds__(256) cast_f16_pad_kernel(const float* __restrict__ src, __half* __restrict__ dst,
                                                           int64_t rows, int c, int c_pad) {
    const int chunks = c_pad >> 3;
    const int64_t total = rows * chunks;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t row = i / chunks;
        const int c0 = (int)(i - row * chunks) * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = c0 + e < c ? src[row * c + c0 + e] : 0.f;
        __half2 h0 = sat_half2(v[0], v[1]), h1 = sat_half2(v[2], v[3]), h2 = sat_half2(v[4], v[5]), h3 = sat_half2(v[6], v[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + i * 8) = pk;
    }
}

// fp32 [rows][ca] ++ fp32 [rows][cb] -> fp16 [rows][ca + cb]: the decoder's skip concatenation (torch.cat((x_skip, x), 1),
// RepMode.py:106) folded into the operand staging of the conv that consumes it -- the concatenated fp32 tensor is never
// written.  One thread per 4 output channels (ca, cb multiples of 4), four loads in flight, saturated like cast_f16_kernel.
__global__ void __launch_bounds__(256) cast_f16_cat_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b,
                                                           int cb, __half* __restrict__ dst, int64_t rows) {
    const int cv = (ca + cb) >> 2, cav = ca >> 2;
    const int64_t nvec = rows * cv;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t k = i + u * stride;
            if (k < nvec) {
                const int64_t row = k / cv;
                const int col = (int)(k - row * cv);
                v[u] = col < cav ? ld_stream(a + (row * cav + col) * 4) : ld_stream(b + (row * (cv - cav) + (col - cav)) * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t k = i + u * stride;
            if (k >= nvec) continue;
            __half2 lo = sat_half2(v[u].x, v[u].y), hi = sat_half2(v[u].z, v[u].w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(dst + k * 4) = pk;
        }
    }
}

__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ src, int64_t n, float* amax) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        m = fmaxf(m, fabsf(src[i]));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));   // m >= 0
}

struct AmaxList { const float* p[8]; long long n[8]; };
// blockIdx.y selects the tensor: one launch covers the five expert tensors of a MoDEConv layer
__global__ void __launch_bounds__(256) amax_multi_kernel(AmaxList L, float* amax) {
    const float* src = L.p[blockIdx.y];
    const long long n = L.n[blockIdx.y];
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        m = fmaxf(m, fabsf(src[i]));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));
}

__global__ void f16_scale_kernel(const float* __restrict__ amax, float target, float* __restrict__ scale2) {
    float sc = 1.f;
    const float a = amax[0];
    if (a > 0.f && isfinite(a)) {
        int e = (int)floorf(log2f(target / a));
        e = max(-60, min(60, e));
        sc = exp2f((float)e);
    }
    scale2[0] = sc;
    scale2[1] = 1.f / sc;
}

// Grid of a grid-stride streaming kernel: enough blocks for the work, at most 8 per SM, and a whole number of blocks per
// SM (a ragged last wave -- e.g. 1024 blocks on 148 SMs -- costs up to half the kernel's time).
// Grid of exactly ONE resident wave of a grid-stride kernel: occupancy x SM count blocks (or fewer when the work is small).
// r2d capture: with 8 blocks per SM on kernels whose registers allow 3-4, the launch runs as two waves of ~10 us each and the
// drain / refill between them costs these 20-40 us streaming kernels a fifth of their time (DRAM throughput 40-50 % of peak
// while a plain copy reaches 82 %).
template <typename K>
static int wave_grid(K kernel, int threads, int64_t want_blocks) {
    static std::mutex mu;
    static std::map<const void*, int> cache;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find((const void*)kernel);
        if (it != cache.end()) per_sm = it->second;
    }
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
        std::lock_guard<std::mutex> lock(mu);
        cache[(const void*)kernel] = per_sm;
    }
    return (int)max((int64_t)1, min(want_blocks, (int64_t)per_sm * sm_count()));
}

static int stream_grid(int64_t work_items, int per_block) {
    const int64_t want = ceil_div(work_items, per_block);
    const int64_t sms = sm_count();
    int64_t g = min(want, sms * 8);
    if (g > sms) g -= g % sms;
    return (int)max((int64_t)1, g);
}

}  // namespace mode

using namespace mode;

extern "C" int mode_bn_stats(const float* y, int64_t M, int32_t C, double* sums, void* stream) {
    if (!y || !sums || M <= 0 || C <= 0 || C > BN_MAXC) MODE_FAIL("mode_bn_stats: bad arguments (C=%d)", C);
    cudaStream_t st = (cudaStream_t)stream;
    const int vpr = C >> 2;
    if ((C & 3) == 0 && vpr <= BN_THREADS && BN_THREADS % vpr == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
        const int rpi = BN_THREADS / vpr;
        bn_stats_kernel<<<stream_grid(M, rpi * 8), BN_THREADS, 0, st>>>(y, M, C, sums);
    } else {
        bn_stats_scalar_kernel<<<C, BN_THREADS, 0, st>>>(y, M, C, sums);
    }
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_bn_finalize_ex(const double* sums, int64_t M, int32_t C, const float* gamma, const float* beta,
                                   float eps, float momentum, float* mean, float* invstd, float* scale, float* shift,
                                   float* running_mean, float* running_var, const mode_peer_gather_t* gather,
                                   void* stream) {
    if ((!sums && !gather) || !scale || !shift || M <= 0 || C <= 0) MODE_FAIL("mode_bn_finalize: bad arguments");
    const PeerGather g = to_peer_gather(gather);
    if (g.on() && (g.world <= 0 || !g.signal || !g.expect)) MODE_FAIL("mode_bn_finalize: incomplete gather descriptor");
    int* ef = device_error_flag();
    if (!ef) MODE_FAIL("mode_bn_finalize: could not allocate the device error flag");
    const unsigned grid = g.on() ? 1u : (unsigned)ceil_div(C, 128);       // the gathering form waits in ONE block
    bn_finalize_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(sums, M, C, gamma, beta, eps, momentum, mean, invstd, scale,
                                                               shift, running_mean, running_var, g, ef);
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_bn_finalize(const double* sums, int64_t M, int32_t C, const float* gamma, const float* beta,
                                float eps, float momentum, float* mean, float* invstd, float* scale, float* shift,
                                float* running_mean, float* running_var, void* stream) {
    return mode_bn_finalize_ex(sums, M, C, gamma, beta, eps, momentum, mean, invstd, scale, shift, running_mean,
                               running_var, nullptr, stream);
}

static int bn_apply_launch(const float* y, int64_t M, int32_t C, const float* scale, const float* shift, int32_t relu,
                           float* out, void* out_f16, float f16_scale, const mode_planes_t* planes, const BnFinalize& fin,
                           const mode_rowmap_t* out_map, void* stream) {
    const RowMap om = to_rowmap(out_map, C);
    if (om.active && (om.pitch & 3)) MODE_FAIL("mode_bn_apply_relu: mapped output needs pitch %% 4 == 0");
    if (om.d2s && ((M & 7) || M / 8 % ((int64_t)om.D * om.H * om.W) != 0))
        MODE_FAIL("mode_bn_apply_relu: depth-to-space map does not match M");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = M * C;
    const bool aligned = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(out_f16) & 7) == 0;
    if ((C & 3) == 0 && aligned) {
        const int64_t want = ceil_div(total / 4, BN_THREADS * 4);
        if (planes && !planes_trivial(planes))
            bn_apply_kernel<true><<<wave_grid(bn_apply_kernel<true>, BN_THREADS, want), BN_THREADS, 0, st>>>(y, total, C, scale, shift, relu, out, (__half*)out_f16,
                                                               f16_scale, to_planes(planes), fin, om);
        else
            bn_apply_kernel<false><<<wave_grid(bn_apply_kernel<false>, BN_THREADS, want), BN_THREADS, 0, st>>>(y, total, C, scale, shift, relu, out, (__half*)out_f16,
                                                                f16_scale, Planes{}, fin, om);
    } else {
        if (planes) MODE_FAIL("mode_bn_apply_relu: plane ranges need C %% 4 == 0 and 16-byte aligned tensors");
        if (om.active) MODE_FAIL("mode_bn_apply_relu: a mapped output needs C %% 4 == 0 and 16-byte aligned tensors");
        if (fin.sums != nullptr) {                 // scalar layout: finalize as its own launch
            bn_finalize_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, st>>>(fin.sums, fin.M, C, fin.gamma, fin.beta, fin.eps,
                                                                           fin.momentum, fin.mean, fin.invstd, fin.scale,
                                                                           fin.shift, fin.running_mean, fin.running_var,
                                                                           PeerGather{nullptr, 0, nullptr, nullptr},
                                                                           device_error_flag());
            MODE_LAUNCH_CHECK();
            scale = fin.scale; shift = fin.shift;
        }
        bn_apply_scalar_kernel<<<stream_grid(total, BN_THREADS * 4), BN_THREADS, 0, st>>>(
            y, total, C, scale, shift, relu, out, (__half*)out_f16, f16_scale);
    }
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_bn_apply_relu(const float* y, int64_t M, int32_t C, const float* scale, const float* shift,
                                  int32_t relu, float* out, void* out_f16, float f16_scale,
                                  const mode_planes_t* planes, void* stream) {
    if (!y || !scale || !shift || (!out && !out_f16) || M <= 0 || C <= 0 || C > BN_MAXC)
        MODE_FAIL("mode_bn_apply_relu: bad arguments (C=%d)", C);
    return bn_apply_launch(y, M, C, scale, shift, relu, out, out_f16, f16_scale, planes, BnFinalize{}, nullptr, stream);
}

extern "C" int mode_bn_finalize_apply_relu(const double* sums, int64_t M_stat, int32_t C, const float* gamma,
                                           const float* beta, float eps, float momentum, float* mean, float* invstd,
                                           float* scale, float* shift, float* running_mean, float* running_var,
                                           const float* y, int64_t M, int32_t relu, float* out, void* out_f16,
                                           float f16_scale, const mode_planes_t* planes, const mode_rowmap_t* out_map,
                                           void* stream) {
    if (!sums || !scale || !shift || !y || (!out && !out_f16) || M <= 0 || M_stat <= 0 || C <= 0 || C > BN_MAXC)
        MODE_FAIL("mode_bn_finalize_apply_relu: bad arguments (C=%d)", C);
    BnFinalize fin{sums, (long long)M_stat, gamma, beta, eps, momentum, mean, invstd, scale, shift, running_mean, running_var};
    return bn_apply_launch(y, M, C, scale, shift, relu, out, out_f16, f16_scale, planes, fin, out_map, stream);
}

// {sum dz, sum dz*xhat}[C] then {max |dz|, max |xhat|}[C], ALL doubles: one contiguous fp64 vector a D-sharded caller can
// all-reduce (sum) in one step -- the sum of the ranks' maxima bounds the global maximum, which is all the fp16 scale needs
extern "C" int64_t mode_bn_bwd_workspace_bytes(int32_t C) { return (int64_t)C * 4 * sizeof(double); }

static int bn_bwd_reduce_launch(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                                const float* beta, const float* mean, const float* invstd,
                                const mode_planes_t* planes, void* workspace_v, const mode_peer_push_t* push,
                                bool workspace_is_zero, const mode_rowmap_t* dout_map, void* stream) {
    if (!y || !dout || !mean || !invstd || !workspace_v || M <= 0 || C <= 0 || C > BN_MAXC)
        MODE_FAIL("mode_bn_relu_bwd_reduce: bad arguments (C=%d)", C);
    const RowMap dm = to_rowmap(dout_map, C);
    if (dm.active && (dm.pitch & 3)) MODE_FAIL("mode_bn_relu_bwd_reduce: mapped dout needs pitch %% 4 == 0");
    const PeerPush pp = to_peer_push(push);
    if (pp.n < 0 || pp.n > 8 || (pp.n > 0 && !pp.ticket)) MODE_FAIL("mode_bn_relu_bwd_reduce: bad push descriptor");
    cudaStream_t st = (cudaStream_t)stream;
    double* workspace = (double*)workspace_v;
    long long* mx = (long long*)(workspace + 2 * (size_t)C);
    if (!workspace_is_zero) MODE_CUDA(cudaMemsetAsync(workspace_v, 0, (size_t)mode_bn_bwd_workspace_bytes(C), st));
    const int vpr = C >> 2;
    const bool aligned = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dout)) & 15) == 0;
    if ((C & 3) == 0 && vpr <= BN_THREADS && BN_THREADS % vpr == 0 && aligned) {
        const int rpi = BN_THREADS / vpr;
        const int64_t want = ceil_div(M, rpi * 4);
        if (planes && !planes_trivial(planes))
            bn_bwd_reduce_vec_kernel<true><<<wave_grid(bn_bwd_reduce_vec_kernel<true>, BN_THREADS, want), BN_THREADS, 0, st>>>(y, dout, M, C, gamma, beta, mean, invstd,
                                                                        workspace, mx, to_planes(planes), pp, dm);
        else
            bn_bwd_reduce_vec_kernel<false><<<wave_grid(bn_bwd_reduce_vec_kernel<false>, BN_THREADS, want), BN_THREADS, 0, st>>>(y, dout, M, C, gamma, beta, mean, invstd,
                                                                         workspace, mx, Planes{}, pp, dm);
    } else {
        if (planes) MODE_FAIL("mode_bn_relu_bwd_reduce: plane ranges need C %% 4 == 0 and 16-byte aligned tensors");
        if (dm.active) MODE_FAIL("mode_bn_relu_bwd_reduce: a mapped dout needs C %% 4 == 0 and 16-byte aligned tensors");
        if (pp.n > 0) MODE_FAIL("mode_bn_relu_bwd_reduce: the fused push needs C %% 4 == 0 and 16-byte aligned tensors");
        const int gx = (int)max((int64_t)1, min(ceil_div(M, BN_THREADS * 8), (int64_t)64));
        bn_bwd_reduce_kernel<<<dim3(gx, C), BN_THREADS, 0, st>>>(y, dout, M, C, gamma, beta, mean, invstd, workspace, mx);
    }
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_bn_relu_bwd_reduce_ex(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                                          const float* beta, const float* mean, const float* invstd,
                                          const mode_planes_t* planes, void* workspace_v, const mode_peer_push_t* push,
                                          void* stream) {
    return bn_bwd_reduce_launch(y, dout, M, C, gamma, beta, mean, invstd, planes, workspace_v, push, false, nullptr, stream);
}

extern "C" int mode_bn_relu_bwd_reduce(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                                       const float* beta, const float* mean, const float* invstd,
                                       const mode_planes_t* planes, void* workspace_v, void* stream) {
    return bn_bwd_reduce_launch(y, dout, M, C, gamma, beta, mean, invstd, planes, workspace_v, nullptr, false, nullptr, stream);
}

extern "C" int mode_bn_relu_bwd_reduce_v2(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                                          const float* beta, const float* mean, const float* invstd,
                                          const mode_planes_t* planes, void* workspace_v, int32_t workspace_is_zero,
                                          const mode_rowmap_t* dout_map, void* stream) {
    return bn_bwd_reduce_launch(y, dout, M, C, gamma, beta, mean, invstd, planes, workspace_v, nullptr,
                                workspace_is_zero != 0, dout_map, stream);
}

static int bn_bwd_apply_launch(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                               const float* beta, const float* mean, const float* invstd, float* dgamma,
                               float* dbeta, float* dy, void* dy_f16, float* dy_scale2,
                               const mode_planes_t* planes, void* workspace_v,
                               const mode_peer_gather_t* gather, const mode_halo_push_t* halo,
                               const mode_rowmap_t* dout_map, void* stream) {
    if (!y || !dout || !mean || !invstd || !workspace_v || (!dy && !dy_f16) || M <= 0 || C <= 0 || C > BN_MAXC)
        MODE_FAIL("mode_bn_relu_bwd_apply: bad arguments (C=%d)", C);
    const RowMap dm = to_rowmap(dout_map, C);
    if (dm.active && (dm.pitch & 3)) MODE_FAIL("mode_bn_relu_bwd_apply: mapped dout needs pitch %% 4 == 0");
    if (dy_f16 && !dy_scale2) MODE_FAIL("mode_bn_relu_bwd_apply: dy_f16 needs dy_scale2");
    cudaStream_t st = (cudaStream_t)stream;
    double* workspace = (double*)workspace_v;
    double* mx = workspace + 2 * (size_t)C;
    const long long m_div = planes ? (long long)planes->m_global : (long long)M;
    const PeerGather g = to_peer_gather(gather);
    const HaloPush hp = to_halo_push(halo);
    if (g.on() && (g.world <= 0 || !g.signal || !g.expect)) MODE_FAIL("mode_bn_relu_bwd_apply: incomplete gather descriptor");
    if (hp.on() && (hp.bytes <= 0 || (hp.bytes & 15) || !hp.ticket)) MODE_FAIL("mode_bn_relu_bwd_apply: bad halo descriptor");
    const bool aligned = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dout)) & 15) == 0;
    const bool vec_ok = (C & 3) == 0 && aligned && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(dy_f16) & 7) == 0);
    // the apply kernel derives the fp16 scale itself unless the sums still have to be gathered from the other ranks
    const bool scale_in_kernel = vec_ok && dy_f16 != nullptr && !g.on() && getenv("REPMODE_BN_SCALE_KERNEL") == nullptr;
    const double* mx_own = scale_in_kernel ? mx : nullptr;
    if ((dy_f16 || g.on()) && !scale_in_kernel) {
        int* ef = device_error_flag();
        if (!ef) MODE_FAIL("mode_bn_relu_bwd_apply: could not allocate the device error flag");
        bn_bwd_scale_kernel<<<1, 256, 0, st>>>(workspace, mx, m_div, C, gamma, invstd, 8192.f, dy_f16 ? dy_scale2 : nullptr,
                                               g, ef);
        MODE_LAUNCH_CHECK();
    }
    if (vec_ok) {
        const int64_t want = ceil_div(M * C / 4, BN_THREADS * 2);
        if (planes && !planes_trivial(planes))
            bn_bwd_apply_vec_kernel<true><<<wave_grid(bn_bwd_apply_vec_kernel<true>, BN_THREADS, want), BN_THREADS, 0, st>>>(y, dout, M, C, gamma, beta, mean, invstd, workspace,
                                                                       dgamma, dbeta, dy, (__half*)dy_f16, dy_scale2,
                                                                       to_planes(planes), m_div, hp, mx_own, dm);
        else
            bn_bwd_apply_vec_kernel<false><<<wave_grid(bn_bwd_apply_vec_kernel<false>, BN_THREADS, want), BN_THREADS, 0, st>>>(y, dout, M, C, gamma, beta, mean, invstd,
                                                                        workspace, dgamma, dbeta, dy, (__half*)dy_f16,
                                                                        dy_scale2, Planes{}, m_div, hp, mx_own, dm);
    } else {
        if (planes) MODE_FAIL("mode_bn_relu_bwd_apply: plane ranges need C %% 4 == 0 and 16-byte aligned tensors");
        if (dm.active) MODE_FAIL("mode_bn_relu_bwd_apply: a mapped dout needs C %% 4 == 0 and 16-byte aligned tensors");
        if (hp.on()) MODE_FAIL("mode_bn_relu_bwd_apply: the fused halo push needs C %% 4 == 0 and 16-byte aligned tensors");
        bn_bwd_apply_kernel<<<stream_grid(M * C, BN_THREADS * 8), BN_THREADS, 0, st>>>(
            y, dout, M, C, gamma, beta, mean, invstd, workspace, dgamma, dbeta, dy, (__half*)dy_f16, dy_scale2);
    }
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_bn_relu_bwd_apply_ex(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                                         const float* beta, const float* mean, const float* invstd, float* dgamma,
                                         float* dbeta, float* dy, void* dy_f16, float* dy_scale2,
                                         const mode_planes_t* planes, void* workspace_v,
                                         const mode_peer_gather_t* gather, const mode_halo_push_t* halo, void* stream) {
    return bn_bwd_apply_launch(y, dout, M, C, gamma, beta, mean, invstd, dgamma, dbeta, dy, dy_f16, dy_scale2, planes,
                               workspace_v, gather, halo, nullptr, stream);
}

extern "C" int mode_bn_relu_bwd_apply_v2(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                                         const float* beta, const float* mean, const float* invstd, float* dgamma,
                                         float* dbeta, float* dy, void* dy_f16, float* dy_scale2,
                                         const mode_planes_t* planes, void* workspace_v, const mode_rowmap_t* dout_map,
                                         void* stream) {
    return bn_bwd_apply_launch(y, dout, M, C, gamma, beta, mean, invstd, dgamma, dbeta, dy, dy_f16, dy_scale2, planes,
                               workspace_v, nullptr, nullptr, dout_map, stream);
}

extern "C" int mode_bn_relu_bwd_apply(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                                      const float* beta, const float* mean, const float* invstd, float* dgamma,
                                      float* dbeta, float* dy, void* dy_f16, float* dy_scale2,
                                      const mode_planes_t* planes, void* workspace_v, void* stream) {
    return mode_bn_relu_bwd_apply_ex(y, dout, M, C, gamma, beta, mean, invstd, dgamma, dbeta, dy, dy_f16, dy_scale2, planes,
                                     workspace_v, nullptr, nullptr, stream);
}

extern "C" int mode_bn_relu_bwd(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                                const float* beta, const float* mean, const float* invstd, float* dgamma,
                                float* dbeta, float* dy, void* dy_f16, float* dy_scale2, void* workspace_v,
                                void* stream) {
    if (mode_bn_relu_bwd_reduce(y, dout, M, C, gamma, beta, mean, invstd, nullptr, workspace_v, stream) != 0) return -1;
    return mode_bn_relu_bwd_apply(y, dout, M, C, gamma, beta, mean, invstd, dgamma, dbeta, dy, dy_f16, dy_scale2,
                                  nullptr, workspace_v, stream);
}

extern "C" int mode_cast_f16_ex(const float* src, void* dst_f16, int64_t n, float scale, const float* scale_dev,
                                const mode_halo_push_t* halo, void* stream) {
    if (!src || !dst_f16 || n <= 0) MODE_FAIL("mode_cast_f16: bad arguments");
    if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst_f16) & 7))
        MODE_FAIL("mode_cast_f16: pointers must be 16-byte (src) / 8-byte (dst) aligned");
    const HaloPush hp = to_halo_push(halo);
    if (hp.on() && ((n & 3) || hp.bytes <= 0 || (hp.bytes & 15) || hp.bytes > n * 2 || !hp.ticket))
        MODE_FAIL("mode_cast_f16: bad halo descriptor (n %% 4 == 0, 0 < bytes <= tensor bytes, bytes %% 16 == 0)");
    cast_f16_kernel<<<wave_grid(cast_f16_kernel, 256, ceil_div(n / 4 + 1, 256 * 4)), 256, 0, (cudaStream_t)stream>>>(
        src, (__half*)dst_f16, n, scale, scale_dev, hp);
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_cast_f16(const float* src, void* dst_f16, int64_t n, float scale, const float* scale_dev,
                             void* stream) {
    return mode_cast_f16_ex(src, dst_f16, n, scale, scale_dev, nullptr, stream);
}

extern "C" int mode_cast_f16_pad(const float* src, void* dst_f16, int64_t rows, int32_t c, int32_t c_pad, void* stream) {
    if (!src || !dst_f16 || rows <= 0 || c <= 0 || c_pad < c || (c_pad & 7)) MODE_FAIL("mode_cast_f16_pad: bad arguments");
    if (reinterpret_cast<uintptr_t>(dst_f16) & 15) MODE_FAIL("mode_cast_f16_pad: dst must be 16-byte aligned");
    const int64_t total = rows * (c_pad >> 3);
    cast_f16_pad_kernel<<<wave_grid(cast_f16_pad_kernel, 256, ceil_div(total, 256)), 256, 0, (cudaStream_t)stream>>>(
        src, (__half*)dst_f16, rows, c, c_pad);
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_cast_f16_cat(const float* a, int32_t ca, const float* b, int32_t cb, void* dst_f16, int64_t rows,
                                 void* stream) {
    if (!a || !b || !dst_f16 || rows <= 0 || ca <= 0 || cb <= 0 || (ca & 3) || (cb & 3))
        MODE_FAIL("mode_cast_f16_cat: bad arguments (ca=%d, cb=%d: multiples of 4)", ca, cb);
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15 || (reinterpret_cast<uintptr_t>(dst_f16) & 7))
        MODE_FAIL("mode_cast_f16_cat: pointers must be 16-byte (sources) / 8-byte (dst) aligned");
    const int64_t nvec = rows * ((ca + cb) >> 2);
    cast_f16_cat_kernel<<<wave_grid(cast_f16_cat_kernel, 256, ceil_div(nvec, 256 * 4)), 256, 0, (cudaStream_t)stream>>>(
        a, ca, b, cb, (__half*)dst_f16, rows);
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_amax(const float* src, int64_t n, float* amax, void* stream) {
    if (!src || !amax || n <= 0) MODE_FAIL("mode_amax: bad arguments");
    amax_kernel<<<stream_grid(n, 256 * 8), 256, 0, (cudaStream_t)stream>>>(src, n, amax);
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_f16_scale(const float* amax, float target, float* scale2, void* stream) {
    if (!amax || !scale2 || !(target > 0.f)) MODE_FAIL("mode_f16_scale: bad arguments");
    f16_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(amax, target, scale2);
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_amax_multi(const float* const* srcs_host, const int64_t* counts_host, int32_t k, float* amax,
                               void* stream) {
    if (!srcs_host || !counts_host || !amax || k <= 0 || k > 8) MODE_FAIL("mode_amax_multi: bad arguments (k=%d)", k);
    AmaxList L;
    int64_t nmax = 0;
    for (int i = 0; i < 8; ++i) {
        L.p[i] = i < k ? srcs_host[i] : nullptr;
        L.n[i] = i < k ? counts_host[i] : 0;
        if (i < k && (!srcs_host[i] || counts_host[i] <= 0)) MODE_FAIL("mode_amax_multi: null / empty tensor %d", i);
        if (i < k) nmax = max(nmax, counts_host[i]);
    }
    amax_multi_kernel<<<dim3(stream_grid(nmax, 256 * 8), k), 256, 0, (cudaStream_t)stream>>>(L, amax);
    MODE_LAUNCH_CHECK();
    return 0;
}
