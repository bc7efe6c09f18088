// SIMT fp32 direct 5x5x5 convolution (forward / dgrad via the flipped pack) and wgrad.
//
// Role on the path: (1) the layers whose channel counts cannot feed a tensor-core tile -- the U-Net stem
// (Ci=1, RepMode.py:27) and head (Co=1, RepMode.py:42), 0.8 % of the FLOPs and HBM-bound -- and
// (2) the exact-fp32 cross-check of the tcgen05 kernels on the GPU.  Replaces F.conv3d(x[i:i+1], w[i],
// padding='same') (RepMode.py:204-210) and its autograd.
//
// Layouts: x/dy NDHWC fp32, weights = K1's packed fp32 layout w[u][tap][k_chunk][n][32].
#include "common.cuh"

namespace mode {

constexpr int TH = 4, TW = 32;           // output tile (rows x cols) of one (n, d) plane per block
constexpr int XR = TH + 4, XC = TW + 4;  // haloed tile
constexpr int XPLANE = XR * XC + 4;      // +4 floats: channel planes land on distinct banks
constexpr int KC8 = 8;                   // channels per stage

// grid (tiles_h*tiles_w, D, N*ceil(Nout/32)), block 256: thread = (2 voxels: w, w+16) x 8 output channels
__global__ void __launch_bounds__(256) conv3d_simt_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const int32_t* __restrict__ sample_u,
                                                          float* __restrict__ y, int D, int H, int W, int K, int Nout,
                                                          float out_scale, const float* __restrict__ out_scale_dev,
                                                          double* __restrict__ bn_sums, int stat_lo, int stat_hi,
                                                          ConvExt ext) {
    __shared__ float xs[KC8 * XPLANE];
    __shared__ __align__(16) float ws[25 * KC8 * 32];
    const int tiles_w = (W + TW - 1) / TW;
    const int th0 = (blockIdx.x / tiles_w) * TH, tw0 = (blockIdx.x % tiles_w) * TW;
    const int d = blockIdx.y;
    const int nob = (Nout + 31) / 32;
    const int n = blockIdx.z / nob, ob = blockIdx.z % nob;
    const int u = sample_u ? sample_u[n] : 0;
    const int tid = threadIdx.x;
    const int cg = tid >> 6, vh = (tid & 63) >> 4, vw = tid & 15;
    const int nkc = (K + 31) / 32;

    float acc0[8], acc1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

    const int nk8 = (K + KC8 - 1) / KC8;
    for (int k8 = 0; k8 < nk8; ++k8) {
        for (int kd = 0; kd < 5; ++kd) {
            const int dz = d + kd - 2 + ext.x_off;    // plane of the (possibly haloed) input tensor
            if (dz < 0 || dz >= ext.Dx) continue;     // zero padding plane (uniform across the block)
            __syncthreads();
            // haloed input tile, channel-planar in shared memory
            for (int idx = tid; idx < XR * XC * KC8; idx += 256) {
                const int ch = idx % KC8, v = idx / KC8;
                const int r = v / XC, cc = v % XC;
                const int hy = th0 + r - 2, wx = tw0 + cc - 2, k = k8 * KC8 + ch;
                float val = 0.f;
                if (hy >= 0 && hy < H && wx >= 0 && wx < W && k < K)
                    val = x[((((size_t)n * ext.Dx + dz) * H + hy) * W + wx) * K + k];
                xs[ch * XPLANE + r * XC + cc] = val;
            }
            // 25 taps x 8 k x 32 n weights of this (kd, k8)
            for (int idx = tid; idx < 25 * 32; idx += 256) {
                const int t = idx >> 5, co = idx & 31;
                const int nn = ob * 32 + co;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (nn < Nout) {
                    const float* src = w + ((((size_t)u * 125 + kd * 25 + t) * nkc + (k8 >> 2)) * Nout + nn) * MODE_KC +
                                       (k8 & 3) * 8;
                    a = *reinterpret_cast<const float4*>(src);
                    b = *reinterpret_cast<const float4*>(src + 4);
                }
                float* dst = ws + t * (KC8 * 32) + co;
                dst[0 * 32] = a.x; dst[1 * 32] = a.y; dst[2 * 32] = a.z; dst[3 * 32] = a.w;
                dst[4 * 32] = b.x; dst[5 * 32] = b.y; dst[6 * 32] = b.z; dst[7 * 32] = b.w;
            }
            __syncthreads();
#pragma unroll 1
            for (int kh = 0; kh < 5; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 5; ++kw) {
                    const float* wt = ws + (kh * 5 + kw) * (KC8 * 32) + cg * 8;
                    const float* xt = xs + (vh + kh) * XC + vw + kw;
#pragma unroll
                    for (int ch = 0; ch < KC8; ++ch) {
                        const float x0 = xt[ch * XPLANE], x1 = xt[ch * XPLANE + 16];
                        const float4 wa = *reinterpret_cast<const float4*>(wt + ch * 32);
                        const float4 wb = *reinterpret_cast<const float4*>(wt + ch * 32 + 4);
                        acc0[0] = fmaf(x0, wa.x, acc0[0]); acc0[1] = fmaf(x0, wa.y, acc0[1]);
                        acc0[2] = fmaf(x0, wa.z, acc0[2]); acc0[3] = fmaf(x0, wa.w, acc0[3]);
                        acc0[4] = fmaf(x0, wb.x, acc0[4]); acc0[5] = fmaf(x0, wb.y, acc0[5]);
                        acc0[6] = fmaf(x0, wb.z, acc0[6]); acc0[7] = fmaf(x0, wb.w, acc0[7]);
                        acc1[0] = fmaf(x1, wa.x, acc1[0]); acc1[1] = fmaf(x1, wa.y, acc1[1]);
                        acc1[2] = fmaf(x1, wa.z, acc1[2]); acc1[3] = fmaf(x1, wa.w, acc1[3]);
                        acc1[4] = fmaf(x1, wb.x, acc1[4]); acc1[5] = fmaf(x1, wb.y, acc1[5]);
                        acc1[6] = fmaf(x1, wb.z, acc1[6]); acc1[7] = fmaf(x1, wb.w, acc1[7]);
                    }
                }
            }
        }
    }

    const int hy = th0 + vh;
    const int c0 = ob * 32 + cg * 8;
    if (out_scale_dev != nullptr) out_scale *= *out_scale_dev;
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int wx = tw0 + vw + half * 16;
        const float* acc = half ? acc1 : acc0;
        if (hy < H && wx < W) {
            const size_t vox = (size_t)hy * W + wx;
            float* dst = y ? y + ((((size_t)n * D + d) * H) * W + vox) * Nout + c0 : nullptr;
            __half* d16 = ext.y16 ? ext.y16 + ((((size_t)n * ext.Dy16 + d + ext.y16_off) * H) * W + vox) * Nout + c0 : nullptr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (c0 + j < Nout) {
                    float v = acc[j] * out_scale;
                    if (ext.ep_scale || ext.ep_shift)          // one fma, like the tensor-core epilogues and bn_apply
                        v = fmaf(v, ext.ep_scale ? ext.ep_scale[c0 + j] : 1.f, ext.ep_shift ? ext.ep_shift[c0 + j] : 0.f);
                    if (ext.relu) v = fmaxf(v, 0.f);
                    if (dst) dst[j] = v;
                    if (d16) d16[j] = __float2half_rn(fminf(fmaxf(v * ext.y16_scale, -65504.f), 65504.f));
                    s1[j] += v;
                    s2[j] = fmaf(v, v, s2[j]);
                }
            }
        }
    }
    if (bn_sums != nullptr && d >= stat_lo && d < stat_hi) {   // warp = 32 voxels-threads of one channel group
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], s);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], s);
            }
        }
        if ((tid & 31) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (c0 + j < Nout) {
                    atomicAdd(bn_sums + c0 + j, (double)s1[j]);
                    atomicAdd(bn_sums + Nout + c0 + j, (double)s2[j]);
                }
            }
        }
    }
    if (ext.push.n > 0) push_vector_from_last_block(bn_sums, 2 * Nout, ext.push);
}

// wgrad: grid (125, splits, N*ob*ib), block 256: thread = (o = tid/8, 4 consecutive i)
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dw, int D, int H, int W, int Ci, int Co,
                                                         float out_scale, const float* __restrict__ out_scale_dev,
                                                         int vox_per_split, int Dx, int x_off) {
    __shared__ float dys[64 * 33];
    __shared__ __align__(16) float xsh[64 * 36];
    const int tap = blockIdx.x, split = blockIdx.y;
    const int nob = (Co + 31) / 32, nib = (Ci + 31) / 32;
    const int n = blockIdx.z / (nob * nib), ob = (blockIdx.z / nib) % nob, ib = blockIdx.z % nib;
    const int kd = tap / 25 - 2, kh = (tap / 5) % 5 - 2, kw = tap % 5 - 2;
    const int tid = threadIdx.x, o = tid >> 3, ig = tid & 7;
    const int64_t nvox = (int64_t)D * H * W;
    const int64_t p0 = (int64_t)split * vox_per_split, p1 = min(nvox, p0 + vox_per_split);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t pb = p0; pb < p1; pb += 64) {
        __syncthreads();
        for (int idx = tid; idx < 64 * 32; idx += 256) {
            const int v = idx >> 5, ch = idx & 31;
            const int64_t p = pb + v;
            float a = 0.f, b = 0.f;
            if (p < p1) {
                const int wx = (int)(p % W), hy = (int)((p / W) % H), dz = (int)(p / ((int64_t)W * H));
                const int oo = ob * 32 + ch, ii = ib * 32 + ch;
                if (oo < Co) a = dy[((size_t)n * nvox + p) * Co + oo];
                const int dz2 = dz + kd + x_off, hy2 = hy + kh, wx2 = wx + kw;      // x may carry halo planes
                if (ii < Ci && dz2 >= 0 && dz2 < Dx && hy2 >= 0 && hy2 < H && wx2 >= 0 && wx2 < W)
                    b = x[((((size_t)n * Dx + dz2) * H + hy2) * W + wx2) * Ci + ii];
            }
            dys[v * 33 + ch] = a;
            xsh[v * 36 + ch] = b;
        }
        __syncthreads();
#pragma unroll 8
        for (int v = 0; v < 64; ++v) {
            const float a = dys[v * 33 + o];
            const float4 b = *reinterpret_cast<const float4*>(xsh + v * 36 + ig * 4);
            acc[0] = fmaf(a, b.x, acc[0]); acc[1] = fmaf(a, b.y, acc[1]);
            acc[2] = fmaf(a, b.z, acc[2]); acc[3] = fmaf(a, b.w, acc[3]);
        }
    }
    const int oo = ob * 32 + o;
    if (out_scale_dev != nullptr) out_scale *= *out_scale_dev;
    if (oo < Co) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ii = ib * 32 + ig * 4 + j;
            if (ii < Ci) {
                float* dst = dw + (((size_t)n * 125 + tap) * Co + oo) * Ci + ii;
                if (gridDim.y == 1) *dst = acc[j] * out_scale;
                else atomicAdd(dst, acc[j] * out_scale);
            }
        }
    }
}

int conv3d_simt(const float* x, const float* w, const int32_t* sample_u, float* y, int N, int D, int H, int W, int K,
                int Nout, float out_scale, const float* out_scale_dev, double* bn_sums, int stat_lo, int stat_hi,
                const ConvExt& ext, cudaStream_t st) {
    const int tiles = (int)(ceil_div(H, TH) * ceil_div(W, TW));
    const int64_t gz = (int64_t)N * ceil_div(Nout, 32);
    if (D > 65535 || gz > 65535) MODE_FAIL("conv3d_simt: grid too large (D=%d, N*ob=%lld)", D, (long long)gz);
    conv3d_simt_kernel<<<dim3(tiles, D, (unsigned)gz), 256, 0, st>>>(x, w, sample_u, y, D, H, W, K, Nout, out_scale,
                                                                      out_scale_dev, bn_sums, stat_lo, stat_hi, ext);
    MODE_LAUNCH_CHECK();
    return 0;
}

int wgrad_simt(const float* x, const float* dy, float* dw, int N, int D, int H, int W, int Ci, int Co, float out_scale,
               const float* out_scale_dev, int Dx, int x_off, cudaStream_t st) {
    const int64_t nvox = (int64_t)D * H * W;
    const int64_t gz = (int64_t)N * ceil_div(Co, 32) * ceil_div(Ci, 32);
    if (gz > 65535) MODE_FAIL("wgrad_simt: grid too large");
    // enough blocks for ~4 waves, at least 512 voxels per split
    int64_t splits = ceil_div((int64_t)sm_count() * 4, 125 * gz);
    splits = max((int64_t)1, min(splits, ceil_div(nvox, 512)));
    const int vps = (int)(ceil_div(ceil_div(nvox, splits), 64) * 64);
    splits = ceil_div(nvox, vps);
    if (splits > 1) MODE_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * 125 * Co * Ci * sizeof(float), st));
    wgrad_simt_kernel<<<dim3(125, (unsigned)splits, (unsigned)gz), 256, 0, st>>>(x, dy, dw, D, H, W, Ci, Co, out_scale,
                                                                               out_scale_dev, vps, Dx, x_off);
    MODE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mode
