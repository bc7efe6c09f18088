// Sliding-window inference glue of Model.predict (fnet/fnet_model.py:149-223 in the reference tree):
// the Gaussian-weighted accumulation of a batch of predicted patches into the volume (reference :207-214, two sliced
// read-modify-write passes over pred_sum and weight_sum per patch) and the final division (:220), as two HBM-bound kernels.
//   accumulate: one thread per volume voxel gathers every patch of the batch that covers it (deterministic: patches are added
//               in batch order, no atomics even where the half-overlapping windows of one batch intersect); voxels no patch
//               of the batch covers are not touched.  Traffic: 8 B per covered voxel + 8 B per (patch, voxel) read.
//   finalize:   out = pred_sum / weight_sum (weight_sum is the same for every channel and kept once).
#include <algorithm>

#include "common.cuh"

namespace mode {

struct BlendParams {
    const float* pred;            // [P][C][pd][ph][pw]
    const int32_t* starts;        // [P][3] window origin (d, h, w) in the volume, device memory
    const float* gauss;           // [gd][gh][gw] importance map, gd >= pd ...
    float* pred_sum;              // [C][D][H][W]
    float* weight_sum;            // [D][H][W]
    int P, C, pd, ph, pw, gh, gw, D, H, W;
};

__global__ void __launch_bounds__(256) blend_accumulate_kernel(const BlendParams B) {
    __shared__ int s_start[64 * 3];
    for (int i = threadIdx.x; i < 3 * B.P; i += 256) s_start[i] = B.starts[i];
    __syncthreads();
    const long long vol = (long long)B.D * B.H * B.W;
    const long long psz = (long long)B.pd * B.ph * B.pw;
    for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < vol; v += (long long)gridDim.x * 256) {
        const int w = (int)(v % B.W);
        const int h = (int)((v / B.W) % B.H);
        const int d = (int)(v / ((long long)B.W * B.H));
        float wacc = 0.f;
        bool hit = false;
        for (int c = 0; c < B.C; ++c) {
            float acc = 0.f;
            for (int p = 0; p < B.P; ++p) {
                const int ld = d - s_start[3 * p], lh = h - s_start[3 * p + 1], lw = w - s_start[3 * p + 2];
                if ((unsigned)ld >= (unsigned)B.pd || (unsigned)lh >= (unsigned)B.ph || (unsigned)lw >= (unsigned)B.pw) continue;
                const float g = __ldg(B.gauss + ((long long)ld * B.gh + lh) * B.gw + lw);
                acc = fmaf(__ldg(B.pred + ((long long)p * B.C + c) * psz + ((long long)ld * B.ph + lh) * B.pw + lw), g, acc);
                if (c == 0) { wacc += g; hit = true; }
            }
            if (hit) B.pred_sum[(long long)c * vol + v] += acc;
        }
        if (hit) B.weight_sum[v] += wacc;
    }
}

__global__ void __launch_bounds__(256) blend_finalize_kernel(const float* __restrict__ pred_sum,
                                                             const float* __restrict__ weight_sum, float* __restrict__ out,
                                                             int C, long long vol) {
    for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < vol; v += (long long)gridDim.x * 256) {
        const float inv_w = weight_sum[v];
        for (int c = 0; c < C; ++c) out[(long long)c * vol + v] = pred_sum[(long long)c * vol + v] / inv_w;
    }
}

}  // namespace mode

using namespace mode;

extern "C" int mode_blend_accumulate(const float* pred, const int32_t* starts_dev, const float* gauss, float* pred_sum,
                                     float* weight_sum, int32_t P, int32_t C, int32_t pd, int32_t ph, int32_t pw, int32_t gh,
                                     int32_t gw, int32_t D, int32_t H, int32_t W, void* stream) {
    if (!pred || !starts_dev || !gauss || !pred_sum || !weight_sum) MODE_FAIL("mode_blend_accumulate: null pointer");
    if (P <= 0 || P > 64) MODE_FAIL("mode_blend_accumulate: 1..64 patches per call (got %d)", P);
    if (C <= 0 || pd <= 0 || ph <= 0 || pw <= 0 || D <= 0 || H <= 0 || W <= 0 || gh < ph || gw < pw)
        MODE_FAIL("mode_blend_accumulate: bad dimensions");
    BlendParams B{pred, starts_dev, gauss, pred_sum, weight_sum, P, C, pd, ph, pw, gh, gw, D, H, W};
    const long long vol = (long long)D * H * W;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ceil_div(vol, 256), 8LL * sm_count()));
    blend_accumulate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(B);
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_blend_finalize(const float* pred_sum, const float* weight_sum, float* out, int32_t C, int64_t voxels,
                                   void* stream) {
    if (!pred_sum || !weight_sum || !out || C <= 0 || voxels <= 0) MODE_FAIL("mode_blend_finalize: bad arguments");
    const int grid = (int)std::max<long long>(1, std::min<long long>(ceil_div(voxels, 256), 8LL * sm_count()));
    blend_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred_sum, weight_sum, out, C, voxels);
    MODE_LAUNCH_CHECK();
    return 0;
}
