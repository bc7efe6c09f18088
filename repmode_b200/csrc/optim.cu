// Step glue of Model.do_train_iter (fnet/fnet_model.py:96-132 in the reference tree): the Adam update
// of all 309 parameter tensors (123.9 M fp32 values) as ONE multi-tensor launch instead of 309 x ~12 elementwise kernels,
// with the GradScaler hand-shake (gradients arrive multiplied by `grad_scale`; a step whose gradients contain inf/nan is
// skipped) folded in, so that scaler.step(optimizer) needs no host synchronisation.
//
// Arithmetic = torch.optim.Adam (amsgrad = False, maximize = False), the optimizer the reference constructs at
// fnet_model.py:55:   g = grad / grad_scale (+ weight_decay * p);  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//                     p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// HBM-bound: 28 B per parameter (read p, g, m, v; write p, m, v); float4 accesses; a chunk table in device memory maps
// blocks to (tensor, offset), so tensors of any size and address take part in the same launch.
#include <algorithm>

#include "common.cuh"

namespace mode {

struct AdamTensor { float* p; const float* g; float* m; float* v; float* step; long long n; };   // 48 bytes
struct AdamChunk { int tensor; int pad; long long off; };                                          // 16 bytes
constexpr int ADAM_CHUNK = 65536;      // elements per block

__global__ void adam_advance_steps_kernel(const AdamTensor* __restrict__ T, int ntensors, const float* found_inf) {
    if (found_inf != nullptr && *found_inf != 0.f) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntensors) *T[i].step += 1.f;
}

__global__ void __launch_bounds__(512) adam_multi_kernel(const AdamTensor* __restrict__ T, const AdamChunk* __restrict__ C,
                                                         double lr_d, double b1_d, double b2_d, float eps, float wd,
                                                         const float* grad_scale, const float* found_inf) {
    if (found_inf != nullptr && *found_inf != 0.f) return;               // GradScaler: skip the whole step
    const AdamChunk c = C[blockIdx.x];
    const AdamTensor t = T[c.tensor];
    const float step = *t.step;                                          // already advanced by adam_advance_steps_kernel
    const float inv_scale = grad_scale != nullptr ? 1.f / *grad_scale : 1.f;
    // bias corrections in double, like the Python scalars of torch.optim.Adam; rounded to fp32 where torch hands them to
    // its fp32 tensor ops
    const float b2 = (float)b2_d;
    const float omb1 = (float)(1.0 - b1_d), omb2 = (float)(1.0 - b2_d);     // torch: lerp weight 1 - beta1, addcmul value 1 - beta2
    const float bc2_sqrt = (float)sqrt(1.0 - pow(b2_d, (double)step));
    const float step_size = (float)(lr_d / (1.0 - pow(b1_d, (double)step)));
    const long long end = min(t.n, c.off + (long long)ADAM_CHUNK);
    const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                       reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
    auto upd = [&](float& p, float g, float& m, float& v) {
        g *= inv_scale;
        if (wd != 0.f) g = fmaf(wd, p, g);
        m = fmaf(omb1, g - m, m);                                  // exp_avg.lerp_(grad, 1 - beta1)
        v = fmaf(omb2 * g, g, v * b2);                              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        p -= step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
    };
    if (vec) {
        const long long n4 = (end - c.off) >> 2;
        float4* p4 = reinterpret_cast<float4*>(t.p + c.off);
        const float4* g4 = reinterpret_cast<const float4*>(t.g + c.off);
        float4* m4 = reinterpret_cast<float4*>(t.m + c.off);
        float4* v4 = reinterpret_cast<float4*>(t.v + c.off);
        for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
            float4 p = p4[i], m = m4[i], v = v4[i];
            const float4 g = __ldcs(g4 + i);
            upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
            p4[i] = p; m4[i] = m; v4[i] = v;
        }
        for (long long i = c.off + (n4 << 2) + threadIdx.x; i < end; i += blockDim.x) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
    } else {
        for (long long i = c.off + threadIdx.x; i < end; i += blockDim.x) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
    }
}

}  // namespace mode

using namespace mode;

extern "C" int64_t mode_adam_chunk_elems(void) { return ADAM_CHUNK; }

// tensors_dev: [ntensors] records of 6 x 8 bytes {p, g, m, v, step (fp32 scalar on the device), n}; chunks_dev: [nchunks]
// records of {int32 tensor, int32 0, int64 offset} covering every tensor in pieces of mode_adam_chunk_elems() elements.
extern "C" int mode_adam_step(const void* tensors_dev, int32_t ntensors, const void* chunks_dev, int32_t nchunks, double lr,
                              double beta1, double beta2, double eps, double weight_decay, const float* grad_scale_dev,
                              const float* found_inf_dev, void* stream) {
    static_assert(sizeof(AdamTensor) == 48 && sizeof(AdamChunk) == 16, "table record sizes are part of the ABI");
    if (!tensors_dev || !chunks_dev || ntensors <= 0 || nchunks <= 0) MODE_FAIL("mode_adam_step: empty tables");
    if (!(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0.))
        MODE_FAIL("mode_adam_step: bad hyper-parameters (beta1=%g beta2=%g eps=%g)", beta1, beta2, eps);
    cudaStream_t st = (cudaStream_t)stream;
    adam_advance_steps_kernel<<<(int)ceil_div(ntensors, 128), 128, 0, st>>>((const AdamTensor*)tensors_dev, ntensors, found_inf_dev);
    MODE_LAUNCH_CHECK();
    adam_multi_kernel<<<nchunks, 512, 0, st>>>((const AdamTensor*)tensors_dev, (const AdamChunk*)chunks_dev, lr, beta1, beta2,
                                               (float)eps, (float)weight_decay, grad_scale_dev, found_inf_dev);
    MODE_LAUNCH_CHECK();
    return 0;
}
