// C-ABI glue: error reporting, capability query and the implementation dispatch for the conv entry points.
// See include/repmode_b200.h for the contract and the reference lines each entry point replaces.
#include <atomic>

#include <stdlib.h>

#include "common.cuh"

namespace mode {

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

char* err_buf() {
    static thread_local char buf[kErrLen] = "";
    return buf;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// One int per device, set by a kernel whose pipeline timed out (never hangs the box); read by mode_poll_error.
int* device_error_flag() {
    static int* flags[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!flags[dev]) {
        int* p = nullptr;
        if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, sizeof(int));
        flags[dev] = p;
    }
    return flags[dev];
}

static long long* g_prof = nullptr;
long long* debug_profile_buffer() { return g_prof; }

// conv_simt.cu
int conv3d_simt(const float* x, const float* w, const int32_t* sample_u, float* y, int N, int D, int H, int W, int K,
                int Nout, float out_scale, const float* out_scale_dev, double* bn_sums, int stat_lo, int stat_hi,
                const ConvExt& ext, cudaStream_t st);
int wgrad_simt(const float* x, const float* dy, float* dw, int N, int D, int H, int W, int Ci, int Co, float out_scale,
               const float* out_scale_dev, int Dx, int x_off, cudaStream_t st);
// conv_umma.cu
bool conv3d_umma_supported(int D, int H, int W, int K, int Nout);
int64_t conv3d_umma_workspace_bytes(int N, int D, int H, int W, int K, int Nout);
int conv3d_umma(const __half* x, const __half* w, const int32_t* sample_u, float* y, int N, int D, int H, int W, int K,
                int Nout, float out_scale, const float* out_scale_dev, double* bn_sums, int stat_lo, int stat_hi,
                const ConvExt& ext, cudaStream_t st);
// conv_pair.cu
bool conv3d_pair_supported(int N, int D, int H, int W, int K, int Nout);
int conv3d_pair(const __half* x, const __half* w, const int32_t* sample_u, float* y, int N, int D, int H, int W, int K,
                int Nout, float out_scale, const float* out_scale_dev, double* bn_sums, int stat_lo, int stat_hi,
                const ConvExt& ext, cudaStream_t st);
// wgrad_split.cu
bool wgrad_split_supported(int D, int H, int W, int Ci, int Co);
int64_t wgrad_split_workspace_bytes(int N, int D, int H, int W, int Ci, int Co);
int wgrad_split(const __half* x, const __half* dy, float* dw, int N, int D, int H, int W, int Ci, int Co,
                float out_scale, const float* out_scale_dev, void* workspace, cudaStream_t st);

// wgrad_deep.cu
bool wgrad_deep_supported(int D, int H, int W, int Ci, int Co);
int64_t wgrad_deep_workspace_bytes(int N, int D, int H, int W, int Ci, int Co, int Dx, int x_off);
void wgrad_deep_layout(int N, int D, int H, int W, int Ci, int Co, int Dx, int x_off, int32_t out[5]);
int wgrad_deep(const __half* x, const __half* dy, float* dw, int N, int D, int H, int W, int Ci, int Co,
               float out_scale, const float* out_scale_dev, void* workspace, int Dx, int x_off, cudaStream_t st, int phase);

// tcgen05 wgrad flavour behind impl = 2: the deep-tile kernel (wgrad_deep.cu; r2a on the headline layer: 112 us against
// 135 us for the split-tap kernel of round 1); REPMODE_WGRAD_SPLIT=1 selects wgrad_split.cu (the A/B arm)
static bool wgrad_use_deep() {
    static const bool split = getenv("REPMODE_WGRAD_SPLIT") != nullptr;
    return !split;
}

}  // namespace mode

using namespace mode;

extern "C" const char* mode_last_error(void) { return err_buf(); }
extern "C" int mode_version(void) { return MODE_ABI_VERSION; }
extern "C" int64_t mode_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" int mode_query(int device, mode_caps_t* caps) {
    if (!caps) MODE_FAIL("mode_query: caps is NULL");
    cudaDeviceProp p;
    MODE_CUDA(cudaGetDeviceProperties(&p, device));
    caps->sm_major = p.major;
    caps->sm_minor = p.minor;
    caps->sm_count = p.multiProcessorCount;
    caps->smem_per_block_optin = (int32_t)p.sharedMemPerBlockOptin;
    caps->tmem_columns = 512;
    caps->abi_version = MODE_ABI_VERSION;
    if (p.major != 10) MODE_FAIL("mode_query: device %d is sm_%d%d; this library contains sm_100a code only", device, p.major, p.minor);
    return 0;
}

extern "C" int mode_poll_error(int32_t* code_host) {
    if (!code_host) MODE_FAIL("mode_poll_error: null");
    int* f = device_error_flag();
    if (!f) MODE_FAIL("mode_poll_error: no flag");
    MODE_CUDA(cudaDeviceSynchronize());
    int v = 0;
    MODE_CUDA(cudaMemcpy(&v, f, sizeof(int), cudaMemcpyDeviceToHost));
    *code_host = v;
    if (v != 0) MODE_CUDA(cudaMemset(f, 0, sizeof(int)));
    return 0;
}

extern "C" int mode_debug_profile(void* buf) {
    g_prof = (long long*)buf;
    return 0;
}

extern "C" int mode_conv3d_ex(const void* x, mode_dtype_t x_dtype, const void* w, const int32_t* sample_u, float* y,
                              int32_t N, int32_t D, int32_t H, int32_t W, int32_t K, int32_t Nout, float out_scale,
                              const float* out_scale_dev, double* bn_sums, int32_t stat_d_lo, int32_t stat_d_hi,
                              int32_t impl, const mode_conv_opts_t* opts, void* stream) {
    ConvExt ext;
    ext.Dx = (opts && opts->Dx > 0) ? opts->Dx : D;
    ext.x_off = opts ? opts->x_off : 0;
    ext.ep_scale = opts ? opts->ep_scale : nullptr;
    ext.ep_shift = opts ? opts->ep_shift : nullptr;
    ext.relu = opts ? opts->relu : 0;
    ext.y16 = opts ? (__half*)opts->y16 : nullptr;
    ext.Dy16 = (opts && opts->Dy16 > 0) ? opts->Dy16 : D;
    ext.y16_off = opts ? opts->y16_off : 0;
    ext.y16_scale = (opts && opts->y16_scale != 0.f) ? opts->y16_scale : 1.f;
    ext.splitk_ws = opts ? opts->splitk_ws : nullptr;
    ext.splitk_ws_bytes = opts ? (long long)opts->splitk_ws_bytes : 0;
    ext.push = to_peer_push(opts ? opts->stats_push : nullptr);
    if (ext.push.n < 0 || ext.push.n > 8 || (ext.push.n > 0 && (!ext.push.ticket || !bn_sums)))
        MODE_FAIL("mode_conv3d: bad stats_push descriptor (needs bn_sums, a ticket and <= 8 destinations)");
    if (!x || !w || (!y && !ext.y16)) MODE_FAIL("mode_conv3d: null pointer");
    if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || K <= 0 || Nout <= 0) MODE_FAIL("mode_conv3d: non-positive dimension");
    if (ext.x_off < 0 || ext.x_off + D > ext.Dx)
        MODE_FAIL("mode_conv3d: haloed input needs 0 <= x_off and x_off + D <= Dx (x_off=%d D=%d Dx=%d)", ext.x_off, D, ext.Dx);
    if (ext.y16 && (ext.y16_off < 0 || ext.y16_off + D > ext.Dy16))
        MODE_FAIL("mode_conv3d: y16 plane window out of range (y16_off=%d D=%d Dy16=%d)", ext.y16_off, D, ext.Dy16);
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == 0) impl = (x_dtype == MODE_F16) ? 2 : 1;
    if (impl == 1) {
        if (x_dtype != MODE_F32) MODE_FAIL("mode_conv3d: the SIMT path takes fp32 operands");
        return conv3d_simt((const float*)x, (const float*)w, sample_u, y, N, D, H, W, K, Nout, out_scale, out_scale_dev, bn_sums, stat_d_lo, stat_d_hi, ext, st);
    }
    if (impl == 2 || impl == 3 || impl == 4) {
        if (x_dtype != MODE_F16) MODE_FAIL("mode_conv3d: the tcgen05 path takes fp16 operands");
        if (!conv3d_umma_supported(D, H, W, K, Nout))
            MODE_FAIL("mode_conv3d: shape D=%d H=%d W=%d K=%d Nout=%d not supported by the tcgen05 path", D, H, W, K, Nout);
        static const bool no_pair = getenv("REPMODE_DISABLE_PAIR") != nullptr;
        // 2: CTA-pair kernel when every cluster gets a long enough march, else the single-CTA kernel; 3 / 4 force one
        // (K > 32 on the pair kernel accumulates chunk by chunk in the fp32 output: not available for an fp16-only result)
        if (impl == 4 || (impl == 2 && !no_pair && (K == 32 || y != nullptr) && conv3d_pair_supported(N, D, H, W, K, Nout)))
            return conv3d_pair((const __half*)x, (const __half*)w, sample_u, y, N, D, H, W, K, Nout, out_scale, out_scale_dev, bn_sums, stat_d_lo, stat_d_hi, ext, st);
        return conv3d_umma((const __half*)x, (const __half*)w, sample_u, y, N, D, H, W, K, Nout, out_scale, out_scale_dev, bn_sums, stat_d_lo, stat_d_hi, ext, st);
    }
    MODE_FAIL("mode_conv3d: unknown impl %d", impl);
}

extern "C" int64_t mode_conv3d_workspace_bytes(int32_t N, int32_t D, int32_t H, int32_t W, int32_t K, int32_t Nout,
                                               mode_dtype_t x_dtype) {
    if (x_dtype != MODE_F16 || N <= 0 || D <= 0 || H <= 0 || W <= 0 || K <= 0 || Nout <= 0) return 0;
    static const bool no_pair = getenv("REPMODE_DISABLE_PAIR") != nullptr;
    if (!conv3d_umma_supported(D, H, W, K, Nout)) return 0;
    if (!no_pair && conv3d_pair_supported(N, D, H, W, K, Nout)) return 0;      // the CTA-pair kernel never splits
    return conv3d_umma_workspace_bytes(N, D, H, W, K, Nout);
}

extern "C" int mode_conv3d(const void* x, mode_dtype_t x_dtype, const void* w, const int32_t* sample_u, float* y,
                           int32_t N, int32_t D, int32_t H, int32_t W, int32_t K, int32_t Nout, float out_scale,
                           const float* out_scale_dev, double* bn_sums, int32_t stat_d_lo, int32_t stat_d_hi,
                           int32_t impl, void* stream) {
    return mode_conv3d_ex(x, x_dtype, w, sample_u, y, N, D, H, W, K, Nout, out_scale, out_scale_dev, bn_sums, stat_d_lo,
                          stat_d_hi, impl, nullptr, stream);
}

extern "C" int64_t mode_conv3d_wgrad_workspace_bytes(int32_t N, int32_t D, int32_t H, int32_t W, int32_t Ci, int32_t Co,
                                                     int32_t impl) {
    if (impl == 6 || (impl == 2 && wgrad_use_deep())) return wgrad_deep_workspace_bytes(N, D, H, W, Ci, Co, D, 0);
    if (impl == 5 || impl == 2) return wgrad_split_workspace_bytes(N, D, H, W, Ci, Co);
    return 0;
}

extern "C" int mode_conv3d_wgrad_ex(const void* x, const void* dy, mode_dtype_t dtype, float* d_weff, int32_t N,
                                    int32_t D, int32_t H, int32_t W, int32_t Ci, int32_t Co, float out_scale,
                                    const float* out_scale_dev, void* workspace, int32_t impl, int32_t Dx, int32_t x_off,
                                    void* stream) {
    if (!x || !dy || !d_weff) MODE_FAIL("mode_conv3d_wgrad: null pointer");
    if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || Ci <= 0 || Co <= 0) MODE_FAIL("mode_conv3d_wgrad: non-positive dimension");
    if (Dx <= 0) { Dx = D; x_off = 0; }
    if (x_off < 0 || x_off + D > Dx)
        MODE_FAIL("mode_conv3d_wgrad: haloed x needs 0 <= x_off and x_off + D <= Dx (x_off=%d D=%d Dx=%d)", x_off, D, Dx);
    const bool halo = Dx != D || x_off != 0;
    cudaStream_t st = (cudaStream_t)stream;
    // bits 8-9 of impl: 0 = the whole wgrad; 1 = the tensor-core part only; 2 = what is left after it (the slab reduce of the
    // deep-tile kernel; nothing for the other kernels) -- a caller may start K3 between the two calls
    const int phase = (impl >> 8) & 3;
    impl &= 0xff;
    if (impl == 0) impl = (dtype == MODE_F16) ? 2 : 1;
    if (phase == 2 && !(impl == 6 || (impl == 2 && (wgrad_use_deep() || halo)))) return 0;
    if (impl == 1) {
        if (dtype != MODE_F32) MODE_FAIL("mode_conv3d_wgrad: the SIMT path takes fp32 operands");
        return wgrad_simt((const float*)x, (const float*)dy, d_weff, N, D, H, W, Ci, Co, out_scale, out_scale_dev, Dx, x_off, st);
    }
    if (impl == 2 || impl == 5 || impl == 6) {
        if (dtype != MODE_F16) MODE_FAIL("mode_conv3d_wgrad: the tcgen05 path takes fp16 operands");
        if (!wgrad_deep_supported(D, H, W, Ci, Co) || !wgrad_split_supported(D, H, W, Ci, Co))
            MODE_FAIL("mode_conv3d_wgrad: shape not supported by the tcgen05 path");
        // 2: deep-tile kernel unless REPMODE_WGRAD_SPLIT; 5 / 6 force the split-tap / the deep-tile kernel
        if (impl == 6 || (impl == 2 && (wgrad_use_deep() || halo)))
            return wgrad_deep((const __half*)x, (const __half*)dy, d_weff, N, D, H, W, Ci, Co, out_scale, out_scale_dev, workspace, Dx, x_off, st, phase);
        if (halo) MODE_FAIL("mode_conv3d_wgrad: the split-tap kernel (impl 5) does not take a haloed x");
        return wgrad_split((const __half*)x, (const __half*)dy, d_weff, N, D, H, W, Ci, Co, out_scale, out_scale_dev, workspace, st);
    }
    MODE_FAIL("mode_conv3d_wgrad: unknown impl %d", impl);
}

extern "C" int mode_conv3d_wgrad_partial_layout(mode_dtype_t dtype, int32_t N, int32_t D, int32_t H, int32_t W, int32_t Ci,
                                                int32_t Co, int32_t impl, int32_t Dx, int32_t x_off, int32_t* layout5) {
    if (!layout5) MODE_FAIL("mode_conv3d_wgrad_partial_layout: layout5 is NULL");
    if (Dx <= 0) { Dx = D; x_off = 0; }
    const bool halo = Dx != D || x_off != 0;
    impl &= 0xff;
    if (impl == 0) impl = (dtype == MODE_F16) ? 2 : 1;
    const bool deep = dtype == MODE_F16 && (impl == 6 || (impl == 2 && (wgrad_use_deep() || halo))) &&
                      wgrad_deep_supported(D, H, W, Ci, Co);
    if (!deep) return 1;                       // not an error: this wgrad does not leave deep-tile partials behind
    wgrad_deep_layout(N, D, H, W, Ci, Co, Dx, x_off, layout5);
    return 0;
}

extern "C" int mode_conv3d_wgrad(const void* x, const void* dy, mode_dtype_t dtype, float* d_weff, int32_t N,
                                 int32_t D, int32_t H, int32_t W, int32_t Ci, int32_t Co, float out_scale,
                                 const float* out_scale_dev, void* workspace, int32_t impl, void* stream) {
    return mode_conv3d_wgrad_ex(x, dy, dtype, d_weff, N, D, H, W, Ci, Co, out_scale, out_scale_dev, workspace, impl, 0, 0, stream);
}
