/* repmode_b200 C ABI -- the drop-in boundary for RepMode's MoDE-conv hot path on B200 (sm_100a).
 *
 * The reference has no native code: its hot path is PyTorch operators called from
 * fnet/nn_modules/RepMode.py (paths relative to the reference tree).  Each entry point below replaces
 * the operator sequence cited next to it.  The host side (repmode_b200/functional.py, Python, because the
 * reference's plugin boundary `fnet.nn_modules.<name>.Net` is Python -- fnet/fnet_model.py:52) binds
 * these symbols with ctypes and passes raw device pointers (tensor.data_ptr()) plus the caller's
 * cudaStream_t.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; mode_last_error() returns a thread-local
 *     message.  No C++ exception crosses the boundary.
 *   - all pointers are DEVICE pointers unless the name ends in _host; the library never allocates or
 *     frees caller-visible memory and keeps no mutable global state besides per-device attribute caches.
 *   - activations are dense NDHWC ("channels_last_3d"): x[n][d][h][w][c], c contiguous.
 *   - int64_t for element counts, int32_t for channel / small counts.
 */
#ifndef REPMODE_B200_H_
#define REPMODE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MODE_ABI_VERSION 1
#define MODE_NUM_EXPERTS 5   /* RepMode.py:22 (num_experts hard-coded), :136-142 */
#define MODE_TAPS 125        /* effective kernel is always 5x5x5, RepMode.py:114-115 */
#define MODE_KC 32           /* K-chunk (channels per 64-byte fp16 row) of the packed weight layout */

typedef enum { MODE_F32 = 0, MODE_F16 = 1 } mode_dtype_t;

/* One MoDEConv layer's parameters, reference layouts (RepMode.py:136-142,153), fp32, device memory. */
typedef struct {
    const float* k5;      /* expert_conv5x5_conv [Co,Ci,5,5,5] */
    const float* k3;      /* expert_conv3x3_conv [Co,Ci,3,3,3] */
    const float* k1;      /* expert_conv1x1_conv [Co,Ci,1,1,1] */
    const float* a3;      /* expert_avg3x3_conv  [Co,Ci,1,1,1] (times the 1/27 pool constant) */
    const float* a5;      /* expert_avg5x5_conv  [Co,Ci,1,1,1] (times the 1/125 pool constant) */
    const float* gate_w;  /* gate.weight [5*Co, T], row index e*Co+o (RepMode.py:199) */
    const float* gate_b;  /* gate.bias   [5*Co] */
    int32_t ci, co, num_tasks;
} mode_layer_t;

typedef struct {
    int32_t sm_major, sm_minor, sm_count;
    int32_t smem_per_block_optin;   /* bytes */
    int32_t tmem_columns;           /* 512 on sm_100 */
    int32_t abi_version;
} mode_caps_t;

const char* mode_last_error(void);
int mode_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches evidence). */
int64_t mode_launch_count(void);
/* Fails unless `device` is compute capability 10.x (this library ships sm_100a SASS only). */
int mode_query(int device, mode_caps_t* caps_host);
/* Synchronises the device and returns (and clears) the pipeline-timeout code a tcgen05 kernel raises
 * instead of hanging (0 = none). Diagnostics only; never called on the hot path. */
int mode_poll_error(int32_t* code_host);
/* Diagnostics: device buffer of int64[8 * 160] (or NULL to switch off) that the tcgen05 conv kernel fills per CTA
 * with MMA-warp cycles {total, wait accumulators, wait weights, wait planes} and globaltimer ns {CTA entry, MMA
 * loop end, all roles done, 0}. Process-global, not thread-safe, never set on the hot path. */
int mode_debug_profile(void* buf);

/* ---- K1: gate softmax + expert re-parameterisation --------------------------------------------------
 * Replaces: Linear + view + Softmax(dim=1) (RepMode.py:198-200) and MoDEConv.routing / trans_kernel
 * (RepMode.py:165-192), for U gate inputs at once (U = distinct tasks in the batch; the reference
 * recomputes per sample).
 *   gate input: either task_ids[U] (one-hot embedding == column gather, RepMode.py:44-49) or a dense
 *               t[U,T] row (the MoDEConv.forward(x, t) signature accepts any t); exactly one non-NULL.
 *   g_out [U,5,Co] fp32: the softmax gates (saved for backward).
 *   w_fwd: packed conv weights, rows = co, k = ci (k within a 32-chunk contiguous, zero padded), tap =
 *          (kd*5+kh)*5+kw:
 *            fp32 pack ("tap-major", SIMT kernels):      w[u][tap][ci_chunk][co][MODE_KC]
 *            fp16 pack ("stage-major", tcgen05 kernel):  w[u][ci_chunk][kh*5+kw][4-kd][co][MODE_KC], each
 *            [rows][MODE_KC] block stored as the 64-byte-swizzled shared-memory image the UMMA B operand reads
 *            (16-byte chunk index ^= (row >> 1) & 3).
 *   w_dgrad (may be NULL): the same packs built from the flipped, io-transposed kernel (tap -> 124-tap,
 *          rows = ci, k = co), same dtype.
 *   w_scale * (w_scale_dev ? *w_scale_dev : 1) multiplies every packed weight before rounding (a power of
 *   two keeps the fp16 operand an exactly-scaled copy); w_scale_dev is a device scalar so the scale can
 *   be produced on-stream (mode_f16_scale) without a host sync.
 */
int mode_reparam_fwd(const mode_layer_t* layer_host, const int32_t* task_ids, const float* t_dense, int32_t U,
                     float* g_out, void* w_fwd, void* w_dgrad, mode_dtype_t w_dtype, float w_scale,
                     const float* w_scale_dev, void* stream);
/* K1 for several layers in ONE launch (plus one grouped dgrad-pack launch): the layers of a U-Net step share the gate
 * inputs (Net.forward hands the same t to every MoDEConv, RepMode.py:51-71).  fp16 packs only; every item needs
 * Ci % 32 == 0 and Co % 32 == 0 (the stem / head layers go through mode_reparam_fwd); w_dgrad may be NULL per item. */
#define MODE_REPARAM_GROUP_MAX 24
typedef struct {
    mode_layer_t layer;
    float* g_out;        /* [U,5,Co] */
    void* w_fwd;         /* U * mode_packed_weight_elems_f16(ci, co) halves */
    void* w_dgrad;       /* U * mode_packed_weight_elems_f16(co, ci) halves, or NULL */
} mode_reparam_item_t;
int mode_reparam_fwd_grouped(const mode_reparam_item_t* items_host, int32_t n_items, const int32_t* task_ids,
                             const float* t_dense, int32_t U, mode_dtype_t w_dtype, float w_scale, void* stream);
int64_t mode_packed_weight_elems(int32_t k_channels, int32_t n_channels);       /* fp32 pack, per gate input u */
/* fp16 pack: rows padded to a multiple of 32 too (pad rows are NOT written: zero-initialise when n % 32 != 0) */
int64_t mode_packed_weight_elems_f16(int32_t k_channels, int32_t n_channels);

/* ---- K1b: backward of K1 ------------------------------------------------------------------------------
 * Replaces autograd through routing/softmax/Linear.  d_weff [N][125][Co][Ci] fp32 is the per-sample
 * gradient of the effective kernel (what K4 writes); sample n used gate input u = sample_u[n].
 * Outputs (all fp32, reference parameter layouts, overwritten): dk5 dk3 dk1 da3 da5, dgate_w [5Co,T],
 * dgate_b [5Co].  workspace: mode_reparam_bwd_workspace_bytes().
 */
int64_t mode_reparam_bwd_workspace_bytes(int32_t ci, int32_t co, int32_t n_samples);
int mode_reparam_bwd(const mode_layer_t* layer_host, const int32_t* task_ids, const float* t_dense, int32_t U,
                     const int32_t* sample_u, int32_t n_samples, const float* g, const float* d_weff,
                     float* dk5, float* dk3, float* dk1, float* da3, float* da5, float* dgate_w, float* dgate_b,
                     void* workspace, void* stream);

/* K1b straight out of K4's work-unit partials: when mode_conv3d_wgrad_ex's tensor-core phase (impl | 0x100) ran ONE slab per
 * unit group -- mode_conv3d_wgrad_partial_layout() returns 0 and layout5 = {SL, SA, SB, nL, nA} with SL = SA = SB = 1 --
 * its reduce phase is a pure re-layout of d_weff, which this entry point performs inside K1b's loads instead: d_weff is
 * never written.  `k4_partials` is the workspace that wgrad call filled; scale * (scale_dev ? *scale_dev : 1) is the
 * out_scale the reduce phase would have applied.  Same outputs, bit for bit, as the reduce phase + mode_reparam_bwd. */
int mode_conv3d_wgrad_partial_layout(mode_dtype_t dtype, int32_t N, int32_t D, int32_t H, int32_t W, int32_t Ci, int32_t Co,
                                     int32_t impl, int32_t Dx, int32_t x_off, int32_t* layout5_host);
int mode_reparam_bwd_partial(const mode_layer_t* layer_host, const int32_t* task_ids, const float* t_dense, int32_t U,
                             const int32_t* sample_u, int32_t n_samples, const float* g, const float* k4_partials,
                             const int32_t* layout5_host, float scale, const float* scale_dev, float* dk5, float* dk3,
                             float* dk1, float* da3, float* da5, float* dgate_w, float* dgate_b, void* workspace,
                             void* stream);

/* ---- K2 / K3: 5x5x5 'same' cross-correlation, stride 1, zero pad 2, no bias ---------------------------
 * Replaces F.conv3d(x[i:i+1], w[i], padding='same') per sample (RepMode.py:204-210) and, called with the
 * w_dgrad pack and dy as input, its autograd dgrad.
 *   x [N,D,H,W,K] (x_dtype), w = packed weights from K1 (same dtype as x), sample_u[N] selects the weight
 *   set per sample (all zeros in eval mode, RepMode.py:209-210), y [N,D,H,W,Nout] fp32.
 *   out_scale * (out_scale_dev ? *out_scale_dev : 1) multiplies the accumulator (undoes operand scaling).
 *   bn_sums (may be NULL): double[2*Nout], += per-channel sum and sum of squares of y (BatchNorm3d
 *   training statistics, RepMode.py:147) -- fused so y is not re-read -- over the d-planes
 *   [stat_d_lo, stat_d_hi) only (the OWNED planes of a D-sharded slab; pass 0, D for the whole tensor).
 *   impl: 0 = auto, 1 = SIMT fp32 direct conv (any shape, fp32 operands only), 2 = tcgen05 implicit GEMM
 *   (fp16 operands, K % 32 == 0, Nout % 32 == 0, W % 8 == 0; picks the CTA-pair kernel for large volumes),
 *   3 = force the single-CTA tcgen05 kernel, 4 = force the CTA-pair (cta_group::2) kernel.
 */
int mode_conv3d(const void* x, mode_dtype_t x_dtype, const void* w, const int32_t* sample_u, float* y,
                int32_t N, int32_t D, int32_t H, int32_t W, int32_t K, int32_t Nout, float out_scale,
                const float* out_scale_dev, double* bn_sums, int32_t stat_d_lo, int32_t stat_d_hi, int32_t impl,
                void* stream);

/* ---- descriptors of exchange steps FUSED into a kernel (D-sharded slabs over peer memory, see the peer section below) -----
 * mode_halo_push_t: the kernel that writes a tensor also stores the tensor's first / last `bytes` (its boundary planes) into
 *   the lower / upper neighbour's halo planes and, when the whole launch is done, increments their counters.
 * mode_peer_push_t: the last block of a kernel broadcasts a small fp64 vector (its per-channel sums) into slot dst[i] of
 *   every destination and increments signal[i].
 * mode_peer_gather_t: the consuming kernel waits until `world` producers have signalled (*signal >= *expect + world), sums
 *   slots[0..world) in rank order instead of reading a local vector, and advances *expect.
 * All pointers are device addresses (dst / signal normally on a peer); `ticket` is a zero-initialised local uint32. */
typedef struct {
    void* lo_dst; void* lo_signal;      /* NULL lo_dst / hi_dst: no neighbour on that side (global face) */
    void* hi_dst; void* hi_signal;
    int64_t bytes;                      /* multiple of 16 */
    void* ticket;
} mode_halo_push_t;
typedef struct {
    int32_t n;                          /* destinations, <= 8 */
    void* dst[8];
    void* signal[8];
    void* ticket;
} mode_peer_push_t;
typedef struct {
    const void* slots;                  /* [world][count] doubles, local */
    int32_t world;
    const void* signal;                 /* local uint32 counter the producers increment */
    void* expect;                       /* local uint32 */
} mode_peer_gather_t;

/* Extended form of mode_conv3d (mode_conv3d == mode_conv3d_ex with opts = NULL).  Two things the plain call cannot say:
 *   (1) a HALOED input (D-sharded slabs): x has Dx >= D planes and output plane q is centred on
 *       input plane q + x_off, y[q] = sum_kd w[kd] * x[q + x_off + kd - 2]; input planes outside [0, Dx) are zero (the conv's
 *       padding at the GLOBAL faces).  Only the D owned output planes are computed -- no redundant halo-plane outputs.
 *       Requires x_off >= 0 and x_off + D <= Dx.
 *   (2) a fused epilogue: y = relu?(acc * out_scale * ep_scale[c] + ep_shift[c]) -- eval-mode BatchNorm3d + ReLU
 *       (RepMode.py:209-212) folded into K2 so y is written once and never re-read -- and/or an fp16 copy of the result
 *       (times y16_scale, saturated to +-65504) written into plane y16_off.. of a [N,Dy16,H,W,Nout] buffer: the next
 *       conv's operand, halo planes left for the exchange.  y may be NULL when y16 is given. */
typedef struct {
    int32_t Dx, x_off;              /* 0, 0 = no halo (Dx = D) */
    const float* ep_scale;          /* [Nout] or NULL */
    const float* ep_shift;          /* [Nout] or NULL */
    int32_t relu;
    void* y16;                      /* fp16 result copy or NULL */
    int32_t Dy16, y16_off;          /* planes of the y16 buffer (0 = D) and the plane output plane 0 lands in */
    float y16_scale;                /* 0 = 1 */
    const mode_peer_push_t* stats_push;   /* HOST pointer or NULL: the last CTA broadcasts bn_sums (2*Nout doubles) */
    void* splitk_ws;                /* device scratch of mode_conv3d_workspace_bytes(...) bytes (16-byte aligned) or NULL */
    int64_t splitk_ws_bytes;
} mode_conv_opts_t;
/* Deep, small-volume layers (the 8x32x32 .. 2x8x8 levels of the U-Net: a handful of 128-voxel tiles against K up to 512)
 * run the tcgen05 conv SPLIT ALONG K over the SMs: every CTA accumulates a slice of the 32-channel chunks into a partial
 * result and a second kernel sums the partials in fixed order and applies the epilogue.  Returns the scratch bytes such a
 * call needs (0 = the shape runs unsplit; then splitk_ws may be NULL).  A call that needs scratch and gets none fails. */
int64_t mode_conv3d_workspace_bytes(int32_t N, int32_t D, int32_t H, int32_t W, int32_t K, int32_t Nout, mode_dtype_t x_dtype);
int mode_conv3d_ex(const void* x, mode_dtype_t x_dtype, const void* w, const int32_t* sample_u, float* y,
                   int32_t N, int32_t D, int32_t H, int32_t W, int32_t K, int32_t Nout, float out_scale,
                   const float* out_scale_dev, double* bn_sums, int32_t stat_d_lo, int32_t stat_d_hi, int32_t impl,
                   const mode_conv_opts_t* opts_host, void* stream);

/* ---- K4: wgrad ------------------------------------------------------------------------------------------
 * d_weff[n][tap][o][i] = out_scale * sum_p dy[n][p][o] * x[n][p + tap - 2][i]   (autograd of RepMode.py:207).
 * x [N,D,H,W,Ci], dy [N,D,H,W,Co] (same dtype), d_weff fp32 (overwritten).
 *   impl: 0 = auto, 1 = SIMT fp32, 2 = tcgen05 (fp16 operands, Ci % 32 == 0, Co % 32 == 0, W % 8 == 0; the
 *   split-tap kernel, 7 MMAs per K step), 3 = force the stacked-tap kernel (10 MMAs per K step), 5 = force the
 *   split-tap kernel, 6 = its experimental deep-tile variant.  workspace:
 *   mode_conv3d_wgrad_workspace_bytes(..., same impl) bytes, 16-byte aligned.
 */
int64_t mode_conv3d_wgrad_workspace_bytes(int32_t N, int32_t D, int32_t H, int32_t W, int32_t Ci, int32_t Co,
                                          int32_t impl);
int mode_conv3d_wgrad(const void* x, const void* dy, mode_dtype_t dtype, float* d_weff, int32_t N, int32_t D,
                      int32_t H, int32_t W, int32_t Ci, int32_t Co, float out_scale, const float* out_scale_dev,
                      void* workspace, int32_t impl, void* stream);

/* Haloed form (D-sharded slabs): dy holds the D OWNED planes, x has Dx planes and dy plane p is centred on x plane p + x_off:
 *   d_weff[n][tap][o][i] = out_scale * sum_{p in [0,D)} dy[n][p][o] * x[n][p + x_off + kd - 2][..]; x planes outside [0, Dx)
 * are zero.  Every rank then holds the partial gradient of its own planes; the caller sums them (gradient all-reduce).
 * Dx = 0 means Dx = D, x_off = 0.  tcgen05: the deep-tile kernel only (impl 0 / 2 / 6). */
int mode_conv3d_wgrad_ex(const void* x, const void* dy, mode_dtype_t dtype, float* d_weff, int32_t N, int32_t D,
                         int32_t H, int32_t W, int32_t Ci, int32_t Co, float out_scale, const float* out_scale_dev,
                         void* workspace, int32_t impl, int32_t Dx, int32_t x_off, void* stream);

/* ---- BatchNorm3d + ReLU on NDHWC fp32 (RepMode.py:146-149,212) ------------------------------------------
 * mode_bn_stats:    sums[2C] (double, must be zeroed by the caller) += sum / sum of squares over M rows.
 * mode_bn_finalize: from sums -> mean, invstd (biased var, eps), scale = gamma*invstd,
 *                   shift = beta - mean*scale; updates running_mean/var (momentum, unbiased var) if non-NULL.
 * mode_bn_apply_relu: out = [relu](y*scale + shift); optional fp16 copy out_f16 scaled by f16_scale.
 * mode_bn_relu_bwd: given y, dout, gamma, beta, mean, invstd: dgamma, dbeta, dy as fp32 and/or as fp16
 *                   scaled by a power of two chosen on the device from a per-channel bound gathered in the
 *                   reduction pass; dy_scale2 (device float[2]) receives {scale, 1/scale}.
 *                   workspace: mode_bn_bwd_workspace_bytes(C).
 */
int64_t mode_bn_bwd_workspace_bytes(int32_t C);
/* Plane bookkeeping of a D-sharded slab [N][D][rows_per_plane][C] (NULL = whole tensor owned and valid):
 * statistics and the mean terms of the backward cover OWNED planes [own_lo, own_hi); planes outside
 * [valid_lo, valid_hi) lie beyond the GLOBAL volume: they are the next conv's zero padding, so the forward
 * writes zeros there and the backward treats their gradient as zero.  m_global = voxels of the global tensor
 * per channel (the statistics' divisor after the caller's all-reduce of the sums). */
typedef struct {
    int64_t rows_per_plane;
    int32_t D, own_lo, own_hi, valid_lo, valid_hi;
    int64_t m_global;
} mode_planes_t;

int mode_bn_stats(const float* y, int64_t M, int32_t C, double* sums, void* stream);
int mode_bn_finalize(const double* sums, int64_t M, int32_t C, const float* gamma, const float* beta, float eps,
                     float momentum, float* mean, float* invstd, float* scale, float* shift,
                     float* running_mean, float* running_var, void* stream);
int mode_bn_apply_relu(const float* y, int64_t M, int32_t C, const float* scale, const float* shift, int32_t relu,
                       float* out, void* out_f16, float f16_scale, const mode_planes_t* planes_host, void* stream);
int mode_bn_relu_bwd(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                     const float* beta, const float* mean, const float* invstd, float* dgamma, float* dbeta,
                     float* dy, void* dy_f16, float* dy_scale2, void* workspace, void* stream);
/* The same in two halves, for D-sharded slabs: _reduce leaves the LOCAL sums {sum dz, sum dz*xhat} (double[2C])
 * at the start of `workspace`; the caller all-reduces them across ranks; _apply finishes (dgamma/dbeta receive
 * the sums found in the workspace; mean terms use planes->m_global). planes_host may be NULL. */
int mode_bn_relu_bwd_reduce(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                            const float* beta, const float* mean, const float* invstd,
                            const mode_planes_t* planes_host, void* workspace, void* stream);
int mode_bn_relu_bwd_apply(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                           const float* beta, const float* mean, const float* invstd, float* dgamma, float* dbeta,
                           float* dy, void* dy_f16, float* dy_scale2, const mode_planes_t* planes_host,
                           void* workspace, void* stream);

/* Launch-structure forms (same arithmetic, fewer kernel boundaries / copies around the BatchNorm of a training step):
 *   mode_bn_finalize_apply_relu: mode_bn_finalize + mode_bn_apply_relu as ONE kernel -- every block derives scale / shift
 *                                from `sums` (M_stat = the statistics' divisor), block 0 writes mean / invstd / scale /
 *                                shift and updates the running statistics; y has M rows.
 *   mode_bn_relu_bwd_reduce_v2:  workspace_is_zero != 0: the caller has ALREADY zeroed the workspace (e.g. during the
 *                                forward, off the critical path): no memset in front of the reduction.
 *   mode_bn_relu_bwd_apply_v2:   mode_bn_relu_bwd_apply with a row map for dout.
 * mode_rowmap_t says where row r of the tensor the kernel WALKS (y) lives in the second tensor (out / dout):
 *   d2s != 0: y is the [voxels][8][C] result of the transposed stride-2 conv as a GEMM (ConvTranspose3d(k=2, s=2),
 *             fnet/nn_modules/RepMode.py:97-101; rows ordered (n, d, h, w, kd, kh, kw), D/H/W = the LOW-resolution grid) and
 *             the mapped tensor is the NDHWC volume [N][2D][2H][2W]: the depth-to-space scatter (forward) / gather
 *             (backward) happens inside the BatchNorm kernels instead of as a permute copy;
 *   pitch:    floats between rows of the mapped tensor (0 = C): a channel range of a wider tensor, e.g. the gradient of
 *             one half of the decoder's concatenated input (RepMode.py:106).
 * NULL = identity.  mode_bn_relu_bwd_apply derives the fp16 scale of dy inside the apply kernel (no separate scale launch)
 * unless a gather descriptor is given. */
typedef struct {
    int32_t d2s;
    int32_t D, H, W;
    int64_t pitch;
} mode_rowmap_t;
int mode_bn_finalize_apply_relu(const double* sums, int64_t M_stat, int32_t C, const float* gamma, const float* beta,
                                float eps, float momentum, float* mean, float* invstd, float* scale, float* shift,
                                float* running_mean, float* running_var, const float* y, int64_t M, int32_t relu,
                                float* out, void* out_f16, float f16_scale, const mode_planes_t* planes_host,
                                const mode_rowmap_t* out_map_host, void* stream);
int mode_bn_relu_bwd_reduce_v2(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                               const float* beta, const float* mean, const float* invstd,
                               const mode_planes_t* planes_host, void* workspace, int32_t workspace_is_zero,
                               const mode_rowmap_t* dout_map_host, void* stream);
int mode_bn_relu_bwd_apply_v2(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                              const float* beta, const float* mean, const float* invstd, float* dgamma, float* dbeta,
                              float* dy, void* dy_f16, float* dy_scale2, const mode_planes_t* planes_host,
                              void* workspace, const mode_rowmap_t* dout_map_host, void* stream);

/* Fused-exchange forms (peer memory, D-sharded slabs).  NULL descriptors give the plain behaviour.
 *   mode_bn_finalize_ex:        `gather` != NULL: the statistics are the rank-ordered sum of the gathered slots (sums ignored).
 *   mode_bn_relu_bwd_reduce_ex: `push` != NULL: the last block broadcasts the 4*C-double workspace vector.
 *   mode_bn_relu_bwd_apply_ex:  `gather`: wait + sum into the workspace first (replaces the caller's all-reduce);
 *                               `halo`: boundary planes of the produced dy (the fp16 copy when dy_f16 != NULL, else the fp32
 *                               tensor) are also stored into the neighbours' halo planes. */
int mode_bn_finalize_ex(const double* sums, int64_t M, int32_t C, const float* gamma, const float* beta, float eps,
                        float momentum, float* mean, float* invstd, float* scale, float* shift, float* running_mean,
                        float* running_var, const mode_peer_gather_t* gather_host, void* stream);
int mode_bn_relu_bwd_reduce_ex(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                               const float* beta, const float* mean, const float* invstd,
                               const mode_planes_t* planes_host, void* workspace, const mode_peer_push_t* push_host,
                               void* stream);
int mode_bn_relu_bwd_apply_ex(const float* y, const float* dout, int64_t M, int32_t C, const float* gamma,
                              const float* beta, const float* mean, const float* invstd, float* dgamma, float* dbeta,
                              float* dy, void* dy_f16, float* dy_scale2, const mode_planes_t* planes_host,
                              void* workspace, const mode_peer_gather_t* gather_host, const mode_halo_push_t* halo_host,
                              void* stream);

/* ---- operand staging -----------------------------------------------------------------------------------
 * fp32 -> fp16 (round to nearest even), value * scale * (scale_dev ? *scale_dev : 1); n elements. */
int mode_cast_f16(const float* src, void* dst_f16, int64_t n, float scale, const float* scale_dev, void* stream);
/* ... with the boundary planes of dst also stored into the neighbours' halo planes (halo_host may be NULL). */
int mode_cast_f16_ex(const float* src, void* dst_f16, int64_t n, float scale, const float* scale_dev,
                     const mode_halo_push_t* halo_host, void* stream);
/* fp32 [rows][c] -> fp16 [rows][c_pad], zero channels appended (c_pad % 8 == 0; the stem layer's Ci = 1 -> 32 operand):
 * replaces the host-side zero-pad copy in front of F.conv3d's operand (fnet/nn_modules/RepMode.py:207). */
int mode_cast_f16_pad(const float* src, void* dst_f16, int64_t rows, int32_t c, int32_t c_pad, void* stream);
/* fp32 [rows][ca] ++ fp32 [rows][cb] -> fp16 [rows][ca + cb] (saturated): the decoder's torch.cat((x_skip, x), 1)
 * (fnet/nn_modules/RepMode.py:106) folded into the staging of the conv operand; ca, cb multiples of 4. */
int mode_cast_f16_cat(const float* a, int32_t ca, const float* b, int32_t cb, void* dst_f16, int64_t rows, void* stream);
/* amax[0] = max(amax[0], max |src|) (amax must be zero-initialised by the caller; device scalar). */
int mode_amax(const float* src, int64_t n, float* amax, void* stream);
/* Same over k <= 8 tensors in one launch (srcs_host / counts_host are HOST arrays of device pointers / sizes). */
int mode_amax_multi(const float* const* srcs_host, const int64_t* counts_host, int32_t k, float* amax, void* stream);
/* scale2[0] = 2^floor(log2(target / amax[0])) (1 if amax is 0), scale2[1] = 1 / scale2[0]; all device. */
int mode_f16_scale(const float* amax, float target, float* scale2, void* stream);

/* ---- peer-memory exchange over NVLink / NVSwitch (D-sharded volumes) ---------------------------
 * The reference has no working multi-GPU data path (fnet/fnet_model.py:40-44 is torch.nn.DataParallel); these entry points
 * implement the exchange steps of the D-axis split -- halo planes, BatchNorm partial sums, gradient sum -- as plain stores
 * into the neighbour's memory plus a counter, instead of one NCCL launch each.  Pointers named *_peer are addresses inside
 * ANOTHER GPU's memory mapped into this process (CUDA IPC; the host side maps them once), everything else is local.
 * Counters (`signal`, `expect`, `ticket`) are zero-initialised uint32 in device memory and only ever increase, so a captured
 * CUDA graph can replay the step.  A wait that is not satisfied within ~10 s raises the device error flag (codes 41 / 42,
 * mode_poll_error) instead of hanging.
 *   mode_peer_put:       n <= 8 segments of `bytes` (multiple of 16) each: dst[i] <- src[i]; once ALL of them are written,
 *                        *signal[i] += 1 for every i (release, system scope).  src/dst/signal are HOST arrays of device
 *                        pointers; dst[i] and signal[i] normally live on a peer.  `ticket`: a local uint32 scratch counter.
 *   mode_peer_wait:      *expect += add; wait until *signal >= *expect (acquire).  One thread; stream-ordered.
 *   mode_peer_sum_slots: one-shot all-reduce tail: wait until `world` producers have signalled, then
 *                        out[i] = slots[0][i] + slots[1][i] + ... in rank order (deterministic); double or float.
 *   mode_peer_enable_access: cudaDeviceEnablePeerAccess(current device -> peer_device); no-op when already enabled. */
int mode_peer_enable_access(int32_t peer_device);
/* Exchange arena of one rank: a zero-filled cudaMalloc block on the current device plus its 64-byte CUDA IPC handle
 * (`handle64_host`, to be sent to the other ranks by any host channel); _open maps another rank's arena into THIS process
 * with the current device as the accessing device (peer mapping over NVLink; also valid when both ranks share one GPU);
 * _close unmaps it, _free releases the owner's block.  Set-up calls: they synchronise. */
int mode_peer_arena_alloc(int64_t bytes, void** ptr_out, void* handle64_host);
int mode_peer_arena_open(const void* handle64_host, void** ptr_out);
int mode_peer_arena_close(void* ptr);
int mode_peer_arena_free(void* ptr);
int mode_peer_put(const void* const* src_host, void* const* dst_peer_host, void* const* signal_peer_host, int32_t n,
                  int64_t bytes, void* ticket, void* stream);
int mode_peer_wait(const void* signal, void* expect, int32_t add, void* stream);
int mode_peer_sum_slots(const void* slots, int32_t world, int64_t n, int32_t is_double, void* out, const void* signal,
                        void* expect, void* ticket, void* stream);

/* ---- Model.predict glue (fnet/fnet_model.py:149-223) ------------------------------------------------
 * Sliding-window inference: the network runs on batches of overlapping patches and the predictions are blended with a
 * Gaussian importance map.  mode_blend_accumulate replaces the per-patch sliced updates of reference :207-214
 *     pred_sum[window] += pred_patch * gauss;  weight_sum[window] += gauss
 * for a whole batch of P <= 64 patches in one pass (deterministic, patches added in batch order); mode_blend_finalize is
 * reference :220, out = pred_sum / weight_sum.
 *   pred       [P][C][pd][ph][pw] fp32 dense (NCDHW: what Net.forward returns for C = 1)
 *   starts_dev [P][3] int32 window origins (d, h, w), DEVICE memory
 *   gauss      [*][gh][gw] fp32 importance map, indexed with the patch-local coordinate (gh >= ph, gw >= pw)
 *   pred_sum   [C][D][H][W] fp32, weight_sum [D][H][W] fp32 (kept once: it is identical for every channel) */
int mode_blend_accumulate(const float* pred, const int32_t* starts_dev, const float* gauss, float* pred_sum,
                          float* weight_sum, int32_t P, int32_t C, int32_t pd, int32_t ph, int32_t pw, int32_t gh, int32_t gw,
                          int32_t D, int32_t H, int32_t W, void* stream);
int mode_blend_finalize(const float* pred_sum, const float* weight_sum, float* out, int32_t C, int64_t voxels, void* stream);

/* ---- Model.do_train_iter glue (fnet/fnet_model.py:96-132) --------------------------------------------
 * torch.optim.Adam (the optimizer built at fnet_model.py:55; amsgrad = False, maximize = False) over ALL parameter tensors in
 * one multi-tensor launch, with the GradScaler hand-shake of fnet_model.py:111-113 on the device: gradients are divided by
 * *grad_scale_dev (NULL = 1) and the whole step -- including the per-tensor step counters -- is skipped when *found_inf_dev is
 * non-zero (NULL = never), so scaler.step() needs no host synchronisation.
 *   tensors_dev [ntensors] x {float* p, const float* g, float* m, float* v, float* step, int64 n}   (48-byte records, device)
 *   chunks_dev  [nchunks]  x {int32 tensor, int32 0, int64 offset}: every tensor cut into pieces of mode_adam_chunk_elems()
 *   step is torch's per-parameter fp32 scalar state['step'] on the device; it is advanced by 1 BEFORE the update, as in torch. */
int64_t mode_adam_chunk_elems(void);
int mode_adam_step(const void* tensors_dev, int32_t ntensors, const void* chunks_dev, int32_t nchunks, double lr, double beta1,
                   double beta2, double eps, double weight_decay, const float* grad_scale_dev, const float* found_inf_dev,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REPMODE_B200_H_ */
