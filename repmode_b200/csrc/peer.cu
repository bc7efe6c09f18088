// Peer-memory exchange steps of the D-sharded MoDE-conv path over NVLink / NVSwitch: halo planes,
// BatchNorm partial sums and parameter gradients move between the GPUs of one box by plain stores into the neighbour's
// memory (CUDA IPC mappings, set up by the host side) plus a release/acquire counter -- no NCCL launch on the data path.
// The reference has no multi-GPU data path to mirror (fnet/fnet_model.py:40-44 is torch.nn.DataParallel); the spec is
// one halo exchange per stage, BatchNorm statistics over owned voxels, gradient sum.
//
// Protocol (all counters are monotonically increasing uint32 in device memory, so a CUDA graph can replay the step):
//   producer: writes the payload into the consumer's buffer, __threadfence_system(), then the LAST block of the launch (a
//             ticket counter in the producer's own memory) does atomicAdd_system(consumer_signal, 1).
//   consumer: `expect` (its own device counter) += number of producers; spins until signal >= expect with acquire loads.
//             A wait that does not complete within ~10 s raises the device error flag (code 41) instead of hanging the box.
// Buffers are reused every step; the write-after-read hazard is closed by the step's own data dependencies (nobody starts
// step s+1 before the gradient exchange of step s, which every rank enters only after its last read of step s' halos).
// HBM/NVLink-bound copies: 16-byte vector accesses, grid sized to the payload.
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "peer.cuh"

namespace mode {

int* device_error_flag();   // mode_abi.cu

// ---- put: up to 8 (src -> dst) segments of equal size, one signal per segment ---------------------------------------
struct PutList { const uint4* src[8]; uint4* dst[8]; uint32_t* sig[8]; int n; long long vec; };
__global__ void __launch_bounds__(256) peer_put_kernel(PutList L, uint32_t* ticket) {
    const int seg = blockIdx.y;
    const uint4* __restrict__ s = L.src[seg];
    uint4* __restrict__ d = L.dst[seg];
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < L.vec; i += 4 * stride) {      // 4 loads in flight
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < L.vec) v[u] = s[i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < L.vec) d[i + u * stride] = v[u];
    }
    if (finish_block_is_last(ticket) && threadIdx.x == 0) signal_all(L.sig, L.n);
}

__global__ void peer_wait_kernel(const uint32_t* signal, uint32_t* expect, uint32_t add, int* error_flag) {
    const uint32_t target = *expect + add;
    *expect = target;
    if (!wait_signal(signal, target)) atomicExch(error_flag, 41);
}

// ---- one-shot all-reduce: every rank holds slots[world][n]; sum in rank order (deterministic) after all arrived ---------
template <typename T>
__global__ void __launch_bounds__(256) peer_sum_slots_kernel(const T* __restrict__ slots, int world, long long n,
                                                             T* __restrict__ out, const uint32_t* signal, uint32_t* expect,
                                                             uint32_t* ticket, int* error_flag) {
    __shared__ uint32_t s_target;
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        s_target = *expect + (uint32_t)world;
        s_ok = wait_signal(signal, s_target) ? 1 : 0;
        if (!s_ok) atomicExch(error_flag, 42);
    }
    __syncthreads();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        T a = slots[i];
        for (int r = 1; r < world; ++r) a += slots[(long long)r * n + i];
        out[i] = a;
    }
    // the last block to finish advances `expect` (all blocks read the same old value above)
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *ticket = 0; *expect = s_target; }
    }
}

}  // namespace mode

using namespace mode;

// ---- exchange arena: one cudaMalloc block per rank, exported / imported through CUDA IPC ------------------------------
// The import happens with the IMPORTING rank's device current and cudaIpcMemLazyEnablePeerAccess, so the mapping is a
// peer (NVLink) mapping usable by kernels of this device.  (r2d2: a mapping opened under the OWNER's device -- what
// torch's storage sharing does -- faults when a kernel of another GPU stores through it.)
extern "C" int mode_peer_arena_alloc(int64_t bytes, void** ptr_out, void* handle64_host) {
    if (bytes <= 0 || !ptr_out || !handle64_host) MODE_FAIL("mode_peer_arena_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    MODE_CUDA(cudaMalloc(&p, (size_t)bytes));
    MODE_CUDA(cudaMemset(p, 0, (size_t)bytes));
    MODE_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        MODE_FAIL("mode_peer_arena_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    }
    memcpy(handle64_host, &h, sizeof(h));
    *ptr_out = p;
    return 0;
}

extern "C" int mode_peer_arena_open(const void* handle64_host, void** ptr_out) {
    if (!handle64_host || !ptr_out) MODE_FAIL("mode_peer_arena_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    void* p = nullptr;
    MODE_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
    return 0;
}

extern "C" int mode_peer_arena_close(void* ptr) {
    if (ptr) MODE_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int mode_peer_arena_free(void* ptr) {
    if (ptr) MODE_CUDA(cudaFree(ptr));
    return 0;
}

extern "C" int mode_peer_enable_access(int32_t peer_device) {
    int dev = 0;
    MODE_CUDA(cudaGetDevice(&dev));
    if (peer_device == dev) return 0;
    int can = 0;
    MODE_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) MODE_FAIL("mode_peer_enable_access: device %d cannot access device %d (no NVLink / P2P path)", dev, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    MODE_CUDA(e);
    return 0;
}

extern "C" int mode_peer_put(const void* const* src_host, void* const* dst_host, void* const* signal_host, int32_t n,
                             int64_t bytes, void* ticket, void* stream) {
    if (!src_host || !dst_host || !signal_host || !ticket || n <= 0 || n > 8 || bytes <= 0 || (bytes & 15))
        MODE_FAIL("mode_peer_put: bad arguments (n=%d bytes=%lld; bytes must be a multiple of 16)", n, (long long)bytes);
    PutList L;
    L.n = n;
    L.vec = bytes / 16;
    for (int i = 0; i < 8; ++i) {
        L.src[i] = i < n ? (const uint4*)src_host[i] : nullptr;
        L.dst[i] = i < n ? (uint4*)dst_host[i] : nullptr;
        L.sig[i] = i < n ? (uint32_t*)signal_host[i] : nullptr;
        if (i < n && (!L.src[i] || !L.dst[i] || !L.sig[i] || ((uintptr_t)L.src[i] & 15) || ((uintptr_t)L.dst[i] & 15)))
            MODE_FAIL("mode_peer_put: null or misaligned segment %d", i);
    }
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(L.vec, 256 * 4), 2 * sm_count() / n + 1));
    peer_put_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(L, (uint32_t*)ticket);
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_peer_wait(const void* signal, void* expect, int32_t add, void* stream) {
    if (!signal || !expect || add <= 0) MODE_FAIL("mode_peer_wait: bad arguments");
    int* ef = device_error_flag();
    if (!ef) MODE_FAIL("mode_peer_wait: could not allocate the device error flag");
    peer_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const uint32_t*)signal, (uint32_t*)expect, (uint32_t)add, ef);
    MODE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mode_peer_sum_slots(const void* slots, int32_t world, int64_t n, int32_t is_double, void* out,
                                   const void* signal, void* expect, void* ticket, void* stream) {
    if (!slots || !out || !signal || !expect || !ticket || world <= 0 || n <= 0) MODE_FAIL("mode_peer_sum_slots: bad arguments");
    int* ef = device_error_flag();
    if (!ef) MODE_FAIL("mode_peer_sum_slots: could not allocate the device error flag");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), sm_count()));
    if (is_double)
        peer_sum_slots_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)slots, world, n, (double*)out,
                                                                            (const uint32_t*)signal, (uint32_t*)expect,
                                                                            (uint32_t*)ticket, ef);
    else
        peer_sum_slots_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)slots, world, n, (float*)out,
                                                                           (const uint32_t*)signal, (uint32_t*)expect,
                                                                           (uint32_t*)ticket, ef);
    MODE_LAUNCH_CHECK();
    return 0;
}
