// Device-side pieces of the peer-memory exchange protocol (csrc/peer.cu has the protocol description), shared by the
// kernels that FUSE an exchange step into their own work: the operand-staging and BatchNorm-backward kernels store their
// boundary planes straight into the neighbours' halo planes, the conv / BatchNorm-reduce kernels broadcast their
// per-channel sums from the last block to finish, and the consumers wait for the counters themselves.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/repmode_b200.h"

namespace mode {

// device copies of the ABI descriptors (resolved pointers)
struct HaloPush {
    uint8_t* lo_dst; uint32_t* lo_sig;      // lower neighbour: receives the FIRST `bytes` of the tensor being written
    uint8_t* hi_dst; uint32_t* hi_sig;      // upper neighbour: receives the LAST `bytes`
    long long bytes;
    uint32_t* ticket;
    __host__ __device__ bool on() const { return lo_dst != nullptr || hi_dst != nullptr; }
};
struct PeerPush {
    int n;
    void* dst[8];
    uint32_t* sig[8];
    uint32_t* ticket;
};
struct PeerGather {
    const void* slots; int world; const uint32_t* signal; uint32_t* expect;
    __host__ __device__ bool on() const { return slots != nullptr; }
};

static inline HaloPush to_halo_push(const mode_halo_push_t* p) {
    HaloPush h{nullptr, nullptr, nullptr, nullptr, 0, nullptr};
    if (p) {
        h.lo_dst = (uint8_t*)p->lo_dst; h.lo_sig = (uint32_t*)p->lo_signal;
        h.hi_dst = (uint8_t*)p->hi_dst; h.hi_sig = (uint32_t*)p->hi_signal;
        h.bytes = p->bytes; h.ticket = (uint32_t*)p->ticket;
    }
    return h;
}
static inline PeerPush to_peer_push(const mode_peer_push_t* p) {
    PeerPush q;
    q.n = p ? p->n : 0;
    for (int i = 0; i < 8; ++i) { q.dst[i] = (p && i < p->n) ? p->dst[i] : nullptr; q.sig[i] = (p && i < p->n) ? (uint32_t*)p->signal[i] : nullptr; }
    q.ticket = p ? (uint32_t*)p->ticket : nullptr;
    return q;
}
static inline PeerGather to_peer_gather(const mode_peer_gather_t* p) {
    PeerGather g{nullptr, 0, nullptr, nullptr};
    if (p) { g.slots = p->slots; g.world = p->world; g.signal = (const uint32_t*)p->signal; g.expect = (uint32_t*)p->expect; }
    return g;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t peer_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *signal has reached `target` (wrap-safe); false after ~10 s (the caller raises the device error flag).
__device__ __forceinline__ bool wait_signal(const uint32_t* signal, uint32_t target) {
    const uint64_t t0 = peer_timer_ns();
    while ((int32_t)(ld_acquire_sys(signal) - target) < 0) {
        __nanosleep(64);
        if (peer_timer_ns() - t0 > 10000000000ull) return false;
    }
    return true;
}

// Called by EVERY thread of EVERY block at the end of a kernel that stored a payload into peer memory.  Each block fences its
// stores and takes a ticket; the block that takes the last ticket resets the counter (the next launch / graph replay starts
// from 0) and increments the consumers' counters.  Returns true in the threads of that last block.
__device__ __forceinline__ bool finish_block_is_last(uint32_t* ticket) {
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        const uint32_t total = gridDim.x * gridDim.y * gridDim.z;
        s_last = (atomicAdd(ticket, 1u) == total - 1) ? 1 : 0;
        if (s_last) *ticket = 0;
    }
    __syncthreads();
    return s_last != 0;
}
__device__ __forceinline__ void signal_all(uint32_t* const* sig, int n) {
    __threadfence_system();
    for (int i = 0; i < n; ++i)
        if (sig[i] != nullptr) atomicAdd_system(sig[i], 1u);
}
__device__ __forceinline__ void halo_finish(const HaloPush& hp) {
    if (finish_block_is_last(hp.ticket) && threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        uint32_t* sg[2] = {hp.lo_dst ? hp.lo_sig : nullptr, hp.hi_dst ? hp.hi_sig : nullptr};
        signal_all(sg, 2);
    }
}
// A vector store that also lands in the neighbours' halo planes when it falls into a boundary region of the tensor.
template <typename V>
__device__ __forceinline__ void halo_store(const HaloPush& hp, long long off, long long total, V v) {
    if (hp.lo_dst != nullptr && off < hp.bytes) *reinterpret_cast<V*>(hp.lo_dst + off) = v;
    if (hp.hi_dst != nullptr && off >= total - hp.bytes) *reinterpret_cast<V*>(hp.hi_dst + (off - (total - hp.bytes))) = v;
}
// Last block of a kernel broadcasts `count` doubles (complete in `local` once every block has passed the ticket) into slot
// dst[i] of every destination and signals.  All threads of all blocks call it; blockDim may be anything.
__device__ __forceinline__ void push_vector_from_last_block(const double* local, int count, const PeerPush& pp) {
    if (!finish_block_is_last(pp.ticket)) return;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    for (int i = tid; i < count; i += nthr) {
        const double v = __ldcg(local + i);                 // the other blocks' atomics live in L2
        for (int q = 0; q < pp.n; ++q) reinterpret_cast<double*>(pp.dst[q])[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) signal_all(pp.sig, pp.n);
}
// One thread waits for `world` producers, everybody then reads: sum over the slots in rank order (deterministic).
// Returns false on timeout.  The caller advances *expect (one thread, after all readers are done).
__device__ __forceinline__ bool gather_wait(const PeerGather& g, uint32_t& target) {
    __shared__ int s_ok;
    __shared__ uint32_t s_target;
    if (threadIdx.x == 0) {
        s_target = *g.expect + (uint32_t)g.world;
        s_ok = wait_signal(g.signal, s_target) ? 1 : 0;
    }
    __syncthreads();
    target = s_target;
    return s_ok != 0;
}
#endif

}  // namespace mode
