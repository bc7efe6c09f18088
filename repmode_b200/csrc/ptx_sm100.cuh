// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld / fences) and the UMMA shared-memory + instruction descriptors.
// Everything here is architecture plumbing shared by the conv kernels and the hardware probe.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `cta_rank` of this cluster (release at cluster scope).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta_rank) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}\n" ::"r"(bar), "r"(cta_rank)
        : "memory");
}
// try_wait with acquire at cluster scope: pairs with mbar_arrive_remote from the peer CTA.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Time-bounded wait: returns false if the barrier has not flipped within ~2 s of wall clock (the kernel
// then bails out and raises an error flag instead of hanging the GPU box). max_spins bounds the number of
// polls between clock checks only.
static __device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity) {
    const uint64_t t0 = globaltimer_ns();
    while (true) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i)
            if (mbar_try_wait(bar, parity)) return true;
        if (globaltimer_ns() - t0 > 2000000000ull) return false;
    }
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, uint32_t /*unused*/ = 0) {
    if (mbar_try_wait(bar, parity)) return true;
    return mbar_wait_slow(bar, parity);
}

static __device__ __noinline__ bool mbar_wait_cluster_slow(uint32_t bar, uint32_t parity) {
    const uint64_t t0 = globaltimer_ns();
    while (true) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i)
            if (mbar_try_wait_cluster(bar, parity)) return true;
        if (globaltimer_ns() - t0 > 2000000000ull) return false;
    }
}
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return true;
    return mbar_wait_cluster_slow(bar, parity);
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// Plain (non-tensor) bulk copy global -> shared, completes on an mbarrier. bytes % 16 == 0.
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// ---- CTA pair (cta_group::2): both CTAs of a 2-CTA cluster allocate / free (same warp id), the leader
// (cluster rank 0) issues the MMAs with M = 256 (128 rows from each CTA) and B split N/2 + N/2 over the pair.
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit of the pair's outstanding MMAs: arrives on the mbarrier at the same shared offset in every CTA of `mask`.
__device__ __forceinline__ void mma_commit_pair(uint32_t bar, uint32_t cta_mask) {
    asm volatile(
        "{\n\t"
        ".reg .b16 lo, hi;\n\t"
        "mov.b32 {lo, hi}, %1;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n\t"
        "}\n" ::"r"(bar), "r"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-uniform variants: executed by ALL lanes of a converged warp with identical (uniform) operands; one lane
// is elected inside the asm.  Keeping the C++ control flow uniform lets ptxas hold descriptors in uniform
// registers instead of the per-thread ELECT / R2UR.BROADCAST loop it emits inside an `if (lane == 0)` region.
__device__ __forceinline__ void mma_f16_ss_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                 uint32_t b_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred q, one;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.eq.u32 one, 1, 1;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, one;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}\n"
        :
        : "r"(bar)
        : "memory");
}

// CTA-pair versions for a warp-uniform issue loop (all 32 lanes run the loop on uniform values; `sel` is the
// per-lane result of ONE elect_one() taken by the caller, so no ELECT is paid per instruction).
__device__ __forceinline__ void mma_f16_ss_pair_sel(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                    uint32_t b_hi, uint32_t idesc, uint32_t sel) {
    asm volatile(
        "{\n\t"
        ".reg .pred q, one;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.eq.u32 one, 1, 1;\n\t"
        "setp.ne.u32 q, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, one;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(sel)
        : "memory");
}
__device__ __forceinline__ void mma_commit_pair_sel(uint32_t bar, uint32_t cta_mask, uint32_t sel) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        ".reg .b16 lo, hi;\n\t"
        "mov.b32 {lo, hi}, %1;\n\t"
        "setp.ne.u32 q, %2, 0;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n\t"
        "}\n" ::"r"(bar), "r"(cta_mask), "r"(sel)
        : "memory");
}
// Single-CTA versions of the same idea (wgrad_split.cu): the accumulate flag is a warp-uniform operand.
__device__ __forceinline__ void mma_f16_ss_sel(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate, uint32_t sel) {
    asm volatile(
        "{\n\t"
        ".reg .pred q, acc;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.u32 acc, %6, 0;\n\t"
        "setp.ne.u32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, acc;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(sel)
        : "memory");
}
__device__ __forceinline__ void mma_commit_sel(uint32_t bar, uint32_t sel) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}\n" ::"r"(bar), "r"(sel)
        : "memory");
}
// Warp-collective, time-bounded mbarrier wait whose RESULT is warp-uniform (vote), so that the code after it stays
// convergent and ptxas can keep descriptors in uniform registers.  cluster = acquire at cluster scope.
template <bool kCluster>
static __device__ __noinline__ bool mbar_wait_warp_slow(uint32_t bar, uint32_t parity) {
    const uint64_t t0 = globaltimer_ns();
    while (true) {
#pragma unroll 1
        for (int i = 0; i < 64; ++i) {
            const bool done = kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity);
            if (__all_sync(0xffffffffu, done)) return true;
        }
        if (__any_sync(0xffffffffu, globaltimer_ns() - t0 > 2000000000ull)) return false;
    }
}
template <bool kCluster>
__device__ __forceinline__ bool mbar_wait_warp(uint32_t bar, uint32_t parity) {
    const bool done = kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity);
    if (__all_sync(0xffffffffu, done)) return true;
    return mbar_wait_warp_slow<kCluster>(bar, parity);
}

// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread t of the warp reads lane base+t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// store 32 registers per thread into 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
          "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]),
          "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (64-bit), sm_100 format:
//   [0,14)  start address >> 4          [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4 [46,48) version = 1        [49,52) base offset
//   [52]    LBO mode = 0                [61,64) swizzle: 0 none, 2 128B, 4 64B, 6 32B
enum : uint32_t { SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };

__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                            uint32_t swizzle, uint32_t base_offset = 0) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(base_offset & 7) << 49;
    d |= static_cast<uint64_t>(swizzle & 7) << 61;
    return d;
}

// Instruction descriptor (32-bit) for kind::f16 / kind::tf32, dense, fp32 accumulate:
//   [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (0 f16, 1 bf16, 2 tf32)
//   [15] A major (0 K, 1 MN)  [16] B major     [17,23) N >> 3       [24,29) M >> 4
enum : uint32_t { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };
__host__ __device__ __forceinline__ uint32_t make_idesc(uint32_t fmt, uint32_t m, uint32_t n, uint32_t a_mn_major,
                                                        uint32_t b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (fmt & 7) << 7;
    d |= (fmt & 7) << 10;
    d |= (a_mn_major & 1) << 15;
    d |= (b_mn_major & 1) << 16;
    d |= ((n >> 3) & 0x3F) << 17;
    d |= ((m >> 4) & 0x1F) << 24;
    return d;
}

}  // namespace sm100
