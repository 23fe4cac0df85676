// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld), proxy fences.  Everything here is hand-written; no CUTLASS/CuTe types are used.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor"
// tables (also summarised in /opt/skills/guides/blackwell_cuda_programming.md).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace b200 {

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
#ifndef B200_MBAR_TIMEOUT_CYCLES
#define B200_MBAR_TIMEOUT_CYCLES (4000000000ll)  // ~2 s at 1.9 GHz
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ffu) == 0u && (clock64() - t0) > B200_MBAR_TIMEOUT_CYCLES) {
            printf("b200: mbarrier timeout block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}

// ----------------------------------------------------------------------------------------------
// proxy fences
// ----------------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (TMA store, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// 4-D box (channels, x, y, image) of a channels-last activation: coordinates may be NEGATIVE or past the end -- the
// out-of-bound part of the box is zero-filled, which is exactly the zero padding of a 3x3 convolution
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// Multicast load: the box lands at the SAME smem offset in every CTA of `cta_mask`, and each destination CTA's
// mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                                  uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
// ---- cta_group::2 (CTA-pair) forms --------------------------------------------------------------------------
// address of the same smem offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// arrive on a (possibly remote) mbarrier given by its shared::cluster address
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// The same without release semantics.  A cluster-scope RELEASE makes the warp drain every global store it has in flight
// first (ERRBAR + membar: 22 % of the GELU GEMM's stall samples in 2-SM mode).  Handing a TMEM accumulator back needs
// no memory ordering at all: the reads have completed (tcgen05.wait::ld) and tcgen05.fence::before_thread_sync orders
// them against the arrive; nothing is communicated through memory.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair: data lands in the issuing CTA's smem, the bytes are credited to the mbarrier
// at `bar_cluster_addr` (the pair leader's barrier)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 split over the pair: each CTA supplies its 128 rows of A and its half
// of B (N/2 rows) at the SAME smem offsets; issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mcast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// L2 prefetch of a tile (no smem destination): lets a later TMA load of the same bytes hit L2 instead of HBM
__device__ __forceinline__ void tma_prefetch_l2_3d(const void* tmap, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// L2 prefetch of a CONTIGUOUS byte range (16-byte aligned address and size): no smem destination, no completion to wait for
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(gptr)), "r"(bytes) : "memory");
}
// wait until at most 1 committed bulk store still has to read its smem source (two staging buffers in rotation)
__device__ __forceinline__ void tma_store_wait_read1() {
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the smem source of all committed bulk stores has been read (buffer reusable)
__device__ __forceinline__ void tma_store_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// wait until all committed bulk stores are complete (globally visible)
__device__ __forceinline__ void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ----------------------------------------------------------------------------------------------
// Whole warp must execute.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4      [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4     [46,48) version = 1 (Blackwell)
//   [49,52) base offset = 0         [52]    lbo mode = 0
//   [61,64) layout: 0 none, 1 128B_base32B, 2 SWIZZLE_128B, 4 SWIZZLE_64B, 6 SWIZZLE_32B
constexpr uint32_t kSwz128 = 2;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout & 7u) << 61;
    return d;
}
// descriptor + byte offset (>> 4): only the 14-bit start-address field changes (cheap on the single MMA-issuing lane)
__device__ __forceinline__ uint64_t desc_off(uint64_t base, uint32_t off16) {
    return (base & 0xffffffff00000000ull) | static_cast<uint64_t>(static_cast<uint32_t>(base) + off16);
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate:
//   [4,6) c_format = 1 (F32)  [7,10) a_format = 1 (BF16)  [10,13) b_format = 1 (BF16)
//   [15] a_major (0 = K-major, 1 = MN-major)  [16] b_major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n, uint32_t a_mn_major,
                                                       uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major & 1u) << 15) | ((b_mn_major & 1u) << 16) |
           ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// tcgen05: mma / commit / ld
// ----------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

// Same, but the arrive is delivered to the mbarrier at this offset in every CTA of `cta_mask` (cluster multicast).
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives lane (base_lane + i), columns [c, c+32).
// A warp may only touch TMEM lanes [32*(warp_id%4), 32*(warp_id%4)+32).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]: the A operand is read from TENSOR MEMORY (lane = row of A, 16-bit elements packed two
// per 32-bit column with K contiguous, K-major only); issued by ONE thread.  Used by the attention forward: the softmax
// warps write the bf16 probabilities back into the columns of S with tcgen05.st and O = P V consumes them from there,
// so P never travels through shared memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> TMEM: thread i of the warp writes lane (base_lane + i), 8 consecutive 32-bit columns starting at taddr's column
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Register re-distribution between the warpgroups of a CTA (all 4 warps of a warpgroup execute it together)
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// ----------------------------------------------------------------------------------------------
// numeric helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits), .y = hi
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// exact-erf GELU, evaluated in fp32 like ATen's GeluCUDAKernelImpl ("none" approximation): libm reference versions
__device__ __forceinline__ float gelu_erf_libm(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad_libm(float x) {
    const float kBeta = 0.39894228040143267794f;  // 1/sqrt(2*pi)   (M_2_SQRTPI * M_SQRT1_2 * 0.5)
    float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    float pdf = expf(-0.5f * x * x) * kBeta;
    return cdf + x * pdf;
}
// Fast erf for the GEMM epilogues (libm erff costs ~30 issue slots / element and made the FF1 GEMM epilogue-bound).
// Measured against ATen's fp32 GELU on 4M bf16 inputs ~ N(0, 1.5): relative L2 difference of the bf16 results
// 3.4e-7 (1-ulp flips only where |gelu| is negligible); derivative max abs error 1.8e-7.
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// 1 + erf(z) without cancellation: erfc(|z|) = t exp(-z^2 + P(t)), t = 1 / (1 + |z| / 2)  (Chebyshev fit of
// Numerical Recipes' erfcc, fractional error < 1.2e-7 everywhere), then 1 + erf(z) = z >= 0 ? 2 - erfc : erfc.
__device__ __forceinline__ float one_plus_erf(float z) {
    const float az = fabsf(z);
    const float t = fast_rcp(fmaf(0.5f, az, 1.0f));
    float p = fmaf(0.17087277f, t, -0.82215223f);
    p = fmaf(p, t, 1.48851587f);
    p = fmaf(p, t, -1.13520398f);
    p = fmaf(p, t, 0.27886807f);
    p = fmaf(p, t, -0.18628806f);
    p = fmaf(p, t, 0.09678418f);
    p = fmaf(p, t, 0.37409196f);
    p = fmaf(p, t, 1.00002368f);
    p = fmaf(p, t, -1.26551223f);
    const float w = t * fast_ex2(fmaf(-az, az, p) * 1.4426950408889634f);
    return z >= 0.f ? 2.0f - w : w;
}
__device__ __forceinline__ float gelu_erf(float x) {
    return (0.5f * x) * one_plus_erf(x * 0.70710678118654752440f);
}
// d/dx gelu(x) = cdf + x * pdf, fp32, same formula as ATen's GeluBackwardCUDAKernelImpl
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * one_plus_erf(x * 0.70710678118654752440f);
    const float e = fast_ex2(x * x * (-0.5f * 1.4426950408889634f));  // exp(-x^2 / 2)
    return fmaf(x * 0.39894228040143267794f, e, cdf);
}
// ---- packed fp32x2 versions (sm_100 FFMA2 / FMUL2 / FADD2: two lanes of fp32 per issue slot) --------------------
// The GELU epilogues are issue-bound (one 128x256 tile = 32768 activations against 6144 cycles of MMA), so the
// polynomial work is done two elements at a time.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float a, float b) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
    f32x2_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// w = erfc(|x| / sqrt 2) for two values, ax = |x|, x2 = x^2 (packed).  erfc(z) = t exp(-z^2 + P(t)), t = 1 / (1 + z / 2):
// degree-6 minimax fit of P (max relative error 5.6e-6 in fp32), coefficients pre-multiplied by log2(e).
__device__ __forceinline__ void erfc_abs2(float x0, float x1, f32x2_t& w, f32x2_t& ax, f32x2_t& x2) {
    ax = pk2(fabsf(x0), fabsf(x1));
    float d0, d1;
    upk2(fma2(ax, pk2(0.35355339059327373f, 0.35355339059327373f), pk2(1.0f, 1.0f)), d0, d1);  // 1 + |x| / (2 sqrt 2)
    const f32x2_t t = pk2(fast_rcp(d0), fast_rcp(d1));
    constexpr float kL = 1.4426950408889634f;
    f32x2_t p = fma2(pk2(-0.0630452529f * kL, -0.0630452529f * kL), t, pk2(0.465173551f * kL, 0.465173551f * kL));
    p = fma2(p, t, pk2(-1.06472952f * kL, -1.06472952f * kL));
    p = fma2(p, t, pk2(0.747652742f * kL, 0.747652742f * kL));
    p = fma2(p, t, pk2(0.14202046f * kL, 0.14202046f * kL));
    p = fma2(p, t, pk2(1.04137768f * kL, 1.04137768f * kL));
    p = fma2(p, t, pk2(-1.26844758f * kL, -1.26844758f * kL));
    x2 = mul2(ax, ax);
    float g0, g1;
    upk2(fma2(x2, pk2(-0.5f * kL, -0.5f * kL), p), g0, g1);  // (-x^2/2 + P(t)) log2 e
    w = mul2(t, pk2(fast_ex2(g0), fast_ex2(g1)));
}
// gelu(x) = x Phi(x) = relu(x) - |x| erfc(|x| / sqrt 2) / 2   (exact identity; no sign-dependent select)
__device__ __forceinline__ void gelu_erf2(float x0, float x1, float& g0, float& g1) {
    f32x2_t w, ax, x2;
    erfc_abs2(x0, x1, w, ax, x2);
    upk2(fma2(mul2(ax, w), pk2(-0.5f, -0.5f), pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), g0, g1);
}
// gelu'(x) = Phi(x) + x phi(x),  Phi(x) = 1/2 + sign(x) (1/2 - erfc(|x| / sqrt 2) / 2),  phi(x) = exp(-x^2/2) / sqrt(2 pi).
// Here erfc(z) = t E Q(t) with E = exp(-z^2) = exp(-x^2/2) taken out of the fit (Q(t) = erfcx(z) / t, degree-6 fit on
// t in [0.2, 1], i.e. |x| <= 11.3, max relative error 9.3e-7; beyond that E < 2e-28 and Q stays in [0.28, 0.35]):
// the SAME exponential serves phi(x), so the derivative costs 2 MUFU ops per element (rcp, ex2) instead of 3 --
// the dGELU epilogue runs 32768 of these per tile against a 16-lane MUFU pipe.  Max abs error vs ATen 3e-7.
__device__ __forceinline__ void gelu_erf_grad2(float x0, float x1, float& g0, float& g1) {
    const f32x2_t ax = pk2(fabsf(x0), fabsf(x1));
    float d0, d1;
    upk2(fma2(ax, pk2(0.35355339059327373f, 0.35355339059327373f), pk2(1.0f, 1.0f)), d0, d1);  // 1 + |x| / (2 sqrt 2)
    const f32x2_t t = pk2(fast_rcp(d0), fast_rcp(d1));
    f32x2_t q = fma2(pk2(0.09331708401441574f, 0.09331708401441574f), t, pk2(-0.37464964389801025f, -0.37464964389801025f));
    q = fma2(q, t, pk2(0.4142274856567383f, 0.4142274856567383f));
    q = fma2(q, t, pk2(0.020182941108942032f, 0.020182941108942032f));
    q = fma2(q, t, pk2(0.2881791591644287f, 0.2881791591644287f));
    q = fma2(q, t, pk2(0.27631810307502747f, 0.27631810307502747f));
    q = fma2(q, t, pk2(0.28242501616477966f, 0.28242501616477966f));
    float n0, n1;
    upk2(mul2(mul2(ax, ax), pk2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f)), n0, n1);
    const f32x2_t e = pk2(fast_ex2(n0), fast_ex2(n1));  // exp(-x^2 / 2)
    float h0, h1;
    upk2(fma2(mul2(mul2(t, q), e), pk2(-0.5f, -0.5f), pk2(0.5f, 0.5f)), h0, h1);  // 1/2 - erfc(|x| / sqrt 2) / 2
    const f32x2_t cdf = pk2(0.5f + copysignf(h0, x0), 0.5f + copysignf(h1, x1));
    upk2(fma2(mul2(pk2(x0, x1), pk2(0.39894228040143267794f, 0.39894228040143267794f)), e, cdf), g0, g1);
}

// CLIP's QuickGELU (cflearn/modules/core/activations.py:151-153): `net * torch.sigmoid(1.702 * net)` on a bf16 tensor.
// Eager runs three elementwise kernels, each rounding its output to bf16; the same three roundings are made here.
__device__ __forceinline__ float fast_sigmoid(float x) {  // 1 / (1 + 2^(-x log2 e)); fp32 accuracy ~1e-7 relative
    return fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * x));
}
__device__ __forceinline__ float quick_gelu_bf16(float x) {  // x: a bf16 value
    const float t = bf16_round(1.702f * x);
    const float sg = bf16_round(fast_sigmoid(t));
    return x * sg;  // (the caller rounds the product to bf16)
}
// autograd of the same three ops, with eager's bf16 rounding after every backward kernel (dy, x: bf16 values):
//   mul:      d_net1 = bf16(dy * s),  d_s = bf16(dy * x)            s = bf16(sigmoid(t)), t = bf16(1.702 x)
//   sigmoid:  d_t = bf16(bf16(d_s * bf16(1 - s)) * s)   -- ATen's CUDA sigmoid_backward evaluates `a * (1 - b) * b` in the
//             tensor's own type (c10::BFloat16 arithmetic rounds after every operator), unlike its opmath forward kernels
//   scale:    d_net2 = bf16(d_t * 1.702)
//   sum:      bf16(d_net1 + d_net2)   (the caller rounds the sum)
__device__ __forceinline__ float quick_gelu_bf16_grad(float x, float dy) {
    const float t = bf16_round(1.702f * x);
    const float sg = bf16_round(fast_sigmoid(t));
    const float d1 = bf16_round(dy * sg);
    const float ds = bf16_round(dy * x);
    const float dt = bf16_round(bf16_round(ds * bf16_round(1.0f - sg)) * sg);
    const float d2 = bf16_round(dt * 1.702f);
    return d1 + d2;
}

}  // namespace b200
