// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (fp32 accumulate in TMEM)
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring: 6 stages of 32 KB in the default
//                 cta_group::2 mode, where each CTA of a pair stores its A rows and half of B; 4 stages of 48 KB otherwise)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma: 256x256x16 per CTA pair, or 128x256x16)
//   warps 2..   : epilogue, 8 warps (bias-only, split-K) or 16 (GELU / dGELU / fp32-residual variants): tcgen05.ld ->
//                 registers -> swizzled smem transpose -> fused math -> coalesced 16-byte global stores; aux operands by
//                 coalesced global loads issued ahead of use
//
// The accumulator is double buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs
// of tile i+1.  Both operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); the
// latter is what the backward GEMMs need (dgrad reads W[N,K] with the reduction over N, wgrad reads dY and X
// with the reduction over tokens) -- no transposed copies are ever materialised.
//
// This one kernel serves every dense layer on the hot path of the reference's transformer block
// (SURVEY.md section 2b, K1/K4/K8/K10/K11/K12/K14):
//   F.linear call sites  cflearn/modules/core/customs.py:85-89, attentions.py:214,277, channel_mixers.py:29-33
//   conv-as-GEMM         cflearn/modules/core/convs/basic.py:155-174 (patch embed, k = s = 16)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include <type_traits>

#include "b200_internal.h"
#include "ptx.cuh"

namespace b200 {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;      // 1-CTA / multicast modes: 4 x (A 16 KB + B 32 KB)
constexpr int STAGES_2SM = 6;  // cta_group::2: each CTA stores only its half of B -> 6 x (16 + 16 KB) in the same 192 KB
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int MN_BOX_BYTES = 64 * BK * 2;   // one MN-major TMA box: 64 (MN) x 64 (K) bf16 = 8 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_STAGE_BYTES = 4096;  // per epilogue warp: 32 rows x 128 B
constexpr int GEMM_THREADS = 64 + NUM_EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;  // 2 accumulator buffers x 256 fp32 columns

constexpr int SMEM_A_OFF = 0;
constexpr int SMEM_EPI_OFF = SMEM_A_OFF + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
static_assert(STAGES_2SM * (A_STAGE_BYTES + B_STAGE_BYTES / 2) == STAGES * (A_STAGE_BYTES + B_STAGE_BYTES), "both ring layouts must end at SMEM_EPI_OFF");
constexpr int SMEM_BIAS_OFF = SMEM_EPI_OFF + NUM_EPI_WARPS * EPI_STAGE_BYTES;  // per epilogue warp: bias of its 128 columns
constexpr int EPI_BIAS_BYTES = 256;
constexpr int SMEM_BAR_OFF = SMEM_BIAS_OFF + NUM_EPI_WARPS * EPI_BIAS_BYTES;
constexpr int SMEM_BAR_BYTES = 256;
constexpr int GEMM_SMEM_USED = SMEM_BAR_OFF + SMEM_BAR_BYTES;
// everything an SM offers (227 KB); the manual 1 KB alignment of the base may take the 768 bytes that are left (in
// practice the dynamic segment starts 1 KB aligned; the kernel traps if the layout ever does not fit)
constexpr int GEMM_SMEM_BYTES = 232448;
static_assert(GEMM_SMEM_USED + 768 <= GEMM_SMEM_BYTES, "GEMM shared-memory layout does not fit");

struct GemmParams {
    int M, N, K;
    int num_m_tiles, num_n_tiles, splits, num_kb;
    int a_mn, b_mn;  // 1 = MN-major operand
    int cluster;     // 1, or 2: CTA pairs work on vertically adjacent tiles and share the B tile
    int two_sm;      // cluster == 2 only.  0: each CTA runs its own 128x256 MMA, the B tile is TMA-multicast to both;
                     //                     1: tcgen05 cta_group::2 -- ONE 256x256 MMA per pair, each CTA holds HALF of B
    const __nv_bfloat16* bias;  // [N] bf16 or nullptr
    void* out0;                 // epilogue outputs / aux operand: plain pointers, leading dimension ldo (elements)
    void* out1;
    const void* aux;
    long long ldo;
    // 3x3 convolution as implicit GEMM (conv != 0): A is a channels-last activation [B, H, W, Cin] behind a 4-D tensor map;
    // an M tile is 128 consecutive pixels (whole image rows), k-block kb = (tap, 64-channel block), and the A tile of a tap is
    // the SAME box shifted by (kx - 1, ky - 1) -- TMA zero-fills what falls outside the image: the padding costs nothing
    int conv, conv_h, conv_w, conv_cin;
    int act;  // GELU / dGELU epilogues: 0 = exact erf GELU (nn.GELU()), 1 = QuickGELU x * sigmoid(1.702 x) (activations.py:151-153)
};

// Work unit u (per cluster) -> (m tile of THIS CTA, n tile, split, k-block range).  With cluster == 2 the two CTAs of
// a cluster take m tiles 2g and 2g+1 of the same (n tile, split): identical k loop, shared B tile.
__device__ __forceinline__ void decode_unit(const GemmParams& p, int u, uint32_t cta_rank, int& m_t, int& n_t, int& sp,
                                            int& kb0, int& kb1) {
    n_t = u % p.num_n_tiles;
    int r = u / p.num_n_tiles;
    sp = r % p.splits;
    m_t = (r / p.splits) * p.cluster + static_cast<int>(cta_rank);
    kb0 = static_cast<int>((static_cast<long long>(p.num_kb) * sp) / p.splits);
    kb1 = static_cast<int>((static_cast<long long>(p.num_kb) * (sp + 1)) / p.splits);
}

// 16-byte row-per-thread accesses into a TMA-swizzled staging tile (bank-conflict free)
__device__ __forceinline__ uint32_t swz128_off(uint32_t row, uint32_t chunk) {  // 128 B rows, 8 chunks
    return row * 128u + ((chunk ^ (row & 7u)) << 4);
}
__device__ __forceinline__ uint32_t swz64_off(uint32_t row, uint32_t chunk) {  // 64 B rows, 4 chunks
    return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4);
}

// Epilogue warps per variant.  The bias-only and split-K epilogues hide behind the main loop with 8 warps (2 per scheduler).
// The GELU / dGELU / fp32-residual epilogues do not: ~19 instructions per element at ~40 % issue efficiency is 7.5 us per
// tile against a 5.3 us main loop.  They are latency-, not throughput-bound (MUFU 4.1k, FMA 2k, issue 4.9k cycles per tile),
// so those variants run 16 epilogue warps (4 per scheduler), each owning 32 rows x 64 columns.  B200_EPI16=0 builds the
// 8-warp epilogue for all variants (A/B timing).
#ifndef B200_EPI16
#define B200_EPI16 1
#endif
template <int EPI>
struct EpiWarps {
    static constexpr int value =
        (B200_EPI16 && (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RESID_F32 || EPI == EPI_DGELU_BF16)) ? 16 : NUM_EPI_WARPS;
};

// TWO_SM is a template parameter (not a runtime flag): a kernel image that contains cta_group::2 instructions can only
// be launched with a cluster size of 2, so the 1-CTA / multicast variants must be separate instantiations.
// QUICK (GELU / dGELU variants only): QuickGELU instead of the exact erf GELU.  A template parameter, not a runtime flag:
// the runtime branch cost the dGELU kernel 40 us (239 -> 278 us) in registers and spills.
template <int EPI, bool TWO_SM, bool QUICK = false>
__global__ void __launch_bounds__(64 + 32 * EpiWarps<EPI>::value, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmBh, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment as an OFFSET from the __shared__ array: a round trip through an integer hides the address space
    // from the compiler and turns every staging access into a generic LD / ST (long-scoreboard latency)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem + SMEM_A_OFF;
    // The ring is what hides the L2 -> smem latency: by Little's law an SM needs (bytes per k-block / 512 MMA cycles) x
    // ~3000 cycles of loaded TMA latency in flight -- 288 KB for 48 KB stages (we have 192 KB: the 1-CTA modes sit at ~65 %
    // tensor-pipe activity), 187 KB for the 32 KB stages of the 2-SM mode, which is why that mode gets six of them.
    constexpr int NSTAGE = TWO_SM ? STAGES_2SM : STAGES;
    constexpr int BSTG = TWO_SM ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;
    uint8_t* sB = smem + SMEM_A_OFF + NSTAGE * A_STAGE_BYTES;
    uint8_t* sEpi = smem + SMEM_EPI_OFF;
    if (smem + GEMM_SMEM_USED > smem_raw + GEMM_SMEM_BYTES) {
        if (threadIdx.x == 0) printf("b200 gemm: dynamic shared memory base %p is not 256-byte aligned\n", smem_raw);
        __trap();
    }
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR_OFF);
    uint64_t* empty_bar = full_bar + NSTAGE;
    uint64_t* tmem_full_bar = empty_bar + NSTAGE;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    constexpr int NEPI = EpiWarps<EPI>::value;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // clusters (of 1 or 2 CTAs) stride over the work units; both CTAs of a pair run the same unit sequence
    const uint32_t cta_rank = p.cluster > 1 ? cluster_ctarank() : 0u;
    const int total_units = ((p.num_m_tiles + p.cluster - 1) / p.cluster) * p.num_n_tiles * p.splits;
    const int unit0 = static_cast<int>(blockIdx.x) / p.cluster;
    const int unit_stride = static_cast<int>(gridDim.x) / p.cluster;
    const uint16_t mc_mask = static_cast<uint16_t>((1u << p.cluster) - 1u);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full_bar[i], 1);
            // multicast mode: released by the MMA warp of EVERY CTA in the cluster; 2-SM mode: by the leader's commit
            mbar_init(&empty_bar[i], TWO_SM ? 1u : static_cast<uint32_t>(p.cluster));
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], TWO_SM ? 2 * NEPI : NEPI);  // 2-SM: both CTAs' epilogues report to the leader
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if constexpr (TWO_SM) {  // pair allocation: the same warp of both CTAs, same destination offset
            tmem_alloc_2sm(tmem_ptr_smem, TMEM_COLS);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_ptr_smem, TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before_sync();
    if (p.cluster > 1) cluster_sync_all();  // the peer's barriers must be initialised before we multicast into them
    else __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = unit0; u < total_units; u += unit_stride) {
                int m_t, n_t, sp, kb0, kb1;
                decode_unit(p, u, cta_rank, m_t, n_t, sp, kb0, kb1);
                const int m0 = m_t * BM, n0 = n_t * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
                    uint8_t* b_dst = sB + stage * BSTG;
                    if constexpr (TWO_SM) {
                        // cta_group::2: every load of BOTH CTAs is credited to the LEADER's full barrier, which the
                        // leader arms once with the bytes of the whole pair (2 x (A 16 KB + half B 16 KB)).
                        const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B_STAGE_BYTES / 2));
                        const int h = static_cast<int>(cta_rank);
                        if (p.conv == 1) {
                            const int cpb = p.conv_cin / BK;
                            const int tap = kb / cpb, c0 = (kb - tap * cpb) * BK;
                            const int ky = tap / 3, kx = tap - ky * 3;
                            const int hw = p.conv_h * p.conv_w;
                            const int bimg = m0 / hw, y0 = (m0 - bimg * hw) / p.conv_w;
                            tma_load_4d_2sm(a_dst, &tmA, lead_full, c0, kx - 1, y0 + ky - 1, bimg);
                        } else if (!p.a_mn) {
                            tma_load_2d_2sm(a_dst, &tmA, lead_full, kb * BK, m0);
                        } else {
#pragma unroll
                            for (int i = 0; i < BM / 64; ++i)
                                tma_load_2d_2sm(a_dst + i * MN_BOX_BYTES, &tmA, lead_full, m0 + i * 64, kb * BK);
                        }
                        if (p.conv == 2) {
                            // weight gradient of the 3x3 convolution: B = the activation behind a 4-D map, k-block = 64 consecutive
                            // pixels, every 64-channel box of the N tile shifted by ITS tap (n = tap * Cin + c)
                            const int hw = p.conv_h * p.conv_w, p0 = kb * BK;
                            const int bimg = p0 / hw, y0 = (p0 - bimg * hw) / p.conv_w;
#pragma unroll
                            for (int i = 0; i < BN / 128; ++i) {
                                const int n = n0 + (2 * h + i) * 64;
                                const int tap = n / p.conv_cin, c0 = n - tap * p.conv_cin;
                                const int ky = tap / 3, kx = tap - ky * 3;
                                // (boxes past N = 9 Cin: tap >= 9 lands outside the image -> zero fill, bytes still credited)
                                tma_load_4d_2sm(b_dst + i * MN_BOX_BYTES, &tmB, lead_full, c0, kx - 1, tap < 9 ? y0 + ky - 1 : -65536, bimg);
                            }
                        } else if (!p.b_mn) {  // this CTA's half of the B tile (128 of the 256 n rows), at offset 0 of the stage
                            tma_load_2d_2sm(b_dst, &tmBh, lead_full, kb * BK, n0 + h * (BN / 2));
                        } else {
#pragma unroll
                            for (int i = 0; i < BN / 128; ++i)
                                tma_load_2d_2sm(b_dst + i * MN_BOX_BYTES, &tmB, lead_full, n0 + (2 * h + i) * 64, kb * BK);
                        }
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
                    if (p.conv == 1) {
                        const int cpb = p.conv_cin / BK;
                        const int tap = kb / cpb, c0 = (kb - tap * cpb) * BK;
                        const int ky = tap / 3, kx = tap - ky * 3;
                        const int hw = p.conv_h * p.conv_w;
                        const int bimg = m0 / hw, y0 = (m0 - bimg * hw) / p.conv_w;
                        tma_load_4d(a_dst, &tmA, &full_bar[stage], c0, kx - 1, y0 + ky - 1, bimg);
                    } else if (!p.a_mn) {
                        tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BK, m0);
                    } else {
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)
                            tma_load_2d(a_dst + i * MN_BOX_BYTES, &tmA, &full_bar[stage], m0 + i * 64, kb * BK);
                    }
                    if (p.conv == 2) {  // (host: cluster == 1 in this mode)
                        const int hw = p.conv_h * p.conv_w, p0 = kb * BK;
                        const int bimg = p0 / hw, y0 = (p0 - bimg * hw) / p.conv_w;
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i) {
                            const int n = n0 + i * 64;
                            const int tap = n / p.conv_cin, c0 = n - tap * p.conv_cin;
                            const int ky = tap / 3, kx = tap - ky * 3;
                            tma_load_4d(b_dst + i * MN_BOX_BYTES, &tmB, &full_bar[stage], c0, kx - 1, tap < 9 ? y0 + ky - 1 : -65536, bimg);
                        }
                    } else if (p.cluster == 1) {
                        if (!p.b_mn) {
                            tma_load_2d(b_dst, &tmB, &full_bar[stage], kb * BK, n0);
                        } else {
#pragma unroll
                            for (int i = 0; i < BN / 64; ++i)
                                tma_load_2d(b_dst + i * MN_BOX_BYTES, &tmB, &full_bar[stage], n0 + i * 64, kb * BK);
                        }
                    } else {
                        // this CTA fetches HALF of the shared B tile and multicasts it into both CTAs' smem (same
                        // offset, each CTA's own full barrier gets the bytes): 32 KB instead of 48 KB of L2 reads per
                        // CTA and k-block.
                        const int h = static_cast<int>(cta_rank);
                        if (!p.b_mn) {
                            tma_load_2d_mcast(b_dst + h * (B_STAGE_BYTES / 2), &tmBh, &full_bar[stage], kb * BK, n0 + h * (BN / 2), mc_mask);
                        } else {
#pragma unroll
                            for (int i = 0; i < BN / 128; ++i)
                                tma_load_2d_mcast(b_dst + (2 * h + i) * MN_BOX_BYTES, &tmB, &full_bar[stage], n0 + (2 * h + i) * 64, kb * BK, mc_mask);
                        }
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        if (elect_one() && (!TWO_SM || cta_rank == 0)) {  // 2-SM mode: only the pair leader issues MMAs
            const uint32_t idesc = make_idesc_bf16(TWO_SM ? 2 * BM : BM, BN, p.a_mn, p.b_mn);
            int stage = 0;
            uint32_t phase = 0;
            int lt = 0;
            for (int u = unit0; u < total_units; u += unit_stride, ++lt) {
                int m_t, n_t, sp, kb0, kb1;
                decode_unit(p, u, cta_rank, m_t, n_t, sp, kb0, kb1);
                const int as = lt & 1;
                const uint32_t aph = (lt >> 1) & 1u;
                mbar_wait(&tmem_empty_bar[as], aph ^ 1u);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after_sync();
                    const uint32_t a_base = smem_u32(sA + stage * A_STAGE_BYTES);
                    const uint32_t b_base = smem_u32(sB + stage * BSTG);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        // K-major  : 8-row groups 1024 B apart (SBO); +32 B per 16-element K step.
                        // MN-major : 8-k-row groups 1024 B apart (SBO); 64-element MN atoms one TMA box
                        //            (8 KB) apart (LBO); +2048 B per 16-row K step.
                        const uint64_t adesc = p.a_mn ? make_smem_desc(a_base + kk * 2048, MN_BOX_BYTES, 1024, kSwz128)
                                                      : make_smem_desc(a_base + kk * 32, 0, 1024, kSwz128);
                        const uint64_t bdesc = p.b_mn ? make_smem_desc(b_base + kk * 2048, MN_BOX_BYTES, 1024, kSwz128)
                                                      : make_smem_desc(b_base + kk * 32, 0, 1024, kSwz128);
                        if constexpr (TWO_SM) umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
                        else umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
                    }
                    // smem slot reusable once these MMAs retire; with multicast the slot is also written by the peer's
                    // TMA, so the release goes to the empty barrier of both CTAs
                    if constexpr (TWO_SM) {
                        umma_commit_2sm_mcast(&empty_bar[stage], mc_mask);
                    } else {
                        if (p.cluster > 1) umma_commit_mcast(&empty_bar[stage], mc_mask);
                        else umma_commit(&empty_bar[stage]);
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
                }
                // accumulator complete -> epilogue (2-SM: each CTA's epilogue owns 128 of the 256 accumulator rows)
                if constexpr (TWO_SM) umma_commit_2sm_mcast(&tmem_full_bar[as], mc_mask);
                else umma_commit(&tmem_full_bar[as]);
            }
        }
    } else if constexpr (NEPI == 16) {
        // ===================================== epilogue warps (16-warp variants) ================================
        // Each warp owns 32 accumulator rows (its TMEM lane quadrant) x 64 columns = 2 chunks of 32 columns.  Same two
        // phases as below, but ONE 32-register accumulator chunk at a time (the 96-register budget of an 18-warp CTA) and a
        // 2 KB staging buffer per warp; the TMEM buffer is handed back after the second chunk has been read, i.e. after
        // half of this warp's work.
        const int e = warp - 2;               // 0..15
        const uint32_t q = warp & 3;          // TMEM lane quadrant this warp may access
        const int part = e >> 2;              // which 64-column quarter of the accumulator
        uint8_t* stg = sEpi + e * 2048;
        uint8_t* sbias = smem + SMEM_BIAS_OFF + e * 128;
        const long long ldo = p.ldo;
        const int prow = lane >> 2, pch = lane & 3;  // phase-2 coordinates: 8 row groups of 4 lanes, 16 B per lane
        int lt = 0;
        for (int u = unit0; u < total_units; u += unit_stride, ++lt) {
            int m_t, n_t, sp, kb0, kb1;
            decode_unit(p, u, cta_rank, m_t, n_t, sp, kb0, kb1);
            const int as = lt & 1;
            const uint32_t aph = (lt >> 1) & 1u;
            const int grow0 = m_t * BM + static_cast<int>(q) * 32;
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + static_cast<uint32_t>(as * BN + part * 64);
            const int gcol0 = n_t * BN + part * 64;
            const bool interior = (m_t * BM + BM <= p.M) && (n_t * BN + BN <= p.N);
            auto tile_epilogue = [&](auto interior_tag) {
            constexpr bool INT = decltype(interior_tag)::value;
            constexpr int NSLOT = EPI == EPI_BIAS_RESID_F32 ? 1 : 2;  // the fp32 residual is 32 registers per chunk
            uint4 aux_h[2][4];
            float4 aux_r[1][4][2];
            auto load_aux = [&](int cc, int slot) {
                const int pc = gcol0 + cc * 32 + pch * 8;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int row = grow0 + it * 8 + prow;
                    const bool ok = INT || (row < p.M && pc + 8 <= p.N);
                    if constexpr (EPI == EPI_DGELU_BF16) {
                        aux_h[slot][it] = make_uint4(0, 0, 0, 0);
                        if (ok) aux_h[slot][it] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + row * ldo + pc);
                    }
                    if constexpr (EPI == EPI_BIAS_RESID_F32) {
                        aux_r[slot][it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                        aux_r[slot][it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok) {
                            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + row * ldo + pc);
                            aux_r[slot][it][0] = src[0];
                            aux_r[slot][it][1] = src[1];
                        }
                    }
                }
            };
            if constexpr (EPI != EPI_DGELU_BF16) {  // bias of this warp's 64 columns -> 128 B of private smem
                if (lane < 8) {
                    const int gc = gcol0 + lane * 8;
                    uint4 t = make_uint4(0, 0, 0, 0);
                    if (p.bias != nullptr && (INT || gc < p.N)) t = *reinterpret_cast<const uint4*>(p.bias + gc);
                    *reinterpret_cast<uint4*>(sbias + lane * 16) = t;
                }
                __syncwarp();
            }
            if constexpr (EPI != EPI_BIAS_GELU_BF16) {
                load_aux(0, 0);
                if constexpr (NSLOT == 2) load_aux(1, 1);
            }
            mbar_wait(&tmem_full_bar[as], aph);
            tc_fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int gcol = gcol0 + c * 32;
                const int pcol = gcol + pch * 8;
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), v);
                tmem_ld_wait();
                if (c == 1) {  // the accumulator has been read completely by this warp
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (TWO_SM) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0));  // leader's barrier
                        else mbar_arrive(&tmem_empty_bar[as]);
                    }
                }
                const bool active = INT || ((gcol < p.N) && (grow0 < p.M));  // warp-uniform
                if (active) {
                    // ---- phase 1: bf16(acc + bias) -> staging (row-per-thread) ----
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        uint4 t = make_uint4(0, 0, 0, 0);
                        if constexpr (EPI != EPI_DGELU_BF16) t = *reinterpret_cast<const uint4*>(sbias + c * 64 + j4 * 16);
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[j4 * 8 + 0]) + bf16lo(t.x), __uint_as_float(v[j4 * 8 + 1]) + bf16hi(t.x));
                        o.y = pack_bf16x2(__uint_as_float(v[j4 * 8 + 2]) + bf16lo(t.y), __uint_as_float(v[j4 * 8 + 3]) + bf16hi(t.y));
                        o.z = pack_bf16x2(__uint_as_float(v[j4 * 8 + 4]) + bf16lo(t.z), __uint_as_float(v[j4 * 8 + 5]) + bf16hi(t.z));
                        o.w = pack_bf16x2(__uint_as_float(v[j4 * 8 + 6]) + bf16lo(t.w), __uint_as_float(v[j4 * 8 + 7]) + bf16hi(t.w));
                        *reinterpret_cast<uint4*>(stg + swz64_off(lane, j4)) = o;
                    }
                    __syncwarp();
                    // ---- phase 2: coalesced layout ----
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = it * 8 + prow;
                        const int row = grow0 + r;
                        const uint4 t = *reinterpret_cast<const uint4*>(stg + swz64_off(r, pch));  // 8 x bf16(acc + bias)
                        if (!INT && row >= p.M) continue;
                        const bool full = INT || pcol + 8 <= p.N;
                        const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
                        if constexpr (EPI == EPI_BIAS_GELU_BF16) {
                            uint32_t gw[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                float g0, g1;
                                if constexpr (!QUICK) {
                                    gelu_erf2(bf16lo(tw[k]), bf16hi(tw[k]), g0, g1);
                                } else {
                                    g0 = quick_gelu_bf16(bf16lo(tw[k]));
                                    g1 = quick_gelu_bf16(bf16hi(tw[k]));
                                }
                                gw[k] = pack_bf16x2(g0, g1);
                            }
                            __nv_bfloat16* d0 = reinterpret_cast<__nv_bfloat16*>(p.out0) + row * ldo + pcol;
                            __nv_bfloat16* d1 = reinterpret_cast<__nv_bfloat16*>(p.out1) + row * ldo + pcol;
                            if (full) {
                                *reinterpret_cast<uint4*>(d0) = t;
                                *reinterpret_cast<uint4*>(d1) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
                            } else {
                                for (int k = 0; k < 8; ++k)
                                    if (pcol + k < p.N) {
                                        d0[k] = __ushort_as_bfloat16(static_cast<unsigned short>(tw[k >> 1] >> ((k & 1) * 16)));
                                        d1[k] = __ushort_as_bfloat16(static_cast<unsigned short>(gw[k >> 1] >> ((k & 1) * 16)));
                                    }
                            }
                        } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
                            float* dst = reinterpret_cast<float*>(p.out0) + row * ldo + pcol;
                            if (full) {
                                float4 a = aux_r[0][it][0], b = aux_r[0][it][1];
                                a.x += bf16lo(tw[0]); a.y += bf16hi(tw[0]); a.z += bf16lo(tw[1]); a.w += bf16hi(tw[1]);
                                b.x += bf16lo(tw[2]); b.y += bf16hi(tw[2]); b.z += bf16lo(tw[3]); b.w += bf16hi(tw[3]);
                                reinterpret_cast<float4*>(dst)[0] = a;
                                reinterpret_cast<float4*>(dst)[1] = b;
                            } else {
                                const float* src = reinterpret_cast<const float*>(p.aux) + row * ldo + pcol;
                                for (int k = 0; k < 8; ++k)
                                    if (pcol + k < p.N) dst[k] = src[k] + ((k & 1) ? bf16hi(tw[k >> 1]) : bf16lo(tw[k >> 1]));
                            }
                        } else {  // EPI_DGELU_BF16
                            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out0) + row * ldo + pcol;
                            if (full) {
                                const uint4 hh = aux_h[c][it];
                                const uint32_t hw[4] = {hh.x, hh.y, hh.z, hh.w};
                                uint32_t ow[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if constexpr (!QUICK) {
                                        float g0, g1;
                                        gelu_erf_grad2(bf16lo(hw[k]), bf16hi(hw[k]), g0, g1);
                                        ow[k] = pack_bf16x2(bf16lo(tw[k]) * g0, bf16hi(tw[k]) * g1);
                                    } else {
                                        ow[k] = pack_bf16x2(quick_gelu_bf16_grad(bf16lo(hw[k]), bf16lo(tw[k])),
                                                            quick_gelu_bf16_grad(bf16hi(hw[k]), bf16hi(tw[k])));
                                    }
                                }
                                *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                            } else {
                                const __nv_bfloat16* hsrc = reinterpret_cast<const __nv_bfloat16*>(p.aux) + row * ldo + pcol;
                                for (int k = 0; k < 8; ++k)
                                    if (pcol + k < p.N) {
                                        const float dv = (k & 1) ? bf16hi(tw[k >> 1]) : bf16lo(tw[k >> 1]);
                                        const float hv = __bfloat162float(hsrc[k]);
                                        dst[k] = __float2bfloat16_rn(!QUICK ? dv * gelu_erf_grad(hv) : quick_gelu_bf16_grad(hv, dv));
                                    }
                            }
                        }
                    }
                    __syncwarp();  // the single staging buffer is rewritten by the next chunk
                }
                if constexpr (EPI == EPI_BIAS_RESID_F32) {
                    if (c == 0) load_aux(1, 0);  // the only residual slot is free again
                }
            }
            };  // tile_epilogue
            if (interior) tile_epilogue(std::true_type{});
            else tile_epilogue(std::false_type{});
        }
    } else {
        // ===================================== epilogue warps ===================================
        // Each warp owns 32 accumulator rows (its TMEM lane quadrant) x 128 columns, processed in 4 chunks of 32
        // columns.  Phase 1 (row-per-thread, the tcgen05.ld layout): acc (+bias) -> bf16 (or fp32) -> swizzled staging
        // smem.  Phase 2 (coalesced layout: 4 lanes x 16 B per 64-byte row segment): staging + aux operand read
        // straight from global -> fused math -> 16-byte global stores.  No asynchronous store completion is waited on
        // (the first version used TMA stores and spent ~1 us per chunk in cp.async.bulk.wait_group.read); aux loads are
        // issued before the TMEM load and were L2-prefetched by the producer warp a whole mainloop earlier.
        const int e = warp - 2;               // 0..7
        const uint32_t q = warp & 3;          // TMEM lane quadrant this warp may access
        const int half = e >> 2;              // which 128-column half of the accumulator
        uint8_t* stg = sEpi + e * EPI_STAGE_BYTES;
        uint8_t* sbias = smem + SMEM_BIAS_OFF + e * EPI_BIAS_BYTES;
        const long long ldo = p.ldo;
        int lt = 0;
        for (int u = unit0; u < total_units; u += unit_stride, ++lt) {
            int m_t, n_t, sp, kb0, kb1;
            decode_unit(p, u, cta_rank, m_t, n_t, sp, kb0, kb1);
            const int as = lt & 1;
            const uint32_t aph = (lt >> 1) & 1u;
            const int grow0 = m_t * BM + static_cast<int>(q) * 32;
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + static_cast<uint32_t>(as * BN);
            // Tiles that lie completely inside the output (every tile of the ViT shapes) run a copy of the epilogue with
            // all bounds predicates compiled out: the ragged-edge checks were ~7 of its ~27 instructions per element.
            const bool interior = (m_t * BM + BM <= p.M) && (n_t * BN + BN <= p.N);
            auto tile_epilogue = [&](auto interior_tag) {
            constexpr bool INT = decltype(interior_tag)::value;
            // The accumulator buffer can only be handed back to the MMA warp once it has been READ completely, and
            // MMA(i+2) waits for that.  So the TMEM reads run two chunks ahead of the processing (register ping-pong
            // va / vb): the buffer is released after chunk 1 instead of after chunk 3, i.e. after about half of the
            // epilogue -- otherwise a K = 768 tile (3.6 us of MMA) stalls behind a ~6 us serial epilogue.
            // aux operand (residual / pre-activation) in the coalesced phase-2 layout, requested one chunk ahead; the first
            // chunk's loads are issued BEFORE the accumulator is ready, so their HBM latency hides behind the mainloop
            uint4 aux_h[2][4];      // DGELU : 8 bf16 of h per row group
            float4 aux_r[2][4][2];  // RESID : 8 fp32 of the residual per row group
            auto load_aux = [&](int cc, int slot) {
                const int gc = n_t * BN + half * 128 + cc * 32;
                const int pc = gc + (lane & 3) * 8;
                const bool act = INT || ((gc < p.N) && (grow0 < p.M));
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int row = grow0 + it * 8 + (lane >> 2);
                    const bool ok = INT || (act && row < p.M && pc + 8 <= p.N);
                    if constexpr (EPI == EPI_DGELU_BF16) {
                        aux_h[slot][it] = make_uint4(0, 0, 0, 0);
                        if (ok) aux_h[slot][it] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + row * ldo + pc);
                    }
                    if constexpr (EPI == EPI_BIAS_RESID_F32) {
                        aux_r[slot][it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                        aux_r[slot][it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok) {
                            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + row * ldo + pc);
                            aux_r[slot][it][0] = src[0];
                            aux_r[slot][it][1] = src[1];
                        }
                    }
                }
            };
            // bias of this warp's 128 columns -> its private 256 B of shared memory, before the accumulator is awaited.
            // (Loaded from global at the point of use, the four 16-byte loads of every chunk exposed their L2 latency.)
            if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RESID_F32) {
                if (lane < 16) {
                    const int gc = n_t * BN + half * 128 + lane * 8;
                    uint4 t = make_uint4(0, 0, 0, 0);
                    if (p.bias != nullptr && (INT || gc < p.N)) t = *reinterpret_cast<const uint4*>(p.bias + gc);
                    *reinterpret_cast<uint4*>(sbias + lane * 16) = t;
                }
                __syncwarp();
            }
            if constexpr (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_DGELU_BF16) load_aux(0, 0);
            mbar_wait(&tmem_full_bar[as], aph);
            tc_fence_after_sync();
            uint32_t va[32], vb[32];
            tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(half * 128), va);
            tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(half * 128 + 32), vb);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t (&v)[32] = (c & 1) ? vb : va;
                const int col = half * 128 + c * 32;
                const int gcol = n_t * BN + col;
                const bool active = INT || ((gcol < p.N) && (grow0 < p.M));  // warp-uniform
                // phase-2 coordinates of this lane inside a bf16-staged chunk: 4 row groups of 8 rows, 4 lanes per row
                const int prow = lane >> 2, pch = lane & 3;
                const int pcol = gcol + pch * 8;

                if constexpr (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_DGELU_BF16) {
                    if (c + 1 < 4) load_aux(c + 1, (c + 1) & 1);
                }
                if (active) {

                if constexpr (EPI == EPI_PARTIAL_F32) {
                    // ---- fp32 split-K partial: stage 32 x 128 B, then 8 lanes x 16 B per row ----
#pragma unroll
                    for (int j8 = 0; j8 < 8; ++j8)
                        *reinterpret_cast<uint4*>(stg + swz128_off(lane, j8)) =
                            make_uint4(v[j8 * 4 + 0], v[j8 * 4 + 1], v[j8 * 4 + 2], v[j8 * 4 + 3]);
                    __syncwarp();
                    float* out = reinterpret_cast<float*>(p.out0) + static_cast<long long>(sp) * p.M * ldo;
                    const int frow = lane >> 3, fch = lane & 7;
                    const int fcol = gcol + fch * 4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + frow;
                        const int row = grow0 + r;
                        const uint4 t = *reinterpret_cast<const uint4*>(stg + swz128_off(r, fch));
                        if (INT || row < p.M) {
                            float* dst = out + row * ldo + fcol;
                            if (INT || fcol + 4 <= p.N) {
                                *reinterpret_cast<uint4*>(dst) = t;
                            } else {
                                const uint32_t tt[4] = {t.x, t.y, t.z, t.w};
                                for (int k = 0; k < 4; ++k)
                                    if (fcol + k < p.N) dst[k] = __uint_as_float(tt[k]);
                            }
                        }
                    }
                    __syncwarp();  // staging is rewritten by the next chunk
                } else {
                    // ---- phase 1: bf16(acc + bias) -> staging half (c & 1) ----
                    uint8_t* const sh = stg + (c & 1) * 2048;
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float bv[8];
                        {
                            uint4 t = make_uint4(0, 0, 0, 0);
                            if constexpr (EPI != EPI_DGELU_BF16) t = *reinterpret_cast<const uint4*>(sbias + c * 64 + j4 * 16);
                            bv[0] = bf16lo(t.x); bv[1] = bf16hi(t.x); bv[2] = bf16lo(t.y); bv[3] = bf16hi(t.y);
                            bv[4] = bf16lo(t.z); bv[5] = bf16hi(t.z); bv[6] = bf16lo(t.w); bv[7] = bf16hi(t.w);
                        }
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[j4 * 8 + 0]) + bv[0], __uint_as_float(v[j4 * 8 + 1]) + bv[1]);
                        o.y = pack_bf16x2(__uint_as_float(v[j4 * 8 + 2]) + bv[2], __uint_as_float(v[j4 * 8 + 3]) + bv[3]);
                        o.z = pack_bf16x2(__uint_as_float(v[j4 * 8 + 4]) + bv[4], __uint_as_float(v[j4 * 8 + 5]) + bv[5]);
                        o.w = pack_bf16x2(__uint_as_float(v[j4 * 8 + 6]) + bv[6], __uint_as_float(v[j4 * 8 + 7]) + bv[7]);
                        *reinterpret_cast<uint4*>(sh + swz64_off(lane, j4)) = o;
                    }
                    __syncwarp();  // (the two halves alternate, so one barrier per chunk also covers the WAR hazard)
                    // ---- phase 2: coalesced layout ----
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = it * 8 + prow;
                        const int row = grow0 + r;
                        const uint4 t = *reinterpret_cast<const uint4*>(sh + swz64_off(r, pch));  // 8 x bf16(acc + bias)
                        if (!INT && row >= p.M) continue;
                        const bool full = INT || pcol + 8 <= p.N;
                        const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
                        if constexpr (EPI == EPI_BIAS_BF16) {
                            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out0) + row * ldo + pcol;
                            if (full) *reinterpret_cast<uint4*>(dst) = t;
                            else for (int k = 0; k < 8; ++k) if (pcol + k < p.N) dst[k] = __ushort_as_bfloat16(static_cast<unsigned short>(tw[k >> 1] >> ((k & 1) * 16)));
                        } else if constexpr (EPI == EPI_BIAS_GELU_BF16) {
                            // out0 = h (rounded pre-activation), out1 = bf16(gelu(h)): GELU sees the ROUNDED h, like eager
                            uint32_t gw[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                float g0, g1;
                                if constexpr (!QUICK) {
                                    gelu_erf2(bf16lo(tw[k]), bf16hi(tw[k]), g0, g1);
                                } else {
                                    g0 = quick_gelu_bf16(bf16lo(tw[k]));
                                    g1 = quick_gelu_bf16(bf16hi(tw[k]));
                                }
                                gw[k] = pack_bf16x2(g0, g1);
                            }
                            __nv_bfloat16* d0 = reinterpret_cast<__nv_bfloat16*>(p.out0) + row * ldo + pcol;
                            __nv_bfloat16* d1 = reinterpret_cast<__nv_bfloat16*>(p.out1) + row * ldo + pcol;
                            if (full) {
                                *reinterpret_cast<uint4*>(d0) = t;
                                *reinterpret_cast<uint4*>(d1) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
                            } else {
                                for (int k = 0; k < 8; ++k)
                                    if (pcol + k < p.N) {
                                        d0[k] = __ushort_as_bfloat16(static_cast<unsigned short>(tw[k >> 1] >> ((k & 1) * 16)));
                                        d1[k] = __ushort_as_bfloat16(static_cast<unsigned short>(gw[k >> 1] >> ((k & 1) * 16)));
                                    }
                            }
                        } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
                            // out0(fp32) = aux(fp32) + float(bf16(acc + bias))     (aux may alias out0)
                            float* dst = reinterpret_cast<float*>(p.out0) + row * ldo + pcol;
                            if (full) {
                                float4 a = aux_r[c & 1][it][0], b = aux_r[c & 1][it][1];
                                a.x += bf16lo(tw[0]); a.y += bf16hi(tw[0]); a.z += bf16lo(tw[1]); a.w += bf16hi(tw[1]);
                                b.x += bf16lo(tw[2]); b.y += bf16hi(tw[2]); b.z += bf16lo(tw[3]); b.w += bf16hi(tw[3]);
                                reinterpret_cast<float4*>(dst)[0] = a;
                                reinterpret_cast<float4*>(dst)[1] = b;
                            } else {
                                const float* src = reinterpret_cast<const float*>(p.aux) + row * ldo + pcol;
                                for (int k = 0; k < 8; ++k)
                                    if (pcol + k < p.N) dst[k] = src[k] + ((k & 1) ? bf16hi(tw[k >> 1]) : bf16lo(tw[k >> 1]));
                            }
                        } else {  // EPI_DGELU_BF16: out0 = bf16( float(bf16(acc)) * gelu'(h) ),  h = aux (bf16)
                            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out0) + row * ldo + pcol;
                            if (full) {
                                const uint32_t hw[4] = {aux_h[c & 1][it].x, aux_h[c & 1][it].y, aux_h[c & 1][it].z, aux_h[c & 1][it].w};
                                uint32_t ow[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if constexpr (!QUICK) {
                                        float g0, g1;
                                        gelu_erf_grad2(bf16lo(hw[k]), bf16hi(hw[k]), g0, g1);
                                        ow[k] = pack_bf16x2(bf16lo(tw[k]) * g0, bf16hi(tw[k]) * g1);
                                    } else {
                                        ow[k] = pack_bf16x2(quick_gelu_bf16_grad(bf16lo(hw[k]), bf16lo(tw[k])),
                                                            quick_gelu_bf16_grad(bf16hi(hw[k]), bf16hi(tw[k])));
                                    }
                                }
                                *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                            } else {
                                const __nv_bfloat16* hsrc = reinterpret_cast<const __nv_bfloat16*>(p.aux) + row * ldo + pcol;
                                for (int k = 0; k < 8; ++k)
                                    if (pcol + k < p.N) {
                                        const float dv = (k & 1) ? bf16hi(tw[k >> 1]) : bf16lo(tw[k >> 1]);
                                        const float hv = __bfloat162float(hsrc[k]);
                                        dst[k] = __float2bfloat16_rn(!QUICK ? dv * gelu_erf_grad(hv) : quick_gelu_bf16_grad(hv, dv));
                                    }
                            }
                        }
                    }
                }
                }  // if (active)
                // ---- refill the register set just consumed with the chunk two ahead; release TMEM after chunk 1 ----
                if (c + 2 < 4) tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(half * 128 + (c + 2) * 32), v);
                if (c == 1) {
                    tmem_ld_wait();  // chunks 2 and 3 are now in registers: the accumulator has been read completely
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (TWO_SM) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0));  // leader's barrier
                        else mbar_arrive(&tmem_empty_bar[as]);
                    }
                }
            }
            };  // tile_epilogue
            if (interior) tile_epilogue(std::true_type{});
            else tile_epilogue(std::false_type{});
        }
    }

    __syncwarp();  // reconverge warps 0 / 1 (only one elected lane ran the role loop): the cluster barrier is .aligned
    tc_fence_before_sync();
    if (p.cluster > 1) cluster_sync_all();  // no CTA may exit while its peer can still multicast into / arrive on its smem
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        if constexpr (TWO_SM) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

// rank-2/3 tiled map; dims/box innermost first; strides in BYTES for dims 1.. (rank-1 entries)
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return set_error(B200_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return set_error(B200_ERR_ALIGN, "tensor base not 16 B aligned");
    for (int i = 0; i + 1 < rank; ++i)
        if (gstr[i] % 16 != 0) return set_error(B200_ERR_ALIGN, "tensor stride not a multiple of 16 B");
    CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[160];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu %llu box %u %u)",
                 static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return set_error(B200_ERR_DRIVER, msg);
    }
    return 0;
}

int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev < 0 || dev >= 64) ? 0 : dev;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

// Upper bound on the grid of every persistent kernel (GEMM, attention version 2), 0 = all SMs.  The data-parallel step lowers it
// by the communicator's CTA count while gradient buckets are in flight: these kernels assign their tiles to CTAs statically, and
// a CTA whose SM is held by an NCCL kernel starts only when that all-reduce is over -- with its full share of tiles still to do.
static int g_persistent_ctas = 0;
static int g_persistent_window = 0;  // > 0: the limit lapses after that many more persistent launches
int persistent_ctas() {  // (called exactly once per persistent launch)
    const int n = num_sms();
    const int r = (g_persistent_ctas > 0 && g_persistent_ctas < n) ? g_persistent_ctas : n;
    if (g_persistent_window > 0 && --g_persistent_window == 0) g_persistent_ctas = 0;
    return r;
}
extern "C" int b200_set_persistent_ctas(int n, int launches) {
    const int old = g_persistent_ctas;
    g_persistent_ctas = n > 0 ? n : 0;
    g_persistent_window = (n > 0 && launches > 0) ? launches : 0;
    return old;
}

static int g_gemm_multicast = 2;  // 0: one CTA per tile; 1: CTA pairs + TMA multicast; 2 (default): CTA pairs + cta_group::2 MMA, 6-stage ring

template <int EPI, bool TWO_SM, bool QUICK>
static int launch_gemm_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBh,
                       const GemmParams& p, int grid, cudaStream_t stream) {
    static bool configured[64] = {};  // the opt-in to > 48 KB of dynamic shared memory is per function AND per device
    cudaError_t e;
    const int dev = current_device_slot();
    if (!configured[dev]) {
        e = cudaFuncSetAttribute(gemm_bf16_kernel<EPI, TWO_SM, QUICK>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        if (e != cudaSuccess) return set_error(B200_ERR_LAUNCH, cudaGetErrorString(e));
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(64 + 32 * EpiWarps<EPI>::value);
    cfg.dynamicSmemBytes = GEMM_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(p.cluster);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<EPI, TWO_SM, QUICK>, tmA, tmB, tmBh, p);
    if (e != cudaSuccess) return set_error(B200_ERR_LAUNCH, cudaGetErrorString(e));
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(B200_ERR_LAUNCH, cudaGetErrorString(e));
    count_launch(1);
    return 0;
}

template <int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBh,
                       const GemmParams& p, int grid, cudaStream_t stream) {
    if constexpr (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_DGELU_BF16) {
        if (p.act != 0)
            return p.two_sm ? launch_gemm_impl<EPI, true, true>(tmA, tmB, tmBh, p, grid, stream)
                            : launch_gemm_impl<EPI, false, true>(tmA, tmB, tmBh, p, grid, stream);
    }
    return p.two_sm ? launch_gemm_impl<EPI, true, false>(tmA, tmB, tmBh, p, grid, stream)
                    : launch_gemm_impl<EPI, false, false>(tmA, tmB, tmBh, p, grid, stream);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_set_gemm_multicast(int enable) {
    const int old = g_gemm_multicast;
    g_gemm_multicast = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
    return old;
}

extern "C" int b200_gemm_pick_splits(int M, int N, int K) {
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int num_kb = (K + BK - 1) / BK;
    const int sms = 148;
    int best = 1;
    double best_score = -1.0;
    for (int s = 1; s <= 16 && s <= num_kb; ++s) {
        const int units = tiles * s;
        const int waves = (units + sms - 1) / sms;
        double eff = static_cast<double>(units) / (static_cast<double>(waves) * sms);
        // each split costs a pipeline fill/drain (~ 8 k-blocks worth) and an extra fp32 partial
        const double kb_per = static_cast<double>(num_kb) / s;
        eff *= kb_per / (kb_per + 8.0);
        if (eff > best_score + 1e-9) {
            best_score = eff;
            best = s;
        }
    }
    return best;
}

struct ConvGeom { int B, H, W, Cin, mode; };  // mode 1: forward / input gradient (A behind the 4-D map); 2: weight gradient (B)

static int gemm_dispatch(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                         int b_mn_major, int M, int N, int K, int epilogue, const void* bias, void* out0,
                         void* out1, const void* aux, long long ldo, int splits, int max_ctas,
                         cudaStream_t stream, const ConvGeom* cg) {
    int act = 0;  // the QuickGELU epilogues are the GELU / dGELU kernels with another activation
    if (epilogue == B200_EPI_BIAS_QGELU_BF16) { epilogue = EPI_BIAS_GELU_BF16; act = 1; }
    if (epilogue == B200_EPI_DQGELU_BF16) { epilogue = EPI_DGELU_BF16; act = 1; }
    if (M <= 0 || N <= 0 || K <= 0) return set_error(B200_ERR_ARG, "gemm: non-positive dimension");
    // N need not be a multiple of 8: the bias is read in 16-byte groups (readable up to roundup(N, 8) elements) and the
    // TMA store clips columns >= N; the leading dimension must keep rows 16-byte aligned (checked in make_tmap).
    if (A == nullptr || B == nullptr || out0 == nullptr) return set_error(B200_ERR_ARG, "gemm: null pointer");
    if (splits < 1) splits = 1;
    if (epilogue != EPI_PARTIAL_F32 && splits != 1) return set_error(B200_ERR_ARG, "gemm: split-K needs EPI_PARTIAL_F32");
    const int num_kb = (K + BK - 1) / BK;
    if (splits > num_kb) splits = num_kb;

    CUtensorMap tmA, tmB, tmBh;
    int rc;
    {
        // A: K-major [M, lda] -> dims {K, M}, box {64, 128};  MN-major [K, lda] -> dims {M, K}, box {64, 64}
        uint64_t d[2], s[1];
        uint32_t bx[2];
        if (cg != nullptr && cg->mode == 1) {
            // channels-last activation [B, H, W, Cin]: dims {Cin, W, H, B}; box = 64 channels x 128 consecutive pixels
            const int hb = cg->H < BM / cg->W ? cg->H : BM / cg->W;      // image rows per tile
            const int nb = BM / (cg->W * hb);                             // images per tile (> 1 only when H * W < 128)
            uint64_t d4[4] = {static_cast<uint64_t>(cg->Cin), static_cast<uint64_t>(cg->W), static_cast<uint64_t>(cg->H), static_cast<uint64_t>(cg->B)};
            uint64_t s4[3] = {static_cast<uint64_t>(cg->Cin) * 2, static_cast<uint64_t>(cg->W) * cg->Cin * 2,
                              static_cast<uint64_t>(cg->H) * cg->W * cg->Cin * 2};
            uint32_t b4[4] = {BK, static_cast<uint32_t>(cg->W), static_cast<uint32_t>(hb), static_cast<uint32_t>(nb)};
            if ((rc = make_tmap(&tmA, A, 2, 4, d4, s4, b4, 128)) != 0) return rc;
        } else {
        if (!a_mn_major) { d[0] = K; d[1] = M; bx[0] = BK; bx[1] = BM; }
        else             { d[0] = M; d[1] = K; bx[0] = 64; bx[1] = BK; }
        s[0] = static_cast<uint64_t>(lda) * 2;
        if ((rc = make_tmap(&tmA, A, 2, 2, d, s, bx, 128)) != 0) return rc;
        }
        if (cg != nullptr && cg->mode == 2) {
            // weight gradient: the activation as the MN-major B operand, box = 64 channels x 64 consecutive pixels (one k-block)
            const int hb = BK / cg->W;
            uint64_t d4[4] = {static_cast<uint64_t>(cg->Cin), static_cast<uint64_t>(cg->W), static_cast<uint64_t>(cg->H), static_cast<uint64_t>(cg->B)};
            uint64_t s4[3] = {static_cast<uint64_t>(cg->Cin) * 2, static_cast<uint64_t>(cg->W) * cg->Cin * 2,
                              static_cast<uint64_t>(cg->H) * cg->W * cg->Cin * 2};
            uint32_t b4[4] = {64, static_cast<uint32_t>(cg->W), static_cast<uint32_t>(hb), 1};
            if ((rc = make_tmap(&tmB, B, 2, 4, d4, s4, b4, 128)) != 0) return rc;
            tmBh = tmB;
        } else {
        if (!b_mn_major) { d[0] = K; d[1] = N; bx[0] = BK; bx[1] = BN; }
        else             { d[0] = N; d[1] = K; bx[0] = 64; bx[1] = BK; }
        s[0] = static_cast<uint64_t>(ldb) * 2;
        if ((rc = make_tmap(&tmB, B, 2, 2, d, s, bx, 128)) != 0) return rc;
        tmBh = tmB;
        if (!b_mn_major) {  // half-tile box for the multicast path (MN-major boxes are 64 wide already)
            bx[1] = BN / 2;
            if ((rc = make_tmap(&tmBh, B, 2, 2, d, s, bx, 128)) != 0) return rc;
        }
        }
    }
    const bool out_f32 = (epilogue == EPI_BIAS_RESID_F32 || epilogue == EPI_PARTIAL_F32);
    {
        // The epilogue writes with plain 16-byte global stores: rows must start 16-byte aligned.
        const int eb = out_f32 ? 4 : 2;
        if ((static_cast<unsigned long long>(ldo) * eb) % 16 != 0) return set_error(B200_ERR_ALIGN, "gemm: ldo rows must be 16-byte aligned");
        if ((reinterpret_cast<uintptr_t>(out0) & 15u) != 0) return set_error(B200_ERR_ALIGN, "gemm: out0 not 16-byte aligned");
        if (epilogue == EPI_BIAS_GELU_BF16 && (out1 == nullptr || (reinterpret_cast<uintptr_t>(out1) & 15u) != 0))
            return set_error(B200_ERR_ARG, "gemm: GELU epilogue needs a 16-byte aligned out1");
        if ((epilogue == EPI_BIAS_RESID_F32 || epilogue == EPI_DGELU_BF16) &&
            (aux == nullptr || (reinterpret_cast<uintptr_t>(aux) & 15u) != 0))
            return set_error(B200_ERR_ARG, "gemm: epilogue needs a 16-byte aligned aux");
    }
    GemmParams p;
    p.act = act;
    p.conv = cg != nullptr ? cg->mode : 0;
    p.conv_h = cg != nullptr ? cg->H : 0;
    p.conv_w = cg != nullptr ? cg->W : 0;
    p.conv_cin = cg != nullptr ? cg->Cin : 0;
    p.M = M; p.N = N; p.K = K;
    p.num_m_tiles = (M + BM - 1) / BM;
    p.num_n_tiles = (N + BN - 1) / BN;
    p.splits = splits;
    p.num_kb = num_kb;
    p.a_mn = a_mn_major ? 1 : 0;
    p.b_mn = b_mn_major ? 1 : 0;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.out0 = out0;
    p.out1 = out1;
    p.aux = aux;
    p.ldo = ldo;
    p.cluster = (g_gemm_multicast && max_ctas != 1) ? 2 : 1;
    p.two_sm = (p.cluster == 2 && g_gemm_multicast == 2) ? 1 : 0;
    if (p.conv == 2 && !p.two_sm) p.cluster = 1;  // the weight-gradient loads exist for the 2-SM and the single-CTA paths only
    const long long units = static_cast<long long>((p.num_m_tiles + p.cluster - 1) / p.cluster) * p.num_n_tiles * splits;
    int grid = persistent_ctas();
    if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
    int nclusters = grid / p.cluster;
    if (nclusters < 1) nclusters = 1;
    if (units < nclusters) nclusters = static_cast<int>(units);
    grid = nclusters * p.cluster;
    switch (epilogue) {
        case EPI_BIAS_BF16: return launch_gemm<EPI_BIAS_BF16>(tmA, tmB, tmBh, p, grid, stream);
        case EPI_BIAS_GELU_BF16: return launch_gemm<EPI_BIAS_GELU_BF16>(tmA, tmB, tmBh, p, grid, stream);
        case EPI_BIAS_RESID_F32: return launch_gemm<EPI_BIAS_RESID_F32>(tmA, tmB, tmBh, p, grid, stream);
        case EPI_DGELU_BF16: return launch_gemm<EPI_DGELU_BF16>(tmA, tmB, tmBh, p, grid, stream);
        case EPI_PARTIAL_F32: return launch_gemm<EPI_PARTIAL_F32>(tmA, tmB, tmBh, p, grid, stream);
        default: return set_error(B200_ERR_ARG, "gemm: unknown epilogue");
    }
}

extern "C" int b200_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                              int b_mn_major, int M, int N, int K, int epilogue, const void* bias, void* out0,
                              void* out1, const void* aux, long long ldo, int splits, int max_ctas,
                              cudaStream_t stream) {
    return gemm_dispatch(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, epilogue, bias, out0, out1, aux, ldo, splits, max_ctas, stream, nullptr);
}

// 3x3 convolution, stride 1, zero padding 1, on channels-last bf16 activations, as an implicit GEMM on the tcgen05 main loop:
//   out[b, y, x, co] = epilogue( sum_{ky,kx,ci} x[b, y+ky-1, x+kx-1, ci] * w[co, ky, kx, ci] )
// x: [B, H, W, Cin]; w_packed: [Cout, 9 * Cin] (tap-major: k = (ky * 3 + kx) * Cin + ci); out / aux: [B*H*W, ldo].
// M = B*H*W pixels, N = Cout, K = 9 * Cin; the im2col matrix is never materialised -- every (tap, 64-channel) k-block is one
// TMA box of the activation shifted by the tap, with the halo zero-filled by the TMA unit.
// Replaces Conv2d.forward -> F.conv2d (cflearn/modules/core/convs/basic.py:155-174) as the SD-v1.5 UNet uses it
// (convs/residual.py:179-186,199-205 conv1 / conv2; multimodal/diffusion/unet.py:213-217,269-272).  Its input gradient is the
// same call with the weights flipped and transposed ([Cin, 9 * Cout], tap (2-ky, 2-kx)).
extern "C" int b200_conv3x3_nhwc_bf16(const void* x, const void* w_packed, const void* bias, void* out0, const void* aux, long long ldo,
                                      int B, int H, int W, int Cin, int Cout, int epilogue, cudaStream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return set_error(B200_ERR_ARG, "conv3x3: non-positive size");
    if (Cin % BK != 0) return set_error(B200_ERR_ARG, "conv3x3: Cin must be a multiple of 64");
    if (W > BM || BM % W != 0) return set_error(B200_ERR_ARG, "conv3x3: the image width must divide 128");
    const int hb = H < BM / W ? H : BM / W;
    if (H % hb != 0 || BM % (W * hb) != 0) return set_error(B200_ERR_ARG, "conv3x3: H must be a multiple of 128 / W (or H * W must divide 128)");
    if (epilogue != EPI_BIAS_BF16 && epilogue != EPI_BIAS_RESID_F32) return set_error(B200_ERR_ARG, "conv3x3: epilogue must be BIAS_BF16 or BIAS_RESID_F32");
    ConvGeom cg{B, H, W, Cin, 1};
    const long long M = static_cast<long long>(B) * H * W;
    if (M > 0x7fffffffll) return set_error(B200_ERR_ARG, "conv3x3: too many pixels");
    return gemm_dispatch(x, Cin, 0, w_packed, 9ll * Cin, 0, static_cast<int>(M), Cout, 9 * Cin, epilogue, bias, out0, nullptr, aux, ldo, 1, 0,
                         stream, &cg);
}

// Weight gradient of the same convolution: dW[co][tap][ci] = sum over pixels of dy[pixel, co] * x[pixel shifted by tap, ci], as ONE
// split-K GEMM with M = Cout, N = 9 * Cin, K = B * H * W: A = dy MN-major (plain 2-D map), B = x MN-major behind the 4-D map, every
// 64-channel box of an N tile shifted by its own tap (halo zero-filled by TMA).  Writes `splits` fp32 partial planes
// [splits][Cout][9 * Cin] (b200_splitk_reduce finishes them); the plane layout IS the packed forward weight layout.
// Replaces the weight half of torch's conv2d backward for convs/basic.py:155-174 under the UNet blocks (convs/residual.py:179-205).
extern "C" int b200_conv3x3_wgrad_nhwc_bf16(const void* dy, long long ld_dy, const void* x, void* partials, int B, int H, int W, int Cin,
                                            int Cout, int splits, cudaStream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return set_error(B200_ERR_ARG, "conv3x3 wgrad: non-positive size");
    if (Cin % 64 != 0) return set_error(B200_ERR_ARG, "conv3x3 wgrad: Cin must be a multiple of 64");
    if (W > BK || BK % W != 0) return set_error(B200_ERR_ARG, "conv3x3 wgrad: the image width must divide 64");
    if ((static_cast<long long>(H) * W) % BK != 0) return set_error(B200_ERR_ARG, "conv3x3 wgrad: H * W must be a multiple of 64");
    const long long K = static_cast<long long>(B) * H * W;
    if (K > 0x7fffffffll) return set_error(B200_ERR_ARG, "conv3x3 wgrad: too many pixels");
    ConvGeom cg{B, H, W, Cin, 2};
    return gemm_dispatch(dy, ld_dy, 1, x, Cin, 1, Cout, 9 * Cin, static_cast<int>(K), EPI_PARTIAL_F32, nullptr, partials, nullptr, nullptr,
                         9ll * Cin, splits, 0, stream, &cg);
}
