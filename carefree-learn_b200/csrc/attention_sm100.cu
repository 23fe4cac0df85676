// Fused multi-head self-attention forward / backward for sm_100a on the PACKED qkv tensor [B, T, 3*H*64].
//
// Replaces, per transformer block of the reference:
//   qkv.chunk(3) + _to_heads permutes + .contiguous() copies   cflearn/modules/core/attentions.py:216,180-185,245
//   F.scaled_dot_product_attention (via sdp_attn)               cflearn/toolkit.py:911-974
//   output.transpose(1,2).contiguous()                          cflearn/modules/core/attentions.py:270
// and their autograd.  T <= 256 (197 for ViT-B/16, 50 / 77 for CLIP) so one KV tile covers the whole
// sequence: no online-softmax rescaling is needed and every score row lives in one TMEM lane.
//
// Forward, one CTA per (batch, head, 128-query tile), 2 CTAs / SM (18 warps):
//   TMA: Q[128x64], K[Tp x64], V[Tp x64] (3-D tensor map over qkv, rows >= T zero-filled)
//   tcgen05.mma  S = Q K^T            (128 x Tp x 64, fp32 in TMEM)
//   8 warps      softmax straight out of TMEM (tcgen05.ld): two threads per score row (one per column half, max / sum
//                exchanged through 2 KB of smem), P -> bf16 -> 128B-swizzled smem
//   tcgen05.mma  O = P V              (A = P K-major from smem, B = V MN-major from smem)
//   8 warps      O / rowsum -> bf16 -> out[B, T, H*64]  (already in the layout the out-projection GEMM reads)
//
// Backward, one CTA per (batch, head): outer loop over 128-key tiles, inner loop over 128-query tiles.
//   S = Q K^T, dP = dO V^T            (TMEM cols [0,128) and [128,256))
//   8 warps: P = exp2(S*c - lse), dS = P * (dP - delta) -> bf16 -> swizzled smem (one tile serves as K-major A
//            for dQ and as MN-major A for dK / dV: the 128B swizzle is purely address based)
//   dV += P^T dO, dK += dS^T Q        (TMEM cols [256,320), [320,384); accumulate over query tiles)
//   dQ_mt += dS K                     (TMEM cols [384,448), [448,512); accumulate over key tiles)
// Numerics mirror the flash kernels torch dispatches to: fp32 scores / softmax / row sums, P and dS rounded to
// bf16 only as MMA operands, softmax scale folded into exp2, dQ / dK scaled in fp32 at the end.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "b200_internal.h"
#include "ptx.cuh"

namespace b200 {

constexpr float kLog2e = 1.4426950408889634f;

// Writes this warp's 32 accumulator rows (TMEM lane == row, NCH chunks of 16 fp32 columns starting at `taddr`) to
// global memory as bf16, scaled by `sc`.  A row-per-thread store would touch 32 different 128-byte lines per
// instruction; instead the rows are transposed through a warp-private staging area (NCH * 1 KB, 16-byte chunks
// XOR-swizzled so both passes are bank-conflict free) and written with whole rows per group of lanes.
// `g0` addresses (row 0, first column); rows >= rows_valid are skipped.
// COLSUM (NCH == 4 only): csum[0..63] receives the fp32 column sums of the valid bf16 rows written (bias gradients).
template <int NCH, bool COLSUM = false>
__device__ __forceinline__ void store_rows_coalesced(uint8_t* stage, __nv_bfloat16* g0, long long row_stride, uint32_t taddr,
                                                     float sc, int rows_valid, int lane, float* csum = nullptr) {
    constexpr int CPR = NCH * 2;    // 16-byte chunks per row
    constexpr int RB = NCH * 32;    // bytes per row
    constexpr int RPI = 32 / CPR;   // rows written per store instruction
    const uint32_t ul = static_cast<uint32_t>(lane);
    const uint32_t sw = NCH == 4 ? (ul & 7u) : ((ul >> 1) & 3u);
    uint8_t* mine = stage + lane * RB;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(c * 16), v);
        tmem_ld_wait();
        uint4 o0, o1;
        o0.x = pack_bf16x2(__uint_as_float(v[0]) * sc, __uint_as_float(v[1]) * sc);
        o0.y = pack_bf16x2(__uint_as_float(v[2]) * sc, __uint_as_float(v[3]) * sc);
        o0.z = pack_bf16x2(__uint_as_float(v[4]) * sc, __uint_as_float(v[5]) * sc);
        o0.w = pack_bf16x2(__uint_as_float(v[6]) * sc, __uint_as_float(v[7]) * sc);
        o1.x = pack_bf16x2(__uint_as_float(v[8]) * sc, __uint_as_float(v[9]) * sc);
        o1.y = pack_bf16x2(__uint_as_float(v[10]) * sc, __uint_as_float(v[11]) * sc);
        o1.z = pack_bf16x2(__uint_as_float(v[12]) * sc, __uint_as_float(v[13]) * sc);
        o1.w = pack_bf16x2(__uint_as_float(v[14]) * sc, __uint_as_float(v[15]) * sc);
        *reinterpret_cast<uint4*>(mine + ((static_cast<uint32_t>(2 * c) ^ sw) << 4)) = o0;
        *reinterpret_cast<uint4*>(mine + ((static_cast<uint32_t>(2 * c + 1) ^ sw) << 4)) = o1;
    }
    __syncwarp();
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int it = 0; it < CPR; ++it) {
        const int row = it * RPI + lane / CPR;
        const uint32_t ch = ul % CPR;
        const uint32_t rsw = NCH == 4 ? (static_cast<uint32_t>(row) & 7u) : ((static_cast<uint32_t>(row) >> 1) & 3u);
        const uint4 val = *reinterpret_cast<const uint4*>(stage + row * RB + ((ch ^ rsw) << 4));
        if (row < rows_valid) {
            *reinterpret_cast<uint4*>(g0 + row * row_stride + ch * 8) = val;
            if constexpr (COLSUM) {
                acc[0] += bf16lo(val.x); acc[1] += bf16hi(val.x); acc[2] += bf16lo(val.y); acc[3] += bf16hi(val.y);
                acc[4] += bf16lo(val.z); acc[5] += bf16hi(val.z); acc[6] += bf16lo(val.w); acc[7] += bf16hi(val.w);
            }
        }
    }
    if constexpr (COLSUM) {
        static_assert(!COLSUM || NCH == 4, "column sums are implemented for 64-column tiles");
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // lanes l, l^8, l^16, l^24 hold the same 8 columns of different rows
            acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
            acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
        }
        if (lane < 8) {
            reinterpret_cast<float4*>(csum + lane * 8)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            reinterpret_cast<float4*>(csum + lane * 8)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
constexpr int AF_THREADS = 288;     // 8 softmax warps (two threads per score row) + 1 TMA/MMA warp
constexpr int AF_SQ = 0;            // 16 KB
constexpr int AF_SK = 16384;        // 32 KB
constexpr int AF_SP = 0;            // 64 KB, aliases Q and K (written only after S = QK^T has retired)
constexpr int AF_SV = 65536;        // 32 KB
constexpr int AF_XCH = 98304;       // 2 KB: per-row max / sum exchanged between the two column halves
constexpr int AF_BAR = AF_XCH + 2048;
constexpr int AF_SMEM = AF_BAR + 64 + 1024;

struct AttnParams {
    int B, T, H, D;   // D = H * 64
    int tp;           // T rounded up to a multiple of 16
    float scale;
    int causal;
    // version-2 kernels: cooperative L2 prefetch.  A head's Q / K / V tile is T pieces of 128 bytes at a stride of 3 * D * 2
    // bytes, and at any moment the 12 heads of an image are being fetched by 12 different SMs: DRAM sees 128-byte requests
    // scattered over many pages.  Each persistent CTA therefore prefetches, one item ahead, ITS 1/H share of the image's
    // CONTIGUOUS qkv rows (and of O / dO in the backward) into L2 -- whole DRAM pages -- and the strided TMA boxes that
    // follow hit in L2.  0 switches it off (A/B timing).
    int prefetch;
    const void* qkv_base;   // packed qkv [B, T, 3 * D] bf16
    const void* o_base;     // backward: O  [B, T, D] bf16
    const void* do_base;    // backward: dO [B, T, D] bf16
};

// 1 / H share `h` of the contiguous block [base + b * bytes_per_image, + bytes_per_image): 16-byte granular
__device__ __forceinline__ void prefetch_share(const void* base, long long bytes_per_image, int b, int h, int H) {
    long long chunk = (bytes_per_image / H + 15) / 16 * 16;
    const long long off = chunk * h;
    if (off >= bytes_per_image) return;
    if (off + chunk > bytes_per_image) chunk = (bytes_per_image - off) / 16 * 16;
    if (chunk <= 0) return;
    const char* p0 = reinterpret_cast<const char*>(base) + static_cast<long long>(b) * bytes_per_image + off;
    while (chunk > 0) {  // (the size operand is 32 bits; keep single requests modest)
        const uint32_t n = chunk > 65536 ? 65536u : static_cast<uint32_t>(chunk);
        bulk_prefetch_l2(p0, n);
        p0 += n;
        chunk -= n;
    }
}

__global__ void __launch_bounds__(AF_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment as an OFFSET from the __shared__ array: a round trip through an integer hides the address space
    // from the compiler and turns every staging access into a generic LD / ST (long-scoreboard latency)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem + AF_SQ;
    uint8_t* sK = smem + AF_SK;
    uint8_t* sP = smem + AF_SP;
    uint8_t* sV = smem + AF_SV;
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + AF_BAR);
    uint64_t* bar_s = bar_load + 1;
    uint64_t* bar_p = bar_load + 2;
    uint64_t* bar_o = bar_load + 3;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_load + 4);
    float* xch_max = reinterpret_cast<float*>(smem + AF_XCH);  // [2][128]
    float* xch_sum = xch_max + 256;                            // [2][128]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_mt = (p.T + 127) / 128;
    const int mt = blockIdx.x % n_mt;
    const int bh = blockIdx.x / n_mt;
    const int h = bh % p.H;
    const int b = bh / p.H;

    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmKV);
            mbar_init(bar_load, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_p, 8);  // one arrival per softmax warp (256 per-thread arrivals serialise on the barrier word)
            mbar_init(bar_o, 1);
            fence_mbar_init();
            // issue the loads right away: they overlap the TMEM allocation and the CTA-wide sync below
            mbar_expect_tx(bar_load, 16384u + 2u * static_cast<uint32_t>(p.tp) * 128u);
            tma_load_3d(sQ, &tmQ, bar_load, h * 64, mt * 128, b);
            tma_load_3d(sK, &tmKV, bar_load, p.D + h * 64, 0, b);
            tma_load_3d(sV, &tmKV, bar_load, 2 * p.D + h * 64, 0, b);
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 256);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 8) {
        if (elect_one()) {
            mbar_wait(bar_load, 0);
            tc_fence_after_sync();
            const uint32_t idesc_s = make_idesc_bf16(128, static_cast<uint32_t>(p.tp), 0, 0);
            const uint32_t q_base = smem_u32(sQ), k_base = smem_u32(sK);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                umma_bf16(tmem_base, make_smem_desc(q_base + kk * 32, 0, 1024, kSwz128),
                          make_smem_desc(k_base + kk * 32, 0, 1024, kSwz128), idesc_s, kk > 0 ? 1u : 0u);
            umma_commit(bar_s);
            mbar_wait(bar_p, 0);
            tc_fence_after_sync();
            const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
            const uint32_t p_base = smem_u32(sP), v_base = smem_u32(sV);
            const int nkk = p.tp / 16;
            for (int kk = 0; kk < nkk; ++kk)
                umma_bf16(tmem_base, make_smem_desc(p_base + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024, kSwz128),
                          make_smem_desc(v_base + kk * 2048, 0, 1024, kSwz128), idesc_o, kk > 0 ? 1u : 0u);
            umma_commit(bar_o);
        }
    } else {
        const uint32_t q = static_cast<uint32_t>(warp & 3);  // TMEM lane quadrant
        const int half = warp >> 2;                           // which half of the score columns this thread owns
        const int r = static_cast<int>(q) * 32 + lane;        // row inside the query tile == TMEM lane
        const int i = mt * 128 + r;                           // query index
        const uint32_t taddr = tmem_base + ((q * 32u) << 16);
        const float sl2 = p.scale * kLog2e;
        int nvalid = p.T;
        if (p.causal && i + 1 < nvalid) nvalid = i + 1;
        const int nchunk = p.tp / 16;
        const int nc0 = (nchunk + 1) / 2;
        // warps whose 32 query rows are all padding (T = 197: 59 of the second tile's 128 rows) only keep the barrier
        // protocol: their P rows feed output rows that are never stored
        const bool dead = mt * 128 + static_cast<int>(q) * 32 >= p.T;
        const int c_begin = dead ? 0 : (half ? nc0 : 0), c_end = dead ? 0 : (half ? nchunk : nc0);
        mbar_wait(bar_s, 0);
        tc_fence_after_sync();
        // pass 1: row max over this thread's column half, then combine the two halves
        float m = -INFINITY;
        for (int c = c_begin; c < c_end; ++c) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(c * 16), v);
            tmem_ld_wait();
            if (!p.causal && c * 16 + 16 <= p.T) {  // warp-uniform: whole chunk valid, no per-element masking
#pragma unroll
                for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(v[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c * 16 + j < nvalid) m = fmaxf(m, __uint_as_float(v[j]));
            }
        }
        xch_max[half * 128 + r] = m;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        m = fmaxf(xch_max[r], xch_max[128 + r]);
        const float m2 = m * sl2;
        // pass 2: p = exp2(s*c - m*c), row sum, P -> smem (bf16, K-major 128B swizzle)
        float sum = 0.f;
        for (int c = c_begin; c < c_end; ++c) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(c * 16), v);
            tmem_ld_wait();
            float pv[16];
            if (!p.causal && c * 16 + 16 <= p.T) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    pv[j] = fast_ex2(fmaf(__uint_as_float(v[j]), sl2, -m2));
                    sum += pv[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float e = fast_ex2(fmaf(__uint_as_float(v[j]), sl2, -m2));
                    pv[j] = (c * 16 + j < nvalid) ? e : 0.f;
                    sum += pv[j];
                }
            }
            const int col = c * 16;
            uint8_t* blk = sP + (col >> 6) * 16384 + r * 128;
            const uint32_t ch = static_cast<uint32_t>((col & 63) >> 3);
            *reinterpret_cast<uint4*>(blk + (((ch) ^ (r & 7u)) << 4)) =
                make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7]));
            *reinterpret_cast<uint4*>(blk + (((ch + 1) ^ (r & 7u)) << 4)) =
                make_uint4(pack_bf16x2(pv[8], pv[9]), pack_bf16x2(pv[10], pv[11]), pack_bf16x2(pv[12], pv[13]), pack_bf16x2(pv[14], pv[15]));
        }
        xch_sum[half * 128 + r] = sum;
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        sum = xch_sum[r] + xch_sum[128 + r];
        // O = P V
        mbar_wait(bar_o, 0);
        tc_fence_after_sync();
        const float inv = 1.0f / sum;
        const bool valid = i < p.T;
        {
            // every MMA reading sQ / sK / sP has retired (bar_o): the front of the buffer is free for staging
            const int row0 = mt * 128 + static_cast<int>(q) * 32;
            __nv_bfloat16* g0 = out + (static_cast<long long>(b) * p.T + row0) * p.D + h * 64 + half * 32;
            if (!dead) store_rows_coalesced<2>(smem + warp * 2048, g0, p.D, taddr + static_cast<uint32_t>(half * 32), inv, p.T - row0, lane);
        }
        if (valid && half == 0) lse_out[(static_cast<long long>(b) * p.H + h) * p.T + i] = m * p.scale + logf(sum);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------
// forward, version 2: persistent + warp-specialised, probabilities kept in TENSOR MEMORY
// ------------------------------------------------------------------------------------------------
// Version 1 above is one short-lived CTA per (batch, head, query tile): every CTA pays launch + TMEM allocation + the
// full TMA latency before its first MMA, and its load -> S -> softmax -> P -> O -> store chain is serial (ncu, round 1:
// tensor pipe 14 %, MUFU 30 %, 2.2 TB/s -- nothing saturated; 132 us against torch SDPA's 102 us).  Version 2:
//   * ONE persistent CTA per SM walks over (batch, head) items; a 2-stage TMA ring prefetches the next item's Q / K / V
//     (K and V are fetched once per item, not once per query tile) while the current one is being processed;
//   * work unit = one 128-query tile; two TMEM slots of 256 columns alternate between units, so the tensor pipe computes
//     S(u+1) = Q K^T while the softmax warps are busy with unit u, and O(u) = P V while they are busy with unit u+1;
//   * 8 softmax warps (two threads per score row, ONE pass over TMEM: each thread keeps its half row in registers --
//     setmaxnreg moves the registers of the idle / light warpgroups to them) write the bf16 probabilities back into the
//     columns of S with tcgen05.st, and O = P V takes its A operand straight from TMEM (tcgen05.mma [d], [a_tmem], b_desc):
//     P never goes through shared memory, which is what frees the 64 KB the second ring stage needs;
//   * 4 epilogue warps drain O (TMEM -> registers -> the unit's own, by then dead, Q tile as staging -> coalesced 16-byte
//     global stores), so the softmax warps never wait for the PV product.
// Warpgroups: 0 = {TMA producer, MMA issuer, 2 idle warps}, 1 and 2 = softmax, 3 = epilogue.
constexpr int F2_THREADS = 512;
constexpr int F2_Q = 0;                    // 2 x 16 KB  (query tiles 0 / 1 of the item)
constexpr int F2_K = 32768;                // 32 KB      (up to 256 keys)
constexpr int F2_V = 65536;                // 32 KB
constexpr int F2_STAGE = 98304;
constexpr int F2_STATS = 2 * F2_STAGE;     // per TMEM slot: xmax[2][128], psum[2][128], rmax[128]  (floats)
constexpr int F2_STATS_SLOT = (256 + 256 + 128) * 4;
constexpr int F2_BAR = F2_STATS + 2 * F2_STATS_SLOT;
constexpr int F2_SMEM = F2_BAR + 256 + 1024;
constexpr uint32_t F2_SLOT_COLS = 256;     // S / P at column 0 of the slot, O at column 192 (P never exceeds 128 columns)
constexpr uint32_t F2_O_COL = 192;

// this warp's 32 accumulator rows (TMEM lane == row) x 64 fp32 columns -> bf16 * sc -> warp-private 4 KB staging area
// (row r at r * 128 B, its eight 16-byte chunks XOR-swizzled with r & 7: both passes are bank-conflict free)
// 32 rows x 64 fp32 columns of tensor memory (this warp's lanes) -> scaled, packed bf16 in registers (one row per lane)
__device__ __forceinline__ void pack_rows64(uint4 (&o)[8], uint32_t taddr, float sc) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {  // two loads in flight per wait (four would need 64 raw registers on top of the packed rows)
        uint32_t v[2][16];
        tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(hh * 32), v[0]);
        tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(hh * 32 + 16), v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = hh * 2 + cc;
            o[2 * c].x = pack_bf16x2(__uint_as_float(v[cc][0]) * sc, __uint_as_float(v[cc][1]) * sc);
            o[2 * c].y = pack_bf16x2(__uint_as_float(v[cc][2]) * sc, __uint_as_float(v[cc][3]) * sc);
            o[2 * c].z = pack_bf16x2(__uint_as_float(v[cc][4]) * sc, __uint_as_float(v[cc][5]) * sc);
            o[2 * c].w = pack_bf16x2(__uint_as_float(v[cc][6]) * sc, __uint_as_float(v[cc][7]) * sc);
            o[2 * c + 1].x = pack_bf16x2(__uint_as_float(v[cc][8]) * sc, __uint_as_float(v[cc][9]) * sc);
            o[2 * c + 1].y = pack_bf16x2(__uint_as_float(v[cc][10]) * sc, __uint_as_float(v[cc][11]) * sc);
            o[2 * c + 1].z = pack_bf16x2(__uint_as_float(v[cc][12]) * sc, __uint_as_float(v[cc][13]) * sc);
            o[2 * c + 1].w = pack_bf16x2(__uint_as_float(v[cc][14]) * sc, __uint_as_float(v[cc][15]) * sc);
        }
    }
}
// ... -> the warp's 4 KB staging tile (128-byte rows, 16-byte chunks XOR-swizzled by the row)
__device__ __forceinline__ void store_rows64(uint8_t* stage, const uint4 (&o)[8], int lane) {
    uint8_t* mine = stage + lane * 128;
    const uint32_t sw = static_cast<uint32_t>(lane) & 7u;
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(mine + ((static_cast<uint32_t>(c) ^ sw) << 4)) = o[c];
}
__device__ __forceinline__ void stage_rows64(uint8_t* stage, uint32_t taddr, float sc, int lane) {
    uint4 o[8];
    pack_rows64(o, taddr, sc);
    store_rows64(stage, o, lane);
}
// staged rows -> global: 8 lanes x 16 B write one whole 128-byte row, 4 rows per instruction
__device__ __forceinline__ void flush_rows64(const uint8_t* stage, __nv_bfloat16* g0, long long row_stride, int rows_valid, int lane) {
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3);
        const uint32_t ch = static_cast<uint32_t>(lane) & 7u;
        const uint4 val = *reinterpret_cast<const uint4*>(stage + row * 128 + ((ch ^ (static_cast<uint32_t>(row) & 7u)) << 4));
        if (row < rows_valid) *reinterpret_cast<uint4*>(g0 + row * row_stride + ch * 8) = val;
    }
    __syncwarp();
}

// One thread's share of the softmax of a unit: `n` chunks of 16 score columns starting at chunk `cb` of TMEM row `taddr`.
// TWO passes over tensor memory (row max, then exp / sum / pack) as real loops with a ~100-instruction body.  History: the first
// cut kept the row in registers with `if (c < nmine)` branches (35 KB body, serial max / sum chains); the second cut was
// straight-line code per compile-time chunk count -- 15 000 SASS instructions (240 KB) for the kernel, and ncu showed the
// softmax warps, which ARE the critical path, at one instruction per ~5.5 cycles with `no instruction` (I-cache miss) as the top
// stall.  Tensor memory is read twice instead (16 TB/s: free), the loops stay resident in the instruction cache, the max uses
// the 3-input max, and scale / subtract / sum run on the packed fp32x2 pipe.
// P goes back over the thread's OWN score columns: packed chunk c -> columns (cb * 16 + c * 8 ...), inside S chunk cb + c / 2,
// which this thread has already consumed when it stores (ascending c) -- no other thread's unread scores are touched, so the
// only cross-thread step is the exchange of the two half-row maxima.  The PV MMA addresses the two P regions separately.
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
    f32x2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// The masked and the unmasked chunk are SEPARATE code paths behind a warp-uniform branch: with the column test inside the
// unrolled element loop ptxas if-converted it and every element paid ISETP + FSEL (ncu, third cut: 15 of 69 M warp instructions).
template <bool MASK>
__device__ __forceinline__ void softmax_max_chunk(const uint32_t (&v)[16], int col0, int nvalid, float& mx0, float& mx1) {
    if (!MASK) {
#pragma unroll
        for (int jj = 0; jj < 16; jj += 4) {
            mx0 = max3(mx0, __uint_as_float(v[jj]), __uint_as_float(v[jj + 1]));
            mx1 = max3(mx1, __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
        }
    } else {
#pragma unroll
        for (int jj = 0; jj < 16; jj += 2) {
            mx0 = fmaxf(mx0, (col0 + jj < nvalid) ? __uint_as_float(v[jj]) : -INFINITY);
            mx1 = fmaxf(mx1, (col0 + jj + 1 < nvalid) ? __uint_as_float(v[jj + 1]) : -INFINITY);
        }
    }
}
template <bool MASK>
__device__ __forceinline__ void softmax_exp_chunk(const uint32_t (&v)[16], int col0, int nvalid, f32x2_t sl22, f32x2_t nm22,
                                                  f32x2_t& sum2a, f32x2_t& sum2b, uint32_t (&pk)[8]) {
#pragma unroll
    for (int jj = 0; jj < 16; jj += 2) {
        float x0, x1;
        upk2(fma2(pk2(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1])), sl22, nm22), x0, x1);
        float e0 = fast_ex2(x0), e1 = fast_ex2(x1);
        if (MASK) {
            if (col0 + jj >= nvalid) e0 = 0.f;
            if (col0 + jj + 1 >= nvalid) e1 = 0.f;
        }
        if (jj & 2) sum2b = add2(sum2b, pk2(e0, e1)); else sum2a = add2(sum2a, pk2(e0, e1));
        pk[jj >> 1] = pack_bf16x2(e0, e1);
    }
}
__device__ __forceinline__ void softmax_rows(const int n, const uint32_t taddr, const int cb, const int nvalid, const int nvalid_warp_min,
                                             const float sl2, float* stats, const int hf, const int r) {
    // chunks [0, nfull) need no mask for any row of this warp
    const int nfull = max(0, min(n, nvalid_warp_min / 16 - cb));
    // ---- pass 1: row max over this thread's columns
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < nfull; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>((cb + c) * 16), v);
        tmem_ld_wait();
        softmax_max_chunk<false>(v, 0, 0, mx0, mx1);
    }
#pragma unroll 1
    for (int c = nfull; c < n; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>((cb + c) * 16), v);
        tmem_ld_wait();
        softmax_max_chunk<true>(v, (cb + c) * 16, nvalid, mx0, mx1);
    }
    float m = fmaxf(mx0, mx1);
    stats[hf * 128 + r] = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    m = fmaxf(stats[r], stats[128 + r]);
    // ---- pass 2: p = exp2(s * sl2 - m * sl2), row sum, bf16 pack, store over the consumed columns
    const f32x2_t sl22 = pk2(sl2, sl2);
    const float nm2 = -m * sl2;
    const f32x2_t nm22 = pk2(nm2, nm2);
    f32x2_t sum2a = pk2(0.f, 0.f), sum2b = pk2(0.f, 0.f);
    const uint32_t pbase = taddr + static_cast<uint32_t>(cb * 16);
#pragma unroll 1
    for (int c = 0; c < nfull; ++c) {
        uint32_t v[16], pk[8];
        tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>((cb + c) * 16), v);
        tmem_ld_wait();
        softmax_exp_chunk<false>(v, 0, 0, sl22, nm22, sum2a, sum2b, pk);
        tmem_st_32x32b_x8(pbase + static_cast<uint32_t>(c * 8), pk);  // P[row, col0 .. col0+16) -> 8 packed columns
    }
#pragma unroll 1
    for (int c = nfull; c < n; ++c) {
        uint32_t v[16], pk[8];
        tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>((cb + c) * 16), v);
        tmem_ld_wait();
        softmax_exp_chunk<true>(v, (cb + c) * 16, nvalid, sl22, nm22, sum2a, sum2b, pk);
        tmem_st_32x32b_x8(pbase + static_cast<uint32_t>(c * 8), pk);
    }
    tmem_st_wait();
    float s0, s1;
    upk2(add2(sum2a, sum2b), s0, s1);
    stats[256 + hf * 128 + r] = s0 + s1;
    if (hf == 0) stats[512 + r] = m;
}

__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out, const AttnParams p, const int n_items) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + F2_BAR);  // [2] stage loaded
    uint64_t* empty_bar = full_bar + 2;                               // [2] stage free (epilogue warps of every unit of the item)
    uint64_t* bar_s = full_bar + 4;                                   // [2] S of the slot's unit is in TMEM
    uint64_t* bar_p = full_bar + 6;                                   // [2] P written (8 softmax warps)
    uint64_t* bar_o = full_bar + 8;                                   // [2] O complete
    uint64_t* bar_stats = full_bar + 10;                              // [2] row max / partial sums in smem (8 softmax warps)
    uint64_t* slot_free = full_bar + 12;                              // [2] O read out of TMEM (4 epilogue warps)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(full_bar + 14);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_mt = (p.T + 127) / 128;
    const int n_my = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int n_units = n_my * n_mt;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], static_cast<uint32_t>(4 * n_mt));
            mbar_init(&bar_s[i], 1);
            mbar_init(&bar_p[i], 8);
            mbar_init(&bar_o[i], 1);
            mbar_init(&bar_stats[i], 8);
            mbar_init(&slot_free[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int wg = warp >> 2;

    if (wg == 0) {
        setmaxnreg_dec<40>();
        if (warp == 0) {
            // ===================================== TMA producer =====================================
            if (elect_one()) {
                const uint32_t bytes = static_cast<uint32_t>(n_mt) * 16384u + 2u * static_cast<uint32_t>(p.tp) * 128u;
                for (int it = 0; it < n_my; ++it) {
                    const int s = it & 1;
                    mbar_wait(&empty_bar[s], static_cast<uint32_t>((it >> 1) & 1) ^ 1u);
                    const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
                    const int h = w % p.H, b = w / p.H;
                    if (p.prefetch && it + 1 < n_my) {  // this CTA's share of the NEXT item's image, as whole contiguous rows
                        const int w1 = w + static_cast<int>(gridDim.x);
                        prefetch_share(p.qkv_base, static_cast<long long>(p.T) * 3 * p.D * 2, w1 / p.H, w1 % p.H, p.H);
                    }
                    uint8_t* st = smem + s * F2_STAGE;
                    mbar_expect_tx(&full_bar[s], bytes);
                    for (int mt = 0; mt < n_mt; ++mt) tma_load_3d(st + F2_Q + mt * 16384, &tmQ, &full_bar[s], h * 64, mt * 128, b);
                    tma_load_3d(st + F2_K, &tmKV, &full_bar[s], p.D + h * 64, 0, b);
                    tma_load_3d(st + F2_V, &tmKV, &full_bar[s], 2 * p.D + h * 64, 0, b);
                }
            }
        } else if (warp == 1) {
            // ===================================== MMA issuer =======================================
            if (elect_one()) {
                const uint32_t idesc_s = make_idesc_bf16(128, static_cast<uint32_t>(p.tp), 0, 0);
                const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
                const int nkk = p.tp / 16;
                const int nc0 = (nkk + 1) / 2;  // chunks of the first half-row thread (see the softmax warps)
                const uint64_t dsc_v = make_smem_desc(smem_u32(smem + F2_V), 0, 1024, kSwz128);
                auto issue_s = [&](int j) {
                    const int it = n_mt == 2 ? (j >> 1) : j, mt = n_mt == 2 ? (j & 1) : 0, s = it & 1, slot = j & 1, k = j >> 1;
                    if (mt == 0) mbar_wait(&full_bar[s], static_cast<uint32_t>((it >> 1) & 1));
                    mbar_wait(&slot_free[slot], static_cast<uint32_t>(k & 1) ^ 1u);  // O of unit j-2 has left the slot
                    tc_fence_after_sync();
                    const uint32_t q_base = smem_u32(smem + s * F2_STAGE + F2_Q + mt * 16384);
                    const uint32_t k_base = smem_u32(smem + s * F2_STAGE + F2_K);
                    const uint32_t d = tmem_base + static_cast<uint32_t>(slot) * F2_SLOT_COLS;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(d, make_smem_desc(q_base + kk * 32, 0, 1024, kSwz128), make_smem_desc(k_base + kk * 32, 0, 1024, kSwz128),
                                  idesc_s, kk > 0 ? 1u : 0u);
                    umma_commit(&bar_s[slot]);
                };
                if (n_units > 0) issue_s(0);
                if (n_units > 1) issue_s(1);
                for (int j = 0; j < n_units; ++j) {
                    const int it = n_mt == 2 ? (j >> 1) : j, s = it & 1, slot = j & 1, k = j >> 1;
                    mbar_wait(&bar_p[slot], static_cast<uint32_t>(k & 1));
                    tc_fence_after_sync();
                    const uint32_t base = tmem_base + static_cast<uint32_t>(slot) * F2_SLOT_COLS;
                    // A = P[128 x 16] from TMEM (8 packed columns per k-step), B = V MN-major.  The two half-row threads of the
                    // softmax each wrote their packed chunks at the start of their own score columns (softmax_rows):
                    // k-steps [0, nc0) at column kk * 8, k-steps [nc0, nkk) at column nc0 * 16 + (kk - nc0) * 8.
                    for (int kk = 0; kk < nkk; ++kk) {
                        const uint32_t pcol = static_cast<uint32_t>(kk < nc0 ? kk * 8 : nc0 * 8 + kk * 8);
                        umma_bf16_ts(base + F2_O_COL, base + pcol, desc_off(dsc_v, static_cast<uint32_t>(s) * (F2_STAGE >> 4) + kk * 128),
                                     idesc_o, kk > 0 ? 1u : 0u);
                    }
                    umma_commit(&bar_o[slot]);
                    if (j + 2 < n_units) issue_s(j + 2);  // runs under the softmax of unit j+1
                }
            }
        }
    } else if (wg == 3) {
        // ===================================== epilogue warps =====================================
        setmaxnreg_dec<72>();
        const uint32_t q = static_cast<uint32_t>(warp & 3);
        const int r = static_cast<int>(q) * 32 + lane;
        for (int j = 0; j < n_units; ++j) {
            const int it = n_mt == 2 ? (j >> 1) : j, mt = n_mt == 2 ? (j & 1) : 0, s = it & 1, slot = j & 1, k = j >> 1;
            const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int h = w % p.H, b = w / p.H;
            const int row0 = mt * 128 + static_cast<int>(q) * 32;
            const bool dead = row0 >= p.T;
            const float* stats = reinterpret_cast<const float*>(smem + F2_STATS + slot * F2_STATS_SLOT);
            mbar_wait(&bar_stats[slot], static_cast<uint32_t>(k & 1));
            const float sum = stats[256 + r] + stats[384 + r];
            const float m = stats[512 + r];
            mbar_wait(&bar_o[slot], static_cast<uint32_t>(k & 1));
            tc_fence_after_sync();
            uint8_t* stage = smem + s * F2_STAGE + F2_Q + mt * 16384 + static_cast<int>(q) * 4096;  // the unit's own Q tile: dead since S
            if (!dead)
                stage_rows64(stage, tmem_base + ((q * 32u) << 16) + static_cast<uint32_t>(slot) * F2_SLOT_COLS + F2_O_COL, 1.0f / sum, lane);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_free[slot]);
            if (!dead) {
                __nv_bfloat16* g0 = out + (static_cast<long long>(b) * p.T + row0) * p.D + h * 64;
                flush_rows64(stage, g0, p.D, p.T - row0, lane);
                const int i = row0 + lane;
                if (i < p.T) lse_out[(static_cast<long long>(b) * p.H + h) * p.T + i] = m * p.scale + logf(sum);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
    } else {
        // ===================================== softmax warps ======================================
        setmaxnreg_inc<200>();
        const uint32_t q = static_cast<uint32_t>(warp & 3);
        const int hf = (warp - 4) >> 2;                    // which half of the score columns this thread owns
        const int r = static_cast<int>(q) * 32 + lane;     // row of the query tile == TMEM lane
        const float sl2 = p.scale * kLog2e;
        const int nchunk = p.tp / 16;
        const int nc0 = (nchunk + 1) / 2;
        const int cb = hf ? nc0 : 0;
        const int nmine = hf ? nchunk - nc0 : nc0;          // <= 8 chunks of 16 columns
        for (int j = 0; j < n_units; ++j) {
            const int it = n_mt == 2 ? (j >> 1) : j, mt = n_mt == 2 ? (j & 1) : 0, slot = j & 1, k = j >> 1;
            const int i = mt * 128 + r;
            const bool dead = mt * 128 + static_cast<int>(q) * 32 >= p.T;  // all 32 rows of this warp are padding
            float* stats = reinterpret_cast<float*>(smem + F2_STATS + slot * F2_STATS_SLOT);
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + static_cast<uint32_t>(slot) * F2_SLOT_COLS;
            int nvalid = p.T;
            if (p.causal && i + 1 < nvalid) nvalid = i + 1;
            mbar_wait(&bar_s[slot], static_cast<uint32_t>(k & 1));
            tc_fence_after_sync();
            if (dead || nmine == 0) {  // no columns of this thread take part: keep the exchange / barrier protocol only
                stats[hf * 128 + r] = -INFINITY;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                stats[256 + hf * 128 + r] = 0.f;
                if (hf == 0) stats[512 + r] = fmaxf(stats[r], stats[128 + r]);
            } else {
                // columns below nvalid_min need no mask for ANY row of this warp (causal: the warp's first row sees the fewest keys)
                const int row_first = mt * 128 + static_cast<int>(q) * 32;
                const int nvalid_min = p.causal ? min(p.T, row_first + 1) : p.T;
                softmax_rows(nmine, taddr, cb, nvalid, nvalid_min, sl2, stats, hf, r);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_p[slot]);
                mbar_arrive(&bar_stats[slot]);
            }
        }
    }
    __syncwarp();
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
constexpr int AB_THREADS = 288;
constexpr int AB_SQ = 0;        // 2 x 16 KB (query tiles 0,1)
constexpr int AB_SDO = 32768;   // 2 x 16 KB
constexpr int AB_SK = 65536;    // 2 x 16 KB (key tiles 0,1: resident, a reload at the tile switch stalled every warp ~1 us)
constexpr int AB_SV = 98304;    // 2 x 16 KB
constexpr int AB_SP = 131072;   // 32 KB: P  [128 q x 128 keys] as two 64-key blocks
constexpr int AB_SDS = 163840;  // 32 KB: dS
constexpr int AB_LSE = 196608;  // 256 floats
constexpr int AB_DELTA = AB_LSE + 1024;
constexpr int AB_CSUM = AB_DELTA + 1024;  // 24 x 64 floats: column sums of dQ / dK / dV per (tile, lane quadrant)
constexpr int AB_BAR = AB_CSUM + 24 * 64 * 4;
constexpr int AB_SMEM = AB_BAR + 128 + 1024;

constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DK = 256, TM_DV = 320, TM_DQ = 384;

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __nv_bfloat16* __restrict__ o_in, const __nv_bfloat16* __restrict__ do_in,
                const float* __restrict__ lse_in, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias_part,
                const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment as an OFFSET from the __shared__ array: a round trip through an integer hides the address space
    // from the compiler and turns every staging access into a generic LD / ST (long-scoreboard latency)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* csum = reinterpret_cast<float*>(smem + AB_CSUM);
    uint8_t* sQ = smem + AB_SQ;
    uint8_t* sDO = smem + AB_SDO;
    uint8_t* sK = smem + AB_SK;
    uint8_t* sV = smem + AB_SV;
    uint8_t* sP = smem + AB_SP;
    uint8_t* sDS = smem + AB_SDS;
    float* lse_s = reinterpret_cast<float*>(smem + AB_LSE);
    float* delta_s = reinterpret_cast<float*>(smem + AB_DELTA);
    uint64_t* bar_qdo = reinterpret_cast<uint64_t*>(smem + AB_BAR);
    uint64_t* bar_kvload = bar_qdo + 1;
    uint64_t* bar_s = bar_qdo + 2;     // S of this (key, query) tile pair is in TMEM
    uint64_t* bar_sdp = bar_qdo + 3;   // ... and so is dP
    uint64_t* bar_pds = bar_qdo + 4;
    uint64_t* bar_kv = bar_qdo + 5;
    uint64_t* bar_kvfree = bar_qdo + 6;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_qdo + 7);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_mt = (p.T + 127) / 128;
    const int n_kt = n_mt;
    const int h = blockIdx.x % p.H;
    const int b = blockIdx.x / p.H;
    const long long D3 = 3ll * p.D;

    // rows of O and dO for delta_i = sum_d dO[i,d] * O[i,d]: the loads are issued before the CTA-wide sync and consumed
    // after it, so their latency overlaps barrier init, the TMEM allocation, the TMA loads and the first S / dP MMAs
    uint4 ro[8], rg[8];
    float ls = 0.f;
    const int irow = threadIdx.x;
    const bool has_row = warp < 8 && irow < p.T;
    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQKV);
            tma_prefetch_desc(&tmDO);
            mbar_init(bar_qdo, 1);
            mbar_init(bar_kvload, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_sdp, 1);
            mbar_init(bar_pds, 8);     // one arrival per compute warp
            mbar_init(bar_kv, 1);
            mbar_init(bar_kvfree, 8);
            fence_mbar_init();
            mbar_expect_tx(bar_qdo, static_cast<uint32_t>(n_mt) * 2u * 16384u);
            for (int mt = 0; mt < n_mt; ++mt) {
                tma_load_3d(sQ + mt * 16384, &tmQKV, bar_qdo, h * 64, mt * 128, b);
                tma_load_3d(sDO + mt * 16384, &tmDO, bar_qdo, h * 64, mt * 128, b);
            }
            mbar_expect_tx(bar_kvload, static_cast<uint32_t>(n_kt) * 2u * 16384u);
            for (int kt = 0; kt < n_kt; ++kt) {
                tma_load_3d(sK + kt * 16384, &tmQKV, bar_kvload, p.D + h * 64, kt * 128, b);
                tma_load_3d(sV + kt * 16384, &tmQKV, bar_kvload, 2 * p.D + h * 64, kt * 128, b);
            }
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    } else if (has_row) {
        const uint4* po = reinterpret_cast<const uint4*>(o_in + (static_cast<long long>(b) * p.T + irow) * p.D + h * 64);
        const uint4* pd = reinterpret_cast<const uint4*>(do_in + (static_cast<long long>(b) * p.T + irow) * p.D + h * 64);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            ro[k] = __ldg(po + k);
            rg[k] = __ldg(pd + k);
        }
        ls = __ldg(lse_in + (static_cast<long long>(b) * p.H + h) * p.T + irow);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 8) {
        if (elect_one()) {
            const uint32_t idesc_tt = make_idesc_bf16(128, 64, 1, 1);   // dV, dK  : A MN-major, B MN-major
            const uint32_t idesc_nt = make_idesc_bf16(128, 64, 0, 1);   // dQ      : A K-major,  B MN-major
            const uint32_t p_base = smem_u32(sP), ds_base = smem_u32(sDS);
            uint32_t it = 0;
            mbar_wait(bar_qdo, 0);
            mbar_wait(bar_kvload, 0);
            tc_fence_after_sync();
            for (int kt = 0; kt < n_kt; ++kt) {
                const uint32_t k_base = smem_u32(sK + kt * 16384), v_base = smem_u32(sV + kt * 16384);
                // only the (16-padded) valid keys / queries of a tile enter the MMAs: T = 197 leaves 80 of 128 in tile 1
                const int nk = min(128, p.tp - kt * 128);
                const int nkc = nk >> 4;
                const uint32_t idesc_nn = make_idesc_bf16(128, static_cast<uint32_t>(nk), 0, 0);  // S, dP: A, B K-major
                for (int mt = 0; mt < n_mt; ++mt, ++it) {
                    const int nqc = min(128, p.tp - mt * 128) >> 4;
                    const uint32_t q_base = smem_u32(sQ + mt * 16384), do_base = smem_u32(sDO + mt * 16384);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem_base + TM_S, make_smem_desc(q_base + kk * 32, 0, 1024, kSwz128),
                                  make_smem_desc(k_base + kk * 32, 0, 1024, kSwz128), idesc_nn, kk > 0 ? 1u : 0u);
                    umma_commit(bar_s);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem_base + TM_DP, make_smem_desc(do_base + kk * 32, 0, 1024, kSwz128),
                                  make_smem_desc(v_base + kk * 32, 0, 1024, kSwz128), idesc_nn, kk > 0 ? 1u : 0u);
                    umma_commit(bar_sdp);
                    mbar_wait(bar_pds, it & 1u);
                    tc_fence_after_sync();
                    if (mt == 0 && kt > 0) {
                        mbar_wait(bar_kvfree, static_cast<uint32_t>(kt - 1) & 1u);  // dK / dV of the previous key tile read out
                        tc_fence_after_sync();
                    }
                    for (int kk = 0; kk < nqc; ++kk) {  // reduction over the queries of this tile, 16 per step
                        const uint64_t a_p = make_smem_desc(p_base + kk * 2048, 16384, 1024, kSwz128);
                        const uint64_t a_ds = make_smem_desc(ds_base + kk * 2048, 16384, 1024, kSwz128);
                        const uint64_t b_do = make_smem_desc(do_base + kk * 2048, 0, 1024, kSwz128);
                        const uint64_t b_q = make_smem_desc(q_base + kk * 2048, 0, 1024, kSwz128);
                        const uint32_t acc = (mt > 0 || kk > 0) ? 1u : 0u;
                        umma_bf16(tmem_base + TM_DV, a_p, b_do, idesc_tt, acc);
                        umma_bf16(tmem_base + TM_DK, a_ds, b_q, idesc_tt, acc);
                    }
                    for (int kk = 0; kk < nkc; ++kk) {  // reduction over the keys of this tile
                        const uint64_t a_ds = make_smem_desc(ds_base + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024, kSwz128);
                        const uint64_t b_k = make_smem_desc(k_base + kk * 2048, 0, 1024, kSwz128);
                        umma_bf16(tmem_base + TM_DQ + static_cast<uint32_t>(mt * 64), a_ds, b_k, idesc_nt, (kt > 0 || kk > 0) ? 1u : 0u);
                    }
                    if (mt == n_mt - 1) umma_commit(bar_kv);
                }
            }
        }
    } else {
        {
            float dl = 0.f;
            if (has_row) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint4 a = ro[k], g = rg[k];
                    dl += bf16lo(a.x) * bf16lo(g.x) + bf16hi(a.x) * bf16hi(g.x) + bf16lo(a.y) * bf16lo(g.y) + bf16hi(a.y) * bf16hi(g.y) +
                          bf16lo(a.z) * bf16lo(g.z) + bf16hi(a.z) * bf16hi(g.z) + bf16lo(a.w) * bf16lo(g.w) + bf16hi(a.w) * bf16hi(g.w);
                }
            }
            lse_s[irow] = has_row ? ls * kLog2e : INFINITY;  // padded query rows: exp2(s - inf) = 0 -> P = dS = 0 for free
            delta_s[irow] = dl;
            for (int k = irow; k < 24 * 64; k += 256) csum[k] = 0.f;  // (slots of absent tiles stay zero)
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        const uint32_t q = static_cast<uint32_t>(warp & 3);
        const int ch = warp >> 2;                       // ch 0 takes the first half of the key chunks, ch 1 the rest
        const int r = static_cast<int>(q) * 32 + lane;  // row in the query tile == TMEM lane
        const uint32_t taddr = tmem_base + ((q * 32u) << 16);
        const float sl2 = p.scale * kLog2e;
        uint8_t* stage = sP + ch * 16384 + static_cast<int>(q) * 4096;  // this warp's rows of block `ch`: epilogue staging
        const uint32_t rsw = static_cast<uint32_t>(r) & 7u;
        uint32_t it = 0;
        for (int kt = 0; kt < n_kt; ++kt) {
            const int nkc = min(128, p.tp - kt * 128) >> 4;
            const int n0 = (nkc + 1) >> 1;
            const int cbeg = ch ? n0 : 0;
            for (int mt = 0; mt < n_mt; ++mt, ++it) {
                const int nq = min(128, p.tp - mt * 128);
                // warps whose 32 query rows are all padding skip the math: their P / dS rows are outside every reduction
                const int nmine = (static_cast<int>(q) * 32 < nq) ? (ch ? nkc - n0 : n0) : 0;
                const int i = mt * 128 + r;
                const float lse2 = lse_s[i];
                const float delta = delta_s[i];
                uint32_t sv[2][16], dv[2][16];
                mbar_wait(bar_s, it & 1u);
                tc_fence_after_sync();
                if (nmine > 0) tmem_ld_32x32b_x16(taddr + TM_S + static_cast<uint32_t>(cbeg * 16), sv[0]);
                mbar_wait(bar_sdp, it & 1u);
                tc_fence_after_sync();
                if (nmine > 0) tmem_ld_32x32b_x16(taddr + TM_DP + static_cast<uint32_t>(cbeg * 16), dv[0]);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    if (cc < nmine) {  // warp-uniform
                        tmem_ld_wait();
                        if (cc + 1 < nmine) {  // next chunk's TMEM reads fly under this chunk's math
                            tmem_ld_32x32b_x16(taddr + TM_S + static_cast<uint32_t>((cbeg + cc + 1) * 16), sv[(cc + 1) & 1]);
                            tmem_ld_32x32b_x16(taddr + TM_DP + static_cast<uint32_t>((cbeg + cc + 1) * 16), dv[(cc + 1) & 1]);
                        }
                        const uint32_t(&s_)[16] = sv[cc & 1];
                        const uint32_t(&d_)[16] = dv[cc & 1];
                        const int col = (cbeg + cc) * 16;
                        const int j0 = kt * 128 + col;
                        float pv[16], ds[16];
                        if (!p.causal && j0 + 16 <= p.T) {  // warp-uniform fast path: every key of the chunk is valid
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) {
                                pv[jj] = fast_ex2(fmaf(__uint_as_float(s_[jj]), sl2, -lse2));
                                ds[jj] = pv[jj] * (__uint_as_float(d_[jj]) - delta);
                            }
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) {
                                const int j = j0 + jj;
                                const bool ok = (j < p.T) && (!p.causal || j <= i);
                                const float e = fast_ex2(fmaf(__uint_as_float(s_[jj]), sl2, -lse2));
                                pv[jj] = ok ? e : 0.f;
                                ds[jj] = ok ? e * (__uint_as_float(d_[jj]) - delta) : 0.f;
                            }
                        }
                        const uint32_t blk = static_cast<uint32_t>(col >> 6) * 16384u + static_cast<uint32_t>(r) * 128u;
                        const uint32_t c16 = static_cast<uint32_t>((col & 63) >> 3);
                        const uint32_t off0 = blk + ((c16 ^ rsw) << 4);
                        const uint32_t off1 = blk + (((c16 + 1) ^ rsw) << 4);
                        *reinterpret_cast<uint4*>(sP + off0) = make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7]));
                        *reinterpret_cast<uint4*>(sP + off1) = make_uint4(pack_bf16x2(pv[8], pv[9]), pack_bf16x2(pv[10], pv[11]), pack_bf16x2(pv[12], pv[13]), pack_bf16x2(pv[14], pv[15]));
                        *reinterpret_cast<uint4*>(sDS + off0) = make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]), pack_bf16x2(ds[4], ds[5]), pack_bf16x2(ds[6], ds[7]));
                        *reinterpret_cast<uint4*>(sDS + off1) = make_uint4(pack_bf16x2(ds[8], ds[9]), pack_bf16x2(ds[10], ds[11]), pack_bf16x2(ds[12], ds[13]), pack_bf16x2(ds[14], ds[15]));
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_pds);
                if (mt == n_mt - 1) {
                    // dK / dV of this key tile are complete (and, on the last key tile, so are all dQ); every MMA that read
                    // sP / sDS has retired, so sP doubles as the staging area of the coalesced stores
                    mbar_wait(bar_kv, static_cast<uint32_t>(kt) & 1u);
                    tc_fence_after_sync();
                    const int key0 = kt * 128 + static_cast<int>(q) * 32;
                    __nv_bfloat16* g0 = dqkv + (static_cast<long long>(b) * p.T + key0) * D3 + h * 64;
                    float* cs = csum + ((1 + ch) * 8 + kt * 4 + static_cast<int>(q)) * 64;
                    if (ch == 0) store_rows_coalesced<4, true>(stage, g0 + p.D, D3, taddr + TM_DK, p.scale, p.T - key0, lane, cs);
                    else         store_rows_coalesced<4, true>(stage, g0 + 2 * p.D, D3, taddr + TM_DV, 1.0f, p.T - key0, lane, cs);
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_kvfree);
                    // the next key tile's P / dS chunks of another warp may land in this warp's staging rows
                    if (kt + 1 < n_kt) asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
        }
        // dQ tiles: warps with ch < n_mt each write query tile `ch`
        if (ch < n_mt) {
            const int row0 = ch * 128 + static_cast<int>(q) * 32;
            __nv_bfloat16* g0 = dqkv + (static_cast<long long>(b) * p.T + row0) * D3 + h * 64;
            store_rows_coalesced<4, true>(stage, g0, D3, taddr + TM_DQ + static_cast<uint32_t>(ch * 64), p.scale, p.T - row0, lane,
                                          csum + (ch * 4 + static_cast<int>(q)) * 64);
        }
        if (dbias_part != nullptr) {
            // per-(batch, head) column sums of dq | dk | dv in a fixed order: the qkv-bias gradient is their sum over the batch
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (irow < 192) {
                const int sec = irow >> 6, c = irow & 63;
                float t = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) t += csum[(sec * 8 + k) * 64 + c];
                dbias_part[static_cast<long long>(b) * D3 + sec * p.D + h * 64 + c] = t;
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// backward, version 2: persistent, transposed scores, P^T / dS^T operands in TENSOR MEMORY
// ------------------------------------------------------------------------------------------------
// Version 1 above spends ~33 k cycles per (batch, head) where the tensor + MUFU work is ~7 k: one CTA per item pays
// launch + TMEM allocation + 128 KB of un-overlapped TMA loads, and inside the item S / dP -> softmax -> dV / dK / dQ is
// one serial chain (all 512 TMEM columns taken, P and dS staged through shared memory).  Version 2:
//   * ONE persistent CTA per SM walks over the items; every operand buffer (K|V per key tile, Q|dO per query tile) has
//     its own full / empty barrier pair, so the next item's tiles stream in as soon as the current item has issued its
//     last MMA on that buffer (K_0 / V_0 are free after the first key tile, ...): no load is ever exposed;
//   * scores are computed TRANSPOSED, S^T = K Q^T and dP^T = V dO^T (TMEM lane = key), in sub-tiles of 64 queries:
//     a sub-tile's S^T | dP^T take 128 columns, so TWO of them fit beside the four 64-column accumulators
//     (dK, dV, dQ_0, dQ_1 = 256 columns) and the tensor pipe works on sub-tile u+1 / u+2 while the compute warps are on u;
//   * the compute warps (lane = key row, two threads per row) write P^T and dS^T back into their OWN S^T / dP^T columns as
//     packed bf16 (tcgen05.st); dV += P^T dO and dK += dS^T Q read their A operand from tensor memory
//     (tcgen05.mma [d], [a_tmem], b_desc), so P never touches shared memory and dS^T goes there only for dQ += dS K,
//     where it is consumed MN-major (M = query) straight from the [key][query] tile the threads wrote;
//   * 4 epilogue warps drain dK / dV after each key tile and dQ after each item (coalesced stores, column sums for the
//     qkv-bias gradient) and, in their idle time, compute delta_i = sum_d dO[i,d] O[i,d] for the NEXT item.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..9 = compute, 10..13 = epilogue.
constexpr int B2_THREADS = 448;
constexpr int B2_SQ = 0;           // 2 x 16 KB   query tiles
constexpr int B2_SDO = 32768;      // 2 x 16 KB
constexpr int B2_SK = 65536;       // 2 x 16 KB   key tiles
constexpr int B2_SV = 98304;       // 2 x 16 KB
constexpr int B2_SDS = 131072;     // 2 pair buffers x 2 blocks x 16 KB: dS^T [128 keys][64 queries] bf16, 128 B rows, swizzled
constexpr int B2_STG = 196608;     // 4 x 4 KB    epilogue staging
constexpr int B2_LSE = B2_STG + 16384;          // [2 items][256] floats (lse * log2 e; +inf for padded queries)
constexpr int B2_DELTA = B2_LSE + 2048;         // [2 items][256]
constexpr int B2_CSUM = B2_DELTA + 2048;        // 24 x 64 floats: column sums of dQ / dK / dV per (tile, lane quadrant)
constexpr int B2_BAR = B2_CSUM + 24 * 64 * 4;
constexpr int B2_SMEM = B2_BAR + 256 + 1024;
constexpr uint32_t T2_BUF = 128;   // per S^T|dP^T buffer: S^T at +0, dP^T at +64
constexpr uint32_t T2_DK = 256, T2_DV = 320, T2_DQ = 384;

// staged rows -> global as in flush_rows64, plus the fp32 column sums of the bf16 values written (qkv-bias gradient)
__device__ __forceinline__ void flush_rows64_colsum(const uint8_t* stage, __nv_bfloat16* g0, long long row_stride, int rows_valid, int lane,
                                                    float* csum) {
    __syncwarp();
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3);
        const uint32_t ch = static_cast<uint32_t>(lane) & 7u;
        const uint4 val = *reinterpret_cast<const uint4*>(stage + row * 128 + ((ch ^ (static_cast<uint32_t>(row) & 7u)) << 4));
        if (row < rows_valid) {
            *reinterpret_cast<uint4*>(g0 + row * row_stride + ch * 8) = val;
            acc[0] += bf16lo(val.x); acc[1] += bf16hi(val.x); acc[2] += bf16lo(val.y); acc[3] += bf16hi(val.y);
            acc[4] += bf16lo(val.z); acc[5] += bf16hi(val.z); acc[6] += bf16lo(val.w); acc[7] += bf16hi(val.w);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // lanes l, l^8, l^16, l^24 hold the same 8 columns of different rows
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    }
    if (lane < 8) {
        reinterpret_cast<float4*>(csum + lane * 8)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        reinterpret_cast<float4*>(csum + lane * 8)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncwarp();
}

// one 16-query chunk of a sub-tile for one key row: S^T / dP^T (fp32, tensor memory) -> P^T / dS^T (packed bf16, written back
// over the chunk's own columns) and dS^T -> shared memory.  nls / ndl hold -lse * log2(e) and -delta of the item's queries.
// The arithmetic runs two queries at a time on the packed fp32x2 pipe (fma.rn.f32x2 / mul.rn.f32x2).
template <bool MASK>
__device__ __forceinline__ void bwd2_chunk(uint32_t taddr, int c, const float* nls, const float* ndl, int i0, float sl2,
                                           bool keyok, bool causal, int j, uint8_t* blk, uint32_t rsw) {
    uint32_t sv[16], dv[16];
    // this thread's columns: chunk c of the sub-tile lives at S^T / dP^T column c * 16
    tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(c * 16), sv);
    tmem_ld_32x32b_x16(taddr + 64 + static_cast<uint32_t>(c * 16), dv);
    float4 la[4], da[4];
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
        la[k4] = *reinterpret_cast<const float4*>(nls + i0 + k4 * 4);
        da[k4] = *reinterpret_cast<const float4*>(ndl + i0 + k4 * 4);
    }
    tmem_ld_wait();
    const f32x2_t sl22 = pk2(sl2, sl2), one2 = pk2(1.0f, 1.0f);
    uint32_t pk[8], dk[8];
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int jj = k4 * 4 + h2 * 2;
            const f32x2_t l2 = h2 == 0 ? pk2(la[k4].x, la[k4].y) : pk2(la[k4].z, la[k4].w);
            const f32x2_t d2 = h2 == 0 ? pk2(da[k4].x, da[k4].y) : pk2(da[k4].z, da[k4].w);
            float x0, x1;
            upk2(fma2(pk2(__uint_as_float(sv[jj]), __uint_as_float(sv[jj + 1])), sl22, l2), x0, x1);
            float e0 = fast_ex2(x0), e1 = fast_ex2(x1);
            if (MASK) {
                if (!(keyok && (!causal || j <= i0 + jj))) e0 = 0.f;
                if (!(keyok && (!causal || j <= i0 + jj + 1))) e1 = 0.f;
            }
            const f32x2_t e2 = pk2(e0, e1);
            float g0, g1;
            upk2(mul2(e2, fma2(pk2(__uint_as_float(dv[jj]), __uint_as_float(dv[jj + 1])), one2, d2)), g0, g1);  // e (dP - delta)
            pk[jj >> 1] = pack_bf16x2(e0, e1);
            dk[jj >> 1] = pack_bf16x2(g0, g1);
        }
    }
    // packed values over the first 8 of the chunk's OWN 16 columns: no other thread's unread S^T / dP^T is touched
    const uint32_t ocol = static_cast<uint32_t>(c * 16);
    tmem_st_32x32b_x8(taddr + ocol, pk);
    tmem_st_32x32b_x8(taddr + 64 + ocol, dk);
    // dS^T[key r][queries i0 .. i0+16) -> 32 bytes of the [key][query] tile (swizzled 16-byte chunks)
    const uint32_t c16 = static_cast<uint32_t>(c * 2);
    *reinterpret_cast<uint4*>(blk + ((c16 ^ rsw) << 4)) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
    *reinterpret_cast<uint4*>(blk + (((c16 + 1) ^ rsw) << 4)) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
}

// walks the (item, key tile, query sub-tile) units of this CTA without divisions (the MMA thread issues ~30 instructions per
// unit on a single lane: every integer division on that path showed up as tensor-pipe idle time)
struct B2Cur {
    int g, it, kt, qs;
    __device__ __forceinline__ void advance(int n_kt, int n_qs) {
        ++g;
        if (++qs == n_qs) {
            qs = 0;
            if (++kt == n_kt) { kt = 0; ++it; }
        }
    }
};

// PP: the eight compute warps work as TWO groups of four (one warp per scheduler each); group 0 owns the sub-tiles in buffer 0,
// group 1 those in buffer 1.  While one group waits for its next S^T | dP^T (its P^T / dS^T must first be consumed by dV / dK,
// then the tensor pipe produces the new scores: ~1500 cycles), the other group has the issue slots to itself.  !PP keeps the
// first cut's schedule (all eight warps on one sub-tile, two threads per key row) for A/B timing.
// (14 warps put 4 on one scheduler: 16 384 registers / 4 warps = 128 per thread is the hardware ceiling for this block size --
// __maxnreg__(144) compiles and then fails to launch)
template <bool PP>
__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                 const __nv_bfloat16* __restrict__ o_in, const __nv_bfloat16* __restrict__ do_in,
                 const float* __restrict__ lse_in, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias_part,
                 const AttnParams p, const int n_items) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* csum = reinterpret_cast<float*>(smem + B2_CSUM);
    uint64_t* full_q = reinterpret_cast<uint64_t*>(smem + B2_BAR);  // [2] Q_mt | dO_mt loaded
    uint64_t* empty_q = full_q + 2;                                 // [2] ... no longer read by any MMA
    uint64_t* full_kv = full_q + 4;                                 // [2] K_kt | V_kt loaded
    uint64_t* empty_kv = full_q + 6;
    uint64_t* bar_sdp = full_q + 8;                                 // [2] S^T | dP^T of the buffer's sub-tile are in TMEM
    uint64_t* bar_pds = full_q + 10;                                // [2] P^T | dS^T written (8 compute warps)
    uint64_t* bar_dsfree = full_q + 12;                             // [2] dS^T pair buffer consumed by its dQ MMAs
    uint64_t* bar_kv = full_q + 14;                                 // dK | dV of the key tile complete
    uint64_t* bar_kvfree = full_q + 15;                             // ... and read out (4 epilogue warps)
    uint64_t* bar_dq = full_q + 16;                                 // dQ of the item complete
    uint64_t* bar_dqfree = full_q + 17;
    uint64_t* bar_delta = full_q + 18;                              // [2] lse / delta of the item (parity buffer) in smem (4 epilogue warps)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(full_q + 20);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_kt = (p.T + 127) / 128;          // key tiles == query tiles (pairs of 64-query sub-tiles)
    const int n_qs = (p.tp + 63) / 64;           // 64-query sub-tiles
    const int U = n_kt * n_qs;                   // sub-iterations per item
    const int n_my = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const long long D3 = 3ll * p.D;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full_q[i], 1);
            mbar_init(&empty_q[i], 1);
            mbar_init(&full_kv[i], 1);
            mbar_init(&empty_kv[i], 1);
            mbar_init(&bar_sdp[i], 1);
            mbar_init(&bar_pds[i], PP ? 4 : 8);
            mbar_init(&bar_dsfree[i], 1);
            mbar_init(&bar_delta[i], 4);
        }
        mbar_init(bar_kv, 1);
        mbar_init(bar_kvfree, 4);
        mbar_init(bar_dq, 1);
        mbar_init(bar_dqfree, 4);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            for (int it = 0; it < n_my; ++it) {
                const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
                const int h = w % p.H, b = w / p.H;
                const uint32_t par = static_cast<uint32_t>(it & 1);
                if (p.prefetch && it + 1 < n_my) {  // this CTA's share of the NEXT item's image: qkv, O and dO as whole contiguous rows
                    const int w1 = w + static_cast<int>(gridDim.x);
                    const int b1 = w1 / p.H, h1 = w1 % p.H;
                    prefetch_share(p.qkv_base, static_cast<long long>(p.T) * 3 * p.D * 2, b1, h1, p.H);
                    prefetch_share(p.o_base, static_cast<long long>(p.T) * p.D * 2, b1, h1, p.H);
                    prefetch_share(p.do_base, static_cast<long long>(p.T) * p.D * 2, b1, h1, p.H);
                }
                for (int t = 0; t < n_kt; ++t) {  // (buffers are released in the order kv0, q0, kv1, q1: wait in that order)
                    mbar_wait(&empty_kv[t], par ^ 1u);
                    mbar_expect_tx(&full_kv[t], 2u * 16384u);
                    tma_load_3d(smem + B2_SK + t * 16384, &tmQKV, &full_kv[t], p.D + h * 64, t * 128, b);
                    tma_load_3d(smem + B2_SV + t * 16384, &tmQKV, &full_kv[t], 2 * p.D + h * 64, t * 128, b);
                    mbar_wait(&empty_q[t], par ^ 1u);
                    mbar_expect_tx(&full_q[t], 2u * 16384u);
                    tma_load_3d(smem + B2_SQ + t * 16384, &tmQKV, &full_q[t], h * 64, t * 128, b);
                    tma_load_3d(smem + B2_SDO + t * 16384, &tmDO, &full_q[t], h * 64, t * 128, b);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        if (elect_one()) {
            const uint32_t idesc_ts = make_idesc_bf16(128, 64, 0, 1);   // dV, dK : A from TMEM (K-major), B MN-major
            const uint32_t idesc_dq = make_idesc_bf16(128, 64, 1, 1);   // dQ     : A MN-major (smem), B MN-major
            const int G = n_my * U;
            // descriptor bases: a 128-byte-row swizzled tile is read K-major (S^T, dP^T) and MN-major (dV, dK, dQ) with the same fields
            const uint64_t dsc_k = make_smem_desc(smem_u32(smem + B2_SK), 0, 1024, kSwz128);
            const uint64_t dsc_v = make_smem_desc(smem_u32(smem + B2_SV), 0, 1024, kSwz128);
            const uint64_t dsc_q = make_smem_desc(smem_u32(smem + B2_SQ), 0, 1024, kSwz128);
            const uint64_t dsc_do = make_smem_desc(smem_u32(smem + B2_SDO), 0, 1024, kSwz128);
            const uint64_t dsc_ds = make_smem_desc(smem_u32(smem + B2_SDS), 16384, 1024, kSwz128);
            auto issue_sdp = [&](const B2Cur& c) {
                const uint32_t par = static_cast<uint32_t>(c.it & 1);
                if (c.qs == 0) mbar_wait(&full_kv[c.kt], par);
                if (c.kt == 0 && (c.qs & 1) == 0) mbar_wait(&full_q[c.qs >> 1], par);
                tc_fence_after_sync();
                const int nq = min(64, p.tp - c.qs * 64);
                const uint32_t idesc_nn = make_idesc_bf16(128, static_cast<uint32_t>(nq), 0, 0);
                const uint32_t kv_off = static_cast<uint32_t>(c.kt) * 1024u;   // 16 KB per key tile
                const uint32_t q_off = static_cast<uint32_t>(c.qs) * 512u;     // 8 KB per 64-query sub-tile (two per query tile)
                const uint32_t d = tmem_base + static_cast<uint32_t>(c.g & 1) * T2_BUF;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(d, desc_off(dsc_k, kv_off + kk * 2), desc_off(dsc_q, q_off + kk * 2), idesc_nn, kk > 0 ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(d + 64, desc_off(dsc_v, kv_off + kk * 2), desc_off(dsc_do, q_off + kk * 2), idesc_nn, kk > 0 ? 1u : 0u);
                umma_commit(&bar_sdp[c.g & 1]);
            };
            // S^T / dP^T run up to two sub-iterations ahead of the compute warps.  Across an item boundary that is only possible
            // when the next item's first tiles can already be resident, i.e. when the current item released K_0|V_0 and
            // Q_0|dO_0 at least two sub-iterations before its end (two key tiles and >= 3 query sub-tiles, e.g. T = 197);
            // otherwise the look-ahead stops at the boundary -- waiting there for loads that need a LATER commit of this
            // thread would deadlock.
            const bool cross = (n_kt == 2 && n_qs >= 3);
            B2Cur nx{0, 0, 0, 0};
            auto pump = [&](int g_cur, int cur_item, bool item_finished) {
                while (nx.g < G && nx.g <= g_cur + 2) {
                    if (!cross && nx.it > cur_item && !item_finished) break;
                    if (!cross && nx.it > cur_item + 1) break;
                    issue_sdp(nx);
                    nx.advance(n_kt, n_qs);
                }
            };
            pump(-1, 0, false);
            int pc = 0;   // dS^T pairs completed so far (pair buffer = pc & 1)
            int kc = 0;   // key tiles completed so far
            for (B2Cur c{0, 0, 0, 0}; c.g < G; c.advance(n_kt, n_qs)) {
                const int g = c.g, it = c.it, kt = c.kt, qs = c.qs, mt = qs >> 1;
                const int nqc = min(64, p.tp - qs * 64) >> 4;   // 16-query steps of the sub-tile
                const int nkc = min(128, p.tp - kt * 128) >> 4; // 16-key steps of the key tile
                const uint32_t buf = tmem_base + static_cast<uint32_t>(g & 1) * T2_BUF;
                mbar_wait(&bar_pds[g & 1], static_cast<uint32_t>((g >> 1) & 1));
                if (qs == 0 && kc > 0) mbar_wait(bar_kvfree, static_cast<uint32_t>((kc - 1) & 1));  // dK / dV of the previous key tile read out
                tc_fence_after_sync();
                const uint32_t q_off = static_cast<uint32_t>(qs) * 512u;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {  // reduction over the queries of the sub-tile, 16 per step
                    if (kk < nqc) {
                        // every 16-query chunk was written back, packed to 8 columns, at the START of its own 16 fp32 columns
                        const uint32_t acol = static_cast<uint32_t>(kk * 16);
                        const uint32_t acc = (qs > 0 || kk > 0) ? 1u : 0u;
                        umma_bf16_ts(tmem_base + T2_DV, buf + acol, desc_off(dsc_do, q_off + kk * 128), idesc_ts, acc);
                        umma_bf16_ts(tmem_base + T2_DK, buf + 64 + acol, desc_off(dsc_q, q_off + kk * 128), idesc_ts, acc);
                    }
                }
                const bool pair_done = (qs & 1) == 1 || qs == n_qs - 1;
                if (pair_done) {
                    if (kt == 0 && mt == 0 && it > 0) mbar_wait(bar_dqfree, static_cast<uint32_t>((it - 1) & 1));  // previous item's dQ read out
                    tc_fence_after_sync();
                    const uint32_t ds_off = static_cast<uint32_t>(pc & 1) * 2048u;  // 32 KB per pair buffer
                    const uint32_t k_off = static_cast<uint32_t>(kt) * 1024u;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)  // reduction over the keys of the tile
                        if (kk < nkc)
                            umma_bf16(tmem_base + T2_DQ + static_cast<uint32_t>(mt * 64), desc_off(dsc_ds, ds_off + kk * 128),
                                      desc_off(dsc_k, k_off + kk * 128), idesc_dq, (kt > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&bar_dsfree[pc & 1]);
                    ++pc;
                    if (kt == n_kt - 1) umma_commit(&empty_q[mt]);  // last MMAs on Q_mt / dO_mt of this item
                }
                if (qs == n_qs - 1) {
                    umma_commit(bar_kv);
                    umma_commit(&empty_kv[kt]);
                    ++kc;
                    if (kt == n_kt - 1) umma_commit(bar_dq);
                }
                pump(g, it, kt == n_kt - 1 && qs == n_qs - 1);
            }
        }
    } else if (warp >= 10) {
        // ===================================== epilogue warps =====================================
        const uint32_t q = static_cast<uint32_t>(warp & 3);
        const int et = (warp - 10) * 32 + lane;  // 0..127
        uint8_t* stage = smem + B2_STG + (warp - 10) * 4096;
        auto prepare = [&](int it) {  // lse (scaled) and delta of item `it` -> its parity buffers
            const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int h = w % p.H, b = w / p.H;
            float* lse_s = reinterpret_cast<float*>(smem + B2_LSE) + (it & 1) * 256;
            float* delta_s = reinterpret_cast<float*>(smem + B2_DELTA) + (it & 1) * 256;
            for (int i = et; i < 256; i += 128) {
                float dl = 0.f, ls = INFINITY;  // padded query columns: exp2(s - inf) = 0 -> P = dS = 0 for free (stored negated)
                if (i < p.T) {
                    const uint4* po = reinterpret_cast<const uint4*>(o_in + (static_cast<long long>(b) * p.T + i) * p.D + h * 64);
                    const uint4* pd = reinterpret_cast<const uint4*>(do_in + (static_cast<long long>(b) * p.T + i) * p.D + h * 64);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint4 a = __ldg(po + k), gg = __ldg(pd + k);
                        dl += bf16lo(a.x) * bf16lo(gg.x) + bf16hi(a.x) * bf16hi(gg.x) + bf16lo(a.y) * bf16lo(gg.y) + bf16hi(a.y) * bf16hi(gg.y) +
                              bf16lo(a.z) * bf16lo(gg.z) + bf16hi(a.z) * bf16hi(gg.z) + bf16lo(a.w) * bf16lo(gg.w) + bf16hi(a.w) * bf16hi(gg.w);
                    }
                    ls = __ldg(lse_in + (static_cast<long long>(b) * p.H + h) * p.T + i) * kLog2e;
                }
                lse_s[i] = -ls;
                delta_s[i] = -dl;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_delta[it & 1]);
        };
        if (n_my > 0) prepare(0);
        int kc = 0;
        for (int it = 0; it < n_my; ++it) {
            const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int h = w % p.H, b = w / p.H;
            for (int k = et; k < 24 * 64; k += 128) csum[k] = 0.f;  // (slots of absent tiles stay zero)
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (it + 1 < n_my) prepare(it + 1);  // the other parity buffer: its previous user (item it-1) has been read out completely
            for (int kt = 0; kt < n_kt; ++kt, ++kc) {
                mbar_wait(bar_kv, static_cast<uint32_t>(kc & 1));
                tc_fence_after_sync();
                const int key0 = kt * 128 + static_cast<int>(q) * 32;
                const uint32_t taddr = tmem_base + ((q * 32u) << 16);
                const bool live = key0 < p.T;
                __nv_bfloat16* g0 = dqkv + (static_cast<long long>(b) * p.T + key0) * D3 + h * 64;
                if (live) {
                    stage_rows64(stage, taddr + T2_DK, p.scale, lane);
                    flush_rows64_colsum(stage, g0 + p.D, D3, p.T - key0, lane, csum + (8 + kt * 4 + static_cast<int>(q)) * 64);
                    stage_rows64(stage, taddr + T2_DV, 1.0f, lane);
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_kvfree);
                if (live) flush_rows64_colsum(stage, g0 + 2 * p.D, D3, p.T - key0, lane, csum + (16 + kt * 4 + static_cast<int>(q)) * 64);
            }
            mbar_wait(bar_dq, static_cast<uint32_t>(it & 1));
            tc_fence_after_sync();
            for (int mt = 0; mt < n_kt; ++mt) {
                const int row0 = mt * 128 + static_cast<int>(q) * 32;
                if (row0 < p.T) {
                    stage_rows64(stage, tmem_base + ((q * 32u) << 16) + T2_DQ + static_cast<uint32_t>(mt * 64), p.scale, lane);
                    if (mt == n_kt - 1) {  // everything of this item has left tensor memory
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_dqfree);
                    }
                    flush_rows64_colsum(stage, dqkv + (static_cast<long long>(b) * p.T + row0) * D3 + h * 64, D3, p.T - row0, lane,
                                        csum + (mt * 4 + static_cast<int>(q)) * 64);
                } else if (mt == n_kt - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_dqfree);
                }
            }
            if (dbias_part != nullptr) {
                // per-(batch, head) column sums of dq | dk | dv in a fixed order: the qkv-bias gradient is their sum over the batch
                asm volatile("bar.sync 2, 128;" ::: "memory");
                for (int idx = et; idx < 192; idx += 128) {
                    const int sec = idx >> 6, c = idx & 63;
                    float t = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) t += csum[(sec * 8 + k) * 64 + c];
                    dbias_part[static_cast<long long>(b) * D3 + sec * p.D + h * 64 + c] = t;
                }
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");  // csum is zeroed again at the top
        }
    } else {
        // ===================================== compute warps ======================================
        const uint32_t q = static_cast<uint32_t>(warp & 3);
        const int grp = (warp - 2) >> 2;                 // PP: which buffer's sub-tiles; !PP: which half of the query columns
        const int r = static_cast<int>(q) * 32 + lane;   // key row inside the key tile == TMEM lane
        const uint32_t rsw = static_cast<uint32_t>(r) & 7u;
        const float sl2 = p.scale * kLog2e;
        const int G = n_my * U;
        const int ppk = (n_qs + 1) >> 1;                 // dS^T pairs per key tile
        B2Cur c{0, 0, 0, 0};
        if (PP && grp == 1) c.advance(n_kt, n_qs);
        while (c.g < G) {
            const int g = c.g, it = c.it, kt = c.kt, qs = c.qs;
            const int pc = (it * n_kt + kt) * ppk + (qs >> 1);   // running index of this sub-tile's dS^T pair
            const int nq = min(64, p.tp - qs * 64);
            const int nk = min(128, p.tp - kt * 128);
            const int nch = nq >> 4;                             // 16-column chunks in this sub-tile (1..4)
            const int cb = PP ? 0 : (grp ? (nch + 1) >> 1 : 0);  // this thread's chunks [cb, ce)
            const int ce = PP ? nch : (grp ? nch : (nch + 1) >> 1);
            const bool dead = static_cast<int>(q) * 32 >= nk;    // all 32 key rows of this warp lie outside the (padded) key tile
            const int j = kt * 128 + r;                          // key index
            const float* nls = reinterpret_cast<const float*>(smem + B2_LSE) + (it & 1) * 256;
            const float* ndl = reinterpret_cast<const float*>(smem + B2_DELTA) + (it & 1) * 256;
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + static_cast<uint32_t>(g & 1) * T2_BUF;
            uint8_t* blk = smem + B2_SDS + (pc & 1) * 32768 + (qs & 1) * 16384 + r * 128;
            if (kt == 0 && qs < (PP ? 2 : 1)) mbar_wait(&bar_delta[it & 1], static_cast<uint32_t>((it >> 1) & 1));
            if (pc >= 2) mbar_wait(&bar_dsfree[pc & 1], static_cast<uint32_t>(((pc >> 1) - 1) & 1));  // pair buffer consumed by its dQ MMAs
            mbar_wait(&bar_sdp[g & 1], static_cast<uint32_t>((g >> 1) & 1));
            tc_fence_after_sync();
            if (!dead) {
                // warp-uniform: every key row of this warp is a real key and there is no causal mask -> no per-element selects
                const bool nomask = !p.causal && (kt * 128 + static_cast<int>(q) * 32 + 31 < p.T);
#pragma unroll
                for (int cc = 0; cc < (PP ? 4 : 2); ++cc) {
                    const int ch = cb + cc;
                    if (ch < ce) {  // warp-uniform
                        const int i0 = qs * 64 + ch * 16;  // first query of the chunk
                        if (nomask) bwd2_chunk<false>(taddr, ch, nls, ndl, i0, sl2, true, false, j, blk, rsw);
                        else bwd2_chunk<true>(taddr, ch, nls, ndl, i0, sl2, j < p.T, p.causal != 0, j, blk, rsw);
                    }
                }
                tmem_st_wait();
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_pds[g & 1]);
            c.advance(n_kt, n_qs);
            if (PP) c.advance(n_kt, n_qs);
        }
    }
    __syncwarp();
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200

using namespace b200;

// 0 = choose per shape (measured on a B200, profiles/r02_op_bench_attention.txt): the persistent version-2 kernels win when
// an item has two query tiles (T > 128: forward 113 us vs 133 us, backward 338 us vs 357 us at B256 T197 H12); with a single,
// mostly padded tile (CLIP: T = 77 / 50) the per-item pipeline hand-offs cost more than version 1's 3-CTAs-per-SM occupancy
// hides (41.6 us vs 34.5 us, 52.2 us vs 42.0 us).  B200_ATTN_FWD / B200_ATTN_BWD = 1 | 2 or b200_set_attention_*_version
// force one kernel (A/B timing, tests of both).
static int g_attn_fwd_version = 0;
static int g_attn_bwd_version = 0;
static inline int attn_pick(int forced, int T) { return forced != 0 ? forced : (T > 128 ? 2 : 1); }
// (backward: 3 = all eight compute warps on one sub-tile measured 276 us against 290 us for the ping-pong groups of 2)
static inline int attn_pick_bwd(int forced, int T) { return forced != 0 ? forced : (T > 128 ? 3 : 1); }
static int g_attn_prefetch = 0;  // measured neutral (tools/probe_layout.py: the kernels are not DRAM-pattern bound), kept as a switch

static int attn_check(int B, int T, int H, int Dh) {
    if (B <= 0 || T <= 0 || H <= 0) return set_error(B200_ERR_ARG, "attention: non-positive size");
    if (Dh != 64) return set_error(B200_ERR_ARG, "attention: head dim must be 64");
    if (T > 256) return set_error(B200_ERR_ARG, "attention: T must be <= 256");
    return 0;
}

extern "C" int b200_attention_fwd(const void* qkv_bf16, void* out_bf16, float* lse, int B, int T, int H, int Dh,
                                  float scale, int causal, cudaStream_t stream) {
    int rc = attn_check(B, T, H, Dh);
    if (rc) return rc;
    const int D = H * 64;
    const int tp = (T + 15) / 16 * 16;
    CUtensorMap tmQ, tmKV;
    uint64_t dims[3] = {static_cast<uint64_t>(3 * D), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(3 * D) * 2, static_cast<uint64_t>(3 * D) * 2 * static_cast<uint64_t>(T)};
    uint32_t boxq[3] = {64, 128, 1};
    uint32_t boxkv[3] = {64, static_cast<uint32_t>(tp), 1};
    if ((rc = make_tmap(&tmQ, qkv_bf16, 2, 3, dims, strides, boxq, 128)) != 0) return rc;
    if ((rc = make_tmap(&tmKV, qkv_bf16, 2, 3, dims, strides, boxkv, 128)) != 0) return rc;
    static bool configured[64] = {};  // per device: the attribute belongs to the (function, device) pair
    const int dev = current_device_slot();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM);
        if (e != cudaSuccess) return set_error(B200_ERR_LAUNCH, cudaGetErrorString(e));
        configured[dev] = true;
    }
    AttnParams p;
    p.B = B; p.T = T; p.H = H; p.D = D; p.tp = tp; p.scale = scale; p.causal = causal;
    p.prefetch = g_attn_prefetch; p.qkv_base = qkv_bf16; p.o_base = nullptr; p.do_base = nullptr;
    const int n_mt = (T + 127) / 128;
    if (attn_pick(g_attn_fwd_version, T) == 1) {
        attn_fwd_kernel<<<B * H * n_mt, AF_THREADS, AF_SMEM, stream>>>(tmQ, tmKV, reinterpret_cast<__nv_bfloat16*>(out_bf16), lse, p);
    } else {
        const int n_items = B * H;
        const int grid = n_items < persistent_ctas() ? n_items : persistent_ctas();
        attn_fwd2_kernel<<<grid, F2_THREADS, F2_SMEM, stream>>>(tmQ, tmKV, reinterpret_cast<__nv_bfloat16*>(out_bf16), lse, p, n_items);
    }
    return check_launch("attention_fwd");
}

// 0 (default): per shape; 2: persistent pipelined forward with P in tensor memory; 1: the round-1 kernel (same results up to
// the summation order of the row sums).  Returns the previous setting.
extern "C" int b200_set_attention_fwd_version(int version) {
    const int old = g_attn_fwd_version;
    g_attn_fwd_version = (version == 1 || version == 2) ? version : 0;
    return old;
}

extern "C" int b200_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse,
                                  void* dqkv_bf16, int B, int T, int H, int Dh, float scale, int causal,
                                  float* dbias_part, cudaStream_t stream) {
    int rc = attn_check(B, T, H, Dh);
    if (rc) return rc;
    const int D = H * 64;
    CUtensorMap tmQKV, tmDO;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(3 * D), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
        uint64_t strides[2] = {static_cast<uint64_t>(3 * D) * 2, static_cast<uint64_t>(3 * D) * 2 * static_cast<uint64_t>(T)};
        uint32_t box[3] = {64, 128, 1};
        if ((rc = make_tmap(&tmQKV, qkv_bf16, 2, 3, dims, strides, box, 128)) != 0) return rc;
    }
    {
        uint64_t dims[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
        uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, static_cast<uint64_t>(D) * 2 * static_cast<uint64_t>(T)};
        uint32_t box[3] = {64, 128, 1};
        if ((rc = make_tmap(&tmDO, dout_bf16, 2, 3, dims, strides, box, 128)) != 0) return rc;
    }
    static bool configured[64] = {};
    const int dev = current_device_slot();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
        if (e != cudaSuccess) return set_error(B200_ERR_LAUNCH, cudaGetErrorString(e));
        configured[dev] = true;
    }
    AttnParams p;
    p.B = B; p.T = T; p.H = H; p.D = D; p.tp = (T + 15) / 16 * 16; p.scale = scale; p.causal = causal;
    p.prefetch = g_attn_prefetch; p.qkv_base = qkv_bf16; p.o_base = out_bf16; p.do_base = dout_bf16;
    const int ver = attn_pick_bwd(g_attn_bwd_version, T);
    if (ver == 1) {
        attn_bwd_kernel<<<B * H, AB_THREADS, AB_SMEM, stream>>>(tmQKV, tmDO, reinterpret_cast<const __nv_bfloat16*>(out_bf16),
                                                                reinterpret_cast<const __nv_bfloat16*>(dout_bf16), lse,
                                                                reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), dbias_part, p);
    } else {
        const int n_items = B * H;
        const int grid = n_items < persistent_ctas() ? n_items : persistent_ctas();
        auto kern = ver == 3 ? attn_bwd2_kernel<false> : attn_bwd2_kernel<true>;
        kern<<<grid, B2_THREADS, B2_SMEM, stream>>>(tmQKV, tmDO, reinterpret_cast<const __nv_bfloat16*>(out_bf16),
                                                    reinterpret_cast<const __nv_bfloat16*>(dout_bf16), lse,
                                                    reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), dbias_part, p, n_items);
    }
    return check_launch("attention_bwd");
}

extern "C" int b200_set_attention_prefetch(int enable) {
    const int old = g_attn_prefetch;
    g_attn_prefetch = enable ? 1 : 0;
    return old;
}

// 0 (default): per shape; 2: persistent backward with transposed scores and P^T / dS^T operands in tensor memory, compute warps
// in two ping-pong groups; 3: the same kernel with all eight compute warps on one sub-tile (A/B timing); 1: the round-1 kernel.
// Returns the previous setting.
extern "C" int b200_set_attention_bwd_version(int version) {
    const int old = g_attn_bwd_version;
    g_attn_bwd_version = (version >= 1 && version <= 3) ? version : 0;
    return old;
}
