// attn_tc.cu -- prefix (image / text rows) attention of the Gemma mixture on the 5th-generation tensor cores.
//
// Reference semantics: forward_mixture_attn, joint_model.py:130-304, for the vlm rows of the prefix pass:
// scores = q k^T / sqrt(256) -> soft-cap 50 (tanh) -> block mask (a vlm row sees the keys j < cnt of its own
// sample, pizero.py:296-300) -> softmax -> P V; multi-query (8 query heads, one K/V head).
//
// Mapping (one CTA per (sample, 16-token tile); 1 CTA per SM):
//   * MQA folds the heads into the MMA M dimension: a tile is 16 tokens x 8 heads = 128 rows that all share the
//     same keys AND the same mask (the mask of a vlm row depends on the sample only).  A 3-D TMA box
//     [64 d][8 heads][16 tokens] lands those rows as a K-major, 128B-swizzled UMMA operand; the output leaves
//     through the same box shape.
//   * keys in chunks of 128: S = Q K_c^T (tcgen05.mma, 128 x nk x 256, fp32 in TMEM, two S buffers) ->
//     8 softmax warps (thread = row, half of the columns each) tcgen05.ld S, soft-cap, exp, mask, write P (bf16)
//     as a K-major swizzled operand into shared memory -> O += P V_c (tcgen05.mma 128 x 256 x nk, V consumed in
//     its natural [key][d] layout as an MN-major operand) -> epilogue: O / rowsum -> bf16 -> TMA store.
//   * no running maximum: the soft-cap bounds every logit to +-50, so exp() cannot overflow fp32 / bf16 and
//     the unnormalised P and its fp32 row sum need no rescaling between key chunks.
//   * TMEM: 2 x 128 columns of S + 256 columns of O = 512.  Shared memory: Q 64 KB (reused as output staging),
//     K chunk 64 KB, V chunk 64 KB, P 32 KB.
//   warp 0: TMA producer, warp 1: MMA issuer, warp 2: TMEM allocator, warps 4-11: softmax + epilogue.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int HD = 256, TOK = 16, NH = 8, ROWS = TOK * NH, KC = 128;
constexpr int THREADS = 12 * 32;
constexpr int SM_WARP0 = 4, SM_THREADS = 256;
constexpr int BLK = 16384;                       // one [128 rows x 64 elements] swizzled block
constexpr int TBLK = 32 * 128;                   // the same for a 32-row box
constexpr int OFF_Q = 0, OFF_K = 4 * BLK, OFF_V = 8 * BLK, OFF_P = 12 * BLK, OFF_MISC = 14 * BLK;
constexpr int SMEM_BYTES = OFF_MISC + 2048 + 1024 /*alignment slack*/;

struct AttnTcParams {
    const int32_t *valid_len;
    int s_vlm, tiles_per_sample;
    float y_scale;      // scale / cap
    float cap_log2e;    // cap * log2(e)
};

// tanh for the soft-cap: |y| <= 0.5 (logits up to +-25) by its odd Taylor polynomial through y^9 on the FMA pipe
// (error < 5e-6, i.e. < 2.5e-4 on a logit); beyond that the exact exponential form
PZ_DEVINL float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
PZ_DEVINL float softcap_tanh(float y) {
    const float y2 = y * y;
    if (y2 <= 0.25f) {
        float r = fmaf(y2, 62.f / 2835.f, -17.f / 315.f);
        r = fmaf(y2, r, 2.f / 15.f);
        r = fmaf(y2, r, -1.f / 3.f);
        r = fmaf(y2, r, 1.f);
        return y * r;
    }
    const float t = __expf(2.f * y);
    return 1.f - __fdividef(2.f, t + 1.f);
}

__global__ void __launch_bounds__(THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
               const __grid_constant__ CUtensorMap map_k32, const __grid_constant__ CUtensorMap map_v32, const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = align_smem(smem_raw, 1024);
    uint8_t *sQ = smem + OFF_Q, *sK = smem + OFF_K, *sV = smem + OFF_V, *sP = smem + OFF_P;
    uint64_t *bars = (uint64_t *)(smem + OFF_MISC);
    uint64_t *q_full = bars, *k_full = bars + 1, *v_full = bars + 2, *s_full = bars + 3 /* [2] */, *p_full = bars + 5,
             *pv_done = bars + 6;
    uint32_t *tmem_slot = (uint32_t *)(bars + 8);
    float *sL = (float *)(smem + OFF_MISC + 1024);   // [2][ROWS] partial row sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.tiles_per_sample, tile = blockIdx.x % p.tiles_per_sample;
    int t0 = tile * TOK;
    if (t0 + TOK > p.s_vlm) t0 = p.s_vlm - TOK;     // the last tile overlaps its predecessor (same values written twice)
    const int row0 = b * p.s_vlm + t0;
    const int vlen = min(max(p.valid_len[b], 0), p.s_vlm);
    const int NC = t0 < vlen ? (vlen + KC - 1) / KC : 0;   // key chunks; a tile of pad rows only writes zeros
    auto nk16 = [&](int c) { int nk = min(vlen - c * KC, KC); return (nk + 15) & ~15; };
    // a chunk of <= 32 keys (the last one: 276 = 2 * 128 + 20) is loaded through 32-row boxes: 4 x 4 KB per operand
    // instead of 4 x 16 KB; its [rows x 64 d] blocks are then TBLK apart
    auto blk_of = [&](int c) { return nk16(c) <= 32 ? TBLK : BLK; };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k32) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v32) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1);
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        mbar_init(p_full, SM_THREADS / 32); mbar_init(pv_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 2 * KC;
    pdl_trigger();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0 && NC > 0) {
            pdl_wait();   // Q, K, V come from the preceding QKV projection
            mbar_expect_tx(q_full, 4 * BLK);
            for (int kd = 0; kd < 4; ++kd) tma_load_3d(&map_q, q_full, sQ + kd * BLK, 64 * kd, 0, row0);
            for (int c = 0; c < NC; ++c) {
                if (c > 0) mbar_wait(&s_full[(c - 1) & 1], ((c - 1) >> 1) & 1);   // S_{c-1} has consumed the K buffer
                const int blk = blk_of(c);
                const CUtensorMap *mk = blk == BLK ? &map_k : &map_k32, *mv = blk == BLK ? &map_v : &map_v32;
                mbar_expect_tx(k_full, 4 * blk);
                for (int kd = 0; kd < 4; ++kd) tma_load_3d(mk, k_full, sK + kd * blk, 64 * kd, KC * c, b);
                if (c > 0) mbar_wait(pv_done, (c - 1) & 1);                       // P V_{c-1} has consumed the V buffer
                mbar_expect_tx(v_full, 4 * blk);
                for (int kd = 0; kd < 4; ++kd) tma_load_3d(mv, v_full, sV + kd * blk, 64 * kd, KC * c, b);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0 && NC > 0) {
            const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
            auto issue_pv = [&](int c) {
                mbar_wait(p_full, c & 1);
                mbar_wait(v_full, c & 1);
                tc_fence_after();
                const uint32_t idesc = umma_idesc_bmn(ROWS, HD);
                const int ks = nk16(c) >> 4;
                for (int k = 0; k < ks; ++k)
                    tc_mma(tmem_o, umma_desc_sw128(aP + (k >> 2) * BLK + (k & 3) * 32), umma_desc_sw128_mn(aV + k * 2048, blk_of(c)), idesc, (c | k) != 0);
                tc_commit(pv_done);
            };
            mbar_wait(q_full, 0);
            for (int c = 0; c < NC; ++c) {
                mbar_wait(k_full, c & 1);
                tc_fence_after();
                const uint32_t idesc = umma_idesc(ROWS, nk16(c));
                const uint32_t d_s = tmem_base + (c & 1) * KC;
                const int kblk = blk_of(c);
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk)
                    tc_mma(d_s, umma_desc_sw128(aQ + (kk >> 2) * BLK + (kk & 3) * 32), umma_desc_sw128(aK + (kk >> 2) * kblk + (kk & 3) * 32), idesc, kk != 0);
                tc_commit(&s_full[c & 1]);
                if (c >= 1) issue_pv(c - 1);
            }
            issue_pv(NC - 1);
        }
    } else if (warp >= SM_WARP0) {
        // ------------------------------------------------------------------ softmax + epilogue
        const int q = warp & 3, half = (warp - SM_WARP0) >> 2;
        const int row = q * 32 + lane;               // tile row = token_local * 8 + head
        const int tok = t0 + (row >> 3);
        const int sw = row & 7;                      // 128B-swizzle phase of this row
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float lsum = 0.f;
        for (int c = 0; c < NC; ++c) {
            const int n16 = nk16(c);
            mbar_wait(&s_full[c & 1], (c >> 1) & 1);
            tc_fence_after();
            if (c > 0) mbar_wait(pv_done, (c - 1) & 1);   // P V_{c-1} has consumed the P buffer
            const int key0 = c * KC;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int col0 = half * 64 + cc * 32;
                if (col0 >= n16) break;
                uint32_t v[32];
                tc_ld32(tmem_base + lane_addr + (c & 1) * KC + col0, v);
                tc_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float p0 = 0.f, p1 = 0.f;
                    if (key0 + col0 + j < vlen) p0 = ex2_approx(p.cap_log2e * softcap_tanh(__uint_as_float(v[j]) * p.y_scale));
                    if (key0 + col0 + j + 1 < vlen) p1 = ex2_approx(p.cap_log2e * softcap_tanh(__uint_as_float(v[j + 1]) * p.y_scale));
                    lsum += p0 + p1;
                    pk[j >> 1] = pack_bf16x2(p0, p1);
                }
                // P[row][key0 + col0 .. +32) -> K-major swizzled operand: atom `half`, 16-byte chunks cc*4 .. cc*4+3
                uint8_t *prow = sP + half * BLK + row * 128;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_shared_v4(prow + (((cc * 4 + i) ^ sw) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
            if (warp == SM_WARP0 && c == NC - 1) {
                // key rows [vlen, n16) of the last V chunk are multiplied by P = 0: they must not hold NaN / Inf bit patterns
                const int nk = vlen - key0;
                if (nk < n16) {
                    mbar_wait(v_full, c & 1);
                    for (int i = lane; i < (n16 - nk) * 32; i += 32) {
                        const int r = nk + (i >> 5), ch = i & 31;   // row, 16-byte chunk across the four d blocks
                        st_shared_v4(sV + (ch >> 3) * blk_of(c) + r * 128 + ((ch & 7) << 4), 0u, 0u, 0u, 0u);
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / rowsum -> bf16 -> staging (the Q buffer, same box layout) -> TMA store
        sL[half * ROWS + row] = lsum;
        if (NC > 0) {
            mbar_wait(pv_done, (NC - 1) & 1);
            tc_fence_after();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");
        const float l = sL[row] + sL[ROWS + row];
        const float inv = (tok < vlen && l > 0.f) ? 1.f / l : 0.f;   // pad rows are written as zeros
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int col0 = half * 128 + cc * 32;
            uint32_t v[32];
            if (NC > 0) {
                tc_ld32(tmem_o + lane_addr + col0, v);
                tc_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            uint8_t *orow = sQ + (col0 >> 6) * BLK + row * 128;
            const int ch0 = (col0 & 63) >> 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t w0 = pack_bf16x2(__uint_as_float(v[8 * i]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
                uint32_t w1 = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
                uint32_t w2 = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
                uint32_t w3 = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
                st_shared_v4(orow + (((ch0 + i) ^ sw) << 4), w0, w1, w2, w3);
            }
        }
        fence_async_smem();
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");
        if (threadIdx.x == SM_WARP0 * 32) {
            if (NC == 0) pdl_wait();   // (the producer waited otherwise) the output buffer may still be read by an earlier kernel
            for (int kd = 0; kd < 4; ++kd) tma_store_3d(&map_o, sQ + kd * BLK, 64 * kd, 0, row0);
            bulk_commit();
            bulk_wait_read0();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// =====================================================================================================
// SigLIP encoder attention (siglip.py:108-166) on tcgen05: 16 heads x head_dim 72, 256 tokens, no mask, no soft-cap.
// One CTA per (image, head); K and V (256 x 72) are loaded once, the two 128-row query tiles run back to back:
//   S = Q K^T (128 x 256 x 80: head_dim padded to 80 by the tensor map's out-of-bounds zero fill, five k-steps) into
//   256 TMEM columns -> 8 softmax warps, exact two-pass softmax (row maximum, then exp and row sum; the scores are
//   unbounded here, unlike the soft-capped Gemma attention) -> P (128 x 256 bf16, K-major swizzled, 64 KB) ->
//   O = P V (128 x 80 x 256; V MN-major, N = 80 spans 1.25 swizzle atoms) into 80 more TMEM columns -> O / rowsum ->
//   TMA store.  Operand blocks keep the 128-byte swizzle: d 64..71 live in a second [rows x 64] block whose d >= 72
//   columns are zeros.
constexpr int V_HDP = 80, V_TOK = 256, V_QT = 128;
constexpr int V_QBLK = V_QT * 128, V_KBLK = V_TOK * 128;                 // bytes of one [rows x 64 d] block
constexpr int V_OFF_Q = 0, V_OFF_K = 2 * V_QBLK, V_OFF_V = V_OFF_K + 2 * V_KBLK, V_OFF_P = V_OFF_V + 2 * V_KBLK,
              V_OFF_MISC = V_OFF_P + 4 * BLK;
constexpr int V_SMEM_BYTES = V_OFF_MISC + 128 + 1024 /*row exchange*/ + 1024 /*alignment slack*/;   // 231 552 <= 232 448

struct AttnVitParams {
    int n_heads;
    float scale_log2e;
};

__global__ void __launch_bounds__(THREADS, 1)
attn_tc_vit_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o, const AttnVitParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = align_smem(smem_raw, 1024);
    uint8_t *sQ = smem + V_OFF_Q, *sK = smem + V_OFF_K, *sV = smem + V_OFF_V, *sP = smem + V_OFF_P;
    uint64_t *bars = (uint64_t *)(smem + V_OFF_MISC);
    uint64_t *q_full = bars, *k_full = bars + 1, *v_full = bars + 2, *s_full = bars + 3, *p_full = bars + 4, *pv_done = bars + 5;
    uint32_t *tmem_slot = (uint32_t *)(bars + 8);
    float *sX = (float *)(smem + V_OFF_MISC + 128);           // [2][128] partial row maxima, then partial row sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x / p.n_heads, head = blockIdx.x % p.n_heads;
    const int row_base = img * V_TOK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1);
        mbar_init(p_full, SM_THREADS / 32); mbar_init(pv_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + V_TOK;
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            pdl_wait();   // q | k | v come from the preceding projection
            mbar_expect_tx(k_full, 2 * V_KBLK);
            for (int kd = 0; kd < 2; ++kd) tma_load_3d(&map_k, k_full, sK + kd * V_KBLK, 64 * kd, head, row_base);
            mbar_expect_tx(q_full, 2 * V_QBLK);
            for (int kd = 0; kd < 2; ++kd) tma_load_3d(&map_q, q_full, sQ + kd * V_QBLK, 64 * kd, head, row_base);
            mbar_expect_tx(v_full, 2 * V_KBLK);
            for (int kd = 0; kd < 2; ++kd) tma_load_3d(&map_v, v_full, sV + kd * V_KBLK, 64 * kd, head, row_base);
            mbar_wait(s_full, 0);   // S of the first query tile has consumed the Q buffer
            mbar_expect_tx(q_full, 2 * V_QBLK);
            for (int kd = 0; kd < 2; ++kd) tma_load_3d(&map_q, q_full, sQ + kd * V_QBLK, 64 * kd, head, row_base + V_QT);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
            mbar_wait(k_full, 0);
            for (int qt = 0; qt < 2; ++qt) {
                mbar_wait(q_full, qt);
                // tile 1: S may be overwritten as soon as the softmax warps have read tile 0's (p_full(0), waited for below
                // in the previous iteration); O is overwritten by P V of tile 1, which waits for p_full(1) -- every softmax
                // warp arrives there only after its epilogue of tile 0 has read O
                tc_fence_after();
                {
                    const uint32_t idesc = umma_idesc(V_QT, V_TOK);
#pragma unroll
                    for (int kk = 0; kk < V_HDP / 16; ++kk)
                        tc_mma(tmem_base, umma_desc_sw128(aQ + (kk >> 2) * V_QBLK + (kk & 3) * 32),
                               umma_desc_sw128(aK + (kk >> 2) * V_KBLK + (kk & 3) * 32), idesc, kk != 0);
                    tc_commit(s_full);
                }
                mbar_wait(p_full, qt);
                if (qt == 0) mbar_wait(v_full, 0);
                tc_fence_after();
                {
                    const uint32_t idesc = umma_idesc_bmn(V_QT, V_HDP);
                    for (int k = 0; k < V_TOK / 16; ++k)
                        tc_mma(tmem_o, umma_desc_sw128(aP + (k >> 2) * BLK + (k & 3) * 32), umma_desc_sw128_mn(aV + k * 2048, V_KBLK), idesc, k != 0);
                    tc_commit(pv_done);
                }
            }
        }
    } else if (warp >= SM_WARP0) {
        const int q = warp & 3, half = (warp - SM_WARP0) >> 2;
        const int row = q * 32 + lane;
        const int sw = row & 7;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        for (int qt = 0; qt < 2; ++qt) {
            mbar_wait(s_full, qt);
            tc_fence_after();
            // pass 1: row maximum over this warp's 128 key columns
            float m = -INFINITY;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t v[32];
                tc_ld32(tmem_base + lane_addr + half * 128 + cc * 32, v);
                tc_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
            }
            sX[half * V_QT + row] = m;
            asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");
            m = fmaxf(sX[row], sX[V_QT + row]);
            asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");   // everybody has read the maxima: sX is reused for the sums
            // pass 2: P = exp((s - m) * scale), row sum
            float lsum = 0.f;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t v[32];
                tc_ld32(tmem_base + lane_addr + half * 128 + cc * 32, v);
                tc_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = ex2_approx((__uint_as_float(v[j]) - m) * p.scale_log2e);
                    const float p1 = ex2_approx((__uint_as_float(v[j + 1]) - m) * p.scale_log2e);
                    lsum += p0 + p1;
                    pk[j >> 1] = pack_bf16x2(p0, p1);
                }
                // keys half*128 + cc*32 .. +32 -> atom (64 keys) half*2 + cc/2, 16-byte chunks (cc & 1)*4 .. +3
                uint8_t *prow = sP + (half * 2 + (cc >> 1)) * BLK + row * 128;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_shared_v4(prow + ((((cc & 1) * 4 + i) ^ sw) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
            sX[half * V_QT + row] = lsum;
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            // ---- epilogue of this query tile
            mbar_wait(pv_done, qt);
            tc_fence_after();
            asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");   // row sums complete; P consumed -> its buffer is the staging area
            const float l = sX[row] + sX[V_QT + row];
            const float inv = 1.f / l;
            // O columns: half 0 -> d 0..63 (block 0), half 1 -> d 64..79 (block 1; only d < 72 is stored)
            const int ncc = half == 0 ? 2 : 1;
            for (int cc = 0; cc < ncc; ++cc) {
                const int col0 = half * 64 + cc * 32;
                uint32_t v[32];
                tc_ld32(tmem_o + lane_addr + col0, v);
                tc_ld_wait();
                uint8_t *orow = sP + (col0 >> 6) * V_QBLK + row * 128;
                const int ch0 = (col0 & 63) >> 3;
                const int nch = half == 0 ? 4 : 2;                         // 16 valid columns in block 1
                for (int i = 0; i < nch; ++i) {
                    uint32_t w0 = pack_bf16x2(__uint_as_float(v[8 * i]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
                    uint32_t w1 = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
                    uint32_t w2 = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
                    uint32_t w3 = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
                    st_shared_v4(orow + (((ch0 + i) ^ sw) << 4), w0, w1, w2, w3);
                }
            }
            fence_async_smem();
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");
            if (threadIdx.x == SM_WARP0 * 32) {
                for (int kd = 0; kd < 2; ++kd) tma_store_3d(&map_o, sP + kd * V_QBLK, 64 * kd, head, row_base + qt * V_QT);
                bulk_commit();
                bulk_wait_read0();
            }
            asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");   // staging (= P buffer) free again
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
// bf16 tensor [d2][d1][d0] (d0 contiguous), strides in elements; 128B-swizzled box
bool make_map3(CUtensorMap *map, const void *base, long d0, long d1, long d2, long stride1, long stride2, int b0, int b1, int b2) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)stride1 * 2, (cuuint64_t)stride2 * 2};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int attn_tc_vit_supported(const AttnArgs &a) {
    static const bool off = [] { const char *e = getenv("PZ_ATTN_TC_VIT"); return e && e[0] == '0'; }();
    if (off) return 0;
    if (a.head_dim <= 64 || a.head_dim > V_HDP || a.head_dim % 8 || a.valid_len || a.n_fresh != 0 || a.q_row0 != 0) return 0;
    if (a.q_rows != V_TOK || a.s_cache != V_TOK || a.softcap != 0.f) return 0;
    if (a.kv_head_stride != a.head_dim || a.q_head_stride != a.head_dim || a.o_head_stride != a.head_dim) return 0;
    if (a.kv_row_stride % 8 || a.q_row_stride % 8 || a.o_row_stride % 8) return 0;
    if (a.q_batch_stride != (long)V_TOK * a.q_row_stride || a.kv_batch_stride != (long)V_TOK * a.kv_row_stride ||
        a.o_batch_stride != (long)V_TOK * a.o_row_stride) return 0;
    if (((uintptr_t)a.Q | (uintptr_t)a.K | (uintptr_t)a.V | (uintptr_t)a.O) & 15) return 0;
    return 1;
}

int launch_attn_tc_vit(const AttnArgs &a, cudaStream_t st) {
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        if (cudaFuncSetAttribute(attn_tc_vit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V_SMEM_BYTES) != cudaSuccess)
            return PZ_ERR_CUDA;
    }
    CUtensorMap mq, mk, mv, mo;
    const long rows = (long)a.batch * V_TOK;
    // [d (head_dim; the 64-wide boxes at d = 64 are zero-filled beyond it)][head][token row]
    if (!make_map3(&mq, a.Q, a.head_dim, a.n_heads, rows, a.q_head_stride, a.q_row_stride, 64, 1, V_QT) ||
        !make_map3(&mo, a.O, a.head_dim, a.n_heads, rows, a.o_head_stride, a.o_row_stride, 64, 1, V_QT) ||
        !make_map3(&mk, a.K, a.head_dim, a.n_heads, rows, a.kv_head_stride, a.kv_row_stride, 64, 1, V_TOK) ||
        !make_map3(&mv, a.V, a.head_dim, a.n_heads, rows, a.kv_head_stride, a.kv_row_stride, 64, 1, V_TOK))
        return PZ_ERR_CUDA;
    AttnVitParams p;
    p.n_heads = a.n_heads;
    p.scale_log2e = a.scale * 1.4426950408889634f;
    launch_k(attn_tc_vit_kernel, dim3((unsigned)(a.batch * a.n_heads)), dim3(THREADS), (size_t)V_SMEM_BYTES, st, mq, mk, mv, mo, p);
    return 0;
}

int attn_tc_supported(const AttnArgs &a) {
    static const bool off = [] { const char *e = getenv("PZ_ATTN_TC"); return e && e[0] == '0'; }();
    if (off) return 0;
    if (a.head_dim != HD || a.n_heads != NH || a.kv_head_stride != 0 || a.n_fresh != 0 || a.q_row0 != 0) return 0;
    if (!a.valid_len || a.q_rows != a.s_vlm || a.s_vlm < TOK || a.s_cache < a.s_vlm) return 0;
    if (!(a.softcap > 0.f && a.softcap <= 80.f)) return 0;   // the max-free softmax needs bounded logits
    if (a.q_head_stride != HD || a.o_head_stride != HD || a.kv_row_stride % 8 || a.q_row_stride % 8 || a.o_row_stride % 8) return 0;
    if (a.q_batch_stride != (long)a.q_rows * a.q_row_stride || a.o_batch_stride != (long)a.q_rows * a.o_row_stride) return 0;
    if (((uintptr_t)a.Q | (uintptr_t)a.K | (uintptr_t)a.V | (uintptr_t)a.O) & 15) return 0;
    return 1;
}

int launch_attn_tc(const AttnArgs &a, cudaStream_t st) {
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        if (cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
            return PZ_ERR_CUDA;
    }
    CUtensorMap mq, mk, mv, mo, mk32, mv32;
    const long rows = (long)a.batch * a.q_rows;
    if (!make_map3(&mq, a.Q, HD, NH, rows, a.q_head_stride, a.q_row_stride, 64, NH, TOK) ||
        !make_map3(&mo, a.O, HD, NH, rows, a.o_head_stride, a.o_row_stride, 64, NH, TOK) ||
        !make_map3(&mk, a.K, HD, a.s_cache, a.batch, a.kv_row_stride, a.kv_batch_stride, 64, KC, 1) ||
        !make_map3(&mv, a.V, HD, a.s_cache, a.batch, a.kv_row_stride, a.kv_batch_stride, 64, KC, 1) ||
        !make_map3(&mk32, a.K, HD, a.s_cache, a.batch, a.kv_row_stride, a.kv_batch_stride, 64, 32, 1) ||
        !make_map3(&mv32, a.V, HD, a.s_cache, a.batch, a.kv_row_stride, a.kv_batch_stride, 64, 32, 1))
        return PZ_ERR_CUDA;
    AttnTcParams p;
    p.valid_len = a.valid_len;
    p.s_vlm = a.s_vlm;
    p.tiles_per_sample = (a.s_vlm + TOK - 1) / TOK;
    p.y_scale = a.scale / a.softcap;
    p.cap_log2e = a.softcap * 1.4426950408889634f;
    launch_k(attn_tc_kernel, dim3((unsigned)(a.batch * p.tiles_per_sample)), dim3(THREADS), (size_t)SMEM_BYTES, st, mq, mk, mv, mo, mk32, mv32, p);
    return 0;
}
