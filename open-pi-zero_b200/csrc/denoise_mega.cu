// denoise_mega.cu -- the whole Euler sampler (pizero.py:454-489) as ONE persistent
// cooperative kernel for small batches (B * horizon <= 16 action rows).
//
// Why: at bs=1 a denoise step is a strictly serial chain of ~110 tiny weight-streaming
// products (629 MB of weights, ~2.5 GFLOP).  As separate kernels every link costs a
// kernel boundary (~5 us even with programmatic dependent launch) and HBM idles between
// them.  Here one CTA per SM stays resident for all 10 steps:
//   * phases (encoder, per layer QKV | attention | o_proj | gate-up | down, decoder) are
//     separated by a grid-wide barrier (one atomic + acquire spin, ~1 us);
//   * weights never wait for a barrier: every thread streams the 16-byte chunks it will
//     later feed to mma.sync through a private cp.async ring in shared memory
//     (4 slots x 32 KB per CTA), prefetching across phase and layer boundaries in the
//     fixed order in which the CTA will consume them, so HBM keeps streaming while the
//     activations synchronise;
//   * every weight item is a 16-row x 1024-k block (8 warps x 128 k), the skinny-kernel
//     MMA mapping (W = A operand, k-permutation, see skinny.cu);
//   * activations cross CTAs through L2 (ld.global.cg) and are staged once per phase as
//     bf16 in shared memory with the deferred epilogue of the producer applied on the fly:
//     RMSNorm for QKV / gate-up / decoder, the split-key softmax combine for o_proj;
//   * attention is one (sample, 64-key tile) item per CTA: Q and the fresh action keys are
//     rotated (RoPE) while staged, S = QK^T, soft-cap, block mask and softmax statistics per
//     tile, unnormalised PV partials to L2 (flash-decoding), combined by the o_proj phase.
// Residual adds are fp32 atomics (o_proj and down are split over K as well as N).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int NT = 256, NW = 8;            // threads / warps per CTA
constexpr int SLOTS = 4;                   // cp.async ring depth (items)
constexpr int ITEM_BYTES = 32768;          // 16 rows x 1024 k x 2 B
constexpr int KI = 1024;                   // k extent of one item
constexpr int MAXM = 16;                   // max action rows (2 MMA n-tiles)
constexpr int LDA = KI + 8;                // staged activation row stride (bf16), conflict-free
constexpr int QROWS = 32, KT = 64, LDQ = 256 + 8;
constexpr int MAX_LAYERS = 24;

struct MegaParams {
    int B, H, M, A, AI, nh, S_v, S_p, S_c, n_layers, n_steps, action_dim, skp, n_splits;
    float dt, clip;
    pz_mix_layer layers[MAX_LAYERS];
    const float *final_norm;
    const bf16 *enc_w1, *enc_w2a, *enc_w3, *dec_w;
    const float *enc_b1, *enc_time_bias, *enc_b3, *dec_b;
    const float *rope_cos, *rope_sin;
    const bf16 *kcache, *vcache;
    long kv_layer_stride, kv_batch_stride;
    const int32_t *valid_len;
    float *act, *xa, *partials, *out;
    bf16 *e1, *z, *qkv, *mlp;
    unsigned int *barrier;   // [0] arrive counter (zeroed before launch), [1] error flag
};

// shared memory map
constexpr int SM_RING = 0;
constexpr int SM_U = SLOTS * ITEM_BYTES;                    // union region
//   GEMV phases
constexpr int SM_ASTAGE = SM_U;                             // bf16 [MAXM][LDA]
constexpr int SM_RED = SM_ASTAGE + MAXM * LDA * 2;          // float [2][NW][16][MAXM + 1] (double buffered)
constexpr int SM_RS = SM_RED + 2 * NW * 16 * (MAXM + 1) * 4;    // float [MAXM]
constexpr int SM_CW = SM_RS + MAXM * 4;                     // float [MAXM][8 heads][8 splits]
constexpr int SM_GEMV_END = SM_CW + MAXM * 8 * 8 * 4;
//   attention phase (aliases the GEMV region)
constexpr int SM_Q = SM_U;                                  // bf16 [32][LDQ]
constexpr int SM_K = SM_Q + QROWS * LDQ * 2;                // bf16 [64][LDQ]
constexpr int SM_V = SM_K + KT * LDQ * 2;
constexpr int SM_S = SM_V + KT * LDQ * 2;                   // float [32][65]
constexpr int SM_P = SM_S + QROWS * (KT + 1) * 4;           // bf16 [32][72]
constexpr int SM_ML = SM_P + QROWS * (KT + 8) * 2;          // float [32][2]
constexpr int SM_ATT_END = SM_ML + QROWS * 2 * 4;
constexpr int SMEM_TOTAL = (SM_ATT_END > SM_GEMV_END ? SM_ATT_END : SM_GEMV_END) + 128;

enum Phase { PH_ENC2 = 0, PH_ENC3, PH_QKV, PH_O, PH_GU, PH_D, PH_DEC };

PZ_DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
PZ_DEVINL void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
PZ_DEVINL void cp_async16_zfill(void *dst, const void *src, bool valid) {
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
PZ_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PZ_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
PZ_DEVINL void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
PZ_DEVINL void ldsm_x4(uint32_t (&r)[4], const void *p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
PZ_DEVINL void ldsm_x4_t(uint32_t (&r)[4], const void *p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
PZ_DEVINL float tanh_fast_acc(float y) { float t = __expf(2.f * y); return 1.f - __fdividef(2.f, t + 1.f); }

// fine-grained trace (CTA 0, thread 0, one layer of one step): tools/stage_times.py prints it
__device__ unsigned long long *g_trace_ptr;
__device__ int g_trace_on;
// Compiled in only with -DPZ_MEGA_TRACE (tools/stage_times.py builds that way): the test of `g_trace_on` is a global
// load (~0.4 us) on thread 0 of CTA 0, which sits on the critical path of every attention / o_proj phase.
PZ_DEVINL void tstamp(int idx) {
#ifndef PZ_MEGA_TRACE
    (void)idx;
    return;
#endif
    if (blockIdx.x == 0 && threadIdx.x == 0 && g_trace_on) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_trace_ptr[idx] = t;
    }
}

// ---- grid barrier: monotonic arrive counter, acquire spin, bounded (never hangs the GPU) ----
template <typename F>
PZ_DEVINL void grid_barrier(unsigned int *bar, unsigned int &target, F between) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
    }
    // work that does not depend on the other CTAs (weight / KV prefetch issue) hides behind the barrier latency
    between();
    if (threadIdx.x == 0) {
        unsigned int v;
        long spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (++spins > (1L << 24)) { atomicExch(bar + 1, 1u); break; }
        } while (v < target);
    }
    __syncthreads();
}

// ---- the CTA's ordered sequence of weight items ------------------------------------------
struct Seq {
    int step, pidx, k;   // pidx indexes the per-step phase list: ENC2, ENC3, (QKV, O, GU, D) x L, DEC
};
PZ_DEVINL int seq_phases(const MegaParams &p) { return 2 + 4 * p.n_layers + 1; }
PZ_DEVINL int seq_type(const MegaParams &p, int pidx) {
    if (pidx < 2) return pidx;                                // ENC2, ENC3
    if (pidx == 2 + 4 * p.n_layers) return PH_DEC;
    return PH_QKV + ((pidx - 2) & 3);
}
PZ_DEVINL int seq_layer(int pidx) { return (pidx - 2) >> 2; }
constexpr int NB_A = KI / 16;   // 16-row blocks of an act_hidden-wide output (act_hidden == KI)
PZ_DEVINL int phase_items(const MegaParams &p, int type) {
    switch (type) {
        case PH_ENC2: case PH_ENC3: return NB_A;
        case PH_QKV: return (p.nh + 2) * 16;
        case PH_O: return NB_A * (p.nh >> 2);
        case PH_GU: return p.AI >> 3;
        case PH_D: return NB_A * (p.AI >> 10);
        default: return 1;                                    // PH_DEC
    }
}
// move to the next valid (step, phase, k) of this CTA; returns false at the end
PZ_DEVINL bool seq_normalize(const MegaParams &p, Seq &s) {
    const int nph = seq_phases(p);
    while (s.step < p.n_steps) {
        int items = phase_items(p, seq_type(p, s.pidx));
        if ((int)blockIdx.x + s.k * (int)gridDim.x < items) return true;
        s.k = 0;
        if (++s.pidx == nph) { s.pidx = 0; ++s.step; }
    }
    return false;
}

// weight rows + k offset of one item for this thread: (row g, row g+8 of the 16-row tile, k base)
struct ItemAddr { const bf16 *w0, *w1; };
PZ_DEVINL ItemAddr item_addr(const MegaParams &p, int type, int layer, int item, int g) {
    const bf16 *W; int K, nb, ks = 0, nrows;
    switch (type) {
        case PH_ENC2: W = p.enc_w2a; K = KI; nb = item; nrows = KI; break;
        case PH_ENC3: W = p.enc_w3; K = KI; nb = item; nrows = KI; break;
        case PH_QKV: W = (const bf16 *)p.layers[layer].w_qkv; K = KI; nb = item; nrows = (p.nh + 2) * 256; break;
        case PH_O: W = (const bf16 *)p.layers[layer].w_o; K = p.nh * 256; nb = item & (NB_A - 1); ks = item / NB_A; nrows = KI; break;
        case PH_GU: W = (const bf16 *)p.layers[layer].w_gate_up; K = KI; nb = item; nrows = 2 * p.AI; break;
        case PH_D: W = (const bf16 *)p.layers[layer].w_down; K = p.AI; nb = item & (NB_A - 1); ks = item / NB_A; nrows = KI; break;
        default: W = p.dec_w; K = KI; nb = 0; nrows = 8; break;
    }
    long r0, r1;
    if (type == PH_GU) {   // 8 gate rows + the 8 matching up rows of the packed [128 gate | 128 up] layout
        int c = nb * 8 + g;
        r0 = (long)(c / PZ_GU_BLOCK) * (2 * PZ_GU_BLOCK) + (c % PZ_GU_BLOCK);
        r1 = r0 + PZ_GU_BLOCK;
    } else if (type == PH_QKV) {
        // item = (head nb / 16, dims d0 = 8 * (nb % 16)): rows d0.. and 128 + d0.. -- the two halves a rotary pair
        // lives in, so the epilogue rotates q and k (model/utils.py:4-16) and attention stages plain copies
        r0 = (long)(nb >> 4) * 256 + (nb & 15) * 8 + g;
        r1 = r0 + 128;
    } else {
        r0 = nb * 16 + g; r1 = r0 + 8;
        if (r0 >= nrows) r0 = nrows - 1;
        if (r1 >= nrows) r1 = nrows - 1;
    }
    ItemAddr a;
    a.w0 = W + r0 * K + (long)ks * KI;
    a.w1 = W + r1 * K + (long)ks * KI;
    return a;
}

// issue the 8 private 16-byte chunks of one item into ring slot `slot`
PZ_DEVINL void prefetch_item(const MegaParams &p, uint8_t *smem, int slot, const Seq &s) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    int type = seq_type(p, s.pidx);
    ItemAddr a = item_addr(p, type, seq_layer(s.pidx), blockIdx.x + s.k * gridDim.x, g);
    uint8_t *dst = smem + SM_RING + slot * ITEM_BYTES + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int u = j >> 2, hk = (j >> 1) & 1;
        int k = warp * 128 + u * 64 + hk * 32 + 8 * t;
        cp_async16(dst + j * (NT * 16), ((j & 1) ? a.w1 : a.w0) + k);
    }
}

// ---- one weight item: 16 rows x 1024 k against the staged activations -----------------------
template <int MT>
PZ_DEVINL void gemv_item(const MegaParams &p, uint8_t *smem, int slot, int type, int layer, int item, int step,
                         int parity) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bf16 *As = reinterpret_cast<const bf16 *>(smem + SM_ASTAGE);
    // the cross-warp reduction scratch alternates between two buffers, so a single barrier per item is
    // enough: buffer `parity` is rewritten two items later, i.e. after the next item's barrier
    float(*red)[16][MAXM + 1] =
        reinterpret_cast<float(*)[16][MAXM + 1]>(smem + SM_RED + parity * (NW * 16 * (MAXM + 1) * 4));
    const uint8_t *src = smem + SM_RING + slot * ITEM_BYTES + threadIdx.x * 16;
    float acc[MT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // q = (u, hk): one 16-byte chunk pair (row g, row g+8)
        uint4 ag = *reinterpret_cast<const uint4 *>(src + (2 * q) * (NT * 16));
        uint4 ag8 = *reinterpret_cast<const uint4 *>(src + (2 * q + 1) * (NT * 16));
        int k = warp * 128 + (q >> 1) * 64 + (q & 1) * 32 + 8 * t;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            int m = mt * 8 + g;
            uint4 x = make_uint4(0, 0, 0, 0);
            if (m < p.M) x = *reinterpret_cast<const uint4 *>(As + m * LDA + k);
            mma_bf16(acc[mt], ag.x, ag8.x, ag.y, ag8.y, x.x, x.y);
            mma_bf16(acc[mt], ag.z, ag8.z, ag.w, ag8.w, x.z, x.w);
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        red[warp][g][mt * 8 + 2 * t] = acc[mt][0];
        red[warp][g][mt * 8 + 2 * t + 1] = acc[mt][1];
        red[warp][g + 8][mt * 8 + 2 * t] = acc[mt][2];
        red[warp][g + 8][mt * 8 + 2 * t + 1] = acc[mt][3];
    }
    __syncthreads();
    const int rsh = (type == PH_GU) ? 3 : 4, nrows = 1 << rsh;
    for (int i = threadIdx.x; i < nrows * p.M; i += NT) {
        int r = i & (nrows - 1), m = i >> rsh;
        float v = 0.f, v2 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            v += red[w][r][m];
            if (type == PH_GU) v2 += red[w][r + 8][m];
            if (type == PH_QKV) v2 += red[w][r ^ 8][m];   // the rotary partner (dim d +- 128)
        }
        switch (type) {
            case PH_ENC2: {   // linear_2 (action half) + per-step time bias, SiLU (vla/modules.py:50-52)
                int n = item * 16 + r;
                p.z[m * p.A + n] = __float2bfloat16_rn(silu(v + p.enc_time_bias[step * p.A + n]));
            } break;
            case PH_ENC3: {   // linear_3, then the sqrt(hidden) embed scale (joint_model.py:348-355)
                int n = item * 16 + r;
                p.xa[m * p.A + n] = (v + p.enc_b3[n]) * sqrtf((float)p.A);
            } break;
            case PH_QKV: {   // q / k rows leave rotated (fp32, table row S_p + token: positions 2.., pizero.py:312-318)
                const int head = item >> 4, d = (item & 15) * 8 + (r & 7);
                float o = v;
                if (head <= p.nh) {
                    const long ti = (long)(p.S_p + m % p.H) * 128 + d;
                    const float cs = __ldg(p.rope_cos + ti), sn = __ldg(p.rope_sin + ti);
                    o = (r < 8) ? v * cs - v2 * sn : v * cs + v2 * sn;
                }
                p.qkv[m * ((p.nh + 2) * 256) + head * 256 + (r < 8 ? d : 128 + d)] = __float2bfloat16_rn(o);
            } break;
            case PH_O: case PH_D: {
                int n = (item & (NB_A - 1)) * 16 + r;
                atomicAdd(p.xa + m * KI + n, v);
            } break;
            case PH_GU: {
                int n = item * 8 + r;
                p.mlp[m * p.AI + n] = __float2bfloat16_rn(gelu_tanh(v) * v2);
            } break;
            default: {        // decoder + Euler update (pizero.py:479-481); the only writer of act
                int n = r;
                if (n < p.action_dim) {
                    float a = p.act[m * p.action_dim + n] + p.dt * (v + p.dec_b[n]);
                    p.act[m * p.action_dim + n] = a;
                    if (step == p.n_steps - 1) {
                        if (p.clip >= 0.f) a = fminf(fmaxf(a, -p.clip), p.clip);
                        p.out[m * p.action_dim + n] = a;
                    }
                }
            } break;
        }
    }
}

// ---- activation staging ------------------------------------------------------------------------
// A[m][0..1024) bf16 in shared memory, built from what the previous phase left in L2.  These loads
// sit on the critical path right after a grid barrier, so every thread issues ALL of its loads
// before it consumes any of them (one L2 round trip per stage instead of one per loop iteration).
PZ_DEVINL void stage_norm(const MegaParams &p, uint8_t *smem, const float *norm_w) {
    // Gemma RMSNorm of the fp32 residual (paligemma/modules.py:13-21); A == 1024: 64 threads per row
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM_ASTAGE);
    float *part = reinterpret_cast<float *>(smem + SM_RS);   // [4 rows][2 warps]
    const int tid = threadIdx.x, lane = tid & 31, c = tid & 63, rsub = tid >> 6;
    float4 w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = __ldg(reinterpret_cast<const float4 *>(norm_w) + c + 64 * j);
    for (int m0 = 0; m0 < p.M; m0 += 4) {
        const int m = m0 + rsub;
        const bool ok = m < p.M;
        float4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            x[j] = ok ? __ldcg(reinterpret_cast<const float4 *>(p.xa + (long)m * KI) + c + 64 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) ss += x[j].x * x[j].x + x[j].y * x[j].y + x[j].z * x[j].z + x[j].w * x[j].w;
        ss = warp_sum(ss);
        if (lane == 0) part[rsub * 2 + ((tid >> 5) & 1)] = ss;
        __syncthreads();
        const float r = rsqrtf((part[rsub * 2] + part[rsub * 2 + 1]) / KI + 1e-6f);
        if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint2 o;
                o.x = pack_bf16x2(x[j].x * r * (1.f + w[j].x), x[j].y * r * (1.f + w[j].y));
                o.y = pack_bf16x2(x[j].z * r * (1.f + w[j].z), x[j].w * r * (1.f + w[j].w));
                *reinterpret_cast<uint2 *>(As + m * LDA + (c + 64 * j) * 4) = o;
            }
        }
        __syncthreads();
    }
}
PZ_DEVINL void stage_copy(const MegaParams &p, uint8_t *smem, const bf16 *src, int ld, int k0) {
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM_ASTAGE);
    const int total = p.M * (KI / 8);
    uint4 v[MAXM * (KI / 8) / NT];
#pragma unroll
    for (int j = 0; j < MAXM * (KI / 8) / NT; ++j) {
        int i = threadIdx.x + j * NT;
        if (i < total) v[j] = __ldcg(reinterpret_cast<const uint4 *>(src + (long)(i / (KI / 8)) * ld + k0 + (i % (KI / 8)) * 8));
    }
#pragma unroll
    for (int j = 0; j < MAXM * (KI / 8) / NT; ++j) {
        int i = threadIdx.x + j * NT;
        if (i < total) *reinterpret_cast<uint4 *>(As + (i / (KI / 8)) * LDA + (i % (KI / 8)) * 8) = v[j];
    }
    __syncthreads();
}
PZ_DEVINL void stage_combine(const MegaParams &p, uint8_t *smem, int ks) {
    // combine of the split-key attention partials for heads [4*ks, 4*ks+4) (k range ks*1024..):
    // A[m][k] = sum_s o_s[k] / sum_s l_s   (no maxima: see attention_item).  One L2 round trip for the
    // denominators, one for the bf16 partials.
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM_ASTAGE);
    float *inv = reinterpret_cast<float *>(smem + SM_CW);   // [m][head 0..3 of this k slice]
    const int rows_total = p.nh * p.H, hpk = KI / 256;
    const uint32_t *po = reinterpret_cast<const uint32_t *>(p.partials);
    const float *pl = p.partials + (long)p.B * p.n_splits * rows_total * 128;
    (void)inv;
    for (int i = threadIdx.x; i < p.M * (KI / 8); i += NT) {
        int m = i / (KI / 8), c = i % (KI / 8);
        int hl = (c * 8) / 256, d = (c * 8) % 256, hh = ks * hpk + hl;
        int b = m / p.H, tok = m % p.H;
        const uint32_t *row = po + (((long)b * p.n_splits) * rows_total + (hh * p.H + tok)) * 128 + (d >> 1);
        const float *lrow = pl + (long)b * p.n_splits * rows_total + hh * p.H + tok;
        const long split_stride = (long)rows_total * 128;
        // denominators and partials in the same L2 round trip (every thread sums its row's few l values itself)
        uint4 v[8];
        float ls[8];
#pragma unroll
        for (int sp = 0; sp < 8; ++sp) {
            const bool ok = sp < p.n_splits;
            v[sp] = ok ? __ldcg(reinterpret_cast<const uint4 *>(row + sp * split_stride)) : make_uint4(0, 0, 0, 0);
            ls[sp] = ok ? __ldcg(lrow + (long)sp * rows_total) : 0.f;
        }
        float o8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, l = 0.f;
#pragma unroll
        for (int sp = 0; sp < 8; ++sp) {
            uint32_t w[4] = {v[sp].x, v[sp].y, v[sp].z, v[sp].w};
            l += ls[sp];
#pragma unroll
            for (int j = 0; j < 4; ++j) { o8[2 * j] += bf16lo(w[j]); o8[2 * j + 1] += bf16hi(w[j]); }
        }
        const float wgt = l > 0.f ? 1.f / l : 0.f;
        *reinterpret_cast<uint4 *>(As + m * LDA + c * 8) =
            make_uint4(pack_bf16x2(o8[0] * wgt, o8[1] * wgt), pack_bf16x2(o8[2] * wgt, o8[3] * wgt),
                       pack_bf16x2(o8[4] * wgt, o8[5] * wgt), pack_bf16x2(o8[6] * wgt, o8[7] * wgt));
    }
    __syncthreads();
}

// ---- attention: one (sample, 64-key tile) item --------------------------------------------------
PZ_DEVINL void rope_pair(const MegaParams &p, bf16 *dst_row, const bf16 *src_row, int c, int pos) {
    uint4 r1 = __ldcg(reinterpret_cast<const uint4 *>(src_row + c * 8));
    uint4 r2 = __ldcg(reinterpret_cast<const uint4 *>(src_row + 128 + c * 8));
    const float4 *cs = reinterpret_cast<const float4 *>(p.rope_cos + (long)pos * 128 + c * 8);
    const float4 *sn = reinterpret_cast<const float4 *>(p.rope_sin + (long)pos * 128 + c * 8);
    float4 c0 = __ldg(cs), c1 = __ldg(cs + 1), s0 = __ldg(sn), s1 = __ldg(sn + 1);
    float cf[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    float sf[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    uint32_t w1[4] = {r1.x, r1.y, r1.z, r1.w}, w2[4] = {r2.x, r2.y, r2.z, r2.w}, o1[4], o2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float x1a = bf16lo(w1[j]), x1b = bf16hi(w1[j]), x2a = bf16lo(w2[j]), x2b = bf16hi(w2[j]);
        o1[j] = pack_bf16x2(x1a * cf[2 * j] - x2a * sf[2 * j], x1b * cf[2 * j + 1] - x2b * sf[2 * j + 1]);
        o2[j] = pack_bf16x2(x2a * cf[2 * j] + x1a * sf[2 * j], x2b * cf[2 * j + 1] + x1b * sf[2 * j + 1]);
    }
    *reinterpret_cast<uint4 *>(dst_row + c * 8) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<uint4 *>(dst_row + 128 + c * 8) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
}

// cached K / V rows of one (sample, key tile): cp.async into the tile buffers (no dependence on the current step)
PZ_DEVINL void attention_prefetch_kv(const MegaParams &p, uint8_t *smem, int layer, int b, int tile) {
    bf16 *sK = reinterpret_cast<bf16 *>(smem + SM_K);
    bf16 *sV = reinterpret_cast<bf16 *>(smem + SM_V);
    const bf16 *Kc = p.kcache + (long)layer * p.kv_layer_stride + (long)b * p.kv_batch_stride;
    const bf16 *Vc = p.vcache + (long)layer * p.kv_layer_stride + (long)b * p.kv_batch_stride;
    for (int i = threadIdx.x; i < KT * 32; i += NT) {
        int r = i >> 5, c = i & 31;
        int j = tile * KT + r;
        if (j < p.S_c) {
            cp_async16(sK + r * LDQ + c * 8, Kc + (long)j * 256 + c * 8);
            cp_async16(sV + r * LDQ + c * 8, Vc + (long)j * 256 + c * 8);
        }
    }
    cp_async_commit();
}

template <bool KV_PREFETCHED = false, bool ROPED = false>   // ROPED: q / k rows of p.qkv are already rotated
PZ_DEVINL void attention_item(const MegaParams &p, uint8_t *smem, int layer, int b, int tile) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    bf16 *sQ = reinterpret_cast<bf16 *>(smem + SM_Q);
    bf16 *sK = reinterpret_cast<bf16 *>(smem + SM_K);
    bf16 *sV = reinterpret_cast<bf16 *>(smem + SM_V);
    float(*sRS)[QROWS] = reinterpret_cast<float(*)[QROWS]>(smem + SM_S);   // [4 key groups][row] partial row sums
    bf16 *sP = reinterpret_cast<bf16 *>(smem + SM_P);
    const int qkvd = (p.nh + 2) * 256, qd = p.nh * 256;
    const int rows_total = p.nh * p.H;
    const int vlen = p.valid_len[b];
    const int n_keys = p.S_c + p.H;
    const bf16 *Kc = p.kcache + (long)layer * p.kv_layer_stride + (long)b * p.kv_batch_stride;
    const bf16 *Vc = p.vcache + (long)layer * p.kv_layer_stride + (long)b * p.kv_batch_stride;
    const bf16 *qkv_b = p.qkv + (long)b * p.H * qkvd;

    // K / V tile: cached rows by cp.async, fresh (action) rows from this step's raw projections
    for (int i = tid; i < KT * 32; i += NT) {
        int r = i >> 5, c = i & 31;
        int j = tile * KT + r;
        bool cached = j < p.S_c;
        bool fresh = !cached && j < n_keys;
        if (KV_PREFETCHED && cached) continue;   // already in flight (issued before the grid barrier)
        if (!fresh) cp_async16_zfill(sK + r * LDQ + c * 8, cached ? Kc + (long)j * 256 + c * 8 : Kc, cached);
        const bf16 *vsrc = cached ? Vc + (long)j * 256 + c * 8 : (fresh ? qkv_b + (long)(j - p.S_c) * qkvd + qd + 256 + c * 8 : Vc);
        cp_async16_zfill(sV + r * LDQ + c * 8, vsrc, cached || fresh);
    }
    cp_async_commit();
    // Q (and fresh keys): row i = (head i / H, token i % H); RoPE at position S_p + token (table row) unless the
    // producer already applied it
    for (int i = tid; i < QROWS * 16; i += NT) {
        int r = i >> 4, c = i & 15;
        bf16 *dst = sQ + r * LDQ;
        if (r < rows_total) {
            int h = r / p.H, tok = r % p.H;
            const bf16 *src = qkv_b + (long)tok * qkvd + h * 256;
            if (ROPED) {   // plain asynchronous copies: everything of this staging is in flight at once
                cp_async16(dst + c * 8, src + c * 8);
                cp_async16(dst + 128 + c * 8, src + 128 + c * 8);
            } else {
                rope_pair(p, dst, src, c, p.S_p + tok);
            }
        } else {
            *reinterpret_cast<uint4 *>(dst + c * 8) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(dst + 128 + c * 8) = make_uint4(0, 0, 0, 0);
        }
    }
    {   // fresh (action) keys of this step
        int first = p.S_c - tile * KT;   // tile-local row of the first fresh key
        for (int i = tid; i < p.H * 16; i += NT) {
            int r = first + (i >> 4), c = i & 15;
            if (r >= 0 && r < KT) {
                const bf16 *src = qkv_b + (long)(i >> 4) * qkvd + qd;
                if (ROPED) {
                    cp_async16(sK + r * LDQ + c * 8, src + c * 8);
                    cp_async16(sK + r * LDQ + 128 + c * 8, src + 128 + c * 8);
                } else {
                    rope_pair(p, sK + r * LDQ, src, c, p.S_p + (i >> 4));
                }
            }
        }
    }
    cp_async_commit();
    cp_async_wait<0>();   // (also drains this thread's ring prefetches; they are far ahead anyway)
    __syncthreads();
    tstamp(11);

    // S = Q K^T : warp -> (16-row tile mt, 16-key group kg)
    {
        const int mt = warp & 1, kg = warp >> 1;
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
            uint32_t qa[4], kb[4];
            ldsm_x4(qa, sQ + (mt * 16 + (lane & 15)) * LDQ + ks * 16 + (lane >> 4) * 8);
            ldsm_x4(kb, sK + (kg * 16 + (lane & 7) + ((lane >> 4) << 3)) * LDQ + ks * 16 + ((lane >> 3) & 1) * 8);
            mma_bf16(s[0], qa[0], qa[1], qa[2], qa[3], kb[0], kb[1]);
            mma_bf16(s[1], qa[0], qa[1], qa[2], qa[3], kb[2], kb[3]);
        }
        const float scale = 0.0625f, cap = 50.f;   // 1/sqrt(256); soft-cap (joint_model.py:139,261-268)
        // The soft-cap bounds every logit to +-50, so exp() cannot overflow and the softmax needs no running
        // maximum: P = exp(logit) goes straight to shared memory (bf16) with per-row partial sums; the split-key
        // partials combine as sum(o) / sum(l) without any rescaling.
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            float pe[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int col = kg * 16 + nt * 8 + 2 * t + (e & 1);
                int j = tile * KT + col;
                bool vis = (j < vlen) || (j >= p.S_v && j < n_keys);
                pe[e] = vis ? __expf(tanh_fast_acc(s[nt][e] * scale * (1.f / cap)) * cap) : 0.f;
            }
            rs0 += pe[0] + pe[1];
            rs1 += pe[2] + pe[3];
            const int col = kg * 16 + nt * 8 + 2 * t;
            *reinterpret_cast<uint32_t *>(sP + (mt * 16 + g) * (KT + 8) + col) = pack_bf16x2(pe[0], pe[1]);
            *reinterpret_cast<uint32_t *>(sP + (mt * 16 + g + 8) * (KT + 8) + col) = pack_bf16x2(pe[2], pe[3]);
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        if (t == 0) { sRS[kg][mt * 16 + g] = rs0; sRS[kg][mt * 16 + g + 8] = rs1; }
    }
    __syncthreads();
    tstamp(12);
    tstamp(13);
    // O = P V : warp -> (16-row tile mt, 64-wide slice of d)
    {
        const int mt = warp & 1, dq = warp >> 1;
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
            uint32_t pa[4];
            ldsm_x4(pa, sP + (mt * 16 + (lane & 15)) * (KT + 8) + kk * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t vb[4];
                ldsm_x4_t(vb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDQ + dq * 64 + dp * 16 + (lane >> 4) * 8);
                mma_bf16(o[2 * dp], pa[0], pa[1], pa[2], pa[3], vb[0], vb[1]);
                mma_bf16(o[2 * dp + 1], pa[0], pa[1], pa[2], pa[3], vb[2], vb[3]);
            }
        }
        // partial layout: bf16 pairs o[b][split][row][128] (unnormalised), then fp32 l[b][split][row]
        uint32_t *po = reinterpret_cast<uint32_t *>(p.partials) + (((long)b * p.n_splits + tile) * rows_total) * 128;
        float *pl = p.partials + (long)p.B * p.n_splits * rows_total * 128 + ((long)b * p.n_splits + tile) * rows_total;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            int row = mt * 16 + g + rr * 8;
            if (row >= rows_total) continue;
            uint32_t *dst = po + (long)row * 128;
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[(dq * 64 + i * 8 + 2 * t) >> 1] = pack_bf16x2(o[i][2 * rr], o[i][2 * rr + 1]);
            if (dq == 0 && t == 0) pl[row] = (sRS[0][row] + sRS[1][row]) + (sRS[2][row] + sRS[3][row]);
        }
    }
    __syncthreads();
}

// ---- the attention item as a stand-alone kernel (decode attention for batches the persistent
// kernel does not cover): grid = batch * key tiles, partials combined by attn_combine_kernel ----------
__global__ void __launch_bounds__(NT, 2) decode_attn_kernel(const __grid_constant__ MegaParams p, int layer) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // attention_item addresses shared memory at SM_U + ...; rebase so that the union region starts at 0
    uint8_t *smem = align_smem(smem_raw, 128) - SM_U;
    pdl_trigger();
    pdl_wait();
    const int it = blockIdx.x;
    attention_item<false, false>(p, smem, layer, it / p.n_splits, it % p.n_splits);
}

// combine of the stand-alone decode attention's partials: one warp per (sample, query row) -> bf16 attention output
__global__ void __launch_bounds__(128) decode_combine_kernel(const __grid_constant__ MegaParams p, bf16 *out, long out_batch_stride,
                                                             int out_row_stride) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int rows_total = p.nh * p.H;
    const long wid = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (wid >= (long)p.B * rows_total) return;
    const int b = wid / rows_total, row = wid % rows_total;
    const uint32_t *po = reinterpret_cast<const uint32_t *>(p.partials) + (((long)b * p.n_splits) * rows_total + row) * 128;
    const float *pl = p.partials + (long)p.B * p.n_splits * rows_total * 128 + (long)b * p.n_splits * rows_total + row;
    float l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int sp = 0; sp < p.n_splits; ++sp) {
        l += pl[(long)sp * rows_total];
        uint4 v = *reinterpret_cast<const uint4 *>(po + (long)sp * rows_total * 128 + lane * 4);   // dims 8*lane .. 8*lane+7
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[2 * j] += bf16lo(w[j]); acc[2 * j + 1] += bf16hi(w[j]); }
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const int h = row / p.H, tok = row % p.H;
    bf16 *dst = out + b * out_batch_stride + (long)tok * out_row_stride + h * 256 + lane * 8;
    *reinterpret_cast<uint4 *>(dst) = make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                                                 pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
}

// ---- the kernel ------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(NT, 1) denoise_mega_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = align_smem(smem_raw, 128);
    unsigned int target = 0;

    // start streaming: the first SLOTS items of this CTA's sequence
    Seq pre = {0, 0, 0};
    int issued = 0;
    for (int s = 0; s < SLOTS; ++s) {
        if (seq_normalize(p, pre)) { prefetch_item(p, smem, s, pre); ++pre.k; ++issued; }
        cp_async_commit();
    }
    int consumed = 0;
    unsigned long long *trace = reinterpret_cast<unsigned long long *>(p.barrier + 32);
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_trace_ptr = trace; g_trace_on = 0; }
    auto stamp = [&](int step, int l, int idx) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && step == 1 && l == 1) {
            g_trace_on = (idx < 10);   // sub-phase stamps (slots 11..) only inside the traced layer
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            trace[idx] = t;
        }
    };
    // Ring refills are not issued between the items of a phase (their address arithmetic and 8 cp.async per
    // thread cost ~0.5 us on the critical path each): they are deferred to `flush_refills`, which runs behind the
    // latency of the next grid barrier.  Every item still gets exactly one commit group, in consumption order.
    int deferred = 0;
    auto flush_refills = [&]() {
        for (; deferred > 0; --deferred) {
            const int slot = (consumed - deferred) % SLOTS;
            if (seq_normalize(p, pre)) { prefetch_item(p, smem, slot, pre); ++pre.k; }
            cp_async_commit();
        }
    };
    auto run_phase = [&](int step, int pidx) {
        int type = seq_type(p, pidx), layer = seq_layer(pidx);
        int items = phase_items(p, type);
        int staged_ks = -1;
        for (int k = 0; (int)blockIdx.x + k * (int)gridDim.x < items; ++k) {
            int item = blockIdx.x + k * gridDim.x;
            // activations for this item (shared by all items of the phase except for the K-split ones)
            int ks = (type == PH_O || type == PH_D) ? item / NB_A : 0;
            if (staged_ks != ks) {
                switch (type) {
                    case PH_ENC2: stage_copy(p, smem, p.e1, p.A, 0); break;
                    case PH_ENC3: stage_copy(p, smem, p.z, p.A, 0); break;
                    case PH_QKV: stage_norm(p, smem, p.layers[layer].norm_in); break;
                    case PH_O: stage_combine(p, smem, ks); tstamp(14); break;
                    case PH_GU: stage_norm(p, smem, p.layers[layer].norm_post); break;
                    case PH_D: stage_copy(p, smem, p.mlp, p.AI, ks * KI); break;
                    default: stage_norm(p, smem, p.final_norm); break;
                }
                staged_ks = ks;
            }
            if (deferred >= SLOTS) flush_refills();   // more items than ring slots in one phase: refill now
            cp_async_wait<0>();                       // every outstanding item (issued one phase or more ago) has landed
            gemv_item<MT>(p, smem, consumed % SLOTS, type, layer, item, step, consumed & 1);
            ++consumed;
            ++deferred;
        }
    };
    auto nothing = [&]() {};
    (void)nothing;

    const int n_att_items = p.B * p.n_splits;
    for (int step = 0; step < p.n_steps; ++step) {
        // ENC1: linear_1 (action_dim -> A) (vla/modules.py:39-41); K is tiny, no streaming needed
        {
            float *sact = reinterpret_cast<float *>(smem + SM_RED);   // [M][action_dim], bf16-rounded like the GEMV input
            for (int i = threadIdx.x; i < p.M * p.action_dim; i += NT)
                sact[i] = __bfloat162float(__float2bfloat16_rn(__ldcg(p.act + i)));
            __syncthreads();
            for (int i = blockIdx.x * NT + threadIdx.x; i < p.M * p.A; i += gridDim.x * NT) {
                int m = i / p.A, n = i % p.A;
                float v = p.enc_b1[n];
                uint4 wv = __ldg(reinterpret_cast<const uint4 *>(p.enc_w1 + (long)n * p.skp));   // skp >= 8
                uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < p.action_dim) v += ((k & 1) ? bf16hi(ww[k >> 1]) : bf16lo(ww[k >> 1])) * sact[m * p.action_dim + k];
                p.e1[i] = __float2bfloat16_rn(v);
            }
        }
        grid_barrier(p.barrier, target, flush_refills);
        run_phase(step, 0);   // ENC2
        grid_barrier(p.barrier, target, flush_refills);
        run_phase(step, 1);   // ENC3
        grid_barrier(p.barrier, target, flush_refills);
        for (int l = 0; l < p.n_layers; ++l) {
            stamp(step, l, 0);
            run_phase(step, 2 + 4 * l);          // QKV
            stamp(step, l, 1);
            // attention CTAs start fetching their cached K / V tile behind the barrier (the tile buffers alias the
            // GEMV staging area, which is idle after the barrier's leading __syncthreads)
            grid_barrier(p.barrier, target, [&]() {
                flush_refills();
                if ((int)blockIdx.x < n_att_items) attention_prefetch_kv(p, smem, l, blockIdx.x / p.n_splits, blockIdx.x % p.n_splits);
            });
            stamp(step, l, 2);
            if ((int)blockIdx.x < n_att_items) attention_item<true, true>(p, smem, l, blockIdx.x / p.n_splits, blockIdx.x % p.n_splits);
            stamp(step, l, 3);
            grid_barrier(p.barrier, target, flush_refills);
            stamp(step, l, 4);
            run_phase(step, 2 + 4 * l + 1);      // O
            stamp(step, l, 5);
            grid_barrier(p.barrier, target, flush_refills);
            stamp(step, l, 6);
            run_phase(step, 2 + 4 * l + 2);      // GU
            stamp(step, l, 7);
            grid_barrier(p.barrier, target, flush_refills);
            stamp(step, l, 8);
            run_phase(step, 2 + 4 * l + 3);      // D
            stamp(step, l, 9);
            grid_barrier(p.barrier, target, flush_refills);
            stamp(step, l, 10);
        }
        run_phase(step, 2 + 4 * p.n_layers);     // DEC (+ Euler; last step also clamps into out)
        grid_barrier(p.barrier, target, flush_refills);
    }
    cp_async_wait<0>();
}

}  // namespace

// host side --------------------------------------------------------------------------------------------
int denoise_mega_supported(const pz_config &c, int B) {
    static const bool off = [] { const char *e = getenv("PZ_MEGA"); return e && e[0] == '0'; }();
    if (off) return 0;
    if (c.dtype != PZ_BF16 || (c.flags & PZ_FLAG_SIMPLE_KERNELS)) return 0;
    if (B * c.horizon > MAXM || c.n_heads * c.horizon > QROWS) return 0;
    if (c.head_dim != 256 || c.n_kv_heads != 1 || c.n_heads > 8) return 0;
    if (c.act_hidden != KI || c.act_inter % KI || (c.n_heads * 256) % KI) return 0;
    if (c.n_layers > MAX_LAYERS || c.action_dim > 8) return 0;
    if ((c.s_vlm + c.cond_steps + c.horizon + KT - 1) / KT > 8) return 0;
    if (B * ((c.s_vlm + c.cond_steps + c.horizon + KT - 1) / KT) > 128) return 0;   // one attention item per CTA
    return 1;
}

int launch_decode_attention(const pz_config &c, const pz_weights &w, const void *qkv, const void *kcache,
                            const void *vcache, int batch_total, const int32_t *valid_len, float *partials, int layer,
                            int B, void *out, long out_batch_stride, int out_row_stride, cudaStream_t st) {
    static PerDeviceOnce attr_once;
    constexpr int smem = SM_ATT_END - SM_U + 256;
    if (attr_once.need()) {
        if (cudaFuncSetAttribute(decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return PZ_ERR_CUDA;
    }
    MegaParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = c.horizon; p.M = B * c.horizon; p.nh = c.n_heads;
    p.S_v = c.s_vlm; p.S_p = c.cond_steps; p.S_c = c.s_vlm + c.cond_steps;
    p.n_splits = (p.S_c + p.H + KT - 1) / KT;
    p.rope_cos = w.rope_act_cos; p.rope_sin = w.rope_act_sin;
    p.kcache = (const bf16 *)kcache; p.vcache = (const bf16 *)vcache;
    p.kv_batch_stride = (long)p.S_c * 256; p.kv_layer_stride = (long)batch_total * p.kv_batch_stride;
    p.valid_len = valid_len; p.partials = partials; p.qkv = (bf16 *)const_cast<void *>(qkv);
    launch_k(decode_attn_kernel, dim3(B * p.n_splits), dim3(NT), smem, st, p, layer);
    // split-key partials -> attention output [B][horizon][n_heads * 256]
    const long warps = (long)B * p.nh * p.H;
    launch_k(decode_combine_kernel, dim3((unsigned)((warps + 3) / 4)), dim3(128), 0, st, p, (bf16 *)out, out_batch_stride,
             out_row_stride);
    return p.n_splits;
}

int decode_attention_supported(const pz_config &c) {
    return c.dtype == PZ_BF16 && !(c.flags & PZ_FLAG_SIMPLE_KERNELS) && c.head_dim == 256 && c.n_kv_heads == 1 &&
           c.n_heads * c.horizon <= QROWS && (c.s_vlm + c.cond_steps + c.horizon + KT - 1) / KT <= 8;
}

int launch_denoise_mega(const pz_config &c, const pz_weights &w, const pz_mix_layer *layers, const MegaBuffers &bf,
                        int B, cudaStream_t st, const char **err) {
    static PerDeviceOnce attr_once;
    const int num_sms = device_sm_count();
    if (attr_once.need()) {
        if (cudaFuncSetAttribute(denoise_mega_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) != cudaSuccess ||
            cudaFuncSetAttribute(denoise_mega_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) != cudaSuccess) {
            if (err) *err = "denoise_mega: cannot set the shared-memory size";
            return PZ_ERR_CUDA;
        }
    }
    MegaParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = c.horizon; p.M = B * c.horizon; p.A = c.act_hidden; p.AI = c.act_inter; p.nh = c.n_heads;
    p.S_v = c.s_vlm; p.S_p = c.cond_steps; p.S_c = c.s_vlm + c.cond_steps; p.n_layers = c.n_layers; p.n_steps = c.n_steps;
    p.action_dim = c.action_dim; p.skp = w.small_k_pad;
    p.n_splits = (p.S_c + p.H + KT - 1) / KT;
    p.dt = (float)(1.0 / c.n_steps); p.clip = c.clip;
    for (int l = 0; l < c.n_layers; ++l) p.layers[l] = layers[l];
    p.final_norm = w.action_final_norm;
    p.enc_w1 = (const bf16 *)w.enc_w1; p.enc_w2a = (const bf16 *)w.enc_w2a; p.enc_w3 = (const bf16 *)w.enc_w3;
    p.dec_w = (const bf16 *)w.dec_w;
    p.enc_b1 = w.enc_b1; p.enc_time_bias = w.enc_time_bias; p.enc_b3 = w.enc_b3; p.dec_b = w.dec_b;
    p.rope_cos = w.rope_act_cos; p.rope_sin = w.rope_act_sin;
    p.kcache = (const bf16 *)bf.kcache; p.vcache = (const bf16 *)bf.vcache;
    p.kv_batch_stride = (long)p.S_c * 256; p.kv_layer_stride = (long)bf.batch_total * p.kv_batch_stride;
    p.valid_len = bf.valid_len;
    p.act = bf.act; p.xa = bf.xa; p.partials = bf.partials; p.out = bf.out;
    p.e1 = (bf16 *)bf.e1; p.z = (bf16 *)bf.z; p.qkv = (bf16 *)bf.qkv; p.mlp = (bf16 *)bf.mlp;
    p.barrier = bf.barrier;
    if (cudaMemsetAsync(bf.barrier, 0, 32 * sizeof(unsigned int), st) != cudaSuccess) {
        if (err) *err = "denoise_mega: memset failed";
        return PZ_ERR_CUDA;
    }
    void *args[] = {&p};
    const void *fn = p.M <= 8 ? (const void *)denoise_mega_kernel<1> : (const void *)denoise_mega_kernel<2>;
    // cooperative launch: all CTAs are guaranteed co-resident (the grid barrier relies on it)
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(num_sms), dim3(NT), args, (size_t)SMEM_TOTAL, st);
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        return PZ_ERR_CUDA;
    }
    count_launch();
    return 0;
}
