// denoise_mega3.cu -- the Euler sampler (pizero.py:454-489) for B * horizon <= 8 action rows as ONE persistent
// kernel without grid barriers: the bs = 1 latency path.
//
// What the first persistent sampler (denoise_mega.cu) spends its time on is not HBM but five grid-wide phase
// boundaries per layer (barrier + reductions + restaging, ~3.2 us each, profiles/r01_probe_barrier*.txt) and
// per-item reductions.  This kernel is organised around what tools/probe_ll.cu measured on B200:
//   * one thread per CTA feeding an mbarrier ring with contiguous 32 KB cp.async.bulk copies streams at the full HBM
//     rate (6.7-7.2 TB/s over 148 CTAs; 16 KB copies: 5.1, 8 KB: 2.5) -- so the action-expert weights are re-packed
//     once (pz_sampler_pack) into one contiguous stream per CTA, in the order that CTA consumes them and in the
//     register-fragment order of mma.sync (no address arithmetic, no bank conflicts, nothing to transpose);
//   * exchanging values through 64-bit {payload, sequence} words needs no barrier and no fence, but polling the
//     whole payload from every CTA saturates L2 (148 x 32 KB per round): gathers re-read only the words that are
//     still missing.
// Structure:
//   * CTAs [0, G): weight streaming.  Per layer five stages, each = gather inputs -> all of the CTA's items of that
//     stage accumulated in registers by 8 warps that split k -> one cross-warp reduction -> publish:
//       QKV (16-row rotary-pair blocks, RoPE in the epilogue) | o_proj (8-row blocks, full K) | gate-up (16-row blocks:
//       8 gate + 8 up rows, GeGLU in the epilogue) | down (8-row blocks, full K = 4096).
//     Every output element has one producer (no split-K across CTAs, no atomics: results are deterministic).  The
//     owner of an 8-column block of the residual stream keeps it in fp32 in shared memory for the whole call.
//   * CTAs [G, G + 16 B): attention, one per (sample, head, half of head_dim).  The layer's cached K (all 256 dims)
//     and its half of V live in shared memory as 128-byte-swizzled tiles loaded by TMA one layer ahead (the prefix
//     KV is step-invariant; L2 evict-last), so attention is q arrival -> S^T = K q^T (keys are the MMA M dimension)
//     -> soft-cap / mask / exp -> O^T = V^T P^T -> publish; no split-key partials, no combine.
//   * every wait is bounded: a lost CTA raises an error flag and the output becomes NaN instead of hanging the GPU.
// Numerics are those of denoise_mega.cu: bf16 operands, fp32 accumulation, fp32 residual stream and action state.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace {

constexpr int NCW = 8, NCT = NCW * 32, NT3 = NCT;        // 8 warps; one elected thread also feeds the weight ring
constexpr int SLOT = 32768;                              // ring slot = one weight item
constexpr int KI = 1024;                                 // act_hidden
constexpr int MAXQ = 3, MAXO = 2, MAXGU = 5, MAXJ = 5;   // items of one stage per CTA
constexpr int MAXKS = (MAXGU + 1) / 2;                   // 16-wide k steps of a CTA's down-projection partial
constexpr int HS_LD = 16 * MAXKS + 8;                    // row stride of the CTA's GeGLU outputs (bf16)
constexpr int MAXP = 160;                                // bound on NP for sizing the exchange workspace
constexpr int MAX_LAYERS = 24;
constexpr int KEYS = 288, KBOX = 144;                    // key rows per attention CTA (two TMA boxes)
constexpr int KT_BYTES = KEYS * 128;                     // one swizzled tile: KEYS rows x 64 dims
constexpr int NATT = 16;                                 // attention CTAs per sample: 8 heads x 2 halves of head_dim
constexpr unsigned long long WAIT_LIMIT_NS = 250ull * 1000 * 1000;   // any single wait: 0.25 s, then the error flag

struct CtaSched {           // what one weight-streaming CTA does in every layer / step (80 bytes)
    int n_qkv, qkv_blk[MAXQ];      // 16-row rotary-pair blocks of the fused q|k|v projection
    int n_o, o_blk[MAXO];          // 8-column blocks of the residual stream this CTA owns (o_proj, down, linear_3 rows)
    int n_gu, gu_tile[MAXGU];      // 8-column tiles of the MLP intermediate (8 gate + 8 up rows)
    int e2_blk;                    // 16-row block of action_encoder.linear_2, or -1
    int is_dec;                    // runs the final norm + action_decoder + Euler update
    int slots_per_step;
    int pid;                       // index among the CTAs that hold gate|up tiles (producers of down partials), or -1
    long long stream_off;          // byte offset of this CTA's item stream
};
static_assert(sizeof(CtaSched) == 80, "CtaSched layout");

enum SlotKind { SK_QKV = 0, SK_GU, SK_O, SK_DP, SK_E2, SK_E3, SK_DEC };
struct SlotDesc { int kind, layer, blk, aux; };

struct Mega3Params {
    int B, H, M, nh, S_v, S_p, S_c, n_layers, n_steps, action_dim, skp, AI;
    int G, NA;
    int NP;                           // CTAs that produce down-projection partials
    int pf_dist;                      // ring items ahead of the copy that are prefetched into L2 (0: none)
    int sentinel;                     // 1: one thread polls one word of an exchange before the CTA gathers all of it
    float dt, clip;
    const float *norm_in[MAX_LAYERS], *norm_post[MAX_LAYERS];
    const float *final_norm;
    const bf16 *enc_w1;
    const float *enc_b1, *enc_time_bias, *enc_b3, *dec_b;
    const float *rope_cos, *rope_sin;
    const int32_t *valid_len;
    const float *noise;
    float *out;
    const uint8_t *stream;
    const CtaSched *sched;
    int batch_total;
    // exchange buffers: 64-bit words {payload, sequence number}, zeroed before every launch
    unsigned long long *ll_act;       // [M][8]       fp32
    unsigned long long *ll_z;         // [M][A/2]     bf16x2
    unsigned long long *ll_x[2];      // [M][A/2]     bf16x2: x (1 + w) for the norm that reads it; residual stream entering layer l
    unsigned long long *ll_x1[2];     // [M][A/2]     bf16x2: the same after o_proj
    unsigned long long *ll_sx[2];     // [A/8][MAXM]  fp32: sum of squares of x over one 8-column block (exact fp32 x)
    unsigned long long *ll_sx1[2];
    unsigned long long *ll_qkv[2];    // [M][1280]    bf16x2, q and k rotated
    unsigned long long *ll_att[2];    // [M][1024]    bf16x2 attention output
    unsigned long long *ll_rs[2];     // [A/8][NP][8][MAXM] fp32: down-projection partials of one producer for one 8-column block
    unsigned int *err;
};

// ------------------------------------------------------------------------------------ PTX helpers ----
PZ_DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
PZ_DEVINL void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
PZ_DEVINL void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
PZ_DEVINL void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
PZ_DEVINL bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
PZ_DEVINL bool err_set(const Mega3Params &p) { return *reinterpret_cast<volatile unsigned int *>(p.err) != 0; }
PZ_DEVINL void raise_err(const Mega3Params &p, unsigned int kind) { atomicCAS(p.err, 0u, 0x80000000u | (kind << 16) | blockIdx.x); }
PZ_DEVINL unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// per-CTA time accounting over the whole call (thread 0, SM cycles), compiled in only with -DPZ_MEGA_TRACE
// (tools/mega3_trace.py): T3(i) charges the time since the previous stamp to bucket i; waits for the weight ring inside
// a GEMV are charged to bucket (current stage + 12)
#ifdef PZ_MEGA_TRACE
__device__ unsigned long long g3_trace[160 * 32];
__shared__ unsigned long long t3_acc[32];
__shared__ unsigned long long t3_last;
__shared__ int t3_cur;
#define T3(idx) do { if (threadIdx.x == 0) { const unsigned long long now_ = clock64(); t3_acc[(idx)] += now_ - t3_last; t3_last = now_; } } while (0)
#define T3CUR(idx) do { if (threadIdx.x == 0) t3_cur = (idx); } while (0)
#define T3WAIT_BEGIN() T3(t3_cur)
#define T3WAIT_END() T3(t3_cur + 12)
#define T3INIT() do { if (threadIdx.x == 0) { for (int i_ = 0; i_ < 32; ++i_) t3_acc[i_] = 0; t3_last = clock64(); t3_cur = 31; } } while (0)
#define T3FLUSH() do { if (threadIdx.x == 0) for (int i_ = 0; i_ < 32; ++i_) g3_trace[blockIdx.x * 32 + i_] = t3_acc[i_]; } while (0)
#else
#define T3(idx) do { } while (0)
#define T3CUR(idx) do { } while (0)
#define T3WAIT_BEGIN() do { } while (0)
#define T3WAIT_END() do { } while (0)
#define T3INIT() do { } while (0)
#define T3FLUSH() do { } while (0)
#endif
// bounded (a bug or a lost CTA must never hang the device): after WAIT_LIMIT_NS, or as soon as somebody else has raised
// the error flag, give up
PZ_DEVINL void mbar_wait(const Mega3Params &p, uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try(bar, parity)) {
        if ((++spins & 0x3F) == 0) {
            if (err_set(p)) return;
            const unsigned long long now = gtime_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > WAIT_LIMIT_NS) { raise_err(p, 1); return; }
        }
    }
}
// wait without the error-flag plumbing of the persistent kernel (bounded: a lost TMA must not hang the device)
PZ_DEVINL void mbar_wait_plain(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) if (++spins > (1u << 26)) break;
}
PZ_DEVINL void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
PZ_DEVINL void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
PZ_DEVINL void tma_load_3d_hint(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy) : "memory");
}
PZ_DEVINL uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
PZ_DEVINL uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
PZ_DEVINL void bar_compute() { __syncthreads(); }
PZ_DEVINL void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
PZ_DEVINL void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
PZ_DEVINL float gelu_fast(float x) {   // hardware tanh: rel. error 2^-11, below the bf16 rounding of the output
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(k0 * (x + k1 * x * x * x)));
    return 0.5f * x * (1.0f + y);
}
// c ? a : 0 as a select: the caller's `a` is always evaluated (straight-line code for the scheduler to interleave)
PZ_DEVINL float sel_or_zero(bool c, float a) {
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tselp.f32 %0, %1, 0f00000000, p;\n\t}" : "=f"(r) : "f"(a), "r"((uint32_t)c));
    return r;
}
PZ_DEVINL float tanh_fast_acc(float y) { float t = __expf(2.f * y); return 1.f - __fdividef(2.f, t + 1.f); }

// ---- exchange words ------------------------------------------------------------------------------------
PZ_DEVINL void ll_store(unsigned long long *dst, uint32_t payload, uint32_t seq) {
    unsigned long long v = ((unsigned long long)seq << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
}
PZ_DEVINL void ll_store2(unsigned long long *dst, uint32_t pay0, uint32_t pay1, uint32_t seq) {   // 16-byte aligned pair
    const unsigned long long a = ((unsigned long long)seq << 32) | pay0, b = ((unsigned long long)seq << 32) | pay1;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(a), "l"(b) : "memory");
}
PZ_DEVINL void ll_load2(const unsigned long long *src, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
}
// N double-words (two exchange words each) per thread, all in flight; only the ones whose sequence numbers are not
// there yet are read again (polling whole payloads from every CTA saturates L2: tools/probe_ll.cu)
template <int N, typename AddrFn>
PZ_DEVINL void ll_gather(const Mega3Params &p, uint32_t seq, uint32_t pend, AddrFn addr, unsigned long long (&v)[2 * N]) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (pend) {
#pragma unroll
        for (int u = 0; u < N; ++u)
            if ((pend >> u) & 1u) ll_load2(addr(u), v[2 * u], v[2 * u + 1]);
#pragma unroll
        for (int u = 0; u < N; ++u)
            if (((pend >> u) & 1u) && (uint32_t)(v[2 * u] >> 32) == seq && (uint32_t)(v[2 * u + 1] >> 32) == seq) pend &= ~(1u << u);
        if (pend) {
            if (++spins > 16) __nanosleep(32);
            if ((spins & 0xFF) == 0) {
                if (err_set(p)) return;
                const unsigned long long now = gtime_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > WAIT_LIMIT_NS) { raise_err(p, 2); return; }
            }
        }
    }
}

// Optional first phase of a big gather: one thread waits for one word (a different one in every CTA), so that the
// other 255 threads do not hammer the lines the producers are still writing (tools/probe_ll.cu: 3.2 -> 2.8 us per stage)
PZ_DEVINL void ll_sentinel(const Mega3Params &p, const unsigned long long *buf, int nwords, uint32_t seq) {
    if (!p.sentinel) return;
    if (threadIdx.x == 0) {
        const unsigned long long *w = buf + (int)((blockIdx.x * 37u) % (unsigned)nwords);
        uint32_t spins = 0;
        unsigned long long t0 = 0, v;
        for (;;) {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
            if ((uint32_t)(v >> 32) == seq) break;
            if ((++spins & 0xFF) == 0) {
                if (err_set(p)) break;
                const unsigned long long now = gtime_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > WAIT_LIMIT_NS) { raise_err(p, 3); break; }
            }
        }
    }
    __syncthreads();
}

// sequence numbers: one per exchange and step
PZ_DEVINL uint32_t seq_of(const Mega3Params &p, int step, int idx) { return 1u + (uint32_t)(step * (3 + 5 * p.n_layers) + idx); }
constexpr int IDX_ACT = 0, IDX_Z = 1, IDX_X0 = 2;
PZ_DEVINL int IDX_QKV(int l) { return 3 + 5 * l; }
PZ_DEVINL int IDX_ATT(int l) { return 4 + 5 * l; }
PZ_DEVINL int IDX_X1(int l) { return 5 + 5 * l; }
PZ_DEVINL int IDX_MLP(int l) { return 6 + 5 * l; }
PZ_DEVINL int IDX_X2(int l) { return 7 + 5 * l; }

// ======================================== weight-streaming role =========================================
template <int MAXM> struct GemvSmem {
    static constexpr int SLOTS = MAXM <= 4 ? 6 : 5;
    static constexpr int RSTRIDE = MAXM + 1;                               // floats per row of the cross-warp reduction buffer
    static constexpr int LDA_MAX = 2048 + 32;
    static constexpr int RING = 0;
    static constexpr int AST = RING + SLOTS * SLOT;                        // bf16 [MAXM][K + 32]
    static constexpr int RED = AST + MAXM * LDA_MAX * 2;                   // float [NCW][MAXJ][16][RSTRIDE]
    static constexpr int MISC = RED + NCW * MAXJ * 16 * RSTRIDE * 4;             // floats: part[NCW][8], xpriv[MAXO][8][8], acts[64]
    static constexpr int XPRIV = 64, SACT = 64 + MAXO * 64;                // float offsets inside MISC
    static constexpr int HS = MISC + (64 + MAXO * 64 + 64) * 4;            // bf16 [8][HS_LD]: this CTA's GeGLU outputs
    static constexpr int BARS = HS + 8 * HS_LD * 2;                        // full[SLOTS]
    static constexpr int SCHED = BARS + SLOTS * 8;
    static constexpr int END = SCHED + (int)sizeof(CtaSched);
};

struct GemvCtx {
    uint8_t *smem;
    uint64_t *full;
    uint32_t cnt;        // ring items consumed so far
    uint32_t issued;     // ring items requested so far (refill thread only)
    uint32_t total;      // items of the whole call
    int in_step;         // position of the next request inside the step's stream
    const uint8_t *base;
    int slots_per_step;
    uint64_t pol;
    int pf_dist;         // L2 prefetch distance in items
    int pf_step;         // position of the next prefetch inside the step's stream
    uint32_t pf_issued;
};
constexpr int REFILL_TID = 7 * 32;   // the elected thread (a warp with little epilogue work)
// Request the items that fit into the ring slots the CTA has finished reading.  Call after a block barrier that
// follows the reads (all warps are past their last use of the consumed slots), from every thread.
template <typename SM>
PZ_DEVINL void ring_refill(GemvCtx &cx) {
    if (threadIdx.x != REFILL_TID) return;
    if (cx.pf_dist > 0) {   // HBM -> L2 runs pf_dist items ahead of the copies, which then hit L2
        const uint32_t want = cx.cnt + SM::SLOTS + (uint32_t)cx.pf_dist;
        while (cx.pf_issued < want && cx.pf_issued < cx.total) {
            bulk_prefetch_l2(cx.base + (long)cx.pf_step * SLOT, SLOT);
            if (++cx.pf_step == cx.slots_per_step) cx.pf_step = 0;
            ++cx.pf_issued;
        }
    }
    while (cx.issued < cx.cnt + SM::SLOTS && cx.issued < cx.total) {
        const int slot = cx.issued % SM::SLOTS;
        mbar_expect_tx(&cx.full[slot], SLOT);
        bulk_g2s(cx.smem + SM::RING + slot * SLOT, cx.base + (long)cx.in_step * SLOT, SLOT, &cx.full[slot], cx.pol);
        if (++cx.in_step == cx.slots_per_step) cx.in_step = 0;
        ++cx.issued;
    }
}

template <typename SM>
PZ_DEVINL float *red_ptr(uint8_t *red, int w, int j, int r) { return reinterpret_cast<float *>(red) + ((w * MAXJ + j) * 16 + r) * SM::RSTRIDE; }

// acc[j] = (16-row item j of this stage) . A^T over this warp's 128-wide k slice; K = 1024
// MID > 0: the slots of the first MID items are handed back (and re-requested) as soon as they are consumed, not after the stage
template <typename SM, int NJ, int MID = 0>
PZ_DEVINL void gemv16(const Mega3Params &p, GemvCtx &cx, int n, float (&acc)[NJ][4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bf16 *As = reinterpret_cast<const bf16 *>(cx.smem + SM::AST);
    constexpr int lda = KI + 32;
    uint4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
        x[u] = g < p.M ? *reinterpret_cast<const uint4 *>(As + g * lda + warp * 128 + u * 32 + 8 * t) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        if (j < n) {
            const int slot = cx.cnt % SM::SLOTS;
            T3WAIT_BEGIN();
            mbar_wait(p, &cx.full[slot], (cx.cnt / SM::SLOTS) & 1);
            T3WAIT_END();
            const uint8_t *w = cx.smem + SM::RING + slot * SLOT + warp * 4096 + lane * 16;
            float a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint4 lo = *reinterpret_cast<const uint4 *>(w + u * 1024);
                const uint4 hi = *reinterpret_cast<const uint4 *>(w + u * 1024 + 512);
                mma_bf16(acc[j], lo.x, hi.x, lo.y, hi.y, x[u].x, x[u].y);
                mma_bf16(a2, lo.z, hi.z, lo.w, hi.w, x[u].z, x[u].w);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] += a2[e];
            cx.cnt += 1;
            if ((NJ > SM::SLOTS && j + 1 == SM::SLOTS && n > SM::SLOTS) ||   // more items than ring slots (CTA-uniform)
                (MID > 0 && j + 1 == MID && n > MID)) {
                __syncthreads();
                    ring_refill<SM>(cx);
            }
        }
    }
}
// 8-row items with the full K (K = 256 * KU: 1024 / 2048 / 4096; 4096 spans two ring slots)
template <typename SM, int KU>
PZ_DEVINL void gemv8(const Mega3Params &p, GemvCtx &cx, int n, float (&acc)[MAXO][4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bf16 *As = reinterpret_cast<const bf16 *>(cx.smem + SM::AST);
    constexpr int K = 256 * KU, lda = K + 32;
    constexpr int NS = KU == 16 ? 2 : 1;
#pragma unroll
    for (int j = 0; j < MAXO; ++j) {
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        if (j < n) {
            const int mine = NS == 2 ? (warp >> 2) : 0;
            const uint32_t c = cx.cnt + mine;
            const int slot = c % SM::SLOTS;
            T3WAIT_BEGIN();
            mbar_wait(p, &cx.full[slot], (c / SM::SLOTS) & 1);
            T3WAIT_END();
            const int unit0 = warp * KU - mine * 64;   // first unit of this warp inside its slot
            const uint8_t *w = cx.smem + SM::RING + slot * SLOT + unit0 * 512 + lane * 16;
            const bf16 *xa = As + g * lda + warp * (KU * 32) + 8 * t;
            float a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
            for (int u = 0; u < KU; ++u) {
                const uint4 a = *reinterpret_cast<const uint4 *>(w + u * 512);
                const uint4 x = g < p.M ? *reinterpret_cast<const uint4 *>(xa + u * 32) : make_uint4(0, 0, 0, 0);
                mma_bf16(acc[j], a.x, 0u, a.y, 0u, x.x, x.y);
                mma_bf16(a2, a.z, 0u, a.w, 0u, x.z, x.w);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] += a2[e];
            cx.cnt += NS;
        }
    }
}
template <typename SM, int NJ>
PZ_DEVINL void red_write(uint8_t *red, int n, const float (&acc)[NJ][4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    if (2 * t >= SM::RSTRIDE - 1) return;   // token columns beyond the kernel's row count are padding
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if (j < n) {
            float *r0 = red_ptr<SM>(red, warp, j, g), *r1 = red_ptr<SM>(red, warp, j, g + 8);
            r0[2 * t] = acc[j][0]; r0[2 * t + 1] = acc[j][1];
            r1[2 * t] = acc[j][2]; r1[2 * t + 1] = acc[j][3];
        }
    }
}
template <typename SM>
PZ_DEVINL float red_sum(uint8_t *red, int j, int r, int m) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) v += red_ptr<SM>(red, w, j, r)[m];
    return v;
}

// This CTA's share of the down projection (paligemma/modules.py:95): out[1024][m] += W_down[:, my columns] h[my columns][m],
// K = this CTA's (up to) 8 MAXGU GeGLU outputs, zero-padded to 16-wide steps.  Ring slot ks = k step ks in fragment order
// [m tile 0..63][lane][a0 a1 a2 a3] (mega3_repack_kernel); warp w owns rows 128 w .. 128 w + 127, so nothing is reduced
// across warps: acc[i] = rows 128 w + 16 i + {g, g + 8}, tokens 2t, 2t + 1.
template <typename SM>
PZ_DEVINL void down_partial(const Mega3Params &p, GemvCtx &cx, int n_ks, float (&acc)[8][4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bf16 *hs = reinterpret_cast<const bf16 *>(cx.smem + SM::HS);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < MAXKS; ++ks) {
        if (ks < n_ks) {
            const int slot = cx.cnt % SM::SLOTS;
            T3WAIT_BEGIN();
            mbar_wait(p, &cx.full[slot], (cx.cnt / SM::SLOTS) & 1);
            T3WAIT_END();
            const uint32_t b0 = g < p.M ? *reinterpret_cast<const uint32_t *>(hs + g * HS_LD + ks * 16 + 2 * t) : 0u;
            const uint32_t b1 = g < p.M ? *reinterpret_cast<const uint32_t *>(hs + g * HS_LD + ks * 16 + 8 + 2 * t) : 0u;
            const uint8_t *w = cx.smem + SM::RING + slot * SLOT + warp * 4096 + lane * 16;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 a = *reinterpret_cast<const uint4 *>(w + i * 512);
                mma_bf16(acc[i], a.x, a.y, a.z, a.w, b0, b1);
            }
            cx.cnt += 1;
        }
    }
}

// ---- activation staging ------------------------------------------------------------------------------
// Residual stream -> As.  The block owners publish bf16 x (1 + w) (w = the weight of the RMSNorm that reads it,
// paligemma/modules.py:13-21) and the fp32 sum of squares of their 8 columns; the per-row factor rsqrt(mean x^2 + eps)
// commutes with the projection, so it is applied to the fp32 accumulators in the epilogue (row_rnorm) and nothing but
// the copy sits between the exchange and the MMAs.  part[warp][m] <- this warp's share of sum x^2.
template <typename SM, int MAXM>
PZ_DEVINL void stage_x(const Mega3Params &p, uint8_t *smem, const unsigned long long *pairs, const unsigned long long *ss, uint32_t seq) {
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM::AST);
    float *part = reinterpret_cast<float *>(smem + SM::MISC);
    constexpr int lda = KI + 32;
    constexpr int NP = MAXM, NS = MAXM / 4;   // double-words per thread: one per row of x; [A/8][MAXM] sums of squares
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long v[2 * (NP + NS)];
    uint32_t pend = ((1u << NS) - 1u) << NP;
#pragma unroll
    for (int u = 0; u < NP; ++u) if (u < p.M) pend |= 1u << u;
    ll_sentinel(p, pairs, p.M * (KI / 2), seq);
    ll_gather<NP + NS>(p, seq, pend, [&](int u) {
        return u < NP ? pairs + (long)u * (KI / 2) + 2 * tid : ss + 2 * ((u - NP) * NCT + tid);
    }, v);
#pragma unroll
    for (int u = 0; u < NP; ++u)
        if (u < p.M) *reinterpret_cast<uint2 *>(As + u * lda + 4 * tid) = make_uint2((uint32_t)v[2 * u], (uint32_t)v[2 * u + 1]);
    // words 2 d, 2 d + 1 of [block][MAXM] with d = s * 256 + tid: rows m0, m0 + 1 with m0 = 2 (tid % (MAXM / 2))
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int u = 0; u < NS; ++u) { s0 += __uint_as_float((uint32_t)v[2 * (NP + u)]); s1 += __uint_as_float((uint32_t)v[2 * (NP + u) + 1]); }
#pragma unroll
    for (int o = MAXM / 2; o < 32; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    if (lane < MAXM / 2) { part[warp * 8 + 2 * lane] = s0; part[warp * 8 + 2 * lane + 1] = s1; }
    bar_compute();
}
// rsqrt(mean x^2 + eps) of row m of what stage_x staged last
template <typename SM>
PZ_DEVINL float row_rnorm(uint8_t *smem, int m) {
    const float *part = reinterpret_cast<const float *>(smem + SM::MISC);
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) t += part[w * 8 + m];
    return rsqrtf(t * (1.f / KI) + 1e-6f);
}
// An owner's epilogue: one (block j, row m, column r) per thread, 8 consecutive lanes = one (j, m).  Publishes bf16 pairs of
// v (1 + wn) and the block's sum of squares of row m (zeros for the pad rows m >= M: the readers wait for whole double-words).
template <int MAXM>
PZ_DEVINL void publish_x(const Mega3Params &p, bool active, bool valid, float v, float wn, int n, int blk, int m,
                         unsigned long long *pairs, unsigned long long *ss, uint32_t seq) {
    const float y = v * (1.f + wn);
    const float y1 = __shfl_down_sync(0xffffffffu, y, 1);
    float q = v * v;
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    q += __shfl_xor_sync(0xffffffffu, q, 4);
    if (valid && !(n & 1)) ll_store(pairs + (long)m * (KI / 2) + (n >> 1), pack_bf16x2(y, y1), seq);
    if (active && !(n & 7)) ll_store(ss + blk * MAXM + m, __float_as_uint(q), seq);
}
// bf16x2 words [M][K/2] -> As[m][k] (row stride K + 32)
template <typename SM, int K>
PZ_DEVINL void stage_pairs(const Mega3Params &p, uint8_t *smem, const unsigned long long *buf, uint32_t seq) {
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM::AST);
    constexpr int lda = K + 32, DPR = K / 4;     // double-words (4 bf16) per row
    const int total = p.M * DPR;
    ll_sentinel(p, buf, 2 * total, seq);
    for (int i0 = 0; i0 < total; i0 += NCT * 8) {
        unsigned long long v[16];
        uint32_t pend = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (i0 + u * NCT + (int)threadIdx.x < total) pend |= 1u << u;
        ll_gather<8>(p, seq, pend, [&](int u) { return buf + 2 * (long)(i0 + u * NCT + threadIdx.x); }, v);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * NCT + threadIdx.x;
            if (i < total) {
                const int m = i / DPR, dw = i % DPR;
                *reinterpret_cast<uint2 *>(As + m * lda + dw * 4) = make_uint2((uint32_t)v[2 * u], (uint32_t)v[2 * u + 1]);
            }
        }
    }
    bar_compute();
}
// action words [M][8] -> linear_1 (vla/modules.py:39-41) -> As[m][0..1024)
template <typename SM>
PZ_DEVINL void stage_enc1(const Mega3Params &p, uint8_t *smem, uint32_t seq) {
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM::AST);
    float *sact = reinterpret_cast<float *>(smem + SM::MISC) + SM::SACT;   // [M][8], bf16-rounded like a GEMM input
    constexpr int lda = KI + 32;
    {
        unsigned long long v[2];
        const bool ok = (int)threadIdx.x < p.M * 4;
        ll_gather<1>(p, seq, ok ? 1u : 0u, [&](int) { return p.ll_act + 2 * threadIdx.x; }, v);
        if (ok) {
            sact[2 * threadIdx.x] = __bfloat162float(__float2bfloat16_rn(__uint_as_float((uint32_t)v[0])));
            sact[2 * threadIdx.x + 1] = __bfloat162float(__float2bfloat16_rn(__uint_as_float((uint32_t)v[1])));
        }
    }
    bar_compute();
    for (int n = threadIdx.x; n < KI; n += NCT) {
        const uint4 wv = __ldg(reinterpret_cast<const uint4 *>(p.enc_w1 + (long)n * p.skp));   // skp >= 8
        const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
        const float b1 = p.enc_b1[n];
        for (int m = 0; m < p.M; ++m) {
            float v = b1;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < p.action_dim) v += ((k & 1) ? bf16hi(ww[k >> 1]) : bf16lo(ww[k >> 1])) * sact[m * 8 + k];
            As[m * lda + n] = __float2bfloat16_rn(v);
        }
    }
    bar_compute();
}

template <int MAXM>
PZ_DEVINL void gemv_role(const Mega3Params &p, uint8_t *smem, const CtaSched &sc) {
    using SM = GemvSmem<MAXM>;
    GemvCtx cx;
    cx.smem = smem; cx.full = reinterpret_cast<uint64_t *>(smem + SM::BARS);
    cx.cnt = 0; cx.issued = 0; cx.in_step = 0;
    cx.total = (uint32_t)(p.n_steps * sc.slots_per_step);
    cx.base = p.stream + sc.stream_off; cx.slots_per_step = sc.slots_per_step;
    cx.pol = policy_evict_first();
    cx.pf_dist = p.pf_dist; cx.pf_step = 0; cx.pf_issued = 0;
    ring_refill<SM>(cx);   // start streaming
    uint8_t *red = smem + SM::RED;
    float *xpriv = reinterpret_cast<float *>(smem + SM::MISC) + SM::XPRIV;   // [MAXO][8 rows m][8 cols]
    const int tid = threadIdx.x;
    for (int i = tid; i < 8 * HS_LD / 2; i += NCT) reinterpret_cast<uint32_t *>(smem + SM::HS)[i] = 0u;   // k padding stays zero
    const int qkvw = (p.nh + 2) * 128;
    // owner epilogues: thread -> (block j, row m, column r); 8 consecutive lanes share (j, m)
    const int ow_r = tid & 7, ow_m = (tid >> 3) % MAXM, ow_j = tid / (8 * MAXM);
    const bool ow_active = ow_j < sc.n_o, ow_valid = ow_active && ow_m < p.M;
    const int ow_blk = ow_active ? sc.o_blk[ow_j] : 0, ow_n = ow_blk * 8 + ow_r;
    const bool ow_warp = (tid & ~31) < sc.n_o * 8 * MAXM;   // warp-uniform: this warp holds owner threads
    // QKV epilogue: (item, m, row) of pass i0 = tid, tid + 256: the rotary factors of q and k never change
    // (table row S_p + token: positions 2.., pizero.py:312-318)
    float rope_c[2] = {1.f, 1.f}, rope_s[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = u * NCT + tid;
        if (i < sc.n_qkv * 16 * p.M) {
            const int r = i & 15, jm = i >> 4, j = jm / p.M, m = jm % p.M;
            const int blk = sc.qkv_blk[j], hh = blk >> 4, d = (blk & 15) * 8 + (r & 7);
            if (hh <= p.nh) {
                const long ti = (long)(p.S_p + (m % p.H)) * 128 + d;
                rope_c[u] = __ldg(p.rope_cos + ti);
                rope_s[u] = __ldg(p.rope_sin + ti);
            }
        }
    }
    // the fp32 action state lives in registers of the decoder CTA: thread i < M*8 owns element (i / 8, i % 8);
    // it starts as the caller's noise (pizero.py:454-458)
    float my_act = 0.f;
    if (sc.is_dec && tid < p.M * 8) {
        const int m = tid >> 3, a = tid & 7;
        my_act = a < p.action_dim ? p.noise[m * p.action_dim + a] : 0.f;
        ll_store(p.ll_act + tid, __float_as_uint(my_act), seq_of(p, 0, IDX_ACT));
    }
    for (int step = 0; step < p.n_steps; ++step) {
        // ---- action encoder (vla/modules.py:39-53): linear_1 on the fly, linear_2 (action half) + per-step time
        //      bias + SiLU, linear_3 + sqrt(hidden) embed scale (joint_model.py:348-355)
        if (sc.e2_blk >= 0) {
            stage_enc1<SM>(p, smem, seq_of(p, step, IDX_ACT));
            float acc[1][4];
            gemv16<SM, 1>(p, cx, 1, acc);
            red_write<SM, 1>(red, 1, acc);
            bar_compute();
            ring_refill<SM>(cx);
            for (int i = tid; i < 8 * p.M; i += NCT) {   // (row pair, m)
                const int rp = i & 7, m = i >> 3, n = sc.e2_blk * 16 + 2 * rp;
                const float v0 = red_sum<SM>(red, 0, 2 * rp, m) + p.enc_time_bias[step * KI + n];
                const float v1 = red_sum<SM>(red, 0, 2 * rp + 1, m) + p.enc_time_bias[step * KI + n + 1];
                ll_store(p.ll_z + (long)m * (KI / 2) + (n >> 1), pack_bf16x2(silu(v0), silu(v1)), seq_of(p, step, IDX_Z));
            }
            bar_compute();
        }
        if (sc.n_o > 0) {
            const float wn = ow_active ? __ldg(p.norm_in[0] + ow_n) : 0.f;
            const float b3 = ow_active ? __ldg(p.enc_b3 + ow_n) : 0.f;
            stage_pairs<SM, KI>(p, smem, p.ll_z, seq_of(p, step, IDX_Z));
            float acc[MAXO][4];
            gemv8<SM, 4>(p, cx, sc.n_o, acc);
            red_write<SM, MAXO>(red, sc.n_o, acc);
            bar_compute();
            ring_refill<SM>(cx);
            if (ow_warp) {
                const float v = ow_valid ? (red_sum<SM>(red, ow_j, ow_r, ow_m) + b3) * sqrtf((float)KI) : 0.f;
                if (ow_active) xpriv[(ow_j * 8 + ow_m) * 8 + ow_r] = v;
                publish_x<MAXM>(p, ow_active, ow_valid, v, wn, ow_n, ow_blk, ow_m, p.ll_x[0], p.ll_sx[0], seq_of(p, step, IDX_X0));
            }
            bar_compute();
        }
        for (int l = 0; l < p.n_layers; ++l) {
            const int pb = l & 1;
            T3(0);
            // ---- QKV: x -> RMSNorm -> fused q|k|v projection (mixture.py:187-215), RoPE on q and k
            if (sc.n_qkv > 0) {
                stage_x<SM, MAXM>(p, smem, p.ll_x[pb], p.ll_sx[pb], seq_of(p, step, l == 0 ? IDX_X0 : IDX_X2(l - 1)));
                T3(1);
                T3CUR(2);
                float acc[MAXQ][4];
                gemv16<SM, MAXQ>(p, cx, sc.n_qkv, acc);
                red_write<SM, MAXQ>(red, sc.n_qkv, acc);
                bar_compute();
                T3(2);
                ring_refill<SM>(cx);
                const uint32_t fo = seq_of(p, step, IDX_QKV(l));
                // one (item, m, row) sum per thread.  Rows 0-7 / 8-15 of an item are the dims d / d + 128 of one head:
                // rotate q and k here (fp32), then neighbouring dims pair up through a shuffle
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = u * NCT + tid;
                    if (u * NCT >= sc.n_qkv * 16 * p.M) break;   // CTA-uniform
                    const bool ok = i < sc.n_qkv * 16 * p.M;
                    const int r = i & 15, jm = ok ? (i >> 4) : 0, j = jm / p.M, m = jm % p.M;
                    float v0 = ok ? red_sum<SM>(red, j, r, m) * row_rnorm<SM>(smem, m) : 0.f;
                    const float other = __shfl_xor_sync(0xffffffffu, v0, 8);
                    const int blk = sc.qkv_blk[j], hh = blk >> 4, d = (blk & 15) * 8 + (r & 7);
                    v0 = (r < 8) ? v0 * rope_c[u] - other * rope_s[u] : v0 * rope_c[u] + other * rope_s[u];
                    const float v1 = __shfl_down_sync(0xffffffffu, v0, 1);
                    const int n = hh * 256 + (r < 8 ? d : 128 + d);
                    if (ok && !(r & 1)) ll_store(p.ll_qkv[pb] + (long)m * qkvw + (n >> 1), pack_bf16x2(v0, v1), fo);
                }
                bar_compute();
                T3(3);
            }
            // ---- o_proj + residual (mixture.py:217-218, joint_model.py:65-75)
            if (sc.n_o > 0) {
                const float wn = ow_active ? __ldg(p.norm_post[l] + ow_n) : 0.f;
                stage_pairs<SM, 2048>(p, smem, p.ll_att[pb], seq_of(p, step, IDX_ATT(l)));
                T3(4);
                T3CUR(5);
                float acc[MAXO][4];
                gemv8<SM, 8>(p, cx, sc.n_o, acc);
                red_write<SM, MAXO>(red, sc.n_o, acc);
                bar_compute();
                T3(5);
                ring_refill<SM>(cx);
                if (ow_warp) {
                    float v = 0.f;
                    if (ow_valid) {
                        v = xpriv[(ow_j * 8 + ow_m) * 8 + ow_r] + red_sum<SM>(red, ow_j, ow_r, ow_m);
                        xpriv[(ow_j * 8 + ow_m) * 8 + ow_r] = v;
                    }
                    publish_x<MAXM>(p, ow_active, ow_valid, v, wn, ow_n, ow_blk, ow_m, p.ll_x1[pb], p.ll_sx1[pb], seq_of(p, step, IDX_X1(l)));
                }
                bar_compute();
                T3(6);
            }
            // ---- gate|up + GeGLU (paligemma/modules.py:86-95)
            if (sc.n_gu > 0) {
                stage_x<SM, MAXM>(p, smem, p.ll_x1[pb], p.ll_sx1[pb], seq_of(p, step, IDX_X1(l)));
                T3(7);
                T3CUR(8);
                float acc[MAXGU][4];
                gemv16<SM, MAXGU, 2>(p, cx, sc.n_gu, acc);
                red_write<SM, MAXGU>(red, sc.n_gu, acc);
                bar_compute();
                T3(8);
                ring_refill<SM>(cx);
                bf16 *hs = reinterpret_cast<bf16 *>(smem + SM::HS);
                for (int i0 = 0; i0 < sc.n_gu * 16 * p.M; i0 += NCT) {
                    const int i = i0 + tid;
                    const bool ok = i < sc.n_gu * 16 * p.M;
                    const int r = i & 15, jm = ok ? (i >> 4) : 0, j = jm / p.M, m = jm % p.M;
                    const float v = ok ? red_sum<SM>(red, j, r, m) * row_rnorm<SM>(smem, m) : 0.f;   // rows 0-7: gate, 8-15: the matching up rows
                    const float u = __shfl_down_sync(0xffffffffu, v, 8);
                    if (ok && r < 8) hs[m * HS_LD + j * 8 + r] = __float2bfloat16_rn(gelu_fast(v) * u);
                }
                bar_compute();
                T3(9);
                T3CUR(16);
                // ---- down projection of this CTA's columns; the fp32 partials go to the owners of the residual blocks
                float dacc[8][4];
                down_partial<SM>(p, cx, (sc.n_gu + 1) >> 1, dacc);
                T3(16);
                if (2 * (tid & 3) < MAXM) {
                    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
                    const uint32_t fo = seq_of(p, step, IDX_MLP(l));
                    unsigned long long *dst = p.ll_rs[pb] + ((long)(warp * 16) * p.NP + sc.pid) * (8 * MAXM) + g * MAXM + 2 * t;
                    const long bstride = (long)p.NP * (8 * MAXM);   // one residual block further
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        ll_store2(dst + (2 * i) * bstride, __float_as_uint(dacc[i][0]), __float_as_uint(dacc[i][1]), fo);
                        ll_store2(dst + (2 * i + 1) * bstride, __float_as_uint(dacc[i][2]), __float_as_uint(dacc[i][3]), fo);
                    }
                }
                bar_compute();
                ring_refill<SM>(cx);
                T3(18);
            }
            // ---- owners: sum the partials of their blocks in producer order (deterministic) + residual
            if (sc.n_o > 0) {
                const float wn = ow_active ? __ldg((l + 1 < p.n_layers ? p.norm_in[l + 1] : p.final_norm) + ow_n) : 0.f;
                constexpr int E = 8 * MAXM, E2 = E / 2, PPU = NCT / E2;   // words / double-words per producer; producers per round
                float *rsum = reinterpret_cast<float *>(red);             // [NCW][MAXO][E]
                const uint32_t fin = seq_of(p, step, IDX_MLP(l));
                const int rounds = (p.NP + PPU - 1) / PPU;
                for (int j = 0; j < sc.n_o; ++j) {
                    const unsigned long long *base = p.ll_rs[pb] + (long)sc.o_blk[j] * p.NP * E;
                    ll_sentinel(p, base, p.NP * E, fin);
                    float s0 = 0.f, s1 = 0.f;
                    for (int u0 = 0; u0 < rounds; u0 += 8) {
                        unsigned long long v[16];
                        uint32_t pend = 0;
#pragma unroll
                        for (int u = 0; u < 8; ++u) if ((u0 + u) * PPU + tid / E2 < p.NP) pend |= 1u << u;
                        ll_gather<8>(p, fin, pend, [&](int u) { return base + 2 * ((long)(u0 + u) * NCT + tid); }, v);
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if ((pend >> u) & 1u) { s0 += __uint_as_float((uint32_t)v[2 * u]); s1 += __uint_as_float((uint32_t)v[2 * u + 1]); }
                    }
                    if (E2 < 32) { s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16); }
                    const int lane = tid & 31, warp = tid >> 5;
                    if (lane < E2) { rsum[(warp * MAXO + j) * E + 2 * lane] = s0; rsum[(warp * MAXO + j) * E + 2 * lane + 1] = s1; }
                }
                bar_compute();
                T3(10);
                if (ow_warp) {
                    float v = 0.f;
                    if (ow_valid) {
                        v = xpriv[(ow_j * 8 + ow_m) * 8 + ow_r];
#pragma unroll
                        for (int w = 0; w < NCW; ++w) v += rsum[(w * MAXO + ow_j) * E + ow_r * MAXM + ow_m];
                        xpriv[(ow_j * 8 + ow_m) * 8 + ow_r] = v;
                    }
                    publish_x<MAXM>(p, ow_active, ow_valid, v, wn, ow_n, ow_blk, ow_m, p.ll_x[(l + 1) & 1], p.ll_sx[(l + 1) & 1],
                                    seq_of(p, step, IDX_X2(l)));
                }
                bar_compute();
                T3(12);
            }
        }
        // ---- final norm + decoder + Euler update (joint_model.py:375-380, pizero.py:479-481)
        if (sc.is_dec) {
            stage_x<SM, MAXM>(p, smem, p.ll_x[p.n_layers & 1], p.ll_sx[p.n_layers & 1], seq_of(p, step, IDX_X2(p.n_layers - 1)));
            float acc[MAXO][4];
            gemv8<SM, 4>(p, cx, 1, acc);
            red_write<SM, MAXO>(red, 1, acc);
            bar_compute();
            ring_refill<SM>(cx);
            if (tid < 8 * p.M) {
                const int a = tid & 7, m = tid >> 3;
                if (a < p.action_dim) my_act += p.dt * (red_sum<SM>(red, 0, a, m) * row_rnorm<SM>(smem, m) + p.dec_b[a]);
                if (step + 1 < p.n_steps) {
                    ll_store(p.ll_act + tid, __float_as_uint(my_act), seq_of(p, step + 1, IDX_ACT));
                } else if (a < p.action_dim) {
                    float v = my_act;
                    if (p.clip >= 0.f) v = fminf(fmaxf(v, -p.clip), p.clip);
                    if (err_set(p)) v = __int_as_float(0x7fc00000);   // a wait ran into its bound: fail loudly
                    p.out[m * p.action_dim + a] = v;
                }
            }
            bar_compute();
        }
    }
}

// ============================================ attention role ============================================
// shared memory: K tiles [4 dim-quarters][KEYS rows][64 dims] and V tiles [2][KEYS][64], 128-byte swizzled (TMA),
// q rows, P, row-sum partials
struct AttSmem {
    static constexpr int K = 0;
    static constexpr int V = K + 4 * KT_BYTES;
    static constexpr int Q = V + 2 * KT_BYTES;             // bf16 [8][264]
    static constexpr int P = Q + 8 * 264 * 2;              // bf16 [8][KEYS + 8]
    static constexpr int LS = P + 8 * (KEYS + 8) * 2;      // float [NCW][8]
    static constexpr int BARS = LS + NCW * 8 * 4;          // kv_full
    static constexpr int END = BARS + 16;
};
PZ_DEVINL uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// one thread: request the cached K (all 256 dims) and this CTA's half of V of (layer, sample b): 12 boxes of 144 keys x 64 dims
PZ_DEVINL void att_request_kv(const Mega3Params &p, uint8_t *smem, const CUtensorMap *kmap, const CUtensorMap *vmap, int l, int b,
                              int dh, uint64_t pol) {
    uint64_t *kv_full = reinterpret_cast<uint64_t *>(smem + AttSmem::BARS);
    mbar_expect_tx(kv_full, 6u * KT_BYTES);
    const int slab = l * p.batch_total + b;
#pragma unroll
    for (int dq = 0; dq < 4; ++dq)
#pragma unroll
        for (int hb = 0; hb < 2; ++hb)
            tma_load_3d_hint(kmap, kv_full, smem + AttSmem::K + dq * KT_BYTES + hb * KBOX * 128, dq * 64, hb * KBOX, slab, pol);
#pragma unroll
    for (int dq = 0; dq < 2; ++dq)
#pragma unroll
        for (int hb = 0; hb < 2; ++hb)
            tma_load_3d_hint(vmap, kv_full, smem + AttSmem::V + dq * KT_BYTES + hb * KBOX * 128, (dh * 2 + dq) * 64, hb * KBOX, slab, pol);
}

PZ_DEVINL void att_role(const Mega3Params &p, uint8_t *smem, const CUtensorMap *kmap, const CUtensorMap *vmap, int b, int head, int dh) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    bf16 *sQ = reinterpret_cast<bf16 *>(smem + AttSmem::Q);
    bf16 *sP = reinterpret_cast<bf16 *>(smem + AttSmem::P);
    float *sLS = reinterpret_cast<float *>(smem + AttSmem::LS);
    uint64_t *kv_full = reinterpret_cast<uint64_t *>(smem + AttSmem::BARS);
    constexpr int LDQ = 264, LDP = KEYS + 8;
    const uint64_t pol = policy_evict_last();
    if (tid == 0) att_request_kv(p, smem, kmap, vmap, 0, b, dh, pol);   // layer 0 of step 0
    const int H = p.H, S_c = p.S_c, n_keys = S_c + H;
    const int vlen = p.valid_len[b];
    const int n_mt = (n_keys + 15) >> 4;               // 16-key tiles
    const int qkvw = (p.nh + 2) * 128;
    const uint32_t sK = smem_u32(smem + AttSmem::K), sV = smem_u32(smem + AttSmem::V);
    // rows no exchange ever writes must be zero: P x (anything finite) stays 0, q pad rows contribute nothing
    for (int i = tid; i < 8 * LDQ / 2; i += NCT) reinterpret_cast<uint32_t *>(sQ)[i] = 0u;
    for (int i = tid; i < 8 * LDP / 2; i += NCT) reinterpret_cast<uint32_t *>(sP)[i] = 0u;
    bar_compute();

    // what this thread fetches from the q|k|v exchange in every layer and where it goes (layer-invariant)
    uint32_t att_pend = 0, src_off[3] = {0, 0, 0}, dst_off[3] = {0, 0, 0};
    {
        const int per_tok = 160, total = H * per_tok;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int i = tid + u * NCT;
            if (i >= total) continue;
            att_pend |= 1u << u;
            const int tok = i / per_tok, e = i % per_tok, row = S_c + tok;
            const uint32_t rbase = (uint32_t)((b * H + tok) * qkvw);
            if (e < 64) {                      // q: dims 4e .. 4e+3
                src_off[u] = rbase + head * 128 + 2 * e;
                dst_off[u] = AttSmem::Q + (tok * LDQ + 4 * e) * 2;
            } else if (e < 128) {              // k
                const int d = 4 * (e - 64);
                src_off[u] = rbase + p.nh * 128 + 2 * (e - 64);
                dst_off[u] = AttSmem::K + (d >> 6) * KT_BYTES + swz(row, (d & 63) >> 3) + (d & 7) * 2;
            } else {                           // v, dims dh*128 + 4(e-128) ..
                const int d = 4 * (e - 128);
                src_off[u] = rbase + (p.nh + 1) * 128 + dh * 64 + 2 * (e - 128);
                dst_off[u] = AttSmem::V + (d >> 6) * KT_BYTES + swz(row, (d & 63) >> 3) + (d & 7) * 2;
            }
        }
    }
    uint32_t it = 0;
    for (int step = 0; step < p.n_steps; ++step) {
        for (int l = 0; l < p.n_layers; ++l, ++it) {
            const int pb = l & 1;
            const uint32_t fin = seq_of(p, step, IDX_QKV(l));
            T3(20);
            // ---- this step's rotated q rows of (sample b, head), the fresh k rows and this CTA's half of the fresh
            //      v rows: per token 64 + 64 + 32 double-words (source offsets / destinations: src_off, dst_off above)
            unsigned long long v[6];
            ll_gather<3>(p, fin, att_pend, [&](int u) { return p.ll_qkv[pb] + src_off[u]; }, v);
            T3(21);
            mbar_wait(p, kv_full, it & 1);
            T3(26);   // the cached rows of this layer (loaded one layer ahead); also orders the
                                             // fresh rows below after the TMA zero fill of rows >= S_c
#pragma unroll
            for (int u = 0; u < 3; ++u)
                if ((att_pend >> u) & 1u)
                    *reinterpret_cast<uint2 *>(smem + dst_off[u]) = make_uint2((uint32_t)v[2 * u], (uint32_t)v[2 * u + 1]);
            bar_compute();
            T3(22);
            // ---- S^T = K q^T: keys are the MMA M dimension (16-key tiles over the warps), the query rows the 8-wide N;
            //      k index permuted so that every thread feeds two MMAs from one 16-byte load (as the GEMV items).  The code
            //      runs at two warps per scheduler, i.e. at the latency of its dependency chains: the warp's (up to) three
            //      tiles advance together (six independent accumulators) and the soft-cap / exp of all twelve scores is
            //      straight-line code with selects instead of branches.
            {
                constexpr int MAXT = (KEYS / 16 + NCW - 1) / NCW;   // 3
                uint4 qf[8];
#pragma unroll
                for (int kc = 0; kc < 8; ++kc)
                    qf[kc] = g < H ? *reinterpret_cast<const uint4 *>(sQ + g * LDQ + kc * 32 + 8 * t) : make_uint4(0, 0, 0, 0);
                // key rows of a tile are visited in the order rho(g) = 0 4 1 5 2 6 3 7: with the 128-byte swizzle
                // (chunk ^ row & 7) the eight lanes of a quarter-warp then hit eight different 16-byte bank groups
                const int rho = ((g & 1) << 2) | (g >> 1);
                float sa[MAXT][4], sb[MAXT][4];
#pragma unroll
                for (int i = 0; i < MAXT; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e) sa[i][e] = sb[i][e] = 0.f;
                const bool third = warp + 2 * NCW < n_mt;   // warp-uniform: rows >= n_keys are zero-filled, tiles >= n_mt skipped
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) {
                    const uint8_t *tile = smem + AttSmem::K + (kc >> 1) * KT_BYTES;
                    const int ch = (kc & 1) * 4 + t;
#pragma unroll
                    for (int i = 0; i < MAXT; ++i) {
                        if (i == MAXT - 1 && !third) continue;
                        const int r0 = (warp + i * NCW) * 16 + rho;
                        const uint4 lo = *reinterpret_cast<const uint4 *>(tile + swz(r0, ch));
                        const uint4 hi = *reinterpret_cast<const uint4 *>(tile + swz(r0 + 8, ch));
                        mma_bf16(sa[i], lo.x, hi.x, lo.y, hi.y, qf[kc].x, qf[kc].y);
                        mma_bf16(sb[i], lo.z, hi.z, lo.w, hi.w, qf[kc].z, qf[kc].w);
                    }
                }
                const float scale = 0.0625f, cap = 50.f;   // 1/sqrt(256); soft-cap (joint_model.py:139,261-268)
                // |logit| <= 50 after the soft-cap: exp() needs no running maximum
                float ls0 = 0.f, ls1 = 0.f;   // partial row sums of query rows 2t, 2t+1
#pragma unroll
                for (int i = 0; i < MAXT; ++i) {
                    if (i == MAXT - 1 && !third) continue;
                    const int r0 = (warp + i * NCW) * 16 + rho, r1 = r0 + 8;
                    const bool vis0 = (r0 < vlen) || (r0 >= p.S_v && r0 < n_keys);
                    const bool vis1 = (r1 < vlen) || (r1 >= p.S_v && r1 < n_keys);
                    float pe[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float val = __expf(tanh_fast_acc((sa[i][e] + sb[i][e]) * (scale / cap)) * cap);
                        pe[e] = sel_or_zero(e < 2 ? vis0 : vis1, val);
                    }
                    ls0 += pe[0] + pe[2];
                    ls1 += pe[1] + pe[3];
                    if (2 * t < H) {
                        sP[(2 * t) * LDP + r0] = __float2bfloat16_rn(pe[0]);
                        sP[(2 * t) * LDP + r1] = __float2bfloat16_rn(pe[2]);
                    }
                    if (2 * t + 1 < H) {
                        sP[(2 * t + 1) * LDP + r0] = __float2bfloat16_rn(pe[1]);
                        sP[(2 * t + 1) * LDP + r1] = __float2bfloat16_rn(pe[3]);
                    }
                }
                T3(25);
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { ls0 += __shfl_xor_sync(0xffffffffu, ls0, o); ls1 += __shfl_xor_sync(0xffffffffu, ls1, o); }
                if (g == 0) { sLS[warp * 8 + 2 * t] = ls0; sLS[warp * 8 + 2 * t + 1] = ls1; }
            }
            bar_compute();
            T3(23);
            // ---- O^T = V^T P^T: warp -> 16 dims of this CTA's 128; V^T fragments by transposing ldmatrix
            {
                float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f}, o2[4] = {0.f, 0.f, 0.f, 0.f}, o3[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t tile = sV + (warp >> 2) * KT_BYTES;
                const int c0 = (warp & 3) * 2;
                const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
                for (int kk = 0; kk < KEYS / 16; ++kk) {   // tiles >= n_mt: P is zero there, V rows are zero-filled
                    if (kk >= n_mt) break;
                    uint32_t a[4];
                    const int key = kk * 16 + (mi >> 1) * 8 + rr;
                    ldsm_x4_t(a, tile + swz(key, c0 + (mi & 1)));
                    const uint32_t b0 = *reinterpret_cast<const uint32_t *>(sP + g * LDP + kk * 16 + 2 * t);
                    const uint32_t b1 = *reinterpret_cast<const uint32_t *>(sP + g * LDP + kk * 16 + 8 + 2 * t);
                    if ((kk & 3) == 0) mma_bf16(o0, a[0], a[1], a[2], a[3], b0, b1);
                    else if ((kk & 3) == 1) mma_bf16(o1, a[0], a[1], a[2], a[3], b0, b1);
                    else if ((kk & 3) == 2) mma_bf16(o2, a[0], a[1], a[2], a[3], b0, b1);
                    else mma_bf16(o3, a[0], a[1], a[2], a[3], b0, b1);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) { o0[e] += o2[e]; o1[e] += o3[e]; }
                float inv0, inv1;   // 1 / row sum of query rows 2t, 2t+1
                {
                    float l0 = 0.f, l1 = 0.f;
#pragma unroll
                    for (int w = 0; w < NCW; ++w) { l0 += sLS[w * 8 + 2 * t]; l1 += sLS[w * 8 + 2 * t + 1]; }
                    inv0 = l0 > 0.f ? 1.f / l0 : 0.f;
                    inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
                }
                // c0, c1: O^T[d = warp*16 + g][m = 2t, 2t+1]; c2, c3: d + 8.  Pair neighbouring dims (lane + 4) into bf16x2
                const uint32_t fo = seq_of(p, step, IDX_ATT(l));
                const float e[4] = {(o0[0] + o1[0]) * inv0, (o0[1] + o1[1]) * inv1, (o0[2] + o1[2]) * inv0, (o0[3] + o1[3]) * inv1};
                float nb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) nb[i] = __shfl_down_sync(0xffffffffu, e[i], 4);
                if (!(g & 1)) {
                    const int n0 = head * 256 + dh * 128 + warp * 16 + g;
#pragma unroll
                    for (int mm = 0; mm < 2; ++mm) {
                        const int tok = 2 * t + mm;
                        if (tok < H) {
                            unsigned long long *row = p.ll_att[pb] + (long)(b * H + tok) * 1024;
                            ll_store(row + (n0 >> 1), pack_bf16x2(e[mm], nb[mm]), fo);
                            ll_store(row + ((n0 + 8) >> 1), pack_bf16x2(e[2 + mm], nb[2 + mm]), fo);
                        }
                    }
                }
            }
            T3(24);
            bar_compute();   // every warp is done with the layer's K / V; q / P / row sums are rewritten by the next layer
            if (tid == 0 && !(step == p.n_steps - 1 && l == p.n_layers - 1)) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the fresh rows were ordinary stores
                att_request_kv(p, smem, kmap, vmap, l + 1 == p.n_layers ? 0 : l + 1, b, dh, pol);   // one layer ahead
            }
        }
    }
}

template <int MAXM>
__global__ void __launch_bounds__(NT3, 1) denoise_mega3_kernel(const __grid_constant__ Mega3Params p,
                                                              const __grid_constant__ CUtensorMap kmap,
                                                              const __grid_constant__ CUtensorMap vmap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = align_smem(smem_raw, 1024);
    using SM = GemvSmem<MAXM>;
    const bool is_att = (int)blockIdx.x >= p.G;
    if (threadIdx.x == 0) {
        if (is_att) {
            mbar_init(reinterpret_cast<uint64_t *>(smem + AttSmem::BARS), 1);
        } else {
            uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM::BARS);
            for (int i = 0; i < SM::SLOTS; ++i) mbar_init(&full[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!is_att) {
        // this CTA's schedule entry -> shared memory
        const int *src = reinterpret_cast<const int *>(p.sched + blockIdx.x);
        int *dst = reinterpret_cast<int *>(smem + SM::SCHED);
        if (threadIdx.x < sizeof(CtaSched) / 4) dst[threadIdx.x] = src[threadIdx.x];
    }
    T3INIT();
    __syncthreads();
    if (is_att) {
        const int a = blockIdx.x - p.G, b = a / NATT, head = (a % NATT) >> 1, dh = a & 1;
        att_role(p, smem, &kmap, &vmap, b, head, dh);
    } else {
        gemv_role<MAXM>(p, smem, *reinterpret_cast<const CtaSched *>(smem + SM::SCHED));
    }
    T3FLUSH();
}

// ------------------------------------------------------------------------------------- re-packing ----
// One block per stream slot: 2048 16-byte chunks in the register-fragment order of mma.sync.m16n8k16 with the weights
// as the A operand and the k index permuted so that one 16-byte load feeds two MMAs (skinny.cu):
//   16-row items (unit = 32 k, 1 KB): chunk [unit][half][lane] = W[row(half, lane / 4)][32 unit + 8 (lane % 4) .. + 8]
//   8-row items  (unit = 32 k, 512 B): chunk [unit][lane]       = W[8 blk + lane / 4][32 unit + 8 (lane % 4) .. + 8]
struct RepackParams {
    pz_mix_layer layers[MAX_LAYERS];
    const bf16 *enc_w2a, *enc_w3, *dec_w;
    const SlotDesc *descs;
    const CtaSched *sched;
    uint8_t *slots;
    int nh, AI;
};
__global__ void __launch_bounds__(256) mega3_repack_kernel(const __grid_constant__ RepackParams p) {
    const SlotDesc d = p.descs[blockIdx.x];
    uint4 *dst = reinterpret_cast<uint4 *>(p.slots + (size_t)blockIdx.x * SLOT);
    const int qd = p.nh * 256;
    for (int ci = threadIdx.x; ci < SLOT / 16; ci += 256) {
        uint4 val = make_uint4(0, 0, 0, 0);
        const bf16 *src = nullptr;
        if (d.kind == SK_QKV || d.kind == SK_GU || d.kind == SK_E2) {
            const int unit = ci >> 6, half = (ci >> 5) & 1, lane = ci & 31, g = lane >> 2, t = lane & 3;
            const int k = unit * 32 + 8 * t;
            long row;
            const bf16 *W;
            if (d.kind == SK_QKV) {          // rows d0.. and 128 + d0.. of one head: a rotary pair lives in one item
                W = (const bf16 *)p.layers[d.layer].w_qkv;
                row = (long)(d.blk >> 4) * 256 + (d.blk & 15) * 8 + g + half * 128;
            } else if (d.kind == SK_GU) {    // 8 gate rows + the 8 matching up rows of the packed [128 gate | 128 up] layout
                W = (const bf16 *)p.layers[d.layer].w_gate_up;
                const int n = d.blk * 8 + g;
                row = (long)(n / PZ_GU_BLOCK) * (2 * PZ_GU_BLOCK) + (n % PZ_GU_BLOCK) + half * PZ_GU_BLOCK;
            } else {
                W = p.enc_w2a;
                row = (long)d.blk * 16 + g + half * 8;
            }
            src = W + row * KI + k;
        } else {
            const int unit = ci >> 5, lane = ci & 31, g = lane >> 2, t = lane & 3;
            if (d.kind == SK_O) {
                src = (const bf16 *)p.layers[d.layer].w_o + (long)(d.blk * 8 + g) * qd + unit * 32 + 8 * t;
            } else if (d.kind == SK_DP) {    // blk = CTA, aux = k step: [m tile][lane][a0 a1 a2 a3] of W_down[:, the CTA's columns]
                const CtaSched &sc = p.sched[d.blk];
                const int mt = ci >> 5;
                const bf16 *W = (const bf16 *)p.layers[d.layer].w_down;
                uint32_t a[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int row = mt * 16 + g + (q & 1) * 8, k = d.aux * 16 + 2 * t + (q >> 1) * 8;   // local column k, k + 1
                    a[q] = 0u;
                    if ((k >> 3) < sc.n_gu)
                        a[q] = *reinterpret_cast<const uint32_t *>(W + (long)row * p.AI + sc.gu_tile[k >> 3] * 8 + (k & 7));
                }
                val = make_uint4(a[0], a[1], a[2], a[3]);
            } else if (unit < 32) {          // K = 1024: half a slot
                const bf16 *W = d.kind == SK_E3 ? p.enc_w3 : p.dec_w;
                src = W + (long)(d.blk * 8 + g) * KI + unit * 32 + 8 * t;
            }
        }
        if (src) val = *reinterpret_cast<const uint4 *>(src);
        dst[ci] = val;
    }
}

// host: who does what.  Balances the bytes per CTA and layer (16 KB granules: QKV item 2, o_proj block 2, gate-up tile 2 + its
// 8 columns of the down projection 1) with a greedy least-loaded assignment.
struct Mega3Plan {
    int G = 0, NA = 0, NP = 0;
    std::vector<CtaSched> sched;
    std::vector<SlotDesc> descs;
};
static bool build_plan(const pz_config &c, int B, int num_sms, Mega3Plan &pl) {
    pl.NA = NATT * B;
    pl.G = num_sms - pl.NA;
    const int G = pl.G;
    const int n_own = c.act_hidden / 8, n_qkv = (c.n_heads + 2) * 16, n_gu = c.act_inter / 8, n_e2 = c.act_hidden / 16;
    if (G < 1 || n_own > MAXO * G || n_e2 > G) return false;
    pl.sched.assign(G, CtaSched{});
    std::vector<int> load(G, 0);
    for (auto &s : pl.sched) s.e2_blk = -1;
    for (int j = 0; j < n_own; ++j) {
        CtaSched &s = pl.sched[j % G];
        s.o_blk[s.n_o++] = j;
        load[j % G] += 2;
    }
    auto least = [&](auto ok) {
        int best = -1;
        for (int i = 0; i < G; ++i)
            if (ok(pl.sched[i]) && (best < 0 || load[i] < load[best])) best = i;
        return best;
    };
    for (int tl = 0; tl < n_gu; ++tl) {
        int i = least([](const CtaSched &s) { return s.n_gu < MAXGU; });
        if (i < 0) return false;
        pl.sched[i].gu_tile[pl.sched[i].n_gu++] = tl;
        load[i] += 3;
    }
    for (int q = 0; q < n_qkv; ++q) {
        int i = least([](const CtaSched &s) { return s.n_qkv < MAXQ; });
        if (i < 0) return false;
        pl.sched[i].qkv_blk[pl.sched[i].n_qkv++] = q;
        load[i] += 2;
    }
    pl.NP = 0;
    for (auto &s : pl.sched) s.pid = s.n_gu > 0 ? pl.NP++ : -1;
    if (pl.NP > MAXP) return false;
    for (int e = 0; e < n_e2; ++e) pl.sched[e].e2_blk = e;
    pl.sched[G - 1].is_dec = 1;
    // slot order = consumption order (gemv_role)
    pl.descs.clear();
    for (int i = 0; i < G; ++i) {
        CtaSched &s = pl.sched[i];
        s.stream_off = (long long)pl.descs.size() * SLOT;
        const size_t first = pl.descs.size();
        if (s.e2_blk >= 0) pl.descs.push_back({SK_E2, 0, s.e2_blk, 0});
        for (int j = 0; j < s.n_o; ++j) pl.descs.push_back({SK_E3, 0, s.o_blk[j], 0});
        for (int l = 0; l < c.n_layers; ++l) {
            for (int j = 0; j < s.n_qkv; ++j) pl.descs.push_back({SK_QKV, l, s.qkv_blk[j], 0});
            for (int j = 0; j < s.n_o; ++j) pl.descs.push_back({SK_O, l, s.o_blk[j], 0});
            for (int j = 0; j < s.n_gu; ++j) pl.descs.push_back({SK_GU, l, s.gu_tile[j], 0});
            for (int ks = 0; ks < (s.n_gu + 1) / 2; ++ks) pl.descs.push_back({SK_DP, l, i, ks});
        }
        if (s.is_dec) pl.descs.push_back({SK_DEC, 0, 0, 0});
        s.slots_per_step = (int)(pl.descs.size() - first);
    }
    return true;
}

constexpr size_t HDR_BYTES = 65536;   // schedule table + slot descriptors' home at the head of the stream buffer

bool make_kv_map(CUtensorMap *map, const void *base, int S_c, long slabs) {
    tcptx::EncodeTiledFn enc = tcptx::get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {256, (cuuint64_t)S_c, (cuuint64_t)slabs};
    cuuint64_t strides[2] = {512, (cuuint64_t)S_c * 512};
    cuuint32_t box[3] = {64, KBOX, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t ll_words(const pz_config &c, int B) {
    const size_t M = (size_t)B * c.horizon, A = c.act_hidden, qkvw = (size_t)(c.n_heads + 2) * 128;
    // act, z, 2 x, 2 x1, 2 + 2 sums of squares, 2 qkv, 2 att, 2 down partials (each rounded up to 16 words)
    auto r = [](size_t w) { return (w + 15) & ~(size_t)15; };
    const size_t MAXM = M <= 4 ? 4 : 8;
    return r(M * 8) + r(M * A / 2) + 4 * r(M * A / 2) + 4 * r(A / 8 * MAXM) + 2 * r(M * qkvw) + 2 * r(M * (size_t)c.n_heads * 128) +
           2 * r(A / 8 * (size_t)MAXP * 8 * MAXM);
}


// ============================ stand-alone decode attention (one CTA per sample) ==========================
// The action rows of the separate-kernel sampler (bs > 2) and of the training forward: joint_model.py:243-282 for the
// `horizon` query tokens of ALL heads of one sample over the sample's cached prefix K/V + the fresh action rows.
// MQA shares K/V between the heads, so one CTA serves the whole sample and reads its K and V exactly once: K (all
// 256 dims, 144 KB, 128-byte swizzled TMA tiles) -> S^T = K q^T for the n_heads x horizon (<= 32) query rows ->
// soft-cap / mask / exp (no running maximum: |logit| <= 50) -> P in shared memory; then V is loaded over K's tiles and
// O^T = V^T P^T.  RoPE of q and of the fresh k rows happens while they are staged (fp32).  No split-key partials, no combine
// kernel (the round-1 pair decode_attn_kernel + decode_combine_kernel: 22 us per layer at bs=64).
struct DA2Params {
    int B, H, nh, S_v, S_p, S_c, layer, batch_total, qkvw;
    const bf16 *qkv;                  // [B * H][qkvw]: q (nh x 256) | k | v, not rotated
    const float *rope_cos, *rope_sin; // [pos][128]
    const int32_t *valid_len;
    bf16 *out;                        // [b][tok][nh x 256]
    long out_batch_stride;
    int out_row_stride;
};
struct DA2Smem {
    static constexpr int LDQ = 264, LDP = KEYS + 8;
    static constexpr int KV = 0;                               // 4 tiles [KEYS][64 dims]: K, later V
    static constexpr int Q = KV + 4 * KT_BYTES;                // bf16 [32][LDQ]
    static constexpr int P = Q + 32 * LDQ * 2;                 // bf16 [32][LDP]
    static constexpr int VF = P + 32 * LDP * 2;                // bf16 [8][256]: the fresh v rows until V's tiles are there
    static constexpr int LS = VF + 8 * 256 * 2;                // float [NCW][32]
    static constexpr int BARS = LS + NCW * 32 * 4;             // k_full, v_full
    static constexpr int END = BARS + 16;
};
PZ_DEVINL void da2_request(const DA2Params &p, uint8_t *smem, const CUtensorMap *map, uint64_t *bar, int b, uint64_t pol) {
    mbar_expect_tx(bar, 4u * KT_BYTES);
    const int slab = p.layer * p.batch_total + b;
#pragma unroll
    for (int dq = 0; dq < 4; ++dq)
#pragma unroll
        for (int hb = 0; hb < 2; ++hb)
            tma_load_3d_hint(map, bar, smem + DA2Smem::KV + dq * KT_BYTES + hb * KBOX * 128, dq * 64, hb * KBOX, slab, pol);
}
// rotate dims (8c .. 8c+7, 128 + 8c ..) of one row (model/utils.py:4-16), fp32 math on the bf16 projections
PZ_DEVINL void da2_rope(const DA2Params &p, const bf16 *src, int c, int pos, uint4 &lo, uint4 &hi) {
    const uint4 r1 = *reinterpret_cast<const uint4 *>(src + c * 8), r2 = *reinterpret_cast<const uint4 *>(src + 128 + c * 8);
    const float4 *cs = reinterpret_cast<const float4 *>(p.rope_cos + (long)pos * 128 + c * 8);
    const float4 *sn = reinterpret_cast<const float4 *>(p.rope_sin + (long)pos * 128 + c * 8);
    const float4 c0 = __ldg(cs), c1 = __ldg(cs + 1), s0 = __ldg(sn), s1 = __ldg(sn + 1);
    const float cf[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, sf[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const uint32_t w1[4] = {r1.x, r1.y, r1.z, r1.w}, w2[4] = {r2.x, r2.y, r2.z, r2.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float x1a = bf16lo(w1[j]), x1b = bf16hi(w1[j]), x2a = bf16lo(w2[j]), x2b = bf16hi(w2[j]);
        o1[j] = pack_bf16x2(x1a * cf[2 * j] - x2a * sf[2 * j], x1b * cf[2 * j + 1] - x2b * sf[2 * j + 1]);
        o2[j] = pack_bf16x2(x2a * cf[2 * j] + x1a * sf[2 * j], x2b * cf[2 * j + 1] + x1b * sf[2 * j + 1]);
    }
    lo = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    hi = make_uint4(o2[0], o2[1], o2[2], o2[3]);
}
PZ_DEVINL uint32_t da2_tile_off(int row, int d) {   // byte offset of dims d .. d+7 (d % 8 == 0) of a K / V row inside the 4 tiles
    return (uint32_t)((d >> 6) * KT_BYTES) + swz(row, (d & 63) >> 3);
}

// NN: 8-row N tiles of query rows per CTA; 32 / (8 NN) CTAs share a sample (each loads K and V: the extra reads are L2 hits)
template <int NN>
__global__ void __launch_bounds__(NCT, 1) decode_attn2_kernel(const __grid_constant__ DA2Params p,
                                                              const __grid_constant__ CUtensorMap kmap,
                                                              const __grid_constant__ CUtensorMap vmap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = align_smem(smem_raw, 1024);
    using S = DA2Smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    constexpr int GROUPS = 4 / NN;
    const int b = blockIdx.x / GROUPS, row_base = (blockIdx.x % GROUPS) * (8 * NN);   // first query row of this CTA
    bf16 *sQ = reinterpret_cast<bf16 *>(smem + S::Q);
    bf16 *sP = reinterpret_cast<bf16 *>(smem + S::P);
    bf16 *sVF = reinterpret_cast<bf16 *>(smem + S::VF);
    float *sLS = reinterpret_cast<float *>(smem + S::LS);
    uint64_t *k_full = reinterpret_cast<uint64_t *>(smem + S::BARS), *v_full = k_full + 1;
    const int H = p.H, S_c = p.S_c, n_keys = S_c + H, n_mt = (n_keys + 15) >> 4, rows = p.nh * H;
    const uint64_t pol = policy_evict_first();
    if (tid == 0) {
        mbar_init(k_full, 1); mbar_init(v_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        da2_request(p, smem, &kmap, k_full, b, pol);   // the cached K does not depend on the previous kernel
    }
    // rows nobody writes must be zero (pad query rows, P beyond the key count)
    for (int i = tid; i < 32 * S::LDQ / 2; i += NCT) reinterpret_cast<uint32_t *>(sQ)[i] = 0u;
    for (int i = tid; i < 32 * S::LDP / 2; i += NCT) reinterpret_cast<uint32_t *>(sP)[i] = 0u;
    pdl_trigger();
    pdl_wait();                                        // q | k | v of this layer come from the QKV projection before us
    __syncthreads();
    const int vlen = p.valid_len[b];
    // ---- stage q (all heads, rotated), keep the fresh k rows (rotated) in registers until K's tiles have landed
    for (int i = tid; i < 8 * NN * 16; i += NCT) {
        const int rl = i >> 4, r = row_base + rl, c = i & 15, head = r / H, tok = r % H;
        if (r >= rows) continue;
        uint4 lo, hi;
        da2_rope(p, p.qkv + (long)(b * H + tok) * p.qkvw + head * 256, c, p.S_p + tok, lo, hi);
        *reinterpret_cast<uint4 *>(sQ + rl * S::LDQ + c * 8) = lo;
        *reinterpret_cast<uint4 *>(sQ + rl * S::LDQ + 128 + c * 8) = hi;
    }
    uint4 klo = make_uint4(0, 0, 0, 0), khi = klo;
    const bool has_k = tid < H * 16;                   // (tok, chunk)
    if (has_k) da2_rope(p, p.qkv + (long)(b * H + (tid >> 4)) * p.qkvw + p.nh * 256, tid & 15, p.S_p + (tid >> 4), klo, khi);
    for (int i = tid; i < H * 32; i += NCT) {          // fresh v rows, plain copy
        const int tok = i >> 5, c = i & 31;
        *reinterpret_cast<uint4 *>(sVF + tok * 256 + c * 8) =
            *reinterpret_cast<const uint4 *>(p.qkv + (long)(b * H + tok) * p.qkvw + (p.nh + 1) * 256 + c * 8);
    }
    mbar_wait_plain(k_full, 0);
    if (has_k) {
        const int row = S_c + (tid >> 4), c = tid & 15;
        *reinterpret_cast<uint4 *>(smem + S::KV + da2_tile_off(row, c * 8)) = klo;
        *reinterpret_cast<uint4 *>(smem + S::KV + da2_tile_off(row, 128 + c * 8)) = khi;
    }
    __syncthreads();
    // ---- S^T = K q^T: 16-key tiles over the warps (M), the 32 query rows as four 8-wide N tiles
    {
        constexpr int MAXT = (KEYS / 16 + NCW - 1) / NCW;   // 3
        const int rho = ((g & 1) << 2) | (g >> 1);          // conflict-free row order under the 128-byte swizzle
        float ls[NN][2];
#pragma unroll
        for (int n = 0; n < NN; ++n) ls[n][0] = ls[n][1] = 0.f;
#pragma unroll 1
        for (int i = 0; i < MAXT; ++i) {
            const int mt = warp + i * NCW;
            if (mt >= n_mt) break;
            const int r0 = mt * 16 + rho, r1 = r0 + 8;
            float sa[NN][4], sb[NN][4];
#pragma unroll
            for (int n = 0; n < NN; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) sa[n][e] = sb[n][e] = 0.f;
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) {
                const uint8_t *tile = smem + S::KV + (kc >> 1) * KT_BYTES;
                const int ch = (kc & 1) * 4 + t;
                const uint4 lo = *reinterpret_cast<const uint4 *>(tile + swz(r0, ch));
                const uint4 hi = *reinterpret_cast<const uint4 *>(tile + swz(r1, ch));
#pragma unroll
                for (int n = 0; n < NN; ++n) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(sQ + (n * 8 + g) * S::LDQ + kc * 32 + 8 * t);
                    mma_bf16(sa[n], lo.x, hi.x, lo.y, hi.y, q.x, q.y);
                    mma_bf16(sb[n], lo.z, hi.z, lo.w, hi.w, q.z, q.w);
                }
            }
            const float scale = 0.0625f, cap = 50.f;   // 1/sqrt(256); soft-cap (joint_model.py:139,261-268)
            const bool vis0 = (r0 < vlen) || (r0 >= p.S_v && r0 < n_keys);
            const bool vis1 = (r1 < vlen) || (r1 >= p.S_v && r1 < n_keys);
#pragma unroll
            for (int n = 0; n < NN; ++n) {
                float pe[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    pe[e] = sel_or_zero(e < 2 ? vis0 : vis1, __expf(tanh_fast_acc((sa[n][e] + sb[n][e]) * (scale / cap)) * cap));
                ls[n][0] += pe[0] + pe[2];
                ls[n][1] += pe[1] + pe[3];
                const int q0 = n * 8 + 2 * t;
                sP[q0 * S::LDP + r0] = __float2bfloat16_rn(pe[0]);
                sP[q0 * S::LDP + r1] = __float2bfloat16_rn(pe[2]);
                sP[(q0 + 1) * S::LDP + r0] = __float2bfloat16_rn(pe[1]);
                sP[(q0 + 1) * S::LDP + r1] = __float2bfloat16_rn(pe[3]);
            }
        }
#pragma unroll
        for (int n = 0; n < NN; ++n) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                ls[n][0] += __shfl_xor_sync(0xffffffffu, ls[n][0], o);
                ls[n][1] += __shfl_xor_sync(0xffffffffu, ls[n][1], o);
            }
            if (g == 0) { sLS[warp * 32 + n * 8 + 2 * t] = ls[n][0]; sLS[warp * 32 + n * 8 + 2 * t + 1] = ls[n][1]; }
        }
    }
    __syncthreads();   // every warp is done with K
    // ---- V over K's tiles
    if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        da2_request(p, smem, &vmap, v_full, b, pol);
    }
    mbar_wait_plain(v_full, 0);
    for (int i = tid; i < H * 32; i += NCT) {
        const int tok = i >> 5, c = i & 31;
        *reinterpret_cast<uint4 *>(smem + S::KV + da2_tile_off(S_c + tok, c * 8)) = *reinterpret_cast<const uint4 *>(sVF + tok * 256 + c * 8);
    }
    __syncthreads();
    // ---- O^T = V^T P^T: warp -> 32 dims (two 16-dim M tiles), the 32 query rows as four N tiles, keys = K dimension
    {
        float o[2][NN][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < NN; ++n) o[m][n][0] = o[m][n][1] = o[m][n][2] = o[m][n][3] = 0.f;
        const uint32_t tile = smem_u32(smem + S::KV) + (warp >> 1) * KT_BYTES;
        const int c00 = (warp & 1) * 4;
        const int mi = lane >> 3, rr = lane & 7;
#pragma unroll 2
        for (int kk = 0; kk < n_mt; ++kk) {
            const int key = kk * 16 + (mi >> 1) * 8 + rr;
            uint32_t a0[4], a1[4];
            ldsm_x4_t(a0, tile + swz(key, c00 + (mi & 1)));
            ldsm_x4_t(a1, tile + swz(key, c00 + 2 + (mi & 1)));
#pragma unroll
            for (int n = 0; n < NN; ++n) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t *>(sP + (n * 8 + g) * S::LDP + kk * 16 + 2 * t);
                const uint32_t b1 = *reinterpret_cast<const uint32_t *>(sP + (n * 8 + g) * S::LDP + kk * 16 + 8 + 2 * t);
                mma_bf16(o[0][n], a0[0], a0[1], a0[2], a0[3], b0, b1);
                mma_bf16(o[1][n], a1[0], a1[1], a1[2], a1[3], b0, b1);
            }
        }
        // c0, c1: O^T[d = 32 warp + 16 m + g][query row 8 n + 2t, + 1]; c2, c3: d + 8.  Neighbouring dims (lane + 4) pair up.
#pragma unroll
        for (int n = 0; n < NN; ++n) {
            float inv[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float l = 0.f;
#pragma unroll
                for (int w = 0; w < NCW; ++w) l += sLS[w * 32 + n * 8 + 2 * t + j];
                inv[j] = l > 0.f ? 1.f / l : 0.f;
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float e[4] = {o[m][n][0] * inv[0], o[m][n][1] * inv[1], o[m][n][2] * inv[0], o[m][n][3] * inv[1]};
                float nb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) nb[i] = __shfl_down_sync(0xffffffffu, e[i], 4);
                if (!(g & 1)) {
                    const int d0 = warp * 32 + m * 16 + g;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int r = row_base + n * 8 + 2 * t + j;
                        if (r < rows) {
                            const int head = r / H, tok = r % H;
                            bf16 *dst = p.out + (long)b * p.out_batch_stride + (long)tok * p.out_row_stride + head * 256;
                            *reinterpret_cast<uint32_t *>(dst + d0) = pack_bf16x2(e[j], nb[j]);
                            *reinterpret_cast<uint32_t *>(dst + d0 + 8) = pack_bf16x2(e[2 + j], nb[2 + j]);
                        }
                    }
                }
            }
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------- host side ----
#ifdef PZ_MEGA_TRACE
extern "C" int pz_debug_mega3_trace(unsigned long long *host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g3_trace, (size_t)n * 8);
}
#endif
int denoise_mega3_supported(const pz_config &c, int B) {
    static const bool off = [] { const char *e = getenv("PZ_MEGA3"); return e && e[0] == '0'; }();
    if (off) return 0;
    if (c.dtype != PZ_BF16 || (c.flags & PZ_FLAG_SIMPLE_KERNELS)) return 0;
    if (B < 1 || B * c.horizon > 8) return 0;
    if (c.head_dim != 256 || c.n_kv_heads != 1 || c.n_heads != 8) return 0;
    if (c.act_hidden != KI || c.act_inter != 4096) return 0;
    if (c.n_layers > MAX_LAYERS || c.action_dim > 8) return 0;
    if (c.s_vlm + c.cond_steps + c.horizon > KEYS) return 0;
    return 1;
}

size_t denoise_mega3_ll_bytes(const pz_config &c, int B) {
    if (!denoise_mega3_supported(c, B)) return 0;
    return ll_words(c, B) * 8 + 256;
}

size_t denoise_mega3_stream_bytes(const pz_config &c, int B, int num_sms) {
    if (!denoise_mega3_supported(c, B)) return 0;
    Mega3Plan pl;
    if (!build_plan(c, B, num_sms, pl)) return 0;
    return HDR_BYTES + pl.descs.size() * sizeof(SlotDesc) + 1024 + pl.descs.size() * (size_t)SLOT;
}

// Builds the schedule for `num_sms` CTAs and re-packs the action expert's weights into the per-CTA item streams.
// Synchronous with respect to the host for the (small) table upload; the re-pack kernel runs on `st`.
int denoise_mega3_pack(const pz_config &c, const pz_weights &w, const pz_mix_layer *layers, int B, int num_sms, void *buf,
                       size_t bytes, Mega3State *state, cudaStream_t st, const char **err) {
    Mega3Plan pl;
    if (!denoise_mega3_supported(c, B) || !build_plan(c, B, num_sms, pl)) {
        if (err) *err = "persistent sampler: configuration / SM count not supported";
        return PZ_ERR_INVALID;
    }
    const size_t desc_bytes = pl.descs.size() * sizeof(SlotDesc);
    const size_t slots_off = (HDR_BYTES + desc_bytes + 1023) & ~(size_t)1023;
    if (bytes < slots_off + pl.descs.size() * (size_t)SLOT || pl.sched.size() * sizeof(CtaSched) > HDR_BYTES) {
        if (err) *err = "persistent sampler: stream buffer too small";
        return PZ_ERR_WORKSPACE;
    }
    if (((uintptr_t)buf) & 1023) {
        if (err) *err = "persistent sampler: stream buffer must be 1 KiB aligned";
        return PZ_ERR_INVALID;
    }
    uint8_t *base = (uint8_t *)buf;
    if (cudaMemcpyAsync(base, pl.sched.data(), pl.sched.size() * sizeof(CtaSched), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(base + HDR_BYTES, pl.descs.data(), desc_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {   // the host vectors go out of scope
        if (err) *err = "persistent sampler: table upload failed";
        return PZ_ERR_CUDA;
    }
    RepackParams rp;
    memset(&rp, 0, sizeof(rp));
    for (int l = 0; l < c.n_layers; ++l) rp.layers[l] = layers[l];
    rp.enc_w2a = (const bf16 *)w.enc_w2a; rp.enc_w3 = (const bf16 *)w.enc_w3; rp.dec_w = (const bf16 *)w.dec_w;
    rp.descs = (const SlotDesc *)(base + HDR_BYTES);
    rp.sched = (const CtaSched *)base;
    rp.slots = base + slots_off;
    rp.nh = c.n_heads; rp.AI = c.act_inter;
    mega3_repack_kernel<<<(unsigned)pl.descs.size(), 256, 0, st>>>(rp);
    if (cudaPeekAtLastError() != cudaSuccess) {
        if (err) *err = "persistent sampler: re-pack launch failed";
        return PZ_ERR_CUDA;
    }
    state->buf = buf; state->bytes = bytes; state->B = B; state->G = pl.G; state->NA = pl.NA; state->NP = pl.NP; state->num_sms = num_sms;
    state->slots_off = slots_off;
    return 0;
}

int launch_denoise_mega3(const pz_config &c, const pz_weights &w, const pz_mix_layer *layers, const Mega3State &state,
                         const Mega3Buffers &bf, int B, cudaStream_t st, const char **err) {
#ifdef PZ_MEGA_TRACE   // the trace counters are static shared memory: the dynamic part then starts 1 KiB-aligned, no slack needed
    constexpr int SLACK = 0;
#else
    constexpr int SLACK = 1024;
#endif
    constexpr int SMEM4 = (GemvSmem<4>::END > AttSmem::END ? GemvSmem<4>::END : AttSmem::END) + SLACK;
    constexpr int SMEM8 = (GemvSmem<8>::END > AttSmem::END ? GemvSmem<8>::END : AttSmem::END) + SLACK;
    static_assert(SMEM4 <= 227 * 1024 && SMEM8 <= 227 * 1024, "shared memory budget");
    const int M = B * c.horizon;
    const void *fn = M <= 4 ? (const void *)denoise_mega3_kernel<4> : (const void *)denoise_mega3_kernel<8>;
    const int smem = M <= 4 ? SMEM4 : SMEM8;
    static PerDeviceOnce attr_once[2];
    if (attr_once[M <= 4 ? 0 : 1].need() &&
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
        if (err) *err = "persistent sampler: cannot set the shared-memory size";
        return PZ_ERR_CUDA;
    }
    if (state.B != B || !state.buf) {
        if (err) *err = "persistent sampler: weights not packed for this batch";
        return PZ_ERR_UNBOUND;
    }
    Mega3Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = c.horizon; p.M = M; p.nh = c.n_heads; p.AI = c.act_inter;
    p.S_v = c.s_vlm; p.S_p = c.cond_steps; p.S_c = c.s_vlm + c.cond_steps; p.n_layers = c.n_layers; p.n_steps = c.n_steps;
    p.action_dim = c.action_dim; p.skp = w.small_k_pad;
    p.G = state.G; p.NA = state.NA; p.NP = state.NP;
    {
        static const int pf = [] { const char *e = getenv("PZ_M3_PF"); return e ? atoi(e) : 0; }();
        p.pf_dist = pf;
        static const int sen = [] { const char *e = getenv("PZ_M3_SENTINEL"); return e ? atoi(e) : 1; }();
        p.sentinel = sen;
    }
    p.dt = (float)(1.0 / c.n_steps); p.clip = c.clip;
    for (int l = 0; l < c.n_layers; ++l) { p.norm_in[l] = layers[l].norm_in; p.norm_post[l] = layers[l].norm_post; }
    p.final_norm = w.action_final_norm;
    p.enc_w1 = (const bf16 *)w.enc_w1;
    p.enc_b1 = w.enc_b1; p.enc_time_bias = w.enc_time_bias; p.enc_b3 = w.enc_b3; p.dec_b = w.dec_b;
    p.rope_cos = w.rope_act_cos; p.rope_sin = w.rope_act_sin;
    p.valid_len = bf.valid_len; p.noise = bf.noise; p.out = bf.out;
    p.stream = (const uint8_t *)state.buf + state.slots_off;
    p.sched = (const CtaSched *)state.buf;
    p.batch_total = bf.batch_total;
    {   // carve the exchange buffers
        unsigned long long *q = (unsigned long long *)bf.ll;
        auto take = [&](size_t words) { unsigned long long *r = q; q += (words + 15) & ~(size_t)15; return r; };
        const size_t Mz = M, A = c.act_hidden, qkvw = (size_t)(c.n_heads + 2) * 128;
        p.ll_act = take(Mz * 8); p.ll_z = take(Mz * A / 2);
        const size_t maxm = M <= 4 ? 4 : 8;
        for (int i = 0; i < 2; ++i) p.ll_x[i] = take(Mz * A / 2);
        for (int i = 0; i < 2; ++i) p.ll_x1[i] = take(Mz * A / 2);
        for (int i = 0; i < 2; ++i) p.ll_sx[i] = take(A / 8 * maxm);
        for (int i = 0; i < 2; ++i) p.ll_sx1[i] = take(A / 8 * maxm);
        for (int i = 0; i < 2; ++i) p.ll_qkv[i] = take(Mz * qkvw);
        for (int i = 0; i < 2; ++i) p.ll_att[i] = take(Mz * (size_t)c.n_heads * 128);
        for (int i = 0; i < 2; ++i) p.ll_rs[i] = take(A / 8 * (size_t)state.NP * 8 * maxm);
        p.err = (unsigned int *)q;
        if ((size_t)((char *)q - (char *)bf.ll) + 256 > bf.ll_bytes) {
            if (err) *err = "persistent sampler: exchange workspace too small";
            return PZ_ERR_WORKSPACE;
        }
        if (cudaMemsetAsync(bf.ll, 0, (size_t)((char *)q - (char *)bf.ll) + 256, st) != cudaSuccess) {   // the part in use
            if (err) *err = "persistent sampler: memset failed";
            return PZ_ERR_CUDA;
        }
    }
    CUtensorMap kmap, vmap;
    const long slabs = (long)c.n_layers * bf.batch_total;
    if (!make_kv_map(&kmap, bf.kcache, p.S_c, slabs) || !make_kv_map(&vmap, bf.vcache, p.S_c, slabs)) {
        if (err) *err = "persistent sampler: cuTensorMapEncodeTiled failed";
        return PZ_ERR_CUDA;
    }
    void *args[] = {&p, &kmap, &vmap};
    // cooperative launch: all CTAs are guaranteed co-resident (the polled exchanges rely on it)
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(state.num_sms), dim3(NT3), args, (size_t)smem, st);
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        return PZ_ERR_CUDA;
    }
    count_launch();
    return 0;
}

// ---------------------------------------------------------------- stand-alone decode attention (host) ----
int decode_attention2_supported(const pz_config &c) {
    static const bool off = [] { const char *e = getenv("PZ_DECODE_ATTN2"); return e && e[0] == '0'; }();
    return !off && c.dtype == PZ_BF16 && !(c.flags & PZ_FLAG_SIMPLE_KERNELS) && c.head_dim == 256 && c.n_kv_heads == 1 &&
           c.n_heads * c.horizon <= 32 && c.horizon <= 8 && c.s_vlm + c.cond_steps + c.horizon <= KEYS;
}

int launch_decode_attention2(const pz_config &c, const pz_weights &w, const void *qkv, const void *kcache, const void *vcache,
                             int batch_total, const int32_t *valid_len, int layer, int B, void *out, long out_batch_stride,
                             int out_row_stride, cudaStream_t st) {
    constexpr int smem = DA2Smem::END + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    // few samples: two CTAs per sample (16 query rows each) so that more SMs share the work
    const int device_sms = device_sm_count();
    const bool two = c.n_heads * c.horizon > 16 && 2 * B <= device_sms;
    static PerDeviceOnce attr_once[2];
    if (attr_once[two ? 1 : 0].need() &&
        cudaFuncSetAttribute(two ? decode_attn2_kernel<2> : decode_attn2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem) != cudaSuccess)
        return PZ_ERR_CUDA;
    DA2Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = c.horizon; p.nh = c.n_heads; p.S_v = c.s_vlm; p.S_p = c.cond_steps; p.S_c = c.s_vlm + c.cond_steps;
    p.layer = layer; p.batch_total = batch_total; p.qkvw = (c.n_heads + 2) * 256;
    p.qkv = (const bf16 *)qkv; p.rope_cos = w.rope_act_cos; p.rope_sin = w.rope_act_sin; p.valid_len = valid_len;
    p.out = (bf16 *)out; p.out_batch_stride = out_batch_stride; p.out_row_stride = out_row_stride;
    CUtensorMap kmap, vmap;
    const long slabs = (long)c.n_layers * batch_total;
    if (!make_kv_map(&kmap, kcache, p.S_c, slabs) || !make_kv_map(&vmap, vcache, p.S_c, slabs)) return PZ_ERR_CUDA;
    if (two) launch_k(decode_attn2_kernel<2>, dim3(2 * B), dim3(NCT), (size_t)smem, st, p, kmap, vmap);
    else launch_k(decode_attn2_kernel<4>, dim3(B), dim3(NCT), (size_t)smem, st, p, kmap, vmap);
    return 0;
}
