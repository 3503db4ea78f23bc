// denoise_mega2.cu -- second-generation persistent Euler sampler (pizero.py:454-489) for
// M = B * horizon <= 8 action rows (bs 1-2 at chunk 4): the latency regime.
//
// What bounds the first generation (denoise_mega.cu) is not HBM but the five all-to-all exchanges
// per layer: grid barrier (~1.5 us) + an L2 round trip to re-stage the activations (~1.3 us), plus
// the per-item reduction/synchronisation of the GEMV phases.  This kernel removes both:
//
//   * No grid barriers.  Every exchanged value travels as one 64-bit word {payload, sequence flag}
//     (the LL protocol of collective libraries): the producer publishes with a single 8-byte store,
//     consumers poll the data words themselves, so an exchange costs one store + one polled load
//     instead of store / fence / atomic / poll / load.  Flags are the global phase number
//     (monotonic), buffers are reused across layers; the dependency chain of the network makes the
//     reuse safe (a writer of phase p of layer l+1 has consumed data that transitively depends on
//     every reader of phase p of layer l).
//   * Role split: the last 2*B CTAs only do attention.  Each owns one half of a sample's keys
//     (<= 144) and keeps that half of the layer's cached K and V resident in shared memory, loaded
//     by TMA bulk copies one layer ahead (the prefix KV is step-invariant and L2-resident), so the
//     attention phase is Q-arrival -> S -> softmax -> PV with no memory latency in it.
//   * The other CTAs stream weights: one producer warp issues 1 KB TMA bulk copies (L2 evict-first)
//     of 8-row x 1024-k items into an 8-slot mbarrier ring, in the fixed order in which the CTA will
//     consume them, running ahead across phases, layers and steps.  Eight consumer warps split k,
//     keep the accumulators of all the CTA's blocks of a phase in registers and reduce through
//     shared memory once per phase.
//   * Every output element has exactly one producer (no split-K partial sums, no atomics): 16-row
//     blocks for QKV / gate-up, 8-row blocks with the full K for o_proj / down, whose owners add the
//     residual they captured while staging and publish the new fp32 residual stream directly.
//
// Numerics are those of denoise_mega.cu (bf16 operands, fp32 accumulation and residual stream),
// except that the two split-key attention partials are exchanged as bf16.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace {

#define PZ_DEVNOINL __device__ __noinline__

constexpr int NCW = 8;                    // compute warps
constexpr int NCT = NCW * 32;             // compute threads
constexpr int NT2 = NCT + 32;             // + one producer warp
constexpr int KI = 1024;                  // k extent of one weight item
constexpr int SLOTS = 8;                  // ring depth (items)
constexpr int ROWB = KI * 2 + 64;         // padded shared-memory row of an item (conflict-free 16-byte reads)
constexpr int SLOT_BYTES = 8 * ROWB;      // 8 rows
constexpr int ITEM_TX = 8 * KI * 2;       // bytes one item brings in
constexpr int MAXM = 8;
constexpr int KMAX = 4096;                // widest staged activation (down_proj input)
constexpr int LDA = KMAX + 32;            // staged activation row stride (bf16): stride % 128 B == 64
constexpr int MAXBLK = 4;                 // blocks of one phase per CTA
constexpr int MAX_LAYERS = 24;
// attention role
constexpr int KS = 144;                   // keys per attention CTA (half of <= 288)
constexpr int QROWS = 32, LDQ = 256 + 8;
constexpr int LDP = KS + 8;

// shared memory map -- GEMV role
constexpr int SM_RING = 0;
constexpr int SM_AST = SM_RING + SLOTS * SLOT_BYTES;               // bf16 [MAXM][LDA]
constexpr int SM_RED = SM_AST + MAXM * LDA * 2;                    // float [NCW][MAXBLK][16][MAXM + 1]
constexpr int SM_MISC = SM_RED + NCW * MAXBLK * 16 * (MAXM + 1) * 4;
//   misc (floats): part[16], rs[8], save_o[8][8], save_d[8][8], act_s[8][8], mlS[2*2*32*2]
constexpr int MISC_PART = 0, MISC_RS = 16, MISC_SAVE_O = 24, MISC_SAVE_D = 88, MISC_ACT = 152, MISC_ML = 216;
constexpr int MISC_FLOATS = MISC_ML + 2 * 2 * 32 * 2;
constexpr int SM_BARS = SM_MISC + MISC_FLOATS * 4;                 // full[SLOTS], empty[SLOTS]
constexpr int SM_GEMV_END = SM_BARS + 2 * SLOTS * 8;
// shared memory map -- attention role
constexpr int SA_K = 0;                                            // bf16 [KS][LDQ]
constexpr int SA_V = SA_K + KS * LDQ * 2;
constexpr int SA_Q = SA_V + KS * LDQ * 2;                          // bf16 [32][LDQ]
constexpr int SA_P = SA_Q + QROWS * LDQ * 2;                       // bf16 [32][LDP]
constexpr int SA_RS = SA_P + QROWS * LDP * 2;                      // float [4][32]
constexpr int SA_BARS = SA_RS + 4 * QROWS * 4;                     // kv_full, kv_empty
constexpr int SM_ATT_END = SA_BARS + 16;
constexpr int SMEM2_TOTAL = (SM_GEMV_END > SM_ATT_END ? SM_GEMV_END : SM_ATT_END) + 128;
static_assert(SMEM2_TOTAL <= 227 * 1024, "shared memory budget");

enum Phase2 { P_ENC2 = 0, P_ENC3, P_QKV, P_O, P_GU, P_D, P_DEC };

struct Mega2Params {
    int B, H, M, A, AI, nh, S_v, S_p, S_c, n_layers, n_steps, action_dim, skp;
    int G;                    // weight-streaming CTAs; CTAs G .. G + 2B - 1 do attention
    int ks;                   // keys per attention CTA
    int lgH;                  // log2(horizon) (power of two required: index math on the exchange path is shifts only)
    float dt, clip;
    pz_mix_layer layers[MAX_LAYERS];
    const float *final_norm;
    const bf16 *enc_w1, *enc_w2a, *enc_w3, *dec_w;
    const float *enc_b1, *enc_time_bias, *enc_b3, *dec_b;
    const float *rope_cos, *rope_sin;
    const bf16 *kcache, *vcache;
    long kv_layer_stride, kv_batch_stride;
    const int32_t *valid_len;
    const float *noise;       // [M][action_dim]: initial action
    float *out;
    // LL exchange buffers (64-bit words {payload, flag}), zeroed before every launch
    unsigned long long *ll_act;   // [M][8]           fp32
    unsigned long long *ll_z;     // [M][A/2]         bf16x2
    unsigned long long *ll_x1;    // [M][A]           fp32 (residual after o_proj)
    unsigned long long *ll_x2;    // [M][A]           fp32 (residual entering a layer)
    unsigned long long *ll_qkv;   // [M][qkvd/2]      bf16x2
    unsigned long long *ll_att;   // [B][2][32][128]  bf16x2 unnormalised partial O
    unsigned long long *ll_l;     // [B][2][32]       fp32 softmax denominators of the partials
    unsigned long long *ll_mlp;   // [M][AI/2]        bf16x2
    unsigned int *err;            // set to 1 if any wait ran into its bound (never hangs the GPU)
    unsigned long long *trace;    // optional globaltimer stamps (CTA 0 and the first attention CTA, step 1 / layer 1)
};

// ---------------------------------------------------------------- PTX helpers ----
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
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// debug: the phase each CTA is in (thread 0 writes it), reported by the first wait that runs into its bound
__device__ unsigned int g_site[256];
PZ_DEVINL void set_site(unsigned int code) { if (threadIdx.x == 0) g_site[blockIdx.x] = code; }
PZ_DEVINL void raise_err(const Mega2Params &p, unsigned int kind) {
    const unsigned int code = 0x80000000u | (kind << 24) | (blockIdx.x << 12) | (g_site[blockIdx.x] & 0xFFF);
    atomicCAS(p.err + kind, 0u, code);   // first failure of each kind (1 mbarrier/compute, 2 mbarrier/producer, 3 spin, 4 gather)
    atomicCAS(p.err, 0u, code);
}
PZ_DEVINL bool err_set(const Mega2Params &p) { return *reinterpret_cast<volatile unsigned int *>(p.err) != 0; }
// bounded wait (a bug or a lost CTA must never hang the device): on timeout raise the error flag
PZ_DEVINL void mbar_wait(const Mega2Params &p, uint64_t *bar, uint32_t parity) {
    long spins = 0;
    while (!mbar_try(bar, parity)) {
        ++spins;
        if ((spins & 0xFFF) == 0 && (spins > (1L << 22) || err_set(p))) { raise_err(p, 1 + (threadIdx.x >= NCT)); return; }
    }
}
PZ_DEVINL void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
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
PZ_DEVINL void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }
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
PZ_DEVINL void ldsm_x2(uint32_t (&r)[2], const void *p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}
PZ_DEVINL void ldsm_x4_t(uint32_t (&r)[4], const void *p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
// gelu-tanh with the hardware tanh (rel. error 2^-11, below the bf16 rounding of the output; same as the
// GeGLU epilogue of the tcgen05 GEMM)
PZ_DEVINL float gelu_fast(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(k0 * (x + k1 * x * x * x)));
    return 0.5f * x * (1.0f + y);
}
PZ_DEVINL float tanh_fast_acc(float y) { float t = __expf(2.f * y); return 1.f - __fdividef(2.f, t + 1.f); }

// ---- LL words ------------------------------------------------------------------------------------
PZ_DEVINL void ll_store(unsigned long long *dst, uint32_t payload, uint32_t flag) {
    unsigned long long v = ((unsigned long long)flag << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
}
PZ_DEVINL unsigned long long ll_load1(const unsigned long long *src) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
    return v;
}
PZ_DEVINL void ll_load2(const unsigned long long *src, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
}
// One thread of the CTA spins on a sentinel word (a word of one of the slowest producers of the phase,
// different for every CTA, so the pollers do not pile up on one L2 line); the gather that follows
// re-reads only the words that are still missing.
PZ_DEVINL void ll_spin(const Mega2Params &p, const unsigned long long *src, uint32_t flag) {
    long spins = 0;
    while ((uint32_t)(ll_load1(src) >> 32) != flag) {
        ++spins;
        if ((spins & 0x3FF) == 0 && (spins > (1L << 21) || err_set(p))) { raise_err(p, 3); return; }
    }
}
PZ_DEVINL void cta_wait(const Mega2Params &p, const unsigned long long *sentinel, uint32_t flag) {
    if (threadIdx.x == 0) ll_spin(p, sentinel, flag);
    bar_compute();
}
// N double-words (16 bytes = two LL words each) per thread, all in flight.  After the sentinel the data
// is almost always complete: one pass + one cheap combined flag test; otherwise everything is re-read.
template <int N, typename AddrFn>
PZ_DEVINL void ll_gather(const Mega2Params &p, uint32_t flag, int n_valid, AddrFn addr, unsigned long long (&v)[2 * N]) {
    long spins = 0;
    while (true) {
#pragma unroll
        for (int u = 0; u < N; ++u)
            if (u < n_valid) ll_load2(addr(u), v[2 * u], v[2 * u + 1]);
        uint32_t diff = 0;
#pragma unroll
        for (int u = 0; u < N; ++u)
            if (u < n_valid) diff |= ((uint32_t)(v[2 * u] >> 32) ^ flag) | ((uint32_t)(v[2 * u + 1] >> 32) ^ flag);
        if (diff == 0) return;
        ++spins;
        if ((spins & 0x3FF) == 0 && (spins > (1L << 21) || err_set(p))) { raise_err(p, 4); return; }
    }
}

// ---- phase geometry (shared by the producer warp and the consumers) -------------------------------
PZ_DEVINL int ph_nblk(const Mega2Params &p, int ph) {
    switch (ph) {
        case P_ENC2: case P_ENC3: case P_O: case P_D: return p.A / 8;
        case P_QKV: return (p.nh + 2) * 16;
        case P_GU: return p.AI / 8;
        default: return 1;
    }
}
PZ_DEVINL int ph_items(const Mega2Params &p, int ph) {   // 8-row x 1024-k items per block
    switch (ph) {
        case P_QKV: case P_GU: return 2;                  // rows g | rows g+8 of the 16-row MMA tile
        case P_O: return (p.nh * 256) / KI;
        case P_D: return p.AI / KI;
        default: return 1;
    }
}
// block b of phase ph runs on CTA (b + off) % G; the offsets put second-round blocks on the CTAs that
// have no o_proj / down block, so the bytes per CTA and layer stay balanced
PZ_DEVINL int ph_off(const Mega2Params &p, int ph) {
    switch (ph) {
        case P_QKV: return p.A / 8 < p.G ? p.A / 8 : 0;
        case P_GU: return p.G / 2;
        case P_DEC: return p.G - 1;
        default: return 0;
    }
}
PZ_DEVINL int ph_first(const Mega2Params &p, int ph, int c) {
    int off = ph_off(p, ph) % p.G;
    return (c + p.G - off) % p.G;
}
// global source of item `it` of block `blk`: row r (0..7), 1024 k starting at the returned pointer
PZ_DEVINL const bf16 *item_row(const Mega2Params &p, int ph, int layer, int blk, int it, int r) {
    switch (ph) {
        case P_ENC2: return p.enc_w2a + (long)(blk * 8 + r) * KI;
        case P_ENC3: return p.enc_w3 + (long)(blk * 8 + r) * KI;
        // block = (head blk / 16, 8 dims d0 = 8 * (blk % 16)): item 0 rows d0.., item 1 rows 128 + d0.. -- the two
        // halves a rotary pair lives in, so the epilogue can rotate (model/utils.py:4-16) without an exchange
        case P_QKV: return (const bf16 *)p.layers[layer].w_qkv + (long)((blk >> 4) * 256 + it * 128 + (blk & 15) * 8 + r) * KI;
        case P_O: return (const bf16 *)p.layers[layer].w_o + (long)(blk * 8 + r) * (p.nh * 256) + it * KI;
        case P_GU: {   // 8 gate rows, then the 8 matching up rows of the packed [128 gate | 128 up] layout
            int n = blk * 8;
            long row = (long)(n / PZ_GU_BLOCK) * (2 * PZ_GU_BLOCK) + (n % PZ_GU_BLOCK) + it * PZ_GU_BLOCK + r;
            return (const bf16 *)p.layers[layer].w_gate_up + row * KI;
        }
        case P_D: return (const bf16 *)p.layers[layer].w_down + (long)(blk * 8 + r) * p.AI + it * KI;
        default: return p.dec_w + (long)r * KI;
    }
}

PZ_DEVINL uint32_t seq_flag(const Mega2Params &p, int step, int idx) { return 1u + (uint32_t)(step * (3 + 5 * p.n_layers) + idx); }
// exchange indices inside a step
PZ_DEVINL int IDX_ACT() { return 0; }
PZ_DEVINL int IDX_Z() { return 1; }
PZ_DEVINL int IDX_X0() { return 2; }
PZ_DEVINL int IDX_QKV(int l) { return 3 + 5 * l; }
PZ_DEVINL int IDX_ATT(int l) { return 4 + 5 * l; }
PZ_DEVINL int IDX_X1(int l) { return 5 + 5 * l; }
PZ_DEVINL int IDX_MLP(int l) { return 6 + 5 * l; }
PZ_DEVINL int IDX_X2(int l) { return 7 + 5 * l; }

// trace layout: [CTA][16] globaltimer words (step 1, layer 1)
PZ_DEVINL void stamp(const Mega2Params &p, bool on, int idx) {
    if (on && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.trace[blockIdx.x * 16 + (idx & 15)] = t;
    }
}

// sentinel block of a phase for CTA c: one of the blocks of the LAST round (its producers finish last)
PZ_DEVINL int sentinel_block(const Mega2Params &p, int ph, int c) {
    const int nblk = ph_nblk(p, ph);
    const int start = ((nblk - 1) / p.G) * p.G;
    return start + c % (nblk - start);
}

// ======================================= weight-streaming role ======================================
struct GemvCtx {
    uint8_t *smem;
    uint64_t *full, *empty;
    uint32_t cnt;     // items consumed so far (ring position)
    int c;            // CTA index among the streaming CTAs
    int first[8];     // first block of each phase on this CTA (ph_first)
    bool tr;          // trace this phase
};

PZ_DEVINL void producer_loop(const Mega2Params &p, uint8_t *smem, int c) {
    const int lane = threadIdx.x & 31;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM_BARS), *empty = full + SLOTS;
    const uint64_t pol = policy_evict_first();
    uint32_t cnt = 0;
    auto phase = [&](int ph, int layer) {
        const int nblk = ph_nblk(p, ph), items = ph_items(p, ph);
        for (int blk = ph_first(p, ph, c); blk < nblk; blk += p.G) {
            for (int it = 0; it < items; ++it, ++cnt) {
                const int slot = cnt % SLOTS;
                if (cnt >= SLOTS) mbar_wait(p, &empty[slot], ((cnt / SLOTS) - 1) & 1);
                if (lane == 0) mbar_expect_tx(&full[slot], ITEM_TX);
                __syncwarp();
                if (lane < 16) {
                    const int r = lane >> 1, hf = lane & 1;
                    bulk_g2s(smem + SM_RING + slot * SLOT_BYTES + r * ROWB + hf * KI, item_row(p, ph, layer, blk, it, r) + hf * (KI / 2),
                             KI, &full[slot], pol);
                }
            }
        }
    };
    for (int step = 0; step < p.n_steps; ++step) {
        phase(P_ENC2, 0);
        phase(P_ENC3, 0);
        for (int l = 0; l < p.n_layers; ++l) {
            phase(P_QKV, l);
            phase(P_O, l);
            phase(P_GU, l);
            phase(P_D, l);
        }
        phase(P_DEC, 0);
    }
}

// cross-warp reduction scratch: red[warp][j][row 0..15][m]
PZ_DEVINL float *red_ptr(uint8_t *smem, int w, int j, int r) {
    return reinterpret_cast<float *>(smem + SM_RED) + ((w * MAXBLK + j) * 16 + r) * (MAXM + 1);
}
PZ_DEVINL void red_write(uint8_t *smem, int nj, const float (&acc)[MAXBLK][4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < MAXBLK; ++j) {
        if (j < nj) {
            float *r0 = red_ptr(smem, warp, j, g), *r1 = red_ptr(smem, warp, j, g + 8);
            r0[2 * t] = acc[j][0]; r0[2 * t + 1] = acc[j][1];
            r1[2 * t] = acc[j][2]; r1[2 * t + 1] = acc[j][3];
        }
    }
}
PZ_DEVINL float red_sum(uint8_t *smem, int j, int r, int m) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) v += red_ptr(smem, w, j, r)[m];
    return v;
}

// accumulate this CTA's blocks of one phase: acc[j] = W_block_j . A^T (partial over this warp's k range)
// 16-row blocks (QKV, gate|up): two ring items per block (rows g | rows g + 8 of the MMA tile)
template <int NJ>
PZ_DEVNOINL int gemv_acc16(const Mega2Params &p, GemvCtx &cx, int ph) {
    float acc[MAXBLK][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bf16 *As = reinterpret_cast<const bf16 *>(cx.smem + SM_AST);
    const int nblk = ph_nblk(p, ph), first = cx.first[ph];
    uint4 x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = warp * 128 + (q >> 1) * 64 + (q & 1) * 32 + 8 * t;
        x[q] = g < p.M ? *reinterpret_cast<const uint4 *>(As + g * LDA + k) : make_uint4(0, 0, 0, 0);
    }
    int nj = 0;
    stamp(p, cx.tr && NJ == 4, 13);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if (first + j * p.G >= nblk) break;
        nj = j + 1;
        float a2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // two independent MMA chains
        const int s0 = cx.cnt % SLOTS, s1 = (cx.cnt + 1) % SLOTS;
        mbar_wait(p, &cx.full[s0], (cx.cnt / SLOTS) & 1);
        mbar_wait(p, &cx.full[s1], ((cx.cnt + 1) / SLOTS) & 1);
        if (j == 0) stamp(p, cx.tr && NJ == 4, 14);
        const uint8_t *w0 = cx.smem + SM_RING + s0 * SLOT_BYTES + g * ROWB + (warp * 128 + 8 * t) * 2;
        const uint8_t *w1 = cx.smem + SM_RING + s1 * SLOT_BYTES + g * ROWB + (warp * 128 + 8 * t) * 2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ko = ((q >> 1) * 64 + (q & 1) * 32) * 2;
            uint4 ag = *reinterpret_cast<const uint4 *>(w0 + ko);
            uint4 ag8 = *reinterpret_cast<const uint4 *>(w1 + ko);
            mma_bf16(a2[0], ag.x, ag8.x, ag.y, ag8.y, x[q].x, x[q].y);
            mma_bf16(a2[1], ag.z, ag8.z, ag.w, ag8.w, x[q].z, x[q].w);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = a2[0][e] + a2[1][e];
        __syncwarp();
        if (lane == 0) { mbar_arrive(&cx.empty[s0]); mbar_arrive(&cx.empty[s1]); }
        cx.cnt += 2;
    }
    stamp(p, cx.tr && NJ == 4, 15);
    red_write(cx.smem, nj, acc);
    return nj;
}
// 8-row blocks with the full K (encoder, o_proj, down, decoder): at most one block per CTA, `items` k-slices
PZ_DEVNOINL int gemv_acc8(const Mega2Params &p, GemvCtx &cx, int ph) {
    float acc[MAXBLK][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bf16 *As = reinterpret_cast<const bf16 *>(cx.smem + SM_AST);
    if (cx.first[ph] >= ph_nblk(p, ph)) return 0;
    const int items = ph_items(p, ph);
    float a4[4][4];   // four independent MMA chains
#pragma unroll
    for (int i = 0; i < 4; ++i) a4[i][0] = a4[i][1] = a4[i][2] = a4[i][3] = 0.f;
    for (int it = 0; it < items; ++it) {
        const int s0 = cx.cnt % SLOTS;
        mbar_wait(p, &cx.full[s0], (cx.cnt / SLOTS) & 1);
        const uint8_t *w0 = cx.smem + SM_RING + s0 * SLOT_BYTES + g * ROWB + (warp * 128 + 8 * t) * 2;
        const bf16 *xa = As + g * LDA + it * KI + warp * 128 + 8 * t;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ko = (q >> 1) * 64 + (q & 1) * 32;
            uint4 ag = *reinterpret_cast<const uint4 *>(w0 + ko * 2);
            uint4 x = g < p.M ? *reinterpret_cast<const uint4 *>(xa + ko) : make_uint4(0, 0, 0, 0);
            mma_bf16(a4[(q & 1) * 2], ag.x, 0u, ag.y, 0u, x.x, x.y);
            mma_bf16(a4[(q & 1) * 2 + 1], ag.z, 0u, ag.w, 0u, x.z, x.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cx.empty[s0]);
        cx.cnt += 1;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[0][e] = (a4[0][e] + a4[1][e]) + (a4[2][e] + a4[3][e]);
    red_write(cx.smem, 1, acc);
    return 1;
}

// ---- activation staging (compute threads only) ----------------------------------------------------
// bf16x2 LL words [M][K/2] -> As[m][k].  DPR = double-words per thread and row (K / 4 / 256): all index math
// is compile-time offsets from one base pointer.
template <int DPR>
PZ_DEVNOINL void stage_pairs(const Mega2Params &p, uint8_t *smem, const unsigned long long *buf, uint32_t flag,
                            const unsigned long long *sentinel) {
    cta_wait(p, sentinel, flag);
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM_AST);
    constexpr int KW = DPR * NCT * 2;           // words per row
    constexpr int RB = 8 / DPR;                 // rows per batch: 8 double-words in flight per thread
    const unsigned long long *src = buf + 2 * threadIdx.x;
    bf16 *dst = As + 4 * threadIdx.x;
    for (int m0 = 0; m0 < p.M; m0 += RB) {
        unsigned long long v[16];
        const int nrow = min(RB, p.M - m0);
        ll_gather<8>(p, flag, nrow * DPR, [&](int u) { return src + (long)(m0 + u / DPR) * KW + (u % DPR) * (2 * NCT); }, v);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (u < nrow * DPR)
                *reinterpret_cast<uint2 *>(dst + (m0 + u / DPR) * LDA + (u % DPR) * (4 * NCT)) = make_uint2((uint32_t)v[2 * u], (uint32_t)v[2 * u + 1]);
    }
    bar_compute();
}
// fp32 LL words [M][A] (A == 1024) -> Gemma RMSNorm (paligemma/modules.py:13-21) -> As; optionally
// captures the raw residual at columns [cap_n0, cap_n0 + 8) into cap[m][8]
PZ_DEVNOINL void stage_norm(const Mega2Params &p, uint8_t *smem, const unsigned long long *buf, const float *norm_w,
                          uint32_t flag, int cap_n0, float *cap, const unsigned long long *sentinel) {
    cta_wait(p, sentinel, flag);
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM_AST);
    float *misc = reinterpret_cast<float *>(smem + SM_MISC);
    float *part = misc + MISC_PART;
    const int tid = threadIdx.x, lane = tid & 31, c = tid & 63, rsub = tid >> 6;
    float4 w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = __ldg(reinterpret_cast<const float4 *>(norm_w) + c + 64 * j);
    for (int m0 = 0; m0 < p.M; m0 += 4) {
        const int m = m0 + rsub;
        const bool ok = m < p.M;
        unsigned long long v[16];
        // columns (c + 64 j) * 4 .. + 3: two double-words per j
        ll_gather<8>(p, flag, ok ? 8 : 0, [&](int u) { return buf + (long)m * KI + (c + 64 * (u >> 1)) * 4 + (u & 1) * 2; }, v);
        float x[16];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { x[i] = ok ? __uint_as_float((uint32_t)v[i]) : 0.f; ss += x[i] * x[i]; }
        ss = warp_sum(ss);
        if (lane == 0) part[rsub * 2 + ((tid >> 5) & 1)] = ss;
        bar_compute();
        const float r = rsqrtf((part[rsub * 2] + part[rsub * 2 + 1]) / KI + 1e-6f);
        if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = (c + 64 * j) * 4;
                uint2 o;
                o.x = pack_bf16x2(x[4 * j] * r * (1.f + w[j].x), x[4 * j + 1] * r * (1.f + w[j].y));
                o.y = pack_bf16x2(x[4 * j + 2] * r * (1.f + w[j].z), x[4 * j + 3] * r * (1.f + w[j].w));
                *reinterpret_cast<uint2 *>(As + m * LDA + col) = o;
                if (cap && (col >> 3) == (cap_n0 >> 3)) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) cap[m * 8 + (col & 7) + e] = x[4 * j + e];
                }
            }
        }
        bar_compute();
    }
}
// the two split-key attention partials -> combined attention output As[m][h*256 + d] = (o0 + o1) / (l0 + l1)
// (the partials are unnormalised sums of exp(logit) * v without a running maximum: the soft-cap bounds
// every logit to +-50, so exp() cannot overflow fp32 / bf16 -- joint_model.py:261-268)
PZ_DEVNOINL void stage_combine(const Mega2Params &p, uint8_t *smem, uint32_t flag, const unsigned long long *sentinel) {
    cta_wait(p, sentinel, flag);
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM_AST);
    float *inv = reinterpret_cast<float *>(smem + SM_MISC) + MISC_ML;   // [b][row]
    if (threadIdx.x < p.B * QROWS) {
        const int b = threadIdx.x / QROWS, row = threadIdx.x % QROWS;
        const unsigned long long *l0 = p.ll_l + (b * 2 + 0) * QROWS + row, *l1 = p.ll_l + (b * 2 + 1) * QROWS + row;
        float l = 0.f;
        if (row < p.nh * p.H) {
            ll_spin(p, l0, flag);
            ll_spin(p, l1, flag);
            l = __uint_as_float((uint32_t)ll_load1(l0)) + __uint_as_float((uint32_t)ll_load1(l1));
        }
        inv[threadIdx.x] = l > 0.f ? 1.f / l : 0.f;
    }
    bar_compute();
    // row m = (b, tok): nh * 128 pair words = nh * 64 double-words; thread handles double-words tid, tid + 256 (nh = 8)
    const int dpr = p.nh * 64 / NCT;           // 2 for 8 heads (checked by denoise_mega2_supported)
    for (int m = 0; m < p.M; m += 2) {
        unsigned long long v[16];
        const int nrow = min(2, p.M - m);
        // u = ((row * 2 + i) * 2 + split)
        ll_gather<8>(p, flag, nrow * dpr * 2, [&](int u) {
            const int mm = m + (u >> 2), i = (u >> 1) & 1, sp = u & 1;
            const int dd = threadIdx.x + i * NCT, h = dd >> 6, dp = (dd & 63) * 2;
            const int b = mm >> p.lgH, tok = mm & (p.H - 1);
            return p.ll_att + (((long)(b * 2 + sp) * QROWS + (h << p.lgH) + tok) << 7) + dp;
        }, v);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (r < nrow && i < dpr) {
                    const int mm = m + r, dd = threadIdx.x + i * NCT, h = dd >> 6;
                    const int b = mm >> p.lgH, tok = mm & (p.H - 1);
                    const float w = inv[b * QROWS + (h << p.lgH) + tok];
                    const int u0 = (r * 2 + i) * 2;
                    uint32_t a0 = (uint32_t)v[2 * u0], a1 = (uint32_t)v[2 * u0 + 1];           // split 0: two pair words
                    uint32_t b0 = (uint32_t)v[2 * u0 + 2], b1 = (uint32_t)v[2 * u0 + 3];       // split 1
                    uint2 o;
                    o.x = pack_bf16x2((bf16lo(a0) + bf16lo(b0)) * w, (bf16hi(a0) + bf16hi(b0)) * w);
                    o.y = pack_bf16x2((bf16lo(a1) + bf16lo(b1)) * w, (bf16hi(a1) + bf16hi(b1)) * w);
                    *reinterpret_cast<uint2 *>(As + mm * LDA + 4 * dd) = o;
                }
            }
        }
    }
    bar_compute();
}
// action -> linear_1 (vla/modules.py:39-41) -> As[m][0..A)
PZ_DEVNOINL void stage_enc1(const Mega2Params &p, uint8_t *smem, uint32_t flag, const unsigned long long *sentinel) {
    cta_wait(p, sentinel, flag);
    bf16 *As = reinterpret_cast<bf16 *>(smem + SM_AST);
    float *sact = reinterpret_cast<float *>(smem + SM_MISC) + MISC_ACT;   // [M][8], bf16-rounded like a GEMM input
    if (threadIdx.x < p.M * 8) {
        ll_spin(p, p.ll_act + threadIdx.x, flag);
        float a = __uint_as_float((uint32_t)ll_load1(p.ll_act + threadIdx.x));
        sact[threadIdx.x] = __bfloat162float(__float2bfloat16_rn(a));
    }
    bar_compute();
    for (int n = threadIdx.x; n < p.A; n += NCT) {
        uint4 wv = __ldg(reinterpret_cast<const uint4 *>(p.enc_w1 + (long)n * p.skp));   // skp >= 8
        uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
        float b1 = p.enc_b1[n];
        for (int m = 0; m < p.M; ++m) {
            float v = b1;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < p.action_dim) v += ((k & 1) ? bf16hi(ww[k >> 1]) : bf16lo(ww[k >> 1])) * sact[m * 8 + k];
            As[m * LDA + n] = __float2bfloat16_rn(v);
        }
    }
    bar_compute();
}

PZ_DEVINL void gemv_role(const Mega2Params &p, uint8_t *smem, int c) {
    GemvCtx cx;
    cx.smem = smem; cx.full = reinterpret_cast<uint64_t *>(smem + SM_BARS); cx.empty = cx.full + SLOTS;
    cx.cnt = 0; cx.c = c; cx.tr = false;
    for (int ph = 0; ph <= P_DEC; ++ph) cx.first[ph] = ph_first(p, ph, c);
    const int first_qkv = cx.first[P_QKV], first_gu = cx.first[P_GU];
    float *misc = reinterpret_cast<float *>(smem + SM_MISC);
    float *save_o = misc + MISC_SAVE_O, *save_d = misc + MISC_SAVE_D;
    const int tid = threadIdx.x;
    const int nA8 = p.A / 8;
    // this CTA's (single) 8-row block of the A-wide outputs, or -1
    const int blk8 = (ph_first(p, P_O, c) < nA8) ? ph_first(p, P_O, c) : -1;
    const bool is_dec = ph_first(p, P_DEC, c) == 0;
    // the fp32 action state lives in registers of the decoder CTA: thread i < M*8 owns element (i / 8, i % 8);
    // it starts as the caller's noise (pizero.py:454-458)
    float my_act = 0.f;
    if (is_dec && tid < p.M * 8) {
        int m = tid >> 3, a = tid & 7;
        my_act = a < p.action_dim ? p.noise[m * p.action_dim + a] : 0.f;
        ll_store(p.ll_act + tid, __float_as_uint(my_act), seq_flag(p, 0, IDX_ACT()));
    }

    // sentinel words of the exchanges this CTA consumes (see sentinel_block)
    const unsigned long long *sen_act = p.ll_act + c % (p.M * 8);
    const unsigned long long *sen_z = p.ll_z + sentinel_block(p, P_ENC2, c) * 4;
    const unsigned long long *sen_x1 = p.ll_x1 + sentinel_block(p, P_O, c) * 8;
    const unsigned long long *sen_x2 = p.ll_x2 + sentinel_block(p, P_D, c) * 8;
    const unsigned long long *sen_mlp = p.ll_mlp + sentinel_block(p, P_GU, c) * 4;
    const unsigned long long *sen_att = p.ll_l + ((c / (p.nh * p.H)) & 1) * QROWS + c % (p.nh * p.H);
    const int qkvw = (p.nh + 2) * 128;

    for (int step = 0; step < p.n_steps; ++step) {
        // ---- action encoder (vla/modules.py:39-53): linear_1 on the fly, linear_2 (action half) + per-step
        //      time bias + SiLU, linear_3 + sqrt(hidden) embed scale (joint_model.py:348-355)
        if (blk8 >= 0) {
            set_site(1 + 512 * (step & 3));
            stage_enc1(p, smem, seq_flag(p, step, IDX_ACT()), sen_act);
            gemv_acc8(p, cx, P_ENC2);
            bar_compute();
            if (tid < 4 * p.M) {
                int rp = tid & 3, m = tid >> 2, n = blk8 * 8 + 2 * rp;
                float v0 = red_sum(smem, 0, 2 * rp, m) + p.enc_time_bias[step * p.A + n];
                float v1 = red_sum(smem, 0, 2 * rp + 1, m) + p.enc_time_bias[step * p.A + n + 1];
                ll_store(p.ll_z + (long)m * (p.A / 2) + (n >> 1), pack_bf16x2(silu(v0), silu(v1)), seq_flag(p, step, IDX_Z()));
            }
            set_site(2 + 512 * (step & 3));
            stage_pairs<1>(p, smem, p.ll_z, seq_flag(p, step, IDX_Z()), sen_z);
            gemv_acc8(p, cx, P_ENC3);
            bar_compute();
            if (tid < 8 * p.M) {
                int r = tid & 7, m = tid >> 3, n = blk8 * 8 + r;
                float v = (red_sum(smem, 0, r, m) + p.enc_b3[n]) * sqrtf((float)p.A);
                ll_store(p.ll_x2 + (long)m * p.A + n, __float_as_uint(v), seq_flag(p, step, IDX_X0()));
            }
        }
        for (int l = 0; l < p.n_layers; ++l) {
            const pz_mix_layer &L = p.layers[l];
            const bool tr = p.trace && step == 1 && l == 1;
            cx.tr = tr;
            stamp(p, tr, 0);
            // ---- QKV: x -> RMSNorm -> fused q|k|v projection (mixture.py:187-215)
            {
                const uint32_t fin = seq_flag(p, step, l == 0 ? IDX_X0() : IDX_X2(l - 1));
                set_site(3 + 16 * l + 512 * (step & 3));
                stage_norm(p, smem, p.ll_x2, L.norm_in, fin, blk8 >= 0 ? blk8 * 8 : -8, blk8 >= 0 ? save_o : nullptr, sen_x2);
                stamp(p, tr, 1);
                const int nj = gemv_acc16<2>(p, cx, P_QKV);
                bar_compute();
                stamp(p, tr, 2);
                const uint32_t fo = seq_flag(p, step, IDX_QKV(l));
                // one (block, m, row) sum per thread.  Rows 0-7 / 8-15 of a block are the dims d / d + 128 of one
                // head: rotate q and k here (fp32, table row S_p + token: positions 2.., pizero.py:312-318),
                // then neighbouring dims pair up through a shuffle
                for (int i0 = 0; i0 < nj * 16 * p.M; i0 += NCT) {
                    const int i = i0 + tid;
                    const bool ok = i < nj * 16 * p.M;
                    int r = i & 15, m = i >> 4, j = 0;
                    while (m >= p.M) { m -= p.M; ++j; }
                    if (!ok) { j = 0; m = 0; }
                    float v0 = ok ? red_sum(smem, j, r, m) : 0.f;
                    const float other = __shfl_xor_sync(0xffffffffu, v0, 8);
                    const int blk = first_qkv + j * p.G, hh = blk >> 4, d = (blk & 15) * 8 + (r & 7);
                    if (hh <= p.nh) {
                        const long ti = (long)(p.S_p + (m & (p.H - 1))) * 128 + d;
                        const float cs = __ldg(p.rope_cos + ti), sn = __ldg(p.rope_sin + ti);
                        v0 = (r < 8) ? v0 * cs - other * sn : v0 * cs + other * sn;
                    }
                    const float v1 = __shfl_down_sync(0xffffffffu, v0, 1);
                    const int n = hh * 256 + (r < 8 ? d : 128 + d);
                    if (ok && !(r & 1)) ll_store(p.ll_qkv + (long)m * qkvw + (n >> 1), pack_bf16x2(v0, v1), fo);
                }
                stamp(p, tr, 3);
            }
            // ---- o_proj + residual (mixture.py:217-218, joint_model.py:65-75)
            if (blk8 >= 0) {
                set_site(4 + 16 * l + 512 * (step & 3));
                stage_combine(p, smem, seq_flag(p, step, IDX_ATT(l)), sen_att);
                stamp(p, tr, 4);
                gemv_acc8(p, cx, P_O);
                bar_compute();
                stamp(p, tr, 5);
                if (tid < 8 * p.M) {
                    int r = tid & 7, m = tid >> 3;
                    float v = save_o[m * 8 + r] + red_sum(smem, 0, r, m);
                    ll_store(p.ll_x1 + (long)m * p.A + blk8 * 8 + r, __float_as_uint(v), seq_flag(p, step, IDX_X1(l)));
                }
                stamp(p, tr, 6);
            }
            // ---- gate|up + GeGLU (paligemma/modules.py:86-95)
            {
                set_site(5 + 16 * l + 512 * (step & 3));
                stage_norm(p, smem, p.ll_x1, L.norm_post, seq_flag(p, step, IDX_X1(l)), blk8 >= 0 ? blk8 * 8 : -8,
                           blk8 >= 0 ? save_d : nullptr, sen_x1);
                stamp(p, tr, 7);
                const int nj = gemv_acc16<4>(p, cx, P_GU);
                bar_compute();
                stamp(p, tr, 8);
                const uint32_t fo = seq_flag(p, step, IDX_MLP(l));
                for (int i0 = 0; i0 < nj * 16 * p.M; i0 += NCT) {
                    const int i = i0 + tid;
                    const bool ok = i < nj * 16 * p.M;
                    int r = i & 15, m = i >> 4, j = 0;
                    while (m >= p.M) { m -= p.M; ++j; }
                    if (!ok) { j = 0; m = 0; }
                    float v = ok ? red_sum(smem, j, r, m) : 0.f;          // rows 0-7: gate, rows 8-15: the matching up rows
                    float u = __shfl_down_sync(0xffffffffu, v, 8);
                    float h0 = gelu_fast(v) * u;
                    float h1 = __shfl_down_sync(0xffffffffu, h0, 1);
                    int blk = first_gu + j * p.G;
                    if (ok && r < 8 && !(r & 1)) ll_store(p.ll_mlp + (long)m * (p.AI / 2) + blk * 4 + (r >> 1), pack_bf16x2(h0, h1), fo);
                }
                stamp(p, tr, 9);
            }
            // ---- down + residual
            if (blk8 >= 0) {
                set_site(6 + 16 * l + 512 * (step & 3));
                stage_pairs<4>(p, smem, p.ll_mlp, seq_flag(p, step, IDX_MLP(l)), sen_mlp);
                stamp(p, tr, 10);
                gemv_acc8(p, cx, P_D);
                bar_compute();
                stamp(p, tr, 11);
                if (tid < 8 * p.M) {
                    int r = tid & 7, m = tid >> 3;
                    float v = save_d[m * 8 + r] + red_sum(smem, 0, r, m);
                    ll_store(p.ll_x2 + (long)m * p.A + blk8 * 8 + r, __float_as_uint(v), seq_flag(p, step, IDX_X2(l)));
                }
                stamp(p, tr, 12);
            }
        }
        // ---- final norm + decoder + Euler update (joint_model.py:375-380, pizero.py:479-481)
        if (is_dec) {
            set_site(7 + 512 * (step & 3));
            stage_norm(p, smem, p.ll_x2, p.final_norm, seq_flag(p, step, IDX_X2(p.n_layers - 1)), -8, nullptr, sen_x2);
            gemv_acc8(p, cx, P_DEC);
            bar_compute();
            if (tid < 8 * p.M) {
                int a = tid & 7, m = tid >> 3;
                if (a < p.action_dim) my_act += p.dt * (red_sum(smem, 0, a, m) + p.dec_b[a]);
                if (step + 1 < p.n_steps) {
                    ll_store(p.ll_act + tid, __float_as_uint(my_act), seq_flag(p, step + 1, IDX_ACT()));
                } else if (a < p.action_dim) {
                    float v = my_act;
                    if (p.clip >= 0.f) v = fminf(fmaxf(v, -p.clip), p.clip);
                    p.out[m * p.action_dim + a] = v;
                }
            }
        }
        bar_compute();   // the reduction scratch / staged activations are rewritten by the next step
    }
}

// ========================================== attention role ==========================================
// Stage this step's (already rotated) q rows of sample b and, if this CTA's key range holds them, the fresh
// action k / v rows, straight from the QKV exchange words into their shared-memory tiles.
PZ_DEVNOINL void att_stage(const Mega2Params &p, uint8_t *smem, int b, int key0, uint32_t fin, const unsigned long long *sentinel) {
    cta_wait(p, sentinel, fin);
    bf16 *sQ = reinterpret_cast<bf16 *>(smem + SA_Q);
    bf16 *sK = reinterpret_cast<bf16 *>(smem + SA_K);
    bf16 *sV = reinterpret_cast<bf16 *>(smem + SA_V);
    const int tid = threadIdx.x;
    const int qkvw = (p.nh + 2) * 128;
    const int r_fresh = p.S_c - key0;                         // local row of the first fresh key
    const bool fresh = r_fresh + p.H > 0 && r_fresh < p.ks;   // (all H fresh rows are in range then: checked on the host)
    // per token row: nh*64 double-words of q (2 per thread for 8 heads), then 128 double-words of k | v
    const int qpt = p.nh * 64 / NCT;
    for (int tok = 0; tok < p.H; tok += 2) {
        unsigned long long v[12];
        const int nrow = min(2, p.H - tok);
        const int per_row = qpt + ((fresh && tid < 128) ? 1 : 0);
        // u = row * 3 + i  (i < qpt: q double-word tid + i*256; i == 2: k|v double-word nh*64 + tid)
        ll_gather<6>(p, fin, nrow * 3, [&](int u) {
            const int rr = u / 3, i = u % 3;
            const int dd = (i < 2) ? tid + (i < qpt ? i : 0) * NCT : p.nh * 64 + (per_row > qpt ? tid : 0);
            return p.ll_qkv + (long)(b * p.H + tok + rr) * qkvw + 2 * dd;
        }, v);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            if (rr >= nrow) break;
            const int tk = tok + rr;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i < qpt) {
                    const int dd = tid + i * NCT, h = dd >> 6, d = (dd & 63) * 4;
                    *reinterpret_cast<uint2 *>(sQ + ((h << p.lgH) + tk) * LDQ + d) =
                        make_uint2((uint32_t)v[2 * (rr * 3 + i)], (uint32_t)v[2 * (rr * 3 + i) + 1]);
                }
            }
            if (per_row > qpt) {
                const int d = (tid & 63) * 4;
                bf16 *dst = (tid < 64 ? sK : sV) + (r_fresh + tk) * LDQ + d;
                *reinterpret_cast<uint2 *>(dst) = make_uint2((uint32_t)v[2 * (rr * 3 + 2)], (uint32_t)v[2 * (rr * 3 + 2) + 1]);
            }
        }
    }
}

PZ_DEVINL void att_producer(const Mega2Params &p, uint8_t *smem, int b, int split) {
    const int lane = threadIdx.x & 31;
    uint64_t *kv_full = reinterpret_cast<uint64_t *>(smem + SA_BARS), *kv_empty = kv_full + 1;
    const uint64_t pol = policy_evict_last();
    const int key0 = split * p.ks;
    const int n_cached = max(0, min(p.ks, p.S_c - key0));   // cached rows of this CTA's key range
    uint32_t it = 0;
    for (int step = 0; step < p.n_steps; ++step) {
        for (int l = 0; l < p.n_layers; ++l, ++it) {
            if (it > 0) mbar_wait(p, kv_empty, (it - 1) & 1);
            if (lane == 0) mbar_expect_tx(kv_full, (uint32_t)n_cached * 1024u);
            __syncwarp();
            const bf16 *Kc = p.kcache + (long)l * p.kv_layer_stride + (long)b * p.kv_batch_stride + (long)key0 * 256;
            const bf16 *Vc = p.vcache + (long)l * p.kv_layer_stride + (long)b * p.kv_batch_stride + (long)key0 * 256;
            for (int r = lane; r < n_cached; r += 32) {
                bulk_g2s(smem + SA_K + r * LDQ * 2, Kc + (long)r * 256, 512, kv_full, pol);
                bulk_g2s(smem + SA_V + r * LDQ * 2, Vc + (long)r * 256, 512, kv_full, pol);
            }
        }
    }
}

PZ_DEVINL void att_role(const Mega2Params &p, uint8_t *smem, int b, int split) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    bf16 *sQ = reinterpret_cast<bf16 *>(smem + SA_Q);
    bf16 *sK = reinterpret_cast<bf16 *>(smem + SA_K);
    bf16 *sV = reinterpret_cast<bf16 *>(smem + SA_V);
    bf16 *sP = reinterpret_cast<bf16 *>(smem + SA_P);
    float(*sRS)[QROWS] = reinterpret_cast<float(*)[QROWS]>(smem + SA_RS);   // [4 key quarters][row] partial row sums
    uint64_t *kv_full = reinterpret_cast<uint64_t *>(smem + SA_BARS), *kv_empty = kv_full + 1;
    const int rows_total = p.nh * p.H;
    const int vlen = p.valid_len[b];
    const int n_keys = p.S_c + p.H;
    const int key0 = split * p.ks;
    const int n_local = max(0, min(p.ks, n_keys - key0));       // keys of this CTA (cached + fresh)
    const int n_cached = max(0, min(p.ks, p.S_c - key0));

    // rows that no copy ever writes must be finite (P = 0 times garbage must stay 0); Q pad rows are zero
    for (int i = tid; i < (KS - n_cached) * 32; i += NCT) {
        int r = n_cached + (i >> 5), cc = i & 31;
        *reinterpret_cast<uint4 *>(sK + r * LDQ + cc * 8) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(sV + r * LDQ + cc * 8) = make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < QROWS * 32; i += NCT)
        *reinterpret_cast<uint4 *>(sQ + (i >> 5) * LDQ + (i & 31) * 8) = make_uint4(0, 0, 0, 0);
    // sentinel of the QKV exchange: a word of a last-round block (see sentinel_block) of this sample's first row
    const int sb = sentinel_block(p, P_QKV, blockIdx.x);
    const unsigned long long *sen_qkv = p.ll_qkv + (long)b * p.H * ((p.nh + 2) * 128) + (((sb >> 4) * 256 + (sb & 15) * 8) >> 1);
    bar_compute();

    // S = Q K^T work split: warp -> 16-row query tile mt, key 8-tiles kq, kq + 4, ... (18 tiles -> 5 or 4 per warp)
    const int mt = warp & 1, kq = warp >> 1;
    const bool fifth = kq + 16 < KS / 8;

    uint32_t it = 0;
    for (int step = 0; step < p.n_steps; ++step) {
        for (int l = 0; l < p.n_layers; ++l, ++it) {
            const bool tr = p.trace && step == 1 && l == 1;
            stamp(p, tr, 0);
            set_site(8 + 16 * l + 512 * (step & 3));
            att_stage(p, smem, b, key0, seq_flag(p, step, IDX_QKV(l)), sen_qkv);
            stamp(p, tr, 1);
            mbar_wait(p, kv_full, it & 1);   // the cached rows of this layer (loaded one layer ahead)
            bar_compute();
            stamp(p, tr, 2);

            // ---- S = Q K^T, soft-cap, mask, exp -> P (bf16) + partial row sums
            {
                float s[5][4];
#pragma unroll
                for (int a = 0; a < 5; ++a) s[a][0] = s[a][1] = s[a][2] = s[a][3] = 0.f;
#pragma unroll 4
                for (int ks = 0; ks < 16; ++ks) {
                    uint32_t q0[4], k01[4], k23[4], k4[2];
                    ldsm_x4(q0, sQ + (mt * 16 + (lane & 15)) * LDQ + ks * 16 + (lane >> 4) * 8);
                    // x4: tile A (d 0-7, d 8-15), tile A + 4 (d 0-7, d 8-15)
                    ldsm_x4(k01, sK + ((kq + (lane >> 4) * 4) * 8 + (lane & 7)) * LDQ + ks * 16 + ((lane >> 3) & 1) * 8);
                    ldsm_x4(k23, sK + ((kq + 8 + (lane >> 4) * 4) * 8 + (lane & 7)) * LDQ + ks * 16 + ((lane >> 3) & 1) * 8);
                    mma_bf16(s[0], q0[0], q0[1], q0[2], q0[3], k01[0], k01[1]);
                    mma_bf16(s[1], q0[0], q0[1], q0[2], q0[3], k01[2], k01[3]);
                    mma_bf16(s[2], q0[0], q0[1], q0[2], q0[3], k23[0], k23[1]);
                    mma_bf16(s[3], q0[0], q0[1], q0[2], q0[3], k23[2], k23[3]);
                    if (fifth) {
                        ldsm_x2(k4, sK + ((kq + 16) * 8 + (lane & 7)) * LDQ + ks * 16 + ((lane >> 3) & 1) * 8);
                        mma_bf16(s[4], q0[0], q0[1], q0[2], q0[3], k4[0], k4[1]);
                    }
                }
                const float scale = 0.0625f, cap = 50.f;   // 1/sqrt(256); soft-cap (joint_model.py:139,261-268)
                float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    if (a == 4 && !fifth) break;
                    const int col = (kq + 4 * a) * 8 + 2 * t;
                    float pe[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int cc = col + (e & 1), j = key0 + cc;
                        const bool vis = cc < n_local && ((j < vlen) || (j >= p.S_v && j < n_keys));
                        // |logit| <= 50 after the soft-cap: exp() needs no running maximum
                        pe[e] = vis ? __expf(tanh_fast_acc(s[a][e] * scale * (1.f / cap)) * cap) : 0.f;
                    }
                    rs0 += pe[0] + pe[1];
                    rs1 += pe[2] + pe[3];
                    *reinterpret_cast<uint32_t *>(sP + (mt * 16 + g) * LDP + col) = pack_bf16x2(pe[0], pe[1]);
                    *reinterpret_cast<uint32_t *>(sP + (mt * 16 + g + 8) * LDP + col) = pack_bf16x2(pe[2], pe[3]);
                }
                rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
                rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
                if (t == 0) { sRS[kq][mt * 16 + g] = rs0; sRS[kq][mt * 16 + g + 8] = rs1; }
            }
            bar_compute();
            stamp(p, tr, 3);
            // ---- O = P V: warp -> (16-row tile mt, 64-wide slice of d); publish the unnormalised partial
            {
                const int dq = warp >> 1;
                float o[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll 3
                for (int kk = 0; kk < KS / 16; ++kk) {
                    uint32_t pa[4];
                    ldsm_x4(pa, sP + (mt * 16 + (lane & 15)) * LDP + kk * 16 + (lane >> 4) * 8);
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {
                        uint32_t vb[4];
                        ldsm_x4_t(vb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDQ + dq * 64 + dp * 16 + (lane >> 4) * 8);
                        mma_bf16(o[2 * dp], pa[0], pa[1], pa[2], pa[3], vb[0], vb[1]);
                        mma_bf16(o[2 * dp + 1], pa[0], pa[1], pa[2], pa[3], vb[2], vb[3]);
                    }
                }
                const uint32_t fo = seq_flag(p, step, IDX_ATT(l));
                unsigned long long *obase = p.ll_att + ((long)(b * 2 + split) * QROWS << 7);
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    int row = mt * 16 + g + rr * 8;
                    if (row >= rows_total) continue;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        ll_store(obase + ((long)row << 7) + ((dq * 64 + i * 8 + 2 * t) >> 1), pack_bf16x2(o[i][2 * rr], o[i][2 * rr + 1]), fo);
                    if (dq == 0 && t == 0)
                        ll_store(p.ll_l + (b * 2 + split) * QROWS + row,
                                 __float_as_uint((sRS[0][row] + sRS[1][row]) + (sRS[2][row] + sRS[3][row])), fo);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(kv_empty);   // this warp is done with the layer's K / V
            stamp(p, tr, 4);
            bar_compute();                           // Q / P / row sums are rewritten by the next layer
        }
    }
}

__global__ void __launch_bounds__(NT2, 1) denoise_mega2_kernel(const __grid_constant__ Mega2Params p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int warp = threadIdx.x >> 5;
    const bool is_att = (int)blockIdx.x >= p.G;
    if (threadIdx.x == 0) {
        if (is_att) {
            uint64_t *kv_full = reinterpret_cast<uint64_t *>(smem + SA_BARS);
            mbar_init(kv_full, 1);
            mbar_init(kv_full + 1, NCW);
        } else {
            uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM_BARS);
            for (int i = 0; i < SLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&full[SLOTS + i], NCW); }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (is_att) {
        const int a = blockIdx.x - p.G, b = a >> 1, split = a & 1;
        if (warp == NCW) att_producer(p, smem, b, split);
        else att_role(p, smem, b, split);
    } else {
        if (warp == NCW) producer_loop(p, smem, blockIdx.x);
        else gemv_role(p, smem, blockIdx.x);
    }
}

}  // namespace

// host side --------------------------------------------------------------------------------------------
size_t denoise_mega2_ll_bytes(const pz_config &c, int B) {
    size_t M = (size_t)B * c.horizon, qkvd = (size_t)(c.n_heads + 2) * c.head_dim;
    size_t words = M * 8 + M * c.act_hidden / 2 + 2 * M * c.act_hidden + M * qkvd / 2 + (size_t)B * 2 * QROWS * 128 +
                   (size_t)B * 2 * QROWS * 2 + M * c.act_inter / 2;
    return words * 8 + 8 * 128 /* per-buffer 128-byte alignment */ + 256 /* error flag */ + 32768 /* trace */;
}

int denoise_mega2_supported(const pz_config &c, int B) {
    // PZ_MEGA: 0 separate kernels, 1 (default) denoise_mega.cu, 2 this kernel.  Opt-in: measured on B200 it does
    // not beat the grid-barrier kernel yet (6.05 vs 5.64 ms at bs=1 under graph replay)
    const char *e = getenv("PZ_MEGA");
    if (!e || atoi(e) != 2) return 0;
    if (c.dtype != PZ_BF16 || (c.flags & PZ_FLAG_SIMPLE_KERNELS)) return 0;
    if (B * c.horizon > MAXM || c.n_heads * c.horizon > QROWS || c.horizon > 4) return 0;
    if (c.head_dim != 256 || c.n_kv_heads != 1 || c.n_heads > 8) return 0;
    if (c.act_hidden != KI || c.act_inter % KI || c.act_inter > KMAX || (c.n_heads * 256) % KI || c.n_heads * 256 > KMAX) return 0;
    if (c.n_layers > MAX_LAYERS || c.action_dim > 8) return 0;
    if (c.s_vlm + c.cond_steps + c.horizon > 2 * KS) return 0;
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (!pow2(c.horizon) || c.n_heads != 8 || c.act_inter != KMAX) return 0;   // compile-time index math of the staging loops
    {   // the fresh (action) keys must not straddle the two attention CTAs of a sample
        int n_keys = c.s_vlm + c.cond_steps + c.horizon, ks = ((n_keys + 1) / 2 + 15) / 16 * 16;
        if (c.s_vlm + c.cond_steps < ks) return 0;
    }
    return 1;
}

int launch_denoise_mega2(const pz_config &c, const pz_weights &w, const pz_mix_layer *layers, const Mega2Buffers &bf,
                         int B, cudaStream_t st, const char **err) {
    static int num_sms = 0;
    static bool attr_set = false;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (!attr_set) {
        if (cudaFuncSetAttribute(denoise_mega2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_TOTAL) != cudaSuccess) {
            if (err) *err = "denoise_mega2: cannot set the shared-memory size";
            return PZ_ERR_CUDA;
        }
        attr_set = true;
    }
    Mega2Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = c.horizon; p.M = B * c.horizon; p.A = c.act_hidden; p.AI = c.act_inter; p.nh = c.n_heads;
    p.S_v = c.s_vlm; p.S_p = c.cond_steps; p.S_c = c.s_vlm + c.cond_steps; p.n_layers = c.n_layers; p.n_steps = c.n_steps;
    p.action_dim = c.action_dim; p.skp = w.small_k_pad;
    p.G = num_sms - 2 * B;
    p.lgH = 0;
    while ((1 << p.lgH) < p.H) ++p.lgH;
    {   // keys per attention CTA: half of the keys, rounded up to the 16-key MMA granularity
        int n_keys = p.S_c + p.H;
        p.ks = ((n_keys + 1) / 2 + 15) / 16 * 16;
        if (p.ks > KS) p.ks = KS;
    }
    if (p.G < c.act_hidden / 8) {
        if (err) *err = "denoise_mega2: too few SMs";
        return PZ_ERR_INVALID;
    }
    p.dt = (float)(1.0 / c.n_steps); p.clip = c.clip;
    for (int l = 0; l < c.n_layers; ++l) p.layers[l] = layers[l];
    p.final_norm = w.action_final_norm;
    p.enc_w1 = (const bf16 *)w.enc_w1; p.enc_w2a = (const bf16 *)w.enc_w2a; p.enc_w3 = (const bf16 *)w.enc_w3;
    p.dec_w = (const bf16 *)w.dec_w;
    p.enc_b1 = w.enc_b1; p.enc_time_bias = w.enc_time_bias; p.enc_b3 = w.enc_b3; p.dec_b = w.dec_b;
    p.rope_cos = w.rope_act_cos; p.rope_sin = w.rope_act_sin;
    p.kcache = (const bf16 *)bf.kcache; p.vcache = (const bf16 *)bf.vcache;
    p.kv_batch_stride = (long)p.S_c * 256; p.kv_layer_stride = (long)bf.batch_total * p.kv_batch_stride;
    p.valid_len = bf.valid_len; p.noise = bf.noise; p.out = bf.out;
    {   // carve the LL buffers
        char *base = (char *)bf.ll;
        size_t off = 0;
        auto take = [&](size_t words) {
            off = (off + 127) & ~(size_t)127;
            unsigned long long *q = (unsigned long long *)(base + off);
            off += words * 8;
            return q;
        };
        size_t M = p.M, qkvd = (size_t)(p.nh + 2) * 256;
        p.ll_act = take(M * 8); p.ll_z = take(M * p.A / 2); p.ll_x1 = take(M * p.A); p.ll_x2 = take(M * p.A);
        p.ll_qkv = take(M * qkvd / 2); p.ll_att = take((size_t)B * 2 * QROWS * 128); p.ll_l = take((size_t)B * 2 * QROWS * 2);
        p.ll_mlp = take(M * p.AI / 2);
        off = (off + 127) & ~(size_t)127;
        p.err = (unsigned int *)(base + bf.ll_bytes - 32768 - 256);   // fixed place (tools read it): just below the trace
        off += 256;
        p.trace = (unsigned long long *)(base + bf.ll_bytes - 32768);
        if (off + 32768 > bf.ll_bytes) {
            if (err) *err = "denoise_mega2: LL workspace too small";
            return PZ_ERR_WORKSPACE;
        }
        if (cudaMemsetAsync(bf.ll, 0, bf.ll_bytes - 32768, st) != cudaSuccess) {
            if (err) *err = "denoise_mega2: memset failed";
            return PZ_ERR_CUDA;
        }
    }
    void *args[] = {&p};
    // cooperative launch: all CTAs are guaranteed co-resident (the polled exchanges rely on it)
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)denoise_mega2_kernel, dim3(num_sms), dim3(NT2), args,
                                                (size_t)SMEM2_TOTAL, st);
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        return PZ_ERR_CUDA;
    }
    count_launch();
    return 0;
}
