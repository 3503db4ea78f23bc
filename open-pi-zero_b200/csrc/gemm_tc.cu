// gemm_tc.cu -- bf16 linear layers on the 5th-generation tensor cores.
//
//   C[M,N] = alpha * epi(A[M,K] . W[N,K]^T + bias)      A, W bf16 K-major
//
// Blackwell-native structure (one persistent CTA per SM):
//   warp 0  : TMA producer  -- cp.async.bulk.tensor 2D loads of a 128 x 64 A tile
//             and a BN x 64 W tile (128-byte swizzle) into a ring of smem stages,
//             completion signalled on mbarriers (complete_tx).
//   warp 1  : MMA issuer    -- one thread issues tcgen05.mma.cta_group::1.kind::f16
//             (128 x BN x 16 per instruction) reading smem through UMMA descriptors,
//             accumulating fp32 in TMEM; tcgen05.commit frees smem stages and
//             publishes finished accumulators.
//   warp 2  : TMEM allocator (tcgen05.alloc / dealloc).
//   warps 4-7: epilogue     -- tcgen05.ld the accumulator (lane = row), apply
//             bias / GELU / SiLU / GeGLU / alpha, write bf16 or accumulate into
//             the fp32 residual stream.  Two accumulator stages in TMEM let the
//             epilogue of tile i overlap the MMAs of tile i+1.
//
// Used for every projection / MLP of the SigLIP encoder and the Gemma prefix
// pass (reference call sites: siglip.py:115-119,164,188-192; mixture.py:187-218;
// paligemma/modules.py:86-95), and for the denoise layers when B*horizon > 64.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;

struct TcParams {
    int M, N, K, ldc;
    const float *bias;
    void *C;
    float alpha;
    int flags;
    int tiles_m, tiles_n;
    int ksplit, kb_per_split;           // split-K (fp32 reduce-add outputs only)
    int n_fastest;                      // tile raster: 1 = consecutive CTAs share the A tile (few N tiles, big A)
    uint64_t a_hint, w_hint;            // L2 eviction priority of the A / W tile loads
    int group_m;                        // M-fastest raster in groups of group_m M tiles (0 = one group): the A rows of a
                                        // group stay L2 resident while the N tiles sweep past them
    // fused RoPE + Q/K/V split (LIN_ROPE): BN = 256 = head_dim, one N tile per head
    const float *rope_cos, *rope_sin;   // [s_x, 128] fp32
    bf16 *k_out, *v_out;                // cache rows: base + b*kv_batch_stride + s*256
    long kv_batch_stride;
    int s_x, n_q_tiles;
    int a_mn, w_mn;                     // operand stored [K][M] / [K][N]: loaded as 64-element x 64-row boxes, MN-major descriptors
};

constexpr int LIN_ROPE = 1 << 10;       // internal flag of this file

constexpr int EPI_WARPS = 8;            // two per TMEM lane quadrant
constexpr int NUM_THREADS2 = (EPI_WARP0 + EPI_WARPS) * 32;
constexpr int STG_BYTES = 4096;         // per-warp staging: 32 rows x 128 B

// CG = 2: a CTA pair (cluster of 2, one TPC) computes a 256 x 256 tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and HALF of the W tile, so a stage is 32 KB instead of 48 KB:
// six stages fit, i.e. 50 % more K in flight per SM and a third less shared-memory / L2 traffic per FLOP.
template <int BN, int CG = 1> struct Cfg {
    // CTA pairs (the big-M, compute-bound configuration) double-buffer the epilogue staging so that a warp never
    // waits for its previous TMA store to drain (64 KB of staging: one mainloop stage less)
    static constexpr int STG_BUFS = CG == 2 ? 2 : 1;
    static constexpr int STAGES = CG == 2 ? 5 : (BN != 256 ? 6 : 4);
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (BN / CG) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;   // two accumulator stages (power of two: 256 / 512)
    static constexpr int STG_OFF = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = STG_OFF + EPI_WARPS * STG_BYTES * STG_BUFS;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 /*align*/ + 256 /*barriers*/;
};

PZ_DEVINL float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// gelu-tanh with the hardware tanh (rel. error 2^-11, far below the bf16 rounding of the output)
PZ_DEVINL float gelu_fast(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    return 0.5f * x * (1.0f + tanh_fast(k0 * (x + k1 * x * x * x)));
}

// EPI: the epilogue flavour is a compile-time parameter -- with run-time flag tests inside the fully unrolled
// per-element loops the epilogue cost ~11 instructions per output element and, for the short-K SigLIP GEMMs
// (7 us of MMA per 256 x 256 tile), became the bottleneck (tensor pipe 40-52 % active; profiles/).
enum { E_PLAIN = 0, E_GELU = 1, E_SILU = 2, E_F32 = 3, E_GEGLU = 4, E_ROPE = 5 };

// tile raster (the workers that run together take consecutive tile indices)
PZ_DEVINL void tile_coords(const TcParams &p, int tmn, int &tm, int &tn) {
    if (p.n_fastest) {
        tm = tmn / p.tiles_n; tn = tmn % p.tiles_n;
    } else if (p.group_m > 0) {
        const int per_group = p.group_m * p.tiles_n;
        const int g = tmn / per_group, r = tmn - g * per_group;
        const int gm0 = g * p.group_m, gsz = min(p.group_m, p.tiles_m - gm0);
        tm = gm0 + r % gsz; tn = r / gsz;
    } else {
        tm = tmn % p.tiles_m; tn = tmn / p.tiles_m;
    }
}

// MN: operand storage of this instantiation -- 0: A [M][K], W [N][K] (every inference GEMM; the MMA-issue loop stays free
// of run-time selects: measured 10 % on the gate|up GEMM when the choice was a kernel parameter), 1: W stored [K][N],
// 3: A stored [K][M] and W stored [K][N] (the two products of the training step's backward)
template <int BN, int CG, int EPI, int MN = 0>
__global__ void __launch_bounds__(NUM_THREADS2, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_c, const TcParams p) {
    using cfg = Cfg<BN, CG>;
    // CTA pair: `crank` 0 is the leader (issues the MMAs); `wid` / `nworkers` enumerate tile workers
    // (single CTAs for CG = 1, CTA pairs for CG = 2)
    const int crank = CG == 2 ? (int)cluster_ctarank() : 0;
    const int wid = blockIdx.x / CG, nworkers = gridDim.x / CG;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = align_smem(smem_raw, 1024);
    uint64_t *full_bar = (uint64_t *)(smem + cfg::BAR_OFF);
    uint64_t *empty_bar = full_bar + cfg::STAGES;
    uint64_t *tfull_bar = empty_bar + cfg::STAGES;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + BK - 1) / BK;
    const int tiles_mn = p.tiles_m * p.tiles_n;
    const int num_tiles = tiles_mn * p.ksplit;   // tile t -> (t % tiles_mn, k-split t / tiles_mn)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < cfg::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], EPI_WARPS * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)cfg::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)cfg::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ----
        if (lane == 0) {
            // one operand tile of a stage: K-major = one box [rows][64 k]; MN-major (operand stored [K][M/N]) = one box
            // [64 k][64 m] per 64 rows of the tile, 8 KB apart (the LBO of the MN-major descriptor)
            auto load_w = [&](uint64_t *bar, uint8_t *dst, int kb, int tn) {
                const int n0 = tn * BN + (CG == 2 ? crank * (BN / 2) : 0);
                if constexpr (!(MN & 1)) {
                    if (CG == 2) tma_load_2d_pair(&map_w, bar, dst, kb * BK, n0, p.w_hint);
                    else tma_load_2d(&map_w, bar, dst, kb * BK, n0, p.w_hint);
                } else {
                    for (int j = 0; j < BN / CG / 64; ++j) {
                        if (CG == 2) tma_load_2d_pair(&map_w, bar, dst + j * 8192, n0 + 64 * j, kb * BK, p.w_hint);
                        else tma_load_2d(&map_w, bar, dst + j * 8192, n0 + 64 * j, kb * BK, p.w_hint);
                    }
                }
            };
            auto load_a = [&](uint64_t *bar, uint8_t *dst, int kb, int tm) {
                const int m0 = (tm * CG + crank) * BM;
                if constexpr (!(MN & 2)) {
                    if (CG == 2) tma_load_2d_pair(&map_a, bar, dst, kb * BK, m0, p.a_hint);
                    else tma_load_2d(&map_a, bar, dst, kb * BK, m0, p.a_hint);
                } else {
                    for (int j = 0; j < BM / 64; ++j) {
                        if (CG == 2) tma_load_2d_pair(&map_a, bar, dst + j * 8192, m0 + 64 * j, kb * BK, p.a_hint);
                        else tma_load_2d(&map_a, bar, dst + j * 8192, m0 + 64 * j, kb * BK, p.a_hint);
                    }
                }
            };
            // PDL: the weight tiles of the first ring pass do not depend on the previous
            // kernel -- issue them before waiting for it; activations (A) only after.
            int pre = 0;
            if (wid < num_tiles) {
                int t = wid;
                int tmn = t % tiles_mn, ks = t / tiles_mn;
                int tm_unused, tn;
                tile_coords(p, tmn, tm_unused, tn);
                int kb0 = ks * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
                pre = min(cfg::STAGES, kb1 - kb0);
                for (int i = 0; i < pre; ++i) {
                    uint8_t *sa = smem + i * cfg::STAGE_BYTES;
                    if (crank == 0) mbar_expect_tx(&full_bar[i], cfg::STAGE_BYTES * CG);
                    load_w(&full_bar[i], sa + cfg::A_BYTES, kb0 + i, tn);
                }
            }
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = wid; t < num_tiles; t += nworkers) {
                int tmn = t % tiles_mn, ks = t / tiles_mn;
                // raster: m fastest -> concurrent CTAs share the W tile in L2; n fastest -> they share the A tile
                int tm, tn;
                tile_coords(p, tmn, tm, tn);
                int kb0 = ks * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    uint8_t *sa = smem + stage * cfg::STAGE_BYTES;
                    if (it >= pre) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (crank == 0) mbar_expect_tx(&full_bar[stage], cfg::STAGE_BYTES * CG);
                        load_w(&full_bar[stage], sa + cfg::A_BYTES, kb, tn);
                    }
                    load_a(&full_bar[stage], sa, kb, tm);
                    if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------- MMA issuer ----
        if (lane == 0 && crank == 0) {
            // bits 15 / 16 of the instruction descriptor: A / B operand MN-major
            constexpr uint32_t idesc = umma_idesc(BM * CG, BN) | ((MN & 2) ? (1u << 15) : 0u) | ((MN & 1) ? (1u << 16) : 0u);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = wid; t < num_tiles; t += nworkers) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                uint32_t d_tmem = tmem_base + acc * BN;
                int ks = t / tiles_mn;
                int kb0 = ks * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    uint32_t sa = smem_u32(smem + stage * cfg::STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(sa);
                    const uint64_t bdesc = umma_desc_sw128(sa + cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // K-major: advance 16 bf16 = 32 B inside the 128 B swizzle atom (+2 in 16 B units);
                        // MN-major: 16 k-rows of 128 B further down, the next 64 M/N elements 8 KB away
                        const uint64_t ad = (MN & 2) ? umma_desc_sw128_mn(sa + k * 2048, 8192) : adesc + 2 * k;
                        const uint64_t bd = (MN & 1) ? umma_desc_sw128_mn(sa + cfg::A_BYTES + k * 2048, 8192) : bdesc + 2 * k;
                        if (CG == 2) tc_mma_pair(d_tmem, ad, bd, idesc, ((kb - kb0) | k) != 0);
                        else tc_mma(d_tmem, ad, bd, idesc, ((kb - kb0) | k) != 0);
                    }
                    if (CG == 2) tc_commit_pair(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);
                    if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                if (CG == 2) tc_commit_pair(&tfull_bar[acc]); else tc_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ---------------------------------------------------- epilogue ----
        // Warp pair (q, half): TMEM lanes [32q, 32q+32) = tile rows, column half `half`.
        // Results are staged in a per-warp 32 x 128 B swizzled buffer and leave through
        // TMA (plain store, or fp32 reduce-add into the residual stream), so HBM/L2 sees
        // full 128-byte lines and out-of-range rows/columns are clipped by the hardware.
        const int q = warp & 3;
        const int half = (warp - EPI_WARP0) >> 2;
        uint8_t *stg0 = smem + cfg::STG_OFF + (warp - EPI_WARP0) * STG_BYTES * cfg::STG_BUFS;
        int sbuf = 0;                            // staging buffer of the next store (double-buffered for CTA pairs)
        int acc = 0;
        uint32_t acc_phase = 0;
        constexpr bool geglu = EPI == E_GEGLU;
        constexpr bool out_f32 = EPI == E_F32;
        constexpr bool rope = EPI == E_ROPE;
        const bool accum = p.flags & LIN_ACCUM;
        const int n_out = geglu ? p.N / 2 : p.N;
        const int sw = lane & 7;                 // 128B-swizzle phase of this thread's staging row
        pdl_wait();                              // C may still be read / written by the previous kernel
        const uint32_t tempty_leader0 = CG == 2 ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0;
        for (int t = wid; t < num_tiles; t += nworkers) {
            int tmn = t % tiles_mn, ksp = t / tiles_mn;
            int tm, tn;
            tile_coords(p, tmn, tm, tn);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int row0 = (tm * CG + crank) * BM + q * 32;       // first row of this warp
            const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
            if (rope) {
                if constexpr (BN == 256) {
                    // ---- fused RoPE (half-split rotation, model/utils.py:4-16) + Q/K/V split ----
                    const int row = row0 + lane;
                    const int bb = row / p.s_x, ss = row % p.s_x;
                    if (tn <= p.n_q_tiles) {
                        const float *cs = p.rope_cos + (long)ss * 128;
                        const float *sn = p.rope_sin + (long)ss * 128;
                        for (int cc = 0; cc < 2; ++cc) {
                            const int c = half * 2 + cc;        // pair index: cols [32c,32c+32) with +128
                            uint32_t r1[32], r2[32];
                            tc_ld32(taddr + c * 32, r1);
                            tc_ld32(taddr + 128 + c * 32, r2);
                            tc_ld_wait();
                            uint32_t o1[16], o2[16];
                            if (row < p.M) {
#pragma unroll
                                for (int i = 0; i < 32; i += 4) {
                                    float4 cv = __ldg(reinterpret_cast<const float4 *>(cs + c * 32 + i));
                                    float4 sv = __ldg(reinterpret_cast<const float4 *>(sn + c * 32 + i));
                                    float cf[4] = {cv.x, cv.y, cv.z, cv.w}, sf[4] = {sv.x, sv.y, sv.z, sv.w};
                                    float a1[4], a2[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        float x1 = __uint_as_float(r1[i + j]), x2 = __uint_as_float(r2[i + j]);
                                        a1[j] = x1 * cf[j] - x2 * sf[j];
                                        a2[j] = x2 * cf[j] + x1 * sf[j];
                                    }
                                    o1[i / 2] = pack_bf16x2(a1[0], a1[1]); o1[i / 2 + 1] = pack_bf16x2(a1[2], a1[3]);
                                    o2[i / 2] = pack_bf16x2(a2[0], a2[1]); o2[i / 2 + 1] = pack_bf16x2(a2[2], a2[3]);
                                }
                                bf16 *dst;
                                if (tn < p.n_q_tiles) dst = (bf16 *)p.C + (long)row * p.ldc + tn * 256;
                                else dst = p.k_out + bb * p.kv_batch_stride + (long)ss * 256;
#pragma unroll
                                for (int i = 0; i < 16; i += 4) {
                                    *reinterpret_cast<uint4 *>(dst + c * 32 + 2 * i) = make_uint4(o1[i], o1[i + 1], o1[i + 2], o1[i + 3]);
                                    *reinterpret_cast<uint4 *>(dst + 128 + c * 32 + 2 * i) = make_uint4(o2[i], o2[i + 1], o2[i + 2], o2[i + 3]);
                                }
                            }
                        }
                    } else {
                        // V tile: plain copy into the cache
                        for (int cc = 0; cc < 4; ++cc) {
                            const int c = half * 4 + cc;
                            uint32_t r1[32];
                            tc_ld32(taddr + c * 32, r1);
                            tc_ld_wait();
                            if (row < p.M) {
                                bf16 *dst = p.v_out + bb * p.kv_batch_stride + (long)ss * 256 + c * 32;
#pragma unroll
                                for (int i = 0; i < 32; i += 8) {
                                    uint4 o;
                                    o.x = pack_bf16x2(__uint_as_float(r1[i]), __uint_as_float(r1[i + 1]));
                                    o.y = pack_bf16x2(__uint_as_float(r1[i + 2]), __uint_as_float(r1[i + 3]));
                                    o.z = pack_bf16x2(__uint_as_float(r1[i + 4]), __uint_as_float(r1[i + 5]));
                                    o.w = pack_bf16x2(__uint_as_float(r1[i + 6]), __uint_as_float(r1[i + 7]));
                                    *reinterpret_cast<uint4 *>(dst + i) = o;
                                }
                            }
                        }
                    }
                }
            } else {
                // output columns of this warp inside the tile
                constexpr int cols_out_tile = geglu ? BN / 2 : BN;
                constexpr int per_store = out_f32 ? 32 : 64;             // output columns per 128-byte staging row
                const int my_cols0 = half * (cols_out_tile / 2);
                for (int oc = 0; oc < cols_out_tile / 2; oc += per_store) {
                    const int col_tile = my_cols0 + oc;                   // first output column (tile-local)
                    const int col_glob = tn * cols_out_tile + col_tile;
                    if (col_glob >= n_out) break;
                    // make sure the TMA store that last used this staging buffer has finished reading it
                    if (lane == 0) { if (cfg::STG_BUFS == 2) bulk_wait_read1(); else bulk_wait_read0(); }
                    __syncwarp();
                    uint8_t *stg = stg0 + sbuf * STG_BYTES;
                    uint8_t *stg_row = stg + lane * 128;
                    constexpr int n_sub = out_f32 ? 1 : 2;                // 32-column TMEM chunks per staging row
                    for (int sb = 0; sb < n_sub; ++sb) {
                        const int ct = col_tile + sb * 32;                // tile-local output column of this chunk
                        uint32_t r[32];
                        float v[32];
                        if (geglu) {
                            uint32_t u[32];
                            tc_ld32(taddr + ct, r);
                            tc_ld32(taddr + BN / 2 + ct, u);
                            tc_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                v[i] = gelu_fast(__uint_as_float(r[i])) * __uint_as_float(u[i]) * p.alpha;
                        } else {
                            tc_ld32(taddr + ct, r);
                            tc_ld_wait();
                            const int cg = tn * BN + ct;
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                float b4[4] = {0.f, 0.f, 0.f, 0.f};
                                if (p.bias && ksp == 0) {
                                    if (cg + i + 3 < p.N) {
                                        float4 bv = __ldg(reinterpret_cast<const float4 *>(p.bias + cg + i));
                                        b4[0] = bv.x; b4[1] = bv.y; b4[2] = bv.z; b4[3] = bv.w;
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 4; ++j) if (cg + i + j < p.N) b4[j] = __ldg(p.bias + cg + i + j);
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float x = __uint_as_float(r[i + j]) + b4[j];
                                    if (EPI == E_GELU) x = gelu_fast(x);
                                    if (EPI == E_SILU) x = silu(x);
                                    if (EPI == E_F32) {   // rare combinations stay run-time here
                                        if (p.flags & LIN_GELU) x = gelu_tanh(x);
                                        if (p.flags & LIN_SILU) x = silu(x);
                                    }
                                    v[i + j] = (EPI == E_F32) ? x * p.alpha : x;
                                }
                            }
                        }
                        // stage: 16-byte chunk index XOR (row & 7) = TMA SWIZZLE_128B
                        if (out_f32) {
#pragma unroll
                            for (int ch = 0; ch < 8; ++ch)
                                st_shared_v4(stg_row + ((ch ^ sw) << 4), __float_as_uint(v[4 * ch]), __float_as_uint(v[4 * ch + 1]),
                                             __float_as_uint(v[4 * ch + 2]), __float_as_uint(v[4 * ch + 3]));
                        } else {
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch)
                                st_shared_v4(stg_row + (((sb * 4 + ch) ^ sw) << 4), pack_bf16x2(v[8 * ch], v[8 * ch + 1]),
                                             pack_bf16x2(v[8 * ch + 2], v[8 * ch + 3]), pack_bf16x2(v[8 * ch + 4], v[8 * ch + 5]),
                                             pack_bf16x2(v[8 * ch + 6], v[8 * ch + 7]));
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (accum) tma_reduce_add_2d(&map_c, stg, col_glob, row0);
                        else tma_store_2d(&map_c, stg, col_glob, row0);
                        bulk_commit();
                    }
                    if (cfg::STG_BUFS == 2) sbuf ^= 1;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_remote(tempty_leader0 + acc * 8);   // the leader's MMA thread owns the accumulators
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) bulk_wait_read0();   // staging smem must outlive the store's read; completion = kernel end
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();   // neither CTA may retire (or free TMEM) while the peer still uses it
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------ host side ----
// 2D tensor [rows, cols] with row stride ld (elements); box = (128 bytes of columns) x box_rows, 128B swizzle
bool make_map(CUtensorMap *map, const void *base, long rows, long cols, long ld, int box_rows, bool f32 = false) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}


template <int BN, int CG, int EPI, int MN = 0>
int launch_epi(const LinearArgs &a, cudaStream_t st, const char **err, const TcParams *extra = nullptr) {
    using cfg = Cfg<BN, CG>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        if (cudaFuncSetAttribute(gemm_tc_kernel<BN, CG, EPI, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 cfg::SMEM_BYTES) != cudaSuccess) {
            if (err) *err = "cudaFuncSetAttribute(max dynamic smem) failed";
            return PZ_ERR_CUDA;
        }
    }
    const int g_num_sms = device_sm_count();
    CUtensorMap ma, mw, mc;
    const bool f32out = a.flags & LIN_OUT_F32;
    const int n_out = (a.flags & LIN_GEGLU) ? a.N / 2 : a.N;
    const bool a_mn = a.flags & LIN_A_MN, w_mn = a.flags & LIN_W_MN;
    // MN-major operands: the tensor is [K rows][M or N columns]; boxes of 64 columns x 64 k-rows
    const bool ok_a = a_mn ? make_map(&ma, a.A, a.K, a.M, a.lda, BK) : make_map(&ma, a.A, a.M, a.K, a.lda, BM);
    const bool ok_w = w_mn ? make_map(&mw, a.W, a.K, a.N, a.ldw, BK) : make_map(&mw, a.W, a.N, a.K, a.K, BN / CG);
    if (!ok_a || !ok_w || !make_map(&mc, a.C, a.M, n_out, a.ldc, 32, f32out)) {
        if (err) *err = "cuTensorMapEncodeTiled failed";
        return PZ_ERR_CUDA;
    }
    TcParams p;
    if (extra) p = *extra; else memset(&p, 0, sizeof(p));
    p.M = a.M; p.N = a.N; p.K = a.K; p.ldc = a.ldc; p.bias = a.bias; p.C = a.C;
    p.alpha = a.alpha; p.flags = a.flags | (extra ? LIN_ROPE : 0);
    p.a_mn = a_mn; p.w_mn = w_mn;
    p.tiles_m = (a.M + BM * CG - 1) / (BM * CG);   // tiles of the worker (CTA or CTA pair)
    p.tiles_n = (a.N + BN - 1) / BN;
    // split K when the output grid cannot fill the machine and the epilogue is a pure fp32
    // accumulate (TMA reduce-add makes the combine free): o_proj / down_proj at small M
    int num_kb = (a.K + BK - 1) / BK;
    int workers_max = g_num_sms / CG;
    p.ksplit = 1;
    if ((a.flags & LIN_ACCUM) && !(a.flags & (LIN_GELU | LIN_SILU | LIN_GEGLU)) && !extra) {
        // as many K slices as it takes to give every SM (pair) a work item, at least 4 k-blocks each: at M = 276
        // (bs=1 prefix) the 128 x 128 tiles of down_proj are MMA-rate bound per CTA (256 clk per k-block), so the
        // K loop must be spread over all SMs, not over a power-of-two subset
        int mn = p.tiles_m * p.tiles_n;
        int by_sms = workers_max / mn, by_k = num_kb / 4;
        p.ksplit = by_sms < by_k ? by_sms : by_k;
        if (p.ksplit < 1) p.ksplit = 1;
    }
    // A bigger than what L2 keeps and only a few N tiles: let the CTAs that run together share the A
    // tile (read from HBM once) instead of the W tile (which then stays L2 resident as a whole)
    {
        double a_bytes = (double)a.M * a.K * 2, w_bytes = (double)a.N * a.K * 2;
        p.n_fastest = (!extra && a_bytes > 48e6 && w_bytes < 96e6 && p.tiles_n <= 32) ? 1 : 0;
        // M-fastest raster: A is swept once per N tile.  A bigger than what one L2 partition keeps (~60 MB: measured
        // 24 % of the gate|up A reads came from DRAM at 72 MB) is cut into groups of <= 24 MB of A rows; W is then read
        // once per group instead of once
        static const int group_mb = [] { const char *e = getenv("PZ_GEMM_GROUP_MB"); return e ? atoi(e) : 24; }();
        p.group_m = 0;
        if (!p.n_fastest && group_mb > 0 && a_bytes > 40e6 && p.tiles_n > 1) {
            double tile_bytes = (double)BM * CG * a.K * 2;
            int g = (int)(group_mb * 1e6 / tile_bytes);
            if (g < 1) g = 1;
            int ngroups = (p.tiles_m + g - 1) / g;
            p.group_m = (p.tiles_m + ngroups - 1) / ngroups;   // equal-sized groups
            if (p.group_m >= p.tiles_m) p.group_m = 0;
        }
        // L2 eviction priorities (experiment, off unless PZ_GEMM_HINTS=1): keep the operand every concurrently running
        // worker re-reads, drop the streaming one first.  Measured: no effect on gate|up (610 vs 611 MB of DRAM reads) and
        // HARMFUL for down_proj (1333 -> 2759 MB: the 8 workers of an M tile no longer find each other's A lines in L2).
        static const bool hints = [] { const char *e = getenv("PZ_GEMM_HINTS"); return e && e[0] == '1'; }();
        p.a_hint = p.w_hint = L2_EVICT_NORMAL;
        if (hints && p.n_fastest && a_bytes > 96e6) { p.a_hint = L2_EVICT_FIRST; p.w_hint = L2_EVICT_LAST; }
        else if (hints && p.group_m > 0) { p.a_hint = L2_EVICT_LAST; }
    }
    p.kb_per_split = (num_kb + p.ksplit - 1) / p.ksplit;
    p.ksplit = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
    int tiles = p.tiles_m * p.tiles_n * p.ksplit;
    int workers = tiles < workers_max ? tiles : workers_max;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(workers * CG); lc.blockDim = dim3(NUM_THREADS2); lc.dynamicSmemBytes = cfg::SMEM_BYTES; lc.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (g_pdl_enabled && !g_pdl_off) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (CG == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    lc.attrs = attr; lc.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&lc, gemm_tc_kernel<BN, CG, EPI, MN>, ma, mw, mc, p);
    count_launch();
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        return PZ_ERR_CUDA;
    }
    return 0;
}

// pick the epilogue instantiation from the run-time flags (bf16 outputs with an activation and a non-unit
// alpha do not occur on this path; they take the fp32-capable flavour's run-time tests)
template <int BN, int CG>
int launch(const LinearArgs &a, cudaStream_t st, const char **err, const TcParams *extra = nullptr) {
    if (extra) {
        if constexpr (BN == 256) return launch_epi<BN, CG, E_ROPE>(a, st, err, extra);
        else return PZ_ERR_INVALID;
    }
    if (a.flags & LIN_GEGLU) {
        if constexpr (BN == 256) return launch_epi<BN, CG, E_GEGLU>(a, st, err);
        else return PZ_ERR_INVALID;
    }
    if (a.flags & (LIN_A_MN | LIN_W_MN)) {
        // the backward's two products: dX (W in place: MN = 1, bf16 or fp32 output) and dW (both in place: MN = 3, fp32 reduce-add)
        const bool both = (a.flags & LIN_A_MN) && (a.flags & LIN_W_MN);
        if (!(a.flags & LIN_W_MN) || (a.flags & (LIN_GELU | LIN_SILU)) || (!(a.flags & LIN_OUT_F32) && a.alpha != 1.f)) {
            if (err) *err = "tcgen05 gemm: this combination of MN-major operands / epilogue is not instantiated";
            return PZ_ERR_INVALID;
        }
        if (both) {
            if (a.flags & LIN_OUT_F32) return launch_epi<BN, CG, E_F32, 3>(a, st, err);
            if (err) *err = "tcgen05 gemm: MN-major A needs an fp32 output";
            return PZ_ERR_INVALID;
        }
        if (a.flags & LIN_OUT_F32) return launch_epi<BN, CG, E_F32, 1>(a, st, err);
        return launch_epi<BN, CG, E_PLAIN, 1>(a, st, err);
    }
    if (a.flags & LIN_OUT_F32) return launch_epi<BN, CG, E_F32>(a, st, err);
    if (a.alpha == 1.f) {
        if (a.flags & LIN_GELU) return launch_epi<BN, CG, E_GELU>(a, st, err);
        if (a.flags & LIN_SILU) return launch_epi<BN, CG, E_SILU>(a, st, err);
        if (!(a.flags & (LIN_GELU | LIN_SILU))) return launch_epi<BN, CG, E_PLAIN>(a, st, err);
    }
    if (err) *err = "tcgen05 gemm: bf16 output with activation and alpha != 1 is not instantiated";
    return PZ_ERR_INVALID;
}

// CTA pairs pay off when there are enough 256-row tiles to keep every pair busy
bool use_pair(const LinearArgs &a) {
    static const int mode = [] { const char *e = getenv("PZ_GEMM_CG"); return e ? atoi(e) : 0; }();
    if (mode == 1) return false;
    return a.M >= 2048 && a.N >= 256;
}

}  // namespace

int gemm_tc_supported(const LinearArgs &a) {
    if (a.flags & (LIN_NORM_A | LIN_COMBINE_A)) return 0;
    if ((a.flags & LIN_W_MN) && (a.ldw % 8 || a.ldw < a.N)) return 0;
    if ((a.flags & LIN_A_MN) && (a.lda < a.M || !(a.flags & LIN_W_MN) || !(a.flags & LIN_OUT_F32))) return 0;
    if ((a.flags & (LIN_A_MN | LIN_W_MN)) && (a.flags & (LIN_GELU | LIN_SILU | LIN_GEGLU))) return 0;
    if (a.M < 1 || (!(a.flags & LIN_W_MN) && a.K % 8) || a.lda % 8) return 0;          // TMA: 16-byte global strides
    if (((uintptr_t)a.A | (uintptr_t)a.W | (uintptr_t)a.C) & 15) return 0;
    int n_out = (a.flags & LIN_GEGLU) ? a.N / 2 : a.N;
    if (n_out % 8 || a.ldc % 8) return 0;                    // TMA store: 16-byte global strides
    if ((a.flags & LIN_GEGLU) && (a.N % 256)) return 0;      // gate|up blocks of 128 pair up inside one 256-wide tile
    if ((a.flags & LIN_ACCUM) && !(a.flags & LIN_OUT_F32)) return 0;
    if (!(a.flags & (LIN_OUT_F32 | LIN_GEGLU)) && a.alpha != 1.f) return 0;   // bf16 epilogues are instantiated without alpha
    return 1;
}

int launch_linear_tc(const LinearArgs &a, cudaStream_t st, const char **err) {
    // 128 x 256 tiles reach the tensor-core issue rate with one CTA per SM; fall back to
    // 128 x 128 when that would leave SMs idle or N is small.
    bool geglu = a.flags & LIN_GEGLU;
    long tiles256 = (long)((a.M + BM - 1) / BM) * ((a.N + 255) / 256);
    bool use256 = geglu || (a.N >= 256 && tiles256 >= 120);
    if (!use256 && a.N >= 256 && a.K >= 8192 && (a.flags & LIN_ACCUM) && !(a.flags & (LIN_GELU | LIN_SILU | LIN_GEGLU))) {
        // split-K capable (fp32 reduce-add), long K and too few tiles: a weight-streaming problem (down_proj at bs=1).  Wider tiles move fewer
        // activation bytes per weight byte through each SM (the per-SM bytes in flight are the limit), as long as
        // the K split still yields a work item for (almost) every SM.
        const int g_num_sms = device_sm_count();
        long tiles128 = (long)((a.M + BM - 1) / BM) * ((a.N + 127) / 128);
        long by_k = ((a.K + BK - 1) / BK) / 4;
        auto items = [&](long tiles) { long s = g_num_sms / tiles; if (s > by_k) s = by_k; if (s < 1) s = 1; return tiles * s; };
        use256 = items(tiles256) * 10 >= items(tiles128) * 9;
    }
    if (use256 && use_pair(a)) return launch<256, 2>(a, st, err);   // (256 x 128 pair tiles for N = 1152 were tried: exact cover, 7.8 instead of 5 waves, but 18.5 vs 17.7 ms for the SigLIP stage -- 50 % more operand bytes per MMA cycle)
    return use256 ? launch<256, 1>(a, st, err) : launch<128, 1>(a, st, err);
}

// Fused QKV projection + RoPE + cache write for the prefix pass (head_dim 256):
// q (post-RoPE) -> q_out [M, n_heads*256]; K (post-RoPE), V -> cache rows
// (b, s) = (m / s_x, m % s_x) at k_out/v_out + b*kv_batch_stride + s*256.
// Replaces mixture.py:187-235 + the kv_cache.update of joint_model.py:195-219.
int launch_qkv_rope_tc(const void *A, int lda, const void *W, void *q_out, void *k_out, void *v_out,
                       long kv_batch_stride, const float *cos_t, const float *sin_t, int M, int K,
                       int n_heads, int s_x, cudaStream_t st, const char **err) {
    LinearArgs a;
    a.A = A; a.W = W; a.bias = nullptr; a.C = q_out;
    a.M = M; a.N = (n_heads + 2) * 256; a.K = K; a.lda = lda; a.ldc = n_heads * 256;
    a.alpha = 1.f; a.flags = 0; a.norm_w = nullptr;
    TcParams ex;
    memset(&ex, 0, sizeof(ex));
    ex.rope_cos = cos_t; ex.rope_sin = sin_t; ex.k_out = (bf16 *)k_out; ex.v_out = (bf16 *)v_out;
    ex.kv_batch_stride = kv_batch_stride; ex.s_x = s_x; ex.n_q_tiles = n_heads;
    if (use_pair(a)) return launch<256, 2>(a, st, err, &ex);
    return launch<256, 1>(a, st, err, &ex);
}
