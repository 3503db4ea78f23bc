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

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;

// ---------------------------------------------------------------- PTX ------
PZ_DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

PZ_DEVINL void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
PZ_DEVINL void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
PZ_DEVINL void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
PZ_DEVINL void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
PZ_DEVINL void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
PZ_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PZ_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
PZ_DEVINL void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
PZ_DEVINL void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
PZ_DEVINL void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
PZ_DEVINL void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle, bf16:
// 8-row core-matrix groups are 1024 B apart (SBO); LBO is unused for swizzled
// K-major tiles (set to 1 like CUTLASS).  Descriptor version 1 = Blackwell.
PZ_DEVINL uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                  // LBO (16 B units)
    d |= (uint64_t)(1024 >> 4) << 32;        // SBO
    d |= (uint64_t)1 << 46;                  // version
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}
// Instruction descriptor: D fp32, A/B bf16, both K-major, M x N.
constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcParams {
    int M, N, K, ldc;
    const float *bias;
    void *C;
    float alpha;
    int flags;
    int tiles_m, tiles_n;
};

template <int BN> struct Cfg {
    static constexpr int STAGES = BN == 256 ? 4 : 6;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;   // two accumulator stages (power of two: 256 / 512)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const TcParams p) {
    using cfg = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = (uint64_t *)(smem + cfg::STAGES * cfg::STAGE_BYTES);
    uint64_t *empty_bar = full_bar + cfg::STAGES;
    uint64_t *tfull_bar = empty_bar + cfg::STAGES;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + BK - 1) / BK;
    const int num_tiles = p.tiles_m * p.tiles_n;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < cfg::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"((uint32_t)cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ----
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int tm = t % p.tiles_m, tn = t / p.tiles_m;   // m fastest: concurrent CTAs share the W tile in L2
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * cfg::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], cfg::STAGE_BYTES);
                    tma_load_2d(&map_a, &full_bar[stage], sa, kb * BK, tm * BM);
                    tma_load_2d(&map_w, &full_bar[stage], sa + cfg::A_BYTES, kb * BK, tn * BN);
                    if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------- MMA issuer ----
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    uint32_t sa = smem_u32(smem + stage * cfg::STAGE_BYTES);
                    uint64_t adesc = umma_desc_sw128(sa);
                    uint64_t bdesc = umma_desc_sw128(sa + cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 B inside the 128 B swizzle atom: +2 in 16 B units
                        tc_mma(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    tc_commit(&empty_bar[stage]);
                    if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ---------------------------------------------------- epilogue ----
        const int q = warp & 3;   // TMEM lane quadrant this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool geglu = p.flags & LIN_GEGLU;
        const bool out_f32 = p.flags & LIN_OUT_F32;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int tm = t % p.tiles_m, tn = t / p.tiles_m;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int row = tm * BM + q * 32 + lane;
            const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
            const int n_chunks = geglu ? BN / 64 : BN / 32;
            for (int c = 0; c < n_chunks; ++c) {
                uint32_t r[32];
                float v[32];
                tc_ld32(taddr + c * 32, r);
                int col0;   // first output column of this chunk
                if (geglu) {
                    uint32_t u[32];
                    tc_ld32(taddr + BN / 2 + c * 32, u);
                    tc_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        v[i] = gelu_tanh(__uint_as_float(r[i])) * __uint_as_float(u[i]) * p.alpha;
                    col0 = tn * (BN / 2) + c * 32;
                } else {
                    tc_ld_wait();
                    col0 = tn * BN + c * 32;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float x = __uint_as_float(r[i]);
                        if (p.bias && col0 + i < p.N) x += __ldg(p.bias + col0 + i);
                        if (p.flags & LIN_GELU) x = gelu_tanh(x);
                        if (p.flags & LIN_SILU) x = silu(x);
                        v[i] = x * p.alpha;
                    }
                }
                const int n_out = geglu ? p.N / 2 : p.N;
                if (row < p.M && col0 < n_out) {
                    int nvalid = n_out - col0 < 32 ? n_out - col0 : 32;   // multiple of 8 (checked on host)
                    if (out_f32) {
                        float *dst = (float *)p.C + (long)row * p.ldc + col0;
                        for (int i = 0; i < nvalid; i += 4) {
                            float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                            if (p.flags & LIN_ACCUM) {
                                float4 old = *reinterpret_cast<float4 *>(dst + i);
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            *reinterpret_cast<float4 *>(dst + i) = o;
                        }
                    } else {
                        bf16 *dst = (bf16 *)p.C + (long)row * p.ldc + col0;
                        for (int i = 0; i < nvalid; i += 8) {
                            uint4 o;
                            o.x = pack_bf16x2(v[i], v[i + 1]);
                            o.y = pack_bf16x2(v[i + 2], v[i + 3]);
                            o.z = pack_bf16x2(v[i + 4], v[i + 5]);
                            o.w = pack_bf16x2(v[i + 6], v[i + 7]);
                            *reinterpret_cast<uint4 *>(dst + i) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)cfg::TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------ host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2D bf16 tensor [rows, cols] with row stride ld (elements); box = 64 cols x box_rows, 128B swizzle
bool make_map(CUtensorMap *map, const void *base, long rows, long cols, long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

int g_num_sms = 0;

template <int BN>
int launch(const LinearArgs &a, cudaStream_t st, const char **err) {
    using cfg = Cfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 cfg::SMEM_BYTES) != cudaSuccess) {
            if (err) *err = "cudaFuncSetAttribute(max dynamic smem) failed";
            return PZ_ERR_CUDA;
        }
        attr_set = true;
    }
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    CUtensorMap ma, mw;
    if (!make_map(&ma, a.A, a.M, a.K, a.lda, BM) || !make_map(&mw, a.W, a.N, a.K, a.K, BN)) {
        if (err) *err = "cuTensorMapEncodeTiled failed";
        return PZ_ERR_CUDA;
    }
    TcParams p;
    p.M = a.M; p.N = a.N; p.K = a.K; p.ldc = a.ldc; p.bias = a.bias; p.C = a.C;
    p.alpha = a.alpha; p.flags = a.flags;
    p.tiles_m = (a.M + BM - 1) / BM;
    p.tiles_n = (a.N + BN - 1) / BN;
    int tiles = p.tiles_m * p.tiles_n;
    int grid = tiles < g_num_sms ? tiles : g_num_sms;
    gemm_tc_kernel<BN><<<grid, NUM_THREADS, cfg::SMEM_BYTES, st>>>(ma, mw, p);
    count_launch();
    return 0;
}

}  // namespace

int gemm_tc_supported(const LinearArgs &a) {
    if (a.M < 1 || a.K % 8 || a.lda % 8) return 0;          // TMA: 16-byte global strides
    if (((uintptr_t)a.A | (uintptr_t)a.W | (uintptr_t)a.C) & 15) return 0;
    int n_out = (a.flags & LIN_GEGLU) ? a.N / 2 : a.N;
    if (n_out % 8 || a.ldc % 8) return 0;                    // vector stores
    if ((a.flags & LIN_GEGLU) && (a.N % 256)) return 0;      // gate|up blocks of 128 pair up inside one 256-wide tile
    if ((a.flags & LIN_ACCUM) && !(a.flags & LIN_OUT_F32)) return 0;
    return 1;
}

int launch_linear_tc(const LinearArgs &a, cudaStream_t st, const char **err) {
    // 128 x 256 tiles reach the tensor-core issue rate with one CTA per SM; fall back to
    // 128 x 128 when that would leave SMs idle or N is small.
    bool geglu = a.flags & LIN_GEGLU;
    long tiles256 = (long)((a.M + BM - 1) / BM) * ((a.N + 255) / 256);
    bool use256 = geglu || (a.N >= 256 && tiles256 >= 120);
    return use256 ? launch<256>(a, st, err) : launch<128>(a, st, err);
}
