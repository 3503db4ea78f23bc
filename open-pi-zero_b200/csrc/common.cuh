// common.cuh -- shared device helpers and host-side launch bookkeeping.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pz_b200.h"

typedef __nv_bfloat16 bf16;

#define PZ_DEVINL __device__ __forceinline__

// ---- scalar conversion -------------------------------------------------
template <typename T> PZ_DEVINL float to_f32(T v);
template <> PZ_DEVINL float to_f32<float>(float v) { return v; }
template <> PZ_DEVINL float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> PZ_DEVINL T from_f32(float v);
template <> PZ_DEVINL float from_f32<float>(float v) { return v; }
template <> PZ_DEVINL bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// Round the dynamic shared-memory base up to `align` bytes without a round trip through an integer: a pointer rebuilt
// from uintptr_t loses its address space, and every access through it compiles to a generic LD.E / ST.E (long-scoreboard,
// slower) instead of LDS / STS.
PZ_DEVINL uint8_t *align_smem(uint8_t *raw, uint32_t align) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(raw);
    return raw + ((align - (a & (align - 1))) & (align - 1));
}
PZ_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
PZ_DEVINL float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
PZ_DEVINL float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// gelu(approximate="tanh"): paligemma/modules.py:94, siglip.py:190
PZ_DEVINL float gelu_tanh(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float u = k0 * (x + k1 * x * x * x);
    return 0.5f * x * (1.0f + tanhf(u));
}
PZ_DEVINL float silu(float x) { return x / (1.0f + __expf(-x)); }

PZ_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
PZ_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum for blockDim.x <= 1024 (scratch: 32 floats of shared memory)
PZ_DEVINL float block_sum(float v, float *scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    float r = (lane < nw) ? scratch[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

// ---- linear-layer argument block (all GEMM flavours share it) ------------
enum : int {
    LIN_GELU = 1,      // gelu_tanh(acc + bias)
    LIN_OUT_F32 = 2,   // C is fp32 (else T)
    LIN_ACCUM = 4,     // C(fp32) += alpha * (acc + bias)
    LIN_GEGLU = 8,     // W rows are [128 gate | 128 up] blocks; C[M, N/2] = gelu(gate)*up
    LIN_SILU = 16,     // silu(acc + bias)
    LIN_NORM_A = 32,   // A is the fp32 residual stream; apply Gemma RMSNorm (norm_w) while loading it
    LIN_COMBINE_A = 64, // A is the split-key attention partials; combine them while loading (skinny only)
    LIN_A_MN = 128,    // A is stored [K][M] (row stride lda): the M dimension is contiguous ("MN-major" UMMA operand);
    LIN_W_MN = 256,    // W is stored [K][N] (row stride ldw).  tcgen05 GEMM only: the backward of the training step reads
                       // dY, X and W in place instead of through transposed copies
};

struct LinearArgs {
    const void *A;      // [M, K] T, row stride lda
    const void *W;      // [N, K] T, row stride K
    const float *bias;  // [N] or nullptr
    void *C;            // [M, N] (T or fp32), row stride ldc
    int M, N, K, lda, ldc;
    float alpha;
    int flags;
    int ldw;            // LIN_W_MN: row stride of the [K][N] storage (otherwise unused)
    const float *norm_w;   // LIN_NORM_A: RMSNorm scale (raw w; 1+w is applied), eps 1e-6
    // LIN_COMBINE_A: A = partials [batch][n_splits][heads*q_rows][hd+2] fp32 (o, m, l) of attn_mma's
    // split-key mode; logical A[m = b*q_rows + tok][k = h*hd + d]
    int cmb_splits, cmb_q_rows, cmb_heads, cmb_hd;
};

// ---- attention argument block ------------------------------------------
// Query rows are (sample b, query token r, head h).  Keys come in two
// segments: a cache segment of s_cache rows (first s_vlm of them are image/text
// positions, valid iff index < valid_len[b]; rows >= s_vlm, the proprio rows,
// are always visible to rows that can see them) and an optional fresh segment
// of n_fresh rows (the action tokens of the current Euler step).
// Row r (global token index q_row0 + r in [vlm | proprio | action] order) sees:
//   r <  s_vlm            : cache keys j < valid_len  (and is itself skipped if r >= valid_len)
//   s_vlm <= r < s_cache  : cache keys j < valid_len or s_vlm <= j < s_cache
//   r >= s_cache          : the above plus every fresh key
// (block mask of pizero.py:271-310).  valid_len == nullptr => everything visible
// (SigLIP, siglip.py:133-152).
struct AttnArgs {
    const void *Q; long q_batch_stride; int q_row_stride, q_head_stride;
    const void *K, *V; long kv_batch_stride; int kv_row_stride, kv_head_stride;
    const void *K2, *V2; long kv2_batch_stride; int kv2_row_stride;
    const int32_t *valid_len;
    void *O; long o_batch_stride; int o_row_stride, o_head_stride;
    int batch, n_heads, head_dim, q_rows, q_row0, s_cache, s_vlm, n_fresh;
    float scale, softcap;
    // optional fp32 scratch for the split-key (flash-decoding) path:
    // batch * key_tiles * n_heads*q_rows * (head_dim + 2) floats
    float *scratch; size_t scratch_bytes;
    // optional fused RoPE (tensor-core kernel, head_dim 256): Q rows and the fresh K2 rows are
    // raw projections; rotate them while staging.  Table row = rope_pos0 + token index.
    const float *rope_cos, *rope_sin; int rope_pos0;
};

// ---- programmatic dependent launch (PDL) ----------------------------------------
// Every kernel of this library is launched with programmatic stream serialisation:
// it may start while its predecessor is still running, does whatever does not depend
// on the predecessor (barrier / TMEM setup, descriptor prefetch, *weight* prefetch),
// and only then waits.  Rule: pdl_wait() before the first access to any buffer a
// previous kernel may read or write.  Without the launch attribute both are no-ops.
PZ_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PZ_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// host-side, per device: one process may drive several GPUs (the reference's callers do `.to(f"cuda:{gpu_id}")`), so
// "done once" function attributes and the SM count are keyed by the current device ordinal
struct PerDeviceOnce {
    unsigned long long done = 0;
    bool need() {   // true the first time it is asked on the current device
        int d = 0;
        cudaGetDevice(&d);
        if (d < 0 || d > 63) return true;
        if ((done >> d) & 1ull) return false;
        done |= 1ull << d;
        return true;
    }
};
static inline int device_sm_count() {
    static int sms[64] = {0};
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d > 63) d = 0;
    if (!sms[d]) cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d);
    return sms[d];
}

// host-side: every launch goes through this counter (bench.py: gpu_launches)
struct LaunchCounter { long long n = 0; };
extern thread_local LaunchCounter *g_launch_counter;
extern int g_pdl_enabled;
// > 0: launch without the PDL attribute (griddepcontrol.wait is then a no-op and plain stream order holds).  The GEMM and
// skinny kernels fetch their WEIGHT operand before griddepcontrol.wait -- correct for weights, wrong when that operand was
// written by the kernel just in front (the transposed operands of the training step's backward, train.cu).
extern thread_local int g_pdl_off;
struct PdlOff {
    PdlOff() { ++g_pdl_off; }
    ~PdlOff() { --g_pdl_off; }
};
static inline void count_launch() { if (g_launch_counter) g_launch_counter->n++; }

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (g_pdl_enabled && !g_pdl_off) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    count_launch();
}
