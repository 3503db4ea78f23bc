// simple.cu -- plain SIMT kernels for the linear layers and the block-masked
// attention.  They are (a) the fp32 path of the library (tensor cores have no
// exact fp32 mode; the fp32 parity bar of 1e-4 needs real fp32 FMAs), and (b)
// the on-device cross-check for the tcgen05 / mma kernels in the unit tests.
// Not the bf16 production path.
#include "common.cuh"
#include "kernels.h"

// ---------------------------------------------------------------- linear --
// C[M,N] = alpha * (A[M,K] . W[N,K]^T + bias) with the LIN_* epilogues.
// 64x64 output tile, 16-wide K slab, 256 threads, 4x4 outputs per thread.
template <typename T, bool GEGLU>
__global__ void __launch_bounds__(256) linear_simple_kernel(LinearArgs a) {
    pdl_trigger();
    pdl_wait();
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float sA[BK][BM + 1];
    __shared__ float sW[GEGLU ? 2 : 1][BK][BN + 1];
    const T *A = (const T *)a.A;
    const T *W = (const T *)a.W;
    int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;  // n0 indexes OUTPUT columns
    float acc[4][4] = {}, acc2[4][4] = {};
    for (int k0 = 0; k0 < a.K; k0 += BK) {
        for (int i = threadIdx.x; i < BM * BK; i += 256) {
            int r = i / BK, c = i % BK;
            int m = m0 + r, k = k0 + c;
            sA[c][r] = (m < a.M && k < a.K) ? to_f32<T>((a.flags & LIN_A_MN) ? A[(long)k * a.lda + m] : A[(long)m * a.lda + k]) : 0.f;
        }
        for (int i = threadIdx.x; i < BN * BK; i += 256) {
            int r = i / BK, c = i % BK;
            int n = n0 + r, k = k0 + c;
            if (!GEGLU) {
                sW[0][c][r] = (n < a.N && k < a.K) ? to_f32<T>((a.flags & LIN_W_MN) ? W[(long)k * a.ldw + n] : W[(long)n * a.K + k]) : 0.f;
            } else {
                // output column n <- gate row g, up row g + PZ_GU_BLOCK
                long g = (long)(n / PZ_GU_BLOCK) * (2 * PZ_GU_BLOCK) + (n % PZ_GU_BLOCK);
                bool ok = (n < a.N / 2) && k < a.K;
                sW[0][c][r] = ok ? to_f32<T>(W[g * a.K + k]) : 0.f;
                sW[1][c][r] = ok ? to_f32<T>(W[(g + PZ_GU_BLOCK) * a.K + k]) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[4], wv[4], wv2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = sA[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                wv[j] = sW[0][k][tx * 4 + j];
                if (GEGLU) wv2[j] = sW[1][k][tx * 4 + j];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
                    if (GEGLU) acc2[i][j] = fmaf(av[i], wv2[j], acc2[i][j]);
                }
        }
        __syncthreads();
    }
    int n_out = GEGLU ? a.N / 2 : a.N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= n_out) continue;
            float v = acc[i][j];
            if (GEGLU) {
                v = gelu_tanh(v) * acc2[i][j];
            } else {
                if (a.bias) v += a.bias[n];
                if (a.flags & LIN_GELU) v = gelu_tanh(v);
                if (a.flags & LIN_SILU) v = silu(v);
            }
            v *= a.alpha;
            long o = (long)m * a.ldc + n;
            if (a.flags & LIN_OUT_F32) {
                float *C = (float *)a.C;
                C[o] = (a.flags & LIN_ACCUM) ? C[o] + v : v;
            } else {
                ((T *)a.C)[o] = from_f32<T>(v);
            }
        }
    }
}

template <typename T>
void launch_linear_simple(const LinearArgs &a, cudaStream_t st) {
    bool geglu = a.flags & LIN_GEGLU;
    int n_out = geglu ? a.N / 2 : a.N;
    dim3 grid((n_out + 63) / 64, (a.M + 63) / 64);
    if (geglu) launch_k(linear_simple_kernel<T, true>, dim3(grid), dim3(256), 0, st, a);
    else launch_k(linear_simple_kernel<T, false>, dim3(grid), dim3(256), 0, st, a);
}
template void launch_linear_simple<float>(const LinearArgs &, cudaStream_t);
template void launch_linear_simple<bf16>(const LinearArgs &, cudaStream_t);

// ------------------------------------------------------------- attention --
// One warp per (sample, head, query row).  Lane j scores keys j, j+32, ...;
// softmax over the visible keys in fp32; then lane d accumulates output dims
// d, d+32, ...  Mirrors joint_model.py:261-282 (scale, tanh soft-cap, mask,
// fp32 softmax, PV) and siglip.py:133-152 (softcap = 0, no mask).
template <typename T>
__global__ void __launch_bounds__(128) attn_simple_kernel(AttnArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float smem[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int nk = a.s_cache + a.n_fresh;
    float *sq = smem + warp * (a.head_dim + nk);
    float *sp = sq + a.head_dim;
    long row_id = (long)blockIdx.x * 4 + warp;
    long total = (long)a.batch * a.n_heads * a.q_rows;
    if (row_id >= total) return;
    int r = row_id % a.q_rows;
    int h = (row_id / a.q_rows) % a.n_heads;
    int b = row_id / ((long)a.q_rows * a.n_heads);
    int tok = a.q_row0 + r;
    int vlen = a.valid_len ? a.valid_len[b] : a.s_cache;
    T *o = (T *)a.O + b * a.o_batch_stride + (long)r * a.o_row_stride + h * a.o_head_stride;
    if (a.valid_len && tok < a.s_vlm && tok >= vlen) {  // pad row: never read by a valid row
        for (int d = lane; d < a.head_dim; d += 32) o[d] = from_f32<T>(0.f);
        return;
    }
    const T *q = (const T *)a.Q + b * a.q_batch_stride + (long)r * a.q_row_stride + h * a.q_head_stride;
    for (int d = lane; d < a.head_dim; d += 32) sq[d] = to_f32<T>(q[d]);
    __syncwarp();
    const T *K = (const T *)a.K + b * a.kv_batch_stride + h * a.kv_head_stride;
    const T *V = (const T *)a.V + b * a.kv_batch_stride + h * a.kv_head_stride;
    const T *K2 = a.K2 ? (const T *)a.K2 + b * a.kv2_batch_stride + h * a.kv_head_stride : nullptr;
    const T *V2 = a.V2 ? (const T *)a.V2 + b * a.kv2_batch_stride + h * a.kv_head_stride : nullptr;
    float mx = -INFINITY;
    for (int j = lane; j < nk; j += 32) {
        bool vis;
        const T *kr;
        if (j < a.s_cache) {
            vis = (j < a.s_vlm) ? (j < vlen) : (!a.valid_len || tok >= a.s_vlm);
            kr = K + (long)j * a.kv_row_stride;
        } else {
            vis = tok >= a.s_cache;
            kr = K2 + (long)(j - a.s_cache) * a.kv2_row_stride;
        }
        float s = -INFINITY;
        if (vis) {
            float acc = 0.f;
            for (int d = 0; d < a.head_dim; ++d) acc = fmaf(sq[d], to_f32<T>(kr[d]), acc);
            s = acc * a.scale;
            if (a.softcap > 0.f) s = tanhf(s / a.softcap) * a.softcap;
        }
        sp[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < nk; j += 32) {
        float p = (sp[j] == -INFINITY) ? 0.f : __expf(sp[j] - mx);
        sp[j] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    __syncwarp();
    for (int d = lane; d < a.head_dim; d += 32) {
        float acc = 0.f;
        for (int j = 0; j < nk; ++j) {
            float p = sp[j];
            if (p == 0.f) continue;
            const T *vr = (j < a.s_cache) ? V + (long)j * a.kv_row_stride
                                          : V2 + (long)(j - a.s_cache) * a.kv2_row_stride;
            acc = fmaf(p, to_f32<T>(vr[d]), acc);
        }
        o[d] = from_f32<T>(acc * inv);
    }
}

template <typename T>
void launch_attn_simple(const AttnArgs &a, cudaStream_t st) {
    long total = (long)a.batch * a.n_heads * a.q_rows;
    int nk = a.s_cache + a.n_fresh;
    size_t smem = 4 * (size_t)(a.head_dim + nk) * sizeof(float);
    launch_k(attn_simple_kernel<T>, dim3((unsigned)((total + 3) / 4)), dim3(128), smem, st, a);
}
template void launch_attn_simple<float>(const AttnArgs &, cudaStream_t);
template void launch_attn_simple<bf16>(const AttnArgs &, cudaStream_t);
