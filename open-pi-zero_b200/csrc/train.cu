// train.cu -- the flow-matching training step (reference: PiZero.forward, src/model/vla/pizero.py:607-661, followed by
// loss.backward(), src/agent/train.py:350-368): forward with every activation the backward needs kept in HBM, then the
// backward written out op by op (statement: oracle/pizero_backward.py, which is pinned to the unmodified reference's
// autograd on all parameter tensors).  Gradients are accumulated in fp32, in the PACKED layouts the forward kernels read
// (fused q|k|v rows, gate|up blocks of 128, padded small matrices), so that an optimizer can update the kernel-side
// weights in place.
//
// B200 design: every matrix product of the backward runs through the same tcgen05 GEMM as the forward, with
// the operands read in place as MN-major UMMA operands (tokens are the contraction dimension of dW = dY^T X; W [out][in] is
// the MN-major B operand of dX = dY W) and the fp32 TMA reduce-add epilogue accumulating straight into the gradient
// buffer; nothing is recomputed except the normalised activations (one norm kernel each); 180 GB of HBM hold the ~30 GB of
// saved activations at 32 samples.  Attention backward: FlashAttention-2 style dQ and dK/dV kernels on mma.sync (bf16),
// a SIMT kernel for the fp32 build.
#include "api_internal.cuh"

namespace {

inline int rup(int x, int m) { return (x + m - 1) / m * m; }

// ============================================================== kernels ====
// dst = T(scale * src)
template <typename T>
__global__ void cast_scale_kernel(const float *__restrict__ src, T *__restrict__ dst, long n, float scale) {
    pdl_trigger();
    pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        dst[i] = from_f32<T>(src[i] * scale);
}
template <typename T>
void cast_scale(const float *src, T *dst, long n, float scale, cudaStream_t st) {
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_k(cast_scale_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, st, src, dst, n, scale);
}

PZ_DEVINL float gelu_tanh_grad(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float th = tanhf(k0 * (x + k1 * x * x * x));
    return 0.5f * (1.f + th) + 0.5f * x * (1.f - th * th) * k0 * (1.f + 3.f * k1 * x * x);
}

// GeGLU on the packed gate|up layout (blocks of PZ_GU_BLOCK gate columns then PZ_GU_BLOCK up columns):
// m = gelu_tanh(g) * u (paligemma/modules.py:86-95)
template <typename T> PZ_DEVINL void store8(T *p, const float *v);
template <> PZ_DEVINL void store8<float>(float *p, const float *v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = v[i];
}
template <> PZ_DEVINL void store8<bf16>(bf16 *p, const float *v) {
    *reinterpret_cast<uint4 *>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
template <typename T> PZ_DEVINL void load8(const T *p, float *o);

// one thread = 8 consecutive output columns (inside one block of PZ_GU_BLOCK); `total` counts groups of 8
template <typename T>
__global__ void geglu_fwd_kernel(const T *__restrict__ gu, T *__restrict__ m, long total, int inter) {
    pdl_trigger();
    pdl_wait();
    const int per_row = inter / 8;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total / 8; i += (long)gridDim.x * blockDim.x) {
        const long r = i / per_row;
        const int c = (int)(i % per_row) * 8;
        const long o = r * 2 * inter + (long)(c / PZ_GU_BLOCK) * 2 * PZ_GU_BLOCK + (c % PZ_GU_BLOCK);
        float g[8], u[8], out[8];
        load8<T>(gu + o, g);
        load8<T>(gu + o + PZ_GU_BLOCK, u);
#pragma unroll
        for (int k = 0; k < 8; ++k) out[k] = gelu_tanh(g[k]) * u[k];
        store8<T>(m + r * inter + c, out);
    }
}
// dg = dm u gelu'(g), du = dm gelu(g), written in the same packed layout
template <typename T>
__global__ void geglu_bwd_kernel(const T *__restrict__ gu, const T *__restrict__ dm, T *__restrict__ dgu, long total, int inter) {
    pdl_trigger();
    pdl_wait();
    const int per_row = inter / 8;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total / 8; i += (long)gridDim.x * blockDim.x) {
        const long r = i / per_row;
        const int c = (int)(i % per_row) * 8;
        const long o = r * 2 * inter + (long)(c / PZ_GU_BLOCK) * 2 * PZ_GU_BLOCK + (c % PZ_GU_BLOCK);
        float g[8], u[8], d[8], dg[8], du[8];
        load8<T>(gu + o, g);
        load8<T>(gu + o + PZ_GU_BLOCK, u);
        load8<T>(dm + r * inter + c, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) { dg[k] = d[k] * u[k] * gelu_tanh_grad(g[k]); du[k] = d[k] * gelu_tanh(g[k]); }
        store8<T>(dgu + o, dg);
        store8<T>(dgu + o + PZ_GU_BLOCK, du);
    }
}
inline unsigned ew_blocks(long total) {
    long b = (total + 255) / 256;
    return (unsigned)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

// act = gelu_tanh(pre) (SigLIP MLP, siglip.py:188-192); the pre-activation is kept for the backward
template <typename T>
__global__ void gelu_fwd_kernel(const T *__restrict__ pre, T *__restrict__ out, long total) {
    pdl_trigger();
    pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        out[i] = from_f32<T>(gelu_tanh(to_f32<T>(pre[i])));
}
// dpre = dpost * act'(pre)
template <typename T>
__global__ void gelu_bwd_kernel(const T *__restrict__ pre, const T *__restrict__ dpost, T *__restrict__ dpre, long total) {
    pdl_trigger();
    pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        dpre[i] = from_f32<T>(to_f32<T>(dpost[i]) * gelu_tanh_grad(to_f32<T>(pre[i])));
}
// action encoder: z = silu(zpre + bias[sample]); dzpre = dz * silu'(.)
template <typename T>
__global__ void silu_bwd_kernel(const float *__restrict__ zpre, const float *__restrict__ bias, const float *__restrict__ dz,
                                T *__restrict__ out, long total, int cols, int rows_per_sample) {
    pdl_trigger();
    pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / cols;
        const int c = (int)(i % cols);
        const float x = zpre[i] + bias[(r / rows_per_sample) * cols + c];
        const float s = 1.f / (1.f + __expf(-x));
        out[i] = from_f32<T>(dz[i] * s * (1.f + x * (1.f - s)));
    }
}

// Norm backward, one WARP per row (the row and its gradient stay in registers, statistics by shuffles -- the first version
// walked rows with a whole CTA and four block-wide reductions per row: 82 us for 8192 x 1152 against 27 us of HBM time).
// A CTA's 8 warps walk rows r = 8 * blockIdx.x + warp, + 8 * gridDim.x, ...; the per-column weight (and bias) gradients are
// kept per lane across the warp's rows, reduced across the 8 warps through shared memory and added with one atomic per
// column and CTA.  NI = float4 chunks per lane (cols <= 128 * NI).  dx is accumulated in place; `dx_cast` (optional) gets the
// model-dtype copy of the new dx (the dY operand of the product that follows).
//   RMSNorm  (paligemma/modules.py:13-21): y = x r (1 + w), r = (mean x^2 + eps)^-1/2
//            dx += r (1 + w) dy - x r^3 mean(x (1 + w) dy);   dw += sum_rows dy x r
//   LayerNorm (siglip.py:211,217,298):     y = xh w + b, xh = (x - mean) rstd
//            dx += rstd (g - mean g - xh mean(g xh)), g = dy w;   dw += sum dy xh;   db += sum dy
template <typename T, int NI, bool LAYER>
__global__ void __launch_bounds__(256) norm_bwd_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                       const float *__restrict__ dy, float *__restrict__ dx, float *__restrict__ dw,
                                                       float *__restrict__ db, long rows, int cols, float eps, T *__restrict__ dx_cast) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float red[];          // [8][cols] (and a second [8][cols] for the bias gradient)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n4 = cols >> 2;
    float4 dwl[NI], dbl[LAYER ? NI : 1];     // (the norm weight is re-read per row from L1: three row-sized register arrays are the budget)
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        dwl[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (LAYER) dbl[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float inv_n = 1.f / (float)cols;
    for (long r = (long)blockIdx.x * 8 + warp; r < rows; r += (long)gridDim.x * 8) {
        const float4 *xr = reinterpret_cast<const float4 *>(x + r * cols), *gr = reinterpret_cast<const float4 *>(dy + r * cols);
        float4 xv[NI], gv[NI];
        float s1 = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = lane + 32 * i;
            xv[i] = c < n4 ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            gv[i] = c < n4 ? gr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            s1 += LAYER ? (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w) : (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
        }
        s1 = warp_sum(s1);
        float mean = 0.f, rstd;
        if (LAYER) {
            mean = s1 * inv_n;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NI; ++i)
                if (lane + 32 * i < n4) {
                    const float d0 = xv[i].x - mean, d1 = xv[i].y - mean, d2 = xv[i].z - mean, d3 = xv[i].w - mean;
                    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
                }
            rstd = rsqrtf(warp_sum(q) * inv_n + eps);
        } else {
            rstd = rsqrtf(s1 * inv_n + eps);
        }
        // g = dy * (w or 1 + w); xh = normalised x; the two row means
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            if (lane + 32 * i < n4) {
                float *xp = reinterpret_cast<float *>(&xv[i]), *gp = reinterpret_cast<float *>(&gv[i]);
                const float4 w4 = __ldg(reinterpret_cast<const float4 *>(w) + lane + 32 * i);
                const float *wp = reinterpret_cast<const float *>(&w4);
                float *dwp = reinterpret_cast<float *>(&dwl[i]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float xh = LAYER ? (xp[k] - mean) * rstd : xp[k] * rstd;
                    dwp[k] += gp[k] * xh;
                    if (LAYER) reinterpret_cast<float *>(&dbl[i])[k] += gp[k];
                    const float g = gp[k] * (LAYER ? wp[k] : 1.f + wp[k]);
                    sg += g; sgx += g * xh;
                    xp[k] = xh; gp[k] = g;
                }
            }
        }
        sgx = warp_sum(sgx) * inv_n;
        if (LAYER) sg = warp_sum(sg) * inv_n; else sg = 0.f;
        float4 *dxr = reinterpret_cast<float4 *>(dx + r * cols);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = lane + 32 * i;
            if (c < n4) {
                float4 o = dxr[c];
                o.x += rstd * (gv[i].x - sg - xv[i].x * sgx); o.y += rstd * (gv[i].y - sg - xv[i].y * sgx);
                o.z += rstd * (gv[i].z - sg - xv[i].z * sgx); o.w += rstd * (gv[i].w - sg - xv[i].w * sgx);
                dxr[c] = o;
                if (dx_cast) {
                    const float ov[8] = {o.x, o.y, o.z, o.w, 0.f, 0.f, 0.f, 0.f};
                    T *dst = dx_cast + r * cols + 4 * c;
#pragma unroll
                    for (int k = 0; k < 4; ++k) dst[k] = from_f32<T>(ov[k]);
                }
            }
        }
    }
    // reduce the per-warp column sums across the CTA, then one atomic per column
    for (int pass = 0; pass < (LAYER ? 2 : 1); ++pass) {
        float *out = pass ? db : dw;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = lane + 32 * i;
            if (c < n4) reinterpret_cast<float4 *>(red + warp * cols)[c] = pass ? dbl[LAYER ? i : 0] : dwl[i];
        }
        __syncthreads();
        if (out)
            for (int c = threadIdx.x; c < cols; c += 256) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) a += red[k * cols + c];
                atomicAdd(out + c, a);
            }
    }
}
template <typename T, bool LAYER>
void norm_bwd(const float *x, const float *w, const float *dy, float *dx, float *dw, float *db, long rows, int cols, cudaStream_t st,
              T *dx_cast) {
    long blocks = (rows + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    const size_t smem = (size_t)8 * cols * sizeof(float);
    auto go = [&](auto kern) {
        // (the three instantiations share one function-pointer type, so a "once" flag inside this lambda would be shared too)
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2048 * 4);
        launch_k(kern, dim3((unsigned)blocks), dim3(256), smem, st, x, w, dy, dx, dw, db, rows, cols, 1e-6f, dx_cast);
    };
    if (cols <= 512) go(norm_bwd_kernel<T, 4, LAYER>);
    else if (cols <= 1280) go(norm_bwd_kernel<T, 10, LAYER>);
    else go(norm_bwd_kernel<T, 16, LAYER>);
}
template <typename T>
void rmsnorm_bwd(const float *x, const float *w, const float *dy, float *dx, float *dw, long rows, int cols, cudaStream_t st,
                 T *dx_cast = nullptr) {
    norm_bwd<T, false>(x, w, dy, dx, dw, nullptr, rows, cols, st, dx_cast);
}
template <typename T>
void layernorm_bwd(const float *x, const float *w, const float *dy, float *dx, float *dw, float *db, long rows, int cols,
                   cudaStream_t st, T *dx_cast = nullptr) {
    norm_bwd<T, true>(x, w, dy, dx, dw, db, rows, cols, st, dx_cast);
}

// out[c] += sum_r in[r][c] (bias gradients); block (32, 8), grid (column tiles, row slabs)
template <typename T>
__global__ void colsum_kernel(const T *__restrict__ in, int ld, float *__restrict__ out, long rows, int cols) {
    pdl_trigger();
    pdl_wait();
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < cols)
        for (long r = (long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long)gridDim.y * 8) acc += to_f32<T>(in[r * ld + c]);
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
        atomicAdd(out + c, s);
    }
}
template <typename T>
void colsum(const T *in, int ld, float *out, long rows, int cols, cudaStream_t st) {
    if (!out) return;
    long slabs = (rows + 255) / 256;
    if (slabs > 64) slabs = 64;
    if (slabs < 1) slabs = 1;
    launch_k(colsum_kernel<T>, dim3((cols + 31) / 32, (unsigned)slabs), dim3(32, 8), 0, st, in, ld, out, rows, cols);
}

// out[g][c] = sum_{r < group} in[g * group + r][c] (T -> T): the time half of the action encoder sees one row per sample
template <typename T>
__global__ void group_sum_kernel(const T *__restrict__ in, T *__restrict__ out, long total, int cols, int group) {
    pdl_trigger();
    pdl_wait();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long g = i / cols;
    const int c = (int)(i % cols);
    float s = 0.f;
    for (int r = 0; r < group; ++r) s += to_f32<T>(in[(g * group + r) * cols + c]);
    out[i] = from_f32<T>(s);
}
// out[p][c] += sum_img in[img * period + p][c] (fp32): position-embedding gradient
__global__ void period_sum_kernel(const float *__restrict__ in, float *__restrict__ out, long total, int cols, int period, int n) {
    pdl_trigger();
    pdl_wait();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float s = 0.f;
    for (int k = 0; k < n; ++k) s += in[(long)k * period * cols + i];
    out[i] += s;
}

// d loss / d v of mean((v - (x1 - (1 - sig) x0))^2) (pizero.py:658-661), as a [rows, 8] matrix (pad columns zero)
template <typename T>
__global__ void loss_bwd_kernel(const float *__restrict__ vel, int vel_ld, const float *__restrict__ x0, const float *__restrict__ x1,
                                T *__restrict__ dv, long rows, int adim, float sig_min, float loss_scale) {
    pdl_trigger();
    pdl_wait();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 8) return;
    const long r = i / 8;
    const int c = (int)(i % 8);
    float g = 0.f;
    if (c < adim) {
        const long j = r * adim + c;
        g = 2.f * (vel[r * vel_ld + c] - (x1[j] - (1.f - sig_min) * x0[j])) / (float)(rows * adim) * loss_scale;
    }
    dv[i] = from_f32<T>(g);
}

// Backward of RoPE + the Q/K/V split (model/utils.py:4-16): y1 = x1 c - x2 s, y2 = x2 c + x1 s  =>
// dx1 = dy1 c + dy2 s, dx2 = dy2 c - dy1 s.  Assembles the gradient of the raw fused projection [q | k | v] of one
// mixture from dq [M, nh*hd] (fp32, may be null = zero) and the rows `row_off ..` of the joint dK / dV [B][S_all][hd].
template <typename T>
__global__ void rope_bwd_merge_kernel(const float *__restrict__ dq, const float *__restrict__ dK, const float *__restrict__ dV,
                                      long dkv_bs, int row_off, const float *__restrict__ cos_t, const float *__restrict__ sin_t,
                                      int pos0, T *__restrict__ out, int s_x, int nh, int hd) {
    pdl_trigger();
    pdl_wait();
    const long row = blockIdx.x;
    const int b = (int)(row / s_x), s = (int)(row % s_x);
    const int half = hd >> 1, qkvd = (nh + 2) * hd;
    const float *cs = cos_t + (long)(pos0 + s) * half, *sn = sin_t + (long)(pos0 + s) * half;
    const float *gk = dK + (long)b * dkv_bs + (long)(row_off + s) * hd;
    const float *gv = dV + (long)b * dkv_bs + (long)(row_off + s) * hd;
    T *o = out + row * qkvd;
    for (int i = threadIdx.x; i < (nh + 1) * half; i += blockDim.x) {
        const int h = i / half, d = i % half;
        const float *g = h < nh ? (dq ? dq + row * nh * hd + (long)h * hd : nullptr) : gk;
        const float g1 = g ? g[d] : 0.f, g2 = g ? g[d + half] : 0.f;
        o[h * hd + d] = from_f32<T>(g1 * cs[d] + g2 * sn[d]);
        o[h * hd + d + half] = from_f32<T>(g2 * cs[d] - g1 * sn[d]);
    }
    for (int i = threadIdx.x; i < hd; i += blockDim.x) o[(nh + 1) * hd + i] = from_f32<T>(gv[i]);
}

// eight consecutive elements as fp32 (16-byte load for bf16 when the address allows it)
template <typename T> PZ_DEVINL void load8(const T *p, float *o);
template <> PZ_DEVINL void load8<float>(const float *p, float *o) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = p[i];
}
template <> PZ_DEVINL void load8<bf16>(const bf16 *p, float *o) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        o[0] = bf16lo(v.x); o[1] = bf16hi(v.x); o[2] = bf16lo(v.y); o[3] = bf16hi(v.y);
        o[4] = bf16lo(v.z); o[5] = bf16hi(v.z); o[6] = bf16lo(v.w); o[7] = bf16hi(v.w);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __bfloat162float(p[i]);
    }
}

// --------------------------------------------------- attention backward ----
// o = P V, P = softmax(cap tanh(s / cap) + mask), s = scale q k^T (joint_model.py:255-283; cap = 0: SigLIP, siglip.py:133-152).
// A CTA owns 32 (query row, head) pairs that share their keys: 4 rows x 8 heads on the single K/V head of the joint
// model (MQA), 32 rows of one head for SigLIP.  Phase 1: logits of the 32 pairs against every key (thread = key);
// phase 2: softmax statistics (warp = 4 pairs); phase 3 (thread = key): dP = dO V^T, dS = P (dP - D) (1 - tanh^2) scale,
// dV_j += sum_p P dO_p and dK_j += sum_p dS q_p as fp32 atomics (the other row tiles add to the same key rows);
// phase 4 (thread = feature): dQ_p = sum_j dS_pj K_j, owned by this CTA.  Rows of a tile never straddle a segment
// (vlm | proprio | action), so key visibility is uniform across the tile (block mask, pizero.py:271-310).
struct AttnBwdArgs {
    int n_seg;
    int seg_rows[3];
    int seg_active[3];
    const void *Q[3], *dO[3], *O[3];   // T
    float *dQ[3];                      // fp32, geometry (dq_rs, head stride hd)
    int q_rs, o_rs, dq_rs;             // row strides in elements
    const void *K, *V;                 // T
    long kv_bs; int kv_rs, kv_hs;
    float *dK, *dV;                    // fp32
    long dkv_bs; int dkv_rs, dkv_hs;
    const int32_t *valid_len;
    int s_v, s_p;                      // mask boundaries (valid_len != null)
    int n_keys, batch, n_heads, hd, hpc;   // hpc: heads per CTA (n_heads when kv_hs == 0 and n_heads divides 32, else 1)
    float scale, softcap;
    int sp;                            // padded key count of the shared-memory score rows (multiple of 4)
};

template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_kernel(const AttnBwdArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float smf[];
    const int hd = a.hd, SP = a.sp;
    float *qs = smf;                    // [32][hd]
    float *dos = qs + 32 * hd;          // [32][hd]
    float *sb = dos + 32 * hd;          // [32][SP]
    float *Dp = sb + 32 * SP;           // [32]
    float *mx = Dp + 32, *inv = mx + 32, *actf = inv + 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int QT = 32 / a.hpc;
    // which segment / row tile
    int seg = 0, tile = blockIdx.x, seg_off = 0;
    for (;;) {
        int tiles = a.seg_active[seg] ? (a.seg_rows[seg] + QT - 1) / QT : 0;
        if (tile < tiles) break;
        tile -= tiles; seg_off += a.seg_rows[seg]; ++seg;
        if (seg >= a.n_seg) return;
    }
    const int b = blockIdx.z, head0 = blockIdx.y * a.hpc;
    const int rows = a.seg_rows[seg];
    const int vl = a.valid_len ? a.valid_len[b] : (1 << 30);
    const T *Qp = (const T *)a.Q[seg], *dOp = (const T *)a.dO[seg], *Op = (const T *)a.O[seg];
    // ---- load q, dO (fp32 in shared memory), D = dO . O, active flags
    for (int p = warp * 4; p < warp * 4 + 4; ++p) {
        const int hi = p / QT, ri = p % QT;
        const int row = tile * QT + ri, head = head0 + hi;
        const bool act = row < rows && !(a.valid_len && seg_off + row < a.s_v && seg_off + row >= vl);
        float dsum = 0.f;
        for (int d = lane; d < hd; d += 32) {
            float qv = 0.f, gv = 0.f, ov = 0.f;
            if (act) {
                const long r = (long)b * rows + row;
                qv = to_f32<T>(Qp[r * a.q_rs + head * hd + d]);
                gv = to_f32<T>(dOp[r * a.o_rs + head * hd + d]);
                ov = to_f32<T>(Op[r * a.o_rs + head * hd + d]);
            }
            qs[p * hd + d] = qv; dos[p * hd + d] = gv;
            dsum += gv * ov;
        }
        dsum = warp_sum(dsum);
        if (lane == 0) { Dp[p] = dsum; actf[p] = act ? 1.f : 0.f; }
    }
    __syncthreads();
    const T *Kb = (const T *)a.K + (long)b * a.kv_bs + (long)head0 * a.kv_hs;
    const T *Vb = (const T *)a.V + (long)b * a.kv_bs + (long)head0 * a.kv_hs;
    auto visible = [&](int j) -> bool {
        if (!a.valid_len) return true;
        if (j < a.s_v) return j < vl;
        if (j < a.s_v + a.s_p) return seg >= 1;
        return seg >= 2;
    };
    // ---- phase 1: logits
    for (int j = tid; j < SP; j += 256) {
        if (j >= a.n_keys || !visible(j)) {
            for (int p = 0; p < 32; ++p) sb[p * SP + j] = -INFINITY;
            continue;
        }
        float acc[32];
#pragma unroll
        for (int p = 0; p < 32; ++p) acc[p] = 0.f;
        const T *kr = Kb + (long)j * a.kv_rs;
        for (int c = 0; c < hd; c += 8) {
            float kv[8];
            load8<T>(kr + c, kv);
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const float4 q0 = *reinterpret_cast<const float4 *>(qs + p * hd + c);
                const float4 q1 = *reinterpret_cast<const float4 *>(qs + p * hd + c + 4);
                acc[p] += kv[0] * q0.x + kv[1] * q0.y + kv[2] * q0.z + kv[3] * q0.w + kv[4] * q1.x + kv[5] * q1.y + kv[6] * q1.z +
                          kv[7] * q1.w;
            }
        }
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            const float s = acc[p] * a.scale;
            sb[p * SP + j] = a.softcap > 0.f ? a.softcap * tanhf(s / a.softcap) : s;
        }
    }
    __syncthreads();
    // ---- phase 2: softmax statistics
    for (int p = warp * 4; p < warp * 4 + 4; ++p) {
        float m = -INFINITY;
        for (int j = lane; j < a.n_keys; j += 32) m = fmaxf(m, sb[p * SP + j]);
        m = warp_max(m);
        float s = 0.f;
        if (m > -INFINITY)
            for (int j = lane; j < a.n_keys; j += 32) s += __expf(sb[p * SP + j] - m);
        s = warp_sum(s);
        if (lane == 0) { mx[p] = m; inv[p] = (s > 0.f && actf[p] > 0.f) ? 1.f / s : 0.f; }
    }
    __syncthreads();
    // ---- phase 3: dP, dS; dV and dK
    for (int j = tid; j < SP; j += 256) {
        if (j >= a.n_keys || !visible(j)) {
            for (int p = 0; p < 32; ++p) sb[p * SP + j] = 0.f;
            continue;
        }
        float acc[32], pr[32];
#pragma unroll
        for (int p = 0; p < 32; ++p) acc[p] = 0.f;
        const T *vr = Vb + (long)j * a.kv_rs;
        for (int c = 0; c < hd; c += 8) {
            float vv[8];
            load8<T>(vr + c, vv);
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const float4 g0 = *reinterpret_cast<const float4 *>(dos + p * hd + c);
                const float4 g1 = *reinterpret_cast<const float4 *>(dos + p * hd + c + 4);
                acc[p] += vv[0] * g0.x + vv[1] * g0.y + vv[2] * g0.z + vv[3] * g0.w + vv[4] * g1.x + vv[5] * g1.y + vv[6] * g1.z +
                          vv[7] * g1.w;
            }
        }
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            const float lg = sb[p * SP + j];
            const float pv = inv[p] > 0.f ? __expf(lg - mx[p]) * inv[p] : 0.f;
            const float th = a.softcap > 0.f ? lg / a.softcap : 0.f;
            pr[p] = pv;
            acc[p] = pv * (acc[p] - Dp[p]) * (1.f - th * th) * a.scale;   // dS
            sb[p * SP + j] = acc[p];
        }
        // MQA: all heads of the tile add into the same key row; SigLIP: hpc == 1
        float *dVr = a.dV + (long)b * a.dkv_bs + (long)j * a.dkv_rs + (long)head0 * a.dkv_hs;
        float *dKr = a.dK + (long)b * a.dkv_bs + (long)j * a.dkv_rs + (long)head0 * a.dkv_hs;
        for (int c = 0; c < hd; c += 8) {
            float av[8], ak[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = ak[i] = 0.f;
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const float4 g0 = *reinterpret_cast<const float4 *>(dos + p * hd + c);
                const float4 g1 = *reinterpret_cast<const float4 *>(dos + p * hd + c + 4);
                const float4 q0 = *reinterpret_cast<const float4 *>(qs + p * hd + c);
                const float4 q1 = *reinterpret_cast<const float4 *>(qs + p * hd + c + 4);
                av[0] += pr[p] * g0.x; av[1] += pr[p] * g0.y; av[2] += pr[p] * g0.z; av[3] += pr[p] * g0.w;
                av[4] += pr[p] * g1.x; av[5] += pr[p] * g1.y; av[6] += pr[p] * g1.z; av[7] += pr[p] * g1.w;
                ak[0] += acc[p] * q0.x; ak[1] += acc[p] * q0.y; ak[2] += acc[p] * q0.z; ak[3] += acc[p] * q0.w;
                ak[4] += acc[p] * q1.x; ak[5] += acc[p] * q1.y; ak[6] += acc[p] * q1.z; ak[7] += acc[p] * q1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { atomicAdd(dVr + c + i, av[i]); atomicAdd(dKr + c + i, ak[i]); }
        }
    }
    __syncthreads();
    // ---- phase 4: dQ
    if (tid < hd) {
        float acc[32];
#pragma unroll
        for (int p = 0; p < 32; ++p) acc[p] = 0.f;
        const T *kc = Kb + tid;
        for (int j = 0; j < a.n_keys; j += 4) {
            float k4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) k4[i] = (j + i < a.n_keys) ? to_f32<T>(kc[(long)(j + i) * a.kv_rs]) : 0.f;
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const float4 s4 = *reinterpret_cast<const float4 *>(sb + p * SP + j);
                acc[p] += s4.x * k4[0] + s4.y * k4[1] + s4.z * k4[2] + s4.w * k4[3];
            }
        }
        float *dQp = a.dQ[seg];
        for (int p = 0; p < 32; ++p) {
            const int hi = p / QT, ri = p % QT;
            const int row = tile * QT + ri;
            if (row < rows) dQp[((long)b * rows + row) * a.dq_rs + (head0 + hi) * hd + tid] = acc[p];
        }
    }
}

template <typename T>
int attn_bwd(AttnBwdArgs a, cudaStream_t st, const char **err) {
    if (a.hd % 8 || a.hd > 256) { *err = "attention backward: head_dim must be a multiple of 8, <= 256"; return PZ_ERR_INVALID; }
    a.hpc = (a.kv_hs == 0 && 32 % a.n_heads == 0) ? a.n_heads : 1;
    a.sp = rup(a.n_keys, 4);
    const int QT = 32 / a.hpc;
    int tiles = 0;
    for (int s = 0; s < a.n_seg; ++s) if (a.seg_active[s]) tiles += (a.seg_rows[s] + QT - 1) / QT;
    if (!tiles) return 0;
    size_t smem = (size_t)(64 * a.hd + 32 * a.sp + 128) * sizeof(float);
    if (smem > 220 * 1024) { *err = "attention backward: too many keys for the shared-memory score tile"; return PZ_ERR_INVALID; }
    static PerDeviceOnce once;
    if (once.need()) cudaFuncSetAttribute(attn_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    launch_k(attn_bwd_kernel<T>, dim3(tiles, a.n_heads / a.hpc, a.batch), dim3(256), smem, st, a);
    return 0;
}



// ------------------------------------- attention backward on tensor cores (bf16) ----
// Two kernels over 2-D views (FlashAttention-2 style split, so that nothing is accumulated with atomics):
//   attn_bwd_dq_kernel : a CTA owns 64 query rows; pass A walks the key blocks for the softmax statistics (log-sum-exp)
//                        and D = rowsum(dO o O); pass B recomputes S, forms dS and accumulates dQ = dS K in registers.
//   attn_bwd_dkv_kernel: a CTA owns 64 keys and walks the query tiles: dV += P^T dO, dK += dS^T Q in registers.
// "Query rows" are (token, head) pairs: with one K/V head (MQA) the 8 heads of a token are 8 consecutive rows of the
// [tokens * heads, head_dim] view of the q buffer, so one CTA serves all heads against one copy of K / V; SigLIP (one K/V
// head per q head) launches one group per head.  All products are bf16 mma.sync.m16n8k16 (fp32 accumulation) fed by
// ldmatrix from shared-memory tiles; P and dS are rounded to bf16 for the second product, as the forward rounds P.

struct AttnBwd2Args {
    const bf16 *Q, *dO, *O;
    float *dQ;
    long q_bs, q_gs; int ld_q;
    long o_bs, o_gs; int ld_o;
    long dq_bs, dq_gs; int ld_dq;
    const bf16 *K, *V;
    long kv_bs, kv_gs; int ld_kv;
    float *dK, *dV;
    long dkv_bs, dkv_gs; int ld_dkv;
    float *lse, *dvec;                 // [batch][groups][NQ]
    const int32_t *valid_len;
    int NQ, NK, hd, hpr;               // hpr: query rows per token (heads folded into the rows)
    int seg, s_v, s_p;                 // segment of the queries (block mask, pizero.py:271-310); valid_len == null: no mask
    float scale, softcap;
    int accumulate, groups;
    int qsplit;                        // dK/dV kernel: query tiles dealt round-robin to this many CTAs per key block, which
                                       // then ADD their partial sums with fp32 atomics (qsplit > 1 needs zeroed dK / dV)
};

constexpr int SLD = 72;               // row pitch of the 64 x 64 score tiles (floats / bf16s)
// hardware tanh (rel. error 2^-11; the forward kernels soft-cap with the same instruction)
PZ_DEVINL float tanh_hw(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// bf16 tiles are [64][HDP + 8]: a row pitch that is an odd multiple of 16 bytes keeps the 8 rows of a fragment load on
// different banks (a 512-byte pitch put them all on the same ones: measured 27 TFLOP/s)
template <int HDP>
PZ_DEVINL void load_rows64(bf16 *__restrict__ dst, const bf16 *__restrict__ src, int ld, int row0, int n_rows, int hd, int hpr,
                           int tok_limit) {
    // dst[r][0..HDP) <- src[(row0 + r) * ld + 0..hd), zero beyond hd, beyond n_rows and for tokens >= tok_limit.
    // All of a thread's 16-byte loads are issued before the first shared-memory store (one L2 latency per tile instead of
    // one per chunk: the stores used to wait for each load in turn), and the row limit is one comparison, no division.
    constexpr int CH = HDP / 8, LDP = HDP + 8, N = 64 * CH, IT = (N + 255) / 256;
    const long lim = (long)tok_limit * hpr;
    const int row_lim = (int)(lim < n_rows ? lim : n_rows);
    uint4 v[IT];
#pragma unroll
    for (int k = 0; k < IT; ++k) {
        const int i = threadIdx.x + k * 256;
        const int r = i / CH, c = (i % CH) * 8;
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        if (i < N && row0 + r < row_lim && c < hd) v[k] = __ldg(reinterpret_cast<const uint4 *>(src + (long)(row0 + r) * ld + c));
    }
#pragma unroll
    for (int k = 0; k < IT; ++k) {
        const int i = threadIdx.x + k * 256;
        const int r = i / CH, c = (i % CH) * 8;
        if (i < N) *reinterpret_cast<uint4 *>(dst + r * LDP + c) = v[k];
    }
}

// two tiles of the same geometry (K and V, or Q and dO): all loads of both in flight together
template <int HDP>
PZ_DEVINL void load_rows64x2(bf16 *__restrict__ dst0, const bf16 *__restrict__ src0, int ld0, bf16 *__restrict__ dst1,
                             const bf16 *__restrict__ src1, int ld1, int row0, int n_rows, int hd, int hpr, int tok_limit) {
    constexpr int CH = HDP / 8, LDP = HDP + 8, N = 64 * CH, IT = (N + 255) / 256;
    const long lim = (long)tok_limit * hpr;
    const int row_lim = (int)(lim < n_rows ? lim : n_rows);
    uint4 v0[IT], v1[IT];
#pragma unroll
    for (int k = 0; k < IT; ++k) {
        const int i = threadIdx.x + k * 256;
        const int r = i / CH, c = (i % CH) * 8;
        v0[k] = v1[k] = make_uint4(0u, 0u, 0u, 0u);
        if (i < N && row0 + r < row_lim && c < hd) {
            v0[k] = __ldg(reinterpret_cast<const uint4 *>(src0 + (long)(row0 + r) * ld0 + c));
            v1[k] = __ldg(reinterpret_cast<const uint4 *>(src1 + (long)(row0 + r) * ld1 + c));
        }
    }
#pragma unroll
    for (int k = 0; k < IT; ++k) {
        const int i = threadIdx.x + k * 256;
        const int r = i / CH, c = (i % CH) * 8;
        if (i < N) {
            *reinterpret_cast<uint4 *>(dst0 + r * LDP + c) = v0[k];
            *reinterpret_cast<uint4 *>(dst1 + r * LDP + c) = v1[k];
        }
    }
}

PZ_DEVINL bool key_visible(const AttnBwd2Args &a, int j, int vl) {
    if (j >= a.NK) return false;
    if (!a.valid_len) return true;
    if (j < a.s_v) return j < vl;
    if (j < a.s_v + a.s_p) return a.seg >= 1;
    return a.seg >= 2;
}


// ---- warp-level bf16 MMA plumbing: ldmatrix + mma.sync.m16n8k16 (the `wmma` API loads its fragments with generic LD
// instructions -- 2.7 per HMMA, 25 % of all issued instructions in the first version of these kernels; ldmatrix is one
// shared-memory instruction per 16 x 16 tile).  Fragment layouts (g = lane / 4, t = lane % 4):
//   A 16x16: a0 (g, 2t..) a1 (g+8, 2t..) a2 (g, 2t+8..) a3 (g+8, 2t+8..);  B 16x8: b0 (k 2t.., n g) b1 (k 2t+8.., n g);
//   C 16x8 : c0 c1 (g, 2t 2t+1), c2 c3 (g+8, 2t 2t+1).  A "16 x 16 output tile" below is two n8 tiles: c[0..3], c[4..7].
PZ_DEVINL void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
PZ_DEVINL void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
PZ_DEVINL uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// A tile (rows r0.., k k0..) of a row-major [row][k] array with pitch `ld` elements
PZ_DEVINL void frag_a(uint32_t (&a)[4], const bf16 *base, int ld, int lane) {
    ldsm_x4(a, smem_addr(base + ((lane & 7) + 8 * ((lane >> 3) & 1)) * ld + 8 * (lane >> 4)));
}
// A tile of an array stored [k][m] (the transpose of what the MMA wants: P^T from P[query][key])
PZ_DEVINL void frag_a_t(uint32_t (&a)[4], const bf16 *base, int ld, int lane) {
    ldsm_x4_t(a, smem_addr(base + ((lane & 7) + 8 * (lane >> 4)) * ld + 8 * ((lane >> 3) & 1)));
}
// B for two n8 tiles (n n0 .. n0+15, k k0 .. k0+15) of an array stored [n][k]: b[0], b[1] = first n8 tile, b[2], b[3] = second
PZ_DEVINL void frag_b_nk(uint32_t (&b)[4], const bf16 *base, int ld, int lane) {
    ldsm_x4(b, smem_addr(base + ((lane & 7) + 8 * (lane >> 4)) * ld + 8 * ((lane >> 3) & 1)));
}
// the same of an array stored [k][n]
PZ_DEVINL void frag_b_kn(uint32_t (&b)[4], const bf16 *base, int ld, int lane) {
    ldsm_x4_t(b, smem_addr(base + ((lane & 7) + 8 * ((lane >> 3) & 1)) * ld + 8 * (lane >> 4)));
}
PZ_DEVINL void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 16 x 16 fp32 output tile (two n8 accumulators) -> row-major memory with pitch `ld` floats
PZ_DEVINL void store_c16(float *base, int ld, const float (&c0)[4], const float (&c1)[4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    *reinterpret_cast<float2 *>(base + g * ld + 2 * t) = make_float2(c0[0], c0[1]);
    *reinterpret_cast<float2 *>(base + (g + 8) * ld + 2 * t) = make_float2(c0[2], c0[3]);
    *reinterpret_cast<float2 *>(base + g * ld + 8 + 2 * t) = make_float2(c1[0], c1[1]);
    *reinterpret_cast<float2 *>(base + (g + 8) * ld + 8 + 2 * t) = make_float2(c1[2], c1[3]);
}

// S[64 x 64] = A[64 x HDP] . B[64 x HDP]^T into the fp32 tile `out` (8 warps: 4 row tiles x 2 column halves)
template <int HDP>
PZ_DEVINL void mma_qkT(const bf16 *A, const bf16 *B, float *out, int warp) {
    constexpr int LDP = HDP + 8;
    const int wr = warp >> 1, wc = warp & 1, lane = threadIdx.x & 31;
    float acc[4][4];      // four n8 tiles: columns wc * 32 + 8 j
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 4
    for (int k = 0; k < HDP / 16; ++k) {
        uint32_t fa[4], fb0[4], fb1[4];
        frag_a(fa, A + wr * 16 * LDP + k * 16, LDP, lane);
        frag_b_nk(fb0, B + (wc * 32) * LDP + k * 16, LDP, lane);
        frag_b_nk(fb1, B + (wc * 32 + 16) * LDP + k * 16, LDP, lane);
        mma16816(acc[0], fa, fb0[0], fb0[1]);
        mma16816(acc[1], fa, fb0[2], fb0[3]);
        mma16816(acc[2], fa, fb1[0], fb1[1]);
        mma16816(acc[3], fa, fb1[2], fb1[3]);
    }
    store_c16(out + wr * 16 * SLD + wc * 32, SLD, acc[0], acc[1], lane);
    store_c16(out + wr * 16 * SLD + wc * 32 + 16, SLD, acc[2], acc[3], lane);
}

template <int HDP>
__global__ void __launch_bounds__(256, HDP <= 128 ? 2 : 1) attn_bwd_dq_kernel(const AttnBwd2Args a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ uint8_t smraw[];
    uint8_t *sm = align_smem(smraw, 128);
    constexpr int LDP = HDP + 8;
    bf16 *Qs = (bf16 *)sm, *dOs = Qs + 64 * LDP, *Ks = dOs + 64 * LDP, *Vs = Ks + 64 * LDP;
    float *Sf = (float *)(Vs + 64 * LDP), *Df = Sf + 64 * SLD;
    bf16 *Pb = (bf16 *)(Df + 64 * SLD);
    float *sD = (float *)(Pb + 64 * SLD), *sL = sD + 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, g = blockIdx.y, row0 = blockIdx.x * 64;
    const int vl = a.valid_len ? a.valid_len[b] : (1 << 30);
    const int tok_limit = (a.valid_len && a.seg == 0) ? vl : (1 << 30);
    const bf16 *Qg = a.Q + b * a.q_bs + g * a.q_gs, *dOg = a.dO + b * a.o_bs + g * a.o_gs, *Og = a.O + b * a.o_bs + g * a.o_gs;
    const bf16 *Kg = a.K + b * a.kv_bs + g * a.kv_gs, *Vg = a.V + b * a.kv_bs + g * a.kv_gs;
    load_rows64x2<HDP>(Qs, Qg, a.ld_q, dOs, dOg, a.ld_o, row0, a.NQ, a.hd, a.hpr, tok_limit);
    __syncthreads();
    // D = rowsum(dO o O): a warp takes 8 rows
    for (int r = warp * 8; r < warp * 8 + 8; ++r) {
        const int row = row0 + r;
        float acc = 0.f;
        if (row < a.NQ && (long)row < (long)tok_limit * a.hpr)
            for (int d = lane; d < a.hd; d += 32) acc += __bfloat162float(dOs[r * LDP + d]) * __bfloat162float(Og[(long)row * a.ld_o + d]);
        acc = warp_sum(acc);
        if (lane == 0) sD[r] = acc;
    }
    const float inv_cap = a.softcap > 0.f ? 1.f / a.softcap : 0.f;
    const int er = tid >> 2, ec0 = (tid & 3) * 16;      // elementwise mapping: 4 threads per row, 16 columns each
    const bool row_ok = row0 + er < a.NQ && (long)(row0 + er) < (long)tok_limit * a.hpr;
    const int n_kb = (a.NK + 63) / 64;
    // Soft-capped logits lie in [-cap, cap]: exp(l - cap / 2) can neither overflow nor leave the normal fp32 range, so the
    // softmax needs no running maximum and dQ is linear in 1 / Z -- ONE pass accumulates the unnormalised sum and Z, the
    // division happens at the end.  Without a cap (SigLIP) pass A computes the log-sum-exp first.
    const bool single = a.softcap > 0.f;
    const float shift = 0.5f * a.softcap;
    // ---- pass A: log-sum-exp of every row
    float m_run = -INFINITY, s_run = 0.f;
    for (int kb = 0; !single && kb < n_kb; ++kb) {
        __syncthreads();
        load_rows64<HDP>(Ks, Kg, a.ld_kv, kb * 64, a.NK, a.hd, 1, 1 << 30);
        __syncthreads();
        mma_qkT<HDP>(Qs, Ks, Sf, warp);
        __syncthreads();
        float lg[16], mloc = -INFINITY;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float sv = Sf[er * SLD + ec0 + c] * a.scale;
            const float l = a.softcap > 0.f ? a.softcap * tanh_hw(sv * inv_cap) : sv;
            lg[c] = key_visible(a, kb * 64 + ec0 + c, vl) ? l : -INFINITY;
            mloc = fmaxf(mloc, lg[c]);
        }
        mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 1));
        mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 2));
        const float m_new = fmaxf(m_run, mloc);
        float sloc = 0.f;
        if (m_new > -INFINITY) {
#pragma unroll
            for (int c = 0; c < 16; ++c) sloc += __expf(lg[c] - m_new);
        }
        sloc += __shfl_xor_sync(0xffffffffu, sloc, 1);
        sloc += __shfl_xor_sync(0xffffffffu, sloc, 2);
        s_run = (m_run > -INFINITY ? s_run * __expf(m_run - m_new) : 0.f) + sloc;
        m_run = m_new;
    }
    float lse = (row_ok && s_run > 0.f) ? m_run + __logf(s_run) : INFINITY;   // +inf: P = exp(l - lse) = 0
    if (single) lse = row_ok ? shift : INFINITY;
    __syncthreads();
    if ((tid & 3) == 0) sL[er] = lse;
    float z_run = 0.f;
    // ---- pass B: dQ = sum over key blocks of dS K
    constexpr int NCT = HDP / 16, HALF = (NCT + 1) / 2;
    const int wr = warp >> 1, wc = warp & 1;
    const int lane_ = tid & 31;
    float accq[HALF][2][4];       // [16-column tile][n8 half][fragment]
#pragma unroll
    for (int t = 0; t < HALF; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) accq[t][0][i] = accq[t][1][i] = 0.f;
    for (int kb = 0; kb < n_kb; ++kb) {
        __syncthreads();
        load_rows64x2<HDP>(Ks, Kg, a.ld_kv, Vs, Vg, a.ld_kv, kb * 64, a.NK, a.hd, 1, 1 << 30);
        __syncthreads();
        mma_qkT<HDP>(Qs, Ks, Sf, warp);
        mma_qkT<HDP>(dOs, Vs, Df, warp);
        __syncthreads();
        const float Dr = sD[er], Lr = sL[er];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float sv = Sf[er * SLD + ec0 + c] * a.scale;
            const float l = a.softcap > 0.f ? a.softcap * tanh_hw(sv * inv_cap) : sv;
            const float th = l * inv_cap;
            const float pv = key_visible(a, kb * 64 + ec0 + c, vl) ? __expf(l - Lr) : 0.f;
            z_run += pv;
            Pb[er * SLD + ec0 + c] = __float2bfloat16_rn(pv * (Df[er * SLD + ec0 + c] - Dr) * (1.f - th * th) * a.scale);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t fa[4];
            frag_a(fa, Pb + wr * 16 * SLD + kk * 16, SLD, lane_);
#pragma unroll
            for (int t = 0; t < HALF; ++t) {
                const int ct = wc * HALF + t;
                if (ct < NCT) {
                    uint32_t fb[4];
                    frag_b_kn(fb, Ks + kk * 16 * LDP + ct * 16, LDP, lane_);
                    mma16816(accq[t][0], fa, fb[0], fb[1]);
                    mma16816(accq[t][1], fa, fb[2], fb[3]);
                }
            }
        }
    }
    // the statistics the dK / dV kernel needs: log-sum-exp and D of every row; single pass: Z is known only now
    z_run += __shfl_xor_sync(0xffffffffu, z_run, 1);
    z_run += __shfl_xor_sync(0xffffffffu, z_run, 2);
    if (single) lse = (row_ok && z_run > 0.f) ? shift + __logf(z_run) : INFINITY;
    __syncthreads();
    if ((tid & 3) == 0) {
        sL[er] = single ? ((row_ok && z_run > 0.f) ? 1.f / z_run : 0.f) : 1.f;      // row scale of the accumulated dQ
        if (row0 + er < a.NQ) {
            const long o = ((long)b * a.groups + g) * a.NQ + row0 + er;
            a.lse[o] = lse;
            a.dvec[o] = sD[er];
        }
    }
    float *stage = (float *)Ks;    // 64 x HDP fp32 = the K and V tiles
#pragma unroll
    for (int t = 0; t < HALF; ++t) {
        const int ct = wc * HALF + t;
        if (ct < NCT) store_c16(stage + wr * 16 * HDP + ct * 16, HDP, accq[t][0], accq[t][1], lane_);
    }
    __syncthreads();
    float *dQg = a.dQ + b * a.dq_bs + g * a.dq_gs;
    for (int i = tid; i < 64 * a.hd; i += 256) {
        const int r = i / a.hd, c = i % a.hd;
        if (row0 + r < a.NQ) dQg[(long)(row0 + r) * a.ld_dq + c] = stage[r * HDP + c] * sL[r];
    }
}

template <int HDP>
__global__ void __launch_bounds__(256, HDP <= 128 ? 2 : 1) attn_bwd_dkv_kernel(const AttnBwd2Args a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ uint8_t smraw[];
    uint8_t *sm = align_smem(smraw, 128);
    constexpr int LDP = HDP + 8;
    bf16 *Qs = (bf16 *)sm, *dOs = Qs + 64 * LDP, *Ks = dOs + 64 * LDP, *Vs = Ks + 64 * LDP;
    float *Sf = (float *)(Vs + 64 * LDP), *Df = Sf + 64 * SLD;
    bf16 *Pb = (bf16 *)(Df + 64 * SLD), *dSb = Pb + 64 * SLD;
    float *sD = (float *)(dSb + 64 * SLD), *sL = sD + 64;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int NS = a.qsplit > 1 ? a.qsplit : 1;
    const int b = blockIdx.z, g = blockIdx.y, key0 = (blockIdx.x / NS) * 64, split = blockIdx.x % NS;
    const int vl = a.valid_len ? a.valid_len[b] : (1 << 30);
    const int tok_limit = (a.valid_len && a.seg == 0) ? vl : (1 << 30);
    const bf16 *Qg = a.Q + b * a.q_bs + g * a.q_gs, *dOg = a.dO + b * a.o_bs + g * a.o_gs;
    const bf16 *Kg = a.K + b * a.kv_bs + g * a.kv_gs, *Vg = a.V + b * a.kv_bs + g * a.kv_gs;
    load_rows64x2<HDP>(Ks, Kg, a.ld_kv, Vs, Vg, a.ld_kv, key0, a.NK, a.hd, 1, 1 << 30);
    constexpr int NCT = HDP / 16, HALF = (NCT + 1) / 2;
    const int kr = warp & 3, wc = warp >> 2;          // 16 keys x one half of the feature tiles per warp
    const int lane_ = tid & 31;
    float acck[HALF][2][4], accv[HALF][2][4];
#pragma unroll
    for (int t = 0; t < HALF; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) acck[t][0][i] = acck[t][1][i] = accv[t][0][i] = accv[t][1][i] = 0.f;
    const float inv_cap = a.softcap > 0.f ? 1.f / a.softcap : 0.f;
    const int er = tid >> 2, ec0 = (tid & 3) * 16;
    const int n_qt = (a.NQ + 63) / 64;
    // keys of this block that any query of this segment sees: skip the block if none
    bool any = false;
    for (int j = key0; j < key0 + 64; ++j) any = any || key_visible(a, j, vl);
    if (split >= n_qt) return;   // no query tile for this CTA: nothing to add
    for (int qt = split; any && qt < n_qt; qt += NS) {
        const int row0 = qt * 64;
        if (a.valid_len && a.seg == 0 && (long)row0 >= (long)vl * a.hpr) break;      // padded tokens only from here on
        __syncthreads();
        load_rows64x2<HDP>(Qs, Qg, a.ld_q, dOs, dOg, a.ld_o, row0, a.NQ, a.hd, a.hpr, tok_limit);
        if (tid < 64) {
            const long o = ((long)b * a.groups + g) * a.NQ + row0 + tid;
            const bool ok = row0 + tid < a.NQ;
            sL[tid] = ok ? a.lse[o] : INFINITY;
            sD[tid] = ok ? a.dvec[o] : 0.f;
        }
        __syncthreads();
        mma_qkT<HDP>(Qs, Ks, Sf, warp);
        mma_qkT<HDP>(dOs, Vs, Df, warp);
        __syncthreads();
        const float Dr = sD[er], Lr = sL[er];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float sv = Sf[er * SLD + ec0 + c] * a.scale;
            const float l = a.softcap > 0.f ? a.softcap * tanh_hw(sv * inv_cap) : sv;
            const float th = l * inv_cap;
            const float pv = key_visible(a, key0 + ec0 + c, vl) ? __expf(l - Lr) : 0.f;
            Pb[er * SLD + ec0 + c] = __float2bfloat16_rn(pv);
            dSb[er * SLD + ec0 + c] = __float2bfloat16_rn(pv * (Df[er * SLD + ec0 + c] - Dr) * (1.f - th * th) * a.scale);
        }
        __syncthreads();
#pragma unroll
        for (int q0 = 0; q0 < 4; ++q0) {
            uint32_t fp[4], fs[4];   // P^T, dS^T: element (key, query) lives at [query][key]
            frag_a_t(fp, Pb + q0 * 16 * SLD + kr * 16, SLD, lane_);
            frag_a_t(fs, dSb + q0 * 16 * SLD + kr * 16, SLD, lane_);
#pragma unroll
            for (int t = 0; t < HALF; ++t) {
                const int ct = wc * HALF + t;
                if (ct < NCT) {
                    uint32_t fo[4], fq[4];
                    frag_b_kn(fo, dOs + q0 * 16 * LDP + ct * 16, LDP, lane_);
                    frag_b_kn(fq, Qs + q0 * 16 * LDP + ct * 16, LDP, lane_);
                    mma16816(accv[t][0], fp, fo[0], fo[1]);
                    mma16816(accv[t][1], fp, fo[2], fo[3]);
                    mma16816(acck[t][0], fs, fq[0], fq[1]);
                    mma16816(acck[t][1], fs, fq[2], fq[3]);
                }
            }
        }
    }
    float *stage = (float *)Qs;    // 64 x HDP fp32 = the Q and dO tiles
    for (int which = 0; which < 2; ++which) {
        __syncthreads();
#pragma unroll
        for (int t = 0; t < HALF; ++t) {
            const int ct = wc * HALF + t;
            if (ct < NCT) {
                if (which) store_c16(stage + kr * 16 * HDP + ct * 16, HDP, acck[t][0], acck[t][1], lane_);
                else store_c16(stage + kr * 16 * HDP + ct * 16, HDP, accv[t][0], accv[t][1], lane_);
            }
        }
        __syncthreads();
        float *dst = (which ? a.dK : a.dV) + b * a.dkv_bs + g * a.dkv_gs;
        if (NS > 1) {
            // partial sums of this CTA's share of the query tiles: 16-byte fp32 atomics (hd and the row pitch are multiples of 4)
            for (int i = tid; any && i < 64 * (a.hd / 4); i += 256) {
                const int r = i / (a.hd / 4), c = (i % (a.hd / 4)) * 4;
                if (key0 + r < a.NK) {
                    const float4 v = *reinterpret_cast<const float4 *>(stage + r * HDP + c);
                    atomicAdd(reinterpret_cast<float4 *>(dst + (long)(key0 + r) * a.ld_dkv + c), v);
                }
            }
        } else {
            for (int i = tid; i < 64 * a.hd; i += 256) {
                const int r = i / a.hd, c = i % a.hd;
                if (key0 + r < a.NK) {
                    float *p = dst + (long)(key0 + r) * a.ld_dkv + c;
                    *p = (a.accumulate ? *p : 0.f) + stage[r * HDP + c];
                }
            }
        }
    }
}

template <int HDP>
int attn_bwd2_launch(const AttnBwd2Args &a, int batch, bool dq, bool dkv, cudaStream_t st) {
    const size_t smem = (size_t)4 * 64 * (HDP + 8) * 2 + 2 * 64 * SLD * 4 + 2 * 64 * SLD * 2 + 2 * 64 * 4 + 128;
    static PerDeviceOnce once;
    if (once.need()) {
        cudaFuncSetAttribute(attn_bwd_dq_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(attn_bwd_dkv_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    if (dq) launch_k(attn_bwd_dq_kernel<HDP>, dim3((a.NQ + 63) / 64, a.groups, batch), dim3(256), smem, st, a);
    if (dkv) {
        AttnBwd2Args k = a;
        const int n_qt = (a.NQ + 63) / 64;
        if (k.qsplit > n_qt) k.qsplit = n_qt;      // (a CTA without a query tile would only add zeros; qsplit stays >= 2 when it was:
        if (a.qsplit > 1 && k.qsplit < 2) k.qsplit = 2;   //  the accumulation into dK / dV must remain atomic)
        launch_k(attn_bwd_dkv_kernel<HDP>, dim3(((a.NK + 63) / 64) * (k.qsplit > 1 ? k.qsplit : 1), a.groups, batch), dim3(256), smem, st, k);
    }
    return 0;
}

// ------------------------------------------------------------ optimizer ----
// sum of squares of a flat fp32 gradient buffer (clip_grad_norm_, train.py:371).  Deterministic: every CTA writes its
// partial, a second one-CTA pass adds them in a fixed order -- data-parallel replicas must compute bit-identical clip
// factors from their bit-identical all-reduced gradients, or their weights drift apart.
__global__ void __launch_bounds__(256) sumsq_kernel(const float *__restrict__ x, size_t n, float *__restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    float acc = 0.f;
    const size_t n4 = n / 4;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = x4[i];
        acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += x[i] * x[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float *__restrict__ partial, int n, float *__restrict__ out) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) *out = acc;
}

// AdamW (torch.optim.AdamW semantics: decoupled weight decay, bias-corrected moments) on fp32 master weights laid out like
// the gradient buffer, with the global-norm clip folded in and the updated value written straight into the PACKED weight
// tensor the kernels read (model dtype).  A CTA's 256 consecutive elements lie inside one entry (entries start on
// multiples of 256), found by bisection over the entry offsets.
struct AdamWArgs {
    float *master, *m, *v, *grad;
    size_t begin, end;                 // element range of this parameter group
    const long long *entry_off;        // [n_entries + 1]
    void *const *entry_dst;            // [n_entries] packed weight tensors
    const long long *entry_n;          // [n_entries] elements of each
    int n_entries, dst_bf16;
    float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;
    const float *sumsq;                // device scalar: sum of squares of the (scaled) gradients, or null = no clipping
    float max_norm, grad_scale;
    int zero_grad;
    int copy_only;                     // no update: just write `master` (any flat buffer of the layout) into the packed tensors
};
__global__ void __launch_bounds__(256) adamw_kernel(const AdamWArgs a) {
    pdl_trigger();
    pdl_wait();
    float gs = a.grad_scale;
    if (a.sumsq) {
        const float norm = sqrtf(*a.sumsq) * a.grad_scale;
        gs *= fminf(1.f, a.max_norm / (norm + 1e-6f));
    }
    const float decay = 1.f - a.lr * a.wd, step_size = a.lr / a.bc1, omb1 = 1.f - a.beta1, omb2 = 1.f - a.beta2;
    // a warp owns 128 consecutive elements per iteration (4 per lane): entries start on multiples of 256, so a warp never
    // straddles two of them; the entry index only moves forward as the warp walks the buffer
    const int lane = threadIdx.x & 31;
    const size_t warp_id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    int e = -1;
    for (size_t base = a.begin + warp_id * 128; base < a.end; base += n_warps * 128) {
        if (e < 0) {
            int lo = 0, hi = a.n_entries;      // last entry with off <= base
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if ((size_t)__ldg(a.entry_off + mid) <= base) lo = mid; else hi = mid;
            }
            e = lo;
        } else {
            while (e + 1 < a.n_entries && (size_t)__ldg(a.entry_off + e + 1) <= base) ++e;
        }
        const size_t i = base + 4 * lane;
        const size_t local = i - (size_t)__ldg(a.entry_off + e);
        const size_t n_e = (size_t)__ldg(a.entry_n + e);
        if (i >= a.end || local >= n_e) continue;
        const float4 p4 = *reinterpret_cast<const float4 *>(a.master + i);
        if (a.copy_only) {
            const int nk_ = (int)(n_e - local < 4 ? n_e - local : 4);
            const float pc[4] = {p4.x, p4.y, p4.z, p4.w};
            if (a.dst_bf16) {
                bf16 *dst = (bf16 *)__ldg(reinterpret_cast<const unsigned long long *>(a.entry_dst) + e) + local;
                for (int k = 0; k < nk_; ++k) dst[k] = __float2bfloat16_rn(pc[k]);
            } else {
                float *dst = (float *)__ldg(reinterpret_cast<const unsigned long long *>(a.entry_dst) + e) + local;
                for (int k = 0; k < nk_; ++k) dst[k] = pc[k];
            }
            continue;
        }
        const float4 g4 = *reinterpret_cast<const float4 *>(a.grad + i);
        const float4 m4 = *reinterpret_cast<const float4 *>(a.m + i);
        const float4 v4 = *reinterpret_cast<const float4 *>(a.v + i);
        float g[4] = {g4.x * gs, g4.y * gs, g4.z * gs, g4.w * gs}, p[4] = {p4.x, p4.y, p4.z, p4.w};
        float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            m[k] = a.beta1 * m[k] + omb1 * g[k];
            v[k] = a.beta2 * v[k] + omb2 * g[k] * g[k];
            p[k] = p[k] * decay - step_size * m[k] / (sqrtf(v[k]) / a.bc2_sqrt + a.eps);
        }
        // (the 0 .. 3 pad elements behind an entry's end are updated too: their gradient is 0, they stay 0)
        *reinterpret_cast<float4 *>(a.master + i) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4 *>(a.m + i) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4 *>(a.v + i) = make_float4(v[0], v[1], v[2], v[3]);
        if (a.zero_grad) *reinterpret_cast<float4 *>(a.grad + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        const int nk = (int)(n_e - local < 4 ? n_e - local : 4);
        if (a.dst_bf16) {
            bf16 *dst = (bf16 *)__ldg(reinterpret_cast<const unsigned long long *>(a.entry_dst) + e) + local;
            if (nk == 4 && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
                *reinterpret_cast<uint2 *>(dst) = make_uint2(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]));
            } else {
                for (int k = 0; k < nk; ++k) dst[k] = __float2bfloat16_rn(p[k]);
            }
        } else {
            float *dst = (float *)__ldg(reinterpret_cast<const unsigned long long *>(a.entry_dst) + e) + local;
            for (int k = 0; k < nk; ++k) dst[k] = p[k];
        }
    }
}


// avg += (x - avg) * weight over a flat fp32 buffer: EMA (weight = 1 - decay) and SWA (weight = 1 / (n_averaged + 1)) of the
// master weights (model_averaging.py:40-66, torch.optim.swa_utils)
__global__ void __launch_bounds__(256) lerp_kernel(float *__restrict__ avg, const float *__restrict__ x, size_t n, float weight) {
    pdl_trigger();
    pdl_wait();
    const size_t n4 = n / 4;
    float4 *a4 = reinterpret_cast<float4 *>(avg);
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = a4[i];
        const float4 v = x4[i];
        a.x += (v.x - a.x) * weight; a.y += (v.y - a.y) * weight; a.z += (v.z - a.z) * weight; a.w += (v.w - a.w) * weight;
        a4[i] = a;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        avg[i] += (x[i] - avg[i]) * weight;
}

// ========================================================= host helpers ====
// one linear layer of the training step: the tensor-core / skinny kernels where the shape has one, otherwise the SIMT
// kernel (the 7- and 8-wide matrices of the action heads and the fp32 build)
template <typename T>
int tlin(pz_handle *h, const LinearArgs &a, cudaStream_t st) {
    if (std::is_same<T, bf16>::value && !(h->cfg.flags & PZ_FLAG_SIMPLE_KERNELS)) {
        if (skinny_supported(a)) return launch_linear_skinny(a, st);
        if (gemm_tc_supported(a)) {
            const char *e = nullptr;
            int rc = launch_linear_tc(a, st, &e);
            if (rc) return fail(h, rc, e ? e : "tcgen05 gemm launch failed");
            return 0;
        }
    }
    launch_linear_simple<T>(a, st);
    return 0;
}

// Backward of y = x W^T (bias handled by the caller through colsum), operands read in place:
//   dX [M,K] (fp32 `dx_flags` = LIN_OUT_F32 [| LIN_ACCUM], or T with 0) = dY . W    -- W [N][K] is the MN-major B operand
//   dW [N,K] fp32 += dY^T . X    -- dY [M][N] and X [M][K] are the MN-major A / B operands, tokens the contraction dimension
// (the first version went through tiled transposes: 14.5 ms per step).  dY: T [M,N] (ld ldy); X: T [M,K] (ld ldx); W: T [N,K].
template <typename T>
int linear_bwd(pz_handle *h, const void *dY, int ldy, const void *X, int ldx, const void *W, void *dX, int ld_dx,
               int dx_flags, float *dW, int M, int N, int K, cudaStream_t st) {
    if (dX) {
        LinearArgs a = lin(dY, ldy, W, nullptr, dX, ld_dx, M, K, N, dx_flags | LIN_W_MN);
        a.ldw = K;
        PZ_TRY(tlin<T>(h, a, st));
    }
    if (dW) {
        PdlOff no_pdl;   // the "weight" operand X was written by the kernel right in front (see PdlOff)
        LinearArgs a = lin(dY, ldy, X, nullptr, dW, K, N, K, M, LIN_OUT_F32 | LIN_ACCUM | LIN_A_MN | LIN_W_MN);
        a.ldw = ldx;
        PZ_TRY(tlin<T>(h, a, st));
    }
    return 0;
}

// ------------------------------------------------------------ workspace ----
struct MixDims { int rows, hidden, inter; };   // rows per sample

struct TrainWs {
    // saved by the forward
    void *patches;                              // T [Mv, kp]
    std::vector<float *> xv;                    // [LV + 1] fp32 [Mv, V]: SigLIP residual stream entering layer i
    std::vector<float *> xv_mid;                // [LV] after attention
    std::vector<void *> qkvv, av, f1, actv;     // [LV] T
    void *hv_post;                              // T [Mv, V] post layernorm output
    float *feats;                               // fp32 [Mv, H]
    std::vector<float *> xin[3];                // [L + 1] fp32 [M_m, H_m]
    std::vector<float *> x1[3];                 // [L]
    std::vector<void *> q[3], att[3], gu[3], mm[3];   // [L] T
    std::vector<void *> K, V;                   // [L] T [B][S_all][hd]
    void *a_in, *e1, *temb, *z, *pp, *hfin;     // T
    float *tbias, *zpre, *vel, *psi;
    // scratch
    void *h, *qkv, *dyb, *d_m, *dgu, *dqkv;
    void *datt[3];
    float *dx[3], *dq[3], *dh, *dK, *dV, *lse, *dvec;
    float *att_scratch; size_t att_scratch_bytes;
    size_t total;
};

TrainWs carve_train(const pz_config &c, int B, void *base) {
    TrainWs w;
    Bump b;
    b.base = (char *)base;
    const size_t es = c.dtype == PZ_BF16 ? 2 : 4;
    const int L = c.n_layers, LV = c.vit_layers;
    const int S_all = c.s_vlm + c.cond_steps + c.horizon;
    const size_t qd = (size_t)c.n_heads * c.head_dim, qkvd = (size_t)(c.n_heads + 2 * c.n_kv_heads) * c.head_dim;
    const size_t Mv = (size_t)B * c.n_images * c.n_img_tokens;
    const int V = c.vit_hidden, VI = c.vit_inter, H = c.vlm_hidden, A = c.act_hidden;
    const MixDims md[3] = {{c.s_vlm, H, c.vlm_inter}, {c.cond_steps, A, c.act_inter}, {c.horizon, A, c.act_inter}};
    w.patches = b.take<void>(Mv * c.patch_k_pad * es);
    w.xv.resize(LV + 1); w.xv_mid.resize(LV); w.qkvv.resize(LV); w.av.resize(LV); w.f1.resize(LV); w.actv.resize(LV);
    for (int i = 0; i <= LV; ++i) w.xv[i] = b.take<float>(Mv * V * 4);
    for (int i = 0; i < LV; ++i) {
        w.xv_mid[i] = b.take<float>(Mv * V * 4);
        w.qkvv[i] = b.take<void>(Mv * 3 * V * es);
        w.av[i] = b.take<void>(Mv * V * es);
        w.f1[i] = b.take<void>(Mv * VI * es);
        w.actv[i] = b.take<void>(Mv * VI * es);
    }
    w.hv_post = b.take<void>(Mv * V * es);
    w.feats = b.take<float>(Mv * H * 4);
    size_t Mmax = 0, MHmax = 0, MImax = 0;
    for (int m = 0; m < 3; ++m) {
        const size_t M = (size_t)B * md[m].rows;
        Mmax = M > Mmax ? M : Mmax;
        MHmax = M * md[m].hidden > MHmax ? M * md[m].hidden : MHmax;
        MImax = M * md[m].inter > MImax ? M * md[m].inter : MImax;
        w.xin[m].resize(L + 1); w.x1[m].resize(L); w.q[m].resize(L); w.att[m].resize(L); w.gu[m].resize(L); w.mm[m].resize(L);
        for (int l = 0; l <= L; ++l) w.xin[m][l] = b.take<float>(M * md[m].hidden * 4);
        for (int l = 0; l < L; ++l) {
            w.x1[m][l] = b.take<float>(M * md[m].hidden * 4);
            w.q[m][l] = b.take<void>(M * qd * es);
            w.att[m][l] = b.take<void>(M * qd * es);
            w.gu[m][l] = b.take<void>(M * 2 * md[m].inter * es);
            w.mm[m][l] = b.take<void>(M * md[m].inter * es);
        }
        w.datt[m] = b.take<void>(M * qd * es);
        w.dx[m] = b.take<float>(M * md[m].hidden * 4);
        w.dq[m] = b.take<float>(M * qd * 4);
    }
    w.K.resize(L); w.V.resize(L);
    for (int l = 0; l < L; ++l) {
        w.K[l] = b.take<void>((size_t)B * S_all * c.head_dim * es);
        w.V[l] = b.take<void>((size_t)B * S_all * c.head_dim * es);
    }
    const size_t Ma = (size_t)B * c.horizon, Mp = (size_t)B * c.cond_steps;
    w.a_in = b.take<void>(Ma * 64 * es);
    w.e1 = b.take<void>(Ma * A * es);
    w.temb = b.take<void>((size_t)rup(B, 8) * A * es);
    w.z = b.take<void>(Ma * A * es);
    w.pp = b.take<void>(Mp * 64 * es);
    w.hfin = b.take<void>(Ma * A * es);
    w.tbias = b.take<float>((size_t)B * A * 4);
    w.zpre = b.take<float>(Ma * A * 4);
    w.vel = b.take<float>(Ma * 8 * 4);
    w.psi = b.take<float>(Ma * c.action_dim * 4);
    // scratch
    size_t hmax = MHmax > Mv * V ? MHmax : Mv * V;
    w.h = b.take<void>(hmax * es);
    w.dyb = b.take<void>((hmax > Mv * (size_t)H ? hmax : Mv * (size_t)H) * es);
    w.dh = b.take<float>((hmax > Mv * (size_t)3 * V ? hmax : Mv * (size_t)3 * V) * 4);
    w.qkv = b.take<void>((Mmax * qkvd > Mv * 3 * V ? Mmax * qkvd : Mv * 3 * V) * es);
    w.dqkv = b.take<void>((Mmax * qkvd > Mv * 3 * V ? Mmax * qkvd : Mv * 3 * V) * es);
    size_t imax = MImax > Mv * VI ? MImax : Mv * VI;
    w.d_m = b.take<void>(imax * es);
    w.dgu = b.take<void>(2 * imax * es);
    w.dK = b.take<float>((size_t)B * S_all * c.head_dim * 4);
    w.dV = b.take<float>((size_t)B * S_all * c.head_dim * 4);
    {
        size_t nj = (size_t)B * (c.s_vlm + c.cond_steps + c.horizon) * c.n_heads, nv = Mv * c.vit_heads;
        w.lse = b.take<float>((nj > nv ? nj : nv) * 4);
        w.dvec = b.take<float>((nj > nv ? nj : nv) * 4);
    }
    {
        size_t rows = (size_t)c.n_heads * (c.horizon > c.cond_steps ? c.horizon : c.cond_steps);
        size_t tiles = (S_all + 63) / 64;
        w.att_scratch_bytes = rows <= 64 ? (size_t)B * tiles * rows * (c.head_dim + 2) * 4 : 0;
        w.att_scratch = b.take<float>(w.att_scratch_bytes);
    }
    w.total = (b.off + 1023) & ~(size_t)1023;
    return w;
}

// the gradient buffers: the same structs as the weights, every non-null pointer an fp32 buffer of the packed shape
struct Grads {
    pz_weights g;
    std::vector<pz_vit_layer> vit;
    std::vector<pz_mix_layer> mix[3];
};
inline float *G(const void *p) { return (float *)const_cast<void *>(p); }

template <typename T>
int forward_backward(pz_handle *h, const int64_t *ids, const void *pixels, const int32_t *valid_len, const float *proprio,
                     const float *actions, const float *noise, const float *t, float sig_min, const Grads *gr, float *loss,
                     float loss_scale, void *wsp, int B, int flags, void *const *events, int n_events, cudaStream_t st) {
    const pz_config &c = h->cfg;
    const pz_weights &w = h->w;
    if (!w.enc_w2t || !w.enc_b2 || !w.time_freq) return fail(h, PZ_ERR_UNBOUND, "enc_w2t / enc_b2 / time_freq not bound");
    TrainWs ws = carve_train(c, B, wsp);
    const int L = c.n_layers, LV = c.vit_layers;
    const int H = c.vlm_hidden, A = c.act_hidden, hd = c.head_dim, nh = c.n_heads;
    const int S_v = c.s_vlm, S_p = c.cond_steps, Hz = c.horizon, S_c = S_v + S_p, S_all = S_c + Hz;
    const int qd = nh * hd, qkvd = (nh + 2 * c.n_kv_heads) * hd;
    const long kv_bs = (long)S_all * hd;
    const int V = c.vit_hidden, VI = c.vit_inter, P = c.n_img_tokens, hdv = V / c.vit_heads;
    const int n_img = B * c.n_images, Mv = n_img * P;
    const MixDims md[3] = {{S_v, H, c.vlm_inter}, {S_p, A, c.act_inter}, {Hz, A, c.act_inter}};
    const pz_mix_layer *mixw[3] = {h->vlm.data(), h->proprio.data(), h->action.data()};
    const int row_off[3] = {0, S_v, S_c};
    const float *rc_[3] = {w.rope_vlm_cos, w.rope_act_cos, w.rope_act_cos};
    const float *rs_[3] = {w.rope_vlm_sin, w.rope_act_sin, w.rope_act_sin};
    const int pos0[3] = {0, 0, S_p};
    const int Ma = B * Hz, Mp = B * S_p;
    const int skp = w.small_k_pad;
    const bool want_grads = gr != nullptr;
    const bool vit_grads = want_grads && !(flags & PZ_TRAIN_FREEZE_VISION);
    const char *err = nullptr;
    // "these gradients are final" marks for the caller's bucketed all-reduce: [0, L) joint layer l, L the encoder / decoder
    // heads, L + 1 + i SigLIP layer i, L + 1 + LV everything else (projector, patch embedding, position table)
    auto mark = [&](int idx) { if (events && idx < n_events && events[idx]) cudaEventRecord((cudaEvent_t)events[idx], st); };
    // bf16: attention backward on the tensor cores (PZ_ATTN_BWD_SIMT=1 keeps the SIMT kernel, the fp32 build's only one)
    static const bool simt_attn = [] { const char *e = getenv("PZ_ATTN_BWD_SIMT"); return e && e[0] == '1'; }();
    const bool tc_attn_bwd = std::is_same<T, bf16>::value && !(c.flags & PZ_FLAG_SIMPLE_KERNELS) && !simt_attn;

    // ================================================================ forward
    // ---- SigLIP (siglip.py:59-78, 220-238, 298) with the residual stream and the GEMM inputs of every layer kept
    if (h->pixel_format == PZ_PIXELS_U8) launch_im2col_u8<T>((const uint8_t *)pixels, (T *)ws.patches, n_img, c.image_size, c.patch_size, c.patch_k_pad, st);
    else launch_im2col<T>((const T *)pixels, (T *)ws.patches, n_img, c.image_size, c.patch_size, c.patch_k_pad, st);
    launch_bcast_rows(ws.xv[0], w.pos_emb, Mv, V, P, st);
    PZ_TRY(tlin<T>(h, lin(ws.patches, c.patch_k_pad, w.patch_w, w.patch_b, ws.xv[0], V, Mv, V, c.patch_k_pad, LIN_OUT_F32 | LIN_ACCUM), st));
    for (int i = 0; i < LV; ++i) {
        const pz_vit_layer &Lw = h->vit[i];
        launch_layernorm<T>(ws.xv[i], Lw.ln1_w, Lw.ln1_b, (T *)ws.h, Mv, V, 1e-6f, st);
        PZ_TRY(tlin<T>(h, lin(ws.h, V, Lw.w_qkv, Lw.b_qkv, ws.qkvv[i], 3 * V, Mv, 3 * V, V), st));
        AttnArgs a;
        memset(&a, 0, sizeof(a));
        a.Q = ws.qkvv[i]; a.q_batch_stride = (long)P * 3 * V; a.q_row_stride = 3 * V; a.q_head_stride = hdv;
        a.K = (const T *)ws.qkvv[i] + V; a.V = (const T *)ws.qkvv[i] + 2 * V;
        a.kv_batch_stride = (long)P * 3 * V; a.kv_row_stride = 3 * V; a.kv_head_stride = hdv;
        a.O = ws.av[i]; a.o_batch_stride = (long)P * V; a.o_row_stride = V; a.o_head_stride = hdv;
        a.batch = n_img; a.n_heads = c.vit_heads; a.head_dim = hdv; a.q_rows = P; a.q_row0 = 0;
        a.s_cache = P; a.s_vlm = P; a.n_fresh = 0;
        a.scale = 1.0f / sqrtf((float)hdv); a.softcap = 0.f;
        PZ_TRY(Ops<T>::attention(h, a, st));
        copy_f32(ws.xv_mid[i], ws.xv[i], (size_t)Mv * V, st);
        PZ_TRY(tlin<T>(h, lin(ws.av[i], V, Lw.w_o, Lw.b_o, ws.xv_mid[i], V, Mv, V, V, LIN_OUT_F32 | LIN_ACCUM), st));
        launch_layernorm<T>(ws.xv_mid[i], Lw.ln2_w, Lw.ln2_b, (T *)ws.h, Mv, V, 1e-6f, st);
        PZ_TRY(tlin<T>(h, lin(ws.h, V, Lw.w_fc1, Lw.b_fc1, ws.f1[i], VI, Mv, VI, V), st));
        launch_k(gelu_fwd_kernel<T>, dim3(ew_blocks((long)Mv * VI)), dim3(256), 0, st, (const T *)ws.f1[i], (T *)ws.actv[i], (long)Mv * VI);
        copy_f32(ws.xv[i + 1], ws.xv_mid[i], (size_t)Mv * V, st);
        PZ_TRY(tlin<T>(h, lin(ws.actv[i], VI, Lw.w_fc2, Lw.b_fc2, ws.xv[i + 1], V, Mv, V, VI, LIN_OUT_F32 | LIN_ACCUM), st));
    }
    launch_layernorm<T>(ws.xv[LV], w.post_ln_w, w.post_ln_b, (T *)ws.hv_post, Mv, V, 1e-6f, st);
    PZ_TRY(tlin<T>(h, lin(ws.hv_post, V, w.proj_w, w.proj_b, ws.feats, H, Mv, H, V, LIN_OUT_F32), st));
    launch_embed_merge<T>(ids, (const T *)w.embed, ws.feats, ws.xin[0][0], B, S_v, H, c.n_images * P, c.image_token_index,
                          c.pad_token_id, sqrtf((float)H), st);
    // ---- proprio encoder, time embedding, action encoder (pizero.py:597-639, vla/modules.py:9-53)
    launch_cast_pad<T>(proprio, (T *)ws.pp, Mp, c.proprio_dim, skp, st);
    PZ_TRY(tlin<T>(h, lin(ws.pp, skp, w.prop_w, w.prop_b, ws.xin[1][0], A, Mp, A, skp, LIN_OUT_F32, sqrtf((float)A)), st));
    launch_psi(noise, actions, t, ws.psi, B, Hz * c.action_dim, sig_min, st);
    cudaMemsetAsync(ws.temb, 0, (size_t)rup(B, 8) * A * sizeof(T), st);
    launch_time_embed<T>(t, w.time_freq, (T *)ws.temb, B, A / 2, st);
    PZ_TRY(tlin<T>(h, lin(ws.temb, A, w.enc_w2t, w.enc_b2, ws.tbias, A, B, A, A, LIN_OUT_F32), st));
    launch_cast_pad<T>(ws.psi, (T *)ws.a_in, Ma, c.action_dim, skp, st);
    PZ_TRY(tlin<T>(h, lin(ws.a_in, skp, w.enc_w1, w.enc_b1, ws.e1, A, Ma, A, skp), st));
    PZ_TRY(tlin<T>(h, lin(ws.e1, A, w.enc_w2a, nullptr, ws.zpre, A, Ma, A, A, LIN_OUT_F32), st));
    launch_rowbias_silu<T>(ws.zpre, ws.tbias, (T *)ws.z, Ma, A, Hz, st);
    PZ_TRY(tlin<T>(h, lin(ws.z, A, w.enc_w3, w.enc_b3, ws.xin[2][0], A, Ma, A, A, LIN_OUT_F32, sqrtf((float)A)), st));
    // ---- joint model, all three mixtures active, no cache (joint_model.py:24-127, 243-305)
    for (int l = 0; l < L; ++l) {
        const bool last = l == L - 1;
        T *Kl = (T *)ws.K[l], *Vl = (T *)ws.V[l];
        for (int m = 0; m < 3; ++m) {
            const int M = B * md[m].rows, Hm = md[m].hidden;
            const pz_mix_layer &Lw = mixw[m][l];
            launch_rmsnorm<T>(ws.xin[m][l], Lw.norm_in, (T *)ws.h, M, Hm, 1e-6f, st);
            PZ_TRY(tlin<T>(h, lin(ws.h, Hm, Lw.w_qkv, nullptr, ws.qkv, qkvd, M, qkvd, Hm), st));
            launch_rope_split<T>((const T *)ws.qkv, qkvd, (T *)ws.q[m][l], (long)md[m].rows * qd, Kl + (size_t)row_off[m] * hd,
                                 Vl + (size_t)row_off[m] * hd, kv_bs, rc_[m], rs_[m], B, md[m].rows, pos0[m], nh, hd, st);
        }
        for (int m = 0; m < 3; ++m) {
            if (last && m < 2) continue;   // joint_model.py:297-299: the vlm / proprio halves of the last layer are dropped
            AttnArgs a;
            memset(&a, 0, sizeof(a));
            a.K = Kl; a.V = Vl; a.kv_batch_stride = kv_bs; a.kv_row_stride = hd; a.kv_head_stride = 0;
            a.valid_len = valid_len;
            a.batch = B; a.n_heads = nh; a.head_dim = hd; a.s_cache = S_c; a.s_vlm = S_v;
            a.scale = 1.0f / sqrtf((float)hd); a.softcap = 50.f;
            a.q_row_stride = qd; a.q_head_stride = hd; a.o_row_stride = qd; a.o_head_stride = hd;
            a.Q = ws.q[m][l]; a.q_batch_stride = (long)md[m].rows * qd;
            a.O = ws.att[m][l]; a.o_batch_stride = (long)md[m].rows * qd;
            a.q_rows = md[m].rows; a.q_row0 = row_off[m];
            a.scratch = ws.att_scratch; a.scratch_bytes = ws.att_scratch_bytes;
            // rows >= valid_len are skipped by the attention kernels: give them a defined (zero) output -- the backward
            // multiplies every saved row by a gradient row, and 0 * uninitialised is not 0
            if (m == 0) cudaMemsetAsync(ws.att[m][l], 0, (size_t)B * md[m].rows * qd * sizeof(T), st);
            if (m == 2) {
                a.n_fresh = Hz;
                a.K2 = Kl + (size_t)S_c * hd; a.V2 = Vl + (size_t)S_c * hd; a.kv2_batch_stride = kv_bs; a.kv2_row_stride = hd;
            }
            PZ_TRY(Ops<T>::attention(h, a, st));
        }
        for (int m = 0; m < 3; ++m) {
            if (last && m < 2) continue;
            const int M = B * md[m].rows, Hm = md[m].hidden, Im = md[m].inter;
            const pz_mix_layer &Lw = mixw[m][l];
            copy_f32(ws.x1[m][l], ws.xin[m][l], (size_t)M * Hm, st);
            PZ_TRY(tlin<T>(h, lin(ws.att[m][l], qd, Lw.w_o, nullptr, ws.x1[m][l], Hm, M, Hm, qd, LIN_OUT_F32 | LIN_ACCUM), st));
            launch_rmsnorm<T>(ws.x1[m][l], Lw.norm_post, (T *)ws.h, M, Hm, 1e-6f, st);
            PZ_TRY(tlin<T>(h, lin(ws.h, Hm, Lw.w_gate_up, nullptr, ws.gu[m][l], 2 * Im, M, 2 * Im, Hm), st));
            launch_k(geglu_fwd_kernel<T>, dim3(ew_blocks((long)M * Im / 8)), dim3(256), 0, st, (const T *)ws.gu[m][l], (T *)ws.mm[m][l],
                     (long)M * Im, Im);
            copy_f32(ws.xin[m][l + 1], ws.x1[m][l], (size_t)M * Hm, st);
            PZ_TRY(tlin<T>(h, lin(ws.mm[m][l], Im, Lw.w_down, nullptr, ws.xin[m][l + 1], Hm, M, Hm, Im, LIN_OUT_F32 | LIN_ACCUM), st));
        }
    }
    // ---- final norm, action decoder, loss (joint_model.py:375-380, pizero.py:657-661)
    launch_rmsnorm<T>(ws.xin[2][L], w.action_final_norm, (T *)ws.hfin, Ma, A, 1e-6f, st);
    PZ_TRY(tlin<T>(h, lin(ws.hfin, A, w.dec_w, w.dec_b, ws.vel, 8, Ma, c.action_dim, A, LIN_OUT_F32), st));
    launch_fm_loss(ws.vel, 8, noise, actions, loss, nullptr, Ma, c.action_dim, sig_min, st);
    if (!want_grads) return 0;

    // =============================================================== backward
    const pz_weights &g = gr->g;
    const pz_mix_layer *mixg[3] = {gr->mix[0].data(), gr->mix[1].data(), gr->mix[2].data()};
    // ---- loss -> action decoder -> final norm
    T *dv = (T *)ws.dyb;
    launch_k(loss_bwd_kernel<T>, dim3((Ma * 8 + 255) / 256), dim3(256), 0, st, (const float *)ws.vel, 8, noise, actions, dv, (long)Ma,
             c.action_dim, sig_min, loss_scale);
    colsum<T>(dv, 8, G(g.dec_b), Ma, c.action_dim, st);
    PZ_TRY(linear_bwd<T>(h, dv, 8, ws.hfin, A, w.dec_w, ws.dh, A, LIN_OUT_F32, G(g.dec_w), Ma, 8, A, st));
    for (int m = 0; m < 3; ++m) cudaMemsetAsync(ws.dx[m], 0, (size_t)B * md[m].rows * md[m].hidden * 4, st);
    rmsnorm_bwd<T>(ws.xin[2][L], w.action_final_norm, ws.dh, ws.dx[2], G(g.action_final_norm), Ma, A, st);
    // ---- layers, last to first
    for (int l = L - 1; l >= 0; --l) {
        const bool last = l == L - 1;
        for (int m = 0; m < 3; ++m) {
            if (last && m < 2) continue;
            const int M = B * md[m].rows, Hm = md[m].hidden, Im = md[m].inter;
            const pz_mix_layer &Lw = mixw[m][l];
            const pz_mix_layer &Lg = mixg[m][l];
            // MLP: x_out = x1 + down(gelu(g) u); the model-dtype copy of dx is the dY operand of both products
            cast_scale<T>(ws.dx[m], (T *)ws.dyb, (long)M * Hm, 1.f, st);
            PZ_TRY(linear_bwd<T>(h, ws.dyb, Hm, ws.mm[m][l], Im, Lw.w_down, ws.d_m, Im, 0, G(Lg.w_down), M, Hm, Im, st));
            launch_k(geglu_bwd_kernel<T>, dim3(ew_blocks((long)M * Im / 8)), dim3(256), 0, st, (const T *)ws.gu[m][l], (const T *)ws.d_m,
                     (T *)ws.dgu, (long)M * Im, Im);
            launch_rmsnorm<T>(ws.x1[m][l], Lw.norm_post, (T *)ws.h, M, Hm, 1e-6f, st);
            PZ_TRY(linear_bwd<T>(h, ws.dgu, 2 * Im, ws.h, Hm, Lw.w_gate_up, ws.dh, Hm, LIN_OUT_F32, G(Lg.w_gate_up), M, 2 * Im, Hm, st));
            // dx becomes d / d x1; its model-dtype copy (the dY of the attention output projection) is written by the same kernel
            rmsnorm_bwd<T>(ws.x1[m][l], Lw.norm_post, ws.dh, ws.dx[m], G(Lg.norm_post), M, Hm, st, (T *)ws.dyb);
            // attention output projection: x1 = x + o_proj(att)
            PZ_TRY(linear_bwd<T>(h, ws.dyb, Hm, ws.att[m][l], qd, Lw.w_o, ws.datt[m], qd, 0, G(Lg.w_o), M, Hm, qd, st));
        }
        // attention
        if (tc_attn_bwd && hd == 256) {
            AttnBwd2Args b2;
            memset(&b2, 0, sizeof(b2));
            b2.K = (const bf16 *)ws.K[l]; b2.V = (const bf16 *)ws.V[l]; b2.kv_bs = kv_bs; b2.ld_kv = hd;
            b2.dK = ws.dK; b2.dV = ws.dV; b2.dkv_bs = kv_bs; b2.ld_dkv = hd;
            b2.lse = ws.lse; b2.dvec = ws.dvec; b2.valid_len = valid_len; b2.s_v = S_v; b2.s_p = S_p;
            b2.NK = S_all; b2.hd = hd; b2.hpr = nh; b2.groups = 1;
            b2.scale = 1.0f / sqrtf((float)hd); b2.softcap = 50.f;
            // dK / dV: every (segment, key block, query split) CTA adds its partial sum with atomics into zeroed buffers, so
            // the three segments are independent: the two small ones (proprio 8, action 32 query rows: latency-bound
            // launches) run on the side stream beside the vlm segment
            cudaMemsetAsync(ws.dK, 0, (size_t)B * S_all * hd * 4, st);
            cudaMemsetAsync(ws.dV, 0, (size_t)B * S_all * hd * 4, st);
            b2.qsplit = 4;
            cudaStream_t side = st;
            if (!last) {
                if (!h->side && cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess)
                    return fail(h, PZ_ERR_CUDA, "cannot create the side stream");
                side = h->side;
                while (h->sync_ev.size() < 2) {
                    cudaEvent_t e;
                    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
                    h->sync_ev.push_back(e);
                }
                cudaEventRecord(h->sync_ev[0], st);
                cudaStreamWaitEvent(side, h->sync_ev[0], 0);
            }
            for (int m = 2; m >= 0; --m) {
                if (last && m < 2) continue;
                // the [tokens, heads * hd] buffers seen as [tokens * heads, hd]: the heads of a token are consecutive query rows
                b2.Q = (const bf16 *)ws.q[m][l]; b2.q_bs = (long)md[m].rows * qd; b2.ld_q = hd;
                b2.dO = (const bf16 *)ws.datt[m]; b2.O = (const bf16 *)ws.att[m][l]; b2.o_bs = (long)md[m].rows * qd; b2.ld_o = hd;
                b2.dQ = ws.dq[m]; b2.dq_bs = (long)md[m].rows * qd; b2.ld_dq = hd;
                b2.NQ = md[m].rows * nh; b2.seg = m;
                // every segment has its own slice of the statistics scratch
                b2.lse = ws.lse + (size_t)B * row_off[m] * nh; b2.dvec = ws.dvec + (size_t)B * row_off[m] * nh;
                attn_bwd2_launch<256>(b2, B, true, true, m == 0 ? st : side);
            }
            if (side != st) {
                cudaEventRecord(h->sync_ev[1], side);
                cudaStreamWaitEvent(st, h->sync_ev[1], 0);
            }
        } else {
        cudaMemsetAsync(ws.dK, 0, (size_t)B * S_all * hd * 4, st);
        cudaMemsetAsync(ws.dV, 0, (size_t)B * S_all * hd * 4, st);
        AttnBwdArgs ab;
        memset(&ab, 0, sizeof(ab));
        ab.n_seg = 3;
        for (int m = 0; m < 3; ++m) {
            ab.seg_rows[m] = md[m].rows;
            ab.seg_active[m] = !(last && m < 2);
            ab.Q[m] = ws.q[m][l]; ab.dO[m] = ws.datt[m]; ab.O[m] = ws.att[m][l]; ab.dQ[m] = ws.dq[m];
        }
        ab.q_rs = ab.o_rs = ab.dq_rs = qd;
        ab.K = ws.K[l]; ab.V = ws.V[l]; ab.kv_bs = kv_bs; ab.kv_rs = hd; ab.kv_hs = 0;
        ab.dK = ws.dK; ab.dV = ws.dV; ab.dkv_bs = kv_bs; ab.dkv_rs = hd; ab.dkv_hs = 0;
        ab.valid_len = valid_len; ab.s_v = S_v; ab.s_p = S_p;
        ab.n_keys = S_all; ab.batch = B; ab.n_heads = nh; ab.hd = hd;
        ab.scale = 1.0f / sqrtf((float)hd); ab.softcap = 50.f;
        {
            int rc = attn_bwd<T>(ab, st, &err);
            if (rc) return fail(h, rc, err ? err : "attention backward failed");
        }
        }
        // q / k / v projections and the input norm
        for (int m = 0; m < 3; ++m) {
            const int M = B * md[m].rows, Hm = md[m].hidden;
            const pz_mix_layer &Lw = mixw[m][l];
            const pz_mix_layer &Lg = mixg[m][l];
            launch_k(rope_bwd_merge_kernel<T>, dim3(M), dim3(256), 0, st, (const float *)((last && m < 2) ? nullptr : ws.dq[m]),
                     (const float *)ws.dK, (const float *)ws.dV, kv_bs, row_off[m], rc_[m], rs_[m], pos0[m], (T *)ws.dqkv, md[m].rows, nh, hd);
            launch_rmsnorm<T>(ws.xin[m][l], Lw.norm_in, (T *)ws.h, M, Hm, 1e-6f, st);
            PZ_TRY(linear_bwd<T>(h, ws.dqkv, qkvd, ws.h, Hm, Lw.w_qkv, ws.dh, Hm, LIN_OUT_F32, G(Lg.w_qkv), M, qkvd, Hm, st));
            rmsnorm_bwd<T>(ws.xin[m][l], Lw.norm_in, ws.dh, ws.dx[m], G(Lg.norm_in), M, Hm, st);
        }
        mark(l);
    }
    // ---- action encoder (vla/modules.py:39-53); dx[2] = d / d (sqrt(A) * linear_3(z))
    {
        T *dy3 = (T *)ws.dyb;
        cast_scale<T>(ws.dx[2], dy3, (long)Ma * A, sqrtf((float)A), st);
        colsum<T>(dy3, A, G(g.enc_b3), Ma, A, st);
        PZ_TRY(linear_bwd<T>(h, dy3, A, ws.z, A, w.enc_w3, ws.dh, A, LIN_OUT_F32, G(g.enc_w3), Ma, A, A, st));
        T *dzp = (T *)ws.dqkv;
        launch_k(silu_bwd_kernel<T>, dim3(ew_blocks((long)Ma * A)), dim3(256), 0, st, (const float *)ws.zpre, (const float *)ws.tbias,
                 (const float *)ws.dh, dzp, (long)Ma * A, A, Hz);
        colsum<T>(dzp, A, G(g.enc_b2), Ma, A, st);
        T *de1 = (T *)ws.d_m;
        PZ_TRY(linear_bwd<T>(h, dzp, A, ws.e1, A, w.enc_w2a, de1, A, 0, G(g.enc_w2a), Ma, A, A, st));
        // time half: one input row per sample, its gradient is the sum over the sample's action tokens
        T *dzs = (T *)ws.dgu;
        launch_k(group_sum_kernel<T>, dim3((B * A + 255) / 256), dim3(256), 0, st, (const T *)dzp, dzs, (long)B * A, A, Hz);
        PZ_TRY(linear_bwd<T>(h, dzs, A, ws.temb, A, w.enc_w2t, nullptr, 0, 0, G(g.enc_w2t), B, A, A, st));
        colsum<T>(de1, A, G(g.enc_b1), Ma, A, st);
        PZ_TRY(linear_bwd<T>(h, de1, A, ws.a_in, skp, w.enc_w1, nullptr, 0, 0, G(g.enc_w1), Ma, A, skp, st));
    }
    // ---- proprio encoder (pizero.py:630)
    {
        T *dyp = (T *)ws.dyb;
        cast_scale<T>(ws.dx[1], dyp, (long)Mp * A, sqrtf((float)A), st);
        colsum<T>(dyp, A, G(g.prop_b), Mp, A, st);
        PZ_TRY(linear_bwd<T>(h, dyp, A, ws.pp, skp, w.prop_w, nullptr, 0, 0, G(g.prop_w), Mp, A, skp, st));
    }
    mark(L);
    if (!vit_grads) {
        mark(L + 1 + LV);
        return 0;
    }
    // ---- embedding merge (pizero.py:376-414): image rows of the merged sequence are the projector output (the
    // 1/sqrt(H) of the merge and the sqrt(H) of the joint model cancel); token embeddings are frozen (pizero.py:243-249)
    const int n_feat = c.n_images * P;
    float *dfe = ws.dh;   // fp32 [Mv, H]
    cudaMemcpy2DAsync(dfe, (size_t)n_feat * H * 4, ws.dx[0], (size_t)S_v * H * 4, (size_t)n_feat * H * 4, B, cudaMemcpyDeviceToDevice, st);
    T *dfb = (T *)ws.dyb;
    cast_scale<T>(dfe, dfb, (long)Mv * H, 1.f, st);
    colsum<T>(dfb, H, G(g.proj_b), Mv, H, st);
    float *dxv = ws.feats;   // the forward's projector output is no longer needed: fp32 [Mv, H] >= [Mv, V]
    PZ_TRY(linear_bwd<T>(h, dfb, H, ws.hv_post, V, w.proj_w, ws.dh, V, LIN_OUT_F32, G(g.proj_w), Mv, H, V, st));
    cudaMemsetAsync(dxv, 0, (size_t)Mv * V * 4, st);
    layernorm_bwd<T>(ws.xv[LV], w.post_ln_w, ws.dh, dxv, G(g.post_ln_w), G(g.post_ln_b), Mv, V, st, (T *)ws.dyb);
    for (int i = LV - 1; i >= 0; --i) {
        const pz_vit_layer &Lw = h->vit[i];
        const pz_vit_layer &Lg = gr->vit[i];
        // MLP (ws.dyb = the model-dtype copy of dxv, written by the norm backward in front)
        colsum<T>((const T *)ws.dyb, V, G(Lg.b_fc2), Mv, V, st);
        PZ_TRY(linear_bwd<T>(h, ws.dyb, V, ws.actv[i], VI, Lw.w_fc2, ws.d_m, VI, 0, G(Lg.w_fc2), Mv, V, VI, st));
        launch_k(gelu_bwd_kernel<T>, dim3(ew_blocks((long)Mv * VI)), dim3(256), 0, st, (const T *)ws.f1[i], (const T *)ws.d_m, (T *)ws.dgu,
                 (long)Mv * VI);
        colsum<T>((const T *)ws.dgu, VI, G(Lg.b_fc1), Mv, VI, st);
        launch_layernorm<T>(ws.xv_mid[i], Lw.ln2_w, Lw.ln2_b, (T *)ws.h, Mv, V, 1e-6f, st);
        PZ_TRY(linear_bwd<T>(h, ws.dgu, VI, ws.h, V, Lw.w_fc1, ws.dh, V, LIN_OUT_F32, G(Lg.w_fc1), Mv, VI, V, st));
        layernorm_bwd<T>(ws.xv_mid[i], Lw.ln2_w, ws.dh, dxv, G(Lg.ln2_w), G(Lg.ln2_b), Mv, V, st, (T *)ws.dyb);
        // attention
        colsum<T>((const T *)ws.dyb, V, G(Lg.b_o), Mv, V, st);
        T *dav = (T *)ws.d_m;
        PZ_TRY(linear_bwd<T>(h, ws.dyb, V, ws.av[i], V, Lw.w_o, dav, V, 0, G(Lg.w_o), Mv, V, V, st));
        float *dqkv32 = ws.dh;   // fp32 [Mv, 3V]: dq | dk | dv
        if (tc_attn_bwd && hdv == 72) {
            AttnBwd2Args b2;
            memset(&b2, 0, sizeof(b2));
            b2.Q = (const bf16 *)ws.qkvv[i]; b2.q_bs = (long)P * 3 * V; b2.q_gs = hdv; b2.ld_q = 3 * V;
            b2.dO = (const bf16 *)dav; b2.O = (const bf16 *)ws.av[i]; b2.o_bs = (long)P * V; b2.o_gs = hdv; b2.ld_o = V;
            b2.dQ = dqkv32; b2.dq_bs = (long)P * 3 * V; b2.dq_gs = hdv; b2.ld_dq = 3 * V;
            b2.K = (const bf16 *)ws.qkvv[i] + V; b2.V = (const bf16 *)ws.qkvv[i] + 2 * V; b2.kv_bs = (long)P * 3 * V; b2.kv_gs = hdv; b2.ld_kv = 3 * V;
            b2.dK = dqkv32 + V; b2.dV = dqkv32 + 2 * V; b2.dkv_bs = (long)P * 3 * V; b2.dkv_gs = hdv; b2.ld_dkv = 3 * V;
            b2.lse = ws.lse; b2.dvec = ws.dvec; b2.valid_len = nullptr;
            b2.NQ = P; b2.NK = P; b2.hd = hdv; b2.hpr = 1; b2.groups = c.vit_heads;
            b2.scale = 1.0f / sqrtf((float)hdv); b2.softcap = 0.f; b2.accumulate = 0;
            attn_bwd2_launch<80>(b2, n_img, true, true, st);
        } else {
        cudaMemsetAsync(dqkv32, 0, (size_t)Mv * 3 * V * 4, st);
        AttnBwdArgs ab;
        memset(&ab, 0, sizeof(ab));
        ab.n_seg = 1; ab.seg_rows[0] = P; ab.seg_active[0] = 1;
        ab.Q[0] = ws.qkvv[i]; ab.dO[0] = dav; ab.O[0] = ws.av[i]; ab.dQ[0] = dqkv32;
        ab.q_rs = 3 * V; ab.o_rs = V; ab.dq_rs = 3 * V;
        ab.K = (const T *)ws.qkvv[i] + V; ab.V = (const T *)ws.qkvv[i] + 2 * V; ab.kv_bs = (long)P * 3 * V; ab.kv_rs = 3 * V; ab.kv_hs = hdv;
        ab.dK = dqkv32 + V; ab.dV = dqkv32 + 2 * V; ab.dkv_bs = (long)P * 3 * V; ab.dkv_rs = 3 * V; ab.dkv_hs = hdv;
        ab.valid_len = nullptr; ab.n_keys = P; ab.batch = n_img; ab.n_heads = c.vit_heads; ab.hd = hdv;
        ab.scale = 1.0f / sqrtf((float)hdv); ab.softcap = 0.f;
        {
            int rc = attn_bwd<T>(ab, st, &err);
            if (rc) return fail(h, rc, err ? err : "attention backward failed");
        }
        }
        cast_scale<T>(dqkv32, (T *)ws.dqkv, (long)Mv * 3 * V, 1.f, st);
        colsum<T>((const T *)ws.dqkv, 3 * V, G(Lg.b_qkv), Mv, 3 * V, st);
        launch_layernorm<T>(ws.xv[i], Lw.ln1_w, Lw.ln1_b, (T *)ws.h, Mv, V, 1e-6f, st);
        PZ_TRY(linear_bwd<T>(h, ws.dqkv, 3 * V, ws.h, V, Lw.w_qkv, ws.dh, V, LIN_OUT_F32, G(Lg.w_qkv), Mv, 3 * V, V, st));
        layernorm_bwd<T>(ws.xv[i], Lw.ln1_w, ws.dh, dxv, G(Lg.ln1_w), G(Lg.ln1_b), Mv, V, st, (T *)ws.dyb);
        mark(L + 1 + i);
    }
    // ---- patch embedding (the convolution as a matrix product over unfolded patches) and the position table
    if (g.pos_emb) launch_k(period_sum_kernel, dim3((P * V + 255) / 256), dim3(256), 0, st, (const float *)dxv, G(g.pos_emb), (long)P * V, V, P, n_img);
    colsum<T>((const T *)ws.dyb, V, G(g.patch_b), Mv, V, st);
    PZ_TRY(linear_bwd<T>(h, ws.dyb, V, ws.patches, c.patch_k_pad, w.patch_w, nullptr, 0, 0, G(g.patch_w), Mv, V, c.patch_k_pad, st));
    mark(L + 1 + LV);
    return 0;
}

}  // namespace

// ------------------------------------------------------------------- ABI ----
extern "C" {

size_t pz_train_workspace_bytes(const pz_handle *h, int batch) {
    if (!h || batch < 1) return 0;
    return carve_train(h->cfg, batch, nullptr).total;
}

int pz_flow_matching_step(pz_handle *h, const int64_t *ids, const void *pixels, const int32_t *valid_len, const float *proprio,
                          const float *actions, const float *noise, const float *t, float sig_min, const pz_weights *grads,
                          float loss_scale, float *loss, void *ws, size_t ws_bytes, int B, int flags, void *const *events, int n_events,
                          void *stream) {
    if (!h) return PZ_ERR_INVALID;
    if (!h->bound) return fail(h, PZ_ERR_UNBOUND, "pz_bind_weights has not been called");
    if (B < 1 || B > h->cfg.max_batch) return fail(h, PZ_ERR_INVALID, "batch out of range (1..max_batch)");
    if (!ids || !pixels || !valid_len || !proprio || !actions || !noise || !t || !loss) return fail(h, PZ_ERR_INVALID, "null input");
    if (!ws || ws_bytes < pz_train_workspace_bytes(h, B)) return fail(h, PZ_ERR_WORKSPACE, "training workspace too small");
    if (((uintptr_t)ws) & 1023) return fail(h, PZ_ERR_WORKSPACE, "workspace must be 1 KiB aligned");
    if (h->cfg.vlm_hidden > 2048 || h->cfg.act_hidden > 2048 || h->cfg.vit_hidden > 2048 || h->cfg.vit_hidden % 4)
        return fail(h, PZ_ERR_INVALID, "training step: hidden sizes up to 2048, multiples of 4");
    Grads gr;
    if (grads) {
        if (!grads->vit || !grads->vlm || !grads->proprio || !grads->action) return fail(h, PZ_ERR_INVALID, "gradient layer tables missing");
        gr.g = *grads;
        gr.vit.assign(grads->vit, grads->vit + h->cfg.vit_layers);
        gr.mix[0].assign(grads->vlm, grads->vlm + h->cfg.n_layers);
        gr.mix[1].assign(grads->proprio, grads->proprio + h->cfg.n_layers);
        gr.mix[2].assign(grads->action, grads->action + h->cfg.n_layers);
    }
    h->lc.n = 0;
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = h->cfg.dtype == PZ_BF16
                 ? forward_backward<bf16>(h, ids, pixels, valid_len, proprio, actions, noise, t, sig_min, grads ? &gr : nullptr, loss,
                                          loss_scale, ws, B, flags, events, n_events, st)
                 : forward_backward<float>(h, ids, pixels, valid_len, proprio, actions, noise, t, sig_min, grads ? &gr : nullptr, loss,
                                           loss_scale, ws, B, flags, events, n_events, st);
    g_launch_counter = nullptr;
    if (rc) return rc;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(h, PZ_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
    }
    return PZ_OK;
}


int pz_grad_sumsq(const float *d_grad, size_t n, float *d_out, void *stream) {
    if (!d_grad || !d_out) return PZ_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > PZ_SUMSQ_SCRATCH) blocks = PZ_SUMSQ_SCRATCH;
    if (blocks < 1) blocks = 1;
    launch_k(sumsq_kernel, dim3((unsigned)blocks), dim3(256), 0, st, d_grad, n, d_out + 1);
    launch_k(sumsq_final_kernel, dim3(1), dim3(256), 0, st, (const float *)(d_out + 1), (int)blocks, d_out);
    return cudaPeekAtLastError() == cudaSuccess ? PZ_OK : PZ_ERR_CUDA;
}

int pz_adamw_step(float *d_master, float *d_grad, float *d_m, float *d_v, size_t begin, size_t end, const long long *d_entry_off,
                  void *const *d_entry_dst, const long long *d_entry_n, int n_entries, int dst_dtype, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, const float *d_sumsq, float max_grad_norm,
                  float grad_scale, int zero_grad, void *stream) {
    if (!d_master || !d_grad || !d_m || !d_v || !d_entry_off || !d_entry_dst || !d_entry_n || n_entries < 1 || step < 1 || end < begin ||
        (begin & 255))
        return PZ_ERR_INVALID;
    if (end == begin) return PZ_OK;
    AdamWArgs a;
    a.master = d_master; a.m = d_m; a.v = d_v; a.grad = d_grad; a.begin = begin; a.end = end;
    a.entry_off = d_entry_off; a.entry_dst = d_entry_dst; a.entry_n = d_entry_n; a.n_entries = n_entries;
    a.dst_bf16 = dst_dtype == PZ_BF16;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
    a.bc1 = 1.f - powf(beta1, (float)step);
    a.bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    a.sumsq = d_sumsq; a.max_norm = max_grad_norm; a.grad_scale = grad_scale; a.zero_grad = zero_grad; a.copy_only = 0;
    size_t blocks = (end - begin + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(adamw_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a);
    return cudaPeekAtLastError() == cudaSuccess ? PZ_OK : PZ_ERR_CUDA;
}


int pz_average_update(float *d_avg, const float *d_x, size_t n, float weight, void *stream) {
    if (!d_avg || !d_x) return PZ_ERR_INVALID;
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    launch_k(lerp_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, d_avg, d_x, n, weight);
    return cudaPeekAtLastError() == cudaSuccess ? PZ_OK : PZ_ERR_CUDA;
}

int pz_write_packed(const float *d_flat, size_t begin, size_t end, const long long *d_entry_off, void *const *d_entry_dst,
                    const long long *d_entry_n, int n_entries, int dst_dtype, void *stream) {
    if (!d_flat || !d_entry_off || !d_entry_dst || !d_entry_n || n_entries < 1 || end < begin || (begin & 255)) return PZ_ERR_INVALID;
    if (end == begin) return PZ_OK;
    AdamWArgs a;
    memset(&a, 0, sizeof(a));
    a.master = const_cast<float *>(d_flat); a.begin = begin; a.end = end;
    a.entry_off = d_entry_off; a.entry_dst = d_entry_dst; a.entry_n = d_entry_n; a.n_entries = n_entries;
    a.dst_bf16 = dst_dtype == PZ_BF16; a.copy_only = 1;
    size_t blocks = (end - begin + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(adamw_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a);
    return cudaPeekAtLastError() == cudaSuccess ? PZ_OK : PZ_ERR_CUDA;
}

}  // extern "C"
