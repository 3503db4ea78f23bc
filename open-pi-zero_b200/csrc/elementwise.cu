// elementwise.cu -- normalisation, layout and sampler kernels (HBM-bound; all
// global access is coalesced and, where alignment allows, 16-byte vectorised).
#include "common.cuh"
#include "kernels.h"

// ---- im2col for the 14x14/14 patch conv (siglip.py:69) ---------------------
// patches[img*P + py*G + px][c*ps*ps + ky*ps + kx] = pix[img][c][py*ps+ky][px*ps+kx]
template <typename T>
__global__ void im2col_kernel(const T *__restrict__ pix, T *__restrict__ patches, int image,
                              int patch, int k_pad, long total) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int k = i % k_pad;
    long row = i / k_pad;
    int G = image / patch, P = G * G;
    int img = row / P, p = row % P;
    int py = p / G, px = p % G;
    int kk = patch * patch;
    T v = from_f32<T>(0.f);
    if (k < 3 * kk) {
        int c = k / kk, r = k % kk, ky = r / patch, kx = r % patch;
        v = pix[(((long)img * 3 + c) * image + (py * patch + ky)) * image + px * patch + kx];
    }
    patches[i] = v;
}
template <typename T>
void launch_im2col(const T *pix, T *patches, int n_images, int image, int patch, int k_pad,
                   cudaStream_t st) {
    int G = image / patch;
    long total = (long)n_images * G * G * k_pad;
    launch_k(im2col_kernel<T>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, pix, patches, image, patch,
                                                                     k_pad, total);
}
// uint8 camera frames: the caller-side normalisation of VLAProcessor (processing.py:27-58,108-113:
// x * (1/255) in fp32, then (x - 0.5) / 0.5) is applied while the patches are gathered, so the host never
// materialises (or copies) float images
template <typename T>
__global__ void im2col_u8_kernel(const uint8_t *__restrict__ pix, T *__restrict__ patches, int image,
                                 int patch, int k_pad, long total) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int k = i % k_pad;
    long row = i / k_pad;
    int G = image / patch, P = G * G;
    int img = row / P, p = row % P;
    int py = p / G, px = p % G;
    int kk = patch * patch;
    float v = 0.f;
    if (k < 3 * kk) {
        int c = k / kk, r = k % kk, ky = r / patch, kx = r % patch;
        float u = (float)pix[(((long)img * 3 + c) * image + (py * patch + ky)) * image + px * patch + kx];
        v = __fdiv_rn(__fsub_rn(__fmul_rn(u, (float)(1.0 / 255.0)), 0.5f), 0.5f);   // no FMA contraction: same roundings as torch
    }
    patches[i] = from_f32<T>(v);
}
template <typename T>
void launch_im2col_u8(const uint8_t *pix, T *patches, int n_images, int image, int patch, int k_pad,
                      cudaStream_t st) {
    int G = image / patch;
    long total = (long)n_images * G * G * k_pad;
    launch_k(im2col_u8_kernel<T>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, pix, patches, image, patch,
                                                                        k_pad, total);
}
template void launch_im2col_u8<float>(const uint8_t *, float *, int, int, int, int, cudaStream_t);
template void launch_im2col_u8<bf16>(const uint8_t *, bf16 *, int, int, int, int, cudaStream_t);
template void launch_im2col<float>(const float *, float *, int, int, int, int, cudaStream_t);
template void launch_im2col<bf16>(const bf16 *, bf16 *, int, int, int, int, cudaStream_t);

// x[r][c] = table[r % period][c]   (position embedding broadcast, siglip.py:76)
__global__ void bcast_rows_kernel(float *__restrict__ x, const float *__restrict__ table,
                                  long total, int cols, int period) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    long r = i / cols;
    int c = i % cols;
    x[i] = table[(r % period) * cols + c];
}
void launch_bcast_rows(float *x, const float *table, long rows, int cols, int period,
                       cudaStream_t st) {
    long total = rows * cols;
    launch_k(bcast_rows_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x, table, total, cols,
                                                                      period);
}

// ---- norms: one warp per row, the row lives in registers (one HBM read), vectorised stores -------------
// (rows of up to 32 * 4 * NI floats; NI float4 per lane, all loads issued before the first use)
template <typename T> PZ_DEVINL void store4(T *o, float a, float b, float c, float d);
template <> PZ_DEVINL void store4<float>(float *o, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(o) = make_float4(a, b, c, d);
}
template <> PZ_DEVINL void store4<bf16>(bf16 *o, float a, float b, float c, float d) {
    *reinterpret_cast<uint2 *>(o) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}

// LayerNorm (siglip.py:211,217,298), fp32 statistics, two-pass variance on the register copy
template <typename T, int NI>
__global__ void __launch_bounds__(256) layernorm_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ b, T *__restrict__ out, long rows,
                                                        int cols, float eps) {
    pdl_trigger();
    pdl_wait();
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4 *xr = reinterpret_cast<const float4 *>(x + row * cols);
    const int n4 = cols >> 2;
    float4 v[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < n4 ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        if (lane + 32 * i < n4) {
            const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
            q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / cols + eps);
    const float4 *wr = reinterpret_cast<const float4 *>(w), *br = reinterpret_cast<const float4 *>(b);
    T *o = out + row * cols;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int c = lane + 32 * i;
        if (c < n4) {
            const float4 g = __ldg(wr + c), h = __ldg(br + c);
            store4<T>(o + 4 * c, (v[i].x - mean) * rstd * g.x + h.x, (v[i].y - mean) * rstd * g.y + h.y,
                      (v[i].z - mean) * rstd * g.z + h.z, (v[i].w - mean) * rstd * g.w + h.w);
        }
    }
}
// any width (scalar, three passes): rows wider than 2048 or not a multiple of 4
template <typename T>
__global__ void __launch_bounds__(256) layernorm_generic_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                const float *__restrict__ b, T *__restrict__ out,
                                                                long rows, int cols, float eps) {
    pdl_trigger();
    pdl_wait();
    long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float *xr = x + row * cols;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += xr[c];
    float mean = warp_sum(s) / cols;
    float v = 0.f;
    for (int c = lane; c < cols; c += 32) { float d = xr[c] - mean; v += d * d; }
    float rstd = rsqrtf(warp_sum(v) / cols + eps);
    T *o = out + row * cols;
    for (int c = lane; c < cols; c += 32) o[c] = from_f32<T>((xr[c] - mean) * rstd * w[c] + b[c]);
}
template <typename T>
void launch_layernorm(const float *x, const float *w, const float *b, T *out, long rows, int cols,
                      float eps, cudaStream_t st) {
    const dim3 grid((unsigned)((rows + 7) / 8)), block(256);
    const int ni = ((cols >> 2) + 31) / 32;
    if (cols % 4 == 0 && ni <= 4) launch_k(layernorm_kernel<T, 4>, grid, block, 0, st, x, w, b, out, rows, cols, eps);
    else if (cols % 4 == 0 && ni <= 9) launch_k(layernorm_kernel<T, 9>, grid, block, 0, st, x, w, b, out, rows, cols, eps);
    else if (cols % 4 == 0 && ni <= 16) launch_k(layernorm_kernel<T, 16>, grid, block, 0, st, x, w, b, out, rows, cols, eps);
    else launch_k(layernorm_generic_kernel<T>, grid, block, 0, st, x, w, b, out, rows, cols, eps);
}
template void launch_layernorm<float>(const float *, const float *, const float *, float *, long,
                                      int, float, cudaStream_t);
template void launch_layernorm<bf16>(const float *, const float *, const float *, bf16 *, long,
                                     int, float, cudaStream_t);

// Gemma RMSNorm (paligemma/modules.py:13-21): x * rsqrt(mean x^2 + eps) * (1 + w)
template <typename T, int NI>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                      T *__restrict__ out, long rows, int cols, float eps) {
    pdl_trigger();
    pdl_wait();
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4 *xr = reinterpret_cast<const float4 *>(x + row * cols);
    const int n4 = cols >> 2;
    float4 v[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < n4 ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) s += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    const float r = rsqrtf(warp_sum(s) / cols + eps);
    const float4 *wr = reinterpret_cast<const float4 *>(w);
    T *o = out + row * cols;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int c = lane + 32 * i;
        if (c < n4) {
            const float4 g = __ldg(wr + c);
            store4<T>(o + 4 * c, v[i].x * r * (1.f + g.x), v[i].y * r * (1.f + g.y), v[i].z * r * (1.f + g.z),
                      v[i].w * r * (1.f + g.w));
        }
    }
}
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_generic_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                              T *__restrict__ out, long rows, int cols, float eps) {
    pdl_trigger();
    pdl_wait();
    long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float *xr = x + row * cols;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += xr[c] * xr[c];
    float r = rsqrtf(warp_sum(s) / cols + eps);
    T *o = out + row * cols;
    for (int c = lane; c < cols; c += 32) o[c] = from_f32<T>(xr[c] * r * (1.f + w[c]));
}
template <typename T>
void launch_rmsnorm(const float *x, const float *w, T *out, long rows, int cols, float eps,
                    cudaStream_t st) {
    const dim3 grid((unsigned)((rows + 7) / 8)), block(256);
    const int ni = ((cols >> 2) + 31) / 32;
    if (cols % 4 == 0 && ni <= 2) launch_k(rmsnorm_kernel<T, 2>, grid, block, 0, st, x, w, out, rows, cols, eps);
    else if (cols % 4 == 0 && ni <= 8) launch_k(rmsnorm_kernel<T, 8>, grid, block, 0, st, x, w, out, rows, cols, eps);
    else if (cols % 4 == 0 && ni <= 16) launch_k(rmsnorm_kernel<T, 16>, grid, block, 0, st, x, w, out, rows, cols, eps);
    else launch_k(rmsnorm_generic_kernel<T>, grid, block, 0, st, x, w, out, rows, cols, eps);
}
template void launch_rmsnorm<float>(const float *, const float *, float *, long, int, float,
                                    cudaStream_t);
template void launch_rmsnorm<bf16>(const float *, const float *, bf16 *, long, int, float,
                                   cudaStream_t);

// ---- embedding merge (pizero.py:385-413 + joint_model.py:348-355) -------------
// x[b][s] = feats[b][rank of s among image tokens]            if ids == image token
//         = 0                                                  if ids == pad
//         = embed[ids] * sqrt(hidden)                          otherwise
// feats is the projector output; the reference divides it by sqrt(hidden) and
// the joint model multiplies every row by sqrt(hidden) again, so image rows
// enter layer 0 unscaled.
template <typename T>
__global__ void __launch_bounds__(256) embed_merge_kernel(
    const int64_t *__restrict__ ids, const T *__restrict__ embed, const float *__restrict__ feats,
    float *__restrict__ x, int s_vlm, int hidden, int n_feat_rows, int image_token, int pad_token,
    float text_scale) {
    pdl_trigger();
    pdl_wait();
    int b = blockIdx.y, s = blockIdx.x;
    const int64_t *row_ids = ids + (long)b * s_vlm;
    long id = row_ids[s];
    float *xr = x + ((long)b * s_vlm + s) * hidden;
    __shared__ int rank_sh;
    if (id == image_token) {
        // rank = number of image tokens before s (block-wide count)
        int cnt = 0;
        for (int j = threadIdx.x; j < s; j += blockDim.x) cnt += (row_ids[j] == image_token);
        __shared__ float scratch[32];
        int rank = (int)(block_sum((float)cnt, scratch) + 0.5f);
        if (threadIdx.x == 0) rank_sh = rank;
        __syncthreads();
        rank = rank_sh;
        if (rank < n_feat_rows) {
            const float *f = feats + ((long)b * n_feat_rows + rank) * hidden;
            for (int c = threadIdx.x; c < hidden; c += blockDim.x) xr[c] = f[c];
        } else {
            for (int c = threadIdx.x; c < hidden; c += blockDim.x) xr[c] = 0.f;
        }
    } else if (id == pad_token) {
        for (int c = threadIdx.x; c < hidden; c += blockDim.x) xr[c] = 0.f;
    } else {
        const T *e = embed + id * hidden;
        for (int c = threadIdx.x; c < hidden; c += blockDim.x) xr[c] = to_f32<T>(e[c]) * text_scale;
    }
}
template <typename T>
void launch_embed_merge(const int64_t *ids, const T *embed, const float *feats, float *x,
                        int batch, int s_vlm, int hidden, int n_feat_rows, int image_token,
                        int pad_token, float text_scale, cudaStream_t st) {
    dim3 grid(s_vlm, batch);
    launch_k(embed_merge_kernel<T>, dim3(grid), dim3(256), 0, st, ids, embed, feats, x, s_vlm, hidden, n_feat_rows,
                                               image_token, pad_token, text_scale);
}
template void launch_embed_merge<float>(const int64_t *, const float *, const float *, float *,
                                        int, int, int, int, int, int, float, cudaStream_t);
template void launch_embed_merge<bf16>(const int64_t *, const bf16 *, const float *, float *, int,
                                       int, int, int, int, int, float, cudaStream_t);

// ---- RoPE + Q/K/V split (mixture.py:187-235, model/utils.py:4-16) ------------
// qkv row = [q (nh*hd) | k (hd) | v (hd)] (one KV head).  Half-split rotation:
// out[i] = x[i]*cos[i] - x[i+hd/2]*sin[i],  out[i+hd/2] = x[i+hd/2]*cos[i] + x[i]*sin[i].
// One block per token row; thread t handles pair index t of each head.
template <typename T>
__global__ void rope_split_kernel(const T *__restrict__ qkv, int qkv_ld, T *__restrict__ q_out,
                                  long q_batch_stride, T *__restrict__ k_out,
                                  T *__restrict__ v_out, long kv_batch_stride,
                                  const float *__restrict__ cos_t, const float *__restrict__ sin_t,
                                  int s_x, int pos0, int n_heads, int head_dim) {
    pdl_trigger();
    pdl_wait();
    int row = blockIdx.x;
    int b = row / s_x, s = row % s_x;
    int half = head_dim >> 1;
    const T *in = qkv + (long)row * qkv_ld;
    const float *cs = cos_t + (long)(pos0 + s) * half;
    const float *sn = sin_t + (long)(pos0 + s) * half;
    T *qo = q_out + b * q_batch_stride + (long)s * n_heads * head_dim;
    T *ko = k_out + b * kv_batch_stride + (long)s * head_dim;
    T *vo = v_out + b * kv_batch_stride + (long)s * head_dim;
    for (int i = threadIdx.x; i < (n_heads + 1) * half; i += blockDim.x) {
        int h = i / half, p = i % half;
        float c = cs[p], sv = sn[p];
        float x1 = to_f32<T>(in[h * head_dim + p]), x2 = to_f32<T>(in[h * head_dim + p + half]);
        float o1 = x1 * c - x2 * sv, o2 = x2 * c + x1 * sv;
        T *dst = (h < n_heads) ? qo + h * head_dim : ko;
        dst[p] = from_f32<T>(o1);
        dst[p + half] = from_f32<T>(o2);
    }
    const T *vin = in + (n_heads + 1) * head_dim;
    for (int i = threadIdx.x; i < head_dim; i += blockDim.x) vo[i] = vin[i];
}
template <typename T>
void launch_rope_split(const T *qkv, int qkv_ld, T *q_out, long q_batch_stride, T *k_out,
                       T *v_out, long kv_batch_stride, const float *cos_t, const float *sin_t,
                       int batch, int s_x, int pos0, int n_heads, int head_dim,
                       cudaStream_t st) {
    launch_k(rope_split_kernel<T>, dim3(batch * s_x), dim3(256), 0, st, qkv, qkv_ld, q_out, q_batch_stride, k_out,
                                                      v_out, kv_batch_stride, cos_t, sin_t, s_x,
                                                      pos0, n_heads, head_dim);
}
template void launch_rope_split<float>(const float *, int, float *, long, float *, float *, long,
                                       const float *, const float *, int, int, int, int, int,
                                       cudaStream_t);
template void launch_rope_split<bf16>(const bf16 *, int, bf16 *, long, bf16 *, bf16 *, long,
                                      const float *, const float *, int, int, int, int, int,
                                      cudaStream_t);

// ---- small casts --------------------------------------------------------------
// optional per-column affine + clip in front of the cast: the caller-side proprio normalisation
// (env_adapter/base.py:8-49, simpler.py:76-90) folded into the first kernel that touches the raw proprio
template <typename T>
__global__ void cast_pad_kernel(const float *__restrict__ src, T *__restrict__ dst, long total,
                                int cols, int cols_pad, const float *__restrict__ scale, const float *__restrict__ shift,
                                int clip) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    long r = i / cols_pad;
    int c = i % cols_pad;
    float v = c < cols ? src[r * cols + c] : 0.f;
    if (scale && c < cols) {
        v = v * scale[c] + shift[c];
        if (clip) v = fminf(fmaxf(v, -1.f), 1.f);
    }
    dst[i] = from_f32<T>(v);
}
template <typename T>
void launch_cast_pad(const float *src, T *dst, long rows, int cols, int cols_pad,
                     cudaStream_t st) {
    long total = rows * cols_pad;
    launch_k(cast_pad_kernel<T>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, src, dst, total, cols,
             cols_pad, (const float *)nullptr, (const float *)nullptr, 0);
}
template <typename T>
void launch_cast_pad_affine(const float *src, T *dst, long rows, int cols, int cols_pad, const float *scale,
                            const float *shift, int clip, cudaStream_t st) {
    long total = rows * cols_pad;
    launch_k(cast_pad_kernel<T>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, src, dst, total, cols,
             cols_pad, scale, shift, clip);
}
template void launch_cast_pad_affine<float>(const float *, float *, long, int, int, const float *, const float *, int, cudaStream_t);
template void launch_cast_pad_affine<bf16>(const float *, bf16 *, long, int, int, const float *, const float *, int, cudaStream_t);
// x[r][c] = x[r][c] * scale[c] + shift[c]: the caller-side action de-normalisation (simpler.py:102-125)
__global__ void affine_cols_kernel(float *__restrict__ x, long total, int cols, const float *__restrict__ scale,
                                   const float *__restrict__ shift) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % cols);
    x[i] = x[i] * scale[c] + shift[c];
}
void launch_affine_cols(float *x, long rows, int cols, const float *scale, const float *shift, cudaStream_t st) {
    long total = rows * cols;
    launch_k(affine_cols_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x, total, cols, scale, shift);
}
template void launch_cast_pad<float>(const float *, float *, long, int, int, cudaStream_t);
template void launch_cast_pad<bf16>(const float *, bf16 *, long, int, int, cudaStream_t);

template <typename T>
__global__ void to_f32_kernel(const T *__restrict__ src, float *__restrict__ dst, long n) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = to_f32<T>(src[i]);
}
template <typename T>
void launch_to_f32(const T *src, float *dst, long n, cudaStream_t st) {
    launch_k(to_f32_kernel<T>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, src, dst, n);
}
template void launch_to_f32<float>(const float *, float *, long, cudaStream_t);
template void launch_to_f32<bf16>(const bf16 *, float *, long, cudaStream_t);

// ---- Euler update (pizero.py:479-481): action += dt * velocity ---------------
__global__ void euler_kernel(float *__restrict__ action, const float *__restrict__ vel,
                             int vel_ld, float dt, long total, int adim,
                             float *__restrict__ vel_capture) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    long r = i / adim;
    int c = i % adim;
    float v = vel[r * vel_ld + c];
    if (vel_capture) vel_capture[i] = v;
    action[i] += dt * v;
}
void launch_euler(float *action, const float *vel, int vel_ld, float dt, long rows, int adim,
                  float *vel_capture, cudaStream_t st) {
    long total = rows * adim;
    launch_k(euler_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, action, vel, vel_ld, dt, total,
                                                                 adim, vel_capture);
}

// final clamp (pizero.py:484-489); clip < 0 => copy only
__global__ void clamp_copy_kernel(const float *__restrict__ src, float *__restrict__ dst, long n,
                                  float clip) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = src[i];
    if (clip >= 0.f) v = fminf(fmaxf(v, -clip), clip);
    dst[i] = v;
}
void launch_clamp_copy(const float *src, float *dst, long n, float clip, cudaStream_t st) {
    launch_k(clamp_copy_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, src, dst, n, clip);
}

// ---- flow-matching training forward (pizero.py:597-661): the small pieces around the joint model ----
// psi_t (pizero.py:597-605): (1 - (1 - sig_min) t_b) x0 + t_b x1
__global__ void psi_kernel(const float *__restrict__ x0, const float *__restrict__ x1, const float *__restrict__ t,
                           float *__restrict__ out, long total, int per_sample, float sig_min) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float tb = t[i / per_sample];
    out[i] = (1.f - (1.f - sig_min) * tb) * x0[i] + tb * x1[i];
}
void launch_psi(const float *x0, const float *x1, const float *t, float *out, long batch, int per_sample, float sig_min,
                cudaStream_t st) {
    long total = batch * per_sample;
    launch_k(psi_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x0, x1, t, out, total, per_sample, sig_min);
}

// SinusoidalPosEmb (vla/modules.py:9-22): out[b] = cat(sin(t_b f), cos(t_b f)), f_i = exp(-i ln(P) / (half - 1)) (fp32 table)
template <typename T>
__global__ void time_embed_kernel(const float *__restrict__ t, const float *__restrict__ freq, T *__restrict__ out, int batch,
                                  int half) {
    pdl_trigger();
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * half) return;
    const int b = i / half, j = i % half;
    float s, c;
    sincosf(t[b] * freq[j], &s, &c);
    out[(long)b * 2 * half + j] = from_f32<T>(s);
    out[(long)b * 2 * half + half + j] = from_f32<T>(c);
}
template <typename T>
void launch_time_embed(const float *t, const float *freq, T *out, int batch, int half, cudaStream_t st) {
    launch_k(time_embed_kernel<T>, dim3((unsigned)((batch * half + 255) / 256)), dim3(256), 0, st, t, freq, out, batch, half);
}
template void launch_time_embed<float>(const float *, const float *, float *, int, int, cudaStream_t);
template void launch_time_embed<bf16>(const float *, const float *, bf16 *, int, int, cudaStream_t);

// ActionEncoder.linear_2 + SiLU with a per-sample time half (vla/modules.py:46-52): the concatenated input
// [time_emb | linear_1(a)] splits the product into a per-sample vector (bias, [batch, cols]) plus the action half (zpre)
template <typename T>
__global__ void rowbias_silu_kernel(const float *__restrict__ zpre, const float *__restrict__ bias, T *__restrict__ out,
                                    long total, int cols, int rows_per_sample) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long r = i / cols;
    const int c = i % cols;
    out[i] = from_f32<T>(silu(zpre[i] + bias[(r / rows_per_sample) * cols + c]));
}
template <typename T>
void launch_rowbias_silu(const float *zpre, const float *bias, T *out, long rows, int cols, int rows_per_sample,
                         cudaStream_t st) {
    long total = rows * cols;
    launch_k(rowbias_silu_kernel<T>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, zpre, bias, out, total, cols,
             rows_per_sample);
}
template void launch_rowbias_silu<float>(const float *, const float *, float *, long, int, int, cudaStream_t);
template void launch_rowbias_silu<bf16>(const float *, const float *, bf16 *, long, int, int, cudaStream_t);

// flow-matching loss (pizero.py:658-661): mean((v - (x1 - (1 - sig_min) x0))^2); one CTA (the tensor is B*horizon*action_dim)
__global__ void __launch_bounds__(1024) fm_loss_kernel(const float *__restrict__ vel, int vel_ld, const float *__restrict__ x0,
                                                       const float *__restrict__ x1, float *__restrict__ loss,
                                                       float *__restrict__ v_out, long total, int adim, float sig_min) {
    __shared__ float scratch[32];
    pdl_trigger();
    pdl_wait();
    float acc = 0.f;
    for (long i = threadIdx.x; i < total; i += blockDim.x) {
        const float v = vel[(i / adim) * vel_ld + (i % adim)];
        if (v_out) v_out[i] = v;
        const float d = v - (x1[i] - (1.f - sig_min) * x0[i]);
        acc += d * d;
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) *loss = acc / (float)total;
}
void launch_fm_loss(const float *vel, int vel_ld, const float *x0, const float *x1, float *loss, float *v_out, long rows,
                    int adim, float sig_min, cudaStream_t st) {
    launch_k(fm_loss_kernel, dim3(1), dim3(1024), 0, st, vel, vel_ld, x0, x1, loss, v_out, rows * adim, adim, sig_min);
}
