// skinny.cu -- weight-streaming linear layer for M <= 64 activation rows (the
// action expert at small batch: M = B * horizon, and the single proprio token
// of the prefix pass).  HBM-bound: every weight byte is read exactly once, in
// 16-byte loads that bypass L1 allocation; the M activation rows stay in L1/L2.
//
//   y^T[N, M] = W[N, K] . x^T[K, M]
//
// The weight matrix is the *A* operand of mma.sync.m16n8k16 (16 weight rows per
// warp-tile), the activations are the 8-wide B operand, so M <= 8 costs one MMA
// per 16 x 16 weight block and tensor throughput is never the limit.  Because a
// dot product is invariant under a permutation of k applied to both operands,
// each thread feeds two MMAs straight from one contiguous 16-byte chunk per
// row (no shared-memory transpose of W).  One CTA = one block of 16 output
// features; its 8 warps split K and reduce through shared memory; for
// fp32-accumulate outputs (o_proj / down_proj + residual) K is additionally
// split across CTAs and combined with fp32 atomics.
//
// Reference call sites: mixture.py:187-218, paligemma/modules.py:86-95,
// vla/modules.py:39-53, pizero.py:436,479.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int NWARPS = 8, NTHREADS = NWARPS * 32, MAX_MT = 8;

PZ_DEVINL uint4 ldg_stream(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
PZ_DEVINL void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                        uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct SkinnyParams {
    const bf16 *A;   // x [M, K], row stride lda  (fp32 when LIN_NORM_A)
    const bf16 *W;   // [N, K]
    const float *norm_w;
    int cmb_splits, cmb_q_rows, cmb_heads, cmb_hd;
    const float *bias;
    void *C;
    int M, N, K, lda, ldc, flags, ksplit, n_mt;
    float alpha;
};

template <int MT>   // number of 8-row activation tiles per CTA (rows per CTA <= 8*MT)
__global__ void __launch_bounds__(NTHREADS, MT <= 2 ? 3 : 1) skinny_kernel(SkinnyParams p) {
    __shared__ float red[NWARPS][16][MT * 8 + 1];
    __shared__ float s_rs[MT * 8];
    __shared__ float s_cw[(MT <= 2) ? MT * 8 * 8 * 16 : 1];   // combine weights [m][head][split] (M<=16, <=8 heads, <=16 splits)
    {   // blockIdx.z selects a chunk of 8*MT activation rows (M > 64 with a tiny N, e.g. the action decoder)
        const int m0 = blockIdx.z * (8 * MT);
        const size_t esz = (p.flags & LIN_NORM_A) ? 4 : 2;
        p.A = reinterpret_cast<const bf16 *>(reinterpret_cast<const char *>(p.A) + (size_t)m0 * p.lda * esz);
        p.C = reinterpret_cast<char *>(p.C) + (size_t)m0 * p.ldc * ((p.flags & LIN_OUT_F32) ? 4 : 2);
        p.M = min(8 * MT, p.M - m0);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const bool geglu = p.flags & LIN_GEGLU;
    const int nblk = blockIdx.x;

    // weight rows of this block: rows g and g+8 of the 16-row tile
    long wrow0, wrow1;
    if (geglu) {   // 8 gate rows + the 8 matching up rows (packed [128 gate | 128 up] blocks)
        int c = nblk * 8 + g;
        wrow0 = (long)(c / PZ_GU_BLOCK) * (2 * PZ_GU_BLOCK) + (c % PZ_GU_BLOCK);
        wrow1 = wrow0 + PZ_GU_BLOCK;
    } else {
        int r0 = nblk * 16 + g, r1 = r0 + 8;
        wrow0 = r0 < p.N ? r0 : p.N - 1;    // clamp; masked at the store
        wrow1 = r1 < p.N ? r1 : p.N - 1;
    }
    const bf16 *w0 = p.W + wrow0 * p.K;
    const bf16 *w1 = p.W + wrow1 * p.K;

    // this warp's K range (multiples of 64 so that every chunk is a full 16-byte load)
    const int parts = NWARPS * p.ksplit;
    const int per = ((p.K + parts - 1) / parts + 63) / 64 * 64;
    const int part = blockIdx.y * NWARPS + warp;
    const int kbeg = part * per;
    const int kend = min(p.K, kbeg + per);

    float acc[MT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    // up to 8 independent 16-byte weight loads in flight per thread
    uint4 a[2][4];
    auto load_w = [&](int k0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int hk = 0; hk < 2; ++hk) {
                int k = k0 + u * 64 + hk * 32 + 8 * t;
                bool ok = k < kend;
                a[u][hk * 2 + 0] = ok ? ldg_stream(w0 + k) : zero4;
                a[u][hk * 2 + 1] = ok ? ldg_stream(w1 + k) : zero4;
            }
        }
    };
    // Weights do not depend on the previous kernel: start streaming them, let the next
    // kernel start its own prefetch, and only then wait for the activations (PDL).
    load_w(kbeg);
    pdl_trigger();
    pdl_wait();
    const bool norm_a = p.flags & LIN_NORM_A;
    if (norm_a) {
        // fused Gemma RMSNorm (paligemma/modules.py:13-21): every CTA recomputes the M row
        // statistics of the fp32 residual (M*K*4 bytes from L2) instead of a separate kernel
        const float *X = reinterpret_cast<const float *>(p.A);
        for (int m = warp; m < p.M; m += NWARPS) {
            const float4 *xr = reinterpret_cast<const float4 *>(X + (long)m * p.lda);
            float ss = 0.f;
            for (int c = lane; c < p.K / 4; c += 32) {
                float4 v = __ldg(xr + c);
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
            ss = warp_sum(ss);
            if (lane == 0) s_rs[m] = rsqrtf(ss / p.K + 1e-6f);
        }
        __syncthreads();
    }
    const bool cmb_a = p.flags & LIN_COMBINE_A;
    if (cmb_a) {
        // softmax-combine weights of the split-key attention partials: w_s = exp(m_s - max) / sum_s l_s exp(m_s - max)
        const float *P = reinterpret_cast<const float *>(p.A);
        const int rows_total = p.cmb_heads * p.cmb_q_rows, stride = p.cmb_hd + 2;
        const int m_base = blockIdx.z * (8 * MT);
        for (int i = threadIdx.x; i < p.M * p.cmb_heads; i += NTHREADS) {
            int m = i / p.cmb_heads, hh = i % p.cmb_heads;
            int gm = m_base + m, b = gm / p.cmb_q_rows, tok = gm % p.cmb_q_rows;
            const float *base = P + ((long)b * p.cmb_splits * rows_total + (hh * p.cmb_q_rows + tok)) * stride;
            float mx = -INFINITY;
            for (int sp = 0; sp < p.cmb_splits; ++sp) mx = fmaxf(mx, base[(long)sp * rows_total * stride + p.cmb_hd]);
            float l = 0.f;
            for (int sp = 0; sp < p.cmb_splits; ++sp) {
                const float *q = base + (long)sp * rows_total * stride;
                float wgt = (q[p.cmb_hd] == -INFINITY) ? 0.f : __expf(q[p.cmb_hd] - mx);
                s_cw[(m * 8 + hh) * 16 + sp] = wgt;
                l += q[p.cmb_hd + 1] * wgt;
            }
            float inv = l > 0.f ? 1.f / l : 0.f;
            for (int sp = 0; sp < p.cmb_splits; ++sp) s_cw[(m * 8 + hh) * 16 + sp] *= inv;
        }
        __syncthreads();
    }
    for (int k0 = kbeg; k0 < kend; k0 += 128) {
        if (k0 > kbeg) load_w(k0);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int hk = 0; hk < 2; ++hk) {
                int k = k0 + u * 64 + hk * 32 + 8 * t;
                bool ok = k < kend;
                const uint4 ag = a[u][hk * 2], ag8 = a[u][hk * 2 + 1];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    int m = mt * 8 + g;
                    uint4 x = zero4;
                    if (ok && m < p.M) {
                        if (norm_a) {
                            const float *xp = reinterpret_cast<const float *>(p.A) + (long)m * p.lda + k;
                            float4 x0 = __ldg(reinterpret_cast<const float4 *>(xp)), x1 = __ldg(reinterpret_cast<const float4 *>(xp + 4));
                            float4 w0v = __ldg(reinterpret_cast<const float4 *>(p.norm_w + k)), w1v = __ldg(reinterpret_cast<const float4 *>(p.norm_w + k + 4));
                            float r = s_rs[m];
                            x.x = pack_bf16x2(x0.x * r * (1.f + w0v.x), x0.y * r * (1.f + w0v.y));
                            x.y = pack_bf16x2(x0.z * r * (1.f + w0v.z), x0.w * r * (1.f + w0v.w));
                            x.z = pack_bf16x2(x1.x * r * (1.f + w1v.x), x1.y * r * (1.f + w1v.y));
                            x.w = pack_bf16x2(x1.z * r * (1.f + w1v.z), x1.w * r * (1.f + w1v.w));
                        } else if (cmb_a) {
                            const float *P = reinterpret_cast<const float *>(p.A);
                            const int rows_total = p.cmb_heads * p.cmb_q_rows, stride = p.cmb_hd + 2;
                            int gm = blockIdx.z * (8 * MT) + m, b = gm / p.cmb_q_rows, tok = gm % p.cmb_q_rows;
                            int hh = k / p.cmb_hd, d = k % p.cmb_hd;
                            const float *src = P + ((long)b * p.cmb_splits * rows_total + (hh * p.cmb_q_rows + tok)) * stride + d;
                            float o8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            for (int sp = 0; sp < p.cmb_splits; ++sp) {
                                const float2 *q = reinterpret_cast<const float2 *>(src + (long)sp * rows_total * stride);
                                float wgt = s_cw[(m * 8 + hh) * 16 + sp];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float2 v = __ldg(q + j);
                                    o8[2 * j] += v.x * wgt; o8[2 * j + 1] += v.y * wgt;
                                }
                            }
                            x.x = pack_bf16x2(o8[0], o8[1]); x.y = pack_bf16x2(o8[2], o8[3]);
                            x.z = pack_bf16x2(o8[4], o8[5]); x.w = pack_bf16x2(o8[6], o8[7]);
                        } else {
                            x = __ldg(reinterpret_cast<const uint4 *>(p.A + (long)m * p.lda + k));
                        }
                    }
                    mma_bf16(acc[mt], ag.x, ag8.x, ag.y, ag8.y, x.x, x.y);
                    mma_bf16(acc[mt], ag.z, ag8.z, ag.w, ag8.w, x.z, x.w);
                }
            }
        }
    }
    // accumulator layout: c0,c1 = (weight row g,   m = mt*8 + 2t, +1); c2,c3 = (weight row g+8, same m)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        red[warp][g][mt * 8 + 2 * t] = acc[mt][0];
        red[warp][g][mt * 8 + 2 * t + 1] = acc[mt][1];
        red[warp][g + 8][mt * 8 + 2 * t] = acc[mt][2];
        red[warp][g + 8][mt * 8 + 2 * t + 1] = acc[mt][3];
    }
    __syncthreads();

    const int n_rows_out = geglu ? 8 : 16;
    for (int i = threadIdx.x; i < n_rows_out * p.M; i += NTHREADS) {
        int r = i % n_rows_out, m = i / n_rows_out;
        float v = 0.f, v2 = 0.f;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) {
            v += red[w][r][m];
            if (geglu) v2 += red[w][r + 8][m];
        }
        int n = geglu ? nblk * 8 + r : nblk * 16 + r;
        int n_out = geglu ? p.N / 2 : p.N;
        if (n >= n_out) continue;
        if (geglu) {
            v = gelu_tanh(v) * v2;
        } else {
            if (p.bias && blockIdx.y == 0) v += p.bias[n];
            if (p.flags & LIN_GELU) v = gelu_tanh(v);
            if (p.flags & LIN_SILU) v = silu(v);
        }
        v *= p.alpha;
        long o = (long)m * p.ldc + n;
        if (p.flags & LIN_OUT_F32) {
            float *C = (float *)p.C;
            if (p.flags & LIN_ACCUM) atomicAdd(C + o, v);
            else C[o] = v;
        } else {
            ((bf16 *)p.C)[o] = __float2bfloat16_rn(v);
        }
    }
}

}  // namespace

int skinny_supported(const LinearArgs &a) {
    // M <= 64: the weight-streaming regime.  Larger M only for tiny N (weights re-read per 64-row
    // chunk stay in L2): the 7-wide action decoder.
    // (above ~16 rows the activation fragments no longer stay in registers / L1 cheaply and the
    // tcgen05 GEMM wins even though its fixed cost is higher)
    if (a.flags & (LIN_A_MN | LIN_W_MN)) return 0;
    if (a.M < 1 || (a.M > 16 && a.N > 64)) return 0;
    if (a.K % 8 || a.lda % 8) return 0;
    if ((a.flags & LIN_NORM_A) && (!a.norm_w || ((uintptr_t)a.norm_w & 15))) return 0;
    if (a.flags & LIN_COMBINE_A) {
        if (a.M > 16 || a.cmb_heads > 8 || a.cmb_splits > 16 || a.cmb_hd % 8 || (a.flags & LIN_NORM_A)) return 0;
        if (a.K != a.cmb_heads * a.cmb_hd) return 0;
        if ((a.K % 8 || ((uintptr_t)a.W & 15))) return 0;
        return 1;
    }
    if (((uintptr_t)a.A | (uintptr_t)a.W) & 15) return 0;
    if ((a.flags & LIN_GEGLU) && (a.N % (2 * PZ_GU_BLOCK))) return 0;
    if ((a.flags & LIN_ACCUM) && !(a.flags & LIN_OUT_F32)) return 0;
    return 1;
}

int launch_linear_skinny(const LinearArgs &a, cudaStream_t st) {
    SkinnyParams p;
    p.A = (const bf16 *)a.A; p.W = (const bf16 *)a.W; p.bias = a.bias; p.C = a.C; p.norm_w = a.norm_w;
    p.cmb_splits = a.cmb_splits; p.cmb_q_rows = a.cmb_q_rows; p.cmb_heads = a.cmb_heads; p.cmb_hd = a.cmb_hd;
    p.M = a.M; p.N = a.N; p.K = a.K; p.lda = a.lda; p.ldc = a.ldc; p.flags = a.flags; p.alpha = a.alpha;
    bool geglu = a.flags & LIN_GEGLU;
    int nblocks = geglu ? a.N / 16 : (a.N + 15) / 16;
    // split K across CTAs only where the epilogue is a pure fp32 accumulate (atomics)
    int ksplit = 1;
    if ((a.flags & LIN_ACCUM) && !(a.flags & (LIN_GELU | LIN_SILU | LIN_GEGLU))) {
        while (nblocks * ksplit < 444 && a.K / (NWARPS * ksplit * 2) >= 64) ksplit *= 2;
    }
    p.ksplit = ksplit;
    int mt = a.M > 8 * MAX_MT ? MAX_MT : (a.M + 7) / 8;
    p.n_mt = mt;
    dim3 grid(nblocks, ksplit, a.M > 8 * MAX_MT ? (a.M + 8 * MAX_MT - 1) / (8 * MAX_MT) : 1);
    switch (mt) {
        case 1: launch_k(skinny_kernel<1>, dim3(grid), dim3(NTHREADS), 0, st, p); break;
        case 2: launch_k(skinny_kernel<2>, dim3(grid), dim3(NTHREADS), 0, st, p); break;
        case 3: case 4: launch_k(skinny_kernel<4>, dim3(grid), dim3(NTHREADS), 0, st, p); break;
        default: launch_k(skinny_kernel<8>, dim3(grid), dim3(NTHREADS), 0, st, p); break;
    }
    return 0;
}
