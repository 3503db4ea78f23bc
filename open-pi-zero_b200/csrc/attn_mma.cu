// attn_mma.cu -- fused block-masked attention on tensor cores (bf16 in, fp32
// softmax and accumulation), flash-style: the score matrix never touches HBM.
//
// Replaces the reference's materialised attention
//   joint_model.py:243-282  (repeat_kv, cat, QK^T/sqrt(d), tanh soft-cap 50, +mask, fp32 softmax, PV)
//   siglip.py:133-152       (QK^T * d^-0.5, fp32 softmax, PV; no mask)
// The block mask of pizero.py:271-310 is applied in registers from the per-sample
// valid length; the dense [B,1,S,S] mask is never read.  MQA
// (one KV head) is folded: (head, token) pairs are just more query rows against
// the same K/V tile (SURVEY.md 8a note 6), so each CTA streams the cached K/V once.
//
// CTA = 4 warps x 16 query rows; keys in tiles of 64; cp.async staging, ldmatrix
// fragments, mma.sync.m16n8k16.  (<1 % of the path's FLOPs: 11.3 of 1268 GFLOP.)
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int KEY_TILE = 64;

PZ_DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
PZ_DEVINL void cp_async16(void *dst, const void *src, bool valid) {
    int sz = valid ? 16 : 0;   // src-size 0 => destination is zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz)
                 : "memory");
}
PZ_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PZ_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
PZ_DEVINL void ldsm_x4(uint32_t (&r)[4], const void *p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
PZ_DEVINL void ldsm_x4_t(uint32_t (&r)[4], const void *p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
PZ_DEVINL void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tanh(y) = 1 - 2 / (exp(2y) + 1) with ex2.approx / fast division: abs. error ~1e-7, i.e. < 1e-5 on
// the soft-capped logit (cap 50) -- the libm tanhf costs more issue slots than the MMAs of a tile
PZ_DEVINL float tanh_fast_acc(float y) {
    float t = __expf(2.f * y);
    return 1.f - __fdividef(2.f, t + 1.f);
}

// visibility classes of a launch (all of its query rows share one)
enum { CLS_ALL = 0, CLS_VLM = 1, CLS_PROPRIO = 2, CLS_ACTION = 3 };

// PRELOAD: all keys fit in shared memory (SigLIP: 256 keys x head_dim 72) -- K and V are staged once and
// the tile loop runs without loads or block barriers.
// NW warps x 16 query rows per CTA.  The preload variant takes 16 warps (256 rows) when a (sample, head) has that
// many rows (SigLIP: 256 patches): K and V are then read once per (sample, head) instead of once per 64 rows.
template <int HD, bool PRELOAD = false, int NW = 4>   // true head_dim; HDP = padded to a multiple of 16
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 2 : 1) attn_mma_kernel(AttnArgs a, int cls, int mqa, int split) {
    constexpr int ROWS_PER_CTA = NW * 16, NTHREADS = NW * 32;
    constexpr int HDP = (HD + 15) / 16 * 16;
    constexpr int LDS = HDP + 8;            // +16 B per row: conflict-free ldmatrix
    constexpr int CHUNKS = HD / 8;          // 16-byte chunks of real data per row
    constexpr int PCHUNKS = HDP / 8;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16 *sQ = (bf16 *)smem_raw;
    bf16 *sK0 = sQ + ROWS_PER_CTA * LDS;
    const int kv_rows = PRELOAD ? (a.s_cache + a.n_fresh + KEY_TILE - 1) / KEY_TILE * KEY_TILE : KEY_TILE;
    bf16 *sV0 = sK0 + kv_rows * LDS;
    bf16 *sK = sK0, *sV = sV0;

    pdl_trigger();
    pdl_wait();     // Q/K/V are produced by the preceding kernels
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.z;
    const int head_y = blockIdx.y;                       // only used when !mqa
    const int rows_total = mqa ? a.n_heads * a.q_rows : a.q_rows;
    const int row0 = blockIdx.x * ROWS_PER_CTA;
    const int vlen = a.valid_len ? a.valid_len[b] : a.s_cache;

    int n_keys;   // keys [0, n_keys) are iterated; the pad gap [vlen, s_vlm) is masked
    if (cls == CLS_ALL) n_keys = a.s_cache;
    else if (cls == CLS_VLM) n_keys = vlen;
    else if (cls == CLS_PROPRIO) n_keys = a.s_cache;
    else n_keys = a.s_cache + a.n_fresh;
    const int n_tiles_all = (n_keys + KEY_TILE - 1) / KEY_TILE;
    // split-key mode (decode): this CTA owns exactly one key tile and writes an
    // unnormalised partial (o, m, l); a second kernel combines the partials
    const int tile_lo = split ? blockIdx.y : 0;
    const int n_tiles = split ? min(n_tiles_all, tile_lo + 1) : n_tiles_all;

    const bf16 *Qb = (const bf16 *)a.Q + b * a.q_batch_stride;
    const bf16 *Kb = (const bf16 *)a.K + b * a.kv_batch_stride + (mqa ? 0 : head_y * a.kv_head_stride);
    const bf16 *Vb = (const bf16 *)a.V + b * a.kv_batch_stride + (mqa ? 0 : head_y * a.kv_head_stride);
    const bf16 *K2b = a.K2 ? (const bf16 *)a.K2 + b * a.kv2_batch_stride : nullptr;
    const bf16 *V2b = a.V2 ? (const bf16 *)a.V2 + b * a.kv2_batch_stride : nullptr;

    // ---- stage Q (once) and the first K tile -----------------------------
    const bool rope = (HD == 256) && a.rope_cos != nullptr;
    // half-split rotation (model/utils.py:4-16) of one 8-element chunk pair (d, d+128)
    auto rope_pair = [&](bf16 *dst_row, const bf16 *src_row, int c, int pos) {
        uint4 r1 = __ldg(reinterpret_cast<const uint4 *>(src_row + c * 8));
        uint4 r2 = __ldg(reinterpret_cast<const uint4 *>(src_row + 128 + c * 8));
        const float4 *cs = reinterpret_cast<const float4 *>(a.rope_cos + (long)pos * 128 + c * 8);
        const float4 *sn = reinterpret_cast<const float4 *>(a.rope_sin + (long)pos * 128 + c * 8);
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
    };
    auto load_kv = [&](bf16 *dst, const bf16 *base, const bf16 *base2, int tile) {
        const bool rope_k = rope && dst == sK0;
        for (int i = tid; i < KEY_TILE * PCHUNKS; i += NTHREADS) {
            int r = i / PCHUNKS, c = i % PCHUNKS;
            int j = tile * KEY_TILE + r;
            bool ok = j < n_keys && c < CHUNKS;
            if (rope_k && ok && j >= a.s_cache) {
                // fresh (action) key: raw projection, rotate while staging
                if (c < 16) rope_pair(dst + r * LDS, base2 + (long)(j - a.s_cache) * a.kv2_row_stride, c,
                                      a.rope_pos0 + (j - a.s_cache));
                continue;
            }
            const bf16 *src = base;
            if (ok) src = (j < a.s_cache) ? base + (long)j * a.kv_row_stride + c * 8
                                          : base2 + (long)(j - a.s_cache) * a.kv2_row_stride + c * 8;
            cp_async16(dst + r * LDS + c * 8, src, ok);
        }
    };
    // K and V of the first tile first (their latency is the critical path); the Q staging
    // work then overlaps with them
    if (PRELOAD) {
        for (int tl = 0; tl < n_tiles; ++tl) {
            load_kv(sK0 + tl * KEY_TILE * LDS, Kb, K2b, tl);
            load_kv(sV0 + tl * KEY_TILE * LDS, Vb, V2b, tl);
        }
        cp_async_commit();
        cp_async_commit();
    } else {
        load_kv(sK, Kb, K2b, tile_lo);
        cp_async_commit();
        load_kv(sV, Vb, V2b, tile_lo);
        cp_async_commit();
    }
    if (rope) {
        for (int i = tid; i < ROWS_PER_CTA * 16; i += NTHREADS) {
            int r = i / 16, c = i % 16;
            int row = row0 + r;
            bf16 *dst = sQ + r * LDS;
            if (row < rows_total) {
                int h = mqa ? row / a.q_rows : head_y;
                int tok = mqa ? row % a.q_rows : row;
                rope_pair(dst, Qb + (long)tok * a.q_row_stride + h * a.q_head_stride, c, a.rope_pos0 + tok);
            } else {
                *reinterpret_cast<uint4 *>(dst + c * 8) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(dst + 128 + c * 8) = make_uint4(0, 0, 0, 0);
            }
        }
    } else {
        for (int i = tid; i < ROWS_PER_CTA * PCHUNKS; i += NTHREADS) {
            int r = i / PCHUNKS, c = i % PCHUNKS;
            int row = row0 + r;
            bool ok = row < rows_total && c < CHUNKS;
            int h = mqa ? row / a.q_rows : head_y;
            int tok = mqa ? row % a.q_rows : row;
            const bf16 *src = ok ? Qb + (long)tok * a.q_row_stride + h * a.q_head_stride + c * 8 : Qb;
            cp_async16(sQ + r * LDS + c * 8, src, ok);
        }
    }
    cp_async_commit();
    float o[HDP / 8][4];
#pragma unroll
    for (int i = 0; i < HDP / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const float scale = a.scale, cap = a.softcap, inv_cap = cap > 0.f ? 1.f / cap : 0.f;

    if (PRELOAD) { cp_async_wait<0>(); __syncthreads(); }
    for (int tile = tile_lo; tile < n_tiles; ++tile) {
        if (PRELOAD) {
            sK = sK0 + tile * KEY_TILE * LDS;
            sV = sV0 + tile * KEY_TILE * LDS;
        } else {
            if (tile > tile_lo) load_kv(sV, Vb, V2b, tile);   // the first tile's V is already in flight
            cp_async_commit();
            cp_async_wait<1>();          // Q and K(tile) have landed
            __syncthreads();
        }

        // ---- S = Q K^T for this warp's 16 rows x 64 keys ------------------
        float s[KEY_TILE / 8][4];
#pragma unroll
        for (int i = 0; i < KEY_TILE / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < HDP / 16; ++ks) {
            uint32_t qa[4];
            ldsm_x4(qa, sQ + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int nt2 = 0; nt2 < KEY_TILE / 16; ++nt2) {
                // 16 keys x 16 d: matrices (keys 0-7,d 0-7), (keys 0-7,d 8-15), (keys 8-15,d 0-7), (keys 8-15,d 8-15)
                uint32_t kb[4];
                ldsm_x4(kb, sK + (nt2 * 16 + (lane & 7) + ((lane >> 4) << 3)) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
                mma_bf16(s[2 * nt2], qa, kb[0], kb[1]);
                mma_bf16(s[2 * nt2 + 1], qa, kb[2], kb[3]);
            }
        }
        if (!PRELOAD) {
            __syncthreads();             // every warp is done reading sK
            if (tile + 1 < n_tiles) load_kv(sK, Kb, K2b, tile + 1);
            cp_async_commit();
        }

        // ---- scale, soft-cap, mask, online softmax --------------------------
        float mx[2] = {-INFINITY, -INFINITY};
        // only a tile that straddles a visibility boundary needs the per-element mask (warp-uniform test)
        const int j_lo = tile * KEY_TILE, j_hi = j_lo + KEY_TILE;
        bool need_mask;
        if (cls == CLS_ALL) need_mask = j_hi > n_keys;
        else if (cls == CLS_VLM) need_mask = j_hi > vlen;
        else need_mask = j_hi > vlen;                       // pad gap and the tail behind n_keys
        if (!need_mask) {
#pragma unroll
            for (int nt = 0; nt < KEY_TILE / 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float v = s[nt][e] * scale;
                    if (cap > 0.f) v = tanh_fast_acc(v * inv_cap) * cap;
                    s[nt][e] = v;
                    mx[e >> 1] = fmaxf(mx[e >> 1], v);
                }
            }
        } else {
#pragma unroll
            for (int nt = 0; nt < KEY_TILE / 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int j = j_lo + nt * 8 + 2 * t + (e & 1);
                    bool vis;
                    if (cls == CLS_ALL) vis = j < n_keys;
                    else if (cls == CLS_VLM) vis = j < vlen;
                    else vis = (j < vlen) || (j >= a.s_vlm && j < n_keys);
                    float v = s[nt][e] * scale;
                    if (cap > 0.f) v = tanh_fast_acc(v * inv_cap) * cap;
                    v = vis ? v : -INFINITY;
                    s[nt][e] = v;
                    mx[e >> 1] = fmaxf(mx[e >> 1], v);
                }
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            float m_new = fmaxf(m_run[r], mx[r]);
            float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
            corr[r] = __expf(m_run[r] - m_safe);     // m_run = -inf -> 0
            m_run[r] = m_new;
            mx[r] = m_safe;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pa[KEY_TILE / 16][4];
#pragma unroll
        for (int nt = 0; nt < KEY_TILE / 8; ++nt) {
            float p0 = __expf(s[nt][0] - mx[0]), p1 = __expf(s[nt][1] - mx[0]);
            float p2 = __expf(s[nt][2] - mx[1]), p3 = __expf(s[nt][3] - mx[1]);
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            // C fragments of two adjacent 8-key tiles form one 16-key A fragment
            pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
        for (int i = 0; i < HDP / 8; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0];
            o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }

        if (!PRELOAD) {
            cp_async_wait<1>();          // V(tile) has landed (K(tile+1) may still be in flight)
            __syncthreads();
        }
        // ---- O += P V -------------------------------------------------------
#pragma unroll
        for (int kk = 0; kk < KEY_TILE / 16; ++kk) {
#pragma unroll
            for (int dt2 = 0; dt2 < HDP / 16; ++dt2) {
                // 16 keys x 16 d, transposed on load: (keys 0-7,d 0-7), (keys 8-15,d 0-7), (keys 0-7,d 8-15), (keys 8-15,d 8-15)
                uint32_t vb[4];
                ldsm_x4_t(vb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dt2 * 16 + (lane >> 4) * 8);
                mma_bf16(o[2 * dt2], pa[kk], vb[0], vb[1]);
                mma_bf16(o[2 * dt2 + 1], pa[kk], vb[2], vb[3]);
            }
        }
        if (!PRELOAD) __syncthreads();   // every warp is done reading sV
    }
    cp_async_wait<0>();

    // ---- normalise and store ------------------------------------------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        l_run[r] = l;
    }
    if (split) {
        // partial layout: [b][tile][row][HD + 2] = o[HD] (unnormalised), m, l
        float *base = a.scratch + (((long)b * gridDim.y + blockIdx.y) * rows_total) * (HD + 2);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int row = row0 + warp * 16 + g + r * 8;
            if (row >= rows_total) continue;
            float *dst = base + (long)row * (HD + 2);
#pragma unroll
            for (int i = 0; i < HDP / 8; ++i) {
                int d = i * 8 + 2 * t;
                if (d < HD) *reinterpret_cast<float2 *>(dst + d) = make_float2(o[i][2 * r], o[i][2 * r + 1]);
            }
            if (t == 0) { dst[HD] = m_run[r]; dst[HD + 1] = l_run[r]; }
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
    bf16 *Ob = (bf16 *)a.O + b * a.o_batch_stride;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int row = row0 + warp * 16 + g + r * 8;
        if (row >= rows_total) continue;
        int h = mqa ? row / a.q_rows : head_y;
        int tok = mqa ? row % a.q_rows : row;
        bf16 *dst = Ob + (long)tok * a.o_row_stride + h * a.o_head_stride;
        // pad rows (token >= valid_len) are never read by a valid row: write zeros
        const float nrm = (cls == CLS_VLM && a.q_row0 + tok >= vlen) ? 0.f : l_run[r];
#pragma unroll
        for (int i = 0; i < HDP / 8; ++i) {
            int d = i * 8 + 2 * t;
            if (d < HD)
                *reinterpret_cast<uint32_t *>(dst + d) = pack_bf16x2(o[i][2 * r] * nrm, o[i][2 * r + 1] * nrm);
        }
    }
}

// combine the split-key partials: one warp per (sample, query row)
template <int HD>
__global__ void __launch_bounds__(128) attn_combine_kernel(AttnArgs a, int n_splits) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int rows_total = a.n_heads * a.q_rows;
    long wid = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (wid >= (long)a.batch * rows_total) return;
    int b = wid / rows_total, row = wid % rows_total;
    const float *base = a.scratch + ((long)b * n_splits * rows_total + row) * (HD + 2);
    const long sstride = (long)rows_total * (HD + 2);
    float mx = -INFINITY;
    for (int s = 0; s < n_splits; ++s) mx = fmaxf(mx, base[s * sstride + HD]);
    float acc[HD / 32];
#pragma unroll
    for (int i = 0; i < HD / 32; ++i) acc[i] = 0.f;
    float l = 0.f;
    for (int s = 0; s < n_splits; ++s) {
        const float *p = base + s * sstride;
        float m = p[HD];
        float w = (m == -INFINITY) ? 0.f : __expf(m - mx);
        l += p[HD + 1] * w;
#pragma unroll
        for (int i = 0; i < HD / 32; ++i) acc[i] += p[i * 32 + lane] * w;
    }
    float inv = l > 0.f ? 1.f / l : 0.f;
    int h = row / a.q_rows, tok = row % a.q_rows;
    bf16 *dst = (bf16 *)a.O + b * a.o_batch_stride + (long)tok * a.o_row_stride + h * a.o_head_stride;
#pragma unroll
    for (int i = 0; i < HD / 32; ++i) dst[i * 32 + lane] = __float2bfloat16_rn(acc[i] * inv);
}

template <int HD>
int launch(const AttnArgs &a, int cls, int mqa, cudaStream_t st, int partials_only = 0) {
    constexpr int HDP = (HD + 15) / 16 * 16;
    constexpr int LDS = HDP + 8;
    constexpr int ROWS_PER_CTA = 64, NTHREADS = 128;   // default: 4 warps
    size_t smem = (size_t)(ROWS_PER_CTA + 2 * KEY_TILE) * LDS * sizeof(bf16);
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        if (cudaFuncSetAttribute(attn_mma_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return PZ_ERR_CUDA;
    }
    int rows_total = mqa ? a.n_heads * a.q_rows : a.q_rows;
    if constexpr (HD % 32 == 0) {
        // decode (action / proprio rows, MQA): too few query rows to fill the machine by rows, so
        // split the keys across CTAs (flash-decoding) and combine
        int n_keys = a.s_cache + (cls == CLS_ACTION ? a.n_fresh : 0);
        int n_splits = (n_keys + KEY_TILE - 1) / KEY_TILE;
        size_t need = (size_t)a.batch * n_splits * rows_total * (HD + 2) * sizeof(float);
        if (mqa && (cls == CLS_ACTION || cls == CLS_PROPRIO) && rows_total <= ROWS_PER_CTA && n_splits > 1 &&
            a.scratch && a.scratch_bytes >= need) {
            dim3 grid(1, n_splits, a.batch);
            launch_k(attn_mma_kernel<HD>, dim3(grid), dim3(NTHREADS), smem, st, a, cls, mqa, 1);
            if (partials_only) return n_splits;   // the consumer (o_proj GEMV) combines
            long warps = (long)a.batch * rows_total;
            launch_k(attn_combine_kernel<HD>, dim3((unsigned)((warps + 3) / 4)), dim3(128), 0, st, a, n_splits);
            return 0;
        }
    }
    if (partials_only) return 0;
    dim3 grid((rows_total + ROWS_PER_CTA - 1) / ROWS_PER_CTA, mqa ? 1 : a.n_heads, a.batch);
    if constexpr (HD <= 128) {
        int kv_rows = (a.s_cache + a.n_fresh + KEY_TILE - 1) / KEY_TILE * KEY_TILE;
        // (only when that still leaves a few CTAs per SM: at bs=1 the 64-row CTAs spread better)
        if (cls == CLS_ALL && rows_total % 256 == 0 && (long)(rows_total / 256) * (mqa ? 1 : a.n_heads) * a.batch >= 296) {
            // all rows of a (sample, head) in one 16-warp CTA
            size_t smem_big = (size_t)(256 + 2 * kv_rows) * LDS * sizeof(bf16);
            if (smem_big <= 200 * 1024) {
                static bool attr3 = false;
                if (!attr3) {
                    if (cudaFuncSetAttribute(attn_mma_kernel<HD, true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
                        return PZ_ERR_CUDA;
                    attr3 = true;
                }
                dim3 grid16(rows_total / 256, mqa ? 1 : a.n_heads, a.batch);
                launch_k(attn_mma_kernel<HD, true, 16>, dim3(grid16), dim3(512), smem_big, st, a, cls, mqa, 0);
                return 0;
            }
        }
        size_t smem_all = (size_t)(ROWS_PER_CTA + 2 * kv_rows) * LDS * sizeof(bf16);
        if (cls == CLS_ALL && smem_all <= 110 * 1024) {
            static bool attr2 = false;
            if (!attr2) {
                if (cudaFuncSetAttribute(attn_mma_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) != cudaSuccess)
                    return PZ_ERR_CUDA;
                attr2 = true;
            }
            launch_k(attn_mma_kernel<HD, true>, dim3(grid), dim3(NTHREADS), smem_all, st, a, cls, mqa, 0);
            return 0;
        }
    }
    if constexpr (HD == 256) {
        // 128 query rows per CTA (K/V tiles read half as often) once the batch fills the machine with such CTAs
        // (59.4 -> 58.9 ms on the bs=64 prefix pass); PZ_ATTN_NW=4 forces the 64-row CTAs
        static const int nw = [] { const char *e = getenv("PZ_ATTN_NW"); return e ? atoi(e) : 8; }();
        if (nw == 8 && rows_total >= 1024 && a.batch >= 16) {
            size_t smem8 = (size_t)(128 + 2 * KEY_TILE) * LDS * sizeof(bf16);
            static bool attr8 = false;
            if (!attr8) {
                if (cudaFuncSetAttribute(attn_mma_kernel<HD, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8) != cudaSuccess)
                    return PZ_ERR_CUDA;
                attr8 = true;
            }
            dim3 grid8((rows_total + 127) / 128, mqa ? 1 : a.n_heads, a.batch);
            launch_k(attn_mma_kernel<HD, false, 8>, dim3(grid8), dim3(256), smem8, st, a, cls, mqa, 0);
            return 0;
        }
    }
    launch_k(attn_mma_kernel<HD>, dim3(grid), dim3(NTHREADS), smem, st, a, cls, mqa, 0);
    return 0;
}

int classify(const AttnArgs &a) {
    if (!a.valid_len) return (a.n_fresh == 0 && a.s_vlm == a.s_cache) ? CLS_ALL : -1;
    int lo = a.q_row0, hi = a.q_row0 + a.q_rows;
    if (hi <= a.s_vlm) return a.n_fresh == 0 ? CLS_VLM : -1;
    if (lo >= a.s_vlm && hi <= a.s_cache) return a.n_fresh == 0 ? CLS_PROPRIO : -1;
    if (lo >= a.s_cache) return CLS_ACTION;
    return -1;
}

}  // namespace

int attn_mma_supported(const AttnArgs &a) {
    if (a.head_dim != 72 && a.head_dim != 256) return 0;
    if (classify(a) < 0) return 0;
    // 16-byte cp.async granularity
    if (a.q_row_stride % 8 || a.q_head_stride % 8 || a.kv_row_stride % 8 || a.kv_head_stride % 8) return 0;
    if (a.q_batch_stride % 8 || a.kv_batch_stride % 8 || a.o_row_stride % 2 || a.o_head_stride % 2) return 0;
    if (a.n_fresh && (a.kv2_row_stride % 8 || a.kv2_batch_stride % 8 || !a.K2 || !a.V2)) return 0;
    if (((uintptr_t)a.Q | (uintptr_t)a.K | (uintptr_t)a.V | (uintptr_t)a.K2 | (uintptr_t)a.V2) & 15) return 0;
    if ((uintptr_t)a.O & 3) return 0;
    return 1;
}

int launch_attn_mma(const AttnArgs &a, cudaStream_t st) {
    int cls = classify(a);
    int mqa = a.kv_head_stride == 0;
    if (a.head_dim == 72) return launch<72>(a, cls, mqa, st);
    return launch<256>(a, cls, mqa, st);
}

int launch_attn_mma_partials(const AttnArgs &a, cudaStream_t st) {
    if (a.head_dim != 256 || a.kv_head_stride != 0) return 0;
    int cls = classify(a);
    if (cls < 0) return 0;
    int r = launch<256>(a, cls, 1, st, 1);
    return r > 0 ? r : 0;
}

int launch_attn_combine(const AttnArgs &a, int n_splits, cudaStream_t st) {
    long warps = (long)a.batch * a.n_heads * a.q_rows;
    launch_k(attn_combine_kernel<256>, dim3((unsigned)((warps + 3) / 4)), dim3(128), 0, st, a, n_splits);
    return 0;
}
