// api.cu -- the C ABI (include/pz_b200.h) and the orchestration of
// PiZero.infer_action (reference: src/model/vla/pizero.py:416-490,
// src/model/vla/joint_model.py:24-383, src/model/paligemma/siglip.py).
//
// Everything here is host code that enqueues kernels on the caller's stream;
// no allocation, no synchronisation, no hidden state besides the handle.
#include "api_internal.cuh"

thread_local LaunchCounter *g_launch_counter = nullptr;
int g_pdl_enabled = [] { const char *e = getenv("PZ_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
thread_local std::string g_create_error;
thread_local int g_pdl_off = 0;

// ------------------------------------------------------------ workspace ---
struct Workspace {
    // persistent between the stages
    void *kcache, *vcache;   // [L][B][S_c][hd] T
    float *x;                // [B][S_v][H] fp32: prefix residual stream
    // SigLIP scratch (per chunk)
    void *patches, *hv, *qkvv, *av, *mv;
    float *xv, *feats;
    // prefix scratch (per chunk)
    void *h, *qkv, *q, *att, *mlp;
    void *pp, *hp, *qkvp, *qp, *attp, *mlpp;
    float *xp;
    // denoise scratch (whole batch)
    void *a_in, *e1, *z, *ha, *qkva, *qa, *ka, *va, *atta, *mlpa;
    float *act, *xa, *vel, *att_scratch;
    unsigned int *mega_barrier;
    void *mega_ll;
    size_t mega_ll_bytes;
    size_t att_scratch_bytes;
    size_t total;
};

static Workspace carve(const pz_config &c, int B, int chunk, void *base) {
    Workspace w;
    Bump b;
    b.base = (char *)base;
    size_t es = c.dtype == PZ_BF16 ? 2 : 4;
    int Bc = B < chunk ? B : chunk;
    size_t S_c = c.s_vlm + c.cond_steps;
    size_t qkvd = (size_t)(c.n_heads + 2 * c.n_kv_heads) * c.head_dim;
    size_t qd = (size_t)c.n_heads * c.head_dim;
    w.kcache = b.take<void>((size_t)c.n_layers * B * S_c * c.head_dim * es);
    w.vcache = b.take<void>((size_t)c.n_layers * B * S_c * c.head_dim * es);
    w.x = b.take<float>((size_t)B * c.s_vlm * c.vlm_hidden * 4);
    size_t Mv = (size_t)Bc * c.n_images * c.n_img_tokens;
    w.patches = b.take<void>(Mv * c.patch_k_pad * es);
    w.xv = b.take<float>(Mv * c.vit_hidden * 4);
    w.hv = b.take<void>(Mv * c.vit_hidden * es);
    w.qkvv = b.take<void>(Mv * 3 * c.vit_hidden * es);
    w.av = b.take<void>(Mv * c.vit_hidden * es);
    w.mv = b.take<void>(Mv * c.vit_inter * es);
    w.feats = b.take<float>(Mv * c.vlm_hidden * 4);
    size_t M = (size_t)Bc * c.s_vlm, Mp = (size_t)Bc * c.cond_steps;
    w.h = b.take<void>(M * c.vlm_hidden * es);
    w.qkv = b.take<void>(M * qkvd * es);
    w.q = b.take<void>(M * qd * es);
    w.att = b.take<void>(M * qd * es);
    w.mlp = b.take<void>(M * c.vlm_inter * es);
    w.pp = b.take<void>(Mp * 64 * es);
    w.xp = b.take<float>(Mp * c.act_hidden * 4);
    w.hp = b.take<void>(Mp * c.act_hidden * es);
    w.qkvp = b.take<void>(Mp * qkvd * es);
    w.qp = b.take<void>(Mp * qd * es);
    w.attp = b.take<void>(Mp * qd * es);
    w.mlpp = b.take<void>(Mp * c.act_inter * es);
    size_t Ma = (size_t)B * c.horizon;
    w.act = b.take<float>(Ma * c.action_dim * 4);
    w.a_in = b.take<void>(Ma * 64 * es);
    w.e1 = b.take<void>(Ma * c.act_hidden * es);
    w.z = b.take<void>(Ma * c.act_hidden * es);
    w.xa = b.take<float>(Ma * c.act_hidden * 4);
    w.ha = b.take<void>(Ma * c.act_hidden * es);
    w.qkva = b.take<void>(Ma * qkvd * es);
    w.qa = b.take<void>(Ma * qd * es);
    w.ka = b.take<void>(Ma * c.head_dim * es);
    w.va = b.take<void>(Ma * c.head_dim * es);
    w.atta = b.take<void>(Ma * qd * es);
    w.mlpa = b.take<void>(Ma * c.act_inter * es);
    w.vel = b.take<float>(Ma * 8 * 4);
    w.mega_barrier = b.take<unsigned int>(128 * sizeof(unsigned int));
    w.mega_ll_bytes = denoise_mega3_ll_bytes(c, B);   // 0 unless the flag-exchange sampler covers this batch
    w.mega_ll = b.take<void>(w.mega_ll_bytes);
    {   // split-key attention partials (decode): [B][key tiles][heads*rows][hd + 2] fp32
        size_t rows = (size_t)c.n_heads * (c.horizon > c.cond_steps ? c.horizon : c.cond_steps);
        size_t tiles = (S_c + c.horizon + 63) / 64;
        w.att_scratch_bytes = rows <= 64 ? (size_t)B * tiles * rows * (c.head_dim + 2) * 4 : 0;
        w.att_scratch = b.take<float>(w.att_scratch_bytes);
    }
    w.total = (b.off + 1023) & ~(size_t)1023;
    return w;
}

// ------------------------------------------------------ stage 1: SigLIP ----
template <typename T>
static int run_embed_prefix(pz_handle *h, const int64_t *ids, const void *pixels, void *wsp, int B,
                            const pz_capture *cap, cudaStream_t st) {
    const pz_config &c = h->cfg;
    const pz_weights &w = h->w;
    Workspace ws = carve(c, B, h->prefix_chunk, wsp);
    const int V = c.vit_hidden, VI = c.vit_inter, H = c.vlm_hidden, P = c.n_img_tokens;
    const int hdv = V / c.vit_heads;
    const size_t img_elems = (size_t)3 * c.image_size * c.image_size;
    for (int b0 = 0; b0 < B; b0 += h->prefix_chunk) {
        int nb = (B - b0 < h->prefix_chunk) ? B - b0 : h->prefix_chunk;
        int n_img = nb * c.n_images;
        int Mv = n_img * P;
        // patch embedding (siglip.py:59-78): conv as GEMM over im2col rows, + bias + position table
        if (h->pixel_format == PZ_PIXELS_U8) {
            const uint8_t *pix = (const uint8_t *)pixels + (size_t)b0 * c.n_images * img_elems;
            launch_im2col_u8<T>(pix, (T *)ws.patches, n_img, c.image_size, c.patch_size, c.patch_k_pad, st);
        } else {
            const T *pix = (const T *)pixels + (size_t)b0 * c.n_images * img_elems;
            launch_im2col<T>(pix, (T *)ws.patches, n_img, c.image_size, c.patch_size, c.patch_k_pad, st);
        }
        launch_bcast_rows(ws.xv, w.pos_emb, Mv, V, P, st);
        PZ_TRY(Ops<T>::linear(h, lin(ws.patches, c.patch_k_pad, w.patch_w, w.patch_b, ws.xv, V, Mv, V,
                                     c.patch_k_pad, LIN_OUT_F32 | LIN_ACCUM), st));
        for (int l = 0; l < c.vit_layers; ++l) {   // siglip.py:220-238
            const pz_vit_layer &L = h->vit[l];
            launch_layernorm<T>(ws.xv, L.ln1_w, L.ln1_b, (T *)ws.hv, Mv, V, 1e-6f, st);
            PZ_TRY(Ops<T>::linear(h, lin(ws.hv, V, L.w_qkv, L.b_qkv, ws.qkvv, 3 * V, Mv, 3 * V, V), st));
            AttnArgs a;
            memset(&a, 0, sizeof(a));
            a.Q = ws.qkvv; a.q_batch_stride = (long)P * 3 * V; a.q_row_stride = 3 * V; a.q_head_stride = hdv;
            a.K = (const T *)ws.qkvv + V; a.V = (const T *)ws.qkvv + 2 * V;
            a.kv_batch_stride = (long)P * 3 * V; a.kv_row_stride = 3 * V; a.kv_head_stride = hdv;
            a.O = ws.av; a.o_batch_stride = (long)P * V; a.o_row_stride = V; a.o_head_stride = hdv;
            a.batch = n_img; a.n_heads = c.vit_heads; a.head_dim = hdv; a.q_rows = P; a.q_row0 = 0;
            a.s_cache = P; a.s_vlm = P; a.n_fresh = 0;
            a.scale = 1.0f / sqrtf((float)hdv); a.softcap = 0.f;
            PZ_TRY(Ops<T>::attention(h, a, st));
            PZ_TRY(Ops<T>::linear(h, lin(ws.av, V, L.w_o, L.b_o, ws.xv, V, Mv, V, V,
                                         LIN_OUT_F32 | LIN_ACCUM), st));
            launch_layernorm<T>(ws.xv, L.ln2_w, L.ln2_b, (T *)ws.hv, Mv, V, 1e-6f, st);
            PZ_TRY(Ops<T>::linear(h, lin(ws.hv, V, L.w_fc1, L.b_fc1, ws.mv, VI, Mv, VI, V, LIN_GELU), st));
            PZ_TRY(Ops<T>::linear(h, lin(ws.mv, VI, L.w_fc2, L.b_fc2, ws.xv, V, Mv, V, VI,
                                         LIN_OUT_F32 | LIN_ACCUM), st));
        }
        launch_layernorm<T>(ws.xv, w.post_ln_w, w.post_ln_b, (T *)ws.hv, Mv, V, 1e-6f, st);
        if (cap && cap->vit_out)
            launch_to_f32<T>((const T *)ws.hv, cap->vit_out + (size_t)b0 * c.n_images * P * V,
                             (long)Mv * V, st);
        // projector (siglip.py:28-31)
        PZ_TRY(Ops<T>::linear(h, lin(ws.hv, V, w.proj_w, w.proj_b, ws.feats, H, Mv, H, V, LIN_OUT_F32), st));
        if (cap && cap->image_features)
            copy_f32(cap->image_features + (size_t)b0 * c.n_images * P * H, ws.feats, (size_t)Mv * H, st);
        float *x = ws.x + (size_t)b0 * c.s_vlm * H;
        launch_embed_merge<T>(ids + (size_t)b0 * c.s_vlm, (const T *)w.embed, ws.feats, x, nb, c.s_vlm,
                              H, c.n_images * P, c.image_token_index, c.pad_token_id,
                              sqrtf((float)H), st);
        if (cap && cap->prefix_embeds)
            copy_f32(cap->prefix_embeds + (size_t)b0 * c.s_vlm * H, x, (size_t)nb * c.s_vlm * H, st);
    }
    return 0;
}

// ------------------------------------- one mixture's post-attention half ----
// x += o_proj(att); x += down(gelu(gate(n)) * up(n)), n = rmsnorm(x)
// (joint_model.py:65-127, mixture.py:217-218, paligemma/modules.py:86-95)
template <typename T>
static int post_attention(pz_handle *h, const pz_mix_layer &L, float *x, void *hbuf, void *att,
                          void *mlp, int M, int hidden, int inter, cudaStream_t st, int cmb_splits = 0,
                          int cmb_q_rows = 0, const float *partials = nullptr) {
    const pz_config &c = h->cfg;
    int qd = c.n_heads * c.head_dim;
    if (cmb_splits > 0) {
        LinearArgs o = lin(partials, qd, L.w_o, nullptr, x, hidden, M, hidden, qd,
                           LIN_OUT_F32 | LIN_ACCUM | LIN_COMBINE_A);
        o.cmb_splits = cmb_splits; o.cmb_q_rows = cmb_q_rows; o.cmb_heads = c.n_heads; o.cmb_hd = c.head_dim;
        PZ_TRY(launch_linear_skinny(o, st));
    } else {
        PZ_TRY(Ops<T>::linear(h, lin(att, qd, L.w_o, nullptr, x, hidden, M, hidden, qd,
                                     LIN_OUT_F32 | LIN_ACCUM), st));
    }
    int tag_gu = hidden == c.vlm_hidden ? TAG_VLM_GATE_UP : TAG_ACT_GATE_UP;
    if (M > 16) launch_rmsnorm<T>(x, L.norm_post, (T *)hbuf, M, hidden, 1e-6f, st);
    tick(h, tag_gu, st);
    if (M > 16) {
        PZ_TRY(Ops<T>::linear(h, lin(hbuf, hidden, L.w_gate_up, nullptr, mlp, inter, M, 2 * inter, hidden,
                                     LIN_GEGLU), st));
    } else {
        PZ_TRY(norm_linear<T>(h, x, L.norm_post, hbuf, lin(nullptr, hidden, L.w_gate_up, nullptr, mlp, inter, M,
                                                            2 * inter, hidden, LIN_GEGLU), hidden, st));
    }
    tick(h, tag_gu, st);
    if (hidden == c.vlm_hidden) tick(h, TAG_VLM_DOWN, st);
    PZ_TRY(Ops<T>::linear(h, lin(mlp, inter, L.w_down, nullptr, x, hidden, M, hidden, inter,
                                 LIN_OUT_F32 | LIN_ACCUM), st));
    if (hidden == c.vlm_hidden) tick(h, TAG_VLM_DOWN, st);
    return 0;
}

// ------------------------------------------------ stage 2: prefix pass ------
// Two dependency chains per layer (joint_model.py:24-127 over the mixtures vlm and proprio):
//   VLM chain     : norm -> QKV+RoPE(+cache write) -> attention over keys < cnt -> o_proj -> norm -> MLP
//   proprio chain : the same with the action-expert-shaped weights on 1 token per sample; its
//                   attention additionally reads the VLM keys of the same layer.
// VLM rows never attend to the proprio token (block mask, pizero.py:296-310), so the VLM chain does
// not depend on the proprio chain at all: the latter runs on a forked side stream and only waits,
// per layer, for the VLM K/V of that layer.
static cudaEvent_t sync_event(pz_handle *h, size_t i) {
    while (h->sync_ev.size() <= i) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        h->sync_ev.push_back(e);
    }
    return h->sync_ev[i];
}

template <typename T>
static int run_prefill(pz_handle *h, const int32_t *valid_len, const float *proprio, void *wsp,
                       int B, const pz_capture *cap, cudaStream_t st, const float *ext_x = nullptr,
                       const float *ext_xp = nullptr) {
    const pz_config &c = h->cfg;
    const pz_weights &w = h->w;
    Workspace ws = carve(c, B, h->prefix_chunk, wsp);
    const int H = c.vlm_hidden, A = c.act_hidden, hd = c.head_dim, nh = c.n_heads;
    const int S_v = c.s_vlm, S_p = c.cond_steps, S_c = S_v + S_p;
    const int qd = nh * hd, qkvd = (nh + 2 * c.n_kv_heads) * hd;
    const long kv_bs = (long)S_c * hd;
    if (!h->side) {
        if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess)
            return fail(h, PZ_ERR_CUDA, "cannot create the side stream");
    }
    static const bool no_side = [] { const char *e = getenv("PZ_NO_SIDE"); return e && e[0] == '1'; }();
    cudaStream_t sp = no_side ? st : h->side;   // PZ_NO_SIDE=1: both chains on one stream (measurement only)
    size_t ev = 0;
    if (ext_x) copy_f32(ws.x, ext_x, (size_t)B * S_v * H, st);   // JointModel.forward entry: embeddings given
    cudaEvent_t e_fork = sync_event(h, ev++);
    cudaEventRecord(e_fork, st);
    cudaStreamWaitEvent(sp, e_fork, 0);
    for (int b0 = 0; b0 < B; b0 += h->prefix_chunk) {
        int nb = (B - b0 < h->prefix_chunk) ? B - b0 : h->prefix_chunk;
        int M = nb * S_v, Mp = nb * S_p;
        float *x = ws.x + (size_t)b0 * S_v * H;
        // proprio encoder (pizero.py:436) and the sqrt(hidden) embed scale (joint_model.py:348-355)
        if (ext_xp) {
            copy_f32(ws.xp, ext_xp + (size_t)b0 * S_p * A, (size_t)Mp * A, sp);
        } else {
            launch_cast_pad_affine<T>(proprio + (size_t)b0 * S_p * c.proprio_dim, (T *)ws.pp, Mp, c.proprio_dim,
                                      w.small_k_pad, h->prop_scale, h->prop_shift, h->prop_clip, sp);
            PZ_TRY(Ops<T>::linear(h, lin(ws.pp, w.small_k_pad, w.prop_w, w.prop_b, ws.xp, A, Mp, A,
                                         w.small_k_pad, LIN_OUT_F32, sqrtf((float)A)), sp));
        }
        for (int l = 0; l < c.n_layers; ++l) {
            bool last = l == c.n_layers - 1;
            T *Kc = (T *)ws.kcache + ((size_t)l * B + b0) * kv_bs;
            T *Vc = (T *)ws.vcache + ((size_t)l * B + b0) * kv_bs;
            // ---- VLM chain (main stream): norm -> fused QKV projection + RoPE; K (post-RoPE), V -> cache
            launch_rmsnorm<T>(x, h->vlm[l].norm_in, (T *)ws.h, M, H, 1e-6f, st);
            if (std::is_same<T, bf16>::value && !(c.flags & PZ_FLAG_SIMPLE_KERNELS) && hd == 256 && H % 8 == 0) {
                // one kernel: tcgen05 GEMM with RoPE, Q/K/V split and the cache write in its epilogue
                const char *e = nullptr;
                int rc = launch_qkv_rope_tc(ws.h, H, h->vlm[l].w_qkv, ws.q, Kc, Vc, kv_bs, w.rope_vlm_cos,
                                            w.rope_vlm_sin, M, H, nh, S_v, st, &e);
                if (rc) return fail(h, rc, e ? e : "fused qkv+rope launch failed");
            } else {
                PZ_TRY(Ops<T>::linear(h, lin(ws.h, H, h->vlm[l].w_qkv, nullptr, ws.qkv, qkvd, M, qkvd, H), st));
                launch_rope_split<T>((const T *)ws.qkv, qkvd, (T *)ws.q, (long)S_v * qd, Kc, Vc, kv_bs,
                                     w.rope_vlm_cos, w.rope_vlm_sin, nb, S_v, 0, nh, hd, st);
            }
            cudaEvent_t e_kv = sync_event(h, ev++);
            cudaEventRecord(e_kv, st);
            // ---- proprio chain (side stream)
            PZ_TRY(norm_linear<T>(h, ws.xp, h->proprio[l].norm_in, ws.hp,
                                  lin(nullptr, A, h->proprio[l].w_qkv, nullptr, ws.qkvp, qkvd, Mp, qkvd, A), A, sp));
            launch_rope_split<T>((const T *)ws.qkvp, qkvd, (T *)ws.qp, (long)S_p * qd, Kc + (size_t)S_v * hd,
                                 Vc + (size_t)S_v * hd, kv_bs, w.rope_act_cos, w.rope_act_sin, nb, S_p,
                                 0, nh, hd, sp);
            // last layer: the reference computes attention for vlm/proprio and drops it
            // (joint_model.py:297-299); only the K/V above are kept.
            if (last) break;
            AttnArgs a;
            memset(&a, 0, sizeof(a));
            a.K = Kc; a.V = Vc; a.kv_batch_stride = kv_bs; a.kv_row_stride = hd; a.kv_head_stride = 0;
            a.valid_len = valid_len + b0;
            a.batch = nb; a.n_heads = nh; a.head_dim = hd; a.s_cache = S_c; a.s_vlm = S_v; a.n_fresh = 0;
            a.scale = 1.0f / sqrtf((float)hd); a.softcap = 50.f;   // joint_model.py:139,261-268
            a.q_row_stride = qd; a.q_head_stride = hd; a.o_row_stride = qd; a.o_head_stride = hd;
            AttnArgs av = a;   // image/text rows
            av.Q = ws.q; av.q_batch_stride = (long)S_v * qd; av.O = ws.att; av.o_batch_stride = (long)S_v * qd;
            av.q_rows = S_v; av.q_row0 = 0;
            PZ_TRY(Ops<T>::attention(h, av, st));
            PZ_TRY(post_attention<T>(h, h->vlm[l], x, ws.h, ws.att, ws.mlp, M, H, c.vlm_inter, st));
            if (cap && cap->prefix_vlm)
                copy_f32(cap->prefix_vlm + ((size_t)l * B + b0) * S_v * H, x, (size_t)M * H, st);
            // proprio rows: need this layer's VLM keys
            cudaStreamWaitEvent(sp, e_kv, 0);
            a.scratch = ws.att_scratch; a.scratch_bytes = ws.att_scratch_bytes;
            AttnArgs ap = a;
            ap.Q = ws.qp; ap.q_batch_stride = (long)S_p * qd; ap.O = ws.attp; ap.o_batch_stride = (long)S_p * qd;
            ap.q_rows = S_p; ap.q_row0 = S_v;
            PZ_TRY(Ops<T>::attention(h, ap, sp));
            PZ_TRY(post_attention<T>(h, h->proprio[l], ws.xp, ws.hp, ws.attp, ws.mlpp, Mp, A, c.act_inter, sp));
            if (cap && cap->prefix_proprio)
                copy_f32(cap->prefix_proprio + ((size_t)l * B + b0) * S_p * A, ws.xp, (size_t)Mp * A, sp);
        }
    }
    cudaEvent_t e_join = sync_event(h, ev++);
    cudaEventRecord(e_join, sp);
    cudaStreamWaitEvent(st, e_join, 0);
    return 0;
}

// ---------------------------------- the action expert's 18 layers over the cached prefix KV ----
// (JointModel.forward with only "action" active, cache_mode="append_non_active":
//  joint_model.py:164-168,358-373; keys = [cached vlm | cached proprio | fresh action])
template <typename T>
static int action_layers(pz_handle *h, Workspace &ws, const int32_t *valid_len, int B, float *cap_layers,
                         cudaStream_t st) {
    const pz_config &c = h->cfg;
    const pz_weights &w = h->w;
    const int A = c.act_hidden, hd = c.head_dim, nh = c.n_heads, Hz = c.horizon;
    const int S_v = c.s_vlm, S_p = c.cond_steps, S_c = S_v + S_p;
    const int qd = nh * hd, qkvd = (nh + 2 * c.n_kv_heads) * hd;
    const long kv_bs = (long)S_c * hd;
    const int Ma = B * Hz;
        for (int l = 0; l < c.n_layers; ++l) {
        const pz_mix_layer &L = h->action[l];
        PZ_TRY(norm_linear<T>(h, ws.xa, L.norm_in, ws.ha, lin(nullptr, A, L.w_qkv, nullptr, ws.qkva, qkvd, Ma, qkvd, A),
                              A, st));
        AttnArgs a;
        memset(&a, 0, sizeof(a));
        a.batch = B; a.n_heads = nh; a.head_dim = hd; a.q_rows = Hz; a.q_row0 = S_c;
        a.s_cache = S_c; a.s_vlm = S_v; a.n_fresh = Hz;
        a.K = (T *)ws.kcache + (size_t)l * B * kv_bs; a.V = (T *)ws.vcache + (size_t)l * B * kv_bs;
        a.kv_batch_stride = kv_bs; a.kv_row_stride = hd; a.kv_head_stride = 0;
        a.Q = ws.qkva; a.q_batch_stride = (long)Hz * qkvd; a.q_row_stride = qkvd; a.q_head_stride = hd;
        a.K2 = (T *)ws.qkva + qd; a.V2 = (T *)ws.qkva + qd + hd;
        a.kv2_batch_stride = (long)Hz * qkvd; a.kv2_row_stride = qkvd;
        a.rope_cos = w.rope_act_cos; a.rope_sin = w.rope_act_sin; a.rope_pos0 = S_p;
        a.valid_len = valid_len;
        bool fused_rope = std::is_same<T, bf16>::value && !(c.flags & PZ_FLAG_SIMPLE_KERNELS) && hd == 256 &&
                          attn_mma_supported(a);
        if (!fused_rope) {
            // separate RoPE + Q/K/V split kernel, attention on the rotated copies
            launch_rope_split<T>((const T *)ws.qkva, qkvd, (T *)ws.qa, (long)Hz * qd, (T *)ws.ka,
                                 (T *)ws.va, (long)Hz * hd, w.rope_act_cos, w.rope_act_sin, B, Hz, S_p,
                                 nh, hd, st);
            a.Q = ws.qa; a.q_batch_stride = (long)Hz * qd; a.q_row_stride = qd;
            a.K2 = ws.ka; a.V2 = ws.va; a.kv2_batch_stride = (long)Hz * hd; a.kv2_row_stride = hd;
            a.rope_cos = a.rope_sin = nullptr;
        }
        a.O = ws.atta; a.o_batch_stride = (long)Hz * qd; a.o_row_stride = qd; a.o_head_stride = hd;
        a.scale = 1.0f / sqrtf((float)hd); a.softcap = 50.f;
        a.scratch = ws.att_scratch; a.scratch_bytes = ws.att_scratch_bytes;
        int n_splits = 0;
        static const bool combine_in_oproj = getenv("PZ_COMBINE_IN_OPROJ") != nullptr;   // experimental
        if (fused_rope && Ma <= 16 && combine_in_oproj) {
            // small batch: leave the split-key partials uncombined; the o_proj GEMV combines them
            // while it loads its activations (one kernel less on the latency-bound chain)
            LinearArgs probe = lin(ws.att_scratch, qd, L.w_o, nullptr, ws.xa, A, Ma, A, qd,
                                   LIN_OUT_F32 | LIN_ACCUM | LIN_COMBINE_A);
            probe.cmb_splits = (S_c + Hz + 63) / 64; probe.cmb_q_rows = Hz; probe.cmb_heads = nh; probe.cmb_hd = hd;
            if (skinny_supported(probe)) n_splits = launch_attn_mma_partials(a, st);
        }
        bool done_attn = false;
        if (n_splits == 0 && fused_rope && decode_attention2_supported(c)) {
            // one CTA per sample: K and V read once for all heads, RoPE fused, no partials, no combine kernel
            int rc = launch_decode_attention2(c, w, ws.qkva, ws.kcache, ws.vcache, B, valid_len, l, B, ws.atta, (long)Hz * qd, qd, st);
            if (rc) return fail(h, rc, "decode attention launch failed");
            done_attn = true;
        }
        if (done_attn) {
        } else if (n_splits == 0 && fused_rope && decode_attention_supported(c) &&
            a.scratch_bytes >= (size_t)B * ((S_c + Hz + 63) / 64) * nh * Hz * (hd + 2) * sizeof(float)) {
            // one CTA per (sample, key tile), all 8 warps busy, RoPE fused; then the combine kernel
            int ns = launch_decode_attention(c, w, ws.qkva, ws.kcache, ws.vcache, B, valid_len, ws.att_scratch, l, B,
                                             ws.atta, (long)Hz * qd, qd, st);
            if (ns < 0) return fail(h, ns, "decode attention launch failed");
        } else if (n_splits == 0) {
            PZ_TRY(Ops<T>::attention(h, a, st));
        }
        PZ_TRY(post_attention<T>(h, L, ws.xa, ws.ha, ws.atta, ws.mlpa, Ma, A, c.act_inter, st, n_splits, Hz,
                                 ws.att_scratch));
        if (cap_layers) copy_f32(cap_layers + (size_t)l * Ma * A, ws.xa, (size_t)Ma * A, st);
    }
    return 0;
}

// ---------------------------------------------- stage 3: Euler sampler -------
template <typename T>
static int run_denoise(pz_handle *h, const int32_t *valid_len, const float *noise, float *out,
                       void *wsp, int B, const pz_capture *cap, cudaStream_t st) {
    const pz_config &c = h->cfg;
    const pz_weights &w = h->w;
    Workspace ws = carve(c, B, h->prefix_chunk, wsp);
    const int A = c.act_hidden, hd = c.head_dim, nh = c.n_heads, Hz = c.horizon;
    const int S_v = c.s_vlm, S_p = c.cond_steps, S_c = S_v + S_p;
    const int qd = nh * hd, qkvd = (nh + 2 * c.n_kv_heads) * hd;
    const long kv_bs = (long)S_c * hd;
    const int Ma = B * Hz;
    const float dt = (float)(1.0 / c.n_steps);
    copy_f32(ws.act, noise, (size_t)Ma * c.action_dim, st);
    const bool no_taps = !(cap && (cap->denoise_action || cap->velocities || cap->action_preclip));
    const int smode = h->sampler;
    if (std::is_same<T, bf16>::value && no_taps && (smode == PZ_SAMPLER_AUTO || smode == PZ_SAMPLER_STREAM) &&
        denoise_mega3_supported(c, B) && B <= 2 && h->mega3[B - 1].buf) {
        // bs 1-2: barrier-free persistent sampler over the re-packed weight streams (denoise_mega3.cu)
        Mega3Buffers mb;
        mb.kcache = ws.kcache; mb.vcache = ws.vcache; mb.batch_total = B; mb.valid_len = valid_len;
        mb.noise = ws.act; mb.out = out; mb.ll = ws.mega_ll; mb.ll_bytes = ws.mega_ll_bytes;
        const char *e = nullptr;
        int rc = launch_denoise_mega3(c, w, h->action.data(), h->mega3[B - 1], mb, B, st, &e);
        if (rc) return fail(h, rc, std::string("persistent sampler launch failed: ") + (e ? e : "?"));
        return 0;
    }
    if (smode == PZ_SAMPLER_STREAM) return fail(h, PZ_ERR_INVALID, "PZ_SAMPLER_STREAM: not packed (pz_sampler_pack) or batch / configuration not covered");
    if (std::is_same<T, bf16>::value && denoise_mega_supported(c, B) && no_taps &&
        (smode == PZ_SAMPLER_AUTO || smode == PZ_SAMPLER_BARRIER)) {
        // small batch: the whole sampler as one persistent cooperative kernel (denoise_mega.cu)
        MegaBuffers mb;
        mb.kcache = ws.kcache; mb.vcache = ws.vcache; mb.batch_total = B; mb.valid_len = valid_len;
        mb.act = ws.act; mb.xa = ws.xa; mb.partials = ws.att_scratch; mb.out = out;
        mb.e1 = ws.e1; mb.z = ws.z; mb.qkv = ws.qkva; mb.mlp = ws.mlpa; mb.barrier = ws.mega_barrier;
        const char *e = nullptr;
        int rc = launch_denoise_mega(c, w, h->action.data(), mb, B, st, &e);
        if (rc) return fail(h, rc, std::string("denoise_mega launch failed: ") + (e ? e : "?"));
        return 0;
    }
    for (int step = 0; step < c.n_steps; ++step) {
        // action encoder (vla/modules.py:39-53); the time half of linear_2 is the
        // per-step constant enc_time_bias[step]
        launch_cast_pad<T>(ws.act, (T *)ws.a_in, Ma, c.action_dim, w.small_k_pad, st);
        PZ_TRY(Ops<T>::linear(h, lin(ws.a_in, w.small_k_pad, w.enc_w1, w.enc_b1, ws.e1, A, Ma, A,
                                     w.small_k_pad), st));
        PZ_TRY(Ops<T>::linear(h, lin(ws.e1, A, w.enc_w2a, w.enc_time_bias + (size_t)step * A, ws.z, A,
                                     Ma, A, A, LIN_SILU), st));
        PZ_TRY(Ops<T>::linear(h, lin(ws.z, A, w.enc_w3, w.enc_b3, ws.xa, A, Ma, A, A, LIN_OUT_F32,
                                     sqrtf((float)A)), st));
        PZ_TRY(action_layers<T>(h, ws, valid_len, B,
                                (cap && cap->denoise_action) ? cap->denoise_action + (size_t)step * c.n_layers * Ma * A : nullptr,
                                st));
        // final norm + decoder + Euler step (joint_model.py:375-380, pizero.py:479-481)
        LinearArgs fused = lin(ws.xa, A, w.dec_w, w.dec_b, ws.act, c.action_dim, Ma, c.action_dim, A,
                               LIN_OUT_F32 | LIN_ACCUM | LIN_NORM_A, dt);
        fused.norm_w = w.action_final_norm;
        if (std::is_same<T, bf16>::value && !(c.flags & PZ_FLAG_SIMPLE_KERNELS) && !(cap && cap->velocities) &&
            skinny_supported(fused)) {
            // one kernel: final norm, decoder, and the Euler update action += dt * (W h + b)
            PZ_TRY(launch_linear_skinny(fused, st));
        } else {
            PZ_TRY(norm_linear<T>(h, ws.xa, w.action_final_norm, ws.ha,
                                  lin(nullptr, A, w.dec_w, w.dec_b, ws.vel, 8, Ma, c.action_dim, A, LIN_OUT_F32), A, st));
            launch_euler(ws.act, ws.vel, 8, dt, Ma, c.action_dim,
                         (cap && cap->velocities) ? cap->velocities + (size_t)step * Ma * c.action_dim : nullptr,
                         st);
        }
    }
    if (cap && cap->action_preclip) copy_f32(cap->action_preclip, ws.act, (size_t)Ma * c.action_dim, st);
    launch_clamp_copy(ws.act, out, (long)Ma * c.action_dim, c.clip, st);
    return 0;
}

// ------------------------- one evaluation of the velocity field at a per-sample time -------
// (the body of the Euler loop, pizero.py:456-479, with time_embedding(t) computed on the device for an
//  arbitrary t per sample instead of the per-step constants; used by the training forward)
template <typename T>
static int run_velocity(pz_handle *h, const int32_t *valid_len, const float *psi, const float *t, float *v_out,
                        int v_ld, void *wsp, int B, cudaStream_t st) {
    const pz_config &c = h->cfg;
    const pz_weights &w = h->w;
    if (!w.enc_w2t || !w.enc_b2 || !w.time_freq) return fail(h, PZ_ERR_UNBOUND, "enc_w2t / enc_b2 / time_freq not bound");
    Workspace ws = carve(c, B, h->prefix_chunk, wsp);
    const int A = c.act_hidden, Hz = c.horizon, Ma = B * Hz;
    float *tbias = (float *)ws.mlpa;   // [B, A] fp32; the MLP scratch is idle until the first layer
    // time_cond = SinusoidalPosEmb(t) (vla/modules.py:9-22); its product with the time half of linear_2 (+ bias)
    launch_time_embed<T>(t, w.time_freq, (T *)ws.ha, B, A / 2, st);
    PZ_TRY(Ops<T>::linear(h, lin(ws.ha, A, w.enc_w2t, w.enc_b2, tbias, A, B, A, A, LIN_OUT_F32), st));
    // action encoder (vla/modules.py:39-53)
    launch_cast_pad<T>(psi, (T *)ws.a_in, Ma, c.action_dim, w.small_k_pad, st);
    PZ_TRY(Ops<T>::linear(h, lin(ws.a_in, w.small_k_pad, w.enc_w1, w.enc_b1, ws.e1, A, Ma, A, w.small_k_pad), st));
    PZ_TRY(Ops<T>::linear(h, lin(ws.e1, A, w.enc_w2a, nullptr, ws.xa, A, Ma, A, A, LIN_OUT_F32), st));
    launch_rowbias_silu<T>(ws.xa, tbias, (T *)ws.z, Ma, A, Hz, st);
    PZ_TRY(Ops<T>::linear(h, lin(ws.z, A, w.enc_w3, w.enc_b3, ws.xa, A, Ma, A, A, LIN_OUT_F32, sqrtf((float)A)), st));
    PZ_TRY(action_layers<T>(h, ws, valid_len, B, nullptr, st));
    // final norm + decoder (joint_model.py:375-380, pizero.py:479)
    PZ_TRY(norm_linear<T>(h, ws.xa, w.action_final_norm, ws.ha,
                          lin(nullptr, A, w.dec_w, w.dec_b, ws.vel, 8, Ma, c.action_dim, A, LIN_OUT_F32), A, st));
    if (v_out) cudaMemcpy2DAsync(v_out, (size_t)v_ld * 4, ws.vel, 8 * 4, (size_t)c.action_dim * 4, Ma,
                                 cudaMemcpyDeviceToDevice, st);
    return 0;
}

// PiZero.forward (pizero.py:607-661), forward value: psi_t, one velocity evaluation, mean squared error
template <typename T>
static int run_flow_loss(pz_handle *h, const int32_t *valid_len, const float *actions, const float *noise,
                         const float *t, float sig_min, float *loss, float *v_psi, void *wsp, int B, cudaStream_t st) {
    const pz_config &c = h->cfg;
    Workspace ws = carve(c, B, h->prefix_chunk, wsp);
    const int Ma = B * c.horizon;
    launch_psi(noise, actions, t, ws.act, B, c.horizon * c.action_dim, sig_min, st);
    PZ_TRY(run_velocity<T>(h, valid_len, ws.act, t, nullptr, 0, wsp, B, st));
    launch_fm_loss(ws.vel, 8, noise, actions, loss, v_psi, Ma, c.action_dim, sig_min, st);
    return 0;
}

// --------------------------------------------- text output (PiZero.infer_text, pizero.py:559-593) ------
// Only the vlm mixture is active, ALL layers run to the end (`final_layer_post_attn_skip_names=[]`), the final norm of the
// vlm mixture and the (tied) lm_head follow.  K / V go to a caller-owned text cache [layers][B][cache_rows][head_dim]
// that later decode steps append to (`cache_mode="append"`, joint_model.py:164-240).  Prompts are not padded
// (pizero.py:346-357 "assume no padding"): every sample has q_len valid tokens, positions 1 .. q_len.
template <typename T>
static int run_text_prefill(pz_handle *h, const int32_t *valid_len, void *kcache, void *vcache, int cache_rows, int q_len,
                            float *logits, int last_only, float *hidden, const float *ext_x, void *wsp, int B, cudaStream_t st) {
    const pz_config &c = h->cfg;
    const pz_weights &w = h->w;
    if (!w.vlm_final_norm || (logits && !w.lm_head)) return fail(h, PZ_ERR_UNBOUND, "text output needs vlm_final_norm and lm_head (use_lm_head / mixture.vlm.use_final_norm)");
    Workspace ws = carve(c, B, h->prefix_chunk, wsp);
    const int H = c.vlm_hidden, hd = c.head_dim, nh = c.n_heads, S_v = c.s_vlm;
    const int qd = nh * hd, qkvd = (nh + 2 * c.n_kv_heads) * hd;
    const long kv_bs = (long)cache_rows * hd;
    if (q_len < 1 || q_len > S_v || cache_rows < S_v) return fail(h, PZ_ERR_INVALID, "text prefill: 1 <= q_len <= max_image_text_tokens <= cache_rows");
    if (ext_x) copy_f32(ws.x, ext_x, (size_t)B * S_v * H, st);   // JointModel.forward entry: (scaled) embeddings given
    for (int b0 = 0; b0 < B; b0 += h->prefix_chunk) {
        const int nb = (B - b0 < h->prefix_chunk) ? B - b0 : h->prefix_chunk;
        const int M = nb * S_v;
        float *x = ws.x + (size_t)b0 * S_v * H;
        for (int l = 0; l < c.n_layers; ++l) {
            T *Kc = (T *)kcache + ((size_t)l * B + b0) * kv_bs;
            T *Vc = (T *)vcache + ((size_t)l * B + b0) * kv_bs;
            launch_rmsnorm<T>(x, h->vlm[l].norm_in, (T *)ws.h, M, H, 1e-6f, st);
            if (std::is_same<T, bf16>::value && !(c.flags & PZ_FLAG_SIMPLE_KERNELS) && hd == 256 && H % 8 == 0) {
                const char *e = nullptr;
                int rc = launch_qkv_rope_tc(ws.h, H, h->vlm[l].w_qkv, ws.q, Kc, Vc, kv_bs, w.rope_vlm_cos, w.rope_vlm_sin, M, H, nh,
                                            S_v, st, &e);
                if (rc) return fail(h, rc, e ? e : "fused qkv+rope launch failed");
            } else {
                PZ_TRY(Ops<T>::linear(h, lin(ws.h, H, h->vlm[l].w_qkv, nullptr, ws.qkv, qkvd, M, qkvd, H), st));
                launch_rope_split<T>((const T *)ws.qkv, qkvd, (T *)ws.q, (long)S_v * qd, Kc, Vc, kv_bs, w.rope_vlm_cos,
                                     w.rope_vlm_sin, nb, S_v, 0, nh, hd, st);
            }
            AttnArgs a;
            memset(&a, 0, sizeof(a));
            a.K = Kc; a.V = Vc; a.kv_batch_stride = kv_bs; a.kv_row_stride = hd; a.kv_head_stride = 0;
            a.valid_len = valid_len + b0;
            a.batch = nb; a.n_heads = nh; a.head_dim = hd; a.s_cache = cache_rows; a.s_vlm = S_v; a.n_fresh = 0;
            a.scale = 1.0f / sqrtf((float)hd); a.softcap = 50.f;
            a.q_row_stride = qd; a.q_head_stride = hd; a.o_row_stride = qd; a.o_head_stride = hd;
            a.Q = ws.q; a.q_batch_stride = (long)S_v * qd; a.O = ws.att; a.o_batch_stride = (long)S_v * qd;
            a.q_rows = S_v; a.q_row0 = 0;
            PZ_TRY(Ops<T>::attention(h, a, st));
            PZ_TRY(post_attention<T>(h, h->vlm[l], x, ws.h, ws.att, ws.mlp, M, H, c.vlm_inter, st));
        }
        // Mixture.forward_norm (mixture.py:68-77): the hidden states JointModel.forward returns
        if (hidden) launch_rmsnorm<float>(x, w.vlm_final_norm, hidden + (size_t)b0 * S_v * H, M, H, 1e-6f, st);
        if (!logits) continue;
        if (last_only) {
            // logits of the last prompt token only: [B, vocab]
            float *xg = (float *)ws.mlp;   // idle now
            cudaMemcpy2DAsync(xg, (size_t)H * 4, x + (size_t)(q_len - 1) * H, (size_t)S_v * H * 4, (size_t)H * 4, nb,
                              cudaMemcpyDeviceToDevice, st);
            PZ_TRY(norm_linear<T>(h, xg, w.vlm_final_norm, ws.h,
                                  lin(nullptr, H, w.lm_head, nullptr, logits + (size_t)b0 * c.vocab_size, c.vocab_size, nb,
                                      c.vocab_size, H, LIN_OUT_F32), H, st));
        } else {
            launch_rmsnorm<T>(x, w.vlm_final_norm, (T *)ws.h, M, H, 1e-6f, st);
            PZ_TRY(Ops<T>::linear(h, lin(ws.h, H, w.lm_head, nullptr, logits + (size_t)b0 * S_v * c.vocab_size, c.vocab_size, M,
                                         c.vocab_size, H, LIN_OUT_F32), st));
        }
    }
    return 0;
}

// One decode step: the embedding of one new token per sample (already scaled by sqrt(hidden), joint_model.py:355) at row
// `cur_len` of the cache, position cur_len + 1; attends rows 0 .. cur_len (no padding).  logits: [B, vocab].
template <typename T>
static int run_text_decode(pz_handle *h, const float *x_in, const int32_t *valid_len1, int cur_len, void *kcache, void *vcache,
                           int cache_rows, float *logits, float *hidden, void *wsp, int B, cudaStream_t st) {
    const pz_config &c = h->cfg;
    const pz_weights &w = h->w;
    if (!w.vlm_final_norm || (logits && !w.lm_head)) return fail(h, PZ_ERR_UNBOUND, "text output needs vlm_final_norm and lm_head");
    if (cur_len < 1 || cur_len >= cache_rows || cur_len >= (w.rope_vlm_rows > 0 ? w.rope_vlm_rows : c.s_vlm)) return fail(h, PZ_ERR_INVALID, "text decode: cache / RoPE table exhausted");
    Workspace ws = carve(c, B, h->prefix_chunk, wsp);
    const int H = c.vlm_hidden, hd = c.head_dim, nh = c.n_heads;
    const int qd = nh * hd, qkvd = (nh + 2 * c.n_kv_heads) * hd;
    const long kv_bs = (long)cache_rows * hd;
    float *x = ws.x;   // [B, H] fp32 residual stream of the new tokens
    copy_f32(x, x_in, (size_t)B * H, st);
    for (int l = 0; l < c.n_layers; ++l) {
        T *Kc = (T *)kcache + (size_t)l * B * kv_bs;
        T *Vc = (T *)vcache + (size_t)l * B * kv_bs;
        PZ_TRY(norm_linear<T>(h, x, h->vlm[l].norm_in, ws.h, lin(nullptr, H, h->vlm[l].w_qkv, nullptr, ws.qkv, qkvd, B, qkvd, H), H, st));
        // rotate q and k at position cur_len + 1 (table row cur_len); k, v -> cache row cur_len
        launch_rope_split<T>((const T *)ws.qkv, qkvd, (T *)ws.q, (long)qd, Kc + (size_t)cur_len * hd, Vc + (size_t)cur_len * hd, kv_bs,
                             w.rope_vlm_cos, w.rope_vlm_sin, B, 1, cur_len, nh, hd, st);
        AttnArgs a;
        memset(&a, 0, sizeof(a));
        a.K = Kc; a.V = Vc; a.kv_batch_stride = kv_bs; a.kv_row_stride = hd; a.kv_head_stride = 0;
        a.valid_len = valid_len1;   // cur_len + 1 for every sample
        a.batch = B; a.n_heads = nh; a.head_dim = hd; a.s_cache = cur_len + 1; a.s_vlm = cur_len + 1; a.n_fresh = 0;
        a.scale = 1.0f / sqrtf((float)hd); a.softcap = 50.f;
        a.q_row_stride = qd; a.q_head_stride = hd; a.o_row_stride = qd; a.o_head_stride = hd;
        a.Q = ws.q; a.q_batch_stride = (long)qd; a.O = ws.att; a.o_batch_stride = (long)qd;
        a.q_rows = 1; a.q_row0 = cur_len;
        a.scratch = ws.att_scratch; a.scratch_bytes = ws.att_scratch_bytes;
        PZ_TRY(Ops<T>::attention(h, a, st));
        PZ_TRY(post_attention<T>(h, h->vlm[l], x, ws.h, ws.att, ws.mlp, B, H, c.vlm_inter, st));
    }
    if (hidden) launch_rmsnorm<float>(x, w.vlm_final_norm, hidden, B, H, 1e-6f, st);
    if (logits)
        PZ_TRY(norm_linear<T>(h, x, w.vlm_final_norm, ws.h,
                              lin(nullptr, H, w.lm_head, nullptr, logits, c.vocab_size, B, c.vocab_size, H, LIN_OUT_F32), H, st));
    return 0;
}

// ------------------------------------------------------------------- ABI ----
extern "C" {

int pz_abi_version(void) { return PZ_ABI_VERSION; }

const char *pz_last_error(const pz_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int pz_create(const pz_config *cfg, pz_handle **out) {
    if (!cfg || !out) return fail(nullptr, PZ_ERR_INVALID, "null argument");
    const pz_config &c = *cfg;
    if (c.dtype != PZ_F32 && c.dtype != PZ_BF16) return fail(nullptr, PZ_ERR_INVALID, "dtype must be 0 (f32) or 1 (bf16)");
    if (c.n_kv_heads != 1) return fail(nullptr, PZ_ERR_INVALID, "only num_key_value_heads == 1 (MQA) is supported");
    if (c.head_dim % 2 || c.head_dim > 256) return fail(nullptr, PZ_ERR_INVALID, "head_dim must be even and <= 256");
    if (c.vit_hidden % c.vit_heads) return fail(nullptr, PZ_ERR_INVALID, "vit hidden_size % num_heads != 0");
    if (c.vlm_hidden % 4 || c.act_hidden % 4) return fail(nullptr, PZ_ERR_INVALID, "hidden sizes must be multiples of 4");
    if (c.vlm_inter % PZ_GU_BLOCK || c.act_inter % PZ_GU_BLOCK)   // packed gate|up layout: blocks of PZ_GU_BLOCK rows
        return fail(nullptr, PZ_ERR_INVALID, "intermediate sizes must be multiples of 128");
    if (c.action_dim > 8 || c.proprio_dim > 64) return fail(nullptr, PZ_ERR_INVALID, "action_dim <= 8 and proprio_dim <= 64 required");
    if (c.n_img_tokens != (c.image_size / c.patch_size) * (c.image_size / c.patch_size))
        return fail(nullptr, PZ_ERR_INVALID, "num_image_tokens must equal (image_size/patch_size)^2");
    if (c.s_vlm < c.n_images * c.n_img_tokens) return fail(nullptr, PZ_ERR_INVALID, "max_image_text_tokens smaller than the image tokens");
    if (c.patch_k_pad < 3 * c.patch_size * c.patch_size || c.patch_k_pad % 8) return fail(nullptr, PZ_ERR_INVALID, "bad patch_k_pad");
    if (c.max_batch < 1) return fail(nullptr, PZ_ERR_INVALID, "max_batch must be >= 1");
    pz_handle *h = new pz_handle();
    h->cfg = c;
    const char *e = getenv("PZ_PREFIX_CHUNK");
    if (e && atoi(e) > 0) h->prefix_chunk = atoi(e);
    *out = h;
    return PZ_OK;
}

void pz_destroy(pz_handle *h) {
    if (!h) return;
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : h->sync_ev) cudaEventDestroy(e);
    if (h->side) cudaStreamDestroy(h->side);
    delete h;
}

int pz_bind_weights(pz_handle *h, const pz_weights *w) {
    if (!h || !w) return PZ_ERR_INVALID;
    if (!w->vit || !w->vlm || !w->proprio || !w->action) return fail(h, PZ_ERR_INVALID, "layer tables missing");
    if (w->small_k_pad < h->cfg.action_dim || w->small_k_pad < h->cfg.proprio_dim || w->small_k_pad > 64 || w->small_k_pad % 8)
        return fail(h, PZ_ERR_INVALID, "bad small_k_pad");
    h->w = *w;
    h->vit.assign(w->vit, w->vit + h->cfg.vit_layers);
    h->vlm.assign(w->vlm, w->vlm + h->cfg.n_layers);
    h->proprio.assign(w->proprio, w->proprio + h->cfg.n_layers);
    h->action.assign(w->action, w->action + h->cfg.n_layers);
    h->w.vit = h->vit.data(); h->w.vlm = h->vlm.data();
    h->w.proprio = h->proprio.data(); h->w.action = h->action.data();
    h->bound = true;
    return PZ_OK;
}

int pz_set_pixel_format(pz_handle *h, int format) {
    if (!h) return PZ_ERR_INVALID;
    if (format != PZ_PIXELS_MODEL_DTYPE && format != PZ_PIXELS_U8) return fail(h, PZ_ERR_INVALID, "unknown pixel format");
    h->pixel_format = format;
    return PZ_OK;
}

static int device_sms(pz_handle *h) {
    h->num_sms = device_sm_count();
    return h->num_sms;
}

int pz_set_sampler(pz_handle *h, int mode) {
    if (!h) return PZ_ERR_INVALID;
    if (mode < PZ_SAMPLER_AUTO || mode > PZ_SAMPLER_STREAM) return fail(h, PZ_ERR_INVALID, "unknown sampler mode");
    h->sampler = mode;
    return PZ_OK;
}

size_t pz_sampler_stream_bytes(pz_handle *h, int batch) {
    if (!h || batch < 1 || batch > 2) return 0;
    return denoise_mega3_stream_bytes(h->cfg, batch, device_sms(h));
}

int pz_sampler_pack(pz_handle *h, int batch, void *d_stream, size_t bytes, void *stream) {
    if (!h) return PZ_ERR_INVALID;
    if (!h->bound) return fail(h, PZ_ERR_UNBOUND, "pz_bind_weights has not been called");
    if (batch < 1 || batch > 2 || !d_stream) return fail(h, PZ_ERR_INVALID, "pz_sampler_pack: batch must be 1 or 2");
    const char *e = nullptr;
    int rc = denoise_mega3_pack(h->cfg, h->w, h->action.data(), batch, device_sms(h), d_stream, bytes, &h->mega3[batch - 1],
                                (cudaStream_t)stream, &e);
    if (rc) return fail(h, rc, e ? e : "pz_sampler_pack failed");
    return PZ_OK;
}

size_t pz_workspace_bytes(const pz_handle *h, int batch) {
    if (!h || batch < 1) return 0;
    return carve(h->cfg, batch, h->prefix_chunk, nullptr).total;
}

size_t pz_debug_trace_offset(const pz_handle *h, int batch) {
    if (!h || batch < 1) return 0;
    Workspace ws = carve(h->cfg, batch, h->prefix_chunk, (void *)0x1000);
    return (size_t)((char *)ws.mega_barrier - (char *)0x1000);
}

int pz_kv_layout(const pz_handle *h, int batch, size_t *k_offset, size_t *v_offset,
                 size_t *layer_stride_elems) {
    if (!h || batch < 1) return PZ_ERR_INVALID;
    Workspace ws = carve(h->cfg, batch, h->prefix_chunk, (void *)0x1000);
    if (k_offset) *k_offset = (size_t)((char *)ws.kcache - (char *)0x1000);
    if (v_offset) *v_offset = (size_t)((char *)ws.vcache - (char *)0x1000);
    if (layer_stride_elems)
        *layer_stride_elems = (size_t)batch * (h->cfg.s_vlm + h->cfg.cond_steps) * h->cfg.head_dim;
    return PZ_OK;
}

static int precheck(pz_handle *h, const void *ws, size_t ws_bytes, int B) {
    if (!h) return PZ_ERR_INVALID;
    if (!h->bound) return fail(h, PZ_ERR_UNBOUND, "pz_bind_weights has not been called");
    if (B < 1 || B > h->cfg.max_batch) return fail(h, PZ_ERR_INVALID, "batch out of range (1..max_batch)");
    if (!ws || ws_bytes < pz_workspace_bytes(h, B)) return fail(h, PZ_ERR_WORKSPACE, "workspace too small");
    if (((uintptr_t)ws) & 1023) return fail(h, PZ_ERR_WORKSPACE, "workspace must be 1 KiB aligned");
    return PZ_OK;
}

static int finish(pz_handle *h, int rc) {
    g_launch_counter = nullptr;
    if (rc) return rc;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(h, PZ_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
    }
    return PZ_OK;
}

int pz_embed_prefix(pz_handle *h, const int64_t *ids, const void *pixels, void *ws, size_t ws_bytes,
                    int B, const pz_capture *cap, void *stream) {
    PZ_TRY(precheck(h, ws, ws_bytes, B));
    if (!ids || !pixels) return fail(h, PZ_ERR_INVALID, "null input");
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = h->cfg.dtype == PZ_BF16 ? run_embed_prefix<bf16>(h, ids, pixels, ws, B, cap, st)
                                     : run_embed_prefix<float>(h, ids, pixels, ws, B, cap, st);
    return finish(h, rc);
}

int pz_prefill(pz_handle *h, const int32_t *valid_len, const float *proprio, void *ws, size_t ws_bytes,
               int B, const pz_capture *cap, void *stream) {
    PZ_TRY(precheck(h, ws, ws_bytes, B));
    if (!valid_len || !proprio) return fail(h, PZ_ERR_INVALID, "null input");
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = h->cfg.dtype == PZ_BF16 ? run_prefill<bf16>(h, valid_len, proprio, ws, B, cap, st)
                                     : run_prefill<float>(h, valid_len, proprio, ws, B, cap, st);
    return finish(h, rc);
}

int pz_denoise(pz_handle *h, const int32_t *valid_len, const float *noise, float *out, void *ws,
               size_t ws_bytes, int B, const pz_capture *cap, void *stream) {
    PZ_TRY(precheck(h, ws, ws_bytes, B));
    if (!valid_len || !noise || !out) return fail(h, PZ_ERR_INVALID, "null input");
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = h->cfg.dtype == PZ_BF16 ? run_denoise<bf16>(h, valid_len, noise, out, ws, B, cap, st)
                                     : run_denoise<float>(h, valid_len, noise, out, ws, B, cap, st);
    // caller-side de-normalisation of the clipped chunk (simpler.py:102-125), when set
    if (!rc && h->act_scale) launch_affine_cols(out, (long)B * h->cfg.horizon, h->cfg.action_dim, h->act_scale, h->act_shift, st);
    return finish(h, rc);
}

int pz_set_io_normalization(pz_handle *h, const float *d_proprio_scale, const float *d_proprio_shift, int proprio_clip,
                            const float *d_action_scale, const float *d_action_shift) {
    if (!h) return PZ_ERR_INVALID;
    if ((d_proprio_scale == nullptr) != (d_proprio_shift == nullptr) || (d_action_scale == nullptr) != (d_action_shift == nullptr))
        return fail(h, PZ_ERR_INVALID, "scale and shift come in pairs");
    h->prop_scale = d_proprio_scale; h->prop_shift = d_proprio_shift; h->prop_clip = proprio_clip;
    h->act_scale = d_action_scale; h->act_shift = d_action_shift;
    return PZ_OK;
}

int pz_joint_prefix(pz_handle *h, const float *x_vlm, const float *x_proprio, const int32_t *valid_len,
                    void *ws, size_t ws_bytes, int B, void *stream) {
    PZ_TRY(precheck(h, ws, ws_bytes, B));
    if (!x_vlm || !x_proprio || !valid_len) return fail(h, PZ_ERR_INVALID, "null input");
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = h->cfg.dtype == PZ_BF16 ? run_prefill<bf16>(h, valid_len, nullptr, ws, B, nullptr, st, x_vlm, x_proprio)
                                     : run_prefill<float>(h, valid_len, nullptr, ws, B, nullptr, st, x_vlm, x_proprio);
    return finish(h, rc);
}

int pz_joint_action(pz_handle *h, const float *x_action, const int32_t *valid_len, float *out_hidden, void *wsp,
                    size_t ws_bytes, int B, void *stream) {
    PZ_TRY(precheck(h, wsp, ws_bytes, B));
    if (!x_action || !valid_len || !out_hidden) return fail(h, PZ_ERR_INVALID, "null input");
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws = carve(h->cfg, B, h->prefix_chunk, wsp);
    const size_t n = (size_t)B * h->cfg.horizon * h->cfg.act_hidden;
    copy_f32(ws.xa, x_action, n, st);
    int rc = h->cfg.dtype == PZ_BF16 ? action_layers<bf16>(h, ws, valid_len, B, nullptr, st)
                                     : action_layers<float>(h, ws, valid_len, B, nullptr, st);
    if (!rc) launch_rmsnorm<float>(ws.xa, h->w.action_final_norm, out_hidden, (long)B * h->cfg.horizon,
                                   h->cfg.act_hidden, 1e-6f, st);   // Mixture.forward_norm, mixture.py:68-77
    return finish(h, rc);
}

int pz_infer_action(pz_handle *h, const int64_t *ids, const void *pixels, const int32_t *valid_len,
                    const float *proprio, const float *noise, float *out, void *ws, size_t ws_bytes,
                    int B, const pz_capture *cap, void *stream) {
    if (h) h->lc.n = 0;
    PZ_TRY(pz_embed_prefix(h, ids, pixels, ws, ws_bytes, B, cap, stream));
    PZ_TRY(pz_prefill(h, valid_len, proprio, ws, ws_bytes, B, cap, stream));
    PZ_TRY(pz_denoise(h, valid_len, noise, out, ws, ws_bytes, B, cap, stream));
    return PZ_OK;
}

int pz_velocity(pz_handle *h, const int32_t *valid_len, const float *psi, const float *t, float *v_out, void *ws,
                size_t ws_bytes, int B, void *stream) {
    PZ_TRY(precheck(h, ws, ws_bytes, B));
    if (!valid_len || !psi || !t || !v_out) return fail(h, PZ_ERR_INVALID, "null input");
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = h->cfg.dtype == PZ_BF16 ? run_velocity<bf16>(h, valid_len, psi, t, v_out, h->cfg.action_dim, ws, B, st)
                                     : run_velocity<float>(h, valid_len, psi, t, v_out, h->cfg.action_dim, ws, B, st);
    return finish(h, rc);
}

int pz_flow_matching_loss(pz_handle *h, const int64_t *ids, const void *pixels, const int32_t *valid_len,
                          const float *proprio, const float *actions, const float *noise, const float *t,
                          float sig_min, float *loss, float *v_psi, void *ws, size_t ws_bytes, int B, void *stream) {
    if (h) h->lc.n = 0;
    PZ_TRY(pz_embed_prefix(h, ids, pixels, ws, ws_bytes, B, nullptr, stream));
    PZ_TRY(pz_prefill(h, valid_len, proprio, ws, ws_bytes, B, nullptr, stream));
    if (!actions || !noise || !t || !loss) return fail(h, PZ_ERR_INVALID, "null input");
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = h->cfg.dtype == PZ_BF16 ? run_flow_loss<bf16>(h, valid_len, actions, noise, t, sig_min, loss, v_psi, ws, B, st)
                                     : run_flow_loss<float>(h, valid_len, actions, noise, t, sig_min, loss, v_psi, ws, B, st);
    return finish(h, rc);
}

int pz_text_prefill(pz_handle *h, const int32_t *valid_len, void *kcache, void *vcache, int cache_rows, int q_len,
                    float *logits, int last_only, float *hidden, const float *x_in, void *ws, size_t ws_bytes, int B, void *stream) {
    PZ_TRY(precheck(h, ws, ws_bytes, B));
    if (!valid_len || !kcache || !vcache) return fail(h, PZ_ERR_INVALID, "null input");
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = h->cfg.dtype == PZ_BF16
                 ? run_text_prefill<bf16>(h, valid_len, kcache, vcache, cache_rows, q_len, logits, last_only, hidden, x_in, ws, B, st)
                 : run_text_prefill<float>(h, valid_len, kcache, vcache, cache_rows, q_len, logits, last_only, hidden, x_in, ws, B, st);
    return finish(h, rc);
}

int pz_text_decode(pz_handle *h, const float *x, const int32_t *valid_len1, int cur_len, void *kcache, void *vcache,
                   int cache_rows, float *logits, float *hidden, void *ws, size_t ws_bytes, int B, void *stream) {
    PZ_TRY(precheck(h, ws, ws_bytes, B));
    if (!x || !valid_len1 || !kcache || !vcache || (!logits && !hidden)) return fail(h, PZ_ERR_INVALID, "null input");
    g_launch_counter = &h->lc;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = h->cfg.dtype == PZ_BF16 ? run_text_decode<bf16>(h, x, valid_len1, cur_len, kcache, vcache, cache_rows, logits, hidden, ws, B, st)
                                     : run_text_decode<float>(h, x, valid_len1, cur_len, kcache, vcache, cache_rows, logits, hidden, ws, B, st);
    return finish(h, rc);
}

int64_t pz_launch_count(const pz_handle *h) { return h ? h->lc.n : 0; }
int64_t pz_fallback_count(const pz_handle *h) { return h ? h->fallbacks : 0; }

int pz_timing_begin(pz_handle *h, int tag) {
    if (!h) return PZ_ERR_INVALID;
    h->timing_tag = tag;
    h->ev_used = 0;
    return PZ_OK;
}

int pz_timing_end(pz_handle *h, double *total_ms, int64_t *launches) {
    if (!h) return PZ_ERR_INVALID;
    double ms = 0;
    for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
        if (cudaEventSynchronize(h->ev[i + 1]) != cudaSuccess) return fail(h, PZ_ERR_CUDA, "event sync failed");
        float t = 0;
        cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]);
        ms += t;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (int64_t)(h->ev_used / 2);
    h->timing_tag = 0;
    h->ev_used = 0;
    return PZ_OK;
}

// ---- single-op entry points -------------------------------------------------
int pz_op_linear(int impl, int dtype, const void *d_a, const void *d_w, const float *d_bias, void *d_c,
                 int M, int N, int K, int lda, int ldc, int flags, float alpha, void *stream) {
    LinearArgs a = lin(d_a, lda, d_w, d_bias, d_c, ldc, M, N, K, flags, alpha);
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == 0) {
        if (dtype == PZ_BF16) launch_linear_simple<bf16>(a, st); else launch_linear_simple<float>(a, st);
    } else if (impl == 1) {
        if (dtype != PZ_BF16 || !gemm_tc_supported(a)) return PZ_ERR_INVALID;
        const char *e = nullptr;
        int rc = launch_linear_tc(a, st, &e);
        if (rc) { g_create_error = e ? e : "tcgen05 gemm failed"; return rc; }
    } else if (impl == 2) {
        if (dtype != PZ_BF16 || !skinny_supported(a)) return PZ_ERR_INVALID;
        int rc = launch_linear_skinny(a, st);
        if (rc) return rc;
    } else {
        return PZ_ERR_INVALID;
    }
    return cudaPeekAtLastError() == cudaSuccess ? PZ_OK : PZ_ERR_CUDA;
}

int pz_op_linear_ex(int impl, int dtype, const void *d_a, const void *d_w, const float *d_bias, void *d_c, int M, int N, int K,
                    int lda, int ldc, int ldw, int flags, float alpha, void *stream) {
    LinearArgs a = lin(d_a, lda, d_w, d_bias, d_c, ldc, M, N, K, flags, alpha);
    a.ldw = ldw;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == 0) {
        if (dtype == PZ_BF16) launch_linear_simple<bf16>(a, st); else launch_linear_simple<float>(a, st);
    } else if (impl == 1) {
        if (dtype != PZ_BF16 || !gemm_tc_supported(a)) return PZ_ERR_INVALID;
        const char *e = nullptr;
        int rc = launch_linear_tc(a, st, &e);
        if (rc) { g_create_error = e ? e : "tcgen05 gemm failed"; return rc; }
    } else {
        return PZ_ERR_INVALID;
    }
    return cudaPeekAtLastError() == cudaSuccess ? PZ_OK : PZ_ERR_CUDA;
}

int pz_op_attention(int impl, int dtype, const void *d_q, const void *d_k, const void *d_v,
                    const void *d_k2, const void *d_v2, const int32_t *d_valid_len, void *d_out,
                    int batch, int n_heads, int head_dim, int q_rows, int q_row0, int s_cache,
                    int s_vlm, int n_fresh, int kv_heads, float scale, float softcap, void *d_scratch,
                    size_t scratch_bytes, void *stream) {
    // Dense test layout: Q,O [B, q_rows, n_heads*hd]; K,V [B, s_cache, kv_heads*hd];
    // K2,V2 [B, n_fresh, kv_heads*hd].
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    int qd = n_heads * head_dim, kd = kv_heads * head_dim;
    a.Q = d_q; a.q_batch_stride = (long)q_rows * qd; a.q_row_stride = qd; a.q_head_stride = head_dim;
    a.K = d_k; a.V = d_v; a.kv_batch_stride = (long)s_cache * kd; a.kv_row_stride = kd;
    a.kv_head_stride = kv_heads == 1 ? 0 : head_dim;
    a.K2 = d_k2; a.V2 = d_v2; a.kv2_batch_stride = (long)n_fresh * kd; a.kv2_row_stride = kd;
    a.valid_len = d_valid_len;
    a.O = d_out; a.o_batch_stride = (long)q_rows * qd; a.o_row_stride = qd; a.o_head_stride = head_dim;
    a.batch = batch; a.n_heads = n_heads; a.head_dim = head_dim; a.q_rows = q_rows; a.q_row0 = q_row0;
    a.s_cache = s_cache; a.s_vlm = s_vlm; a.n_fresh = n_fresh; a.scale = scale; a.softcap = softcap;
    a.scratch = (float *)d_scratch; a.scratch_bytes = scratch_bytes;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == 0) {
        if (dtype == PZ_BF16) launch_attn_simple<bf16>(a, st); else launch_attn_simple<float>(a, st);
    } else if (impl == 1) {
        if (dtype != PZ_BF16 || !attn_mma_supported(a)) return PZ_ERR_INVALID;
        int rc = launch_attn_mma(a, st);
        if (rc) return rc;
    } else if (impl == 3) {
        if (dtype != PZ_BF16) return PZ_ERR_INVALID;
        int rc;
        if (attn_tc_supported(a)) rc = launch_attn_tc(a, st);
        else if (attn_tc_vit_supported(a)) rc = launch_attn_tc_vit(a, st);
        else return PZ_ERR_INVALID;
        if (rc) return rc;
    } else {
        return PZ_ERR_INVALID;
    }
    return cudaPeekAtLastError() == cudaSuccess ? PZ_OK : PZ_ERR_CUDA;
}

}  // extern "C"
