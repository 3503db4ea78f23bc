// kernels.h -- host-callable launchers (one per kernel family).
#pragma once
#include "common.cuh"

// simple.cu
template <typename T> void launch_linear_simple(const LinearArgs &a, cudaStream_t st);
template <typename T> void launch_attn_simple(const AttnArgs &a, cudaStream_t st);

// elementwise.cu
template <typename T>
void launch_im2col(const T *pix, T *patches, int n_images, int image, int patch, int k_pad,
                   cudaStream_t st);
template <typename T>
void launch_im2col_u8(const uint8_t *pix, T *patches, int n_images, int image, int patch, int k_pad,
                      cudaStream_t st);
void launch_bcast_rows(float *x, const float *table, long rows, int cols, int period,
                       cudaStream_t st);
template <typename T>
void launch_layernorm(const float *x, const float *w, const float *b, T *out, long rows, int cols,
                      float eps, cudaStream_t st);
template <typename T>
void launch_rmsnorm(const float *x, const float *w, T *out, long rows, int cols, float eps,
                    cudaStream_t st);
template <typename T>
void launch_embed_merge(const int64_t *ids, const T *embed, const float *feats, float *x,
                        int batch, int s_vlm, int hidden, int n_feat_rows, int image_token,
                        int pad_token, float text_scale, cudaStream_t st);
template <typename T>
void launch_rope_split(const T *qkv, int qkv_ld, T *q_out, long q_batch_stride, T *k_out,
                       T *v_out, long kv_batch_stride, const float *cos_t, const float *sin_t,
                       int batch, int s_x, int pos0, int n_heads, int head_dim,
                       cudaStream_t st);
template <typename T>
void launch_cast_pad(const float *src, T *dst, long rows, int cols, int cols_pad, cudaStream_t st);
template <typename T>
void launch_cast_pad_affine(const float *src, T *dst, long rows, int cols, int cols_pad, const float *scale,
                            const float *shift, int clip, cudaStream_t st);
void launch_affine_cols(float *x, long rows, int cols, const float *scale, const float *shift, cudaStream_t st);
template <typename T>
void launch_to_f32(const T *src, float *dst, long n, cudaStream_t st);
void launch_euler(float *action, const float *vel, int vel_ld, float dt, long rows, int adim,
                  float *vel_capture, cudaStream_t st);
void launch_clamp_copy(const float *src, float *dst, long n, float clip, cudaStream_t st);
// flow-matching training forward (pizero.py:597-661)
void launch_psi(const float *x0, const float *x1, const float *t, float *out, long batch, int per_sample, float sig_min,
                cudaStream_t st);
template <typename T>
void launch_time_embed(const float *t, const float *freq, T *out, int batch, int half, cudaStream_t st);
template <typename T>
void launch_rowbias_silu(const float *zpre, const float *bias, T *out, long rows, int cols, int rows_per_sample,
                         cudaStream_t st);
void launch_fm_loss(const float *vel, int vel_ld, const float *x0, const float *x1, float *loss, float *v_out, long rows,
                    int adim, float sig_min, cudaStream_t st);

// gemm_tc.cu (tcgen05 + TMA, bf16)
int gemm_tc_supported(const LinearArgs &a);
int launch_linear_tc(const LinearArgs &a, cudaStream_t st, const char **err);
int launch_qkv_rope_tc(const void *A, int lda, const void *W, void *q_out, void *k_out, void *v_out,
                       long kv_batch_stride, const float *cos_t, const float *sin_t, int M, int K,
                       int n_heads, int s_x, cudaStream_t st, const char **err);

// skinny.cu (weight streaming for M <= 64, bf16)
int skinny_supported(const LinearArgs &a);
int launch_linear_skinny(const LinearArgs &a, cudaStream_t st);

// attn_mma.cu (tensor-core attention, bf16)
int attn_mma_supported(const AttnArgs &a);
int launch_attn_mma(const AttnArgs &a, cudaStream_t st);
// split-key decode only: returns the number of key splits (> 1) if the partials path applies and
// launches just the partial kernel (no combine); 0 if it does not apply (nothing launched).
int launch_attn_mma_partials(const AttnArgs &a, cudaStream_t st);

// attn_tc.cu (tcgen05 attention for the vlm rows of the prefix pass: bf16, head_dim 256, 8 query heads on one K/V head)
int attn_tc_supported(const AttnArgs &a);
int launch_attn_tc(const AttnArgs &a, cudaStream_t st);
// SigLIP encoder attention on tcgen05 (256 tokens, head_dim 72, no mask)
int attn_tc_vit_supported(const AttnArgs &a);
int launch_attn_tc_vit(const AttnArgs &a, cudaStream_t st);

// denoise_mega.cu (persistent cooperative sampler for B * horizon <= 16)
struct MegaBuffers {
    const void *kcache, *vcache;   // [L][batch_total][S_c][256] bf16
    int batch_total;
    const int32_t *valid_len;
    float *act, *xa, *partials, *out;
    void *e1, *z, *qkv, *mlp;
    unsigned int *barrier;         // 2 words
};
int denoise_mega_supported(const pz_config &c, int B);
int launch_denoise_mega(const pz_config &c, const pz_weights &w, const pz_mix_layer *layers, const MegaBuffers &bf,
                        int B, cudaStream_t st, const char **err);

// denoise_mega3.cu (barrier-free persistent sampler for B * horizon <= 8: exchanges through {payload, sequence}
// words, per-CTA weight streams re-packed in MMA fragment order and fed by 32 KB bulk copies, attention CTAs with
// the layer's K / V resident in shared memory)
struct Mega3State {      // filled by denoise_mega3_pack
    void *buf = nullptr;
    size_t bytes = 0, slots_off = 0;
    int B = 0, G = 0, NA = 0, NP = 0, num_sms = 0;
};
struct Mega3Buffers {
    const void *kcache, *vcache;   // [L][batch_total][S_c][256] bf16
    int batch_total;
    const int32_t *valid_len;
    const float *noise;            // [B*horizon][action_dim] initial action
    float *out;
    void *ll; size_t ll_bytes;     // exchange buffers, >= denoise_mega3_ll_bytes()
};
int denoise_mega3_supported(const pz_config &c, int B);
size_t denoise_mega3_ll_bytes(const pz_config &c, int B);
size_t denoise_mega3_stream_bytes(const pz_config &c, int B, int num_sms);
int denoise_mega3_pack(const pz_config &c, const pz_weights &w, const pz_mix_layer *layers, int B, int num_sms, void *buf,
                       size_t bytes, Mega3State *state, cudaStream_t st, const char **err);
int launch_denoise_mega3(const pz_config &c, const pz_weights &w, const pz_mix_layer *layers, const Mega3State &state,
                         const Mega3Buffers &bf, int B, cudaStream_t st, const char **err);

// stand-alone decode attention (one CTA per (sample, 64-key tile), RoPE fused, 8 warps): writes split-key
// partials (bf16 o [B][splits][heads*horizon][256], then fp32 l) and combines them into `out` [B][horizon][heads*256];
// returns the number of splits (or < 0 on error)
int decode_attention_supported(const pz_config &c);
int launch_decode_attention(const pz_config &c, const pz_weights &w, const void *qkv, const void *kcache,
                            const void *vcache, int batch_total, const int32_t *valid_len, float *partials, int layer,
                            int B, void *out, long out_batch_stride, int out_row_stride, cudaStream_t st);
int launch_attn_combine(const AttnArgs &a, int n_splits, cudaStream_t st);
// one CTA per sample, K and V read once, no split-key partials (denoise_mega3.cu)
int decode_attention2_supported(const pz_config &c);
int launch_decode_attention2(const pz_config &c, const pz_weights &w, const void *qkv, const void *kcache, const void *vcache,
                             int batch_total, const int32_t *valid_len, int layer, int B, void *out, long out_batch_stride,
                             int out_row_stride, cudaStream_t st);
