/*
 * pz_b200.h -- C ABI of the B200-native `PiZero.infer_action` hot path.
 *
 * The reference (shroglck/open-pi-zero) is pure Python/PyTorch and exposes no
 * FFI layer; its boundary for this path is the Python class surface
 * (SURVEY.md section 8b).  This header is what the Python look-alike in
 * `open-pi-zero_b200/pizero.py` binds through ctypes, and what any other host
 * language would bind.  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer named `d_*` / inside pz_weights is a
 *     DEVICE pointer owned by the caller (torch allocations); the library never
 *     frees, reallocates or retains them beyond the handle's lifetime.
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as
 *     void*); no entry point synchronises or allocates device memory unless
 *     its comment says so.
 *   - return value: 0 = ok, negative = error (see pz_status); the message is
 *     available through pz_last_error().  Nothing throws or aborts.
 *   - one handle may be used by one host thread at a time.
 */
#ifndef PZ_B200_H
#define PZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PZ_ABI_VERSION 8

typedef enum pz_status {
    PZ_OK = 0,
    PZ_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
    PZ_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed */
    PZ_ERR_UNBOUND = -3,      /* weights not bound */
    PZ_ERR_WORKSPACE = -4     /* workspace too small */
} pz_status;

typedef enum pz_dtype { PZ_F32 = 0, PZ_BF16 = 1 } pz_dtype;

/* Model dimensions: config/train/bridge.yaml:85-181 of the reference, as read
 * by PiZero.__init__ (src/model/vla/pizero.py:30-103) and JointModel.__init__
 * (src/model/vla/joint_model.py:309-323). */
typedef struct pz_config {
    int32_t dtype;              /* pz_dtype: storage type of weights/activations */
    int32_t vocab_size, pad_token_id, image_token_index;
    int32_t s_vlm;              /* max_image_text_tokens (276) */
    int32_t n_img_tokens;       /* tokens per image (256) */
    int32_t n_images;           /* 1 (3 for the pi0-paper shape) */
    int32_t cond_steps, horizon, action_dim, proprio_dim;
    int32_t n_steps;            /* num_inference_steps (10) */
    float   clip;               /* final_action_clip_value; < 0 = none */
    int32_t n_layers, n_heads, n_kv_heads, head_dim;
    int32_t vlm_hidden, vlm_inter, act_hidden, act_inter;
    int32_t vit_hidden, vit_inter, vit_layers, vit_heads, image_size, patch_size;
    int32_t patch_k_pad;        /* K of the packed patch-embedding matrix (>= 3*p*p, mult of 8) */
    int32_t max_batch;          /* largest B any call will use (sizes the workspace) */
    int32_t flags;              /* PZ_FLAG_* */
} pz_config;

#define PZ_FLAG_SIMPLE_KERNELS 1   /* debug: run every op through the plain SIMT kernels */
#define PZ_FLAG_ALLOW_FALLBACK 2   /* bf16: a shape without a tensor-core kernel runs on the SIMT kernel instead of failing */

/* One SigLIP encoder layer (src/model/paligemma/siglip.py:197-238).  Matrices
 * are `[out,in]` row-major in the handle dtype; vectors are fp32. */
typedef struct pz_vit_layer {
    const float *ln1_w, *ln1_b;
    const void  *w_qkv;  const float *b_qkv;   /* [3*V, V]: q | k | v rows */
    const void  *w_o;    const float *b_o;     /* [V, V] */
    const float *ln2_w, *ln2_b;
    const void  *w_fc1;  const float *b_fc1;   /* [VI, V] */
    const void  *w_fc2;  const float *b_fc2;   /* [V, VI] */
} pz_vit_layer;

/* One mixture decoder layer (src/model/vla/mixture.py:80-218,
 * src/model/paligemma/modules.py:70-95).  No biases. */
typedef struct pz_mix_layer {
    const float *norm_in;      /* input_layernorm.weight [hidden] (raw w; kernels use 1+w) */
    const void  *w_qkv;        /* [(nh+2*nkv)*hd, hidden]: q | k | v rows */
    const void  *w_o;          /* [hidden, nh*hd] */
    const float *norm_post;    /* post_attention_layernorm.weight */
    const void  *w_gate_up;    /* [2*inter, hidden], blocks of PZ_GU_BLOCK gate rows then PZ_GU_BLOCK up rows */
    const void  *w_down;       /* [hidden, inter] */
} pz_mix_layer;

#define PZ_GU_BLOCK 128

typedef struct pz_weights {
    const void  *embed;                 /* embed_tokens.weight [vocab, H] */
    const void  *patch_w;               /* patch_embedding.weight as [V, patch_k_pad] (c,ky,kx order, zero pad) */
    const float *patch_b;               /* [V] */
    const float *pos_emb;               /* position_embedding.weight [n_img_tokens, V] fp32 */
    const pz_vit_layer *vit;            /* HOST array [vit_layers] */
    const float *post_ln_w, *post_ln_b;
    const void  *proj_w; const float *proj_b;      /* multi_modal_projector [H, V] */
    const pz_mix_layer *vlm, *proprio, *action;    /* HOST arrays [n_layers] */
    const float *action_final_norm;     /* joint_model.mixtures.action.norm.weight */
    const void  *enc_w1; const float *enc_b1;      /* action_encoder.linear_1 [A, action_dim_pad] */
    const void  *enc_w2a;               /* action_encoder.linear_2.weight[:, A:] -> [A, A] (action half) */
    const float *enc_time_bias;         /* [n_steps, A] fp32: W2[:, :A] . time_emb(t_i) + b2 (constant per step) */
    const void  *enc_w3; const float *enc_b3;      /* [A, A] */
    const void  *prop_w; const float *prop_b;      /* proprio_encoder [A, proprio_dim_pad] */
    const void  *dec_w;  const float *dec_b;       /* action_decoder [action_dim_pad8, A] */
    const float *rope_vlm_cos, *rope_vlm_sin;      /* [s_vlm, hd/2] fp32, positions 1..s_vlm */
    const float *rope_act_cos, *rope_act_sin;      /* [cond+horizon, hd/2] fp32, positions 1.. */
    int32_t small_k_pad;                /* padded K of enc_w1 / prop_w (>= action_dim, proprio_dim; mult of 8) */
    /* ABI 4: the time half of the action encoder for an arbitrary per-sample t (training forward, pz_velocity) */
    const void  *enc_w2t;               /* action_encoder.linear_2.weight[:, :A] -> [A, A] (time half) */
    const float *enc_b2;                /* action_encoder.linear_2.bias [A] */
    const float *time_freq;             /* [A/2] fp32: exp(-i ln(time_max_period) / (A/2 - 1)), vla/modules.py:15-19 */
    /* ABI 6: optional text output (pizero.py:105-112, 559-593); NULL / 0 when the model has no lm_head */
    const float *vlm_final_norm;        /* joint_model.mixtures.vlm.norm.weight [H] (mixture.vlm.use_final_norm) */
    const void  *lm_head;               /* [vocab, H], usually the embed pointer (tied, pizero.py:112) */
    int32_t rope_vlm_rows;              /* rows of rope_vlm_cos / _sin (0 = s_vlm): decode steps need positions > s_vlm */
} pz_weights;

/* Optional capture taps for per-layer parity tests (all fp32 device buffers,
 * any may be NULL).  They mirror what wrapping
 * `joint_model.forward_mixture_layers` records on the reference side. */
typedef struct pz_capture {
    float *vit_out;          /* [B*n_images*n_img_tokens, V] after post_layernorm */
    float *image_features;   /* [B*n_images*n_img_tokens, H] projector output (before /sqrt(H)) */
    float *prefix_embeds;    /* [B, s_vlm, H] residual stream entering layer 0 (= merged embedding * sqrt(H)) */
    float *prefix_vlm;       /* [n_layers-1, B, s_vlm, H] residual stream after each layer */
    float *prefix_proprio;   /* [n_layers-1, B, cond, A] */
    float *denoise_action;   /* [n_steps, n_layers, B, horizon, A] */
    float *velocities;       /* [n_steps, B, horizon, action_dim] */
    float *action_preclip;   /* [B, horizon, action_dim] */
} pz_capture;

typedef struct pz_handle pz_handle;

int pz_abi_version(void);

/* PiZero.__init__ (pizero.py:30-103): validates the configuration. Host only. */
int pz_create(const pz_config *cfg, pz_handle **out);
void pz_destroy(pz_handle *h);
const char *pz_last_error(const pz_handle *h);   /* h may be NULL: last create error */

/* load_state_dict (eval.py:182-188) after host-side packing: records device
 * pointers; builds TMA descriptors (host only, no device work). */
int pz_bind_weights(pz_handle *h, const pz_weights *w);

/* Bytes of scratch `pz_infer_action` needs for batch B (<= max_batch). */
size_t pz_workspace_bytes(const pz_handle *h, int batch);

/* Format of `d_pixel_values` for the following calls on this handle.
 *   PZ_PIXELS_MODEL_DTYPE (default): normalised floats in the handle dtype, as the reference's
 *                                    `pixel_values` argument (pizero.py:419)
 *   PZ_PIXELS_U8: raw uint8 [B, (n_images,) 3, H, W] camera frames; the normalisation the reference's caller
 *                 does on the host (VLAProcessor, src/model/vla/processing.py:27-58,108-113:
 *                 x/255, then (x - 0.5)/0.5) is fused into the patch-gather kernel. */
#define PZ_PIXELS_MODEL_DTYPE 0
#define PZ_PIXELS_U8 1
int pz_set_pixel_format(pz_handle *h, int format);

/* Where the prefix KV cache lives inside the workspace (for parity tests):
 * K and V are each [n_layers][B][s_vlm+cond][head_dim] in the handle dtype
 * (replaces src/model/kv_cache.py: list-of-tensors growing by torch.cat). */
int pz_kv_layout(const pz_handle *h, int batch, size_t *k_offset, size_t *v_offset,
                 size_t *layer_stride_elems);

/* Debug: workspace offset of the persistent sampler's barrier / phase-timestamp words. */
size_t pz_debug_trace_offset(const pz_handle *h, int batch);
/* The bs = 1 / 2 Euler sampler (pizero.py:454-489 at batch <= 2) runs as one persistent kernel that streams the
 * action expert's weights from a per-SM re-packed copy (csrc/denoise_mega3.cu).  The copy depends on the device's SM
 * count and on the batch (1 or 2), lives in caller-owned device memory and is optional: without it the call uses the
 * grid-barrier sampler (csrc/denoise_mega.cu).
 *   pz_sampler_stream_bytes : bytes of the copy for `batch` on the current device (0: configuration not covered)
 *   pz_sampler_pack         : builds it in `d_stream` (1 KiB aligned) on `stream`; call again after re-binding weights.
 *                             Synchronises `stream` once (table upload). */
size_t pz_sampler_stream_bytes(pz_handle *h, int batch);
/* Which implementation of the Euler loop pz_denoise uses (all give the same result within bf16 rounding):
 *   AUTO     the stream sampler when packed and the batch is covered, else the grid-barrier kernel (B * horizon <= 16),
 *            else one kernel per op
 *   KERNELS  always one kernel per op (the path large batches take)
 *   BARRIER  csrc/denoise_mega.cu or an error
 *   STREAM   csrc/denoise_mega3.cu or an error */
#define PZ_SAMPLER_AUTO 0
#define PZ_SAMPLER_KERNELS 1
#define PZ_SAMPLER_BARRIER 2
#define PZ_SAMPLER_STREAM 3
int pz_set_sampler(pz_handle *h, int mode);
int pz_sampler_pack(pz_handle *h, int batch, void *d_stream, size_t bytes, void *stream);

/* PiZero.infer_action (pizero.py:416-490), whole call.
 *   d_input_ids  int64 [B, s_vlm]
 *   d_pixels     handle dtype [B*n_images, 3, image, image], already normalised
 *   d_valid_len  int32 [B]: number of image+text tokens per sample (what the
 *                dense masks of build_causal_mask_and_position_ids encode)
 *   d_proprio    fp32 [B, cond, proprio_dim]
 *   d_noise      fp32 [B, horizon, action_dim]  (the reference's torch.randn, pizero.py:454)
 *   d_action_out fp32 [B, horizon, action_dim]
 *   cap          optional taps (NULL in production)                              */
int pz_infer_action(pz_handle *h, const int64_t *d_input_ids, const void *d_pixels,
                    const int32_t *d_valid_len, const float *d_proprio, const float *d_noise,
                    float *d_action_out, void *d_workspace, size_t workspace_bytes, int batch,
                    const pz_capture *cap, void *stream);

/* The three stages of the call, separately (bench / profiling / tests).
 * pz_embed_prefix  = _forward_siglip_and_text_embedding (pizero.py:376-414)
 * pz_prefill       = proprio_encoder + prefix JointModel.forward (pizero.py:436-451)
 * pz_denoise       = the Euler loop (pizero.py:454-489)
 * They communicate through the workspace and must run in this order. */
int pz_embed_prefix(pz_handle *h, const int64_t *d_input_ids, const void *d_pixels,
                    void *d_workspace, size_t workspace_bytes, int batch,
                    const pz_capture *cap, void *stream);
int pz_prefill(pz_handle *h, const int32_t *d_valid_len, const float *d_proprio,
               void *d_workspace, size_t workspace_bytes, int batch,
               const pz_capture *cap, void *stream);
int pz_denoise(pz_handle *h, const int32_t *d_valid_len, const float *d_noise,
               float *d_action_out, void *d_workspace, size_t workspace_bytes, int batch,
               const pz_capture *cap, void *stream);

/* JointModel.forward (joint_model.py:328-383) for the two call patterns of the path, from
 * caller-provided embeddings (fp32, already multiplied by sqrt(hidden) as joint_model.py:348-355 does):
 *   pz_joint_prefix : vlm + proprio active, fills the workspace KV cache (return_caches=True)
 *   pz_joint_action : action active, cache_mode="append_non_active"; writes the action hidden state
 *                     after the mixture's final norm, fp32 [B, horizon, act_hidden]               */
int pz_joint_prefix(pz_handle *h, const float *d_x_vlm, const float *d_x_proprio, const int32_t *d_valid_len,
                    void *d_workspace, size_t workspace_bytes, int batch, void *stream);
int pz_joint_action(pz_handle *h, const float *d_x_action, const int32_t *d_valid_len, float *d_out_hidden,
                    void *d_workspace, size_t workspace_bytes, int batch, void *stream);

/* One evaluation of the velocity field v(psi, t) over the prefix KV that pz_prefill left in the workspace:
 * time_embedding (vla/modules.py:9-22) + action_encoder (:25-53) + the action expert's layers + final norm +
 * action_decoder, i.e. the body of the Euler loop (pizero.py:456-479) for an arbitrary per-sample time.
 *   d_psi fp32 [B, horizon, action_dim]; d_t fp32 [B]; d_v_out fp32 [B, horizon, action_dim] */
int pz_velocity(pz_handle *h, const int32_t *d_valid_len, const float *d_psi, const float *d_t, float *d_v_out,
                void *d_workspace, size_t workspace_bytes, int batch, void *stream);

/* PiZero.forward (pizero.py:607-661): the flow-matching training loss, FORWARD VALUE ONLY (no gradients).
 * The reference runs one joint pass with all three mixtures active under the full block mask and no cache;
 * vlm / proprio rows never see action keys (pizero.py:271-310), so that equals the prefix pass followed by one
 * action pass over the cached prefix: the call runs pz_embed_prefix, pz_prefill, psi_t (pizero.py:597-605),
 * pz_velocity and the mean-squared error.
 *   d_actions fp32 [B, horizon, action_dim]  x1, the ground-truth chunk
 *   d_noise   fp32 [B, horizon, action_dim]  x0 (the reference's torch.randn_like, pizero.py:622)
 *   d_t       fp32 [B]                       flow-matching time per sample (train.py:239-247 samples it)
 *   sig_min                                  cfg.flow_sig_min (pizero.py:58, default 0.001)
 *   d_loss    fp32 [1]                       mean((v_psi - (x1 - (1 - sig_min) x0))^2)
 *   d_v_psi   fp32 [B, horizon, action_dim]  optional (may be NULL): the predicted velocity */
int pz_flow_matching_loss(pz_handle *h, const int64_t *d_input_ids, const void *d_pixels,
                          const int32_t *d_valid_len, const float *d_proprio, const float *d_actions,
                          const float *d_noise, const float *d_t, float sig_min, float *d_loss, float *d_v_psi,
                          void *d_workspace, size_t workspace_bytes, int batch, void *stream);

/* Text output, replaces PiZero.infer_text (pizero.py:559-593; only the vlm mixture, all layers, final norm, lm_head).
 * pz_text_prefill runs AFTER pz_embed_prefix on the same workspace (it consumes the merged embeddings) and fills the
 * caller-owned text cache d_kcache / d_vcache [n_layers][batch][cache_rows][head_dim] (model dtype; K post-RoPE).
 * Prompts are not padded (pizero.py:346-357): q_len valid tokens per sample, d_valid_len[b] = q_len.
 *   d_logits  fp32 [batch, s_vlm, vocab] (rows >= q_len are meaningless), or [batch, vocab] = the last prompt token with
 *             last_only != 0, or NULL (cache only)
 * pz_text_decode appends ONE token per sample at cache row cur_len (position cur_len + 1, `cache_mode="append"`,
 * joint_model.py:164-240): d_x fp32 [batch, H] = embedding * sqrt(H); d_valid_len1[b] = cur_len + 1; d_logits [batch, vocab].
 * d_hidden (either call, optional): fp32 hidden states after the vlm mixture's final norm, [batch, s_vlm, H] / [batch, H] -- what
 * `JointModel.forward(embeds_all={"vlm": ...}, cache_mode="append", final_layer_post_attn_skip_names=[])` returns
 * (joint_model.py:375-380).  d_x_in (prefill, optional): fp32 [batch, s_vlm, H] scaled embeddings to use instead of the
 * pz_embed_prefix result (the JointModel.forward entry). */
int pz_text_prefill(pz_handle *h, const int32_t *d_valid_len, void *d_kcache, void *d_vcache, int cache_rows, int q_len,
                    float *d_logits, int last_only, float *d_hidden, const float *d_x_in, void *d_workspace, size_t workspace_bytes,
                    int batch, void *stream);
int pz_text_decode(pz_handle *h, const float *d_x, const int32_t *d_valid_len1, int cur_len, void *d_kcache, void *d_vcache,
                   int cache_rows, float *d_logits, float *d_hidden, void *d_workspace, size_t workspace_bytes, int batch, void *stream);

/* Caller-side normalisation folded into the path (env_adapter/base.py:8-49, simpler.py:76-125): the raw proprio
 * is mapped x -> x * scale[c] + shift[c] (then clipped to [-1, 1] if proprio_clip) inside the kernel that first reads it, and
 * the clipped action chunk a -> a * scale[c] + shift[c] after the sampler.  Device fp32 vectors [proprio_dim] / [action_dim]
 * owned by the caller; NULL pairs switch a side off (the default).  Affects pz_prefill / pz_denoise / pz_infer_action. */
int pz_set_io_normalization(pz_handle *h, const float *d_proprio_scale, const float *d_proprio_shift, int proprio_clip,
                            const float *d_action_scale, const float *d_action_shift);

/* The flow-matching training step, replaces PiZero.forward + loss.backward() (pizero.py:607-661, train.py:350-368):
 * forward with the activations kept in the training workspace, then the backward (statement: oracle/pizero_backward.py).
 *   d_actions, d_noise  fp32 [batch, horizon, action_dim];  d_t fp32 [batch] (TrainAgent.sample_fm_time, train.py:239-247)
 *   grads     NULL = loss only; else the SAME structs as the weights with every non-NULL pointer an fp32 device buffer of
 *             the packed shape of that weight (fused q|k|v rows, gate|up blocks of PZ_GU_BLOCK, padded small matrices,
 *             dec_b [8]); d loss / d weight * loss_scale is ACCUMULATED into them (micro-batches add up, as under DDP
 *             no_sync, train.py:350-356).  Fields without a gradient (embed: frozen, pizero.py:243-249; RoPE / time tables)
 *             are ignored.  If the proprio table aliases the action table (tied weights) both gradients land in one buffer.
 *   d_loss    fp32 [1]: mean squared error of this batch (unscaled)
 *   flags     PZ_TRAIN_FREEZE_VISION: stop at the projector output (no SigLIP / projector gradients)
 *   events    optional array of cudaEvent_t (NULL entries skipped) recorded on `stream` when a group of gradients is final, so
 *             that a data-parallel caller can all-reduce that slice on another stream while the backward goes on (DDP bucket
 *             overlap, train.py:121): [l] joint layer l (vlm / action / proprio entries of that layer), [n_layers] the encoder /
 *             decoder heads and the final norm, [n_layers + 1 + i] SigLIP layer i, [n_layers + 1 + vit_layers] the rest
 * Image tokens must be the first n_images * n_img_tokens positions of every sequence (what VLAProcessor builds,
 * processing.py:63-136).  Workspace: pz_train_workspace_bytes(batch), 1 KiB aligned. */
#define PZ_TRAIN_FREEZE_VISION 1
size_t pz_train_workspace_bytes(const pz_handle *h, int batch);
int pz_flow_matching_step(pz_handle *h, const int64_t *d_input_ids, const void *d_pixels, const int32_t *d_valid_len,
                          const float *d_proprio, const float *d_actions, const float *d_noise, const float *d_t, float sig_min,
                          const pz_weights *grads, float loss_scale, float *d_loss, void *d_workspace, size_t workspace_bytes,
                          int batch, int flags, void *const *events, int n_events, void *stream);

/* Optimizer step on the flat gradient buffer (replaces clip_grad_norm_ + AdamW8bit.step, train.py:371-379; the reference's
 * 8-bit AdamW is the third-party bitsandbytes 0.4x optimizer, absent from /root/reference -- the update rule here is
 * torch.optim.AdamW's on fp32 master weights and fp32 moments).
 * pz_grad_sumsq: d_out[0] = sum of squares of d_grad[0..n), summed in a fixed order (bit-reproducible across replicas);
 *   d_out must hold 1 + PZ_SUMSQ_SCRATCH floats (the per-CTA partials live behind the result).
 * pz_adamw_step: elements [begin, end) (one parameter group; begin a multiple of 256) of the flat master / grad / m / v buffers;
 *   g = grad * grad_scale * min(1, max_grad_norm / (sqrt(*d_sumsq) * grad_scale + 1e-6))   (d_sumsq NULL = no clipping)
 *   and the updated weight is written, rounded to dst_dtype, into the packed weight tensor of its entry:
 *   entry e covers flat elements [d_entry_off[e], d_entry_off[e] + d_entry_n[e]) -> d_entry_dst[e][0 .. d_entry_n[e]);
 *   entry offsets are multiples of 256.  zero_grad != 0 clears the gradient elements it consumed.  step counts from 1. */
#define PZ_SUMSQ_SCRATCH 1184
int pz_grad_sumsq(const float *d_grad, size_t n, float *d_out, void *stream);
int pz_adamw_step(float *d_master, float *d_grad, float *d_m, float *d_v, size_t begin, size_t end, const long long *d_entry_off,
                  void *const *d_entry_dst, const long long *d_entry_n, int n_entries, int dst_dtype, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, const float *d_sumsq, float max_grad_norm,
                  float grad_scale, int zero_grad, void *stream);

/* Weight averaging on the flat master buffer (replaces torch.optim.swa_utils.AveragedModel, model_averaging.py:40-66):
 * d_avg += (d_x - d_avg) * weight; EMA: weight = 1 - ema_decay, SWA: weight = 1 / (n_averaged + 1).
 * pz_write_packed writes any flat fp32 buffer of the gradient layout (the averaged weights for validation, the master
 * weights to restore) into the packed weight tensors, rounded to dst_dtype; same entry tables as pz_adamw_step. */
int pz_average_update(float *d_avg, const float *d_x, size_t n, float weight, void *stream);
int pz_write_packed(const float *d_flat, size_t begin, size_t end, const long long *d_entry_off, void *const *d_entry_dst,
                    const long long *d_entry_n, int n_entries, int dst_dtype, void *stream);

/* Number of kernels the last call on this handle launched (bench: gpu_launches). */
int64_t pz_launch_count(const pz_handle *h);

/* Ops that ran on the plain SIMT kernels because no tensor-core kernel covers their shape, since pz_create.  Always 0
 * unless PZ_FLAG_ALLOW_FALLBACK is set (without it such a shape fails the call with PZ_ERR_INVALID). */
int64_t pz_fallback_count(const pz_handle *h);

/* CUDA-event timing of one kernel family inside real calls (bench.py's roofline
 * object).  tag: 1 = VLM gate|up GEMM, 2 = VLM down GEMM, 3 = action gate|up.
 * pz_timing_begin arms it; every tagged launch of the following calls is
 * bracketed by events on the launching stream; pz_timing_end synchronises on
 * those events and returns the summed duration and the number of launches. */
int pz_timing_begin(pz_handle *h, int tag);
int pz_timing_end(pz_handle *h, double *total_ms, int64_t *launches);

/* ---- single-op entry points (unit tests of the hand-written kernels) ---- */

/* C = alpha * epilogue(A[M,K] . W[N,K]^T + bias[N]).  impl: 0 = SIMT reference
 * kernel, 1 = tcgen05/TMA kernel (bf16 only), 2 = skinny weight-streaming
 * kernel (bf16, M <= 64).  flags: 1 gelu-tanh, 2 C is fp32 (else `dtype`),
 * 4 accumulate into fp32 C, 8 GeGLU over [128 gate | 128 up] row blocks of W
 * (C is [M, N/2]), 16 SiLU. */
int pz_op_linear(int impl, int dtype, const void *d_a, const void *d_w, const float *d_bias,
                 void *d_c, int M, int N, int K, int lda, int ldc, int flags, float alpha,
                 void *stream);

/* Block-masked attention over [cache | fresh] keys (csrc/common.cuh: AttnArgs).
 * Dense test layout: Q,O [B, q_rows, n_heads*hd]; K,V [B, s_cache, kv_heads*hd];
 * K2,V2 [B, n_fresh, kv_heads*hd].  impl: 0 = SIMT reference, 1 = mma.sync tensor-core
 * kernel (bf16), 3 = tcgen05 kernel (prefix vlm rows only; -1 if the shape is not covered).  d_scratch (optional fp32) enables the split-key decode path. */
/* pz_op_linear with MN-major operands: flags may carry 128 (A stored [K][M], row stride lda) and 256 (W stored [K][N], row
 * stride ldw) -- how the training step's backward reads dY, X and W in place (impl 0 = SIMT, 1 = tcgen05). */
int pz_op_linear_ex(int impl, int dtype, const void *d_a, const void *d_w, const float *d_bias, void *d_c, int M, int N, int K,
                    int lda, int ldc, int ldw, int flags, float alpha, void *stream);
int pz_op_attention(int impl, int dtype, const void *d_q, const void *d_k, const void *d_v,
                    const void *d_k2, const void *d_v2, const int32_t *d_valid_len, void *d_out,
                    int batch, int n_heads, int head_dim, int q_rows, int q_row0, int s_cache,
                    int s_vlm, int n_fresh, int kv_heads, float scale, float softcap,
                    void *d_scratch, size_t scratch_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PZ_B200_H */
