"""TEST INFRASTRUCTURE ONLY -- CPU oracle for `PiZero.infer_action`.

A functional, state-dict driven restatement (plain PyTorch CPU ops, fp32 by
default) of the reference's action-inference path, written to be read next to
the reference (shroglck/open-pi-zero; every function cites the file:line it
follows).  It is the checker for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  The product package never does (it fails loudly when the
CUDA library is missing instead of falling back to this).

Parity pinning: `tests/golden/*.pt` were produced by the *unmodified*
reference, imported through `oracle/ref_shims.py` (script:
`oracle/make_golden.py`).  `tests/test_oracle_golden.py` checks this file
against those vectors on every run, and -- where `/root/reference` exists --
against the live reference on fresh seeds.

Inputs follow the reference's call: `input_ids [B,S_v] int64`,
`pixel_values [B,3,H,W]` (or `[B,n_img,3,H,W]` for the multi-image extension,
SURVEY.md F10), `attention_mask [B,S_v]` (1 = image/text token, 0 = pad; what
the tokenizer returns and `build_causal_mask_and_position_ids` consumes),
`proprios [B,cond,P]`, and the initial `noise [B,H,A]` (the reference draws it
internally, pizero.py:454; here it is an argument so both sides can share it).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

ATTN_SOFTCAP = 50.0  # hard-coded default argument, joint_model.py:139


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def gemma_rms_norm(x, w, eps=1e-6):
    """paligemma/modules.py:13-21 -- fp32 norm, scale by (1 + w), cast back."""
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (y * (1.0 + w.float())).type_as(x)


def rope_cos_sin(position_ids, head_dim, theta, dtype):
    """paligemma/modules.py:36-67 -- fp32 table, cat(freqs, freqs), cast."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = position_ids[:, :, None].float() * inv_freq[None, None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def apply_rope(x, cos, sin):
    """model/utils.py:4-16 -- half-split (HF) convention; x is [B,heads,S,hd]."""
    half = x.shape[-1] // 2
    rot = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
    return x * cos[:, None] + rot * sin[:, None]


def gemma_mlp(x, gate_w, up_w, down_w):
    """paligemma/modules.py:86-95 -- down(gelu_tanh(gate(x)) * up(x)), no bias."""
    return F.linear(F.gelu(F.linear(x, gate_w), approximate="tanh") * F.linear(x, up_w), down_w)


def sinusoidal_time_embedding(t, dim, max_period):
    """vla/modules.py:15-22 -- arange in the model dtype, cat(sin, cos)."""
    half = dim // 2
    k = math.log(max_period) / (half - 1)
    f = torch.exp(torch.arange(half, dtype=t.dtype) * -k)
    e = t[:, None] * f[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def action_encoder(sd, action, time_emb):
    """vla/modules.py:39-53 with time_cond=True (pizero.py:86-90): linear_1,
    cat([time, emb]) (time first), linear_2, SiLU, linear_3."""
    e = F.linear(action, sd["action_encoder.linear_1.weight"], sd["action_encoder.linear_1.bias"])
    tfull = time_emb[:, None, :].expand(-1, action.shape[1], -1)
    e = torch.cat([tfull, e], dim=-1)
    e = F.silu(F.linear(e, sd["action_encoder.linear_2.weight"], sd["action_encoder.linear_2.bias"]))
    return F.linear(e, sd["action_encoder.linear_3.weight"], sd["action_encoder.linear_3.bias"])


# --------------------------------------------------------------------------
# SigLIP + projector + embedding merge
# --------------------------------------------------------------------------
def siglip_forward(sd, dims, pixel_values, capture=None):
    """siglip.py:59-78 (patch conv + position table), :220-238 x layers
    (pre-LN attention and MLP blocks), :298 (post LN)."""
    p = "vision_tower.vision_model."
    x = F.conv2d(pixel_values, sd[p + "embeddings.patch_embedding.weight"],
                 sd[p + "embeddings.patch_embedding.bias"], stride=dims["patch_size"])
    x = x.flatten(2).transpose(1, 2)
    x = x + sd[p + "embeddings.position_embedding.weight"][None]
    B, S, D = x.shape
    nh = dims["vit_heads"]
    hd = D // nh
    for i in range(dims["vit_layers"]):
        q = p + f"encoder.layers.{i}."
        h = F.layer_norm(x, (D,), sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], 1e-6)
        qs = F.linear(h, sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.q_proj.bias"])
        ks = F.linear(h, sd[q + "self_attn.k_proj.weight"], sd[q + "self_attn.k_proj.bias"])
        vs = F.linear(h, sd[q + "self_attn.v_proj.weight"], sd[q + "self_attn.v_proj.bias"])
        qs, ks, vs = (t.view(B, S, nh, hd).transpose(1, 2) for t in (qs, ks, vs))
        w = torch.matmul(qs, ks.transpose(2, 3)) * hd ** -0.5          # siglip.py:133-135
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(qs.dtype)     # :144-146
        a = torch.matmul(w, vs).transpose(1, 2).reshape(B, S, D)
        a = F.linear(a, sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"])
        x = x + a
        h = F.layer_norm(x, (D,), sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], 1e-6)
        h = F.linear(h, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"])
        h = F.gelu(h, approximate="tanh")
        h = F.linear(h, sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
        x = x + h
        if capture is not None:
            capture.setdefault("vit_layers", []).append(x.clone())
    return F.layer_norm(x, (D,), sd[p + "post_layernorm.weight"], sd[p + "post_layernorm.bias"], 1e-6)


def embed_prefix(sd, dims, input_ids, pixel_values, capture=None):
    """pizero.py:376-414 -- text embedding gather, SigLIP + projector,
    `/ sqrt(hidden)`, scatter into a `[B,S_v,hidden]` tensor filled with
    pad_token_id (= 0).  Multi-image: pixel_values
    `[B,n_img,3,H,W]` -> vision tower on `[B*n_img,...]`, features laid out
    image after image."""
    dtype = pixel_values.dtype
    B, S = input_ids.shape
    if pixel_values.dim() == 5:
        n_img = pixel_values.shape[1]
        pixel_values = pixel_values.reshape(B * n_img, *pixel_values.shape[2:])
    else:
        n_img = 1
    tok = F.embedding(input_ids, sd["embed_tokens.weight"])
    vit = siglip_forward(sd, dims, pixel_values, capture)
    feats = F.linear(vit, sd["multi_modal_projector.linear.weight"],
                     sd["multi_modal_projector.linear.bias"])
    if capture is not None:
        capture["vit_out"] = vit.clone()
        capture["image_features"] = feats.clone()
    feats = feats / (dims["vlm_hidden"] ** 0.5)
    feats = feats.reshape(B, n_img * feats.shape[1], feats.shape[2])
    out = torch.full((B, S, feats.shape[-1]), dims["pad_token_id"], dtype=dtype)
    is_img = input_ids == dims["image_token_index"]
    is_txt = (~is_img) & (input_ids != dims["pad_token_id"])
    out[is_txt] = tok[is_txt]
    for b in range(B):
        idx = is_img[b].nonzero(as_tuple=True)[0]
        out[b, idx] = feats[b, : len(idx)]
    return out


# --------------------------------------------------------------------------
# masks / positions
# --------------------------------------------------------------------------
def build_masks_and_positions(dims, attention_mask, dtype):
    """pizero.py:271-336 -- block mask (finfo.min / 0), positions start at 1,
    then the two sub-masks `infer_action` consumes."""
    B = attention_mask.shape[0]
    Sv, Sp, H = dims["max_image_text_tokens"], dims["cond_steps"], dims["horizon_steps"]
    S = Sv + Sp + H
    cnt = attention_mask.sum(dim=1)
    m = torch.full((B, S, S), torch.finfo(dtype).min, dtype=dtype)
    for b in range(B):
        c = int(cnt[b])
        m[b, :c, :c] = 0
        m[b, Sv:, :c] = 0
    m[:, Sv:Sv + Sp, Sv:Sv + Sp] = 0
    m[:, Sv + Sp:, Sv:] = 0
    m = m[:, None]
    pos = dict(
        vlm=torch.arange(1, Sv + 1).repeat(B, 1),
        proprio=torch.arange(1, Sp + 1).repeat(B, 1),
        action=torch.arange(Sp + 1, Sp + H + 1).repeat(B, 1),
    )
    prefix_mask = m[..., : Sv + Sp, : Sv + Sp]
    action_mask = m[..., -H:, :]
    return m, prefix_mask, action_mask, pos


# --------------------------------------------------------------------------
# joint mixture-of-transformers
# --------------------------------------------------------------------------
def _theta(dims, name):
    return dims["vlm_rope_theta"] if name == "vlm" else dims["act_rope_theta"]


def joint_forward(sd, dims, attention_mask, position_ids, embeds, kv_caches, *,
                  skip_last=("vlm", "proprio"), capture=None, capture_key=None):
    """joint_model.py:328-383 in `append_non_active` mode.

    `embeds` is an ordered dict name -> [B,S_x,hidden_x] of the *active*
    mixtures.  `kv_caches` is name -> list of per-layer (K,V); a cached
    mixture that is active and lacks the layer gets it filled (prefix pass,
    joint_model.py:195-219); a cached mixture that is not active contributes
    its K,V first, in dict order (denoise pass, joint_model.py:164-168)."""
    L, nh, nkv, hd = dims["num_layers"], dims["num_heads"], dims["num_kv_heads"], dims["head_dim"]
    names = list(embeds.keys())
    x = {}
    for n in names:  # joint_model.py:348-355 (normaliser rounded to the model dtype)
        e = embeds[n]
        x[n] = e * torch.tensor(e.shape[-1] ** 0.5, dtype=e.dtype)
    B = x[names[0]].shape[0]
    for li in range(L):
        last = li == L - 1
        skip = tuple(skip_last) if last else ()
        pre = "joint_model.mixtures.{}.layers.%d." % li
        q_all, k_all, v_all = {}, {}, {}
        for n, cache in kv_caches.items():          # non-active cached mixtures first
            if n not in names:
                k_all[n], v_all[n] = cache[li]
        for n in names:
            p = pre.format(n)
            h = gemma_rms_norm(x[n], sd[p + "input_layernorm.weight"])      # joint_model.py:41-48
            S = h.shape[1]
            q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(B, S, nh, hd).transpose(1, 2)
            k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
            v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
            cos, sin = rope_cos_sin(position_ids[n], hd, _theta(dims, n), h.dtype)
            k = apply_rope(k, cos, sin)                                     # K cached post-RoPE
            q = apply_rope(q, cos, sin)
            if n in kv_caches and len(kv_caches[n]) <= li:
                kv_caches[n].append((k, v))
            q_all[n], k_all[n], v_all[n] = q, k, v
        rep = nh // nkv
        K = torch.cat([k_all[n].repeat_interleave(rep, dim=1) for n in k_all], dim=2)
        V = torch.cat([v_all[n].repeat_interleave(rep, dim=1) for n in v_all], dim=2)
        Q = torch.cat([q_all[n] for n in names], dim=2)
        w = torch.matmul(Q, K.transpose(2, 3)) / math.sqrt(hd)              # joint_model.py:261-263
        w = torch.tanh(w / ATTN_SOFTCAP) * ATTN_SOFTCAP                     # :266-268
        w = w + attention_mask                                              # :271
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(Q.dtype)           # :273-275
        o = torch.matmul(w, V).transpose(1, 2).reshape(B, Q.shape[2], nh * hd)
        outs = dict(zip(names, torch.split(o, [x[n].shape[1] for n in names], dim=1)))
        new_x = {}
        for n in names:
            if n in skip:                                                   # joint_model.py:297-299
                new_x[n] = None
                continue
            p = pre.format(n)
            x1 = x[n] + F.linear(outs[n], sd[p + "self_attn.o_proj.weight"])
            h2 = gemma_rms_norm(x1, sd[p + "post_attention_layernorm.weight"])
            new_x[n] = x1 + gemma_mlp(h2, sd[p + "mlp.gate_proj.weight"],
                                      sd[p + "mlp.up_proj.weight"], sd[p + "mlp.down_proj.weight"])
        x = new_x
        if capture is not None:
            capture.setdefault(capture_key, []).append(
                {n: (None if t is None else t.clone()) for n, t in x.items()})
    out = {}
    for n in names:                                                         # joint_model.py:375-380
        if n not in skip_last:
            key = f"joint_model.mixtures.{n}.norm.weight"
            out[n] = gemma_rms_norm(x[n], sd[key]) if key in sd else None
    return out


# --------------------------------------------------------------------------
# the hot call
# --------------------------------------------------------------------------
@torch.no_grad()
def infer_action(sd, dims, input_ids, pixel_values, attention_mask, proprios, noise,
                 capture=None):
    """pizero.py:416-490.  Returns the clamped action `[B,H,A]`; when `capture`
    is a dict it is filled with: vit_layers, vit_out, image_features,
    prefix_embeds, prefix_layers (list of {vlm,proprio}), kv (name -> list of
    (K,V)), denoise_layers (list over steps x layers of {action}), velocities,
    action_preclip, action."""
    dtype = pixel_values.dtype
    B = input_ids.shape[0]
    _, prefix_mask, action_mask, pos = build_masks_and_positions(dims, attention_mask, dtype)
    kv = {"vlm": [], "proprio": []}                                         # joint_model.py:325
    emb = embed_prefix(sd, dims, input_ids, pixel_values, capture)
    if capture is not None:
        capture["prefix_embeds"] = emb.clone()
    pe = F.linear(proprios, sd["proprio_encoder.weight"], sd["proprio_encoder.bias"])
    joint_forward(sd, dims, prefix_mask, {"vlm": pos["vlm"], "proprio": pos["proprio"]},
                  {"vlm": emb, "proprio": pe}, kv, capture=capture, capture_key="prefix_layers")
    if capture is not None:
        capture["kv"] = kv
    action = noise.to(dtype).clone()
    n_steps = dims["num_inference_steps"]
    dt = 1.0 / n_steps
    t = torch.zeros(B, dtype=dtype)
    for _ in range(n_steps):
        temb = sinusoidal_time_embedding(t, dims["act_hidden"], dims["time_max_period"])
        ae = action_encoder(sd, action, temb)
        h = joint_forward(sd, dims, action_mask, {"action": pos["action"]}, {"action": ae}, kv,
                          capture=capture, capture_key="denoise_layers")["action"]
        vel = F.linear(h, sd["action_decoder.weight"], sd["action_decoder.bias"])
        if capture is not None:
            capture.setdefault("velocities", []).append(vel.clone())
        action += dt * vel
        t += dt
    if capture is not None:
        capture["action_preclip"] = action.clone()
    clip = dims["final_action_clip_value"]
    if clip is not None:
        action = torch.clamp(action, -clip, clip)
    if capture is not None:
        capture["action"] = action.clone()
    return action


# --------------------------------------------------------------------------
# text output
# --------------------------------------------------------------------------
@torch.no_grad()
def infer_text(sd, dims, input_ids, pixel_values, attention_mask, kv_cache=None):
    """pizero.py:559-593 (`infer_text`) with its mask / position builder (:338-372; the batch size it needs is taken from
    `attention_mask` -- the reference reads an undefined global `bsz`; tests/golden/text_tiny.pt was produced
    by the unmodified reference with that one name injected into its module).  Only the vlm mixture is active, in
    `cache_mode="append"` (joint_model.py:164-240): nothing is masked, every layer runs to the end
    (`final_layer_post_attn_skip_names=[]`), then the vlm mixture's final norm and the tied lm_head.
    `kv_cache`: None, or a list of per-layer (K, V) `[B, 1, S, hd]` that is extended in place.  Returns {"logits"}."""
    L, nh, nkv, hd = dims["num_layers"], dims["num_heads"], dims["num_kv_heads"], dims["head_dim"]
    B, q_len = input_ids.shape
    cached = kv_cache is not None and len(kv_cache) > 0
    if cached:
        assert q_len == 1, "Using KV cache so should only use one single token"
        position_ids = attention_mask.cumsum(-1)[:, -1:]
    else:
        position_ids = attention_mask.cumsum(-1).masked_fill(attention_mask == 0, 1)
    embeds = embed_prefix(sd, dims, input_ids, pixel_values)      # pizero.py:569 (no image token in a decode step: token embedding)
    x = embeds * torch.tensor(embeds.shape[-1] ** 0.5, dtype=embeds.dtype)   # joint_model.py:348-355
    for li in range(L):
        p = "joint_model.mixtures.vlm.layers.%d." % li
        h = gemma_rms_norm(x, sd[p + "input_layernorm.weight"])
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(B, q_len, nh, hd).transpose(1, 2)
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(B, q_len, nkv, hd).transpose(1, 2)
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(B, q_len, nkv, hd).transpose(1, 2)
        cos, sin = rope_cos_sin(position_ids, hd, dims["vlm_rope_theta"], h.dtype)
        k = apply_rope(k, cos, sin)
        q = apply_rope(q, cos, sin)
        if kv_cache is not None:                                   # joint_model.py:195-240: append, then use old + new
            if len(kv_cache) > li:
                k = torch.cat((kv_cache[li][0], k), dim=-2)
                v = torch.cat((kv_cache[li][1], v), dim=-2)
                kv_cache[li] = (k, v)
            else:
                kv_cache.append((k, v))
        rep = nh // nkv
        K, V = k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1)
        w = torch.matmul(q, K.transpose(2, 3)) / math.sqrt(hd)
        w = torch.tanh(w / ATTN_SOFTCAP) * ATTN_SOFTCAP
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)  # the text mask is all zeros
        o = torch.matmul(w, V).transpose(1, 2).reshape(B, q_len, nh * hd)
        x1 = x + F.linear(o, sd[p + "self_attn.o_proj.weight"])
        h2 = gemma_rms_norm(x1, sd[p + "post_attention_layernorm.weight"])
        x = x1 + gemma_mlp(h2, sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"], sd[p + "mlp.down_proj.weight"])
    out = gemma_rms_norm(x, sd["joint_model.mixtures.vlm.norm.weight"])
    return {"logits": F.linear(out, sd.get("lm_head.weight", sd["embed_tokens.weight"]))}


@torch.no_grad()
def infer_action_naive(sd, dims, input_ids, pixel_values, attention_mask, proprios, noise):
    """pizero.py:492-557 -- no KV cache: every Euler step re-runs the whole
    joint model with all three mixtures active under the full block mask.
    Used only to cross-check `infer_action`."""
    dtype = pixel_values.dtype
    B = input_ids.shape[0]
    full_mask, _, _, pos = build_masks_and_positions(dims, attention_mask, dtype)
    emb = embed_prefix(sd, dims, input_ids, pixel_values)
    pe = F.linear(proprios, sd["proprio_encoder.weight"], sd["proprio_encoder.bias"])
    action = noise.to(dtype).clone()
    n_steps = dims["num_inference_steps"]
    dt = 1.0 / n_steps
    t = torch.zeros(B, dtype=dtype)
    for _ in range(n_steps):
        temb = sinusoidal_time_embedding(t, dims["act_hidden"], dims["time_max_period"])
        ae = action_encoder(sd, action, temb)
        h = joint_forward(sd, dims, full_mask, pos,
                          {"vlm": emb.clone(), "proprio": pe.clone(), "action": ae}, {})["action"]
        vel = F.linear(h, sd["action_decoder.weight"], sd["action_decoder.bias"])
        action += dt * vel
        t += dt
    clip = dims["final_action_clip_value"]
    if clip is not None:
        action = torch.clamp(action, -clip, clip)
    return action


# --------------------------------------------------------------------------
# flow-matching training forward (value only)
# --------------------------------------------------------------------------
@torch.no_grad()
def flow_matching_loss(sd, dims, input_ids, pixel_values, attention_mask, proprios, actions, t, noise,
                       capture=None):
    """pizero.py:597-661 (`psi_t` + `PiZero.forward`): one joint pass with all three
    mixtures active under the full block mask, no KV cache; returns the scalar loss.
    `noise` is x0 (the reference draws it with torch.randn_like, pizero.py:622).
    `capture` (dict) receives psi_t, v_psi and d_psi."""
    dtype = pixel_values.dtype
    sig_min = dims.get("flow_sig_min", 0.001)
    full_mask, _, _, pos = build_masks_and_positions(dims, attention_mask, dtype)
    x0, x1 = noise.to(dtype), actions.to(dtype)
    tt = t.to(dtype)[:, None, None]
    psi = (1 - (1 - sig_min) * tt) * x0 + tt * x1                            # pizero.py:597-605
    emb = embed_prefix(sd, dims, input_ids, pixel_values)
    pe = F.linear(proprios, sd["proprio_encoder.weight"], sd["proprio_encoder.bias"])
    temb = sinusoidal_time_embedding(t.to(dtype), dims["act_hidden"], dims["time_max_period"])
    ae = action_encoder(sd, psi, temb)
    h = joint_forward(sd, dims, full_mask, pos, {"vlm": emb, "proprio": pe, "action": ae}, {})["action"]
    v_psi = F.linear(h, sd["action_decoder.weight"], sd["action_decoder.bias"])
    d_psi = x1 - (1 - sig_min) * x0                                          # pizero.py:659
    if capture is not None:
        capture.update(psi_t=psi.clone(), v_psi=v_psi.clone(), d_psi=d_psi.clone())
    return torch.mean((v_psi - d_psi) ** 2)


def flow_matching_loss_and_grads(sd, dims, input_ids, pixel_values, attention_mask, proprios, actions, t, noise):
    """The training step's backward, as autograd of the restatement above (the checker for the backward kernels of
    SURVEY 8f-1; the reference's own backward is `normalized_loss.backward()`, train.py:355-368).  Returns
    (loss, {key: d loss / d parameter}); parameters the loss does not depend on (the vlm / proprio post-attention
    half of the last layer, joint_model.py:297-299; `embed_tokens` rows never looked up) get zeros."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        loss = flow_matching_loss.__wrapped__(leaf, dims, input_ids, pixel_values, attention_mask, proprios, actions, t, noise)
        loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return loss.detach(), grads
