"""Synthetic weights and inputs of the reference's shapes (there is no network
for checkpoints or datasets).

`init_state_dict` fills the reference's 938-key state dict
(`PiZero.state_dict()`) with the same *distributions* the
reference's default constructors use (nn.Linear / nn.Conv2d: U(-1/sqrt(fan_in),
1/sqrt(fan_in)) for weight and bias; nn.Embedding: N(0,1) with a zero padding
row, `pizero.py:61-65`; LayerNorm ones/zeros; GemmaRMSNorm zeros,
`paligemma/modules.py:11`) from one seeded CPU generator, so the identical
tensors can be rebuilt on the GPU box without the reference installed.
`randomize_norms=True` perturbs the norm scales/biases so that `(1+w)` and LN
affine handling is actually exercised by parity tests.

`make_inputs` follows the processor's layout (`src/model/vla/processing.py:22`,
`:109-114`): `<image>`x256 (per image), one bos-class id, n text ids, pad 0;
pixels are uint8 U{0..255} mapped through (x/255-0.5)/0.5.
"""
from __future__ import annotations

import math

import torch


def state_dict_spec(d: dict) -> list:
    """[(key, shape, kind, fan_in)] in the reference's key order."""
    spec = []

    def lin(key, out_f, in_f, bias):
        spec.append((key + ".weight", (out_f, in_f), "uniform", in_f))
        if bias:
            spec.append((key + ".bias", (out_f,), "uniform", in_f))

    H, I = d["vlm_hidden"], d["vlm_inter"]
    A, AI = d["act_hidden"], d["act_inter"]
    V, VI = d["vit_hidden"], d["vit_inter"]
    qd = d["num_heads"] * d["head_dim"]
    kvd = d["num_kv_heads"] * d["head_dim"]
    ps = d["patch_size"]
    spec.append(("embed_tokens.weight", (d["vocab_size"], H), "embedding", None))
    vp = "vision_tower.vision_model."
    spec.append((vp + "embeddings.patch_embedding.weight", (V, 3, ps, ps), "uniform", 3 * ps * ps))
    spec.append((vp + "embeddings.patch_embedding.bias", (V,), "uniform", 3 * ps * ps))
    spec.append((vp + "embeddings.position_embedding.weight", (d["num_image_tokens"], V), "normal", None))
    for i in range(d["vit_layers"]):
        p = vp + f"encoder.layers.{i}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(p + "self_attn." + nm, V, V, True)
        spec.append((p + "layer_norm1.weight", (V,), "ones", None))
        spec.append((p + "layer_norm1.bias", (V,), "zeros", None))
        lin(p + "mlp.fc1", VI, V, True)
        lin(p + "mlp.fc2", V, VI, True)
        spec.append((p + "layer_norm2.weight", (V,), "ones", None))
        spec.append((p + "layer_norm2.bias", (V,), "zeros", None))
    spec.append((vp + "post_layernorm.weight", (V,), "ones", None))
    spec.append((vp + "post_layernorm.bias", (V,), "zeros", None))
    lin("multi_modal_projector.linear", H, V, True)
    for name, hid, inter in (("vlm", H, I), ("proprio", A, AI), ("action", A, AI)):
        for i in range(d["num_layers"]):
            p = f"joint_model.mixtures.{name}.layers.{i}."
            lin(p + "self_attn.q_proj", qd, hid, False)
            lin(p + "self_attn.k_proj", kvd, hid, False)
            lin(p + "self_attn.v_proj", kvd, hid, False)
            lin(p + "self_attn.o_proj", hid, qd, False)
            lin(p + "mlp.gate_proj", inter, hid, False)
            lin(p + "mlp.up_proj", inter, hid, False)
            lin(p + "mlp.down_proj", hid, inter, False)
            spec.append((p + "input_layernorm.weight", (hid,), "rms", None))
            spec.append((p + "post_attention_layernorm.weight", (hid,), "rms", None))
        if name != "vlm" or d.get("vlm_use_final_norm", False):
            spec.append((f"joint_model.mixtures.{name}.norm.weight", (hid,), "rms", None))
    lin("action_encoder.linear_1", A, d["action_dim"], True)
    lin("action_encoder.linear_2", A, 2 * A, True)
    lin("action_encoder.linear_3", A, A, True)
    lin("proprio_encoder", A, d["proprio_dim"], True)
    lin("action_decoder", d["action_dim"], A, True)
    if d.get("use_lm_head", False):   # tied to embed_tokens.weight (pizero.py:105-112): the same tensor under a second key
        spec.append(("lm_head.weight", (d["vocab_size"], H), "tied:embed_tokens.weight", None))
    return spec


def init_state_dict(d: dict, seed: int = 42, randomize_norms: bool = False,
                    tie_proprio: bool = True, dtype=torch.float32) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind, fan_in in state_dict_spec(d):
        if tie_proprio and key.startswith("joint_model.mixtures.proprio."):
            continue
        if kind == "uniform":
            b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind in ("normal", "embedding"):
            t = torch.randn(shape, generator=g)
            if kind == "embedding":
                t[d["pad_token_id"]] = 0
        elif kind == "ones":
            t = torch.ones(shape)
            if randomize_norms:
                t = t + 0.1 * torch.randn(shape, generator=g)
        elif kind == "zeros":
            t = torch.zeros(shape)
            if randomize_norms:
                t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "rms":
            t = torch.zeros(shape)
            if randomize_norms:
                t = 0.1 * torch.randn(shape, generator=g)
        elif kind.startswith("tied:"):
            sd[key] = sd[kind[5:]]
            continue
        else:
            raise ValueError(kind)
        sd[key] = t.to(dtype)
    if tie_proprio:
        # the reference's checkpoints hold both prefixes with equal values
        # (tie_action_proprio_weights, pizero.py:262-264)
        for key in list(sd):
            if key.startswith("joint_model.mixtures.action."):
                sd[key.replace(".action.", ".proprio.", 1)] = sd[key]
    # re-order to the spec order
    return {k: sd[k] for k, _, _, _ in state_dict_spec(d)}


def make_inputs(d: dict, batch: int, seed: int = 0, min_text: int | None = None,
                max_text: int | None = None) -> dict:
    """CPU tensors: input_ids, attention_mask, pixel_values (fp32, normalised),
    pixel_u8, proprios, noise (fp32), valid_len."""
    g = torch.Generator().manual_seed(seed)
    n_img = d.get("num_images", 1)
    n_img_tok = n_img * d["num_image_tokens"]
    Sv = d["max_image_text_tokens"]
    room = Sv - n_img_tok - 1            # text ids after bos
    if max_text is None:
        max_text = room
    if min_text is None:
        min_text = min(4, max_text)
    ids = torch.full((batch, Sv), d["pad_token_id"], dtype=torch.int64)
    ids[:, :n_img_tok] = d["image_token_index"]
    ids[:, n_img_tok] = 2                # bos-class id
    n_text = torch.randint(min_text, max_text + 1, (batch,), generator=g)
    hi = min(d["vocab_size"], d["image_token_index"])
    for b in range(batch):
        n = int(n_text[b])
        ids[b, n_img_tok + 1: n_img_tok + 1 + n] = torch.randint(3, hi, (n,), generator=g)
    attn = (ids != d["pad_token_id"]).to(torch.int64)
    shape = (batch, 3, d["image_size"], d["image_size"]) if n_img == 1 else \
        (batch, n_img, 3, d["image_size"], d["image_size"])
    u8 = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    pix = (u8.float() * (1 / 255.0) - 0.5) / 0.5
    proprios = torch.rand((batch, d["cond_steps"], d["proprio_dim"]), generator=g) * 2 - 1
    gn = torch.Generator().manual_seed(seed + 1)
    noise = torch.randn((batch, d["horizon_steps"], d["action_dim"]), generator=gn)
    return dict(input_ids=ids, attention_mask=attn, pixel_values=pix, pixel_u8=u8,
                proprios=proprios, noise=noise, valid_len=attn.sum(1).to(torch.int32))


@torch.no_grad()
def fill_random_(module, d: dict, seed: int = 42) -> None:
    """In-place random init of a PiZero's parameters on their own device, with the
    same distributions as `init_state_dict` (fast path for benchmarks: no 13 GB
    host copy).  Proprio and action experts get equal values."""
    params = dict(module.named_parameters(remove_duplicate=False))
    dev = next(iter(params.values())).device
    g = torch.Generator(device=dev).manual_seed(seed)
    for key, shape, kind, fan_in in state_dict_spec(d):
        if key.startswith("joint_model.mixtures.proprio."):
            continue
        p = params[key]
        if kind == "uniform":
            b = 1.0 / math.sqrt(fan_in)
            p.copy_((torch.rand(shape, generator=g, device=dev) * 2 - 1) * b)
        elif kind in ("normal", "embedding"):
            p.copy_(torch.randn(shape, generator=g, device=dev))
            if kind == "embedding":
                p[d["pad_token_id"]] = 0
        elif kind == "ones":
            p.fill_(1.0)
        else:
            p.zero_()
    for key in params:
        if key.startswith("joint_model.mixtures.proprio."):
            params[key].copy_(params[key.replace(".proprio.", ".action.", 1)])
