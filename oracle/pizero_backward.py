"""TEST INFRASTRUCTURE ONLY -- the backward pass of the flow-matching training step
(`PiZero.forward` + `loss.backward()`, pizero.py:607-661 / train.py:350-368) written
out op by op, without autograd: the statement the CUDA backward kernels of SURVEY 8f-1
will follow.

  flow_matching_backward       loss, `action_decoder`, the joint mixture-of-transformers
                               (all three mixtures active, full block mask, no cache),
                               `action_encoder`; returns d loss / d embeds for the rest
  flow_matching_backward_full  + `proprio_encoder`, the embedding merge / token embedding,
                               projector and SigLIP (patch matmul, LayerNorm, attention, MLP)

Each formula is the derivative of the forward statement in `oracle/pizero_oracle.py`
(which cites the reference lines); `tests/test_flow_matching.py` checks the result
against d loss / d parameter of the unmodified reference's own `loss.backward()`
(tests/golden/fm_tiny.pt: every tensor; fm_width2.pt: norms and leading rows at the real
widths).  fp32, CPU, small cases.
"""
import math

import torch
import torch.nn.functional as F

from oracle import pizero_oracle as O

EPS = 1e-6


# ---------------------------------------------------------------- primitives
def linear_bwd(x, w, gy, bias=False):
    """y = x w^T (+ b): gx = gy w, gw = gy^T x, gb = sum gy."""
    gx = gy @ w
    gw = gy.reshape(-1, gy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])
    gb = gy.reshape(-1, gy.shape[-1]).sum(0) if bias else None
    return gx, gw, gb


def rms_norm_fwd(x, w):
    r = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + EPS)
    return x * r * (1.0 + w), r


def rms_norm_bwd(x, w, r, gy):
    """y = x r (1 + w), r = (mean x^2 + eps)^-1/2 (paligemma/modules.py:13-21):
    gx = r (1+w) gy - x r^3 mean(x (1+w) gy);  gw = sum_rows gy x r."""
    gs = gy * (1.0 + w)
    gx = r * gs - x * r.pow(3) * (x * gs).mean(-1, keepdim=True)
    gw = (gy * x * r).reshape(-1, x.shape[-1]).sum(0)
    return gx, gw


def rope_bwd(gy, cos, sin):
    """y = x cos + rot(x) sin with rot(x) = cat(-x2, x1) (model/utils.py:4-16): gx = gy cos + rot^T(gy sin),
    rot^T(z) = cat(z2, -z1)."""
    half = gy.shape[-1] // 2
    z = gy * sin[:, None]
    return gy * cos[:, None] + torch.cat((z[..., half:], -z[..., :half]), dim=-1)


def gelu_tanh_grad(x):
    k0, k1 = 0.7978845608028654, 0.044715
    u = k0 * (x + k1 * x ** 3)
    th = torch.tanh(u)
    return 0.5 * (1 + th) + 0.5 * x * (1 - th * th) * k0 * (1 + 3 * k1 * x * x)


def silu_grad(x):
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s))


# ---------------------------------------------------------------- joint model
def joint_forward_saved(sd, dims, attention_mask, position_ids, embeds, skip_last=("vlm", "proprio")):
    """`pizero_oracle.joint_forward` (joint_model.py:328-383, every mixture active, no cache) keeping what the backward
    needs.  Returns ({name: final hidden}, saved)."""
    L, nh, nkv, hd = dims["num_layers"], dims["num_heads"], dims["num_kv_heads"], dims["head_dim"]
    names = list(embeds.keys())
    x = {n: embeds[n] * (embeds[n].shape[-1] ** 0.5) for n in names}
    B = x[names[0]].shape[0]
    rep = nh // nkv
    layers = []
    rope = {n: O.rope_cos_sin(position_ids[n], hd, O._theta(dims, n), torch.float32) for n in names}
    for li in range(L):
        skip = tuple(skip_last) if li == L - 1 else ()
        pre = "joint_model.mixtures.{}.layers.%d." % li
        sv = dict(x=dict(x), skip=skip, r_in={}, h={}, q={}, k={}, v={})
        for n in names:
            p = pre.format(n)
            h, r = rms_norm_fwd(x[n], sd[p + "input_layernorm.weight"])
            S = h.shape[1]
            q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(B, S, nh, hd).transpose(1, 2)
            k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
            v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
            cos, sin = rope[n]
            sv["r_in"][n], sv["h"][n] = r, h
            sv["q"][n], sv["k"][n], sv["v"][n] = O.apply_rope(q, cos, sin), O.apply_rope(k, cos, sin), v
        Q = torch.cat([sv["q"][n] for n in names], dim=2)
        K = torch.cat([sv["k"][n].repeat_interleave(rep, dim=1) for n in names], dim=2)
        V = torch.cat([sv["v"][n].repeat_interleave(rep, dim=1) for n in names], dim=2)
        s = torch.matmul(Q, K.transpose(2, 3)) / math.sqrt(hd)
        th = torch.tanh(s / O.ATTN_SOFTCAP)
        pr = F.softmax(th * O.ATTN_SOFTCAP + attention_mask, dim=-1)
        o = torch.matmul(pr, V).transpose(1, 2).reshape(B, Q.shape[2], nh * hd)
        sv.update(Q=Q, K=K, V=V, th=th, pr=pr)
        sizes = [x[n].shape[1] for n in names]
        outs = dict(zip(names, torch.split(o, sizes, dim=1)))
        sv["o"] = outs
        new_x = {}
        sv.update(x1={}, r_post={}, h2={}, g={}, u={}, m={})
        for n in names:
            if n in skip:
                new_x[n] = None
                continue
            p = pre.format(n)
            x1 = x[n] + F.linear(outs[n], sd[p + "self_attn.o_proj.weight"])
            h2, r2 = rms_norm_fwd(x1, sd[p + "post_attention_layernorm.weight"])
            g = F.linear(h2, sd[p + "mlp.gate_proj.weight"])
            u = F.linear(h2, sd[p + "mlp.up_proj.weight"])
            m = F.gelu(g, approximate="tanh") * u
            new_x[n] = x1 + F.linear(m, sd[p + "mlp.down_proj.weight"])
            sv["x1"][n], sv["r_post"][n], sv["h2"][n], sv["g"][n], sv["u"][n], sv["m"][n] = x1, r2, h2, g, u, m
        layers.append(sv)
        x = new_x
    out, fin = {}, {}
    for n in names:
        key = f"joint_model.mixtures.{n}.norm.weight"
        if n not in skip_last and key in sd:
            out[n], fin[n] = rms_norm_fwd(x[n], sd[key])
    saved = dict(layers=layers, names=names, x_final=x, r_final=fin, rope=rope, sizes=[embeds[n].shape[1] for n in names])
    return out, saved


def joint_backward(sd, dims, saved, g_out):
    """Backward of `joint_forward_saved`.  `g_out`: {name: d loss / d final hidden}.  Returns
    ({parameter key: grad}, {name: d loss / d embeds[name]})."""
    L, nh, nkv, hd = dims["num_layers"], dims["num_heads"], dims["num_kv_heads"], dims["head_dim"]
    names, rep = saved["names"], nh // nkv
    grads = {}
    gx = {n: None for n in names}
    for n, g in g_out.items():
        key = f"joint_model.mixtures.{n}.norm.weight"
        gx[n], grads[key] = rms_norm_bwd(saved["x_final"][n], sd[key], saved["r_final"][n], g)
    for li in reversed(range(L)):
        sv = saved["layers"][li]
        pre = "joint_model.mixtures.{}.layers.%d." % li
        B = sv["Q"].shape[0]
        g_o = {}
        g_x = {}
        for n in names:
            S = sv["x"][n].shape[1]
            if n in sv["skip"] or gx[n] is None:
                # post-attention half discarded (joint_model.py:297-299): nothing flows into its attention output
                g_o[n] = torch.zeros(B, S, nh * hd)
                g_x[n] = torch.zeros_like(sv["x"][n])
                continue
            p = pre.format(n)
            g_xn = gx[n]                                           # d / d x_next = x1 + down(m)
            g_m, grads[p + "mlp.down_proj.weight"], _ = linear_bwd(sv["m"][n], sd[p + "mlp.down_proj.weight"], g_xn)
            g_g = g_m * sv["u"][n] * gelu_tanh_grad(sv["g"][n])   # m = gelu(g) u
            g_u = g_m * F.gelu(sv["g"][n], approximate="tanh")
            a, grads[p + "mlp.gate_proj.weight"], _ = linear_bwd(sv["h2"][n], sd[p + "mlp.gate_proj.weight"], g_g)
            b, grads[p + "mlp.up_proj.weight"], _ = linear_bwd(sv["h2"][n], sd[p + "mlp.up_proj.weight"], g_u)
            g_x1n, grads[p + "post_attention_layernorm.weight"] = rms_norm_bwd(
                sv["x1"][n], sd[p + "post_attention_layernorm.weight"], sv["r_post"][n], a + b)
            g_x1 = g_xn + g_x1n                                    # residual
            g_o[n], grads[p + "self_attn.o_proj.weight"], _ = linear_bwd(sv["o"][n], sd[p + "self_attn.o_proj.weight"], g_x1)
            g_x[n] = g_x1                                          # x1 = x + o_proj(o): identity branch
        # attention: o = pr V, pr = softmax(cap tanh(s / cap) + mask), s = Q K^T / sqrt(hd)   (joint_model.py:255-283)
        g_oall = torch.cat([g_o[n] for n in names], dim=1).view(B, -1, nh, hd).transpose(1, 2)   # [B, nh, S, hd]
        g_V = torch.matmul(sv["pr"].transpose(2, 3), g_oall)
        g_pr = torch.matmul(g_oall, sv["V"].transpose(2, 3))
        g_c = sv["pr"] * (g_pr - (g_pr * sv["pr"]).sum(-1, keepdim=True))      # softmax
        g_s = g_c * (1.0 - sv["th"] * sv["th"])                                 # cap tanh(s / cap)
        g_Q = torch.matmul(g_s, sv["K"]) / math.sqrt(hd)
        g_K = torch.matmul(g_s.transpose(2, 3), sv["Q"]) / math.sqrt(hd)
        off = 0
        for n, S in zip(names, saved["sizes"]):
            p = pre.format(n)
            cos, sin = saved["rope"][n]
            gq = rope_bwd(g_Q[:, :, off:off + S], cos, sin)                                         # [B, nh, S, hd]
            gk = rope_bwd(g_K[:, :, off:off + S].reshape(B, nkv, rep, S, hd).sum(2), cos, sin)      # repeat_kv backward
            gv = g_V[:, :, off:off + S].reshape(B, nkv, rep, S, hd).sum(2)
            off += S
            h = sv["h"][n]
            a, grads[p + "self_attn.q_proj.weight"], _ = linear_bwd(h, sd[p + "self_attn.q_proj.weight"], gq.transpose(1, 2).reshape(B, S, nh * hd))
            b, grads[p + "self_attn.k_proj.weight"], _ = linear_bwd(h, sd[p + "self_attn.k_proj.weight"], gk.transpose(1, 2).reshape(B, S, nkv * hd))
            c, grads[p + "self_attn.v_proj.weight"], _ = linear_bwd(h, sd[p + "self_attn.v_proj.weight"], gv.transpose(1, 2).reshape(B, S, nkv * hd))
            g_in, grads[p + "input_layernorm.weight"] = rms_norm_bwd(sv["x"][n], sd[p + "input_layernorm.weight"], sv["r_in"][n], a + b + c)
            gx[n] = g_x[n] + g_in
    g_embeds = {n: gx[n] * (gx[n].shape[-1] ** 0.5) for n in names}   # embeds * sqrt(hidden), joint_model.py:348-355
    return grads, g_embeds


# ---------------------------------------------------------------- the training step
def flow_matching_backward(sd, dims, input_ids, pixel_values, attention_mask, proprios, actions, t, noise):
    """loss and its gradients w.r.t. the parameters downstream of the joint model's inputs (`action_decoder`,
    `joint_model.*`, `action_encoder`), plus d loss / d (vlm embeds, proprio embeds) for everything upstream."""
    sig_min = dims.get("flow_sig_min", 0.001)
    full_mask, _, _, pos = O.build_masks_and_positions(dims, attention_mask, torch.float32)
    x0, x1 = noise.float(), actions.float()
    tt = t.float()[:, None, None]
    psi = (1 - (1 - sig_min) * tt) * x0 + tt * x1
    with torch.no_grad():
        emb = O.embed_prefix(sd, dims, input_ids, pixel_values)
        pe = F.linear(proprios, sd["proprio_encoder.weight"], sd["proprio_encoder.bias"])
        temb = O.sinusoidal_time_embedding(t.float(), dims["act_hidden"], dims["time_max_period"])
        # action encoder forward (vla/modules.py:39-53), keeping the intermediates
        w1, b1 = sd["action_encoder.linear_1.weight"], sd["action_encoder.linear_1.bias"]
        w2, b2 = sd["action_encoder.linear_2.weight"], sd["action_encoder.linear_2.bias"]
        w3, b3 = sd["action_encoder.linear_3.weight"], sd["action_encoder.linear_3.bias"]
        e1 = F.linear(psi, w1, b1)
        cat = torch.cat([temb[:, None, :].expand(-1, psi.shape[1], -1), e1], dim=-1)
        z_pre = F.linear(cat, w2, b2)
        z = F.silu(z_pre)
        ae = F.linear(z, w3, b3)
        out, saved = joint_forward_saved(sd, dims, full_mask, pos, {"vlm": emb, "proprio": pe, "action": ae})
        hfin = out["action"]
        wd, bd = sd["action_decoder.weight"], sd["action_decoder.bias"]
        v_psi = F.linear(hfin, wd, bd)
        d_psi = x1 - (1 - sig_min) * x0
        diff = v_psi - d_psi
        loss = (diff ** 2).mean()
        grads = {}
        g_v = 2.0 * diff / diff.numel()                                               # mean squared error
        g_h, grads["action_decoder.weight"], grads["action_decoder.bias"] = linear_bwd(hfin, wd, g_v, bias=True)
        jg, g_emb = joint_backward(sd, dims, saved, {"action": g_h})
        grads.update(jg)
        g_z, grads["action_encoder.linear_3.weight"], grads["action_encoder.linear_3.bias"] = linear_bwd(z, w3, g_emb["action"], bias=True)
        g_cat, grads["action_encoder.linear_2.weight"], grads["action_encoder.linear_2.bias"] = linear_bwd(cat, w2, g_z * silu_grad(z_pre), bias=True)
        A = dims["act_hidden"]
        _, grads["action_encoder.linear_1.weight"], grads["action_encoder.linear_1.bias"] = linear_bwd(psi, w1, g_cat[..., A:], bias=True)
    return loss, grads, {"vlm": g_emb["vlm"], "proprio": g_emb["proprio"]}


# ---------------------------------------------------------------- upstream of the joint model
def layer_norm_fwd(x, w, b, eps=1e-6):
    mu = x.mean(-1, keepdim=True)
    xc = x - mu
    rstd = torch.rsqrt(xc.pow(2).mean(-1, keepdim=True) + eps)
    xh = xc * rstd
    return xh * w + b, xh, rstd


def layer_norm_bwd(xh, rstd, w, gy):
    """y = xh w + b, xh = (x - mean) rstd: gx = rstd (g - mean g - xh mean(g xh)), g = gy w."""
    g = gy * w
    gx = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    D = xh.shape[-1]
    return gx, (gy * xh).reshape(-1, D).sum(0), gy.reshape(-1, D).sum(0)


def siglip_forward_saved(sd, dims, pixel_values):
    """`pizero_oracle.siglip_forward` (siglip.py:59-78, 220-238, 298) with the patch convolution as a matrix product over
    unfolded patches (what the CUDA path does) and everything the backward needs."""
    p = "vision_tower.vision_model."
    ps = dims["patch_size"]
    wconv = sd[p + "embeddings.patch_embedding.weight"]
    D = wconv.shape[0]
    patches = F.unfold(pixel_values, kernel_size=ps, stride=ps).transpose(1, 2)           # [B, S, 3*ps*ps], (c, ky, kx) order
    x = patches @ wconv.reshape(D, -1).t() + sd[p + "embeddings.patch_embedding.bias"]
    x = x + sd[p + "embeddings.position_embedding.weight"][None]
    B, S, _ = x.shape
    nh = dims["vit_heads"]
    hd = D // nh
    layers = []
    for i in range(dims["vit_layers"]):
        q = p + f"encoder.layers.{i}."
        sv = dict(x=x)
        h, sv["xh1"], sv["rstd1"] = layer_norm_fwd(x, sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"])
        sv["h1"] = h
        qs, ks, vs = (F.linear(h, sd[q + f"self_attn.{n}_proj.weight"], sd[q + f"self_attn.{n}_proj.bias"])
                      .view(B, S, nh, hd).transpose(1, 2) for n in "qkv")
        pr = F.softmax(torch.matmul(qs, ks.transpose(2, 3)) * hd ** -0.5, dim=-1)
        a = torch.matmul(pr, vs).transpose(1, 2).reshape(B, S, D)
        sv.update(q=qs, k=ks, v=vs, pr=pr, a=a)
        x = x + F.linear(a, sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"])
        sv["x_mid"] = x
        h, sv["xh2"], sv["rstd2"] = layer_norm_fwd(x, sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"])
        sv["h2"] = h
        f1 = F.linear(h, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"])
        sv["f1"] = f1
        sv["act"] = F.gelu(f1, approximate="tanh")
        x = x + F.linear(sv["act"], sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
        layers.append(sv)
    y, xh, rstd = layer_norm_fwd(x, sd[p + "post_layernorm.weight"], sd[p + "post_layernorm.bias"])
    return y, dict(layers=layers, patches=patches, xh=xh, rstd=rstd, B=B, S=S, D=D, nh=nh, hd=hd)


def siglip_backward(sd, dims, saved, gy):
    p = "vision_tower.vision_model."
    B, S, D, nh, hd = saved["B"], saved["S"], saved["D"], saved["nh"], saved["hd"]
    grads = {}
    gx, grads[p + "post_layernorm.weight"], grads[p + "post_layernorm.bias"] = layer_norm_bwd(
        saved["xh"], saved["rstd"], sd[p + "post_layernorm.weight"], gy)
    for i in reversed(range(dims["vit_layers"])):
        q = p + f"encoder.layers.{i}."
        sv = saved["layers"][i]
        g_act, grads[q + "mlp.fc2.weight"], grads[q + "mlp.fc2.bias"] = linear_bwd(sv["act"], sd[q + "mlp.fc2.weight"], gx, bias=True)
        g_h2, grads[q + "mlp.fc1.weight"], grads[q + "mlp.fc1.bias"] = linear_bwd(
            sv["h2"], sd[q + "mlp.fc1.weight"], g_act * gelu_tanh_grad(sv["f1"]), bias=True)
        g_mid, grads[q + "layer_norm2.weight"], grads[q + "layer_norm2.bias"] = layer_norm_bwd(
            sv["xh2"], sv["rstd2"], sd[q + "layer_norm2.weight"], g_h2)
        gx = gx + g_mid
        g_a, grads[q + "self_attn.out_proj.weight"], grads[q + "self_attn.out_proj.bias"] = linear_bwd(
            sv["a"], sd[q + "self_attn.out_proj.weight"], gx, bias=True)
        g_a = g_a.view(B, S, nh, hd).transpose(1, 2)
        g_v = torch.matmul(sv["pr"].transpose(2, 3), g_a)
        g_pr = torch.matmul(g_a, sv["v"].transpose(2, 3))
        g_s = sv["pr"] * (g_pr - (g_pr * sv["pr"]).sum(-1, keepdim=True)) * hd ** -0.5
        g_q = torch.matmul(g_s, sv["k"])
        g_k = torch.matmul(g_s.transpose(2, 3), sv["q"])
        g_h1 = 0
        for n, g in (("q", g_q), ("k", g_k), ("v", g_v)):
            a, grads[q + f"self_attn.{n}_proj.weight"], grads[q + f"self_attn.{n}_proj.bias"] = linear_bwd(
                sv["h1"], sd[q + f"self_attn.{n}_proj.weight"], g.transpose(1, 2).reshape(B, S, D), bias=True)
            g_h1 = g_h1 + a
        g_in, grads[q + "layer_norm1.weight"], grads[q + "layer_norm1.bias"] = layer_norm_bwd(
            sv["xh1"], sv["rstd1"], sd[q + "layer_norm1.weight"], g_h1)
        gx = gx + g_in
    grads[p + "embeddings.position_embedding.weight"] = gx.sum(0)
    wconv = sd[p + "embeddings.patch_embedding.weight"]
    _, gw, gb = linear_bwd(saved["patches"], wconv.reshape(D, -1), gx, bias=True)
    grads[p + "embeddings.patch_embedding.weight"] = gw.view_as(wconv)
    grads[p + "embeddings.patch_embedding.bias"] = gb
    return grads


def flow_matching_backward_full(sd, dims, input_ids, pixel_values, attention_mask, proprios, actions, t, noise):
    """`flow_matching_backward` plus everything upstream of the joint model: proprio_encoder, the embedding merge
    (pizero.py:376-414: text rows <- embed_tokens, image rows <- projector(SigLIP) / sqrt(hidden)), projector, SigLIP.
    Returns (loss, {parameter key: grad}) for every parameter the loss depends on."""
    loss, grads, g_emb = flow_matching_backward(sd, dims, input_ids, pixel_values, attention_mask, proprios, actions, t, noise)
    with torch.no_grad():
        A, H = dims["act_hidden"], dims["vlm_hidden"]
        _, grads["proprio_encoder.weight"], grads["proprio_encoder.bias"] = linear_bwd(proprios.float(), sd["proprio_encoder.weight"],
                                                                                       g_emb["proprio"], bias=True)
        gv = g_emb["vlm"]                                                 # [B, S_v, H]
        is_img = input_ids == dims["image_token_index"]
        is_txt = (~is_img) & (input_ids != dims["pad_token_id"])
        ge = torch.zeros_like(sd["embed_tokens.weight"])
        ge.index_add_(0, input_ids[is_txt], gv[is_txt])                   # F.embedding backward (text rows only)
        grads["embed_tokens.weight"] = ge
        B = input_ids.shape[0]
        vit, saved = siglip_forward_saved(sd, dims, pixel_values)
        n_img_tok = vit.shape[1]
        g_feat = torch.stack([gv[b][is_img[b]][:n_img_tok] for b in range(B)]) / H ** 0.5
        g_vit, grads["multi_modal_projector.linear.weight"], grads["multi_modal_projector.linear.bias"] = linear_bwd(
            vit, sd["multi_modal_projector.linear.weight"], g_feat, bias=True)
        grads.update(siglip_backward(sd, dims, saved, g_vit))
    return loss, grads
