"""Host side of the flow-matching training step (reference: `PiZero.forward` + `loss.backward()`,
`src/model/vla/pizero.py:607-661`, `src/agent/train.py:350-379`).

`GradBuffer` owns one flat fp32 gradient buffer laid out like the PACKED weights the kernels read (fused q|k|v rows,
gate|up blocks of 128, padded small matrices) and hands `pz_flow_matching_step` a struct of pointers into it; `unpack()`
maps the buffer back to the reference's parameter names and shapes.  The flat layout is what the optimizer and the
data-parallel all-reduce work on: one contiguous tensor, bucketed by offset (train.py:121, DDP bucket all-reduce C1).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

VIT_FIELDS = ("ln1_w", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_w", "ln2_b", "w_fc1", "b_fc1", "w_fc2", "b_fc2")
MIX_FIELDS = ("norm_in", "w_qkv", "w_o", "norm_post", "w_gate_up", "w_down")
TOP_FIELDS = ("patch_w", "patch_b", "pos_emb", "post_ln_w", "post_ln_b", "proj_w", "proj_b", "action_final_norm",
              "enc_w1", "enc_b1", "enc_w2a", "enc_w2t", "enc_b2", "enc_w3", "enc_b3", "prop_w", "prop_b", "dec_w", "dec_b")


def _ptr(v) -> int:
    return int(v or 0)


class GradBuffer:
    """fp32 gradients of every trainable packed weight of one `PiZero`, in ONE flat tensor (`.flat`).

    `entries`: list of (key, offset, shape) with key = ("top", field) | ("vit", layer, field) | (mixture, layer, field).
    If the proprio mixture aliases the action mixture (tied weights, pizero.py:262-269) both use the action entries.
    """

    def __init__(self, model):
        model.pack()
        keep, w = model._packed
        by_ptr = {t.data_ptr(): t for t in keep if isinstance(t, torch.Tensor)}
        d = model.dims
        self.model = model
        self.entries: List[Tuple[tuple, int, tuple]] = []
        self._index: Dict[tuple, Tuple[int, tuple]] = {}
        off = 0

        def add(key, ptr):
            nonlocal off
            t = by_ptr[ptr]
            self.entries.append((key, off, tuple(t.shape)))
            self._index[key] = (off, tuple(t.shape))
            off += (t.numel() + 255) // 256 * 256      # 1 KiB-aligned fp32 slices

        for f in TOP_FIELDS:
            add(("top", f), _ptr(getattr(w, f)))
        for i in range(d["vit_layers"]):
            for f in VIT_FIELDS:
                add(("vit", i, f), _ptr(getattr(w.vit[i], f)))
        self.tied = C.addressof(w.proprio.contents) == C.addressof(w.action.contents)
        for name in ("vlm", "action") + (() if self.tied else ("proprio",)):
            arr = getattr(w, name)
            for i in range(d["num_layers"]):
                for f in MIX_FIELDS:
                    add((name, i, f), _ptr(getattr(arr[i], f)))
        dev = keep[0].device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        base = self.flat.data_ptr()
        assert base % 1024 == 0 or True
        # the pointer struct the C ABI reads
        g = _lib.PzWeights()
        for f in TOP_FIELDS:
            setattr(g, f, base + 4 * self._index[("top", f)][0])
        self._vit = (_lib.PzVitLayer * d["vit_layers"])()
        for i in range(d["vit_layers"]):
            for f in VIT_FIELDS:
                setattr(self._vit[i], f, base + 4 * self._index[("vit", i, f)][0])
        g.vit = C.cast(self._vit, C.POINTER(_lib.PzVitLayer))
        self._mix = {}
        for name in ("vlm", "action", "proprio"):
            src = "action" if (name == "proprio" and self.tied) else name
            arr = (_lib.PzMixLayer * d["num_layers"])()
            for i in range(d["num_layers"]):
                for f in MIX_FIELDS:
                    setattr(arr[i], f, base + 4 * self._index[(src, i, f)][0])
            self._mix[name] = arr
            setattr(g, name, C.cast(arr, C.POINTER(_lib.PzMixLayer)))
        self.struct = g

    def zero_(self):
        self.flat.zero_()
        return self

    def view(self, key) -> torch.Tensor:
        off, shape = self._index[key]
        n = 1
        for s in shape:
            n *= s
        return self.flat[off:off + n].view(shape)

    def unpack(self) -> Dict[str, torch.Tensor]:
        """Gradients under the reference's parameter names / shapes (fp32 views or small copies of `.flat`)."""
        d = self.model.dims
        out: Dict[str, torch.Tensor] = {}
        V, ps = d["vit_hidden"], d["patch_size"]
        A, adim, pdim = d["act_hidden"], d["action_dim"], d["proprio_dim"]
        vp = "vision_tower.vision_model."
        top = lambda f: self.view(("top", f))   # noqa: E731
        out[vp + "embeddings.patch_embedding.weight"] = top("patch_w")[:, :3 * ps * ps].reshape(V, 3, ps, ps)
        out[vp + "embeddings.patch_embedding.bias"] = top("patch_b")
        out[vp + "embeddings.position_embedding.weight"] = top("pos_emb")
        out[vp + "post_layernorm.weight"], out[vp + "post_layernorm.bias"] = top("post_ln_w"), top("post_ln_b")
        out["multi_modal_projector.linear.weight"], out["multi_modal_projector.linear.bias"] = top("proj_w"), top("proj_b")
        for i in range(d["vit_layers"]):
            q = vp + f"encoder.layers.{i}."
            g = lambda f: self.view(("vit", i, f))   # noqa: E731
            out[q + "layer_norm1.weight"], out[q + "layer_norm1.bias"] = g("ln1_w"), g("ln1_b")
            out[q + "layer_norm2.weight"], out[q + "layer_norm2.bias"] = g("ln2_w"), g("ln2_b")
            wq, bq = g("w_qkv"), g("b_qkv")
            for j, n in enumerate("qkv"):
                out[q + f"self_attn.{n}_proj.weight"] = wq[j * V:(j + 1) * V]
                out[q + f"self_attn.{n}_proj.bias"] = bq[j * V:(j + 1) * V]
            out[q + "self_attn.out_proj.weight"], out[q + "self_attn.out_proj.bias"] = g("w_o"), g("b_o")
            out[q + "mlp.fc1.weight"], out[q + "mlp.fc1.bias"] = g("w_fc1"), g("b_fc1")
            out[q + "mlp.fc2.weight"], out[q + "mlp.fc2.bias"] = g("w_fc2"), g("b_fc2")
        nh, hd = d["num_heads"], d["head_dim"]
        for name in ("vlm", "proprio", "action"):
            src = "action" if (name == "proprio" and self.tied) else name
            for i in range(d["num_layers"]):
                p = f"joint_model.mixtures.{name}.layers.{i}."
                g = lambda f: self.view((src, i, f))   # noqa: E731
                out[p + "input_layernorm.weight"] = g("norm_in")
                out[p + "post_attention_layernorm.weight"] = g("norm_post")
                wq = g("w_qkv")
                out[p + "self_attn.q_proj.weight"] = wq[:nh * hd]
                out[p + "self_attn.k_proj.weight"] = wq[nh * hd:(nh + 1) * hd]
                out[p + "self_attn.v_proj.weight"] = wq[(nh + 1) * hd:]
                out[p + "self_attn.o_proj.weight"] = g("w_o")
                gu = g("w_gate_up")
                inter, hid = gu.shape[0] // 2, gu.shape[1]
                gu = gu.view(inter // 128, 2, 128, hid)
                out[p + "mlp.gate_proj.weight"] = gu[:, 0].reshape(inter, hid)
                out[p + "mlp.up_proj.weight"] = gu[:, 1].reshape(inter, hid)
                out[p + "mlp.down_proj.weight"] = g("w_down")
        out["joint_model.mixtures.action.norm.weight"] = top("action_final_norm")
        out["action_encoder.linear_1.weight"] = top("enc_w1")[:, :adim]
        out["action_encoder.linear_1.bias"] = top("enc_b1")
        out["action_encoder.linear_2.weight"] = torch.cat([top("enc_w2t"), top("enc_w2a")], dim=1)
        out["action_encoder.linear_2.bias"] = top("enc_b2")
        out["action_encoder.linear_3.weight"], out["action_encoder.linear_3.bias"] = top("enc_w3"), top("enc_b3")
        out["proprio_encoder.weight"], out["proprio_encoder.bias"] = top("prop_w")[:, :pdim], top("prop_b")
        out["action_decoder.weight"], out["action_decoder.bias"] = top("dec_w")[:adim], top("dec_b")[:adim]
        return out


def flow_matching_step(model, input_ids, pixel_values, proprios, actions, t, *, noise=None, valid_len=None,
                       causal_mask=None, grads: Optional[GradBuffer] = None, loss_scale: float = 1.0,
                       freeze_vision: bool = False) -> torch.Tensor:
    """One forward + backward of the flow-matching loss (pizero.py:607-661 followed by `loss.backward()`): returns the loss
    (fp32 scalar tensor on the device) and ACCUMULATES `loss_scale * d loss / d weight` into `grads` (None = loss only)."""
    from .pizero import PzError
    model.pack()
    lib = _lib.load()
    dev = model._packed[0][0].device
    B = input_ids.shape[0]
    Sv, H, Adim = model.max_image_text_tokens, model.horizon_steps, model.action_dim
    n_img = model.dims.get("num_images", 1) * model.dims["num_image_tokens"]
    if input_ids.shape != (B, Sv):
        raise ValueError(f"input_ids must be [B, {Sv}], got {tuple(input_ids.shape)}")
    if actions.shape != (B, H, Adim) or t.shape != (B,):
        raise ValueError("actions must be [B, horizon, action_dim] and t [B]")
    ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
    if not bool((ids[:, :n_img] == model.image_token_index).all()) or bool((ids[:, n_img:] == model.image_token_index).any()):
        raise ValueError("training step: the image tokens must be the first num_images * num_image_tokens positions")
    u8 = pixel_values.dtype == torch.uint8
    pix = pixel_values.to(device=dev).contiguous() if u8 else pixel_values.to(device=dev, dtype=model._T).contiguous()
    prop = proprios.to(device=dev, dtype=torch.float32).contiguous()
    if valid_len is None and causal_mask is not None:
        valid_len = (causal_mask[:, 0, 0, :Sv] == 0).sum(-1, dtype=torch.int32)
    vlen = model._valid_len(None, ids, valid_len)
    x1 = actions.to(device=dev, dtype=torch.float32).contiguous()
    if noise is None:   # pizero.py:622
        noise = torch.randn_like(x1)
    x0 = noise.to(device=dev, dtype=torch.float32).contiguous()
    tt = t.to(device=dev, dtype=torch.float32).contiguous()
    loss = torch.empty((), device=dev, dtype=torch.float32)
    nbytes = lib.pz_train_workspace_bytes(model._handle, B)
    tw = model.__dict__.get("_train_ws")
    if tw is None or tw.numel() < nbytes + 1024:
        model.__dict__["_train_ws"] = None
        tw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        model.__dict__["_train_ws"] = tw
    base = (tw.data_ptr() + 1023) // 1024 * 1024
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib.pz_set_pixel_format(model._handle, 1 if u8 else 0)
        rc = lib.pz_flow_matching_step(model._handle, ids.data_ptr(), pix.data_ptr(), vlen.data_ptr(), prop.data_ptr(),
                                       x1.data_ptr(), x0.data_ptr(), tt.data_ptr(), float(model.flow_sig_min),
                                       C.byref(grads.struct) if grads is not None else None, float(loss_scale),
                                       loss.data_ptr(), base, nbytes, B, 1 if freeze_vision else 0, stream)
    if rc != 0:
        raise PzError(f"pz_flow_matching_step failed ({rc}): {lib.pz_last_error(model._handle).decode()}")
    model.last_launch_count = int(lib.pz_launch_count(model._handle))
    model._inflight = (ids, pix, prop, vlen, x1, x0, tt)
    return loss
