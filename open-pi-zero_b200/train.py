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
# flat order = the reference's two optimizer groups (pizero.py:114-129, train.py:171-199): [vision tower, projector, vlm
# mixture] then [action encoder / decoder, proprio encoder, action (and proprio) mixture]
TOP_VLM = ("patch_w", "patch_b", "pos_emb", "post_ln_w", "post_ln_b", "proj_w", "proj_b")
TOP_ACTION = ("action_final_norm", "enc_w1", "enc_b1", "enc_w2a", "enc_w2t", "enc_b2", "enc_w3", "enc_b3", "prop_w", "prop_b",
              "dec_w", "dec_b")
TOP_FIELDS = TOP_VLM + TOP_ACTION


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

        self._ptrs: Dict[tuple, int] = {}

        def add(key, ptr):
            nonlocal off
            t = by_ptr[ptr]
            self._ptrs[key] = ptr
            self.entries.append((key, off, tuple(t.shape)))
            self._index[key] = (off, tuple(t.shape))
            off += (t.numel() + 255) // 256 * 256      # 1 KiB-aligned fp32 slices

        def add_mixture(name):
            arr = getattr(w, name)
            for i in range(d["num_layers"]):
                for f in MIX_FIELDS:
                    add((name, i, f), _ptr(getattr(arr[i], f)))

        for f in TOP_VLM:
            add(("top", f), _ptr(getattr(w, f)))
        for i in range(d["vit_layers"]):
            for f in VIT_FIELDS:
                add(("vit", i, f), _ptr(getattr(w.vit[i], f)))
        add_mixture("vlm")
        self.action_begin = off                      # first element of the action-expert group
        for f in TOP_ACTION:
            add(("top", f), _ptr(getattr(w, f)))
        self.tied = C.addressof(w.proprio.contents) == C.addressof(w.action.contents)
        add_mixture("action")
        if not self.tied:
            add_mixture("proprio")
        self.numel = off
        self.packed = {key: by_ptr[p] for key, p in self._ptrs.items()}   # key -> packed weight tensor (model dtype)
        dev = keep[0].device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        base = self.flat.data_ptr()
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

    def view(self, key, flat: Optional[torch.Tensor] = None) -> torch.Tensor:
        off, shape = self._index[key]
        n = 1
        for s in shape:
            n *= s
        return (self.flat if flat is None else flat)[off:off + n].view(shape)

    def unpack(self, flat: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """The buffer (default: the gradients; any flat tensor of the same layout, e.g. the optimizer's master weights)
        under the reference's parameter names / shapes (fp32 views or small copies)."""
        d = self.model.dims
        _view = self.view
        self_view = lambda key: _view(key, flat)   # noqa: E731
        out: Dict[str, torch.Tensor] = {}
        V, ps = d["vit_hidden"], d["patch_size"]
        A, adim, pdim = d["act_hidden"], d["action_dim"], d["proprio_dim"]
        vp = "vision_tower.vision_model."
        top = lambda f: self_view(("top", f))   # noqa: E731
        out[vp + "embeddings.patch_embedding.weight"] = top("patch_w")[:, :3 * ps * ps].reshape(V, 3, ps, ps)
        out[vp + "embeddings.patch_embedding.bias"] = top("patch_b")
        out[vp + "embeddings.position_embedding.weight"] = top("pos_emb")
        out[vp + "post_layernorm.weight"], out[vp + "post_layernorm.bias"] = top("post_ln_w"), top("post_ln_b")
        out["multi_modal_projector.linear.weight"], out["multi_modal_projector.linear.bias"] = top("proj_w"), top("proj_b")
        for i in range(d["vit_layers"]):
            q = vp + f"encoder.layers.{i}."
            g = lambda f: self_view(("vit", i, f))   # noqa: E731
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
                g = lambda f: self_view((src, i, f))   # noqa: E731
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
                       freeze_vision: bool = False, overlap: Optional["OverlappedAllReduce"] = None) -> torch.Tensor:
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
                                       loss.data_ptr(), base, nbytes, B, 1 if freeze_vision else 0,
                                       overlap.handles if overlap is not None else None, overlap.n if overlap is not None else 0,
                                       stream)
        if overlap is not None and rc == 0:
            overlap.launch()
    if rc != 0:
        raise PzError(f"pz_flow_matching_step failed ({rc}): {lib.pz_last_error(model._handle).decode()}")
    model.last_launch_count = int(lib.pz_launch_count(model._handle))
    model._inflight = (ids, pix, prop, vlen, x1, x0, tt)
    return loss


def release_training_workspace(model) -> None:
    """Free the activation workspace `flow_matching_step` keeps on the model between steps (29 GB at 32 samples of the bridge
    shape), e.g. before a large-batch validation `infer_action`."""
    model.__dict__.pop("_train_ws", None)


class FusedAdamW:
    """clip_grad_norm_ + the two AdamW optimizers of the reference's training loop (train.py:171-199, 371-379) as ONE fused
    pass per parameter group over the flat buffers: fp32 master weights and moments in the GradBuffer layout, the updated
    weights written straight into the packed tensors the forward kernels read (no re-pack after a step).  The update rule
    is `torch.optim.AdamW`'s (the reference's bitsandbytes AdamW8bit keeps 8-bit moments)."""

    def __init__(self, grads: GradBuffer, action_lr: float = 5e-5, vlm_lr: float = 5e-5, action_weight_decay: float = 0.0,
                 vlm_weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8, max_grad_norm: Optional[float] = 1.0,
                 train_vlm: bool = True):
        self.grads = grads
        self.model = grads.model
        dev = grads.flat.device
        self.master = torch.zeros_like(grads.flat)
        for key, _, _ in grads.entries:     # master weights <- the packed weights (exact for an fp32 model)
            grads.view(key, self.master).copy_(grads.packed[key].to(torch.float32))
        self.m = torch.zeros_like(grads.flat)
        self.v = torch.zeros_like(grads.flat)
        self.sumsq = torch.zeros(1 + 1184, dtype=torch.float32, device=dev)   # [0] the result, then pz_grad_sumsq's per-CTA partials
        offs = [off for _, off, _ in grads.entries] + [grads.numel]
        self._off = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._n = torch.tensor([grads.packed[k].numel() for k, _, _ in grads.entries], dtype=torch.int64, device=dev)
        self._dst = torch.tensor([grads.packed[k].data_ptr() for k, _, _ in grads.entries], dtype=torch.int64, device=dev)
        self.groups = {"vlm": dict(begin=0, end=grads.action_begin, lr=vlm_lr, weight_decay=vlm_weight_decay),
                       "action": dict(begin=grads.action_begin, end=grads.numel, lr=action_lr, weight_decay=action_weight_decay)}
        self.betas, self.eps, self.max_grad_norm, self.train_vlm = betas, eps, max_grad_norm, train_vlm
        self.step_count = 0
        self._packed_owner = self.model._packed     # the dst pointers are only valid for this packing

    def step(self, grad_scale: float = 1.0, zero_grad: bool = True):
        """One update from the accumulated gradients.  `grad_scale` multiplies every gradient first (1 / world_size after a
        sum all-reduce).  The gradient norm is measured over the trained parameters (train.py:371)."""
        from .pizero import PzError
        if self.model._packed is not self._packed_owner:
            raise PzError("the model was re-packed since this optimizer was built (parameters edited outside of it)")
        lib = _lib.load()
        g = self.grads
        dev = g.flat.device
        self.step_count += 1
        names = ("vlm", "action") if self.train_vlm else ("action",)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            lo = 0 if self.train_vlm else g.action_begin
            sumsq = None
            if self.max_grad_norm is not None:
                rc = lib.pz_grad_sumsq(g.flat.data_ptr() + 4 * lo, g.numel - lo, self.sumsq.data_ptr(), st)
                if rc != 0:
                    raise PzError(f"pz_grad_sumsq failed ({rc})")
                sumsq = self.sumsq.data_ptr()
            for name in names:
                grp = self.groups[name]
                rc = lib.pz_adamw_step(self.master.data_ptr(), g.flat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), grp["begin"],
                                       grp["end"], self._off.data_ptr(), self._dst.data_ptr(), self._n.data_ptr(), len(g.entries),
                                       _lib.PZ_BF16 if self.model._T == torch.bfloat16 else _lib.PZ_F32, float(grp["lr"]),
                                       float(self.betas[0]), float(self.betas[1]), float(self.eps), float(grp["weight_decay"]),
                                       self.step_count, sumsq, float(self.max_grad_norm or 0.0), float(grad_scale),
                                       1 if zero_grad else 0, st)
                if rc != 0:
                    raise PzError(f"pz_adamw_step failed ({rc})")
        self.model._graphs = {}     # captured graphs replay kernels on the same packed tensors: still valid, but drop for safety
        self._params_stale = True

    def grad_norm(self) -> torch.Tensor:
        """sqrt of the last measured sum of squares (before grad_scale): the value clip_grad_norm_ returns."""
        return self.sumsq[0].sqrt()

    @torch.no_grad()
    def sync_parameters(self):
        """Copy the master weights back into the reference-layout nn.Parameters (for state_dict() / checkpoints / EMA,
        train.py:390-403).  Not needed between steps: the kernels read the packed tensors the optimizer updates."""
        named = dict(self.model.named_parameters())
        for k, t in self.grads.unpack(self.master).items():
            if k in named:
                named[k].copy_(t.to(named[k].dtype))
        self.model._packed_key = self.model._param_key()      # the packed weights already hold these values
        self._params_stale = False


def allreduce_gradients(grads: GradBuffer, bucket_mb: int = 256, group=None, async_op: bool = False):
    """Sum the flat gradient buffer over the data-parallel ranks in buckets of `bucket_mb` (DDP's C1, train.py:119-126,
    350-368; NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU tests).  Divide by the world size through
    `FusedAdamW.step(grad_scale=1 / world_size)`.  Returns the work handles when `async_op`."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return []
    n = grads.flat.numel()
    step = max(1024, bucket_mb * (1 << 20) // 4)
    works = []
    for lo in range(0, n, step):
        w = dist.all_reduce(grads.flat[lo:lo + step], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def bucket_ranges(entries, numel: int, L: int, LV: int, tied: bool):
    """The slices of the flat gradient buffer that become final together, indexed like the events of
    `pz_flow_matching_step` ([l] joint layer l, [L] the action heads, [L + 1 + i] SigLIP layer i, [L + 1 + LV] the rest), and
    the order in which the backward finishes them.  Pure layout logic (CPU-testable): every element belongs to exactly one."""
    offs = {key: off for key, off, _ in entries}
    ends = [off for _, off, _ in entries][1:] + [numel]
    end_of = {key: end for (key, _, _), end in zip(entries, ends)}

    def span(first, last):
        return (offs[first], end_of[last])

    n = L + LV + 2
    ranges = [[] for _ in range(n)]
    mixes = ("vlm", "action") + (() if tied else ("proprio",))
    for l in range(L):
        ranges[l] = [span((m, l, MIX_FIELDS[0]), (m, l, MIX_FIELDS[-1])) for m in mixes]
    ranges[L] = [span(("top", TOP_ACTION[0]), ("top", TOP_ACTION[-1]))]
    for i in range(LV):
        ranges[L + 1 + i] = [span(("vit", i, VIT_FIELDS[0]), ("vit", i, VIT_FIELDS[-1]))]
    ranges[L + 1 + LV] = [span(("top", TOP_VLM[0]), ("top", TOP_VLM[-1]))]
    order = list(range(L - 1, -1, -1)) + [L] + list(range(L + LV, L, -1)) + [L + 1 + LV]   # completion order
    assert sum(hi - lo for r in ranges for lo, hi in r) == numel
    return ranges, order


class OverlappedAllReduce:
    """The gradient all-reduce of the data-parallel step, overlapped with the backward (what DDP's bucket hooks do for the
    reference, train.py:119-126): `pz_flow_matching_step` records one event per group of finished gradients (joint layer
    17 .. 0, the action heads, SigLIP layer 26 .. 0, the rest); a side stream waits for each event and all-reduces that
    slice of the flat buffer (NCCL over NVLink / NVSwitch) while the main stream keeps computing.  `wait()` makes the
    current stream wait for all of them (call it before the optimizer step)."""

    def __init__(self, grads: GradBuffer, group=None, wire_dtype: torch.dtype = torch.float32):
        """`wire_dtype=torch.bfloat16`: every slice is rounded to bf16 before it goes over NVLink and widened again after the
        sum (half the bytes; what DDP moves for the reference's bf16 parameters) -- fp32 (the default) keeps the sum exact."""
        d = grads.model.dims
        L, LV = d["num_layers"], d["vit_layers"]
        self.grads, self.group = grads, group
        self.wire = None if wire_dtype == torch.float32 else torch.empty(grads.numel, dtype=wire_dtype, device=grads.flat.device)
        self.n = L + LV + 2
        dev = grads.flat.device
        self.stream = torch.cuda.Stream(device=dev)
        self.events = [torch.cuda.Event() for _ in range(self.n)]
        with torch.cuda.device(dev):
            for e in self.events:       # an event's handle exists after its first record
                e.record()
        self.handles = (C.c_void_p * self.n)(*[e.cuda_event for e in self.events])
        self.ranges, self.order = bucket_ranges(grads.entries, grads.numel, L, LV, grads.tied)
        self._works = []

    def launch(self):
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        flat = self.grads.flat
        with torch.cuda.stream(self.stream):
            for idx in self.order:
                self.stream.wait_event(self.events[idx])
                for lo, hi in self.ranges[idx]:
                    if self.wire is None:
                        self._works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                    else:
                        w = self.wire[lo:hi]
                        w.copy_(flat[lo:hi])
                        dist.all_reduce(w, op=dist.ReduceOp.SUM, group=self.group, async_op=True).wait()   # orders the side stream
                        flat[lo:hi].copy_(w)
            self._done = torch.cuda.Event()
            self._done.record(self.stream)

    def wait(self):
        for w in self._works:
            w.wait()
        self._works = []
        done = self.__dict__.pop("_done", None)
        if done is not None:
            torch.cuda.current_stream(self.grads.flat.device).wait_event(done)


class CosineAnnealingWarmupRestarts:
    """The reference's learning-rate schedule (`src/utils/optim.py:31-160`, stepped once per optimizer update,
    train.py:376-379) for one parameter group of a `FusedAdamW`: linear warm-up from `min_lr` to `max_lr`, cosine decay
    back to `min_lr`, restarts with `cycle_mult` / `gamma`.  Same constructor arguments, `step()`, `get_lr()`,
    `state_dict()` / `load_state_dict()`; `optimizer` is a `FusedAdamW` and `group` the name of its group."""

    def __init__(self, optimizer: "FusedAdamW", group: str, first_cycle_steps: int, cycle_mult: float = 1.0, max_lr: float = 0.1,
                 min_lr: float = 0.001, warmup_steps: int = 0, gamma: float = 1.0, last_epoch: int = -1):
        assert warmup_steps < first_cycle_steps
        self.first_cycle_steps, self.cycle_mult = first_cycle_steps, cycle_mult
        self.base_max_lr = self.max_lr = max_lr
        self.min_lr, self.warmup_steps, self.gamma = min_lr, warmup_steps, gamma
        self.cur_cycle_steps, self.cycle = first_cycle_steps, 0
        self.step_in_cycle = self.last_epoch = last_epoch
        self.optimizer, self.group = optimizer, group
        self.base_lr = min_lr
        optimizer.groups[group]["lr"] = min_lr          # init_lr(): the group starts at min_lr

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, state_dict):
        self.__dict__.update(state_dict)

    def get_lr(self) -> float:
        import math
        if self.step_in_cycle == -1:
            return self.base_lr
        if self.step_in_cycle < self.warmup_steps:
            return (self.max_lr - self.base_lr) * self.step_in_cycle / self.warmup_steps + self.base_lr
        return self.base_lr + (self.max_lr - self.base_lr) * (
            1 + math.cos(math.pi * (self.step_in_cycle - self.warmup_steps) / (self.cur_cycle_steps - self.warmup_steps))) / 2

    def step(self):
        import math
        epoch = self.last_epoch + 1
        self.step_in_cycle += 1
        if self.step_in_cycle >= self.cur_cycle_steps:
            self.cycle += 1
            self.step_in_cycle -= self.cur_cycle_steps
            self.cur_cycle_steps = int((self.cur_cycle_steps - self.warmup_steps) * self.cycle_mult) + self.warmup_steps
        self.max_lr = self.base_max_lr * (self.gamma ** self.cycle)
        self.last_epoch = math.floor(epoch)
        self.optimizer.groups[self.group]["lr"] = self.get_lr()


def _optimizer_state_dict(self) -> dict:
    """What `TrainAgent.save_model` stores per optimizer (train.py:520-560): here one entry for both groups."""
    return dict(master=self.master, exp_avg=self.m, exp_avg_sq=self.v, step=self.step_count,
                groups={k: dict(lr=g["lr"], weight_decay=g["weight_decay"]) for k, g in self.groups.items()})


def _optimizer_load_state_dict(self, sd: dict):
    """Resume: restores the master weights, both moments and the step count, and rewrites the packed weights from the
    master weights (the model then computes with exactly the weights the checkpointed optimizer held)."""
    self.master.copy_(sd["master"])
    self.m.copy_(sd["exp_avg"])
    self.v.copy_(sd["exp_avg_sq"])
    self.step_count = int(sd["step"])
    for k, g in sd.get("groups", {}).items():
        self.groups[k].update(g)
    self.write_packed(self.master)


def _optimizer_write_packed(self, flat: torch.Tensor):
    """Write a flat fp32 buffer of the gradient layout (master weights, averaged weights) into the packed weight tensors."""
    from .pizero import PzError
    if self.model._packed is not self._packed_owner:
        raise PzError("the model was re-packed since this optimizer was built")
    lib = _lib.load()
    dev = flat.device
    with torch.cuda.device(dev):
        rc = lib.pz_write_packed(flat.data_ptr(), 0, self.grads.numel, self._off.data_ptr(), self._dst.data_ptr(), self._n.data_ptr(),
                                 len(self.grads.entries), _lib.PZ_BF16 if self.model._T == torch.bfloat16 else _lib.PZ_F32,
                                 torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
        raise PzError(f"pz_write_packed failed ({rc})")
    self.model._graphs = {}


FusedAdamW.state_dict = _optimizer_state_dict
FusedAdamW.load_state_dict = _optimizer_load_state_dict
FusedAdamW.write_packed = _optimizer_write_packed


class ModelAveraging:
    """EMA / SWA of the trained weights (`src/agent/model_averaging.py:9-90`: `torch.optim.swa_utils.AveragedModel` with
    the EMA or the equal-weight SWA rule) kept as ONE flat fp32 buffer in the optimizer's layout and updated by one fused
    pass (`pz_average_update`).  `averaged_weights()` is a context manager that lets the model compute with the averaged
    weights (what validation does through `get_model_module()`, train.py:418-433) and puts the trained weights back."""

    def __init__(self, optimizer: FusedAdamW, use_ema: bool = False, use_swa: bool = False, ema_start: int = 0, ema_decay: float = 0.99,
                 ema_freq: int = 1, swa_start: int = 0, swa_freq: int = 1):
        assert not (use_ema and use_swa), "Cannot use both EMA and SWA at once"
        self.opt = optimizer
        self.use_ema, self.use_swa = use_ema, use_swa
        self.ema_start, self.ema_decay, self.ema_freq = ema_start, ema_decay, ema_freq
        self.swa_start, self.swa_freq = swa_start, swa_freq
        self.avg: Optional[torch.Tensor] = None
        self.n_averaged = 0

    def maybe_initialize(self, cnt_update: int):
        if (self.use_swa and cnt_update == self.swa_start) or (self.use_ema and cnt_update == self.ema_start):
            self.avg = self.opt.master.clone()       # AveragedModel copies the model; the first update_parameters() also copies
            self.n_averaged = 0

    def maybe_update(self, cnt_update: int):
        if self.avg is None:
            return
        due = (self.use_ema and cnt_update % self.ema_freq == 0) or (self.use_swa and cnt_update % self.swa_freq == 0)
        if not due:
            return
        if self.n_averaged == 0:
            self.avg.copy_(self.opt.master)
        else:
            w = (1.0 - self.ema_decay) if self.use_ema else 1.0 / (self.n_averaged + 1)
            lib = _lib.load()
            dev = self.avg.device
            with torch.cuda.device(dev):
                lib.pz_average_update(self.avg.data_ptr(), self.opt.master.data_ptr(), self.avg.numel(), float(w),
                                      torch.cuda.current_stream(dev).cuda_stream)
        self.n_averaged += 1

    def averaged_weights(self):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            if self.avg is None:
                yield self.opt.model
                return
            self.opt.write_packed(self.avg)
            try:
                yield self.opt.model
            finally:
                self.opt.write_packed(self.opt.master)
        return ctx()

    def state_dict(self) -> dict:
        if self.avg is None:
            return {}
        return dict(averaged=self.avg, n_averaged=self.n_averaged, model_type="ema" if self.use_ema else "swa")


def preprocess_batch(model, processor, batch: dict, dtype: torch.dtype, device, split_mask: bool, sample_fm_time: bool,
                     time_sampler=None, dense_masks: bool = True) -> dict:
    """`TrainAgent.run`'s `preprocess_batch` closure (train.py:271-313): the dataloader's batch (`observation.image_primary`
    [B, T, H, W, C] uint8, `observation.proprio`, `action`, `task.language_instruction` as bytes) -> the keyword arguments of
    `PiZero.forward` (`split_mask=False`) or `infer_action` (`split_mask=True`).  `dense_masks=False` is the B200-native form:
    instead of the O(B S^2) additive masks and the three position-id tensors (which the kernels reduce back to one integer
    per sample) it returns `valid_len` [B] -- what `flow_matching_step(valid_len=...)` / `infer_action(valid_len=...)` take.
    `processor`: `processing.VLAProcessor` (with `keep_uint8=True` the frames stay uint8 and are normalised on the device)."""
    images = batch["observation"]["image_primary"]
    proprios = batch["observation"]["proprio"]
    actions = batch["action"].squeeze(1)                                   # remove the time dimension
    texts = [t.decode("utf-8") if isinstance(t, (bytes, bytearray)) else str(t) for t in batch["task"]["language_instruction"]]
    B, T, Hh, Ww, Cc = images.shape
    images = images.permute(0, 1, 4, 2, 3).reshape(B, T * Cc, Hh, Ww)      # "B T H W C -> B (T C) H W"
    model_inputs = processor(text=texts, images=images)
    pix = model_inputs["pixel_values"]
    inputs = {"input_ids": model_inputs["input_ids"],
              "pixel_values": pix if pix.dtype == torch.uint8 else pix.to(dtype),
              "proprios": proprios.to(dtype), "actions": actions.to(dtype)}
    if dense_masks:
        causal_mask, vlm_pos, proprio_pos, action_pos = model.build_causal_mask_and_position_ids(model_inputs["attention_mask"], dtype)
        inputs.update(vlm_position_ids=vlm_pos, proprio_position_ids=proprio_pos, action_position_ids=action_pos)
        if split_mask:
            inputs["image_text_proprio_mask"], inputs["action_mask"] = model.split_full_mask_into_submasks(causal_mask)
        else:
            inputs["causal_mask"] = causal_mask
    else:
        inputs["valid_len"] = model_inputs["attention_mask"].sum(dim=1).to(torch.int32)
    if sample_fm_time:
        inputs["t"] = (time_sampler or _default_time_sampler()).sample_fm_time(len(texts)).to(dtype)
    return {k: v.to(device) for k, v in inputs.items()}


def _default_time_sampler():
    from .flow import FlowTimeSampler
    return FlowTimeSampler("beta")
