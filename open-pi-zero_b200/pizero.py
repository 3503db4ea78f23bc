"""Drop-in `PiZero` / `PiZeroInference` for the `infer_action` path.

Same constructor contract, method names, keyword arguments and `state_dict()`
keys as the reference's `src/model/vla/pizero.py` (class `PiZero` :28,
`infer_action` :416, `build_causal_mask_and_position_ids` :271,
`split_full_mask_into_submasks` :326, `tie_action_proprio_weights` :262,
`PiZeroInference` :664), so a reference checkpoint loads with
`load_state_dict(strict=True)` and `EvalAgent` (`src/agent/eval.py:32-125`) can
call `self.model(**inputs)` unchanged.

What is different underneath: the modules below only *hold parameters* under
the reference's names.  All arithmetic of `infer_action` runs in the
hand-written sm_100a kernels of `libpz_b200.so` (C ABI: `include/pz_b200.h`),
fed from packed weight buffers built once per (weights, dtype, device).  There
is no PyTorch/CPU fallback: without the CUDA library or a CUDA device the call
raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from .config import AttrDict, cfg_from_dims, dims_from_cfg
from .synth import state_dict_spec


class PzError(RuntimeError):
    pass


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _Holder(nn.Module):
    """Parameter container: nested modules named after the reference's
    attribute paths so that state_dict() reproduces its 938 keys."""

    def forward(self, *a, **k):  # pragma: no cover
        raise PzError("this module only holds parameters; call PiZero.infer_action")


def _attach(root: nn.Module, key: str, tensor: torch.Tensor) -> None:
    parts = key.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Holder())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class KVCache:
    """Read-only view of one mixture's prefix KV inside the workspace, with the
    reference's `KVCache` accessors (`src/model/kv_cache.py:6-46`).  The
    storage is one `[layers, B, S_cache, head_dim]` tensor per K and V
    (K post-RoPE, V raw, pad positions included) instead of Python lists that
    grow by `torch.cat`."""

    def __init__(self, k: torch.Tensor, v: torch.Tensor, lo: int, hi: int, filled: int):
        self._k, self._v, self._lo, self._hi, self._filled = k, v, lo, hi, filled

    def has_item(self, layer_idx) -> bool:
        return layer_idx < self._filled

    def num_items(self) -> int:
        return 0 if self._filled == 0 else self._hi - self._lo

    def get(self, layer_idx) -> Tuple[torch.Tensor, torch.Tensor]:
        # [B, num_kv_heads = 1, S, head_dim], as the reference returns
        return (self._k[layer_idx, :, self._lo:self._hi].unsqueeze(1),
                self._v[layer_idx, :, self._lo:self._hi].unsqueeze(1))

    def update(self, key_states, value_states, layer_idx):
        raise PzError("the prefix KV cache is written by the prefill kernels, not from Python")


class TextKVCache:
    """KV cache of `PiZero.infer_text`: one `[layers, B, capacity, head_dim]` tensor per K (post-RoPE) and V that decode
    steps append to in place (`cache_mode="append"`, joint_model.py:164-240; the reference grows Python lists with
    `torch.cat`, kv_cache.py:6-46).  Same accessors as the reference's `KVCache`."""

    def __init__(self):
        self.k = self.v = None
        self.length = 0          # tokens per sample held so far (prompts are not padded, pizero.py:346-357)
        self.n_layers = 0

    def _allocate(self, n_layers, batch, capacity, head_dim, dtype, device):
        self.k = torch.zeros((n_layers, batch, capacity, head_dim), dtype=dtype, device=device)
        self.v = torch.zeros_like(self.k)
        self.n_layers, self.length = n_layers, 0

    @property
    def capacity(self) -> int:
        return 0 if self.k is None else self.k.shape[2]

    def has_item(self, layer_idx) -> bool:
        return self.k is not None and self.length > 0 and layer_idx < self.n_layers

    def num_items(self) -> int:
        return self.length

    def get(self, layer_idx):
        """(K, V) of one layer as `[B, 1, length, head_dim]` views (kv_cache.py:28-33)."""
        return self.k[layer_idx, :, None, : self.length], self.v[layer_idx, :, None, : self.length]


class JointModel(_Holder):
    """Holds `mixtures.{vlm,proprio,action}` parameters and mirrors the reference's
    `JointModel` surface (`joint_model.py:308-383`): attributes, `build_mixture_caches`
    and `forward` for the two call patterns `infer_action` uses."""

    def __init__(self, dims: dict):
        super().__init__()
        self.num_hidden_layers = dims["num_layers"]
        self.mixture_names = ["vlm", "proprio", "action"]
        self.cache_names = ["vlm", "proprio"]
        self.num_mixture = 3
        self._owner = None   # weakref to the PiZero that owns the packed weights / workspace

    def build_mixture_caches(self):
        """joint_model.py:325 -- empty caches; `forward(..., return_caches=True)` fills them."""
        return {name: KVCache(None, None, 0, 0, 0) for name in self.cache_names}

    @torch.no_grad()
    def forward(self, attention_mask, position_ids_all, embeds_all, time_cond=None,
                final_layer_post_attn_skip_names=("vlm", "proprio"), kv_caches={},
                cache_mode="append_non_active", return_caches=False):
        """joint_model.py:328-383 for (a) the prefix pass (vlm + proprio active, caches filled,
        nothing returned for the skipped mixtures), (b) the action pass over the cached prefix
        (`cache_mode="append_non_active"`), (c) all three mixtures active with no cache (the training forward,
        pizero.py:637-652) or with `cache_mode="no_append"` caches (infer_action_naive, pizero.py:529-544: filled by the first
        call, read by the later ones) and (d) the vlm mixture alone in `cache_mode="append"` with a `TextKVCache`
        (infer_text, pizero.py:571-583): the prompt first, then one token per call.
        Like the reference, `embeds_all[*]` is scaled by sqrt(hidden) IN PLACE (joint_model.py:355)."""
        assert cache_mode in ["no_append", "append", "append_non_active"], f"Invalid cache mode: {cache_mode}"
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise PzError("JointModel.forward needs the owning PiZero (construct it through PiZero)")
        names = list(embeds_all.keys())
        skip = tuple(final_layer_post_attn_skip_names)
        # decide BEFORE touching the caller's tensors whether this call pattern is one of the three the kernels cover
        pattern = None
        if names == ["vlm", "proprio"] and skip == ("vlm", "proprio"):
            pattern = "prefix"
        elif names == ["action"] and cache_mode == "append_non_active":
            pattern = "action"
        elif names == ["vlm", "proprio", "action"] and skip == ("vlm", "proprio") and (
                not kv_caches or cache_mode == "no_append"):
            # no caches: the training forward; "no_append" with caches: infer_action_naive (pizero.py:529-544) -- the vlm /
            # proprio K / V are cached by the first call and only READ afterwards (joint_model.py:176-196)
            pattern = "joint"
        elif names == ["vlm"] and cache_mode == "append" and skip == ():
            pattern = "text"    # infer_text's call (pizero.py:571-583): the vlm mixture alone, K / V appended to its cache
        if pattern is None:
            raise NotImplementedError(
                f"JointModel.forward with active mixtures {names} / cache_mode {cache_mode!r} / skip {skip} is outside the "
                "infer_action / training-forward paths (\"append\" / \"no_append\" text generation is not built)")
        if time_cond is not None:
            raise NotImplementedError("adaLN time conditioning (time_cond) is not built: action_expert_adaptive_mode must be None")
        if owner.check_inputs not in ("0", "off") and position_ids_all is not None and pattern != "text":
            B_ = embeds_all[names[0]].shape[0]
            owner._check_position_ids(B_, position_ids_all.get("vlm"), position_ids_all.get("proprio"),
                                      position_ids_all.get("action"))
        for n in names:
            e = embeds_all[n]
            e *= torch.tensor(e.shape[-1] ** 0.5, dtype=e.dtype, device=e.device)
        owner.pack()
        lib = _lib.load()
        d = owner.dims
        B = embeds_all[names[0]].shape[0]
        Sv = d["max_image_text_tokens"]
        dev = owner._packed[0][0].device
        vlen = None
        if pattern != "text":
            vlen = (attention_mask[:, 0, 0, :Sv] == 0).sum(-1, dtype=torch.int32).to(dev).contiguous()
        ws, ws_bytes = owner._ensure_workspace(B)
        stream = torch.cuda.current_stream(dev).cuda_stream
        f32 = lambda t: t.to(torch.float32).contiguous()   # noqa: E731
        # launches go into a stream of `dev`: it must be the current device
        with torch.cuda.device(dev):
            if pattern == "prefix":
                xv, xp = f32(embeds_all["vlm"]), f32(embeds_all["proprio"])
                rc = lib.pz_joint_prefix(owner._handle, xv.data_ptr(), xp.data_ptr(), vlen.data_ptr(), ws, ws_bytes,
                                         B, stream)
                if rc != 0:
                    raise PzError(f"pz_joint_prefix failed ({rc}): {lib.pz_last_error(owner._handle).decode()}")
                kv_caches.update(owner.kv_caches(B))
                out = {}
                return (out, kv_caches) if return_caches else out
            if pattern == "action":
                if not all(isinstance(kv_caches.get(n), KVCache) and kv_caches[n].has_item(0) for n in self.cache_names):
                    raise PzError("the action pass needs the caches a prefix JointModel.forward(return_caches=True) filled")
                xa = f32(embeds_all["action"])
                out = torch.empty_like(xa)
                rc = lib.pz_joint_action(owner._handle, xa.data_ptr(), vlen.data_ptr(), out.data_ptr(), ws, ws_bytes,
                                         B, stream)
                if rc != 0:
                    raise PzError(f"pz_joint_action failed ({rc}): {lib.pz_last_error(owner._handle).decode()}")
                res = {"action": out.to(embeds_all["action"].dtype)}
                return (res, kv_caches) if return_caches else res
            if pattern == "text":
                return self._forward_text(owner, lib, embeds_all["vlm"], kv_caches, ws, ws_bytes, stream, dev, return_caches)
            if pattern == "joint" and kv_caches and all(
                    isinstance(kv_caches.get(n), KVCache) and kv_caches[n].has_item(0) for n in self.cache_names):
                # "no_append", caches already filled: only the action rows produce anything new
                xa = f32(embeds_all["action"])
                out = torch.empty_like(xa)
                rc = lib.pz_joint_action(owner._handle, xa.data_ptr(), vlen.data_ptr(), out.data_ptr(), ws, ws_bytes, B, stream)
                if rc != 0:
                    raise PzError(f"pz_joint_action failed ({rc}): {lib.pz_last_error(owner._handle).decode()}")
                res = {"action": out.to(embeds_all["action"].dtype)}
                return (res, kv_caches) if return_caches else res
            if pattern == "joint":
                # the training call pattern (pizero.py:637-652) and infer_action_naive's (pizero.py:529-544): all three
                # mixtures active under the full block mask, no cache.  vlm / proprio rows never attend to action keys
                # (pizero.py:271-310), so this is the prefix pass followed by the action pass over the prefix's K/V.
                xv, xp, xa = f32(embeds_all["vlm"]), f32(embeds_all["proprio"]), f32(embeds_all["action"])
                rc = lib.pz_joint_prefix(owner._handle, xv.data_ptr(), xp.data_ptr(), vlen.data_ptr(), ws, ws_bytes,
                                         B, stream)
                if rc != 0:
                    raise PzError(f"pz_joint_prefix failed ({rc}): {lib.pz_last_error(owner._handle).decode()}")
                out = torch.empty_like(xa)
                rc = lib.pz_joint_action(owner._handle, xa.data_ptr(), vlen.data_ptr(), out.data_ptr(), ws, ws_bytes,
                                         B, stream)
                if rc != 0:
                    raise PzError(f"pz_joint_action failed ({rc}): {lib.pz_last_error(owner._handle).decode()}")
                if cache_mode == "no_append" and kv_caches is not None and len(kv_caches) > 0:
                    kv_caches.update(owner.kv_caches(B))     # the reference fills the caches it was handed on the first call
                res = {"action": out.to(embeds_all["action"].dtype)}
                return (res, kv_caches) if return_caches else res
        raise AssertionError("unreachable")

    def _forward_text(self, owner, lib, emb, kv_caches, ws, ws_bytes, stream, dev, return_caches):
        """`cache_mode="append"`, vlm only (joint_model.py:164-240, 375-380): returns {"vlm": final-norm hidden states}."""
        d = owner.dims
        B, q_len, Hd = emb.shape
        Sv = d["max_image_text_tokens"]
        cache = kv_caches.get("vlm") if kv_caches else None
        if cache is None:
            cache = TextKVCache()
        if not isinstance(cache, TextKVCache):
            raise TypeError('kv_caches["vlm"] must be an open_pi_zero_b200 TextKVCache for cache_mode="append"')
        if owner.__dict__.get("_packed") is None or owner._packed[1].vlm_final_norm is None:
            raise PzError('the vlm-only pass needs mixture.vlm.use_final_norm=True')
        x = emb.to(torch.float32)
        if cache.num_items() == 0:
            if q_len > Sv:
                raise ValueError(f"the prompt has {q_len} tokens, max_image_text_tokens is {Sv}")
            cap = Sv + owner.text_max_new_tokens
            if cache.capacity < cap or cache.k.shape[1] != B or cache.k.dtype != owner._T:
                cache._allocate(d["num_layers"], B, cap, d["head_dim"], owner._T, dev)
            xin = torch.zeros((B, Sv, Hd), dtype=torch.float32, device=dev)
            xin[:, :q_len] = x
            vlen = torch.full((B,), q_len, dtype=torch.int32, device=dev)
            hid = torch.empty((B, Sv, Hd), dtype=torch.float32, device=dev)
            rc = lib.pz_text_prefill(owner._handle, vlen.data_ptr(), cache.k.data_ptr(), cache.v.data_ptr(), cache.capacity, q_len,
                                     None, 0, hid.data_ptr(), xin.data_ptr(), ws, ws_bytes, B, stream)
            out = hid[:, :q_len]
        else:
            if q_len != 1:
                raise ValueError("Using KV cache so should only use one single token")
            cur = cache.num_items()
            if cur + 1 > cache.capacity:
                raise PzError(f"the text KV cache is full ({cache.capacity} rows)")
            xin = x[:, 0].contiguous()
            vlen = torch.full((B,), cur + 1, dtype=torch.int32, device=dev)
            hid = torch.empty((B, Hd), dtype=torch.float32, device=dev)
            rc = lib.pz_text_decode(owner._handle, xin.data_ptr(), vlen.data_ptr(), cur, cache.k.data_ptr(), cache.v.data_ptr(),
                                    cache.capacity, None, hid.data_ptr(), ws, ws_bytes, B, stream)
            out = hid[:, None]
        if rc != 0:
            raise PzError(f"vlm-only pass failed ({rc}): {lib.pz_last_error(owner._handle).decode()}")
        cache.length = cache.num_items() + q_len
        owner._inflight = (xin, vlen)
        if kv_caches is not None:
            kv_caches["vlm"] = cache
        res = {"vlm": out.to(emb.dtype)}
        return (res, kv_caches) if return_caches else res


class PiZero(nn.Module):
    def __init__(self, cfg, use_ddp: bool = False, *, device=None, dtype=None,
                 init: str = "reference", max_batch: int = 64):
        """`cfg`: the reference's config tree (OmegaConf DictConfig, AttrDict or
        dict with the fields of config/train/bridge.yaml:85-181) or the flat
        dims dict of `open-pi-zero_b200/config.py`.

        Extra keyword-only arguments (not in the reference): `device`/`dtype`
        to create parameters in place, `init` in {"reference", "empty"},
        `max_batch` (workspace sizing; grows on demand)."""
        super().__init__()
        self.dims = dims_from_cfg(cfg)
        self.cfg = cfg if not isinstance(cfg, dict) or "mixture" in cfg else cfg_from_dims(self.dims)
        self.use_ddp = use_ddp
        d = self.dims
        self.vocab_size = d["vocab_size"]
        self.pad_token_id = d["pad_token_id"]
        self.image_token_index = d["image_token_index"]
        self.max_image_text_tokens = d["max_image_text_tokens"]
        self.num_proprio_tokens = d["cond_steps"]
        self.num_action_tokens = d["horizon_steps"]
        self.total_num_tokens = (self.max_image_text_tokens + self.num_proprio_tokens
                                 + self.num_action_tokens)
        self.image_text_hidden_size = d["vlm_hidden"]
        self.proprio_hidden_size = d["act_hidden"]
        self.action_hidden_size = d["act_hidden"]
        self.num_inference_steps = d["num_inference_steps"]
        self.horizon_steps = d["horizon_steps"]
        self.action_dim = d["action_dim"]
        self.proprio_dim = d["proprio_dim"]
        self.final_action_clip_value = d["final_action_clip_value"]
        self.flow_sig_min = float(d.get("flow_sig_min", 0.001))   # pizero.py:58

        self.joint_model = JointModel(d)
        import weakref
        self.joint_model._owner = weakref.ref(self)
        dtype = dtype or torch.float32
        tied_keys = {k: kind[5:] for k, _, kind, _ in state_dict_spec(d) if kind.startswith("tied:")}
        if init == "reference":
            from .synth import init_state_dict
            sd = init_state_dict(d, seed=int(torch.initial_seed() % (2 ** 31)), tie_proprio=False)
            for k, t in sd.items():
                if k not in tied_keys:
                    _attach(self, k, t.to(device=device, dtype=dtype))
        elif init == "empty":
            for k, shape, kind, _ in state_dict_spec(d):
                if k not in tied_keys:
                    _attach(self, k, torch.empty(shape, device=device, dtype=dtype))
        else:
            raise ValueError(f"init must be 'reference' or 'empty', got {init!r}")
        self.use_lm_head = bool(d.get("use_lm_head", False))
        if self.use_lm_head:   # pizero.py:105-112: lm_head.weight IS embed_tokens.weight
            self.add_module("lm_head", _Holder())
            self.lm_head.weight = self.embed_tokens.weight
        self._tied = False
        self._max_batch = max_batch
        self._handle = None
        self._packed = None
        self._packed_key = None
        self._workspace = None
        self._ws_batch = 0
        # kernel-selection flags (include/pz_b200.h): SIMT kernels for every op (debug), or SIMT only for shapes no tensor-core
        # kernel covers (otherwise such a shape is an error, never a silent 50x slow-down)
        self._flags = (_lib.PZ_FLAG_SIMPLE_KERNELS if os.environ.get("PZ_SIMPLE_KERNELS") == "1" else 0) | \
                      (_lib.PZ_FLAG_ALLOW_FALLBACK if os.environ.get("PZ_ALLOW_FALLBACK") == "1" else 0)
        self.use_cuda_graph = os.environ.get("PZ_CUDA_GRAPH", "1") != "0"
        # which Euler-loop implementation (include/pz_b200.h PZ_SAMPLER_*): 0 auto, 1 kernels, 2 barrier, 3 stream
        self._sampler_mode = int(os.environ.get("PZ_SAMPLER", "0"))
        self._sampler_pack_batches = tuple(int(b) for b in os.environ.get("PZ_SAMPLER_BATCHES", "1,2").split(",") if b)
        self._sampler_batches = []
        self.text_max_new_tokens = int(os.environ.get("PZ_TEXT_MAX_NEW_TOKENS", "256"))   # capacity of a text KV cache / RoPE table
        self.max_graphs = int(os.environ.get("PZ_MAX_GRAPHS", "8"))
        # dtype of the action chunk infer_action returns: None = the reference's (the dtype of the pixel values, i.e. the
        # model dtype, pizero.py:454-456,484-490); torch.float32 = the sampler's own fp32 state, unrounded
        self.action_dtype = None
        # position ids are compared with the canonical ones on every call (three tiny device comparisons); the dense masks'
        # block structure only with PZ_CHECK_INPUTS=1 (the kernels apply the canonical block mask)
        self.check_inputs = os.environ.get("PZ_CHECK_INPUTS", "pos")
        self._graphs = {}
        self._timing_armed = False
        self.last_launch_count = 0
        self.eval()

    # ------------------------------------------------------------------ misc
    def no_sync(self):   # NoSyncBase parity (src/utils/decorator.py:14-28); inert at inference
        import contextlib
        return contextlib.nullcontext()

    def freeze_all_weights(self):
        for p in self.parameters():
            p.requires_grad = False

    def tie_action_proprio_weights(self):
        """pizero.py:262-264: proprio uses the action expert's weights."""
        self.joint_model.mixtures._modules["proprio"] = self.joint_model.mixtures._modules["action"]
        self._tied = True
        self._packed_key = None
        self.__dict__.pop("_param_list", None)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        # checkpoints written from a torch.compile'd model carry "_orig_mod." (eval.py:184-188)
        sd = {k.replace("_orig_mod.", ""): v for k, v in state_dict.items()}
        out = super().load_state_dict(sd, strict=strict, assign=assign)
        self._packed_key = None
        self.__dict__.pop("_param_list", None)
        return out

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._packed_key = None
        self.__dict__.pop("_param_list", None)
        return out

    # ------------------------------------------------------ input preparation
    def build_causal_mask_and_position_ids(self, attention_mask: torch.Tensor, dtype: torch.dtype):
        """Same outputs as pizero.py:271-324 (block mask with finfo.min, position
        ids starting at 1), built without the per-sample Python loop."""
        bsz = attention_mask.size(0)
        dev = attention_mask.device
        Sv, Sp, H = self.max_image_text_tokens, self.num_proprio_tokens, self.num_action_tokens
        S = self.total_num_tokens
        cnt = attention_mask.sum(dim=1).view(bsz, 1, 1)
        idx = torch.arange(S, device=dev)
        row, col = idx.view(1, S, 1), idx.view(1, 1, S)
        col_valid = col < cnt
        vis = (row < cnt) & col_valid                                   # image/text block
        vis = vis | ((row >= Sv) & col_valid)                           # proprio/action -> image/text
        vis = vis | ((row >= Sv) & (row < Sv + Sp) & (col >= Sv) & (col < Sv + Sp))
        vis = vis | ((row >= Sv + Sp) & (col >= Sv))
        mask = torch.full((bsz, S, S), torch.finfo(dtype).min, dtype=dtype, device=dev)
        mask = mask.masked_fill(vis, 0).unsqueeze(1)
        vlm_pos = torch.arange(1, Sv + 1, device=dev).repeat(bsz, 1)
        proprio_pos = torch.arange(1, Sp + 1, device=dev).repeat(bsz, 1)
        action_pos = torch.arange(Sp + 1, Sp + H + 1, device=dev).repeat(bsz, 1)
        return mask, vlm_pos, proprio_pos, action_pos

    def split_full_mask_into_submasks(self, causal_mask: torch.Tensor):
        """pizero.py:326-336."""
        n = self.max_image_text_tokens + self.num_proprio_tokens
        return causal_mask[..., :n, :n], causal_mask[..., -self.num_action_tokens:, :]

    # ---------------------------------------------------------------- packing
    def _param_key(self):
        """Staleness key of the packed weights, evaluated on every call: dtype, device, storage address and the SUM OF
        ALL parameters' version counters (any in-place edit of any tensor -- optimizer step, EMA update as in the
        reference's train.py:419,433, manual surgery -- bumps one of them).  Walking a cached list of the 938
        parameters costs ~80 us of host time; `.to()` / `load_state_dict` / `_apply` drop the cached list."""
        ps = self.__dict__.get("_param_list")
        if ps is None:
            ps = list(self.parameters())
            self.__dict__["_param_list"] = ps
        if not ps or ps[0].numel() == 0:
            if self._handle is None:
                raise PzError("the parameters were released (release_unpacked_parameters) and the packed weights are "
                              "gone: reload the state dict before calling again")
            return self._packed_key
        return (ps[0].dtype, ps[0].device, sum(p._version for p in ps), ps[0].data_ptr(), self._tied, self._flags)

    @torch.no_grad()
    def pack(self, force: bool = False):
        """Build the kernel-side weight buffers (fused QKV, interleaved gate|up,
        padded small matrices, fp32 vectors, RoPE and time tables) and bind them."""
        key = self._param_key()
        if not force and self._packed_key == key and self._handle is not None:
            return
        sd = dict(self.state_dict())
        p0 = sd["embed_tokens.weight"]
        if p0.device.type != "cuda":
            raise PzError("PiZero.infer_action needs the model on a CUDA device (no CPU path)")
        if p0.dtype not in (torch.float32, torch.bfloat16):
            raise PzError(f"unsupported parameter dtype {p0.dtype}")
        lib = _lib.load()
        d, dev, T = self.dims, p0.device, p0.dtype
        f32 = lambda t: t.detach().to(torch.float32).contiguous()   # noqa: E731
        mat = lambda t: t.detach().to(T).contiguous()               # noqa: E731
        keep = []   # owns every packed tensor

        def own(t):
            keep.append(t)
            return t.data_ptr()

        A, H = d["act_hidden"], d["vlm_hidden"]
        V, ps = d["vit_hidden"], d["patch_size"]
        kp = _round_up(3 * ps * ps, 64)
        skp = _round_up(max(d["action_dim"], d["proprio_dim"]), 8)
        w = _lib.PzWeights()
        w.embed = own(mat(sd["embed_tokens.weight"]))
        vpfx = "vision_tower.vision_model."
        pw = torch.zeros((V, kp), dtype=T, device=dev)
        pw[:, :3 * ps * ps] = sd[vpfx + "embeddings.patch_embedding.weight"].reshape(V, -1).to(T)
        w.patch_w = own(pw)
        w.patch_b = own(f32(sd[vpfx + "embeddings.patch_embedding.bias"]))
        w.pos_emb = own(f32(sd[vpfx + "embeddings.position_embedding.weight"]))
        vit = (_lib.PzVitLayer * d["vit_layers"])()
        for i in range(d["vit_layers"]):
            q = vpfx + f"encoder.layers.{i}."
            L = vit[i]
            L.ln1_w, L.ln1_b = own(f32(sd[q + "layer_norm1.weight"])), own(f32(sd[q + "layer_norm1.bias"]))
            L.w_qkv = own(mat(torch.cat([sd[q + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)))
            L.b_qkv = own(f32(torch.cat([sd[q + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)))
            L.w_o, L.b_o = own(mat(sd[q + "self_attn.out_proj.weight"])), own(f32(sd[q + "self_attn.out_proj.bias"]))
            L.ln2_w, L.ln2_b = own(f32(sd[q + "layer_norm2.weight"])), own(f32(sd[q + "layer_norm2.bias"]))
            L.w_fc1, L.b_fc1 = own(mat(sd[q + "mlp.fc1.weight"])), own(f32(sd[q + "mlp.fc1.bias"]))
            L.w_fc2, L.b_fc2 = own(mat(sd[q + "mlp.fc2.weight"])), own(f32(sd[q + "mlp.fc2.bias"]))
        keep.append(vit)
        w.vit = C.cast(vit, C.POINTER(_lib.PzVitLayer))
        w.post_ln_w, w.post_ln_b = own(f32(sd[vpfx + "post_layernorm.weight"])), own(f32(sd[vpfx + "post_layernorm.bias"]))
        w.proj_w = own(mat(sd["multi_modal_projector.linear.weight"]))
        w.proj_b = own(f32(sd["multi_modal_projector.linear.bias"]))

        GU = 128

        def pack_mixture(name):
            arr = (_lib.PzMixLayer * d["num_layers"])()
            for i in range(d["num_layers"]):
                p = f"joint_model.mixtures.{name}.layers.{i}."
                L = arr[i]
                L.norm_in = own(f32(sd[p + "input_layernorm.weight"]))
                L.w_qkv = own(mat(torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)))
                L.w_o = own(mat(sd[p + "self_attn.o_proj.weight"]))
                L.norm_post = own(f32(sd[p + "post_attention_layernorm.weight"]))
                g, u = sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]
                inter, hid = g.shape
                gu = torch.stack([g.reshape(inter // GU, GU, hid), u.reshape(inter // GU, GU, hid)], 1)
                L.w_gate_up = own(mat(gu.reshape(2 * inter, hid)))
                L.w_down = own(mat(sd[p + "mlp.down_proj.weight"]))
            keep.append(arr)
            return arr

        vlm = pack_mixture("vlm")
        action = pack_mixture("action")
        same = self._tied or all(
            torch.equal(sd[k], sd[k.replace(".proprio.", ".action.", 1)])
            for k in sd if k.startswith("joint_model.mixtures.proprio."))
        proprio = action if same else pack_mixture("proprio")   # alias only if bit-equal (SURVEY 8a note 4)
        w.vlm = C.cast(vlm, C.POINTER(_lib.PzMixLayer))
        w.action = C.cast(action, C.POINTER(_lib.PzMixLayer))
        w.proprio = C.cast(proprio, C.POINTER(_lib.PzMixLayer))
        w.action_final_norm = own(f32(sd["joint_model.mixtures.action.norm.weight"]))

        def pad_k(t, k):
            out = torch.zeros((t.shape[0], k), dtype=T, device=dev)
            out[:, :t.shape[1]] = t.to(T)
            return out

        w.enc_w1 = own(pad_k(sd["action_encoder.linear_1.weight"], skp))
        w.enc_b1 = own(f32(sd["action_encoder.linear_1.bias"]))
        w2 = sd["action_encoder.linear_2.weight"]
        w.enc_w2a = own(mat(w2[:, A:]))
        w.enc_w2t = own(mat(w2[:, :A]))     # time half, for an arbitrary per-sample t (training forward)
        w.enc_b2 = own(f32(sd["action_encoder.linear_2.bias"]))
        # time conditioning is a per-step constant: t_i accumulates
        # dt in fp32 exactly as pizero.py:460-481 does in an fp32 run; the embedding
        # follows vla/modules.py:15-22.
        n_steps = d["num_inference_steps"]
        half = A // 2
        freq = torch.exp(torch.arange(half, device=dev, dtype=torch.float32)
                         * -(math.log(d["time_max_period"]) / (half - 1)))
        w.time_freq = own(freq.contiguous())
        t = torch.zeros(1, device=dev, dtype=torch.float32)
        tb = []
        for _ in range(n_steps):
            e = t[:, None] * freq[None, :]
            temb = torch.cat((e.sin(), e.cos()), dim=-1)[0]
            tb.append(w2[:, :A].float() @ temb + sd["action_encoder.linear_2.bias"].float())
            t = t + 1.0 / n_steps
        w.enc_time_bias = own(torch.stack(tb).contiguous())
        w.enc_w3 = own(mat(sd["action_encoder.linear_3.weight"]))
        w.enc_b3 = own(f32(sd["action_encoder.linear_3.bias"]))
        w.prop_w = own(pad_k(sd["proprio_encoder.weight"], skp))
        w.prop_b = own(f32(sd["proprio_encoder.bias"]))
        dw = torch.zeros((8, A), dtype=T, device=dev)
        dw[:d["action_dim"]] = sd["action_decoder.weight"].to(T)
        w.dec_w = own(dw)
        db = torch.zeros(8, dtype=torch.float32, device=dev)
        db[:d["action_dim"]] = sd["action_decoder.bias"].float()
        w.dec_b = own(db)

        def rope(theta, n_pos):   # paligemma/modules.py:36-67, positions start at 1
            hd = d["head_dim"]
            inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
            fr = torch.arange(1, n_pos + 1, dtype=torch.float32)[:, None] * inv[None, :]
            return fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()

        # text decode steps run past the prompt: positions up to max_image_text_tokens + text_max_new_tokens
        has_text = self.use_lm_head or "joint_model.mixtures.vlm.norm.weight" in sd
        vlm_rope_rows = d["max_image_text_tokens"] + (self.text_max_new_tokens if has_text else 0)
        c1, s1 = rope(d["vlm_rope_theta"], vlm_rope_rows)
        c2, s2 = rope(d["act_rope_theta"], d["cond_steps"] + d["horizon_steps"])
        w.rope_vlm_cos, w.rope_vlm_sin = own(c1), own(s1)
        w.rope_act_cos, w.rope_act_sin = own(c2), own(s2)
        w.small_k_pad = skp
        w.rope_vlm_rows = vlm_rope_rows
        if "joint_model.mixtures.vlm.norm.weight" in sd:
            w.vlm_final_norm = own(f32(sd["joint_model.mixtures.vlm.norm.weight"]))
        if self.use_lm_head:
            lm = sd["lm_head.weight"]
            w.lm_head = w.embed if lm.data_ptr() == sd["embed_tokens.weight"].data_ptr() else own(mat(lm))

        cfg = _lib.PzConfig()
        cfg.dtype = _lib.PZ_BF16 if T == torch.bfloat16 else _lib.PZ_F32
        cfg.vocab_size, cfg.pad_token_id = d["vocab_size"], d["pad_token_id"]
        cfg.image_token_index = d["image_token_index"]
        cfg.s_vlm, cfg.n_img_tokens, cfg.n_images = d["max_image_text_tokens"], d["num_image_tokens"], d.get("num_images", 1)
        cfg.cond_steps, cfg.horizon = d["cond_steps"], d["horizon_steps"]
        cfg.action_dim, cfg.proprio_dim, cfg.n_steps = d["action_dim"], d["proprio_dim"], n_steps
        clip = d["final_action_clip_value"]
        cfg.clip = -1.0 if clip is None else float(clip)
        cfg.n_layers, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim = d["num_layers"], d["num_heads"], d["num_kv_heads"], d["head_dim"]
        cfg.vlm_hidden, cfg.vlm_inter, cfg.act_hidden, cfg.act_inter = H, d["vlm_inter"], A, d["act_inter"]
        cfg.vit_hidden, cfg.vit_inter, cfg.vit_layers, cfg.vit_heads = V, d["vit_inter"], d["vit_layers"], d["vit_heads"]
        cfg.image_size, cfg.patch_size, cfg.patch_k_pad = d["image_size"], ps, kp
        cfg.max_batch = 1 << 20
        cfg.flags = self._flags
        self._destroy_handle()
        hnd = C.c_void_p()
        rc = lib.pz_create(C.byref(cfg), C.byref(hnd))
        if rc != 0:
            raise PzError(f"pz_create failed ({rc}): {lib.pz_last_error(None).decode()}")
        rc = lib.pz_bind_weights(hnd, C.byref(w))
        if rc != 0:
            msg = lib.pz_last_error(hnd).decode()
            lib.pz_destroy(hnd)
            raise PzError(f"pz_bind_weights failed ({rc}): {msg}")
        # bs 1 / 2: per-SM re-packed copy of the action expert for the stream sampler (csrc/denoise_mega3.cu)
        self._sampler_batches = []
        if T == torch.bfloat16 and not (self._flags & _lib.PZ_FLAG_SIMPLE_KERNELS):
            with torch.cuda.device(dev):
                st = torch.cuda.current_stream(dev).cuda_stream
                for b in self._sampler_pack_batches:
                    nbytes = lib.pz_sampler_stream_bytes(hnd, b)
                    if nbytes == 0:
                        continue
                    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
                    base = (buf.data_ptr() + 1023) // 1024 * 1024
                    rc = lib.pz_sampler_pack(hnd, b, base, nbytes, st)
                    if rc != 0:
                        msg = lib.pz_last_error(hnd).decode()
                        lib.pz_destroy(hnd)
                        raise PzError(f"pz_sampler_pack failed ({rc}): {msg}")
                    keep.append(buf)
                    self._sampler_batches.append(b)
        lib.pz_set_sampler(hnd, self._sampler_mode)
        self._handle, self._packed, self._packed_key = hnd, (keep, w), key
        self._apply_io_normalization()
        self._T = T
        self._workspace, self._ws_batch = None, 0
        self._graphs = {}

    # ------------------------------------------------ caller-side normalisation folded into the kernels
    def set_io_normalization(self, dataset_statistics: Optional[dict], action_normalization_type: str = "bound",
                             proprio_normalization_type: str = "bound"):
        """After this call `infer_action` takes RAW proprios and returns DE-NORMALISED actions: what
        `SimplerAdapter.preprocess` / `postprocess` do on the host around the model (simpler.py:76-90, 102-125;
        base.py:8-49) runs inside the kernels instead (adapter.py).  `dataset_statistics` as in the reference's
        `dataset_statistics.json`: {"proprio": {"p01", "p99", "mean", "std"}, "action": {...}}; None switches it off."""
        from .adapter import action_affine, proprio_affine
        if dataset_statistics is None:
            self.__dict__["_io_norm"] = None
        else:
            ps, pb, clip = proprio_affine(dataset_statistics["proprio"], proprio_normalization_type)
            as_, ab = action_affine(dataset_statistics["action"], action_normalization_type)
            if len(ps) != self.proprio_dim or len(as_) != self.action_dim:
                raise ValueError("dataset statistics do not match proprio_dim / action_dim")
            self.__dict__["_io_norm"] = (ps, pb, clip, as_, ab)
        self._graphs = {}
        if self._handle is not None:
            self._apply_io_normalization()

    def _apply_io_normalization(self):
        io = self.__dict__.get("_io_norm")
        lib = _lib.load()
        if io is None:
            self.__dict__["_io_norm_dev"] = None
            lib.pz_set_io_normalization(self._handle, None, None, 0, None, None)
            return
        dev = self._packed[0][0].device
        ps, pb, clip, as_, ab = io
        t = [torch.tensor(v, dtype=torch.float32, device=dev) for v in (ps, pb, as_, ab)]
        self.__dict__["_io_norm_dev"] = t
        lib.pz_set_io_normalization(self._handle, t[0].data_ptr(), t[1].data_ptr(), 1 if clip else 0, t[2].data_ptr(), t[3].data_ptr())

    def _destroy_handle(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().pz_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:
            pass

    def release_unpacked_parameters(self):
        """Free the reference-layout parameters after packing (inference-only
        deployments: halves the resident weight memory).  state_dict() is empty
        afterwards."""
        self.pack()
        for mod in self.modules():
            for name in list(mod._parameters):
                mod._parameters[name] = nn.Parameter(torch.empty(0, device="cpu"), requires_grad=False)
        self.__dict__.pop("_param_list", None)
        self._packed_key = self._param_key()

    def _ensure_workspace(self, batch: int):
        lib = _lib.load()
        if self._workspace is None or batch > self._ws_batch:
            cap = batch
            nbytes = lib.pz_workspace_bytes(self._handle, cap)
            dev = self._packed[0][0].device
            self._workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
            self._ws_batch = cap
            self._ws_bytes = nbytes
        # every call uses the layout of its own batch size (a prefix of the buffer)
        base = (self._workspace.data_ptr() + 1023) // 1024 * 1024
        return base, lib.pz_workspace_bytes(self._handle, batch)

    # -------------------------------------------------------------- inference
    def _check_position_ids(self, B, vlm_pos, proprio_pos, action_pos):
        """The kernels rotate with the canonical positions (1.., pizero.py:312-318); anything else must not silently give
        the canonical result.  One device comparison per tensor, one host read for all of them."""
        Sv, Sp, H = self.max_image_text_tokens, self.num_proprio_tokens, self.num_action_tokens
        ok = None
        for got, lo, n, nm in ((vlm_pos, 1, Sv, "vlm"), (proprio_pos, 1, Sp, "proprio"), (action_pos, Sp + 1, H, "action")):
            if got is None:
                continue
            if tuple(got.shape) != (B, n):
                raise ValueError(f"{nm}_position_ids must be [{B}, {n}], got {tuple(got.shape)}")
            good = (got == torch.arange(lo, lo + n, device=got.device, dtype=got.dtype)).all()
            ok = good if ok is None else (ok & good.to(ok.device))
        if ok is not None and not bool(ok):
            raise ValueError("position ids differ from build_causal_mask_and_position_ids (vlm 1.., proprio 1.., action "
                             "after the proprio positions): non-canonical position ids are not supported")

    def _check_masks(self, B, image_text_proprio_mask, action_mask):
        """PZ_CHECK_INPUTS=1: the dense masks must be the block masks of build_causal_mask_and_position_ids for some
        per-sample valid length (the kernels only take the valid length from them)."""
        Sv = self.max_image_text_tokens
        cnt = (image_text_proprio_mask[:, 0, 0, :Sv] == 0).sum(-1)
        am = (torch.arange(Sv, device=cnt.device)[None, :] < cnt[:, None]).to(torch.int64)
        full, _, _, _ = self.build_causal_mask_and_position_ids(am, torch.float32)
        pm, acm = self.split_full_mask_into_submasks(full)
        if not torch.equal(image_text_proprio_mask == 0, pm.to(image_text_proprio_mask.device) == 0):
            raise ValueError("image_text_proprio_mask is not a block mask of build_causal_mask_and_position_ids")
        if action_mask is not None and not torch.equal(action_mask == 0, acm.to(action_mask.device) == 0):
            raise ValueError("action_mask is not a block mask of build_causal_mask_and_position_ids")

    def _valid_len(self, image_text_proprio_mask, input_ids, valid_len):
        if valid_len is not None:
            return valid_len.to(device=input_ids.device, dtype=torch.int32).contiguous()
        if image_text_proprio_mask is not None:
            # row 0 is an image token: its visible columns are exactly the valid image/text
            # positions (pizero.py:296-300)
            Sv = self.max_image_text_tokens
            return (image_text_proprio_mask[:, 0, 0, :Sv] == 0).sum(-1, dtype=torch.int32).to(input_ids.device).contiguous()
        return (input_ids != self.pad_token_id).sum(-1, dtype=torch.int32).contiguous()

    @torch.no_grad()
    def infer_action(
        self,
        input_ids: torch.LongTensor,
        pixel_values: torch.FloatTensor,
        image_text_proprio_mask: Optional[torch.FloatTensor] = None,
        action_mask: Optional[torch.FloatTensor] = None,
        vlm_position_ids: Optional[torch.LongTensor] = None,
        proprio_position_ids: Optional[torch.LongTensor] = None,
        action_position_ids: Optional[torch.LongTensor] = None,
        proprios: Optional[torch.FloatTensor] = None,
        *,
        noise: Optional[torch.Tensor] = None,
        valid_len: Optional[torch.Tensor] = None,
        capture: Optional[dict] = None,
    ) -> torch.FloatTensor:
        """pizero.py:416-490.  The eight reference arguments keep their names and
        meaning.  The dense masks are only used to read the per-sample valid
        length (the block structure is applied inside the attention kernels,
        SURVEY.md F8); position ids are the canonical ones
        `build_causal_mask_and_position_ids` returns (checked when
        PZ_CHECK_INPUTS=1).  Extras: `noise` ([B,H,A]; default: torch.randn as
        the reference), `valid_len` (int32 [B], skips the masks entirely),
        `capture` (dict filled with per-layer tensors for parity tests; `capture["action"]` is always the fp32 state).
        Returns `[B, horizon, action_dim]` in the dtype of `pixel_values` like the reference (pizero.py:454-456), or in
        `self.action_dtype` when that is set (torch.float32: the sampler's fp32 state, unrounded).
        Run-to-run: the prefill's split-K reductions are fp32 atomics / TMA reduce-adds, so two calls on the same inputs
        may differ in the last bits of the cache (<= ~2e-3 in the action); the bs <= 2 stream sampler itself is
        bit-reproducible (tests/test_gpu_sampler.py)."""
        if proprios is None:
            raise TypeError("infer_action() missing required argument: 'proprios'")
        self.pack()
        lib = _lib.load()
        d = self.dims
        dev = self._packed[0][0].device
        B = input_ids.shape[0]
        Sv, H, Adim = self.max_image_text_tokens, self.horizon_steps, self.action_dim
        if input_ids.shape != (B, Sv):
            raise ValueError(f"input_ids must be [B, {Sv}], got {tuple(input_ids.shape)}")
        n_img = d.get("num_images", 1)
        pix_elems = n_img * 3 * d["image_size"] ** 2
        if pixel_values.shape[0] != B or pixel_values[0].numel() != pix_elems:
            raise ValueError(f"pixel_values must be [B, {'%d, ' % n_img if n_img > 1 else ''}3, "
                             f"{d['image_size']}, {d['image_size']}], got {tuple(pixel_values.shape)}")
        if proprios.shape != (B, self.num_proprio_tokens, self.proprio_dim):
            raise ValueError(f"proprios must be [B, {self.num_proprio_tokens}, {self.proprio_dim}]")
        if self.check_inputs not in ("0", "off"):
            self._check_position_ids(B, vlm_position_ids, proprio_position_ids, action_position_ids)
            if self.check_inputs == "1" and image_text_proprio_mask is not None:
                self._check_masks(B, image_text_proprio_mask, action_mask)
        ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
        # uint8 camera frames are normalised on the device while the patches are gathered
        # (processing.py:27-58,108-113 fused into the im2col kernel); floats are the reference's argument
        u8 = pixel_values.dtype == torch.uint8
        pix = pixel_values.to(device=dev).contiguous() if u8 else pixel_values.to(device=dev, dtype=self._T).contiguous()
        prop = proprios.to(device=dev, dtype=torch.float32).contiguous()
        vlen = self._valid_len(image_text_proprio_mask, ids, valid_len)
        if noise is None:   # pizero.py:454-456
            noise = torch.randn((B, H, Adim), device=dev, dtype=self._T if u8 else pixel_values.dtype)
        nz = noise.to(device=dev, dtype=torch.float32).contiguous()
        out_dtype = self.action_dtype
        if out_dtype is None:   # pizero.py:454-456: the action chunk has the dtype of the pixel values
            out_dtype = pixel_values.dtype if pixel_values.is_floating_point() else self._T
        if capture is None and self.use_cuda_graph and not self._timing_armed:
            return self._replay_graph(B, ids, pix, vlen, prop, nz).to(out_dtype)
        out = torch.empty((B, H, Adim), device=dev, dtype=torch.float32)
        ws, ws_bytes = self._ensure_workspace(B)
        cap_struct, cap_bufs = None, None
        if capture is not None:
            cap_struct, cap_bufs = self._make_capture(B, dev)
        self._launch(ids, pix, vlen, prop, nz, out, ws, ws_bytes, B, cap_struct)
        if capture is not None:
            capture.update(cap_bufs)
            capture["action"] = out
            capture["kv"] = self.kv_caches(B)
        # keep the inputs alive until the stream has consumed them
        self._inflight = (ids, pix, prop, vlen, nz)
        return out.to(out_dtype)

    # ------------------------------------------------------------- text output
    def build_causal_mask_and_position_ids_for_text(self, q_len: int, attention_mask: torch.Tensor, kv_cache=None):
        """pizero.py:338-372 with the batch size taken from `attention_mask` (the reference reads an undefined `bsz`,
        SURVEY F11): nothing is masked (no padding), positions = running count of the mask, 1 for pad tokens."""
        dtype, device = attention_mask.dtype, attention_mask.device
        bsz = attention_mask.shape[0]
        if kv_cache is None or kv_cache.num_items() == 0:
            causal_mask = torch.full((bsz, q_len, q_len), 0, dtype=dtype, device=device)
        else:
            assert q_len == 1, "Using KV cache so should only use one single token"
            causal_mask = torch.full((bsz, q_len, kv_cache.num_items() + q_len), 0, dtype=dtype, device=device)
        causal_mask = causal_mask.unsqueeze(1)
        if kv_cache is not None and kv_cache.num_items() > 0:
            position_ids = attention_mask.cumsum(-1)[:, -1:]
        else:
            position_ids = (attention_mask.cumsum(-1)).masked_fill_((attention_mask == 0), 1)
        return causal_mask, position_ids

    @torch.no_grad()
    def infer_text(self, input_ids: torch.LongTensor, pixel_values: torch.FloatTensor, attention_mask: torch.Tensor,
                   kv_cache: Optional[TextKVCache] = None, *, last_token_only: bool = False) -> dict:
        """pizero.py:559-593: image + text through the vlm mixture alone (all layers, final norm) and the tied lm_head.
        Needs `use_lm_head` and `mixture.vlm.use_final_norm` in the config (what the reference's `--text_only` run sets,
        pizero.py:712-714).  First call (no cache, or an empty `TextKVCache`): the whole prompt, `q_len <=
        max_image_text_tokens` unpadded tokens per sample; `logits` is `[B, q_len, vocab]` fp32 (or `[B, 1, vocab]`, the
        last prompt token, with `last_token_only=True`: 1 GB less of lm_head output per sample at the full vocabulary).
        Later calls: ONE new token per sample (`input_ids [B, 1]`, `attention_mask` covering prompt + generated tokens),
        appended to the cache in place; `logits [B, 1, vocab]`.  Returns {"logits": ..., "kv_cache": ...} (the cache only
        when one was passed, like the reference)."""
        if not self.use_lm_head or "joint_model.mixtures.vlm.norm.weight" not in dict(self.named_parameters()):
            raise PzError("infer_text needs use_lm_head=True and mixture.vlm.use_final_norm=True in the config")
        if kv_cache is not None and not isinstance(kv_cache, TextKVCache):
            raise TypeError("kv_cache must be an open_pi_zero_b200 TextKVCache (or None)")
        self.pack()
        lib = _lib.load()
        d = self.dims
        dev = self._packed[0][0].device
        B, q_len = input_ids.shape
        Sv, Hd, V = self.max_image_text_tokens, d["vlm_hidden"], d["vocab_size"]
        ws, ws_bytes = self._ensure_workspace(B)
        cache = kv_cache
        if cache is None or cache.num_items() == 0:
            if q_len > Sv:
                raise ValueError(f"the prompt has {q_len} tokens, max_image_text_tokens is {Sv}")
            if not bool((attention_mask != 0).all()):
                raise ValueError("infer_text assumes unpadded prompts (pizero.py:346-357)")
            if cache is None:
                cache = TextKVCache()
            cap = Sv + self.text_max_new_tokens
            if cache.capacity < cap or cache.k.shape[1] != B or cache.k.dtype != self._T:
                cache._allocate(d["num_layers"], B, cap, d["head_dim"], self._T, dev)
            ids = torch.full((B, Sv), self.pad_token_id, dtype=torch.int64, device=dev)
            ids[:, :q_len] = input_ids.to(dev)
            u8 = pixel_values.dtype == torch.uint8
            pix = pixel_values.to(device=dev).contiguous() if u8 else pixel_values.to(device=dev, dtype=self._T).contiguous()
            vlen = torch.full((B,), q_len, dtype=torch.int32, device=dev)
            logits = torch.empty((B, 1 if last_token_only else Sv, V), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                lib.pz_set_pixel_format(self._handle, 1 if u8 else 0)
                rc = lib.pz_embed_prefix(self._handle, ids.data_ptr(), pix.data_ptr(), ws, ws_bytes, B, None, stream)
                if rc == 0:
                    rc = lib.pz_text_prefill(self._handle, vlen.data_ptr(), cache.k.data_ptr(), cache.v.data_ptr(), cache.capacity,
                                             q_len, logits.data_ptr(), 1 if last_token_only else 0, None, None, ws, ws_bytes, B, stream)
            if rc != 0:
                raise PzError(f"infer_text prefill failed ({rc}): {lib.pz_last_error(self._handle).decode()}")
            cache.length = q_len
            self._inflight = (ids, pix, vlen)
            out = {"logits": logits if last_token_only else logits[:, :q_len]}
        else:
            if q_len != 1:
                raise ValueError("Using KV cache so should only use one single token")
            cur = cache.num_items()
            if cur + 1 > cache.capacity:
                raise PzError(f"the text KV cache is full ({cache.capacity} rows): raise PZ_TEXT_MAX_NEW_TOKENS / text_max_new_tokens")
            # joint_model.py:348-355: embedding scaled by sqrt(hidden) (rounded to the model dtype like the reference)
            emb = self.embed_tokens.weight[input_ids[:, 0].to(dev)]
            x = (emb * torch.tensor(Hd ** 0.5, dtype=emb.dtype, device=dev)).to(torch.float32).contiguous()
            vlen1 = torch.full((B,), cur + 1, dtype=torch.int32, device=dev)
            logits = torch.empty((B, 1, V), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                rc = lib.pz_text_decode(self._handle, x.data_ptr(), vlen1.data_ptr(), cur, cache.k.data_ptr(), cache.v.data_ptr(),
                                        cache.capacity, logits.data_ptr(), None, ws, ws_bytes, B, stream)
            if rc != 0:
                raise PzError(f"infer_text decode failed ({rc}): {lib.pz_last_error(self._handle).decode()}")
            cache.length = cur + 1
            self._inflight = (x, vlen1)
            out = {"logits": logits}
        self.last_launch_count = int(lib.pz_launch_count(self._handle))
        if kv_cache is not None:
            out["kv_cache"] = cache
        else:
            out["_cache"] = cache   # not in the reference's dict: lets a caller continue decoding without pre-building a cache
        return out

    @torch.no_grad()
    def generate_text(self, input_ids, pixel_values, attention_mask, max_new_tokens: int = 16, eos_token_id: Optional[int] = None):
        """Greedy decoding on top of infer_text (the loop of the reference's `--text_only` run, pizero.py:770-800):
        returns the generated ids `[B, n]`."""
        cache = TextKVCache()
        out = self.infer_text(input_ids, pixel_values, attention_mask, cache, last_token_only=True)
        mask = attention_mask
        new = []
        for _ in range(max_new_tokens):
            nxt = out["logits"][:, -1].argmax(-1, keepdim=True)
            new.append(nxt)
            if eos_token_id is not None and bool((nxt == eos_token_id).all()):
                break
            mask = torch.cat([mask, torch.ones_like(mask[:, :1])], dim=-1)
            out = self.infer_text(nxt, pixel_values, mask, cache)
        return torch.cat(new, dim=1)

    @torch.no_grad()
    def infer_action_naive(
        self,
        input_ids: torch.LongTensor,
        pixel_values: torch.FloatTensor,
        causal_mask: torch.FloatTensor,
        vlm_position_ids: Optional[torch.LongTensor] = None,
        proprio_position_ids: Optional[torch.LongTensor] = None,
        action_position_ids: Optional[torch.LongTensor] = None,
        proprios: Optional[torch.FloatTensor] = None,
        *,
        noise: Optional[torch.Tensor] = None,
    ) -> torch.FloatTensor:
        """pizero.py:492-550: the reference's second entry point (same seven arguments, the full
        `[B,1,S,S]` mask instead of the two sub-masks).  The reference re-runs the VLM in every Euler step
        there ("which is unnecessary", pizero.py:517); under the block mask the VLM / proprio rows do not
        depend on the action rows, so the result is the cached path's (1.5e-7 apart in fp32) and
        the same kernels serve it.  The valid length is read from row 0 of the mask."""
        Sv = self.max_image_text_tokens
        vlen = (causal_mask[:, 0, 0, :Sv] == 0).sum(-1, dtype=torch.int32).contiguous()
        return self.infer_action(input_ids, pixel_values, None, None, vlm_position_ids, proprio_position_ids,
                                 action_position_ids, proprios, noise=noise, valid_len=vlen)

    def _launch(self, ids, pix, vlen, prop, nz, out, ws, ws_bytes, B, cap_struct=None):
        lib = _lib.load()
        dev = out.device
        for t in (ids, pix, vlen, prop, nz):
            if t.device != dev:
                raise PzError(f"internal: input on {t.device}, model on {dev}")
        # kernels are launched into a stream of `dev`: it must be the current device (the reference's callers do
        # `.to(f"cuda:{gpu_id}")` without set_device, eval.py / try_checkpoint_in_simpler.py)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            lib.pz_set_pixel_format(self._handle, 1 if pix.dtype == torch.uint8 else 0)
            rc = lib.pz_infer_action(self._handle, ids.data_ptr(), pix.data_ptr(), vlen.data_ptr(),
                                     prop.data_ptr(), nz.data_ptr(), out.data_ptr(), ws, ws_bytes, B,
                                     C.byref(cap_struct) if cap_struct is not None else None, stream)
        if rc != 0:
            raise PzError(f"pz_infer_action failed ({rc}): {lib.pz_last_error(self._handle).decode()}")
        self.last_launch_count = int(lib.pz_launch_count(self._handle))

    # ------------------------------------------------------------- CUDA graphs
    def _replay_graph(self, B, ids, pix, vlen, prop, nz):
        """The whole call (~2k kernel launches at bs=1) is captured once per batch
        size into a CUDA graph over static buffers and replayed: launch overhead
        leaves the critical path (eager dispatch is launch-bound)."""
        key = (B, pix.dtype)
        g = self._graphs.pop(key, None)
        if g is not None:
            self._graphs[key] = g          # most recently used last
        if g is None:
            while len(self._graphs) >= max(self.max_graphs, 1):   # least recently used first (dicts keep insertion order)
                old = self._graphs.pop(next(iter(self._graphs)))
                old.clear()
            lib = _lib.load()
            dev = ids.device
            # static buffers must be ordinary tensors even if the first call runs under torch.inference_mode(): a later
            # call under plain no_grad copies into them in place
            with torch.inference_mode(False):
                st = dict(ids=torch.empty_like(ids), pix=torch.empty_like(pix), vlen=torch.empty_like(vlen),
                          prop=torch.empty_like(prop), nz=torch.empty_like(nz),
                          out=torch.empty((B, self.horizon_steps, self.action_dim), device=dev, dtype=torch.float32))
                nbytes = lib.pz_workspace_bytes(self._handle, B)
                st["ws_t"] = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
            st["ws"] = (st["ws_t"].data_ptr() + 1023) // 1024 * 1024
            st["ws_bytes"] = nbytes
            for k, v in (("ids", ids), ("pix", pix), ("vlen", vlen), ("prop", prop), ("nz", nz)):
                st[k].copy_(v)
            # one eager run first: one-time function attributes / lazy module loading must not
            # happen inside a capture
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self._launch(st["ids"], st["pix"], st["vlen"], st["prop"], st["nz"], st["out"], st["ws"],
                             st["ws_bytes"], B)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._launch(st["ids"], st["pix"], st["vlen"], st["prop"], st["nz"], st["out"], st["ws"],
                             st["ws_bytes"], B)
            st["graph"] = graph
            st["launches"] = self.last_launch_count
            self._graphs[key] = g = st
        for k, v in (("ids", ids), ("pix", pix), ("vlen", vlen), ("prop", prop), ("nz", nz)):
            g[k].copy_(v, non_blocking=True)
        g["graph"].replay()
        self.last_launch_count = g["launches"]
        self._last_graph_B = B
        return g["out"].clone()

    def _make_capture(self, B, dev):
        d = self.dims
        L, T_, Sv, Sp, Hz = d["num_layers"], d["num_inference_steps"], d["max_image_text_tokens"], d["cond_steps"], d["horizon_steps"]
        Mv = B * d.get("num_images", 1) * d["num_image_tokens"]
        z = lambda *s: torch.zeros(s, device=dev, dtype=torch.float32)   # noqa: E731
        bufs = dict(
            vit_out=z(Mv, d["vit_hidden"]), image_features=z(Mv, d["vlm_hidden"]),
            prefix_embeds=z(B, Sv, d["vlm_hidden"]), prefix_vlm=z(max(L - 1, 1), B, Sv, d["vlm_hidden"]),
            prefix_proprio=z(max(L - 1, 1), B, Sp, d["act_hidden"]),
            denoise_action=z(T_, L, B, Hz, d["act_hidden"]), velocities=z(T_, B, Hz, d["action_dim"]),
            action_preclip=z(B, Hz, d["action_dim"]))
        cs = _lib.PzCapture()
        for k, t in bufs.items():
            setattr(cs, k, t.data_ptr())
        return cs, bufs

    def kv_caches(self, batch: int) -> dict:
        """The prefix KV of the last call as reference-style caches
        (`joint_model.build_mixture_caches`, joint_model.py:325)."""
        lib = _lib.load()
        d = self.dims
        ko, vo, ls = C.c_size_t(), C.c_size_t(), C.c_size_t()
        lib.pz_kv_layout(self._handle, batch, C.byref(ko), C.byref(vo), C.byref(ls))
        base = (self._workspace.data_ptr() + 1023) // 1024 * 1024 - self._workspace.data_ptr()
        es = 2 if self._T == torch.bfloat16 else 4
        Sc = d["max_image_text_tokens"] + d["cond_steps"]
        n = d["num_layers"] * batch * Sc * d["head_dim"]

        def view(off):
            raw = self._workspace[base + off: base + off + n * es]
            return raw.view(self._T).view(d["num_layers"], batch, Sc, d["head_dim"])

        k, v = view(ko.value), view(vo.value)
        Sv = d["max_image_text_tokens"]
        return {"vlm": KVCache(k, v, 0, Sv, d["num_layers"]),
                "proprio": KVCache(k, v, Sv, Sc, d["num_layers"])}

    # kernel-family timing taps (bench.py roofline object)
    TAG_VLM_GATE_UP, TAG_VLM_DOWN, TAG_ACT_GATE_UP = 1, 2, 3

    def timing_begin(self, tag: int):
        """Arms event timing of one kernel family; calls run eagerly (not from a
        captured graph) until timing_end()."""
        self.pack()
        self._timing_armed = True
        _lib.load().pz_timing_begin(self._handle, tag)

    def timing_end(self):
        ms, n = C.c_double(), C.c_int64()
        rc = _lib.load().pz_timing_end(self._handle, C.byref(ms), C.byref(n))
        self._timing_armed = False
        if rc != 0:
            raise PzError("pz_timing_end failed")
        return ms.value, n.value

    # ------------------------------------------------- flow-matching training forward (value only)
    def psi_t(self, x: torch.Tensor, x1: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """pizero.py:597-605 (conditional flow); host-side helper, the loss call computes it on the device."""
        t = t[:, None, None]
        return (1 - (1 - self.flow_sig_min) * t) * x + t * x1

    @torch.no_grad()
    def forward(
        self,
        input_ids: torch.LongTensor,
        pixel_values: torch.Tensor,
        causal_mask: Optional[torch.Tensor] = None,
        vlm_position_ids: Optional[torch.LongTensor] = None,
        proprio_position_ids: Optional[torch.LongTensor] = None,
        action_position_ids: Optional[torch.LongTensor] = None,
        proprios: Optional[torch.Tensor] = None,
        actions: Optional[torch.Tensor] = None,
        t: Optional[torch.Tensor] = None,
        *,
        noise: Optional[torch.Tensor] = None,
        valid_len: Optional[torch.Tensor] = None,
        return_velocity: bool = False,
    ):
        """pizero.py:607-661: the flow-matching loss `mean((v_psi - (x1 - (1 - sig_min) x0))^2)` for a batch,
        same nine arguments.  FORWARD VALUE ONLY: the result carries no autograd graph (the backward pass of the
        training step, SURVEY 8f-1, is not built) -- it serves validation / loss monitoring and pins the joint
        all-mixtures-active pass against the reference.  The reference's single joint pass (no cache, full block
        mask) is computed as prefix pass + one action pass over the cached prefix, which is the same function
        because vlm / proprio rows never attend to action keys (pizero.py:271-310).
        Extras: `noise` (x0; default torch.randn_like(actions) as pizero.py:622), `valid_len` (int32 [B], instead
        of reading it from row 0 of `causal_mask`), `return_velocity` (also return v_psi [B,H,A])."""
        if proprios is None or actions is None or t is None:
            raise TypeError("forward() needs proprios, actions and t")
        self.pack()
        lib = _lib.load()
        dev = self._packed[0][0].device
        B = input_ids.shape[0]
        Sv, H, Adim = self.max_image_text_tokens, self.horizon_steps, self.action_dim
        if input_ids.shape != (B, Sv):
            raise ValueError(f"input_ids must be [B, {Sv}], got {tuple(input_ids.shape)}")
        if actions.shape != (B, H, Adim):
            raise ValueError(f"actions must be [B, {H}, {Adim}], got {tuple(actions.shape)}")
        if t.shape != (B,):
            raise ValueError(f"t must be [B], got {tuple(t.shape)}")
        if proprios.shape != (B, self.num_proprio_tokens, self.proprio_dim):
            raise ValueError(f"proprios must be [B, {self.num_proprio_tokens}, {self.proprio_dim}]")
        ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
        u8 = pixel_values.dtype == torch.uint8
        pix = pixel_values.to(device=dev).contiguous() if u8 else pixel_values.to(device=dev, dtype=self._T).contiguous()
        prop = proprios.to(device=dev, dtype=torch.float32).contiguous()
        if valid_len is None and causal_mask is not None:
            valid_len = (causal_mask[:, 0, 0, :Sv] == 0).sum(-1, dtype=torch.int32)
        vlen = self._valid_len(None, ids, valid_len)
        x1 = actions.to(device=dev, dtype=torch.float32).contiguous()
        if noise is None:   # pizero.py:622
            noise = torch.randn_like(x1)
        x0 = noise.to(device=dev, dtype=torch.float32).contiguous()
        tt = t.to(device=dev, dtype=torch.float32).contiguous()
        loss = torch.empty((), device=dev, dtype=torch.float32)
        vel = torch.empty((B, H, Adim), device=dev, dtype=torch.float32) if return_velocity else None
        ws, ws_bytes = self._ensure_workspace(B)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            lib.pz_set_pixel_format(self._handle, 1 if u8 else 0)
            rc = lib.pz_flow_matching_loss(self._handle, ids.data_ptr(), pix.data_ptr(), vlen.data_ptr(), prop.data_ptr(),
                                           x1.data_ptr(), x0.data_ptr(), tt.data_ptr(), float(self.flow_sig_min),
                                           loss.data_ptr(), vel.data_ptr() if vel is not None else None, ws, ws_bytes, B,
                                           stream)
        if rc != 0:
            raise PzError(f"pz_flow_matching_loss failed ({rc}): {lib.pz_last_error(self._handle).decode()}")
        self.last_launch_count = int(lib.pz_launch_count(self._handle))
        self._inflight = (ids, pix, prop, vlen, x1, x0, tt)
        return (loss, vel) if return_velocity else loss


class PiZeroInference(PiZero):
    """pizero.py:664-686: `forward` is `infer_action`, so `model(**inputs)` works."""

    def forward(self, input_ids, pixel_values, image_text_proprio_mask=None, action_mask=None,
                vlm_position_ids=None, proprio_position_ids=None, action_position_ids=None,
                proprios=None, **extra):
        return super().infer_action(input_ids, pixel_values, image_text_proprio_mask, action_mask,
                                    vlm_position_ids, proprio_position_ids, action_position_ids,
                                    proprios, **extra)
