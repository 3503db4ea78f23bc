"""Model dimensions for the `PiZero.infer_action` path.

Every value restates `config/train/bridge.yaml:85-181` of the reference
(shroglck/open-pi-zero); the YAML/hydra machinery itself is not rebuilt.  Two
views of the same data are offered:

* a flat ``dims`` dict (what the kernels, the oracle and the tests use), and
* ``cfg_from_dims`` -- the nested attribute tree with the reference's field
  names (`cfg.mixture.vlm.hidden_size`, `cfg.joint.config.head_dim`, ...) so
  that ``PiZero(cfg)`` keeps the reference's constructor contract
  (`src/model/vla/pizero.py:30-103`).  ``dims_from_cfg`` goes the other way and
  accepts an OmegaConf DictConfig just as well as our AttrDict.
"""
from __future__ import annotations

import copy

BRIDGE_DIMS = dict(
    # tokens / sequence (bridge.yaml:88-94,129-131)
    vocab_size=257216, pad_token_id=0, image_token_index=257152,
    max_image_text_tokens=276, num_image_tokens=256, num_images=1,
    cond_steps=1, horizon_steps=4, action_dim=7, proprio_dim=7,
    # sampler (bridge.yaml:85-86)
    num_inference_steps=10, final_action_clip_value=1.0,
    # flow matching (pizero.py:58: cfg.get("flow_sig_min", 0.001))
    flow_sig_min=0.001,
    # joint transformer (bridge.yaml:174-179)
    num_layers=18, num_heads=8, num_kv_heads=1, head_dim=256,
    # mixtures (bridge.yaml:96-126)
    vlm_hidden=2048, vlm_inter=16384, vlm_rope_theta=10000.0,
    act_hidden=1024, act_inter=4096, act_rope_theta=100.0,
    time_max_period=100.0,
    # SigLIP-So400m/14 (bridge.yaml:136-145)
    vit_hidden=1152, vit_inter=4304, vit_layers=27, vit_heads=16,
    image_size=224, patch_size=14,
    # optional text output (pizero.py:37,105-112; the reference's --text_only run sets both, pizero.py:712-714)
    use_lm_head=False, vlm_use_final_norm=False,
)

# Pi0-paper shape (BASELINE.json configs[3]): 3 images, 48 text tokens, chunk 50.
PI0_PAPER_DIMS = dict(BRIDGE_DIMS, num_images=3, max_image_text_tokens=3 * 256 + 48,
                      horizon_steps=50)


def make_dims(base: dict | None = None, **overrides) -> dict:
    d = copy.deepcopy(BRIDGE_DIMS if base is None else base)
    unknown = set(overrides) - set(d)
    if unknown:
        raise KeyError(f"unknown dims: {sorted(unknown)}")
    d.update(overrides)
    return d


class AttrDict(dict):
    """dict with attribute access and `.get`, recursive -- same access surface
    the reference uses on its DictConfig."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, AttrDict):
                self[k] = AttrDict(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def cfg_from_dims(d: dict) -> AttrDict:
    def mix(hidden, inter, final_norm, cache, theta):
        return dict(hidden_size=hidden, intermediate_size=inter, use_final_norm=final_norm,
                    cache=cache, use_quantize=False, use_lora=False, adaptive_mode=None,
                    rope_theta=theta)

    mixture = dict(
        vlm=mix(d["vlm_hidden"], d["vlm_inter"], bool(d.get("vlm_use_final_norm", False)), True, d["vlm_rope_theta"]),
        proprio=mix(d["act_hidden"], d["act_inter"], True, True, d["act_rope_theta"]),
        action=mix(d["act_hidden"], d["act_inter"], True, False, d["act_rope_theta"]),
    )
    return AttrDict(
        vocab_size=d["vocab_size"], pad_token_id=d["pad_token_id"],
        image_token_index=d["image_token_index"],
        max_image_text_tokens=d["max_image_text_tokens"], max_seq_len=d["max_image_text_tokens"],
        cond_steps=d["cond_steps"], horizon_steps=d["horizon_steps"],
        num_inference_steps=d["num_inference_steps"], action_dim=d["action_dim"],
        proprio_dim=d["proprio_dim"], final_action_clip_value=d["final_action_clip_value"],
        action_expert_adaptive_mode=None, time_hidden_size=256,
        time_max_period=d["time_max_period"], num_images=d.get("num_images", 1),
        flow_sig_min=d.get("flow_sig_min", 0.001),
        use_lm_head=bool(d.get("use_lm_head", False)),
        mixture=mixture,
        vision=dict(config=dict(hidden_size=d["vit_hidden"], intermediate_size=d["vit_inter"],
                                num_hidden_layers=d["vit_layers"],
                                num_attention_heads=d["vit_heads"], num_channels=3,
                                image_size=d["image_size"], patch_size=d["patch_size"],
                                layer_norm_eps=1e-6, attention_dropout=0.0,
                                num_image_tokens=d["num_image_tokens"]),
                    use_quantize=False, use_lora=False),
        vision_projector=dict(config=dict(vision_config=dict(hidden_size=d["vit_hidden"],
                                                             projection_dim=d["vlm_hidden"])),
                              use_quantize=False, use_lora=False),
        joint=dict(config=dict(action_expert_adaptive_mode=None, time_hidden_size=256,
                               mixture=mixture, num_hidden_layers=d["num_layers"],
                               num_attention_heads=d["num_heads"],
                               num_key_value_heads=d["num_kv_heads"], head_dim=d["head_dim"],
                               rms_norm_eps=1e-6, attention_bias=False, attention_dropout=0.0,
                               pad_token_id=d["pad_token_id"])),
    )


def _g(node, key, default=None):
    if hasattr(node, "get"):
        v = node.get(key, default)
    else:
        v = getattr(node, key, default)
    return v


def dims_from_cfg(cfg) -> dict:
    """Flatten a reference-style config tree (DictConfig / AttrDict / dict)."""
    if "vlm_hidden" in cfg:  # already flat
        return make_dims(**{k: cfg[k] for k in cfg if k in BRIDGE_DIMS})
    mixture = cfg["mixture"]
    vis = cfg["vision"]["config"]
    joint = cfg["joint"]["config"]
    for name in ("vlm", "proprio", "action"):
        m = mixture[name]
        if _g(m, "use_quantize", False) or _g(m, "use_lora", False):
            raise NotImplementedError("LoRA / 4-bit layers are out of scope (SURVEY.md section 2)")
        if _g(m, "adaptive_mode", None):
            raise NotImplementedError("adaLN action expert is out of scope (SURVEY.md 8f-3)")
    if _g(cfg, "action_expert_adaptive_mode", None):
        raise NotImplementedError("adaLN action expert is out of scope (SURVEY.md 8f-3)")
    if mixture["proprio"]["hidden_size"] != mixture["action"]["hidden_size"]:
        raise ValueError("proprio and action experts must share their width")
    n_img_tok = _g(vis, "num_image_tokens", (vis["image_size"] // vis["patch_size"]) ** 2)
    return make_dims(
        vocab_size=cfg["vocab_size"], pad_token_id=cfg["pad_token_id"],
        image_token_index=cfg["image_token_index"],
        max_image_text_tokens=cfg["max_image_text_tokens"], num_image_tokens=n_img_tok,
        num_images=_g(cfg, "num_images", 1),
        cond_steps=cfg["cond_steps"], horizon_steps=cfg["horizon_steps"],
        action_dim=cfg["action_dim"], proprio_dim=cfg["proprio_dim"],
        num_inference_steps=cfg["num_inference_steps"],
        final_action_clip_value=cfg["final_action_clip_value"],
        flow_sig_min=float(_g(cfg, "flow_sig_min", 0.001)),
        num_layers=joint["num_hidden_layers"], num_heads=joint["num_attention_heads"],
        num_kv_heads=joint["num_key_value_heads"], head_dim=joint["head_dim"],
        vlm_hidden=mixture["vlm"]["hidden_size"], vlm_inter=mixture["vlm"]["intermediate_size"],
        vlm_rope_theta=float(mixture["vlm"]["rope_theta"]),
        act_hidden=mixture["action"]["hidden_size"],
        act_inter=mixture["action"]["intermediate_size"],
        act_rope_theta=float(mixture["action"]["rope_theta"]),
        time_max_period=float(cfg["time_max_period"]),
        vit_hidden=vis["hidden_size"], vit_inter=vis["intermediate_size"],
        vit_layers=vis["num_hidden_layers"], vit_heads=vis["num_attention_heads"],
        image_size=vis["image_size"], patch_size=vis["patch_size"],
        use_lm_head=bool(_g(cfg, "use_lm_head", False)),
        vlm_use_final_norm=bool(_g(mixture["vlm"], "use_final_norm", False)),
    )
