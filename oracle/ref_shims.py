"""TEST INFRASTRUCTURE ONLY -- import shims that let the *unmodified* reference
(`/root/reference`, shroglck/open-pi-zero) be imported in this image.

The reference cannot be imported as shipped: `hydra`, `omegaconf` and
`bitsandbytes` are absent.  None of the three carries arithmetic
on the default inference path, so three non-arithmetic stand-ins are installed
in `sys.modules`:

* `omegaconf.OmegaConf.merge`  -> shallow dict merge, second argument wins
  (reference call site: src/model/vla/joint_model.py:321)
* `hydra.utils.instantiate`    -> import `_target_`, call with the other keys
  (reference call sites: src/model/vla/pizero.py:68,69,72)
* `bitsandbytes.nn.{Params4bit,Linear4bit}` -> empty subclasses, only needed
  because src/model/lora.py:214,236 subclass them at import time.

Nothing here is used by the product path.  It is used by
`oracle/make_golden.py` (to generate `tests/golden/*`) and by the CPU tests
that pin `oracle/pizero_oracle.py` against the reference when
`/root/reference` exists (it does not exist on the GPU box).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PZ_REFERENCE_ROOT", "/root/reference")


class RefCfg(dict):
    """dict with attribute access, applied recursively (stands in for DictConfig)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, RefCfg):
                self[k] = RefCfg(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "model", "vla"))


def install() -> None:
    """Install the three shims and put the reference on sys.path (idempotent)."""
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class OmegaConf:
            @staticmethod
            def merge(a, b):
                out = RefCfg(dict(a))
                out.update(dict(b))
                return RefCfg(out)

        oc.OmegaConf = OmegaConf
        oc.DictConfig = RefCfg
        sys.modules["omegaconf"] = oc
    if "hydra" not in sys.modules:
        hy = types.ModuleType("hydra")
        hu = types.ModuleType("hydra.utils")

        def instantiate(node):
            modname, clsname = node["_target_"].rsplit(".", 1)
            cls = getattr(importlib.import_module(modname), clsname)
            return cls(**{k: v for k, v in node.items() if k != "_target_"})

        hu.instantiate = instantiate
        hy.utils = hu
        sys.modules["hydra"] = hy
        sys.modules["hydra.utils"] = hu
    if "bitsandbytes" not in sys.modules:
        import torch

        bnb = types.ModuleType("bitsandbytes")
        bnn = types.ModuleType("bitsandbytes.nn")

        class Params4bit(torch.nn.Parameter):
            pass

        class Linear4bit(torch.nn.Linear):
            pass

        bnn.Params4bit = Params4bit
        bnn.Linear4bit = Linear4bit
        bnb.nn = bnn
        sys.modules["bitsandbytes"] = bnb
        sys.modules["bitsandbytes.nn"] = bnn
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_cfg(dims: dict) -> RefCfg:
    """Build the already-resolved config tree the reference's PiZero reads
    (fields: config/train/bridge.yaml:85-181; pizero.py:32-103) from the flat
    `dims` dict used throughout this repo (see open-pi-zero_b200/config.py)."""
    d = dims
    mixture = {
        "vlm": dict(hidden_size=d["vlm_hidden"], intermediate_size=d["vlm_inter"],
                    use_final_norm=bool(d.get("vlm_use_final_norm", False)), cache=True, use_quantize=False, use_lora=False,
                    adaptive_mode=None, rope_theta=d["vlm_rope_theta"]),
        "proprio": dict(hidden_size=d["act_hidden"], intermediate_size=d["act_inter"],
                        use_final_norm=True, cache=True, use_quantize=False, use_lora=False,
                        adaptive_mode=None, rope_theta=d["act_rope_theta"]),
        "action": dict(hidden_size=d["act_hidden"], intermediate_size=d["act_inter"],
                       use_final_norm=True, cache=False, use_quantize=False, use_lora=False,
                       adaptive_mode=None, rope_theta=d["act_rope_theta"]),
    }
    return RefCfg(
        vocab_size=d["vocab_size"], pad_token_id=d["pad_token_id"],
        image_token_index=d["image_token_index"],
        max_image_text_tokens=d["max_image_text_tokens"],
        cond_steps=d["cond_steps"], horizon_steps=d["horizon_steps"],
        num_inference_steps=d["num_inference_steps"], action_dim=d["action_dim"],
        proprio_dim=d["proprio_dim"], final_action_clip_value=d["final_action_clip_value"],
        flow_sig_min=0.001, use_lm_head=bool(d.get("use_lm_head", False)),
        action_expert_adaptive_mode=None, time_hidden_size=256,
        time_max_period=d["time_max_period"], mixture=mixture,
        vision=dict(_target_="src.model.paligemma.siglip.SiglipVisionModel",
                    config=dict(hidden_size=d["vit_hidden"], intermediate_size=d["vit_inter"],
                                num_hidden_layers=d["vit_layers"],
                                num_attention_heads=d["vit_heads"], num_channels=3,
                                image_size=d["image_size"], patch_size=d["patch_size"],
                                layer_norm_eps=1e-6, attention_dropout=0.0,
                                num_image_tokens=d["num_image_tokens"]),
                    use_quantize=False, use_lora=False),
        vision_projector=dict(
            _target_="src.model.paligemma.siglip.PaliGemmaMultiModalProjector",
            config=dict(vision_config=dict(hidden_size=d["vit_hidden"],
                                           projection_dim=d["vlm_hidden"])),
            use_quantize=False, use_lora=False),
        joint=dict(_target_="src.model.vla.joint_model.JointModel",
                   config=dict(action_expert_adaptive_mode=None, time_hidden_size=256,
                               mixture=mixture, num_hidden_layers=d["num_layers"],
                               num_attention_heads=d["num_heads"],
                               num_key_value_heads=d["num_kv_heads"], head_dim=d["head_dim"],
                               rms_norm_eps=1e-6, attention_bias=False, attention_dropout=0.0,
                               pad_token_id=d["pad_token_id"])),
    )


def build_reference_model(dims: dict):
    """Instantiate the reference's own PiZero (fp32, CPU, eval)."""
    install()
    from src.model.vla.pizero import PiZero  # noqa: the reference's class

    model = PiZero(reference_cfg(dims))
    model.eval()
    return model
