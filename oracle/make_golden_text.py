"""TEST INFRASTRUCTURE ONLY -- `tests/golden/text_tiny.pt`: logits of the UNMODIFIED reference's `PiZero.infer_text`
(pizero.py:559-593) for a prompt prefill and three greedy decode steps on a tiny configuration.

`infer_text` is broken as shipped: `build_causal_mask_and_position_ids_for_text` reads an undefined name `bsz`
(pizero.py:349,355).  Nothing of the reference is edited: the name is INJECTED into the reference module's
globals (`src.model.vla.pizero.bsz = batch`) for the duration of the call, which is exactly the value the author meant.

    python oracle/make_golden_text.py
"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


config = _load("pz_config", "open-pi-zero_b200/config.py")
synth = _load("pz_synth", "open-pi-zero_b200/synth.py")

DIMS = config.make_dims(
    vocab_size=320, image_token_index=300, max_image_text_tokens=10, num_image_tokens=4,
    num_layers=2, num_heads=4, num_kv_heads=1, head_dim=16, vlm_hidden=64, vlm_inter=128,
    act_hidden=32, act_inter=64, vit_hidden=32, vit_inter=64, vit_layers=2, vit_heads=2,
    image_size=28, patch_size=14, use_lm_head=True, vlm_use_final_norm=True)


# the parity tests' kernel-compatible small shape (tests/helpers.py SMALL: real attention geometry) with the text head on
SMALL_TEXT = config.make_dims(
    vocab_size=1024, image_token_index=1000, image_size=56, num_image_tokens=16,
    max_image_text_tokens=24, num_layers=3, vlm_hidden=256, vlm_inter=512, act_hidden=128,
    act_inter=256, vit_hidden=144, vit_inter=256, vit_layers=2, vit_heads=2, use_lm_head=True, vlm_use_final_norm=True)


def make(case, DIMS, B, q_len, steps, store_weights):
    ref_shims.install()
    import src.model.vla.pizero as ref_pizero
    from src.model.kv_cache import KVCache
    sd = synth.init_state_dict(DIMS, seed=5, randomize_norms=True)
    model = ref_shims.build_reference_model(DIMS)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    g = torch.Generator().manual_seed(9)
    n_img, px = DIMS["num_image_tokens"], DIMS["image_size"]
    ids = torch.cat([torch.full((B, n_img), DIMS["image_token_index"]),
                     torch.randint(1, DIMS["image_token_index"] - 10, (B, q_len - n_img), generator=g)], 1)
    pix = (torch.randint(0, 256, (B, 3, px, px), generator=g, dtype=torch.uint8) * (1 / 255.0) - 0.5) / 0.5
    mask = torch.ones((B, q_len), dtype=torch.int64)
    ref_pizero.bsz = B   # the undefined name (see the module docstring)
    logits, tokens = [], []
    try:
        with torch.inference_mode():
            cache = KVCache()
            out = model.infer_text(ids, pix, mask, cache)
            logits.append(out["logits"].clone())
            for _ in range(steps):
                nxt = logits[-1][:, -1].argmax(-1, keepdim=True)
                tokens.append(nxt.clone())
                mask = torch.cat([mask, torch.ones_like(mask[:, :1])], 1)
                out = model.infer_text(nxt, pix, mask, cache)
                logits.append(out["logits"].clone())
            kv = [(k.clone(), v.clone()) for k, v in zip(cache.key_cache, cache.value_cache)]
    finally:
        del ref_pizero.bsz
    fx = dict(case=case, dims=DIMS, seed=5, randomize_norms=True, input_ids=ids, pixel_values=pix,
              q_len=q_len, logits=logits, tokens=tokens, kv=kv, reference="shroglck/open-pi-zero", torch=torch.__version__)
    if store_weights:   # otherwise rebuilt from the seed by synth.init_state_dict
        fx["state_dict"] = sd
    path = os.path.join(ROOT, "tests", "golden", case + ".pt")
    torch.save(fx, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB); prefill logits {tuple(logits[0].shape)}, decode {tuple(logits[1].shape)}")


def main():
    make("text_tiny", DIMS, B=2, q_len=8, steps=3, store_weights=True)
    make("text_small", SMALL_TEXT, B=3, q_len=21, steps=4, store_weights=False)   # the CUDA kernels' smallest shape


if __name__ == "__main__":
    main()
