"""TEST INFRASTRUCTURE ONLY -- generate `tests/golden/*.pt` from the UNMODIFIED
reference (`/root/reference`, imported via `oracle/ref_shims.py`).

Run here (the authoring container), not on the GPU box:
    python oracle/make_golden.py [tiny] [width2] [bridge]

Cases
  tiny    every dimension shrunk; the fixture holds the full state dict, the
          inputs and every captured intermediate -> pins oracle/pizero_oracle.py
          against the reference with no reference installed.
  width2  real widths (2048/16384, 1024/4096, 1152/4304, head_dim 256), two
          SigLIP layers and two joint layers, small vocabulary.  Weights are
          NOT stored: they are rebuilt from the seed by
          `open-pi-zero_b200/synth.py::init_state_dict` (same torch build on
          both boxes).  Stored: actions, velocities and row-subsampled hidden
          states / KV from the reference.
  bridge  the full bridge config (27 SigLIP + 18 joint layers, vocab 257216,
          3.24 B parameters), same storage policy as width2.
  bridge64  the same model at batch 64 (the batch bench.py times): the actions of
          every sample; hidden-state / KV rows of the samples listed in `samples`.

How outputs are captured without touching the reference:
`torch.randn` is patched for the duration of the call to return the shared
noise (pizero.py:454 draws it through the module-global `torch`), and
`src.model.vla.joint_model.forward_mixture_layers` is wrapped to record the
per-layer hidden dicts it returns.
"""
import importlib.util
import os
import sys
import time

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

TINY = config.make_dims(
    vocab_size=320, image_token_index=300, max_image_text_tokens=10, num_image_tokens=4,
    num_layers=2, num_heads=4, num_kv_heads=1, head_dim=16, vlm_hidden=64, vlm_inter=128,
    act_hidden=32, act_inter=64, vit_hidden=32, vit_inter=64, vit_layers=2, vit_heads=2,
    image_size=28, patch_size=14)
WIDTH2 = config.make_dims(vocab_size=1024, image_token_index=1000, num_layers=2, vit_layers=2)
BRIDGE = config.make_dims()

CASES = {
    "tiny": dict(dims=TINY, batch=3, seed=7, store_weights=True, randomize_norms=True),
    "width2": dict(dims=WIDTH2, batch=2, seed=11, store_weights=False, randomize_norms=True),
    "bridge": dict(dims=BRIDGE, batch=2, seed=42, store_weights=False, randomize_norms=False),
    # the batch the bench times (BASELINE configs[1]): actions of all 64 samples, sampled rows of four of them
    "bridge64": dict(dims=BRIDGE, batch=64, seed=44, store_weights=False, randomize_norms=True,
                     samples=[0, 21, 42, 63]),
}


def run_reference(dims, sd, inp):
    """Run the reference's own PiZero.infer_action with injected noise and
    per-layer capture.  Returns a dict of captured tensors."""
    ref_shims.install()
    import src.model.vla.joint_model as jm

    model = ref_shims.build_reference_model(dims)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    dtype = torch.float32
    cm, vpos, ppos, apos = model.build_causal_mask_and_position_ids(inp["attention_mask"], dtype)
    pmask, amask = model.split_full_mask_into_submasks(cm)
    cap = {"layers": []}
    orig_fml = jm.forward_mixture_layers

    def fml(*a, **kw):
        out = orig_fml(*a, **kw)
        cap["layers"].append({k: (None if v is None else v.clone()) for k, v in out.items()})
        return out

    orig_randn = torch.randn

    def randn(*a, **kw):
        return inp["noise"].to(kw.get("dtype", torch.float32)).clone()

    cache_holder = {}
    orig_build = model.joint_model.build_mixture_caches

    def build():
        c = orig_build()
        cache_holder.setdefault("kv", c)  # keep infer_action's caches, not the naive run's
        return c

    model.joint_model.build_mixture_caches = build
    model.final_action_clip_value = None  # capture the pre-clamp action (pizero.py:484)
    jm.forward_mixture_layers = fml
    torch.randn = randn
    try:
        with torch.inference_mode():
            t0 = time.time()
            pre = model.infer_action(
                input_ids=inp["input_ids"], pixel_values=inp["pixel_values"],
                image_text_proprio_mask=pmask, action_mask=amask, vlm_position_ids=vpos,
                proprio_position_ids=ppos, action_position_ids=apos, proprios=inp["proprios"])
            dt = time.time() - t0
            # also the reference's own uncached path
            naive = model.infer_action_naive(
                input_ids=inp["input_ids"], pixel_values=inp["pixel_values"], causal_mask=cm,
                vlm_position_ids=vpos, proprio_position_ids=ppos, action_position_ids=apos,
                proprios=inp["proprios"]) if dims["num_layers"] <= 2 else None
            vit = model.vision_tower(inp["pixel_values"])
            feats = model.multi_modal_projector(vit)
            emb = model._forward_siglip_and_text_embedding(inp["input_ids"], inp["pixel_values"])
    finally:
        jm.forward_mixture_layers = orig_fml
        torch.randn = orig_randn
    L = dims["num_layers"]
    out = dict(
        action_preclip=pre.clone(), vit_out=vit.clone(), image_features=feats.clone(),
        prefix_embeds=emb.clone(), prefix_layers=cap["layers"][:L],
        denoise_layers=cap["layers"][L:L + L * dims["num_inference_steps"]],
        kv={n: [(k.clone(), v.clone()) for k, v in zip(c.key_cache, c.value_cache)]
            for n, c in cache_holder["kv"].items()},
        naive_preclip=naive, seconds=dt)
    clip = dims["final_action_clip_value"]
    out["action"] = pre.clamp(-clip, clip) if clip is not None else pre.clone()
    return out


def sample_rows(valid_len, Sv):
    """Row indices stored for the big cases: a few image rows, the bos row,
    the last valid row of each sample."""
    rows = []
    for c in valid_len.tolist():
        rows.append(sorted({0, 1, 100 % Sv, 255 % Sv, min(256, c - 1), c - 1}))
    return rows


def main(argv):
    names = argv or ["tiny", "width2"]
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in names:
        case = CASES[name]
        dims = case["dims"]
        t0 = time.time()
        sd = synth.init_state_dict(dims, seed=case["seed"], randomize_norms=case["randomize_norms"])
        inp = synth.make_inputs(dims, case["batch"], seed=case["seed"] + 100)
        print(f"[{name}] weights+inputs built in {time.time()-t0:.1f}s", flush=True)
        ref = run_reference(dims, sd, inp)
        print(f"[{name}] reference infer_action {ref['seconds']:.2f}s", flush=True)
        fx = dict(case=name, dims=dims, seed=case["seed"], batch=case["batch"],
                  randomize_norms=case["randomize_norms"], reference="shroglck/open-pi-zero",
                  torch=torch.__version__)
        if case["store_weights"]:
            fx["state_dict"] = sd
            fx["inputs"] = inp
            fx["ref"] = ref
        else:
            Sv = dims["max_image_text_tokens"]
            samples = case.get("samples") or list(range(case["batch"]))
            all_rows = sample_rows(inp["valid_len"], Sv)
            rows = [all_rows[b] for b in samples]
            L, T = dims["num_layers"], dims["num_inference_steps"]
            sub = lambda t: [t[b, all_rows[b]].clone() for b in samples]  # noqa: E731
            H = dims["horizon_steps"]
            fx["inputs_seed"] = case["seed"] + 100
            fx["rows"] = rows
            fx["samples"] = samples
            fx["ref"] = dict(
                action=ref["action"], action_preclip=ref["action_preclip"],
                naive_preclip=ref["naive_preclip"],
                vit_rows=[0, 1, 100, 255],
                vit_out_rows=ref["vit_out"][samples][:, [0, 1, 100, 255]].clone(),
                image_features_rows=ref["image_features"][samples][:, [0, 1, 100, 255]].clone(),
                prefix_embeds_rows=sub(ref["prefix_embeds"]),
                prefix_layers_rows=[
                    dict(vlm=None if l["vlm"] is None else sub(l["vlm"]),
                         proprio=None if l["proprio"] is None else l["proprio"][samples].clone())
                    for l in ref["prefix_layers"]],
                kv_rows=dict(
                    vlm=[(sub(k[:, 0]), sub(v[:, 0])) for k, v in ref["kv"]["vlm"]],
                    proprio=[(k[samples].clone(), v[samples].clone()) for k, v in ref["kv"]["proprio"]]),
                denoise_layers={s: [ref["denoise_layers"][s * L + l]["action"][samples].clone()
                                    for l in range(L)] for s in (0, T - 1)},
                # norms of every prefix layer's valid rows, for a cheap whole-tensor check
                prefix_layer_norms=[
                    None if l["vlm"] is None else
                    torch.stack([l["vlm"][b, : int(inp["valid_len"][b])].norm()
                                 for b in range(case["batch"])])
                    for l in ref["prefix_layers"]],
            )
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(fx, path)
        print(f"[{name}] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB)", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
