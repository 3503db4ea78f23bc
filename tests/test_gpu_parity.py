"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the
reference's golden vectors.  Tolerances (BASELINE.json north_star): fp32 final
action <= 1e-4 abs; bf16 per-layer hidden states <= 2e-2 relative (Frobenius,
valid rows only) and final action <= 1e-2 abs."""
import os

import pytest
import torch

from helpers import SMALL, max_abs, pz, rel_err, valid_rows
from oracle import pizero_oracle as O

pytestmark = pytest.mark.gpu

FP32_ACTION_TOL = 1e-4
FP32_LAYER_TOL = 1e-4      # relative
BF16_LAYER_TOL = 2e-2      # relative
BF16_ACTION_TOL = 1e-2     # absolute


def _model(d, sd, dtype):
    from open_pi_zero_b200.pizero import PiZeroInference
    m = PiZeroInference(pz.cfg_from_dims(d), init="empty")
    m.load_state_dict(sd, strict=True)
    m = m.to(dtype).to("cuda")
    m.action_dtype = torch.float32      # compare the sampler's fp32 state (the default is the reference's: pixel dtype)
    return m


def _run(m, d, inp, capture=True):
    mask, vp, pp, ap = m.build_causal_mask_and_position_ids(inp["attention_mask"].cuda(), torch.float32)
    pm, am = m.split_full_mask_into_submasks(mask)
    cap = {} if capture else None
    dt = next(m.parameters()).dtype
    out = m(input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda().to(dt),
            image_text_proprio_mask=pm, action_mask=am, vlm_position_ids=vp,
            proprio_position_ids=pp, action_position_ids=ap, proprios=inp["proprios"].cuda().to(dt),
            noise=inp["noise"].cuda(), capture=cap)
    torch.cuda.synchronize()
    return out, cap


def _compare_all(d, inp, cap, ocap, layer_tol, report):
    vl = inp["valid_len"]
    H = d["vlm_hidden"]
    worst = 0.0
    e = rel_err(cap["vit_out"].view_as(ocap["vit_out"]), ocap["vit_out"]); report["vit_out"] = e; worst = max(worst, e)
    e = rel_err(cap["image_features"].view_as(ocap["image_features"]), ocap["image_features"]); report["image_features"] = e; worst = max(worst, e)
    e = rel_err(valid_rows(cap["prefix_embeds"], vl), valid_rows(ocap["prefix_embeds"] * H ** 0.5, vl)); report["prefix_embeds"] = e; worst = max(worst, e)
    for l, want in enumerate(ocap["prefix_layers"]):
        if want["vlm"] is None:
            continue
        e = rel_err(valid_rows(cap["prefix_vlm"][l], vl), valid_rows(want["vlm"], vl)); report[f"prefix_vlm_{l}"] = e; worst = max(worst, e)
        e = rel_err(cap["prefix_proprio"][l], want["proprio"]); report[f"prefix_proprio_{l}"] = e; worst = max(worst, e)
    L = d["num_layers"]
    for i, want in enumerate(ocap["denoise_layers"]):
        e = rel_err(cap["denoise_action"][i // L, i % L], want["action"]); report[f"denoise_{i // L}_{i % L}"] = e; worst = max(worst, e)
    kv = cap["kv"]
    for n in ("vlm", "proprio"):
        for l in range(L):
            k, v = kv[n].get(l)
            k2, v2 = ocap["kv"][n][l]
            if n == "vlm":
                k, v, k2, v2 = (valid_rows(t[:, 0], vl) for t in (k, v, k2, v2))
            e = max(rel_err(k.float(), k2), rel_err(v.float(), v2)); report[f"kv_{n}_{l}"] = e; worst = max(worst, e)
    assert worst < layer_tol, {k: v for k, v in report.items() if v >= layer_tol}
    return worst


@pytest.mark.parametrize("dtype,layer_tol,act_tol", [
    (torch.float32, FP32_LAYER_TOL, FP32_ACTION_TOL),
    (torch.bfloat16, BF16_LAYER_TOL, BF16_ACTION_TOL)])
def test_small_every_layer_vs_oracle(dtype, layer_tol, act_tol):
    d = SMALL
    sd = pz.init_state_dict(d, seed=3, randomize_norms=True)
    inp = pz.make_inputs(d, 5, seed=21, min_text=0)
    ocap = {}
    want = O.infer_action(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                          inp["proprios"], inp["noise"], capture=ocap)
    m = _model(d, sd, dtype)
    out, cap = _run(m, d, inp)
    report = {}
    worst = _compare_all(d, inp, cap, ocap, layer_tol, report)
    pre = max_abs(cap["action_preclip"], ocap["action_preclip"])
    print(f"[small {dtype}] worst layer rel {worst:.3e}; preclip max-abs {pre:.3e}; "
          f"clamped {max_abs(out, want):.3e}; launches {m.last_launch_count}")
    assert pre < act_tol
    assert max_abs(out, want) < act_tol
    assert m.last_launch_count > 0


def test_small_fractal_config_proprio_dim8_vs_oracle():
    """The reference's second shipped robot config (config/eval/fractal_apple.yaml:49): proprio_dim = 8."""
    d = pz.make_dims(SMALL, proprio_dim=8)
    sd = pz.init_state_dict(d, seed=6, randomize_norms=True)
    inp = pz.make_inputs(d, 3, seed=12)
    assert inp["proprios"].shape[-1] == 8
    ocap = {}
    want = O.infer_action(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                          inp["proprios"], inp["noise"], capture=ocap)
    m = _model(d, sd, torch.bfloat16)
    out, cap = _run(m, d, inp)
    for l, w in enumerate(ocap["prefix_layers"]):
        if w["proprio"] is not None:
            assert rel_err(cap["prefix_proprio"][l], w["proprio"]) < BF16_LAYER_TOL
    assert max_abs(cap["action_preclip"], ocap["action_preclip"]) < BF16_ACTION_TOL
    assert max_abs(out, want) < BF16_ACTION_TOL


@pytest.mark.parametrize("batch", [1, 3, 6])   # 1, 3: persistent sampler (MT = 1 / 2); 6: separate kernels
def test_small_no_clip_five_steps_vs_oracle(batch):
    """Config knobs off the bridge defaults: `final_action_clip_value: null` (pizero.py:484-489 skipped) and
    `num_inference_steps: 5` (the per-step time constants and dt follow), through both sampler implementations."""
    d = pz.make_dims(SMALL, final_action_clip_value=None, num_inference_steps=5)
    sd = pz.init_state_dict(d, seed=8, randomize_norms=True)
    inp = pz.make_inputs(d, batch, seed=14)
    inp["noise"] = inp["noise"] * 2.0          # push some components beyond +-1 so that a wrong clamp would show
    want = O.infer_action(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                          inp["proprios"], inp["noise"])
    assert float(want.abs().max()) > 1.0
    m = _model(d, sd, torch.bfloat16)
    out, _ = _run(m, d, inp, capture=False)
    assert max_abs(out, want) < 2 * BF16_ACTION_TOL   # values up to ~4: same relative bar as the clamped case


def test_small_eval_yaml_rope_and_time_periods_vs_oracle():
    """The released checkpoints' parameter set (config/eval/bridge.yaml:72-73, README.md:151): action / proprio RoPE theta and
    the time embedding's max period are 10 000 instead of the training yaml's 100."""
    d = pz.make_dims(SMALL, act_rope_theta=10000.0, time_max_period=10000.0)
    sd = pz.init_state_dict(d, seed=10, randomize_norms=True)
    inp = pz.make_inputs(d, 2, seed=16)
    ocap = {}
    want = O.infer_action(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                          inp["proprios"], inp["noise"], capture=ocap)
    m = _model(d, sd, torch.bfloat16)
    out, cap = _run(m, d, inp)
    L = d["num_layers"]
    for i, w in enumerate(ocap["denoise_layers"]):
        assert rel_err(cap["denoise_action"][i // L, i % L], w["action"]) < BF16_LAYER_TOL
    assert max_abs(out, want) < BF16_ACTION_TOL
    fast, _ = _run(m, d, inp, capture=False)      # the persistent sampler reads the same tables
    assert max_abs(fast, want) < BF16_ACTION_TOL


def test_small_simple_kernels_match_fast_path(monkeypatch):
    """The tcgen05 / mma / skinny kernels against the plain SIMT kernels, same bf16 inputs."""
    d = SMALL
    sd = pz.init_state_dict(d, seed=5, randomize_norms=True)
    inp = pz.make_inputs(d, 4, seed=8)
    fast = _model(d, sd, torch.bfloat16)
    a, cap_a = _run(fast, d, inp)
    monkeypatch.setenv("PZ_SIMPLE_KERNELS", "1")
    slow = _model(d, sd, torch.bfloat16)
    b, cap_b = _run(slow, d, inp)
    assert rel_err(cap_a["prefix_vlm"], cap_b["prefix_vlm"]) < 1e-2
    assert max_abs(cap_a["action_preclip"], cap_b["action_preclip"]) < 1e-2


def _golden(golden_dir, name):
    path = os.path.join(golden_dir, name + ".pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} missing")
    return torch.load(path, weights_only=False)


def _check_golden(fx, dtype, layer_tol, act_tol, tag):
    """CUDA vs tensors the UNMODIFIED reference produced (weights rebuilt from the seed)."""
    d = fx["dims"]
    sd = pz.init_state_dict(d, seed=fx["seed"], randomize_norms=fx["randomize_norms"])
    inp = pz.make_inputs(d, fx["batch"], seed=fx["inputs_seed"])
    m = _model(d, sd, dtype)
    del sd
    out, cap = _run(m, d, inp)
    ref = fx["ref"]
    worst = 0.0
    e = rel_err(cap["vit_out"].view(fx["batch"], -1, d["vit_hidden"])[:, ref["vit_rows"]], ref["vit_out_rows"])
    worst = max(worst, e)
    for b, rows in enumerate(fx["rows"]):
        for l, want in enumerate(ref["prefix_layers_rows"]):
            if want["vlm"] is not None:
                worst = max(worst, rel_err(cap["prefix_vlm"][l, b, rows], want["vlm"][b]))
        for l, (k2, v2) in enumerate(ref["kv_rows"]["vlm"]):
            k, v = cap["kv"]["vlm"].get(l)
            worst = max(worst, rel_err(k[b, 0, rows].float(), k2[b]), rel_err(v[b, 0, rows].float(), v2[b]))
    for s, layers in ref["denoise_layers"].items():
        for l, want in enumerate(layers):
            worst = max(worst, rel_err(cap["denoise_action"][s, l], want))
    pre = max_abs(cap["action_preclip"], ref["action_preclip"])
    clamped = max_abs(out, ref["action"])
    print(f"[{tag} {dtype}] worst sampled-layer rel {worst:.3e}; preclip max-abs {pre:.3e}; clamped {clamped:.3e}")
    assert worst < layer_tol
    assert pre < act_tol and clamped < act_tol
    # production path (no capture taps): CUDA-graph replay; at this batch size the sampler runs as the
    # persistent cooperative kernel (denoise_mega.cu) in bf16
    for _ in range(2):
        prod, _ = _run(m, d, inp, capture=False)
    e_prod = max_abs(prod, ref["action"])
    print(f"[{tag} {dtype}] production path (graph replay) clamped max-abs vs reference {e_prod:.3e}")
    assert e_prod < act_tol


@pytest.mark.parametrize("dtype,layer_tol,act_tol", [
    (torch.float32, 2e-4, FP32_ACTION_TOL), (torch.bfloat16, BF16_LAYER_TOL, BF16_ACTION_TOL)])
def test_width2_vs_reference_golden(golden_dir, dtype, layer_tol, act_tol):
    _check_golden(_golden(golden_dir, "width2"), dtype, layer_tol, act_tol, "width2")


def test_bridge_full_size_bf16_vs_reference_golden(golden_dir):
    """Full bridge config (3.24 B parameters, 27 + 18 layers) in bf16."""
    _check_golden(_golden(golden_dir, "bridge"), torch.bfloat16, BF16_LAYER_TOL, BF16_ACTION_TOL, "bridge")


def test_bridge64_production_path_vs_reference_golden(golden_dir):
    """The batch and the path bench.py times (BASELINE configs[1]: bs = 64, CTA-pair tcgen05 GEMMs with the grouped raster,
    64-sample prefix chunk, one kernel per op in the sampler, CUDA-graph replay, all 27 + 18 layers) against actions the
    UNMODIFIED reference produced for the same 64 observations (oracle/make_golden.py bridge64), plus sampled hidden-state
    and K/V rows of four of the samples through the eager path."""
    fx = _golden(golden_dir, "bridge64")
    d, B = fx["dims"], fx["batch"]
    sd = pz.init_state_dict(d, seed=fx["seed"], randomize_norms=fx["randomize_norms"])
    inp = pz.make_inputs(d, B, seed=fx["inputs_seed"])
    m = _model(d, sd, torch.bfloat16)
    del sd
    ref = fx["ref"]
    # production path first: graph capture on the first call, replay on the second
    for _ in range(2):
        prod, _ = _run(m, d, inp, capture=False)
    assert any(k[0] == B for k in m._graphs)
    e_prod = max_abs(prod, ref["action"])
    # eager path with capture taps
    out, cap = _run(m, d, inp)
    worst = 0.0
    for i, b in enumerate(fx["samples"]):
        rows = fx["rows"][i]
        worst = max(worst, rel_err(cap["vit_out"].view(B, -1, d["vit_hidden"])[b, ref["vit_rows"]], ref["vit_out_rows"][i]))
        for l, want in enumerate(ref["prefix_layers_rows"]):
            if want["vlm"] is not None:
                worst = max(worst, rel_err(cap["prefix_vlm"][l, b, rows], want["vlm"][i]))
        for l, (k2, v2) in enumerate(ref["kv_rows"]["vlm"]):
            k, v = cap["kv"]["vlm"].get(l)
            worst = max(worst, rel_err(k[b, 0, rows].float(), k2[i]), rel_err(v[b, 0, rows].float(), v2[i]))
    for s_, layers in ref["denoise_layers"].items():
        for l, want in enumerate(layers):
            worst = max(worst, rel_err(cap["denoise_action"][s_, l][fx["samples"]], want))
    pre = max_abs(cap["action_preclip"], ref["action_preclip"])
    print(f"[bridge64 bf16] graph-replay clamped max-abs vs reference {e_prod:.3e}; eager preclip {pre:.3e}; "
          f"worst sampled-layer rel {worst:.3e}")
    assert e_prod < BF16_ACTION_TOL
    assert pre < BF16_ACTION_TOL and max_abs(out, ref["action"]) < BF16_ACTION_TOL
    assert worst < BF16_LAYER_TOL


def test_pi0_paper_shape_multi_image_chunk50():
    """BASELINE configs[3] geometry: 3 images (768 image tokens), 48 text tokens, chunk 50 --
    a capability the reference lacks; oracle = the reference's vision tower applied
    per image + its unmodified joint model (oracle.embed_prefix, 5-D pixel_values)."""
    d = pz.make_dims(pz.PI0_PAPER_DIMS, vocab_size=1024, image_token_index=1000, num_layers=2, vit_layers=2)
    sd = pz.init_state_dict(d, seed=13, randomize_norms=True)
    inp = pz.make_inputs(d, 2, seed=5, min_text=3)
    ocap = {}
    want = O.infer_action(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                          inp["proprios"], inp["noise"], capture=ocap)
    for dtype, layer_tol, act_tol in ((torch.float32, FP32_LAYER_TOL, FP32_ACTION_TOL),
                                      (torch.bfloat16, BF16_LAYER_TOL, BF16_ACTION_TOL)):
        m = _model(d, sd, dtype)
        out, cap = _run(m, d, inp)
        report = {}
        worst = _compare_all(d, inp, cap, ocap, layer_tol, report)
        pre = max_abs(cap["action_preclip"], ocap["action_preclip"])
        print(f"[pi0-shape {dtype}] worst layer rel {worst:.3e}; preclip max-abs {pre:.3e}")
        assert pre < act_tol and max_abs(out, want) < act_tol
        del m


def test_cuda_graph_replay_matches_eager(monkeypatch):
    """The production path (CUDA-graph replay, no capture taps) against the eager call."""
    d = SMALL
    sd = pz.init_state_dict(d, seed=2, randomize_norms=True)
    m = _model(d, sd, torch.bfloat16)
    for seed in (1, 2):
        inp = pz.make_inputs(d, 3, seed=seed)
        eager, _ = _run(m, d, inp, capture=True)      # capture taps force the eager path
        eager = eager.clone()
        replay, _ = _run(m, d, inp, capture=False)    # graph (captured on first use, replayed after)
        assert max_abs(eager, replay) < 2e-3           # fp32 atomics in the split-K GEMV: order may differ
    assert any(k[0] == 3 for k in m._graphs)


def test_uint8_frames_match_host_normalisation():
    """raw uint8 camera frames, normalised inside the patch-gather kernel, give exactly the
    result of the reference's host-side VLAProcessor normalisation (processing.py:27-58,108-113) followed
    by the float path.  The kernel applies the same fp32 operations in the same order (u * f32(1/255), - 0.5,
    / 0.5, no FMA contraction), so the pixels are bit-identical; the comparison of the final actions still
    needs a small tolerance because two runs of the path differ in the order of their fp32 reduce-adds."""
    d = SMALL
    sd = pz.init_state_dict(d, seed=4, randomize_norms=True)
    inp = pz.make_inputs(d, 3, seed=9)
    g = torch.Generator().manual_seed(11)
    u8 = torch.randint(0, 256, tuple(inp["pixel_values"].shape), dtype=torch.uint8, generator=g)
    host = (u8 * (1 / 255.0) - 0.5) / 0.5           # rescale() then normalize() of processing.py
    for dtype in (torch.float32, torch.bfloat16):
        m = _model(d, sd, dtype)
        kw = dict(input_ids=inp["input_ids"].cuda(), proprios=inp["proprios"].cuda().to(dtype),
                  noise=inp["noise"].cuda(), valid_len=inp["valid_len"].cuda())
        for _ in range(2):                           # second call: CUDA-graph replay for both formats
            want = m(pixel_values=host.cuda().to(dtype), **kw).clone()
            got = m(pixel_values=u8.cuda(), **kw).clone()
            torch.cuda.synchronize()
            assert torch.isfinite(got).all()
            assert max_abs(got, want) < 2e-3, max_abs(got, want)
        del m


def test_infer_action_naive_entry_point():
    """pizero.py:492-550 (full-mask signature) against the oracle's own uncached implementation."""
    d = SMALL
    sd = pz.init_state_dict(d, seed=6, randomize_norms=True)
    inp = pz.make_inputs(d, 2, seed=12)
    want = O.infer_action_naive(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                                inp["proprios"], inp["noise"])
    m = _model(d, sd, torch.float32)
    mask, vp, pp, ap = m.build_causal_mask_and_position_ids(inp["attention_mask"].cuda(), torch.float32)
    got = m.infer_action_naive(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), mask, vp, pp, ap,
                               inp["proprios"].cuda(), noise=inp["noise"].cuda())
    assert max_abs(got, want) < FP32_ACTION_TOL


def test_large_m_pair_gemm_and_fused_rope_vs_oracle():
    """B = 9 at the real widths gives M = 2484 prefix rows: the CTA-pair (cta_group::2) GEMM variant,
    including its fused RoPE / KV-cache epilogue, against the oracle."""
    fx_dims = pz.make_dims(vocab_size=1024, image_token_index=1000, num_layers=2, vit_layers=2)
    sd = pz.init_state_dict(fx_dims, seed=31, randomize_norms=True)
    inp = pz.make_inputs(fx_dims, 9, seed=6)
    ocap = {}
    want = O.infer_action(sd, fx_dims, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                          inp["proprios"], inp["noise"], capture=ocap)
    m = _model(fx_dims, sd, torch.bfloat16)
    out, cap = _run(m, fx_dims, inp)
    report = {}
    worst = _compare_all(fx_dims, inp, cap, ocap, BF16_LAYER_TOL, report)
    print(f"[pair-gemm width2 B=9] worst layer rel {worst:.3e}; clamped {max_abs(out, want):.3e}")
    assert max_abs(out, want) < BF16_ACTION_TOL


def test_prefix_chunking_matches_unchunked(monkeypatch):
    """Large batches run the prefix pass in sub-batches (PZ_PREFIX_CHUNK, default 64) against one
    KV cache; a ragged last chunk and the chunk boundaries must not change any sample."""
    d = SMALL
    sd = pz.init_state_dict(d, seed=23, randomize_norms=True)
    inp = pz.make_inputs(d, 7, seed=12, min_text=0)
    whole = _model(d, sd, torch.bfloat16)
    a, cap_a = _run(whole, d, inp)
    monkeypatch.setenv("PZ_PREFIX_CHUNK", "3")      # 7 samples -> chunks of 3, 3, 1
    chunked = _model(d, sd, torch.bfloat16)
    b, cap_b = _run(chunked, d, inp)
    vl = inp["valid_len"]
    # tolerance: one bf16 ulp here and there (fp32 atomics in the split-K GEMV sum in arrival order)
    assert rel_err(valid_rows(cap_a["prefix_vlm"][-1], vl), valid_rows(cap_b["prefix_vlm"][-1], vl)) < 5e-3
    for l in range(d["num_layers"]):
        ka, va = cap_a["kv"]["proprio"].get(l)
        kb, vb = cap_b["kv"]["proprio"].get(l)
        assert rel_err(ka.float(), kb.float()) < 5e-3 and rel_err(va.float(), vb.float()) < 5e-3
    assert max_abs(a, b) < 2e-3
    # production path (graph replay) with chunking
    c, _ = _run(chunked, d, inp, capture=False)
    assert max_abs(a, c) < 2e-3


def test_joint_model_forward_api():
    """The inner boundary: JointModel.forward for the prefix pass (fills the caches)
    and the action pass over them, against the oracle's joint_forward on the same embeddings."""
    d = SMALL
    sd = pz.init_state_dict(d, seed=17, randomize_norms=True)
    m = _model(d, sd, torch.bfloat16)
    jm = m.joint_model
    B, Sv, Sp, Hz = 3, d["max_image_text_tokens"], d["cond_steps"], d["horizon_steps"]
    g = torch.Generator().manual_seed(0)
    ev = torch.randn(B, Sv, d["vlm_hidden"], generator=g) * 0.05
    ep = torch.randn(B, Sp, d["act_hidden"], generator=g) * 0.05
    ea = torch.randn(B, Hz, d["act_hidden"], generator=g) * 0.05
    attn = torch.ones(B, Sv, dtype=torch.int64)
    attn[0, 18:] = 0
    attn[2, 21:] = 0
    for b in range(B):
        ev[b, int(attn[b].sum()):] = 0
    _, pmask, amask, pos = O.build_masks_and_positions(d, attn, torch.float32)
    okv = {"vlm": [], "proprio": []}
    O.joint_forward(sd, d, pmask, {"vlm": pos["vlm"], "proprio": pos["proprio"]},
                    {"vlm": ev.clone(), "proprio": ep.clone()}, okv)
    want = O.joint_forward(sd, d, amask, {"action": pos["action"]}, {"action": ea.clone()}, okv)["action"]
    caches = jm.build_mixture_caches()
    assert not caches["vlm"].has_item(0)
    ev_c, ep_c = ev.cuda(), ep.cuda()
    out, caches = jm(attention_mask=pmask.cuda(), position_ids_all={"vlm": pos["vlm"].cuda(), "proprio": pos["proprio"].cuda()},
                     embeds_all={"vlm": ev_c, "proprio": ep_c}, kv_caches=caches, return_caches=True)
    assert out == {} and caches["vlm"].has_item(d["num_layers"] - 1) and caches["vlm"].num_items() == Sv
    assert max_abs(ev_c, ev * d["vlm_hidden"] ** 0.5) < 1e-4          # scaled in place, like the reference
    vl = attn.sum(1)
    for l in range(d["num_layers"]):
        k, v = caches["vlm"].get(l)
        k2, v2 = okv["vlm"][l]
        assert rel_err(valid_rows(k[:, 0].float(), vl), valid_rows(k2[:, 0], vl)) < BF16_LAYER_TOL
        assert rel_err(valid_rows(v[:, 0].float(), vl), valid_rows(v2[:, 0], vl)) < BF16_LAYER_TOL
    got = jm(attention_mask=amask.cuda(), position_ids_all={"action": pos["action"].cuda()},
             embeds_all={"action": ea.cuda()}, kv_caches=caches, cache_mode="append_non_active")["action"]
    torch.cuda.synchronize()
    e = rel_err(got, want)
    print(f"[JointModel.forward] action hidden rel err {e:.3e}")
    assert e < BF16_LAYER_TOL
    # (c) all three mixtures active, full mask, no cache: the training-forward / infer_action_naive call pattern
    full_mask, _, _, _ = O.build_masks_and_positions(d, attn, torch.float32)
    got3 = jm(attention_mask=full_mask.cuda(), position_ids_all={n: p.cuda() for n, p in pos.items()},
              embeds_all={"vlm": ev.cuda(), "proprio": ep.cuda(), "action": ea.cuda()}, kv_caches={})["action"]
    want3 = O.joint_forward(sd, d, full_mask, pos, {"vlm": ev.clone(), "proprio": ep.clone(), "action": ea.clone()}, {})["action"]
    torch.cuda.synchronize()
    e3 = rel_err(got3, want3)
    print(f"[JointModel.forward, all active] action hidden rel err {e3:.3e}")
    assert e3 < BF16_LAYER_TOL
    with pytest.raises(NotImplementedError):
        jm(attention_mask=amask.cuda(), position_ids_all={}, embeds_all={"vlm": ev.cuda(), "action": ea.cuda()})
    with pytest.raises(AssertionError):
        jm(attention_mask=amask.cuda(), position_ids_all={}, embeds_all={"action": ea.cuda()}, cache_mode="bogus")


def test_batch_invariance_and_ragged_lengths():
    """Each sample's result must not depend on its neighbours or on pad content:
    run B=6 with ragged lengths, then each sample alone."""
    d = SMALL
    sd = pz.init_state_dict(d, seed=9, randomize_norms=True)
    inp = pz.make_inputs(d, 6, seed=4, min_text=0)
    m = _model(d, sd, torch.bfloat16)
    full, _ = _run(m, d, inp, capture=False)
    full = full.clone()
    for b in (0, 3, 5):
        one = {k: (v[b:b + 1] if torch.is_tensor(v) else v) for k, v in inp.items()}
        out, _ = _run(m, d, one, capture=False)
        assert max_abs(out, full[b:b + 1]) < 2e-3


def test_full_size_bs1024_rollout_batch_properties():
    """BASELINE configs[2] at its full size on one GPU: the 3.24 B-parameter bridge model, bs = 1024, uint8 frames.
    No oracle finishes this in seconds, so the checks are size-independent properties: a sample's chunk does not
    depend on the batch it rides in (bs 1024 vs the same samples at bs 64 and, through the persistent sampler,
    bs 1), the clamp bound holds, and nothing is NaN/Inf."""
    from open_pi_zero_b200.pizero import PiZeroInference
    from open_pi_zero_b200.synth import fill_random_
    d = pz.make_dims()
    m = PiZeroInference(pz.cfg_from_dims(d), init="empty", device="cuda", dtype=torch.bfloat16)
    fill_random_(m, d, seed=42)
    m.action_dtype = torch.float32
    B = 1024
    inp = pz.make_inputs(d, B, seed=77)
    dev = {k: inp[k].cuda() for k in ("input_ids", "proprios", "noise", "valid_len")}
    dev["pixel_values"] = inp["pixel_u8"].cuda()
    big = m(**dev).clone()
    torch.cuda.synchronize()
    assert big.shape == (B, d["horizon_steps"], d["action_dim"])
    assert bool(torch.isfinite(big).all())
    assert float(big.abs().max()) <= d["final_action_clip_value"] + 1e-6
    for lo in (0, 960):
        part = m(**{k: v[lo:lo + 64].contiguous() for k, v in dev.items()}).clone()
        e = max_abs(part, big[lo:lo + 64])
        print(f"[bs1024] samples {lo}..{lo + 63}: max |bs1024 - bs64| = {e:.3e}")
        assert e < BF16_ACTION_TOL
    for b in (5, 1023):
        one = m(**{k: v[b:b + 1].contiguous() for k, v in dev.items()}).clone()
        e = max_abs(one, big[b:b + 1])
        print(f"[bs1024] sample {b}: max |bs1024 - bs1 (persistent sampler)| = {e:.3e}")
        assert e < BF16_ACTION_TOL


def test_drop_in_defaults_dtype_position_ids_and_inference_mode():
    """What a stock caller of the reference sees: the action chunk comes back in the dtype of the pixel values
    (pizero.py:454-456,484-490); position ids other than the canonical ones raise instead of silently giving the
    canonical result; a first call under torch.inference_mode() followed by one under no_grad works (static graph
    buffers are ordinary tensors); the graph cache is bounded."""
    from open_pi_zero_b200.pizero import PiZeroInference
    d = SMALL
    sd = pz.init_state_dict(d, seed=12, randomize_norms=True)
    m = PiZeroInference(pz.cfg_from_dims(d), init="empty")
    m.load_state_dict(sd, strict=True)
    m = m.to(torch.bfloat16).to("cuda")
    inp = pz.make_inputs(d, 2, seed=3)
    mask, vp, pp, ap = m.build_causal_mask_and_position_ids(inp["attention_mask"], torch.bfloat16)   # CPU masks, as the processor builds them
    pm, am = m.split_full_mask_into_submasks(mask)
    kw = dict(input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda().bfloat16(),
              image_text_proprio_mask=pm, action_mask=am, vlm_position_ids=vp.cuda(), proprio_position_ids=pp.cuda(),
              action_position_ids=ap.cuda(), proprios=inp["proprios"].cuda().bfloat16())
    with torch.inference_mode():
        a = m(**kw, noise=inp["noise"].cuda())
    assert a.dtype == torch.bfloat16 and a.shape == (2, d["horizon_steps"], d["action_dim"])
    with torch.no_grad():
        b = m(**kw, noise=inp["noise"].cuda())
    assert max_abs(a.float(), b.float()) < 1e-2
    m.action_dtype = torch.float32
    c = m(**kw, noise=inp["noise"].cuda())
    assert c.dtype == torch.float32 and max_abs(c, a.float()) < 1e-2
    bad = dict(kw, action_position_ids=ap.cuda() + 1)
    with pytest.raises(ValueError, match="position ids"):
        m(**bad)
    m.max_graphs = 2
    for bsz in (1, 2, 3, 4):
        i2 = pz.make_inputs(d, bsz, seed=bsz)
        m(input_ids=i2["input_ids"].cuda(), pixel_values=i2["pixel_values"].cuda().bfloat16(), proprios=i2["proprios"].cuda(),
          valid_len=i2["valid_len"].cuda())
    assert len(m._graphs) <= 2 and any(k[0] == 4 for k in m._graphs)


def test_unsupported_bf16_shape_is_an_error_not_a_silent_simt_fallback(monkeypatch):
    """A bf16 configuration with a shape no tensor-core kernel covers (head_dim 128) fails loudly; with PZ_ALLOW_FALLBACK=1 it runs on the SIMT kernels, is counted, and matches the oracle."""
    from open_pi_zero_b200 import _lib
    from open_pi_zero_b200.pizero import PiZeroInference, PzError
    d = pz.make_dims(SMALL, head_dim=128)    # real widths, but the attention kernels are built for head_dim 256
    sd = pz.init_state_dict(d, seed=2, randomize_norms=True)
    inp = pz.make_inputs(d, 2, seed=1)
    kw = dict(input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda().bfloat16(),
              proprios=inp["proprios"].cuda(), noise=inp["noise"].cuda(), valid_len=inp["valid_len"].cuda())

    def build():
        m = PiZeroInference(pz.cfg_from_dims(d), init="empty")
        m.load_state_dict(sd, strict=True)
        m = m.to(torch.bfloat16).to("cuda")
        m.use_cuda_graph = False
        m.action_dtype = torch.float32
        return m

    with pytest.raises(PzError, match="no tensor-core kernel"):
        build()(**kw)
    monkeypatch.setenv("PZ_ALLOW_FALLBACK", "1")
    m = build()
    out = m(**kw)
    assert _lib.load().pz_fallback_count(m._handle) > 0
    want = O.infer_action(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], inp["proprios"], inp["noise"])
    assert max_abs(out, want) < 2e-2


def test_errors_are_python_exceptions():
    from open_pi_zero_b200.pizero import PzError
    d = SMALL
    sd = pz.init_state_dict(d, seed=1)
    m = _model(d, sd, torch.bfloat16)
    inp = pz.make_inputs(d, 2, seed=1)
    with pytest.raises(ValueError):
        m(input_ids=inp["input_ids"][:, :-1].cuda(), pixel_values=inp["pixel_values"].cuda(),
          proprios=inp["proprios"].cuda())
    with pytest.raises(TypeError):
        m(input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda())
    from open_pi_zero_b200.pizero import PiZero
    with pytest.raises(TypeError):   # the training forward needs proprios, actions and t (pizero.py:607-618)
        PiZero.forward(m, input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda())
    with pytest.raises(ValueError):
        PiZero.forward(m, input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda(),
                       proprios=inp["proprios"].cuda(), actions=torch.zeros(2, 1, 1).cuda(), t=torch.zeros(2).cuda())
    assert PzError is not None


@pytest.mark.parametrize("kind", ["bound", "gaussian"])
def test_io_normalization_folded_into_the_kernels(kind):
    """set_io_normalization: raw proprio in, de-normalised actions out == the host-side adapter arithmetic
    (simpler.py:76-90, 102-125) around the plain call; eager and CUDA-graph paths."""
    import numpy as np
    from open_pi_zero_b200.adapter import BaseEnvAdapter
    d = SMALL
    sd = pz.init_state_dict(d, seed=3, randomize_norms=True)
    inp = pz.make_inputs(d, 3, seed=8)
    m = _model(d, sd, torch.bfloat16)
    rng = np.random.default_rng(1)
    stats = {k: dict(p01=rng.normal(size=7) - 2, p99=rng.normal(size=7) + 2, mean=rng.normal(size=7), std=rng.uniform(0.5, 2, 7))
             for k in ("proprio", "action")}
    raw = torch.tensor(rng.normal(size=(3, 1, 7)) * 2, dtype=torch.float32)
    ad = BaseEnvAdapter()
    if kind == "bound":
        norm = ad.normalize_bound(raw.numpy(), stats["proprio"]["p01"], stats["proprio"]["p99"])
    else:
        norm = ad.normalize_gaussian(raw.numpy(), stats["proprio"]["mean"], stats["proprio"]["std"])
    kw = dict(input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda().bfloat16(), noise=inp["noise"].cuda(),
              valid_len=inp["valid_len"].cuda())
    plain = m(proprios=torch.tensor(norm, dtype=torch.float32).cuda(), **kw).float().cpu().numpy()
    if kind == "bound":
        want = ad.denormalize_bound(plain[..., :-1], stats["action"]["p01"][:-1], stats["action"]["p99"][:-1])
    else:
        want = ad.denormalize_gaussian(plain[..., :-1], stats["action"]["mean"][:-1], stats["action"]["std"][:-1])
    want = np.concatenate([want, plain[..., -1:]], -1)
    m.set_io_normalization(stats, action_normalization_type=kind, proprio_normalization_type=kind)
    for _ in range(3):      # capture + replay
        got = m(proprios=raw.cuda(), **kw).float().cpu().numpy()
        assert np.abs(got - want).max() < 2e-2 * max(1.0, np.abs(want).max())
    m.set_io_normalization(None)
    back = m(proprios=torch.tensor(norm, dtype=torch.float32).cuda(), **kw).float().cpu().numpy()
    assert np.abs(back - plain).max() < 1e-2


def test_joint_model_forward_no_append_mode():
    """`cache_mode="no_append"` with all three mixtures active (infer_action_naive's call, pizero.py:529-544): the first call
    fills the vlm / proprio caches it was handed, later calls read them; both give what the cache-free joint pass gives."""
    d = SMALL
    sd = pz.init_state_dict(d, seed=6, randomize_norms=True)
    inp = pz.make_inputs(d, 2, seed=12)
    m = _model(d, sd, torch.float32)
    mask, vp, pp, ap = m.build_causal_mask_and_position_ids(inp["attention_mask"].cuda(), torch.float32)
    g = torch.Generator().manual_seed(1)
    B, Sv, H, A = 2, d["max_image_text_tokens"], d["vlm_hidden"], d["act_hidden"]
    ev = (torch.randn((B, Sv, H), generator=g) * 0.05).cuda()
    ep = (torch.randn((B, 1, A), generator=g) * 0.05).cuda()
    ea = [(torch.randn((B, d["horizon_steps"], A), generator=g) * 0.05).cuda() for _ in range(2)]
    pos = {"vlm": vp, "proprio": pp, "action": ap}
    fresh = [m.joint_model(attention_mask=mask, position_ids_all=pos,
                           embeds_all={"vlm": ev.clone(), "proprio": ep.clone(), "action": e.clone()})["action"] for e in ea]
    caches = m.joint_model.build_mixture_caches()
    got = []
    for e in ea:
        got.append(m.joint_model(attention_mask=mask, position_ids_all=pos,
                                 embeds_all={"vlm": ev.clone(), "proprio": ep.clone(), "action": e.clone()}, kv_caches=caches,
                                 cache_mode="no_append")["action"])
    assert all(caches[n].has_item(0) for n in ("vlm", "proprio"))
    for a, b in zip(got, fresh):
        assert max_abs(a, b) < 1e-5
