"""CPU: host-side logic of the drop-in boundary and the C-ABI library surface."""
import ctypes
import os
import re

import pytest
import torch

from helpers import ROOT, SMALL, pz
from oracle import pizero_oracle as O


def test_library_exports_every_declared_symbol():
    from open_pi_zero_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "pz_b200.h")).read()
    declared = set(re.findall(r"\b(pz_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pz_abi_version() == _lib.PZ_ABI_VERSION


def test_create_rejects_bad_config_without_gpu():
    from open_pi_zero_b200 import _lib
    lib = _lib.load()
    cfg = _lib.PzConfig()
    h = ctypes.c_void_p()
    assert lib.pz_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"" != lib.pz_last_error(None)
    assert lib.pz_workspace_bytes(None, 4) == 0


def test_state_dict_contract_938_keys():
    from open_pi_zero_b200.pizero import PiZero
    m = PiZero(pz.cfg_from_dims(pz.make_dims()), init="empty", device="meta")
    sd = m.state_dict()
    assert len(sd) == 938
    assert sd["embed_tokens.weight"].shape == (257216, 2048)
    assert sd["vision_tower.vision_model.embeddings.patch_embedding.weight"].shape == (1152, 3, 14, 14)
    assert sd["joint_model.mixtures.vlm.layers.17.mlp.down_proj.weight"].shape == (2048, 16384)
    assert sd["joint_model.mixtures.action.layers.0.self_attn.q_proj.weight"].shape == (2048, 1024)
    assert sd["joint_model.mixtures.proprio.norm.weight"].shape == (1024,)
    assert "joint_model.mixtures.vlm.norm.weight" not in sd
    assert sd["action_encoder.linear_2.weight"].shape == (1024, 2048)
    assert sd["action_decoder.weight"].shape == (7, 1024)
    n = sum(v.numel() for k, v in sd.items() if ".proprio." not in k)
    assert abs(n - 3.238e9) < 2e6    # 3.238 B unique parameters


def test_load_state_dict_strict_and_orig_mod_prefix():
    from open_pi_zero_b200.pizero import PiZero
    sd = pz.init_state_dict(SMALL, seed=0)
    m = PiZero(SMALL, init="empty")
    m.load_state_dict({"_orig_mod." + k: v for k, v in sd.items()}, strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])
    bad = dict(sd)
    bad.pop("action_decoder.bias")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)
    m.tie_action_proprio_weights()
    assert len(m.state_dict()) == len(sd)     # tied: both prefixes are still emitted


def test_mask_and_position_builder_matches_reference_semantics():
    from open_pi_zero_b200.pizero import PiZero
    m = PiZero(SMALL, init="empty", device="meta")
    inp = pz.make_inputs(SMALL, 7, seed=3, min_text=0)
    for dtype in (torch.float32, torch.bfloat16):
        mask, vp, pp, ap = m.build_causal_mask_and_position_ids(inp["attention_mask"], dtype)
        full, pm, am, pos = O.build_masks_and_positions(SMALL, inp["attention_mask"], dtype)
        assert torch.equal(mask, full)
        assert torch.equal(vp, pos["vlm"]) and torch.equal(pp, pos["proprio"]) and torch.equal(ap, pos["action"])
        a, b = m.split_full_mask_into_submasks(mask)
        assert torch.equal(a, pm) and torch.equal(b, am)
    vl = m._valid_len(pm, inp["input_ids"], None)
    assert torch.equal(vl.to(torch.int64), inp["attention_mask"].sum(1))


def test_config_round_trip_and_scope_guards():
    d = pz.make_dims()
    assert pz.dims_from_cfg(pz.cfg_from_dims(d)) == d
    cfg = pz.cfg_from_dims(d)
    cfg.mixture.vlm.use_lora = True
    with pytest.raises(NotImplementedError):
        pz.dims_from_cfg(cfg)
    cfg = pz.cfg_from_dims(d)
    cfg.action_expert_adaptive_mode = "adaLN"
    with pytest.raises(NotImplementedError):
        pz.dims_from_cfg(cfg)


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA -- never route through the oracle."""
    from open_pi_zero_b200.pizero import PiZero, PzError
    m = PiZero(SMALL, init="empty")
    inp = pz.make_inputs(SMALL, 1)
    with pytest.raises(PzError):
        m.infer_action(inp["input_ids"], inp["pixel_values"], proprios=inp["proprios"])
    import open_pi_zero_b200.pizero as mod
    src = open(mod.__file__).read()
    assert "oracle" not in src.replace("oracle/", "")


def test_synthetic_inputs_follow_processor_layout():
    d = pz.make_dims()
    inp = pz.make_inputs(d, 8, seed=0)
    ids = inp["input_ids"]
    assert ids.shape == (8, 276) and (ids[:, :256] == 257152).all() and (ids[:, 256] == 2).all()
    assert inp["pixel_values"].abs().max() <= 1.0
    vl = inp["valid_len"]
    assert vl.min() >= 257 + 4 and vl.max() <= 276 and len(set(vl.tolist())) > 1
    for b in range(8):
        assert (ids[b, int(vl[b]):] == 0).all() and (ids[b, : int(vl[b])] != 0).all()


def test_packed_weight_staleness_key_is_cheap_and_sees_moves():
    """The per-call staleness key (on the bs=1 latency path) samples the version counters of a few parameters;
    `.to()`, `load_state_dict` and in-place edits of a sampled parameter must all change it."""
    from open_pi_zero_b200.pizero import PiZero
    d = SMALL
    m = PiZero(pz.cfg_from_dims(d), init="empty")
    k0 = m._param_key()
    assert m._param_key() == k0 and "_param_list" in m.__dict__
    m.to(torch.bfloat16)
    assert "_param_list" not in m.__dict__          # dropped by _apply
    k1 = m._param_key()
    assert k1 != k0
    m.load_state_dict(m.state_dict())
    assert "_param_list" not in m.__dict__
    with torch.no_grad():
        next(iter(m.parameters())).add_(1.0)         # the first parameter is always in the sample
    assert m._param_key() != k1


def test_flow_time_sampler_matches_reference_formulas():
    """train.py:217-247: beta sampling t = (1 - sig_min)(1 - Beta(1.5, 1)) and the stratified uniform variant --
    same torch RNG stream, so the draws are bit-equal to the reference's expressions."""
    import torch
    from open_pi_zero_b200.flow import FlowTimeSampler
    s = FlowTimeSampler.from_cfg(pz.cfg_from_dims(pz.make_dims()))
    torch.manual_seed(5)
    t = s.sample_fm_time(4096)
    torch.manual_seed(5)
    want = (1 - 0.001) * (1 - torch.distributions.Beta(1.5, 1).sample((4096,)))
    assert torch.equal(t, want)
    assert 0.0 <= float(t.min()) and float(t.max()) <= 0.999
    assert abs(float(t.mean()) - 0.999 * (1 - 1.5 / 2.5)) < 0.02      # E[Beta(a, b)] = a / (a + b)
    u = FlowTimeSampler("uniform")
    torch.manual_seed(7)
    tu = u.sample_fm_time(8)
    torch.manual_seed(7)
    wu = (torch.rand(1) + torch.arange(8) / 8) % (1 - 1e-5)
    assert torch.equal(tu, wu)
    import pytest
    with pytest.raises(AssertionError):
        FlowTimeSampler("gaussian")


def test_flow_sig_min_travels_through_the_config():
    d = pz.make_dims(flow_sig_min=0.01)
    cfg = pz.cfg_from_dims(d)
    assert cfg.flow_sig_min == 0.01
    assert pz.dims_from_cfg(cfg)["flow_sig_min"] == 0.01
    plain = pz.cfg_from_dims(pz.make_dims())
    del plain["flow_sig_min"]                      # the reference's yaml does not carry it (pizero.py:58 default)
    assert pz.dims_from_cfg(plain)["flow_sig_min"] == 0.001


def test_joint_model_forward_rejects_unsupported_patterns_before_touching_inputs():
    """JointModel.forward scales `embeds_all` in place like the reference (joint_model.py:355) -- but only once it knows
    the call pattern is one the kernels cover; an unsupported one leaves the caller's tensors alone."""
    from open_pi_zero_b200.pizero import PiZero
    d = pz.make_dims(vocab_size=320, image_token_index=300, max_image_text_tokens=10, num_image_tokens=4,
                     num_layers=2, num_heads=4, num_kv_heads=1, head_dim=16, vlm_hidden=64, vlm_inter=128,
                     act_hidden=32, act_inter=64, vit_hidden=32, vit_inter=64, vit_layers=2, vit_heads=2,
                     image_size=28, patch_size=14)
    m = PiZero(pz.cfg_from_dims(d), init="empty")
    x = torch.ones(2, 10, 64)
    keep = x.clone()
    mask = torch.zeros(2, 1, 10, 10)
    with pytest.raises(NotImplementedError):
        m.joint_model.forward(mask, {"vlm": None}, {"vlm": x}, cache_mode="append")
    assert torch.equal(x, keep)
    a = torch.ones(2, 4, 32)
    with pytest.raises(NotImplementedError, match="adaLN"):
        m.joint_model.forward(mask, {"action": None}, {"action": a}, time_cond=torch.zeros(2, 32),
                              cache_mode="append_non_active")
    assert torch.equal(a, torch.ones(2, 4, 32))


def test_adapter_formulas_match_reference_base_adapter():
    """adapter.py restates env_adapter/base.py:8-49; the affine forms handed to the kernels reproduce SimplerAdapter's
    preprocess / postprocess arithmetic (simpler.py:76-90, 102-125)."""
    import numpy as np
    from open_pi_zero_b200.adapter import BaseEnvAdapter, action_affine, proprio_affine
    rng = np.random.default_rng(0)
    stats = {k: dict(p01=rng.normal(size=7) - 2, p99=rng.normal(size=7) + 2, mean=rng.normal(size=7), std=rng.uniform(0.5, 2, 7))
             for k in ("proprio", "action")}
    ours = BaseEnvAdapter()
    x = rng.normal(size=(5, 7)) * 3
    ref_path = "/root/reference/src/agent/env_adapter/base.py"
    import os
    if os.path.exists(ref_path):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_base_adapter", ref_path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ref = mod.BaseEnvAdapter()
        for name, args in (("normalize_bound", (stats["proprio"]["p01"], stats["proprio"]["p99"])),
                           ("denormalize_bound", (stats["action"]["p01"], stats["action"]["p99"])),
                           ("normalize_gaussian", (stats["proprio"]["mean"], stats["proprio"]["std"])),
                           ("denormalize_gaussian", (stats["action"]["mean"], stats["action"]["std"]))):
            assert np.array_equal(getattr(ours, name)(x, *args), getattr(ref, name)(x, *args)), name
    for kind in ("bound", "gaussian"):
        s, b, clip = proprio_affine(stats["proprio"], kind)
        got = x * s + b
        if clip:
            got = np.clip(got, -1, 1)
        want = (ours.normalize_bound(x, stats["proprio"]["p01"], stats["proprio"]["p99"]) if kind == "bound"
                else ours.normalize_gaussian(x, stats["proprio"]["mean"], stats["proprio"]["std"]))
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
        s, b = action_affine(stats["action"], kind)
        a = np.clip(x, -1, 1)
        got = a * s + b
        want = (ours.denormalize_bound(a[:, :-1], stats["action"]["p01"][:-1], stats["action"]["p99"][:-1]) if kind == "bound"
                else ours.denormalize_gaussian(a[:, :-1], stats["action"]["mean"][:-1], stats["action"]["std"][:-1]))
        assert np.allclose(got[:, :-1], want, rtol=1e-12, atol=1e-12)
        assert np.array_equal(got[:, -1], a[:, -1])          # the gripper dimension is passed through


def test_cosine_warmup_schedule_matches_reference():
    """train.CosineAnnealingWarmupRestarts against src/utils/optim.py:31-160 (the schedule both optimizers of the reference's
    training loop step once per update, train.py:376-379), including restarts with cycle_mult / gamma."""
    import importlib.util
    import os
    from types import SimpleNamespace
    from open_pi_zero_b200.train import CosineAnnealingWarmupRestarts as Ours
    ref_path = "/root/reference/src/utils/optim.py"
    if not os.path.exists(ref_path):
        pytest.skip("/root/reference not present")
    spec = importlib.util.spec_from_file_location("ref_optim", ref_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for kw in (dict(first_cycle_steps=50, max_lr=5e-5, min_lr=1e-8, warmup_steps=10),
               dict(first_cycle_steps=20, cycle_mult=1.5, max_lr=1e-3, min_lr=1e-5, warmup_steps=3, gamma=0.7)):
        ref_opt = SimpleNamespace(param_groups=[{"lr": 1.0}])
        ref = mod.CosineAnnealingWarmupRestarts(ref_opt, **kw)
        mine_opt = SimpleNamespace(groups={"action": {"lr": 1.0}})
        mine = Ours(mine_opt, "action", **kw)
        assert mine_opt.groups["action"]["lr"] == ref_opt.param_groups[0]["lr"]
        for _ in range(130):
            ref.step()
            mine.step()
            assert abs(mine_opt.groups["action"]["lr"] - ref_opt.param_groups[0]["lr"]) <= 1e-18 + 1e-12 * ref_opt.param_groups[0]["lr"]
        assert mine.state_dict()["cycle"] == ref.state_dict()["cycle"]


class _StubTokenizer:
    """A HuggingFace-shaped tokenizer (only what VLAProcessor touches): special / added tokens get ids, everything else is
    tokenised per character; right padding to max_length with id 0."""
    bos_token = "<bos>"

    def __init__(self):
        self.vocab = {"<pad>": 0, "<bos>": 2}
        self.special = []
        self.add_bos_token = self.add_eos_token = True
        self.calls = []

    def _add(self, toks):
        for t in toks:
            if t not in self.vocab:
                self.vocab[t] = 1000 + len(self.vocab)
                self.special.append(t)

    def add_special_tokens(self, d):
        self._add(d["additional_special_tokens"])

    def add_tokens(self, toks):
        self._add(toks)

    def convert_tokens_to_ids(self, t):
        return self.vocab[t]

    def __call__(self, strings, return_tensors, max_length, padding, truncation):
        import torch
        self.calls.append(dict(return_tensors=return_tensors, max_length=max_length, padding=padding, truncation=truncation))
        specials = sorted(self.special + ["<bos>"], key=len, reverse=True)
        ids = []
        for s in strings:
            row, i = [], 0
            while i < len(s):
                for sp in specials:
                    if s.startswith(sp, i):
                        row.append(self.vocab[sp]); i += len(sp)
                        break
                else:
                    row.append(3 + ord(s[i]) % 200); i += 1
            row = row[:max_length] if truncation else row
            ids.append(row)
        n = max_length if padding == "max_length" else max(len(r) for r in ids)
        mask = [[1] * len(r) + [0] * (n - len(r)) for r in ids]
        ids = [r + [0] * (n - len(r)) for r in ids]
        return {"input_ids": torch.tensor(ids), "attention_mask": torch.tensor(mask)}


def test_vla_processor_matches_reference():
    """processing.VLAProcessor against src/model/vla/processing.py:63-136 with the same stub tokenizer: identical prompt strings /
    ids / masks / pixel values and identical tokenizer arguments; keep_uint8 hands the frames through untouched."""
    import importlib.util
    import os
    import torch
    from open_pi_zero_b200.processing import VLAProcessor as Ours
    ref_path = "/root/reference/src/model/vla/processing.py"
    if not os.path.exists(ref_path):
        pytest.skip("/root/reference not present")
    spec = importlib.util.spec_from_file_location("ref_processing", ref_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (2, 3, 28, 28), generator=g, dtype=torch.uint8)
    text = ["put the carrot on the plate", "open the drawer"]
    for padding in ("max_length", "longest"):
        t_ref, t_ours = _StubTokenizer(), _StubTokenizer()
        ref = mod.VLAProcessor(t_ref, num_image_tokens=4, max_seq_len=48, tokenizer_padding=padding)
        ours = Ours(t_ours, num_image_tokens=4, max_seq_len=48, tokenizer_padding=padding)
        a, b = ref(text=text, images=images), ours(text=text, images=images)
        assert set(a) == set(b) and ours.image_token_id == ref.image_token_id
        for k in a:
            assert torch.equal(a[k], b[k]), k
        assert t_ref.calls == t_ours.calls and t_ours.add_bos_token is False and t_ours.add_eos_token is False
        assert int((b["input_ids"][0] == ours.image_token_id).sum()) == 4 and bool((b["input_ids"][:, :4] == ours.image_token_id).all())
    raw = Ours(_StubTokenizer(), num_image_tokens=4, max_seq_len=48, keep_uint8=True)(text=text, images=images)
    assert raw["pixel_values"].dtype == torch.uint8 and torch.equal(raw["pixel_values"], images)


def test_action_accuracy_matches_reference_metric():
    import importlib.util
    import os
    import torch
    from open_pi_zero_b200.metric import eval_stats, get_action_accuracy
    ref_path = "/root/reference/src/utils/metric.py"
    g = torch.Generator().manual_seed(0)
    gt = torch.rand((6, 4, 7), generator=g) * 2 - 1
    pred = gt + 0.12 * torch.randn((6, 4, 7), generator=g)
    ours = get_action_accuracy(gt, pred, [0.1, 0.2, 0.5])
    if os.path.exists(ref_path):
        spec = importlib.util.spec_from_file_location("ref_metric", ref_path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert torch.equal(ours, mod.get_action_accuracy(gt, pred, [0.1, 0.2, 0.5]))
    assert ours[0] <= ours[1] <= ours[2] and 0.0 < float(ours[2]) <= 1.0
    acc, l1 = eval_stats([pred, pred], [gt, gt], [0.1, 0.2, 0.5])
    assert torch.allclose(acc, ours) and abs(float(l1) - float((pred - gt).abs().mean())) < 1e-7


def test_preprocess_batch_mirrors_the_training_loop_closure():
    """train.preprocess_batch (train.py:271-313): same keys / shapes / values as composing the processor, the mask builder and
    the time sampler by hand; `dense_masks=False` returns valid_len = the number of image + text tokens."""
    import torch
    from open_pi_zero_b200.pizero import PiZero
    from open_pi_zero_b200.processing import VLAProcessor
    from open_pi_zero_b200.train import preprocess_batch
    # a parameter-free stand-in for the model: only the mask builders are used
    class M:
        max_image_text_tokens, num_proprio_tokens, num_action_tokens, total_num_tokens = 48, 1, 4, 53
        build_causal_mask_and_position_ids = PiZero.build_causal_mask_and_position_ids
        split_full_mask_into_submasks = PiZero.split_full_mask_into_submasks
    m = M()
    proc = VLAProcessor(_StubTokenizer(), num_image_tokens=4, max_seq_len=48)
    g = torch.Generator().manual_seed(0)
    B = 3
    batch = {"observation": {"image_primary": torch.randint(0, 256, (B, 1, 28, 28, 3), generator=g, dtype=torch.uint8),
                             "proprio": torch.rand((B, 1, 7), generator=g)},
             "action": torch.rand((B, 1, 4, 7), generator=g),
             "task": {"language_instruction": [b"pick up the spoon", b"open the drawer", b"move left"]}}
    out = preprocess_batch(m, proc, batch, torch.float32, "cpu", split_mask=False, sample_fm_time=True)
    assert set(out) == {"input_ids", "pixel_values", "vlm_position_ids", "proprio_position_ids", "action_position_ids", "proprios",
                        "actions", "causal_mask", "t"}
    assert out["pixel_values"].shape == (B, 3, 28, 28) and out["actions"].shape == (B, 4, 7) and out["t"].shape == (B,)
    assert out["causal_mask"].shape == (B, 1, 53, 53)
    want = proc(text=["pick up the spoon", "open the drawer", "move left"], images=batch["observation"]["image_primary"][:, 0].permute(0, 3, 1, 2))
    assert torch.equal(out["input_ids"], want["input_ids"]) and torch.equal(out["pixel_values"], want["pixel_values"])
    inf = preprocess_batch(m, proc, batch, torch.float32, "cpu", split_mask=True, sample_fm_time=False)
    assert "image_text_proprio_mask" in inf and inf["image_text_proprio_mask"].shape == (B, 1, 49, 49) and inf["action_mask"].shape == (B, 1, 4, 53)
    lean = preprocess_batch(m, VLAProcessor(_StubTokenizer(), 4, 48, keep_uint8=True), batch, torch.float32, "cpu", split_mask=True,
                            sample_fm_time=False, dense_masks=False)
    assert lean["pixel_values"].dtype == torch.uint8 and "action_mask" not in lean
    assert torch.equal(lean["valid_len"], want["attention_mask"].sum(1).to(torch.int32))
    # the dense mask's row 0 encodes exactly that count
    assert torch.equal((out["causal_mask"][:, 0, 0, :48] == 0).sum(-1).to(torch.int32), lean["valid_len"])


def test_c_abi_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/pz_b200.h is the drop-in boundary: it must compile as C99 and as C++17, and a C program that references every
    declared entry point must link against libpz_b200.so (no compute call: no GPU here)."""
    import re
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    from open_pi_zero_b200 import _lib
    header = os.path.join(ROOT, "include", "pz_b200.h")
    for cmd in (["gcc", "-std=c99", "-fsyntax-only", "-x", "c", header], ["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", header]):
        assert subprocess.run(cmd, capture_output=True).returncode == 0, cmd
    lib = os.path.join(ROOT, "open-pi-zero_b200", "libpz_b200.so")
    if not os.path.exists(lib):
        pytest.skip("library not built")
    names = sorted(set(re.findall(r"\b(pz_[a-z0-9_]+)\s*\(", open(header).read())))
    src = tmp_path / "link.c"
    src.write_text('#include "pz_b200.h"\n#include <stdio.h>\nint main(void) {\n  const void *p[] = {' +
                   ", ".join(f"(const void *)&{n}" for n in names) +
                   '};\n  printf("%d %d\\n", (int)(sizeof(p) / sizeof(p[0])), pz_abi_version());\n  return 0;\n}\n')
    exe = tmp_path / "link"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), lib,
                        "-Wl,-rpath," + os.path.dirname(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    n, ver = out.stdout.split()
    assert int(n) == len(names) == len(_lib.EXPORTS) and int(ver) == _lib.PZ_ABI_VERSION
