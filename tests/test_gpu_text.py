"""GPU parity of the text output (`PiZero.infer_text`, reference pizero.py:559-593, `cache_mode="append"`
joint_model.py:164-240) through the C ABI (`pz_embed_prefix` -> `pz_text_prefill`, then `pz_text_decode` per token) against
logits the UNMODIFIED reference produced (tests/golden/text_small.pt, oracle/make_golden_text.py) and the CPU oracle.
Tolerances: fp32 logits <= 2e-4 relative; bf16 logits <= 2e-2 relative (Frobenius), K/V rows <= 2e-2 relative."""
import os

import pytest
import torch

from helpers import max_abs, pz, rel_err
from oracle import pizero_oracle as O

pytestmark = pytest.mark.gpu


def _fixture(golden_dir):
    path = os.path.join(golden_dir, "text_small.pt")
    if not os.path.exists(path):
        pytest.skip("text_small.pt missing")
    return torch.load(path, weights_only=False)


def _model(d, sd, dtype):
    from open_pi_zero_b200.pizero import PiZeroInference
    m = PiZeroInference(pz.cfg_from_dims(d), init="empty")
    m.load_state_dict(sd, strict=True)
    return m.to(dtype).to("cuda")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 2e-2)])
def test_infer_text_prefill_and_decode_vs_reference_golden(golden_dir, dtype, tol):
    from open_pi_zero_b200.pizero import TextKVCache
    fx = _fixture(golden_dir)
    d = fx["dims"]
    sd = pz.init_state_dict(d, seed=fx["seed"], randomize_norms=fx["randomize_norms"])
    m = _model(d, sd, dtype)
    ids, pix = fx["input_ids"].cuda(), fx["pixel_values"].cuda().to(dtype)
    mask = torch.ones_like(ids)
    cache = TextKVCache()
    out = m.infer_text(ids, pix, mask, cache)
    assert out["kv_cache"] is cache and cache.num_items() == fx["q_len"]
    assert out["logits"].shape == fx["logits"][0].shape
    worst = rel_err(out["logits"], fx["logits"][0])
    # teacher-forced with the reference's own greedy tokens, so that every step is compared on the same history
    for i, tok in enumerate(fx["tokens"]):
        if dtype == torch.float32:
            assert torch.equal(out["logits"][:, -1].argmax(-1, keepdim=True).cpu(), tok)
        mask = torch.cat([mask, torch.ones_like(mask[:, :1])], 1)
        out = m.infer_text(tok.cuda(), pix, mask, cache)
        assert out["logits"].shape == fx["logits"][i + 1].shape
        worst = max(worst, rel_err(out["logits"], fx["logits"][i + 1]))
    assert cache.num_items() == fx["q_len"] + len(fx["tokens"])
    for l, (k2, v2) in enumerate(fx["kv"]):
        k, v = cache.get(l)
        assert k.shape == k2.shape
        worst = max(worst, rel_err(k.float(), k2), rel_err(v.float(), v2))
    print(f"[text_small {dtype}] worst rel err (logits of prefill + {len(fx['tokens'])} decode steps, K/V) {worst:.3e}")
    assert worst < tol


def test_infer_text_last_token_only_and_generate_match_full_prefill(golden_dir):
    fx = _fixture(golden_dir)
    d = fx["dims"]
    sd = pz.init_state_dict(d, seed=fx["seed"], randomize_norms=fx["randomize_norms"])
    m = _model(d, sd, torch.float32)
    ids, pix = fx["input_ids"].cuda(), fx["pixel_values"].cuda()
    mask = torch.ones_like(ids)
    full = m.infer_text(ids, pix, mask)["logits"]
    last = m.infer_text(ids, pix, mask, last_token_only=True)["logits"]
    assert last.shape == (ids.shape[0], 1, d["vocab_size"])
    assert max_abs(full[:, -1:], last) < 1e-4
    gen = m.generate_text(ids, pix, mask, max_new_tokens=len(fx["tokens"]))
    assert torch.equal(gen.cpu(), torch.cat(fx["tokens"], 1))


def test_infer_text_short_prompt_and_errors(golden_dir):
    """A prompt shorter than the fixture's (rows past q_len of the merged buffer are pad tokens that nothing attends to)
    against the oracle; the error behaviour of the reference's asserts."""
    from open_pi_zero_b200.pizero import PzError, TextKVCache
    fx = _fixture(golden_dir)
    d = fx["dims"]
    sd = pz.init_state_dict(d, seed=fx["seed"], randomize_norms=fx["randomize_norms"])
    m = _model(d, sd, torch.float32)
    q = d["num_image_tokens"] + 2
    ids, pix = fx["input_ids"][:2, :q], fx["pixel_values"][:2]
    want = O.infer_text(sd, d, ids, pix, torch.ones_like(ids))["logits"]
    got = m.infer_text(ids.cuda(), pix.cuda(), torch.ones_like(ids).cuda())["logits"]
    assert rel_err(got, want) < 2e-4
    cache = TextKVCache()
    m.infer_text(ids.cuda(), pix.cuda(), torch.ones_like(ids).cuda(), cache)
    with pytest.raises(ValueError):     # "Using KV cache so should only use one single token" (pizero.py:352)
        m.infer_text(ids.cuda(), pix.cuda(), torch.ones_like(ids).cuda(), cache)
    with pytest.raises(ValueError):     # padded prompts are outside infer_text's contract (pizero.py:346-357)
        bad = torch.ones_like(ids); bad[0, -1] = 0
        m.infer_text(ids.cuda(), pix.cuda(), bad.cuda())
    # a model without the text head refuses
    d2 = dict(d, use_lm_head=False, vlm_use_final_norm=False)
    m2 = _model(d2, pz.init_state_dict(d2, seed=1), torch.float32)
    with pytest.raises(PzError):
        m2.infer_text(ids.cuda(), pix.cuda(), torch.ones_like(ids).cuda())


def test_joint_model_forward_append_mode_matches_infer_text(golden_dir):
    """`JointModel.forward(embeds_all={"vlm": ...}, cache_mode="append", final_layer_post_attn_skip_names=[])` -- the call
    infer_text makes (pizero.py:571-583): final-norm hidden states whose lm_head product is the reference's logits, for the
    prompt and for appended single tokens."""
    from open_pi_zero_b200.pizero import TextKVCache
    fx = _fixture(golden_dir)
    d = fx["dims"]
    sd = pz.init_state_dict(d, seed=fx["seed"], randomize_norms=fx["randomize_norms"])
    m = _model(d, sd, torch.float32)
    ids, pix = fx["input_ids"].cuda(), fx["pixel_values"].cuda()
    B, q_len = ids.shape
    # embeddings of the prompt exactly as infer_text builds them: through the public infer_text once (prefix embeds are internal),
    # here rebuilt from the oracle's embed_prefix
    emb = O.embed_prefix(sd, d, fx["input_ids"], fx["pixel_values"]).cuda()
    mask = torch.zeros((B, 1, q_len, q_len), device="cuda")
    caches = {"vlm": TextKVCache()}
    out = m.joint_model(attention_mask=mask, position_ids_all={"vlm": torch.arange(1, q_len + 1, device="cuda").repeat(B, 1)},
                        embeds_all={"vlm": emb.clone()}, kv_caches=caches, cache_mode="append",
                        final_layer_post_attn_skip_names=[])["vlm"]
    W = sd["embed_tokens.weight"].cuda()
    assert rel_err(out @ W.t(), fx["logits"][0]) < 2e-4
    assert caches["vlm"].num_items() == q_len
    tok = fx["tokens"][0].cuda()
    e1 = W[tok[:, 0]][:, None].clone()
    out1 = m.joint_model(attention_mask=torch.zeros((B, 1, 1, q_len + 1), device="cuda"),
                         position_ids_all={"vlm": torch.full((B, 1), q_len + 1, device="cuda")}, embeds_all={"vlm": e1},
                         kv_caches=caches, cache_mode="append", final_layer_post_attn_skip_names=[])["vlm"]
    assert out1.shape == (B, 1, d["vlm_hidden"])
    assert rel_err(out1 @ W.t(), fx["logits"][1]) < 2e-4
    assert caches["vlm"].num_items() == q_len + 1
