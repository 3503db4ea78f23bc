"""GPU parity of the flow-matching TRAINING STEP (`PiZero.forward` + `loss.backward()`, reference pizero.py:607-661 /
train.py:350-368) through the C ABI (`pz_flow_matching_step`): loss and d loss / d parameter for every trainable tensor.

Checkers: (1) `oracle.pizero_backward.flow_matching_backward_full` (the hand-written backward, itself pinned to the
unmodified reference's autograd in tests/test_flow_matching.py) on the kernels' small shape; (2) the gradient norms and
leading rows the UNMODIFIED reference's `loss.backward()` produced at the real widths (tests/golden/fm_width2.pt,
oracle/make_golden_fm.py).  Tolerances, per tensor, relative Frobenius error: fp32 <= 2e-3; bf16 <= 6e-2 (bf16 operands
in every product of a 2 x (27 + 18)-layer backward chain; the loss itself <= 2e-2)."""
import os

import pytest
import torch

from helpers import SMALL, max_abs, pz, rel_err
from oracle import pizero_backward as Bk

pytestmark = pytest.mark.gpu


def _model(d, sd, dtype):
    from open_pi_zero_b200.pizero import PiZero
    m = PiZero(pz.cfg_from_dims(d), init="empty")
    m.load_state_dict(sd, strict=True)
    return m.to(dtype).to("cuda")


def _targets(d, B, seed):
    g = torch.Generator().manual_seed(seed)
    actions = torch.rand((B, d["horizon_steps"], d["action_dim"]), generator=g) * 2 - 1
    noise = torch.randn((B, d["horizon_steps"], d["action_dim"]), generator=g)
    t = torch.rand((B,), generator=g)
    return actions, noise, t


def _step(m, inp, actions, noise, t, grads, **kw):
    from open_pi_zero_b200.train import flow_matching_step
    dt = next(m.parameters()).dtype
    loss = flow_matching_step(m, inp["input_ids"].cuda(), inp["pixel_values"].cuda().to(dt), inp["proprios"].cuda(),
                              actions.cuda(), t.cuda(), noise=noise.cuda(), valid_len=inp["valid_len"].cuda(), grads=grads, **kw)
    torch.cuda.synchronize()
    return loss


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 6e-2)])
def test_small_every_gradient_vs_oracle(dtype, tol):
    from open_pi_zero_b200.train import GradBuffer
    d = SMALL
    B = 3
    sd = pz.init_state_dict(d, seed=13, randomize_norms=True, tie_proprio=False)
    inp = pz.make_inputs(d, B, seed=31, min_text=0)
    actions, noise, t = _targets(d, B, 5)
    want_loss, want = Bk.flow_matching_backward_full(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                                                     inp["proprios"], actions, t, noise)
    m = _model(d, sd, dtype)
    gb = GradBuffer(m)
    assert not gb.tied
    loss = _step(m, inp, actions, noise, t, gb)
    assert abs(float(loss) - float(want_loss)) < (1e-4 if dtype == torch.float32 else 2e-2) * max(1.0, float(want_loss))
    got = gb.unpack()
    worst, worst_key, checked, errs = 0.0, None, 0, {}
    for k, g in want.items():
        if k == "embed_tokens.weight":      # frozen in the reference's training (pizero.py:243-249); not produced
            continue
        assert k in got, k
        assert got[k].shape == g.shape, (k, got[k].shape, g.shape)
        if float(g.abs().max()) == 0.0:     # the discarded last-layer vlm / proprio half
            assert float(got[k].abs().max()) == 0.0, k
            continue
        if k.endswith("self_attn.k_proj.bias"):
            # softmax is invariant to a constant added to every key: this gradient is zero in exact arithmetic and what is
            # left is rounding noise -- compare it with the scale of the q bias gradient instead of with itself
            scale = float(want[k.replace("k_proj", "q_proj")].norm())
            assert float(g.norm()) < 1e-3 * scale and float(got[k].norm()) < (1e-3 if dtype == torch.float32 else 5e-2) * scale, k
            continue
        e = rel_err(got[k], g)
        errs[k] = e
        if e > worst:
            worst, worst_key = e, k
        checked += 1
    print(f"[train step small {dtype}] loss {float(loss):.6f} vs {float(want_loss):.6f}; {checked} gradient tensors, worst "
          f"relative error {worst:.3e} ({worst_key}); launches {m.last_launch_count}")
    bad = {k: round(e, 5) for k, e in errs.items() if not e < tol}
    assert checked >= 90
    assert not bad, bad


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 6e-2)])
def test_width2_gradients_vs_reference_autograd(golden_dir, dtype, tol):
    """Real widths (2048 / 16384, 1024 / 4096, 1152 / 4304, head_dim 256; 2 + 2 layers): gradient norms and leading rows
    of the UNMODIFIED reference's loss.backward()."""
    from open_pi_zero_b200.train import GradBuffer
    path = os.path.join(golden_dir, "fm_width2.pt")
    if not os.path.exists(path):
        pytest.skip("fm_width2.pt missing")
    fx = torch.load(path, weights_only=False)
    if "grad_norms" not in fx["ref"]:
        pytest.skip("fixture predates the gradient capture")
    d = fx["dims"]
    sd = pz.init_state_dict(d, seed=fx["seed"], randomize_norms=fx["randomize_norms"], tie_proprio=fx["tie_proprio"])
    inp = pz.make_inputs(d, fx["batch"], seed=fx["inputs_seed"])
    m = _model(d, sd, dtype)
    del sd
    gb = GradBuffer(m)
    loss = _step(m, inp, fx["actions"], fx["noise"], fx["t"], gb)
    ref = fx["ref"]
    assert abs(float(loss) - float(ref["loss"])) < (1e-4 if dtype == torch.float32 else 2e-2) * max(1.0, float(ref["loss"]))
    got = gb.unpack()
    worst_n, worst_h, checked = 0.0, 0.0, 0
    for k, want_n in ref["grad_norms"].items():
        want_n = float(want_n)
        if k not in got:   # frozen embedding (pizero.py:243-249) or a parameter the loss does not depend on (proprio final norm)
            assert k in ("embed_tokens.weight", "lm_head.weight") or want_n == 0.0, k
            continue
        g = got[k]
        if want_n == 0.0:
            if ".17." in k or f".{d['num_layers'] - 1}." in k:
                assert float(g.abs().max()) == 0.0, k
            continue
        if k.endswith("self_attn.k_proj.bias"):   # zero in exact arithmetic (see above): rounding noise on both sides
            scale = float(ref["grad_norms"][k.replace("k_proj", "q_proj")])
            assert want_n < 1e-3 * scale and float(g.norm()) < (1e-3 if dtype == torch.float32 else 5e-2) * scale, k
            continue
        en = abs(float(g.double().norm()) - want_n) / want_n
        head = g.reshape(g.shape[0], -1)[:4, :64] if g.dim() > 1 else g[:64]
        want_h = ref["grad_heads"][k]
        eh = max_abs(head, want_h) / max(float(want_h.abs().max()), want_n / g.numel() ** 0.5)
        worst_n, worst_h = max(worst_n, en), max(worst_h, eh)
        assert en < tol and eh < 4 * tol, (k, en, eh)   # (NaN fails both)
        checked += 1
    print(f"[train step width2 {dtype}] loss {float(loss):.6f} vs reference {float(ref['loss']):.6f}; {checked} tensors; worst norm "
          f"error {worst_n:.3e}, worst leading-rows error {worst_h:.3e}")
    assert checked >= 80


def test_accumulation_loss_scale_and_frozen_vision():
    """Two calls at loss_scale 0.5 accumulate to one call at 1 (micro-batches under no_sync, train.py:350-356);
    `freeze_vision` leaves the SigLIP / projector gradients untouched and everything else unchanged."""
    from open_pi_zero_b200.train import GradBuffer
    d = SMALL
    B = 2
    sd = pz.init_state_dict(d, seed=3, randomize_norms=True, tie_proprio=True)
    inp = pz.make_inputs(d, B, seed=7)
    actions, noise, t = _targets(d, B, 9)
    m = _model(d, sd, torch.float32)
    m.tie_action_proprio_weights()
    a, b, c = GradBuffer(m), GradBuffer(m), GradBuffer(m)
    assert a.tied
    l1 = _step(m, inp, actions, noise, t, a)
    _step(m, inp, actions, noise, t, b, loss_scale=0.5)
    _step(m, inp, actions, noise, t, b, loss_scale=0.5)
    assert rel_err(b.flat, a.flat) < 1e-4      # fp32 atomics: not bit-equal
    l3 = _step(m, inp, actions, noise, t, c, freeze_vision=True)
    assert float(l1) == float(l3)
    ga, gc = a.unpack(), c.unpack()
    for k in ga:
        if k.startswith("vision_tower.") or k.startswith("multi_modal_projector."):
            assert float(gc[k].abs().max()) == 0.0, k
        else:
            assert rel_err(gc[k], ga[k]) < 1e-4 or float(ga[k].abs().max()) == 0.0, k
    # loss only
    l4 = _step(m, inp, actions, noise, t, None)
    assert float(l4) == float(l1)


def test_fused_adamw_matches_torch_adamw_and_updates_the_packed_weights():
    """Two optimizer steps (global-norm clip + AdamW, two parameter groups) against torch.optim.AdamW +
    clip_grad_norm_ applied to the unpacked gradients in the reference's layout (train.py:371-379); the loss of the next
    forward must see the new weights without a re-pack; sync_parameters() writes them back under the reference's names."""
    from open_pi_zero_b200.train import FusedAdamW, GradBuffer
    d = SMALL
    B = 2
    sd = pz.init_state_dict(d, seed=21, randomize_norms=True, tie_proprio=False)
    inp = pz.make_inputs(d, B, seed=2)
    actions, noise, t = _targets(d, B, 4)
    m = _model(d, sd, torch.float32)
    gb = GradBuffer(m)
    opt = FusedAdamW(gb, action_lr=3e-3, vlm_lr=1e-3, action_weight_decay=0.1, vlm_weight_decay=0.0, max_grad_norm=0.5)
    ref_params = {k: v.detach().clone().float().cuda().requires_grad_(True) for k, v in sd.items()}
    act_keys = [k for k in ref_params if k.startswith(("action_encoder.", "action_decoder.", "proprio_encoder.",
                                                       "joint_model.mixtures.action.", "joint_model.mixtures.proprio."))]
    vlm_keys = [k for k in ref_params if k not in act_keys and k != "embed_tokens.weight"]
    topt = torch.optim.AdamW([dict(params=[ref_params[k] for k in vlm_keys], lr=1e-3, weight_decay=0.0),
                              dict(params=[ref_params[k] for k in act_keys], lr=3e-3, weight_decay=0.1)], betas=(0.9, 0.999), eps=1e-8)
    losses = []
    for it in range(2):
        losses.append(float(_step(m, inp, actions, noise, t, gb)))
        got = gb.unpack()
        for k, p in ref_params.items():
            p.grad = got[k].detach().clone() if k in got else None
        trained = [ref_params[k] for k in vlm_keys + act_keys if ref_params[k].grad is not None]
        tn = torch.nn.utils.clip_grad_norm_(trained, max_norm=0.5)
        topt.step()
        opt.step()
        torch.cuda.synchronize()
        assert abs(float(opt.grad_norm()) - float(tn)) < 1e-4 * float(tn)
        assert float(gb.flat.abs().max()) == 0.0          # zero_grad fused into the update
    l3 = float(_step(m, inp, actions, noise, t, None))
    assert l3 < losses[0]                                    # the kernels read the updated packed weights
    opt.sync_parameters()
    now = dict(m.named_parameters())
    worst = 0.0
    for k in vlm_keys + act_keys:
        if ref_params[k].grad is None:
            continue
        worst = max(worst, max_abs(now[k], ref_params[k]) / max(1e-6, float(ref_params[k].abs().max())))
    print(f"[fused AdamW] losses {losses} -> {l3:.6f}; worst relative parameter difference vs torch.optim.AdamW {worst:.3e}")
    assert worst < 1e-5
    # after the sync the packed weights are not rebuilt and the loss is unchanged
    packed = m._packed
    assert abs(float(_step(m, inp, actions, noise, t, None)) - l3) < 1e-6 and m._packed is packed


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_training_loop_overfits_a_fixed_batch(dtype):
    """End to end: 40 optimizer steps (forward + backward + clip + AdamW through the C ABI) on one fixed batch must drive the
    flow-matching loss down -- the gradients point the right way and the optimizer writes the weights the kernels read."""
    from open_pi_zero_b200.train import FusedAdamW, GradBuffer
    d = SMALL
    B = 4
    sd = pz.init_state_dict(d, seed=17, randomize_norms=False, tie_proprio=True)
    inp = pz.make_inputs(d, B, seed=3)
    actions, noise, t = _targets(d, B, 11)
    m = _model(d, sd, dtype)
    m.tie_action_proprio_weights()
    gb = GradBuffer(m)
    opt = FusedAdamW(gb, action_lr=2e-3, vlm_lr=5e-4, max_grad_norm=1.0)
    losses = []
    for _ in range(40):
        losses.append(float(_step(m, inp, actions, noise, t, gb)))
        opt.step()
    final = float(_step(m, inp, actions, noise, t, None))
    print(f"[overfit {dtype}] loss {losses[0]:.4f} -> {final:.4f} (min over the run {min(losses):.4f})")
    assert final < 0.35 * losses[0], (losses[0], final)
    assert all(l == l for l in losses)      # no NaN on the way


def test_optimizer_checkpoint_resume_and_weight_averaging():
    """FusedAdamW.state_dict / load_state_dict: a run resumed from the checkpoint continues bit for bit; ModelAveraging (EMA)
    against the torch.optim.swa_utils rule the reference uses (model_averaging.py:40-66), and the averaged weights are what
    the kernels compute with inside `averaged_weights()`."""
    from open_pi_zero_b200.train import FusedAdamW, GradBuffer, ModelAveraging
    d = SMALL
    B = 2
    sd = pz.init_state_dict(d, seed=23, randomize_norms=True, tie_proprio=True)
    inp = pz.make_inputs(d, B, seed=5)
    actions, noise, t = _targets(d, B, 6)

    def fresh():
        m = _model(d, sd, torch.float32)
        m.tie_action_proprio_weights()
        gb = GradBuffer(m)
        return m, gb, FusedAdamW(gb, action_lr=1e-3, vlm_lr=1e-3)

    m, gb, opt = fresh()
    avg = ModelAveraging(opt, use_ema=True, ema_start=1, ema_decay=0.9)
    ema_ref = None
    for it in range(1, 4):
        _step(m, inp, actions, noise, t, gb)
        opt.step()
        avg.maybe_initialize(it)
        avg.maybe_update(it)
        cur = opt.master.clone()
        ema_ref = cur if ema_ref is None else 0.9 * ema_ref + 0.1 * cur      # get_ema_multi_avg_fn: the first update copies
    assert rel_err(avg.avg, ema_ref) < 1e-6
    ckpt = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state_dict().items()}
    l_cont = []
    for _ in range(2):
        l_cont.append(float(_step(m, inp, actions, noise, t, gb)))
        opt.step()
    # the averaged weights are a different function than the trained ones, and the trained ones come back afterwards
    l_train = float(_step(m, inp, actions, noise, t, None))
    with avg.averaged_weights() as me:
        l_avg = float(_step(me, inp, actions, noise, t, None))
    assert l_avg != l_train and float(_step(m, inp, actions, noise, t, None)) == l_train
    # resume in a fresh process-equivalent: new model from the ORIGINAL state dict + optimizer state from the checkpoint
    m2, gb2, opt2 = fresh()
    opt2.load_state_dict(ckpt)
    l_res = []
    for _ in range(2):
        l_res.append(float(_step(m2, inp, actions, noise, t, gb2)))
        opt2.step()
    assert max(abs(a - b) for a, b in zip(l_cont, l_res)) < 1e-5 * max(l_cont), (l_cont, l_res)
    assert rel_err(opt2.master, opt.master) < 1e-6


def test_multi_image_chunk10_joint_gradients_vs_oracle():
    """Two images per observation and a 10-step action chunk (the geometry family of BASELINE configs[3]): loss and the
    gradients of everything downstream of the embeddings (joint model, action encoder / decoder) against
    `oracle.pizero_backward.flow_matching_backward` (its SigLIP part is single-image)."""
    from open_pi_zero_b200.train import GradBuffer
    d = pz.make_dims(SMALL, num_images=2, max_image_text_tokens=40, horizon_steps=10)
    B = 2
    sd = pz.init_state_dict(d, seed=29, randomize_norms=True, tie_proprio=False)
    inp = pz.make_inputs(d, B, seed=41, min_text=2)
    actions, noise, t = _targets(d, B, 15)
    want_loss, want, _ = Bk.flow_matching_backward(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                                                   inp["proprios"], actions, t, noise)
    m = _model(d, sd, torch.float32)
    gb = GradBuffer(m)
    loss = _step(m, inp, actions, noise, t, gb)
    assert abs(float(loss) - float(want_loss)) < 1e-4 * max(1.0, float(want_loss))
    got = gb.unpack()
    bad, checked = {}, 0
    for k, g in want.items():
        if float(g.abs().max()) == 0.0:
            assert float(got[k].abs().max()) == 0.0, k
            continue
        e = rel_err(got[k], g)
        checked += 1
        if not e < 2e-3:
            bad[k] = e
    print(f"[train step 2 images, chunk 10] loss {float(loss):.6f} vs {float(want_loss):.6f}; {checked} tensors")
    assert checked >= 50 and not bad, bad
