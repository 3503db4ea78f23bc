"""TEST INFRASTRUCTURE ONLY -- `tests/golden/fm_*.pt`: the flow-matching training
forward (`PiZero.forward`, pizero.py:607-661) of the UNMODIFIED reference
(`/root/reference`, imported via `oracle/ref_shims.py`), fp32 on CPU.

Run here (the authoring container), not on the GPU box:
    python oracle/make_golden_fm.py [tiny] [width2] [bridge]

x0 is injected by patching `torch.randn_like` for the duration of the call
(pizero.py:622 draws it through the module-global `torch`); v_psi is captured
with a forward hook on `action_decoder`.  The proprio and action experts carry
DIFFERENT weights here (training does not tie them), so the fixture also pins
the separate-parameter-set path.  Weights of the big cases are rebuilt from
the seed by `open-pi-zero_b200/synth.py::init_state_dict`.
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from oracle.make_golden import BRIDGE, TINY, WIDTH2, synth  # noqa: E402

CASES = {
    "tiny": dict(dims=TINY, batch=3, seed=17, store_weights=True, randomize_norms=True),
    "width2": dict(dims=WIDTH2, batch=2, seed=21, store_weights=False, randomize_norms=True),
    "bridge": dict(dims=BRIDGE, batch=2, seed=43, store_weights=False, randomize_norms=False),
}


def make_targets(dims, batch, seed):
    g = torch.Generator().manual_seed(seed)
    H, A = dims["horizon_steps"], dims["action_dim"]
    actions = torch.rand((batch, H, A), generator=g) * 2 - 1
    noise = torch.randn((batch, H, A), generator=g)
    t = torch.rand((batch,), generator=g) * 0.98 + 0.01
    return actions, noise, t


def run_reference(dims, sd, inp, actions, noise, t):
    ref_shims.install()
    model = ref_shims.build_reference_model(dims)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    cm, vpos, ppos, apos = model.build_causal_mask_and_position_ids(inp["attention_mask"], torch.float32)
    cap = {}
    hook = model.action_decoder.register_forward_hook(lambda m, i, o: cap.__setitem__("v_psi", o.detach().clone()))
    orig = torch.randn_like
    torch.randn_like = lambda x, **kw: noise.to(kw.get("dtype", x.dtype)).clone()
    try:
        with torch.inference_mode():
            t0 = time.time()
            loss = model.forward(input_ids=inp["input_ids"], pixel_values=inp["pixel_values"], causal_mask=cm,
                                 vlm_position_ids=vpos, proprio_position_ids=ppos, action_position_ids=apos,
                                 proprios=inp["proprios"], actions=actions, t=t)
            dt = time.time() - t0
    finally:
        torch.randn_like = orig
        hook.remove()
    return dict(loss=loss.clone(), v_psi=cap["v_psi"], seconds=dt)


def run_reference_grads(dims, sd, inp, actions, noise, t):
    """d loss / d parameter of the unmodified reference (its own autograd), fp32 CPU; small configs only."""
    ref_shims.install()
    model = ref_shims.build_reference_model(dims)
    model.load_state_dict(sd, strict=True)
    model.train()
    for p in model.parameters():
        p.requires_grad_(True)
    cm, vpos, ppos, apos = model.build_causal_mask_and_position_ids(inp["attention_mask"], torch.float32)
    orig = torch.randn_like
    torch.randn_like = lambda x, **kw: noise.to(kw.get("dtype", x.dtype)).clone()
    try:
        loss = model.forward(input_ids=inp["input_ids"], pixel_values=inp["pixel_values"], causal_mask=cm,
                             vlm_position_ids=vpos, proprio_position_ids=ppos, action_position_ids=apos,
                             proprios=inp["proprios"], actions=actions, t=t)
        loss.backward()
    finally:
        torch.randn_like = orig
    return {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}


def main(argv):
    for name in argv or ["tiny", "width2"]:
        case = CASES[name]
        dims = case["dims"]
        sd = synth.init_state_dict(dims, seed=case["seed"], randomize_norms=case["randomize_norms"], tie_proprio=False)
        inp = synth.make_inputs(dims, case["batch"], seed=case["seed"] + 100)
        actions, noise, t = make_targets(dims, case["batch"], case["seed"] + 200)
        ref = run_reference(dims, sd, inp, actions, noise, t)
        print(f"[fm_{name}] reference forward {ref['seconds']:.2f}s, loss {float(ref['loss']):.6f}", flush=True)
        fx = dict(case=name, dims=dims, seed=case["seed"], batch=case["batch"], randomize_norms=case["randomize_norms"],
                  tie_proprio=False, inputs_seed=case["seed"] + 100, targets_seed=case["seed"] + 200,
                  actions=actions, noise=noise, t=t, ref=dict(loss=ref["loss"], v_psi=ref["v_psi"]),
                  reference="shroglck/open-pi-zero", torch=torch.__version__)
        if case["store_weights"]:
            fx["state_dict"] = sd
            fx["inputs"] = inp
            fx["ref"]["grads"] = run_reference_grads(dims, sd, inp, actions, noise, t)   # the training step's backward
        elif name == "width2":
            # real widths: per-tensor Frobenius norm of the reference's gradient plus the first 4 rows (or elements) of each
            # tensor -- enough to pin a backward at these shapes without storing 2.6 GB
            g = run_reference_grads(dims, sd, inp, actions, noise, t)
            fx["ref"]["grad_norms"] = {k: v.double().norm().float() for k, v in g.items()}
            fx["ref"]["grad_heads"] = {k: v.reshape(v.shape[0], -1)[:4, :64].clone() if v.dim() > 1 else v[:64].clone()
                                       for k, v in g.items()}
            del g
        path = os.path.join(ROOT, "tests", "golden", f"fm_{name}.pt")
        torch.save(fx, path)
        print(f"[fm_{name}] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB)", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
