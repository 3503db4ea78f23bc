"""Per-tensor report of the CUDA training step against the fm_width2 reference gradients / the oracle (debug aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import SMALL, pz, rel_err, max_abs
from open_pi_zero_b200.pizero import PiZero
from open_pi_zero_b200.train import GradBuffer, flow_matching_step

which = sys.argv[1] if len(sys.argv) > 1 else "width2"
dtype = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
if which == "width2":
    fx = torch.load(os.path.join(ROOT, "tests/golden/fm_width2.pt"), weights_only=False)
    d = fx["dims"]
    sd = pz.init_state_dict(d, seed=fx["seed"], randomize_norms=fx["randomize_norms"], tie_proprio=fx["tie_proprio"])
    inp = pz.make_inputs(d, fx["batch"], seed=fx["inputs_seed"])
    actions, noise, t = fx["actions"], fx["noise"], fx["t"]
else:
    from oracle import pizero_backward as Bk
    d = SMALL; B = 3
    sd = pz.init_state_dict(d, seed=13, randomize_norms=True, tie_proprio=False)
    inp = pz.make_inputs(d, B, seed=31, min_text=0)
    g = torch.Generator().manual_seed(5)
    actions = torch.rand((B, d["horizon_steps"], d["action_dim"]), generator=g) * 2 - 1
    noise = torch.randn((B, d["horizon_steps"], d["action_dim"]), generator=g)
    t = torch.rand((B,), generator=g)
m = PiZero(pz.cfg_from_dims(d), init="empty"); m.load_state_dict(sd, strict=True); m = m.to(dtype).to("cuda")
gb = GradBuffer(m)
loss = flow_matching_step(m, inp["input_ids"].cuda(), inp["pixel_values"].cuda().to(dtype), inp["proprios"].cuda(), actions.cuda(),
                          t.cuda(), noise=noise.cuda(), valid_len=inp["valid_len"].cuda(), grads=gb)
torch.cuda.synchronize()
got = gb.unpack()
print("loss", float(loss))
if which == "width2":
    print("ref loss", float(fx["ref"]["loss"]))
    for k, wn in fx["ref"]["grad_norms"].items():
        if k in got:
            print(f"{k:90s} ref {float(wn):.4e} got {float(got[k].double().norm()):.4e}")
else:
    wl, want = Bk.flow_matching_backward_full(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], inp["proprios"], actions, t, noise)
    print("oracle loss", float(wl))
    for k, g in want.items():
        if k in got:
            print(f"{k:90s} ref {float(g.norm()):.4e} got {float(got[k].double().norm()):.4e} rel {rel_err(got[k], g):.3e}")
