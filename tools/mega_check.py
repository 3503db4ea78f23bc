"""Compare / time the three implementations of the Euler loop at small batch: one kernel per op (PZ_SAMPLER_KERNELS),
the grid-barrier persistent kernel (denoise_mega.cu) and the stream sampler (denoise_mega3.cu).
  python tools/mega_check.py [B ...] [--layers N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import open_pi_zero_b200 as pz
from open_pi_zero_b200 import _lib
from open_pi_zero_b200.pizero import PiZeroInference
from open_pi_zero_b200.synth import fill_random_

args = sys.argv[1:]
kw = {}
if "--layers" in args:
    i = args.index("--layers")
    kw = dict(num_layers=int(args[i + 1]), vit_layers=2)
    del args[i:i + 2]
dims = pz.make_dims(**kw)
dev = torch.device("cuda")
m = PiZeroInference(pz.cfg_from_dims(dims), init="empty", device=dev, dtype=torch.bfloat16)
fill_random_(m, dims)
m.pack()
print("stream sampler packed for batches", m._sampler_batches)
lib = _lib.load()
NAMES = {1: "kernels", 2: "barrier", 3: "stream"}
for B in [int(a) for a in args] or [1, 2]:
    inp = pz.make_inputs(dims, B, seed=0)
    ids = inp["input_ids"].to(dev); pix = inp["pixel_values"].to(dev, torch.bfloat16)
    prop = inp["proprios"].to(dev); nz = inp["noise"].to(dev); vlen = inp["valid_len"].to(dev)
    nbytes = lib.pz_workspace_bytes(m._handle, B)
    ws_t = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    ws = (ws_t.data_ptr() + 1023) // 1024 * 1024
    st = torch.cuda.current_stream().cuda_stream
    assert lib.pz_embed_prefix(m._handle, ids.data_ptr(), pix.data_ptr(), ws, nbytes, B, None, st) == 0
    assert lib.pz_prefill(m._handle, vlen.data_ptr(), prop.data_ptr(), ws, nbytes, B, None, st) == 0
    torch.cuda.synchronize()
    outs = {}
    for mode in (1, 2, 3):
        if mode == 3 and B not in m._sampler_batches:
            continue
        assert lib.pz_set_sampler(m._handle, mode) == 0
        out = torch.zeros(B, dims["horizon_steps"], dims["action_dim"], device=dev)

        def run():
            rc = lib.pz_denoise(m._handle, vlen.data_ptr(), nz.data_ptr(), out.data_ptr(), ws, nbytes, B, None,
                                torch.cuda.current_stream().cuda_stream)
            assert rc == 0, lib.pz_last_error(m._handle)
        run(); torch.cuda.synchronize()
        outs[mode] = out.clone()
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        gb = dims["num_inference_steps"] * (629.3e6 * dims["num_layers"] / 18 + 5.11e6 * B * dims["num_layers"] / 18)
        print(f"B={B} {NAMES[mode]:8s}: {ms:.3f} ms per {dims['num_inference_steps']}-step denoise "
              f"({gb / ms / 1e6:.0f} GB/s algorithmic); max|out - kernels| = {float((out - outs[1]).abs().max()):.3e}; "
              f"finite={bool(torch.isfinite(out).all())}", flush=True)
    lib.pz_set_sampler(m._handle, 0)
