"""Does the bs=64 sampler (1.5 k small launches, latency-bound) gain from running sub-batches as concurrent chains on
several streams?  Times pz_denoise for B=64 on one stream against 2 x 32 / 4 x 16 on 2 / 4 streams, each as one CUDA graph.
    python tools/split_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import open_pi_zero_b200 as pz
from open_pi_zero_b200 import _lib
from open_pi_zero_b200.pizero import PiZeroInference
from open_pi_zero_b200.synth import fill_random_

dims = pz.make_dims()
dev = torch.device("cuda")
m = PiZeroInference(pz.cfg_from_dims(dims), init="empty", device=dev, dtype=torch.bfloat16)
fill_random_(m, dims)
m.pack()
lib = _lib.load()
B = 64
inp = pz.make_inputs(dims, B, seed=0)


def prep(lo, hi):
    n = hi - lo
    ids = inp["input_ids"][lo:hi].to(dev); pix = inp["pixel_values"][lo:hi].to(dev, torch.bfloat16)
    prop = inp["proprios"][lo:hi].to(dev); nz = inp["noise"][lo:hi].to(dev).contiguous(); vlen = inp["valid_len"][lo:hi].to(dev)
    nbytes = lib.pz_workspace_bytes(m._handle, n)
    ws_t = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    ws = (ws_t.data_ptr() + 1023) // 1024 * 1024
    st = torch.cuda.current_stream().cuda_stream
    assert lib.pz_embed_prefix(m._handle, ids.data_ptr(), pix.data_ptr(), ws, nbytes, n, None, st) == 0
    assert lib.pz_prefill(m._handle, vlen.data_ptr(), prop.data_ptr(), ws, nbytes, n, None, st) == 0
    out = torch.zeros(n, dims["horizon_steps"], dims["action_dim"], device=dev)
    torch.cuda.synchronize()
    return dict(n=n, nz=nz, vlen=vlen, ws=ws, nbytes=nbytes, out=out, keep=(ws_t, ids, pix, prop))


def denoise(p):
    rc = lib.pz_denoise(m._handle, p["vlen"].data_ptr(), p["nz"].data_ptr(), p["out"].data_ptr(), p["ws"], p["nbytes"], p["n"], None,
                        torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.pz_last_error(m._handle)


ref = None
for parts in (1, 2, 4):
    n = B // parts
    ps = [prep(i * n, (i + 1) * n) for i in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts - 1)]

    def run_all():
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        denoise(ps[0])
        for s, p in zip(streams, ps[1:]):
            with torch.cuda.stream(s):
                denoise(p)
        for s in streams:
            cur.wait_stream(s)

    run_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run_all()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    out = torch.cat([p["out"] for p in ps])
    if ref is None:
        ref = out.clone()
    print(f"B=64 sampler as {parts} chain(s) of {n}: {e0.elapsed_time(e1) / 10:.3f} ms; max|out - single| = {float((out - ref).abs().max()):.3e}",
          flush=True)
    del ps
    torch.cuda.empty_cache()
