"""Compare / time the persistent samplers: PZ_MEGA=1 (grid-barrier kernel) vs PZ_MEGA=2 (flag-exchange kernel)."""
import sys, os
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
for B in [int(a) for a in sys.argv[1:]] or [1, 2]:
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
    for ver in ("0", "1", "2"):
        os.environ["PZ_MEGA"] = ver
        out = torch.zeros(B, 4, 7, device=dev)
        def run():
            rc = lib.pz_denoise(m._handle, vlen.data_ptr(), nz.data_ptr(), out.data_ptr(), ws, nbytes, B, None,
                                torch.cuda.current_stream().cuda_stream)
            assert rc == 0, lib.pz_last_error(m._handle)
        run(); torch.cuda.synchronize()
        outs[ver] = out.clone()
        if ver == "0":
            continue
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} PZ_MEGA={ver}: {e0.elapsed_time(e1) / 20:.3f} ms per 10-step denoise; "
              f"max|out - separate kernels| = {float((out - outs['0']).abs().max()):.3e}; finite={bool(torch.isfinite(out).all())}")
    off = lib.pz_debug_ll_trace_offset(m._handle, B) + (ws - ws_t.data_ptr())
    raw = ws_t[off: off + 148 * 16 * 8].view(torch.int64).cpu().view(148, 16)
    G = 148 - 2 * B
    for errw in [int(x) & 0xFFFFFFFF for x in ws_t[off - 256: off - 256 + 20].view(torch.int32).cpu()]:
      if errw:
        print(f"ERR word {errw:#x}: kind={(errw >> 24) & 0x7f} cta={(errw >> 12) & 0xfff} site={errw & 0xf} layer={(errw >> 4) & 0x1f} step&3={(errw >> 9) & 3}")
    if int(raw[0, 0]) > 0:
        t0 = int(raw[:G, 0].min())
        g = (raw[:G].double() - t0) / 1e3
        n = ["start", "QKV.stage", "QKV.acc", "QKV.st", "O.stage", "O.acc", "O.st", "GU.stage", "GU.acc", "GU.st", "D.stage", "D.acc", "D.st", "GU.acc.enter", "GU.acc.firstfull", "GU.acc.mmadone"]
        print("event: min / median / max over streaming CTAs (us since first CTA entered the layer)")
        for i, name in enumerate(n):
            col = g[:, i][raw[:G, i] > 0]
            if len(col): print(f"  {name:10s} {col.min():7.2f} {col.median():7.2f} {col.max():7.2f}  (n={len(col)})")
        a = (raw[G:G + 2 * B].double() - t0) / 1e3
        an = ["start", "stage", "kv", "S+P", "PV+st"]
        for r in range(2 * B):
            print(f"  ATT cta {r}: " + " ".join(f"{an[i]}={a[r, i]:.2f}" for i in range(5)))
    print(f"B={B} max|v2 - v1| = {float((outs['2'] - outs['1']).abs().max()):.3e}")
