"""Per-CTA stage stamps of the stream sampler (csrc/denoise_mega3.cu), step 1 / layer 1.  Needs a trace build:
    PZ_NVCC_EXTRA=-DPZ_MEGA_TRACE python open-pi-zero_b200/build.py --force
    python tools/mega3_trace.py [B]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import open_pi_zero_b200 as pz
from open_pi_zero_b200 import _lib
from open_pi_zero_b200.pizero import PiZeroInference
from open_pi_zero_b200.synth import fill_random_

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dims = pz.make_dims()
dev = torch.device("cuda")
m = PiZeroInference(pz.cfg_from_dims(dims), init="empty", device=dev, dtype=torch.bfloat16)
fill_random_(m, dims)
m.pack()
lib = _lib.load()
inp = pz.make_inputs(dims, B, seed=0)
ids = inp["input_ids"].to(dev); pix = inp["pixel_values"].to(dev, torch.bfloat16)
prop = inp["proprios"].to(dev); nz = inp["noise"].to(dev); vlen = inp["valid_len"].to(dev)
out = torch.empty(B, 4, 7, device=dev)
nbytes = lib.pz_workspace_bytes(m._handle, B)
ws_t = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
ws = (ws_t.data_ptr() + 1023) // 1024 * 1024
st = torch.cuda.current_stream().cuda_stream
assert lib.pz_embed_prefix(m._handle, ids.data_ptr(), pix.data_ptr(), ws, nbytes, B, None, st) == 0
assert lib.pz_prefill(m._handle, vlen.data_ptr(), prop.data_ptr(), ws, nbytes, B, None, st) == 0
assert lib.pz_set_sampler(m._handle, 3) == 0
for _ in range(3):
    assert lib.pz_denoise(m._handle, vlen.data_ptr(), nz.data_ptr(), out.data_ptr(), ws, nbytes, B, None, st) == 0
torch.cuda.synchronize()
raw_lib = C.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else lib
fn = raw_lib.pz_debug_mega3_trace
fn.argtypes = [C.c_void_p, C.c_int]
fn.restype = C.c_int
N = 160 * 32
buf = (C.c_ulonglong * N)()
assert fn(buf, N) == 0
t = np.array(buf, dtype=np.int64).reshape(160, 32)
sms = torch.cuda.get_device_properties(0).multi_processor_count
G = sms - 16 * B
g = t[:G]
a = t[G:sms]
t0 = g[:, 0][g[:, 0] > 0].min()
names = {1: "x gathered+normed", 2: "qkv gemv", 3: "qkv published", 4: "att gathered", 5: "o gemv", 6: "x1 published",
         7: "x1 gathered+normed", 8: "gu gemv", 9: "mlp published", 10: "mlp gathered", 11: "down gemv", 12: "x published"}
print(f"B={B}: {G} streaming CTAs, {sms - G} attention CTAs; times in us relative to the earliest layer start")
print(f"{'event':24s} {'min':>8s} {'mean':>8s} {'max':>8s}   (n CTAs)")
for e in range(0, 13):
    col = g[:, e]
    ok = col > 0
    if not ok.any():
        continue
    r = (col[ok] - t0) / 1e3
    print(f"{e:2d} {names.get(e, 'layer start'):21s} {r.min():8.2f} {r.mean():8.2f} {r.max():8.2f}   ({ok.sum()})")
an = {20: "loop top", 21: "qkv gathered", 22: "kv ready + staged", 25: "S tiles (thread 0)", 23: "S + softmax", 24: "PV + published"}
for e in (20, 21, 22, 25, 23, 24):
    col = a[:, e]
    ok = col > 0
    if not ok.any():
        continue
    r = (col[ok] - t0) / 1e3
    print(f"{e:2d} {an[e]:21s} {r.min():8.2f} {r.mean():8.2f} {r.max():8.2f}   ({ok.sum()})")
