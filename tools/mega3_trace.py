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
khz = torch.cuda.get_device_properties(0).clock_rate   # kHz (max SM clock; the kernel runs at it: no power cap at bs 1)
n_it = dims["num_layers"] * dims["num_inference_steps"]
g = t[:G].astype(np.float64) / khz * 1e3 / n_it          # us per (layer, step)
a = t[G:sms].astype(np.float64) / khz * 1e3 / n_it
names = {0: "(between layers)", 1: "x gather", 2: "qkv gemv", 14: "qkv ring wait", 3: "qkv epilogue+publish",
         4: "att gather", 5: "o gemv", 17: "o ring wait", 6: "o epilogue+publish",
         7: "x1 gather", 8: "gu gemv", 20: "gu ring wait", 9: "GeGLU epilogue",
         16: "down partial mma", 28: "down ring wait", 18: "publish partials",
         10: "partials gather+sum", 12: "residual + publish x", 31: "encoder/decoder/other"}
print(f"B={B}: {G} streaming CTAs, {sms - G} attention CTAs; us per (layer, step), thread 0 of each CTA")
print(f"{'bucket':26s} {'min':>7s} {'mean':>7s} {'max':>7s}")
tot = 0.0
for e in (0, 1, 2, 14, 3, 4, 5, 17, 6, 7, 8, 20, 9, 16, 28, 18, 10, 12, 31):
    col = g[:, e]
    if col.max() == 0:
        continue
    print(f"{e:2d} {names.get(e, '?'):23s} {col.min():7.2f} {col.mean():7.2f} {col.max():7.2f}")
    tot += col.mean()
print(f"   sum of means {tot:.2f} us")
an = {20: "(loop top)", 21: "qkv gather", 26: "kv tiles wait", 22: "stage q/k/v", 25: "S tiles", 23: "S tail + barrier", 24: "PV + publish"}
tot = 0.0
for e in (20, 21, 26, 22, 25, 23, 24):
    col = a[:, e]
    print(f"{e:2d} {an[e]:23s} {col.min():7.2f} {col.mean():7.2f} {col.max():7.2f}")
    tot += col.mean()
print(f"   sum of means {tot:.2f} us")
