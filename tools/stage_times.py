"""Time the three stages of infer_action separately (each captured in its own CUDA graph).
The per-phase stamps of the persistent sampler need a trace build:
    PZ_NVCC_EXTRA=-DPZ_MEGA_TRACE python open-pi-zero_b200/build.py --force"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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

def stage(i):
    st = torch.cuda.current_stream().cuda_stream
    if i == 0: rc = lib.pz_embed_prefix(m._handle, ids.data_ptr(), pix.data_ptr(), ws, nbytes, B, None, st)
    elif i == 1: rc = lib.pz_prefill(m._handle, vlen.data_ptr(), prop.data_ptr(), ws, nbytes, B, None, st)
    else: rc = lib.pz_denoise(m._handle, vlen.data_ptr(), nz.data_ptr(), out.data_ptr(), ws, nbytes, B, None, st)
    assert rc == 0, lib.pz_last_error(m._handle)

for i in range(3): stage(i)
torch.cuda.synchronize()
names = ["siglip+embed", "prefix joint", "denoise x10"]
tot = 0
for i in range(3):
    n0 = lib.pz_launch_count(m._handle)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        stage(i)
    n1 = lib.pz_launch_count(m._handle)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tot += ms
    print(f"B={B} {names[i]:14s}: {ms:8.3f} ms  ({n1 - n0} launches, {1e3 * ms / max(n1 - n0, 1):.2f} us/launch)")
print(f"B={B} total {tot:.3f} ms")

# phase timestamps of the persistent sampler (CTA 0, step 1, layer 1)
off = lib.pz_debug_trace_offset(m._handle, B)
base = (ws - ws_t.data_ptr()) + off + 128
raw = ws_t[base: base + 24 * 8].view(torch.int64).cpu().tolist()
if raw[0] > 0:
    names2 = ["QKV", "bar", "ATT", "bar", "O", "bar", "GU", "bar", "D", "bar"]
    print("mega trace (us):", " ".join(f"{n}={(raw[i+1]-raw[i])/1e3:.2f}" for i, n in enumerate(names2)), f"layer={(raw[10]-raw[0])/1e3:.2f}")
    print("GU items (us): stage->", f"{(raw[11]-raw[6])/1e3:.2f}", " ".join(
        f"[wait={(raw[12+4*k]-raw[11+4*k])/1e3:.2f} gemv={(raw[13+4*k]-raw[12+4*k])/1e3:.2f} prefetch={(raw[14+4*k]-raw[13+4*k])/1e3:.2f}]" for k in range(2)))
    if raw[11] > 0:
        print(f"  attention (us): stage={(raw[11]-raw[2])/1e3:.2f} S={(raw[12]-raw[11])/1e3:.2f} softmax={(raw[13]-raw[12])/1e3:.2f} PV+store={(raw[3]-raw[13])/1e3:.2f};"
              f"  o_proj: combine-stage={(raw[14]-raw[4])/1e3:.2f} item={(raw[5]-raw[14])/1e3:.2f}")
