"""bf16 linear kernels on the shapes the backward of the training step produces (tokens as the K dimension, small / odd K)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_err
from open_pi_zero_b200 import _lib
lib = _lib.load()
OUT_F32, ACCUM = 2, 4

def run(impl, M, N, K, flags, ldc=None):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    c = torch.randn(M, N, device="cuda", generator=g) if flags & OUT_F32 else torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    c0 = c.clone().float()
    rc = lib.pz_op_linear(impl, 1, a.data_ptr(), w.data_ptr(), None, c.data_ptr(), M, N, K, K, N, flags, 1.0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc}"
    want = a.float() @ w.float().t()
    if flags & ACCUM:
        want = want + c0
    return f"{rel_err(c.float(), want):.2e}"

for (M, N, K, fl) in [(8, 128, 16, OUT_F32 | ACCUM), (128, 256, 16, OUT_F32 | ACCUM), (512, 128, 16, OUT_F32 | ACCUM), (128, 128, 16, OUT_F32 | ACCUM),
                      (12, 256, 128, 0), (12, 128, 512, OUT_F32), (12, 128, 8, OUT_F32), (256, 512, 72, OUT_F32 | ACCUM), (1024, 256, 72, OUT_F32 | ACCUM),
                      (72, 512, 256, 0), (72, 256, 1024, OUT_F32), (2560, 256, 72, OUT_F32 | ACCUM), (432, 144, 48, OUT_F32 | ACCUM),
                      (3, 128, 128, OUT_F32), (128, 64, 8, OUT_F32 | ACCUM), (128, 128, 8, OUT_F32 | ACCUM), (2560, 128, 8, OUT_F32 | ACCUM),
                      (2048, 16384, 552, OUT_F32 | ACCUM), (32768, 2048, 552, OUT_F32 | ACCUM), (552, 16384, 2048, 0), (552, 2048, 32768, OUT_F32),
                      (1024, 4096, 8, OUT_F32 | ACCUM), (8, 1024, 8192, OUT_F32)]:
    print(f"M={M:6d} N={N:6d} K={K:6d} flags={fl}: simple {run(0, M, N, K, fl)}  tc {run(1, M, N, K, fl)}  skinny {run(2, M, N, K, fl)}")


# MN-major operands (LIN_A_MN = 128: A stored [K][M]; LIN_W_MN = 256: W stored [K][N]) through pz_op_linear_ex
A_MN, W_MN = 128, 256
def run_mn(impl, M, N, K, flags):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    a_st = a.t().contiguous() if flags & A_MN else a
    w_st = w.t().contiguous() if flags & W_MN else w
    c = torch.randn(M, N, device="cuda", generator=g) if flags & OUT_F32 else torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    c0 = c.clone().float()
    lda = M if flags & A_MN else K
    rc = lib.pz_op_linear_ex(impl, 1, a_st.data_ptr(), w_st.data_ptr(), None, c.data_ptr(), M, N, K, lda, N, N, flags, 1.0,
                             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc}"
    want = a.float() @ w.float().t()
    if flags & ACCUM:
        want = want + c0
    return f"{rel_err(c.float(), want):.2e}"

for (M, N, K, fl) in [(552, 2048, 2560, W_MN), (552, 16384, 2048, W_MN | OUT_F32), (2560, 2048, 552, A_MN | W_MN | OUT_F32 | ACCUM),
                      (32768, 2048, 552, A_MN | W_MN | OUT_F32 | ACCUM), (2048, 16384, 8832, A_MN | W_MN | OUT_F32 | ACCUM),
                      (8832, 2048, 32768, W_MN | OUT_F32), (8832, 16384, 2048, W_MN), (128, 256, 16, A_MN | W_MN | OUT_F32 | ACCUM),
                      (3456, 1152, 8192, A_MN | W_MN | OUT_F32 | ACCUM), (8192, 1152, 4304, W_MN | OUT_F32), (8, 1024, 128, A_MN | W_MN | OUT_F32 | ACCUM),
                      (300, 200, 72, A_MN | OUT_F32), (1152, 640, 8192, A_MN | W_MN | OUT_F32 | ACCUM)]:
    print(f"MN M={M:6d} N={N:6d} K={K:6d} flags={fl}: simple {run_mn(0, M, N, K, fl)}  tc {run_mn(1, M, N, K, fl)}")
