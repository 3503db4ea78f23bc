"""Launch the four SigLIP GEMM shapes at bs=64 (M = 16384) a few times each (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_pi_zero_b200 import _lib
GELU, OUT_F32, ACCUM = 1, 2, 4
lib = _lib.load()
bf = torch.bfloat16
M = 16384
for (N, K, flags) in ((3456, 1152, 0), (1152, 1152, OUT_F32 | ACCUM), (4304, 1152, GELU), (1152, 4304, OUT_F32 | ACCUM)):
    a = torch.randn(M, K, device="cuda").to(bf)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(bf)
    b = torch.randn(N, device="cuda")
    c = torch.zeros(M, N, device="cuda", dtype=torch.float32 if flags & OUT_F32 else bf)
    for _ in range(3):
        rc = lib.pz_op_linear(1, 1, a.data_ptr(), w.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, K, N, flags, 1.0,
                              torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    torch.cuda.synchronize()
