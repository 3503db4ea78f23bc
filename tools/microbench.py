"""Micro-benchmarks of single kernels through the C ABI (CUDA-event timed, back to back)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_pi_zero_b200 import _lib

GELU, OUT_F32, ACCUM, GEGLU, SILU = 1, 2, 4, 8, 16
lib = _lib.load()


def bench_linear(impl, M, N, K, flags, iters=200, nbuf=8):
    bf = torch.bfloat16
    a = torch.randn(M, K, device="cuda").to(bf)
    # rotate over several weight buffers so that weights come from HBM, as in the model
    ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).to(bf) for _ in range(nbuf)]
    n_out = N // 2 if flags & GEGLU else N
    c = torch.zeros(M, n_out, device="cuda", dtype=torch.float32 if flags & OUT_F32 else bf)
    st = torch.cuda.current_stream().cuda_stream
    def run(i):
        nonlocal st
        rc = lib.pz_op_linear(impl, 1, a.data_ptr(), ws[i % nbuf].data_ptr(), None, c.data_ptr(), M, N, K, K, n_out, flags, 1.0, st)
        assert rc == 0, rc
    for i in range(10): run(i)
    torch.cuda.synchronize()
    # capture the launches into a CUDA graph so the GPU-side time is measured, not the host launch rate
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        for i in range(iters): run(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    gb = (N * K * 2) / 1e9
    tf = 2.0 * M * N * K / 1e12
    print(f"linear impl={impl} M={M:5d} N={N:5d} K={K:5d} flags={flags:2d}: {us:8.1f} us  {gb/us*1e6:7.0f} GB/s(W)  {tf/us*1e6:7.1f} TF/s")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("floor",):   # what a launch costs when there is next to nothing to compute
        bench_linear(1, 128, 128, 64, 0)
        bench_linear(1, 128, 128, 64, OUT_F32 | ACCUM)
        bench_linear(1, 128, 256, 512, 0)
        bench_linear(1, 256, 2560, 1024, 0)
        bench_linear(1, 256, 1024, 2048, OUT_F32 | ACCUM)
        bench_linear(1, 256, 8192, 1024, GEGLU)
        bench_linear(1, 256, 1024, 4096, OUT_F32 | ACCUM)
        bench_linear(1, 276, 2560, 2048, 0, iters=50)
        bench_linear(1, 276, 2048, 2048, OUT_F32 | ACCUM, iters=50)
        bench_linear(1, 276, 32768, 2048, GEGLU, iters=20)
        bench_linear(1, 276, 2048, 16384, OUT_F32 | ACCUM, iters=20)
    if which in ("all", "tc"):
        for M in (256,):
            bench_linear(1, M, 2560, 1024, 0)
            bench_linear(1, M, 1024, 2048, OUT_F32 | ACCUM)
            bench_linear(1, M, 8192, 1024, GEGLU)
            bench_linear(1, M, 1024, 4096, OUT_F32 | ACCUM)
        for M in (276, 17664):
            bench_linear(1, M, 2560, 2048, 0, iters=50)
            bench_linear(1, M, 2048, 2048, OUT_F32 | ACCUM, iters=50)
            bench_linear(1, M, 32768, 2048, GEGLU, iters=20)
            bench_linear(1, M, 2048, 16384, OUT_F32 | ACCUM, iters=20)
        for M in (256, 16384):
            bench_linear(1, M, 3456, 1152, 0, iters=50)
            bench_linear(1, M, 1152, 1152, OUT_F32 | ACCUM, iters=50)
            bench_linear(1, M, 4304, 1152, GELU, iters=50)
            bench_linear(1, M, 1152, 4304, OUT_F32 | ACCUM, iters=50)
    if which in ("all", "skinny"):
        for M in (4, 8, 32, 64):
            bench_linear(2, M, 2560, 1024, 0)
            bench_linear(2, M, 1024, 2048, OUT_F32 | ACCUM)
            bench_linear(2, M, 8192, 1024, GEGLU)
            bench_linear(2, M, 1024, 4096, OUT_F32 | ACCUM)
    if which == "floor":
        # fixed cost of one tcgen05 GEMM launch in a PDL chain: tiny K, then growing K / N
        for (M, N, K) in ((128, 128, 64), (256, 1024, 64), (256, 2560, 64), (256, 2560, 256), (256, 2560, 512),
                          (256, 2560, 1024), (256, 1024, 2048), (256, 1024, 4096)):
            bench_linear(1, M, N, K, 0)
        for (M, N, K) in ((256, 1024, 64), (256, 1024, 1024), (256, 1024, 2048), (256, 1024, 4096)):
            bench_linear(1, M, N, K, OUT_F32 | ACCUM)
