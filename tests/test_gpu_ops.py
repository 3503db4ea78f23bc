"""GPU: each hand-written kernel through its single-op C-ABI entry point,
against a plain PyTorch fp32 reference of the same op on the same bf16 inputs."""
import ctypes as C

import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu

GELU, OUT_F32, ACCUM, GEGLU, SILU = 1, 2, 4, 8, 16


def _lib():
    from open_pi_zero_b200 import _lib
    return _lib.load()


def _linear_ref(a, w, bias, flags, alpha, c0):
    y = a.float() @ w.float().t()
    if flags & GEGLU:
        n = w.shape[0]
        y = y.view(a.shape[0], n // 256, 2, 128)
        y = (torch.nn.functional.gelu(y[:, :, 0], approximate="tanh") * y[:, :, 1]).reshape(a.shape[0], n // 2)
    else:
        if bias is not None:
            y = y + bias
        if flags & GELU:
            y = torch.nn.functional.gelu(y, approximate="tanh")
        if flags & SILU:
            y = torch.nn.functional.silu(y)
    y = y * alpha
    if flags & ACCUM:
        y = y + c0
    return y


def _run_linear(impl, M, N, K, flags=0, bias=True, alpha=1.0, seed=0):
    lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g)).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g) if (bias and not flags & GEGLU) else None
    n_out = N // 2 if flags & GEGLU else N
    if flags & OUT_F32:
        c = torch.randn(M, n_out, device="cuda", generator=g)
    else:
        c = torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16)
    c0 = c.clone().float()
    rc = lib.pz_op_linear(impl, 1, a.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                          c.data_ptr(), M, N, K, K, n_out, flags, alpha,
                          torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    want = _linear_ref(a, w, b, flags, alpha, c0)
    return rel_err(c.float(), want)


SHAPES = [
    # M, N, K, flags, bias, alpha
    (128, 256, 64, 0, False, 1.0),
    (128, 128, 128, 0, True, 1.0),
    (300, 512, 256, 0, True, 1.0),
    (1000, 4304, 1152, GELU, True, 1.0),            # SigLIP fc1: N tail (4304 = 16*256 + 208)
    (777, 1152, 4304, OUT_F32 | ACCUM, True, 1.0),  # SigLIP fc2: K tail (4304 = 67*64 + 16)
    (513, 3456, 1152, 0, True, 1.0),                # SigLIP qkv
    (552, 2560, 2048, 0, False, 1.0),               # Gemma qkv (2 samples)
    (552, 32768, 2048, GEGLU, False, 1.0),          # Gemma gate|up + GeGLU
    (552, 2048, 16384, OUT_F32 | ACCUM, False, 1.0),  # Gemma down + residual
    (4096, 2048, 1152, OUT_F32, True, 1.0),         # projector
    (256, 1024, 1024, OUT_F32, True, 32.0),         # action encoder linear_3 * sqrt(hidden)
    (256, 1024, 1024, SILU, True, 1.0),
    (2048, 640, 640, OUT_F32 | ACCUM, True, 1.0),   # patch embedding (K padded 588 -> 640)
    # M >= 2048: the cta_group::2 (CTA pair, 256 x 256 tiles) variant
    (2208, 4096, 256, GEGLU, False, 1.0),
    (2304, 4304, 1152, GELU, True, 1.0),
    (2500, 2560, 2048, 0, False, 1.0),               # M tail inside a pair tile (2500 = 9*256 + 196)
    (2049, 1152, 4304, OUT_F32 | ACCUM, True, 1.0),  # second CTA of the last pair almost empty
    # N = 1152 at M >= 2048 (4.5 pair tiles of 256 columns)
    (4096, 1152, 1152, OUT_F32 | ACCUM, True, 1.0),  # SigLIP out_proj + residual
    (2304, 1152, 1152, 0, True, 1.0),
    (2100, 1152, 640, GELU, True, 1.0),
    # A = 72 MB (> 40 MB): the M-fastest raster walks M in groups of <= 24 MB of A rows (3 groups of 23 pair tiles)
    (17664, 768, 2048, 0, True, 1.0),
    (17000, 512, 2048, GEGLU, False, 1.0),           # last group smaller than the others, M tail
]


@pytest.mark.parametrize("M,N,K,flags,bias,alpha", SHAPES)
def test_linear_tcgen05(M, N, K, flags, bias, alpha):
    e = _run_linear(1, M, N, K, flags, bias, alpha)
    print(f"tcgen05 linear M={M} N={N} K={K} flags={flags}: rel err {e:.3e}")
    assert e < 6e-3   # bf16 output rounding (2^-9) dominates; fp32 outputs are ~1e-6


@pytest.mark.parametrize("M,N,K,flags,bias,alpha", [
    (4, 2560, 1024, 0, False, 1.0), (4, 1024, 2048, OUT_F32 | ACCUM, False, 1.0),
    (4, 8192, 1024, GEGLU, False, 1.0), (4, 1024, 4096, OUT_F32 | ACCUM, False, 1.0),
    (1, 2560, 1024, 0, False, 1.0), (16, 1024, 1024, SILU, True, 1.0), (13, 1024, 1024, OUT_F32, True, 32.0),
    (8, 7, 1024, OUT_F32, True, 1.0), (12, 1024, 8, 0, True, 1.0),
    (256, 7, 1024, OUT_F32, True, 1.0)])   # M > 16 only for the tiny-N action decoder (skinny_supported)
def test_linear_skinny(M, N, K, flags, bias, alpha):
    lib = _lib()
    e = _run_linear(2, M, N, K, flags, bias, alpha)
    print(f"skinny linear M={M} N={N} K={K} flags={flags}: rel err {e:.3e}")
    assert e < 6e-3


def _attn_ref(q, k, v, k2, v2, vlen, q_row0, s_vlm, scale, softcap):
    B, R, nh, hd = q.shape
    K = k if k2 is None else torch.cat([k, k2], 1)
    V = v if v2 is None else torch.cat([v, v2], 1)
    s_cache = k.shape[1]
    kvh = K.shape[2]
    out = torch.zeros(B, R, nh, hd, device=q.device)
    for b in range(B):
        for r in range(R):
            tok = q_row0 + r
            j = torch.arange(K.shape[1], device=q.device)
            if vlen is None:
                vis = torch.ones_like(j, dtype=torch.bool)
            else:
                L = int(vlen[b])
                if tok < s_vlm and tok >= L:
                    continue
                vis = j < L
                if tok >= s_vlm:
                    vis = vis | ((j >= s_vlm) & (j < s_cache))
                if tok >= s_cache:
                    vis = vis | (j >= s_cache)
            for h in range(nh):
                kh = h if kvh > 1 else 0
                s = (K[b, :, kh].float() @ q[b, r, h].float()) * scale
                if softcap > 0:
                    s = torch.tanh(s / softcap) * softcap
                s = s.masked_fill(~vis, float("-inf"))
                out[b, r, h] = torch.softmax(s, 0) @ V[b, :, kh].float()
    return out


ATTN_CASES = [
    # name, B, nh, hd, q_rows, q_row0, s_cache, s_vlm, n_fresh, kv_heads, softcap, use_len
    ("siglip", 2, 16, 72, 256, 0, 256, 256, 0, 16, 0.0, False),
    ("prefix_vlm", 3, 8, 256, 276, 0, 277, 276, 0, 1, 50.0, True),
    ("prefix_proprio", 3, 8, 256, 1, 276, 277, 276, 0, 1, 50.0, True),
    ("denoise", 3, 8, 256, 4, 277, 277, 276, 4, 1, 50.0, True),
]


@pytest.mark.parametrize("impl", [0, 1, 2, 3])   # 2 = mma.sync kernel with the split-key scratch, 3 = tcgen05 kernel
@pytest.mark.parametrize("case", ATTN_CASES, ids=[c[0] for c in ATTN_CASES])
def test_attention(impl, case):
    name, B, nh, hd, R, q0, s_cache, s_vlm, n_fresh, kvh, softcap, use_len = case
    lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(1)
    bf = torch.bfloat16
    q = torch.randn(B, R, nh, hd, device="cuda", generator=g).to(bf)
    k = torch.randn(B, s_cache, kvh, hd, device="cuda", generator=g).to(bf)
    v = torch.randn(B, s_cache, kvh, hd, device="cuda", generator=g).to(bf)
    k2 = torch.randn(B, n_fresh, kvh, hd, device="cuda", generator=g).to(bf) if n_fresh else None
    v2 = torch.randn(B, n_fresh, kvh, hd, device="cuda", generator=g).to(bf) if n_fresh else None
    vlen = torch.tensor([258, 276, 267][:B], device="cuda", dtype=torch.int32) if use_len else None
    out = torch.full((B, R, nh, hd), 7.0, device="cuda", dtype=bf)
    scale = hd ** -0.5
    scratch = torch.empty(B * 8 * nh * R * (hd + 2), device="cuda") if impl == 2 else None
    if impl == 2 and R * nh > 64:
        pytest.skip("split-key path is for decode-sized query blocks")
    rc = lib.pz_op_attention(3 if impl == 3 else min(impl, 1), 1, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                             k2.data_ptr() if n_fresh else None, v2.data_ptr() if n_fresh else None,
                             vlen.data_ptr() if use_len else None, out.data_ptr(), B, nh, hd, R, q0,
                             s_cache, s_vlm, n_fresh, kvh, scale, softcap,
                             scratch.data_ptr() if scratch is not None else None,
                             scratch.numel() * 4 if scratch is not None else 0,
                             torch.cuda.current_stream().cuda_stream)
    if impl >= 1 and rc == -1:
        pytest.skip("tensor-core attention does not cover this shape")
    assert rc == 0, rc
    torch.cuda.synchronize()
    want = _attn_ref(q, k, v, k2, v2, vlen, q0, s_vlm, scale, softcap)
    e = rel_err(out.float(), want)
    print(f"attention[{name}] impl={impl}: rel err {e:.3e}")
    assert e < 8e-3


@pytest.mark.parametrize("lens", [[16, 100, 128, 129], [256, 257, 272, 276], [1, 15, 17, 144]])
def test_attention_tcgen05_chunk_boundaries(lens):
    """tcgen05 prefix attention (attn_tc.cu) at valid lengths around the 16-key MMA granule, the 128-key chunk and the
    16-token tile boundaries; pad rows must come back as zeros; poisoned (NaN) K/V rows beyond the valid length must not leak."""
    lib = _lib()
    B, nh, hd, R, s_cache = len(lens), 8, 256, 276, 277
    g = torch.Generator(device="cuda").manual_seed(3)
    bf = torch.bfloat16
    q = torch.randn(B, R, nh, hd, device="cuda", generator=g).to(bf)
    k = torch.randn(B, s_cache, 1, hd, device="cuda", generator=g).to(bf)
    v = torch.randn(B, s_cache, 1, hd, device="cuda", generator=g).to(bf)
    vlen = torch.tensor(lens, device="cuda", dtype=torch.int32)
    want = _attn_ref(q, k, v, None, None, vlen, 0, R, hd ** -0.5, 50.0)
    for b, L in enumerate(lens):          # rows no valid query may read
        k[b, L:] = float("nan")
        v[b, L:] = float("nan")
    out = torch.full((B, R, nh, hd), 7.0, device="cuda", dtype=bf)
    rc = lib.pz_op_attention(3, 1, q.data_ptr(), k.data_ptr(), v.data_ptr(), None, None, vlen.data_ptr(), out.data_ptr(),
                             B, nh, hd, R, 0, s_cache, R, 0, 1, hd ** -0.5, 50.0, None, 0,
                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out.float()).all())
    e = rel_err(out.float(), want)
    print(f"attention tcgen05 lens={lens}: rel err {e:.3e}")
    assert e < 8e-3
    for b, L in enumerate(lens):
        assert float(out[b, L:].float().abs().max()) == 0.0 if L < R else True
