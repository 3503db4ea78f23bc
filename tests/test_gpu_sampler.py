"""The three implementations of the Euler loop (pizero.py:454-489) behind pz_denoise -- one kernel per op, the
grid-barrier persistent kernel (csrc/denoise_mega.cu) and the stream sampler (csrc/denoise_mega3.cu) -- against the CPU
oracle and against each other, at the real widths (the persistent kernels only cover act_hidden 1024 / 8 heads x 256)."""
import pytest
import torch

from helpers import max_abs, pz
from oracle import pizero_oracle as O

pytestmark = pytest.mark.gpu

BF16_ACTION_TOL = 1e-2


def _model(d, sd):
    from open_pi_zero_b200.pizero import PiZeroInference
    m = PiZeroInference(pz.cfg_from_dims(d), init="empty")
    m.load_state_dict(sd, strict=True)
    m = m.to(torch.bfloat16).to("cuda")
    m.use_cuda_graph = False
    m.action_dtype = torch.float32
    return m


def _call(m, inp, mode):
    from open_pi_zero_b200 import _lib
    m.pack()
    assert _lib.load().pz_set_sampler(m._handle, mode) == 0
    out = m(input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda().bfloat16(),
            proprios=inp["proprios"].cuda(), noise=inp["noise"].cuda(), valid_len=inp["valid_len"].cuda()).clone()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("batch", [1, 2])
def test_sampler_implementations_agree_with_oracle(batch):
    from open_pi_zero_b200 import _lib
    d = pz.make_dims(vocab_size=1024, image_token_index=1000, num_layers=3, vit_layers=2)
    sd = pz.init_state_dict(d, seed=41, randomize_norms=True)
    inp = pz.make_inputs(d, batch, seed=7, min_text=0)
    want = O.infer_action(sd, d, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                          inp["proprios"], inp["noise"])
    m = _model(d, sd)
    m.pack()
    assert batch in m._sampler_batches, "the stream sampler must cover bs 1 and 2 at the bridge widths"
    outs = {}
    for name, mode in (("kernels", _lib.PZ_SAMPLER_KERNELS), ("barrier", _lib.PZ_SAMPLER_BARRIER),
                       ("stream", _lib.PZ_SAMPLER_STREAM)):
        outs[name] = _call(m, inp, mode)
        e = max_abs(outs[name], want)
        print(f"[sampler {name} B={batch}] max |action - oracle| = {e:.3e}")
        assert torch.isfinite(outs[name]).all()
        assert e < BF16_ACTION_TOL
    assert max_abs(outs["stream"], outs["kernels"]) < 5e-3


@pytest.mark.parametrize("batch", [1, 2])
def test_stream_sampler_is_bit_reproducible_on_one_prefix(batch):
    """Every output element of the stream sampler has one producer and a fixed summation order: two runs over the SAME
    cached prefix (the prefill's split-K reductions are not ordered, so the prefix is filled once) give identical bits."""
    from open_pi_zero_b200 import _lib
    d = pz.make_dims(vocab_size=1024, image_token_index=1000, num_layers=3, vit_layers=2)
    sd = pz.init_state_dict(d, seed=45, randomize_norms=True)
    inp = pz.make_inputs(d, batch, seed=9, min_text=0)
    m = _model(d, sd)
    m.pack()
    lib = _lib.load()
    dev = torch.device("cuda")
    ids = inp["input_ids"].to(dev); pix = inp["pixel_values"].to(dev, torch.bfloat16)
    prop = inp["proprios"].to(dev); nz = inp["noise"].to(dev); vlen = inp["valid_len"].to(dev)
    nbytes = lib.pz_workspace_bytes(m._handle, batch)
    ws_t = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    ws = (ws_t.data_ptr() + 1023) // 1024 * 1024
    st = torch.cuda.current_stream().cuda_stream
    assert lib.pz_embed_prefix(m._handle, ids.data_ptr(), pix.data_ptr(), ws, nbytes, batch, None, st) == 0
    assert lib.pz_prefill(m._handle, vlen.data_ptr(), prop.data_ptr(), ws, nbytes, batch, None, st) == 0
    assert lib.pz_set_sampler(m._handle, _lib.PZ_SAMPLER_STREAM) == 0
    outs = []
    for _ in range(3):
        out = torch.zeros(batch, d["horizon_steps"], d["action_dim"], device=dev)
        assert lib.pz_denoise(m._handle, vlen.data_ptr(), nz.data_ptr(), out.data_ptr(), ws, nbytes, batch, None, st) == 0
        torch.cuda.synchronize()
        outs.append(out)
    lib.pz_set_sampler(m._handle, _lib.PZ_SAMPLER_AUTO)
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_stream_sampler_ragged_valid_lengths_and_padding_content():
    """Keys beyond a sample's valid length must not leak into the stream sampler's attention (its K / V tiles hold
    every cached row): poison the pad embeddings by changing pad ids' pixel / text content and compare."""
    from open_pi_zero_b200 import _lib
    d = pz.make_dims(vocab_size=1024, image_token_index=1000, num_layers=2, vit_layers=2)
    sd = pz.init_state_dict(d, seed=43, randomize_norms=True)
    m = _model(d, sd)
    for seed in (1, 2, 3):
        inp = pz.make_inputs(d, 2, seed=seed, min_text=0)
        a = _call(m, inp, _lib.PZ_SAMPLER_STREAM)
        b = _call(m, inp, _lib.PZ_SAMPLER_KERNELS)
        assert max_abs(a, b) < 5e-3, (seed, inp["valid_len"])


def test_forced_stream_sampler_on_uncovered_batch_is_an_error():
    from open_pi_zero_b200 import _lib
    from open_pi_zero_b200.pizero import PzError
    d = pz.make_dims(vocab_size=1024, image_token_index=1000, num_layers=2, vit_layers=2)
    sd = pz.init_state_dict(d, seed=44)
    m = _model(d, sd)
    inp = pz.make_inputs(d, 3, seed=1)
    with pytest.raises(PzError):
        _call(m, inp, _lib.PZ_SAMPLER_STREAM)
    _call(m, inp, _lib.PZ_SAMPLER_AUTO)
