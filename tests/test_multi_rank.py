"""CPU, world_size 2 (gloo): the N > 1 path -- contiguous batch shards, no data-path
collective, gather in sample order.  The per-rank "replica" here is the CPU oracle on
the tiny golden fixture (the CUDA replica needs a GPU); what is under test is the
host-side shard/gather logic that bench.py and multi-GPU callers use."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, pz


def _worker(rank, world, port, golden, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from open_pi_zero_b200.shard import gather_actions, shard_bounds, shard_inputs
        from oracle import pizero_oracle as O
        fx = torch.load(golden, weights_only=False)
        d, sd, inp = fx["dims"], fx["state_dict"], fx["inputs"]
        B = inp["input_ids"].shape[0]
        mine = shard_inputs(inp, world, rank)
        lo, hi = shard_bounds(B, world, rank)
        assert mine["input_ids"].shape[0] == hi - lo
        local = O.infer_action(sd, d, mine["input_ids"], mine["pixel_values"], mine["attention_mask"],
                               mine["proprios"], mine["noise"])
        dist.barrier()
        full = gather_actions(local, B)
        if rank == 0:
            q.put(float((full - fx["ref"]["action"]).abs().max()))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather(golden_dir):
    golden = os.path.join(golden_dir, "tiny.pt")
    if not os.path.exists(golden):
        pytest.skip("tiny golden fixture missing")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, golden, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # batch 3 over 2 ranks (ragged: 2 + 1), result identical to the reference's full-batch output
    assert q.get(timeout=5) < 2e-5


def test_shard_bounds_cover_batch_exactly():
    from open_pi_zero_b200.shard import shard_bounds
    for batch in (1, 3, 64, 1024, 1027):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


# ------------------------------------------------------------------ training step: gradient all-reduce (N > 1)
def _allreduce_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace

        from open_pi_zero_b200.train import allreduce_gradients
        n = 3 * 1024 + 17
        g = torch.Generator().manual_seed(rank)
        flat = torch.randn(n, generator=g)
        want = sum(torch.randn(n, generator=torch.Generator().manual_seed(r)) for r in range(world))
        allreduce_gradients(SimpleNamespace(flat=flat), bucket_mb=0)      # bucket_mb 0 -> the smallest bucket: many buckets
        if rank == 0:
            q.put(float((flat - want).abs().max()))
    finally:
        dist.destroy_process_group()


def test_two_rank_bucketed_gradient_allreduce():
    """train.py:119-126 (DDP): the flat gradient buffer is summed over the ranks bucket by bucket."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-6


def _ddp_step_worker(rank, world, port, q):
    """Both ranks share cuda:0 (gloo moves CUDA tensors through the host): rank r trains on its own micro-batch with the
    all-reduce overlapped layer by layer; the result must be the sum of the two local gradients."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import SMALL
        from open_pi_zero_b200.pizero import PiZero
        from open_pi_zero_b200.train import FusedAdamW, GradBuffer, OverlappedAllReduce, flow_matching_step
        d = SMALL
        sd = pz.init_state_dict(d, seed=13, randomize_norms=True, tie_proprio=False)
        m = PiZero(pz.cfg_from_dims(d), init="empty")
        m.load_state_dict(sd, strict=True)
        m = m.to(torch.float32).to("cuda")

        def batch(r):
            inp = pz.make_inputs(d, 2, seed=40 + r)
            g = torch.Generator().manual_seed(60 + r)
            a = torch.rand((2, d["horizon_steps"], d["action_dim"]), generator=g) * 2 - 1
            n = torch.randn((2, d["horizon_steps"], d["action_dim"]), generator=g)
            t = torch.rand((2,), generator=g)
            return inp, a, n, t

        def run(r, gb, ov=None):
            inp, a, n, t = batch(r)
            return flow_matching_step(m, inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["proprios"].cuda(), a.cuda(),
                                      t.cuda(), noise=n.cuda(), valid_len=inp["valid_len"].cuda(), grads=gb, overlap=ov)

        want = GradBuffer(m)
        for r in range(world):          # the sum of every rank's local gradient, computed locally without communication
            run(r, want)
        gb = GradBuffer(m)
        ov = OverlappedAllReduce(gb)
        run(rank, gb, ov)
        ov.wait()
        torch.cuda.synchronize()
        err = float((gb.flat - want.flat).norm() / want.flat.norm())
        # the optimizer step with grad_scale 1 / world leaves identical weights on every rank
        opt = FusedAdamW(gb, action_lr=1e-3, vlm_lr=1e-3)
        opt.step(grad_scale=1.0 / world)
        chk = opt.master.double().sum().reshape(1).cpu()
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        q.put((rank, err, float(both[0]) == float(both[1])))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_data_parallel_training_step_overlapped_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ddp_step_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for _ in range(2):
        rank, err, same = q.get(timeout=5)
        assert err < 1e-5, (rank, err)
        assert same


def test_gradient_buckets_cover_the_flat_buffer_exactly_once():
    """bucket_ranges (what OverlappedAllReduce all-reduces per 'gradients final' event): every element of the flat buffer in
    exactly one bucket, buckets in the order the backward finishes them, for tied and untied proprio / action weights."""
    from open_pi_zero_b200.train import MIX_FIELDS, TOP_ACTION, TOP_VLM, VIT_FIELDS, bucket_ranges
    L, LV = 3, 2
    for tied in (True, False):
        entries, off = [], 0

        def add(key, n):
            nonlocal off
            entries.append((key, off, (n,)))
            off += (n + 255) // 256 * 256

        for f in TOP_VLM:
            add(("top", f), 100)
        for i in range(LV):
            for f in VIT_FIELDS:
                add(("vit", i, f), 300 + i)
        for mix in ("vlm",):
            for l in range(L):
                for f in MIX_FIELDS:
                    add((mix, l, f), 1000 + l)
        for f in TOP_ACTION:
            add(("top", f), 64)
        for mix in ("action",) + (() if tied else ("proprio",)):
            for l in range(L):
                for f in MIX_FIELDS:
                    add((mix, l, f), 500 + l)
        ranges, order = bucket_ranges(entries, off, L, LV, tied)
        assert len(ranges) == L + LV + 2 and sorted(order) == list(range(L + LV + 2))
        assert order[:L] == list(range(L - 1, -1, -1)) and order[L] == L and order[-1] == L + LV + 1
        cover = torch.zeros(off, dtype=torch.int32)
        for r in ranges:
            for lo, hi in r:
                cover[lo:hi] += 1
        assert bool((cover == 1).all())
        assert len(ranges[0]) == (2 if tied else 3)
