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
