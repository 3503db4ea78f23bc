"""Flow-matching training step at BASELINE configs[4] (bridge shape, per-GPU batch 32, bf16): this repo's
`pz_flow_matching_step` + fused clip/AdamW (and, under torchrun, the NCCL gradient all-reduce) timed with CUDA events, max over
ranks; `--reference` times the UNMODIFIED reference (baseline/_ref) doing forward + loss.backward() + clip_grad_norm_ +
torch.optim.AdamW eagerly in bf16 on the same GPU (bitsandbytes' AdamW8bit, train.py:171, is not in this image).
Prints one JSON object per arm.

    python tools/train_bench.py [--batch 32] [--steps 5] [--warmup 2] [--reference] [--layers L]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_bench.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import open_pi_zero_b200 as pz


def targets(d, B, seed):
    g = torch.Generator().manual_seed(seed)
    actions = torch.rand((B, d["horizon_steps"], d["action_dim"]), generator=g) * 2 - 1
    noise = torch.randn((B, d["horizon_steps"], d["action_dim"]), generator=g)
    t = pz.FlowTimeSampler("beta").sample_fm_time(B)
    return actions, noise, t


def ours(args, d, dev, rank, world):
    from open_pi_zero_b200.pizero import PiZero
    from open_pi_zero_b200.train import FusedAdamW, GradBuffer, OverlappedAllReduce, allreduce_gradients, flow_matching_step
    import torch.distributed as dist
    m = PiZero(pz.cfg_from_dims(d), init="empty")
    m.load_state_dict(pz.init_state_dict(d, seed=42), strict=True)
    m = m.to(torch.bfloat16).to(dev)
    m.tie_action_proprio_weights()
    B = args.batch
    inp = pz.make_inputs(d, B, seed=rank)
    actions, noise, t = targets(d, B, 100 + rank)
    ids, pix, prop = inp["input_ids"].to(dev), inp["pixel_values"].to(dev, torch.bfloat16), inp["proprios"].to(dev)
    actions, noise, t, vlen = actions.to(dev), noise.to(dev), t.to(dev), inp["valid_len"].to(dev)
    gb = GradBuffer(m)
    opt = FusedAdamW(gb)
    ov = OverlappedAllReduce(gb, wire_dtype=torch.bfloat16 if args.wire_bf16 else torch.float32) if (world > 1 and not args.no_overlap) else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    parts = {"fwd_bwd": 0.0, "allreduce": 0.0, "optimizer": 0.0}
    losses = []

    def step(timed):
        if timed:
            ev[0].record()
        loss = flow_matching_step(m, ids, pix, prop, actions, t, noise=noise, valid_len=vlen, grads=gb, overlap=ov)
        if timed:
            ev[1].record()
        if ov is not None:
            ov.wait()
        else:
            allreduce_gradients(gb)
        if timed:
            ev[2].record()
        opt.step(grad_scale=1.0 / world)
        if timed:
            ev[3].record()
            ev[3].synchronize()
            parts["fwd_bwd"] += ev[0].elapsed_time(ev[1])
            parts["allreduce"] += ev[1].elapsed_time(ev[2])
            parts["optimizer"] += ev[2].elapsed_time(ev[3])
        return loss

    for _ in range(args.warmup):
        losses.append(float(step(False)))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        loss = step(True)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    losses.append(float(loss))
    if rank == 0:
        print(json.dumps({
            "metric": "flow-matching training step (forward + backward + gradient all-reduce + clip + AdamW), samples/s",
            "impl": "ours", "value": world * B / (float(ms) / 1e3), "unit": "samples/s", "n_gpus": world, "ms_per_step": float(ms),
            "per_gpu_batch": B, "dtype": "bf16", "layers": [d["vit_layers"], d["num_layers"]],
            "parts_ms": {k: v / args.steps for k, v in parts.items()}, "allreduce": ("overlapped per layer" + (", bf16 on the wire" if args.wire_bf16 else ", fp32")) if ov is not None else "after backward", "launches_fwd_bwd": m.last_launch_count,
            "losses_first_last": [losses[0], losses[-1]], "grad_buffer_gb": gb.flat.numel() * 4 / 1e9,
            "train_workspace_gb": m._train_ws.numel() / 1e9, "max_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9}))


def reference(args, d, dev):
    from oracle import ref_shims
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "src", "model", "vla")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is missing"}))
        return
    ref_shims.REFERENCE_ROOT = ref_root
    model = ref_shims.build_reference_model(d)
    model.load_state_dict(pz.init_state_dict(d, seed=42), strict=True)
    model = model.to(torch.bfloat16).to(dev).train()
    model.tie_action_proprio_weights()
    model.freeze_unused_weights()
    B = args.batch
    inp = pz.make_inputs(d, B, seed=0)
    actions, noise, t = targets(d, B, 100)
    cm, vpos, ppos, apos = model.build_causal_mask_and_position_ids(inp["attention_mask"], torch.bfloat16)
    kw = dict(input_ids=inp["input_ids"].to(dev), pixel_values=inp["pixel_values"].to(dev, torch.bfloat16), causal_mask=cm.to(dev),
              vlm_position_ids=vpos.to(dev), proprio_position_ids=ppos.to(dev), action_position_ids=apos.to(dev),
              proprios=inp["proprios"].to(dev, torch.bfloat16), actions=actions.to(dev, torch.bfloat16), t=t.to(dev, torch.bfloat16))
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.0)

    def step():
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            loss = model(**kw)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, max_norm=1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        loss = step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    print(json.dumps({"metric": "flow-matching training step, samples/s", "impl": "reference",
                      "what": "unmodified reference (baseline/_ref): autocast bf16 forward + loss.backward() + clip_grad_norm_ + "
                              "torch.optim.AdamW, eager, this GPU", "value": B / (ms / 1e3), "unit": "samples/s", "n_gpus": 1,
                      "ms_per_step": ms, "per_gpu_batch": B, "loss": float(loss),
                      "max_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--wire-bf16", action="store_true", help="round the gradients to bf16 for the all-reduce (half the NVLink bytes)")
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce after the backward instead of layer by layer beside it")
    ap.add_argument("--layers", type=int, default=0, help="debug: shrink to this many Gemma / SigLIP layers")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    d = pz.make_dims() if not args.layers else pz.make_dims(num_layers=args.layers, vit_layers=args.layers)
    if args.reference:
        if rank == 0:
            reference(args, d, dev)
    else:
        ours(args, d, dev, rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
