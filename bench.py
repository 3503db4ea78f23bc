#!/usr/bin/env python
"""bench.py -- `PiZero.infer_action` throughput / latency on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

One step = one `infer_action` call (SigLIP + Gemma prefix + 10 Euler steps) over
one batch of synthetic observations of the bridge shape (BASELINE.json
configs[1]): 224 px image, 276 image+text tokens, 1 proprio token, chunk 4.
Per-GPU batch 64, one full replica per GPU, batch-sharded, no collective
(weak scaling).  Prints ONE JSON line on rank 0.

    --config bridge64     (default, the driver's line) BASELINE configs[1]
    --config rollout1024  BASELINE configs[2]: bs = 1024 in total, 1024 / N per GPU (strong scaling)
    --config pi0paper     BASELINE configs[3]: 3 images, 48 text tokens, chunk 50, bs = 256 in total, 256 / N per GPU
    --ref-compiled        also time the unmodified reference under torch.compile(mode="default") on this GPU
                          (eval.py:38-40; the first compiled call takes minutes)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import open_pi_zero_b200 as pz  # noqa: E402

METRIC = "action chunks/s per box (bridge-shape infer_action, bf16); p50 infer_action latency at bs=1"
UNIT = "action_chunks/s"
PER_GPU_BATCH = int(os.environ.get("PZ_BENCH_BATCH", "64"))
# DRAM bytes of one VLM gate|up GEMM launch at bs=64 (cta_group::2 kernel), from the committed ncu capture
GATE_UP_DRAM_BYTES = 609.5e6 + 552.6e6   # dram read + write of one launch (profiles/r01_ncu_gemm_dram_raster.txt)
# DRAM bytes of the bs=1 sampler launch (10 Euler steps): profiles/r02_ncu_mega3_stream.txt (read + write; the writes are the
# exchange buffers and their memset)
MEGA_BS1_DRAM_BYTES = 6.343199e9 + 99.813632e6


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p["bf16_tflops"],
                    tflops_sustained=p["bf16_tflops_sustained"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def build_model(dims, device, cpu_state_dict=None):
    from open_pi_zero_b200.pizero import PiZeroInference
    if cpu_state_dict is not None:
        m = PiZeroInference(pz.cfg_from_dims(dims), init="empty")
        m.load_state_dict(cpu_state_dict, strict=True)
        m = m.to(torch.bfloat16).to(device)
    else:
        m = PiZeroInference(pz.cfg_from_dims(dims), init="empty", device=device, dtype=torch.bfloat16)
        from open_pi_zero_b200.synth import fill_random_
        fill_random_(m, dims, seed=42)
    m.pack()
    return m


def device_inputs(m, inp, device):
    return dict(input_ids=inp["input_ids"].to(device),
                pixel_values=inp["pixel_values"].to(device=device, dtype=torch.bfloat16),
                proprios=inp["proprios"].to(device), noise=inp["noise"].to(device),
                valid_len=inp["valid_len"].to(device))


def run_cpu_oracle(dims, sd, steps, warmup, batch=1):
    """The reference's algorithm on the host cores (oracle port, fp32, all threads)."""
    from oracle import pizero_oracle as O
    inp = pz.make_inputs(dims, batch, seed=0)
    times, out = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.infer_action(sd, dims, inp["input_ids"], inp["pixel_values"], inp["attention_mask"],
                             inp["proprios"], inp["noise"])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, out, inp


def run_cpu_reference(dims, sd, steps, warmup, batch=1):
    """The UNMODIFIED reference (`baseline/_ref`, installed by `pip install --no-deps --target
    baseline/_ref` from /root/reference) through its own `PiZero.infer_action`,
    fp32, all host threads.  The three missing third-party imports (omegaconf, hydra,
    bitsandbytes -- no arithmetic on this path) are the stand-ins of oracle/ref_shims.py.
    Returns None when the install is not there (then the oracle port is timed instead)."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "src", "model", "vla")):
        return None
    from oracle import ref_shims
    ref_shims.REFERENCE_ROOT = ref_root
    try:
        model = ref_shims.build_reference_model(dims)
        model.load_state_dict(sd, strict=True)
    except Exception as e:   # pragma: no cover - environment dependent
        print(f"[bench] reference import failed ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
        return None
    inp = pz.make_inputs(dims, batch, seed=0)
    cm, vpos, ppos, apos = model.build_causal_mask_and_position_ids(inp["attention_mask"], torch.float32)
    pmask, amask = model.split_full_mask_into_submasks(cm)
    orig_randn = torch.randn

    def randn(*a, **kw):   # same initial noise as every other arm (pizero.py:454)
        return inp["noise"].to(kw.get("dtype", torch.float32)).clone()

    times, out = [], None
    torch.randn = randn
    try:
        with torch.inference_mode():
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                out = model.infer_action(input_ids=inp["input_ids"], pixel_values=inp["pixel_values"],
                                         image_text_proprio_mask=pmask, action_mask=amask, vlm_position_ids=vpos,
                                         proprio_position_ids=ppos, action_position_ids=apos, proprios=inp["proprios"])
                dt = time.perf_counter() - t0
                if i >= warmup:
                    times.append(dt)
    finally:
        torch.randn = orig_randn
    return times, out.float(), inp


def reference_on_gpu(dims, sd, device, B, steps=3):
    """Extra (not the reference arm): the unmodified reference's eager PyTorch bf16 `infer_action` on this
    same B200 (the honest same-box bar).  None if baseline/_ref is absent."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "src", "model", "vla")):
        return None
    from oracle import ref_shims
    ref_shims.REFERENCE_ROOT = ref_root
    try:
        model = ref_shims.build_reference_model(dims)
        model.load_state_dict(sd, strict=True)
        model = model.to(torch.bfloat16).to(device)
    except Exception as e:   # pragma: no cover
        print(f"[bench] reference-on-GPU skipped ({type(e).__name__}: {e})", file=sys.stderr)
        return None
    res = {}
    with torch.inference_mode():
        for b, n in ((B, steps), (1, 10)):
            inp = pz.make_inputs(dims, b, seed=0)
            cm, vpos, ppos, apos = model.build_causal_mask_and_position_ids(inp["attention_mask"], torch.bfloat16)
            pmask, amask = model.split_full_mask_into_submasks(cm)
            kw = dict(input_ids=inp["input_ids"].to(device), pixel_values=inp["pixel_values"].to(device, torch.bfloat16),
                      image_text_proprio_mask=pmask.to(device), action_mask=amask.to(device),
                      vlm_position_ids=vpos.to(device), proprio_position_ids=ppos.to(device),
                      action_position_ids=apos.to(device), proprios=inp["proprios"].to(device, torch.bfloat16))
            for _ in range(2):
                model.infer_action(**kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                model.infer_action(**kw)
            e1.record()
            torch.cuda.synchronize()
            res[b] = e0.elapsed_time(e1) / n
    del model
    torch.cuda.empty_cache()
    return dict(what="unmodified reference (baseline/_ref), eager PyTorch bf16 on this GPU, inputs resident",
                batch=B, ms_per_step=res[B], value=B / (res[B] * 1e-3), unit=UNIT, latency_bs1_ms=res[1])


def cpu_arm(dims, sd, steps, warmup):
    """(times, action, inputs, kind): the reference itself when installed, else the oracle port."""
    r = run_cpu_reference(dims, sd, steps, warmup)
    if r is not None:
        return r + ("reference",)
    return run_cpu_oracle(dims, sd, steps, warmup) + ("port",)


def denoise_bs1_roofline(model, dims, device, peaks):
    """HBM roofline of the bs=1 sampler (the second regime BASELINE.json names): the 10 Euler steps as one
    persistent kernel, captured in a CUDA graph and timed with CUDA events on the launching stream.
    Algorithmic bytes per launch: 10 x (629.3 MB of action-expert weights + 5.11 MB prefix KV)."""
    from open_pi_zero_b200 import _lib
    lib = _lib.load()
    inp = pz.make_inputs(dims, 1, seed=7)
    ids = inp["input_ids"].to(device); pix = inp["pixel_values"].to(device, torch.bfloat16)
    prop = inp["proprios"].to(device); nz = inp["noise"].to(device); vlen = inp["valid_len"].to(device)
    out = torch.empty(1, dims["horizon_steps"], dims["action_dim"], device=device)
    nbytes = lib.pz_workspace_bytes(model._handle, 1)
    ws_t = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    ws = (ws_t.data_ptr() + 1023) // 1024 * 1024

    def st():
        return torch.cuda.current_stream().cuda_stream

    assert lib.pz_embed_prefix(model._handle, ids.data_ptr(), pix.data_ptr(), ws, nbytes, 1, None, st()) == 0
    assert lib.pz_prefill(model._handle, vlen.data_ptr(), prop.data_ptr(), ws, nbytes, 1, None, st()) == 0

    def den():
        rc = lib.pz_denoise(model._handle, vlen.data_ptr(), nz.data_ptr(), out.data_ptr(), ws, nbytes, 1, None, st())
        assert rc == 0, lib.pz_last_error(model._handle)

    n0 = lib.pz_launch_count(model._handle)
    den()
    n_launch = lib.pz_launch_count(model._handle) - n0
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        den()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n_steps = dims["num_inference_steps"]
    bytes_per_launch = n_steps * (629.3e6 + 5.11e6)
    achieved = bytes_per_launch / (ms * 1e-3) / 1e9
    return dict(bound="hbm", kernel="denoise_mega3_kernel (10 Euler steps, persistent stream sampler: bulk-copy weight ring, flag exchanges)", achieved=achieved,
                peak=peaks["hbm_gbs"], unit="GB/s", frac=achieved / peaks["hbm_gbs"], traffic=MEGA_BS1_DRAM_BYTES,
                traffic_source="profiles/r02_ncu_mega3_stream.txt (dram read + write of one launch); algorithmic bytes "
                               f"{bytes_per_launch / 1e9:.3f} GB",
                avg_launch_ms=ms, launches_timed=reps, kernels_per_launch=int(n_launch),
                peak_source=peaks["source"] + ", copy bandwidth")


def stage_split_bs1(model, dims, device):
    """bs=1 latency by stage: SigLIP + embedding merge | Gemma prefix (fills the KV cache) | 10 Euler steps, each captured
    in its own CUDA graph and replayed 20 times between CUDA events."""
    from open_pi_zero_b200 import _lib
    lib = _lib.load()
    inp = pz.make_inputs(dims, 1, seed=7)
    ids = inp["input_ids"].to(device); pix = inp["pixel_values"].to(device, torch.bfloat16)
    prop = inp["proprios"].to(device); nz = inp["noise"].to(device); vlen = inp["valid_len"].to(device)
    out = torch.empty(1, dims["horizon_steps"], dims["action_dim"], device=device)
    nbytes = lib.pz_workspace_bytes(model._handle, 1)
    ws_t = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    ws = (ws_t.data_ptr() + 1023) // 1024 * 1024

    def stage(i):
        st = torch.cuda.current_stream().cuda_stream
        if i == 0:
            rc = lib.pz_embed_prefix(model._handle, ids.data_ptr(), pix.data_ptr(), ws, nbytes, 1, None, st)
        elif i == 1:
            rc = lib.pz_prefill(model._handle, vlen.data_ptr(), prop.data_ptr(), ws, nbytes, 1, None, st)
        else:
            rc = lib.pz_denoise(model._handle, vlen.data_ptr(), nz.data_ptr(), out.data_ptr(), ws, nbytes, 1, None, st)
        assert rc == 0, lib.pz_last_error(model._handle)

    lib.pz_set_pixel_format(model._handle, 0)
    for i in range(3):
        stage(i)
    torch.cuda.synchronize()
    res = {}
    for i, name in enumerate(("siglip_embed", "prefix", "denoise_x10")):
        n0 = lib.pz_launch_count(model._handle)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            stage(i)
        n1 = lib.pz_launch_count(model._handle)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            g.replay()
        b.record()
        b.synchronize()
        res[name] = dict(ms=a.elapsed_time(b) / 20, launches=int(n1 - n0))
    return res


def train_step_extra(model, dims, dev_in, device, Bt, steps=3):
    """BASELINE configs[4] on one GPU: `pz_flow_matching_step` (forward + backward) and the fused clip + AdamW, CUDA-event
    timed; next to it the unmodified reference's eager step measured by tools/train_bench.py --reference on the same GPU
    type (quoted from profiles/, it needs its own process: 65 GB of autograd state)."""
    from open_pi_zero_b200.train import FusedAdamW, GradBuffer, flow_matching_step
    g = torch.Generator().manual_seed(7)
    acts = (torch.rand((Bt, dims["horizon_steps"], dims["action_dim"]), generator=g) * 2 - 1).to(device)
    x0 = torch.randn((Bt, dims["horizon_steps"], dims["action_dim"]), generator=g).to(device)
    tt = pz.FlowTimeSampler("beta").sample_fm_time(Bt).to(device)
    gb = GradBuffer(model)
    opt = FusedAdamW(gb)
    kw = dict(noise=x0, valid_len=dev_in["valid_len"][:Bt], grads=gb)
    args_ = (dev_in["input_ids"][:Bt], dev_in["pixel_values"][:Bt], dev_in["proprios"][:Bt], acts, tt)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    flow_matching_step(model, *args_, **kw)
    opt.step()
    torch.cuda.synchronize()
    fb = op = 0.0
    for _ in range(steps):
        ev[0].record()
        loss = flow_matching_step(model, *args_, **kw)
        ev[1].record()
        opt.step()
        ev[2].record()
        ev[2].synchronize()
        fb += ev[0].elapsed_time(ev[1])
        op += ev[1].elapsed_time(ev[2])
    ms = (fb + op) / steps
    out = dict(what="flow-matching training step (BASELINE configs[4]): forward + backward (pz_flow_matching_step) + global-norm clip "
                    "+ AdamW on fp32 master weights (pz_adamw_step), one GPU, no all-reduce", batch=Bt, ms=ms,
               fwd_bwd_ms=fb / steps, optimizer_ms=op / steps, samples_per_s=Bt / (ms / 1e3), loss=float(loss),
               launches_fwd_bwd=model.last_launch_count, parameters_trained=int(gb.flat.numel()))
    # roofline of the step: 3 x the forward's FLOPs (1268.2 GF prefix + 2.68 GF for one velocity evaluation per
    # sample; backward = 2 x) against the measured sustained bf16 rate -- the step is GEMM-bound by construction
    flops = 3.0 * (1268.2e9 + 26.8e9 / 10) * Bt
    peaks = measured_peaks()
    out["roofline"] = dict(bound="tensor", flops_per_step=flops, achieved=flops / (ms * 1e-3) / 1e12, unit="TFLOP/s",
                           peak=peaks["tflops_sustained"], frac=flops / (ms * 1e-3) / 1e12 / peaks["tflops_sustained"],
                           gemm_share="95 ms of GEMM per step in profiles/r02_train_step_launches.txt = 1284 TFLOP/s")
    ref_path = os.path.join(ROOT, "profiles", "r02_train_bench_reference.json")
    if os.path.exists(ref_path):
        with open(ref_path) as fh:
            ref = json.load(fh)
        out["reference_eager_same_gpu_type"] = dict(ms=ref.get("ms_per_step"), samples_per_s=ref.get("value"), source="profiles/r02_train_bench_reference.json")
    del opt, gb
    return out


def reference_compiled(live):
    """The bar the reference deploys (eval.py:38-40): its model under torch.compile(mode="default").  The first compiled
    call takes minutes, so the default run quotes the committed measurement of tools/ref_compiled.py; --ref-compiled
    measures it live."""
    if live:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_compiled.py"), "--batches", "1"],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            return dict(json.loads(r.stdout.strip().splitlines()[-1]), source="measured in this run")
        except Exception:
            return dict(error=(r.stderr or r.stdout)[-300:])
    path = os.path.join(ROOT, "profiles", "r02_ref_compiled.json")
    if os.path.exists(path):
        with open(path) as f:
            return dict(json.load(f), source="profiles/r02_ref_compiled.json (tools/ref_compiled.py, separate run on a B200 of this pool)")
    return None


CONFIGS = {
    # name: (BASELINE.json configs index, total batch or None = PER_GPU_BATCH per GPU, scaling)
    "bridge64": (1, None, "weak"),
    "rollout1024": (2, 1024, "strong"),
    "pi0paper": (3, 256, "strong"),
}


def config_dims(name):
    return pz.make_dims(pz.PI0_PAPER_DIMS) if name == "pi0paper" else pz.make_dims()


def per_gpu_batch(name, world):
    total = CONFIGS[name][1]
    if total is None:
        return PER_GPU_BATCH
    if total % world:
        raise SystemExit(f"--config {name}: total batch {total} does not divide over {world} GPUs")
    return total // world


def active_switches(model=None):
    """Every PZ_* environment switch that is set (12 of them change kernel selection at run time) + the sampler mode."""
    sw = {k: v for k, v in sorted(os.environ.items()) if k.startswith("PZ_")}
    if model is not None:
        sw["sampler_mode"] = {0: "auto", 1: "kernels", 2: "barrier", 3: "stream"}.get(getattr(model, "_sampler_mode", 0), "?")
        sw["stream_sampler_batches"] = list(getattr(model, "_sampler_batches", []))
        sw["cuda_graph"] = bool(model.use_cuda_graph)
    return sw


def workload_config(world, B, name="bridge64"):
    if name == "pi0paper":
        what = (f"Pi0-paper shape infer_action (BASELINE configs[3]): bs={B} per GPU ({world * B} in total), 3 x 224px images "
                "(768 image tokens) + 48 text tokens, 1 proprio, chunk 50, 10 Euler steps, random-init 3.24B-parameter model")
    elif name == "rollout1024":
        what = (f"batched rollout infer_action (BASELINE configs[2]): bs={world * B} in total, {B} per GPU, 224px image, "
                "276 image+text tokens, 1 proprio, chunk 4, 10 Euler steps, random-init 3.24B-parameter model")
    else:
        what = (f"bridge infer_action (BASELINE configs[1]): bs={B} per GPU, 224px image, "
                "276 image+text tokens, 1 proprio, chunk 4, 10 Euler steps, random-init "
                "3.24B-parameter model")
    return dict(workload=what, name=name, global_batch=world * B, per_gpu_batch=B,
                parallelism=f"replica x{world}, batch-sharded, no collective",
                l2="working set per step (6.5 GB bf16 weights + activations) >> 126 MB L2; no flush")


def reference_arm(args, rank):
    if rank != 0:
        return
    dims = config_dims(args.config)
    torch.set_num_threads(os.cpu_count() or 1)
    sd = pz.init_state_dict(dims, seed=42)
    if args.config == "pi0paper":   # the reference has no multi-image path: its algorithm through the oracle port
        times, _, _ = run_cpu_oracle(dims, sd, args.steps, max(args.warmup, 1))
        kind = "port"
    else:
        times, _, _, kind = cpu_arm(dims, sd, args.steps, max(args.warmup, 1))
    total = sum(times)
    value = len(times) / total
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * total / len(times), higher_is_better=True,
                scaling=CONFIGS[args.config][2], vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload_config(args.gpus, per_gpu_batch(args.config, args.gpus), args.config),
                            reference_sample="each step = one infer_action at bs=1 of the same workload, fp32, "
                                             + ("the unmodified reference (baseline/_ref) " if kind == "reference"
                                                else "reference algorithm (oracle port) ") + "on the host cores"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=torch.get_num_threads(), kind=kind,
                                  sample=f"{len(times)} infer_action calls at bs=1, fp32, after "
                                         f"{max(args.warmup, 1)} warm-up"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only")
    ap.add_argument("--skip-latency", action="store_true", help="profiling runs only")
    ap.add_argument("--config", default="bridge64", choices=sorted(CONFIGS))
    ap.add_argument("--ref-compiled", action="store_true",
                    help="also time the unmodified reference under torch.compile(mode='default') (minutes)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the infer_action path has no CPU fallback")
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's own "NCCL version ..." banner (printed by the C library while the
        # communicator is created) goes to stderr instead
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        torch.cuda.set_device(local_rank)
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    warmup = max(args.warmup, 3)
    dims = config_dims(args.config)
    B = per_gpu_batch(args.config, world)
    extras = args.config == "bridge64"          # latency / roofline / CPU legs belong to the headline configuration
    peaks = measured_peaks()

    do_cpu = (world == 1 and not args.no_cpu_baseline and extras)
    sd = pz.init_state_dict(dims, seed=42) if do_cpu else None
    model = build_model(dims, device, sd)
    inp = pz.make_inputs(dims, B, seed=1000 + rank)
    dev_in = device_inputs(model, inp, device)

    def step():
        return model(**dev_in)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        sync_all()
    clocks = clk.summary()
    ms_total = e0.elapsed_time(e1)
    launches = model.last_launch_count * args.steps
    # dominant kernel: the same K steps once more, launched eagerly (CUDA events cannot be
    # read back from a replayed graph) with the library's event taps around every VLM gate|up
    # GEMM on the launching stream
    model.timing_begin(model.TAG_VLM_GATE_UP)
    step()
    sync_all()
    model.timing_end()
    model.timing_begin(model.TAG_VLM_GATE_UP)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        step()
    e3.record()
    sync_all()
    ms_eager = e2.elapsed_time(e3)
    gu_ms, gu_n = model.timing_end()
    t = torch.tensor([ms_total], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end through the public API with host buffers ----------------
    host = {k: v.pin_memory() for k, v in dict(
        input_ids=inp["input_ids"], pixel_values=inp["pixel_values"].to(torch.bfloat16),
        proprios=inp["proprios"], noise=inp["noise"], valid_len=inp["valid_len"]).items()}
    out_host = torch.empty((B, dims["horizon_steps"], dims["action_dim"]), dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()

    def e2e_step():
        dv = {k: v.to(device, non_blocking=True) for k, v in host.items()}
        out = model(**dv)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.synchronize()

    e2e_s = float("nan")
    if not args.skip_e2e:
        for _ in range(2):
            e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())

    # ---- the same, through the reference's stock 8-keyword signature: the two dense block masks and the three position-id
    #      tensors travel from pinned host memory every step, as a drop-in caller (eval.py) ships them
    e2e_stock = None
    if extras and not args.skip_e2e:
        cm, vpos, ppos, apos = model.build_causal_mask_and_position_ids(inp["attention_mask"], torch.bfloat16)
        pmask, amask = model.split_full_mask_into_submasks(cm)
        host2 = {k: v.contiguous().pin_memory() for k, v in dict(
            input_ids=inp["input_ids"], pixel_values=inp["pixel_values"].to(torch.bfloat16),
            image_text_proprio_mask=pmask, action_mask=amask, vlm_position_ids=vpos, proprio_position_ids=ppos,
            action_position_ids=apos, proprios=inp["proprios"].to(torch.bfloat16)).items()}
        h2d2 = sum(v.numel() * v.element_size() for v in host2.values())
        out_host2 = torch.empty((B, dims["horizon_steps"], dims["action_dim"]), dtype=torch.bfloat16).pin_memory()

        def stock_step():
            dv = {k: v.to(device, non_blocking=True) for k, v in host2.items()}
            out = model(**dv)                         # noise drawn inside, like the reference (pizero.py:454)
            out_host2.copy_(out, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            stock_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            stock_step()
        s2 = time.perf_counter() - t0
        t = torch.tensor([s2], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_stock = dict(value=world * B * args.steps / float(t.item()), unit=UNIT, h2d_bytes_per_step=h2d2,
                         d2h_bytes_per_step=out_host2.numel() * out_host2.element_size(),
                         what="model(input_ids, pixel_values, image_text_proprio_mask, action_mask, vlm/proprio/action "
                              "position ids, proprios) with host tensors: masks + position ids copied and checked every step")

    # ---- bs=1 latency (p50) --------------------------------------------------
    lat = None
    if rank == 0 and not args.skip_latency and extras:
        one = {k: v[:1].contiguous() for k, v in dev_in.items()}
        for _ in range(20):
            model(**one)
        torch.cuda.synchronize()
        ts = []
        for _ in range(200):   # p50 of >= 200 calls after 20 warm-up
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            model(**one)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        lat = dict(p50_ms=statistics.median(ts), min_ms=min(ts), launches=model.last_launch_count,
                   stages_ms=stage_split_bs1(model, dims, device))
    sync_all()

    # ---- extra: forward half of BASELINE configs[4] (flow-matching loss value, per-GPU bs=32; no backward) ----
    fm = None
    if rank == 0 and not args.skip_latency and extras:
        from open_pi_zero_b200.pizero import PiZero as _PiZero
        Bt = min(32, B)
        g = torch.Generator().manual_seed(7)
        acts = (torch.rand((Bt, dims["horizon_steps"], dims["action_dim"]), generator=g) * 2 - 1).to(device)
        x0 = torch.randn((Bt, dims["horizon_steps"], dims["action_dim"]), generator=g).to(device)
        tt = torch.rand((Bt,), generator=g).to(device)
        fm_in = dict(input_ids=dev_in["input_ids"][:Bt], pixel_values=dev_in["pixel_values"][:Bt], proprios=dev_in["proprios"][:Bt],
                     valid_len=dev_in["valid_len"][:Bt], actions=acts, t=tt, noise=x0)
        for _ in range(2):
            loss = _PiZero.forward(model, **fm_in)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            loss = _PiZero.forward(model, **fm_in)
        b.record()
        b.synchronize()
        fm = dict(what="PiZero.forward flow-matching loss, FORWARD ONLY (no backward / optimizer): the forward half of "
                       "BASELINE configs[4], per-GPU batch 32, eager launches", batch=Bt, ms=a.elapsed_time(b) / 5,
                  samples_per_s=Bt * 5 / (a.elapsed_time(b) / 1e3), loss=float(loss), launches=model.last_launch_count)
    sync_all()

    # ---- roofline of the dominant kernel (VLM gate|up GEMM, tensor-bound) -----
    Mchunk = min(B, 64) * dims["max_image_text_tokens"]
    flops_per_launch = 2.0 * Mchunk * (2 * dims["vlm_inter"]) * dims["vlm_hidden"]
    roof = None
    if gu_n > 0:
        avg_ms = gu_ms / gu_n
        achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12
        roof = dict(bound="tensor", kernel="gemm_tc_kernel<256, cta_group::2> (VLM gate|up + GeGLU)", achieved=achieved,
                    peak=peaks["tflops_sustained"], unit="TFLOP/s", frac=achieved / peaks["tflops_sustained"],
                    traffic=GATE_UP_DRAM_BYTES, traffic_source="profiles/r01_ncu_gemm_dram_raster.txt (ncu, "
                    "dram__bytes_read.sum + dram__bytes_write.sum of one launch, grouped raster; 2.6 GB before it); algorithmic bytes 785 MB",
                    avg_launch_ms=avg_ms, launches_timed=gu_n,
                    share_of_step=gu_ms / ms_eager, eager_ms_per_step=ms_eager / args.steps, peak_source=peaks["source"] + ", sustained figure")

    roof_denoise = None
    if rank == 0 and not args.skip_latency and extras:
        roof_denoise = denoise_bs1_roofline(model, dims, device, peaks)
    sync_all()

    cpu = None
    ref_gpu = None
    ref_compiled = None
    if do_cpu and rank == 0:
        torch.set_num_threads(os.cpu_count() or 1)
        times, want, cinp, kind = cpu_arm(dims, sd, steps=3, warmup=1)
        got = model(input_ids=cinp["input_ids"].to(device), pixel_values=cinp["pixel_values"].to(device),
                    proprios=cinp["proprios"].to(device), noise=cinp["noise"].to(device),
                    valid_len=cinp["valid_len"].to(device)).cpu()
        ref_gpu = reference_on_gpu(dims, sd, device, B)
        ref_compiled = reference_compiled(args.ref_compiled)
        cpu = dict(value=len(times) / sum(times), unit=UNIT, cores=torch.get_num_threads(), kind=kind,
                   sample="3 infer_action calls at bs=1 (fp32, " + ("the unmodified reference from baseline/_ref"
                          if kind == "reference" else "oracle port of the reference") + ", all host threads) after 1 warm-up",
                   max_abs_gpu_vs_cpu=float((got - want).abs().max()))

    # ---- extra: the full training step of BASELINE configs[4] (forward + backward + clip + AdamW, per-GPU batch 32);
    # LAST, because the optimizer updates the packed weights in place
    train = None
    if rank == 0 and world == 1 and not args.skip_latency and extras and args.config == "bridge64":
        try:
            train = train_step_extra(model, dims, dev_in, device, min(32, B))
        except Exception as e:   # noqa: BLE001 -- an extra must not take the headline line down
            train = dict(error=repr(e)[:300])
    sync_all()

    if rank == 0:
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=warmup,
            ms_per_step=ms_total / args.steps, higher_is_better=True, scaling=CONFIGS[args.config][2], vs_baseline=None,
            dtype="bf16", data="synthetic",
            config=workload_config(world, B, args.config),
            clocks=clocks, e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
            e2e_stock_signature=e2e_stock,
            gpu_launches=launches, latency_bs1=lat, roofline=roof, roofline_denoise_bs1=roof_denoise,
            cpu_baseline=cpu, reference_gpu_eager=ref_gpu, reference_gpu_compiled=ref_compiled, train_forward=fm, train_step=train,
            switches=active_switches(model))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
