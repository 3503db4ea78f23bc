"""The same-box bar the reference actually deploys (src/agent/eval.py:38-40): the UNMODIFIED reference from baseline/_ref,
bf16, wrapped in torch.compile(mode="default"), timed at bs=1 (p50 of 50 calls) and at bs=64 on this GPU, next to its
eager numbers.  Prints one JSON object.   python tools/ref_compiled.py [--batches 1,64]"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import open_pi_zero_b200 as pz
from oracle import ref_shims


def main():
    batches = [1, 64]
    if "--batches" in sys.argv:
        batches = [int(b) for b in sys.argv[sys.argv.index("--batches") + 1].split(",")]
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "src", "model", "vla")):
        print(json.dumps({"unavailable": "baseline/_ref is missing"}))
        return
    ref_shims.REFERENCE_ROOT = ref_root
    dims = pz.make_dims()
    dev = torch.device("cuda")
    sd = pz.init_state_dict(dims, seed=42)
    model = ref_shims.build_reference_model(dims)
    model.load_state_dict(sd, strict=True)
    del sd
    model = model.to(torch.bfloat16).to(dev).eval()
    model.forward = model.infer_action          # PiZeroInference.forward (pizero.py: class PiZeroInference)
    out = {"what": "unmodified reference (baseline/_ref), bf16, this GPU, inputs resident", "torch": torch.__version__}

    def inputs(b):
        inp = pz.make_inputs(dims, b, seed=0)
        cm, vpos, ppos, apos = model.build_causal_mask_and_position_ids(inp["attention_mask"], torch.bfloat16)
        pmask, amask = model.split_full_mask_into_submasks(cm)
        return dict(input_ids=inp["input_ids"].to(dev), pixel_values=inp["pixel_values"].to(dev, torch.bfloat16),
                    image_text_proprio_mask=pmask.to(dev), action_mask=amask.to(dev), vlm_position_ids=vpos.to(dev),
                    proprio_position_ids=ppos.to(dev), action_position_ids=apos.to(dev),
                    proprios=inp["proprios"].to(dev, torch.bfloat16))

    def time_calls(fn, kw, n):
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(**kw)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts), min(ts)

    with torch.inference_mode():
        for b in batches:
            kw = inputs(b)
            for _ in range(3):
                model.infer_action(**kw)
            torch.cuda.synchronize()
            p50, mn = time_calls(model.infer_action, kw, 50 if b == 1 else 5)
            out[f"eager_bs{b}"] = dict(p50_ms=p50, min_ms=mn, chunks_per_s=b / (p50 * 1e-3))
        compiled = torch.compile(model, mode="default")
        for b in batches:
            kw = inputs(b)
            t0 = time.perf_counter()
            try:
                compiled(**kw)
                torch.cuda.synchronize()
            except Exception as e:   # pragma: no cover - environment dependent
                out[f"compiled_bs{b}"] = dict(error=f"{type(e).__name__}: {str(e)[:300]}")
                continue
            compile_s = time.perf_counter() - t0
            for _ in range(3):
                compiled(**kw)
            torch.cuda.synchronize()
            p50, mn = time_calls(compiled, kw, 50 if b == 1 else 5)
            out[f"compiled_bs{b}"] = dict(p50_ms=p50, min_ms=mn, chunks_per_s=b / (p50 * 1e-3), first_call_s=compile_s)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
