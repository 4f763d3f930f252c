#!/usr/bin/env python
"""Per-kernel time table of the APE-L_D step (CUPTI via torch.profiler; development aid, not a bench).

    python tests/profile_step.py [--dtype fp16] [--no-graphs] [--out gpurun_out/kernels_step.json]

Prints kernels sorted by total device time per step, with launch counts, split into kernels of
libape_b200.so ("own") and library kernels, so the next optimisation target is the top line."""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OWN_MARKERS = ("ape_", "msda_", "gemm_tc", "layernorm_", "rope_", "groupnorm_", "vlf_", "nms_", "attn_fwd", "conv3x3",
               "ape::")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernels_step.json"))
    args = ap.parse_args()
    from ape_b200 import configs, synthetic
    from ape_b200.modeling import build_model

    dev = torch.device("cuda", 0)
    tdt = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
    from bench import bench_spec, bench_weights  # same spec / weights / threshold as the benchmarked step

    model = bench_weights(build_model(bench_spec(), num_text=1203)).to(dev)
    model.engine_dtype = tdt
    model.use_cuda_graphs = tdt != torch.float32 and not args.no_graphs
    g = torch.Generator().manual_seed(0)
    imgs = [torch.randint(0, 256, (3, 1024, 1024), generator=g).to(torch.float32).to(dev) for _ in range(2)]

    def step(i):
        return model([{"image": imgs[i % 2], "height": 1024, "width": 1024}])

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(10):
        step(i)
    b.record()
    torch.cuda.synchronize()
    step_ms = a.elapsed_time(b) / 10
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            step(i)
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type is not None and str(ev.device_type).endswith("CUDA"):
            r = rows.setdefault(ev.name, [0.0, 0])
            r[0] += ev.device_time
            r[1] += 1
    tab = sorted(((n, t / args.steps, c / args.steps) for n, (t, c) in rows.items()), key=lambda r: -r[1])
    total = sum(r[1] for r in tab)
    own = sum(r[1] for r in tab if any(m in r[0] for m in OWN_MARKERS))
    print(f"step {step_ms:.3f} ms (events, 10 steps) | kernel time {total / 1e3:.3f} ms/step, own {own / 1e3:.3f} ms, "
          f"library {(total - own) / 1e3:.3f} ms, launches/step {sum(r[2] for r in tab):.0f}")
    for n, t, c in tab[:70]:
        print(f"{t:10.1f} us {c:7.1f}x  {'OWN' if any(m in n for m in OWN_MARKERS) else 'lib'}  {n[:150]}")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    by_grid = []
    try:  # the chrome trace carries grid / block of every kernel record: split the table by launch geometry
        tmp = args.out + ".trace.json"
        prof.export_chrome_trace(tmp)
        g = {}
        events = json.load(open(tmp))["traceEvents"]
        os.remove(tmp)
        t0 = None
        for ev in events:
            if ev.get("cat") == "kernel":
                a = ev.get("args", {})
                key = (ev["name"], str(a.get("grid")), str(a.get("block")))
                r = g.setdefault(key, [0.0, 0])
                r[0] += ev["dur"]
                r[1] += 1
                t0 = ev["ts"] if t0 is None else min(t0, ev["ts"])
        by_grid = sorted(([n, gr, bl, t / args.steps, c / args.steps] for (n, gr, bl), (t, c) in g.items()), key=lambda r: -r[3])
        print("-- by launch geometry (us per step, launches per step, us per launch)")
        for n, gr, bl, t, c in by_grid[:60]:
            print(f"{t:10.1f} us {c:6.1f}x {t / c:8.2f}  {gr:>16s} {bl:>14s}  {n[:110]}")
    except Exception as ex:  # noqa: BLE001
        print("no per-geometry table:", ex)
    json.dump({"step_ms": step_ms, "kernel_us_per_step": total, "own_us_per_step": own, "pdl": os.environ.get("APE_PDL", "1"),
               "kernels": [{"name": n, "us_per_step": t, "launches_per_step": c} for n, t, c in tab],
               "by_geometry": [{"name": n, "grid": gr, "block": bl, "us_per_step": t, "launches_per_step": c} for n, gr, bl, t, c in by_grid]},
              open(args.out, "w"), indent=0)


if __name__ == "__main__":
    main()
