#!/usr/bin/env python
"""bench.py — driver-facing benchmark (contract in the task statement, §④).

    python bench.py --gpus N --steps K --warmup W [--workload msda|...] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of
synthetic input.  Workloads:

  msda     ms_deform_attn forward at the APE-L_D 1024² encoder shape (B=1 per GPU, Q=S=87 296,
           5 levels, 8 heads x 32, 4 points, fp32) — BASELINE.json's "ms_deform_attn HBM GB/s"
           half of the metric; algorithmic bytes per call as SURVEY.md §8(d).

`value`   : device-resident inputs, CUDA events on the launching stream, max over ranks.
`e2e`     : same metric through the public operator (torch.ops.ape.ms_deform_attn_forward) with
            HOST (pinned) buffers: H2D of value/loc/attn and D2H of the output inside the timed region.
`roofline`: algorithmic bytes per launch / mean launch duration (CUDA events inside the timed
            region) against MEASURED_PEAKS.json's hbm_gbs (fallback 6650 GB/s, said so).
`cpu_baseline`: the oracle port of the reference's CPU path timed on this box's host cores
            (rank 0, N=1 only).  `--impl reference` runs only that arm.
Multi-GPU: the path shards over images with no data-path collective ("weak" scaling: one image
per GPU); launched by torchrun, NCCL is used only for the barrier and the max-over-ranks reduce.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L5_1024 = [(256, 256), (128, 128), (64, 64), (32, 32), (16, 16)]
H, D, P = 8, 32, 4


def msda_bytes(B, S, Q, L, esize):
    """read value once, loc + attn once, write out once (+ level tables)."""
    return esize * (B * S * H * D + B * Q * H * L * P * 2 + B * Q * H * L * P + B * Q * H * D) + 24 * L


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference_arm(steps, warmup):
    """The reference's CPU path for ms_deform_attn = multi_scale_deformable_attn_pytorch
    (ape/layers/multi_scale_deform_attn.py:84-124); /root/reference does not exist on the GPU box,
    so the oracle's port of it (oracle/msda.py:msda_torch) is timed, all host threads."""
    from oracle import msda as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    S = sum(h * w for h, w in L5_1024)
    ins = O.make_inputs(1, S, H, D, L5_1024, P, seed=3)
    value, ss, st, loc, attn = ins
    for _ in range(max(1, min(warmup, 1))):
        O.msda_torch(value, ss, loc, attn)
    n = max(1, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(n):
        O.msda_torch(value, ss, loc, attn)
    dt = (time.perf_counter() - t0) / n
    gbs = msda_bytes(1, S, S, 5, 4) / dt / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"{n} full encoder-shape calls (B=1,Q=S={S},L=5,fp32) of the oracle's port of "
                      "multi_scale_deformable_attn_pytorch, all host threads", "ms_per_step": dt * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="msda")
    ap.add_argument("--impl", default="ape_b200", choices=["ape_b200", "reference"])
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "fp16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    S = sum(h * w for h, w in L5_1024)
    L = len(L5_1024)
    tdt = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
    esize = 4 if args.dtype == "fp32" else 2
    config = {"workload": "ms_deform_attn_forward APE-L_D 1024^2 encoder shape (B=1/GPU, Q=S=87296, L=5, H=8, D=32, P=4)",
              "loc": "uniform(0,1) seed 3 (SURVEY 8d)", "l2": "4 rotating input sets (1.4 GB fp32) > 126 MB L2",
              "parallelism": f"dp{args.gpus} (one image per GPU, no data-path collective)"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(args.steps, args.warmup)
        line = {"impl": "reference", "metric": "ms_deform_attn_algorithmic_GBps", "value": cb["value"], "unit": "GB/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in ape_b200)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    import ape_b200
    from ape_b200 import ops

    # --- inputs: 4 rotating sets so consecutive steps never hit a warm L2 -------------------
    NSETS = 4
    g = torch.Generator().manual_seed(3 + rank)
    ss = torch.tensor(L5_1024, dtype=torch.int64)
    areas = ss[:, 0] * ss[:, 1]
    st = torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
    ss_d, st_d = ss.to(dev), st.to(dev)
    host_sets, dev_sets = [], []
    for i in range(NSETS):
        value = torch.randn(1, S, H, D, generator=g).to(tdt)
        loc = torch.rand(1, S, H, L, P, 2, generator=g).to(tdt)
        attn = torch.randn(1, S, H, L * P, generator=g).softmax(-1).view(1, S, H, L, P).to(tdt)
        host_sets.append(tuple(t.pin_memory() for t in (value, loc, attn)))
        dev_sets.append(tuple(t.to(dev) for t in (value, loc, attn)))

    def step(i):
        v, lo, at = dev_sets[i % NSETS]
        return ops.ms_deform_attn_forward(v, ss_d, st_d, lo, at, 64)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = ape_b200._lib.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for i in range(args.steps):
        evs[i][0].record()
        step(i)
        evs[i][1].record()
    t_end.record()
    barrier()
    launches = ape_b200._lib.launch_count() - n0
    total_ms = t_start.elapsed_time(t_end)
    kern_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps

    # --- e2e: host buffers in, host result out, through the public operator ----------------
    out_host = torch.empty((1, S, H * D), dtype=tdt).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step(i):
        hv, hl, ha = host_sets[i % NSETS]
        v = hv.to(dev, non_blocking=True)
        lo = hl.to(dev, non_blocking=True)
        at = ha.to(dev, non_blocking=True)
        out = torch.ops.ape.ms_deform_attn_forward(v, ss_d, st_d, lo, at, 64)
        out_host.copy_(out, non_blocking=True)

    e2e_step(0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        e2e_step(i)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    t = torch.tensor([total_ms, kern_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_ms = t.tolist()

    if rank == 0:
        nbytes = msda_bytes(1, S, S, L, esize)
        ms_per_step = total_ms / args.steps
        value_gbs = world * nbytes / (ms_per_step * 1e-3) / 1e9
        peak, peak_src = peaks()
        achieved = nbytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"msda_enc_{args.dtype}")
        h2d = sum(t.numel() * t.element_size() for t in host_sets[0])
        d2h = out_host.numel() * out_host.element_size()
        line = {
            "metric": "ms_deform_attn_algorithmic_GBps", "value": value_gbs, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "fp16": "f16", "bf16": "bf16"}[args.dtype],
            "data": "synthetic", "config": config,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "msda_fwd_kernel",
                         "algorithmic_bytes_per_launch": nbytes, "launch_ms": kern_ms},
            "e2e": {"value": world * nbytes / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_arm(2, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
