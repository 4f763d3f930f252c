#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== all gpu tests (not slow)"; timeout 900 python -m pytest tests -q -m "gpu and not slow" -x 2>&1 | tail -5
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r27.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['launch_ms'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'], d['clocks']); print(json.dumps(d['own_kernel_ms_per_step']))"
echo "== profile"; timeout 300 python tests/profile_step.py --out gpurun_out/kernels_step_r27.json 2>&1 | grep -v Warn | head -8
