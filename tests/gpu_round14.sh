#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== new-kernel tests"; timeout 900 python -m pytest tests/test_nms_gpu.py tests/test_norm_rope_gpu.py tests/test_vlf_gpu.py -q -m gpu -x 2>&1 | tail -15
echo "== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m "gpu and not slow" -x 2>&1 | tail -25
echo "== profile step (graphs)"; timeout 600 python tests/profile_step.py --out gpurun_out/kernels_step_r14.json 2>&1 | tail -45
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r14.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['launch_ms'], d['roofline']['frac'], d['stage_ms']); print(json.dumps(d['own_kernel_ms_per_step']))"
