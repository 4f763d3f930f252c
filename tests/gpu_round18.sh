#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== msda tests"; timeout 600 python -m pytest tests/test_msda_gpu.py -q -m gpu -x 2>&1 | tail -12
echo "== perf"; timeout 300 python tests/perf_msda_self.py 2>&1 | tail -14
echo "== model tests"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m "gpu and not slow" -x 2>&1 | tail -4
echo "== profile"; timeout 600 python tests/profile_step.py --out gpurun_out/kernels_step_r18.json 2>&1 | grep -v Warn | head -8
