#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m "gpu and not slow" -x 2>&1 | tail -30
echo "== msda tests"; timeout 600 python -m pytest tests/test_msda_gpu.py -q -m gpu -x 2>&1 | tail -3
echo "== profile"; timeout 600 python tests/profile_step.py --out gpurun_out/kernels_step_r21.json 2>&1 | grep -v Warn | head -12
