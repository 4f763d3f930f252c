#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gemm+rope, vlf, model tests"; timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_vlf_gpu.py tests/test_model_gpu.py -q -m "gpu and not slow" -x 2>&1 | tail -12
echo "== profile"; timeout 300 python tests/profile_step.py --out gpurun_out/kernels_step_r26.json 2>&1 | grep -v Warn | head -14
