#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gemm/norm/nms tests"; timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_norm_rope_gpu.py tests/test_nms_gpu.py -q -m gpu 2>&1 | tail -15
echo "== model tests (no slow)"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m "gpu and not slow" 2>&1 | tail -15
echo "== bench fp16 engine + graphs"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_fp16_graphs.log | cut -c1-2500
