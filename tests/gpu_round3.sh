#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gemm.log
echo "== norm/rope tests"; timeout 300 python -m pytest tests/test_norm_rope_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "== launches fp16"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 8000 --csv --log-file gpurun_out/launches_model_fp16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --dtype fp16 > gpurun_out/bench_under_ncu16.log 2>&1; tail -1 gpurun_out/bench_under_ncu16.log | cut -c1-120
