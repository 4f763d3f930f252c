#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "== model tests (no slow)"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m "gpu and not slow" 2>&1 | tail -30
echo "== bench fp16 engine + graphs"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_fp16_graphs.log | cut -c1-300
echo "== bench fp16 engine no graphs"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graphs 2>&1 | tail -1 | cut -c1-200
echo "== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 2600 --csv --log-file gpurun_out/launches_model_fp16_engine3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/bench_under_ncu16e3.log 2>&1; tail -1 gpurun_out/bench_under_ncu16e3.log | cut -c1-120
