#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== model tests (no slow)"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m "gpu and not slow" 2>&1 | tail -30
echo "== bench fp16 engine"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_fp16_engine2.log | cut -c1-300
echo "== launches fp16 engine"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 3000 --csv --log-file gpurun_out/launches_model_fp16_engine2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu16e2.log 2>&1; tail -1 gpurun_out/bench_under_ncu16e2.log | cut -c1-120
