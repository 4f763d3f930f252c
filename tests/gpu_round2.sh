#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench fp32"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_fp32.log
echo "== bench fp16"; timeout 900 python bench.py --steps 10 --warmup 3 --dtype fp16 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_fp16.log
echo "== bench bf16"; timeout 900 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_bf16.log
echo "== launches fp32"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 6000 --csv --log-file gpurun_out/launches_model_fp32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
ls -la gpurun_out | head -30
