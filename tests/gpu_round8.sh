#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== slow full-size test"; timeout 1500 python -m pytest tests/test_model_gpu.py -q -m "gpu and slow" 2>&1 | tail -25 | tee gpurun_out/pytest_slow.log
echo "== launches (graph mode)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_graph.log 2>&1; tail -1 gpurun_out/bench_under_ncu_graph.log | cut -c1-100; wc -l gpurun_out/launches_graph.csv
echo "== ncu full msda fused"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_fused_fwd_kernel -s 24 -c 2 -o gpurun_out/msda_fused_model -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_full_msda.log 2>&1; tail -1 gpurun_out/ncu_full_msda.log | cut -c1-100
echo "== ncu full gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 300 -c 4 -o gpurun_out/gemm_model -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_full_gemm.log 2>&1; tail -1 gpurun_out/ncu_full_gemm.log | cut -c1-100
ls -la gpurun_out | tail -8
