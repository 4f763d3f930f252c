#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gemm tests (cluster multicast)"; timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -12
echo "== model tests incl slow"; timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu -s 2>&1 | grep -E "agreement|passed|failed|Error|assert|^E " | tail -25
echo "== bench fp16"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp16_graphs3.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['stage_ms'])"
