#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== msda tests"; timeout 900 python -m pytest tests/test_msda_gpu.py -q -m gpu -x 2>&1 | tail -8
echo "== model tests (no slow)"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m "gpu and not slow" 2>&1 | tail -5
echo "== sweep quick"; timeout 600 python tests/perf_msda_sweep.py --quick 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l)
    if r['variant'] in ('reference_kernel_sm100a','ht1_u4','ht1_u2','ht8_u2','default'): print(r['case'],r['dtype'],r['variant'],r['ms_median'],r['alg_GBps'],r['frac_hbm_peak'])"
echo "== bench fp16"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp16_graphs4.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['stage_ms'], d['own_kernel_ms_per_step'])"
