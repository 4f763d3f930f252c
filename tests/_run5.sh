#!/bin/bash
# round-2 GPU run 4: new GEMM epilogues (tests first), GEMM sweep with clocks, micro-benchmarks, bench
set -u
mkdir -p gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | cut -c1-220
echo "== model tests"; timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_model_ld_gpu.py tests/test_vlf_gpu.py tests/test_text_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | cut -c1-220
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench_r4.json | cut -c1-300
echo "== micro: ex2"; timeout 120 tests/micro/ex2_rate 2>&1 | tee gpurun_out/micro_ex2.txt
echo "== micro: mma"; timeout 120 tests/micro/mma_rate 2>&1 | tee gpurun_out/micro_mma.txt
echo "== gemm phases"; timeout 900 python tests/perf_gemm2.py > gpurun_out/gemm_phases2.jsonl 2>gpurun_out/gemm_phases2.err; tail -3 gpurun_out/gemm_phases2.err
python - <<'PY'
import json
for l in open('gpurun_out/gemm_phases2.jsonl'):
    r = json.loads(l)
    c = r.get('cycles') or {}
    print(f"{r['what']:12s} {r['variant']:6s} warm {r.get('us_warm')} cold {r.get('us_cold')} TF {r.get('tflops_cold')} sust {r.get('us_sustained')} {r.get('clocks_sustained')} | "
          + " ".join(f"{k}={v}" for k, v in c.items()), r.get('error', ''))
PY
echo "== CUPTI per geometry, PDL off"; APE_PDL=0 timeout 400 python tests/profile_step.py --out gpurun_out/kernels_step_nopdl2.json 2>&1 | grep -A40 "by launch geometry" | cut -c1-200
