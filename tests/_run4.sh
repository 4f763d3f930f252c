#!/bin/bash
# round-2 GPU run 3: whole suite after the fixes, bench, per-geometry CUPTI table without PDL overlap, GEMM phase dissection
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== previously failing files"
for f in tests/test_msda_bwd_gpu.py tests/test_text_gpu.py tests/test_nms_gpu.py; do
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | cut -c1-220
done
echo "== whole suite in ONE process"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; tail -25 gpurun_out/pytest_all.log | cut -c1-220
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench_r3.json | cut -c1-300
echo "== CUPTI per geometry, PDL off"; APE_PDL=0 timeout 400 python tests/profile_step.py --out gpurun_out/kernels_step_nopdl.json 2>&1 | grep -v Warn | sed -n 1,140p | cut -c1-230
echo "== gemm phases"; timeout 600 python tests/perf_gemm2.py > gpurun_out/gemm_phases.jsonl 2>gpurun_out/gemm_phases.err; tail -3 gpurun_out/gemm_phases.err; cat gpurun_out/gemm_phases.jsonl | cut -c1-600
