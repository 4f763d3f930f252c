#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gemm tests"; timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -6
echo "== sweep"; timeout 600 python tests/perf_gemm.py > gpurun_out/gemm_sweep2.jsonl 2> gpurun_out/gemm_sweep2.err; tail -3 gpurun_out/gemm_sweep2.err
python - <<'PY'
import json
for l in open('gpurun_out/gemm_sweep2.jsonl'):
    r=json.loads(l)
    if 'error' in r: print(r); continue
    print(f"{r['what']:16s} {r['M']:6d} {r['N']:5d} {r['K']:5d} {r['variant']:9s} {r['us']:8.1f} us {r['tflops']:7.1f} TF/s  diff {r['max_diff_vs_first']:.1e}  {r['ms_per_step']} ms/step")
PY
echo "== norm tests"; timeout 300 python -m pytest tests/test_norm_rope_gpu.py tests/test_model_gpu.py -q -m "gpu and not slow" -x 2>&1 | tail -4
echo "== profile"; timeout 600 python tests/profile_step.py --out gpurun_out/kernels_step_r17.json 2>&1 | grep -v Warn | head -16
