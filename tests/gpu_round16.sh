#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pair-kernel smoke (short timeout: a protocol bug must not hang the box)"
timeout 120 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x -k "plain_gemm" 2>&1 | tail -8
rc=$?
if [ $rc -ne 0 ]; then echo "pair smoke failed rc=$rc"; fi
echo "== all gemm tests"; timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu 2>&1 | tail -8
echo "== sweep"; timeout 600 python tests/perf_gemm.py > gpurun_out/gemm_sweep.jsonl 2> gpurun_out/gemm_sweep.err; tail -3 gpurun_out/gemm_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/gemm_sweep.jsonl'):
    r=json.loads(l)
    if 'error' in r: print(r); continue
    print(f"{r['what']:16s} {r['M']:6d} {r['N']:5d} {r['K']:5d} {r['variant']:8s} {r['us']:8.1f} us {r['tflops']:7.1f} TF/s  diff {r['max_diff_vs_first']:.1e}  {r['ms_per_step']} ms/step")
PY
echo "== model tests + nms + norm"; timeout 900 python -m pytest tests/test_model_gpu.py tests/test_nms_gpu.py tests/test_norm_rope_gpu.py -q -m "gpu and not slow" -x 2>&1 | tail -6
echo "== profile"; timeout 600 python tests/profile_step.py --out gpurun_out/kernels_step_r16.json 2>&1 | grep -v Warn | head -24
