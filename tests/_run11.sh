#!/bin/bash
# round-2 GPU run 11: single-pass VLF pooling, lean rope epilogue, CTA-pair policy: tests + whole-step A/B
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_vlf_gpu.py tests/test_model_gpu.py tests/test_model_ld_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | cut -c1-220
ab() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-microbench 2>gpurun_out/b11_$name.err | tail -1 > gpurun_out/b11_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/b11_{n}.json").read())
    print(n, "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "own", d.get("own_kernel_ms_per_step"))
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/b11_{n}.err").read()[-1500:])
PY
}
ab base APE_NOP=1
ab rope APE_FUSED_ROPE=1
ab single APE_GEMM_POLICY=single
echo "== CUPTI per geometry, PDL off, fused rope"; APE_FUSED_ROPE=1 APE_PDL=0 timeout 400 python tests/profile_step.py --out gpurun_out/kernels_step_nopdl11.json 2>&1 | grep -v Warn | head -24 | cut -c1-200
