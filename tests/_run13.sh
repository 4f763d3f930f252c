#!/bin/bash
# round-2 GPU run 13: attention f32x2, conv on the CTA pair, fused reference update, decoder buffers: tests + A/B
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_attn_gpu.py tests/test_norm_rope_gpu.py tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_model_ld_gpu.py tests/test_text_gpu.py tests/test_preprocess_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | cut -c1-220
echo "== attention"; timeout 300 python tests/perf_attn.py 2>&1 | tee gpurun_out/perf_attn13.txt | cut -c1-160
ab() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-microbench 2>gpurun_out/b13_$name.err | tail -1 > gpurun_out/b13_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/b13_{n}.json").read())
    print(n, "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "pred", round(d["e2e_predictor"]["ms_per_step"], 3), "launches", d["gpu_launches"], "own", d.get("own_kernel_ms_per_step"))
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/b13_{n}.err").read()[-1500:])
PY
}
ab base APE_NOP=1
ab convsingle APE_CONV_PAIR=0
ab cudnnconv APE_CONV3X3=0
echo "== predictor"; timeout 300 python tests/perf_predictor.py 2>&1 | tail -11
