#!/bin/bash
# round-2 GPU run 8: GEMM resident-operand schedules + epilogue prefetch: tests, per-shape timings, whole step
set -u
mkdir -p gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | cut -c1-220
echo "== gemm shapes (graph-timed)"; PERF_GEMM_VARIANTS=${PERF_GEMM_VARIANTS:-single,pair} timeout 900 python tests/perf_gemm2.py > gpurun_out/gemm_phases8.jsonl 2>gpurun_out/gemm_phases8.err; tail -2 gpurun_out/gemm_phases8.err
python - <<'PY'
import json
for l in open('gpurun_out/gemm_phases8.jsonl'):
    r = json.loads(l)
    c = r.get('cycles') or {}
    print(f"{r['what']:12s} {r['variant']:6s} warm {r.get('us_warm')} cold {r.get('us_cold')} TF {r.get('tflops_cold')} | "
          + " ".join(f"{k}={v}" for k, v in c.items()), r.get('error', ''))
PY
echo "== model tests"; timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_model_ld_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | cut -c1-220
for v in base:APE_NOP=1 nores:APE_GEMM_RESIDENT=0; do
  name=${v%%:*}; envv=${v##*:}
  echo "== bench $name"; env $envv timeout 600 python bench.py --no-cpu-baseline --no-microbench 2>gpurun_out/b8_$name.err | tail -1 > gpurun_out/b8_$name.json
  python - $name <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/b8_{n}.json").read())
    print(n, "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "own", d.get("own_kernel_ms_per_step"))
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/b8_{n}.err").read()[-1500:])
PY
done
echo "== CUPTI per geometry, PDL off"; APE_PDL=0 timeout 400 python tests/profile_step.py --out gpurun_out/kernels_step_nopdl8.json 2>&1 | grep -v Warn | head -14 | cut -c1-200
echo "== config 3 (masks + sem-seg)"; timeout 900 python bench.py --workload ape_l_d_masks --no-cpu-baseline --no-microbench --steps 10 --warmup 3 2>gpurun_out/bench_masks8.err | tail -1 > gpurun_out/bench_masks8.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_masks8.json").read())
print("masks ms/step", d["ms_per_step"], "e2e", d["e2e"], "stages", d.get("stage_ms_eager_profile"))
PY
