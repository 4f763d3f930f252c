#!/bin/bash
# round-2 GPU run 7: new kernels' tests, CUPTI table without PDL overlap, whole-step A/B of the selectable variants, isolated
# GEMM / attention timings, config 3 / config 4 workloads
set -u
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_mask_post_gpu.py tests/test_panoptic_gpu.py tests/test_model_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest7.log 2>&1; tail -3 gpurun_out/pytest7.log | cut -c1-200
grep -E "^(FAILED|ERROR)|^  [a-z]" gpurun_out/pytest7.log | cut -c1-220 | head -40
echo "== CUPTI per geometry, PDL off"; APE_PDL=0 timeout 400 python tests/profile_step.py --out gpurun_out/kernels_step_nopdl.json 2>&1 | grep -v Warn | head -30 | cut -c1-200
ab() {  # name, env assignments...
  local name=$1; shift
  echo "== bench A/B: $name"
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-microbench --steps 20 --warmup 5 2>gpurun_out/ab_$name.err | tail -1 > gpurun_out/ab_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{n}.json").read())
    print(n, "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "own", d.get("own_kernel_ms_per_step"), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print(n, "FAILED", e)
PY
}
ab base APE_NOP=1
ab nopdl APE_PDL=0
ab gemm_mc APE_GEMM_POLICY=mc
ab gemm_single APE_GEMM_POLICY=single
ab attn1 APE_ATTN_VARIANT=1
ab msda_pair APE_MSDA_PAIR=2048
ab conv3x3 APE_CONV3X3=1
echo "== attention: own variants vs SDPA"; timeout 300 python tests/perf_attn.py 2>&1 | tee gpurun_out/perf_attn.txt | cut -c1-200
echo "== gemm shapes (graph-timed, cold operands)"; timeout 900 python tests/perf_gemm2.py > gpurun_out/gemm_phases.jsonl 2>gpurun_out/gemm_phases.err; tail -2 gpurun_out/gemm_phases.err
python - <<'PY'
import json
for l in open('gpurun_out/gemm_phases.jsonl'):
    r = json.loads(l)
    c = r.get('cycles') or {}
    print(f"{r['what']:12s} {r['variant']:6s} warm {r.get('us_warm')} cold {r.get('us_cold')} TF {r.get('tflops_cold')} | "
          + " ".join(f"{k}={v}" for k, v in c.items()), r.get('error', ''))
PY
echo "== config 3 (masks + sem-seg)"; timeout 900 python bench.py --workload ape_l_d_masks --no-cpu-baseline --no-microbench --steps 10 --warmup 3 2>gpurun_out/bench_masks.err | tail -1 | tee gpurun_out/bench_masks.json | cut -c1-1500
tail -2 gpurun_out/bench_masks.err
echo "== config 4 (1536^2, 5000 phrases, batch 4)"; timeout 900 python bench.py --workload ape_l_d_1536_phrase --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/bench_phrase.err | tail -1 | tee gpurun_out/bench_phrase.json | cut -c1-1500
tail -2 gpurun_out/bench_phrase.err
