#!/bin/bash
# round-2 GPU run 12: state check after the library evictions (conv3x3 engine default, own GEMV / language LayerNorm): suite as the
# driver runs it, smoke, bench lines of configs 2 / 3, CUPTI table, conv A/B
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== whole suite in ONE process"; timeout 2400 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log | cut -c1-200
grep -E "^(==|  [a-z]|full-size|mask logits|sem_seg|proposal set|text tower)" gpurun_out/pytest.log | cut -c1-260 > gpurun_out/parity_report.txt
grep -E "^(FAILED|ERROR)" gpurun_out/pytest.log | cut -c1-200 | head -40
echo "== bench (full line)"; timeout 1200 python bench.py 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json | cut -c1-300
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read())
for k in ("ms_per_step", "e2e", "e2e_predictor", "roofline_gemm", "roofline_attention", "own_kernel_ms_per_step", "clocks"):
    print(k, json.dumps(d.get(k))[:400])
print("roofline", {k: d["roofline"][k] for k in ("frac", "launch_ms", "lsu_frac", "share_of_step")})
PY
echo "== bench conv3x3 library A/B"; APE_CONV3X3=0 timeout 600 python bench.py --no-cpu-baseline --no-microbench 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cudnn conv: ms/step', d['ms_per_step'])"
echo "== per-kernel profile (CUPTI under graph replay, PDL off)"; APE_PDL=0 timeout 300 python tests/profile_step.py --out gpurun_out/kernels_step_final.json 2>&1 | grep -v Warn | head -30 | cut -c1-180
echo "== config 3"; timeout 900 python bench.py --workload ape_l_d_masks --no-cpu-baseline --no-microbench --steps 10 --warmup 3 2>gpurun_out/bench_masks.err | tail -1 > gpurun_out/bench_masks.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_masks.json").read())
print("masks ms/step", d["ms_per_step"], "e2e", d["e2e"], "own", d.get("own_kernel_ms_per_step"))
PY
ls -la gpurun_out/kernels_step_final.json
