#!/bin/bash
# Everything that needs a B200, as run at the end of round 2 (through gpurun: `gpurun --timeout 2700 -- 'bash tests/gpu_all.sh'`):
# the driver's sequence (smoke, suite in one process, bench, reference arm) + configs 3 / 4, CUPTI table, predictor timings,
# ncu --set full of the attention / GEMM kernels.  Outputs land in gpurun_out/; what is kept goes to profiles/ (see its README).
# A/B switches for whole-step comparisons (python bench.py --no-cpu-baseline --no-microbench under each): APE_PDL, APE_GEMM_POLICY,
# APE_GEMM_LEAN, APE_ATTN_VARIANT, APE_MSDA_PAIR, APE_CONV3X3, APE_CONV_PAIR, APE_FUSED_ROPE.
# Multi-GPU: gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
#   --master-port 29517 bench.py --gpus N'.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== whole suite in ONE process"; timeout 2400 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log | cut -c1-200
grep -E "^(==|  [a-z]|full-size|mask logits|sem_seg|proposal set|text tower)" gpurun_out/pytest.log | cut -c1-260 > gpurun_out/parity_report.txt
grep -E "^(FAILED|ERROR)" gpurun_out/pytest.log | cut -c1-200 | head -40
echo "== bench (full line)"; timeout 1200 python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read())
for k in ("value", "ms_per_step", "e2e", "e2e_predictor", "gpu_launches", "own_kernel_ms_per_step", "clocks"):
    print(k, json.dumps(d.get(k))[:300])
print("roofline", {k: d["roofline"][k] for k in ("frac", "launch_ms", "lsu_frac", "share_of_step", "decoder_launch_ms")})
print("gemm", {k: d["roofline_gemm"][k] for k in ("frac", "achieved", "ms_per_step")}, "attn", {k: d["roofline_attention"][k] for k in ("frac", "achieved", "ms_per_step")})
print("cpu_baseline", d.get("cpu_baseline"))
PY
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-300
echo "== per-kernel profile (CUPTI under graph replay, PDL off)"; APE_PDL=0 timeout 300 python tests/profile_step.py --out gpurun_out/kernels_step_final.json 2>&1 | grep -v Warn | head -16 | cut -c1-180
echo "== config 3"; timeout 900 python bench.py --workload ape_l_d_masks --no-cpu-baseline --no-microbench --steps 10 --warmup 3 2>gpurun_out/bench_masks.err | tail -1 > gpurun_out/bench_masks.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_masks.json").read())
print("masks ms/step", d["ms_per_step"], "e2e", d["e2e"])
PY
echo "== config 4"; timeout 900 python bench.py --workload ape_l_d_1536_phrase --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/bench_phrase.err | tail -1 > gpurun_out/bench_phrase.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_phrase.json").read())
print("phrase ms/step", d["ms_per_step"], "value", d["value"], "roofline", d["roofline"]["frac"], "own", d.get("own_kernel_ms_per_step"))
PY
echo "== predictor"; timeout 300 python tests/perf_predictor.py 2>&1 | tail -11 | tee gpurun_out/perf_predictor.txt
echo "== attention"; timeout 300 python tests/perf_attn.py 2>&1 | tee gpurun_out/perf_attn.txt | cut -c1-160
for t in attn_global:attn_fwd gemm_qkv:gemm_ gemm_ffn1:gemm_; do
  name=${t%%:*}; k=${t##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02f_$name python tests/ncu_targets.py $name > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-120
done
