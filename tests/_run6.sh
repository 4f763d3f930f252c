#!/bin/bash
# round-2 GPU run 6 (after the container was re-created): the driver's own sequence (suite in one process, smoke, bench), then
# the evidence for profiles/: CUPTI per-kernel table, ncu launch list of bench.py, ncu --set full of the three top kernels
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== whole suite in ONE process"; timeout 2400 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log | cut -c1-200
grep -E "^(==|  [a-z]|full-size|mask logits|sem_seg|proposal set|text tower)" gpurun_out/pytest.log | cut -c1-260 > gpurun_out/parity_report.txt
grep -E "^(FAILED|ERROR)" gpurun_out/pytest.log | cut -c1-200 | head -40
echo "== bench (full line)"; timeout 1200 python bench.py 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json | cut -c1-600
tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-400
echo "== per-kernel profile (CUPTI under graph replay)"; timeout 300 python tests/profile_step.py --out gpurun_out/kernels_step.json 2>&1 | grep -v Warn | head -60
echo "== ncu launch list of bench.py"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-microbench > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
wc -l gpurun_out/launches_bench.csv
for t in gemm_qkv:gemm_tc attn_global:attn_fwd attn_window:attn_fwd msda_pair:msda_pair gemm_proj:gemm_tc; do
  name=${t%%:*}; k=${t##*:}
  echo "== ncu --set full $name"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_$name python tests/ncu_targets.py $name > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-160
done
ls -la gpurun_out | head -40
