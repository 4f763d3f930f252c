#!/bin/bash
# One gpurun call: parity tests, smoke, bench, sweep, ncu captures.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "== sweep" ; timeout 1200 python tests/perf_msda_sweep.py > gpurun_out/msda_sweep.jsonl 2> gpurun_out/msda_sweep.err; tail -3 gpurun_out/msda_sweep.err; wc -l gpurun_out/msda_sweep.jsonl
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_fwd_kernel -s 3 -c 2 -o gpurun_out/msda_enc_fp32 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
