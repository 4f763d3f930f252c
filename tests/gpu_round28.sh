#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== default bench (driver command)"; /usr/bin/time -v timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err | head -1; grep -E "Elapsed|Maximum resident" gpurun_out/bench_final.err; python -c "
import json
d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline'), d.get('detections_per_image'))"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 | cut -c1-400
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/launches_bench.csv
echo "== ncu full msda"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fused_fwd -s 20 -c 1 -f -o gpurun_out/msda_fused_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; ls -la gpurun_out/msda_fused_final.ncu-rep
