#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (msda)"; timeout 900 python -m pytest tests/test_msda_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "== sweep quick"; timeout 600 python tests/perf_msda_sweep.py --quick > gpurun_out/msda_sweep_v2.jsonl 2> gpurun_out/msda_sweep_v2.err; tail -3 gpurun_out/msda_sweep_v2.err
python - <<'PY'
import json
for l in open('gpurun_out/msda_sweep_v2.jsonl'):
    r=json.loads(l); print(r['case'],r['dtype'],r['variant'],r['ms_median'],r['alg_GBps'],r['frac_hbm_peak'],r['max_abs_diff_vs_first'])
PY
