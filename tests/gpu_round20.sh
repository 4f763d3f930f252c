#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== msda tests"; timeout 600 python -m pytest tests/test_msda_gpu.py -q -m gpu -x 2>&1 | tail -12
echo "== perf"; timeout 300 python tests/perf_msda_self.py 2>&1 | head -4
echo "== ncu both kernels (instruction counts)"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:msda -c 12 --csv --log-file gpurun_out/ncu_msda_r20.csv python tests/perf_msda_self.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_msda_r20.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
out={}
for r in rows[1:]:
    out.setdefault((r[iid], r[ik][:60]), {})[r[im]]=r[iv]
for k,v in list(out.items())[:6]: print(k, v)
PY
