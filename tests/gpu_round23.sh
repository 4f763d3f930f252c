#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== attn tests"; timeout 180 python -m pytest tests/test_attn_gpu.py -q -m gpu -x 2>&1 | tail -4
echo "== perf"; timeout 120 python tests/perf_attn.py 2>&1 | tail -5
echo "== ncu"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/attn_r23 python tests/perf_attn.py > /dev/null 2>&1; ls -la gpurun_out/attn_r23.ncu-rep
