#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== attn tests"; timeout 120 python -m pytest tests/test_attn_gpu.py -q -m gpu -x 2>&1 | tail -4
echo "== perf"; timeout 60 python tests/perf_attn.py 2>&1 | tail -5
