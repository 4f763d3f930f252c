#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== attn tests (short timeout)"; timeout 180 python -m pytest tests/test_attn_gpu.py -q -m gpu -x 2>&1 | tail -15
echo "== perf"; timeout 120 python tests/perf_attn.py 2>&1 | tail -5
