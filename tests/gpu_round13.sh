#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== profile step (graphs)"; timeout 600 python tests/profile_step.py --out gpurun_out/kernels_step_graphs.json 2>&1 | tail -75
echo "== all gpu tests"; timeout 1200 python -m pytest tests -q -m "gpu and not slow" -x 2>&1 | tail -6
