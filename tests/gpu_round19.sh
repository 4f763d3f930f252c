#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== vit test"; timeout 300 python -m pytest tests/test_model_gpu.py -q -m "gpu and not slow" -k "vit_engine or pyramid" 2>&1 | tail -25
echo "== ncu msda_self"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_self -c 1 -f -o gpurun_out/msda_self_r19 python tests/perf_msda_self.py > gpurun_out/ncu_msda_self.log 2>&1; tail -3 gpurun_out/ncu_msda_self.log
ls -la gpurun_out/*.ncu-rep
