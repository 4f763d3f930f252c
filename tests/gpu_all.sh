#!/bin/bash
# Everything that needs a B200, in the order used during development (run through gpurun):
#   bash tests/gpu_all.sh [quick]        tests + smoke + bench + per-kernel profile
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== gpu tests"; timeout 2400 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 > gpurun_out/pytest.log; tail -40 gpurun_out/pytest.log | cut -c1-300
grep -E "^(==|  [a-z]|full-size|mask logits|sem_seg|proposal set)" gpurun_out/pytest.log | cut -c1-260 > gpurun_out/parity_report.txt
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json | cut -c1-600
tail -5 gpurun_out/bench.err
echo "== per-kernel profile (CUPTI under graph replay)"; timeout 300 python tests/profile_step.py --out gpurun_out/kernels_step.json 2>&1 | grep -v Warn | head -45
