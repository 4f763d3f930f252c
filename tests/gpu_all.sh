#!/bin/bash
# Everything that needs a B200, in the order used during development (run through gpurun):
#   bash tests/gpu_all.sh        smoke + GPU tests (one pytest process per file: a CUDA fault stays local) + bench + profile
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== gpu tests"; : > gpurun_out/pytest.log
for f in tests/test_*gpu*.py tests/test_abi.py; do
  timeout 1500 python -m pytest $f -q -m gpu -s -p no:cacheprovider >> gpurun_out/pytest.log 2>&1
  echo "$f: $(tail -1 gpurun_out/pytest.log | cut -c1-120)"
done
grep -E "^(==|  [a-z]|full-size|mask logits|sem_seg|proposal set|text tower)" gpurun_out/pytest.log | cut -c1-260 > gpurun_out/parity_report.txt
grep -E "^(FAILED|ERROR)" gpurun_out/pytest.log | cut -c1-200 | head -40
echo "== whole suite in ONE process (as the driver runs it)"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -3 | cut -c1-200
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
echo "== per-kernel profile (CUPTI under graph replay)"; timeout 300 python tests/profile_step.py --out gpurun_out/kernels_step.json 2>&1 | grep -v Warn | head -40
