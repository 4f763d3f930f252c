#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_nms_gpu.py tests/test_norm_rope_gpu.py tests/test_model_gpu.py -q -m "gpu and not slow" -x 2>&1 | tail -15
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r15.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['launch_ms'], d['roofline']['frac'], d['stage_ms']); print(json.dumps(d['own_kernel_ms_per_step']))"
python - <<'PY'
import json
d=json.load(open('gpurun_out/own_kernel_detail.json'))
for k,v in list(d.items())[:45]:
    name=k.split()
    extra=''
    if name[0]=='gemm_tn':
        M,N,K=map(int,name[1:4]); extra=f" {2*M*N*K/ (v[0]/v[1]*1e-3)/1e12:.0f} TF/s"
    print(f"{v[0]:8.3f} ms {v[1]:5.1f}x  {v[0]/v[1]*1e3:8.1f} us/call  {k}{extra}")
PY
