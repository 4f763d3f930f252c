#!/bin/bash
# round-2 GPU run 9: GEMM epilogue prefetch / resident-operand A/B per shape, ncu source view of the K = 256 GEMM
set -u
mkdir -p gpurun_out
for cfg in "p0r0:APE_GEMM_PREFETCH=0 APE_GEMM_RESIDENT=0" "p1r0:APE_GEMM_PREFETCH=1 APE_GEMM_RESIDENT=0" "p3r0:APE_GEMM_PREFETCH=3 APE_GEMM_RESIDENT=0" "p1r3:APE_GEMM_PREFETCH=1 APE_GEMM_RESIDENT=3"; do
  name=${cfg%%:*}; envv=${cfg##*:}
  echo "== gemm shapes $name"
  env $envv PERF_GEMM_VARIANTS=single timeout 600 python tests/perf_gemm2.py > gpurun_out/gemm9_$name.jsonl 2>gpurun_out/gemm9_$name.err
  python - $name <<'PY'
import json, sys
for l in open(f'gpurun_out/gemm9_{sys.argv[1]}.jsonl'):
    r = json.loads(l)
    c = r.get('cycles') or {}
    print(f"{r['what']:12s} warm {r.get('us_warm')} cold {r.get('us_cold')} | " + " ".join(f"{k}={v}" for k, v in c.items() if k in ('load','issue','tail','total')), r.get('error', ''))
PY
done
echo "== ncu ffn1"
APE_GEMM_PREFETCH=1 APE_GEMM_RESIDENT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/r02_gemm_ffn1 python tests/ncu_targets.py gemm_ffn1 > gpurun_out/ncu_gemm_ffn1.log 2>&1; tail -1 gpurun_out/ncu_gemm_ffn1.log | cut -c1-160
