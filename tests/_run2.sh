bash tests/gpu_all.sh
echo "== msda pair sweep"; timeout 600 python tests/perf_msda_pair.py > gpurun_out/msda_pair_sweep.txt 2>&1; grep -E "generic|pairing" gpurun_out/msda_pair_sweep.txt; grep "1024 float16" gpurun_out/msda_pair_sweep.txt | sort -k 12 -n | head -8
echo "== bench with APE_GEMM_POLICY=mc"; APE_GEMM_POLICY=mc timeout 600 python bench.py --no-cpu-baseline --no-microbench 2>/dev/null | tail -1 | tee gpurun_out/bench_mc.json | cut -c1-200
python - <<'PY'
import json
for f in ("gpurun_out/bench.json","gpurun_out/bench_mc.json"):
    try:
        b=json.loads(open(f).read()); print(f, b["value"], b["ms_per_step"], b.get("roofline_gemm",{}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
