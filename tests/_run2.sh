bash tests/gpu_all.sh
echo "== msda pair sweep"; timeout 600 python tests/perf_msda_pair.py > gpurun_out/msda_pair_sweep.txt 2>&1; grep -E "generic|pairing" gpurun_out/msda_pair_sweep.txt; grep "1024 float16" gpurun_out/msda_pair_sweep.txt | grep pair | sort -t' ' -k 12 -n | awk '{print}' | sort -k12 -n | head -8
echo "== A/B: APE_GEMM_POLICY=mc | APE_PDL=0 | no LN fold"
APE_GEMM_POLICY=mc timeout 600 python bench.py --no-cpu-baseline --no-microbench 2>/dev/null | tail -1 > gpurun_out/bench_mc.json
APE_PDL=0 timeout 600 python bench.py --no-cpu-baseline --no-microbench 2>/dev/null | tail -1 > gpurun_out/bench_nopdl.json
python - <<'PY'
import json
for f in ("gpurun_out/bench.json","gpurun_out/bench_mc.json","gpurun_out/bench_nopdl.json"):
    try:
        b=json.loads(open(f).read()); print(f, round(b["value"],2), "img/s", round(b["ms_per_step"],3), "ms; gemm ms", round(b.get("roofline_gemm",{}).get("ms_per_step",0),3), "msda ms", round(b["roofline"]["launch_ms"],4))
    except Exception as e: print(f, "ERR", e)
PY
APE_ATTN_VARIANT=1 timeout 600 python bench.py --no-cpu-baseline --no-microbench 2>/dev/null | tail -1 > gpurun_out/bench_attn5.json
python - <<'PY'
import json
try:
    b=json.loads(open("gpurun_out/bench_attn5.json").read()); print("APE_ATTN_VARIANT=1:", round(b["value"],2), "img/s", round(b["ms_per_step"],3), "ms; attention ms", round(b.get("roofline_attention",{}).get("ms_per_step",0),3))
except Exception as e: print("attn5 bench ERR", e)
PY
echo "== attention vs SDPA"; timeout 300 python tests/perf_attn.py 2>&1 | tail -4
echo "== config 4 (1536^2, 5000 phrases, batch 4), 2 steps"; timeout 900 python bench.py --workload ape_l_d_1536_phrase --steps 2 --warmup 1 2>gpurun_out/bench_phrase.err | tail -1 | tee gpurun_out/bench_phrase.json | cut -c1-700; tail -3 gpurun_out/bench_phrase.err
echo "== config 3 (masks + semantic), 5 steps"; timeout 900 python bench.py --workload ape_l_d_masks --steps 5 --warmup 3 --no-cpu-baseline --no-microbench 2>gpurun_out/bench_masks.err | tail -1 | tee gpurun_out/bench_masks.json | cut -c1-300; tail -3 gpurun_out/bench_masks.err
echo "== conv3x3 engine A/B"; APE_CONV3X3=1 timeout 600 python bench.py --no-cpu-baseline --no-microbench 2>/dev/null | tail -1 | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('conv3x3 engine:', round(b['value'],2), 'img/s', b['own_kernel_ms_per_step'].get('conv3x3'))"
echo "== ncu captures"; bash tests/_run3.sh
