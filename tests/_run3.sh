# ncu captures of the shipped kernels (one launch each, --set full) + the launch list of the bench command
set -u
mkdir -p gpurun_out
for t in attn_window attn_global gemm_qkv gemm_proj msda_pair xattn; do
  case $t in attn_*) k=attn_fwd_kernel;; gemm_*) k=gemm_tc_kernel;; msda_pair) k=msda_pair_fused;; xattn) k=attn_xfwd;; esac
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02_ncu_$t python tests/ncu_targets.py $t > gpurun_out/ncu_$t.log 2>&1
  echo "$t rc=$? $(ls -la gpurun_out/r02_ncu_$t.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-microbench --no-graphs > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$? $(wc -l < gpurun_out/r02_launches_bench.csv) lines"
