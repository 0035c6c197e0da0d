# One GPU visit: full GPU test-suite, N=1 bench with per-kernel breakdown, ncu launch list, ncu --set full of the top kernels.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 python scripts/bench_metric.py > gpurun_out/bench_metric.json 2> gpurun_out/bench_metric.err; cut -c1-300 gpurun_out/bench_metric.json
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 $BENCH > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 360 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# gemm_tc2_kernel (CTA pair) launches per block: mlp_2, mlp_3, mlp_4 -> the first two of block 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 0 -c 2 -f -o gpurun_out/prof_gemm $BENCH > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
# gemm_tc_kernel launches in one forward: proj(0) | per block: qkv, out, mlp_1, mlp_5 -> index 3 = mlp_1 of block 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_mlp1 $BENCH > gpurun_out/ncu_gemm_mlp1.log 2>&1
echo "ncu gemm mlp1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_attn_tc $BENCH > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"patchify_kernel|head_tail_kernel|head_slots_kernel|mlp_tail_kernel" -c 4 -f -o gpurun_out/prof_rowops $BENCH > gpurun_out/ncu_rowops.log 2>&1
echo "ncu rowops rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"map_" -c 8 -f -o gpurun_out/prof_metric python scripts/bench_metric.py > gpurun_out/ncu_metric.log 2>&1
echo "ncu metric rc=$?"
ls -la gpurun_out/ | head -30
