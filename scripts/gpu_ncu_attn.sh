mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 $BENCH > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn_tc $BENCH > gpurun_out/ncu_attn_tc.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_attn_tc.log
