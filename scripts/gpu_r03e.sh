# r03e: ncu --set full of the small kernels of the default step (QKV, output projection, mlp_1, mlp_5, mlp_4, MLP tail) and of the
# ViT-B-width LayerNorm, to look for fixable stalls.
O=gpurun_out; mkdir -p $O
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 4 -f -o $O/prof_small_tc $BENCH > $O/ncu_small_tc.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 2 -c 1 -f -o $O/prof_small_tc2 $BENCH > $O/ncu_small_tc2.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tail_kernel -s 1 -c 1 -f -o $O/prof_tail $BENCH > $O/ncu_tail.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 1 -c 1 -f -o $O/prof_ln_vitb python bench.py --variant vitb --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > $O/ncu_ln.log 2>&1; echo rc=$?
ls -la $O/*.ncu-rep
