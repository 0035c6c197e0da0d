# r03b: A/B of three builds (A = product: 4 patch rows in flight; B = 9; C = 17 + FMNMX3 row maximum in the attention kernel),
# the wide slot-projection kernel on the ViT-B-width variant, parity tests on C, and an ncu capture of the 28 -> 3584 layer.
O=gpurun_out; mkdir -p $O
P=vision_transformer_detector_b200
cp $P/libvitdet_b200.so /tmp/lib_keep.so
summ='
import json,sys; l=json.loads(sys.stdin.read()); b=l.get("breakdown",{})
print(sys.argv[1], "%.3f ms  %.0f img/s  clk %s | "%(l["ms_per_step"], l["value"], l["clocks"]["sm_mhz"]) + "  ".join("%s %.3f"%(k, v["ms_per_step"]) for k,v in sorted(b.items(), key=lambda kv:-kv[1]["ms_per_step"])))'
for rep in 1 2; do
for v in A B C; do
  cp $P/libvitdet_b200_$v.so $P/libvitdet_b200.so
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "lib $v rep$rep"
done; done 2>&1 | tee $O/r03b_ab.log
cp $P/libvitdet_b200_C.so $P/libvitdet_b200.so
for w in 0 1 0 1; do
  VITDET_SLOTS_WIDE=$w timeout 300 python bench.py --variant vitb --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "vitb wide=$w"
done 2>&1 | tee $O/r03b_vitb.log
timeout 300 python bench.py --variant hires --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "hires libC" | tee $O/r03b_hires.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bf16_faithful.py -q -x > $O/r03b_tests_C.log 2>&1; tail -5 $O/r03b_tests_C.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o $O/prof_gemm_mlp1 $BENCH > $O/ncu_gemm_mlp1.log 2>&1
echo "ncu rc=$?"
cp /tmp/lib_keep.so $P/libvitdet_b200.so
