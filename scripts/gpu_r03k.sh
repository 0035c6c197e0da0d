# r03k: a fraction of the Mish epilogue's exponentials on the FMA pipe (A = none, B = 1 of 4, C = 1 of 2): breakdown A/B + parity on C
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
done; done 2>&1 | tee $O/r03k_ab.log
cp $P/libvitdet_b200_C.so $P/libvitdet_b200.so
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bf16_faithful.py -q -x > $O/r03k_tests_C.log 2>&1; tail -5 $O/r03k_tests_C.log
cp /tmp/lib_keep.so $P/libvitdet_b200.so
