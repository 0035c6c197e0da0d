# r03s: MLP tail kernel: residual row fetched at the start of the tile: parity + breakdown
O=gpurun_out; mkdir -p $O
summ='
import json,sys; l=json.loads(sys.stdin.read()); b=l.get("breakdown",{})
print(sys.argv[1], "%.3f ms  %.0f img/s  clk %s | "%(l["ms_per_step"], l["value"], l["clocks"]["sm_mhz"]) + "  ".join("%s %.3f"%(k, v["ms_per_step"]) for k,v in sorted(b.items(), key=lambda kv:-kv[1]["ms_per_step"])))'
for r in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "default"
done | tee $O/r03s_default.log
timeout 900 python -m pytest tests/test_gpu_bf16_faithful.py tests/test_gpu_model.py -q -x > $O/r03s_tests.log 2>&1; tail -3 $O/r03s_tests.log
