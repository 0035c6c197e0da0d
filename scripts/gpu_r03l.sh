# r03l: heads wider than 64 (two boxes per head tile): parity (operator, bf16-faithful, model knobs) + default breakdown (NB = 1 unchanged?)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bf16_faithful.py tests/test_gpu_model.py -q -x > $O/r03l_tests.log 2>&1; tail -25 $O/r03l_tests.log
summ='
import json,sys; l=json.loads(sys.stdin.read()); b=l.get("breakdown",{})
print(sys.argv[1], "%.3f ms  %.0f img/s  clk %s | "%(l["ms_per_step"], l["value"], l["clocks"]["sm_mhz"]) + "  ".join("%s %.3f"%(k, v["ms_per_step"]) for k,v in sorted(b.items(), key=lambda kv:-kv[1]["ms_per_step"])))'
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "default" | tee $O/r03l_default.log
