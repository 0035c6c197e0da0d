O=gpurun_out; mkdir -p $O
summ='
import json,sys; l=json.loads(sys.stdin.read()); b=l.get("breakdown",{})
print(sys.argv[1], "%.3f ms  %.0f img/s  clk %s | "%(l["ms_per_step"], l["value"], l["clocks"]["sm_mhz"]) + "  ".join("%s %.3f"%(k, v["ms_per_step"]) for k,v in sorted(b.items(), key=lambda kv:-kv[1]["ms_per_step"])))'
timeout 300 python bench.py --mode fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "fp32" | tee $O/r03n_fp32.log
timeout 300 python bench.py --variant hires --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "hires" | tee -a $O/r03n_fp32.log
