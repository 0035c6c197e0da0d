# r03m: head GEMMs (M = 1088) on the single-CTA kernel instead of CTA pairs (A/B by env), wide LayerNorm with 32 resident warps
O=gpurun_out; mkdir -p $O
summ='
import json,sys; l=json.loads(sys.stdin.read()); b=l.get("breakdown",{})
print(sys.argv[1], "%.3f ms  %.0f img/s  clk %s | "%(l["ms_per_step"], l["value"], l["clocks"]["sm_mhz"]) + "  ".join("%s %.3f"%(k, v["ms_per_step"]) for k,v in sorted(b.items(), key=lambda kv:-kv[1]["ms_per_step"])))'
for r in 1 2; do for mm in 0 4096; do
  VITDET_PAIR_MIN_M=$mm timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "pair_min_m=$mm"
done; done 2>&1 | tee $O/r03m_ab.log
timeout 300 python bench.py --variant vitb --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "vitb" | tee -a $O/r03m_ab.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "layernorm or dense" > $O/r03m_tests.log 2>&1; tail -3 $O/r03m_tests.log
