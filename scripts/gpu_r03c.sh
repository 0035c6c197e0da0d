# r03c: operator parity on the current tree, ViT-B-width breakdown (wide slot projection v2, 16-byte gamma/beta loads in the
# LayerNorm), encoder micro-batch sweep on the default config (does an L2-resident activation chunk pay?).
O=gpurun_out; mkdir -p $O
summ='
import json,sys; l=json.loads(sys.stdin.read()); b=l.get("breakdown",{})
print(sys.argv[1], "%.3f ms  %.0f img/s  clk %s | "%(l["ms_per_step"], l["value"], l["clocks"]["sm_mhz"]) + "  ".join("%s %.3f"%(k, v["ms_per_step"]) for k,v in sorted(b.items(), key=lambda kv:-kv[1]["ms_per_step"])))'
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bf16_faithful.py -q -x > $O/r03c_tests.log 2>&1; tail -5 $O/r03c_tests.log
for w in 0 1 0 1; do
  VITDET_SLOTS_WIDE=$w timeout 300 python bench.py --variant vitb --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "$summ" "vitb wide=$w"
done 2>&1 | tee $O/r03c_vitb.log
for c in 64 32 16 8 64 32; do
  timeout 300 python bench.py --chunk $c --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants 2>/dev/null | python -c "$summ" "chunk=$c"
done 2>&1 | tee $O/r03c_chunk.log
