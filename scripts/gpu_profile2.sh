# bench + source-level ncu of the first MLP GEMM and the attention kernel + variant benches
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/bench_n1.json'))
print({k:l[k] for k in ('value','ms_per_step','model_tflops')}, l['e2e']['value'])
for k,v in l.get('breakdown',{}).items(): print('  %-26s %8.3f ms  x%d'%(k,v['ms_per_step'],v['launches_per_step']))
PY
for v in hires vitb; do
timeout 900 python bench.py --variant $v --steps 5 --warmup 3 --breakdown --no-cpu-baseline --no-e2e > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
python - <<PY
import json
l=json.load(open('gpurun_out/bench_$v.json'))
print('$v', {k:l[k] for k in ('value','ms_per_step','model_tflops')})
for k,v in l.get('breakdown',{}).items(): print('  %-26s %8.3f ms  x%d'%(k,v['ms_per_step'],v['launches_per_step']))
PY
tail -2 gpurun_out/bench_$v.err
done
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_mlp1 $BENCH > gpurun_out/ncu_mlp1.log 2>&1
echo "ncu mlp1 rc=$?"
