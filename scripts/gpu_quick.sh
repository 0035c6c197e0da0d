# Quick GPU visit: full GPU test-suite + N=1 bench with per-kernel breakdown (no ncu).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown ${BENCH_ARGS} > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/bench_n1.json'))
print({k:l[k] for k in ('value','ms_per_step','model_tflops','gpu_launches')}, l['e2e'], l['clocks'])
print(l['roofline']); print(l['cpu_baseline'])
for k,v in l.get('breakdown',{}).items(): print('  %-26s %8.3f ms  x%d'%(k,v['ms_per_step'],v['launches_per_step']))
PY
tail -3 gpurun_out/bench_n1.err
