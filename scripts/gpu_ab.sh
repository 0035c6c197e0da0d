# A/B of env switches inside one box: bash scripts/gpu_ab.sh "VAR=0" "VAR=1" ...
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "$@"; do
  env $cfg python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$cfg', 'rep$rep', '%.3f ms  %.0f img/s  clk %s'%(l['ms_per_step'], l['value'], l['clocks']['sm_mhz']))"
done; done
