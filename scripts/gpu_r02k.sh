# r02k: row-staged patchify (parity + A/B + ncu), set_weight without per-tensor sync, final-tree sanity
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q -m gpu > $O/r02k_tests.log 2>&1; tail -3 $O/r02k_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown"
run() { env $1 timeout 300 $B $2 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$2 $1', '%.3f ms  %.0f img/s  patchify %.4f ms  clk %s'%(l['ms_per_step'], l['value'], l['breakdown']['patchify']['ms_per_step'], l['clocks']['sm_mhz']))" | tee -a $O/r02k_ab.log; }
for rep in 1 2; do run VITDET_PATCHIFY=direct; run VITDET_PATCHIFY=staged; done
run VITDET_PATCHIFY=direct "--variant hires"; run VITDET_PATCHIFY=staged "--variant hires"
NB="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
timeout 600 ncu --set full --clock-control none -k regex:"patchify" -c 2 -f -o $O/prof_patchify $NB > $O/ncu_patchify.log 2>&1
echo "ncu rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02k_bench.json 2> $O/r02k_bench.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms',l['ms_per_step'], 'traffic', l['roofline']['traffic'], l['roofline'].get('traffic_note'))
for k in ('e2e','e2e_sync','e2e_pageable','e2e_pageable_sync','e2e_uint8'):
    print(k, round(l[k]['value'],1), round(l[k]['value']/l['value'],3))
print('fp32', l['fp32']['value']); print('attn', l['roofline_attention']['avg_launch_ms'], l['roofline_attention'].get('sfu'))
PY
