# r02f: diagnostics (bf16-store epilogue vs oracle), new bench.py keys, host path
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python scripts/debug_store_epilogue.py > $O/r02f_debug_store.log 2>&1; tail -40 $O/r02f_debug_store.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02f_bench.json 2> $O/r02f_bench.err; tail -5 $O/r02f_bench.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms',l['ms_per_step'])
for k in ('e2e','e2e_sync','e2e_pageable','e2e_pageable_sync','e2e_uint8'):
    print(k, l[k]['value'] if l.get(k) else None)
print('fp32', l['fp32'])
for k,v in (l.get('variants') or {}).items(): print(k, v['value'], v['ms_per_step'], v.get('e2e',{}).get('value'), v['dominant_kernel'])
print('roofline', l['roofline']['frac'] if l['roofline'] else None, 'attn', l['roofline_attention'])
PY
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_metric.py -x -q -m gpu > $O/r02f_model_tests.log 2>&1; tail -3 $O/r02f_model_tests.log
