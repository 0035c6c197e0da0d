# N = 2: the driver's launch line; checks the C-ABI gather, multi_gpu_parity and the new keys
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r03_bench_n$N.json 2> gpurun_out/r03_bench_n$N.err
echo "rc=$?"; tail -5 gpurun_out/r03_bench_n$N.err; python - <<PY
import json
l=json.loads(open('gpurun_out/r03_bench_n$N.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms',l['ms_per_step'],'parity',l.get('multi_gpu_parity'),'launches',l['gpu_launches'])
for k in ('e2e','e2e_sync','e2e_pageable','e2e_pageable_sync','e2e_uint8'):
    print(k, l[k]['value'] if l.get(k) else None)
print('fp32', l['fp32']['value'] if l.get('fp32') else None)
for k,v in (l.get('variants') or {}).items(): print(k, v['value'], v['ms_per_step'], v.get('e2e',{}).get('value'), v['dominant_kernel'])
PY
