# N = 8 (or $N): the driver's launch line + the pure H2D microbench next to it
mkdir -p gpurun_out
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r03_bench_n$N.json 2> gpurun_out/r03_bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r03_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/h2d_microbench.py > gpurun_out/r03_h2d_n$N.json 2> gpurun_out/r03_h2d_n$N.err
cat gpurun_out/r03_h2d_n$N.json
python - <<PY
import json
l=json.loads(open('gpurun_out/r03_bench_n$N.json').read().strip().splitlines()[-1])
print('value',l['value'],'ms',l['ms_per_step'],'parity',l.get('multi_gpu_parity'),'launches',l['gpu_launches'], l['clocks'])
for k in ('e2e','e2e_sync','e2e_pageable','e2e_pageable_sync','e2e_uint8'):
    print(k, l[k]['value'] if l.get(k) else None)
print('fp32', l['fp32']['value'] if l.get('fp32') else None)
for k,v in (l.get('variants') or {}).items(): print(k, v['value'], v['ms_per_step'], v.get('e2e',{}).get('value'), v['dominant_kernel'])
PY
