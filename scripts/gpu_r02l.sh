mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2; do
for f in "--no-e2e" ""; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants --breakdown $f 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); b=l['breakdown']; print('[$f]', '%.3f ms %.0f img/s'%(l['ms_per_step'], l['value']), 'attn-roof %.4f'%l['roofline_attention']['avg_launch_ms'], {k: round(v['ms_per_step'],3) for k,v in b.items() if k in ('attention','gemm_mlp_1','gemm_mlp_2','gemm_qkv')}, l['clocks']['sm_mhz'], l['clocks']['power_w_max'])" | tee -a $O/r02l.log
done; done
nvidia-smi --query-gpu=power.limit,power.max_limit,power.default_limit,clocks.max.sm --format=csv | tee -a $O/r02l.log
