# r02b: attention A/B after the ring split (two CTAs per SM again)
mkdir -p gpurun_out
O=gpurun_out
echo "== VITDET_ATTN=8" > $O/r02b_attn_tests.log
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" >> $O/r02b_attn_tests.log 2>&1; tail -2 $O/r02b_attn_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --breakdown"
for rep in 1 2; do
for v in "VITDET_ATTN=4" "VITDET_ATTN=8"; do
  env $v timeout 300 $B 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$v rep$rep', '%.3f ms  %.0f img/s  attn %.3f ms  clk %s'%(l['ms_per_step'], l['value'], l['breakdown']['attention']['ms_per_step'], l['clocks']['sm_mhz']))" | tee -a $O/r02b_ab.log
done; done
for v in "VITDET_ATTN=4" "VITDET_ATTN=8"; do
  env $v timeout 300 $B --variant hires 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('hires $v', '%.3f ms  %.0f img/s  attn %.3f ms  clk %s'%(l['ms_per_step'], l['value'], l['breakdown']['attention']['ms_per_step'], l['clocks']['sm_mhz']))" | tee -a $O/r02b_ab.log
  env $v timeout 300 $B --variant vitb 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('vitb $v', '%.3f ms  %.0f img/s  attn %.3f ms  clk %s'%(l['ms_per_step'], l['value'], l['breakdown']['attention']['ms_per_step'], l['clocks']['sm_mhz']))" | tee -a $O/r02b_ab.log
done
NB="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc8_kernel -s 1 -c 1 -f -o $O/r02b_attn8 $NB > $O/r02b_ncu_attn8.log 2>&1
echo "ncu rc=$?"
