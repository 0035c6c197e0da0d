# r02e: ping-pong attention kernel
mkdir -p gpurun_out
O=gpurun_out
echo "== VITDET_ATTN=2" > $O/r02e_attn_tests.log
VITDET_ATTN=2 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" >> $O/r02e_attn_tests.log 2>&1; tail -1 $O/r02e_attn_tests.log
VITDET_ATTN=2 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "full_size or default_model or variant" >> $O/r02e_attn_tests.log 2>&1; tail -1 $O/r02e_attn_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown"
run() { env $1 $2 timeout 300 $B $3 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$3 $1 $2', '%.3f ms  %.0f img/s  attn %.3f ms  clk %s'%(l['ms_per_step'], l['value'], l['breakdown']['attention']['ms_per_step'], l['clocks']['sm_mhz']))" | tee -a $O/r02e_ab.log; }
for rep in 1 2; do
  run VITDET_ATTN=4 VITDET_ATTN_POLY=0
  run VITDET_ATTN=2 VITDET_ATTN_POLY=0
  run VITDET_ATTN=2 VITDET_ATTN_POLY=1
done
for a in 4 2; do run VITDET_ATTN=$a VITDET_ATTN_POLY=0 "--variant hires"; run VITDET_ATTN=$a VITDET_ATTN_POLY=0 "--variant vitb"; done
NB="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
VITDET_ATTN=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_pp_kernel -s 1 -c 1 -f -o $O/r02e_attnpp $NB > $O/r02e_ncu_attnpp.log 2>&1
echo "ncu rc=$?"
