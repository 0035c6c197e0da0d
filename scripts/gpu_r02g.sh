# r02g: software-pipelined attention kernel; recalibrated faithful suite
mkdir -p gpurun_out
O=gpurun_out
echo "== VITDET_ATTN=1" > $O/r02g_attn_tests.log
VITDET_ATTN=1 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" >> $O/r02g_attn_tests.log 2>&1; tail -1 $O/r02g_attn_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown"
run() { env $1 $2 timeout 300 $B $3 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$3 $1 $2', '%.3f ms  %.0f img/s  attn %.3f ms  clk %s'%(l['ms_per_step'], l['value'], l['breakdown']['attention']['ms_per_step'], l['clocks']['sm_mhz']))" | tee -a $O/r02g_ab.log; }
for rep in 1 2; do
  run VITDET_ATTN=4 VITDET_ATTN_POLY=0
  run VITDET_ATTN=1 VITDET_ATTN_POLY=0
  run VITDET_ATTN=1 VITDET_ATTN_POLY=1
  run VITDET_ATTN=1 VITDET_ATTN_POLY=2
done
for a in 4 1; do run VITDET_ATTN=$a VITDET_ATTN_POLY=0 "--variant hires"; run VITDET_ATTN=$a VITDET_ATTN_POLY=0 "--variant vitb"; done
timeout 1500 python -m pytest tests/test_gpu_bf16_faithful.py -q -m gpu -s > $O/r02g_faithful.log 2>&1; tail -5 $O/r02g_faithful.log
NB="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
VITDET_ATTN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_sw_kernel -s 1 -c 1 -f -o $O/r02g_attnsw $NB > $O/r02g_ncu_attnsw.log 2>&1
echo "ncu rc=$?"
