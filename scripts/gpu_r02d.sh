# r02d: persistent split-row kernel and polynomial exp2 variants; recalibrated bf16-faithful suite
mkdir -p gpurun_out
O=gpurun_out
for v in "VITDET_ATTN=80" "VITDET_ATTN=80 VITDET_ATTN_POLY=2" "VITDET_ATTN=40 VITDET_ATTN_POLY=1" "VITDET_ATTN=40 VITDET_ATTN_POLY=4"; do
  echo "== $v" >> $O/r02d_attn_tests.log
  env $v timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" >> $O/r02d_attn_tests.log 2>&1; tail -1 $O/r02d_attn_tests.log
done
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --breakdown"
run() { env $1 $2 timeout 300 $B $3 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$3 $1 $2', '%.3f ms  %.0f img/s  attn %.3f ms  clk %s'%(l['ms_per_step'], l['value'], l['breakdown']['attention']['ms_per_step'], l['clocks']['sm_mhz']))" | tee -a $O/r02d_ab.log; }
for rep in 1 2; do
  run VITDET_ATTN=4 VITDET_ATTN_POLY=0
  for p in 0 1 2 3; do run VITDET_ATTN=40 VITDET_ATTN_POLY=$p; done
  for p in 0 1 2 3; do run VITDET_ATTN=80 VITDET_ATTN_POLY=$p; done
done
for a in 40 80; do for p in 0 2; do run VITDET_ATTN=$a VITDET_ATTN_POLY=$p "--variant hires"; run VITDET_ATTN=$a VITDET_ATTN_POLY=$p "--variant vitb"; done; done
timeout 1500 python -m pytest tests/test_gpu_bf16_faithful.py -q -m gpu -s > $O/r02d_faithful.log 2>&1; tail -5 $O/r02d_faithful.log
NB="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
VITDET_ATTN=80 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc8p_kernel -s 1 -c 1 -f -o $O/r02d_attn8p $NB > $O/r02d_ncu_attn8p.log 2>&1
echo "ncu rc=$?"
