# r02h: attention: three CTAs per SM with 64-key tiles; software-pipelined kernel with the QK-first MMA order
mkdir -p gpurun_out
O=gpurun_out
for v in 3 1; do
echo "== VITDET_ATTN=$v" >> $O/r02h_attn_tests.log
VITDET_ATTN=$v timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bf16_faithful.py -x -q -m gpu -k "attention" >> $O/r02h_attn_tests.log 2>&1; tail -1 $O/r02h_attn_tests.log
done
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown"
run() { env $1 $2 timeout 300 $B $3 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$3 $1 $2', '%.3f ms  %.0f img/s  attn %.3f ms  clk %s'%(l['ms_per_step'], l['value'], l['breakdown']['attention']['ms_per_step'], l['clocks']['sm_mhz']))" | tee -a $O/r02h_ab.log; }
for rep in 1 2; do
  run VITDET_ATTN=4 VITDET_ATTN_POLY=0
  run VITDET_ATTN=3 VITDET_ATTN_POLY=0
  run VITDET_ATTN=1 VITDET_ATTN_POLY=0
done
for a in 4 3; do run VITDET_ATTN=$a VITDET_ATTN_POLY=0 "--variant hires"; run VITDET_ATTN=$a VITDET_ATTN_POLY=0 "--variant vitb"; done
NB="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
VITDET_ATTN=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc3_kernel -s 1 -c 1 -f -o $O/r02h_attn3 $NB > $O/r02h_ncu_attn3.log 2>&1
echo "ncu rc=$?"
# fp32-accumulate mode on the tensor cores
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "dense" > $O/r02h_fp32_ops.log 2>&1; tail -3 $O/r02h_fp32_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "tiny_model or knobs or fp32 or default_model" > $O/r02h_fp32_model.log 2>&1; tail -3 $O/r02h_fp32_model.log
timeout 600 python bench.py --mode fp32 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --breakdown > $O/r02h_bench_fp32.json 2> $O/r02h_bench_fp32.err; tail -25 $O/r02h_bench_fp32.err; cut -c1-300 $O/r02h_bench_fp32.json
VITDET_FP32=simt timeout 600 python bench.py --mode fp32 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --breakdown > $O/r02h_bench_fp32_simt.json 2> $O/r02h_bench_fp32_simt.err; tail -25 $O/r02h_bench_fp32_simt.err
