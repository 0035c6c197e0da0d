# r03a: (1) MUFU-rate micro-benchmark of the attention exponential loop; (2) in-box A/B of three library builds
# (A = patch rows one at a time, B = 4 patch rows in flight per lane, C = 6 rows + one shared reciprocal per Mish pair)
# by per-category times (bench.py --breakdown); (3) operator / bf16-faithful parity tests on build C.
O=gpurun_out; mkdir -p $O
P=vision_transformer_detector_b200
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_rate experiments/microbench/mufu_rate.cu && timeout 120 /tmp/mufu_rate > $O/r03a_mufu.log 2>&1
cat $O/r03a_mufu.log
cp $P/libvitdet_b200.so /tmp/lib_keep.so
for rep in 1 2; do
for v in A B C; do
  cp $P/libvitdet_b200_$v.so $P/libvitdet_b200.so
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); b=l.get('breakdown',{})
print('lib $v rep$rep  %.3f ms  %.0f img/s  clk %s | '%(l['ms_per_step'], l['value'], l['clocks']['sm_mhz']) + '  '.join('%s %.3f'%(k, v['ms_per_step']) for k,v in sorted(b.items(), key=lambda kv:-kv[1]['ms_per_step'])))"
done; done 2>&1 | tee $O/r03a_ab.log
cp $P/libvitdet_b200_C.so $P/libvitdet_b200.so
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bf16_faithful.py -q -x > $O/r03a_tests_C.log 2>&1; tail -5 $O/r03a_tests_C.log
cp /tmp/lib_keep.so $P/libvitdet_b200.so
