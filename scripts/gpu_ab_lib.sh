# A/B of two builds inside one box: libvitdet_b200_A.so vs libvitdet_b200_B.so (copied over the product library in turn)
mkdir -p gpurun_out
P=vision_transformer_detector_b200
for rep in ${REPS:-1 2 3}; do
for v in A B; do
  cp $P/libvitdet_b200_$v.so $P/libvitdet_b200.so
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $EXTRA 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('lib $v', 'rep$rep', '%.3f ms  %.0f img/s  clk %s'%(l['ms_per_step'], l['value'], l['clocks']['sm_mhz']))"
done; done
cp $P/libvitdet_b200_B.so $P/libvitdet_b200.so
