# r02i: fp32 tensor-core mode parity, one-MUFU Mish A/B (A = two MUFU, B = one MUFU + Newton), faithful suite on B
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/r02i_model.log 2>&1; tail -3 $O/r02i_model.log
timeout 1500 python -m pytest tests/test_gpu_bf16_faithful.py -q -m gpu -s > $O/r02i_faithful.log 2>&1; tail -4 $O/r02i_faithful.log; grep "mish fast form" $O/r02i_faithful.log
REPS="1 2 3" EXTRA="--no-variants --breakdown" bash scripts/gpu_ab_lib.sh 2>&1 | tee $O/r02i_mish_ab.log
P=vision_transformer_detector_b200
for v in A B; do cp $P/libvitdet_b200_$v.so $P/libvitdet_b200.so; python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --breakdown 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); b=l['breakdown']; print('lib $v', {k: round(v['ms_per_step'],3) for k,v in b.items() if k.startswith('gemm_mlp') or k=='gemm_head'})" | tee -a $O/r02i_mish_ab.log; done
cp $P/libvitdet_b200_B.so $P/libvitdet_b200.so
timeout 600 python scripts/error_growth.py r02 > $O/r02i_error_growth.log 2>&1; tail -30 $O/r02i_error_growth.log
