# Runs the GPU parity tests one group per process (a trapped kernel kills its CUDA context, so groups are isolated).
mkdir -p gpurun_out
log=gpurun_out/gpu_tests_split.log; : > $log
for k in "dense_matches and bf16" "epilogue" ; do
  echo "=== ops: $k" >> $log
  timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$k" -x 2>&1 | tail -25 >> $log
done
for k in "weight_table or api_surface or unset" "tiny_model" "configuration_knobs" "default_model_fp32" "default_model_bf16" "device_tensor" "golden or transform_predictions"; do
  echo "=== model: $k" >> $log
  timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "$k" -x 2>&1 | tail -30 >> $log
done
echo "=== smoke" >> $log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -15 >> $log
tail -200 $log
