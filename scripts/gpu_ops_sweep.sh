mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt 2>&1
nproc >> gpurun_out/r1_smi.txt
for k in "patchify" "layernorm" "dense_matches and fp32" "dense_matches and bf16" "epilogue" "wide_range" "attention_matches and fp32" "attention_matches and bf16" "key_shift"; do
  echo "=== $k" >> gpurun_out/r1_ops.log
  timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$k" -x 2>&1 | tail -25 >> gpurun_out/r1_ops.log
done
tail -120 gpurun_out/r1_ops.log
