mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "dense" -x 2>&1 | tail -25
echo "=== model"
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -8
bash scripts/gpu_quick.sh 2>&1 | grep -v "^\.\.\." | tail -28
