mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" -x 2>&1 | tail -30
echo "=== rest"
timeout 1200 python -m pytest tests -x -q -m gpu -k "not attention" 2>&1 | tail -5
bash scripts/gpu_quick.sh 2>&1 | grep -v "^\.\.\." | tail -30
