# r02j: fp32-accumulate attention on the tensor cores (attention_tcs.cu)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/r02j_ops.log 2>&1; tail -3 $O/r02j_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/r02j_model.log 2>&1; tail -3 $O/r02j_model.log
timeout 600 python bench.py --mode fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --breakdown > $O/r02j_bench_fp32.json 2> $O/r02j_bench_fp32.err; tail -22 $O/r02j_bench_fp32.err; cut -c1-200 $O/r02j_bench_fp32.json
timeout 600 python scripts/error_growth.py r02j > $O/r02j_error_growth.log 2>&1; grep -A14 "weights: spread" $O/r02j_error_growth.log
