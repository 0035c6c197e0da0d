# Final verification of a tree: full GPU suite, smoke(), default bench, reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-600 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2>gpurun_out/bench_ref_n1.err; cut -c1-300 gpurun_out/bench_ref_n1.json
