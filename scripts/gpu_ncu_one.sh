# usage: gpu_ncu_one.sh <kernel regex> <skip> <out name>
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 $BENCH > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/$3 $BENCH > gpurun_out/ncu_$3.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$3.log | cut -c1-200
