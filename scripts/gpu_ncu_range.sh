# usage: gpu_ncu_range.sh <kernel regex> <skip> <count> <out name>   (ncu --set full on a range of launches of bench.py)
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 $BENCH > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/$4 $BENCH > gpurun_out/ncu_$4.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$4.log | cut -c1-200
