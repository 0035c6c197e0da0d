# SM clock and board power under single kernels run back to back (evidence for "the step is power-limited").
O=gpurun_out; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vision_transformer_detector_b200/csrc -I include -o /tmp/power_probe experiments/microbench/power_probe.cu -L vision_transformer_detector_b200 -lvitdet_b200 -Xlinker -rpath -Xlinker $PWD/vision_transformer_detector_b200 || exit 1
: > $O/r03_power_probe.log
for k in mlp_2 mlp_3 mlp_1 qkv; do
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.sw_power_cap,temperature.gpu --format=csv,noheader,nounits -lms 200 > /tmp/smi_$k.csv &
  SMI=$!
  sleep 0.5
  timeout 60 /tmp/power_probe $k 4 >> $O/r03_power_probe.log
  kill $SMI; wait $SMI 2>/dev/null
  python - "$k" >> $O/r03_power_probe.log <<'PY'
import sys, statistics
rows = [l.strip().split(", ") for l in open("/tmp/smi_%s.csv" % sys.argv[1]) if l.strip()]
rows = rows[4:-1] or rows
clk = [float(r[0]) for r in rows]; pw = [float(r[1]) for r in rows]
cap = sum(1 for r in rows if r[2].strip().lower() in ("active", "1"))
print("   nvidia-smi under %s: SM clock median %.0f MHz (min %.0f, max %.0f), power median %.0f W (max %.0f), sw_power_cap active in %d of %d samples, %s C"
      % (sys.argv[1], statistics.median(clk), min(clk), max(clk), statistics.median(pw), max(pw), cap, len(rows), rows[-1][3]))
PY
done
cat $O/r03_power_probe.log
