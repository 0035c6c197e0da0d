# r03p: ncu --set full of the ViT-B-width row kernels on the final tree (LayerNorm D = 768 after the occupancy change, wide slot projection)
O=gpurun_out; mkdir -p $O
rm -f $O/prof_*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layernorm_kernel|head_slots_wide_kernel" -s 1 -c 1 -f -o $O/prof_ln_vitb python bench.py --variant vitb --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > $O/ncu_ln.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none -k regex:"head_slots_wide_kernel" -c 1 -f -o $O/prof_slots_vitb python bench.py --variant vitb --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > $O/ncu_slots.log 2>&1; echo rc=$?
