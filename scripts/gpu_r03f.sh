O=gpurun_out; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_store_rate experiments/microbench/tma_store_rate.cu -lcuda && timeout 120 /tmp/tma_store_rate 2>&1 | tee $O/r03f_tma_store.log
