O=gpurun_out; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vision_transformer_detector_b200/csrc -I include -o /tmp/gemm_sweep experiments/microbench/gemm_sweep.cu -L vision_transformer_detector_b200 -lvitdet_b200 -Xlinker -rpath -Xlinker $PWD/vision_transformer_detector_b200 && timeout 200 /tmp/gemm_sweep 2>&1 | tee $O/r03h_gemm_sweep.log
