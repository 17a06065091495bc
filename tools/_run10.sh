mkdir -p gpurun_out/r2
B=65536 MASKS=0,1,2,3,4,8,16,32,36,5,7,12,28,44,47,63 SHOW=eq_dft,eq_dense3,eq_conv7x64_phaseeq,eq_dense5,rx_demod_gemm timeout 400 python tools/ablate_gemm.py > gpurun_out/r2/ablate_f16.log 2>&1
cat gpurun_out/r2/ablate_f16.log | tail -20
