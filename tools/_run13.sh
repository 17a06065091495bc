mkdir -p gpurun_out/r2e
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e/launches.csv python tools/prof_step.py > gpurun_out/r2e/launch_run.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|chain_tc_kernel" -s 9 -c 9 -f -o gpurun_out/r2e/gemm python tools/prof_step.py > gpurun_out/r2e/gemm_run.log 2>&1; echo gemm rc=$?
ncu -i gpurun_out/r2e/gemm.ncu-rep --page raw --csv > gpurun_out/r2e/gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/r2e/gemm.ncu-rep --page details > gpurun_out/r2e/gemm_details.txt 2>/dev/null
rm -f gpurun_out/r2e/gemm.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"tx_fade_kernel|awgn_kernel|moments_partial_kernel|prep_kernel|head_kernel|bit_source" -s 6 -c 6 -f -o gpurun_out/r2e/hbm python tools/prof_step.py > gpurun_out/r2e/hbm_run.log 2>&1; echo hbm rc=$?
ncu -i gpurun_out/r2e/hbm.ncu-rep --page raw --csv > gpurun_out/r2e/hbm_raw.csv 2>/dev/null
ncu -i gpurun_out/r2e/hbm.ncu-rep --page details > gpurun_out/r2e/hbm_details.txt 2>/dev/null
rm -f gpurun_out/r2e/hbm.ncu-rep
ls -la gpurun_out/r2e
