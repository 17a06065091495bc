mkdir -p gpurun_out/r2
timeout 300 python tools/r2_probe2.py > gpurun_out/r2/probe3.log 2>&1
tail -4 gpurun_out/r2/probe3.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_forms.py -m gpu -x -q > gpurun_out/r2/gputest_h.log 2>&1
tail -5 gpurun_out/r2/gputest_h.log
