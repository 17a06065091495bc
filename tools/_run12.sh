timeout 300 python tools/chain_probe.py > gpurun_out/chain_probe8.log 2>&1; grep -E "burst|chain vs|p99" gpurun_out/chain_probe8.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2e.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu_r2e.log
