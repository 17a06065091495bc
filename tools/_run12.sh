DCCN_LIB=$PWD/dl_ofdm_b200/libdccn_chtrace.so timeout 200 python tools/chain_trace.py > gpurun_out/chain_trace5.log 2>&1; echo rc=$?
timeout 300 python tools/chain_probe.py > gpurun_out/chain_probe5.log 2>&1; echo rc=$?
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "equalizer or config3 or full_size or subgraph" > gpurun_out/pytest_chain5.log 2>&1; echo rc=$?; tail -3 gpurun_out/pytest_chain5.log
grep -E "burst|chain vs|p99" gpurun_out/chain_probe5.log
head -24 gpurun_out/chain_trace5.log; grep -A 26 "tail chain" gpurun_out/chain_trace5.log
