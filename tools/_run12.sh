for v in fix1 r152; do
  L=$PWD/dl_ofdm_b200/libdccn_$v.so
  echo "== $v chain=0"; DCCN_LIB=$L B=65536 CHUNKS=0,0 SEED=5 DCCN_CHAIN=0 timeout 200 python tools/chunk_diff.py 2>&1 | tail -4 | cut -d' ' -f2-9
done
echo "== product"; B=65536 CHUNKS=0,0 SEED=5 timeout 200 python tools/chunk_diff.py 2>&1 | tail -4 | cut -d' ' -f2-9
timeout 1400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2d.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu_r2d.log
timeout 300 python tools/chain_probe.py time > gpurun_out/chain_probe7.log 2>&1; grep burst gpurun_out/chain_probe7.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo bench rc=$?; cut -c1-400 gpurun_out/bench_r2d.json
