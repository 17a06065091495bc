for v in rb2 rb3; do
  L=$PWD/dl_ofdm_b200/libdccn_$v.so
  echo "== $v chain=0"; DCCN_LIB=$L B=65536 CHUNKS=0,0 SEED=5 DCCN_CHAIN=0 timeout 200 python tools/chunk_diff.py 2>&1 | tail -4 | cut -d' ' -f2-9
done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2f.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu_r2f.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo bench rc=$?; cut -c1-200 gpurun_out/bench_r2f.json
