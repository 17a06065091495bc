mkdir -p gpurun_out/r2
timeout 300 python tools/r2_probe2.py > gpurun_out/r2/probe2.log 2>&1
for slot in 3 8 11 6; do
  DCCN_TRACE_SLOT=$slot B=65536 N=60 timeout 200 python tools/trace_gemm.py > gpurun_out/r2/trace_slot$slot.log 2>&1
done
tail -5 gpurun_out/r2/probe2.log
