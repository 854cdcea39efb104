#!/bin/bash
# DRAM bytes of the forward encoder GEMMs / fused pooling with and without the serpentine row order, L2 state carried over
# between kernels (ncu --cache-control none; three metrics = one pass, no replay).
cd /root/repo
mkdir -p gpurun_out
P="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph --no-secondary"
for s in 0 1; do
  MURCL_SERPENTINE=$s timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --kernel-name-base demangled -k "regex:gemm_tc_kernel..int.256, .bool.0|attnpool_fwd_kernel|pack_gather" -s 60 -c 66 --csv --log-file gpurun_out/r2_serp$s.csv $P > gpurun_out/ncu_serp$s.log 2>&1
  wc -l gpurun_out/r2_serp$s.csv
done
