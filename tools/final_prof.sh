#!/bin/bash
# Profile evidence for profiles/ (round 2): (1) ncu launch list of ~one eager step, (2) --set full captures of the dominant
# GEMM kernels and of the fused pooling kernels, (3) --set full of the HBM- / latency-bound kernels (tools/prof_hbm.sh).
# Run on the GPU box:  gpurun -- 'bash tools/final_prof.sh'
cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph --no-secondary"
# exactly ONE eager step: bench.py brackets its roofline-probe step with the NVTX range "murcl_probe_step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "murcl_probe_step/" --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_r2_launches.log 2>&1
wc -l gpurun_out/r2_launches.csv
cap() {  # name, regex, skip, count
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/r2_full_$1 $B > gpurun_out/ncu_r2_full_$1.log 2>&1
  # the raw page travels back (gpurun_out/ is capped at 64 MiB per call); the report itself stays on the box
  ncu -i gpurun_out/r2_full_$1.ncu-rep --page raw --csv > gpurun_out/r2_full_$1_raw.csv 2>/dev/null
  ls -la gpurun_out/r2_full_$1.ncu-rep gpurun_out/r2_full_$1_raw.csv
  rm -f gpurun_out/r2_full_$1.ncu-rep
}
cap fwd   'gemm_tc_kernel..int.256, .bool.0, .bool.0, .int.0, __nv_bfloat16, .int.2, .bool.1.' 20 1
cap dgrad 'gemm_tc_kernel..int.256, .bool.0, .bool.1, .int.1, __nv_bfloat16, .int.2, .bool.1.' 20 2
cap wgrad 'gemm_tc_kernel..int.256, .bool.1, .bool.1, .int.2, float, .int.[12], .bool.0.' 60 4
bash tools/prof_hbm.sh
