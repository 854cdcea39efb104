#!/bin/bash
# ncu --set full capture of the HBM- / latency-bound kernels (packer, fused pooling fwd/bwd, segmented, NT-Xent) in one run.
# Run on the GPU box:  gpurun -- 'bash tools/prof_hbm.sh'; then here: ncu -i gpurun_out/r2_hbm.ncu-rep --page raw --csv
cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:pack_|seg_topk|seg_argmax|pool_bwd_scores|attn_score_bwd|ntx|attnpool|adam_step' -f -o gpurun_out/r2_hbm \
  python tools/prof_hbm_kernels.py > gpurun_out/prof_hbm.log 2>&1
ncu -i gpurun_out/r2_hbm.ncu-rep --page raw --csv > gpurun_out/r2_hbm_raw.csv 2>/dev/null
ls -la gpurun_out/r2_hbm.ncu-rep gpurun_out/r2_hbm_raw.csv
rm -f gpurun_out/r2_hbm.ncu-rep        # gpurun_out/ is capped at 64 MiB per call: the raw page is what tools/ncu_summary.py reads
tail -3 gpurun_out/prof_hbm.log
