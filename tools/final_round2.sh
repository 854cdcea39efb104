#!/bin/bash
# Final pass of the round on one B200: A/B of the two-wave weight-gradient plan, the whole GPU test suite, smoke(), the launch
# list of one eager step, a --set full capture of the weight-gradient kernel, and the default bench line.
cd /root/repo
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); f=d['roofline']['families']; print('$1', d['ms_per_step'], 'ms/step; wgrad ms', f['linear_bwd_weight']['ms'])"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e"
MURCL_WGRAD_WAVES=1 $B 2>/dev/null | show "wgrad one wave (18 splits)"
MURCL_WGRAD_WAVES=2 $B 2>/dev/null | show "wgrad two waves (37 splits)"
MURCL_WGRAD_WAVES=1 $B 2>/dev/null | show "wgrad one wave again"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -3
P="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph --no-secondary"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_all.csv $P > gpurun_out/ncu_r2_launches.log 2>&1
wc -l gpurun_out/r2_launches_all.csv
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc_kernel..int.256, .bool.1, .bool.1, .int.2, float, .int.[12], .bool.0." -s 60 -c 4 -f -o gpurun_out/r2_full_wgrad $P > gpurun_out/ncu_r2_full_wgrad.log 2>&1
ncu -i gpurun_out/r2_full_wgrad.ncu-rep --page raw --csv > gpurun_out/r2_full_wgrad_raw.csv 2>/dev/null
rm -f gpurun_out/r2_full_wgrad.ncu-rep
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python tools/show_bench.py gpurun_out/r2_bench_final.json 2>/dev/null | head -3
du -sh gpurun_out
