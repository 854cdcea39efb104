#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "accumulate" 2>&1 | tail -3
show() { python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); f=d['roofline']['families']; print('$1', d['ms_per_step'], 'ms/step; head_bwd_input ms', f['head_bwd_input']['ms'], f['head_bwd_input']['launches'])"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e"
MURCL_TAPE_ACCUM_DGRAD=0 $B 2>/dev/null | show "tape W_hh dgrad: plain launch + bf16 carry"
MURCL_TAPE_ACCUM_DGRAD=1 $B 2>/dev/null | show "tape W_hh dgrad: split-K atomics into the fp32 carry"
MURCL_TAPE_ACCUM_DGRAD=0 $B 2>/dev/null | show "plain again"
MURCL_TAPE_ACCUM_DGRAD=1 $B 2>/dev/null | show "accum again"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
