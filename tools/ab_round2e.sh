#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show() { python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', d['ms_per_step'], 'ms/step,', d['gpu_launches']/d['steps'], 'lib launches/step')"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e"
MURCL_TAPE_DECODER=0 $B 2>gpurun_out/e_a.err | show "decoder per patch-step (autograd)"
MURCL_TAPE_DECODER=1 $B 2>gpurun_out/e_b.err | show "decoder on the tape (one batched backward)"
MURCL_TAPE_DECODER=0 $B 2>/dev/null | show "per patch-step again"
MURCL_TAPE_DECODER=1 $B 2>/dev/null | show "tape again"
tail -n 3 gpurun_out/e_b.err
