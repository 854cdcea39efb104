#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py tests/test_gpu_attnpool.py -m gpu -x -q 2>&1 | tail -2
MURCL_SERPENTINE=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
show() { python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); f=d['roofline']['families']; print('$1', d['ms_per_step'], 'ms/step; fwd ms', f['linear_fwd']['ms'], 'attnpool_fwd ms', f['attnpool_fwd']['ms'])"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e"
MURCL_SERPENTINE=0 $B 2>/dev/null | show "same direction"
MURCL_SERPENTINE=1 $B 2>/dev/null | show "serpentine (fwd)"
MURCL_SERPENTINE=0 $B 2>/dev/null | show "same direction again"
MURCL_SERPENTINE=1 $B 2>/dev/null | show "serpentine again"
