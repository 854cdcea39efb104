#!/bin/bash
# A/B of the step time: stage 3 (actor-chosen windows) vs stage 1 (random windows), and with the head tape disabled.
cd /root/repo
show() { python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', d['ms_per_step'], 'ms/step,', d['gpu_launches']/d['steps'], 'library launches/step')"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e"
$B --stage 3 2>/dev/null | show "stage 3"
$B --stage 1 2>/dev/null | show "stage 1"
MURCL_DISABLE_HEADTAPE=1 $B --stage 3 2>/dev/null | show "stage 3, no head tape"
$B --stage 3 2>/dev/null | show "stage 3 (repeat)"
