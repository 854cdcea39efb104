#!/bin/bash
# A/B of the side-stream overlap: projection head + loss chain beside the actor / window selection (forward), head
# weight-gradient GEMMs beside the start of the aggregators' backward.
cd /root/repo
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', d['ms_per_step'], 'ms/step, graph', d['config'].get('cuda_graph'))"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e"
MURCL_OVERLAP_HEADS=0 $B 2>gpurun_out/ov_a.err | show "no overlap"
MURCL_OVERLAP_HEADS=1 MURCL_OVERLAP_HEAD_WGRAD=0 $B 2>gpurun_out/ov_b.err | show "fwd head/loss chain on the side stream"
MURCL_OVERLAP_HEADS=1 MURCL_OVERLAP_HEAD_WGRAD=1 $B 2>gpurun_out/ov_c.err | show "+ head wgrads on the side stream"
MURCL_OVERLAP_HEADS=1 $B --stage 1 2>gpurun_out/ov_d.err | show "stage 1, overlap"
MURCL_OVERLAP_HEADS=0 $B --stage 1 2>gpurun_out/ov_e.err | show "stage 1, no overlap"
for f in gpurun_out/ov_*.err; do echo "== $f"; tail -n 4 $f; done
