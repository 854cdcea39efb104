#!/bin/bash
# A/B: split-K partial tiles added to dw with vector atomics (no reduce launch) and the fused pooling forward's L2 prefetch.
cd /root/repo
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); f=d['roofline']['families']; print('$1', d['ms_per_step'], 'ms/step; wgrad ms', f['linear_bwd_weight']['ms'], 'attnpool_fwd ms', f['attnpool_fwd']['ms'])"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e"
MURCL_WGRAD_ATOMIC=0 MURCL_ATTNPOOL_L2PF=0 $B 2>gpurun_out/c_a.err | show "A  base"
MURCL_WGRAD_ATOMIC=1 MURCL_ATTNPOOL_L2PF=0 $B 2>gpurun_out/c_b.err | show "B  wgrad atomics"
MURCL_WGRAD_ATOMIC=0 MURCL_ATTNPOOL_L2PF=1 $B 2>gpurun_out/c_c.err | show "C  attnpool L2 prefetch"
MURCL_WGRAD_ATOMIC=1 MURCL_ATTNPOOL_L2PF=1 $B 2>gpurun_out/c_d.err | show "D  both"
MURCL_WGRAD_ATOMIC=0 MURCL_ATTNPOOL_L2PF=0 $B 2>gpurun_out/c_e.err | show "A' base again"
for pf in 0 1; do MURCL_ATTNPOOL_L2PF=$pf python tools/bench_attnpool.py 2>&1 | tail -4 | sed "s/^/l2pf=$pf /"; done
for at in 0 1; do MURCL_WGRAD_ATOMIC=$at python tools/bench_gemm_shapes.py 262144,512,512 262144,128,512 2>&1 | sed "s/^/atomic=$at /" | cut -c1-200; done
MURCL_WGRAD_ATOMIC=1 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
