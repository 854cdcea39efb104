#!/bin/bash
# L2-prefetch distance sweep of the tcgen05 GEMM producers (MURCL_GEMM_L2PF: tiles ahead, B-stationary fwd / dgrad;
# MURCL_GEMM_L2PF_KB: k-blocks ahead, streaming wgrad), standalone shapes and the whole step.
cd /root/repo
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d.get('roofline') or {}; f=r.get('families',{}); print('$1', d['ms_per_step'], 'ms/step; fwd/dgrad/wgrad TF', f.get('linear_fwd',{}).get('tflops'), f.get('linear_bwd_input',{}).get('tflops'), f.get('linear_bwd_weight',{}).get('tflops'))"; }
for pf in 0 1 2; do for kb in 0 8 16; do
  MURCL_GEMM_L2PF=$pf MURCL_GEMM_L2PF_KB=$kb python tools/bench_gemm_shapes.py 262144,512,512 2>&1 | sed "s/^/pf=$pf kb=$kb /"
done; done
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e"
MURCL_GEMM_L2PF=0 MURCL_GEMM_L2PF_KB=0 $B 2>gpurun_out/pf_a.err | show "step pf=0 kb=0"
MURCL_GEMM_L2PF=1 MURCL_GEMM_L2PF_KB=0 $B 2>gpurun_out/pf_b.err | show "step pf=1 kb=0"
MURCL_GEMM_L2PF=2 MURCL_GEMM_L2PF_KB=0 $B 2>gpurun_out/pf_c.err | show "step pf=2 kb=0"
MURCL_GEMM_L2PF=1 MURCL_GEMM_L2PF_KB=8 $B 2>gpurun_out/pf_d.err | show "step pf=1 kb=8"
MURCL_GEMM_L2PF=1 MURCL_GEMM_L2PF_KB=16 $B 2>gpurun_out/pf_e.err | show "step pf=1 kb=16"
