#!/bin/bash
# A/B of this session's changes on one B200: CTA-pair weight gradient, fused Adam, batched step draws, split-precision pairs.
# Run on the GPU box:  gpurun -- 'bash tools/ab_round2b.sh'
cd /root/repo
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d.get('roofline') or {}; f=r.get('families',{}); print('$1', d['ms_per_step'], 'ms/step,', d['gpu_launches']/d['steps'], 'lib launches/step; fwd/dgrad/wgrad TF', f.get('linear_fwd',{}).get('tflops'), f.get('linear_bwd_input',{}).get('tflops'), f.get('linear_bwd_weight',{}).get('tflops'))"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e"
MURCL_WGRAD_PAIR=0 MURCL_TORCH_ADAM=1 $B --rng reference 2>gpurun_out/ab_a.err | show "A  base (no pair, torch Adam, per-step rng)"
MURCL_WGRAD_PAIR=1 MURCL_TORCH_ADAM=1 $B --rng reference 2>gpurun_out/ab_b.err | show "B  + wgrad CTA pair"
MURCL_WGRAD_PAIR=1 MURCL_TORCH_ADAM=0 $B --rng reference 2>gpurun_out/ab_c.err | show "C  + fused Adam"
MURCL_WGRAD_PAIR=1 MURCL_TORCH_ADAM=0 $B --rng batched 2>gpurun_out/ab_d.err | show "D  + batched draws"
MURCL_WGRAD_PAIR=0 python tools/bench_gemm_shapes.py 262144,512,512 262144,128,512 2>&1 | sed 's/^/pair=0 /'
MURCL_WGRAD_PAIR=1 python tools/bench_gemm_shapes.py 262144,512,512 262144,128,512 2>&1 | sed 's/^/pair=1 /'
F="python bench.py --precision fp32 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --no-e2e --no-graph"
MURCL_SPLIT_PAIR=0 MURCL_WGRAD_PAIR=0 $F 2>gpurun_out/ab_f0.err | show "fp32 mode, one-CTA split GEMMs"
MURCL_SPLIT_PAIR=1 MURCL_WGRAD_PAIR=1 $F 2>gpurun_out/ab_f1.err | show "fp32 mode, CTA-pair split GEMMs"
tail -3 gpurun_out/ab_*.err
