for d in 0 1 2 4 7; do echo "== MURCL_DEBUG_EPI=$d"; MURCL_DEBUG_EPI=$d timeout 100 python tools/bench_gemm.py bf16 2>&1 | sed -n 1,7p; done
