MURCL_DEBUG_EPI=8 timeout 100 python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from murcl_b200 import ops
M,N,K=131072,512,512
x=torch.randn(M,K,device='cuda').bfloat16(); w=(torch.randn(N,K,device='cuda')/K**0.5).bfloat16(); b=torch.randn(N,device='cuda')
for _ in range(3):
    ops.linear_fwd(x,w,b,ops.ACT_RELU)
torch.cuda.synchronize()
PY
