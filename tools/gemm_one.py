"""One K1a GEMM, a few launches of one engine — the target of `ncu -k regex:gemm_tc_kernel` captures.  usage: gemm_one.py MxKxN engine [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeprecommendation_b200 import ops

M, K, N = (int(v) for v in sys.argv[1].split('x'))
eng = sys.argv[2]
x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') / K ** 0.5; b = torch.randn(N, device='cuda')
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    y = ops.linear_raw(x, w, b, engine=eng)
torch.cuda.synchronize()
print(float(y.abs().max()))
