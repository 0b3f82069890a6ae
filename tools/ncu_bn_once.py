"""One launch of every BN kernel on one large shape (for `ncu --set full`).  usage: ncu_bn_once.py [M C]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from convnet.pytorch_b200 import ops
M, C = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256 * 56 * 56, 256)
bf16 = torch.bfloat16
z = torch.randn(M, C, device='cuda').to(bf16); dy = torch.randn(M, C, device='cuda').to(bf16)
y = torch.randn(M, C, device='cuda').to(bf16); out = torch.empty_like(z); g = torch.empty_like(z)
gamma = torch.ones(C, device='cuda'); beta = torch.zeros(C, device='cuda')
rm = torch.zeros(C, device='cuda'); rv = torch.ones(C, device='cuda'); nbt = torch.zeros((), dtype=torch.int64, device='cuda')
mean, invstd, scale, shift = [torch.empty(C, device='cuda') for _ in range(4)]
sums = torch.empty(2 * C, device='cuda'); dg = torch.zeros(C, device='cuda'); db = torch.zeros(C, device='cuda')
ws = torch.zeros(ops.bn_workspace_floats(C), device='cuda')
ops.bn_stats(z, gamma, beta, 1e-5, 0.1, rm, rv, nbt, mean, invstd, scale, shift, ws)
ops.bn_apply(z, scale, shift, 1, out=out)
ops.bn_apply(z, scale, shift, 1, residual=y, out=out)
ops.bn_bwd_reduce(dy, None, z, 1, mean, invstd, gamma, beta, sums, dg, db, ws)
ops.bn_bwd_dx(dy, None, z, 1, mean, invstd, gamma, beta, sums, dz=out)
ops.bn_bwd_dx(dy, y, z, 1, mean, invstd, gamma, beta, sums, dz=out, g_out=g)
torch.cuda.synchronize()
