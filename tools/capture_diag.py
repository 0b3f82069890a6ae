"""Which forward/loss/backward formulations survive CUDA-graph capture?  usage: capture_diag.py <A|B|C|D>
A: Runtime.train_step (no autograd)            B: autograd + fused CE + backward(grad_tensors=[device scalar])
C: autograd + fused CE + loss.backward()       D: autograd + torch F.cross_entropy + backward(grad_tensors=[...])"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from convnet.pytorch_b200 import models
from convnet.pytorch_b200.engine import convert_b200
from convnet.pytorch_b200.utils.cross_entropy import CrossEntropyLoss

v = sys.argv[1]
torch.manual_seed(0)
m = convert_b200(models.resnet(dataset='cifar10', depth=20), 'cuda').train()
rt = m._b200
x = torch.randn(32, 3, 32, 32, device='cuda'); y = torch.randint(0, 10, (32,), device='cuda')
crit = CrossEntropyLoss()
up = torch.ones((), device='cuda')


def step():
    if v == 'A':
        return rt.train_step(x, y, 0.0, up)[1]
    out = m(x)
    if v == 'D':
        loss = F.cross_entropy(out, y)
    else:
        loss = crit(out, y)
    if v == 'C':
        loss.backward()
    else:
        torch.autograd.backward(loss, grad_tensors=[up])
    return loss.detach()


for _ in range(2):
    rt.arena.zero_grad_force(); step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    rt.arena.zero_grad_force()
    with torch.cuda.graph(g):
        loss = step()
    g.replay(); torch.cuda.synchronize()
    print('CAPTURE %s OK loss %.4f' % (v, float(loss)))
except Exception as e:
    print('CAPTURE %s FAIL %s' % (v, str(e).splitlines()[0][:200]))
