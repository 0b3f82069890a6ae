import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch, torch.nn.functional as F
from test_gpu_engine import _pair, _rel, _cos, _emulate_bf16_storage
from convnet.pytorch_b200.models import mobilenet_v2

def factory(**cfg):
    m = mobilenet_v2(**cfg); m.classifier[0].p = 0.0; return m
ref, mine, x, y = _pair(factory, dict(dataset='imagenet'), (3, 96, 96), 1000, steps=3, batch=16)
xq = x.to(torch.bfloat16).float()
mine.train(); mine._b200.arena.zero_grad()
lo = mine(x); F.cross_entropy(lo, y).backward(); torch.cuda.synchronize()
emu = _emulate_bf16_storage(copy.deepcopy(ref)).train(); emu.zero_grad()
le = emu(xq); F.cross_entropy(le, y).backward()
ref.train(); ref.zero_grad(); lr_ = ref(xq); F.cross_entropy(lr_, y).backward()
print('logits: mine-fp32 %.3e  emu-fp32 %.3e  mine-emu %.3e' % (_rel(lo, lr_), _rel(le, lr_), _rel(lo, le)))
rows = []
for (n, p), (_, q), (_, e) in zip(mine.named_parameters(), ref.named_parameters(), emu.named_parameters()):
    if float(q.grad.norm()) > 0:
        rows.append((n, _cos(p.grad, q.grad), _cos(e.grad, q.grad), _cos(p.grad, e.grad), float(q.grad.norm()), float(p.grad.norm())))
for r in rows:
    flag = ' <<<' if r[1] < r[2] - 0.05 else ''
    print('%-45s mine/fp32 %.4f  emu/fp32 %.4f  mine/emu %.4f  |ref| %.2e |mine| %.2e%s' % (r + (flag,)))
