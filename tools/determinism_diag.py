"""Runs the same training step repeatedly with the caching allocator poisoned by NaNs in between: any kernel that
reads memory it did not write shows up as run-to-run differences or NaNs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import torch.nn.functional as F
from test_gpu_engine import _pair, _rel, _cos
from convnet.pytorch_b200.models import resnet
from oracle import ref_model

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 18
size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ref, mine, x, y = _pair(resnet, dict(dataset='imagenet', depth=depth), (3, size, size), 1000, batch=8)
sd = {k: v.detach().cpu().clone() for k, v in ref.state_dict().items()}
ol, oloss, og, ob = ref_model.loss_and_grads(sd, x.cpu(), y.cpu(), quant=True)
go = torch.cat([og[n].flatten() for n, _ in mine.named_parameters()])


def poison():
    bufs = [torch.full((n,), float('nan'), device='cuda', dtype=torch.float32) for n in (1 << 20, 1 << 22, 1 << 24, 3 << 24)]
    bufs += [torch.full((n,), float('nan'), device='cuda', dtype=torch.bfloat16) for n in (1 << 18, 1 << 21, 1 << 23, 1 << 25)]
    torch.cuda.synchronize()
    del bufs


def step():
    mine.train(); mine._b200.arena.zero_grad()
    lo = mine(x); loss = F.cross_entropy(lo, y); loss.backward(); torch.cuda.synchronize()
    g = torch.cat([p.grad.cpu().flatten() for _, p in mine.named_parameters()]).clone()
    return lo.detach().cpu().clone(), g


prev = None
for it in range(4):
    if it >= 1:
        poison()
    lo, g = step()
    print('run %d: logits vs oracle %.3e  grad vs oracle rel %.3e  nan=%d' % (it, _rel(lo, ol), _rel(g, go), int(torch.isnan(g).sum())))
    if prev is not None:
        print('       vs run0: logits %.3e grads %.3e' % (_rel(lo, prev[0]), _rel(g, prev[1])))
        if _rel(g, prev[1]) > 1e-4:
            off = 0
            for n, p in mine.named_parameters():
                k = p.numel()
                r = _rel(g[off:off + k], prev[1][off:off + k])
                if r > 1e-3:
                    print('         differs: %-40s rel %.3e' % (n, r))
                off += k
    else:
        prev = (lo, g)
