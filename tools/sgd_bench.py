"""Microbenchmark of b200_fused_sgd on a ResNet-50-sized arena: with / without the folded gradient clearing, and the
separate-memset alternative.  usage: python tools/sgd_bench.py [n_params]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from convnet.pytorch_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25_557_032
dev = torch.device('cuda')
p32 = torch.randn(n, device=dev)
g32 = torch.randn(n, device=dev)
m32 = torch.zeros(n, device=dev)
p16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def run(label, fn, iters=10):
    ts = []
    for _ in range(iters):
        flush.zero_()                      # evict the arenas from L2 between runs
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print('%-34s median %7.1f us  min %7.1f us' % (label, ts[len(ts) // 2], ts[0]))


def sgd(zero):
    ops.fused_sgd(p32, g32, m32, p16, n, n - 1000, 0.1, 0.9, 0.0, 1e-4, 1.0, None, False, zero_grad=zero)


run('fused_sgd zero_grad=0', lambda: sgd(False))
run('fused_sgd zero_grad=1', lambda: sgd(True))
run('fused_sgd zero_grad=0 + memset', lambda: (sgd(False), g32.zero_()))
run('memset g32 alone', lambda: g32.zero_())
