"""Micro-benchmark of the BN kernels on the biggest ResNet-50 activation shapes (L2 flushed between launches)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from convnet.pytorch_b200 import ops
bf16 = torch.bfloat16
SHAPES = [(256 * 56 * 56, 64), (256 * 56 * 56, 256), (256 * 28 * 28, 512), (256 * 14 * 14, 1024), (256 * 7 * 7, 2048)]
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[2]


for M, C in SHAPES:
    z = torch.randn(M, C, device='cuda').to(bf16); dy = torch.randn(M, C, device='cuda').to(bf16)
    y = torch.randn(M, C, device='cuda').to(bf16); out = torch.empty_like(z); g = torch.empty_like(z)
    gamma = torch.ones(C, device='cuda'); beta = torch.zeros(C, device='cuda')
    rm = torch.zeros(C, device='cuda'); rv = torch.ones(C, device='cuda'); nbt = torch.zeros((), dtype=torch.int64, device='cuda')
    mean, invstd, scale, shift = [torch.empty(C, device='cuda') for _ in range(4)]
    sums = torch.empty(2 * C, device='cuda'); dg = torch.zeros(C, device='cuda'); db = torch.zeros(C, device='cuda')
    ws = torch.zeros(ops.bn_workspace_floats(C), device='cuda')
    nb = M * C * 2
    res = {}
    t = timeit(lambda: ops.bn_stats(z, gamma, beta, 1e-5, 0.1, rm, rv, nbt, mean, invstd, scale, shift, ws)); res['stats'] = (t, 1)
    t = timeit(lambda: ops.bn_apply(z, scale, shift, 1, out=out)); res['apply'] = (t, 2)
    t = timeit(lambda: ops.bn_apply(z, scale, shift, 1, residual=y, out=out)); res['apply_res'] = (t, 3)
    t = timeit(lambda: ops.bn_bwd_reduce(dy, None, z, 1, mean, invstd, gamma, beta, sums, dg, db, ws)); res['red'] = (t, 2)
    t = timeit(lambda: ops.bn_bwd_reduce(dy, y, z, 1, mean, invstd, gamma, beta, sums, dg, db, ws)); res['red_y'] = (t, 3)
    t = timeit(lambda: ops.bn_bwd_dx(dy, None, z, 1, mean, invstd, gamma, beta, sums, dz=out)); res['dx'] = (t, 3)
    t = timeit(lambda: ops.bn_bwd_dx(dy, y, z, 1, mean, invstd, gamma, beta, sums, dz=out, g_out=g)); res['dx_y_g'] = (t, 5)
    print('M=%d C=%d  ' % (M, C) + '  '.join('%s %.0fus %.1fTB/s' % (k, v[0] * 1e3, v[1] * nb / v[0] / 1e9) for k, v in res.items()))
