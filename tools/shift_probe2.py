"""Runs tools/libshiftprobe2.so: row-shifted UMMA descriptors with 32/64/128-byte rows (K-major A, MN-major B)."""
import ctypes, os, sys, json
import torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libshiftprobe2.so'))
lib.shift_probe2.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
torch.manual_seed(0)
bad = 0
for mode in (0, 1):
    for rb in (32, 64, 128):
        ch = rb // 2
        N = 64 if mode == 0 else ch
        if mode == 0:
            a = torch.randn(256, ch, device='cuda').to(torch.bfloat16)
            b = torch.randn(N, ch, device='cuda').to(torch.bfloat16)
        else:
            a = torch.randn(96, 128, device='cuda').to(torch.bfloat16)
            b = torch.randn(160, ch, device='cuda').to(torch.bfloat16)
        for shift in (0, 8, 1, 2, 3, 5, 9, 30, 58, 67, 95) + ((117, 119) if mode == 0 else ()):
            out = torch.zeros(128, N, device='cuda')
            rc = lib.shift_probe2(a.data_ptr(), b.data_ptr(), mode, rb, shift, N, out.data_ptr())
            if rc != 0:
                print(json.dumps(dict(mode=mode, rb=rb, shift=shift, rc=rc))); sys.exit(1)
            if mode == 0:
                ref = a[shift:shift + 128].double() @ b.double().t()
            else:
                ref = a[:64].double().t() @ b[shift:shift + 64].double()
            err = float((out.double() - ref).abs().max() / ref.abs().max())
            bad += err > 1e-3
            print(json.dumps(dict(mode=mode, rb=rb, N=N, shift=shift, rel_err=round(err, 6))))
print('PROBE2', 'ALL_OK' if bad == 0 else 'MISMATCHES=%d' % bad)
