"""Runs tools/libshiftprobe.so: are row-shifted 128B-swizzle UMMA descriptors usable (base_offset or not)?"""
import ctypes, os, sys, json
import torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libshiftprobe.so'))
lib.shift_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
torch.manual_seed(0)
res = []
for mode in (0, 1):
    for N in (64, 128):
        if mode == 0:
            a = torch.randn(256, 64, device='cuda').to(torch.bfloat16)
            b = torch.randn(N, 64, device='cuda').to(torch.bfloat16)
        else:
            a = torch.randn(96, 128, device='cuda').to(torch.bfloat16)
            b = torch.randn(96, N, device='cuda').to(torch.bfloat16)
        for shift in (0, 8, 1, 3, 9, 30 if mode else 58, 31 if mode else 117):
            for bo in (0, 1):
                out = torch.zeros(128, N, device='cuda')
                rc = lib.shift_probe(a.data_ptr(), b.data_ptr(), mode, shift, bo, N, out.data_ptr())
                if rc != 0:
                    res.append(dict(mode=mode, N=N, shift=shift, base_off=bo, rc=rc)); print(res[-1]); sys.exit(1)
                if mode == 0:
                    ref = a[shift:shift + 128].double() @ b.double().t()
                else:
                    ref = a[shift:shift + 64].double().t() @ b[:64].double()
                err = float((out.double() - ref).abs().max() / ref.abs().max())
                res.append(dict(mode=mode, N=N, shift=shift, base_off=bo, rel_err=round(err, 6)))
                print(json.dumps(res[-1]))
