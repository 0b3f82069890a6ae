"""Runs tools/libpairprobe.so (build line in pair_probe.cu): does the cta_group::2 UMMA path produce A @ B^T?"""
import ctypes, os, json
import torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libpairprobe.so'))
lib.pair_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
torch.manual_seed(0)
for N in (64, 128, 256):
    a = torch.randn(256, 64, device='cuda').to(torch.bfloat16)
    b = torch.randn(N, 64, device='cuda').to(torch.bfloat16)
    out = torch.zeros(256, N, device='cuda')
    rc = lib.pair_probe(a.data_ptr(), b.data_ptr(), N, out.data_ptr())
    ref = a.double() @ b.double().t()
    err = float((out.double() - ref).abs().max() / ref.abs().max()) if rc == 0 else None
    # which half-swaps would explain a mismatch?
    alt = float((out.double() - torch.cat([ref[:, N // 2:], ref[:, :N // 2]], 1)).abs().max() / ref.abs().max()) if rc == 0 else None
    print(json.dumps(dict(N=N, rc=rc, rel_err=err, rel_err_if_B_halves_swapped=alt)))
