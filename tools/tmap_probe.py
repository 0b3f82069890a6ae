"""Does cuTensorMapEncode{Tiled,Im2col} accept OVERLAPPING rows (stride of dim1 smaller than the extent of dim0)?"""
import ctypes
import torch
torch.zeros(1, device='cuda')   # creates and binds the primary context
cu = ctypes.CDLL('libcuda.so.1')
print('cuInit', cu.cuInit(0))
tm = (ctypes.c_uint64 * 16)()
u64x = lambda *v: (ctypes.c_uint64 * len(v))(*v)
u32x = lambda *v: (ctypes.c_uint32 * len(v))(*v)
i32x = lambda *v: (ctypes.c_int * len(v))(*v)
base = ctypes.c_void_p(0x7f0000000000)
BF16, SW128, L2_128, NONE = 9, 3, 2, 0
# tiled 2D: 64 elements per "row" (128 B) but rows start every 32 B
r = cu.cuTensorMapEncodeTiled(tm, BF16, 2, base, u64x(64, 1000), u64x(32), u32x(64, 128), u32x(1, 1), NONE, SW128, L2_128, NONE)
print('tiled 2D overlapping (stride 32B, extent 128B):', r)
r = cu.cuTensorMapEncodeTiled(tm, BF16, 2, base, u64x(64, 1000), u64x(128), u32x(64, 128), u32x(1, 1), NONE, SW128, L2_128, NONE)
print('tiled 2D dense control:', r)
# im2col 4D: C=64 (128B) per pixel, pixel stride 32 B, W=112, H=115, N=4
r = cu.cuTensorMapEncodeIm2col(tm, BF16, 4, base, u64x(64, 112, 115, 4), u64x(32, 115 * 32, 115 * 115 * 32),
                               i32x(0, 0), i32x(0, -3), 64, 128, u32x(1, 1, 1, 1), NONE, SW128, L2_128, NONE)
print('im2col 4D overlapping wide pixels:', r)
r = cu.cuTensorMapEncodeTiled(tm, BF16, 4, base, u64x(64, 112, 115, 4), u64x(32, 115 * 32, 115 * 115 * 32),
                              u32x(64, 112, 1, 1), u32x(1, 1, 1, 1), NONE, SW128, L2_128, NONE)
print('tiled 4D overlapping wide pixels:', r)
