"""Tensor-level wrappers over the C ABI (one Python function per entry point of include/b200conv.h).

Tensors are torch CUDA tensors used purely as device-memory handles: activations are contiguous
[N,H,W,C] bf16, conv weights [K, R*S, C] bf16, statistics / master weights / gradients fp32.
Every call is asynchronous on the current torch CUDA stream.
"""
import ctypes

import torch

from . import lib as _l
from .lib import ACT_NONE, ACT_RELU, ACT_RELU6, ConvDesc, Epilogue  # noqa: F401

bf16 = torch.bfloat16


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ---- optional per-call device timing (bench.py's kernel-class breakdown): CUDA events on the launch stream
_TIMING = None


class _T(object):
    __slots__ = ('name', 'flops', 'nbytes', 'e0')

    def __init__(self, name, flops=0, nbytes=0):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        if _TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _TIMING is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _TIMING.append((self.name, self.flops, self.nbytes, self.e0, e1))
        return False


def start_timing():
    global _TIMING
    _TIMING = []


def stop_timing():
    """-> {class: {'ms', 'calls', 'flops', 'bytes'}} for the calls since start_timing()."""
    global _TIMING
    rec, _TIMING = _TIMING, None
    torch.cuda.synchronize()
    out = {}
    for name, flops, nbytes, e0, e1 in rec or []:
        c = out.setdefault(name, {'ms': 0.0, 'calls': 0, 'flops': 0, 'bytes': 0})
        c['ms'] += e0.elapsed_time(e1)
        c['calls'] += 1
        c['flops'] += flops
        c['bytes'] += nbytes
    return out


def _conv_flops(d):
    """ALGORITHMIC FLOPs of the convolution the reference runs (bench.py's roofline numerators): descriptors of
    padded / re-laid problems (space-to-depth stem 7x7x3 -> 4x4x16, class count padded to 8, grouped convolutions
    executed on a block-diagonal dense weight) carry the true MACs per output pixel in ``algo_macs``."""
    macs = getattr(d, 'algo_macs', None)
    if macs is None:
        macs = d.K * d.R * d.S * d.C
    return 2 * d.N * d.P * d.Q * macs


def _conv_bytes(x, w, out, residual=None):
    """ALGORITHMIC bytes of one convolution launch: every operand tensor once (input, weights, output, residual)."""
    n = x.numel() * x.element_size() + w.numel() * w.element_size() + out.numel() * out.element_size()
    if residual is not None:
        n += residual.numel() * residual.element_size()
    return n


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise _l.B200Error("%s must be a CUDA tensor (no CPU fallback exists)" % name)
    if t.dtype != dtype:
        raise _l.B200Error("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise _l.B200Error("%s must be contiguous" % name)


def out_size(h, r, stride, pad_lo, pad_hi=None):
    pad_hi = pad_lo if pad_hi is None else pad_hi
    return (h + pad_lo + pad_hi - r) // stride + 1


def make_desc(N, H, W, C, K, R, S, stride, pad, P=None, Q=None, x_strides=(0, 0, 0), algo_macs=None, window=0):
    pad_h, pad_w = (pad, pad) if isinstance(pad, int) else pad
    P = out_size(H, R, stride, pad_h) if P is None else P
    Q = out_size(W, S, stride, pad_w) if Q is None else Q
    d = ConvDesc(N, H, W, C, K, R, S, stride, pad_h, pad_w, P, Q, *x_strides, int(window))
    if algo_macs is not None:
        d.algo_macs = int(algo_macs)      # python-side annotation only (not part of the C struct)
    return d


# ------------------------------------------------------------------------------------ convolution
def can_fuse_bn_stats(K):
    """b200_conv_fprop can accumulate BN statistics in its epilogue when the output tile is 64/128/256 wide."""
    n_tiles = (K + 255) // 256
    block_n = ((K + n_tiles - 1) // n_tiles + 15) // 16 * 16
    return K % 64 == 0 and K % block_n == 0 and 256 % block_n == 0


def conv_fprop(x, w, desc, out=None, bias=None, residual=None, act=ACT_NONE, out_fp32=False, bn_stats_ws=None):
    """x [N,H,W,C] bf16, w [K,R*S,C] bf16 -> y [N,P,Q,K] (bf16, or fp32 if out_fp32).
    bn_stats_ws: BN workspace into which the epilogue accumulates per-channel sum / sum^2 (finish: bn_finalize)."""
    _chk(x, bf16, "x"); _chk(w, bf16, "w"); _chk(bias, torch.float32, "bias"); _chk(residual, bf16, "residual")
    if out is None:
        out = torch.empty((desc.N, desc.P, desc.Q, desc.K), device=x.device,
                          dtype=torch.float32 if out_fp32 else bf16)
    ep = Epilogue(_l.ptr(bias), _l.ptr(residual), int(act), int(bool(out_fp32)), _l.ptr(bn_stats_ws))
    with _T('conv_fprop', _conv_flops(desc), _conv_bytes(x, w, out, residual)):
        _l.check(_l.load().b200_conv_fprop(ctypes.byref(desc), x.data_ptr(), w.data_ptr(), out.data_ptr(),
                                           ctypes.byref(ep), _stream()), "b200_conv_fprop")
    return out


def conv_dgrad(dy, wt, desc, out=None, residual=None):
    """dy [N,P,Q,K] bf16, wt [C,R*S,K] bf16 -> dx [N,H,W,C] bf16 (+ residual)."""
    _chk(dy, bf16, "dy"); _chk(wt, bf16, "wt"); _chk(residual, bf16, "residual")
    if out is None:
        out = torch.empty((desc.N, desc.H, desc.W, desc.C), device=dy.device, dtype=bf16)
    with _T('conv_dgrad', _conv_flops(desc), _conv_bytes(dy, wt, out, residual)):
        _l.check(_l.load().b200_conv_dgrad(ctypes.byref(desc), dy.data_ptr(), wt.data_ptr(), out.data_ptr(),
                                           _l.ptr(residual), _stream()), "b200_conv_dgrad")
    return out


_WGRAD_WS = {}


def _wgrad_workspace(device):
    """per-device split-K scratch (allocated once; the library states its size)."""
    ws = _WGRAD_WS.get(device)
    if ws is None:
        ws = torch.empty(int(_l.load().b200_conv_wgrad_workspace_bytes()), device=device, dtype=torch.uint8)
        _WGRAD_WS[device] = ws
    return ws


def conv_wgrad(x, dy, desc, dw):
    """dw [K,R*S,C] fp32 += dy^T (*) x."""
    _chk(x, bf16, "x"); _chk(dy, bf16, "dy"); _chk(dw, torch.float32, "dw")
    ws = _wgrad_workspace(x.device)
    with _T('conv_wgrad', _conv_flops(desc), _conv_bytes(x, dy, dw)):
        _l.check(_l.load().b200_conv_wgrad(ctypes.byref(desc), x.data_ptr(), dy.data_ptr(), dw.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _stream()), "b200_conv_wgrad")
    return dw


def dwconv_fprop(x, w, desc, out=None):
    _chk(x, bf16, "x"); _chk(w, bf16, "w")
    if out is None:
        out = torch.empty((desc.N, desc.P, desc.Q, desc.K), device=x.device, dtype=bf16)
    with _T('dw_fprop', 0, _conv_bytes(x, w, out)):       # HBM-bound class (name does not start with conv_)
        _l.check(_l.load().b200_dwconv_fprop(ctypes.byref(desc), x.data_ptr(), w.data_ptr(), out.data_ptr(), _stream()),
                 "b200_dwconv_fprop")
    return out


def dwconv_dgrad(dy, w, desc, out=None):
    _chk(dy, bf16, "dy"); _chk(w, bf16, "w")
    if out is None:
        out = torch.empty((desc.N, desc.H, desc.W, desc.C), device=dy.device, dtype=bf16)
    with _T('dw_dgrad', 0, _conv_bytes(dy, w, out)):
        _l.check(_l.load().b200_dwconv_dgrad(ctypes.byref(desc), dy.data_ptr(), w.data_ptr(), out.data_ptr(), _stream()),
                 "b200_dwconv_dgrad")
    return out


def dwconv_wgrad(x, dy, desc, dw, workspace):
    _chk(x, bf16, "x"); _chk(dy, bf16, "dy"); _chk(dw, torch.float32, "dw"); _chk(workspace, torch.float32, "ws")
    with _T('dw_wgrad', 0, _conv_bytes(x, dy, dw)):
        _l.check(_l.load().b200_dwconv_wgrad(ctypes.byref(desc), x.data_ptr(), dy.data_ptr(), dw.data_ptr(),
                                             workspace.data_ptr(), workspace.numel() * 4, _stream()), "b200_dwconv_wgrad")
    return dw


# ------------------------------------------------------------------------------------ batch norm
def bn_workspace_floats(C):
    return int(_l.load().b200_bn_workspace_floats(int(C)))


def bn_act_mask_bytes(M, C):
    """bytes of the 1-bit activation mask of an [M, C] tensor (rows padded to 8: the kernels read whole row-quad words)"""
    return (int(M) + 7) // 8 * 8 * (int(C) // 8)


def bn_stats(z, gamma, beta, eps, momentum, running_mean, running_var, nbt, mean, invstd, scale, shift, workspace):
    C = z.shape[-1]
    M = z.numel() // C
    _chk(z, bf16, "z")
    mom = -1.0 if momentum is None else float(momentum)
    with _T('bn_stats', 0, 2 * z.numel()):
        _l.check(_l.load().b200_bn_stats(z.data_ptr(), M, C, _l.ptr(gamma), _l.ptr(beta), float(eps), mom,
                                         _l.ptr(running_mean), _l.ptr(running_var), _l.ptr(nbt), mean.data_ptr(),
                                         invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), workspace.data_ptr(),
                                         _stream()), "b200_bn_stats")


def bn_finalize(M, C, gamma, beta, eps, momentum, running_mean, running_var, nbt, mean, invstd, scale, shift, workspace):
    mom = -1.0 if momentum is None else float(momentum)
    with _T('bn_stats', 0, 0):
        _l.check(_l.load().b200_bn_finalize(int(M), int(C), _l.ptr(gamma), _l.ptr(beta), float(eps), mom,
                                            _l.ptr(running_mean), _l.ptr(running_var), _l.ptr(nbt), mean.data_ptr(),
                                            invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                            workspace.data_ptr(), _stream()), "b200_bn_finalize")


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps, scale, shift):
    C = running_mean.numel()
    _l.check(_l.load().b200_bn_eval_coeffs(C, _l.ptr(gamma), _l.ptr(beta), running_mean.data_ptr(),
                                           running_var.data_ptr(), float(eps), scale.data_ptr(), shift.data_ptr(),
                                           _stream()), "b200_bn_eval_coeffs")


def bn_apply(z, scale, shift, act=ACT_NONE, residual=None, z2=None, scale2=None, shift2=None, out=None, act_mask=None):
    """act_mask: optional uint8 tensor of z.numel()/8 bytes receiving one act'(.) bit per element (bn_bwd_* read it
    instead of y)."""
    C = z.shape[-1]
    M = z.numel() // C
    _chk(z, bf16, "z"); _chk(residual, bf16, "residual"); _chk(z2, bf16, "z2"); _chk(act_mask, torch.uint8, "act_mask")
    if act_mask is not None and act_mask.numel() != bn_act_mask_bytes(M, C):
        raise _l.B200Error("act_mask must hold bn_act_mask_bytes(M, C) bytes")
    if out is None:
        out = torch.empty_like(z)
    with _T('bn_apply', 0, 2 * z.numel() * (2 + (residual is not None) + (z2 is not None))):
        _l.check(_l.load().b200_bn_apply(z.data_ptr(), M, C, scale.data_ptr(), shift.data_ptr(), _l.ptr(residual),
                                         _l.ptr(z2), _l.ptr(scale2), _l.ptr(shift2), int(act), out.data_ptr(),
                                         _l.ptr(act_mask), _stream()),
                 "b200_bn_apply")
    return out


def bn_bwd_reduce(dy, y, z, act, mean, invstd, gamma, beta, sums, dgamma_acc, dbeta_acc, workspace, act_mask=None):
    """y=None and act_mask=None: the activation mask is recomputed from z (valid when nothing was added before the
    activation); act_mask (bits written by bn_apply) takes precedence over y."""
    C = z.shape[-1]
    M = z.numel() // C
    _chk(dy, bf16, "dy"); _chk(y, bf16, "y"); _chk(z, bf16, "z"); _chk(act_mask, torch.uint8, "act_mask")
    with _T('bn_bwd_reduce', 0, 2 * z.numel() * (2 + (y is not None and act_mask is None))):
        _l.check(_l.load().b200_bn_bwd_reduce(dy.data_ptr(), _l.ptr(y), _l.ptr(act_mask), z.data_ptr(), M, C, int(act),
                                              mean.data_ptr(),
                                              invstd.data_ptr(), _l.ptr(gamma), _l.ptr(beta), sums.data_ptr(),
                                              _l.ptr(dgamma_acc), _l.ptr(dbeta_acc), workspace.data_ptr(), _stream()),
                 "b200_bn_bwd_reduce")


def bn_bwd_dx(dy, y, z, act, mean, invstd, gamma, beta, sums, dz=None, g_out=None, act_mask=None):
    C = z.shape[-1]
    M = z.numel() // C
    if dz is None:
        dz = torch.empty_like(z)
    with _T('bn_bwd_dx', 0, 2 * z.numel() * (3 + (y is not None and act_mask is None) + (g_out is not None))):
        _l.check(_l.load().b200_bn_bwd_dx(dy.data_ptr(), _l.ptr(y), _l.ptr(act_mask), z.data_ptr(), M, C, int(act),
                                          mean.data_ptr(),
                                          invstd.data_ptr(), _l.ptr(gamma), _l.ptr(beta), sums.data_ptr(),
                                          dz.data_ptr(), _l.ptr(g_out), _stream()), "b200_bn_bwd_dx")
    return dz


# ------------------------------------------------------------------------------------ pooling
def maxpool_fwd(x, want_argmax=True):
    N, H, W, C = x.shape
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((N, OH, OW, C), device=x.device, dtype=bf16)
    am = torch.empty((N, OH, OW, C), device=x.device, dtype=torch.uint8) if want_argmax else None
    with _T('maxpool_fwd', 0, 2 * x.numel() + 3 * y.numel()):
        _l.check(_l.load().b200_maxpool3x3s2_fwd(x.data_ptr(), N, H, W, C, y.data_ptr(), _l.ptr(am), _stream()),
                 "b200_maxpool3x3s2_fwd")
    return y, am


def maxpool_bwd(dy, argmax, in_shape):
    N, H, W, C = in_shape
    dx = torch.empty(in_shape, device=dy.device, dtype=bf16)
    with _T('maxpool_bwd', 0, 2 * dx.numel() + 3 * dy.numel()):
        _l.check(_l.load().b200_maxpool3x3s2_bwd(dy.data_ptr(), argmax.data_ptr(), N, H, W, C, dx.data_ptr(), _stream()),
                 "b200_maxpool3x3s2_bwd")
    return dx


def bn_apply_maxpool(z, scale, shift, act=ACT_RELU, want_argmax=True):
    """maxpool3x3s2(bf16(act(z*scale+shift))) in one pass (the ImageNet stem tail); returns (pooled, argmax bytes)."""
    N, H, W, C = z.shape
    _chk(z, bf16, "z")
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((N, OH, OW, C), device=z.device, dtype=bf16)
    am = torch.empty((N, OH, OW, C), device=z.device, dtype=torch.uint8) if want_argmax else None
    with _T('bn_apply', 0, 2 * z.numel() + 3 * y.numel()):
        _l.check(_l.load().b200_bn_apply_maxpool3x3s2(z.data_ptr(), N, H, W, C, scale.data_ptr(), shift.data_ptr(),
                                                      int(act), y.data_ptr(), _l.ptr(am), _stream()),
                 "b200_bn_apply_maxpool3x3s2")
    return y, am


def bn_bwd_pooled(dp, argmax, z, act, mean, invstd, gamma, beta, sums, dgamma_acc, dbeta_acc, workspace, dz=None,
                  sums_hook=None):
    """BatchNorm(+activation) backward whose incoming gradient is the gradient of the max-pooled tensor: both kernels
    gather it through the argmax bytes (no materialised pre-pool gradient).  sums_hook(sums) runs between the two
    kernels (SyncBatchNorm all-reduce).  Returns dz."""
    N, H, W, C = z.shape
    _chk(dp, bf16, "dp"); _chk(z, bf16, "z"); _chk(argmax, torch.uint8, "argmax")
    if dz is None:
        dz = torch.empty_like(z)
    lib = _l.load()
    with _T('bn_bwd_reduce', 0, 2 * z.numel() + 3 * dp.numel()):
        _l.check(lib.b200_bn_bwd_reduce_pooled(dp.data_ptr(), argmax.data_ptr(), z.data_ptr(), N, H, W, C, int(act),
                                               mean.data_ptr(), invstd.data_ptr(), _l.ptr(gamma), _l.ptr(beta),
                                               sums.data_ptr(), _l.ptr(dgamma_acc), _l.ptr(dbeta_acc),
                                               workspace.data_ptr(), _stream()), "b200_bn_bwd_reduce_pooled")
    if sums_hook is not None:
        sums_hook(sums)
    with _T('bn_bwd_dx', 0, 4 * z.numel() + 3 * dp.numel()):
        _l.check(lib.b200_bn_bwd_dx_pooled(dp.data_ptr(), argmax.data_ptr(), z.data_ptr(), N, H, W, C, int(act),
                                           mean.data_ptr(), invstd.data_ptr(), _l.ptr(gamma), _l.ptr(beta),
                                           sums.data_ptr(), dz.data_ptr(), _stream()), "b200_bn_bwd_dx_pooled")
    return dz


def avgpool_fwd(x):
    N, H, W, C = x.shape
    y = torch.empty((N, 1, 1, C), device=x.device, dtype=bf16)
    with _T('avgpool', 0, 2 * x.numel()):
        _l.check(_l.load().b200_avgpool_fwd(x.data_ptr(), N, H * W, C, y.data_ptr(), _stream()), "b200_avgpool_fwd")
    return y


def avgpool_bwd(dy, in_shape):
    N, H, W, C = in_shape
    dx = torch.empty(in_shape, device=dy.device, dtype=bf16)
    with _T('avgpool', 0, 2 * dx.numel()):
        _l.check(_l.load().b200_avgpool_bwd(dy.data_ptr(), N, H * W, C, dx.data_ptr(), _stream()), "b200_avgpool_bwd")
    return dx


# ------------------------------------------------------------------------------------ squeeze-and-excitation
def se_pool(r):
    N, H, W, C = r.shape
    out = torch.empty((N, 1, 1, C), device=r.device, dtype=bf16)
    with _T('se', 0, 2 * r.numel()):
        _l.check(_l.load().b200_se_pool(r.data_ptr(), N, H * W, C, out.data_ptr(), _stream()), "b200_se_pool")
    return out


def se_scale_fwd(r, logit):
    N, H, W, C = r.shape
    _chk(logit, torch.float32, "logit")
    out = torch.empty_like(r)
    with _T('se', 0, 4 * r.numel()):
        _l.check(_l.load().b200_se_scale_fwd(r.data_ptr(), logit.data_ptr(), N, H * W, C, out.data_ptr(), _stream()),
                 "b200_se_scale_fwd")
    return out


def se_bwd_reduce(g, r, logit):
    N, H, W, C = r.shape
    out = torch.empty((N, 1, 1, C), device=r.device, dtype=bf16)
    with _T('se', 0, 4 * r.numel()):
        _l.check(_l.load().b200_se_bwd_reduce(g.data_ptr(), r.data_ptr(), logit.data_ptr(), N, H * W, C, out.data_ptr(),
                                              _stream()), "b200_se_bwd_reduce")
    return out


def se_bwd_dx(g, logit, dmean):
    N, H, W, C = g.shape
    out = torch.empty_like(g)
    with _T('se', 0, 4 * g.numel()):
        _l.check(_l.load().b200_se_bwd_dx(g.data_ptr(), logit.data_ptr(), dmean.data_ptr(), N, H * W, C, out.data_ptr(),
                                          _stream()), "b200_se_bwd_dx")
    return out


def act_bwd(dy, y, act):
    out = torch.empty_like(dy)
    _l.check(_l.load().b200_act_bwd(dy.data_ptr(), y.data_ptr(), dy.numel(), int(act), out.data_ptr(), _stream()),
             "b200_act_bwd")
    return out


# ------------------------------------------------------------------------------------ layout / casts
def input_prep(x_nchw, cpad, s2d=False, border=False):
    """NCHW fp32 -> NHWC bf16 (channels zero-padded to cpad) or its 2x2 space-to-depth form (optionally with
    the physical zero border of mode 2: +2 low / +1 high in H and W)."""
    _chk(x_nchw, torch.float32, "x")
    N, C, H, W = x_nchw.shape
    if s2d and border:
        shape = (N, H // 2 + 3, W // 2 + 3, cpad)
    else:
        shape = (N, H // 2, W // 2, cpad) if s2d else (N, H, W, cpad)
    out = torch.empty(shape, device=x_nchw.device, dtype=bf16)
    with _T('input_prep', 0, 4 * x_nchw.numel() + 2 * out.numel()):
        _l.check(_l.load().b200_input_prep(x_nchw.data_ptr(), N, C, H, W, cpad, (2 if border else 1) if s2d else 0, out.data_ptr(),
                                           _stream()), "b200_input_prep")
    return out


def input_prep_u8(x_nhwc_u8, cpad, mean, std, s2d=False, border=False):
    """uint8 NHWC [N,H,W,C] -> the layouts of input_prep, normalised as (u8/255 - mean)/std in the same pass."""
    _chk(x_nhwc_u8, torch.uint8, "x")
    N, H, W, C = x_nhwc_u8.shape
    if s2d and border:
        shape = (N, H // 2 + 3, W // 2 + 3, cpad)
    else:
        shape = (N, H // 2, W // 2, cpad) if s2d else (N, H, W, cpad)
    out = torch.empty(shape, device=x_nhwc_u8.device, dtype=bf16)
    scale = (ctypes.c_float * C)(*[1.0 / (255.0 * float(s)) for s in std])
    bias = (ctypes.c_float * C)(*[-float(m) / float(s) for m, s in zip(mean, std)])
    with _T('input_prep', 0, x_nhwc_u8.numel() + 2 * out.numel()):
        _l.check(_l.load().b200_input_prep_u8(x_nhwc_u8.data_ptr(), N, C, H, W, cpad, (2 if border else 1) if s2d else 0,
                                              scale, bias, out.data_ptr(), _stream()), "b200_input_prep_u8")
    return out


def weight_transpose(w, out=None):
    """bf16 [K,T,C] -> [C,T,K]."""
    K, T, C = w.shape
    if out is None:
        out = torch.empty((C, T, K), device=w.device, dtype=bf16)
    with _T('weight_transpose', 0, 4 * w.numel()):
        _l.check(_l.load().b200_weight_transpose(w.data_ptr(), out.data_ptr(), K, T, C, _stream()),
                 "b200_weight_transpose")
    return out


def transpose_jobs(shapes_and_offsets, device):
    """[(src_off, dst_off, K, T, C), ...] -> (int32 device tensor [n, 6], total_tiles) for weight_transpose_batched."""
    rows, tiles = [], 0
    for src_off, dst_off, K, T, C in shapes_and_offsets:
        rows.append([src_off, dst_off, K, T, C, tiles])
        tiles += T * ((K + 31) // 32) * ((C + 31) // 32)
    return torch.tensor(rows, dtype=torch.int32, device=device), tiles


def weight_transpose_batched(src_base, dst_base, jobs, total_tiles):
    """every job: bf16 [K,T,C] at src_base[src_off:] -> [C,T,K] at dst_base[dst_off:], one launch."""
    _chk(src_base, bf16, "src_base"); _chk(dst_base, bf16, "dst_base"); _chk(jobs, torch.int32, "jobs")
    with _T('weight_transpose', 0, 4 * src_base.numel()):
        _l.check(_l.load().b200_weight_transpose_batched(src_base.data_ptr(), dst_base.data_ptr(), jobs.data_ptr(),
                                                         int(jobs.shape[0]), int(total_tiles), _stream()),
                 "b200_weight_transpose_batched")
    return dst_base


def stem_weight_to_s2d(w_f32, K, C, cpad, out):
    _l.check(_l.load().b200_stem_weight_to_s2d(w_f32.data_ptr(), K, C, cpad, out.data_ptr(), _stream()),
             "b200_stem_weight_to_s2d")
    return out


def stem_wgrad_from_s2d(dw_s2d, K, C, cpad, dw):
    _l.check(_l.load().b200_stem_wgrad_from_s2d(dw_s2d.data_ptr(), K, C, cpad, dw.data_ptr(), _stream()),
             "b200_stem_wgrad_from_s2d")
    return dw


def group_weight_pack(w32_grouped, K, T, C, groups, window, transpose=False, out=None):
    """fp32 [K,T,C/g] -> bf16 [K,T,window] (or [C,T,window] transposed): the block-diagonal operand of a grouped
    convolution at window granularity (window == C: dense expansion)."""
    if out is None:
        out = torch.empty((C if transpose else K, T, window), device=w32_grouped.device, dtype=bf16)
    with _T('weight_transpose', 0, 2 * out.numel()):
        _l.check(_l.load().b200_group_weight_pack(w32_grouped.data_ptr(), K, T, C, groups, int(window), int(bool(transpose)),
                                                  out.data_ptr(), _stream()), "b200_group_weight_pack")
    return out


def group_wgrad_unpack(dw_win, K, T, C, groups, window, dw_grouped):
    with _T('weight_transpose', 0, 4 * dw_grouped.numel()):
        _l.check(_l.load().b200_group_wgrad_unpack(dw_win.data_ptr(), K, T, C, groups, int(window), dw_grouped.data_ptr(),
                                                   _stream()), "b200_group_wgrad_unpack")
    return dw_grouped


def cast_bf16(src, dst):
    with _T('cast', 0, 6 * src.numel()):
        _l.check(_l.load().b200_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()),
                 "b200_cast_f32_to_bf16")
    return dst


# ------------------------------------------------------------------------------------ loss / optimizer
def softmax_ce(logits, target, classes, smooth_eps, loss=None, row_loss=None, dlogits=None, grad_scale=1.0,
               grad_scale_dev=None):
    """logits fp32 [B, ld] (ld >= classes), target int64 [B].  loss fp32[3] (+ row_loss fp32[2B] scratch): mean loss,
    top-1 %, top-5 % -- overwritten; dlogits bf16 [B, ld] = grad_scale * (*grad_scale_dev) / B * dloss/dlogits (pad
    columns zeroed)."""
    if loss is not None and (loss.numel() < 3 or row_loss is None or row_loss.numel() < 2 * logits.shape[0]):
        raise _l.B200Error("softmax_ce: loss needs 3 floats and row_loss 2*B floats")
    B, ld = logits.shape
    _chk(logits, torch.float32, "logits"); _chk(target, torch.int64, "target"); _chk(dlogits, bf16, "dlogits")
    _chk(loss, torch.float32, "loss"); _chk(row_loss, torch.float32, "row_loss")
    _chk(grad_scale_dev, torch.float32, "grad_scale_dev")
    with _T('softmax_ce', 0, 4 * logits.numel()):
        _l.check(_l.load().b200_softmax_ce(logits.data_ptr(), target.data_ptr(), B, int(classes), int(ld),
                                           float(smooth_eps or 0.0), float(grad_scale), _l.ptr(grad_scale_dev),
                                           _l.ptr(loss), _l.ptr(row_loss), _l.ptr(dlogits), _stream()),
                 "b200_softmax_ce")


def colsum_bf16(m, out):
    B, K = m.shape
    _l.check(_l.load().b200_colsum_bf16(m.data_ptr(), B, K, out.data_ptr(), _stream()), "b200_colsum_bf16")


def fused_sgd(p32, g32, m32, p16, n, wd_count, lr, momentum, dampening, weight_decay, inv_scale, clip_coef,
              first_step, zero_grad=False):
    with _T('fused_sgd', 0, (26 if zero_grad else 22) * int(n)):   # 12 B read + 10 B written (+4 B memset) per parameter
        _l.check(_l.load().b200_fused_sgd(p32.data_ptr(), g32.data_ptr(), _l.ptr(m32), _l.ptr(p16), int(n),
                                          int(wd_count), float(lr), float(momentum), float(dampening),
                                          float(weight_decay), float(inv_scale), _l.ptr(clip_coef), int(bool(first_step)),
                                          int(bool(zero_grad)), _stream()), "b200_fused_sgd")


def sumsq(g, n, out, workspace):
    _l.check(_l.load().b200_sumsq(g.data_ptr(), int(n), out.data_ptr(), workspace.data_ptr(), _stream()),
             "b200_sumsq")


def grad_coef(sumsq_t, inv_scale, mode, max_norm, momentum, state, coef_out, norm_out):
    _l.check(_l.load().b200_grad_coef(sumsq_t.data_ptr(), float(inv_scale), int(mode), float(max_norm),
                                      float(momentum), _l.ptr(state), coef_out.data_ptr(), _l.ptr(norm_out),
                                      _stream()), "b200_grad_coef")
