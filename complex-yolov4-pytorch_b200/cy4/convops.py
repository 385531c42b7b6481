"""Thin Python wrappers over the conv C-ABI: used by the step engine (engine.py) and by tests."""
import ctypes

import torch

from . import _lib
from ._sigs_engine import CONV_A_MATRIX, CONV_ACCUM, CONV_OUT_F32, CONV_STATS, CONV_ZERO_ACC, ConvDesc


def rup(x, m):
    return (x + m - 1) // m * m


def conv_desc(B, Hi, Wi, Cin, Cout, k, stride, pad, ldx, ldy, flags=0):
    d = ConvDesc()
    d.B, d.Hi, d.Wi, d.Cin = B, Hi, Wi, Cin
    d.Ho = (Hi + 2 * pad - k) // stride + 1
    d.Wo = (Wi + 2 * pad - k) // stride + 1
    d.Cout, d.ksize, d.stride, d.pad = Cout, k, stride, pad
    d.ldx, d.ldy, d.flags = ldx, ldy, flags
    return d


def pack_fprop(w_oihw, cin_pad=None):
    """OIHW fp32 parameter -> [Cout_pad][kh][kw][cin_pad] fp16 (K-major rows)."""
    L = _lib.lib()
    Cout, Cin, k, _ = w_oihw.shape
    cin_pad = cin_pad or Cin
    out = torch.empty(rup(Cout, 32), k * k * cin_pad, device=w_oihw.device, dtype=torch.float16)
    w = w_oihw.detach().contiguous().float()
    _lib.check(L.cy4_pack_weight_fprop(w.data_ptr(), Cout, Cin, k, cin_pad, out.data_ptr(), _lib.stream()), "pack_fprop")
    return out


def pack_dgrad(w_oihw):
    """OIHW fp32 parameter -> [Cin_pad][kh][kw][Cout] fp16."""
    L = _lib.lib()
    Cout, Cin, k, _ = w_oihw.shape
    out = torch.empty(rup(Cin, 32), k * k * Cout, device=w_oihw.device, dtype=torch.float16)
    w = w_oihw.detach().contiguous().float()
    _lib.check(L.cy4_pack_weight_dgrad(w.data_ptr(), Cout, Cin, k, out.data_ptr(), _lib.stream()), "pack_dgrad")
    return out


def conv_fwd(x_nhwc, w_packed, Cout, k, stride, pad, out=None, out_f32=False, bias=None, stats=None, a_matrix=False,
             accumulate=False):
    """x_nhwc [B,H,W,C] fp16 (last-dim stride 1, channel stride = x.stride(2)).  Returns y [B,Ho,Wo,ld]."""
    L = _lib.lib()
    B, Hi, Wi, Cin = x_nhwc.shape
    ldx = x_nhwc.stride(2)
    Ho = (Hi + 2 * pad - k) // stride + 1
    Wo = (Wi + 2 * pad - k) // stride + 1
    if out is None:
        out = torch.empty(B, Ho, Wo, rup(Cout, 32), device=x_nhwc.device, dtype=torch.float32 if out_f32 else torch.float16)
    flags = (CONV_OUT_F32 if out_f32 else 0) | (CONV_STATS if stats is not None else 0) | (CONV_A_MATRIX if a_matrix else 0) | \
            (CONV_ACCUM if accumulate else 0)
    d = conv_desc(B, Hi, Wi, Cin, Cout, k, stride, pad, ldx, out.stride(2), flags)
    _lib.check(L.cy4_conv_fwd(ctypes.byref(d), x_nhwc.data_ptr(), w_packed.data_ptr(), out.data_ptr(),
                              bias.data_ptr() if bias is not None else None,
                              stats[0].data_ptr() if stats is not None else None,
                              stats[1].data_ptr() if stats is not None else None, _lib.stream()), "conv_fwd")
    return out


def conv_fwd_fused(x_nhwc, w_packed, Cout, k, stride, pad, shift, act, residual=None, out=None, a_matrix=False):
    """Eval-mode fused layer: out = act(conv(x, w_folded) + shift[c]) (+ residual), fp16 NHWC (cy4_conv_fwd_fused)."""
    L = _lib.lib()
    B, Hi, Wi, Cin = x_nhwc.shape
    Ho = (Hi + 2 * pad - k) // stride + 1
    Wo = (Wi + 2 * pad - k) // stride + 1
    if out is None:
        out = torch.empty(B, Ho, Wo, rup(Cout, 32), device=x_nhwc.device, dtype=torch.float16)
    d = conv_desc(B, Hi, Wi, Cin, Cout, k, stride, pad, x_nhwc.stride(2), out.stride(2), CONV_A_MATRIX if a_matrix else 0)
    _lib.check(L.cy4_conv_fwd_fused(ctypes.byref(d), x_nhwc.data_ptr(), w_packed.data_ptr(), out.data_ptr(), shift.data_ptr(), int(act),
                                    residual.data_ptr() if residual is not None else None,
                                    residual.stride(2) if residual is not None else 0, _lib.stream()), "conv_fwd_fused")
    return out


def conv_dgrad(dy_nhwc, w_dgrad, Hi, Wi, Cin, k, stride, pad, out=None, accumulate=False, fuse=None):
    """dy [B,Ho,Wo,Cout] fp16 -> dx [B,Hi,Wi,Cin] fp16.
    fuse = (y_producer [B,Hi,Wi,Cin] fp16, scale [Cin], shift [Cin], act, sum_dz [Cin], sum_dzy [Cin]): cy4_conv_dgrad_fused --
    dx receives dz = dx_total * act'(scale*Y + shift) and the two per-channel sums are accumulated."""
    L = _lib.lib()
    B, Ho, Wo, Cout = dy_nhwc.shape
    if out is None:
        out = torch.empty(B, Hi, Wi, Cin, device=dy_nhwc.device, dtype=torch.float16)
    d = conv_desc(B, Hi, Wi, Cin, Cout, k, stride, pad, out.stride(2), dy_nhwc.stride(2), CONV_ACCUM if accumulate else 0)
    assert (d.Ho, d.Wo) == (Ho, Wo)
    if fuse is None:
        _lib.check(L.cy4_conv_dgrad(ctypes.byref(d), dy_nhwc.data_ptr(), w_dgrad.data_ptr(), out.data_ptr(), _lib.stream()), "conv_dgrad")
    else:
        yp, sc, sh, act, s1, s2 = fuse
        _lib.check(L.cy4_conv_dgrad_fused(ctypes.byref(d), dy_nhwc.data_ptr(), w_dgrad.data_ptr(), out.data_ptr(), yp.data_ptr(), yp.stride(2),
                                          sc.data_ptr(), sh.data_ptr(), int(act), s1.data_ptr(), s2.data_ptr(), _lib.stream()), "conv_dgrad_fused")
    return out


def stem_im2col(x_nchw, k, stride, pad):
    L = _lib.lib()
    B, C, H, W = x_nchw.shape
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    cols = torch.empty(B, Ho, Wo, 32, device=x_nchw.device, dtype=torch.float16)
    x = x_nchw.detach().contiguous().float()
    _lib.check(L.cy4_stem_im2col(x.data_ptr(), B, C, H, W, k, stride, pad, cols.data_ptr(), _lib.stream()), "stem_im2col")
    return cols


def conv_wgrad(x_nhwc, dy_nhwc, Cin, Cout, k, stride, pad, acc=None, a_matrix=False):
    """dw_acc [Cout_pad][k*k][Cin] fp32 (+)= dy^T im2col(x).  x/dy may be wider than Cin/Cout
    (ld >= channels rounded up to 64 is required by the kernel)."""
    L = _lib.lib()
    B, Hi, Wi, _ = x_nhwc.shape
    if acc is None:
        acc = torch.zeros(rup(Cout, 32), k * k, Cin, device=x_nhwc.device, dtype=torch.float32)
    d = conv_desc(B, Hi, Wi, Cin, Cout, k, stride, pad, x_nhwc.stride(2), dy_nhwc.stride(2), CONV_A_MATRIX if a_matrix else 0)
    _lib.check(L.cy4_conv_wgrad(ctypes.byref(d), x_nhwc.data_ptr(), dy_nhwc.data_ptr(), acc.data_ptr(), _lib.stream()), "conv_wgrad")
    return acc


def unpack_wgrad(acc, Cout, Cin, k, scale=1.0, out=None, accumulate=False):
    L = _lib.lib()
    if out is None:
        out = torch.empty(Cout, Cin, k, k, device=acc.device, dtype=torch.float32)
    _lib.check(L.cy4_unpack_wgrad(acc.data_ptr(), Cout, Cin, k, acc.shape[-1], float(scale), None, 1 if accumulate else 0,
                                  out.data_ptr(), _lib.stream()), "unpack_wgrad")
    return out
