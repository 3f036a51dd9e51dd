"""Thin torch-tensor wrappers over the op-level C-ABI entry points (tests and the Python-side glue use
these; the detector engine calls the same kernels natively)."""
import ctypes

import torch

from . import _lib


def conv2d_nhwc(x, w, bias=None, residual=None, stride=1, relu=False, residual_mode=0, out_fp32=False):
    """``pe_conv2d_fwd``: x [N,H,W,Cin] bf16, w [Cout,KH,KW,Cin] bf16 -> y [N,Ho,Wo,Cout] (bf16 | fp32).
    Replaces detectron2.layers.Conv2d(+FrozenBN folded) / nn.Linear (H=1) calls of the reference."""
    lib = _lib.load()
    _lib.require_cuda(x, w, bias, residual)
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise RuntimeError("probenb200.conv2d_nhwc: x and w must be bfloat16")
    N, H, W, Cin = x.shape
    Cout, KH, KW, Cin2 = w.shape
    if Cin2 != Cin:
        raise RuntimeError("probenb200.conv2d_nhwc: channel mismatch")
    Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if stride == 2 else (H, W)
    y = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=x.device)
    d = _lib.ConvDesc(N, H, W, Cin, Cout, KH, KW, stride, int(relu), int(residual_mode), int(out_fp32))
    st = lib.pe_conv2d_fwd(ctypes.byref(d), _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y),
                           _lib.current_stream_ptr(x.device))
    _lib.check(st, "pe_conv2d_fwd")
    return y


def linear(x, w, bias=None, relu=False, out_fp32=False):
    """x [M,K] bf16, w [N,K] bf16 -> [M,N]; the H=1 case of the conv kernel."""
    M, K = x.shape
    y = conv2d_nhwc(x.view(1, 1, M, K), w.view(w.shape[0], 1, 1, K), bias, None, 1, relu, 0, out_fp32)
    return y.view(M, w.shape[0])
