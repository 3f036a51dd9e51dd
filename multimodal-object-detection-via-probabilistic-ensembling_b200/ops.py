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
    if x.dtype != w.dtype or x.dtype not in (torch.bfloat16, torch.float16):
        raise RuntimeError("probenb200.conv2d_nhwc: x and w must both be bfloat16 (or both float16)")
    N, H, W, Cin = x.shape
    Cout, KH, KW, Cin2 = w.shape
    if Cin2 != Cin:
        raise RuntimeError("probenb200.conv2d_nhwc: channel mismatch")
    Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if stride == 2 else (H, W)
    y = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=x.device)
    d = _lib.ConvDesc(N, H, W, Cin, Cout, KH, KW, stride, int(relu), int(residual_mode), int(out_fp32),
                      int(x.dtype == torch.float16))
    st = lib.pe_conv2d_fwd(ctypes.byref(d), _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y),
                           _lib.current_stream_ptr(x.device))
    _lib.check(st, "pe_conv2d_fwd")
    return y


def conv1x1_dual_nhwc(x, x2, w, bias=None, stride2=1, relu=False):
    """``pe_conv1x1_dual_fwd``: y = act([x | x2 at stride2] . w + bias) with x [N,H,W,Cin], x2 [N,H2,W2,Cin2], w [Cout, Cin + Cin2]
    (bottleneck conv3 + projection shortcut as one GEMM, resnet.py:205-221)."""
    lib = _lib.load()
    _lib.require_cuda(x, x2, w, bias)
    if not (x.dtype == x2.dtype == w.dtype == torch.bfloat16):
        raise RuntimeError("probenb200.conv1x1_dual_nhwc: bfloat16 operands expected")
    N, H, W, Cin = x.shape
    _, H2, W2, Cin2 = x2.shape
    Cout = w.shape[0]
    if w.numel() != Cout * (Cin + Cin2):
        raise RuntimeError("probenb200.conv1x1_dual_nhwc: weight must be [Cout, Cin + Cin2]")
    y = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device=x.device)
    d = _lib.ConvDesc(N, H, W, Cin, Cout, 1, 1, 1, int(relu), 0, 0, 0)
    st = lib.pe_conv1x1_dual_fwd(ctypes.byref(d), _lib.ptr(x), _lib.ptr(x2), Cin2, H2, W2, int(stride2), _lib.ptr(w), _lib.ptr(bias),
                                 _lib.ptr(y), _lib.current_stream_ptr(x.device))
    _lib.check(st, "pe_conv1x1_dual_fwd")
    return y


def conv1x1_chain_nhwc(x, w, bias, wc, bias_c, residual=None, x2=None, stride2=1, relu=True, relu_c=True):
    """``pe_conv1x1_chain_fwd``: y = act([x | x2] . w + bias (+ residual)); y_c = act_c(y . wc + bias_c) computed from y's
    shared-memory tiles in the same kernel.  Returns (y, y_c)."""
    lib = _lib.load()
    _lib.require_cuda(x, w, bias, wc, bias_c, residual, x2)
    N, H, W, Cin = x.shape
    Cout, Nc = w.shape[0], wc.shape[0]
    y = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device=x.device)
    yc = torch.empty((N, H, W, Nc), dtype=torch.bfloat16, device=x.device)
    d = _lib.ConvDesc(N, H, W, Cin, Cout, 1, 1, 1, int(relu), 1 if residual is not None else 0, 0, 0)
    c2, h2, w2 = (x2.shape[3], x2.shape[1], x2.shape[2]) if x2 is not None else (0, 0, 0)
    st = lib.pe_conv1x1_chain_fwd(ctypes.byref(d), _lib.ptr(x), _lib.ptr(x2), c2, h2, w2, int(stride2), _lib.ptr(w), _lib.ptr(bias),
                                  _lib.ptr(residual), _lib.ptr(y), _lib.ptr(wc), _lib.ptr(bias_c), Nc, int(relu_c), _lib.ptr(yc),
                                  _lib.current_stream_ptr(x.device))
    _lib.check(st, "pe_conv1x1_chain_fwd")
    return y, yc


def conv_rpn_head_nhwc(x, w, bias, wc, bias_c, relu=True):
    """``pe_conv_rpn_head_fwd``: y_c = relu(conv3x3(x, w) + bias) . wc + bias_c, x [N,H,W,Cin] bf16, w [256,3,3,Cin] bf16,
    wc [16,256] bf16 -> y_c [N,H,W,16] float32; the hidden 256-channel tensor never leaves the SM (rpn.py:74-85)."""
    lib = _lib.load()
    _lib.require_cuda(x, w, bias, wc, bias_c)
    N, H, W, Cin = x.shape
    yc = torch.empty((N, H, W, 16), dtype=torch.float32, device=x.device)
    d = _lib.ConvDesc(N, H, W, Cin, w.shape[0], 3, 3, 1, int(relu), 0, 0, 0)
    st = lib.pe_conv_rpn_head_fwd(ctypes.byref(d), _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(wc), _lib.ptr(bias_c), _lib.ptr(yc),
                                  _lib.current_stream_ptr(x.device))
    _lib.check(st, "pe_conv_rpn_head_fwd")
    return yc


def linear(x, w, bias=None, relu=False, out_fp32=False):
    """x [M,K] bf16, w [N,K] bf16 -> [M,N]; the H=1 case of the conv kernel."""
    M, K = x.shape
    y = conv2d_nhwc(x.view(1, 1, M, K), w.view(w.shape[0], 1, 1, K), bias, None, 1, relu, 0, out_fp32)
    return y.view(M, w.shape[0])


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def _int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


MAX_PROPOSALS = 1000


def rpn_proposals(rpn_out, image_size, pre_nms_topk=1000, post_nms_topk=1000, nms_thresh=0.7):
    """``pe_rpn_proposals``: rpn_out = list of 5 fp32 tensors [B, H_l, W_l, 16] (3 logits | pad | 12 deltas).
    Returns (proposals [B,1000,4], counts [B]).  find_top_rpn_proposals of the reference."""
    lib = _lib.load()
    _lib.require_cuda(*rpn_out)
    B = rpn_out[0].shape[0]
    dev = rpn_out[0].device
    props = torch.zeros((B, MAX_PROPOSALS, 4), dtype=torch.float32, device=dev)
    counts = torch.zeros((B,), dtype=torch.int32, device=dev)
    nbytes = int(lib.pe_rpn_proposals_workspace_bytes(B))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = lib.pe_rpn_proposals(_ptr_array(rpn_out), _int_array([t.shape[1] for t in rpn_out]), _int_array([t.shape[2] for t in rpn_out]),
                              B, pre_nms_topk, post_nms_topk, float(nms_thresh), float(image_size[0]), float(image_size[1]),
                              _lib.ptr(props), _lib.ptr(counts), _lib.ptr(ws), nbytes, _lib.current_stream_ptr(dev))
    _lib.check(st, "pe_rpn_proposals")
    return props, counts


def roi_align_fpn(features, proposals, counts):
    """``pe_roi_align_fwd``: features = [p2..p5] bf16 NHWC; proposals [B, P, 4]; -> [B*P, 49, C] bf16.
    ROIPooler(7x7, scales 1/4..1/32, sampling_ratio 0, ROIAlignV2) of the reference."""
    lib = _lib.load()
    _lib.require_cuda(proposals, counts, *features)
    B, P, _ = proposals.shape
    C = features[0].shape[3]
    out = torch.empty((B * P, 49, C), dtype=torch.bfloat16, device=proposals.device)
    st = lib.pe_roi_align_fwd(_ptr_array(features), _int_array([f.shape[1] for f in features]), _int_array([f.shape[2] for f in features]),
                              C, _lib.ptr(proposals), _lib.ptr(counts), B, P, _lib.ptr(out), _lib.current_stream_ptr(proposals.device))
    _lib.check(st, "pe_roi_align_fwd")
    return out


def head_postprocess(head_out, proposals, counts, K, image_size, out_size, score_thresh=0.5, nms_thresh=0.5,
                     detections_per_image=100):
    """``pe_head_postprocess``: head_out [B*P, npad] fp32 -> DetectionBuffers (fast_rcnn_inference +
    detector_postprocess of the reference)."""
    from .detector import DetectionBuffers
    lib = _lib.load()
    _lib.require_cuda(head_out, proposals, counts)
    B, P, _ = proposals.shape
    out = DetectionBuffers(B, K, proposals.device)
    det = out.struct()
    st = lib.pe_head_postprocess(_lib.ptr(head_out), head_out.shape[1], _lib.ptr(proposals), _lib.ptr(counts), B, P, K,
                                 float(image_size[0]), float(image_size[1]), float(out_size[0]), float(out_size[1]),
                                 float(score_thresh), float(nms_thresh), int(detections_per_image), ctypes.byref(det),
                                 _lib.current_stream_ptr(proposals.device))
    _lib.check(st, "pe_head_postprocess")
    return out


def resize_frames(frames_u8, dst_hw, round_u8=True, out=None):
    """``pe_resize_frames``: uint8 [B,H,W,C] CUDA frames -> float32 [B,C,dst_h,dst_w] (DefaultPredictor's
    ResizeShortestEdge + HWC->CHW float32, engine/defaults.py:186-192)."""
    lib = _lib.load()
    _lib.require_cuda(frames_u8)
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4:
        raise RuntimeError("probenb200.resize_frames: expected uint8 [B,H,W,C]")
    B, H, W, C = frames_u8.shape
    if out is None:
        out = torch.empty((B, C, dst_hw[0], dst_hw[1]), dtype=torch.float32, device=frames_u8.device)
    st = lib.pe_resize_frames(_lib.ptr(frames_u8), _lib.ptr(out), B, C, H, W, int(dst_hw[0]), int(dst_hw[1]), int(round_u8),
                              _lib.current_stream_ptr(frames_u8.device))
    _lib.check(st, "pe_resize_frames")
    return out
