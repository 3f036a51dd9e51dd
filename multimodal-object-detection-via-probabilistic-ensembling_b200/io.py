"""Input side of the detector path on the GPU.

The reference reads every validation pair on the CPU (demo/FLIR/demo_FLIR_save_predictions.py:93-121): ``cv2.imread``
of ``RGB/<name>.jpg`` and ``thermal_8_bit/<name>.jpeg``, ``cv2.resize`` of the RGB frame to the thermal frame's size
(the ``cv2.INTER_CUBIC`` it passes lands in the ``dst`` slot, so the interpolation is the default INTER_LINEAR) and
the assembly of the 3-channel (thermal_only / rgb_only), 4-channel BGRT (early_fusion) or 6-channel BGRTTT
(middle_fusion) array.  Here the JPEG byte strings are decoded by nvJPEG straight into HBM and resized / assembled by
``pe_resize_u8_cv`` (cv2's 8-bit bilinear arithmetic, bit for bit); the result feeds ``Detector.forward_frames_device``.
"""
import ctypes

import numpy as np
import torch

from . import _lib


class JpegDecoder:
    """nvJPEG decoder handle (``pe_jpeg_*``).  ``decode(list_of_bytes)`` -> uint8 CUDA tensor [n, H, W, 3] in BGR,
    ``cv2.imread``'s layout; all images of one call must have the same size."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pe_jpeg_create(ctypes.byref(h)), "jpeg_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pe_jpeg_destroy(self._h)
            self._h = None

    __del__ = close

    def image_info(self, data):
        buf = np.frombuffer(data, np.uint8)
        h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(self._lib.pe_jpeg_image_info(self._h, ctypes.c_void_p(buf.ctypes.data), buf.size, ctypes.byref(h),
                                                ctypes.byref(w), ctypes.byref(c)), "jpeg_image_info")
        return h.value, w.value, c.value

    def decode(self, datas, out=None):
        n = len(datas)
        if n == 0:
            return torch.empty((0, 0, 0, 3), dtype=torch.uint8, device=self.device)
        bufs = [np.frombuffer(d, np.uint8) for d in datas]
        H, W, _ = self.image_info(datas[0])
        if out is None:
            out = torch.empty((n, H, W, 3), dtype=torch.uint8, device=self.device)
        ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (ctypes.c_size_t * n)(*[b.size for b in bufs])
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pe_jpeg_decode_batch(self._h, ptrs, sizes, n, _lib.ptr(out), H, W,
                                                      _lib.current_stream_ptr(self.device)), "jpeg_decode_batch")
        return out


def resize_u8(src, dst_hw, out=None, src_c0=0, dst_c0=0, channels=None):
    """``cv2.resize(src, (w, h))`` for uint8 NHWC CUDA tensors (INTER_LINEAR, OpenCV's fixed-point arithmetic).  Copies
    ``channels`` channels starting at ``src_c0`` into channels ``dst_c0..`` of ``out`` (allocated [B, h, w, channels]
    when omitted)."""
    _lib.require_cuda(src)
    if src.dtype != torch.uint8 or src.dim() != 4:
        raise RuntimeError("resize_u8 expects a uint8 [B, H, W, C] tensor")
    B, Hs, Ws, Cs = src.shape
    channels = Cs - src_c0 if channels is None else channels
    Hd, Wd = int(dst_hw[0]), int(dst_hw[1])
    if out is None:
        out = torch.empty((B, Hd, Wd, dst_c0 + channels), dtype=torch.uint8, device=src.device)
    _lib.require_cuda(out)
    lib = _lib.load()
    with torch.cuda.device(src.device):
        _lib.check(lib.pe_resize_u8_cv(_lib.ptr(src), Cs, src_c0, _lib.ptr(out), out.shape[3], dst_c0, channels, B, Hs, Ws,
                                       Hd, Wd, _lib.current_stream_ptr(src.device)), "resize_u8_cv")
    return out


def assemble_input(method, rgb, thermal):
    """demo_FLIR_save_predictions.py:100-121 on device tensors: ``rgb`` [B, Hr, Wr, 3] and ``thermal`` [B, Ht, Wt, 3]
    uint8 BGR frames -> the uint8 [B, Ht, Wt, C] network input of ``method`` (C = 3, 4 or 6)."""
    if method == "thermal_only":
        return thermal
    B, Ht, Wt, _ = thermal.shape
    if method == "rgb_only":
        return resize_u8(rgb, (Ht, Wt))
    if method == "early_fusion":
        out = torch.empty((B, Ht, Wt, 4), dtype=torch.uint8, device=thermal.device)
        resize_u8(rgb, (Ht, Wt), out=out, channels=3)
        resize_u8(thermal, (Ht, Wt), out=out, src_c0=0, dst_c0=3, channels=1)
        return out
    if method == "middle_fusion":
        out = torch.empty((B, Ht, Wt, 6), dtype=torch.uint8, device=thermal.device)
        resize_u8(rgb, (Ht, Wt), out=out, channels=3)
        resize_u8(thermal, (Ht, Wt), out=out, src_c0=0, dst_c0=3, channels=3)
        return out
    raise ValueError("unknown fusion method %r" % (method,))


def load_pair_batch(decoder, rgb_files, thermal_files, method):
    """Reads the JPEG files of a batch of pairs and returns the assembled uint8 device input of ``method``."""
    thermal = decoder.decode([open(f, "rb").read() for f in thermal_files])
    if method == "thermal_only":
        return thermal
    rgb = decoder.decode([open(f, "rb").read() for f in rgb_files])
    return assemble_input(method, rgb, thermal)
