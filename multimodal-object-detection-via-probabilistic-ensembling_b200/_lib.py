"""ctypes binding of the C-ABI library ``csrc/libprobenb200.so`` (include/probenb200.h).

There is no fallback: if the library is missing or a call returns a non-zero status a RuntimeError is
raised (mirrors the reference, whose ``detectron2._C`` ops raise ``RuntimeError`` through c10::Error).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libprobenb200.so")

c_void_p, c_int, c_float, c_size_t, c_char_p = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                                 ctypes.c_size_t, ctypes.c_char_p)

# name -> (restype, argtypes); must list every PE_API symbol of include/probenb200.h
SIGNATURES = {
    "pe_status_string": (c_char_p, [c_int]),
    "pe_abi_version": (c_int, []),
    "pe_last_error_string": (c_char_p, []),
    "pe_fuse_workspace_bytes": (c_size_t, [c_int]),
    "pe_fuse_max_dets_per_image": (c_int, []),
    "pe_fuse_batch": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_float, c_int, c_int, c_float, c_float] +
                      [c_void_p] * 5 + [c_size_t, c_void_p]),
}



class ConvDesc(ctypes.Structure):
    """struct pe_conv_desc (include/probenb200.h)."""
    _fields_ = [(n, c_int) for n in ("N", "H", "W", "Cin", "Cout", "KH", "KW", "stride", "relu", "residual_mode", "out_fp32", "in_fp16")]


SIGNATURES["pe_conv2d_fwd"] = (c_int, [ctypes.POINTER(ConvDesc)] + [c_void_p] * 6)
SIGNATURES["pe_conv1x1_chain_fwd"] = (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p])
SIGNATURES["pe_conv_rpn_head_fwd"] = (c_int, [ctypes.POINTER(ConvDesc)] + [c_void_p] * 7)
SIGNATURES["pe_conv1x1_dual_fwd"] = (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                              c_void_p, c_void_p])


class DetectorConfig(ctypes.Structure):
    """struct pe_detector_config."""
    _fields_ = ([(n, c_int) for n in ("depth", "in_channels", "middle_fusion", "num_classes", "max_batch", "canvas_h", "canvas_w")] +
                [("pixel_mean", c_float * 8), ("pixel_std", c_float * 8)] +
                [(n, c_float) for n in ("score_thresh", "nms_thresh", "rpn_nms_thresh")] +
                [(n, c_int) for n in ("pre_nms_topk", "post_nms_topk", "detections_per_image")])


class ParamInfo(ctypes.Structure):
    """struct pe_param_info."""
    _fields_ = [("name", ctypes.c_char * 96)] + [(n, c_int) for n in ("kind", "cout", "kh", "kw", "cin")] + \
               [("weight_offset", c_size_t), ("bias_offset", c_size_t)]


class Detections(ctypes.Structure):
    """struct pe_detections (device pointers)."""
    _fields_ = [(n, c_void_p) for n in ("boxes", "scores", "classes", "class_logits", "probs", "vars", "roi_index", "counts")]


SIGNATURES.update({
    "pe_detector_create": (c_int, [ctypes.POINTER(DetectorConfig), ctypes.POINTER(c_void_p)]),
    "pe_detector_destroy": (None, [c_void_p]),
    "pe_detector_num_params": (c_int, [c_void_p]),
    "pe_detector_param_info": (c_int, [c_void_p, c_int, ctypes.POINTER(ParamInfo)]),
    "pe_detector_weight_bytes": (c_size_t, [c_void_p]),
    "pe_detector_workspace_bytes": (c_size_t, [c_void_p]),
    "pe_detector_buffer_info": (c_int, [c_void_p, c_char_p, ctypes.POINTER(c_size_t), ctypes.POINTER(c_int * 4), ctypes.POINTER(c_int)]),
    "pe_detector_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float,
                                    ctypes.POINTER(Detections), c_void_p, c_size_t, c_void_p]),
    "pe_detector_forward_stages": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float,
                                           ctypes.POINTER(Detections), c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "pe_detector_forward_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                           ctypes.POINTER(Detections), c_void_p, c_size_t, c_void_p]),
    "pe_pack_detections": (c_int, [ctypes.POINTER(Detections), c_int, c_int, c_int] + [c_void_p] * 7),
    "pe_detector_set_profiling": (c_int, [c_void_p, c_int]),
    "pe_detector_set_stagger_event": (c_int, [c_void_p, c_void_p, c_int]),
    "pe_detector_last_profile": (c_int, [c_void_p, ctypes.POINTER(c_float), ctypes.POINTER(c_float), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "pe_detector_profile_launches": (c_int, [c_void_p, ctypes.POINTER(c_float), ctypes.POINTER(ctypes.c_double),
                                             ctypes.POINTER(ctypes.c_double), c_int]),
    "pe_detector_profile_kernels": (c_int, [c_void_p, ctypes.c_char_p, ctypes.POINTER(c_float), c_int]),
    "pe_resize_frames": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pe_rpn_proposals_workspace_bytes": (c_size_t, [c_int]),
    "pe_rpn_proposals": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int, c_int, c_int,
                                 c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pe_roi_align_fwd": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int, c_void_p, c_void_p,
                                 c_int, c_int, c_void_p, c_void_p]),
    "pe_head_postprocess": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int] + [c_float] * 6 + [c_int] +
                            [ctypes.POINTER(Detections), c_void_p]),
})

SIGNATURES.update({
    "pe_batched_nms_workspace_bytes": (c_size_t, [c_int]),
    "pe_batched_nms_max_boxes": (c_int, []),
    "pe_batched_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pe_roi_align_forward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_float, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p]),
})

SIGNATURES.update({
    "pe_jpeg_create": (c_int, [ctypes.POINTER(c_void_p)]),
    "pe_jpeg_destroy": (None, [c_void_p]),
    "pe_jpeg_image_info": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "pe_jpeg_decode_batch": (c_int, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_size_t), c_int, c_void_p, c_int, c_int, c_void_p]),
    "pe_resize_u8_cv": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
})

SIGNATURES.update({
    "pe_coco_match_max_gt": (c_int, []),
    "pe_coco_match": (c_int, [c_void_p] * 7 + [c_int, c_void_p, c_int, c_void_p, c_int, ctypes.c_longlong, ctypes.c_longlong, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p]),
})

_lib = None


def load():
    """Returns the loaded CDLL; raises RuntimeError if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "probenb200: %s is missing - run `python -m probenb200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        lib = load()
        name = lib.pe_status_string(status).decode()
        detail = lib.pe_last_error_string().decode()
        raise RuntimeError("probenb200.%s failed: %s %s" % (what, name, detail))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream_ptr(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("probenb200: expected CUDA tensors (no CPU path exists)")
        if t is not None and not t.is_contiguous():
            raise RuntimeError("probenb200: expected contiguous tensors")
