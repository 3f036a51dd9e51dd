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
    _fields_ = [(n, c_int) for n in ("N", "H", "W", "Cin", "Cout", "KH", "KW", "stride", "relu", "residual_mode", "out_fp32")]


SIGNATURES["pe_conv2d_fwd"] = (c_int, [ctypes.POINTER(ConvDesc)] + [c_void_p] * 6)

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
