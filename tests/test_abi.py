"""The C-ABI library loads and exports every symbol include/probenb200.h declares (no compute calls)."""
import ctypes
import os
import re

from probenb200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "probenb200.h")).read()
    return sorted(set(re.findall(r"PE_API\s+[\w\s\*]+?\b(pe_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert "pe_fuse_batch" in syms and len(syms) >= 6
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
        assert s in _lib.SIGNATURES, "no ctypes signature for %s" % s
    assert set(_lib.SIGNATURES) <= set(syms)


def test_status_strings_and_argument_validation():
    lib = _lib.load()
    assert lib.pe_status_string(0) == b"PE_OK"
    assert lib.pe_status_string(-2) == b"PE_ERR_UNSUPPORTED"
    assert lib.pe_abi_version() >= 1
    assert lib.pe_fuse_workspace_bytes(10) >= 44
    assert lib.pe_fuse_max_dets_per_image() >= 300
    # invalid arguments are rejected before any CUDA call (safe without a GPU)
    null = ctypes.c_void_p(None)
    st = lib.pe_fuse_batch(null, null, null, null, null, null, 1, 0, 3, 0.5, 0, 0, 640.0, 512.0,
                           null, null, null, null, null, 0, null)
    assert st == -1
    st = lib.pe_fuse_batch(null, null, null, null, null, null, 0, 2, 3, 0.5, 0, 0, 640.0, 512.0,
                           null, null, null, null, null, 0, null)
    assert st == 0  # empty batch is a no-op
