"""Host-side packing logic of probenb200.fusion (CPU only)."""
import numpy as np
import pytest

import proben_cases as pc
from probenb200 import fusion, synth


def test_pack_roundtrip():
    dets = synth.synth_model_detections(12, 3, seed=3)
    images = [[synth.image_info(d, i) for d in dets] for i in range(12)]
    p = fusion.pack_detections(images)
    assert p["B"] == 12 and p["M"] == 3 and p["K"] == 3
    assert p["offsets"][0] == 0 and p["offsets"][-1] == len(p["scores"])
    back = pc.packed_to_images(p)
    for a, b in zip(images, back):
        for ia, ib in zip(a, b):
            assert np.array_equal(np.float32(ia["bbox"]).reshape(-1, 4), np.float32(ib["bbox"]).reshape(-1, 4))
            assert ia["class"] == ib["class"]


def test_synth_packed_layout():
    p = synth.synth_packed(1000, num_models=2, seed=1)
    o = p["offsets"]
    assert len(o) == 2001 and o[-1] == len(p["scores"]) and (np.diff(o) >= 0).all()
    assert p["boxes"].dtype == np.float32 and p["classes"].dtype == np.int32
    assert (p["scores"] == p["probs"].max(1)).all()


def test_bad_method_rejected():
    with pytest.raises(ValueError):
        fusion._method_codes(["median", "avg"])


def test_missing_library_is_loud(monkeypatch):
    from probenb200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libprobenb200.so")
    with pytest.raises(RuntimeError):
        _lib.load()


def test_operator_dropins_refuse_cpu_tensors():
    """No CPU path exists behind the detectron2.layers drop-ins or the input-side ops: CPU tensors raise RuntimeError
    (the reference's ops dispatch on `input.is_cuda()`; here the other branch does not exist)."""
    import pytest
    import torch
    from probenb200 import io as pio
    from probenb200 import layers
    with pytest.raises(RuntimeError):
        layers.batched_nms(torch.zeros(3, 4), torch.zeros(3), torch.zeros(3, dtype=torch.int64), 0.5)
    with pytest.raises(RuntimeError):
        layers.ROIAlign((7, 7), 1.0, 0)(torch.zeros(1, 1, 8, 8), torch.zeros(1, 5))
    with pytest.raises(RuntimeError):
        pio.resize_u8(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), (4, 4))
    # shape errors are reported before any device work, like the reference's asserts
    with pytest.raises((RuntimeError, AssertionError)):
        layers.ROIAlign((7, 7), 1.0, 0)(torch.zeros(1, 1, 8, 8), torch.zeros(1, 4))
    # empty inputs short-circuit without touching the GPU
    assert layers.nms(torch.zeros(0, 4), torch.zeros(0), 0.5).shape == (0,)
    assert layers.ROIAlign((7, 7), 1.0, 0)(torch.zeros(0, 3, 10, 10), torch.zeros(0, 5)).shape == (0, 3, 7, 7)
