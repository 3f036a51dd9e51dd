"""Pins oracle/detector_oracle.py to the reference's GeneralizedRCNN: (a) committed golden outputs produced by
the unmodified reference modeling code (tests/golden/make_golden_detector.py), (b) the live reference when
/root/reference is present, (c) the reference's own known-answer tables (tests/test_roi_align.py:26-39,
tests/test_anchor_generator.py, tests/test_box2box_transform.py semantics)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import detector_oracle as D
from probenb200 import weights

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_detector as G  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_run(name):
    c, fmt, mean, mid, seed = G.CASES[name]
    sd = weights.random_state_dict(50, 3 if mid else c, 3, seed=seed, middle_fusion=mid)
    cfg = D.DetCfg(in_channels=c, pixel_mean=mean, pixel_std=(1.0,) * c, middle_fusion=mid)
    return [D.detector_forward([img], [G.OUT_HW], sd, cfg)[0] for img in G.case_inputs(name)]


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_reproduces_reference_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "detector_golden.npz"))
    for i, r in enumerate(oracle_run(name)):
        key = "%s/%d/" % (name, i)
        assert len(r["scores"]) == len(g[key + "scores"])
        assert np.array_equal(r["pred_classes"].numpy(), g[key + "classes"])
        # same torch build -> bit identical; across builds allow fp32 round-off
        for a, b in ((r["pred_boxes"], "boxes"), (r["scores"], "scores"), (r["class_logits"], "class_logits"),
                     (r["prob_score"], "probs"), (r["vars"], "vars")):
            assert np.allclose(a.numpy(), g[key + b], rtol=1e-5, atol=1e-5), b


def test_roi_align_known_answers():
    """detectron2 tests/test_roi_align.py:12-45: 5x5 arange image, box [1,1,3,3] -> 4x4."""
    from torchvision.ops import roi_align
    x = torch.arange(25, dtype=torch.float32).reshape(1, 1, 5, 5)
    rois = torch.tensor([[0, 1, 1, 3, 3]], dtype=torch.float32)
    old = roi_align(x, rois, (4, 4), 1.0, 0, False)[0, 0]
    want_old = torch.tensor([[7.5, 8, 8.5, 9], [10, 10.5, 11, 11.5], [12.5, 13, 13.5, 14], [15, 15.5, 16, 16.5]])
    assert torch.allclose(old, want_old)
    new = roi_align(x, rois, (4, 4), 1.0, 0, True)[0, 0]
    want_new = torch.tensor([[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]])
    assert torch.allclose(new, want_new)


def test_anchor_and_delta_known_answers():
    """tests/test_anchor_generator.py:14-43 (sizes 32/64, ratios .25/1/4 style grid) and the
    apply_deltas(get_deltas) round trip of tests/test_box2box_transform.py:16-36."""
    a = D.grid_anchors(1, 2, 4, 32)  # H=1, W=2, stride 4, size 32, ratios .5/1/2
    assert a.shape == (6, 4)
    assert torch.allclose(a[1], torch.tensor([-16.0, -16.0, 16.0, 16.0]))
    assert torch.allclose(a[4], torch.tensor([-12.0, -16.0, 20.0, 16.0]))
    w = 32 / (0.5 ** 0.5)
    assert torch.allclose(a[0], torch.tensor([-w / 2, -0.5 * w / 2, w / 2, 0.5 * w / 2]))
    src = torch.tensor([[10.0, 20.0, 50.0, 80.0]])
    dst = torch.tensor([[12.0, 18.0, 61.0, 70.0]])
    sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
    dw, dh = dst[:, 2] - dst[:, 0], dst[:, 3] - dst[:, 1]
    deltas = torch.stack([10 * ((dst[:, 0] + .5 * dw) - (src[:, 0] + .5 * sw)) / sw, 10 * ((dst[:, 1] + .5 * dh) - (src[:, 1] + .5 * sh)) / sh,
                          5 * torch.log(dw / sw), 5 * torch.log(dh / sh)], 1)
    assert torch.allclose(D.apply_deltas(deltas, src, (10.0, 10.0, 5.0, 5.0)), dst, atol=1e-4)
    lv = D.assign_levels(torch.tensor([[0, 0, 10, 10], [0, 0, 224, 224], [0, 0, 500, 500], [0, 0, 1000, 800.0]]))
    assert lv.tolist() == [0, 2, 3, 3]


@pytest.mark.needs_reference
def test_oracle_matches_live_reference_bit_exact():
    import ref_loader
    ns = ref_loader.load_reference_detector_modules()
    saved = {k: sys.modules.get(k) for k in ns.modules}
    sys.modules.update(ns.modules)
    try:
        c, fmt, mean, mid, seed = G.CASES["thermal_only"]
        cfg = ns.make_cfg(depth=50, num_classes=3, in_channels=c, input_format=fmt, pixel_mean=mean)
        model = ns.GeneralizedRCNN(cfg).eval()
        sd = weights.random_state_dict(50, c, 3, seed=9)
        model.load_state_dict(sd, strict=False)
        img = torch.rand(3, 128, 160, generator=torch.Generator().manual_seed(3)) * 255
        with torch.no_grad():
            inst = model([{"image": img, "height": 96, "width": 120}])[0]["instances"]
        r = D.detector_forward([img], [(96, 120)], sd, D.DetCfg())[0]
        assert torch.equal(inst.pred_boxes.tensor, r["pred_boxes"]) and torch.equal(inst.scores, r["scores"])
        assert torch.equal(inst.pred_classes, r["pred_classes"]) and torch.equal(inst.vars, r["vars"])
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
