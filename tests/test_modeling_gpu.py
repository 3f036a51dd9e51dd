"""Module-registry seams (modeling/backbone/build.py:20-33, meta_arch/build.py:12-19, rcnn.py:39-69, 219-267) and
``DefaultPredictor`` (engine/defaults.py:161-198) against the one-call engine path: a model assembled from the registries
must reproduce ``Detector.forward`` exactly, each module must honour its forward contract, and modules must accept
tensors that did not come from their sibling (features / proposals supplied by the caller)."""
import numpy as np
import pytest
import torch

from probenb200 import detector, modeling, weights
from probenb200.structures import Boxes, Instances

pytestmark = pytest.mark.gpu


def _cfg(sd, B=2, canvas=(224, 256)):
    cfg = modeling.get_cfg()
    cfg.ENGINE.MAX_BATCH, cfg.ENGINE.CANVAS, cfg.ENGINE.STATE_DICT = B, canvas, sd
    return cfg


def _same(a, b):
    assert len(a) == len(b)
    assert torch.equal(a.pred_classes, b.pred_classes)
    assert torch.equal(a.pred_boxes.tensor, b.pred_boxes.tensor) and torch.equal(a.scores, b.scores)
    assert torch.equal(a.class_logits, b.class_logits) and torch.equal(a.prob_score, b.prob_score) and torch.equal(a.vars, b.vars)


def test_registry_built_model_equals_engine_forward():
    sd = weights.random_state_dict(50, 3, 3, seed=1)
    g = torch.Generator().manual_seed(5)
    imgs = [torch.rand(3, 200, 250, generator=g) * 255 for _ in range(2)]
    inputs = [{"image": im, "height": 128, "width": 160} for im in imgs]
    want = detector.Detector(sd, depth=50, num_classes=3, max_batch=2, canvas=(224, 256))(inputs)
    cfg = _cfg(sd)
    assert cfg.MODEL.META_ARCHITECTURE in modeling.META_ARCH_REGISTRY and cfg.MODEL.BACKBONE.NAME in modeling.BACKBONE_REGISTRY
    model = modeling.build_model(cfg)
    got = model(inputs)
    assert sum(len(o["instances"]) for o in want) > 0
    for w, g_ in zip(want, got):
        _same(w["instances"], g_["instances"])
        assert g_["instances"].image_size == (128, 160)


def test_each_module_honours_its_forward_contract():
    sd = weights.random_state_dict(50, 3, 3, seed=1)
    cfg = _cfg(sd)
    backbone = modeling.build_backbone(cfg)
    rpn = modeling.build_proposal_generator(cfg, backbone.output_shape())
    heads = modeling.build_roi_heads(cfg, backbone.output_shape())
    assert backbone.size_divisibility == 32 and set(backbone.output_shape()) == {"p2", "p3", "p4", "p5", "p6"}
    assert backbone.output_shape()["p3"].stride == 8 and backbone.output_shape()["p3"].channels == 256
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(2, 3, 224, 256, generator=g) * 255 - 110).cuda()
    feats = backbone(x)
    assert [tuple(feats["p%d" % l].shape) for l in (2, 3, 4, 5, 6)] == [(2, 256, 56, 64), (2, 256, 28, 32), (2, 256, 14, 16), (2, 256, 7, 8), (2, 256, 4, 4)]
    assert feats["p2"].dtype == torch.float32
    images = modeling.ImageList(x, [(200, 250), (200, 250)])
    props, losses = rpn(images, feats, None)
    assert losses == {} and len(props) == 2 and all(isinstance(p, Instances) and 0 < len(p) <= 1000 for p in props)
    b = props[0].proposal_boxes.tensor
    assert float(b[:, 0::2].max()) <= 250 and float(b[:, 1::2].max()) <= 200 and float(b.min()) >= 0
    res, losses = heads(images, feats, props, None)
    assert losses == {} and len(res) == 2
    assert set(res[0].get_fields()) == {"pred_boxes", "scores", "pred_classes", "class_logits", "prob_score", "vars"}
    # the same stages driven with CALLER-supplied tensors (copies, so nothing is recognised as "already in the workspace"):
    feats2 = {k: v.clone() for k, v in feats.items()}
    props2 = []
    for p in props:
        q = Instances(p.image_size)
        q.proposal_boxes = Boxes(p.proposal_boxes.tensor.clone())
        props2.append(q)
    res2, _ = heads(images, feats2, props2, None)
    for a, b_ in zip(res, res2):
        _same(a, b_)
    props3, _ = rpn(images, feats2, None)
    for a, b_ in zip(props, props3):
        assert torch.equal(a.proposal_boxes.tensor, b_.proposal_boxes.tensor)


def test_default_predictor_equals_frames_path():
    """engine/defaults.py:177-198: one BGR uint8 frame in -> resize shortest edge -> model -> Instances in the frame's own
    coordinates; must equal the batched ``forward_frames_device`` path the CLIs use."""
    sd = weights.random_state_dict(50, 3, 3, seed=20)
    det = detector.Detector(sd, depth=50, num_classes=3, max_batch=1, canvas=(224, 256))
    rng = np.random.default_rng(3)
    frame = rng.integers(0, 256, (128, 160, 3), dtype=np.uint8)
    pred = detector.DefaultPredictor(det, min_size=200, max_size=333)
    out = pred(frame)["instances"]
    nh, nw = detector.resize_shortest_edge_shape(128, 160, 200, 333)
    assert (nh, nw) == (200, 250)
    buf = det.forward_frames_device(torch.from_numpy(frame)[None].cuda(), (nh, nw), out=detector.DetectionBuffers(1, 3, det.device))
    want = buf.to_instances([(128, 160)])[0]
    assert len(out) > 0 and out.image_size == (128, 160)
    _same(out, want)
