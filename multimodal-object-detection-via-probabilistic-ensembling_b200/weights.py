"""State dicts for the detector with the reference's parameter names (SURVEY.md §5 "Checkpoint"):
``backbone.bottom_up.{stem,res2..res5}``, ``backbone.fpn_{lateral,output}{2..5}``,
``proposal_generator.rpn_head.{conv,objectness_logits,anchor_deltas}``, ``roi_heads.box_head.fc{1,2}``,
``roi_heads.box_predictor.{cls_score,bbox_pred,var_pred}``.

``random_state_dict`` makes seeded synthetic weights for tests / benchmarks (no checkpoints are available
offline).  Plain detectron2 init overflows through the residual stack, so the last
norm of each block is down-scaled and the heads are widened until the detector emits proposals and
detections like a trained model does.  ``load_checkpoint`` reads a ``.pth`` (bare state dict or {"model": ...}) or a
Detectron2 model-zoo ``.pkl`` the way detectron2/checkpoint/detection_checkpoint.py:26-45 does.
"""
import math
import pickle

import torch

STAGE_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}


def _conv_w(g, cout, cin, k, gain=1.0):
    std = gain * math.sqrt(2.0 / (cin * k * k))
    return torch.randn(cout, cin, k, k, generator=g) * std


def _bn(g, c, sd, name, wscale=1.0):
    sd[name + ".weight"] = (0.8 + 0.4 * torch.rand(c, generator=g)) * wscale
    sd[name + ".bias"] = 0.05 * torch.randn(c, generator=g)
    sd[name + ".running_mean"] = 0.05 * torch.randn(c, generator=g)
    sd[name + ".running_var"] = 0.8 + 0.4 * torch.rand(c, generator=g)


def random_state_dict(depth=50, in_channels=3, num_classes=3, seed=0, middle_fusion=False, head_gain=1.0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    p = "backbone.bottom_up"
    sd[p + ".stem.conv1.weight"] = _conv_w(g, 64, in_channels, 7) * (1.0 / 60.0)  # inputs are 0..255 - mean
    _bn(g, 64, sd, p + ".stem.conv1.norm")
    cin = 64
    for si, nb in enumerate(STAGE_BLOCKS[depth]):
        mid, cout = 64 * 2 ** si, 256 * 2 ** si
        for b in range(nb):
            q = "%s.res%d.%d" % (p, si + 2, b)
            if b == 0:
                sd[q + ".shortcut.weight"] = _conv_w(g, cout, cin, 1, 0.7)
                _bn(g, cout, sd, q + ".shortcut.norm")
            sd[q + ".conv1.weight"] = _conv_w(g, mid, cin, 1)
            _bn(g, mid, sd, q + ".conv1.norm")
            sd[q + ".conv2.weight"] = _conv_w(g, mid, mid, 3)
            _bn(g, mid, sd, q + ".conv2.norm")
            sd[q + ".conv3.weight"] = _conv_w(g, cout, mid, 1)
            _bn(g, cout, sd, q + ".conv3.norm", wscale=0.25)
            cin = cout
    for lvl, c in zip((2, 3, 4, 5), (256, 512, 1024, 2048)):
        sd["backbone.fpn_lateral%d.weight" % lvl] = _conv_w(g, 256, c, 1, 0.7)
        sd["backbone.fpn_lateral%d.bias" % lvl] = 0.02 * torch.randn(256, generator=g)
        sd["backbone.fpn_output%d.weight" % lvl] = _conv_w(g, 256, 256, 3, 0.7)
        sd["backbone.fpn_output%d.bias" % lvl] = 0.02 * torch.randn(256, generator=g)
    fc = 512 if middle_fusion else 256
    r = "proposal_generator.rpn_head"
    sd[r + ".conv.weight"] = _conv_w(g, fc, fc, 3)
    sd[r + ".conv.bias"] = 0.02 * torch.randn(fc, generator=g)
    sd[r + ".objectness_logits.weight"] = _conv_w(g, 3, fc, 1, 0.3)
    sd[r + ".objectness_logits.bias"] = torch.zeros(3)
    sd[r + ".anchor_deltas.weight"] = _conv_w(g, 12, fc, 1, 0.15)
    sd[r + ".anchor_deltas.bias"] = torch.zeros(12)
    h = "roi_heads.box_head"
    sd[h + ".fc1.weight"] = torch.randn(1024, fc * 49, generator=g) * math.sqrt(2.0 / (fc * 49))
    sd[h + ".fc1.bias"] = 0.02 * torch.randn(1024, generator=g)
    sd[h + ".fc2.weight"] = torch.randn(1024, 1024, generator=g) * math.sqrt(2.0 / 1024)
    sd[h + ".fc2.bias"] = 0.02 * torch.randn(1024, generator=g)
    q = "roi_heads.box_predictor"
    K = num_classes
    sd[q + ".cls_score.weight"] = torch.randn(K + 1, 1024, generator=g) * (0.08 * head_gain)
    sd[q + ".cls_score.bias"] = torch.zeros(K + 1)
    sd[q + ".bbox_pred.weight"] = torch.randn(4 * K, 1024, generator=g) * 0.02
    sd[q + ".bbox_pred.bias"] = torch.zeros(4 * K)
    sd[q + ".var_pred.weight"] = torch.randn(1, 1024, generator=g) * 0.02
    sd[q + ".var_pred.bias"] = torch.zeros(1)
    return sd


class _NumpyOnlyUnpickler(pickle.Unpickler):
    """Model-zoo ``.pkl`` files are plain dicts of numpy arrays and strings; nothing else is allowed to be constructed."""

    _ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"),
                ("numpy", "dtype"), ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
                ("collections", "OrderedDict"), ("_codecs", "encode")}

    def find_class(self, module, name):
        if (module, name) not in self._ALLOWED:
            raise pickle.UnpicklingError("refusing to unpickle %s.%s from a checkpoint" % (module, name))
        return super().find_class(module, name)


def _load_pkl(path):
    """detectron2/checkpoint/detection_checkpoint.py:26-39: a ``.pkl`` in the Detectron2 model-zoo format is
    ``{"model": {name: ndarray}, "__author__": ...}`` with the same parameter names as a ``.pth`` state dict.  Caffe2 /
    Detectron1 files (no ``__author__``; names matched by heuristics, c2_model_loading.py) are outside the path."""
    with open(path, "rb") as f:
        data = _NumpyOnlyUnpickler(f, encoding="latin1").load()
    if not (isinstance(data, dict) and "model" in data and "__author__" in data):
        raise RuntimeError("%s is not a Detectron2 model-zoo .pkl (Caffe2 / Detectron1 checkpoints need name-matching "
                           "heuristics that are not part of this path; convert them with the reference first)" % path)
    return data["model"]


def load_checkpoint(path, num_classes=None):
    """``.pth`` (bare state dict or ``{"model": ...}``) or model-zoo ``.pkl`` -> {name: float tensor}
    (detection_checkpoint.py:26-45).  ``.pth`` files are read with ``weights_only=True``: a checkpoint cannot run code.
    A zoo model has no ``roi_heads.box_predictor.var_pred`` (the fork's variance head): the reference leaves that layer at
    its unseeded random init (fast_rcnn.py:509-512), so its ``vars`` output for such a model is noise; here the layer is
    zero-filled (variance 1) and a note is printed."""
    if str(path).endswith(".pkl"):
        obj = _load_pkl(path)
    else:
        obj = torch.load(path, map_location="cpu", weights_only=True)
        if isinstance(obj, dict) and "model" in obj and isinstance(obj["model"], dict):
            obj = obj["model"]
    sd = {k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(v)) for k, v in obj.items() if not isinstance(v, (str, bytes))}
    q = "roi_heads.box_predictor.var_pred"
    if q + ".weight" not in sd and "roi_heads.box_predictor.cls_score.weight" in sd:
        print("checkpoint %s has no %s: variance head zero-filled (vars = 1)" % (path, q))
        sd[q + ".weight"] = torch.zeros(1, sd["roi_heads.box_predictor.cls_score.weight"].shape[1])
        sd[q + ".bias"] = torch.zeros(1)
    if num_classes is not None and "roi_heads.box_predictor.cls_score.weight" in sd:
        k = sd["roi_heads.box_predictor.cls_score.weight"].shape[0] - 1
        if k != num_classes:
            raise RuntimeError("checkpoint %s has %d classes, the model was configured for %d" % (path, k, num_classes))
    return sd
