"""Module seams of the reference's model registries, backed by the native engine.

The reference assembles ``GeneralizedRCNN`` from three registries (detectron2/modeling/backbone/build.py:20-33,
proposal_generator/build.py, roi_heads/roi_heads.py ``build_roi_heads``, meta_arch/build.py:12-19) and calls

    features            = backbone(images.tensor)                                  # dict "p2".."p6" -> [N, C, H_l, W_l]
    proposals, _        = proposal_generator(images, features, None)               # list[Instances(proposal_boxes)]
    results, _          = roi_heads(images, features, proposals, None)             # list[Instances(pred_boxes, scores, ...)]
    return GeneralizedRCNN._postprocess(results, batched_inputs, image_sizes)      # rcnn.py:219-267, 288-302

(rcnn.py:39-69 builds them).  The same objects exist here with the same forward contracts; each one runs its stage of the
engine's plan (``pe_detector_forward_stages``) and the stages hand their tensors over through the engine's named workspace
buffers.  A module that receives the tensors another module of the same engine just produced runs in place; tensors that
come from elsewhere (a user's own backbone, hand-made proposals) are copied into the workspace first, so the seams can be
driven independently - that is what tests/test_modeling_gpu.py does.

Limits (stated, not hidden): inference only; all images of a batch share one size (FLIR / KAIST frames do; the reference
pads ragged batches, structures/image_list.py:51-102); proposals carry ``proposal_boxes`` only (the objectness logits stay
inside the engine).
"""
import ctypes
import types

import torch

from . import _lib
from .detector import FLIR_PIXEL_MEAN, DetectionBuffers, Detector
from .structures import Boxes, Instances


class Registry:
    """fvcore.common.registry.Registry as the reference uses it: ``@REG.register()`` and ``REG.get(name)``."""

    def __init__(self, name):
        self._name, self._map = name, {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._map[o.__name__] = o
                return o
            return deco
        self._map[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self._map:
            raise KeyError("No object named '%s' found in '%s' registry!" % (name, self._name))
        return self._map[name]

    def __contains__(self, name):
        return name in self._map


BACKBONE_REGISTRY = Registry("BACKBONE")
PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
META_ARCH_REGISTRY = Registry("META_ARCH")


def get_cfg():
    """The configuration keys this path reads, with the values the FLIR demos end up with (config/defaults.py,
    configs/Base-RCNN-FPN.yaml, demo_FLIR_save_predictions.py:49-73)."""
    ns = types.SimpleNamespace
    return ns(
        INPUT=ns(NUM_IN_CHANNELS=3, FORMAT="BGR", MIN_SIZE_TEST=800, MAX_SIZE_TEST=1333),
        MODEL=ns(META_ARCHITECTURE="GeneralizedRCNN", WEIGHTS="", DEVICE="cuda", PIXEL_MEAN=list(FLIR_PIXEL_MEAN), PIXEL_STD=[1.0, 1.0, 1.0],
                 BACKBONE=ns(NAME="build_resnet_fpn_backbone"), RESNETS=ns(DEPTH=50),
                 PROPOSAL_GENERATOR=ns(NAME="RPN"),
                 RPN=ns(PRE_NMS_TOPK_TEST=1000, POST_NMS_TOPK_TEST=1000, NMS_THRESH=0.7),
                 ROI_HEADS=ns(NAME="StandardROIHeads", NUM_CLASSES=3, SCORE_THRESH_TEST=0.5, NMS_THRESH_TEST=0.5, ENABLE_GAUSSIANNLLOSS=True),
                 ROI_BOX_HEAD=ns(OUTPUT_LOGITS=True)),
        TEST=ns(DETECTIONS_PER_IMAGE=100),
        # not in the reference: the engine allocates its scratch for a fixed canvas / batch
        ENGINE=ns(MAX_BATCH=1, CANVAS=(800, 1024), STATE_DICT=None))


class _Engine:
    """One native engine shared by the three modules of a model (they exchange tensors through its workspace)."""

    def __init__(self, cfg):
        m = cfg.MODEL
        middle = cfg.INPUT.FORMAT == "BGRTTT"
        sd = cfg.ENGINE.STATE_DICT
        if sd is None and m.WEIGHTS:
            from . import weights
            sd = weights.load_checkpoint(m.WEIGHTS, num_classes=m.ROI_HEADS.NUM_CLASSES)
        self.det = Detector(sd, depth=m.RESNETS.DEPTH, num_classes=m.ROI_HEADS.NUM_CLASSES, in_channels=cfg.INPUT.NUM_IN_CHANNELS,
                            middle_fusion=middle, pixel_mean=tuple(m.PIXEL_MEAN), pixel_std=tuple(m.PIXEL_STD),
                            max_batch=cfg.ENGINE.MAX_BATCH, canvas=tuple(cfg.ENGINE.CANVAS), score_thresh=m.ROI_HEADS.SCORE_THRESH_TEST,
                            nms_thresh=m.ROI_HEADS.NMS_THRESH_TEST, rpn_nms_thresh=m.RPN.NMS_THRESH, pre_nms_topk=m.RPN.PRE_NMS_TOPK_TEST,
                            post_nms_topk=m.RPN.POST_NMS_TOPK_TEST, detections_per_image=cfg.TEST.DETECTIONS_PER_IMAGE, device=m.DEVICE)
        self.middle = middle
        self.fc = 512 if middle else 256
        self.token = None  # identity of the feature dict / proposal list currently held by the workspace

    def level_buffer(self, lvl):
        name = ("p%d" % lvl) if (self.middle or lvl == 6) else ("pout%d_0" % lvl)
        raw, dims, _ = self.det.buffer(name)
        return raw.view(torch.bfloat16).view(*dims)

    def run(self, stages, images=None, B=1, img_hw=(1, 1), out_hw=(1.0, 1.0), out=None, prenormalized=True):
        d = self.det
        det = out.struct() if out is not None else None
        with torch.cuda.device(d.device):
            st = d.lib.pe_detector_forward_stages(d.handle, _lib.ptr(d.weights), _lib.ptr(images), B, int(img_hw[0]), int(img_hw[1]),
                                                  float(out_hw[0]), float(out_hw[1]), ctypes.byref(det) if det is not None else None,
                                                  _lib.ptr(d.workspace), d.ws_bytes, stages, int(prenormalized),
                                                  _lib.current_stream_ptr(d.device))
        _lib.check(st, "pe_detector_forward_stages")


def _engine_of(cfg):
    if getattr(cfg, "_engine", None) is None:
        cfg._engine = _Engine(cfg)
    return cfg._engine


class ShapeSpec(types.SimpleNamespace):
    pass


class FPNBackbone:
    """``Backbone`` contract (modeling/backbone/backbone.py, fpn.py:110-145): ``forward(x) -> {"p2": ..., "p6": ...}`` with
    x = ImageList.tensor (normalised, zero-padded to a multiple of ``size_divisibility``), NCHW float32 in and out."""

    size_divisibility = 32

    def __init__(self, cfg, input_shape=None):
        self.engine = _engine_of(cfg)

    def output_shape(self):
        return {"p%d" % l: ShapeSpec(channels=self.engine.fc, stride=2 ** l) for l in range(2, 7)}

    def forward(self, x):
        e = self.engine
        _lib.require_cuda(x)
        B, C, H, W = x.shape
        if (H, W) != e.det.canvas or C != e.det.in_channels or B > e.det.max_batch:
            raise RuntimeError("probenb200.FPNBackbone: expected [<=%d, %d, %d, %d], got %s" % (e.det.max_batch, e.det.in_channels,
                                                                                            e.det.canvas[0], e.det.canvas[1], tuple(x.shape)))
        e.run(_lib_stage("backbone"), images=x.float().contiguous(), B=B, img_hw=(H, W), prenormalized=True)
        feats = {"p%d" % l: e.level_buffer(l)[:B].permute(0, 3, 1, 2).float() for l in range(2, 7)}
        e.token = ("features", id(feats), B)
        e._held = feats
        return feats

    __call__ = forward


def _lib_stage(name):
    return {"backbone": 1, "rpn": 2, "roi_heads": 4}[name]


def _load_features(e, features, B):
    """Makes the engine's workspace hold ``features``: a no-op for the dict its own backbone just returned."""
    if e.token == ("features", id(features), B) or e.token == ("proposals+features", id(features), B):
        return
    for l in range(2, 7):
        f = features["p%d" % l]
        e.level_buffer(l)[:B].copy_(f.permute(0, 2, 3, 1).to(torch.bfloat16))
    e.token = ("features", id(features), B)


@BACKBONE_REGISTRY.register()
def build_resnet_fpn_backbone(cfg, input_shape=None):
    return FPNBackbone(cfg, input_shape)


def build_backbone(cfg, input_shape=None):
    """modeling/backbone/build.py:20-33."""
    if input_shape is None:
        input_shape = ShapeSpec(channels=len(cfg.MODEL.PIXEL_MEAN))
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)


@PROPOSAL_GENERATOR_REGISTRY.register()
class RPN:
    """proposal_generator/rpn.py:88-185: ``forward(images, features, gt_instances=None) -> (list[Instances], losses)``; the
    instances carry ``proposal_boxes`` in network-input coordinates, clipped to each image's size, at most 1000."""

    def __init__(self, cfg, input_shape=None):
        self.engine = _engine_of(cfg)

    def forward(self, images, features, gt_instances=None):
        e = self.engine
        sizes = images.image_sizes
        B = len(sizes)
        if any(tuple(s) != tuple(sizes[0]) for s in sizes):
            raise RuntimeError("probenb200.RPN: all images of a batch must share one size")
        _load_features(e, features, B)
        e.run(_lib_stage("rpn"), B=B, img_hw=sizes[0])
        props = e.det.buffer("proposals")[0].view(torch.float32).view(-1, 1000, 4)[:B]
        counts = e.det.buffer("prop_count")[0].view(torch.int32)[:B].tolist()
        out = []
        for b in range(B):
            inst = Instances(tuple(sizes[b]))
            inst.proposal_boxes = Boxes(props[b, : counts[b]].clone())
            out.append(inst)
        e.token = ("proposals+features", id(features), B)
        e._props = out
        return out, {}

    __call__ = forward


def build_proposal_generator(cfg, input_shape=None):
    return PROPOSAL_GENERATOR_REGISTRY.get(cfg.MODEL.PROPOSAL_GENERATOR.NAME)(cfg, input_shape)


@ROI_HEADS_REGISTRY.register()
class StandardROIHeads:
    """roi_heads/roi_heads.py:595-631 at inference: ``forward(images, features, proposals, targets=None) ->
    (list[Instances], {})`` with the fork's fields (pred_boxes, scores, pred_classes, class_logits, prob_score, vars) in
    network-input coordinates (``detector_postprocess`` is the meta-architecture's job)."""

    def __init__(self, cfg, input_shape=None):
        self.engine = _engine_of(cfg)

    def forward(self, images, features, proposals, targets=None):
        e = self.engine
        sizes = images.image_sizes
        B = len(sizes)
        own = e.token == ("proposals+features", id(features), B) and getattr(e, "_props", None) is proposals
        if not own:
            _load_features(e, features, B)
            pb = e.det.buffer("proposals")[0].view(torch.float32).view(-1, 1000, 4)
            pc = e.det.buffer("prop_count")[0].view(torch.int32)
            for b, inst in enumerate(proposals):
                t = inst.proposal_boxes.tensor[:1000].to(pb.device, torch.float32)
                pb[b, : len(t)] = t
                pc[b] = len(t)
        buf = DetectionBuffers(e.det.max_batch, e.det.num_classes, e.det.device)
        h, w = sizes[0]
        e.run(_lib_stage("roi_heads"), B=B, img_hw=(h, w), out_hw=(float(h), float(w)), out=buf)
        return buf.to_instances([tuple(s) for s in sizes]), {}

    __call__ = forward


def build_roi_heads(cfg, input_shape=None):
    return ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)


class ImageList:
    """structures/image_list.py:8-102 for same-size images: ``tensor`` [N, C, Hpad, Wpad] and ``image_sizes``."""

    def __init__(self, tensor, image_sizes):
        self.tensor, self.image_sizes = tensor, image_sizes

    @staticmethod
    def from_tensors(tensors, size_divisibility=0, pad_value=0.0):
        h, w = tensors[0].shape[-2:]
        if any(t.shape[-2:] != (h, w) for t in tensors):
            raise RuntimeError("probenb200.ImageList: ragged batches are not supported (FLIR / KAIST frames share one size)")
        d = max(1, size_divisibility)
        H, W = (h + d - 1) // d * d, (w + d - 1) // d * d
        out = tensors[0].new_full((len(tensors), tensors[0].shape[0], H, W), pad_value)
        for i, t in enumerate(tensors):
            out[i, :, :h, :w] = t
        return ImageList(out, [(h, w)] * len(tensors))

    def __len__(self):
        return len(self.image_sizes)


def detector_postprocess(results, output_height, output_width):
    """modeling/postprocessing.py:8-52: rescale to the requested output size, clip, drop empty boxes."""
    sx, sy = output_width / results.image_size[1], output_height / results.image_size[0]
    out = Instances((output_height, output_width), **results.get_fields())
    b = out.pred_boxes.tensor.clone()
    b[:, 0::2] *= sx
    b[:, 1::2] *= sy
    b[:, 0::2] = b[:, 0::2].clamp(min=0, max=output_width)
    b[:, 1::2] = b[:, 1::2].clamp(min=0, max=output_height)
    out.pred_boxes = Boxes(b)
    keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    return out[keep]


@META_ARCH_REGISTRY.register()
class GeneralizedRCNN:
    """meta_arch/rcnn.py:26-302 at inference, assembled from the registries exactly like the reference (:39-69)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.device = torch.device(cfg.MODEL.DEVICE)
        self.backbone = build_backbone(cfg)
        self.proposal_generator = build_proposal_generator(cfg, self.backbone.output_shape())
        self.roi_heads = build_roi_heads(cfg, self.backbone.output_shape())
        self.pixel_mean = torch.tensor(cfg.MODEL.PIXEL_MEAN, device=self.device).view(-1, 1, 1)
        self.pixel_std = torch.tensor(cfg.MODEL.PIXEL_STD, device=self.device).view(-1, 1, 1)

    def eval(self):
        return self

    def preprocess_image(self, batched_inputs):
        """rcnn.py:269-286."""
        images = [(x["image"].to(self.device).float() - self.pixel_mean) / self.pixel_std for x in batched_inputs]
        canvas = self.cfg.ENGINE.CANVAS
        il = ImageList.from_tensors(images, self.backbone.size_divisibility)
        if tuple(il.tensor.shape[-2:]) != tuple(canvas):  # the engine's scratch is sized for one canvas
            t = il.tensor.new_zeros((len(images), il.tensor.shape[1], canvas[0], canvas[1]))
            t[:, :, : il.tensor.shape[2], : il.tensor.shape[3]] = il.tensor
            il = ImageList(t, il.image_sizes)
        return il

    def inference(self, batched_inputs):
        images = self.preprocess_image(batched_inputs)
        features = self.backbone(images.tensor)
        proposals, _ = self.proposal_generator(images, features, None)
        results, _ = self.roi_heads(images, features, proposals, None)
        out = []
        for r, inp, size in zip(results, batched_inputs, images.image_sizes):
            out.append({"instances": detector_postprocess(r, inp.get("height", size[0]), inp.get("width", size[1]))})
        return out

    forward = inference
    __call__ = inference


def build_model(cfg):
    """meta_arch/build.py:12-19."""
    return META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
