"""Load pieces of the UNMODIFIED reference (/root/reference) by file path with stub modules.

Only usable in the build container (the GPU box has no /root/reference).  Used by the
``make_golden_*.py`` scripts to generate the committed fixtures and by tests marked
``needs_reference`` to pin the oracle against the real reference code (SURVEY.md §8c recipe).
Nothing under the product package imports this file.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("PROBEN_REFERENCE_ROOT", "/root/reference")


def have_reference():
    return os.path.isfile(os.path.join(REF, "demo", "FLIR", "demo_probEn.py"))


def _load(dotted, relpath):
    spec = importlib.util.spec_from_file_location(dotted, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    return mod


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_PROBEN = None


def load_reference_proben():
    """Returns the reference's demo/FLIR/demo_probEn.py as a module (fusion, nms_bayesian, ...)."""
    global _PROBEN
    if _PROBEN is not None:
        return _PROBEN
    saved = {k: v for k, v in sys.modules.items() if k == "detectron2" or k.startswith("detectron2.")}
    try:
        _stub("detectron2")
        _stub("detectron2.config", get_cfg=lambda: None)
        _stub("detectron2.data", DatasetCatalog=None, MetadataCatalog=None)
        _stub("detectron2.data.datasets", register_coco_instances=lambda *a, **k: None)
        _stub("detectron2.structures", Instances=None, Boxes=None)
        _stub("detectron2.evaluation", FLIREvaluator=None)
        _stub("detectron2.layers")
        _stub("detectron2.utils")
        _stub("detectron2.utils.opt", config_parser=lambda *a, **k: None)
        _load("detectron2.layers.nms", "detectron2/layers/nms.py")
        _PROBEN = _load("_reference_demo_probEn", "demo/FLIR/demo_probEn.py")
    finally:
        for k in [k for k in sys.modules if k == "detectron2" or k.startswith("detectron2.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return _PROBEN


# ------------------------------------------------------------------------------------------------------
# The reference detector: detectron2's GeneralizedRCNN built from the reference's OWN modeling files, loaded
# by path under their real dotted names.  Only infrastructure is stubbed (registries, config node, fvcore
# initialisers, logging); the single substituted arithmetic is detectron2.layers.ROIAlign ->
# torchvision.ops.roi_align (detectron2._C does not build against torch 2.11; identical on the reference's
# golden tables).
class _CfgNode(dict):
    def __init__(self, init=None, **kw):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = _CfgNode(v) if isinstance(v, dict) and not isinstance(v, _CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy
        return copy.deepcopy(self)

    def merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge(v)
            else:
                self[k] = _CfgNode(v) if isinstance(v, dict) else v


class _Registry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self[o.__name__] = o
                return o
            return deco
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        return self[name]


def _configurable(init_func):
    import functools

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        if args and isinstance(args[0], _CfgNode):
            init_func(self, **type(self).from_config(*args, **kwargs))
        elif isinstance(kwargs.get("cfg"), _CfgNode):
            init_func(self, **type(self).from_config(*args, **kwargs))
        else:
            init_func(self, *args, **kwargs)
    return wrapped


_DET = None


def load_reference_detector_modules():
    """Loads the reference modeling stack; returns a namespace with GeneralizedRCNN, default cfg factory, ..."""
    global _DET
    if _DET is not None:
        return _DET
    import logging
    import torch
    import torchvision
    import yaml

    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("detectron2", "fvcore")}
    for k in saved:
        del sys.modules[k]
    ns = types.SimpleNamespace()
    try:
        noop = lambda *a, **k: None
        _stub("fvcore")
        _stub("fvcore.common")
        _stub("fvcore.common.registry", Registry=_Registry)
        _stub("fvcore.nn", smooth_l1_loss=noop)
        wi = _stub("fvcore.nn.weight_init")

        def c2_msra_fill(m):
            torch.nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                torch.nn.init.constant_(m.bias, 0)

        def c2_xavier_fill(m):
            torch.nn.init.kaiming_uniform_(m.weight, a=1)
            if m.bias is not None:
                torch.nn.init.constant_(m.bias, 0)
        wi.c2_msra_fill, wi.c2_xavier_fill = c2_msra_fill, c2_xavier_fill
        sys.modules["fvcore.nn"].weight_init = wi

        _stub("detectron2")
        _stub("detectron2.utils")
        _stub("detectron2.utils.registry", Registry=_Registry)
        _stub("detectron2.utils.events", get_event_storage=noop)
        _stub("detectron2.utils.logger", log_first_n=noop)
        _stub("detectron2.utils.memory", retry_if_cuda_oom=lambda f: f)
        _stub("detectron2.utils.comm", get_world_size=lambda: 1)
        _stub("detectron2.utils.env", TORCH_VERSION=(2, 0))
        _stub("detectron2.config", configurable=_configurable)
        _stub("detectron2.config.config", CfgNode=_CfgNode)
        defaults = _load("detectron2.config.defaults", "detectron2/config/defaults.py")

        lay = _stub("detectron2.layers")
        wr = _load("detectron2.layers.wrappers", "detectron2/layers/wrappers.py")
        ss = _load("detectron2.layers.shape_spec", "detectron2/layers/shape_spec.py")
        nm = _load("detectron2.layers.nms", "detectron2/layers/nms.py")
        for n in ("cat", "Conv2d", "Linear", "interpolate", "BatchNorm2d", "ConvTranspose2d"):
            if hasattr(wr, n):
                setattr(lay, n, getattr(wr, n))
        lay.ShapeSpec = ss.ShapeSpec
        lay.batched_nms = nm.batched_nms
        bn = _load("detectron2.layers.batch_norm", "detectron2/layers/batch_norm.py")
        lay.FrozenBatchNorm2d, lay.get_norm, lay.NaiveSyncBatchNorm = bn.FrozenBatchNorm2d, bn.get_norm, bn.NaiveSyncBatchNorm
        lay.DeformConv = lay.ModulatedDeformConv = lay.ROIAlignRotated = None
        lay.paste_masks_in_image = noop

        class ROIAlign(torch.nn.Module):  # detectron2/layers/roi_align.py:46-105 semantics via torchvision
            def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
                super().__init__()
                self.output_size, self.spatial_scale = output_size, spatial_scale
                self.sampling_ratio, self.aligned = sampling_ratio, aligned

            def forward(self, input, rois):
                return torchvision.ops.roi_align(input, rois, self.output_size, self.spatial_scale,
                                                 self.sampling_ratio, self.aligned)
        lay.ROIAlign = ROIAlign

        st = _stub("detectron2.structures")
        bx = _load("detectron2.structures.boxes", "detectron2/structures/boxes.py")
        ins = _load("detectron2.structures.instances", "detectron2/structures/instances.py")
        il = _load("detectron2.structures.image_list", "detectron2/structures/image_list.py")
        st.Boxes, st.BoxMode, st.pairwise_iou = bx.Boxes, bx.BoxMode, bx.pairwise_iou
        st.Instances, st.ImageList, st.RotatedBoxes = ins.Instances, il.ImageList, None

        _stub("detectron2.modeling")
        _load("detectron2.modeling.box_regression", "detectron2/modeling/box_regression.py")
        _load("detectron2.modeling.matcher", "detectron2/modeling/matcher.py")
        _load("detectron2.modeling.sampling", "detectron2/modeling/sampling.py")
        bb = _stub("detectron2.modeling.backbone")
        _load("detectron2.modeling.backbone.backbone", "detectron2/modeling/backbone/backbone.py")
        bbuild = _load("detectron2.modeling.backbone.build", "detectron2/modeling/backbone/build.py")
        _load("detectron2.modeling.backbone.resnet", "detectron2/modeling/backbone/resnet.py")
        _load("detectron2.modeling.backbone.fpn", "detectron2/modeling/backbone/fpn.py")
        bb.build_backbone = bbuild.build_backbone
        _load("detectron2.modeling.anchor_generator", "detectron2/modeling/anchor_generator.py")
        pg = _stub("detectron2.modeling.proposal_generator")
        PG = _Registry("PROPOSAL_GENERATOR")
        _stub("detectron2.modeling.proposal_generator.build", PROPOSAL_GENERATOR_REGISTRY=PG)
        _stub("detectron2.modeling.proposal_generator.proposal_utils", add_ground_truth_to_proposals=noop)
        _load("detectron2.modeling.proposal_generator.rpn_outputs", "detectron2/modeling/proposal_generator/rpn_outputs.py")
        _load("detectron2.modeling.proposal_generator.rpn", "detectron2/modeling/proposal_generator/rpn.py")
        pg.build_proposal_generator = lambda cfg, shape: PG.get(cfg.MODEL.PROPOSAL_GENERATOR.NAME)(cfg, shape)
        _load("detectron2.modeling.poolers", "detectron2/modeling/poolers.py")
        rh = _stub("detectron2.modeling.roi_heads")
        _stub("detectron2.modeling.roi_heads.keypoint_head", build_keypoint_head=noop)
        _stub("detectron2.modeling.roi_heads.mask_head", build_mask_head=noop)
        _load("detectron2.modeling.roi_heads.box_head", "detectron2/modeling/roi_heads/box_head.py")
        _load("detectron2.modeling.roi_heads.fast_rcnn", "detectron2/modeling/roi_heads/fast_rcnn.py")
        rhm = _load("detectron2.modeling.roi_heads.roi_heads", "detectron2/modeling/roi_heads/roi_heads.py")
        rh.build_roi_heads = rhm.build_roi_heads
        _load("detectron2.modeling.postprocessing", "detectron2/modeling/postprocessing.py")
        _stub("detectron2.modeling.meta_arch")
        _stub("detectron2.modeling.meta_arch.build", META_ARCH_REGISTRY=_Registry("META_ARCH"))
        _stub("detectron2.modeling.meta_arch.gaussian_blur", gaussian_blur=noop)
        rc = _load("detectron2.modeling.meta_arch.rcnn", "detectron2/modeling/meta_arch/rcnn.py")

        def make_cfg(depth=50, num_classes=3, in_channels=3, input_format="BGR", pixel_mean=None):
            cfg = defaults._C.clone()
            for rel in ("configs/Base-RCNN-FPN.yaml", "configs/COCO-Detection/faster_rcnn_R_50_FPN_3x.yaml"):
                y = yaml.safe_load(open(os.path.join(REF, rel)))
                y.pop("_BASE_", None)
                cfg.merge(y)
            cfg.MODEL.DEVICE = "cpu"
            cfg.MODEL.RESNETS.DEPTH = depth
            cfg.MODEL.ROI_HEADS.NUM_CLASSES = num_classes
            cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST = 0.5          # demo_FLIR_save_predictions.py:51-57
            cfg.MODEL.ROI_BOX_HEAD.OUTPUT_LOGITS = True
            cfg.MODEL.ROI_HEADS.ENABLE_GAUSSIANNLLOSS = True
            cfg.MODEL.BACKBONE.FREEZE_AT = 3
            cfg.INPUT.FORMAT = input_format
            cfg.INPUT.NUM_IN_CHANNELS = in_channels
            if pixel_mean is not None:
                cfg.MODEL.PIXEL_MEAN = list(pixel_mean)
                cfg.MODEL.PIXEL_STD = [1.0] * len(pixel_mean)
            return cfg

        ns.GeneralizedRCNN = rc.GeneralizedRCNN
        ns.make_cfg = make_cfg
        ns.modules = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("detectron2", "fvcore")}
        logging.getLogger("detectron2").setLevel(logging.ERROR)
        _DET = ns
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("detectron2", "fvcore")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return _DET


def load_reference_cocoeval():
    """Returns (COCO, COCOeval) of the reference's vendored detectron2/pycocotools (coco.py, cocoeval.py loaded by
    path).  Stubs: matplotlib (plotting only), and ``pycocotools._mask`` whose one function used by the bbox
    protocol, ``iou``, is the published maskApi.c ``bbIou`` restated in numpy (the Cython extension is not built
    in this image).  numpy >= 1.24 dropped the ``np.float`` alias cocoeval.py:379 uses; it is restored as float."""
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float

    def bb_iou(dt, gt, iscrowd):
        dt = np.asarray(dt, np.float64).reshape(-1, 4)
        gt = np.asarray(gt, np.float64).reshape(-1, 4)
        out = np.zeros((len(dt), len(gt)))
        for g in range(len(gt)):
            ga = gt[g, 2] * gt[g, 3]
            for d in range(len(dt)):
                da = dt[d, 2] * dt[d, 3]
                w = min(dt[d, 2] + dt[d, 0], gt[g, 2] + gt[g, 0]) - max(dt[d, 0], gt[g, 0])
                if w <= 0:
                    continue
                h = min(dt[d, 3] + dt[d, 1], gt[g, 3] + gt[g, 1]) - max(dt[d, 1], gt[g, 1])
                if h <= 0:
                    continue
                i = w * h
                out[d, g] = i / (da if iscrowd[g] else da + ga - i)
        return out

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.patches"):
        if name not in sys.modules:
            _stub(name, PatchCollection=None, Polygon=None)
    _stub("pycocotools")
    _stub("pycocotools._mask", iou=bb_iou, merge=None, frPyObjects=None)  # masks are never touched for bbox
    _stub("_ref_pycocotools")
    _load("_ref_pycocotools.mask", "detectron2/pycocotools/mask.py")
    coco = _load("_ref_pycocotools.coco", "detectron2/pycocotools/coco.py")
    cocoeval = _load("_ref_pycocotools.cocoeval", "detectron2/pycocotools/cocoeval.py")
    return coco.COCO, cocoeval.COCOeval
