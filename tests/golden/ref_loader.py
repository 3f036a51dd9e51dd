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
