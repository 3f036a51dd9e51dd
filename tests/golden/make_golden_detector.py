"""Generates tests/golden/detector_golden.npz: outputs of the UNMODIFIED reference GeneralizedRCNN (loaded by file
path with infrastructure stubs, ref_loader.load_reference_detector_modules) on seeded inputs and seeded synthetic
weights (probenb200.weights.random_state_dict, reference parameter names, strict key match) for the three
input formats of demo_FLIR_save_predictions.py:58-73.  Run in the build container:
    python tests/golden/make_golden_detector.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_loader  # noqa: E402
from probenb200 import weights  # noqa: E402

CASES = {
    # name: (in_channels, input_format, pixel_mean, middle_fusion, weight seed)
    "thermal_only": (3, "BGR", (103.530, 116.280, 123.675), False, 1),
    "early_fusion": (4, "BGRT", (103.530, 116.280, 123.675, 135.438), False, 2),
    "middle_fusion": (6, "BGRTTT", (103.530, 116.280, 123.675, 135.438, 135.438, 135.438), True, 3),
}
IMG_HW = (160, 200)
OUT_HW = (128, 160)


def case_inputs(name):
    c = CASES[name][0]
    g = torch.Generator().manual_seed(100 + c)
    return [torch.rand(c, *IMG_HW, generator=g) * 255 for _ in range(2)]


def main():
    ns = ref_loader.load_reference_detector_modules()
    sys.modules.update(ns.modules)
    out = {}
    for name, (c, fmt, mean, mid, seed) in CASES.items():
        cfg = ns.make_cfg(depth=50, num_classes=3, in_channels=c, input_format=fmt, pixel_mean=mean)
        model = ns.GeneralizedRCNN(cfg).eval()
        sd = weights.random_state_dict(50, 3 if mid else c, 3, seed=seed, middle_fusion=mid)
        if mid:  # backbone_2 exists in the module tree but is unused at inference (rcnn.py:243-244)
            sd.update({k.replace("backbone.", "backbone_2.", 1): v for k, v in list(sd.items()) if k.startswith("backbone.")})
        missing = model.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys, missing.unexpected_keys
        assert all("anchor_generator" in k for k in missing.missing_keys), missing.missing_keys
        for i, img in enumerate(case_inputs(name)):
            with torch.no_grad():  # the reference runs batch 1 (demo_FLIR_save_predictions.py:93-133)
                inst = model([{"image": img, "height": OUT_HW[0], "width": OUT_HW[1]}])[0]["instances"]
            key = "%s/%d/" % (name, i)
            out[key + "boxes"] = inst.pred_boxes.tensor.numpy()
            out[key + "scores"] = inst.scores.numpy()
            out[key + "classes"] = inst.pred_classes.numpy()
            out[key + "class_logits"] = inst.class_logits.numpy()
            out[key + "probs"] = inst.prob_score.numpy()
            out[key + "vars"] = inst.vars.numpy()
            print(name, i, "detections:", len(inst))
    path = os.path.join(HERE, "detector_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
