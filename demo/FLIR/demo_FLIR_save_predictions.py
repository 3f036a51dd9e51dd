#!/usr/bin/env python
"""Drop-in for the reference's demo/FLIR/demo_FLIR_save_predictions.py (same flags, same output file / schema):

    python demo/FLIR/demo_FLIR_save_predictions.py --dataset_path /path/to/FLIR/val \
        --fusion_method thermal_only --model_path out_model_thermal_only.pth

Runs the B200 detector engine over the validation pairs and writes ``<outfolder>/val_<method>_predictions.json``
with the keys of demo_FLIR_save_predictions.py:166-176 (image, boxes, scores, classes, image_id, class_logits,
probs, vars; detections with class > 2 dropped, :148-164).  Differences from the reference, all on the input
side (not yet on the GPU): frames are batched (``--batch``) instead of one by one.
"""
import json
import os
import sys
from os import listdir
from os.path import isfile, join

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from probenb200 import detector, weights  # noqa: E402
from probenb200.opt import config_parser  # noqa: E402


def load_frame(t_path, rgb_path, name, method):
    """Input assembly of demo_FLIR_save_predictions.py:100-121 (needs OpenCV)."""
    import cv2
    stem = name.split(".")[0]
    thermal = cv2.imread(join(t_path, stem + ".jpeg"))
    if method == "thermal_only":
        return thermal.astype(np.float32)
    rgb = cv2.imread(join(rgb_path, name))
    rgb = cv2.resize(rgb, (thermal.shape[1], thermal.shape[0]))  # default interpolation = bilinear
    if method == "rgb_only":
        return rgb.astype(np.float32)
    if method == "early_fusion":
        return np.concatenate([rgb, thermal[:, :, :1]], axis=2).astype(np.float32)
    return np.concatenate([rgb, thermal], axis=2).astype(np.float32)  # middle_fusion: BGRTTT


def resize_like_predictor(img, short=800, max_size=1333):
    """DefaultPredictor's ResizeShortestEdge (engine/defaults.py:186-190; cv2 bilinear as transform.py:81-99 uses
    for float / >3-channel inputs)."""
    import cv2
    h, w = img.shape[:2]
    nh, nw = detector.resize_shortest_edge_shape(h, w, short, max_size)
    return cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)


def save_predictions(args, batch=8, depth=101):
    val_folder = args.dataset_path
    val_json_path = val_folder + "/FLIR_thermal_RGBT_pairs_val.json"
    rgb_path, t_path = val_folder + "/RGB/", val_folder + "/thermal_8_bit/"
    method = args.fusion_method
    print("==========================")
    print("model:", method)
    print("==========================")
    data = json.load(open(val_json_path, "r"))
    name_to_id = {im["file_name"].split("/")[1].split(".")[0]: im["id"] for im in data["images"]}
    files = [f for f in listdir(rgb_path) if isfile(join(rgb_path, f))]
    if not os.path.exists(args.outfolder):
        os.mkdir(args.outfolder)
    mcfg = detector.fusion_method_config(method)
    sd = weights.load_checkpoint(args.model_path)
    print("model loaded:", args.model_path)
    first = resize_like_predictor(load_frame(t_path, rgb_path, files[0], method))
    canvas = ((first.shape[0] + 31) // 32 * 32, (first.shape[1] + 31) // 32 * 32)
    det = detector.Detector(sd, depth=depth, num_classes=80 if method == "rgb_only" else 3, max_batch=batch, canvas=canvas,
                            score_thresh=0.5, **mcfg)
    out = {k: [] for k in ("image", "boxes", "scores", "classes", "image_id", "class_logits", "probs", "vars")}
    for i0 in range(0, len(files), batch):
        names = files[i0:i0 + batch]
        frames = [load_frame(t_path, rgb_path, n, method) for n in names]
        h0, w0 = frames[0].shape[:2]
        x = torch.stack([torch.from_numpy(resize_like_predictor(f).transpose(2, 0, 1).copy()) for f in frames])
        res = det.forward_device(x.cuda(), (h0, w0)).to_instances([(h0, w0)] * len(names))
        for n, inst in zip(names, res):
            keep = inst.pred_classes <= 2
            inst = inst[keep]
            out["image"].append(n)
            out["image_id"].append(name_to_id.get(n.split(".")[0], -1))
            out["boxes"].append(inst.pred_boxes.tensor.tolist())
            out["scores"].append(inst.scores.tolist())
            out["classes"].append(inst.pred_classes.tolist())
            out["class_logits"].append(inst.class_logits.tolist())
            out["probs"].append(inst.prob_score.tolist())
            out["vars"].append(inst.vars.tolist())
    path = join(args.outfolder, "val_" + method + "_predictions.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=2)
    print("saved", path)


if __name__ == "__main__":
    save_predictions(config_parser())
