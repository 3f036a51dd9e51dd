#!/usr/bin/env python
"""Drop-in for the reference's demo/FLIR/demo_FLIR_save_predictions.py (same flags, same output file / schema):

    python demo/FLIR/demo_FLIR_save_predictions.py --dataset_path /path/to/FLIR/val \
        --fusion_method thermal_only --model_path out_model_thermal_only.pth

Runs the B200 detector engine over the validation pairs and writes ``<outfolder>/val_<method>_predictions.json``
with the keys of demo_FLIR_save_predictions.py:166-176 (image, boxes, scores, classes, image_id, class_logits,
probs, vars; detections with class > 2 dropped, :148-164).  Rows follow ``data['images']`` like the reference's
loop (:93-99); the ``image`` column repeats its quirk of listing ``os.listdir(RGB)`` names by position (:42,157).

The input side runs on the GPU: the JPEG files are read as bytes, decoded by nvJPEG into HBM,
the RGB frame is resized to the thermal frame's size with OpenCV's 8-bit bilinear arithmetic and the 3-/4-/6-channel
input is assembled on the device (``probenb200.io``); DefaultPredictor's ResizeShortestEdge is fused into the
detector's stem staging.  Extra flags (not in the reference): ``--batch N`` frames per launch, ``--depth {50,101}``,
``--cpu_decode`` to read and assemble the frames with cv2 exactly as the reference does - REQUIRED for bit-level decoder
parity with the reference: nvJPEG and libjpeg-turbo differ by 1-2 grey levels.  ``rgb_only`` runs the 80-class COCO
zoo model from ``trained_models/Detectron2_pretrained/model_final_f6e8b1.pkl`` like the reference (falling back to
``--model_path`` when that file is absent); ``.pkl`` (model-zoo format) and ``.pth`` checkpoints are both accepted.
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


# rgb_only: the COCO model-zoo R101-FPN 3x checkpoint, 80 classes (demo_FLIR_save_predictions.py:58-60)
ZOO_RGB_MODEL = "trained_models/Detectron2_pretrained/model_final_f6e8b1.pkl"


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


def extra_flags(argv):
    """Flags this CLI adds to the reference's (stripped before the reference parser sees the command line)."""
    import argparse
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--depth", type=int, default=101, choices=[50, 101])
    ap.add_argument("--cpu_decode", action="store_true")
    return ap.parse_known_args(argv)


def save_predictions(args, batch=8, depth=101, cpu_decode=False):
    from probenb200 import io as pio
    val_folder = args.dataset_path
    val_json_path = val_folder + "/FLIR_thermal_RGBT_pairs_val.json"
    rgb_path, t_path = val_folder + "/RGB/", val_folder + "/thermal_8_bit/"
    method = args.fusion_method
    print("==========================")
    print("model:", method)
    print("==========================")
    data = json.load(open(val_json_path, "r"))
    stems = [im["file_name"].split("/")[1].split(".")[0] for im in data["images"]]
    name_to_id = {st: im["id"] for st, im in zip(stems, data["images"])}
    files_names = [f for f in listdir(rgb_path) if isfile(join(rgb_path, f))]
    if not os.path.exists(args.outfolder):
        os.mkdir(args.outfolder)
    mcfg = detector.fusion_method_config(method)
    K = 80 if method == "rgb_only" else 3
    model_path = args.model_path
    if method == "rgb_only" and os.path.isfile(ZOO_RGB_MODEL):
        model_path = ZOO_RGB_MODEL  # the reference ignores --model_path for rgb_only (:58-60)
    sd = weights.load_checkpoint(model_path, num_classes=K)
    print("model loaded:", model_path)
    decoder = None if cpu_decode else pio.JpegDecoder()

    def load_batch(names):
        """uint8 device frames [B, H, W, C] of the pairs `names` (file stems), C = 3 / 4 / 6."""
        if cpu_decode:
            return torch.from_numpy(np.stack([load_frame(t_path, rgb_path, n + ".jpg", method).astype(np.uint8) for n in names])).cuda()
        return pio.load_pair_batch(decoder, [join(rgb_path, n + ".jpg") for n in names], [join(t_path, n + ".jpeg") for n in names],
                                   method)

    first = load_batch(stems[:1])
    h0, w0 = int(first.shape[1]), int(first.shape[2])
    net_hw = detector.resize_shortest_edge_shape(h0, w0)
    canvas = ((net_hw[0] + 31) // 32 * 32, (net_hw[1] + 31) // 32 * 32)
    det = detector.Detector(sd, depth=depth, num_classes=K, max_batch=batch, canvas=canvas,
                            score_thresh=0.5, **mcfg)
    out = {k: [] for k in ("image", "boxes", "scores", "classes", "image_id", "class_logits", "probs", "vars")}
    for i0 in range(0, len(stems), batch):
        names = stems[i0:i0 + batch]
        frames = load_batch(names)
        # 3-channel uint8 frames take Pillow's resize in the reference, 4-/6-channel arrays cv2's float path
        res = det.forward_frames_device(frames, net_hw, round_u8=frames.shape[3] == 3).to_instances([(h0, w0)] * len(names))
        for j, (n, inst) in enumerate(zip(names, res)):
            inst = inst[inst.pred_classes <= 2]
            out["image"].append(files_names[i0 + j] if i0 + j < len(files_names) else n + ".jpg")
            out["image_id"].append(name_to_id[n])
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
    # the same columns as a binary columnar file; demo_probEn.py prefers it when present
    from probenb200 import detfile
    detfile.DetFile.from_json_dict(out, K=K).save(path[:-5] + ".pedet")
    return path


if __name__ == "__main__":
    extra, rest = extra_flags(sys.argv[1:])
    save_predictions(config_parser(rest), batch=extra.batch, depth=extra.depth, cpu_decode=extra.cpu_decode)
