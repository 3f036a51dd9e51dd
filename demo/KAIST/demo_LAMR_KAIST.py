#!/usr/bin/env python
"""Drop-in for the reference's demo/KAIST/demo_LAMR_KAIST.py: KAIST test-set inference with a Faster R-CNN R50-FPN
person detector (NUM_CLASSES = 1, score threshold 0.5, :45-51) and the two output files of the reference:

* ``<outfolder>/KAIST_<method>_gnll.txt`` - one line per detection ``frame,x,y,w,h,score`` with 1-based frame numbers in
  split-file order and xywh boxes (:127-142), the input format of the KAIST log-average-miss-rate evaluation script
  (external to the reference repository, :88,146);
* ``<outfolder>/KAIST_<method>_variance.npz`` - the predicted box variances per frame (:91-92,124-125,145).

The reference hard-codes its paths and the modality in the file (:21-28,90-94); here they are flags:

    python demo/KAIST/demo_LAMR_KAIST.py --dataset_path /data/KAIST/test --split_file test-all-20.txt \
        --fusion_method middle_fusion --model_path out_model_middle_fusion.pth [--outfolder out/] [--batch 8]

Frames are ``<dataset_path>/<set>/<V>/{lwir,visible}/<frame>.jpg`` for every ``set/V/frame`` line of the split file
(:100-106).  Decode (nvJPEG), input assembly (:108-123; visible and lwir frames have the same size) and the predictor's
resize run on the GPU.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from probenb200 import detector, weights  # noqa: E402
from probenb200 import io as pio  # noqa: E402


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset_path", required=True)
    ap.add_argument("--split_file", required=True)
    ap.add_argument("--fusion_method", default="middle_fusion", choices=["rgb_only", "thermal_only", "early_fusion", "middle_fusion"])
    ap.add_argument("--model_path", required=True)
    ap.add_argument("--outfolder", default="out/box_predictions/KAIST/")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--depth", type=int, default=50, choices=[50, 101])
    return ap.parse_args(argv)


def lamr_lines(frame_no, boxes_xyxy, scores):
    """The reference's writer (:131-141): numpy float32 values through ``str`` (shortest round-trip repr)."""
    lines = []
    for b, s in zip(np.asarray(boxes_xyxy, np.float32), np.asarray(scores, np.float32)):
        b = b.copy()
        b[2] -= b[0]
        b[3] -= b[1]
        lines.append(str(frame_no) + "," + ",".join(str(c) for c in b) + "," + str(s) + "\n")
    return lines


def main(argv=None):
    args = parse(argv)
    method = args.fusion_method
    frames = [l.strip() for l in open(args.split_file) if l.strip()]
    os.makedirs(args.outfolder, exist_ok=True)
    out_txt = os.path.join(args.outfolder, "KAIST_" + method + "_gnll.txt")
    out_npz = os.path.join(args.outfolder, "KAIST_" + method + "_variance.npz")

    def paths(entry):
        s, v, n = entry.split("/")[:3]
        d = os.path.join(args.dataset_path, s, v)
        return os.path.join(d, "visible", n + ".jpg"), os.path.join(d, "lwir", n + ".jpg")

    dec = pio.JpegDecoder()
    h0, w0, _ = dec.image_info(open(paths(frames[0])[1], "rb").read())
    net_hw = detector.resize_shortest_edge_shape(h0, w0)
    canvas = ((net_hw[0] + 31) // 32 * 32, (net_hw[1] + 31) // 32 * 32)
    det = detector.Detector(weights.load_checkpoint(args.model_path), depth=args.depth, num_classes=1, max_batch=args.batch,
                            canvas=canvas, score_thresh=0.5, **detector.fusion_method_config(method))
    var_dict = {}
    with open(out_txt, "w") as f:
        for i0 in range(0, len(frames), args.batch):
            chunk = frames[i0:i0 + args.batch]
            rgb_files, th_files = zip(*[paths(e) for e in chunk])
            x = pio.load_pair_batch(dec, list(rgb_files), list(th_files), method)
            res = det.forward_frames_device(x, net_hw, round_u8=x.shape[3] == 3).to_instances([(h0, w0)] * len(chunk))
            for j, inst in enumerate(res):
                var_dict[i0 + j + 1] = inst.vars.numpy()
                f.writelines(lamr_lines(i0 + j + 1, inst.pred_boxes.tensor.numpy(), inst.scores.numpy()))
    np.savez(out_npz, vars=np.array(var_dict, dtype=object))
    print("saved", out_txt, out_npz)
    return out_txt, out_npz


if __name__ == "__main__":
    main()
