#!/usr/bin/env python
"""ProbEn late fusion of KAIST person detectors (BASELINE.json configs[3]).

The reference ships no KAIST ProbEn script (demo_probEn.py hard-codes FLIR: K = 3, 640 x 512); the paper's KAIST
numbers come from the same algorithm with one foreground class, i.e. probability rows ``[p, 1 - p]``
(bayesian_fusion_multiclass, demo_probEn.py:32-42, with K = 1).  Inputs are the files demo_LAMR_KAIST.py writes per
modality - ``KAIST_<method>_gnll.txt`` (``frame,x,y,w,h,score``) and ``KAIST_<method>_variance.npz`` - and the output is
one fused txt in the same format, ready for the KAIST log-average-miss-rate evaluation script:

    python demo/KAIST/demo_probEn_KAIST.py --prediction_path out/box_predictions/KAIST/ \
        --methods thermal_only rgb_only [--score_fusion probEn --box_fusion v-avg] [--out fused.txt]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from probenb200 import detfile, fusion  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from demo_LAMR_KAIST import lamr_lines  # noqa: E402


def read_modality(folder, method, n_frames=None):
    """txt + npz of one modality -> ``detfile.DetFile`` with K = 1 (class 0 = person, prob = score)."""
    txt = os.path.join(folder, "KAIST_" + method + "_gnll.txt")
    rows = np.loadtxt(txt, delimiter=",", ndmin=2, dtype=np.float64) if os.path.getsize(txt) else np.zeros((0, 6))
    var = np.load(os.path.join(folder, "KAIST_" + method + "_variance.npz"), allow_pickle=True)["vars"].item()
    n_frames = n_frames or max([int(rows[:, 0].max()) if len(rows) else 0] + list(var.keys()))
    frame = rows[:, 0].astype(np.int64)
    order = np.argsort(frame, kind="stable")
    rows, frame = rows[order], frame[order]
    counts = np.bincount(frame, minlength=n_frames + 1)[1:n_frames + 1]
    offsets = np.zeros(n_frames + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    boxes = rows[:, 1:5].copy()
    boxes[:, 2] += boxes[:, 0]
    boxes[:, 3] += boxes[:, 1]
    scores = rows[:, 5]
    vs = np.concatenate([np.asarray(var.get(i + 1, np.zeros((0, 1)))).reshape(-1) for i in range(n_frames)]) if n_frames else np.zeros(0)
    if len(vs) != len(scores):
        raise ValueError("%s: %d detections but %d variances" % (txt, len(scores), len(vs)))
    logits = np.stack([np.log(np.clip(scores, 1e-12, 1)), np.log(np.clip(1 - scores, 1e-12, 1))], 1)
    return detfile.DetFile(1, np.arange(1, n_frames + 1), offsets, boxes, scores, np.zeros(len(scores), np.int32), logits,
                           scores.reshape(-1, 1), vs)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--prediction_path", required=True)
    ap.add_argument("--methods", nargs="+", default=["thermal_only", "rgb_only"])
    ap.add_argument("--score_fusion", default="probEn", choices=["avg", "max", "probEn"])
    ap.add_argument("--box_fusion", default="v-avg", choices=["avg", "s-avg", "v-avg", "argmax"])
    ap.add_argument("--img_w", type=float, default=640.0)
    ap.add_argument("--img_h", type=float, default=512.0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args(argv)
    files = [read_modality(args.prediction_path, m) for m in args.methods]
    n = max(f.n_images for f in files)
    files = [read_modality(args.prediction_path, m, n) for m in args.methods]
    packed = detfile.pack_models(files)
    buf = fusion.fuse_packed(fusion.to_device(packed), [args.score_fusion, args.box_fusion], img_w=args.img_w, img_h=args.img_h)
    out = args.out or os.path.join(args.prediction_path, "KAIST_probEn_%s_%s.txt" % (args.score_fusion, args.box_fusion))
    with open(out, "w") as f:
        for i, r in enumerate(fusion.unpack_results(packed, buf)):
            if r is None:
                continue
            keep = r[2] == 0  # the fused class can be background (K): not a person
            f.writelines(lamr_lines(i + 1, r[0][keep], r[1][keep]))
    print("saved", out)
    return out


if __name__ == "__main__":
    main()
