"""Precision bisect for the mAP-parity harness (TEST INFRASTRUCTURE): which stage's bf16 arithmetic moves the detections?

Runs the harness detectors (make_map_harness.py) through the CPU oracle with ``detector_oracle.EMULATE`` set to one stage
group at a time - the oracle then restates the engine's arithmetic for those stages (BN folded, bf16 operands, fp32
accumulate, one bf16 rounding per stored activation) - and compares the detections with the fp32 oracle's stored ones:
box-level match rate (same class, IoU > 0.9, |score diff| < 0.05) and COCO AP on the first ``n`` evaluation scenes.

    python tests/golden/bisect_bf16.py [n_scenes=16] [model=1]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import make_map_harness as H  # noqa: E402
from oracle import detector_oracle as D  # noqa: E402
from probenb200 import evaluation  # noqa: E402

GROUPS = [("fp32", set()), ("stem+res2", {"stem", "res2"}), ("res3", {"res3"}), ("res4", {"res4"}), ("res5", {"res5"}), ("fpn", {"fpn"}),
          ("rpn", {"rpn"}), ("head", {"head"}), ("all (engine arithmetic)", {"stem", "res2", "res3", "res4", "res5", "fpn", "rpn", "head"})]


def main(n=16, m=1, heads_file=None, groups=GROUPS):
    import test_map_parity_gpu as T
    gold = np.load(os.path.join(HERE, "map_harness_oracle.npz"))
    heads = np.load(heads_file or os.path.join(HERE, "map_harness_heads.npz"))
    sd = H.fitted_state_dict(m, heads)
    cfg = D.DetCfg()
    anns = [a for a in T._coco_gt(gold) if a["image_id"] < n]
    frames = [H.resized_input(H.scene(i, 1)[m]) for i in range(n)]
    base = None
    rows = []
    for name, stages in groups:
        D.EMULATE = set(stages)
        got = []
        for i in range(n):
            r = D.detector_forward([frames[i]], [H.FRAME_HW], sd, cfg)[0]
            got.append((r["pred_boxes"].numpy(), r["scores"].numpy(), r["pred_classes"].numpy()))
        D.EMULATE = set()
        if base is None:
            base = got
        dets = []
        for i in range(n):
            dets += evaluation.instances_to_coco_json(*got[i], i)
        ap = T._ap(anns, dets, n)
        rate, tot = T._match_rate(base, got)
        rows.append((name, rate, ap["AP"], ap["AP50"], sum(len(g[1]) for g in got)))
        print("%-26s match %.3f  AP %.2f  AP50 %.2f  detections %d" % rows[-1], flush=True)
    return rows


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
