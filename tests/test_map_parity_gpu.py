"""Does bf16 tensor-core inference move COCO AP?  (north_star: "identical COCO mAP to two decimals";
FLIR_evaluation.py:496-563, fast_rcnn.py:86-147, demo_probEn.py:198-298.)

The harness (tests/golden/make_map_harness.py, which also documents how the detector was fitted) stores the fp32
oracle's detections of two R50-FPN detectors on 256 held-out synthetic RGB+thermal pairs with ground truth, and their
ProbEn fusion.  Here the B200 engine runs the SAME models on the SAME uint8 frames at the benchmarked shape
(512x640 frames -> 800x1000 -> 800x1024 canvas, batch 16); its detections and their fusion are scored with the
COCOeval restatement against the same ground truth.

Tolerances (0..100 scale):
  * ProbEn-fused output (what the pipeline delivers): |AP_gpu - AP_oracle| < 0.75 for COCO mAP = AP@[.5:.95].  Measured on
    a B200 with the shipped kernels: 0.385 on the 256-scene set (60.60 vs 60.99: both 0.61 to two decimals); over five
    kernel generations that differ ONLY in fp32 summation order (separate / chained conv1, 128- / 256-wide chained tiles):
    0.01, 0.08, 0.39, 0.68 - always 0.61 at two decimals, never a systematic sign (profiles/r02_map_parity*.json).
  * each model alone: < 1.5.  The harness detector is chaotic at the level of single detections: changing only the fp32
    summation ORDER of one conv (separate conv1 launch vs the chained conv3 -> conv1 kernel, same operands, same precision)
    moved one model's AP by 0.7 on 96 scenes, while the fused AP moved by 0.07; tests/golden/bisect_bf16.py shows the same
    for the fp32 oracle re-run with one stage in the engine's arithmetic (no sign, +-1 AP on 16 scenes).  Late fusion
    averages that noise away, which is why the fused number is the pinned one.
  * the single-threshold slices AP50 / AP75: < 1.5 and < 4.0 (borderline boxes flip across ONE IoU threshold; AP75 moved by up to
    3.9 between kernel generations).
  * >= 85 % of the oracle's detections have a same-class GPU detection with IoU > 0.5 (measured 0.88 .. 0.92; the rest are
    low-score duplicates / false positives whose survival of the 0.5 score threshold or of NMS flips either way); the
    strict rate (IoU > 0.9 and |score diff| < 0.05) is reported: which of several near-duplicate candidates survives NMS
    is chaotic under any perturbation (0.62 for the bf16-emulating oracle against the fp32 oracle, same as the engine).
The measured deltas are written to gpurun_out/map_parity.json (committed under profiles/).
"""
import json
import os

import numpy as np
import pytest
import torch

import make_map_harness as H
from probenb200 import detector, evaluation, pipeline

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _coco_gt(gold):
    anns = []
    off = gold["gt_offsets"]
    for i in range(len(off) - 1):
        for g in gold["gt"][off[i]: off[i + 1]]:
            w, h = float(g[2] - g[0]), float(g[3] - g[1])
            anns.append({"id": len(anns) + 1, "image_id": i, "category_id": int(g[4]), "bbox": [float(g[0]), float(g[1]), w, h],
                         "area": w * h, "iscrowd": 0})
    return anns


def _ap(anns, dets, n_img):
    if not dets:
        return {"AP": 0.0, "AP50": 0.0, "AP75": 0.0}
    return evaluation.COCOBBoxEval(anns, dets, image_ids=list(range(n_img))).evaluate()


def _match_rate(want, got, iou_thr=0.9, ds=0.05):
    """fraction of `want` rows (boxes, scores, classes per image) with a same-class `got` box at IoU > iou_thr, |ds| < ds"""
    from torchvision.ops import box_iou
    hit = tot = 0
    for (wb, ws, wc), (gb, gs, gc) in zip(want, got):
        tot += len(ws)
        if len(ws) == 0 or len(gs) == 0:
            continue
        iou = box_iou(torch.as_tensor(wb, dtype=torch.float32).reshape(-1, 4), torch.as_tensor(gb, dtype=torch.float32).reshape(-1, 4))
        ok = (iou > iou_thr) & (torch.as_tensor(wc).reshape(-1, 1) == torch.as_tensor(gc).reshape(1, -1)) & \
             ((torch.as_tensor(ws, dtype=torch.float32).reshape(-1, 1) - torch.as_tensor(gs, dtype=torch.float32).reshape(1, -1)).abs() < ds)
        hit += int(ok.any(dim=1).sum())
    return hit / max(1, tot), tot


def test_bf16_engine_keeps_coco_ap_of_fp32_oracle():
    gold = np.load(os.path.join(GOLD, "map_harness_oracle.npz"))
    heads = np.load(os.path.join(GOLD, "map_harness_heads.npz"))
    n_img, B = H.N_EVAL, 16
    anns = _coco_gt(gold)
    scenes = [H.scene(i, 1) for i in range(n_img)]
    for i, sc in enumerate(scenes):  # the generator is seeded: the stored ground truth must be reproduced exactly
        assert np.array_equal(sc[2], gold["gt"][gold["gt_offsets"][i]: gold["gt_offsets"][i + 1]])
    frames = [np.stack([sc[m] for sc in scenes]) for m in range(2)]
    dets = [detector.Detector(H.fitted_state_dict(m, heads), depth=50, num_classes=3, max_batch=B, canvas=(800, 1024)) for m in range(2)]
    pipe = pipeline.ProbEnPipeline(dets, ("probEn", "v-avg"), frame_size=H.FRAME_HW)
    per_model = [[], []]
    fused = []
    for i0 in range(0, n_img, B):
        dev = [torch.from_numpy(f[i0:i0 + B]).cuda() for f in frames]
        out = pipe.forward_device(dev, net_hw=H.NET_HW)
        torch.cuda.synchronize()
        fused += pipeline.FusedOutput.split(out.flat, B, 2)
        for m in range(2):
            per_model[m] += pipe.dets[m].to_instances([H.FRAME_HW] * B)
    report = {"images": n_img, "gt_boxes": len(anns), "shape": "512x640 -> 800x1000 (canvas 800x1024), batch 16, R50-FPN x2"}
    # ---- single models
    for m in range(2):
        off = gold["m%d_offsets" % m]
        want = [(gold["m%d_boxes" % m][off[i]: off[i + 1]], gold["m%d_scores" % m][off[i]: off[i + 1]], gold["m%d_classes" % m][off[i]: off[i + 1]])
                for i in range(n_img)]
        got = [(inst.pred_boxes.tensor.numpy(), inst.scores.numpy(), inst.pred_classes.numpy()) for inst in per_model[m]]
        d_want, d_got = [], []
        for i in range(n_img):
            d_want += evaluation.instances_to_coco_json(*want[i], i)
            d_got += evaluation.instances_to_coco_json(*got[i], i)
        a_w, a_g = _ap(anns, d_want, n_img), _ap(anns, d_got, n_img)
        rate, tot = _match_rate(want, got)
        loose, _ = _match_rate(want, got, 0.5, 2.0)
        report["model%d" % m] = {"same_object_rate": loose, "oracle_AP": a_w["AP"], "gpu_AP": a_g["AP"], "oracle_AP50": a_w["AP50"], "gpu_AP50": a_g["AP50"],
                                 "oracle_AP75": a_w["AP75"], "gpu_AP75": a_g["AP75"],
                                 "oracle_detections": tot, "gpu_detections": int(sum(len(g[1]) for g in got)), "box_match_rate": rate}
    # ---- ProbEn fusion of the two models
    off = gold["fused_offsets"]
    want = [(gold["fused_boxes"][off[i]: off[i + 1]], gold["fused_scores"][off[i]: off[i + 1]], gold["fused_classes"][off[i]: off[i + 1]])
            for i in range(n_img)]
    got = [(np.zeros((0, 4)), np.zeros(0), np.zeros(0)) if f is None else tuple(t.numpy() for t in f) for f in fused]
    d_want, d_got = [], []
    for i in range(n_img):
        d_want += evaluation.instances_to_coco_json(*want[i], i)
        d_got += evaluation.instances_to_coco_json(*got[i], i)
    a_w, a_g = _ap(anns, d_want, n_img), _ap(anns, d_got, n_img)
    rate, tot = _match_rate(want, got)
    loose, _ = _match_rate(want, got, 0.5, 2.0)
    report["proben_fused"] = {"same_object_rate": loose, "oracle_AP": a_w["AP"], "gpu_AP": a_g["AP"], "oracle_AP50": a_w["AP50"], "gpu_AP50": a_g["AP50"],
                              "oracle_AP75": a_w["AP75"], "gpu_AP75": a_g["AP75"],
                              "oracle_detections": tot, "gpu_detections": int(sum(len(g[1]) for g in got)), "box_match_rate": rate}
    for k in ("model0", "model1", "proben_fused"):
        r = report[k]
        r["abs_dAP"], r["abs_dAP50"] = abs(r["gpu_AP"] - r["oracle_AP"]), abs(r["gpu_AP50"] - r["oracle_AP50"])
        r["abs_dAP75"] = abs(r["gpu_AP75"] - r["oracle_AP75"])
        r["mAP_two_decimals"] = ["%.2f" % (r["oracle_AP"] / 100), "%.2f" % (r["gpu_AP"] / 100)]
    print(json.dumps(report, indent=1))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(report, open(os.path.join(ROOT, "gpurun_out", "map_parity.json"), "w"), indent=1)
    except OSError:
        pass
    for k in ("model0", "model1", "proben_fused"):
        r = report[k]
        assert r["oracle_AP"] > 5.0, (k, r)                      # the harness model must actually detect something
        assert r["abs_dAP"] < (0.75 if k == "proben_fused" else 1.5), (k, r)  # 0..100 scale
        assert r["abs_dAP50"] < 1.5 and r["abs_dAP75"] < 4.0, (k, r)
        assert r["same_object_rate"] >= 0.85, (k, r)
    fused = report["proben_fused"]
    assert fused["mAP_two_decimals"][0] == fused["mAP_two_decimals"][1], fused  # "identical COCO mAP to two decimals"
