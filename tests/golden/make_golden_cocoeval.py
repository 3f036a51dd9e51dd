"""Generates tests/golden/cocoeval_golden.json: seeded synthetic ground truth + detections in the FLIR schema and the
numbers the UNMODIFIED reference evaluator produces for them (detectron2/pycocotools/cocoeval.py COCOeval 'bbox' as
driven by FLIR_evaluation.py:496-528, summary of _derive_coco_results :249-310).  Run in the build container only:

    python tests/golden/make_golden_cocoeval.py
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def synth_case(seed, n_img=30, n_cat=3, crowd_frac=0.05):
    rng = np.random.default_rng(seed)
    images, anns, dets = [], [], []
    aid = 1
    for i in range(n_img):
        iid = 100 + i
        images.append({"id": iid, "width": 640, "height": 512, "file_name": "%05d.jpg" % i})
        for _ in range(rng.poisson(5)):
            w, h = rng.uniform(6, 220, 2)
            x, y = rng.uniform(0, 640 - w), rng.uniform(0, 512 - h)
            cat = int(rng.integers(0, n_cat)) + 1
            crowd = int(rng.random() < crowd_frac)
            anns.append({"id": aid, "image_id": iid, "category_id": cat, "bbox": [float(x), float(y), float(w), float(h)],
                         "area": float(w * h), "iscrowd": crowd})
            aid += 1
            if rng.random() < 0.85:  # detected, with jitter; sometimes twice, sometimes with the wrong class
                for _ in range(1 + int(rng.random() < 0.15)):
                    j = rng.normal(0, 0.08, 4) * np.array([w, h, w, h])
                    dc = cat if rng.random() < 0.9 else int(rng.integers(0, n_cat)) + 1
                    dets.append({"image_id": iid, "category_id": dc,
                                 "bbox": [float(x + j[0]), float(y + j[1]), float(max(1.0, w + j[2])), float(max(1.0, h + j[3]))],
                                 "score": float(np.round(rng.uniform(0.3, 1.0), 3))})  # rounded -> score ties occur
        for _ in range(rng.poisson(2)):  # false positives
            w, h = rng.uniform(6, 150, 2)
            dets.append({"image_id": iid, "category_id": int(rng.integers(0, n_cat)) + 1,
                         "bbox": [float(rng.uniform(0, 640 - w)), float(rng.uniform(0, 512 - h)), float(w), float(h)],
                         "score": float(np.round(rng.uniform(0.05, 0.9), 3))})
    cats = [{"id": c + 1, "name": "c%d" % (c + 1)} for c in range(n_cat)]
    return {"images": images, "annotations": anns, "categories": cats}, dets


def reference_numbers(gt, dets):
    COCO, COCOeval = ref_loader.load_reference_cocoeval()
    with contextlib.redirect_stdout(io.StringIO()):
        api = COCO()
        api.dataset = json.loads(json.dumps(gt))
        api.createIndex()
        dt = api.loadRes(json.loads(json.dumps(dets)))
        ev = COCOeval(api, dt, "bbox")
        ev.evaluate()
        ev.accumulate()
        ev.summarize()
    stats = [float(s) for s in ev.stats]
    per_cat = []
    for k in range(ev.eval["precision"].shape[2]):
        p = ev.eval["precision"][:, :, k, 0, -1]
        p = p[p > -1]
        per_cat.append(float(np.mean(p) * 100) if p.size else float("nan"))
    return stats, per_cat


def main():
    cases = []
    for seed, kw in [(0, {}), (1, {"n_img": 12, "crowd_frac": 0.3}), (2, {"n_img": 50, "n_cat": 1, "crowd_frac": 0.0})]:
        gt, dets = synth_case(seed, **kw)
        stats, per_cat = reference_numbers(gt, dets)
        cases.append({"seed": seed, "gt": gt, "dets": dets, "stats": stats, "per_category_ap": per_cat})
        print("seed", seed, "AP %.4f AP50 %.4f" % (stats[0], stats[1]), "per-cat", per_cat)
    with open(os.path.join(HERE, "cocoeval_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden_cocoeval.py", "cases": cases}, f)


if __name__ == "__main__":
    main()
