"""COCO bbox matching on the GPU (pe_coco_match, csrc/coco_eval.cu) against the host evaluator (probenb200/evaluation.py, itself
pinned to the reference's vendored cocoeval.py to 1e-9): the per-(image, category, IoU threshold, area range) match / ignore
decisions are float64 comparisons in the same operation order, so precision, recall and every summary number must be EQUAL."""
import json
import os

import numpy as np
import pytest

from probenb200 import evaluation

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _random_set(seed, n_img=40, crowd=True):
    rng = np.random.default_rng(seed)
    anns, dets, aid = [], [], 1
    for i in range(n_img):
        for _ in range(int(rng.integers(0, 9))):
            x, y = rng.uniform(0, 560), rng.uniform(0, 440)
            w, h = rng.uniform(4, 160), rng.uniform(4, 130)
            c = int(rng.integers(0, 3))
            anns.append({"id": aid, "image_id": i, "category_id": c, "bbox": [x, y, w, h], "area": w * h * rng.uniform(0.5, 1.0),
                         "iscrowd": int(crowd and rng.random() < 0.1)})
            aid += 1
            for _ in range(int(rng.integers(0, 4))):  # detections around the box: duplicates, shifted, other class
                j = rng.normal(0, 6, 4)
                dets.append({"image_id": i, "category_id": c if rng.random() < 0.85 else int(rng.integers(0, 3)),
                             "bbox": [x + j[0], y + j[1], max(1.0, w + j[2]), max(1.0, h + j[3])], "score": float(np.round(rng.random(), 2))})
        for _ in range(int(rng.integers(0, 5))):      # false positives; rounded scores create ties (stable-sort order matters)
            dets.append({"image_id": i, "category_id": int(rng.integers(0, 3)),
                         "bbox": [rng.uniform(0, 560), rng.uniform(0, 440), rng.uniform(4, 200), rng.uniform(4, 200)],
                         "score": float(np.round(rng.random(), 2))})
    return anns, dets


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gpu_matching_equals_host_evaluator(seed):
    anns, dets = _random_set(seed)
    host = evaluation.COCOBBoxEval(anns, dets, image_ids=list(range(40)))
    want = host.evaluate()
    gpu = evaluation.COCOBBoxEval(anns, dets, image_ids=list(range(40)))
    got = gpu.evaluate(device="cuda")
    assert np.array_equal(host.precision, gpu.precision) and np.array_equal(host.recall, gpu.recall)
    assert want.keys() == got.keys()
    for k in want:
        assert want[k] == got[k] or (np.isnan(want[k]) and np.isnan(got[k])), k


def test_many_detections_per_image_cap_and_empty_groups():
    """More than maxDets = 100 detections in one (image, category) group, images without ground truth, categories without
    detections: the 100-cap, the -1 conventions and the empty slices must come out as on the host."""
    rng = np.random.default_rng(7)
    anns = [{"id": 1, "image_id": 0, "category_id": 0, "bbox": [50, 50, 100, 80], "area": 8000.0, "iscrowd": 0},
            {"id": 2, "image_id": 2, "category_id": 1, "bbox": [10, 10, 20, 20], "area": 400.0, "iscrowd": 0}]
    dets = [{"image_id": 0, "category_id": 0, "bbox": [50 + rng.normal(0, 20), 50 + rng.normal(0, 20), 100, 80], "score": float(rng.random())}
            for _ in range(150)]
    dets += [{"image_id": 1, "category_id": 0, "bbox": [0, 0, 30, 30], "score": 0.9}]
    a = evaluation.COCOBBoxEval(anns, dets, image_ids=[0, 1, 2])
    b = evaluation.COCOBBoxEval(anns, dets, image_ids=[0, 1, 2])
    want, got = a.evaluate(), b.evaluate(device="cuda")
    assert np.array_equal(a.precision, b.precision) and np.array_equal(a.recall, b.recall)
    assert all(want[k] == got[k] or (np.isnan(want[k]) and np.isnan(got[k])) for k in want)


def test_gpu_evaluator_reproduces_reference_goldens():
    """The reference's own COCOeval numbers (tests/golden/cocoeval_golden.json, generated from the vendored cocoeval.py)."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cocoeval_golden.json")))
    for case in gold["cases"]:
        gt = case["gt"]
        ev = evaluation.COCOBBoxEval(gt["annotations"], case["dets"], category_ids=[c["id"] for c in gt["categories"]],
                                     image_ids=[im["id"] for im in gt["images"]])
        res = ev.evaluate(device="cuda")
        for key, idx in (("AP", 0), ("AP50", 1), ("AP75", 2), ("APs", 3), ("APm", 4), ("APl", 5)):
            assert res[key] == pytest.approx(case["stats"][idx] * 100, abs=1e-9), key
