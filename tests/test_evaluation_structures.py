"""Host-side pieces: COCO bbox evaluation restatement, Boxes/Instances schema, CLI flag parser, pipeline helpers."""
import numpy as np
import pytest
import torch

from probenb200 import evaluation
from probenb200.opt import config_parser
from probenb200.structures import Boxes, Instances


def _gt(n_img=6, seed=0):
    rng = np.random.default_rng(seed)
    anns, aid = [], 1
    for i in range(n_img):
        for _ in range(rng.integers(1, 5)):
            x, y = rng.uniform(0, 500), rng.uniform(0, 400)
            w, h = rng.uniform(20, 120), rng.uniform(20, 100)
            anns.append({"id": aid, "image_id": i, "category_id": int(rng.integers(0, 3)), "bbox": [x, y, w, h], "area": w * h, "iscrowd": 0})
            aid += 1
    return anns


def test_perfect_detections_give_ap_100():
    anns = _gt()
    dets = [{"image_id": a["image_id"], "category_id": a["category_id"], "bbox": a["bbox"], "score": 0.9} for a in anns]
    res = evaluation.COCOBBoxEval(anns, dets).evaluate()
    assert abs(res["AP"] - 100) < 1e-9 and abs(res["AP50"] - 100) < 1e-9


def test_half_shifted_detections_and_false_positives():
    anns = [{"id": 1, "image_id": 0, "category_id": 0, "bbox": [0, 0, 100, 100], "area": 1e4, "iscrowd": 0},
            {"id": 2, "image_id": 0, "category_id": 0, "bbox": [200, 200, 100, 100], "area": 1e4, "iscrowd": 0}]
    dets = [{"image_id": 0, "category_id": 0, "bbox": [0, 0, 100, 100], "score": 0.9},       # IoU 1
            {"image_id": 0, "category_id": 0, "bbox": [400, 0, 50, 50], "score": 0.8},       # false positive
            {"image_id": 0, "category_id": 0, "bbox": [200, 200, 100, 80], "score": 0.7}]    # IoU 0.8
    res = evaluation.COCOBBoxEval(anns, dets).evaluate()
    # AP50: precision envelope 1.0 up to recall .5, then 2/3 up to recall 1 -> mean over 101 points
    want50 = (51 * 1.0 + 50 * (2 / 3)) / 101 * 100
    assert abs(res["AP50"] - want50) < 1e-6
    # second match counts for thresholds <= 0.8 (7 of 10); others keep only the first
    want = (7 * want50 + 3 * (51 * 1.0) / 101 * 100) / 10
    assert abs(res["AP"] - want) < 1e-6


def test_bbox_iou_and_crowd():
    iou = evaluation.bbox_iou([[0, 0, 10, 10]], [[5, 0, 10, 10], [0, 0, 20, 20]], [0, 1])
    assert abs(iou[0, 0] - 50 / 150) < 1e-12 and abs(iou[0, 1] - 1.0) < 1e-12


def test_instances_to_coco_json_drops_background_and_remaps():
    out = evaluation.instances_to_coco_json([[1, 2, 11, 22]] * 4, [.9, .8, .7, .6], [0, 3, 5, 7], 42)
    assert [d["category_id"] for d in out] == [0, 2, 2] and out[0]["bbox"] == [1, 2, 10, 20] and out[0]["image_id"] == 42


def test_boxes_and_instances_schema():
    b = Boxes([np.array([1.0, 2.0, 3.0, 4.0]), np.array([0.0, 0.0, 700.0, 600.0])])
    assert b.tensor.dtype == torch.float32 and len(b) == 2
    b.clip((512, 640))
    assert b.tensor[1].tolist() == [0, 0, 640, 512]
    assert Boxes([]).tensor.shape == (0, 4)
    assert b.nonempty().tolist() == [True, True] and abs(float(b.area()[0]) - 4.0) < 1e-6
    inst = Instances((512, 640))
    inst.pred_boxes = b
    inst.scores = torch.tensor([.9, .8])
    inst.pred_classes = torch.tensor([0, 1])
    assert len(inst) == 2 and inst.image_size == (512, 640) and inst.has("scores")
    sub = inst[inst.scores > .85]
    assert len(sub) == 1 and sub.pred_boxes.tensor.shape == (1, 4)
    with pytest.raises(AssertionError):
        inst.bad = torch.zeros(3)
    cat = Instances.cat([inst, inst])
    assert len(cat) == 4


def test_cli_flags_match_reference():
    a = config_parser(["--dataset_path", "D", "--prediction_path", "P/"])
    assert (a.outfolder, a.dataset_name, a.fusion_method, a.score_fusion, a.box_fusion) == ("out", "FLIR", "middle_fusion", "probEn", "v-avg")
    with pytest.raises(SystemExit):
        config_parser(["--score_fusion", "median"])


def test_shard_range_matches_inference_sampler_rule():
    from probenb200.pipeline import shard_range
    got = [shard_range(10, r, 4) for r in range(4)]
    assert got == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [shard_range(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]


def test_resize_shape_rule():
    from probenb200.detector import resize_shortest_edge_shape
    assert resize_shortest_edge_shape(512, 640) == (800, 1000)
    assert resize_shortest_edge_shape(480, 1920) == (333, 1333)


def test_category_ids_are_unmapped_like_flirevaluator():
    """FLIR_evaluation.py:163-175 / coco.py:87-88: contiguous class index -> sorted dataset category ids (1-based in
    FLIR-style files); without the un-mapping every class is off by one and AP collapses."""
    anns = [dict(a, category_id=a["category_id"] + 1) for a in _gt()]
    cats = [{"id": 3, "name": "car"}, {"id": 1, "name": "person"}, {"id": 2, "name": "bicycle"}]  # unsorted on purpose
    assert evaluation.contiguous_to_dataset_ids(cats) == [1, 2, 3]
    dets = []
    for a in anns:
        b = a["bbox"]
        dets += evaluation.instances_to_coco_json([[b[0], b[1], b[0] + b[2], b[1] + b[3]]], [0.9], [a["category_id"] - 1], a["image_id"])
    wrong = evaluation.COCOBBoxEval(anns, [dict(d) for d in dets]).evaluate()
    assert wrong["AP"] < 50
    evaluation.unmap_category_ids(dets, cats)
    assert abs(evaluation.COCOBBoxEval(anns, dets).evaluate()["AP"] - 100) < 1e-9
    with pytest.raises(AssertionError):
        evaluation.unmap_category_ids([{"category_id": 3}], cats)
    assert evaluation.unmap_category_ids([{"category_id": 2}], None) == [{"category_id": 2}]


def test_stale_pedet_sidecar_is_not_used(tmp_path):
    import importlib.util
    import os
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("cli_demo_probEn_host", os.path.join(root, "demo", "FLIR", "demo_probEn.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    j, b = str(tmp_path / "a.json"), str(tmp_path / "a.pedet")
    open(b, "w").write("x")
    assert cli._fresh_sidecars([j], [b], False)            # no JSON beside it: the sidecar is the only source
    open(j, "w").write("{}")
    os.utime(b, (time.time() - 100, time.time() - 100))
    assert not cli._fresh_sidecars([j], [b], False)        # JSON regenerated after the sidecar
    assert cli._fresh_sidecars([j], [b], True)             # --binary forces it
    os.utime(b, None)
    assert cli._fresh_sidecars([j], [b], False)
    assert not cli._fresh_sidecars([j], [str(tmp_path / "missing.pedet")], True)
