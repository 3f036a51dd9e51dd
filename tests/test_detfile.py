"""Binary columnar detections file: lossless round trip with the reference's JSON schema and identical packing to
the JSON path (fusion.pack_detections)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probenb200 import detfile, fusion, synth  # noqa: E402


def _json_dict(seed, n_img=12):
    d = synth.synth_model_detections(n_img, 3, seed=seed)[seed % 3]
    return d


def test_round_trip_json_binary_json(tmp_path):
    d = _json_dict(1)
    f = detfile.DetFile.from_json_dict(d)
    p = f.save(str(tmp_path / "val_thermal_only_predictions.pedet"))
    g = detfile.DetFile.load(p)
    assert g.n_images == len(d["image_id"]) and g.K == 3
    back = g.to_json_dict()
    for key in ("image_id", "classes"):
        assert back[key] == [list(map(int, x)) if isinstance(x, list) else int(x) for x in d[key]]
    for key in ("boxes", "scores", "probs", "class_logits"):
        for a, b in zip(back[key], d[key]):
            assert np.array_equal(np.asarray(a, np.float32), np.asarray(b, np.float32))
    # the JSON text itself survives: float32 values print/parse exactly
    again = detfile.DetFile.from_json_dict(json.loads(json.dumps(back)))
    assert np.array_equal(again.boxes, g.boxes) and np.array_equal(again.vars, g.vars) and np.array_equal(again.probs, g.probs)


def test_pack_models_equals_json_packing(tmp_path):
    dets = synth.synth_model_detections(20, 3, seed=4)
    files = [detfile.DetFile.load(detfile.DetFile.from_json_dict(d).save(str(tmp_path / ("m%d.pedet" % i)))) for i, d in enumerate(dets)]
    a = detfile.pack_models(files)
    images = [[synth.image_info(d, i) for d in dets] for i in range(20)]
    b = fusion.pack_detections(images)
    assert a["B"] == b["B"] and a["M"] == b["M"] and a["K"] == b["K"]
    for k in ("offsets", "boxes", "scores", "classes", "probs", "vars"):
        assert np.array_equal(a[k], b[k]), k


def test_empty_images_and_empty_file(tmp_path):
    d = {"image": ["a", "b"], "boxes": [[], []], "scores": [[], []], "classes": [[], []], "image_id": [5, 6],
         "class_logits": [[], []], "probs": [[], []], "vars": [[], []]}
    f = detfile.DetFile.load(detfile.DetFile.from_json_dict(d).save(str(tmp_path / "e.pedet")))
    assert f.n_images == 2 and f.boxes.shape == (0, 4) and f.to_json_dict()["boxes"] == [[], []]
    p = detfile.pack_models([f, f])
    assert p["offsets"].tolist() == [0, 0, 0, 0, 0]
