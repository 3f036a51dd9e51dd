"""The two drop-in CLIs end to end on a tiny synthetic FLIR-layout dataset (JPEG pairs + COCO annotations + random
checkpoints): demo_FLIR_save_predictions.py (GPU decode and --cpu_decode) -> prediction JSONs in the reference's
schema (demo_FLIR_save_predictions.py:166-176) -> demo_probEn.py (late fusion + COCO bbox evaluation)."""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest
import torch

from probenb200 import weights
from probenb200.opt import config_parser

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_cli(name):
    spec = importlib.util.spec_from_file_location("cli_" + name, os.path.join(ROOT, "demo", "FLIR", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _make_dataset(root, n=5):
    rng = np.random.default_rng(0)
    os.makedirs(os.path.join(root, "RGB"))
    os.makedirs(os.path.join(root, "thermal_8_bit"))
    images, anns = [], []
    for i in range(n):
        stem = "FLIR_%05d" % i
        yy, xx = np.mgrid[0:128, 0:160].astype(np.float32)
        th = 120 + 60 * np.sin(0.05 * xx + i) * np.cos(0.07 * yy) + rng.normal(0, 4, (128, 160))
        th = np.clip(th, 0, 255).astype(np.uint8)
        rgb = np.clip(cv2.resize(np.dstack([th, th // 2 + 40, 255 - th]), (260, 200)).astype(np.float32) + rng.normal(0, 3, (200, 260, 3)), 0, 255)
        cv2.imwrite(os.path.join(root, "thermal_8_bit", stem + ".jpeg"), th)
        cv2.imwrite(os.path.join(root, "RGB", stem + ".jpg"), rgb.astype(np.uint8))
        images.append({"id": 1000 + i, "file_name": "thermal_8_bit/%s.jpeg" % stem, "height": 128, "width": 160})
        for k in range(3):
            x, y = rng.uniform(0, 100), rng.uniform(0, 80)
            anns.append({"id": len(anns) + 1, "image_id": 1000 + i, "category_id": int(rng.integers(1, 4)),
                         "bbox": [float(x), float(y), 40.0, 30.0], "area": 1200.0, "iscrowd": 0})
    # FLIR-style 1-based dataset category ids: the evaluator must un-map contiguous class indices (FLIR_evaluation.py:163-175)
    cats = [{"id": c + 1, "name": n_} for c, n_ in enumerate(["person", "bicycle", "car"])]
    json.dump({"images": images, "annotations": anns, "categories": cats}, open(os.path.join(root, "FLIR_thermal_RGBT_pairs_val.json"), "w"))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("flir"))
    _make_dataset(os.path.join(d, "val"))
    return d


SCHEMA = ("image", "boxes", "scores", "classes", "image_id", "class_logits", "probs", "vars")


def test_save_predictions_then_proben(workdir):
    save = _load_cli("demo_FLIR_save_predictions")
    out = os.path.join(workdir, "out") + "/"
    counts = {}
    for m, (method, cin, mid) in enumerate([("thermal_only", 3, False), ("early_fusion", 4, False), ("middle_fusion", 3, True)]):
        ck = os.path.join(workdir, method + ".pth")
        torch.save({"model": weights.random_state_dict(50, cin, 3, seed=20 + m, middle_fusion=mid)}, ck)
        args = config_parser(["--dataset_path", os.path.join(workdir, "val"), "--fusion_method", method, "--model_path", ck,
                              "--outfolder", out])
        path = save.save_predictions(args, batch=2, depth=50)
        pred = json.load(open(path))
        assert tuple(pred.keys()) == SCHEMA
        assert pred["image_id"] == [1000 + i for i in range(5)] and len(pred["boxes"]) == 5
        for i in range(5):
            n = len(pred["boxes"][i])
            assert len(pred["scores"][i]) == n and len(pred["probs"][i]) == n and len(pred["vars"][i]) == n
            assert all(c <= 2 for c in pred["classes"][i])
            assert all(len(p) == 3 for p in pred["probs"][i]) and all(len(l) == 4 for l in pred["class_logits"][i])
        counts[method] = sum(len(b) for b in pred["boxes"])
    assert sum(counts.values()) > 0, counts
    proben = _load_cli("demo_probEn")
    res = proben.main(["--dataset_path", os.path.join(workdir, "val"), "--prediction_path", out, "--score_fusion", "probEn",
                       "--box_fusion", "v-avg", "--outfolder", out])
    assert res is not None and set(("AP", "AP50", "AP75", "APs", "APm", "APl")) <= set(res)
    fused = json.load(open(os.path.join(out, "probEn_probEn_v-avg_fused.json")))
    assert all(set(d) == {"image_id", "category_id", "bbox", "score"} for d in fused)
    assert all(d["category_id"] in (1, 2, 3) for d in fused)  # dataset ids, not contiguous indices
    # that run read the binary columnar files written next to the JSONs; the JSON path must give the same result
    for f in os.listdir(out):
        if f.endswith(".pedet"):
            os.remove(os.path.join(out, f))
    res_json = proben.main(["--dataset_path", os.path.join(workdir, "val"), "--prediction_path", out, "--score_fusion", "probEn",
                            "--box_fusion", "v-avg", "--outfolder", out])
    assert res_json == pytest.approx(res, nan_ok=True)
    assert json.load(open(os.path.join(out, "probEn_probEn_v-avg_fused.json"))) == fused


def test_rgb_only_with_model_zoo_pkl(workdir):
    """The reference CLI's first branch (demo_FLIR_save_predictions.py:58-60): rgb_only = the 80-class COCO zoo model,
    read from a Detectron2 model-zoo ``.pkl``; only classes <= 2 reach the JSON (:148-164), probs rows have 80 entries."""
    import pickle
    save = _load_cli("demo_FLIR_save_predictions")
    sd = weights.random_state_dict(50, 3, 80, seed=27, head_gain=0.2)
    # make the three FLIR-relevant COCO classes win so that rows survive the class <= 2 filter
    sd["roi_heads.box_predictor.cls_score.bias"][:3] += torch.tensor([5.0, 7.0, 6.0])
    ck = os.path.join(workdir, "model_final_zoo.pkl")
    pickle.dump({"model": {k: v.numpy() for k, v in sd.items() if "var_pred" not in k}, "__author__": "Detectron2 Model Zoo"}, open(ck, "wb"))
    out = os.path.join(workdir, "out_rgb") + "/"
    args = config_parser(["--dataset_path", os.path.join(workdir, "val"), "--fusion_method", "rgb_only", "--model_path", ck, "--outfolder", out])
    pred = json.load(open(save.save_predictions(args, batch=2, depth=50)))
    assert tuple(pred.keys()) == SCHEMA and len(pred["boxes"]) == 5
    n = 0
    for i in range(5):
        assert all(c <= 2 for c in pred["classes"][i])
        assert all(len(p) == 80 for p in pred["probs"][i]) and all(len(l) == 81 for l in pred["class_logits"][i])
        assert all(v == [1.0] for v in pred["vars"][i])  # zoo models carry no variance head: zero-filled -> exp(0)
        n += len(pred["boxes"][i])
    assert n > 0


def test_single_model_map_cli(workdir):
    """demo_mAP_FLIR.py: one detector over the validation pairs -> COCO bbox numbers."""
    cli = _load_cli("demo_mAP_FLIR")
    ck = os.path.join(workdir, "thermal_map.pth")
    torch.save({"model": weights.random_state_dict(50, 3, 3, seed=20)}, ck)
    out = os.path.join(workdir, "out_map") + "/"
    res = cli.main(["--dataset_path", os.path.join(workdir, "val"), "--fusion_method", "thermal_only", "--model_path", ck,
                    "--outfolder", out, "--batch", "2", "--depth", "50"])
    assert set(("AP", "AP50", "AP75", "APs", "APm", "APl")) <= set(res)
    assert json.load(open(os.path.join(out, "FLIR_thermal_only_mAP.json"))).keys() == res.keys()


def test_gpu_decode_agrees_with_cpu_decode(workdir):
    """Same checkpoint, same pairs: frames decoded by nvJPEG + device assembly vs cv2 (the reference's input code).
    Grey-scale thermal JPEGs decode within +-2 grey levels, so the detections must be the same objects."""
    save = _load_cli("demo_FLIR_save_predictions")
    ck = os.path.join(workdir, "thermal_cmp.pth")
    torch.save({"model": weights.random_state_dict(50, 3, 3, seed=20)}, ck)
    preds = []
    for cpu in (False, True):
        out = os.path.join(workdir, "out_cpu" if cpu else "out_gpu") + "/"
        args = config_parser(["--dataset_path", os.path.join(workdir, "val"), "--fusion_method", "thermal_only", "--model_path", ck,
                              "--outfolder", out])
        preds.append(json.load(open(save.save_predictions(args, batch=3, depth=50, cpu_decode=cpu))))
    a, b = preds
    assert a["image_id"] == b["image_id"]
    na, nb = sum(len(x) for x in a["boxes"]), sum(len(x) for x in b["boxes"])
    assert na > 0 and abs(na - nb) <= max(3, 0.25 * na), (na, nb)  # +-2 grey levels can flip detections near the 0.5 score threshold


def test_kaist_lamr_then_binary_proben(tmp_path):
    """demo/KAIST: per-modality LAMR txt + variance npz (reference format, demo_LAMR_KAIST.py:127-145), then K = 1
    ProbEn of two modalities; the fused txt must equal the oracle's fusion of the same files."""
    from oracle import proben_oracle as O
    root = str(tmp_path / "kaist")
    rng = np.random.default_rng(3)
    entries = []
    for i in range(4):
        e = "set06/V000/I%05d" % i
        entries.append(e)
        for sub in ("visible", "lwir"):
            os.makedirs(os.path.join(root, "set06", "V000", sub), exist_ok=True)
        yy, xx = np.mgrid[0:128, 0:160].astype(np.float32)
        th = np.clip(110 + 70 * np.sin(0.06 * xx + i) * np.cos(0.05 * yy) + rng.normal(0, 4, (128, 160)), 0, 255).astype(np.uint8)
        cv2.imwrite(os.path.join(root, "set06", "V000", "lwir", "I%05d.jpg" % i), th)
        cv2.imwrite(os.path.join(root, "set06", "V000", "visible", "I%05d.jpg" % i), np.dstack([th, 255 - th, th // 2 + 30]))
    split = os.path.join(root, "test-all-20.txt")
    open(split, "w").write("\n".join(entries) + "\n")
    spec = importlib.util.spec_from_file_location("cli_kaist", os.path.join(ROOT, "demo", "KAIST", "demo_LAMR_KAIST.py"))
    kaist = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kaist)
    out = str(tmp_path / "pred") + "/"
    for m, method in enumerate(["thermal_only", "rgb_only"]):
        ck = str(tmp_path / (method + ".pth"))
        torch.save({"model": weights.random_state_dict(50, 3, 1, seed=40 + m, head_gain=3.0)}, ck)
        txt, npz = kaist.main(["--dataset_path", root, "--split_file", split, "--fusion_method", method, "--model_path", ck,
                               "--outfolder", out, "--batch", "2"])
        rows = [l.strip().split(",") for l in open(txt)]
        assert all(len(r) == 6 and 1 <= int(r[0]) <= 4 for r in rows)
        v = np.load(npz, allow_pickle=True)["vars"].item()
        assert sorted(v.keys()) == [1, 2, 3, 4] and sum(len(x) for x in v.values()) == len(rows)
    sys.path.insert(0, os.path.join(ROOT, "demo", "KAIST"))
    spec = importlib.util.spec_from_file_location("cli_kaist_pe", os.path.join(ROOT, "demo", "KAIST", "demo_probEn_KAIST.py"))
    pe = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pe)
    fused = pe.main(["--prediction_path", out, "--methods", "thermal_only", "rgb_only", "--img_w", "160", "--img_h", "128"])
    got = np.loadtxt(fused, delimiter=",", ndmin=2) if os.path.getsize(fused) else np.zeros((0, 6))
    files = [pe.read_modality(out, m, 4) for m in ("thermal_only", "rgb_only")]
    want = []
    for i in range(4):
        infos = []
        for f in files:
            lo, hi = int(f.offsets[i]), int(f.offsets[i + 1])
            infos.append({"bbox": f.boxes[lo:hi].astype(np.float64).tolist(), "score": f.scores[lo:hi].astype(np.float64).tolist(),
                          "class": f.classes[lo:hi].tolist(), "prob": f.probs[lo:hi].astype(np.float64).tolist(),
                          "vars": f.vars[lo:hi].astype(np.float64).tolist()})
        r = O.late_fusion_dispatch(("probEn", "v-avg"), infos, img_w=160, img_h=128)
        if r is None:
            continue
        b, s, c = (np.asarray(t, np.float64) for t in r)
        for k in range(len(s)):
            if int(c[k]) == 0:
                want.append([i + 1, b[k][0], b[k][1], b[k][2] - b[k][0], b[k][3] - b[k][1], s[k]])
    want = np.asarray(want, np.float64).reshape(-1, 6)
    assert got.shape == want.shape and got.shape[0] > 0
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-3)
