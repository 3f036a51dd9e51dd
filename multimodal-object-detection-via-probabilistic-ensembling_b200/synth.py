"""Synthetic detection sets in the reference's saved-prediction schema.

The reference ships no detection JSONs, so parity and benchmarks run on generated ones: per image a
Poisson number of ground-truth boxes, each detected by every model with probability ``p_det`` with
Gaussian jitter, plus Poisson false positives; class logits -> float32 softmax -> first K columns;
variance = float32 exp(N(0, .5)); kept when the top foreground probability exceeds 0.5 (the detector's
SCORE_THRESH_TEST, demo_FLIR_save_predictions.py:51).  Field names follow the JSON written at
demo_FLIR_save_predictions.py:166-176.
"""
import numpy as np


def _softmax32(z):
    z = z.astype(np.float32)
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def synth_model_detections(num_images, num_models=3, seed=0, K=3, img_w=640, img_h=512, gt_mean=8.0,
                           p_det=0.8, fp_mean=2.0, force_count=None):
    """Returns ``dets[m]`` = dict(image, image_id, boxes, scores, classes, class_logits, probs, vars) of
    per-image python lists (the reference JSON layout).  ``force_count`` pins every model to exactly that
    many detections per image (stress sets)."""
    rng = np.random.default_rng(seed)
    dets = [dict(image=[], image_id=[], boxes=[], scores=[], classes=[], class_logits=[], probs=[], vars=[])
            for _ in range(num_models)]
    for i in range(num_images):
        g = int(rng.poisson(gt_mean)) if force_count is None else int(force_count)
        gxy = rng.uniform([0, 0], [img_w - 80, img_h - 72], size=(g, 2))
        gwh = rng.uniform(10, 120, size=(g, 2))
        gcls = rng.integers(0, K, size=g)
        for m in range(num_models):
            if force_count is None:
                hit = rng.random(g) < p_det
                nfp = int(rng.poisson(fp_mean))
            else:
                hit = np.ones(g, bool)
                nfp = 0
            xy = gxy[hit] + rng.normal(0, 2.0, size=(hit.sum(), 2))
            wh = np.maximum(gwh[hit] + rng.normal(0, 2.0, size=(hit.sum(), 2)), 2.0)
            cls_true = gcls[hit]
            fxy = rng.uniform([0, 0], [img_w - 80, img_h - 72], size=(nfp, 2))
            fwh = rng.uniform(10, 120, size=(nfp, 2))
            xy = np.concatenate([xy, fxy])
            wh = np.concatenate([wh, fwh])
            cls_true = np.concatenate([cls_true, rng.integers(0, K, size=nfp)])
            n = len(cls_true)
            logits = rng.normal(0, 1, size=(n, K + 1))
            logits[np.arange(n), cls_true] += rng.uniform(2, 6, size=n)
            logits = logits.astype(np.float32)
            probs = _softmax32(logits)[:, :K]
            boxes = np.concatenate([xy, xy + wh], axis=1)
            boxes[:, 0::2] = np.clip(boxes[:, 0::2], 0, img_w)
            boxes[:, 1::2] = np.clip(boxes[:, 1::2], 0, img_h)
            boxes = boxes.astype(np.float32)
            var = np.exp(rng.normal(0, 0.5, size=n)).astype(np.float32)
            keep = probs.max(axis=1) > 0.5 if force_count is None else np.ones(n, bool)
            d = dets[m]
            d["image"].append("FLIR_%05d.jpg" % i)
            d["image_id"].append(i)
            d["boxes"].append(boxes[keep].astype(np.float64).tolist())
            d["scores"].append(probs[keep].max(axis=1).astype(np.float64).tolist())
            d["classes"].append(probs[keep].argmax(axis=1).tolist())
            d["class_logits"].append(logits[keep].astype(np.float64).tolist())
            d["probs"].append(probs[keep].astype(np.float64).tolist())
            d["vars"].append(var[keep].astype(np.float64)[:, None].tolist())
    return dets


def image_info(det, i):
    """The ``info_k`` dict demo_probEn.py:205-234 builds for image i of one model's predictions."""
    return {"img_name": det["image"][i], "bbox": det["boxes"][i], "score": det["scores"][i],
            "class": det["classes"][i], "class_logits": det["class_logits"][i], "prob": det["probs"][i],
            "vars": det["vars"][i]}


def synth_packed(num_images, num_models=2, mean_dets=7.5, seed=0, K=3, img_w=640, img_h=512, force_count=None):
    """Fast vectorised generator for benchmark-scale batches (millions of images): returns the packed SoA
    numpy arrays of ``fusion.pack_detections`` directly.  Per (image, model) Poisson(mean_dets) detections;
    consecutive models re-detect the same objects with jitter so that cross-model clusters form.
    ``force_count``: every model detects exactly that many objects per image (the detector's 100-per-image regime)."""
    rng = np.random.default_rng(seed)
    B, M = num_images, num_models
    if force_count is None:
        cnt_obj = rng.poisson(mean_dets / 0.8, size=B)
        hit = [rng.random(int(cnt_obj.sum())) < 0.8 for _ in range(M)]
    else:
        cnt_obj = np.full(B, int(force_count))
        hit = [np.ones(int(cnt_obj.sum()), bool) for _ in range(M)]
    obj_img = np.repeat(np.arange(B), cnt_obj)
    oxy = rng.uniform([0, 0], [img_w - 130, img_h - 130], size=(len(obj_img), 2))
    owh = rng.uniform(10, 120, size=(len(obj_img), 2))
    ocls = rng.integers(0, K, size=len(obj_img))
    rows = []
    for m in range(M):
        sel = np.nonzero(hit[m])[0]
        n = len(sel)
        xy = oxy[sel] + rng.normal(0, 2.0, size=(n, 2))
        wh = np.maximum(owh[sel] + rng.normal(0, 2.0, size=(n, 2)), 2.0)
        logits = rng.normal(0, 1, size=(n, K + 1)).astype(np.float32)
        logits[np.arange(n), ocls[sel]] += rng.uniform(2.5, 6, size=n).astype(np.float32)
        probs = _softmax32(logits)[:, :K]
        boxes = np.clip(np.concatenate([xy, xy + wh], axis=1), 0, [img_w, img_h, img_w, img_h]).astype(np.float32)
        var = np.exp(rng.normal(0, 0.5, size=n)).astype(np.float32)
        rows.append((obj_img[sel], np.full(n, m), boxes, probs, var))
    img = np.concatenate([r[0] for r in rows])
    mod = np.concatenate([r[1] for r in rows])
    order = np.lexsort((mod, img))
    boxes = np.concatenate([r[2] for r in rows])[order]
    probs = np.concatenate([r[3] for r in rows])[order]
    var = np.concatenate([r[4] for r in rows])[order]
    counts = np.bincount(img * M + mod, minlength=B * M)
    offsets = np.zeros(B * M + 1, np.int32)
    np.cumsum(counts, out=offsets[1:])
    return {"boxes": np.ascontiguousarray(boxes), "scores": probs.max(axis=1).astype(np.float32),
            "classes": probs.argmax(axis=1).astype(np.int32), "probs": np.ascontiguousarray(probs),
            "vars": var, "offsets": offsets, "B": B, "M": M, "K": K}
