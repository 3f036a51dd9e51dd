"""CPU oracle for the ProbEn late-fusion stage.  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, the algorithm of the reference script
``demo/FLIR/demo_probEn.py`` (paths relative to /root/reference):

  * ``fusion``                      <- demo_probEn.py:189-196   (dispatch nms_1 / nms_bayesian)
  * ``cluster_and_fuse``            <- demo_probEn.py:92-187    (nms_bayesian)
  * ``probEn_multiclass``           <- demo_probEn.py:32-42     (bayesian_fusion_multiclass)
  * ``probEn_binary``               <- demo_probEn.py:24-30     (bayesian_fusion; KAIST K=1 form)
  * ``nms_max_argmax``              <- demo_probEn.py:44-71     (nms_1) + detectron2/layers/nms.py:9-26
                                       + torchvision 0.13 ``ops.boxes.batched_nms`` (coordinate trick) / ``nms``
  * ``late_fusion_dispatch``        <- demo_probEn.py:236-267   (0 / 1 / 2 / 3 non-empty models)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import it.  The product (``probenb200``) never does: it fails loudly without its CUDA library.

Parity status: PINNED.  ``tests/golden/make_golden_proben.py`` ran the unmodified reference functions
(loaded by path) in the build container and stored their outputs in
``tests/golden/proben_golden.npz``; ``tests/test_oracle_proben.py`` checks this restatement against those
vectors (and, when /root/reference is present, against the live reference) for every score x box mode.

Documented deviation: the reference orders detections with ``scores.argsort()[::-1]`` (numpy's default
unstable sort).  For exactly tied scores numpy's order is implementation defined for n > 16; the oracle
(and the CUDA kernel) define it as "descending score, ties -> higher concatenated index first", which is
what numpy produces for n <= 16 (insertion sort is stable, then reversed).
"""
import numpy as np

SCORE_MODES = ("probEn", "avg", "max")
BOX_MODES = ("v-avg", "s-avg", "avg", "argmax")


def probEn_multiclass(member_probs):
    """demo_probEn.py:32-42.  member_probs (m, K) float64 -> (score, class index in [0, K])."""
    m, k = member_probs.shape
    table = np.empty((m, k + 1), dtype=np.float64)
    table[:, :k] = member_probs
    table[:, k] = 1 - np.sum(member_probs, axis=1)
    with np.errstate(all="ignore"):
        joint = np.exp(np.sum(np.log(table), axis=0))
        post = joint / np.sum(joint)
    return np.max(post), int(np.argmax(post))


def probEn_binary(member_scores):
    """demo_probEn.py:24-30 (defined in the reference, never called; the K=1 special case)."""
    s = np.asarray(member_scores, dtype=np.float64)
    with np.errstate(all="ignore"):
        pos = np.exp(np.sum(np.log(s)))
        neg = np.exp(np.sum(np.log(1 - s)))
        return pos / (pos + neg)


def descending_order(scores):
    """demo_probEn.py:106 with the documented tie rule (see module docstring)."""
    return np.argsort(scores, kind="stable")[::-1]


def cluster_and_fuse(boxes, scores, classes, probs, variances, score_mode, box_mode,
                     iou_thr=0.5, img_w=640, img_h=512):
    """Greedy score-ordered clustering + per-cluster fusion (demo_probEn.py:92-187).

    All inputs float64 arrays of the CONCATENATED detections of the contributing models (model order
    preserved, demo_probEn.py:79-90).  Returns (head_indices, out_boxes (n,4) f64, out_scores f64 (n,),
    out_classes f64 (n,)); the reference then casts scores/classes to float32 tensors (:182-183).
    """
    boxes = np.asarray(boxes, np.float64).reshape(-1, 4)
    scores = np.asarray(scores, np.float64)
    classes = np.asarray(classes, np.float64)
    probs = np.asarray(probs, np.float64).reshape(len(scores), -1)
    variances = np.asarray(variances, np.float64).reshape(len(scores), -1)
    # class-offset trick, :100-103 ; legacy "+1" areas, :105
    ox1 = boxes[:, 0] + classes * img_w
    oy1 = boxes[:, 1] + classes * img_h
    ox2 = boxes[:, 2] + classes * img_w
    oy2 = boxes[:, 3] + classes * img_h
    area = (ox2 - ox1 + 1) * (oy2 - oy1 + 1)
    remaining = descending_order(scores)
    heads, fb, fs, fc = [], [], [], []
    while remaining.size:
        h, rest = remaining[0], remaining[1:]
        iw = np.maximum(0.0, np.minimum(ox2[h], ox2[rest]) - np.maximum(ox1[h], ox1[rest]) + 1)
        ih = np.maximum(0.0, np.minimum(oy2[h], oy2[rest]) - np.maximum(oy1[h], oy1[rest]) + 1)
        inter = iw * ih
        with np.errstate(all="ignore"):
            iou = inter / (area[h] + area[rest] - inter)
        hit = iou > iou_thr
        members = rest[hit]
        heads.append(int(h))
        if members.size:
            grp = np.concatenate([members, [h]])         # matches first, head LAST (:139-141)
            if score_mode == "probEn":
                s, c = probEn_multiclass(probs[grp])
            elif score_mode == "avg":
                s, c = np.mean(scores[grp]), classes[h]
            elif score_mode == "max":
                s, c = np.max(probs[grp]), classes[h]     # max over the whole m x K matrix (:151)
            else:
                raise ValueError(score_mode)
            if box_mode == "v-avg":
                w = 1.0 / variances[grp, 0]
                b = np.sum(boxes[grp] * (w / np.sum(w))[:, None], axis=0)
            elif box_mode == "s-avg":
                w = scores[grp]
                b = np.sum(boxes[grp] * (w / np.sum(w))[:, None], axis=0)
            elif box_mode == "avg":
                b = np.sum(boxes[grp], axis=0) / len(grp)
            elif box_mode == "argmax":
                b = boxes[grp[int(np.argmax(scores[grp]))]]
            else:
                raise ValueError(box_mode)
        else:
            s, c, b = scores[h], classes[h], boxes[h]
        fb.append(b)
        fs.append(s)
        fc.append(c)
        remaining = rest[~hit]                            # survivors: ovr <= thresh (:125,170)
    return (np.asarray(heads, np.int64), np.asarray(fb, np.float64).reshape(-1, 4),
            np.asarray(fs, np.float64), np.asarray(fc, np.float64))


def greedy_nms_f32(boxes, scores, iou_thr):
    """torchvision ``ops.nms`` CPU kernel semantics in float32: stable descending sort, area (x2-x1)(y2-y1),
    suppress j when inter / (area_i + area_j - inter) > thr.  Returns kept indices in score order."""
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    scores = np.asarray(scores, np.float32)
    n = len(scores)
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    area = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, bool)
    thr = np.float32(iou_thr)
    keep = []
    for a in range(n):
        i = order[a]
        if dead[i]:
            continue
        keep.append(i)
        js = order[a + 1:]
        w = np.maximum(np.float32(0), np.minimum(x2[i], x2[js]) - np.maximum(x1[i], x1[js]))
        h = np.maximum(np.float32(0), np.minimum(y2[i], y2[js]) - np.maximum(y1[i], y1[js]))
        inter = w * h
        with np.errstate(all="ignore"):
            ovr = inter / (area[i] + area[js] - inter)
        dead[js[ovr > thr]] = True
    return np.asarray(keep, np.int64)


def batched_nms_f32(boxes, scores, idxs, iou_thr):
    """detectron2/layers/nms.py:9-26 -> torchvision batched_nms, coordinate-offset form
    (``boxes + idx * (max_coordinate + 1)`` in float32, one nms call)."""
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    if boxes.shape[0] == 0:
        return np.zeros((0,), np.int64)
    idxs = np.asarray(idxs)
    off = idxs.astype(np.float32) * (boxes.max() + np.float32(1))
    return greedy_nms_f32(boxes + off[:, None], scores, iou_thr)


def nms_max_argmax(boxes, scores, classes, iou_thr=0.5):
    """demo_probEn.py:44-71: float32 tensors, per-class NMS, rows gathered in keep order."""
    b = np.asarray(boxes, np.float32).reshape(-1, 4)
    s = np.asarray(scores, np.float32)
    c = np.asarray(classes, np.float32)
    keep = batched_nms_f32(b, s, c, iou_thr)
    return keep, b[keep], s[keep], c[keep]


def _cat(infos, key, width=None):
    parts = [np.asarray(i[key], np.float64) for i in infos]
    if width is not None:
        parts = [p.reshape(-1, width) for p in parts]
    return np.concatenate(parts, axis=0)


def fusion(method, info_1, info_2, info_3="", iou_thr=0.5, img_w=640, img_h=512):
    """demo_probEn.py:189-196.  ``info_k`` dicts with keys bbox, score, class, prob, vars (python lists).
    Returns (boxes, scores float32, classes float32); boxes float64 (n,4) for the bayesian path,
    float32 for the ('max','argmax') NMS path - the dtypes the reference hands to ``Boxes``."""
    infos = [info_1, info_2] + ([info_3] if info_3 else [])
    if method[0] == "max" and method[1] == "argmax":
        _, b, s, c = nms_max_argmax(_cat(infos, "bbox", 4), _cat(infos, "score"), _cat(infos, "class"), iou_thr)
        return b, s, c
    k = np.asarray(infos[0]["prob"]).reshape(len(infos[0]["score"]), -1).shape[1]
    _, b, s, c = cluster_and_fuse(_cat(infos, "bbox", 4), _cat(infos, "score"), _cat(infos, "class"),
                                  _cat(infos, "prob", k), _cat(infos, "vars", 1),
                                  method[0], method[1], iou_thr, img_w, img_h)
    return b, s.astype(np.float32), c.astype(np.float32)


def late_fusion_dispatch(method, infos, iou_thr=0.5, img_w=640, img_h=512):
    """Per-image dispatch of demo_probEn.py:236-267 for M = len(infos) in {2, 3} models.
    Returns None when no model has detections (the reference ``continue``s, the image never reaches the
    evaluator), else (boxes, scores float32, classes float32)."""
    live = [i for i in infos if len(i["bbox"]) > 0]
    if not live:
        return None
    if len(live) == 1:
        i = live[0]
        return (np.asarray(i["bbox"], np.float64).reshape(-1, 4), np.asarray(i["score"], np.float32),
                np.asarray(i["class"], np.float32))
    return fusion(method, *live, iou_thr=iou_thr, img_w=img_w, img_h=img_h)
