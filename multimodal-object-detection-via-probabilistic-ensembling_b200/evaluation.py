"""COCO-style bbox evaluation used by the ProbEn demo (parity harness).

Host-side numpy restatement of what ``FLIREvaluator`` does with the fused detections
(detectron2/evaluation/FLIR_evaluation.py:313-382 ``instances_to_coco_json``: XYXY->XYWH, keep classes
{0,1,2,5,7,16}, remap 5/7 -> 2, so a fused background class 3 is dropped) followed by the stock COCOeval bbox
protocol (detectron2/pycocotools/cocoeval.py:124-420: IoU .50:.05:.95, 101 recall points, areas all/small/
medium/large, maxDets 1/10/100, crowd handling) and the summary of ``_derive_coco_results``
(FLIR_evaluation.py:249-310).  pycocotools' C ``bbIou`` (not installed here) is restated as plain xywh IoU with
the crowd rule inter / area_dt.
"""
import numpy as np

KEEP_CLASSES = (0, 1, 2, 5, 7, 16)
REMAP = {5: 2, 7: 2}


def instances_to_coco_json(boxes, scores, classes, image_id):
    """FLIR_evaluation.py:313-382."""
    out = []
    boxes = np.asarray(boxes, np.float64).reshape(-1, 4)
    for b, s, c in zip(boxes, np.asarray(scores, np.float64), np.asarray(classes)):
        c = int(c)
        if c not in KEEP_CLASSES:
            continue
        c = REMAP.get(c, c)
        out.append({"image_id": image_id, "category_id": c, "bbox": [b[0], b[1], b[2] - b[0], b[3] - b[1]], "score": float(s)})
    return out


def contiguous_to_dataset_ids(categories):
    """Reverse of ``thing_dataset_id_to_contiguous_id``: ``load_coco_json`` maps the sorted dataset category ids to
    0..C-1 (data/datasets/coco.py:87-88) and the evaluator un-maps every prediction before COCOeval
    (FLIR_evaluation.py:163-175).  Returns the list ``dataset_id[contiguous]``; ``None`` when the annotation file
    carries no category table (ids are then taken as they are)."""
    if not categories:
        return None
    return sorted(int(c["id"]) for c in categories)


def unmap_category_ids(detections, categories):
    """FLIR_evaluation.py:163-175: contiguous class index -> dataset category id, in place; an index without a
    category is an error exactly as in the reference (``assert category_id in reverse_id_mapping``)."""
    ids = contiguous_to_dataset_ids(categories)
    if ids is None:
        return detections
    for d in detections:
        c = d["category_id"]
        assert 0 <= c < len(ids), "A prediction has category_id={}, which is not available in the dataset.".format(c)
        d["category_id"] = ids[c]
    return detections


def bbox_iou(dt, gt, iscrowd):
    """maskApi bbIou: dt (D,4) xywh, gt (G,4) xywh -> (D,G); crowd gt uses the detection area as union."""
    dt = np.asarray(dt, np.float64).reshape(-1, 4)
    gt = np.asarray(gt, np.float64).reshape(-1, 4)
    if len(dt) == 0 or len(gt) == 0:
        return np.zeros((len(dt), len(gt)))
    da, ga = dt[:, 2] * dt[:, 3], gt[:, 2] * gt[:, 3]
    w = np.minimum(dt[:, None, 0] + dt[:, None, 2], gt[None, :, 0] + gt[None, :, 2]) - np.maximum(dt[:, None, 0], gt[None, :, 0])
    h = np.minimum(dt[:, None, 1] + dt[:, None, 3], gt[None, :, 1] + gt[None, :, 3]) - np.maximum(dt[:, None, 1], gt[None, :, 1])
    inter = np.clip(w, 0, None) * np.clip(h, 0, None)
    union = np.where(np.asarray(iscrowd, bool)[None, :], da[:, None], da[:, None] + ga[None, :] - inter)
    return inter / union


class COCOBBoxEval:
    IOU_THRS = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)
    REC_THRS = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
    AREA_RNG = [[0 ** 2, 1e5 ** 2], [0 ** 2, 32 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e5 ** 2]]
    MAX_DETS = [1, 10, 100]

    def __init__(self, annotations, detections, category_ids=None, image_ids=None):
        """annotations: COCO 'annotations' dicts (image_id, category_id, bbox xywh, area, iscrowd);
        detections: dicts from instances_to_coco_json."""
        self.gts, self.dts = {}, {}
        for a in annotations:
            a = dict(a)
            a["ignore"] = int(a.get("iscrowd", 0))
            a.setdefault("area", a["bbox"][2] * a["bbox"][3])
            self.gts.setdefault((a["image_id"], a["category_id"]), []).append(a)
        for d in detections:
            d = dict(d)
            d["area"] = d["bbox"][2] * d["bbox"][3]
            self.dts.setdefault((d["image_id"], d["category_id"]), []).append(d)
        cats = set(k[1] for k in self.gts) | set(k[1] for k in self.dts)
        self.cat_ids = sorted(category_ids if category_ids is not None else cats)
        imgs = set(k[0] for k in self.gts) | set(k[0] for k in self.dts)
        self.img_ids = sorted(image_ids if image_ids is not None else imgs)

    def _evaluate_img(self, img, cat, rng, max_det):
        gt = self.gts.get((img, cat), [])
        dt = self.dts.get((img, cat), [])
        if not gt and not dt:
            return None
        g_ign = np.array([g["ignore"] or g["area"] < rng[0] or g["area"] > rng[1] for g in gt], bool)
        gind = np.argsort(g_ign, kind="mergesort")
        gt = [gt[i] for i in gind]
        g_ign = g_ign[gind]
        dind = np.argsort([-d["score"] for d in dt], kind="mergesort")[:max_det]
        dt = [dt[i] for i in dind]
        iscrowd = [int(g.get("iscrowd", 0)) for g in gt]
        ious = bbox_iou([d["bbox"] for d in dt], [g["bbox"] for g in gt], iscrowd)
        T, G, Dn = len(self.IOU_THRS), len(gt), len(dt)
        gtm = -np.ones((T, G), int)
        dtm = -np.ones((T, Dn), int)
        dt_ig = np.zeros((T, Dn), bool)
        for ti, t in enumerate(self.IOU_THRS):
            for di in range(Dn):
                iou = min(t, 1 - 1e-10)
                m = -1
                for gi in range(G):
                    if gtm[ti, gi] >= 0 and not iscrowd[gi]:
                        continue
                    if m > -1 and not g_ign[m] and g_ign[gi]:
                        break
                    if ious[di, gi] < iou:
                        continue
                    iou = ious[di, gi]
                    m = gi
                if m == -1:
                    continue
                dt_ig[ti, di] = g_ign[m]
                dtm[ti, di] = m
                gtm[ti, m] = di
        d_area = np.array([d["area"] < rng[0] or d["area"] > rng[1] for d in dt], bool).reshape(1, Dn)
        dt_ig = dt_ig | ((dtm < 0) & np.repeat(d_area, T, 0))
        return {"dtm": dtm, "dt_scores": np.array([d["score"] for d in dt]), "g_ign": g_ign, "dt_ig": dt_ig}

    def _match_host(self):
        """evaluateImg for every (category, area range, image) on the host (numpy): {(ki, ai): [per-image results]}."""
        out = {}
        md = self.MAX_DETS[-1]
        for ki, cat in enumerate(self.cat_ids):
            for ai, rng in enumerate(self.AREA_RNG):
                out[(ki, ai)] = [e for e in (self._evaluate_img(i, cat, rng, md) for i in self.img_ids) if e is not None]
        return out

    def _match_gpu(self, device):
        """The same matching on the GPU (``pe_coco_match``, csrc/coco_eval.cu): one block per (image, category) group, one thread
        per (IoU threshold, area range); float64 IoUs in the host evaluator's operation order, so every decision is identical."""
        import ctypes

        import torch

        from . import _lib
        lib = _lib.load()
        md = self.MAX_DETS[-1]
        groups, gt_rows, dt_rows, gt_off, dt_off = [], [], [], [0], [0]
        for ki, cat in enumerate(self.cat_ids):
            for img in self.img_ids:
                gt = self.gts.get((img, cat), [])
                dt = self.dts.get((img, cat), [])
                if not gt and not dt:
                    continue
                order = np.argsort([-d["score"] for d in dt], kind="mergesort")[:md]
                dt = [dt[i] for i in order]
                groups.append((ki, len(gt), [d["score"] for d in dt]))
                gt_rows += [(g["bbox"][0], g["bbox"][1], g["bbox"][2], g["bbox"][3], g["area"], int(g.get("iscrowd", 0)), g["ignore"]) for g in gt]
                dt_rows += [(d["bbox"][0], d["bbox"][1], d["bbox"][2], d["bbox"][3], d["area"]) for d in dt]
                gt_off.append(len(gt_rows))
                dt_off.append(len(dt_rows))
        T, A, P = len(self.IOU_THRS), len(self.AREA_RNG), len(groups)
        G, D = len(gt_rows), len(dt_rows)
        gt_np = np.asarray(gt_rows, np.float64).reshape(-1, 7)
        dt_np = np.asarray(dt_rows, np.float64).reshape(-1, 5)
        if np.any(gt_np[:, 5] != gt_np[:, 6]):
            raise RuntimeError("probenb200.COCOBBoxEval: 'ignore' differs from 'iscrowd' (the GPU path assumes cocoeval.py:246)")
        max_g = max([g[1] for g in groups] + [0])
        if max_g > lib.pe_coco_match_max_gt():
            raise RuntimeError("probenb200.COCOBBoxEval: %d ground-truth boxes in one (image, category) exceed the supported %d"
                               % (max_g, lib.pe_coco_match_max_gt()))
        dev = torch.device(device)

        def up(a, dtype):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype)
        t_gt_box, t_gt_area = up(gt_np[:, :4], torch.float64), up(gt_np[:, 4], torch.float64)
        t_gt_crowd = up(gt_np[:, 5].astype(np.uint8), torch.uint8)
        t_dt_box, t_dt_area = up(dt_np[:, :4], torch.float64), up(dt_np[:, 4], torch.float64)
        t_gt_off, t_dt_off = up(np.asarray(gt_off, np.int32), torch.int32), up(np.asarray(dt_off, np.int32), torch.int32)
        t_thr = up(self.IOU_THRS, torch.float64)
        t_rng = up(np.asarray(self.AREA_RNG, np.float64).reshape(-1), torch.float64)
        matched = torch.zeros((A, T, max(D, 1)), dtype=torch.uint8, device=dev)
        dt_ig = torch.zeros((A, T, max(D, 1)), dtype=torch.uint8, device=dev)
        gt_ig = torch.zeros((A, max(G, 1)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = lib.pe_coco_match(_lib.ptr(t_gt_box), _lib.ptr(t_gt_area), _lib.ptr(t_gt_crowd), _lib.ptr(t_gt_off), _lib.ptr(t_dt_box),
                                   _lib.ptr(t_dt_area), _lib.ptr(t_dt_off), P, _lib.ptr(t_thr), T, _lib.ptr(t_rng), A, max(D, 1), max(G, 1),
                                   max_g, _lib.ptr(matched), _lib.ptr(dt_ig), _lib.ptr(gt_ig), _lib.current_stream_ptr(dev))
        _lib.check(st, "pe_coco_match")
        matched, dt_ig, gt_ig = matched.cpu().numpy(), dt_ig.cpu().numpy().astype(bool), gt_ig.cpu().numpy().astype(bool)
        out = {(ki, ai): [] for ki in range(len(self.cat_ids)) for ai in range(A)}
        for p, (ki, n_gt, scores) in enumerate(groups):
            d0, d1, g0 = dt_off[p], dt_off[p + 1], gt_off[p]
            for ai in range(A):
                out[(ki, ai)].append({"dtm": np.where(matched[ai, :, d0:d1] > 0, 1, -1), "dt_scores": np.asarray(scores, np.float64),
                                      "g_ign": gt_ig[ai, g0:g0 + n_gt], "dt_ig": dt_ig[ai, :, d0:d1]})
        return out

    def evaluate(self, device=None):
        """``device``: None = match on the host (numpy); a CUDA device = match on the GPU (identical decisions)."""
        T, R, K, A, M = len(self.IOU_THRS), len(self.REC_THRS), len(self.cat_ids), len(self.AREA_RNG), len(self.MAX_DETS)
        precision = -np.ones((T, R, K, A, M))
        recall = -np.ones((T, K, A, M))
        matches = self._match_host() if device is None else self._match_gpu(device)
        for ki, cat in enumerate(self.cat_ids):
            for ai, rng in enumerate(self.AREA_RNG):
                evs = matches[(ki, ai)]
                if not evs:
                    continue
                for mi, md in enumerate(self.MAX_DETS):
                    scores = np.concatenate([e["dt_scores"][:md] for e in evs])
                    order = np.argsort(-scores, kind="mergesort")
                    dtm = np.concatenate([e["dtm"][:, :md] for e in evs], 1)[:, order]
                    dig = np.concatenate([e["dt_ig"][:, :md] for e in evs], 1)[:, order]
                    g_ign = np.concatenate([e["g_ign"] for e in evs])
                    npig = int(np.count_nonzero(~g_ign))
                    if npig == 0:
                        continue
                    tps = (dtm >= 0) & ~dig
                    fps = (dtm < 0) & ~dig
                    tp_sum = np.cumsum(tps, 1).astype(float)
                    fp_sum = np.cumsum(fps, 1).astype(float)
                    for ti in range(T):
                        tp, fp = tp_sum[ti], fp_sum[ti]
                        nd = len(tp)
                        rc = tp / npig
                        pr = tp / (fp + tp + np.spacing(1))
                        recall[ti, ki, ai, mi] = rc[-1] if nd else 0
                        pr = pr.tolist()
                        for i in range(nd - 1, 0, -1):
                            if pr[i] > pr[i - 1]:
                                pr[i - 1] = pr[i]
                        inds = np.searchsorted(rc, self.REC_THRS, side="left")
                        q = np.zeros(R)
                        for ri, pi in enumerate(inds):
                            if pi < nd:
                                q[ri] = pr[pi]
                        precision[ti, :, ki, ai, mi] = q
        self.precision, self.recall = precision, recall
        return self.summarize()

    def _stat(self, ap=True, iou=None, area=0, md=2, cat=None):
        s = self.precision if ap else self.recall
        if iou is not None:
            s = s[np.where(np.isclose(self.IOU_THRS, iou))[0]]
        s = s[:, :, :, area, md] if ap else s[:, :, area, md]
        if cat is not None:
            s = s[:, :, cat] if ap else s[:, cat]
        s = s[s > -1]
        return float(np.mean(s)) if s.size else -1.0

    def summarize(self):
        """The six numbers _derive_coco_results reports (x100), plus per-category AP."""
        res = {"AP": self._stat(), "AP50": self._stat(iou=.5), "AP75": self._stat(iou=.75),
               "APs": self._stat(area=1), "APm": self._stat(area=2), "APl": self._stat(area=3)}
        # COCOeval reports -1 for an empty slice and _derive_coco_results multiplies it like any other value
        # (FLIR_evaluation.py:274), so an empty area range reads -100; only the per-category AP becomes NaN (:296)
        res = {k: v * 100 for k, v in res.items()}
        for ci, cat in enumerate(self.cat_ids):
            v = self._stat(cat=ci)
            res["AP-%s" % cat] = v * 100 if v >= 0 else float("nan")
        return res
