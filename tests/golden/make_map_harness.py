"""mAP-agreement harness fixture (TEST INFRASTRUCTURE): does bf16 tensor-core inference move COCO AP relative to the
fp32 reference path?  (north_star: "identical COCO mAP to two decimals"; FLIR_evaluation.py:496-563, fast_rcnn.py:86-147.)

No FLIR images or checkpoints exist offline, so the harness builds a detector whose scores are bimodal like a trained
model's:

  1. ``scene(i)``: seeded synthetic 512x640 RGB+thermal pairs with structured content - warm "person" (tall), "bicycle"
     (two rings) and "car" (wide, with a darker cabin band) objects over a smooth background, plus clutter - and
     their ground-truth boxes.
  2. The R50-FPN backbone, FPN, RPN conv and fc1/fc2 stay at the seeded random values of ``weights.random_state_dict``
     (regenerated from the seed, never stored).  Only the small read-out layers are FITTED on ``N_TRAIN`` scenes, with the
     CPU oracle providing the features: RPN objectness / anchor deltas (256 -> 3 + 12) by logistic / ridge regression on
     sampled anchors (labels as rpn.py: IoU >= 0.7 fg, < 0.3 bg), then the box predictor (1024 -> 4 + 12 + 1) on the
     fitted RPN's proposals (labels as roi_heads.py: IoU >= 0.5) by softmax regression, ridge regression of the box
     deltas and a log-variance fit of the squared box residual.
  3. The fitted read-outs (~20 K floats per model) are stored in ``map_harness_heads.npz``; the fp32 oracle detections
     of both models on ``N_EVAL`` held-out scenes, their ProbEn fusion and the ground truth are stored in
     ``map_harness_oracle.npz``.

tests/test_map_parity_gpu.py rebuilds the same state dicts, runs the B200 engine on the same uint8 frames and compares
COCO AP (probenb200.evaluation.COCOBBoxEval) of GPU vs oracle detections against the same ground truth.

Run:  python tests/golden/make_map_harness.py [--eval-only]   (about 30 minutes on 8 cores)
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import detector_oracle as D  # noqa: E402
from oracle import proben_oracle as O  # noqa: E402
from oracle import resize_oracle as R  # noqa: E402
from probenb200 import weights  # noqa: E402

N_TRAIN, N_EVAL = 48, 256
SEEDS = (11, 12)            # RGB model, thermal model (the bench's seeds)
FRAME_HW = (512, 640)
NET_HW = (800, 1000)
FITTED_KEYS = ("proposal_generator.rpn_head.objectness_logits.weight", "proposal_generator.rpn_head.objectness_logits.bias",
               "proposal_generator.rpn_head.anchor_deltas.weight", "proposal_generator.rpn_head.anchor_deltas.bias",
               "roi_heads.box_predictor.cls_score.weight", "roi_heads.box_predictor.cls_score.bias",
               "roi_heads.box_predictor.bbox_pred.weight", "roi_heads.box_predictor.bbox_pred.bias",
               "roi_heads.box_predictor.var_pred.weight", "roi_heads.box_predictor.var_pred.bias")


# ------------------------------------------------------------------------------------------------ scenes
def scene(i, split):
    """-> rgb uint8 (512,640,3) BGR, thermal uint8 (512,640,3) replicated plane, gt float (n,5) x1,y1,x2,y2,class."""
    rng = np.random.default_rng(1000003 * (1 + split) + i)
    H, W = FRAME_HW
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    th = 70 + 25 * np.sin(xx / 97.0 + rng.uniform(0, 6)) * np.cos(yy / 71.0 + rng.uniform(0, 6)) + 0.04 * yy
    rgb = np.stack([60 + 30 * np.sin(xx / 131.0 + k + rng.uniform(0, 6)) + 0.05 * yy for k in range(3)], -1)
    gt = []
    n_obj = int(rng.integers(3, 9))
    tries = 0
    while len(gt) < n_obj and tries < 200:
        tries += 1
        c = int(rng.integers(0, 3))
        s = float(rng.uniform(0.6, 2.2))
        w, h = {0: (26 * s, 68 * s), 1: (62 * s, 40 * s), 2: (96 * s, 46 * s)}[c]
        x1, y1 = float(rng.uniform(4, W - w - 4)), float(rng.uniform(4, H - h - 4))
        box = np.array([x1, y1, x1 + w, y1 + h], np.float32)
        if any(_iou(box, g[:4]) > 0.05 for g in gt):
            continue
        heat = float(rng.uniform(150, 235))
        col = rng.uniform(120, 240, size=3).astype(np.float32)
        _draw(th, rgb, box, c, heat, col, xx, yy)
        gt.append(np.concatenate([box, [c]]).astype(np.float32))
    for _ in range(int(rng.integers(2, 6))):  # clutter: faint blobs that are not objects
        cx, cy, r = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(6, 22)
        m = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))
        th += 35 * m
        rgb += 30 * m[..., None]
    th += rng.normal(0, 3.0, th.shape)
    rgb += rng.normal(0, 4.0, rgb.shape)
    th8 = np.clip(th, 0, 255).astype(np.uint8)
    return np.clip(rgb, 0, 255).astype(np.uint8), np.repeat(th8[..., None], 3, axis=2), np.stack(gt)


def _iou(a, b):
    iw = min(a[2], b[2]) - max(a[0], b[0])
    ih = min(a[3], b[3]) - max(a[1], b[1])
    if iw <= 0 or ih <= 0:
        return 0.0
    inter = iw * ih
    return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter)


def _draw(th, rgb, box, c, heat, col, xx, yy):
    x1, y1, x2, y2 = box
    w, h = x2 - x1, y2 - y1
    u, v = (xx - x1) / w, (yy - y1) / h  # object coordinates in [0,1]
    inside = (u >= 0) & (u <= 1) & (v >= 0) & (v <= 1)
    if c == 0:    # person: head disc + torso + legs
        m = (((u - 0.5) ** 2 / 0.05 + (v - 0.1) ** 2 / 0.01) <= 1) | ((abs(u - 0.5) <= 0.42) & (v >= 0.2) & (v <= 0.62)) | \
            ((abs(u - 0.27) <= 0.17) & (v > 0.62)) | ((abs(u - 0.73) <= 0.17) & (v > 0.62))
    elif c == 1:  # bicycle: two rings + frame bar
        r1 = np.sqrt(((u - 0.22) * w) ** 2 + ((v - 0.68) * h) ** 2) / (0.3 * h)
        r2 = np.sqrt(((u - 0.78) * w) ** 2 + ((v - 0.68) * h) ** 2) / (0.3 * h)
        m = ((r1 <= 1) & (r1 >= 0.6)) | ((r2 <= 1) & (r2 >= 0.6)) | ((abs(v - 0.35) <= 0.08) & (u >= 0.2) & (u <= 0.8)) | \
            ((abs(u - 0.5) <= 0.05) & (v >= 0.0) & (v <= 0.7))
    else:         # car: body + cabin, cooler window band
        m = ((v >= 0.4) & (v <= 0.92)) | ((v < 0.4) & (u >= 0.22) & (u <= 0.78))
        m &= ~((v >= 0.12) & (v <= 0.34) & (u >= 0.3) & (u <= 0.7))
    m = m & inside
    th[m] = heat * (0.85 + 0.15 * v[m])
    for k in range(3):
        rgb[..., k][m] = col[k] * (0.8 + 0.2 * u[m])


def resized_input(frame_u8):
    """DefaultPredictor's Pillow BILINEAR resize of a 3-channel uint8 frame (oracle/resize_oracle.py, bit-identical to
    Pillow and to the engine's fused staging kernel) -> float CHW tensor."""
    out = R.pil_bilinear_resize_u8(frame_u8, NET_HW[0], NET_HW[1])
    return torch.from_numpy(out.astype(np.float32)).permute(2, 0, 1).contiguous()


# ------------------------------------------------------------------------------------------------ fitting
def pairwise_iou(a, b):
    from torchvision.ops import box_iou
    return box_iou(a, b)


def box_deltas(src, dst, wts):
    sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
    sx, sy = src[:, 0] + 0.5 * sw, src[:, 1] + 0.5 * sh
    dw, dh = dst[:, 2] - dst[:, 0], dst[:, 3] - dst[:, 1]
    dx, dy = dst[:, 0] + 0.5 * dw, dst[:, 1] + 0.5 * dh
    return torch.stack([wts[0] * (dx - sx) / sw, wts[1] * (dy - sy) / sh, wts[2] * torch.log(dw / sw), wts[3] * torch.log(dh / sh)], 1)


def ridge(X, Y, lam):
    """min |X w + b - Y|^2 + lam |w|^2  -> (w [out, in], b [out])."""
    Xm, Ym = X.mean(0, keepdim=True), Y.mean(0, keepdim=True)
    Xc, Yc = (X - Xm).double(), (Y - Ym).double()
    A = Xc.t() @ Xc + lam * len(X) * torch.eye(X.shape[1], dtype=torch.float64)
    w = torch.linalg.solve(A, Xc.t() @ Yc).float()
    return w.t().contiguous(), (Ym - Xm @ w).reshape(-1)


def softmax_fit(X, y, n_out, steps=400, lr=0.05, wd=1e-4, class_weight=None):
    mu, sd_ = X.mean(0, keepdim=True), X.std(0, keepdim=True) + 1e-6
    Xn = (X - mu) / sd_
    w = torch.zeros(n_out, X.shape[1], requires_grad=True)
    b = torch.zeros(n_out, requires_grad=True)
    opt = torch.optim.Adam([w, b], lr=lr)
    for _ in range(steps):
        opt.zero_grad()
        loss = F.cross_entropy(Xn @ w.t() + b, y, weight=class_weight) + wd * (w ** 2).sum()
        loss.backward()
        opt.step()
    w = (w.detach() / sd_).contiguous()          # fold the standardisation back: logits = w (x - mu) / sd + b
    return w, (b.detach() - (w * mu).sum(1)).contiguous()


def scale_gt(gt):
    s = NET_HW[0] / FRAME_HW[0]
    return torch.from_numpy(gt[:, :4]) * s, torch.from_numpy(gt[:, 4]).long()


@torch.no_grad()
def backbone_feats(x, sd, cfg):
    canvas, sizes = D.preprocess([x], cfg)
    feats = D.fpn(D.resnet(canvas, sd, cfg.depth), sd)
    return [feats["p%d" % l] for l in range(2, 7)], sizes


def fit_model(sd, modality, log):
    cfg = D.DetCfg()
    g = torch.Generator().manual_seed(5)
    # ---- pass 1: RPN read-outs on sampled anchors
    Xs, As, Ys, Ts = [], [], [], []
    cache = []
    for i in range(N_TRAIN):
        rgb, th, gt = scene(i, 0)
        x = resized_input(rgb if modality == 0 else th)
        with torch.no_grad():
            plist, sizes = backbone_feats(x, sd, cfg)
        gtb, gtc = scale_gt(gt)
        cache.append((plist[:4], sizes, gtb, gtc))
        for li, p in enumerate(plist):
            with torch.no_grad():
                t = D._conv(p, sd, "proposal_generator.rpn_head.conv", padding=1, relu=True)[0]  # (256, H, W)
            Hh, Ww = t.shape[1:]
            anchors = D.grid_anchors(Hh, Ww, D.STRIDES[li], D.ANCHOR_SIZES[li])               # (H*W*3, 4), a fastest
            iou = pairwise_iou(anchors, gtb)
            best, arg = iou.max(1)
            # rpn.py / matcher: IoU >= 0.7 fg, < 0.3 bg, plus the best anchor of every GT
            lab = torch.full((len(anchors),), -1, dtype=torch.long)
            lab[best < 0.3] = 0
            lab[best >= 0.7] = 1
            lab[iou.argmax(0)] = 1
            pos = (lab == 1).nonzero().squeeze(1)
            neg = (lab == 0).nonzero().squeeze(1)
            neg = neg[torch.randperm(len(neg), generator=g)[: max(256, 3 * len(pos))]]
            sel = torch.cat([pos, neg])
            cell, a = sel // 3, sel % 3
            feat = t.reshape(256, -1).t()[cell]
            Xs.append(feat); As.append(a); Ys.append(lab[sel])
            d = torch.zeros(len(sel), 4)
            d[: len(pos)] = box_deltas(anchors[pos], gtb[arg[pos]], (1.0, 1.0, 1.0, 1.0))
            Ts.append(d)
        log("  rpn features %d/%d" % (i + 1, N_TRAIN))
    X, A, Y, T = torch.cat(Xs), torch.cat(As), torch.cat(Ys), torch.cat(Ts)
    wo, bo = torch.zeros(3, 256), torch.zeros(3)
    wd_, bd_ = torch.zeros(12, 256), torch.zeros(12)
    for a in range(3):
        m = A == a
        w2, b2 = softmax_fit(X[m], Y[m], 2, steps=300)
        wo[a], bo[a] = w2[1] - w2[0], b2[1] - b2[0]           # two-class softmax -> one logit
        mp = m & (Y == 1)
        if int(mp.sum()) >= 8:
            w4, b4 = ridge(X[mp], T[mp], 1e-3)
            wd_[4 * a: 4 * a + 4], bd_[4 * a: 4 * a + 4] = w4, b4
    r = "proposal_generator.rpn_head"
    sd[r + ".objectness_logits.weight"] = wo.view(3, 256, 1, 1).clone()
    sd[r + ".objectness_logits.bias"] = bo.clone()
    sd[r + ".anchor_deltas.weight"] = wd_.view(12, 256, 1, 1).clone()
    sd[r + ".anchor_deltas.bias"] = bd_.clone()
    # ---- pass 2: box predictor on the fitted RPN's proposals (+ jittered GT boxes, as roi_heads.py appends GT)
    Fs, Ls, Ds, Ps, Gs = [], [], [], [], []
    for i, (p4, sizes, gtb, gtc) in enumerate(cache):
        with torch.no_grad():
            p6 = F.max_pool2d(p4[3], kernel_size=1, stride=2, padding=0)
            lg, dl = D.rpn_head(list(p4) + [p6], sd)
            props, _ = D.find_top_proposals(lg, dl, sizes, cfg)
        pb = props[0][0]
        # jittered copies of the ground truth (roi_heads.py appends the GT boxes to the proposals during training; many
        # jitters per box give the 1024-feature ridge regression enough foreground rows to pull near-duplicate proposals onto
        # the SAME refined box - without that, which duplicate survives NMS decides the output box and detections become chaotic)
        reps = 30
        sig = torch.cat([torch.full((len(gtb) * (reps // 2), 1), 0.05), torch.full((len(gtb) * (reps - reps // 2), 1), 0.15)])
        wh = torch.cat([gtb[:, 2:] - gtb[:, :2], gtb[:, 2:] - gtb[:, :2]], 1).repeat(reps, 1)
        jit = gtb.repeat(reps, 1) + sig * wh * torch.randn(len(gtb) * reps, 4, generator=g)
        jit = jit[(jit[:, 2] > jit[:, 0] + 4) & (jit[:, 3] > jit[:, 1] + 4)]
        pb = torch.cat([pb, D.clip_boxes(jit, sizes[0])])
        with torch.no_grad():
            pooled = D.roi_pool(list(p4), [pb])
            x = pooled.flatten(1)
            x = F.relu(F.linear(x, sd["roi_heads.box_head.fc1.weight"], sd["roi_heads.box_head.fc1.bias"]))
            x = F.relu(F.linear(x, sd["roi_heads.box_head.fc2.weight"], sd["roi_heads.box_head.fc2.bias"]))
        iou = pairwise_iou(pb, gtb)
        best, arg = iou.max(1)
        lab = torch.where(best >= 0.5, gtc[arg], torch.full_like(arg, 3))
        Fs.append(x); Ls.append(lab); Ps.append(pb); Gs.append(gtb[arg])
        Ds.append(box_deltas(pb, gtb[arg], (10.0, 10.0, 5.0, 5.0)))
        log("  roi features %d/%d (fg %d of %d)" % (i + 1, N_TRAIN, int((lab < 3).sum()), len(lab)))
    X, L, Dl, P, G = torch.cat(Fs), torch.cat(Ls), torch.cat(Ds), torch.cat(Ps), torch.cat(Gs)
    wc, bc = softmax_fit(X, L, 4, steps=500, lr=0.03)
    wb, bb = torch.zeros(12, 1024), torch.zeros(12)
    for c in range(3):
        m = L == c
        if int(m.sum()) >= 16:
            w4, b4 = ridge(X[m], Dl[m], 1e-2)
            wb[4 * c: 4 * c + 4], bb[4 * c: 4 * c + 4] = w4, b4
    q = "roi_heads.box_predictor"
    sd[q + ".cls_score.weight"], sd[q + ".cls_score.bias"] = wc.clone(), bc.clone()
    sd[q + ".bbox_pred.weight"], sd[q + ".bbox_pred.bias"] = wb.clone(), bb.clone()
    # variance head: log of the mean squared corner residual (frame pixels) of the fitted box regression on fg ROIs
    fg = L < 3
    pred = torch.stack([D.apply_deltas((X[fg] @ wb.t() + bb)[:, 4 * c: 4 * c + 4], P[fg], (10.0, 10.0, 5.0, 5.0)) for c in range(3)], 1)
    pred = pred[torch.arange(int(fg.sum())), L[fg]]
    res2 = (((pred - G[fg]) * (FRAME_HW[0] / NET_HW[0])) ** 2).mean(1).clamp(min=1e-2)
    wv, bv = ridge(X[fg], torch.log(res2)[:, None], 1e-2)
    sd[q + ".var_pred.weight"], sd[q + ".var_pred.bias"] = wv.clone(), bv.clone()
    return sd


def fitted_state_dict(model_index, heads=None):
    """The harness model m (0: RGB, 1: thermal): seeded random trunk + fitted read-outs from map_harness_heads.npz."""
    sd = weights.random_state_dict(50, 3, 3, seed=SEEDS[model_index])
    heads = heads if heads is not None else np.load(os.path.join(HERE, "map_harness_heads.npz"))
    for k in FITTED_KEYS:
        sd[k] = torch.from_numpy(heads["m%d.%s" % (model_index, k)]).clone()
    return sd


def infos_from(res):
    return {"bbox": res["pred_boxes"].double().tolist(), "score": res["scores"].double().tolist(),
            "class": res["pred_classes"].tolist(), "prob": res["prob_score"].double().tolist(), "vars": res["vars"].double().tolist()}


def main():
    t0 = time.time()

    def log(msg):
        print("[%6.0fs] %s" % (time.time() - t0, msg), flush=True)

    torch.set_num_threads(os.cpu_count() or 1)
    heads = {}
    sds = []
    if "--eval-only" in sys.argv:  # keep the fitted read-outs, regenerate the oracle detections only
        sds = [fitted_state_dict(m) for m in range(2)]
    else:
        for m in range(2):
            log("fitting model %d (seed %d)" % (m, SEEDS[m]))
            sd = fit_model(weights.random_state_dict(50, 3, 3, seed=SEEDS[m]), m, log)
            for k in FITTED_KEYS:
                heads["m%d.%s" % (m, k)] = sd[k].numpy().astype(np.float32)
            sds.append(sd)
        np.savez_compressed(os.path.join(HERE, "map_harness_heads.npz"), **heads)
    cfg = D.DetCfg()
    out = {"gt_offsets": [0], "gt": []}
    for m in range(2):
        for k in ("boxes", "scores", "classes", "probs", "vars", "logits"):
            out["m%d_%s" % (m, k)] = []
        out["m%d_offsets" % m] = [0]
    for k in ("boxes", "scores", "classes"):
        out["fused_" + k] = []
    out["fused_offsets"] = [0]
    for i in range(N_EVAL):
        rgb, th, gt = scene(i, 1)
        out["gt"].append(gt)
        out["gt_offsets"].append(out["gt_offsets"][-1] + len(gt))
        infos = []
        for m, fr in enumerate((rgb, th)):
            r = D.detector_forward([resized_input(fr)], [FRAME_HW], sds[m], cfg)[0]
            infos.append(infos_from(r))
            out["m%d_boxes" % m].append(r["pred_boxes"].numpy()); out["m%d_scores" % m].append(r["scores"].numpy())
            out["m%d_classes" % m].append(r["pred_classes"].numpy().astype(np.int32)); out["m%d_probs" % m].append(r["prob_score"].numpy())
            out["m%d_vars" % m].append(r["vars"].numpy().reshape(-1)); out["m%d_logits" % m].append(r["class_logits"].numpy())
            out["m%d_offsets" % m].append(out["m%d_offsets" % m][-1] + len(r["scores"]))
        f = O.late_fusion_dispatch(("probEn", "v-avg"), infos)
        fb, fs, fc = (np.zeros((0, 4)), np.zeros(0), np.zeros(0)) if f is None else (np.asarray(f[0], np.float64).reshape(-1, 4), np.asarray(f[1], np.float64), np.asarray(f[2], np.float64))
        out["fused_boxes"].append(fb); out["fused_scores"].append(fs); out["fused_classes"].append(fc)
        out["fused_offsets"].append(out["fused_offsets"][-1] + len(fs))
        log("eval scene %d/%d: %d + %d detections -> %d fused, %d gt" % (i + 1, N_EVAL, len(infos[0]["score"]), len(infos[1]["score"]), len(fs), len(gt)))
    packed = {}
    for k, v in out.items():
        packed[k] = np.asarray(v, np.int64) if k.endswith("offsets") else np.concatenate(v)
    np.savez_compressed(os.path.join(HERE, "map_harness_oracle.npz"), **packed)
    log("done")


if __name__ == "__main__":
    main()
