"""CPU oracle for the per-modality Faster R-CNN detector (ResNet-FPN + RPN + StandardROIHeads with the
fork's variance head).  TEST INFRASTRUCTURE ONLY - fp32 PyTorch on the CPU.

Functional restatement (from a state dict with the reference's parameter names) of, paths relative to
/root/reference:

  preprocess            <- detectron2/modeling/meta_arch/rcnn.py:269-286, structures/image_list.py:51-102
  frozen_bn / conv      <- layers/batch_norm.py:45-64, layers/wrappers.py (Conv2d + norm + activation)
  resnet                <- modeling/backbone/resnet.py:369-384 (BasicStem), :160-221 (BottleneckBlock,
                           stride_in_1x1=True), :474-568 (build_resnet_backbone: R50 [3,4,6,3], R101 [3,4,23,3])
  fpn                   <- modeling/backbone/fpn.py:110-145, :166-178 (LastLevelMaxPool)
  rpn_head / anchors    <- modeling/proposal_generator/rpn.py:74-85, modeling/anchor_generator.py:130-199
  apply_deltas          <- modeling/box_regression.py:78-115
  find_top_proposals    <- modeling/proposal_generator/rpn_outputs.py:52-162, :409-451
  roi_pool              <- modeling/poolers.py:13-81,180-235 (+ csrc/ROIAlign/ROIAlign_cpu.cpp:20-218 through
                           torchvision.ops.roi_align, bit-identical on tests/test_roi_align.py's tables)
  box_head / predictor  <- modeling/roi_heads/box_head.py:73-81, fast_rcnn.py:531-545 (exp(var_pred))
  fast_rcnn_inference   <- modeling/roi_heads/fast_rcnn.py:86-147,345-360,417-452 (incl. the ``vars = variance[keep]``
                           indexing quirk)
  postprocess           <- modeling/postprocessing.py:8-52
  middle fusion         <- rcnn.py:240-248 (shared backbone on both halves, channel concat)

Third-party arithmetic: torchvision ``ops.boxes.batched_nms`` / ``ops.roi_align`` (reference pins 0.13, installed
0.26; same algorithms).  Parity status: PINNED against the reference's own ``GeneralizedRCNN`` loaded by file path
(tests/golden/make_golden_detector.py -> tests/golden/detector_golden.npz; tests/test_oracle_detector.py).
"""
import math
import sys

import torch
import torch.nn.functional as F
from torchvision.ops import boxes as box_ops
from torchvision.ops import roi_align

SCALE_CLAMP = math.log(1000.0 / 16)
STAGE_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}
STRIDES = (4, 8, 16, 32, 64)
ANCHOR_SIZES = (32, 64, 128, 256, 512)
ANCHOR_RATIOS = (0.5, 1.0, 2.0)


class DetCfg:
    """The configuration keys the inference path reads (config/defaults.py + Base-RCNN-FPN.yaml +
    demo_FLIR_save_predictions.py:49-73 overrides)."""

    def __init__(self, depth=50, in_channels=3, num_classes=3, pixel_mean=(103.530, 116.280, 123.675),
                 pixel_std=(1.0, 1.0, 1.0), score_thresh=0.5, nms_thresh=0.5, detections_per_image=100,
                 pre_nms_topk=1000, post_nms_topk=1000, rpn_nms_thresh=0.7, middle_fusion=False):
        self.depth, self.in_channels, self.num_classes = depth, in_channels, num_classes
        self.pixel_mean, self.pixel_std = tuple(pixel_mean), tuple(pixel_std)
        self.score_thresh, self.nms_thresh, self.detections_per_image = score_thresh, nms_thresh, detections_per_image
        self.pre_nms_topk, self.post_nms_topk, self.rpn_nms_thresh = pre_nms_topk, post_nms_topk, rpn_nms_thresh
        self.middle_fusion = middle_fusion


# ---- precision bisect (test infrastructure for tests/golden/bisect_bf16.py and the mAP-parity harness) ---------------------------
# EMULATE names the stages whose arithmetic is restated the way the B200 engine computes it: FrozenBN folded into the weights,
# operands rounded to bf16 (stem: fp16), fp32 accumulation, bias / residual / ReLU in fp32, ONE rounding of the stored activation
# to bf16.  Stages: "stem", "res2".."res5", "fpn", "rpn", "head".  Empty set = the fp32 reference arithmetic (the pinned oracle).
EMULATE = set()


def _bf(t):
    return t.bfloat16().float()


def _stage_of(name):
    for st in ("stem", "res2", "res3", "res4", "res5"):
        if "." + st + "." in name:
            return st
    if "fpn_" in name:
        return "fpn"
    if "rpn_head" in name:
        return "rpn"
    return "head"


def _conv_emulated(x, sd, name, stride, padding, relu, add=None, round_out=True):
    w = sd[name + ".weight"]
    if name + ".norm.weight" in sd:
        scale = sd[name + ".norm.weight"] / torch.sqrt(sd[name + ".norm.running_var"] + 1e-5)
        b = sd[name + ".norm.bias"] - sd[name + ".norm.running_mean"] * scale
        w = w * scale.view(-1, 1, 1, 1)
    else:
        b = sd.get(name + ".bias")
    if ".stem." in name:
        x, w = x.half().float(), w.half().float()
    else:
        x, w = _bf(x), _bf(w)
    y = F.conv2d(x, w, b, stride=stride, padding=padding)
    if add is not None:
        y = y + add
    if relu:
        y = F.relu(y)
    return _bf(y) if round_out else y


def _conv(x, sd, name, stride=1, padding=0, relu=False):
    if _stage_of(name) in EMULATE:
        return _conv_emulated(x, sd, name, stride, padding, relu, round_out="rpn_head.objectness" not in name and "rpn_head.anchor" not in name)
    y = F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)
    if name + ".norm.weight" in sd:  # FrozenBatchNorm2d, eps 1e-5
        y = F.batch_norm(y, sd[name + ".norm.running_mean"], sd[name + ".norm.running_var"],
                         sd[name + ".norm.weight"], sd[name + ".norm.bias"], training=False, eps=1e-5)
    return F.relu(y) if relu else y


def resnet(x, sd, depth, prefix="backbone.bottom_up"):
    x = _conv(x, sd, prefix + ".stem.conv1", stride=2, padding=3, relu=True)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = {}
    if EMULATE:
        x = _bf(x)
    for si, nblocks in enumerate(STAGE_BLOCKS[depth]):
        stage = "res%d" % (si + 2)
        for b in range(nblocks):
            p = "%s.%s.%d" % (prefix, stage, b)
            stride = 2 if (b == 0 and si > 0) else 1
            if stage in EMULATE:  # engine: conv3 accumulator + bias (+ shortcut GEMM / residual tile) -> ReLU -> one bf16 rounding
                y = _conv(x, sd, p + ".conv1", stride=stride, relu=True)
                y = _conv(y, sd, p + ".conv2", padding=1, relu=True)
                if (p + ".shortcut.weight") in sd:
                    short = _conv_emulated(x, sd, p + ".shortcut", stride, 0, False, round_out=False)
                else:
                    short = x
                x = _conv_emulated(y, sd, p + ".conv3", 1, 0, True, add=short)
                continue
            short = _conv(x, sd, p + ".shortcut", stride=stride) if (p + ".shortcut.weight") in sd else x
            y = _conv(x, sd, p + ".conv1", stride=stride, relu=True)     # stride lives in the 1x1 (stride_in_1x1)
            y = _conv(y, sd, p + ".conv2", padding=1, relu=True)
            y = _conv(y, sd, p + ".conv3")
            x = F.relu(y + short)
        outs[stage] = x
    return outs


def fpn(c, sd, prefix="backbone"):
    prev = _conv(c["res5"], sd, prefix + ".fpn_lateral5")
    p = {"p5": _conv(prev, sd, prefix + ".fpn_output5", padding=1)}
    for lvl in (4, 3, 2):
        if "fpn" in EMULATE:  # engine: lateral accumulator + bias + nearest-2x top-down tile, one bf16 rounding
            prev = _conv_emulated(c["res%d" % lvl], sd, prefix + ".fpn_lateral%d" % lvl, 1, 0, False,
                                  add=F.interpolate(prev, scale_factor=2, mode="nearest"))
        else:
            prev = _conv(c["res%d" % lvl], sd, prefix + ".fpn_lateral%d" % lvl) + F.interpolate(prev, scale_factor=2, mode="nearest")
        p["p%d" % lvl] = _conv(prev, sd, prefix + ".fpn_output%d" % lvl, padding=1)
    p["p6"] = F.max_pool2d(p["p5"], kernel_size=1, stride=2, padding=0)
    return p


def cell_anchors(size):
    rows = []
    for r in ANCHOR_RATIOS:
        w = math.sqrt(size ** 2.0 / r)
        h = r * w
        rows.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(rows)


def grid_anchors(H, W, stride, size):
    sx = torch.arange(0, W * stride, step=stride, dtype=torch.float32)
    sy = torch.arange(0, H * stride, step=stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
    return (shifts.view(-1, 1, 4) + cell_anchors(size).view(1, -1, 4)).reshape(-1, 4)


def apply_deltas(deltas, boxes, weights):
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * widths
    cy = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
    dw = torch.clamp(deltas[:, 2::4] / ww, max=SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3::4] / wh, max=SCALE_CLAMP)
    pcx = dx * widths[:, None] + cx[:, None]
    pcy = dy * heights[:, None] + cy[:, None]
    pw = torch.exp(dw) * widths[:, None]
    ph = torch.exp(dh) * heights[:, None]
    out = torch.zeros_like(deltas)
    out[:, 0::4] = pcx - 0.5 * pw
    out[:, 1::4] = pcy - 0.5 * ph
    out[:, 2::4] = pcx + 0.5 * pw
    out[:, 3::4] = pcy + 0.5 * ph
    return out


def clip_boxes(b, hw):
    h, w = hw
    b = b.clone()
    b[:, 0::2] = b[:, 0::2].clamp(min=0, max=w)
    b[:, 1::2] = b[:, 1::2].clamp(min=0, max=h)
    return b


def rpn_head(feats, sd, prefix="proposal_generator.rpn_head"):
    logits, deltas = [], []
    for x in feats:
        t = _conv(x, sd, prefix + ".conv", padding=1, relu=True)
        logits.append(_conv(t, sd, prefix + ".objectness_logits"))
        deltas.append(_conv(t, sd, prefix + ".anchor_deltas"))
    return logits, deltas


def find_top_proposals(logits, deltas, image_sizes, cfg):
    """Returns per image (proposal_boxes (n,4), objectness_logits (n,)) plus the per-level intermediates."""
    N = logits[0].shape[0]
    lvl_scores, lvl_boxes, lvl_ids = [], [], []
    for li, (lg, dl) in enumerate(zip(logits, deltas)):
        _, A, H, W = lg.shape
        anchors = grid_anchors(H, W, STRIDES[li], ANCHOR_SIZES[li])
        d = dl.view(N, A, 4, H, W).permute(0, 3, 4, 1, 2).reshape(-1, 4)
        props = apply_deltas(d, anchors.repeat(N, 1), (1.0, 1.0, 1.0, 1.0)).view(N, -1, 4)
        s = lg.permute(0, 2, 3, 1).reshape(N, -1)
        k = min(cfg.pre_nms_topk, s.shape[1])
        s_sorted, idx = s.sort(descending=True, dim=1)
        lvl_scores.append(s_sorted[:, :k])
        lvl_boxes.append(torch.gather(props, 1, idx[:, :k, None].expand(-1, -1, 4)))
        lvl_ids.append(torch.full((k,), li, dtype=torch.int64))
    scores = torch.cat(lvl_scores, 1)
    boxes = torch.cat(lvl_boxes, 1)
    lvls = torch.cat(lvl_ids)
    out = []
    for n in range(N):
        b, s, l = boxes[n], scores[n], lvls
        ok = torch.isfinite(b).all(dim=1) & torch.isfinite(s)
        b, s, l = b[ok], s[ok], l[ok]
        b = clip_boxes(b, image_sizes[n])
        keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
        b, s, l = b[keep], s[keep], l[keep]
        k = box_ops.batched_nms(b, s, l, cfg.rpn_nms_thresh)[: cfg.post_nms_topk]
        out.append((b[k], s[k]))
    return out, {"topk_scores": scores, "topk_boxes": boxes}


def assign_levels(boxes):
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    lv = torch.floor(4 + torch.log2(torch.sqrt(area) / 224 + sys.float_info.epsilon))
    return torch.clamp(lv, min=2, max=5).to(torch.int64) - 2


def roi_pool(feats, boxes_per_image):
    """feats: [p2..p5] NCHW; returns (R, C, 7, 7) in image-major order."""
    rois = torch.cat([torch.cat([torch.full((len(b), 1), float(i)), b], 1) for i, b in enumerate(boxes_per_image)])
    lv = assign_levels(rois[:, 1:])
    out = torch.zeros((len(rois), feats[0].shape[1], 7, 7))
    for li, f in enumerate(feats):
        idx = torch.nonzero(lv == li).squeeze(1)
        if len(idx):
            out[idx] = roi_align(f, rois[idx], (7, 7), spatial_scale=1.0 / STRIDES[li], sampling_ratio=0, aligned=True)
    return out


def box_head(pooled, sd, prefix="roi_heads"):
    x = pooled.flatten(1)
    p = prefix + ".box_predictor"
    if "head" in EMULATE:  # engine: bf16 ROI features / fc operands, fp32 accumulate, bf16 fc1 / fc2 outputs, fp32 predictor outputs
        x = _bf(x)
        x = _bf(F.relu(F.linear(x, _bf(sd[prefix + ".box_head.fc1.weight"]), sd[prefix + ".box_head.fc1.bias"])))
        x = _bf(F.relu(F.linear(x, _bf(sd[prefix + ".box_head.fc2.weight"]), sd[prefix + ".box_head.fc2.bias"])))
        logits = F.linear(x, _bf(sd[p + ".cls_score.weight"]), sd[p + ".cls_score.bias"])
        deltas = F.linear(x, _bf(sd[p + ".bbox_pred.weight"]), sd[p + ".bbox_pred.bias"])
        var = torch.exp(F.linear(x, _bf(sd[p + ".var_pred.weight"]), sd[p + ".var_pred.bias"]))
        return logits, deltas, var
    x = F.relu(F.linear(x, sd[prefix + ".box_head.fc1.weight"], sd[prefix + ".box_head.fc1.bias"]))
    x = F.relu(F.linear(x, sd[prefix + ".box_head.fc2.weight"], sd[prefix + ".box_head.fc2.bias"]))
    logits = F.linear(x, sd[p + ".cls_score.weight"], sd[p + ".cls_score.bias"])
    deltas = F.linear(x, sd[p + ".bbox_pred.weight"], sd[p + ".bbox_pred.bias"])
    var = torch.exp(F.linear(x, sd[p + ".var_pred.weight"], sd[p + ".var_pred.bias"]))
    return logits, deltas, var


def fast_rcnn_inference_image(boxes, probs, logits, variance, image_size, cfg):
    """One image, batch-1 indexing semantics (fast_rcnn.py:86-147)."""
    ok = torch.isfinite(boxes).all(dim=1) & torch.isfinite(probs).all(dim=1)
    boxes, probs = boxes[ok], probs[ok]
    scores = probs[:, :-1]
    K = boxes.shape[1] // 4
    boxes = clip_boxes(boxes.reshape(-1, 4), image_size).view(-1, K, 4)
    mask = scores > cfg.score_thresh
    inds = mask.nonzero()
    cand_logits = logits[inds[:, 0]]
    cand_probs = scores[inds[:, 0]]
    cand_boxes = boxes[mask]
    cand_scores = scores[mask]
    keep = box_ops.batched_nms(cand_boxes, cand_scores, inds[:, 1], cfg.nms_thresh)[: cfg.detections_per_image]
    return {"pred_boxes": cand_boxes[keep], "scores": cand_scores[keep], "pred_classes": inds[keep, 1],
            "class_logits": cand_logits[keep], "prob_score": cand_probs[keep],
            "vars": variance[keep],  # sic: candidate-list indices into the per-ROI tensor (quirk 1)
            "roi_index": inds[keep, 0]}


def postprocess(det, image_size, out_h, out_w):
    sx, sy = out_w / image_size[1], out_h / image_size[0]
    b = det["pred_boxes"].clone()
    b[:, 0::2] *= sx
    b[:, 1::2] *= sy
    b = clip_boxes(b, (out_h, out_w))
    keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    out = {k: v[keep] for k, v in det.items()}
    out["pred_boxes"] = b[keep]
    return out


def preprocess(images, cfg):
    """images: list of (C,h,w) float tensors (already resized).  Returns canvas (N,C,H32,W32) + sizes."""
    mean = torch.tensor(cfg.pixel_mean).view(-1, 1, 1)
    std = torch.tensor(cfg.pixel_std).view(-1, 1, 1)
    normed = [(x - mean) / std for x in images]
    H = max(x.shape[1] for x in normed)
    W = max(x.shape[2] for x in normed)
    H, W = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    canvas = torch.zeros((len(normed), normed[0].shape[0], H, W))
    for i, x in enumerate(normed):
        canvas[i, :, : x.shape[1], : x.shape[2]] = x
    return canvas, [(x.shape[1], x.shape[2]) for x in normed]


@torch.no_grad()
def detector_forward(images, out_sizes, sd, cfg, return_intermediates=False):
    """GeneralizedRCNN.inference (rcnn.py:219-267) with per-image (batch-1) head semantics.
    images: list of (C,h,w) float32 tensors; out_sizes: list of (height, width) to rescale to."""
    canvas, sizes = preprocess(images, cfg)
    if cfg.middle_fusion:
        fa = fpn(resnet(canvas[:, :3], sd, cfg.depth), sd)
        fb = fpn(resnet(canvas[:, 3:], sd, cfg.depth), sd)
        feats = {k: torch.cat((fa[k], fb[k]), 1) for k in fa}
    else:
        feats = fpn(resnet(canvas, sd, cfg.depth), sd)
    plist = [feats["p%d" % i] for i in range(2, 7)]
    logits, deltas = rpn_head(plist, sd)
    proposals, rpn_dbg = find_top_proposals(logits, deltas, sizes, cfg)
    pooled = roi_pool(plist[:4], [p[0] for p in proposals])
    cls_logits, box_deltas, var = box_head(pooled, sd)
    results, start = [], 0
    for n, (pb, _) in enumerate(proposals):
        r = len(pb)
        lg, dl, vr = cls_logits[start:start + r], box_deltas[start:start + r], var[start:start + r]
        start += r
        boxes = apply_deltas(dl, pb, (10.0, 10.0, 5.0, 5.0))
        probs = F.softmax(lg, dim=-1)
        det = fast_rcnn_inference_image(boxes, probs, lg, vr, sizes[n], cfg)
        results.append(postprocess(det, sizes[n], out_sizes[n][0], out_sizes[n][1]))
    if return_intermediates:
        return results, {"canvas": canvas, "features": feats, "rpn_logits": logits, "rpn_deltas": deltas,
                         "proposals": proposals, "pooled": pooled, "cls_logits": cls_logits,
                         "box_deltas": box_deltas, "var": var, **rpn_dbg}
    return results
