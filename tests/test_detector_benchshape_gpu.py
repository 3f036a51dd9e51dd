"""Detector parity AT THE BENCHMARKED SHAPE: 8 frames of 512x640 resized to 800x1000 on the 800x1024 canvas, R50-FPN and
R101-FPN, engine (bf16 tensor cores) vs the fp32 CPU oracle (oracle/detector_oracle.py, pinned to the reference's
GeneralizedRCNN).  Compared: FPN outputs p2..p5, RPN objectness logits and anchor deltas on all five levels, and the box
head's outputs (class logits, box deltas, log-variance) on the ENGINE's own proposals (the oracle pools its fp32 features
over the same boxes, so RPN-NMS tie flips - SURVEY.md quirk 9 - cannot leak into the comparison).

Tolerances are relative L2 errors of whole tensors, the noise floor of ~50 (R50) / ~100 (R101) bf16-rounded layers:
features and RPN outputs 4e-2 (R50) / 6e-2 (R101); head outputs 6e-2 / 8e-2.  The detections themselves are compared
through COCO AP in tests/test_map_parity_gpu.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import make_map_harness as H
from oracle import detector_oracle as D
from probenb200 import detector, weights

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-12))


@pytest.mark.parametrize("depth", [50, 101])
def test_features_rpn_and_head_at_bench_shape(depth):
    B, K = 8, 3
    sd = weights.random_state_dict(depth, 3, K, seed=11)
    cfg = D.DetCfg(depth=depth)
    frames = np.stack([H.scene(i, 2)[i % 2] for i in range(B)])          # RGB and thermal frames alternate
    imgs = [H.resized_input(f) for f in frames]                           # Pillow-exact resize, bit-identical to the engine's
    torch.set_num_threads(max(1, torch.get_num_threads()))
    res, inter = D.detector_forward(imgs, [H.FRAME_HW] * B, sd, cfg, return_intermediates=True)
    det = detector.Detector(sd, depth=depth, num_classes=K, max_batch=B, canvas=(800, 1024))
    out = det.forward_frames_device(torch.from_numpy(frames).cuda(), H.NET_HW, round_u8=True)
    torch.cuda.synchronize()
    tol_f, tol_h = (4e-2, 6e-2) if depth == 50 else (6e-2, 8e-2)
    errs = {}
    for l in (2, 3, 4, 5):
        raw, dims, _ = det.buffer("pout%d_0" % l)
        got = raw.view(torch.bfloat16).view(*dims).float().cpu().permute(0, 3, 1, 2)
        want = inter["features"]["p%d" % l]
        assert got.shape == want.shape == (B, 256, 800 >> l, 1024 >> l)
        errs["p%d" % l] = rel_err(got, want)
        assert errs["p%d" % l] < tol_f, errs
    for l in range(2, 7):
        raw, dims, _ = det.buffer("rpn_out%d" % l)
        got = raw.view(torch.float32).view(*dims).cpu()
        errs["rpn_logits%d" % l] = rel_err(got[..., :3], inter["rpn_logits"][l - 2].permute(0, 2, 3, 1))
        errs["rpn_deltas%d" % l] = rel_err(got[..., 4:16], inter["rpn_deltas"][l - 2].permute(0, 2, 3, 1))
        assert errs["rpn_logits%d" % l] < tol_f and errs["rpn_deltas%d" % l] < tol_f, errs
    # box head on the engine's proposals
    raw, dims, _ = det.buffer("head_out")
    head = raw.view(torch.float32).view(dims[0], dims[3]).cpu()
    props = det.buffer("proposals")[0].view(torch.float32).view(B, 1000, 4).cpu()
    pcount = det.buffer("prop_count")[0].view(torch.int32).cpu().tolist()
    assert min(pcount) > 100, pcount
    plist = [inter["features"]["p%d" % l] for l in (2, 3, 4, 5)]
    pooled = D.roi_pool(plist, [props[n, : pcount[n]] for n in range(B)])
    lg, dl, var = D.box_head(pooled, sd)
    got = torch.cat([head[n * 1000: n * 1000 + pcount[n]] for n in range(B)])
    errs["cls_logits"] = rel_err(got[:, : K + 1], lg)
    errs["box_deltas"] = rel_err(got[:, K + 1: K + 1 + 4 * K], dl)
    errs["log_var"] = rel_err(got[:, K + 1 + 4 * K], torch.log(var[:, 0]))
    print("bench-shape parity R%d:" % depth, {k: round(v, 5) for k, v in errs.items()})
    assert errs["cls_logits"] < tol_h and errs["box_deltas"] < tol_h and errs["log_var"] < tol_h, errs
    # proposals: same count regime as the oracle, and the top-scoring oracle proposals are found by the engine
    from torchvision.ops import box_iou
    for n in range(B):
        want_boxes = inter["proposals"][n][0]
        assert abs(pcount[n] - len(want_boxes)) <= 0.05 * len(want_boxes) + 5, (pcount[n], len(want_boxes))
        top = want_boxes[:100]
        iou = box_iou(top, props[n, : pcount[n]])
        assert float((iou.max(dim=1).values > 0.9).float().mean()) >= 0.75  # RPN-NMS membership is not bit-reproducible
    assert int(out.counts.sum()) >= 0
