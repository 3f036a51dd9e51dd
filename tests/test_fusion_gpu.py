"""GPU parity of pe_fuse_batch (through the C ABI) against the golden vectors from the reference, the
oracle on fresh seeded inputs, and size-independent properties at benchmark scale."""
import numpy as np
import pytest
import torch

import proben_cases as pc
from oracle import proben_oracle as O
from probenb200 import fusion, synth

pytestmark = pytest.mark.gpu
TOL = 1e-4  # BASELINE.json: scores / boxes within 1e-4 of the reference on identical saved predictions


@pytest.mark.parametrize("name", pc.SETS)
def test_cuda_matches_reference_golden(golden_proben, name):
    packed = pc.golden_inputs(golden_proben, name)
    dev = fusion.to_device(packed)
    for sm in pc.SCORES:
        for bm in pc.BOXES:
            buf = fusion.fuse_packed(dev, (sm, bm))
            got = fusion.unpack_results(packed, buf)
            want = pc.golden_outputs(golden_proben, name, sm, bm)
            for b in range(packed["B"]):
                pc.assert_same_detections(got[b], want[b], TOL, "%s %s/%s img %d" % (name, sm, bm, b))


@pytest.mark.parametrize("M,seed", [(2, 101), (3, 102)])
def test_cuda_matches_oracle_seeded(M, seed):
    dets = synth.synth_model_detections(300, M, seed=seed)
    images = [[synth.image_info(d, i) for d in dets] for i in range(300)]
    for sm, bm in (("probEn", "v-avg"), ("avg", "s-avg"), ("max", "avg"), ("max", "argmax"), ("probEn", "argmax")):
        got = fusion.late_fusion_batch((sm, bm), images)
        for b, infos in enumerate(images):
            want = O.late_fusion_dispatch((sm, bm), infos)
            pc.assert_same_detections(got[b], want, TOL, "M%d %s/%s img %d" % (M, sm, bm, b))


@pytest.mark.parametrize("M,per_model", [(2, 17), (2, 100), (2, 128), (3, 43), (3, 85), (2, 129)])
def test_block_kernels_match_oracle(M, per_model):
    """33..256 detections per image take fuse_mid_kernel (thread per detection, bit fixed-point clustering), 257..1024 the
    block kernel with the serial head scan: both against the oracle at the pipeline's regime (100 per model) and at the
    boundaries of the two ranges (34, 256, 129, 255, 258 detections)."""
    dets = synth.synth_model_detections(3, M, seed=300 + per_model, force_count=per_model)
    images = [[synth.image_info(d, i) for d in dets] for i in range(3)]
    for sm, bm in (("probEn", "v-avg"), ("avg", "s-avg"), ("max", "avg"), ("max", "argmax"), ("probEn", "argmax")):
        got = fusion.late_fusion_batch((sm, bm), images)
        for b, infos in enumerate(images):
            want = O.late_fusion_dispatch((sm, bm), infos)
            pc.assert_same_detections(got[b], want, TOL, "M%d n%d %s/%s img %d" % (M, per_model, sm, bm, b))


def test_mid_kernel_cross_class_border_contact():
    """fuse_mid_kernel walks same-class partners only; boxes of DIFFERENT classes that touch through the +1 border of the
    reference's class-offset tiles (a box ending at the far corner of its 640 x 512 tile, a box starting at the near corner of
    the next: demo_probEn.py:100-123) must still be clustered as the reference does.  40+ detections per image so that the
    block path runs; degenerate corner boxes of classes 0/1 and 1/2 plus near misses injected into both models."""
    dets = synth.synth_model_detections(3, 2, seed=77, force_count=20)
    images = [[synth.image_info(d, i) for d in dets] for i in range(3)]
    extra = [  # (model, box, class, score)
        (0, [640.0, 512.0, 640.0, 512.0], 0, 0.93), (1, [0.0, 0.0, 0.0, 0.0], 1, 0.81),          # IoU 1 through the border
        (0, [639.5, 511.5, 640.0, 512.0], 1, 0.77), (1, [0.0, 0.0, 0.25, 0.25], 2, 0.66),        # overlap 1 px^2, IoU < 0.5
        (1, [638.0, 510.0, 640.0, 512.0], 0, 0.71), (0, [0.0, 0.0, 1.0, 1.0], 2, 0.62),          # class gap 2: tiles not adjacent
        (0, [0.0, 0.0, 640.0, 512.0], 1, 0.58), (1, [0.0, 0.0, 640.0, 512.0], 2, 0.57),          # full-frame boxes: both flags
    ]
    for infos in images:
        for m, box, cls, score in extra:
            p = [(1.0 - score) / 3.0] * 3
            p[cls] = score
            infos[m]["bbox"].append(box); infos[m]["score"].append(score); infos[m]["class"].append(cls)
            infos[m]["prob"].append(p); infos[m]["vars"].append([1.5])
            if "class_logits" in infos[m]:
                infos[m]["class_logits"].append([0.0] * 4)
    for sm, bm in (("probEn", "v-avg"), ("avg", "s-avg"), ("max", "argmax")):
        got = fusion.late_fusion_batch((sm, bm), images)
        for b, infos in enumerate(images):
            want = O.late_fusion_dispatch((sm, bm), infos)
            assert len(infos[0]["bbox"]) + len(infos[1]["bbox"]) > 32
            pc.assert_same_detections(got[b], want, TOL, "border contact %s/%s img %d" % (sm, bm, b))


@pytest.mark.parametrize("per_model", [30, 100])
def test_block_kernel_binary_form(per_model):
    """K = 1 (KAIST: one class, rows [p, 1 - p]) through fuse_mid_kernel<1> at 60 / 200 detections per image - the kaist32
    workload of bench.py (100 detections per model)."""
    dets = synth.synth_model_detections(2, 2, seed=500 + per_model, K=1, force_count=per_model)
    images = [[synth.image_info(d, i) for d in dets] for i in range(2)]
    for sm, bm in (("probEn", "v-avg"), ("avg", "avg"), ("max", "argmax")):
        got = fusion.late_fusion_batch((sm, bm), images, K=1)
        for b, infos in enumerate(images):
            pc.assert_same_detections(got[b], O.late_fusion_dispatch((sm, bm), infos), TOL, "K1 n%d %s/%s img %d" % (per_model, sm, bm, b))


def test_kaist_binary_form():
    """K = 1: rows [p, 1-p]; checked against the oracle's K-generic restatement."""
    rng = np.random.default_rng(9)
    images = []
    for _ in range(64):
        infos = []
        g = rng.integers(1, 6)
        xy = rng.uniform(0, 500, size=(g, 2)); wh = rng.uniform(20, 100, size=(g, 2))
        for m in range(2):
            bx = np.concatenate([xy, xy + wh], 1) + rng.normal(0, 2, size=(g, 4))
            p = rng.uniform(0.5, 0.99, size=(g, 1)).astype(np.float32)
            infos.append({"bbox": bx.astype(np.float32).astype(np.float64).tolist(), "score": p[:, 0].astype(np.float64).tolist(),
                          "class": [0] * g, "prob": p.astype(np.float64).tolist(),
                          "vars": rng.uniform(.5, 2, size=(g, 1)).astype(np.float32).astype(np.float64).tolist()})
        images.append(infos)
    got = fusion.late_fusion_batch(("probEn", "v-avg"), images, K=1)
    for b, infos in enumerate(images):
        pc.assert_same_detections(got[b], O.late_fusion_dispatch(("probEn", "v-avg"), infos), TOL, "kaist %d" % b)


def test_drop_in_fusion_signature():
    a = {"bbox": [[10, 10, 50, 50]], "score": [.9], "class": [0], "prob": [[.9, .05, .03]], "vars": [[1.0]]}
    b = {"bbox": [[12, 11, 52, 49]], "score": [.8], "class": [0], "prob": [[.8, .1, .05]], "vars": [[3.0]]}
    boxes, scores, classes = fusion.fusion(["probEn", "v-avg"], a, b)
    assert boxes.dtype == torch.float32 and scores.dtype == torch.float32 and classes.dtype == torch.float32
    assert torch.allclose(boxes, torch.tensor([[10.5, 10.25, 50.5, 49.75]]), atol=TOL)
    assert abs(float(scores[0]) - 0.9896907) < TOL and float(classes[0]) == 0.0
    boxes, scores, classes = fusion.fusion(["max", "argmax"], a, b, info_3="")
    assert boxes.tolist() == [[10.0, 10.0, 50.0, 50.0]] and abs(float(scores[0]) - .9) < 1e-7


def test_full_size_properties():
    """Benchmark-scale batch (2^18 images): properties that hold for any input.
    (1) ('max','argmax') output is a subset of the input rows, in descending score order;
    (2) fusing is permutation-consistent: counts never exceed inputs, every image with >=1 detection yields >=1;
    (3) avg/avg of a batch whose second model duplicates the first returns the first model's NMS-clustered boxes."""
    p = synth.synth_packed(1 << 18, num_models=2, seed=4)
    dev = fusion.to_device(p)
    buf = fusion.fuse_packed(dev, ("max", "argmax"))
    counts = buf.out_counts[: p["B"]].cpu().numpy()
    o = p["offsets"]
    n_in = o[2::2] - o[:-2:2]
    assert (counts >= 0).all() and (counts <= n_in).all() and ((counts > 0) == (n_in > 0)).all()
    os_ = buf.out_scores.cpu().numpy()
    ob = buf.out_boxes.cpu().numpy()
    for b in np.random.default_rng(0).integers(0, p["B"], size=2000):
        lo, n = int(o[2 * b]), int(counts[b])
        s = os_[lo:lo + n]
        if o[2 * b + 1] > o[2 * b] and o[2 * b + 2] > o[2 * b + 1]:  # both models live (else pass-through order)
            assert (np.diff(s) <= 0).all()
        rows = {tuple(r) for r in p["boxes"][lo:lo + int(n_in[b])].tolist()}
        assert all(tuple(r) in rows for r in ob[lo:lo + n].tolist())
    buf2 = fusion.fuse_packed(dev, ("probEn", "v-avg"))
    c2 = buf2.out_counts[: p["B"]].cpu().numpy()
    assert (c2 <= n_in).all() and ((c2 > 0) == (n_in > 0)).all()
    # checksum-of-checksums against the oracle on a 512-image slice
    images = pc.packed_to_images({**p, "B": 512, "offsets": o[: 512 * 2 + 1]})
    s2 = buf2.out_scores.cpu().numpy(); b2 = buf2.out_boxes.cpu().numpy(); k2 = buf2.out_classes.cpu().numpy()
    for b, infos in enumerate(images):
        n = int(c2[b]); lo = int(o[2 * b])
        got = None if n == 0 else (b2[lo:lo + n], s2[lo:lo + n], k2[lo:lo + n])
        pc.assert_same_detections(got, O.late_fusion_dispatch(("probEn", "v-avg"), infos), TOL, "slice %d" % b)
