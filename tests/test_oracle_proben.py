"""Pins oracle/proben_oracle.py against (a) the committed golden vectors produced by the unmodified
reference, (b) the live reference when /root/reference is present, (c) the known answers of SURVEY §8c."""
import numpy as np
import pytest

import proben_cases as pc
from oracle import proben_oracle as O


@pytest.mark.parametrize("name", pc.SETS)
def test_oracle_matches_golden_bit_exact(golden_proben, name):
    packed = pc.golden_inputs(golden_proben, name)
    images = pc.packed_to_images(packed)
    for sm in pc.SCORES:
        for bm in pc.BOXES:
            want = pc.golden_outputs(golden_proben, name, sm, bm)
            for b, infos in enumerate(images):
                got = O.late_fusion_dispatch((sm, bm), infos)
                pc.assert_same_detections(got, want[b], 0.0, "%s %s/%s img %d" % (name, sm, bm, b), exact_boxes=True)


def test_known_answers():
    s, c = O.probEn_multiclass(np.array([[.7, .1, .1], [.6, .2, .1]]))
    assert abs(s - 0.9130434782608695) < 1e-15 and c == 0
    a = {"bbox": [[10, 10, 50, 50]], "score": [.9], "class": [0], "prob": [[.9, .05, .03]], "vars": [[1.0]]}
    b = {"bbox": [[12, 11, 52, 49]], "score": [.8], "class": [0], "prob": [[.8, .1, .05]], "vars": [[3.0]]}
    bx, sc, cl = O.fusion(["probEn", "v-avg"], a, b)
    assert np.allclose(bx, [[10.5, 10.25, 50.5, 49.75]]) and abs(sc[0] - 0.9897) < 1e-4 and cl[0] == 0
    bx, sc, cl = O.fusion(["max", "argmax"], a, b)
    assert np.array_equal(bx, np.float32([[10, 10, 50, 50]])) and sc[0] == np.float32(.9)
    assert list(O.descending_order(np.array([.9, .8, .9, .8, .9]))) == [4, 2, 0, 3, 1]
    assert abs(O.probEn_binary([.9, .8]) - (.72 / (.72 + .02))) < 1e-12


def test_batched_nms_matches_torchvision():
    import torch
    from torchvision.ops import boxes as box_ops
    rng = np.random.default_rng(5)
    for n in (1, 7, 64, 300):
        xy = rng.uniform(0, 500, size=(n, 2))
        wh = rng.uniform(5, 150, size=(n, 2))
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        scores = rng.random(n).astype(np.float32)
        scores[: n // 3] = scores[0]  # ties
        idxs = rng.integers(0, 3, size=n)
        want = box_ops.batched_nms(torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(idxs), 0.5).numpy()
        got = O.batched_nms_f32(boxes, scores, idxs, 0.5)
        assert np.array_equal(got, want)


@pytest.mark.needs_reference
def test_oracle_matches_live_reference():
    import torch
    import ref_loader
    from probenb200 import synth
    ref = ref_loader.load_reference_proben()
    dets = synth.synth_model_detections(25, 3, seed=77)
    for sm in pc.SCORES:
        for bm in pc.BOXES:
            for i in range(25):
                infos = [synth.image_info(d, i) for d in dets]
                if any(len(x["bbox"]) == 0 for x in infos):
                    continue
                rb, rs, rc = ref.fusion([sm, bm], *infos)
                rb = rb.numpy() if isinstance(rb, torch.Tensor) else np.asarray([np.asarray(x) for x in rb])
                ob, os_, oc = O.fusion([sm, bm], *infos)
                assert np.array_equal(rb.reshape(-1, 4), ob) and np.array_equal(rs.numpy(), os_, equal_nan=True)
                assert np.array_equal(rc.numpy(), oc)
