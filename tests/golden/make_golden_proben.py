"""Generates tests/golden/proben_golden.npz by running the UNMODIFIED reference ``fusion()`` of
/root/reference/demo/FLIR/demo_probEn.py (loaded by path, see ref_loader.py) on seeded synthetic and
hand-built adversarial detection sets.  Run in the build container:  python tests/golden/make_golden_proben.py

The per-image dispatch (0 / 1 / >=2 non-empty models) is the one of demo_probEn.py:236-267; only the
``fusion`` call itself is reference code (the surrounding loop needs the dataset + evaluator).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_loader  # noqa: E402
from probenb200 import synth  # noqa: E402
from probenb200.fusion import pack_detections  # noqa: E402

SCORES = ("probEn", "avg", "max")
BOXES = ("v-avg", "s-avg", "avg", "argmax")


def info(boxes, probs, var=None, classes=None):
    probs = np.asarray(probs, np.float32).reshape(len(boxes), 3)
    var = np.ones(len(boxes), np.float32) if var is None else np.asarray(var, np.float32)
    cls = probs.argmax(axis=1) if classes is None else np.asarray(classes)
    return {"img_name": "x.jpg", "bbox": np.asarray(boxes, np.float32).astype(np.float64).reshape(-1, 4).tolist(),
            "score": probs.max(axis=1).astype(np.float64).tolist() if len(boxes) else [],
            "class": [int(c) for c in cls], "class_logits": [[0.0] * (probs.shape[1] + 1)] * len(boxes),
            "prob": probs.astype(np.float64).tolist(), "vars": var.astype(np.float64)[:, None].tolist()}


EMPTY = info(np.zeros((0, 4)), np.zeros((0, 3)))


def adversarial_images():
    nf = np.nextafter
    f32 = np.float32
    imgs = []
    # known-answer pair of SURVEY.md §8c
    imgs.append([info([[10, 10, 50, 50]], [[.9, .05, .03]], [1.0]), info([[12, 11, 52, 49]], [[.8, .1, .05]], [3.0])])
    # exact IoU == 0.5 with the +1 convention (not a match), and one ulp either side
    a = [0, 0, 9, 9]
    for y2 in (f32(4), nf(f32(4), f32(9)), nf(f32(4), f32(0))):
        imgs.append([info([a], [[.9, .05, .03]]), info([[0, 0, 9, y2]], [[.8, .1, .05]], [2.0])])
    # tied scores (n <= 16 keeps numpy's order deterministic), overlapping chain a~b, b~c, a!~c
    imgs.append([info([[100, 100, 200, 200], [130, 100, 230, 200], [300, 300, 360, 380]],
                      [[.9, .05, .03], [.8, .1, .05], [.9, .02, .03]], [1.0, 2.0, .5]),
                 info([[160, 100, 260, 200], [302, 301, 361, 379], [100, 101, 199, 202]],
                      [[.8, .1, .05], [.9, .04, .03], [.9, .01, .01]], [1.5, .7, 3.0])])
    # sum(p) == 1 exactly (bg = 0 -> -inf log) and sum(p) > 1 (bg < 0 -> NaN posterior, quirk 4)
    imgs.append([info([[10, 10, 60, 60]], [[.5, .25, .25]]), info([[11, 11, 61, 61]], [[.75, .125, .125]])])
    imgs.append([info([[10, 10, 60, 60]], [[f32(.6), f32(.3), f32(.1)]]),
                 info([[11, 11, 61, 61]], [[.8, .1, .05]])])
    # same geometry, different classes: no match through the class-offset trick
    imgs.append([info([[50, 50, 150, 150]], [[.9, .05, .03]]), info([[50, 50, 150, 150]], [[.05, .9, .03]])])
    # boxes at the right/bottom border vs left/top border of the next class tile (touch through +1)
    imgs.append([info([[600, 480, 640, 512]], [[.9, .05, .03]]), info([[0, 0, 40, 30]], [[.05, .9, .03]])])
    # posterior flips to background (class 3) when both models are unsure
    imgs.append([info([[10, 10, 60, 60]], [[.4, .05, .05]]), info([[11, 11, 61, 61]], [[.42, .04, .05]])])
    # three models, one / two / all empty
    one = info([[20, 20, 80, 90], [200, 200, 260, 280]], [[.9, .05, .03], [.1, .8, .05]], [1.0, 2.0])
    two = info([[22, 21, 83, 88]], [[.7, .2, .05]], [.5])
    imgs.append([one, EMPTY, two])
    imgs.append([EMPTY, two, one])
    imgs.append([EMPTY, EMPTY, one])
    imgs.append([EMPTY, EMPTY, EMPTY])
    imgs.append([one, two, info([[19, 22, 79, 91], [400, 100, 470, 190]], [[.6, .3, .05], [.02, .03, .9]], [.25, 4.0])])
    return imgs


def run_reference(ref, method, infos):
    live = [i for i in infos if len(i["bbox"]) > 0]
    if not live:
        return None
    if len(live) == 1:
        i = live[0]
        return (np.array(i["bbox"], np.float64), torch.Tensor(i["score"]).numpy(), torch.Tensor(i["class"]).numpy())
    b, s, c = ref.fusion(list(method), *live)
    b = b.numpy().astype(np.float64) if isinstance(b, torch.Tensor) else np.asarray([np.asarray(x) for x in b], np.float64)
    return b.reshape(-1, 4), s.numpy(), c.numpy()


def build_sets():
    sets = {}
    d3 = synth.synth_model_detections(40, 3, seed=11)
    sets["real3"] = [[synth.image_info(d, i) for d in d3] for i in range(40)]
    d2 = synth.synth_model_detections(40, 2, seed=12)
    sets["real2"] = [[synth.image_info(d, i) for d in d2] for i in range(40)]
    s40 = synth.synth_model_detections(3, 3, seed=13, force_count=40)     # N = 120 -> block path
    sets["stress120"] = [[synth.image_info(d, i) for d in s40] for i in range(3)]
    s100 = synth.synth_model_detections(2, 3, seed=14, force_count=100)   # N = 300 = 3 x DETECTIONS_PER_IMAGE
    sets["stress300"] = [[synth.image_info(d, i) for d in s100] for i in range(2)]
    adv = adversarial_images()
    sets["adv2"] = [im for im in adv if len(im) == 2]
    sets["adv3"] = [im for im in adv if len(im) == 3]
    return sets


def main():
    ref = ref_loader.load_reference_proben()
    out = {}
    for name, images in build_sets().items():
        packed = pack_detections(images, K=3)
        for k in ("boxes", "scores", "classes", "probs", "vars", "offsets"):
            out["%s/in/%s" % (name, k)] = packed[k]
        out["%s/in/BM" % name] = np.asarray([packed["B"], packed["M"]], np.int32)
        for sm in SCORES:
            for bm in BOXES:
                counts, bb, ss, cc = [], [], [], []
                for infos in images:
                    r = run_reference(ref, (sm, bm), infos)
                    if r is None:
                        counts.append(0)
                        continue
                    counts.append(len(r[1]))
                    bb.append(r[0]); ss.append(r[1]); cc.append(r[2])
                key = "%s/out/%s/%s" % (name, sm, bm)
                out[key + "/counts"] = np.asarray(counts, np.int32)
                out[key + "/boxes"] = np.concatenate(bb).astype(np.float64) if bb else np.zeros((0, 4))
                out[key + "/scores"] = np.concatenate(ss).astype(np.float32) if ss else np.zeros(0, np.float32)
                out[key + "/classes"] = np.concatenate(cc).astype(np.float32) if cc else np.zeros(0, np.float32)
    path = os.path.join(HERE, "proben_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
