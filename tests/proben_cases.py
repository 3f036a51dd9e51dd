"""Helpers shared by the ProbEn tests: unpack the golden npz back into per-image info dicts."""
import numpy as np

SCORES = ("probEn", "avg", "max")
BOXES = ("v-avg", "s-avg", "avg", "argmax")
SETS = ("real3", "real2", "stress120", "stress300", "adv2", "adv3")


def golden_inputs(g, name):
    packed = {k: g["%s/in/%s" % (name, k)] for k in ("boxes", "scores", "classes", "probs", "vars", "offsets")}
    B, M = (int(v) for v in g["%s/in/BM" % name])
    packed.update(B=B, M=M, K=packed["probs"].shape[1])
    return packed


def packed_to_images(p):
    """Inverse of fusion.pack_detections: list over images of lists over models of info dicts."""
    B, M, K = p["B"], p["M"], p["K"]
    o = p["offsets"]
    images = []
    for b in range(B):
        infos = []
        for m in range(M):
            lo, hi = int(o[b * M + m]), int(o[b * M + m + 1])
            infos.append({"img_name": "%d.jpg" % b,
                          "bbox": p["boxes"][lo:hi].astype(np.float64).tolist(),
                          "score": p["scores"][lo:hi].astype(np.float64).tolist(),
                          "class": p["classes"][lo:hi].tolist(),
                          "prob": p["probs"][lo:hi].astype(np.float64).reshape(hi - lo, K).tolist(),
                          "vars": p["vars"][lo:hi].astype(np.float64)[:, None].tolist()})
        images.append(infos)
    return images


def golden_outputs(g, name, sm, bm):
    key = "%s/out/%s/%s" % (name, sm, bm)
    counts = g[key + "/counts"]
    starts = np.concatenate([[0], np.cumsum(counts)])
    res = []
    for b, n in enumerate(counts):
        if n == 0:
            res.append(None)
        else:
            lo = starts[b]
            res.append((g[key + "/boxes"][lo:lo + n], g[key + "/scores"][lo:lo + n], g[key + "/classes"][lo:lo + n]))
    return res


def assert_same_detections(got, want, tol, what, exact_boxes=False):
    """got / want: (boxes, scores, classes) or None.  classes & counts exact, boxes/scores within tol."""
    assert (got is None) == (want is None), what
    if want is None:
        return
    gb, gs, gc = (np.asarray(x) for x in got)
    wb, ws, wc = (np.asarray(x) for x in want)
    assert len(gs) == len(ws), "%s: count %d != %d" % (what, len(gs), len(ws))
    assert np.array_equal(gc.astype(np.float64), wc.astype(np.float64)), "%s: classes %s vs %s" % (what, gc, wc)
    assert np.array_equal(np.isnan(gs), np.isnan(ws)), "%s: NaN pattern" % what
    ok = ~np.isnan(ws)
    if exact_boxes:
        assert np.array_equal(gb, wb), what
        assert np.array_equal(gs[ok], ws[ok]), what
    else:
        assert np.max(np.abs(gs[ok].astype(np.float64) - ws[ok]), initial=0) <= tol, "%s: scores" % what
        assert np.max(np.abs(gb.astype(np.float64).reshape(-1, 4) - wb.reshape(-1, 4)), initial=0) <= tol, "%s: boxes" % what
