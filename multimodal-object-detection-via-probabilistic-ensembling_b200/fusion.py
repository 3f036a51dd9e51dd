"""ProbEn late fusion - host side of ``pe_fuse_batch`` (include/probenb200.h).

Mirrors the reference's ``demo/FLIR/demo_probEn.py`` interface:

  * ``fusion(method, info_1, info_2, info_3='')``  <- demo_probEn.py:189-196 (same argument meaning:
    ``method = [score_fusion, box_fusion]``, ``info_k`` dicts with python lists ``bbox, score, class,
    prob, vars``), returns ``(boxes, scores, classes)``;
  * ``late_fusion_batch`` runs the per-image dispatch of demo_probEn.py:236-267 for a whole prediction
    set in ONE kernel launch;
  * ``fuse_packed`` is the device-resident entry (no host copies) used by the detector pipeline.

All compute happens in the CUDA library; this module only packs/unpacks buffers.
"""
import numpy as np
import torch

from . import _lib

SCORE_MODES = {"probEn": 0, "avg": 1, "max": 2}
BOX_MODES = {"v-avg": 0, "s-avg": 1, "avg": 2, "argmax": 3}


def _method_codes(method):
    try:
        return SCORE_MODES[method[0]], BOX_MODES[method[1]]
    except KeyError:
        raise ValueError("unknown fusion method %r (score: %s, box: %s)" %
                         (list(method), list(SCORE_MODES), list(BOX_MODES)))


def pack_detections(images, K=None):
    """``images``: list over images of lists over models of info dicts -> packed SoA numpy arrays.

    Rows of one image are the models' detections concatenated in model order (prepare_data,
    demo_probEn.py:79-90); ``offsets`` has B*M+1 entries."""
    B = len(images)
    M = len(images[0]) if B else 0
    counts = np.zeros(B * M, np.int64)
    for b, infos in enumerate(images):
        if len(infos) != M:
            raise ValueError("every image needs the same number of models")
        for m, info in enumerate(infos):
            counts[b * M + m] = len(info["bbox"])
    offsets = np.zeros(B * M + 1, np.int32)
    np.cumsum(counts, out=offsets[1:])
    N = int(offsets[-1])
    if K is None:
        K = 3
        for infos in images:
            for info in infos:
                if len(info["bbox"]):
                    K = np.asarray(info["prob"]).reshape(len(info["bbox"]), -1).shape[1]
                    break
    boxes = np.zeros((N, 4), np.float32)
    scores = np.zeros(N, np.float32)
    classes = np.zeros(N, np.int32)
    probs = np.zeros((N, K), np.float32)
    var = np.ones(N, np.float32)
    for b, infos in enumerate(images):
        for m, info in enumerate(infos):
            lo, hi = offsets[b * M + m], offsets[b * M + m + 1]
            if hi > lo:
                boxes[lo:hi] = np.asarray(info["bbox"], np.float64).reshape(-1, 4)
                scores[lo:hi] = np.asarray(info["score"], np.float64)
                classes[lo:hi] = np.asarray(info["class"]).astype(np.int32)
                probs[lo:hi] = np.asarray(info["prob"], np.float64).reshape(hi - lo, K)
                var[lo:hi] = np.asarray(info["vars"], np.float64).reshape(hi - lo)
    return {"boxes": boxes, "scores": scores, "classes": classes, "probs": probs, "vars": var,
            "offsets": offsets, "B": B, "M": M, "K": K}


def to_device(packed, device="cuda", pinned=False):
    out = dict(packed)
    for k in ("boxes", "scores", "classes", "probs", "vars", "offsets"):
        t = torch.from_numpy(np.ascontiguousarray(packed[k]))
        if pinned:
            t = t.pin_memory()
        out[k] = t.to(device, non_blocking=pinned)
    return out


class FuseBuffers:
    """Reusable output + workspace buffers for ``fuse_packed`` (sized for N rows / B images)."""

    def __init__(self, N, B, device):
        lib = _lib.load()
        self.N, self.B = N, B
        self.out_boxes = torch.empty((max(N, 1), 4), dtype=torch.float32, device=device)
        self.out_scores = torch.empty(max(N, 1), dtype=torch.float32, device=device)
        self.out_classes = torch.empty(max(N, 1), dtype=torch.int32, device=device)
        self.out_counts = torch.empty(max(B, 1), dtype=torch.int32, device=device)
        self.ws_bytes = int(lib.pe_fuse_workspace_bytes(B))
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)


def fuse_packed(dev, method, iou_thr=0.5, img_w=640.0, img_h=512.0, buffers=None):
    """Launch ``pe_fuse_batch`` on device-resident packed detections (dict from ``to_device``).
    Asynchronous on the current stream.  Returns the ``FuseBuffers`` holding the outputs."""
    lib = _lib.load()
    sm, bm = _method_codes(method)
    B, M, K = dev["B"], dev["M"], dev["K"]
    N = dev["boxes"].shape[0]
    _lib.require_cuda(dev["boxes"], dev["scores"], dev["classes"], dev["probs"], dev["vars"], dev["offsets"])
    if dev["classes"].dtype != torch.int32 or dev["offsets"].dtype != torch.int32:
        raise RuntimeError("probenb200: classes/offsets must be int32")
    if buffers is None or buffers.N < N or buffers.B < B:
        buffers = FuseBuffers(N, B, dev["boxes"].device)
    if B == 0:
        return buffers
    # a zero-row batch still needs valid pointers
    p = lambda t: _lib.ptr(t if t.numel() else buffers.out_boxes)
    with torch.cuda.device(dev["boxes"].device):  # the library launches on the current device
        st = lib.pe_fuse_batch(p(dev["boxes"]), p(dev["scores"]), p(dev["classes"]), p(dev["probs"]), p(dev["vars"]),
                               _lib.ptr(dev["offsets"]), B, M, K, float(iou_thr), sm, bm, float(img_w), float(img_h),
                               _lib.ptr(buffers.out_boxes), _lib.ptr(buffers.out_scores), _lib.ptr(buffers.out_classes),
                               _lib.ptr(buffers.out_counts), _lib.ptr(buffers.workspace), buffers.ws_bytes,
                               _lib.current_stream_ptr(dev["boxes"].device))
    _lib.check(st, "pe_fuse_batch")
    return buffers


def unpack_results(packed, buffers):
    """Device -> host; returns list over images of None (image skipped by the reference) or
    (boxes float32 (n,4), scores float32 (n,), classes float32 (n,)) numpy arrays."""
    B, M = packed["B"], packed["M"]
    counts = buffers.out_counts[:B].cpu().numpy()
    if (counts < 0).any():
        raise RuntimeError("probenb200: an image exceeds %d detections" % _lib.load().pe_fuse_max_dets_per_image())
    N = packed["boxes"].shape[0]
    ob = buffers.out_boxes[:N].cpu().numpy()
    os_ = buffers.out_scores[:N].cpu().numpy()
    oc = buffers.out_classes[:N].cpu().numpy()
    offs = packed["offsets"]
    offs = offs.cpu().numpy() if isinstance(offs, torch.Tensor) else offs
    out = []
    for b in range(B):
        n = int(counts[b])
        if n == 0:
            out.append(None)
            continue
        lo = int(offs[b * M])
        out.append((ob[lo:lo + n].copy(), os_[lo:lo + n].copy(), oc[lo:lo + n].astype(np.float32)))
    return out


def late_fusion_batch(method, images, iou_thr=0.5, img_w=640.0, img_h=512.0, device="cuda", K=None):
    """Whole-set equivalent of the loop body demo_probEn.py:204-267: ``images[b]`` is the list of the M
    models' info dicts for image b.  One H2D, one kernel launch, one D2H."""
    packed = pack_detections(images, K=K)
    dev = to_device(packed, device)
    buf = fuse_packed(dev, method, iou_thr, img_w, img_h)
    return unpack_results(packed, buf)


def fusion(method, info_1, info_2, info_3="", iou_thr=0.5, img_w=640.0, img_h=512.0, device="cuda"):
    """Drop-in for the reference ``fusion`` (demo_probEn.py:189-196).

    Returns ``(boxes, scores, classes)`` as CPU torch tensors: boxes float32 (n,4) (the reference returns
    float64 rows that ``Boxes`` immediately casts to float32, structures/boxes.py:145), scores float32,
    classes float32 (demo_probEn.py:182-183).  Like the reference it assumes it is only called when at
    least two of the models have detections."""
    infos = [info_1, info_2] + ([info_3] if info_3 else [])
    if sum(len(i["bbox"]) > 0 for i in infos) < 2:
        # the reference's np.concatenate of a (0,) and an (n,4) array raises here
        raise ValueError("fusion() needs detections from at least two models (see late_fusion_batch)")
    res = late_fusion_batch(method, [infos], iou_thr, img_w, img_h, device)[0]
    b, s, c = res
    return torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(c)
