"""TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the
product package).

CPU restatement of the frame resize of the reference's predictor:
``DefaultPredictor.__call__`` (detectron2/engine/defaults.py:186-190) -> ``ResizeShortestEdge.get_transform``
(data/transforms/transform_gen.py:192-213) -> ``ResizeTransform.apply_image`` (data/transforms/transform.py:81-99),
which for 3-channel uint8 frames calls ``PIL.Image.resize(..., BILINEAR)``.

Third-party arithmetic: Pillow (reference environment pins 9.2.0, probEn.yml:146; installed here: 12.2).  Its 8-bit
resampler (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc,
ImagingResampleVertical_8bpc; unchanged between those versions) is restated below in numpy: triangle weights in
double over a support of max(1, scale), normalised, rounded to 22-bit fixed point; rows are filtered first into a
rounded uint8 image, then columns.  Parity status: PINNED - tests/test_oracle_resize.py checks this restatement bit
for bit against the installed Pillow on upscales, downscales and odd sizes.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0), box = whole image.
    Returns (lo[out], n[out], k[out, ksize] int64)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    lo = np.zeros(out_size, np.int64)
    cnt = np.zeros(out_size, np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize)
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w[:xmax] = w[:xmax] / ww
        for x in range(xmax):
            kk[xx, x] = int(0.5 + w[x] * (1 << PRECISION_BITS))
        lo[xx], cnt[xx] = xmin, xmax
    return lo, cnt, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255)


def pil_bilinear_resize_u8(img, out_h, out_w):
    """img: (H, W, C) uint8 -> (out_h, out_w, C) uint8, identical to ``Image.fromarray(img).resize((out_w, out_h), BILINEAR)``."""
    img = np.asarray(img, np.uint8)
    H, W, C = img.shape
    src = img.astype(np.int64)
    if out_w != W:
        lo, cnt, kk = pil_bilinear_coeffs(W, out_w)
        acc = np.full((H, out_w, C), 1 << (PRECISION_BITS - 1), np.int64)
        for t in range(kk.shape[1]):
            idx = np.minimum(lo + t, W - 1)
            acc += src[:, idx, :] * np.where(t < cnt, kk[:, t], 0)[None, :, None]
        src = _clip8(acc)
    if out_h != H:
        lo, cnt, kk = pil_bilinear_coeffs(H, out_h)
        acc = np.full((out_h, src.shape[1], C), 1 << (PRECISION_BITS - 1), np.int64)
        for t in range(kk.shape[1]):
            idx = np.minimum(lo + t, H - 1)
            acc += src[idx, :, :] * np.where(t < cnt, kk[:, t], 0)[:, None, None]
        src = _clip8(acc)
    return src.astype(np.uint8)


def resize_shortest_edge_shape(h, w, short=800, max_size=1333):
    """ResizeShortestEdge.get_transform (transform_gen.py:192-213): scale so the short side is `short`, cap the long
    side at `max_size`, round half up."""
    scale = short * 1.0 / min(h, w)
    if h < w:
        newh, neww = short, scale * w
    else:
        newh, neww = scale * h, short
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    return int(newh + 0.5), int(neww + 0.5)


# ---- cv2.resize(uint8, INTER_LINEAR) ---------------------------------------------------------------------------------
# demo/FLIR/demo_FLIR_save_predictions.py:108,117 resizes the RGB frame to the thermal frame's size with
# ``cv2.resize(rgb_img, (w, h), cv2.INTER_CUBIC)`` - the third positional argument of cv2.resize is ``dst``, so the
# interpolation stays at its default, INTER_LINEAR.  Third-party arithmetic: OpenCV (reference pins 4.6.0, probEn.yml:140;
# installed 4.13), modules/imgproc/src/resize.cpp, 8-bit linear path: 11-bit fixed-point weights from float32
# fractions, horizontal pass in int, vertical pass with two truncating shifts.  Parity status: PINNED -
# tests/test_oracle_resize.py checks this restatement bit for bit against the installed cv2 for down- and upscales
# of 1-, 3- and 4-channel images.

def _cv_coeffs(ssize, dsize, clamp_fraction):
    scale = 1.0 / (float(dsize) / float(ssize))
    i0 = np.zeros(dsize, np.int64)
    i1 = np.zeros(dsize, np.int64)
    w0 = np.zeros(dsize, np.int64)
    w1 = np.zeros(dsize, np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_fraction:  # columns: the fraction is zeroed at the borders
            if s < 0:
                f, s = np.float32(0), 0
            if s >= ssize - 1:
                f, s = np.float32(0), ssize - 1
        i0[d] = min(max(s, 0), ssize - 1)
        i1[d] = min(max(s + 1, 0), ssize - 1)
        w0[d] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(2048)))
        w1[d] = int(np.rint(f * np.float32(2048)))
    return i0, i1, w0, w1


def cv2_linear_resize_u8(img, out_h, out_w):
    """img: (H, W, C) uint8 -> (out_h, out_w, C) uint8, identical to ``cv2.resize(img, (out_w, out_h))``."""
    img = np.asarray(img, np.uint8)
    if img.ndim == 2:
        img = img[:, :, None]
    H, W, _ = img.shape
    x0, x1, a0, a1 = _cv_coeffs(W, out_w, True)
    y0, y1, b0, b1 = _cv_coeffs(H, out_h, False)
    src = img.astype(np.int64)
    rows = src[:, x0, :] * a0[None, :, None] + src[:, x1, :] * a1[None, :, None]
    out = (((b0[:, None, None] * (rows[y0] >> 4)) >> 16) + ((b1[:, None, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def assemble_input(method, rgb, thermal):
    """demo_FLIR_save_predictions.py:100-121: the array handed to the predictor (uint8 here; the reference stores the
    4-/6-channel cases in float64 arrays holding the same integers)."""
    if method == "thermal_only":
        return thermal
    rgb_r = cv2_linear_resize_u8(rgb, thermal.shape[0], thermal.shape[1])
    if method in ("rgb_only", "RGB"):
        return rgb_r if method == "rgb_only" else rgb
    if method == "early_fusion":
        return np.concatenate([rgb_r, thermal[:, :, :1]], axis=2)
    if method == "middle_fusion":
        return np.concatenate([rgb_r, thermal], axis=2)
    raise ValueError(method)
