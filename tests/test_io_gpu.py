"""Input side on the GPU: nvJPEG decode against cv2.imdecode (the reference's cv2.imread,
demo_FLIR_save_predictions.py:100-117), the cv2-exact uint8 resize and the 3-/4-/6-channel assembly against the
oracle (oracle/resize_oracle.py, pinned to OpenCV) and against OpenCV itself."""
import numpy as np
import pytest
import torch

from oracle import resize_oracle as R
from probenb200 import io as pio

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu


def _smooth_image(h, w, c, seed):
    """Photo-like content (JPEG decoders differ most on noise; the datasets are natural images)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, c), np.float32)
    for ch in range(c):
        for _ in range(6):
            fx, fy, ph = rng.uniform(0.005, 0.08), rng.uniform(0.005, 0.08), rng.uniform(0, 6.28)
            img[:, :, ch] += rng.uniform(10, 40) * np.sin(fx * xx + fy * yy + ph)
        img[:, :, ch] += 128 + rng.normal(0, 3, (h, w))
    cv2.rectangle(img, (w // 4, h // 4), (w // 2, h // 2), (200, 60, 30)[:c] if c == 3 else 220, -1)
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("src,dst,C", [((100, 130), (64, 80), 3), ((1600, 1800), (512, 640), 3), ((64, 80), (100, 125), 4),
                                       ((90, 70), (31, 33), 1), ((48, 48), (48, 48), 3)])
def test_resize_u8_is_opencv_exact(src, dst, C):
    rng = np.random.default_rng(src[0] + dst[1] + C)
    img = rng.integers(0, 256, (2, src[0], src[1], C), dtype=np.uint8)
    got = pio.resize_u8(torch.from_numpy(img).cuda(), dst).cpu().numpy()
    for b in range(2):
        assert np.array_equal(got[b], R.cv2_linear_resize_u8(img[b], dst[0], dst[1]))
        assert np.array_equal(got[b], cv2.resize(img[b], (dst[1], dst[0])).reshape(dst[0], dst[1], C))


@pytest.mark.parametrize("method,C", [("thermal_only", 3), ("rgb_only", 3), ("early_fusion", 4), ("middle_fusion", 6)])
def test_assemble_input_matches_reference_recipe(method, C):
    rng = np.random.default_rng(C)
    rgb = rng.integers(0, 256, (2, 200, 260, 3), dtype=np.uint8)
    th = np.repeat(rng.integers(0, 256, (2, 128, 160, 1), dtype=np.uint8), 3, axis=3)
    got = pio.assemble_input(method, torch.from_numpy(rgb).cuda(), torch.from_numpy(th).cuda()).cpu().numpy()
    assert got.shape == (2, 128, 160, C)
    for b in range(2):
        assert np.array_equal(got[b], R.assemble_input(method, rgb[b], th[b]))
    if method == "early_fusion":  # the reference's own lines (:104-111) with cv2
        want = np.zeros((128, 160, 4))
        want[:, :, 0:3] = cv2.resize(rgb[0], (160, 128))
        want[:, :, -1] = th[0][:, :, 0]
        assert np.array_equal(got[0].astype(np.float64), want)


def test_nvjpeg_decode_matches_cv2_imdecode():
    dec = pio.JpegDecoder()
    # thermal_8_bit frames are grey-scale JPEGs: cv2.imread replicates the plane into B, G, R
    grey = [_smooth_image(128, 160, 1, s)[:, :, 0] for s in range(3)]
    datas = [cv2.imencode(".jpeg", g, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for g in grey]
    assert dec.image_info(datas[0])[:2] == (128, 160)
    got = dec.decode(datas).cpu().numpy()
    assert got.shape == (3, 128, 160, 3)
    for i, d in enumerate(datas):
        want = cv2.imdecode(np.frombuffer(d, np.uint8), cv2.IMREAD_COLOR)
        diff = np.abs(got[i].astype(int) - want.astype(int))
        assert diff.max() <= 2 and diff.mean() < 0.3, (diff.max(), diff.mean())  # IDCT rounding differs between decoders
    # colour JPEGs without chroma subsampling: only IDCT / colour-conversion rounding
    col = [_smooth_image(96, 112, 3, 10 + s) for s in range(2)]
    datas = [cv2.imencode(".jpg", c, [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                      cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])[1].tobytes() for c in col]
    got = dec.decode(datas).cpu().numpy()
    for i, d in enumerate(datas):
        want = cv2.imdecode(np.frombuffer(d, np.uint8), cv2.IMREAD_COLOR)
        diff = np.abs(got[i].astype(int) - want.astype(int))
        assert diff.max() <= 4 and diff.mean() < 0.6, (diff.max(), diff.mean())
    # 4:2:0 files (what cameras write): decoders also differ in chroma upsampling; stay close on smooth content
    datas = [cv2.imencode(".jpg", c, [cv2.IMWRITE_JPEG_QUALITY, 92])[1].tobytes() for c in col]
    got = dec.decode(datas).cpu().numpy()
    for i, d in enumerate(datas):
        want = cv2.imdecode(np.frombuffer(d, np.uint8), cv2.IMREAD_COLOR)
        diff = np.abs(got[i].astype(int) - want.astype(int))
        assert diff.mean() < 3.0, diff.mean()  # measured 1.6 grey levels on this content (sharp colour edge)
    dec.close()


def test_nvjpeg_rejects_mismatched_sizes_and_garbage():
    dec = pio.JpegDecoder()
    a = cv2.imencode(".jpg", _smooth_image(64, 64, 3, 1))[1].tobytes()
    b = cv2.imencode(".jpg", _smooth_image(64, 80, 3, 2))[1].tobytes()
    with pytest.raises(RuntimeError):
        dec.decode([a, b])
    with pytest.raises(RuntimeError):
        dec.decode([b"not a jpeg at all"])
