"""The resize oracle (oracle/resize_oracle.py) against the installed Pillow - the library the reference's
DefaultPredictor delegates 3-channel uint8 frames to (data/transforms/transform.py:92-96)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resize_oracle as R  # noqa: E402

Image = pytest.importorskip("PIL.Image")

CASES = [((64, 80), (100, 125)), ((512, 640), (800, 1000)), ((37, 53), (91, 60)), ((120, 90), (45, 77)),
         ((200, 300), (67, 100)), ((33, 33), (33, 70)), ((50, 64), (13, 16))]


@pytest.mark.parametrize("src,dst", CASES)
def test_restatement_equals_pillow(src, dst):
    rng = np.random.default_rng(src[0] * 1000 + dst[1])
    img = rng.integers(0, 256, (src[0], src[1], 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((dst[1], dst[0]), Image.BILINEAR))
    got = R.pil_bilinear_resize_u8(img, dst[0], dst[1])
    assert np.array_equal(got, want)


def test_extreme_values_and_gradients():
    img = np.zeros((40, 40, 3), np.uint8)
    img[::2] = 255
    img[:, ::3, 1] = 128
    want = np.asarray(Image.fromarray(img).resize((63, 63), Image.BILINEAR))
    assert np.array_equal(R.pil_bilinear_resize_u8(img, 63, 63), want)


def test_resize_shortest_edge_rule():
    assert R.resize_shortest_edge_shape(512, 640) == (800, 1000)
    assert R.resize_shortest_edge_shape(480, 1920) == (333, 1333)
    assert R.resize_shortest_edge_shape(1000, 600) == (1333, 800)


cv2 = pytest.importorskip("cv2")

CV_CASES = [((1600, 1800), (512, 640)), ((100, 130), (64, 80)), ((64, 80), (100, 125)), ((513, 777), (512, 640)),
            ((512, 640), (800, 1000)), ((37, 41), (80, 90)), ((90, 70), (31, 33)), ((48, 48), (48, 48))]


@pytest.mark.parametrize("src,dst", CV_CASES)
@pytest.mark.parametrize("C", [1, 3, 4])
def test_cv2_restatement_equals_opencv(src, dst, C):
    rng = np.random.default_rng(src[1] * 7 + dst[0] + C)
    img = rng.integers(0, 256, (src[0], src[1], C), dtype=np.uint8)
    want = cv2.resize(img, (dst[1], dst[0])).reshape(dst[0], dst[1], C)
    assert np.array_equal(R.cv2_linear_resize_u8(img, dst[0], dst[1]), want)


def test_third_positional_argument_of_cv2_resize_is_not_the_interpolation():
    """The reference passes cv2.INTER_CUBIC as the third positional argument (demo_FLIR_save_predictions.py:108): that
    slot is `dst`, so the result is the default INTER_LINEAR resize."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (90, 120, 3), dtype=np.uint8)
    try:
        as_reference = cv2.resize(img, (64, 48), cv2.INTER_CUBIC)
    except (cv2.error, TypeError):
        pytest.skip("this OpenCV build rejects an int in the dst slot")
    assert np.array_equal(as_reference, cv2.resize(img, (64, 48)))
