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
