"""Pins oracle/nv12_ref.py byte for byte to the installed OpenCV (cv2.cvtColor COLOR_YUV2RGB_NV12)."""
import numpy as np
import pytest

from oracle import nv12_ref
from synth import structured_frames

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("h,w,seed", [(2, 2, 0), (64, 96, 1), (270, 480, 2), (1080, 1920, 3), (722, 1282, 4)])
def test_random_bytes_match_cv2(h, w, seed):
    nv = np.random.default_rng(seed).integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)   # includes out-of-range Y/U/V
    assert np.array_equal(nv12_ref.nv12_to_rgb(nv), cv2.cvtColor(nv, cv2.COLOR_YUV2RGB_NV12))


def test_extremes_saturate_like_cv2():
    vals = np.array([0, 1, 15, 16, 17, 127, 128, 129, 234, 235, 236, 240, 254, 255], np.uint8)
    y, u, v = np.meshgrid(vals, vals, vals, indexing="ij")
    n = y.size                                         # one 2x2 block per (Y, U, V) combination
    nv = np.empty((3, 2 * n), np.uint8)
    nv[0] = nv[1] = np.repeat(y.ravel(), 2)
    nv[2, 0::2], nv[2, 1::2] = u.ravel(), v.ravel()
    assert np.array_equal(nv12_ref.nv12_to_rgb(nv), cv2.cvtColor(nv, cv2.COLOR_YUV2RGB_NV12))


def test_structured_frames_round_trip_is_close():
    f = structured_frames(2, 360, 640, seed=5)
    for img in f:
        nv = nv12_ref.rgb_to_nv12(img)
        assert nv.shape == (540, 640)
        rgb = nv12_ref.nv12_to_rgb(nv)
        assert np.array_equal(rgb, cv2.cvtColor(nv, cv2.COLOR_YUV2RGB_NV12))
        assert np.abs(rgb.astype(int) - img.astype(int)).mean() < 6.0      # chroma subsampling loss only


def test_odd_sizes_are_rejected():
    with pytest.raises(ValueError):
        nv12_ref.nv12_to_rgb(np.zeros((5, 4), np.uint8))
