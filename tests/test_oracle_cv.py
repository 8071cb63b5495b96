"""The OpenCV 2.4.9 arithmetic restated in oracle/cvmath.h, checked against the only OpenCV in this
image (cv2 4.13, AVX2/FMA build, different kernel normalisation) -- a plausibility bound, the
reference pins nothing here (parity unpinned)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import synth


@pytest.mark.parametrize("sigma", [1.2262737, 1.5198685, 2.4525473])
def test_blur_close_to_cv2(oracle, sigma):
    img = synth.blob_image(300, 200, seed=5)
    size = int(2.0 * 3.0 * sigma + 1.0); size += (size % 2 == 0)
    ref = cv2.GaussianBlur(img, (size, size), sigma, sigmaY=sigma, borderType=cv2.BORDER_REPLICATE)
    out = oracle.gaussian_blur(img, sigma)
    assert np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-5


def test_resize_half_close_to_cv2(oracle):
    for (w, h) in ((300, 200), (135, 101), (17, 33)):
        img = synth.blob_image(w, h, seed=w)
        ref = cv2.resize(img, (0, 0), fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR)
        out = oracle.resize_half(img)
        assert out.shape == ref.shape
        assert np.max(np.abs(out - ref)) < 1e-3
