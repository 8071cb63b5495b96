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


@pytest.mark.parametrize("tilt,phi", [(2, 0.0), (2, 0.7), (4, 1.2), (8, 0.3)])
def test_synth_view_close_to_cv2(oracle, tilt, phi):
    """GenerateSynthImageCorr rebuilt from cv2 4.13's warpAffine / GaussianBlur (synth-detection.cpp:344-427): the oracle's
    fixed-point warpAffine (1/32 px coordinates, float bilinear table) and reflect-101 blur agree to float rounding."""
    img = synth.blob_image(320, 240, seed=5)
    h, w = img.shape
    c, s = np.cos(phi), np.sin(phi)
    wr, hr = int(np.floor(0.5 + c * w + s * h)), int(np.floor(0.5 + s * w + c * h))
    M = np.array([[c, s, 0], [-s, c, np.floor(0.5 + s * w)]])
    rot = cv2.warpAffine(img, M, (wr, hr), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(128, 128, 128))
    sx, sy = 0.5 * tilt / 2, 0.25
    kx = int(np.floor(6 * sx + 1)); kx += (kx % 2 == 0); kx = max(kx, 3)
    ky = int(np.floor(6 * sy + 1)); ky += (ky % 2 == 0); ky = max(ky, 3)
    rot = cv2.GaussianBlur(rot, (kx, ky), sx, sigmaY=sy)
    wn, hn = int(np.floor((0.5 + c * w + s * h) / tilt)), int(np.floor(0.5 + s * w + c * h))
    ref = cv2.warpAffine(rot, np.array([[1.0 / tilt, 0, 0], [0, 1.0, 0]]), (wn, hn), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                         borderValue=(128, 128, 128))
    out, H, ident = oracle.synth_view(img, tilt, phi, 1.0)
    assert out.shape == ref.shape and not ident
    assert np.max(np.abs(out - ref)) < 2e-4 and np.mean(out == ref) > 0.5
