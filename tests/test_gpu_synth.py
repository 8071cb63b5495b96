"""View synthesis on the GPU through the C ABI: GenerateSynthImageCorr and one full (detector, view)
pass of SynthDetectDescribeKeypoints for tilted / rotated / zoomed views, bit-exact against the CPU oracle."""
import numpy as np
import pytest

import mods_b200 as mb
import synth

pytestmark = pytest.mark.gpu
VIEWS = [(2, 0.0, 1.0), (2, 0.7, 1.0), (4, 2.2, 1.0), (1, 0.0, 0.5), (6, 1.2, 0.25), (-2, 0.5, 1.0), (9, 3.0, 1.0), (8, 1.5707, 1.0)]


@pytest.mark.parametrize("view", VIEWS)
def test_synth_view_bit_exact(ctx, oracle, view):
    img = synth.blob_image(320, 240, seed=5)
    g, Hg, ig = ctx.synth_view(img, *view)
    o, Ho, io = oracle.synth_view(img, *view)
    assert g.shape == o.shape and np.array_equal(g, o) and np.array_equal(Hg, Ho) and ig == io


def test_synth_view_identity_and_odd_sizes(ctx, oracle):
    img = synth.blob_image(135, 101, seed=9)
    g, H, ident = ctx.synth_view(img, 1.0, 0.0, 1.0)
    assert ident and np.array_equal(g, img) and np.array_equal(H, np.eye(3))
    for view in ((3, 0.9, 1.0), (1, 0.0, 0.33)):
        g, Hg, _ = ctx.synth_view(img, *view)
        o, Ho, _ = oracle.synth_view(img, *view)
        assert np.array_equal(g, o) and np.array_equal(Hg, Ho)


@pytest.mark.parametrize("detector", ["hess", "mser"])
@pytest.mark.parametrize("view", [(2, 0.6, 1.0), (4, 2.0, 1.0), (1, 0.0, 0.5)])
def test_synth_view_pipeline_bit_exact(ctx, oracle, detector, view):
    """synthesise -> detect on the view -> orientation -> reprojection to the original frame -> RootSIFT on the view."""
    img = synth.blob_image(640, 480, seed=3)
    det = mb.MserParams.default() if detector == "mser" else mb.HessaffParams.default()
    gd, gr, gu = ctx.detect_describe_synth_view(img, *view, det=det, slot=6)
    od, orp, ou = oracle.view_pipeline_synth(img, *view, detector=3 if detector == "mser" else 0)
    assert len(od) > 20
    assert np.array_equal(gd, od) and np.array_equal(gr, orp) and np.array_equal(gu.astype(np.float32), ou)


def test_synth_view_full_size_properties(ctx):
    """4096x3072, tilt 2: view size as the reference computes it, deterministic, border value where the rotation leaves the image."""
    img = synth.blob_image(4096, 3072, seed=1, n_blobs=20000)
    a, H, _ = ctx.synth_view(img, 2.0, 0.5, 1.0)
    b, _, _ = ctx.synth_view(img, 2.0, 0.5, 1.0)
    c, s = np.cos(0.5), np.sin(0.5)
    assert a.shape == (int(np.floor(0.5 + s * 4096 + c * 3072)), int(np.floor((0.5 + c * 4096 + s * 3072) / 2.0)))
    assert np.array_equal(a, b) and a[0, 0] == 128.0 and a[-1, -1] == 128.0 and 0 <= a.min() and a.max() <= 255.0
