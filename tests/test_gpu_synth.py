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


def test_mods_pair_with_view_tiers_equals_per_view_composition(ctx):
    """One mods.cpp step with view tiers (HessianAffine: identity + two tilted views, MSER: identity + a zoomed-out view) through the
    C++ host mirror (SynthDetectDescribeKeypoints appends the views of a detector in view order, MatchImgReps per detector, joint
    verification) == the same step composed from per-view C-ABI calls by the view-sharded driver (world size 1)."""
    from mods_b200 import sharding
    A = synth.blob_image(480, 360, seed=21, n_blobs=500)
    B = synth.warp_image(A, synth.gt_homography(480, 360), seed=22)
    hess = [(1.0, 0.0, 1.0, 0.2), (2.0, 0.0, 1.0, 0.2), (2.0, 1.5707963267948966, 1.0, 0.2)]
    mser = [(1.0, 0.0, 1.0, 0.8), (1.0, 0.0, 0.5, 0.8)]
    cfg = mb.PairConfig.default()
    cfg.seed = 77
    cfg.use_mser = 1
    cfg.set_views(hess, mser)
    res, ver = ctx.mods_pair(A, B, cfg, capacity=8192)
    tiers = {"HessianAffine": [(v[2], v[0], v[1], v[3]) for v in hess], "MSER": [(v[2], v[0], v[1], v[3]) for v in mser]}   # rows (zoom, tilt, phi, sigma)
    units, costs = sharding.iters_units(480, 360, tiers)
    compute, match = sharding.gpu_workers(ctx, A, B, cfg)
    groups = sharding.pair_views_sharded(compute, match, units, costs)
    n1 = sum(len(g[0]) for g in groups.values()); n2 = sum(len(g[2]) for g in groups.values())
    assert (res.regions1, res.regions2) == (n1, n2) and res.mser_regions1 == len(groups["MSER"][0]) > 20
    assert res.tentatives == sum(len(g[4]) for g in groups.values()) and res.mser_tentatives == len(groups["MSER"][4])
    frames, keys = sharding.frames_and_keys(groups)
    r2, v2 = ctx.verify(frames, keys, cfg, capacity=8192)
    assert (res.unique_tentatives, res.ransac_inliers, res.verified) == (r2.unique_tentatives, r2.ransac_inliers, r2.verified)
    assert np.array_equal(ver, v2) and res.verified >= 8
    # more views than the identity alone must have added regions
    cfg1 = mb.PairConfig.default(); cfg1.seed = 77; cfg1.use_mser = 1
    res1, _ = ctx.mods_pair(A, B, cfg1)
    assert res.regions1 > res1.regions1 and res.tentatives > res1.tentatives


def test_views_sharded_driver_equals_pair_driver(ctx):
    """mb2_views_sharded_pair at world size 1 (views -> 184-byte device records -> ordered sets -> row-range matching -> verification) gives
    exactly what mb2_mods_pair gives for the same view tiers: counts and the verified list."""
    import mods_b200 as mb
    from synth import blob_image, warp_image, gt_homography
    A = blob_image(640, 480, seed=51, n_blobs=700); B = warp_image(A, gt_homography(640, 480), seed=52)
    cfg = mb.PairConfig.default(); cfg.use_mser = 1; cfg.seed = 5
    hess = [(1.0, 0.0, 1.0, 0.2), (2.0, 0.0, 1.0, 0.2), (4.0, 0.0, 1.0, 0.2), (4.0, np.pi / 2, 1.0, 0.2)]
    mser = [(1.0, 0.0, 1.0, 0.8), (1.0, 0.0, 0.25, 0.8)]
    cfg.set_views(hess, mser)
    r1, v1 = ctx.mods_pair(A, B, cfg, capacity=1 << 15)
    r2, v2, dig, st = ctx.views_sharded_pair(A, B, cfg, capacity=1 << 15)
    f = lambda r: (r.regions1, r.regions2, r.mser_regions1, r.mser_regions2, r.tentatives, r.mser_tentatives, r.unique_tentatives, r.ransac_inliers, r.verified)
    assert f(r1) == f(r2) and r1.verified > 100
    assert np.array_equal(v1, v2) and np.array_equal(np.array(r1.H), np.array(r2.H))
    r3, v3, dig3, _ = ctx.views_sharded_pair(A, B, cfg, capacity=1 << 15)
    assert dig == dig3 and st["units"] == 12


def test_views_sharded_dataset_call_equals_single_calls(ctx):
    """mb2_views_sharded_pairs (pair k verified on a helper thread / context while pair k + 1 is in its views) == mb2_views_sharded_pair per pair:
    counts, H, verified lists and digests, for pairs of different sizes."""
    import mods_b200 as mb
    from synth import blob_image, warp_image, gt_homography
    pairs = []
    for k in range(3):
        A = blob_image(480 + 32 * k, 360 + 16 * k, seed=71 + k, n_blobs=450 + 60 * k)
        pairs.append((A, warp_image(A, gt_homography(A.shape[1], A.shape[0]), seed=81 + k)))
    cfg = mb.PairConfig.default(); cfg.use_mser = 1; cfg.seed = 7
    cfg.set_views([(1.0, 0.0, 1.0, 0.2), (2.0, 0.0, 1.0, 0.2), (2.0, np.pi / 2, 1.0, 0.2)], [(1.0, 0.0, 1.0, 0.8)])
    single = [ctx.views_sharded_pair(a, b, cfg, capacity=1 << 14) for a, b in pairs]
    res, ver, dig, st = ctx.views_sharded_pairs(pairs, cfg, capacity=1 << 14)
    f = lambda r: (r.regions1, r.regions2, r.mser_regions1, r.mser_regions2, r.tentatives, r.mser_tentatives, r.unique_tentatives, r.ransac_inliers, r.verified)
    for (r1, v1, d1, _), r2, v2, d2 in zip(single, res, ver, dig):
        assert f(r1) == f(r2) and r1.verified > 50
        assert np.array_equal(v1, v2) and np.array_equal(np.array(r1.H), np.array(r2.H)) and d1 == d2
    assert ctx.views_sharded_pairs([], cfg)[0] == []


def test_mods_multi_and_feature_cache_callers(ctx, tmp_path):
    """mods_multi.cpp (1-to-N, query described once) == mb2_mods_pair per pair; extract_features + the read_pre_extracted flow: the same
    tentatives from the cache files (descriptors are integers; positions carry the 6 significant digits of the reference's text format)."""
    import mods_b200 as mb
    from synth import blob_image, warp_image, gt_homography
    A = blob_image(512, 384, seed=61, n_blobs=500)
    Bs = [warp_image(A, gt_homography(512, 384), seed=62 + k) for k in range(2)]
    cfg = mb.PairConfig.default(); cfg.use_mser = 1; cfg.seed = 3
    multi, vers = ctx.mods_multi(A, Bs, cfg, capacity=1 << 14)
    f = lambda r: (r.regions1, r.regions2, r.mser_regions1, r.mser_regions2, r.tentatives, r.unique_tentatives, r.ransac_inliers, r.verified)
    for k, B in enumerate(Bs):
        r, v = ctx.mods_pair(A, B, cfg, capacity=1 << 14)
        assert f(r) == f(multi[k]) and r.verified > 100 and np.array_equal(v, vers[k])
    c1, c2 = tmp_path / "a.txt", tmp_path / "b.txt"
    n1 = ctx.extract_features(A, c1, cfg); n2 = ctx.extract_features(Bs[0], c2, cfg)
    assert (n1, n2) == (multi[0].regions1, multi[0].regions2)
    rc, vc = ctx.mods_pair_cached(c1, c2, cfg, capacity=1 << 14)
    assert (rc.regions1, rc.regions2, rc.tentatives) == (multi[0].regions1, multi[0].regions2, multi[0].tentatives)
    assert abs(rc.verified - multi[0].verified) <= 0.03 * multi[0].verified
