"""Parity of the CUDA path (through the C ABI, libmods_b200.so) against the CPU oracle and the
committed reference vectors.  Bit-exact for keypoints, patches, descriptors, tentatives and
residuals; the MSAC sum J to 1e-9 (fixed-order tree sum vs serial sum)."""
import os
import sys

import numpy as np
import pytest

import mods_b200 as mb
import synth

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.fixture(scope="module")
def img():
    return synth.blob_image(640, 480, seed=3)


# ---- detection ---------------------------------------------------------------------------------
def test_pyramid_levels_bit_exact(ctx, oracle, img):
    ctx.hessaff_detect(img, as_regions=False)
    P = oracle.pyramid(img)
    assert len(P["levels"]) >= 25
    for lv in P["levels"]:
        g, _ = ctx.pyramid_level(lv["octave"], lv["level"], False)
        r, _ = ctx.pyramid_level(lv["octave"], lv["level"], True)
        assert np.array_equal(g, lv["blur"]), (lv["octave"], lv["level"])
        assert np.array_equal(r[1:-1, 1:-1], lv["resp"][1:-1, 1:-1]), (lv["octave"], lv["level"])


@pytest.mark.parametrize("as_regions", [False, True])
def test_hessaff_keys_bit_exact(ctx, oracle, img, as_regions):
    g = ctx.hessaff_detect(img, as_regions=as_regions)
    o = oracle.hessaff_detect(img, raw=not as_regions)
    assert len(o) > 1000 and np.array_equal(g, o)


@pytest.mark.parametrize("wh", [(135, 101), (299, 150), (64, 37), (33, 200), (13, 13), (12, 40)])
def test_hessaff_odd_and_tiny_sizes(ctx, oracle, wh):
    """cv::resize rounding (135 -> 68, 299 -> 150), ragged tiles, images too small for one octave."""
    im = synth.blob_image(wh[0], wh[1], seed=wh[0])
    assert np.array_equal(ctx.hessaff_detect(im), oracle.hessaff_detect(im))


def test_hessaff_flat_image_yields_nothing(ctx):
    assert len(ctx.hessaff_detect(np.full((200, 300), 77.0, np.float32))) == 0


def test_hessaff_matches_reference_golden(ctx):
    im = synth.blob_image(320, 240, seed=int(G["s_seed"]))
    assert np.array_equal(ctx.hessaff_detect(im, as_regions=False), G["s_raw"])
    assert np.array_equal(ctx.hessaff_detect(im, as_regions=True), G["s_reg"])


@pytest.mark.parametrize("mode,regs,rel", [(2, 100, -1.0), (4, 2000, -1.0), (4, 50, -1.0), (1, -1, 0.3), (3, -1, 0.25)])
def test_hessaff_detector_modes(ctx, oracle, mode, regs, rel):
    """DetectorMode != FIXED_TH (scale-space-detector.hpp:127-198): all extrema localised + Baumberg, std::sort by |response|, truncation."""
    from oracle.pyoracle import HessParams
    im = synth.blob_image(300, 200, seed=8)
    gp = mb.HessaffParams.default(); gp.mode = mode; gp.reg_number = regs; gp.rel_threshold = rel; gp.rel_reg_number = rel if mode == 3 else -1.0
    hp = HessParams.default(); hp.mode = mode; hp.reg_number = regs; hp.rel_threshold = rel; hp.rel_reg_number = rel if mode == 3 else -1.0
    for as_regions in (False, True):
        g = ctx.hessaff_detect(im, gp, as_regions=as_regions)
        o = oracle.hessaff_detect(im, hp, raw=not as_regions)
        assert len(o) > 20 and np.array_equal(g, o)


def test_hessaff_no_baumberg(ctx, oracle, img):
    from oracle.pyoracle import HessParams
    p = mb.HessaffParams.default(); p.doBaumberg = 0
    po = HessParams.default(); po.doBaumberg = 0
    assert np.array_equal(ctx.hessaff_detect(img, p), oracle.hessaff_detect(img, po))


# ---- orientation -------------------------------------------------------------------------------
@pytest.mark.parametrize("max_angles", [1, 3, 5])
def test_orientation_bit_exact(ctx, oracle, img, max_angles):
    k = oracle.hessaff_detect(img)
    g = ctx.detect_orientation(img, k, mb.OrientationParams(1.0, 41, max_angles, 0.8))
    o = oracle.detect_orientation(img, k, maxAngles=max_angles)
    assert len(o) > 1000 and np.array_equal(g, o)


def test_orientation_empty_and_zero_angles(ctx, img):
    assert len(ctx.detect_orientation(img, np.zeros((0, 9)))) == 0
    k = np.array([[320.0, 240.0, 1, 0, 0, 1, 3.0, 50.0, 1]])
    assert len(ctx.detect_orientation(img, k, mb.OrientationParams(1.0, 41, 0, 0.8))) == 0


# ---- description -------------------------------------------------------------------------------
@pytest.mark.parametrize("root", [1, 0])
def test_describe_bit_exact(ctx, oracle, img, root):
    k = oracle.detect_orientation(img, oracle.hessaff_detect(img))
    k = oracle.reproject(k, np.eye(3), img.shape[1], img.shape[0], 0)[0]
    gd, gp = ctx.describe_sift(img, k, mb.SiftParams(5.1962, 41, 1, root, 0), want_patches=True)
    od, op = oracle.describe(img, k, rootsift=bool(root), want_patches=True)
    assert np.array_equal(gp, op)
    assert np.array_equal(gd.astype(np.float32), od)


def test_describe_large_and_tiny_regions(ctx, oracle):
    """Direct path (P/41 <= 0.4), the 82x82 'needed samples' path up to P ~ 400, no photoNorm, fast extraction."""
    im = synth.blob_image(900, 700, seed=9)
    rng = np.random.default_rng(0)
    ks = []
    for s in (0.4, 0.9, 1.4, 2.0, 3.7, 8.0, 15.0, 24.0, 38.0):
        a = rng.uniform(0, np.pi); r = rng.uniform(1.0, 2.0)
        A = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]) @ np.diag([r, 1 / r])
        ks.append([450.0 + rng.uniform(-20, 20), 350.0 + rng.uniform(-20, 20), A[0, 0], A[0, 1], A[1, 0], A[1, 1], s, 10.0, 1])
    ks.append([5.0, 6.0, 1, 0, 0, 1, 4.0, 10.0, 0])        # sticks out of the image: zero-filled samples
    k = np.array(ks)
    for par, kw in ((mb.SiftParams(5.1962, 41, 1, 1, 0), dict(rootsift=True)),
                    (mb.SiftParams(5.1962, 41, 0, 0, 0), dict(rootsift=False, photoNorm=False)),
                    (mb.SiftParams(5.1962, 41, 1, 1, 1), dict(rootsift=True, fast=True))):
        gd, gp = ctx.describe_sift(im, k, par, want_patches=True)
        od, op = oracle.describe(im, k, want_patches=True, **kw)
        assert np.array_equal(gp, op)
        assert np.array_equal(gd.astype(np.float32), od)


# ---- one view end to end -----------------------------------------------------------------------
def test_view_pipeline_bit_exact(ctx, oracle, img):
    g = ctx.detect_describe_view(img)
    o = oracle.view_pipeline(img)
    assert len(o[0]) > 1000
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
    assert np.array_equal(g[2].astype(np.float32), o[2])


def test_view_pipeline_matches_reference_golden(ctx):
    im = synth.blob_image(320, 240, seed=int(G["s_seed"]))
    det, rep, desc = ctx.detect_describe_view(im)
    assert np.array_equal(det, G["s_det"]) and np.array_equal(rep, G["s_rep"]) and np.array_equal(desc, G["s_desc"])
    if "cat_gray" in G:
        det, rep, desc = ctx.detect_describe_view(G["cat_gray"])
        assert np.array_equal(det, G["cat_det"]) and np.array_equal(rep, G["cat_rep"]) and np.array_equal(desc, G["cat_desc"])


def test_view_pipeline_is_deterministic_at_scale(ctx):
    """1920x1080 (BASELINE config 2 size): run twice, identical; descriptors are valid RootSIFT."""
    im = synth.blob_image(1920, 1080, seed=5, n_blobs=6000)
    a = ctx.detect_describe_view(im)
    b = ctx.detect_describe_view(im)
    assert len(a[0]) > 3000
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    n2 = (a[2].astype(np.float64) ** 2).sum(1)
    assert np.all(np.abs(np.sqrt(n2) - 512.0) < 12.0)  # sqrt(L1-normalised) has unit L2 norm, x512, rounded


# ---- matching ----------------------------------------------------------------------------------
def _match_case(nq, nt, seed):
    t_desc, _ = synth.random_descriptors(nt, seed)
    q_desc, _ = synth.random_descriptors(nq, seed + 100, dup_of=t_desc, dup_frac=0.5)
    rng = np.random.default_rng(seed)
    txy = rng.uniform(0, 1000, size=(nt, 2))
    for i in range(0, nt - 1, 7):  # near-duplicate trains close in the image: consistent 2nd NNs
        t_desc[i + 1] = np.clip(t_desc[i].astype(int) + rng.integers(-2, 3, 128), 0, 255)
        txy[i + 1] = txy[i] + rng.uniform(-5, 5, 2)
    for i in range(3, nt - 1, 11):  # exact duplicates far apart: ties + inconsistent neighbours
        t_desc[i + 1] = t_desc[i]
    return q_desc, t_desc, txy


@pytest.mark.parametrize("impl", ["tc", "simt"])
@pytest.mark.parametrize("shape", [(300, 257, 1), (1000, 1500, 2), (2500, 3100, 3), (1, 60, 4), (40, 7, 5)])
def test_fginn_bit_exact(ctx, oracle, impl, shape, monkeypatch):
    nq, nt, seed = shape
    q, t, txy = _match_case(nq, nt, seed)
    monkeypatch.setenv("MB2_NN_IMPL", impl)
    g = ctx.match_fginn(q, t, txy)
    o = oracle.match_fginn(q.astype(np.float32), t.astype(np.float32), txy)
    assert np.array_equal(g, o)


@pytest.mark.parametrize("ratio,cd", [(0.8, 30.0), (0.85, 10.0), (0.6, 1e9), (0.95, 0.0)])
def test_fginn_thresholds(ctx, oracle, ratio, cd):
    q, t, txy = _match_case(800, 900, 9)
    g = ctx.match_fginn(q, t, txy, ratio=ratio, contradDist=cd)
    o = oracle.match_fginn(q.astype(np.float32), t.astype(np.float32), txy, ratio=ratio, contradDist=cd)
    assert np.array_equal(g, o)


def test_fginn_empty_and_unsupported(ctx):
    q, t, txy = _match_case(10, 20, 1)
    assert len(ctx.match_fginn(q[:0], t, txy)) == 0
    assert len(ctx.match_fginn(q, t[:0], txy[:0])) == 0


@pytest.mark.parametrize("ratio,cd,shape", [(1.0, 30.0, (700, 900, 21)), (1.2, 10.0, (300, 55, 22)), (1.0, 1e9, (200, 400, 23)), (1.0, 30.0, (50, 30, 24))])
def test_fginn_all_points_branch(ctx, oracle, ratio, cd, shape):
    """matchRatio >= 1 (matching.cpp:397-428): NN + first geometrically inconsistent neighbour, or neighbour nn - 1; fewer trains than nn."""
    q, t, txy = _match_case(*shape)
    g = ctx.match_fginn(q, t, txy, ratio=ratio, contradDist=cd)
    o = oracle.match_fginn(q.astype(np.float32), t.astype(np.float32), txy, ratio=ratio, contradDist=cd)
    assert np.array_equal(g, o) and (len(o) > 0 or shape[1] < 50)


def test_fginn_full_size_properties(ctx):
    """30k x 30k (BASELINE config 3 size), where the CPU oracle would take minutes: tcgen05 and the
    SIMT cross-check kernel must agree exactly, planted matches must be found, indices are valid."""
    nt = nq = 30000
    t_desc, _ = synth.random_descriptors(nt, 1)
    q_desc, src = synth.random_descriptors(nq, 2, dup_of=t_desc, dup_frac=0.4, jitter=4)
    txy = np.random.default_rng(3).uniform(0, 4096, size=(nt, 2))
    a = ctx.match_fginn(q_desc, t_desc, txy)
    os.environ["MB2_NN_IMPL"] = "simt"
    try:
        b = ctx.match_fginn(q_desc, t_desc, txy)
    finally:
        os.environ.pop("MB2_NN_IMPL", None)
    assert np.array_equal(a, b)
    planted = np.flatnonzero(src >= 0)
    found = {int(r[0]): int(r[1]) for r in a}
    hits = sum(1 for qi in planted if found.get(int(qi)) == int(src[qi]))
    assert hits > 0.9 * len(planted)
    assert a[:, 1].max() < nt and a[:, 2].max() < nt and np.all(a[:, 4] <= a[:, 6]) and np.all(a[:, 6] <= a[:, 5])


def test_match_slots_equals_host_path(ctx):
    A = synth.blob_image(640, 480, seed=3)
    B = synth.warp_image(A, synth.gt_homography(640, 480))
    da, ra, ua = ctx.detect_describe_view(A, slot=0)
    db, rb, ub = ctx.detect_describe_view(B, slot=1)
    g = ctx.match_slots(0, 1)
    h = ctx.match_fginn(ua, ub, np.ascontiguousarray(rb[:, :2]))
    assert len(g) > 50 and np.array_equal(g, h)


# ---- verification ------------------------------------------------------------------------------
@pytest.mark.parametrize("which", range(6))
def test_batched_scorer(ctx, oracle, which):
    rng = np.random.default_rng(5)
    n, K = 3000, 33
    u = np.zeros((n, 6)); u[:, 0:2] = rng.random((n, 2)) * 1000; u[:, 2] = 1; u[:, 5] = 1
    Hgt = synth.gt_homography(1000, 1000)
    p = (Hgt @ u[:, 0:3].T).T; u[:, 3:5] = p[:, :2] / p[:, 2:3] + rng.normal(size=(n, 2))
    u[n // 2:, 3:5] = rng.random((n - n // 2, 2)) * 1000
    M0 = np.linalg.inv(Hgt).T.ravel()
    models = np.stack([M0 * (1 + 1e-3 * rng.normal(size=9)) for _ in range(K)])
    I, J, R = ctx.score_models(which, u, models, 9.0, want_resid=True)
    Ro = np.stack([oracle.score(which, u, m) for m in models])
    assert np.array_equal(R, Ro)                                   # residuals: bit-exact f64
    assert np.array_equal(I, (Ro <= 9.0).sum(1))                   # inlier counts: exact
    Jo = np.where(Ro >= 9.0 * 9 / 4, 0.0, 1 - (Ro / (9.0 * 9 / 4))).sum(1)
    assert np.allclose(J, Jo, rtol=1e-12, atol=1e-9)               # MSAC sum: summation order only


def test_scorer_matches_reference_golden(ctx):
    for which in range(5):
        _, _, R = ctx.score_models(which, G["r_u"], G["r_M"][None, :], 9.0, want_resid=True)
        assert np.array_equal(R[0], G["r_scores"][which])


# ---- LO-RANSAC homography ----------------------------------------------------------------------
def _h_close(a, b):
    a = np.asarray(a) / np.linalg.norm(a); b = np.asarray(b) / np.linalg.norm(b)
    return min(np.abs(a - b).max(), np.abs(a + b).max())


def test_ransac_h_matches_reference_golden(ctx):
    """exp_ransacHcustom of the reference (seed 12345 through its srand(time(NULL)) hook) on a planted
    homography: same hypothesis sequence => same sample / LO / rejection counts, same inlier mask."""
    res = ctx.ransac_h(G["r_u"], seed=12345)
    I, samples, lo, rej = [int(v) for v in G["r_stats"]]
    assert (res["I"], res["samples"], res["lo"], res["rejected"]) == (I, samples, lo, rej)
    assert np.array_equal(res["inl"], G["r_inl"])
    assert _h_close(res["H"], G["r_H"]) < 1e-9
    assert abs(res["J"] - float(G["r_J"])) < 1e-9


@pytest.mark.parametrize("seed,inl_frac,n", [(1, 0.6, 500), (2, 0.3, 800), (3, 0.9, 200), (7, 0.15, 1500), (11, 0.5, 20)])
def test_ransac_h_vs_reference_build(ctx, reference, seed, inl_frac, n):
    rng = np.random.default_rng(seed)
    u = np.zeros((n, 6)); u[:, 0:2] = rng.random((n, 2)) * 800; u[:, 2] = 1; u[:, 5] = 1
    Hgt = synth.gt_homography(800, 800)
    p = (Hgt @ u[:, 0:3].T).T; u[:, 3:5] = p[:, :2] / p[:, 2:3] + rng.normal(size=(n, 2))
    k = int(n * inl_frac)
    u[k:, 3:5] = rng.random((n - k, 2)) * 800
    for et in (0, 2):
        r = reference.exp_ransacH(u, seed=seed * 17, errorType=et)
        g = ctx.ransac_h(u, seed=seed * 17, errorType=et)
        assert (g["I"], g["samples"], g["lo"], g["rejected"]) == (r["I"], r["samples"], r["lo"], r["rejected"])
        assert np.array_equal(g["inl"], r["inl"])
        assert _h_close(g["H"], r["H"]) < 1e-8


def test_ransac_h_degenerate_inputs(ctx):
    assert ctx.ransac_h(np.zeros((0, 6)))["I"] == 0
    u = np.zeros((3, 6)); u[:, 2] = 1; u[:, 5] = 1
    assert ctx.ransac_h(u)["I"] == 0


# ---- DoG flavour of the scale-space detector -----------------------------------------------------------
@pytest.mark.parametrize("mode,regs", [(0, 3000), (4, 200), (2, 150)])
def test_dog_detector_vs_oracle_and_golden(ctx, oracle, mode, regs):
    """mb2_hessaff_params.detectorType = 1 (DET_DOG): response = level - GaussianBlur(level, sigma^2) with up to ~100 taps, un-squared
    threshold, sign point types; keys bit-exact against the oracle (== compiled reference) and the reference's golden vectors."""
    from oracle.pyoracle import HessParams
    im = synth.blob_image(480, 360, seed=33)
    hp = HessParams.dog(); hp.mode = mode; hp.reg_number = regs
    gp = mb.HessaffParams.dog(); gp.mode = mode; gp.reg_number = regs
    a = ctx.hessaff_detect(im, gp, as_regions=False)
    assert len(a) > 100 and np.array_equal(a, oracle.hessaff_detect(im, hp, raw=True)) and set(a[:, 8].astype(int)) <= {10, 11}
    if mode == 0:
        GD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dog_vectors.npz"))
        img = GD["image"].astype(np.float32)
        assert np.array_equal(ctx.hessaff_detect(img, mb.HessaffParams.dog(), as_regions=False), GD["raw_fixed_th"])
        v = ctx.detect_describe_view(img, det=mb.HessaffParams.dog())
        assert np.array_equal(v[0], GD["view_det"]) and np.array_equal(v[2], GD["view_desc"])
        g, o = ctx.detect_describe_view(im, det=gp), oracle.view_pipeline(im, hp=hp)
        assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and np.array_equal(g[2].astype(np.float32), o[2])
        # pyramid levels: the response plane is the level minus its wide blur, bit for bit
        p = oracle.pyramid(im, hp)
        ctx.hessaff_detect(im, gp)
        for lv in p["levels"][:7]:
            assert np.array_equal(ctx.pyramid_level(lv["octave"], lv["level"], want_resp=True)[0], lv["resp"])


# ---- DSPSIFT ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scales,start,end,photo", [(3, 0.5, 1.5, 1), (4, 0.6, 1.2, 0)])
def test_dspsift_vs_oracle_and_golden(ctx, oracle, scales, start, end, photo):
    """mb2_sift_params.dspScales > 0 (imagerepresentation.cpp:1547-1598): raw votes at dspScales + 1 region sizes summed in float, float
    SIFTnorm; bit-exact against the oracle (== compiled reference) and the reference's golden vectors."""
    im = synth.blob_image(480, 360, seed=35)
    k = oracle.detect_orientation(im, oracle.hessaff_detect(im))
    sp = mb.SiftParams.dspsift(scales, start, end); sp.photoNorm = photo
    g = ctx.describe_sift(im, k, sp)
    o = oracle.describe_dsp(im, k, numScales=scales, startCoef=start, endCoef=end, photoNorm=bool(photo))
    assert len(k) > 500 and np.array_equal(g.astype(np.float32), o)
    if photo:
        GD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dsp_vectors.npz"))
        assert np.array_equal(ctx.describe_sift(GD["image"].astype(np.float32), GD["keys"], mb.SiftParams.dspsift()), GD["desc_default"])


# ---- Harris flavour of the scale-space detector ---------------------------------------------------------
@pytest.mark.parametrize("mode,regs", [(0, 1000), (4, 200), (2, 150)])
def test_harris_detector_vs_oracle_and_golden(ctx, oracle, mode, regs):
    """mb2_hessaff_params.detectorType = 2 (DET_HARRIS, pyramid.cpp:283-305): gradient products, three blurs, Harris measure; un-squared
    threshold, sign point types 30 / 31; keys bit-exact against the oracle (== compiled reference) and the reference's golden vectors."""
    from oracle.pyoracle import HessParams
    im = synth.blob_image(480, 360, seed=34)
    hp = HessParams.harris(); hp.mode = mode; hp.reg_number = regs
    gp = mb.HessaffParams.harris(); gp.mode = mode; gp.reg_number = regs
    a = ctx.hessaff_detect(im, gp, as_regions=False)
    assert len(a) > 100 and np.array_equal(a, oracle.hessaff_detect(im, hp, raw=True)) and set(a[:, 8].astype(int)) <= {30, 31}
    if mode == 0:
        GD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "harris_vectors.npz"))
        img = GD["image"].astype(np.float32)
        assert np.array_equal(ctx.hessaff_detect(img, mb.HessaffParams.harris(), as_regions=False), GD["raw_fixed_th"])
        v = ctx.detect_describe_view(img, det=mb.HessaffParams.harris())
        assert np.array_equal(v[0], GD["view_det"]) and np.array_equal(v[2], GD["view_desc"])
        p = oracle.pyramid(im, hp)
        ctx.hessaff_detect(im, gp)
        for lv in p["levels"][:7]:
            assert np.array_equal(ctx.pyramid_level(lv["octave"], lv["level"], want_resp=True)[0], lv["resp"])


# ---- HalfRootSIFT (WxBS tiers) -----------------------------------------------------------------------
def test_half_root_sift_vs_oracle_and_golden(ctx, oracle):
    """Orientations modulo pi (mb2_orientation_params.doHalfSIFT) and the folded 64-entry descriptor (mb2_sift_params.doHalfSIFT):
    bit-exact against the oracle on a fresh image and against the reference's golden vectors; the per-view pass with both flags."""
    GH = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "half_sift_vectors.npz"))
    im = GH["image"].astype(np.float32)
    for maxA in (1, 5):
        op = mb.OrientationParams.default(); op.maxAngles = maxA; op.doHalfSIFT = 1
        assert np.array_equal(ctx.detect_orientation(im, GH["kps"], op), GH["oriented_half_a%d" % maxA])
    sp = mb.SiftParams.default(); sp.doHalfSIFT = 1
    assert np.array_equal(ctx.describe_sift(im, GH["oriented_half_a1"], sp), GH["desc_half"])
    op = mb.OrientationParams.default(); op.doHalfSIFT = 1
    v = ctx.detect_describe_view(im, ori=op, desc=sp)
    assert np.array_equal(v[0], GH["view_det"]) and np.array_equal(v[2], GH["view_desc"])
    # fresh image, both detectors, and RootSIFT described on half-oriented regions (a view listing RootSIFT,HalfRootSIFT)
    im2 = synth.blob_image(480, 360, seed=31)
    for det, dflag in ((mb.HessaffParams.default(), 0), (mb.MserParams.default(), 3)):
        for flags, half_desc in ((7, 1), (5, 0)):
            sp = mb.SiftParams.default(); sp.doHalfSIFT = half_desc
            g = ctx.detect_describe_view(im2, det=det, ori=op, desc=sp)
            o = oracle.view_pipeline(im2, detector=dflag, desc=(5.1962, 41, True, flags))
            assert len(g[0]) > 100 and np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and np.array_equal(g[2].astype(np.float32), o[2])
    # HalfSIFT without RootSIFT is undefined in the reference (SIFTnorm reads past the 64-entry vector): refused
    sp = mb.SiftParams.default(); sp.doHalfSIFT = 1; sp.rootSIFT = 0
    with pytest.raises(mb.Mb2Error):
        ctx.describe_sift(im, GH["oriented_half_a1"], sp)
    # matching two HalfRootSIFT sets: the zero upper halves leave the L2 distances of the 64 entries unchanged
    sp = mb.SiftParams.default(); sp.doHalfSIFT = 1
    B = synth.warp_image(im2, synth.gt_homography(480, 360), seed=32)
    ga = ctx.detect_describe_view(im2, ori=op, desc=sp, slot=0); gb = ctx.detect_describe_view(B, ori=op, desc=sp, slot=1)
    gm = ctx.match_slots(0, 1)
    om = oracle.match_fginn(ga[2].astype(np.float32), gb[2].astype(np.float32), np.ascontiguousarray(gb[1][:, :2]))
    assert len(gm) > 20 and np.array_equal(gm, om)


# ---- LO-RANSAC fundamental matrix (DEGENSAC) -------------------------------------------------------
def _load_f_golden():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_f
    return make_golden_f, np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ransac_f_vectors.npz"))


def test_ransac_f_matches_reference_golden(ctx):
    """exp_ransacFcustom of the reference on four scenes x seeds x error types x inlLimit forms (tests/golden/make_golden_f.py): same
    inlier count, sample count, LO count, best-homography support, inlier mask; F up to scale."""
    mg, GF = _load_f_golden()
    for name, seed, et, lim in mg.CASES:
        key = "%s_s%d_e%d_l%s" % (name, seed, et, lim)
        u = GF["u_" + name]
        r = ctx.ransac_f(u, seed=seed, errorType=et, inlLimit=None if lim == "n" else 0)
        assert [r["I"], r["samples"], r["lo"], r["Ih"]] == GF["stats_" + key].tolist(), key
        assert np.array_equal(r["inl"], np.unpackbits(GF["inl_" + key])[:len(u)]), key
        assert _h_close(r["F"], GF["F_" + key]) < 1e-7, key
        assert r["launches"] > 0


@pytest.mark.parametrize("cfg", [dict(n=300, n_out=200, noise=1.0), dict(n=600, n_out=200, planar_frac=0.8), dict(n=2000, n_out=1200), dict(n=30, n_out=8)])
def test_ransac_f_vs_reference_build(ctx, reference, cfg):
    mg, _ = _load_f_golden()
    u = mg.general_scene(21, **cfg)
    for seed in (5, 6):
        for et in (0, 1):
            for lim in (None, 0):
                a = reference.exp_ransacF(u, seed=seed, errorType=et, inlLimit=lim); b = ctx.ransac_f(u, seed=seed, errorType=et, inlLimit=lim)
                assert [a[k] for k in ("I", "samples", "lo", "Ih")] == [b[k] for k in ("I", "samples", "lo", "Ih")], (cfg, seed, et, lim)
                assert np.array_equal(a["inl"], b["inl"]) and _h_close(a["F"], b["F"]) < 1e-7


def test_ransac_f_full_size_properties(ctx):
    """BASELINE C3 size (30k tentatives, 40 % outliers): no oracle at this size in seconds, so size-independent properties -- the
    mask is exactly {FDs(F) <= th} recomputed through the scorer, it recovers the planted inliers, F has rank 2 and the run is
    repeatable."""
    mg, _ = _load_f_golden()
    n, n_out = 30000, 12000
    u = mg.general_scene(4, n=n, n_out=n_out, noise=0.5)
    r = ctx.ransac_f(u, seed=9, inlLimit=0)
    _, _, R = ctx.score_models(3, u, r["F"][None, :], 9.0, want_resid=True)
    assert np.array_equal(r["inl"], (R[0] <= 9.0).astype(np.uint8))
    assert r["I"] == int(r["inl"].sum()) and r["I"] > 0.9 * (n - n_out)
    assert np.linalg.svd(r["F"].reshape(3, 3))[1][2] < 1e-9 * np.linalg.svd(r["F"].reshape(3, 3))[1][0]
    r2 = ctx.ransac_f(u, seed=9, inlLimit=0)
    assert np.array_equal(r["inl"], r2["inl"]) and np.array_equal(r["F"], r2["F"])


def test_ransac_f_degenerate_inputs(ctx):
    assert ctx.ransac_f(np.zeros((0, 6)))["I"] == 0
    assert ctx.ransac_f(np.ones((20, 6)))["inl"].sum() == 0


# ---- one mods.cpp iteration through the host mirror (libmods_host.so) -------------------------------
def _dup_filter_bruteforce(xy, key, r):
    """matching.cpp:2983-3047: stable sort by key, drop an entry when a kept one is within r in both images."""
    order = np.argsort(key, kind="stable")
    kept = []
    for j in order:
        ok = True
        for i in kept:
            if (xy[i, 0] - xy[j, 0]) ** 2 + (xy[i, 1] - xy[j, 1]) ** 2 <= r * r and (xy[i, 2] - xy[j, 2]) ** 2 + (xy[i, 3] - xy[j, 3]) ** 2 <= r * r:
                ok = False
                break
        if ok:
            kept.append(j)
    return np.array(kept, dtype=np.int64)


def test_mods_pair_equals_stage_composition(ctx, oracle):
    import mods_b200 as mb
    from synth import blob_image, warp_image, gt_homography
    A = blob_image(480, 360, seed=21, n_blobs=500)
    B = warp_image(A, gt_homography(480, 360), seed=22)
    cfg = mb.PairConfig.default()
    cfg.seed = 4242
    res, ver = ctx.mods_pair(A, B, cfg, capacity=4096)
    oa, ob = oracle.view_pipeline(A), oracle.view_pipeline(B)
    assert (res.regions1, res.regions2) == (len(oa[0]), len(ob[0]))
    om = oracle.match_fginn(oa[2], ob[2], np.ascontiguousarray(ob[1][:, :2]), ratio=cfg.matchRatio, contradDist=cfg.contradDist)
    assert res.tentatives == len(om)
    qi, ti = om[:, 0].astype(int), om[:, 1].astype(int)
    xy = np.concatenate([oa[1][qi, :2], ob[1][ti, :2]], axis=1)
    key = np.abs(np.sqrt((om[:, 4].astype(np.float32) / om[:, 5].astype(np.float32)).astype(np.float64)))
    kept = _dup_filter_bruteforce(xy, key, cfg.duplicateDist)
    assert res.unique_tentatives == len(kept)
    u = np.ones((len(kept), 6)); u[:, 0:2] = xy[kept, 0:2]; u[:, 3:5] = xy[kept, 2:4]
    r = ctx.ransac_h(u, th=cfg.err_threshold ** 2, conf=cfg.confidence, max_sam=cfg.max_samples if len(kept) > 20 else 1000,
                     errorType=cfg.errorType, doSymCheck=cfg.doSymmCheck, seed=cfg.seed)
    assert res.ransac_inliers == int(r["inl"].sum())
    assert 8 <= res.verified <= res.ransac_inliers
    # verified pairs obey the ground-truth homography
    Hgt = gt_homography(480, 360)
    p = np.c_[ver[:, :2], np.ones(len(ver))] @ Hgt.T
    assert np.median(np.hypot(p[:, 0] / p[:, 2] - ver[:, 2], p[:, 1] / p[:, 2] - ver[:, 3])) < 1.5
    # the same pair verified in epipolar mode (RANSACPars.useF: exp_ransacFcustom with inlLimit 0 + F_LAF_check, matching.cpp:875-971)
    cfg.useF = 1
    resF, verF = ctx.mods_pair(A, B, cfg, capacity=4096)
    assert (resF.tentatives, resF.unique_tentatives) == (res.tentatives, res.unique_tentatives)
    rf = ctx.ransac_f(u, th=cfg.err_threshold ** 2, conf=cfg.confidence, max_sam=cfg.max_samples, errorType=0 if cfg.errorType == 0 else 1,
                      doSymCheck=cfg.doSymmCheck, seed=cfg.seed, do_lo=cfg.localOptimization, inlLimit=0)
    assert resF.ransac_inliers == int(rf["inl"].sum()) and _h_close(np.array(resF.H), rf["F"]) < 1e-12
    assert 8 <= resF.verified <= resF.ransac_inliers
    p = np.c_[verF[:, :2], np.ones(len(verF))] @ Hgt.T
    assert np.median(np.hypot(p[:, 0] / p[:, 2] - verF[:, 2], p[:, 1] / p[:, 2] - verF[:, 3])) < 1.5


def test_mods_pair_two_streams_equals_one(ctx):
    """Both images on two contexts / host threads (default) == everything on one stream (profiling mode)."""
    import mods_b200 as mb
    from synth import blob_image, warp_image, gt_homography
    A = blob_image(640, 480, seed=31, n_blobs=900)
    B = warp_image(A, gt_homography(640, 480), seed=32)
    cfg = mb.PairConfig.default()
    cfg.seed = 7
    r2, v2 = ctx.mods_pair(A, B, cfg, capacity=8192)
    ctx.profile_begin()
    r1, v1 = ctx.mods_pair(A, B, cfg, capacity=8192)
    prof = ctx.profile_end()
    assert any("k_nn_tc" in k or "k_nn_simt" in k for k in prof)
    for f in ("regions1", "regions2", "tentatives", "unique_tentatives", "ransac_inliers", "verified"):
        assert getattr(r1, f) == getattr(r2, f), f
    assert np.array_equal(v1, v2) and np.array_equal(np.array(r1.H[:]), np.array(r2.H[:]))


def test_slot_move(ctx):
    import mods_b200 as mb
    from synth import blob_image
    A = blob_image(320, 240, seed=11)
    c2 = mb.Context(0)
    try:
        ga = ctx.detect_describe_view(A, slot=0)
        gb = c2.detect_describe_view(A, slot=3)
        c2.slot_move_to(ctx, 1, 3)
        m = ctx.match_slots(0, 1, ratio=0.8, contradDist=0.0)
        assert len(gb[0]) == len(ga[0]) and len(m) > 0 and np.array_equal(m[:, 0], m[:, 1])  # identical sets: every match is i -> i
    finally:
        c2.close()


def test_mods_pairs_pipeline_equals_single_calls(ctx):
    import mods_b200 as mb
    from synth import blob_image, warp_image, gt_homography
    pairs = []
    for k in range(4):
        A = blob_image(400 + 40 * k, 300 + 8 * k, seed=41 + k, n_blobs=350 + 50 * k)
        pairs.append((A, warp_image(A, gt_homography(A.shape[1], A.shape[0]), seed=51 + k)))
    cfg = mb.PairConfig.default()
    cfg.seed = 99
    single = [ctx.mods_pair(a, b, cfg, capacity=4096) for a, b in pairs]
    res, ver = ctx.mods_pairs(pairs, cfg, capacity=4096)
    assert len(res) == 4
    for (r1, v1), r2, v2 in zip(single, res, ver):
        for f in ("regions1", "regions2", "tentatives", "unique_tentatives", "ransac_inliers", "verified"):
            assert getattr(r1, f) == getattr(r2, f), f
        assert r1.verified >= 8 and np.array_equal(v1, v2) and np.array_equal(np.array(r1.H[:]), np.array(r2.H[:]))
    assert ctx.mods_pairs([], cfg)[0] == []


# ---- Hamming matcher (MatchFLANNDistance, matching.cpp:607-666) --------------------------------------------------------
def _hamming_rows(o):
    """oracle rows (q idx0 idx1 d0 d1 ratio) in the C ABI's 7-column layout (q idx0 idx1 idx1 d0 d1 d1)"""
    return o[:, [0, 1, 2, 2, 3, 4, 4]]


def test_hamming_matcher_golden_and_oracle(ctx, oracle):
    """Bit-exact against the compiled reference's golden vectors (descriptor lengths 16 / 20 / 32 / 64 bytes, duplicates among the
    trains, only two trains, thresholds 64 / 30.5 / 0) and against the oracle on fresh, larger inputs."""
    import sys
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hamming_vectors.npz"))
    for name in sorted(k[2:] for k in G.files if k.startswith("q_")):
        q, t = G["q_" + name], G["t_" + name]
        for th in (64.0, 30.5, 0.0):
            want = G["rows_%s_%g" % (name, th)]            # q second.id d1 d2 ratio
            got = ctx.match_hamming(q, t, th)
            assert np.array_equal(got[:, [0, 1, 4, 5]], want[:, :4]), (name, th)
            with np.errstate(divide="ignore", invalid="ignore"):
                assert np.array_equal(got[:, 4] / got[:, 5], want[:, 4], equal_nan=True)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_hamming import case
    for nq, nt, nb, th in ((5000, 7000, 32, 64.0), (3000, 20000, 64, 150.0), (129, 513, 7, 9.0), (20000, 300, 32, 80.0)):
        q, t = case(nq, nt, nb, seed=nq)
        g = ctx.match_hamming(q, t, th)
        o = oracle.match_hamming(q.astype(np.float32), t.astype(np.float32), th)
        assert len(o) > 0 and np.array_equal(g, _hamming_rows(o)), (nq, nt, nb)


def test_hamming_matcher_edges(ctx):
    q = np.zeros((4, 32), np.uint8); t = np.zeros((3, 32), np.uint8)
    assert len(ctx.match_hamming(q[:0], t)) == 0 and len(ctx.match_hamming(q, t[:0])) == 0
    with pytest.raises(mb.Mb2Error):
        ctx.match_hamming(q, t[:1])                         # knnSearch with knn = 2 on one train: undefined in the reference, refused
    with pytest.raises(mb.Mb2Error):
        ctx.match_hamming(np.zeros((4, 65), np.uint8), np.zeros((3, 65), np.uint8))
    r = ctx.match_hamming(q, t, 0.0)                        # all equal: distance 0, first two trains in index order
    assert np.array_equal(r[:, 1], np.zeros(4)) and np.array_equal(r[:, 2], np.ones(4)) and not r[:, 4:].any()
