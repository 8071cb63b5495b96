"""The CPU restatement (oracle/) against vectors produced by the reference's own code
(tests/golden/make_golden.py ran oracle/_ref/libmods_ref.so == /root/reference compiled in place).
Bit-exact everywhere: these are the pins that make the oracle trustworthy on the GPU box, where
/root/reference does not exist."""
import os

import numpy as np
import pytest

import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def test_hessaff_keys_match_reference(oracle):
    img = synth.blob_image(320, 240, seed=int(G["s_seed"]))
    assert np.array_equal(oracle.hessaff_detect(img, raw=True), G["s_raw"])
    assert np.array_equal(oracle.hessaff_detect(img, raw=False), G["s_reg"])


def test_orientation_matches_reference(oracle):
    img = synth.blob_image(320, 240, seed=int(G["s_seed"]))
    assert np.array_equal(oracle.detect_orientation(img, G["s_reg"]), G["s_ori"])


def test_view_pipeline_matches_reference(oracle):
    img = synth.blob_image(320, 240, seed=int(G["s_seed"]))
    det, rep, desc = oracle.view_pipeline(img)
    assert np.array_equal(det, G["s_det"]) and np.array_equal(rep, G["s_rep"])
    assert np.array_equal(desc.astype(np.uint8), G["s_desc"]) and desc.max() <= 255


def test_cat_crop_matches_reference(oracle):
    """300x300 crop of the reference's example image build/examples/cat.png."""
    if "cat_gray" not in G:
        pytest.skip("cat crop not in golden file")
    det, rep, desc = oracle.view_pipeline(G["cat_gray"])
    assert np.array_equal(det, G["cat_det"]) and np.array_equal(rep, G["cat_rep"])
    assert np.array_equal(desc.astype(np.uint8), G["cat_desc"])


@pytest.mark.parametrize("which", range(5))
def test_scorers_match_reference(oracle, which):
    """HDs, HDsSym, HDsSymMax (Htools.c), FDs, FDsSym (Ftools.c): closed-form f64, bit-exact."""
    assert np.array_equal(oracle.score(which, G["r_u"], G["r_M"]), G["r_scores"][which])


# ---- MSER (row a9): vectors from the reference's own detectors/mser sources (tests/golden/make_golden_mser.py)
GM = np.load(os.path.join(os.path.dirname(__file__), "golden", "mser_vectors.npz"))


def test_mser_cat_crop_matches_reference(oracle):
    assert np.array_equal(oracle.mser_detect(G["cat_gray"]), G["cat_mser"])
    assert np.array_equal(oracle.mser_regions(G["cat_gray"]), GM["cat_regions"])
    assert np.array_equal(oracle.mser_detect(G["cat_gray"]), GM["cat_keys"])


def test_mser_synthetic_matches_reference(oracle):
    img = synth.blob_image(320, 240, seed=11)
    assert np.array_equal(oracle.mser_regions(img), GM["s_regions"]) and len(GM["s_regions"]) > 100
    assert np.array_equal(oracle.mser_detect(img), GM["s_keys"])


def test_mser_plateaus_and_noise_match_reference(oracle):
    """Equal-size merges, regions born inside a level, tiny min_size: the order-dependent corners of getExtrema.cpp."""
    p = GM["p_img"].astype(np.float32)
    assert np.array_equal(oracle.mser_regions(p, max_area=0.3, min_size=8, min_margin=1.0), GM["p_regions"])
    assert np.array_equal(oracle.mser_detect(p, max_area=0.3, min_size=8, min_margin=1.0), GM["p_keys"])
    n = GM["n_img"].astype(np.float32)
    assert np.array_equal(oracle.mser_regions(n, min_size=5, min_margin=2.0), GM["n_regions"])


def test_half_root_sift_matches_reference(oracle):
    """HalfRootSIFT (orientations modulo pi + folded 64-entry descriptor): tests/golden/make_golden_half.py."""
    GH = np.load(os.path.join(os.path.dirname(__file__), "golden", "half_sift_vectors.npz"))
    im = GH["image"].astype(np.float32)
    for maxA in (1, 5):
        assert np.array_equal(oracle.detect_orientation(im, GH["kps"], maxAngles=maxA, doHalfSIFT=1), GH["oriented_half_a%d" % maxA])
    assert np.array_equal(oracle.describe(im, GH["oriented_half_a1"], rootsift=3).astype(np.uint8), GH["desc_half"])
    v = oracle.view_pipeline(im, detector=0, desc=(5.1962, 41, True, 7))
    assert np.array_equal(v[0], GH["view_det"]) and np.array_equal(v[2].astype(np.uint8), GH["view_desc"])


def test_dog_detector_matches_reference(oracle):
    """DET_DOG flavour of the scale-space detector: tests/golden/make_golden_dog.py."""
    from oracle.pyoracle import HessParams
    GD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dog_vectors.npz"))
    im = GD["image"].astype(np.float32)
    hp = HessParams.dog()
    assert np.array_equal(oracle.hessaff_detect(im, hp, raw=True), GD["raw_fixed_th"]) and len(GD["raw_fixed_th"]) > 50
    hp.mode = 4; hp.reg_number = 80
    assert np.array_equal(oracle.hessaff_detect(im, hp, raw=True), GD["raw_not_less_80"])
    v = oracle.view_pipeline(im, hp=HessParams.dog())
    assert np.array_equal(v[0], GD["view_det"]) and np.array_equal(v[2].astype(np.uint8), GD["view_desc"])


def test_harris_detector_matches_reference(oracle):
    """DET_HARRIS flavour of the scale-space detector (pyramid.cpp:283-305): tests/golden/make_golden_harris.py."""
    from oracle.pyoracle import HessParams
    GD = np.load(os.path.join(os.path.dirname(__file__), "golden", "harris_vectors.npz"))
    im = GD["image"].astype(np.float32)
    hp = HessParams.harris()
    assert np.array_equal(oracle.hessaff_detect(im, hp, raw=True), GD["raw_fixed_th"]) and len(GD["raw_fixed_th"]) > 50
    hp.mode = 4; hp.reg_number = 80
    assert np.array_equal(oracle.hessaff_detect(im, hp, raw=True), GD["raw_not_less_80"])
    v = oracle.view_pipeline(im, hp=HessParams.harris())
    assert np.array_equal(v[0], GD["view_det"]) and np.array_equal(v[2].astype(np.uint8), GD["view_desc"])


def test_dspsift_matches_reference(oracle):
    """DSPSIFT (imagerepresentation.cpp:1547-1598): tests/golden/make_golden_dsp.py."""
    GD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dsp_vectors.npz"))
    im = GD["image"].astype(np.float32)
    assert len(GD["keys"]) > 100
    assert np.array_equal(oracle.describe_dsp(im, GD["keys"]).astype(np.uint8), GD["desc_default"])
    assert np.array_equal(oracle.describe_dsp(im, GD["keys"], numScales=5, startCoef=0.7, endCoef=1.3, photoNorm=False).astype(np.uint8), GD["desc_5_07_13_nophoto"])


def test_hamming_matcher_golden(oracle):
    """MatchFLANNDistance (matching.cpp:607-666): the oracle reproduces the compiled reference's tentatives on the committed vectors
    (tests/golden/make_golden_hamming.py): descriptor lengths 16 / 20 / 32 / 64 bytes, duplicates among the trains, two trains only."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hamming_vectors.npz"))
    names = sorted(k[2:] for k in G.files if k.startswith("q_"))
    assert len(names) == 5
    for name in names:
        q, t = G["q_" + name].astype(np.float32), G["t_" + name].astype(np.float32)
        for th in (64.0, 30.5, 0.0):
            want = G["rows_%s_%g" % (name, th)]
            got = oracle.match_hamming(q, t, th)
            assert np.array_equal(got[:, [0, 1, 3, 4, 5]], want, equal_nan=True), (name, th)
