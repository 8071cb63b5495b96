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
