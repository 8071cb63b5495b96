"""Host test of the MSER decision logic shared with the CUDA kernels (mods_b200/csrc/mser_logic.cuh): the canonical
component tree is built sequentially by tests/native/mser_tree_cpu.cpp, then survivors, births, equal-size merges,
stability thresholds and run moments go through the same functions the GPU compiles.  Compared with the oracle (which
follows the reference's pixel-by-pixel algorithm) on the inputs where the two formulations could disagree."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def tree():
    src = os.path.join(HERE, "native", "mser_tree_cpu.cpp")
    so = os.path.join(HERE, "native", "libmser_tree_cpu.so")
    hdr = os.path.join(HERE, "..", "mods_b200", "csrc", "mser_logic.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    lib = C.CDLL(so)

    def run(img, max_area=0.05, min_size=30, min_margin=8.0, max_out=100000):
        img = np.ascontiguousarray(img, np.float32)
        out = np.zeros((max_out, 13)); st = (C.c_int * 4)()
        n = lib.mser_tree_regions(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], C.c_double(max_area), C.c_int(min_size),
                                  C.c_double(min_margin), out.ctypes.data_as(C.c_void_p), max_out, st)
        return out[:n].copy(), list(st)
    return run


def test_blobs(tree, oracle):
    for seed in (1, 2):
        img = synth.blob_image(200, 150, seed=seed)
        t, st = tree(img)
        assert np.array_equal(t, oracle.mser_regions(img)) and len(t) > 50 and st[2] == 0 and st[3] == 0


def test_order_dependent_corners(tree, oracle):
    rng = np.random.default_rng(7)
    ties = 0
    for trial in range(120):
        kind = trial % 4
        if kind == 0:
            k = int(rng.integers(1, 5))
            img = (np.kron(rng.integers(0, 5, (int(rng.integers(2, 9)), int(rng.integers(2, 9)))), np.ones((k, k))) * int(rng.integers(1, 50))).astype(np.float32)
            kw = dict(min_margin=float(rng.integers(1, 4)), min_size=int(rng.integers(2, 10)), max_area=0.9)
        elif kind == 1:
            img = rng.integers(0, 256, (int(rng.integers(5, 50)), int(rng.integers(5, 50)))).astype(np.float32)
            kw = dict(min_margin=float(rng.integers(1, 6)), min_size=int(rng.integers(2, 12)), max_area=0.5)
        elif kind == 2:
            img = (rng.integers(0, 8, (int(rng.integers(5, 40)), int(rng.integers(5, 40)))) * 30).astype(np.float32)
            kw = dict(min_margin=1.0, min_size=int(rng.integers(2, 20)), max_area=0.9)
        else:
            a = rng.random((int(rng.integers(4, 30)), int(rng.integers(4, 30))))
            img = np.kron(np.floor(a * int(rng.integers(2, 6))) * 40, np.ones((2, 3))).astype(np.float32)
            img[rng.random(img.shape) < 0.1] = 255
            kw = dict(min_margin=1.0, min_size=4, max_area=0.7)
        t, st = tree(img, **kw)
        ties += st[1]
        assert np.array_equal(t, oracle.mser_regions(img, **kw)), (trial, img.shape, kw)
    assert ties > 50  # the equal-size branch was really exercised


def test_golden_plateaus(tree):
    GM = np.load(os.path.join(HERE, "golden", "mser_vectors.npz"))
    t, _ = tree(GM["p_img"].astype(np.float32), max_area=0.3, min_size=8, min_margin=1.0)
    assert np.array_equal(t, GM["p_regions"])
