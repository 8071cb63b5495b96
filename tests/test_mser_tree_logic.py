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


def _lib():
    src = os.path.join(HERE, "native", "mser_tree_cpu.cpp")
    so = os.path.join(HERE, "native", "libmser_tree_cpu.so")
    hdrs = [os.path.join(HERE, "..", "mods_b200", "csrc", h) for h in ("mser_logic.cuh", "mser_tree_build.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-o", so, src])
    return C.CDLL(so)


# 0: the sequential level-by-level builder; 1 / 2: the lock-free merge the GPU runs (mser_tree_build.cuh) -- sequentially with tiles of
# 8 and 64 pixels, and on 4 host threads with real compare-and-swap races
@pytest.fixture(scope="module", params=[(0, 64, 1), (1, 8, 1), (1, 64, 1), (2, 16, 4)], ids=["levelwise", "merge-t8", "merge-t64", "merge-4threads"])
def tree(request):
    lib = _lib()
    mode, tile, threads = request.param

    def run(img, max_area=0.05, min_size=30, min_margin=8.0, max_out=100000):
        lib.mser_tree_set_mode(mode, tile, threads, 12345)
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


def test_lockfree_merge_builds_the_same_tree():
    """mser_tree_build.cuh against the level-by-level builder, node by node (level, area, inner edges of every pixel's node and of its
    parent node): random edge orders, tile sizes, plateaus, noise, and 8 host threads racing on the same words."""
    lib = _lib()
    rng = np.random.default_rng(11)
    n = 0
    for trial in range(60):
        h, w = int(rng.integers(3, 90)), int(rng.integers(3, 90))
        kind = trial % 5
        if kind == 0:
            img = rng.integers(0, 256, (h, w))
        elif kind == 1:
            img = rng.integers(0, 4, (h, w)) * 60                       # big plateaus
        elif kind == 2:
            img = np.full((h, w), 128) + rng.integers(-6, 7, (h, w))     # the bench image's background: few levels, one giant component
        elif kind == 3:
            img = synth.blob_image(w + 10, h + 10, seed=trial)[:h, :w]
        else:
            img = np.kron(rng.integers(0, 256, ((h + 3) // 4, (w + 3) // 4)), np.ones((4, 4)))[:h, :w]
        img = np.ascontiguousarray(img, np.float32)
        for mode, tile, threads in ((1, 4, 1), (1, 64, 1), (2, 8, 8), (2, 1 << 20, 8)):
            lib.mser_tree_set_mode(mode, tile, threads, 1000 + trial)
            for pol in (0, 1):
                assert lib.mser_tree_compare(img.ctypes.data_as(C.c_void_p), w, h, pol) == 0, (trial, mode, tile, threads, pol)
                n += 1
    lib.mser_tree_set_mode(0, 64, 1, 1)
    assert n == 480
