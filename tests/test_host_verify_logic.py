"""Host arithmetic of the verification stage (mods_b200/host/mods_host.cpp: duplicate filter -> LORANSACFiltering glue -> NaiveHCheck /
H_LAF_check / F_LAF_check) without a GPU: tests/native/host_verify_cpu.cpp includes the host mirror unchanged, plants the model and
inlier mask that the RANSAC driver would return and lets the ORACLE's residual functions stand in for mb2_score_models.  Expected
values are a numpy restatement of matching.cpp:193-309, 806-980, 1171-1200, 2983-3047."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mods_b200 as mb
import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def hv(oracle):
    mb.build()
    src = os.path.join(HERE, "native", "host_verify_cpu.cpp")
    so = os.path.join(HERE, "native", "libhost_verify_cpu.so")
    deps = [src, os.path.join(ROOT, "mods_b200", "host", "mods_host.cpp"), os.path.join(ROOT, "mods_b200", "host", "mods_host.hpp"),
            os.path.join(ROOT, "include", "mods_b200.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
        libdir = os.path.join(ROOT, "mods_b200")
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src, "-L" + libdir, "-lmods_b200",
                               "-Wl,-rpath," + libdir])
    lib = C.CDLL(so)
    score = getattr(oracle.lib, oracle.prefix + "score")

    def run(frames, keys, cfg, model, mask):
        frames = np.ascontiguousarray(frames, np.float64); keys = np.ascontiguousarray(keys, np.float64)
        model = np.ascontiguousarray(model, np.float64); mask = np.ascontiguousarray(mask, np.uint8)
        lib.t_plant(score, _p(model), _p(mask), C.c_int(len(mask)))
        res = mb.PairResult(); out = np.zeros((max(1, len(keys)), 4))
        dummy_ctx = C.c_void_p(1)   # never dereferenced: every service that would use it is replaced
        k = lib.mb2_host_verify(dummy_ctx, _p(frames), _p(keys), C.c_int(len(keys)), C.byref(cfg), C.byref(res), _p(out), C.c_int(len(out)))
        last = np.zeros(3, np.int32); lib.t_last_call(_p(last))
        return k, res, out[:max(k, 0)], last
    return run


def make_frames(n=400, n_dup=60, seed=3):
    """Tentatives obeying a homography (plus outliers and near-duplicates), with local affine frames that obey its Jacobian."""
    rng = np.random.default_rng(seed)
    Hgt = synth.gt_homography(800, 600)
    f = np.zeros((n, 14))
    f[:, 0] = rng.random(n) * 700 + 50; f[:, 1] = rng.random(n) * 500 + 50
    p = np.c_[f[:, 0:2], np.ones(n)] @ Hgt.T
    f[:, 7:9] = p[:, :2] / p[:, 2:3] + rng.normal(size=(n, 2)) * 0.7
    for i in range(n):
        a = rng.normal(size=(2, 2)) * 0.3 + np.eye(2); s = 3 + 10 * rng.random()
        f[i, 2:6] = a.ravel(); f[i, 6] = s
        x, y = f[i, 0], f[i, 1]; w = Hgt[2, 0] * x + Hgt[2, 1] * y + Hgt[2, 2]
        Jm = (Hgt[:2, :2] - np.outer(p[i, :2] / p[i, 2], Hgt[2, :2])) / w
        b = Jm @ (a * s); s2 = np.sqrt(abs(np.linalg.det(b))); f[i, 9:13] = (b / s2).ravel(); f[i, 13] = s2
    out = rng.choice(n, n // 4, replace=False)
    f[out, 7:9] = rng.random((len(out), 2)) * [800, 600]
    bad_frames = rng.choice(n, n // 8, replace=False)          # right position, wrong local frame: the LAF checks must drop them
    f[bad_frames, 9:13] = rng.normal(size=(len(bad_frames), 4)) * 2
    f[:n_dup] = f[n_dup:2 * n_dup] + rng.normal(size=(n_dup, 14)) * 0.05
    keys = rng.random(n) * 0.8
    return f, keys, Hgt


def dup_filter(xy, key, r):   # matching.cpp:2983-3047, stable order
    kept = []
    for j in np.argsort(key, kind="stable"):
        if not any((xy[i, 0] - xy[j, 0]) ** 2 + (xy[i, 1] - xy[j, 1]) ** 2 <= r * r and (xy[i, 2] - xy[j, 2]) ** 2 + (xy[i, 3] - xy[j, 3]) ** 2 <= r * r
                   for i in kept):
            kept.append(j)
    return np.array(kept, dtype=np.int64)


def laf_points(f):   # matching.cpp:209-232 / 268-291, k_sigma = 3
    u = np.ones((len(f), 3, 6))
    u[:, 0, 0:2] = f[:, 0:2]; u[:, 0, 3:5] = f[:, 7:9]
    u[:, 1, 0] = f[:, 0] + 3.0 * f[:, 3] * f[:, 6]; u[:, 1, 1] = f[:, 1] + 3.0 * f[:, 5] * f[:, 6]
    u[:, 1, 3] = f[:, 7] + 3.0 * f[:, 10] * f[:, 13]; u[:, 1, 4] = f[:, 8] + 3.0 * f[:, 12] * f[:, 13]
    u[:, 2, 0] = f[:, 0] + 3.0 * f[:, 2] * f[:, 6]; u[:, 2, 1] = f[:, 1] + 3.0 * f[:, 4] * f[:, 6]
    u[:, 2, 3] = f[:, 7] + 3.0 * f[:, 9] * f[:, 13]; u[:, 2, 4] = f[:, 8] + 3.0 * f[:, 11] * f[:, 13]
    return u.reshape(-1, 6)


@pytest.mark.parametrize("errorType", [0, 1, 2])
def test_homography_mode_post_checks(hv, oracle, errorType):
    f, keys, Hgt = make_frames()
    cfg = mb.PairConfig.default(); cfg.errorType = errorType
    kept = dup_filter(np.c_[f[:, 0:2], f[:, 7:9]], keys, cfg.duplicateDist)
    fk = f[kept]
    p = np.c_[fk[:, 0:2], np.ones(len(fk))] @ Hgt.T
    mask = (np.hypot(p[:, 0] / p[:, 2] - fk[:, 7], p[:, 1] / p[:, 2] - fk[:, 8]) < 3).astype(np.uint8)
    Hloran = np.linalg.inv(Hgt).T.ravel()                       # DEGENSAC's convention: column-wise, second image -> first
    k, res, ver, last = hv(f, keys, cfg, Hloran, mask)
    assert (res.tentatives, res.unique_tentatives, res.ransac_inliers) == (len(f), len(kept), int(mask.sum()))
    assert tuple(last[:2]) == (errorType, cfg.max_samples)
    assert np.allclose(np.array(res.H).reshape(3, 3), np.linalg.inv(Hloran.reshape(3, 3).T), rtol=1e-12)
    cand = fk[mask == 1]
    e = oracle.score(2, laf_points(cand), Hloran).reshape(-1, 3)   # H_LAF_check always scores with HDsSymMax (matching.cpp:951)
    good = ~(np.sqrt(e.sum(1)) > 3.0 * cfg.HLAFCoef * cfg.err_threshold)
    exp = cand[good]
    assert 50 < len(exp) < len(cand) and k == len(exp)
    assert np.array_equal(ver, np.c_[exp[:, 0:2], exp[:, 7:9]])


@pytest.mark.parametrize("errorType", [0, 2])
def test_epipolar_mode_post_checks(hv, oracle, errorType):
    f, keys, Hgt = make_frames(seed=5)
    cfg = mb.PairConfig.default(); cfg.errorType = errorType; cfg.useF = 1
    kept = dup_filter(np.c_[f[:, 0:2], f[:, 7:9]], keys, cfg.duplicateDist)
    fk = f[kept]
    # a fundamental matrix compatible with the homography: F = [e']x H for an arbitrary epipole e', stored the way FDs reads it
    e2 = np.array([900.0, 300.0, 1.0]); Ex = np.array([[0, -e2[2], e2[1]], [e2[2], 0, -e2[0]], [-e2[1], e2[0], 0]])
    Fm = (Ex @ Hgt).ravel()                 # FDs evaluates x2^T M x1 with M = F row-major
    which = 3 if errorType == 0 else 4
    u = np.ones((len(fk), 6)); u[:, 0:2] = fk[:, 0:2]; u[:, 3:5] = fk[:, 7:9]
    mask = (oracle.score(which, u, Fm) <= cfg.err_threshold ** 2).astype(np.uint8)
    k, res, ver, last = hv(f, keys, cfg, Fm, mask)
    assert tuple(last) == (0 if errorType == 0 else 1, cfg.max_samples, 0)      # inlLimit 0, as LORANSACFiltering passes it
    assert res.ransac_inliers == int(mask.sum()) and np.array_equal(np.array(res.H), Fm)
    cand = fk[mask == 1]
    e = oracle.score(which, laf_points(cand), Fm).reshape(-1, 3)
    good = ~(np.sqrt(e).sum(1) > cfg.LAFCoef * cfg.err_threshold)
    exp = cand[good]
    assert 50 < len(exp) < len(cand) and k == len(exp)
    assert np.array_equal(ver, np.c_[exp[:, 0:2], exp[:, 7:9]])


def test_too_few_tentatives_and_failed_checks(hv):
    f, keys, Hgt = make_frames(n=40, n_dup=0)
    cfg = mb.PairConfig.default()
    k, res, _, _ = hv(f[:7], keys[:7], cfg, np.linalg.inv(Hgt).T.ravel(), np.ones(7, np.uint8))      # < MIN_POINTS: nothing runs
    assert k == 0 and res.ransac_inliers == 0
    k, res, _, _ = hv(f, keys, cfg, np.zeros(9), np.ones(40, np.uint8))                                # singular model: cleared (matching.cpp:925-930)
    assert k == 0


# ---- the same stage against the reference's own matching.cpp, compiled in place (oracle/_ref) ---------------------------------------
@pytest.mark.parametrize("seed,errorType,ties", [(3, 0, False), (4, 0, True), (5, 1, True), (6, 2, False)])
def test_verify_stage_vs_compiled_reference(hv, reference, seed, errorType, ties):
    """DuplicateFiltering + LORANSACFiltering (inv(H^T), NaiveHCheck, H_LAF_check) of the host mirror == matching.cpp:2983-3047, 806-980 run
    by the compiled reference on the same tentatives: the RANSAC model / mask planted into the host mirror are the ones the reference's own
    exp_ransacHcustom returns for the seed, so every list (unique rows incl. the unstable std::sort's tie order, verified rows) must agree."""
    f, keys, Hgt = make_frames(n=600, n_dup=150, seed=seed)
    if ties:
        keys = np.round(keys * 12) / 12          # many equal ratios: the order of equal keys is std::sort's
    cfg = mb.PairConfig.default(); cfg.errorType = errorType; cfg.seed = seed
    kept = reference.duplicate_filter(f, keys, r=cfg.duplicateDist)
    fk = f[kept]
    u = np.ones((len(fk), 6)); u[:, 0:2] = fk[:, 0:2]; u[:, 3:5] = fk[:, 7:9]
    rr = reference.exp_ransacH(u, th=cfg.err_threshold ** 2, conf=cfg.confidence, max_sam=cfg.max_samples, errorType=errorType,
                               doSymCheck=cfg.doSymmCheck, seed=seed)
    exp = reference.loransac_filtering(fk, err_threshold=cfg.err_threshold, confidence=cfg.confidence, max_samples=cfg.max_samples,
                                       HLAFCoef=cfg.HLAFCoef, LAFCoef=cfg.LAFCoef, errorType=errorType, doSymmCheck=cfg.doSymmCheck, seed=seed)
    assert np.array_equal(exp["inl"], rr["inl"])                      # LORANSACFiltering ran the same RANSAC
    k, res, ver, _ = hv(f, keys, cfg, rr["H"], rr["inl"])
    assert res.unique_tentatives == len(kept) and res.ransac_inliers == int(rr["inl"].sum())
    rows = fk[exp["verified"]]
    assert 50 < len(rows) <= int(rr["inl"].sum()) and k == len(rows)
    assert np.array_equal(ver, np.c_[rows[:, 0:2], rows[:, 7:9]])
    assert np.allclose(np.array(res.H), exp["H"], rtol=1e-12, atol=0)


def test_duplicate_filter_tie_order_vs_compiled_reference(hv, reference):
    """All keys equal: the surviving rows are decided by std::sort's permutation of equal elements alone."""
    f, keys, Hgt = make_frames(n=500, n_dup=200, seed=9)
    keys = np.full(len(f), 0.5)
    cfg = mb.PairConfig.default()
    kept = reference.duplicate_filter(f, keys, r=cfg.duplicateDist)
    assert not np.array_equal(kept, np.sort(kept))                    # the reference's order is NOT the stable one
    k, res, ver, _ = hv(f, keys, cfg, np.linalg.inv(Hgt).T.ravel(), np.ones(len(kept), np.uint8))
    assert res.unique_tentatives == len(kept)
    exp = reference.loransac_filtering(f[kept], err_threshold=cfg.err_threshold, HLAFCoef=cfg.HLAFCoef, seed=1)
    # planted all-inlier mask: compare only the order-sensitive part -- the verified rows are a subsequence of f[kept] in the reference's order
    pos = {tuple(r): i for i, r in enumerate(np.c_[f[kept][:, 0:2], f[kept][:, 7:9]].tolist())}
    idx = [pos[tuple(r)] for r in ver.tolist()]
    assert idx == sorted(idx) and len(idx) > 50
