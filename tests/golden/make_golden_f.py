"""Golden vectors of the reference's F-matrix LO-RANSAC (degensac/exp_ranF.c:795, DEGENSAC branch on) with fixed seeds, produced by
the reference's own sources compiled in place (oracle/_ref, lapack_int pinned to 32 bits -- see oracle/ref_preinclude_c.h).  They pin
the behaviour of the F driver (mods_b200/csrc/ransac_f_logic.hpp).  Scenes: "deg" = a shallow scene whose best
samples are H-degenerate (checksample -> innerH -> plane-and-parallax rFtH decide), "gen" = general 3-D scene with 50 % outliers
(local optimisation runs), "plane" = 50 % of the inliers on one plane, "small" = 60 tentatives.  lim: "n" = inlLimit len (no
limit), "0" = inlLimit 0, which is what LORANSACFiltering passes (matching.cpp:883).
Run in the build container only:  python tests/golden/make_golden_f.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def two_view_scene(seed=5, n=400, n_out=100):
    rng = np.random.default_rng(seed)
    X = np.c_[rng.random((n, 2)) * 4 - 2, rng.random(n) * 4 + 4]
    K = np.array([[800, 0, 400], [0, 800, 300], [0, 0, 1.0]])
    Rm = np.array([[0.98, 0, 0.199], [0, 1, 0], [-0.199, 0, 0.98]]); t = np.array([0.5, 0.05, 0.1])
    x1 = (K @ X.T).T; x1 = x1[:, :2] / x1[:, 2:]
    x2 = (K @ (Rm @ X.T + t[:, None])).T; x2 = x2[:, :2] / x2[:, 2:]
    u = np.ones((n, 6)); u[:, 0:2] = x1 + rng.normal(size=(n, 2)) * 0.5; u[:, 3:5] = x2 + rng.normal(size=(n, 2)) * 0.5
    u[n - n_out:, 3:5] = rng.random((n_out, 2)) * 600
    return u


def general_scene(seed, n=400, n_out=200, noise=0.5, planar_frac=0.0, depth=(2, 12)):
    """Wide depth range, outliers shuffled among the inliers, optionally a fraction of the inliers on one plane."""
    rng = np.random.default_rng(seed)
    X = np.c_[rng.random((n, 2)) * 6 - 3, rng.random(n) * (depth[1] - depth[0]) + depth[0]]
    npl = int(planar_frac * n)
    X[:npl, 2] = 6 + 0.1 * X[:npl, 0]
    K = np.array([[800, 0, 400], [0, 800, 300], [0, 0, 1.0]])
    a = 0.15; Rm = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]); t = np.array([0.8, 0.1, 0.2])
    x1 = (K @ X.T).T; x1 = x1[:, :2] / x1[:, 2:]
    x2 = (K @ (Rm @ X.T + t[:, None])).T; x2 = x2[:, :2] / x2[:, 2:]
    u = np.ones((n, 6)); u[:, 0:2] = x1 + rng.normal(size=(n, 2)) * noise; u[:, 3:5] = x2 + rng.normal(size=(n, 2)) * noise
    u[n - n_out:, 3:5] = rng.random((n_out, 2)) * 800
    return np.ascontiguousarray(u[rng.permutation(n)])


SCENES = {"deg": lambda: two_view_scene(), "gen": lambda: general_scene(1), "plane": lambda: general_scene(2, n=500, n_out=350, planar_frac=0.5),
          "small": lambda: general_scene(1, n=60, n_out=30)}
CASES = [(name, seed, et, lim) for name in SCENES for seed in (1, 3) for et in (0, 1) for lim in ("n", "0")]


def main():
    from oracle.pyoracle import Reference
    R = Reference()
    out = {}
    for name, make in SCENES.items():
        out["u_" + name] = make()
    for name, seed, et, lim in CASES:
        r = R.exp_ransacF(out["u_" + name], seed=seed, errorType=et, inlLimit=None if lim == "n" else 0)
        key = "%s_s%d_e%d_l%s" % (name, seed, et, lim)
        out["F_" + key] = r["F"]; out["inl_" + key] = np.packbits(r["inl"])
        out["stats_" + key] = np.array([r["I"], r["samples"], r["lo"], r["Ih"]])
        print(key, out["stats_" + key].tolist())
    np.savez_compressed(os.path.join(HERE, "ransac_f_vectors.npz"), **out)


if __name__ == "__main__":
    main()
