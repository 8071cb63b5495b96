"""Golden vectors of the reference's F-matrix LO-RANSAC (degensac/exp_ranF.c:795, DEGENSAC branch on) with a fixed seed, produced by
the reference's own sources compiled in place (oracle/_ref).  They pin the behaviour the F driver (not built yet)
has to reproduce.  Run in the build container only:  python tests/golden/make_golden_f.py"""
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


def main():
    from oracle.pyoracle import Reference
    R = Reference()
    u = two_view_scene()
    out = {"u": u}
    for seed in (1, 2):
        for et in (0, 1):
            r = R.exp_ransacF(u, seed=seed, errorType=et)
            out["F_s%d_e%d" % (seed, et)] = r["F"]; out["inl_s%d_e%d" % (seed, et)] = r["inl"]
            out["stats_s%d_e%d" % (seed, et)] = np.array([r["I"], r["samples"], r["lo"], r["Ih"]])
    np.savez_compressed(os.path.join(HERE, "ransac_f_vectors.npz"), **out)
    print({k: (v.tolist() if k.startswith("stats") else v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
