"""Golden vectors for the Hamming matcher (MatchFLANNDistance, matching/matching.cpp:607-666) from the reference's own matching.cpp
compiled in place (oracle/_ref; cv::flann::Index answered by the shim's exact linear scan on the bit count, lower train index first).
ORB itself is outside the hot path: the descriptors are random bit strings with planted noisy copies, exact duplicates among the
trains (tie order) and several descriptor lengths.
Run in the build container only:  python tests/golden/make_golden_hamming.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def case(nq, nt, nbytes, seed, flip=0.08):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (nt, nbytes), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, nbytes), dtype=np.uint8)
    m = min(nq, nt) // 2
    noise = np.packbits(rng.random((m, nbytes * 8)) < flip, axis=1)
    q[:m] = t[rng.permutation(nt)[:m]] ^ noise          # noisy copies of trains
    t[nt // 2:nt // 2 + nt // 8] = t[:nt // 8]           # exact duplicates among the trains: equal distances
    if nq > m + 1:
        q[m] = t[0]                                      # distance 0 twice when t[0] has a duplicate: ratio = 0 / 0 (nan, as in the reference)
        q[m + 1] = t[nt - 1]                             # distance 0 once
    return q, t


CASES = {"orb32": (600, 700, 32, 1), "brisk64": (300, 500, 64, 2), "short16": (200, 300, 16, 3), "odd20": (150, 90, 20, 4), "two_trains": (40, 2, 32, 5)}
THRESHOLDS = (64.0, 30.5, 0.0)


def main():
    from oracle.pyoracle import Reference
    R = Reference()
    out = {}
    for name, (nq, nt, nb, seed) in CASES.items():
        q, t = case(nq, nt, nb, seed)
        out["q_" + name] = q; out["t_" + name] = t
        for th in THRESHOLDS:
            out["rows_%s_%g" % (name, th)] = R.match_hamming(q.astype(np.float32), t.astype(np.float32), th)
    np.savez_compressed(os.path.join(HERE, "hamming_vectors.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
