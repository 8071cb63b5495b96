"""Host-side logic of the C++ mirror that needs no GPU."""
import ctypes as C

import numpy as np

import mods_b200 as mb


def _brute(xy, ratio, r):
    """matching.cpp:2983-3047 verbatim: (stable) sort by ratio, O(T^2) sweep."""
    order = np.argsort(np.abs(ratio), kind="stable")
    P = xy[order]
    keep = np.ones(len(P), bool)
    for i in range(len(P)):
        if not keep[i]:
            continue
        d1 = ((P[i + 1:, 0] - P[i, 0]) ** 2 + (P[i + 1:, 1] - P[i, 1]) ** 2) <= r * r
        d2 = ((P[i + 1:, 2] - P[i, 2]) ** 2 + (P[i + 1:, 3] - P[i, 3]) ** 2) <= r * r
        keep[i + 1:] &= ~(d1 & d2)
    return order[keep]


def test_duplicate_filter_equals_reference_double_loop():
    mb.build()
    H = mb.host_lib()
    rng = np.random.default_rng(0)
    for n, span in ((0, 10), (1, 10), (500, 40), (3000, 200), (3000, 4000)):
        xy = np.zeros((n, 4))
        xy[:, :2] = rng.uniform(0, span, (n, 2)); xy[:, 2:] = xy[:, :2] + rng.normal(0, 1.5, (n, 2))
        ratio = rng.uniform(0, 1, n)
        if n > 10:
            ratio[5] = ratio[6]  # tie in the sort key: input order must be kept
        out = np.zeros(max(1, n), np.int32)
        k = H.mb2_host_duplicate_filter(xy.ctypes.data_as(C.c_void_p), ratio.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_double(2.0),
                                        C.c_int(1), out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out[:k], _brute(xy, ratio, 2.0))
