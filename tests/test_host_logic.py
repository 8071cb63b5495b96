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


def _set_vs_pars(H, scales, tilts, phi, prev):
    scales = np.asarray(scales, np.float64); tilts = np.asarray(tilts, np.float64)
    prev = np.ascontiguousarray(np.asarray(prev, np.float64).reshape(-1, 3))
    out = np.zeros((512, 3))
    n = H.mb2_host_set_vs_pars(scales.ctypes.data_as(C.c_void_p), C.c_int(len(scales)), tilts.ctypes.data_as(C.c_void_p), C.c_int(len(tilts)),
                               C.c_double(phi), prev.ctypes.data_as(C.c_void_p), C.c_int(len(prev)), out.ctypes.data_as(C.c_void_p), C.c_int(512))
    return out[:n].copy()


def test_set_vs_pars_view_counts_of_iters_mods_cviu():
    """SetVSPars (synth-detection.cpp:103-234) on the [MSER2], [MSER3], [HessianAffine4..6] tiers of build/iters_mods_cviu.ini:
    3, 24, 11, 20, 30 new views per step after the de-duplication against earlier steps."""
    mb.build()
    H = mb.host_lib()
    m2 = _set_vs_pars(H, [1, 0.25, 0.125], [1], 360, [])
    m3 = _set_vs_pars(H, [1, 0.25, 0.125], [1, 3, 6, 9], 360, m2)
    assert (len(m2), len(m3)) == (3, 24)
    h4 = _set_vs_pars(H, [1], [1, 2, 4, 6, 8], 360, [])
    h5 = _set_vs_pars(H, [1], [1, 2, 4, 6, 8], 120, h4)
    h6 = _set_vs_pars(H, [1], [1, 2, 4, 6, 8], 60, np.concatenate([h4, h5]))
    assert (len(h4), len(h5), len(h6)) == (11, 20, 30)
    # rows are (zoom, tilt, phi): tilt t gets floor(180 t / Phi) rotations spaced pi / n apart, the un-tilted view none
    assert np.allclose(h4[0], (1, 1, 0)) and np.allclose(h4[1], (1, 2, 0)) and np.allclose(h4[2:4, 2], (0, np.pi / 2))
    v = _set_vs_pars(H, [1], [-3], 360, [])   # negative tilt in the set: floor(180 * -3 / 360) < 0 -> "no rotation" mode, two views (:144-169)
    assert len(v) == 2 and np.allclose(v[0], (1, 3, 0)) and np.allclose(v[1], (1, -3, 0))


def test_feature_cache_format_and_round_trip(tmp_path):
    """ImageRepresentation::SaveRegions / LoadRegions (imagerepresentation.cpp:2139-2215): the reference's text format, record by
    record (saveAR :89-99 = ids, det_kp, reproj_kp as `x y a11 a12 a21 a22 pyramid_scale octave_number s sub_type`, descriptor size,
    descriptor values; C++ default stream formatting = %g with 6 significant digits), and the values survive a round trip."""
    mb.build()
    H = mb.host_lib()
    rng = np.random.default_rng(3)
    n = 7
    det = np.zeros((n, 9)); rep = np.zeros((n, 9))
    for a in (det, rep):
        a[:, 0:2] = rng.random((n, 2)) * 4000; a[:, 2:6] = rng.normal(size=(n, 4)); a[:, 6] = rng.random(n) * 30 + 1
        a[:, 7] = rng.normal(size=n) * 100; a[:, 8] = rng.integers(0, 22, n)
    desc = rng.integers(0, 256, (n, 128)).astype(np.uint8)
    f = str(tmp_path / "regions.txt").encode()
    assert H.mb2_host_save_regions(f, b"MSER", b"RootSIFT", C.c_int(n), det.ctypes.data_as(C.c_void_p), rep.ctypes.data_as(C.c_void_p),
                                   desc.ctypes.data_as(C.c_void_p)) == n
    lines = open(f.decode()).read().split("\n")
    assert lines[0] == "1" and lines[1] == "MSER 1" and lines[2] == "RootSIFT %d" % n and lines[3] == "128"

    def kp(v):
        return " ".join("%g" % x for x in v[:6]) + " 0 0 " + "%g" % v[6] + " %d " % int(v[8])
    for i in range(n):
        want = "%d 0 0 0 " % i + kp(det[i]) + kp(rep[i]) + " 128 " + " ".join("%d" % d for d in desc[i]) + " "
        assert lines[4 + i] == want, i
    d2 = np.zeros((n, 9)); r2 = np.zeros((n, 9)); u2 = np.zeros((n, 128), np.uint8)
    assert H.mb2_host_load_regions(f, b"MSER", b"RootSIFT", C.c_int(n), d2.ctypes.data_as(C.c_void_p), r2.ctypes.data_as(C.c_void_p),
                                   u2.ctypes.data_as(C.c_void_p)) == n
    assert np.array_equal(u2, desc)
    for got, src in ((d2, det), (r2, rep)):
        assert np.allclose(got[:, :7], src[:, :7], rtol=1e-5) and np.array_equal(got[:, 8], src[:, 8]) and np.all(got[:, 7] == 0)   # response is not in the format
    # a file as the reference writes it: a "None" list without descriptors next to the described one
    txt = "1\nHessianAffine 2\nNone 1\n0\n0 0 0 0 1 2 1 0 0 1 0 0 3 1 1 2 1 0 0 1 0 0 3 1  0 \nRootSIFT 1\n128\n" + lines[4] + "\n"
    g = str(tmp_path / "ref_style.txt")
    open(g, "w").write(txt)
    assert H.mb2_host_load_regions(g.encode(), b"HessianAffine", b"RootSIFT", C.c_int(n), d2.ctypes.data_as(C.c_void_p), r2.ctypes.data_as(C.c_void_p),
                                   u2.ctypes.data_as(C.c_void_p)) == 1
    assert np.array_equal(u2[0], desc[0])
