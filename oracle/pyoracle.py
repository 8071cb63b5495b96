"""TEST INFRASTRUCTURE: ctypes doors into the CPU restatement (oracle/libmods_oracle.so, prefix
``orc_``) and, when it has been built, the reference compiled in place
(oracle/_ref/libmods_ref.so, prefix ``ref_``).  Both export the same call shapes, so tests loop
over them.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KP = 9  # x y a11 a12 a21 a22 s response sub_type


class HessParams(C.Structure):
    """[HessianAffine] of config_iter_mods_cviu.ini (reference defaults: structures.hpp:146-165, affine.h:52-63)."""
    _fields_ = [("threshold", C.c_float), ("numberOfScales", C.c_int), ("initialSigma", C.c_float),
                ("edgeEigenValueRatio", C.c_float), ("border", C.c_int), ("maxIterations", C.c_int),
                ("convergenceThreshold", C.c_float), ("smmWindowSize", C.c_int), ("doBaumberg", C.c_int),
                ("mode", C.c_int), ("reg_number", C.c_int), ("rel_threshold", C.c_float),
                ("rel_reg_number", C.c_float), ("patchSize", C.c_int), ("mrSize", C.c_float), ("detectorType", C.c_int)]

    @staticmethod
    def default():
        return HessParams(5.3333, 3, 1.6, 10.0, 5, 16, 0.05, 19, 1, 0, 2000, -1.0, -1.0, 41, 3.0 * 3.0 ** 0.5, 0)

    @staticmethod
    def harris():
        """[HarrisAffine] of config_iter_mods_cviu.ini:28-44 (mode FixedTh, threshold 15, Baumberg with convergence threshold 0.1)."""
        return HessParams(15.0, 3, 1.6, 10.0, 5, 16, 0.1, 19, 1, 0, 1000, 0.1, 0.5, 41, 3.0 * 3.0 ** 0.5, 2)

    @staticmethod
    def dog():
        """[DoG] of config_iter_mods_cviu_wxbs.ini:45-59 with mode FixedTh: threshold 8, no Baumberg iteration."""
        return HessParams(8.0, 3, 1.6, 10.0, 5, 16, 0.05, 19, 0, 0, 3000, 0.01, 0.5, 41, 3.0 * 3.0 ** 0.5, 1)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Lib:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix

    def fn(self, name, restype=C.c_int):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    # ---- image helpers
    def gaussian_blur(self, img, sigma):
        img = _f32(img); out = np.empty_like(img)
        self.fn("gaussian_blur", None)(_p(img), _p(out), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_float(sigma))
        return out

    def hessian_response(self, img, norm):
        img = _f32(img); out = np.zeros_like(img)
        self.fn("hessian_response", None)(_p(img), _p(out), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_float(norm))
        return out

    def atan2(self, y, x):
        return self.fn("atan2LUTff", C.c_float)(C.c_float(y), C.c_float(x))

    def interpolate(self, img, ox, oy, a11, a12, a21, a22, ow, oh):
        img = _f32(img); out = np.zeros((oh, ow), np.float32)
        r = self.fn("interpolate")(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_float(ox), C.c_float(oy),
                                   C.c_float(a11), C.c_float(a12), C.c_float(a21), C.c_float(a22), _p(out), C.c_int(ow), C.c_int(oh))
        return out, r

    # ---- detectors
    def hessaff_detect(self, img, hp=None, raw=False, max_out=400000):
        img = _f32(img); hp = hp or HessParams.default()
        out = np.zeros((max_out, KP))
        n = self.fn("hessaff_detect")(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.byref(hp), C.c_int(int(raw)),
                                      _p(out), C.c_int(max_out))
        assert n <= max_out
        return out[:n].copy()

    def mser_detect(self, img, max_area=0.05, min_size=30, min_margin=8.0, mode=0, reg_number=500, raw=False, max_out=200000):
        img = _f32(img); out = np.zeros((max_out, KP))
        n = self.fn("mser_detect")(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_double(max_area), C.c_int(min_size),
                                   C.c_double(min_margin), C.c_int(mode), C.c_int(reg_number), C.c_int(int(raw)), _p(out), C.c_int(max_out))
        return out[:n].copy()

    def mser_regions(self, img, max_area=0.05, min_size=30, min_margin=8.0, max_out=400000):
        """rows: polarity minI maxI threshold margin area border nruns cx cy sxx sxy syy"""
        img = _f32(img); out = np.zeros((max_out, 13))
        n = self.fn("mser_regions")(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_double(max_area), C.c_int(min_size),
                                    C.c_double(min_margin), _p(out), C.c_int(max_out))
        assert n <= max_out
        return out[:n].copy()

    def detect_orientation(self, img, kps, mrSize=1.0, patchSize=41, maxAngles=1, th=0.8, doHalfSIFT=0):
        img = _f32(img); kps = _f64(kps); n = len(kps)
        out = np.zeros((max(1, n * max(1, maxAngles) * 2), KP))
        m = self.fn("detect_orientation_half")(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), _p(kps), C.c_int(n), C.c_double(mrSize),
                                               C.c_int(patchSize), C.c_int(maxAngles), C.c_double(th), C.c_int(doHalfSIFT), _p(out), C.c_int(len(out)))
        return out[:m].copy()

    def reproject(self, kps, H, w, h, which=0, mrSize=5.1962):
        kps = _f64(kps); H = _f64(H).ravel(); n = len(kps)
        od = np.zeros((max(1, n), KP)); orp = np.zeros((max(1, n), KP))
        m = self.fn("reproject")(_p(kps), C.c_int(n), _p(H), C.c_int(w), C.c_int(h), C.c_int(which), C.c_double(mrSize), _p(od), _p(orp))
        return od[:m].copy(), orp[:m].copy()

    def describe(self, img, kps, mrSize=5.1962, patchSize=41, fast=False, photoNorm=True, rootsift=True, want_patches=False):
        img = _f32(img); kps = _f64(kps); n = len(kps)
        desc = np.zeros((max(1, n), 128), np.float32)
        args = [_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), _p(kps), C.c_int(n), C.c_double(mrSize), C.c_int(patchSize),
                C.c_int(int(fast)), C.c_int(int(photoNorm)), C.c_int(int(rootsift)), _p(desc)]
        patches = None
        if self.prefix == "orc_":
            patches = np.zeros((max(1, n), patchSize, patchSize), np.float32) if want_patches else None
            args.append(_p(patches) if want_patches else C.c_void_p(0))
        self.fn("describe", None)(*args)
        return (desc[:n].copy(), patches[:n].copy()) if want_patches else desc[:n].copy()

    def describe_dsp(self, img, kps, mrSize=5.1962, patchSize=41, fast=False, photoNorm=True, numScales=3, startCoef=0.5, endCoef=1.5):
        """DSPSIFT (imagerepresentation.cpp:1547-1598; DomainSizePolingParams defaults siftdesc.h:23-29)."""
        img = _f32(img); kps = _f64(kps); n = len(kps)
        desc = np.zeros((max(1, n), 128), np.float32)
        self.fn("describe_dsp", None)(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), _p(kps), C.c_int(n), C.c_double(mrSize), C.c_int(patchSize),
                                      C.c_int(int(fast)), C.c_int(int(photoNorm)), C.c_int(numScales), C.c_double(startCoef), C.c_double(endCoef), _p(desc))
        return desc[:n].copy()

    def sift_patch(self, patch, rootsift=True):
        patch = _f32(patch); out = np.zeros(128, np.float32)
        self.fn("sift_patch", None)(_p(patch), C.c_int(int(rootsift)), _p(out))
        return out

    def view_pipeline(self, img, detector=0, hp=None, mser=(0.05, 30, 8.0), ori=(1.0, 41, 1, 0.8), desc=(5.1962, 41, True, True),
                      max_out=400000):
        img = _f32(img); hp = hp or HessParams.default()
        det = np.zeros((max_out, KP)); rep = np.zeros((max_out, KP)); d = np.zeros((max_out, 128), np.float32)
        n = self.fn("view_pipeline")(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(detector), C.byref(hp),
                                     C.c_double(mser[0]), C.c_int(mser[1]), C.c_double(mser[2]),
                                     C.c_double(ori[0]), C.c_int(ori[1]), C.c_int(ori[2]), C.c_double(ori[3]),
                                     C.c_double(desc[0]), C.c_int(desc[1]), C.c_int(int(desc[2])), C.c_int(int(desc[3])),
                                     _p(det), _p(rep), _p(d), C.c_int(max_out))
        assert 0 <= n <= max_out, n
        return det[:n].copy(), rep[:n].copy(), d[:n].copy()

    def synth_view(self, img, tilt, phi, zoom, InitSigma=0.5, doBlur=1):
        """GenerateSynthImageCorr: returns (pixels, H, is_identity)."""
        img = _f32(img); h, w = img.shape
        cap = int((w + h) ** 2 * max(1.0, zoom) ** 2) + 16
        out = np.zeros(cap, np.float32); ow, oh = C.c_int(), C.c_int(); H = np.zeros(9)
        ident = self.fn("synth_view")(_p(img), C.c_int(w), C.c_int(h), C.c_double(tilt), C.c_double(phi), C.c_double(zoom), C.c_double(InitSigma),
                                      C.c_int(doBlur), _p(out), C.c_int(cap), C.byref(ow), C.byref(oh), _p(H))
        return out[: ow.value * oh.value].reshape(oh.value, ow.value).copy(), H.reshape(3, 3), bool(ident)

    def view_pipeline_synth(self, img, tilt, phi, zoom, detector=0, hp=None, mser=(0.05, 30, 8.0), ori=(1.0, 41, 1, 0.8),
                            desc=(5.1962, 41, True, True), InitSigma=0.5, doBlur=1, max_out=400000):
        img = _f32(img); hp = hp or HessParams.default()
        det = np.zeros((max_out, KP)); rep = np.zeros((max_out, KP)); d = np.zeros((max_out, 128), np.float32)
        n = self.fn("view_pipeline_synth")(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(detector), C.byref(hp),
                                           C.c_double(mser[0]), C.c_int(mser[1]), C.c_double(mser[2]),
                                           C.c_double(ori[0]), C.c_int(ori[1]), C.c_int(ori[2]), C.c_double(ori[3]),
                                           C.c_double(desc[0]), C.c_int(desc[1]), C.c_int(int(desc[2])), C.c_int(int(desc[3])),
                                           C.c_double(tilt), C.c_double(phi), C.c_double(zoom), C.c_double(InitSigma), C.c_int(doBlur),
                                           _p(det), _p(rep), _p(d), C.c_int(max_out))
        assert 0 <= n <= max_out, n
        return det[:n].copy(), rep[:n].copy(), d[:n].copy()

    # ---- matching / verification
    def score(self, which, u, M):
        u = _f64(u); M = _f64(M).ravel(); n = len(u)
        d = np.zeros(n)
        self.fn("score", None)(C.c_int(which), _p(u), _p(M), _p(d), C.c_int(n))
        return d


class Oracle(Lib):
    def __init__(self):
        path = os.path.join(HERE, "libmods_oracle.so")
        if not os.path.exists(path):
            build()
        super().__init__(path, "orc_")

    def match_fginn(self, q, t, txy, ratio=0.8, contradDist=30.0, nn=50):
        q = _f32(q); t = _f32(t); txy = _f64(txy)
        out = np.zeros((max(1, len(q)), 7))
        n = self.fn("match_fginn")(_p(q), C.c_int(len(q)), _p(t), C.c_int(len(t)), _p(txy), C.c_double(ratio), C.c_double(contradDist),
                                   C.c_int(nn), _p(out), C.c_int(len(out)))
        return out[:n].copy()

    def match_hamming(self, q, t, max_distance=64.0):
        """MatchFLANNDistance (matching.cpp:607) restated: rows q idx0 idx1 d0 d1 ratio."""
        q = _f32(q); t = _f32(t)
        out = np.zeros((max(1, len(q)), 6))
        n = self.fn("match_hamming")(_p(q), C.c_int(len(q)), _p(t), C.c_int(len(t)), C.c_int(q.shape[1]), C.c_double(max_distance), _p(out),
                                     C.c_int(len(out)))
        return out[:n].copy()

    def resize_half(self, img):
        img = _f32(img)
        oh, ow = int(np.rint(img.shape[0] * 0.5)), int(np.rint(img.shape[1] * 0.5))
        out = np.zeros((oh, ow), np.float32)
        self.fn("resize_half", None)(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), _p(out))
        return out

    def pyramid(self, img, hp=None):
        """All (octave, level) blurred images + Hessian responses and the localized points."""
        img = _f32(img); hp = hp or HessParams.default()
        f = self.fn("pyramid_build", C.c_void_p)
        h = C.c_void_p(f(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.byref(hp)))
        levels = []
        for i in range(self.fn("pyramid_num_levels")(h)):
            o, l, r, c, s = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_float()
            self.fn("pyramid_level_info", None)(h, C.c_int(i), C.byref(o), C.byref(l), C.byref(r), C.byref(c), C.byref(s))
            blur = np.zeros((r.value, c.value), np.float32); resp = np.zeros_like(blur)
            self.fn("pyramid_level_data", None)(h, C.c_int(i), _p(blur), _p(resp))
            levels.append(dict(octave=o.value, level=l.value, sigma=s.value, blur=blur, resp=resp))
        e, lz = C.c_int(), C.c_int()
        nkeys = self.fn("pyramid_counts")(h, C.byref(e), C.byref(lz))
        loc = np.zeros((max(1, lz.value), 8), np.float32)
        self.fn("pyramid_localized", None)(h, _p(loc))
        self.fn("pyramid_free", None)(h)
        return dict(levels=levels, extrema=e.value, localized=loc[:lz.value].copy(), nkeys=nkeys)


class Reference(Lib):
    """The reference's own code (oracle/_ref).  Present in the build container and, as a prebuilt
    .so, on the GPU box; absent => tests that need it skip."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libmods_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _preload_blas_deps()
        super().__init__(path, "ref_")

    def exp_ransacH(self, u, th=9.0, conf=0.99, max_sam=100000, errorType=0, doSymCheck=1, seed=1):
        u = _f64(u); n = len(u)
        H = np.zeros(9); inl = np.zeros(n, np.uint8); out4 = np.zeros(4, np.int32)
        f = self.fn("exp_ransacH", C.c_double)
        J = f(_p(u), C.c_int(n), C.c_double(th), C.c_double(conf), C.c_int(max_sam), C.c_int(errorType), C.c_int(doSymCheck),
              C.c_long(seed), _p(H), _p(inl), _p(out4))
        return dict(H=H, inl=inl, I=int(out4[0]), samples=int(out4[1]), lo=int(out4[2]), rejected=int(out4[3]), J=J)

    def exp_ransacF(self, u, th=9.0, conf=0.99, max_sam=100000, errorType=0, doSymCheck=1, seed=1, inlLimit=None, do_lo=1):
        """Reference only (oracle/_ref): exp_ransacFcustom with a fixed seed.  inlLimit None = len (no limit); LORANSACFiltering passes 0."""
        u = _f64(u); n = len(u)
        F = np.zeros(9); inl = np.zeros(n, np.uint8); out4 = np.zeros(4, np.int32)
        I = self.fn("exp_ransacF3")(_p(u), C.c_int(n), C.c_double(th), C.c_double(conf), C.c_int(max_sam), C.c_int(errorType), C.c_int(doSymCheck),
                                    C.c_long(seed), C.c_uint(n if inlLimit is None else inlLimit), C.c_int(do_lo), _p(F), _p(inl), _p(out4))
        return dict(F=F, inl=inl, I=int(I), samples=int(out4[1]), lo=int(out4[2]), Ih=int(out4[3]))

    # ---- matching/matching.cpp compiled in place (FLANN answered by the shim's exact linear k-NN)
    def match_fginn(self, q, t, t_kps, ratio=0.8, contradDist=30.0, nn=50):
        """MatchFlannFGINN (matching.cpp:357).  t_kps: nt x 9 regions or nt x 2 centres.  Rows: q idx0 idxJ idx1 d0 dJ d1."""
        q = _f32(q); t = _f32(t); t_kps = _f64(t_kps)
        if t_kps.shape[1] == 2:
            k = np.zeros((len(t), KP)); k[:, :2] = t_kps; t_kps = k
        out = np.zeros((max(1, len(q)), 7))
        n = self.fn("match_fginn")(_p(q), C.c_int(len(q)), _p(t), C.c_int(len(t)), _p(t_kps), C.c_double(ratio), C.c_double(contradDist),
                                   C.c_int(nn), _p(out), C.c_int(len(out)))
        return out[:n].copy()

    def match_hamming(self, q, t, max_distance=64.0):
        """MatchFLANNDistance (matching.cpp:607), Hamming 2-NN.  Rows: q second.id d1 d2 ratio."""
        q = _f32(q); t = _f32(t)
        out = np.zeros((max(1, len(q)), 5))
        n = self.fn("match_hamming")(_p(q), C.c_int(len(q)), _p(t), C.c_int(len(t)), C.c_int(q.shape[1]), C.c_double(max_distance), _p(out),
                                     C.c_int(len(out)))
        return out[:n].copy()

    def duplicate_filter(self, frames14, ratio, r=2.0, mode=1):
        """DuplicateFiltering (matching.cpp:2983), MODE_FGINN = 1: surviving row numbers in output order."""
        frames14 = _f64(frames14); ratio = _f64(ratio); n = len(frames14)
        kept = np.zeros(max(1, n), np.int32)
        k = self.fn("duplicate_filter")(_p(frames14), _p(ratio), C.c_int(n), C.c_double(r), C.c_int(mode), _p(kept))
        return kept[:k].copy()

    def loransac_filtering(self, frames14, useF=0, err_threshold=3.0, confidence=0.99, max_samples=100000, localOptimization=1, LAFCoef=3.0,
                           HLAFCoef=10.0, errorType=0, doSymmCheck=1, seed=1):
        """LORANSACFiltering (matching.cpp:806) incl. NaiveHCheck / H_LAF_check / F_LAF_check.  errorType: 0 Sampson, 1 SymmMax, 2 SymmSum."""
        frames14 = _f64(frames14); n = len(frames14)
        ver = np.zeros(max(1, n), np.int32); inl = np.zeros(max(1, n), np.uint8); H = np.zeros(9)
        k = self.fn("loransac_filtering")(_p(frames14), C.c_int(n), C.c_int(useF), C.c_double(err_threshold), C.c_double(confidence), C.c_int(max_samples),
                                          C.c_int(localOptimization), C.c_double(LAFCoef), C.c_double(HLAFCoef), C.c_int(errorType), C.c_int(doSymmCheck),
                                          C.c_long(seed), _p(ver), _p(inl), _p(H))
        return dict(verified=ver[:k].copy(), inl=inl[:n].copy(), H=H)

    def pair_back(self, groups, contradDist=30.0, nn=50, duplicateDist=2.0, useF=0, err_threshold=3.0, confidence=0.99, max_samples=100000,
                  localOptimization=1, LAFCoef=3.0, HLAFCoef=10.0, errorType=0, doSymmCheck=1, seed=1):
        """mods.cpp:290-351 on (q_kps, q_desc, t_kps, t_desc, ratio) groups: match every group, append, duplicate-filter, LO-RANSAC + checks.
        Returns dict(tent = rows (group q idx0 idxJ idx1 d0 dJ d1), kept, verified (rows of tent), H, counts)."""
        G = len(groups)
        keep = []
        PD = C.c_void_p * G; I = C.c_int * G
        qk, qd, tk, td = PD(), PD(), PD(), PD(); nq, nt = I(), I(); ratio = (C.c_double * G)()
        for g, (a, b, c, d, r) in enumerate(groups):
            a = _f64(a); b = _f32(b); c = _f64(c); d = _f32(d); keep += [a, b, c, d]
            qk[g] = a.ctypes.data; qd[g] = b.ctypes.data; tk[g] = c.ctypes.data; td[g] = d.ctypes.data; nq[g] = len(a); nt[g] = len(c); ratio[g] = r
        cap = max(1, sum(len(g[0]) for g in groups))
        tent = np.zeros((cap, 8)); kept = np.zeros(cap, np.int32); ver = np.zeros(cap, np.int32); H = np.zeros(9); counts = np.zeros(4, np.int32)
        self.fn("pair_back")(C.c_int(G), nq, qk, qd, nt, tk, td, ratio, C.c_double(contradDist), C.c_int(nn), C.c_double(duplicateDist), C.c_int(useF),
                             C.c_double(err_threshold), C.c_double(confidence), C.c_int(max_samples), C.c_int(localOptimization), C.c_double(LAFCoef),
                             C.c_double(HLAFCoef), C.c_int(errorType), C.c_int(doSymmCheck), C.c_long(seed), _p(tent), C.c_int(cap), _p(kept), _p(ver),
                             _p(H), _p(counts))
        return dict(tent=tent[:counts[0]].copy(), kept=kept[:counts[1]].copy(), verified=ver[:counts[3]].copy(), H=H, counts=counts.tolist())

    def u2h(self, u, idx):
        u = _f64(u); idx = np.ascontiguousarray(idx, np.int32); H = np.zeros(9)
        self.fn("u2h", None)(_p(u), _p(idx), C.c_int(len(idx)), _p(H))
        return H


def _preload_blas_deps():
    """The reference links LAPACK from the OpenBLAS bundled with opencv-python-headless; that
    library needs its sibling libgfortran / libquadmath, which are not on the loader path."""
    import glob
    import site
    for sp in site.getsitepackages():
        d = os.path.join(sp, "opencv_python_headless.libs")
        for pat in ("libquadmath*.so*", "libgfortran*.so*", "libopenblas*.so*"):
            for f in sorted(glob.glob(os.path.join(d, pat))):
                try:
                    C.CDLL(f, mode=C.RTLD_GLOBAL)
                except OSError:
                    pass


def have_reference():
    return os.path.exists(os.path.join(HERE, "_ref", "libmods_ref.so"))


def build(with_ref=True):
    """Compile the restatement and (when /root/reference is present) the reference."""
    subprocess.check_call(["make", "-C", HERE, "libmods_oracle.so"], stdout=subprocess.DEVNULL)
    if with_ref:
        subprocess.check_call([os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
