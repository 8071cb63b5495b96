"""mods_b200: B200-native (sm_100a) implementation of the MODS per-view feature pipeline.

The product is ``libmods_b200.so`` (hand-written CUDA behind the C ABI in ``include/mods_b200.h``).
This module is only the ctypes door used by the tests, ``bench.py`` and ``__graft_entry__``; it
never falls back to a CPU path: if the library or a GPU is missing, calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmods_b200.so")
HOST_LIB_PATH = os.path.join(HERE, "libmods_host.so")
KP = 9
EXPORTS = [
    "mb2_ctx_create", "mb2_ctx_destroy", "mb2_last_error", "mb2_ctx_sync", "mb2_ctx_stream", "mb2_ctx_launch_count",
    "mb2_hessaff_detect", "mb2_detect_orientation", "mb2_describe_sift", "mb2_detect_describe_view", "mb2_view_fetch",
    "mb2_match_fginn", "mb2_match_hamming", "mb2_match_slots", "mb2_score_models", "mb2_ransac_h", "mb2_ransac_f", "mb2_debug_pyramid_level", "mb2_ctx_profile_begin", "mb2_ctx_profile_end", "mb2_ctx_device", "mb2_ctx_profiling", "mb2_slot_move",
    "mb2_mser_detect", "mb2_mser_regions", "mb2_detect_describe_view_mser", "mb2_mser_detect_pair", "mb2_describe_view_of_pair", "mb2_synth_view", "mb2_detect_describe_synth_view", "mb2_ctx_tree_epoch", "mb2_ctx_wait_tree", "mb2_ctx_create_prio",
    "mb2_view_pack", "mb2_slot_from_records", "mb2_match_slots_range", "mb2_is_device_pointer", "mb2_dev_alloc", "mb2_dev_free", "mb2_dev_copy", "mb2_ctx_make_current", "mb2_debug_fp64_peak", "mb2_records_gather_frames", "mb2_records_checksum",
]


class Mb2Error(RuntimeError):
    pass


class HessaffParams(C.Structure):
    """mb2_hessaff_params == [HessianAffine] of config_iter_mods_cviu.ini."""
    _fields_ = [("threshold", C.c_float), ("numberOfScales", C.c_int), ("initialSigma", C.c_float),
                ("edgeEigenValueRatio", C.c_float), ("border", C.c_int), ("maxIterations", C.c_int),
                ("convergenceThreshold", C.c_float), ("smmWindowSize", C.c_int), ("doBaumberg", C.c_int),
                ("mode", C.c_int), ("reg_number", C.c_int), ("rel_threshold", C.c_float),
                ("rel_reg_number", C.c_float), ("patchSize", C.c_int), ("mrSize", C.c_float), ("detectorType", C.c_int)]

    @staticmethod
    def default():
        return HessaffParams(5.3333, 3, 1.6, 10.0, 5, 16, 0.05, 19, 1, 0, 2000, -1.0, -1.0, 41, 3.0 * 3.0 ** 0.5, 0)

    @staticmethod
    def harris():
        """[HarrisAffine] of config_iter_mods_cviu.ini:28-44 (mode FixedTh, threshold 15, Baumberg with convergence threshold 0.1)."""
        return HessaffParams(15.0, 3, 1.6, 10.0, 5, 16, 0.1, 19, 1, 0, 1000, 0.1, 0.5, 41, 3.0 * 3.0 ** 0.5, 2)

    @staticmethod
    def dog():
        """[DoG] of config_iter_mods_cviu_wxbs.ini:45-59 with mode FixedTh: threshold 8, no Baumberg iteration."""
        return HessaffParams(8.0, 3, 1.6, 10.0, 5, 16, 0.05, 19, 0, 0, 3000, 0.01, 0.5, 41, 3.0 * 3.0 ** 0.5, 1)


class MserParams(C.Structure):
    """mb2_mser_params == [MSER] of config_iter_mods_cviu.ini (extrema::ExtremaParams)."""
    _fields_ = [("max_area", C.c_double), ("min_size", C.c_int), ("min_margin", C.c_double), ("relative", C.c_int),
                ("mode", C.c_int), ("reg_number", C.c_int), ("rel_threshold", C.c_float), ("rel_reg_number", C.c_float)]

    @staticmethod
    def default():
        return MserParams(0.05, 30, 8.0, 0, 0, -1, -1.0, -1.0)


class ViewParams(C.Structure):
    """mb2_view_params == one ViewSynthParameters entry (tilt, rotation phi [rad], zoom, InitSigma, doBlur)."""
    _fields_ = [("tilt", C.c_double), ("phi", C.c_double), ("zoom", C.c_double), ("InitSigma", C.c_double), ("doBlur", C.c_int)]


class OrientationParams(C.Structure):
    _fields_ = [("mrSize", C.c_double), ("patchSize", C.c_int), ("maxAngles", C.c_int), ("threshold", C.c_double),
                ("doHalfSIFT", C.c_int), ("reserved", C.c_int)]

    @staticmethod
    def default():
        return OrientationParams(1.0, 41, 1, 0.8, 0, 0)


class SiftParams(C.Structure):
    _fields_ = [("mrSize", C.c_double), ("patchSize", C.c_int), ("photoNorm", C.c_int), ("rootSIFT", C.c_int),
                ("fastPatchExtraction", C.c_int), ("doHalfSIFT", C.c_int), ("dspScales", C.c_int), ("dspStartCoef", C.c_double), ("dspEndCoef", C.c_double)]

    @staticmethod
    def default():
        return SiftParams(5.1962, 41, 1, 1, 0, 0, 0, 0.5, 1.5)

    @staticmethod
    def dspsift(numScales=3, startCoef=0.5, endCoef=1.5):
        """DSPSIFT (imagerepresentation.cpp:1547-1598) with the DomainSizePolingParams defaults (siftdesc.h:19-30)."""
        return SiftParams(5.1962, 41, 1, 0, 0, 0, numScales, startCoef, endCoef)


def build(force=False):
    """Compile libmods_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.check_call(["make", "-C", HERE, "all"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mb2Error("libmods_b200.so is not built (run python -c 'import __graft_entry__ as g; g.build()'); "
                           "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.mb2_last_error.restype = C.c_char_p
        _lib.mb2_ctx_stream.restype = C.c_void_p
        _lib.mb2_ctx_launch_count.restype = C.c_longlong
    return _lib


_host = None


def host_lib():
    """libmods_host.so: the C++ mirror of the reference's plugin surface (mods_b200/host)."""
    global _host
    if _host is None:
        lib()
        if not os.path.exists(HOST_LIB_PATH):
            raise Mb2Error("libmods_host.so is not built")
        _host = C.CDLL(HOST_LIB_PATH)
        _host.mb2_mods_launch_count.restype = C.c_longlong
        _host.mb2_mods_launch_count.argtypes = [C.c_void_p]
        _host.mb2_mods_release.argtypes = [C.c_void_p]
        _host.mb2_mods_release.restype = None
        _host.mb2_dist_comm_destroy.restype = None
        _host.mb2_shard_view_cost.restype = C.c_double
    return _host


class PairConfig(C.Structure):
    _fields_ = [("det", HessaffParams), ("ori", OrientationParams), ("desc", SiftParams),
                ("matchRatio", C.c_double), ("contradDist", C.c_double), ("duplicateDist", C.c_double),
                ("err_threshold", C.c_double), ("confidence", C.c_double), ("HLAFCoef", C.c_double),
                ("max_samples", C.c_int), ("errorType", C.c_int), ("doSymmCheck", C.c_int), ("seed", C.c_long),
                ("use_mser", C.c_int), ("mser", MserParams), ("mserMatchRatio", C.c_double),
                ("n_hess_views", C.c_int), ("n_mser_views", C.c_int), ("hess_views", ViewParams * 64), ("mser_views", ViewParams * 64),
                ("useF", C.c_int), ("localOptimization", C.c_int), ("LAFCoef", C.c_double),
                ("halfRootSIFT", C.c_int), ("reserved", C.c_int)]

    def set_views(self, hess=None, mser=None):
        """View tiers of the step: lists of (tilt, phi, zoom[, InitSigma]) as SetVSPars produces them; None = identity view only."""
        for name, views in (("hess", hess), ("mser", mser)):
            views = views or []
            assert len(views) <= 64
            setattr(self, "n_%s_views" % name, len(views))
            arr = getattr(self, "%s_views" % name)
            for i, v in enumerate(views):
                arr[i] = ViewParams(v[0], v[1], v[2], v[3] if len(v) > 3 else 0.5, 1)

    @staticmethod
    def default():
        c = PairConfig()
        host_lib().mb2_pair_config_default(C.byref(c))
        return c


def set_vs_pars(scales, tilts, phi, prev=()):
    """SetVSPars (synth-detection.cpp:103-234) through the host mirror: rows (zoom, tilt, phi) of one step, de-duplicated against `prev`."""
    scales = np.asarray(scales, np.float64); tilts = np.asarray(tilts, np.float64)
    prev = np.ascontiguousarray(np.asarray(prev, np.float64).reshape(-1, 3)); out = np.zeros((512, 3))
    n = host_lib().mb2_host_set_vs_pars(scales.ctypes.data_as(C.c_void_p), C.c_int(len(scales)), tilts.ctypes.data_as(C.c_void_p), C.c_int(len(tilts)),
                                        C.c_double(phi), prev.ctypes.data_as(C.c_void_p), C.c_int(len(prev)), out.ctypes.data_as(C.c_void_p), C.c_int(512))
    return out[:n].copy()


def iters_mods_cviu_views(which="c4"):
    """View tiers of build/iters_mods_cviu.ini as (tilt, phi, zoom, InitSigma) lists for PairConfig.set_views: "small" = [HessianAffine4]
    (11 views) + [MSER2] (3); "c4" = every SIFT tier, [HessianAffine4..6] (11 + 20 + 30) and [MSER2..3] (3 + 24) -- BASELINE config C4."""
    m2 = set_vs_pars([1, 0.25, 0.125], [1], 360)
    h4 = set_vs_pars([1], [1, 2, 4, 6, 8], 360)
    if which == "small":
        hess, mser = h4, m2
    else:
        m3 = set_vs_pars([1, 0.25, 0.125], [1, 3, 6, 9], 360, m2)
        h5 = set_vs_pars([1], [1, 2, 4, 6, 8], 120, h4); h6 = set_vs_pars([1], [1, 2, 4, 6, 8], 60, np.concatenate([h4, h5]))
        hess, mser = np.concatenate([h4, h5, h6]), np.concatenate([m2, m3])
    return [(r[1], r[2], r[0], 0.2) for r in hess], [(r[1], r[2], r[0], 0.8) for r in mser]


class PairResult(C.Structure):
    _fields_ = [("regions1", C.c_int), ("regions2", C.c_int), ("tentatives", C.c_int), ("unique_tentatives", C.c_int),
                ("ransac_inliers", C.c_int), ("verified", C.c_int), ("H", C.c_double * 9),
                ("ms_detect_describe", C.c_double), ("ms_match", C.c_double), ("ms_duplicate", C.c_double),
                ("ms_ransac", C.c_double), ("ms_total", C.c_double),
                ("mser_regions1", C.c_int), ("mser_regions2", C.c_int), ("mser_tentatives", C.c_int)]


def _ptr(a):
    """numpy array -> host pointer; torch CUDA tensor / int -> device pointer."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.data_ptr())  # torch tensor


class Context:
    """One mb2_ctx (one GPU, one stream)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        rc = lib().mb2_ctx_create(C.c_int(device), C.byref(self.h))
        if rc != 0:
            raise Mb2Error("mb2_ctx_create(device=%d) failed with %d: no usable CUDA device (no CPU fallback)" % (device, rc))
        self.device = device

    def close(self):
        if self.h:
            if _host is not None:
                _host.mb2_mods_release(self.h)
            lib().mb2_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise Mb2Error("%s failed (%d): %s" % (what, rc, lib().mb2_last_error(self.h).decode()))
        return rc

    def sync(self):
        self._check(lib().mb2_ctx_sync(self.h), "sync")

    @property
    def stream(self):
        return lib().mb2_ctx_stream(self.h)

    @property
    def launches(self):
        """Kernels launched so far (including the helper context mods_pair keeps for the second image)."""
        if _host is not None:
            return _host.mb2_mods_launch_count(self.h)
        return lib().mb2_ctx_launch_count(self.h)

    def fp64_peak(self):
        """Measured FP64 FMA throughput (TFLOP/s) of this device."""
        v = C.c_double()
        self._check(lib().mb2_debug_fp64_peak(self.h, C.byref(v)), "fp64_peak")
        return v.value

    def profile_begin(self):
        self._check(lib().mb2_ctx_profile_begin(self.h), "profile_begin")

    def profile_end(self):
        buf = C.create_string_buffer(1 << 16)
        self._check(lib().mb2_ctx_profile_end(self.h, buf, C.c_int(len(buf))), "profile_end")
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split("\t")
            out[name] = (int(n), float(ms))
        return out

    def pyramid_level(self, octave, level, want_resp=False):
        r, c = C.c_int(), C.c_int()
        n_oct = self._check(lib().mb2_debug_pyramid_level(self.h, C.c_int(octave), C.c_int(level), C.c_int(int(want_resp)),
                                                          C.c_void_p(0), C.byref(r), C.byref(c)), "pyramid_level")
        out = np.zeros((r.value, c.value), np.float32)
        self._check(lib().mb2_debug_pyramid_level(self.h, C.c_int(octave), C.c_int(level), C.c_int(int(want_resp)), _ptr(out),
                                                  C.byref(r), C.byref(c)), "pyramid_level")
        return out, n_oct

    # ---- detection
    def hessaff_detect(self, img, par=None, as_regions=True, capacity=None, shape=None):
        h, w = shape if shape is not None else img.shape
        par = par or HessaffParams.default()
        capacity = capacity or max(4096, (h * w) // 16)
        out = np.zeros((capacity, KP))
        n = self._check(lib().mb2_hessaff_detect(self.h, _ptr(img), C.c_int(w), C.c_int(h), C.byref(par), C.c_double(1.0),
                                                 C.c_double(1.0), C.c_int(int(as_regions)), _ptr(out), C.c_int(capacity)), "hessaff_detect")
        return out[:n].copy()

    def mser_detect(self, img, par=None, as_regions=True, capacity=None, shape=None, tilt=1.0, zoom=1.0):
        h, w = shape if shape is not None else img.shape
        par = par or MserParams.default()
        capacity = capacity or max(4096, (h * w) // 16)
        out = np.zeros((capacity, KP))
        n = self._check(lib().mb2_mser_detect(self.h, _ptr(img), C.c_int(w), C.c_int(h), C.byref(par), C.c_double(tilt),
                                              C.c_double(zoom), C.c_int(int(as_regions)), _ptr(out), C.c_int(capacity)), "mser_detect")
        return out[:n].copy()

    def mser_regions(self, img, par=None, capacity=None, shape=None):
        """rows: polarity minI maxI threshold margin area border nruns cx cy sxx sxy syy"""
        h, w = shape if shape is not None else img.shape
        par = par or MserParams.default()
        capacity = capacity or max(4096, (h * w) // 16)
        out = np.zeros((capacity, 13))
        n = self._check(lib().mb2_mser_regions(self.h, _ptr(img), C.c_int(w), C.c_int(h), C.byref(par), _ptr(out), C.c_int(capacity)),
                        "mser_regions")
        return out[:n].copy()

    def mser_pair_views(self, img1, img2, det=None, ori=None, desc=None, slots=(2, 3), shape=None):
        """Both images of a pair through MSER detection in one pass, then orientation / reprojection / description per image.
        Returns [(det_kp, reproj_kp, desc_u8), (...)]."""
        h, w = shape if shape is not None else img1.shape
        det = det or MserParams.default(); ori = ori or OrientationParams.default(); desc = desc or SiftParams.default()
        n1, n2 = C.c_int(), C.c_int()
        self._check(lib().mb2_mser_detect_pair(self.h, _ptr(img1), _ptr(img2), C.c_int(w), C.c_int(h), C.byref(det), C.byref(n1), C.byref(n2)),
                    "mser_detect_pair")
        H = np.ascontiguousarray(np.eye(3), np.float64)
        out = []
        for which, n in ((0, n1.value), (1, n2.value)):
            cap = max(1, n * max(1, ori.maxAngles))
            dk = np.zeros((cap, KP)); rk = np.zeros((cap, KP)); du = np.zeros((cap, 128), np.uint8)
            k = self._check(lib().mb2_describe_view_of_pair(self.h, C.c_void_p(0), C.c_int(which), _ptr(H), C.c_int(w), C.c_int(h), C.byref(ori), C.byref(desc),
                                                            C.c_int(slots[which]), C.c_int(0), _ptr(dk), _ptr(rk), _ptr(du), C.c_int(cap)),
                            "describe_view_of_pair")
            out.append((dk[:k].copy(), rk[:k].copy(), du[:k].copy()))
        return out

    def detect_orientation(self, img, kps, par=None, shape=None):
        h, w = shape if shape is not None else img.shape
        par = par or OrientationParams.default()
        kps = np.ascontiguousarray(kps, np.float64)
        cap = max(1, len(kps) * max(1, par.maxAngles))
        out = np.zeros((cap, KP))
        n = self._check(lib().mb2_detect_orientation(self.h, _ptr(img), C.c_int(w), C.c_int(h), _ptr(kps), C.c_int(len(kps)),
                                                     C.byref(par), _ptr(out), C.c_int(cap)), "detect_orientation")
        return out[:n].copy()

    def describe_sift(self, img, kps, par=None, want_patches=False, shape=None):
        h, w = shape if shape is not None else img.shape
        par = par or SiftParams.default()
        kps = np.ascontiguousarray(kps, np.float64)
        n = len(kps)
        desc = np.zeros((max(1, n), 128), np.uint8)
        patches = np.zeros((max(1, n), 41, 41), np.float32) if want_patches else None
        self._check(lib().mb2_describe_sift(self.h, _ptr(img), C.c_int(w), C.c_int(h), _ptr(kps), C.c_int(n), C.byref(par),
                                            _ptr(desc), _ptr(patches)), "describe_sift")
        return (desc[:n].copy(), patches[:n].copy()) if want_patches else desc[:n].copy()

    def detect_describe_view(self, img, H=None, orig_shape=None, det=None, ori=None, desc=None, slot=0, append=False,
                             capacity=None, want_host=True, shape=None):
        h, w = shape if shape is not None else img.shape
        oh, ow = orig_shape or (h, w)
        H = np.ascontiguousarray(np.eye(3) if H is None else H, np.float64)
        det = det or HessaffParams.default(); ori = ori or OrientationParams.default(); desc = desc or SiftParams.default()
        capacity = capacity or max(4096, (h * w) // 16)
        if want_host:
            dk = np.zeros((capacity, KP)); rk = np.zeros((capacity, KP)); du = np.zeros((capacity, 128), np.uint8)
        else:
            dk = rk = du = None
        fn = lib().mb2_detect_describe_view_mser if isinstance(det, MserParams) else lib().mb2_detect_describe_view
        n = self._check(fn(self.h, _ptr(img), C.c_int(w), C.c_int(h), _ptr(H), C.c_int(ow), C.c_int(oh),
                           C.byref(det), C.byref(ori), C.byref(desc), C.c_int(slot), C.c_int(int(append)),
                           _ptr(dk), _ptr(rk), _ptr(du), C.c_int(capacity)), "detect_describe_view")
        if want_host:
            return dk[:n].copy(), rk[:n].copy(), du[:n].copy()
        return n

    def synth_view(self, img, tilt, phi, zoom, InitSigma=0.5, doBlur=1, shape=None):
        """GenerateSynthImageCorr on the GPU: returns (pixels, H, is_identity)."""
        h, w = shape if shape is not None else img.shape
        vp = ViewParams(tilt, phi, zoom, InitSigma, doBlur)
        cap = int((w + h) ** 2 * max(1.0, zoom) ** 2) + 16
        out = np.zeros(cap, np.float32); ow, oh = C.c_int(), C.c_int(); H = np.zeros(9)
        ident = self._check(lib().mb2_synth_view(self.h, _ptr(img), C.c_int(w), C.c_int(h), C.byref(vp), _ptr(out), C.c_int(cap), C.byref(ow),
                                                 C.byref(oh), _ptr(H)), "synth_view")
        return out[: ow.value * oh.value].reshape(oh.value, ow.value).copy(), H.reshape(3, 3), bool(ident)

    def detect_describe_synth_view(self, img, tilt, phi, zoom, det=None, ori=None, desc=None, slot=0, append=False, capacity=None,
                                   InitSigma=0.5, doBlur=1, shape=None, want_host=True):
        h, w = shape if shape is not None else img.shape
        vp = ViewParams(tilt, phi, zoom, InitSigma, doBlur)
        det = det or HessaffParams.default(); ori = ori or OrientationParams.default(); desc = desc or SiftParams.default()
        is_mser = isinstance(det, MserParams)
        capacity = capacity or max(4096, (h * w) // 16)
        if want_host:
            dk = np.zeros((capacity, KP)); rk = np.zeros((capacity, KP)); du = np.zeros((capacity, 128), np.uint8)
        else:
            dk = rk = du = None
        n = self._check(lib().mb2_detect_describe_synth_view(self.h, _ptr(img), C.c_int(w), C.c_int(h), C.byref(vp), C.c_int(3 if is_mser else 0),
                                                             None if is_mser else C.byref(det), C.byref(det) if is_mser else None,
                                                             C.byref(ori), C.byref(desc), C.c_int(slot), C.c_int(int(append)),
                                                             _ptr(dk), _ptr(rk), _ptr(du), C.c_int(capacity)), "detect_describe_synth_view")
        if want_host:
            return dk[:n].copy(), rk[:n].copy(), du[:n].copy()
        return n

    # ---- matching
    def match_fginn(self, q_desc, t_desc, t_xy, ratio=0.8, contradDist=30.0, nn=50, nq=None, nt=None):
        nq = len(q_desc) if nq is None else nq
        nt = len(t_desc) if nt is None else nt
        out = np.zeros((max(1, nq), 7))
        n = self._check(lib().mb2_match_fginn(self.h, _ptr(q_desc), C.c_int(nq), _ptr(t_desc), C.c_int(nt), _ptr(t_xy),
                                              C.c_double(ratio), C.c_double(contradDist), C.c_int(nn), _ptr(out), C.c_int(len(out))),
                        "match_fginn")
        return out[:n].copy()

    def match_hamming(self, q_desc, t_desc, max_distance=64.0):
        """MatchFLANNDistance (matching.cpp:607-666): exact Hamming 2-NN of byte descriptors.  Rows: q idx0 idx1 idx1 d0 d1 d1."""
        q_desc = np.ascontiguousarray(q_desc, np.uint8); t_desc = np.ascontiguousarray(t_desc, np.uint8)
        nq, nt = len(q_desc), len(t_desc)
        nbytes = q_desc.shape[1] if q_desc.ndim == 2 else t_desc.shape[1]
        out = np.zeros((max(1, nq), 7))
        n = self._check(lib().mb2_match_hamming(self.h, _ptr(q_desc), C.c_int(nq), _ptr(t_desc), C.c_int(nt), C.c_int(nbytes),
                                                C.c_double(max_distance), _ptr(out), C.c_int(len(out))), "match_hamming")
        return out[:n].copy()

    def match_slots(self, q_slot, t_slot, ratio=0.8, contradDist=30.0, nn=50, capacity=1 << 20):
        out = np.zeros((capacity, 7))
        n = self._check(lib().mb2_match_slots(self.h, C.c_int(q_slot), C.c_int(t_slot), C.c_double(ratio), C.c_double(contradDist),
                                              C.c_int(nn), _ptr(out), C.c_int(capacity)), "match_slots")
        return out[:n].copy()

    def slot_move_to(self, dst, dst_slot, src_slot):
        """Hand the device-resident region set in self's src_slot over to context dst (no copy)."""
        self._check(lib().mb2_slot_move(dst.h, C.c_int(dst_slot), self.h, C.c_int(src_slot)), "slot_move")

    # ---- verification
    def score_models(self, which, u, models, th, want_resid=False, len_=None, K=None):
        n = len(u) if len_ is None else len_
        K = (np.asarray(models).size // 9) if K is None else K
        I = np.zeros(max(1, K), np.int32); J = np.zeros(max(1, K))
        resid = np.zeros((max(1, K), max(1, n))) if want_resid else None
        self._check(lib().mb2_score_models(self.h, C.c_int(which), _ptr(u), C.c_int(n), _ptr(models), C.c_int(K), C.c_double(th),
                                           _ptr(resid), _ptr(I), _ptr(J)), "score_models")
        return (I[:K], J[:K], resid) if want_resid else (I[:K], J[:K])

    def mods_pair(self, img1, img2, cfg=None, shape1=None, shape2=None, capacity=0):
        """One MODS iteration (mods.cpp:229-415) on a pair through the C++ host mirror."""
        h1, w1 = shape1 if shape1 is not None else img1.shape
        h2, w2 = shape2 if shape2 is not None else img2.shape
        cfg = cfg or PairConfig.default()
        res = PairResult()
        out = np.zeros((max(1, capacity), 4)) if capacity else None
        n = self._check(host_lib().mb2_mods_pair(self.h, _ptr(img1), C.c_int(w1), C.c_int(h1), _ptr(img2), C.c_int(w2), C.c_int(h2),
                                                 C.byref(cfg), C.byref(res), _ptr(out), C.c_int(capacity)), "mods_pair")
        return res, (out[:min(n, capacity)] if capacity else None)

    def mods_pairs(self, pairs, cfg=None, shapes=None, capacity=0):
        """A list of independent pairs through the pipelined driver (mb2_mods_pairs).  pairs: [(img1, img2), ...] (numpy
        arrays, or torch tensors / device pointers together with shapes=[((h1, w1), (h2, w2)), ...]).
        Returns ([PairResult], [verified arrays or None])."""
        n = len(pairs)
        cfg = cfg or PairConfig.default()
        P = C.c_void_p * max(1, n); I = C.c_int * max(1, n)
        p1, p2, w1, h1, w2, h2 = P(), P(), I(), I(), I(), I()
        for k, (a, b) in enumerate(pairs):
            (ha, wa), (hb, wb) = shapes[k] if shapes is not None else (a.shape, b.shape)
            p1[k] = _ptr(a).value; p2[k] = _ptr(b).value
            w1[k], h1[k], w2[k], h2[k] = wa, ha, wb, hb
        res = (PairResult * max(1, n))()
        outs = [np.zeros((capacity, 4)) for _ in range(n)] if capacity else None
        vo = P(*[o.ctypes.data for o in outs]) if capacity and n else None
        caps = I(*([capacity] * n)) if capacity and n else None
        self._check(host_lib().mb2_mods_pairs(self.h, C.c_int(n), p1, w1, h1, p2, w2, h2, C.byref(cfg), res, vo, caps), "mods_pairs")
        results = [res[k] for k in range(n)]
        return results, ([outs[k][:min(results[k].verified, capacity)] for k in range(n)] if capacity else None)

    def mods_multi(self, img1, imgs2, cfg=None, capacity=0):
        """mods_multi.cpp:232-330: one query image against N images (the query is described once).  Returns ([PairResult], [verified rows])."""
        cfg = cfg or PairConfig.default()
        n = len(imgs2)
        P = C.c_void_p * max(1, n); I = C.c_int * max(1, n)
        p2, w2, h2 = P(), I(), I()
        for k, b in enumerate(imgs2):
            p2[k] = _ptr(b).value; h2[k], w2[k] = b.shape
        res = (PairResult * max(1, n))()
        outs = [np.zeros((capacity, 4)) for _ in range(n)] if capacity else None
        vo = P(*[o.ctypes.data for o in outs]) if capacity and n else None
        caps = I(*([capacity] * n)) if capacity and n else None
        self._check(host_lib().mb2_mods_multi(self.h, _ptr(img1), C.c_int(img1.shape[1]), C.c_int(img1.shape[0]), C.c_int(n), p2, w2, h2, C.byref(cfg), res, vo, caps),
                    "mods_multi")
        results = [res[k] for k in range(n)]
        return results, ([outs[k][:min(results[k].verified, capacity)] for k in range(n)] if capacity else None)

    def extract_features(self, img, fname, cfg=None):
        """extract_features.cpp: describe the image and write the reference's text feature cache.  Returns the region count."""
        cfg = cfg or PairConfig.default()
        return self._check(host_lib().mb2_extract_features(self.h, _ptr(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.byref(cfg), str(fname).encode()),
                           "extract_features")

    def mods_pair_cached(self, cache1, cache2, cfg=None, capacity=0):
        """mods.cpp's read_pre_extracted flow: LoadRegions for both images, then match + verify."""
        cfg = cfg or PairConfig.default()
        res = PairResult(); out = np.zeros((max(1, capacity), 4)) if capacity else None
        n = self._check(host_lib().mb2_mods_pair_cached(self.h, str(cache1).encode(), str(cache2).encode(), C.byref(cfg), C.byref(res), _ptr(out), C.c_int(capacity)),
                        "mods_pair_cached")
        return res, (out[:min(n, capacity)] if capacity else None)

    def views_sharded_pair(self, img1, img2, cfg, comm=None, rank=0, world=1, shape1=None, shape2=None, capacity=0):
        """One pair with its synthesised views dealt out over the ranks (mb2_views_sharded_pair).  comm: dist_comm_create() handle
        (None when world == 1).  Returns (PairResult, verified rows or None, digest (4 ints), stats dict)."""
        h1, w1 = shape1 if shape1 is not None else img1.shape
        h2, w2 = shape2 if shape2 is not None else img2.shape
        res = PairResult(); dig = (C.c_ulonglong * 4)(); st = (C.c_double * 8)()
        out = np.zeros((max(1, capacity), 4)) if capacity else None
        f = host_lib().mb2_views_sharded_pair
        n = self._check(f(self.h, C.c_void_p(comm or 0), C.c_int(rank), C.c_int(world), _ptr(img1), C.c_int(w1), C.c_int(h1), _ptr(img2), C.c_int(w2),
                          C.c_int(h2), C.byref(cfg), C.byref(res), _ptr(out), C.c_int(capacity), dig, st), "views_sharded_pair")
        stats = dict(zip(("ms_views", "ms_gather", "ms_match", "ms_tentative_gather", "ms_verify", "allgather_bytes_per_rank", "regions", "units"), list(st)))
        return res, (out[:min(n, capacity)] if capacity else None), [int(v) for v in dig], stats

    def views_sharded_pairs(self, pairs, cfg, comm=None, rank=0, world=1, shapes=None, capacity=0):
        """A list of pairs through the view-sharded dataset call (mb2_views_sharded_pairs): pair k is verified by rank k % world while all
        ranks go on with pair k + 1.  Returns ([PairResult] -- complete on every rank --, [verified rows or None] -- filled on rank
        k % world --, [digest (4 ints)], [stats dict])."""
        n = len(pairs)
        P = C.c_void_p * max(1, n); I = C.c_int * max(1, n)
        p1, p2, w1, h1, w2, h2 = P(), P(), I(), I(), I(), I()
        for k, (a, b) in enumerate(pairs):
            (ha, wa), (hb, wb) = shapes[k] if shapes is not None else (a.shape, b.shape)
            p1[k] = _ptr(a).value; p2[k] = _ptr(b).value
            w1[k], h1[k], w2[k], h2[k] = wa, ha, wb, hb
        res = (PairResult * max(1, n))()
        dig = (C.c_ulonglong * (4 * max(1, n)))(); st = (C.c_double * (8 * max(1, n)))()
        outs = [np.zeros((capacity, 4)) for _ in range(n)] if capacity else None
        vo = P(*[o.ctypes.data for o in outs]) if capacity and n else None
        caps = I(*([capacity] * n)) if capacity and n else None
        f = host_lib().mb2_views_sharded_pairs
        self._check(f(self.h, C.c_void_p(comm or 0), C.c_int(rank), C.c_int(world), C.c_int(n), p1, w1, h1, p2, w2, h2, C.byref(cfg), res, vo, caps, dig, st),
                    "views_sharded_pairs")
        names = ("ms_views", "ms_gather", "ms_match", "ms_tentative_gather", "ms_verify", "allgather_bytes_per_rank", "regions", "units")
        results = [res[k] for k in range(n)]
        ver = [outs[k][:min(results[k].verified, capacity)] if (k % world) == rank else None for k in range(n)] if capacity else None
        return results, ver, [[int(v) for v in dig[4 * k:4 * k + 4]] for k in range(n)], [dict(zip(names, list(st[8 * k:8 * k + 8]))) for k in range(n)]

    def dist_comm_create(self, rank, world):
        """ncclComm_t for the view-sharded driver: rank 0 draws the unique id, torch.distributed (already initialised by the caller) hands
        it to every rank.  Returns an opaque handle for views_sharded_pair / dist_comm_destroy."""
        import torch
        import torch.distributed as dist
        idb = (C.c_ubyte * 128)()
        if rank == 0:
            self._check(host_lib().mb2_dist_unique_id(idb), "dist_unique_id")
        t = torch.tensor(list(idb), dtype=torch.uint8, device="cuda:%d" % self.device if dist.get_backend() == "nccl" else "cpu")
        dist.broadcast(t, src=0)
        idb = (C.c_ubyte * 128)(*[int(v) for v in t.cpu().tolist()])
        comm = C.c_void_p()
        self._check(host_lib().mb2_dist_comm_create(self.h, C.c_int(rank), C.c_int(world), idb, C.byref(comm)), "dist_comm_create")
        return comm.value

    def dist_comm_destroy(self, comm):
        if comm:
            host_lib().mb2_dist_comm_destroy(C.c_void_p(comm))

    def verify(self, frames14, keys, cfg=None, capacity=0):
        """DuplicateFiltering + LORANSACFiltering of gathered tentatives (mb2_host_verify).  Returns (PairResult, verified rows)."""
        cfg = cfg or PairConfig.default()
        frames14 = np.ascontiguousarray(frames14, np.float64); keys = np.ascontiguousarray(keys, np.float64)
        res = PairResult()
        out = np.zeros((max(1, capacity), 4)) if capacity else None
        n = self._check(host_lib().mb2_host_verify(self.h, _ptr(frames14), _ptr(keys), C.c_int(len(keys)), C.byref(cfg), C.byref(res), _ptr(out),
                                                   C.c_int(capacity)), "host_verify")
        return res, (out[:min(n, capacity)] if capacity else None)

    def ransac_f(self, u, th=9.0, conf=0.99, max_sam=100000, errorType=0, doSymCheck=1, seed=1, do_lo=1, inlLimit=None):
        """exp_ransacFcustom (degensac/exp_ranF.c:795).  inlLimit None = len (no limit); LORANSACFiltering passes 0."""
        u = np.ascontiguousarray(u, np.float64)
        n = len(u)
        F = np.zeros(9); inl = np.zeros(max(1, n), np.uint8); data = np.zeros(4, np.int32); J = C.c_double()
        I = self._check(lib().mb2_ransac_f(self.h, _ptr(u), C.c_int(n), C.c_double(th), C.c_double(conf), C.c_int(max_sam),
                                           C.c_int(errorType), C.c_int(doSymCheck), C.c_int(do_lo), C.c_uint(n if inlLimit is None else inlLimit),
                                           C.c_long(seed), _ptr(F), _ptr(inl), _ptr(data), C.byref(J)), "ransac_f")
        return dict(F=F, inl=inl[:n], I=I, samples=int(data[0]), lo=int(data[1]), Ih=int(data[2]), launches=int(data[3]), J=J.value)

    def ransac_h(self, u, th=9.0, conf=0.99, max_sam=100000, errorType=0, doSymCheck=1, seed=1):
        u = np.ascontiguousarray(u, np.float64)
        n = len(u)
        H = np.zeros(9); inl = np.zeros(max(1, n), np.uint8); data = np.zeros(3, np.int32); J = C.c_double()
        I = self._check(lib().mb2_ransac_h(self.h, _ptr(u), C.c_int(n), C.c_double(th), C.c_double(conf), C.c_int(max_sam),
                                           C.c_int(errorType), C.c_int(doSymCheck), C.c_long(seed), _ptr(H), _ptr(inl), _ptr(data),
                                           C.byref(J)), "ransac_h")
        return dict(H=H, inl=inl[:n], I=I, samples=int(data[0]), lo=int(data[1]), rejected=int(data[2]), J=J.value)
