"""First-light diagnostics on a B200: every stage of the CUDA path against the CPU oracle, with
enough printing to localise a mismatch from one gpurun round trip.  (Not a test; tests/ has the
assertions.)  Usage: python tools/gpu_diag.py [stage ...]
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mods_b200 as mb  # noqa: E402
from oracle.pyoracle import Oracle, HessParams  # noqa: E402
import synth  # noqa: E402


def cmp(name, a, b, tol=0.0):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print("  [%s] SHAPE MISMATCH gpu %s oracle %s" % (name, a.shape, b.shape)); return False
    if a.size == 0:
        print("  [%s] both empty" % name); return True
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    bad = d > tol
    print("  [%s] shape %s exact=%s maxabs=%.3e nbad=%d/%d" % (name, a.shape, bool((a == b).all()), d.max(), int(bad.sum()), a.size))
    if bad.any():
        idx = np.argwhere(bad)[:5]
        for i in idx:
            i = tuple(i); print("     at %s gpu=%r oracle=%r" % (i, a[i], b[i]))
    return not bad.any()


def stage_pyramid(ctx, O, img):
    print("== pyramid levels (blur / response) vs oracle")
    ctx.hessaff_detect(img, as_regions=False)
    P = O.pyramid(img)
    ok = True
    for lv in P["levels"]:
        g, _ = ctx.pyramid_level(lv["octave"], lv["level"], False)
        r, _ = ctx.pyramid_level(lv["octave"], lv["level"], True)
        ok &= cmp("blur o%d l%d" % (lv["octave"], lv["level"]), g, lv["blur"])
        ok &= cmp("resp o%d l%d" % (lv["octave"], lv["level"]), r[1:-1, 1:-1], lv["resp"][1:-1, 1:-1])
    print("  oracle: extrema %d localized %d keys %d" % (P["extrema"], len(P["localized"]), P["nkeys"]))
    return ok


def stage_detect(ctx, O, img):
    print("== hessaff detect (raw + regions) vs oracle")
    ok = True
    for raw in (True, False):
        t = time.time(); g = ctx.hessaff_detect(img, as_regions=not raw); tg = time.time() - t
        t = time.time(); o = O.hessaff_detect(img, raw=raw); to = time.time() - t
        print("  raw=%s gpu %d (%.1f ms) oracle %d (%.1f ms)" % (raw, len(g), tg * 1e3, len(o), to * 1e3))
        ok &= cmp("keys raw=%s" % raw, g, o, 1e-12)
        if len(g) != len(o) and len(g) and len(o):
            # which keys are missing / extra (match on x,y,s)
            so = {tuple(np.round(r[[0, 1, 6]], 4)) for r in o}; sg = {tuple(np.round(r[[0, 1, 6]], 4)) for r in g}
            print("   only oracle:", list(so - sg)[:5], " only gpu:", list(sg - so)[:5])
    return ok


def stage_orient(ctx, O, img):
    print("== orientation vs oracle")
    k = O.hessaff_detect(img)
    ok = True
    for maxA in (1, 3):
        par = mb.OrientationParams(1.0, 41, maxA, 0.8)
        g = ctx.detect_orientation(img, k, par); o = O.detect_orientation(img, k, maxAngles=maxA)
        print("  maxAngles=%d in %d gpu %d oracle %d" % (maxA, len(k), len(g), len(o)))
        ok &= cmp("oriented maxA=%d" % maxA, g, o, 1e-11)
    return ok


def stage_describe(ctx, O, img):
    print("== describe (patch + SIFT) vs oracle")
    k = O.hessaff_detect(img)
    k = O.detect_orientation(img, k)
    k = O.reproject(k, np.eye(3), img.shape[1], img.shape[0], 0)[0]
    ok = True
    for root in (1, 0):
        par = mb.SiftParams(5.1962, 41, 1, root, 0)
        t = time.time(); gd, gp = ctx.describe_sift(img, k, par, want_patches=True); tg = time.time() - t
        t = time.time(); od, op = O.describe(img, k, rootsift=bool(root), want_patches=True); to = time.time() - t
        print("  root=%d n=%d gpu %.1f ms oracle %.1f ms  max scale %.1f" % (root, len(k), tg * 1e3, to * 1e3, k[:, 6].max() if len(k) else 0))
        ok &= cmp("patches", gp, op)
        ok &= cmp("desc root=%d" % root, gd.astype(np.float32), od)
    return ok


def stage_view(ctx, O, img):
    print("== full view pipeline vs oracle")
    t = time.time(); g = ctx.detect_describe_view(img); tg = time.time() - t
    t = time.time(); g = ctx.detect_describe_view(img); tg2 = time.time() - t
    t = time.time(); o = O.view_pipeline(img); to = time.time() - t
    print("  gpu %d regions (%.1f ms first, %.1f ms second), oracle %d (%.1f ms)" % (len(g[0]), tg * 1e3, tg2 * 1e3, len(o[0]), to * 1e3))
    ok = cmp("det_kp", g[0], o[0], 1e-11)
    ok &= cmp("reproj_kp", g[1], o[1], 1e-11)
    ok &= cmp("desc", g[2].astype(np.float32), o[2])
    return ok


def stage_match(ctx, O, img):
    print("== FGINN matching (tcgen05 and SIMT) vs oracle")
    ok = True
    for (nq, nt, seed) in ((300, 257, 1), (1000, 1500, 2), (3000, 2900, 3)):
        t_desc, _ = synth.random_descriptors(nt, seed)
        q_desc, src = synth.random_descriptors(nq, seed + 100, dup_of=t_desc, dup_frac=0.5)
        rng = np.random.default_rng(seed)
        txy = rng.uniform(0, 1000, size=(nt, 2))
        # clusters of near-identical trains close in space (geometrically consistent 2nd NNs)
        for i in range(0, nt - 1, 7):
            t_desc[i + 1] = np.clip(t_desc[i].astype(int) + rng.integers(-2, 3, 128), 0, 255); txy[i + 1] = txy[i] + rng.uniform(-5, 5, 2)
        o = O.match_fginn(q_desc.astype(np.float32), t_desc.astype(np.float32), txy)
        for impl in ("tc", "simt"):
            os.environ["MB2_NN_IMPL"] = impl
            try:
                t = time.time(); g = ctx.match_fginn(q_desc, t_desc, txy); tg = time.time() - t
                print("  %s nq=%d nt=%d: gpu %d tentatives (%.1f ms), oracle %d" % (impl, nq, nt, len(g), tg * 1e3, len(o)))
                ok &= cmp("tentatives %s" % impl, g, o)
            except Exception as e:  # keep going: the other implementation is the cross-check
                ok = False; print("  %s FAILED: %s" % (impl, e))
        os.environ.pop("MB2_NN_IMPL", None)
    return ok


def stage_score(ctx, O, img):
    print("== batched scorer vs oracle")
    rng = np.random.default_rng(5)
    n, K = 2000, 64
    u = np.zeros((n, 6)); u[:, 0:2] = rng.random((n, 2)) * 1000; u[:, 2] = 1; u[:, 5] = 1
    Hgt = synth.gt_homography(1000, 1000)
    p = (Hgt @ u[:, 0:3].T).T; u[:, 3:5] = p[:, :2] / p[:, 2:3] + rng.normal(size=(n, 2))
    u[n // 2:, 3:5] = rng.random((n - n // 2, 2)) * 1000
    M0 = np.linalg.inv(Hgt).T.ravel()
    models = np.stack([M0 * (1 + 1e-3 * rng.normal(size=9)) for _ in range(K)])
    ok = True
    for which in range(5):
        I, J, R = ctx.score_models(which, u, models, 9.0, want_resid=True)
        Ro = np.stack([O.score(which, u, m) for m in models])
        Io = (Ro <= 9.0).sum(1)
        Jo = np.array([sum((0.0 if (9.0 == 0 or e >= 9.0 * 9 / 4) else 1 - (e / (9.0 * 9 / 4))) for e in r) for r in Ro])
        ok &= cmp("resid which=%d" % which, R, Ro)
        ok &= cmp("I which=%d" % which, I, Io)
        ok &= cmp("J which=%d" % which, J, Jo, 1e-9)
    return ok


STAGES = dict(pyramid=stage_pyramid, detect=stage_detect, orient=stage_orient, describe=stage_describe, view=stage_view,
              match=stage_match, score=stage_score)

if __name__ == "__main__":
    want = sys.argv[1:] or list(STAGES)
    O = Oracle()
    ctx = mb.Context(0)
    img = synth.blob_image(640, 480, seed=3)
    results = {}
    for s in want:
        try:
            results[s] = STAGES[s](ctx, O, img)
        except Exception:
            traceback.print_exc(); results[s] = False
    print("SUMMARY", results, "launches", ctx.launches)
