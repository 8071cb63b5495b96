"""Timing of the F-matrix LO-RANSAC driver (mb2_ransac_f) at BASELINE C3 size next to the compiled reference (oracle/_ref, one host
thread: exp_ransacFcustom is sequential) on the same tentatives and seed, and a check that both return the same result."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np
import mods_b200 as mb
from make_golden_f import general_scene
ctx = mb.Context(0)
try:
    from oracle.pyoracle import Reference
    R = Reference()
except Exception as e:  # noqa
    R = None; print("no reference build:", e)
for n, n_out, planar in ((30000, 12000, 0.0), (30000, 18000, 0.5), (5000, 2500, 0.0)):
    u = general_scene(4, n=n, n_out=n_out, noise=0.5, planar_frac=planar)
    for lim in (0,):
        ctx.ransac_f(u, seed=9, inlLimit=lim)
        t0 = time.perf_counter(); g = ctx.ransac_f(u, seed=9, inlLimit=lim); tg = time.perf_counter() - t0
        line = "n %d outliers %d planar %.1f inlLimit %s: GPU %.1f ms (I %d samples %d LO %d Ih %d launches %d)" % (
            n, n_out, planar, lim, 1e3 * tg, g["I"], g["samples"], g["lo"], g["Ih"], g["launches"])
        if R is not None:
            t0 = time.perf_counter(); r = R.exp_ransacF(u, seed=9, inlLimit=lim); tr = time.perf_counter() - t0
            same = [r[k] for k in ("I", "samples", "lo", "Ih")] == [g[k] for k in ("I", "samples", "lo", "Ih")] and np.array_equal(r["inl"], g["inl"])
            line += " | reference CPU %.1f ms (x%.1f) identical=%s" % (1e3 * tr, tr / tg, same)
        print(line, flush=True)
    th0 = time.perf_counter(); h = ctx.ransac_h(u, seed=9); th = time.perf_counter() - th0
    print("   (H driver on the same tentatives: %.1f ms, I %d)" % (1e3 * th, h["I"]))
