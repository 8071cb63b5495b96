"""Timing probe for the MSER path at BASELINE sizes (run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mods_b200 as mb
from mods_b200 import synth
w, h = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096x3072").split("x"))
img = synth.blob_image(w, h, seed=1, n_blobs=int(1.5e-3 * w * h))
ctx = mb.Context(0)
d = torch.from_numpy(img).cuda()
for it in range(3):
    t0 = time.perf_counter(); k = ctx.mser_detect(d, shape=(h, w), capacity=400000); t1 = time.perf_counter()
    print("mser_detect %dx%d: %d keys, %.1f ms" % (w, h, len(k), 1e3 * (t1 - t0)))
ctx.profile_begin(); ctx.mser_detect(d, shape=(h, w), capacity=400000); p = ctx.profile_end()
tot = sum(v[1] for v in p.values())
for name, (n, ms) in sorted(p.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOPK", "16"))]:
    print("  %-22s %5d launches %8.3f ms  %5.1f%%" % (name, n, ms, 100 * ms / tot))
hist = np.bincount(img.astype(np.uint8).ravel(), minlength=256)
print("  pixels per level (top):", sorted([(int(c), i) for i, c in enumerate(hist)], reverse=True)[:16])
print("  total kernel ms (profiled, serialised): %.2f" % tot)
t0 = time.perf_counter(); k2 = ctx.hessaff_detect(d, shape=(h, w), capacity=400000); print("hessaff_detect: %d keys %.1f ms" % (len(k2), 1e3 * (time.perf_counter() - t0)))
# pair-batched detection (both images stacked)
img2 = synth.warp_image(img, synth.gt_homography(w, h), seed=2)
d2 = torch.from_numpy(img2).cuda()
import ctypes as C
n1, n2 = C.c_int(), C.c_int()
par = mb.MserParams.default()
for it in range(3):
    t0 = time.perf_counter()
    rc = mb.lib().mb2_mser_detect_pair(ctx.h, C.c_void_p(d.data_ptr()), C.c_void_p(d2.data_ptr()), C.c_int(w), C.c_int(h), C.byref(par), C.byref(n1), C.byref(n2))
    ctx.sync(); t1 = time.perf_counter()
    print("mser_detect_pair: rc %d, %d + %d keys, %.1f ms" % (rc, n1.value, n2.value, 1e3 * (t1 - t0)))
ctx.profile_begin()
mb.lib().mb2_mser_detect_pair(ctx.h, C.c_void_p(d.data_ptr()), C.c_void_p(d2.data_ptr()), C.c_int(w), C.c_int(h), C.byref(par), C.byref(n1), C.byref(n2))
p = ctx.profile_end()
tot = sum(v[1] for v in p.values())
for name, (n, ms) in sorted(p.items(), key=lambda kv: -kv[1][1])[:10]:
    print("  %-22s %5d launches %8.3f ms  %5.1f%%" % (name, n, ms, 100 * ms / tot))
print("  total kernel ms (pair, profiled): %.2f" % tot)
t0 = time.perf_counter(); v = ctx.mser_pair_views(d, d2, shape=(h, w)); print("mser_pair_views (detect + orient + describe both): %.1f ms, %d + %d regions" % (1e3 * (time.perf_counter() - t0), len(v[0][0]), len(v[1][0])))
