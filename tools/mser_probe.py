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
