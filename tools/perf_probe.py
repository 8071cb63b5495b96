"""Stage timings at BASELINE config-3 size (4096x3072 pair) -- a probe, not the bench."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mods_b200 as mb
import synth

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 3072)
nb = int(float(sys.argv[3]) * W * H) if len(sys.argv) > 3 else None
t = time.time(); A = synth.blob_image(W, H, seed=1, n_blobs=nb); print("gen A %.1fs" % (time.time() - t))
t = time.time(); B = synth.warp_image(A, synth.gt_homography(W, H)); print("gen B %.1fs" % (time.time() - t))
ctx = mb.Context(0)
REPS = int(os.environ.get("PROBE_REPS", "3"))
def timed(name, f, reps=None):
    reps = reps or REPS
    out = None
    for i in range(reps):
        ctx.sync(); t = time.time(); out = f(); ctx.sync(); dt = time.time() - t
        print("  %-28s run %d: %8.2f ms" % (name, i, dt * 1e3))
    return out
k = timed("hessaff_detect", lambda: ctx.hessaff_detect(A))
print("keys", len(k), "max s", k[:, 6].max())
ko = timed("detect_orientation", lambda: ctx.detect_orientation(A, k))
print("oriented", len(ko))
d = timed("describe_sift", lambda: ctx.describe_sift(A, ko))
va = timed("view A", lambda: ctx.detect_describe_view(A, slot=0))
vb = timed("view B", lambda: ctx.detect_describe_view(B, slot=1))
print("regions", len(va[0]), len(vb[0]))
m = timed("match_slots", lambda: ctx.match_slots(0, 1))
print("tentatives", len(m))
os.environ["MB2_NN_IMPL"] = "simt"
m2 = timed("match_slots simt", lambda: ctx.match_slots(0, 1), reps=1)
print("simt equal:", np.array_equal(m, m2))
print("launches", ctx.launches)
os.environ.pop("MB2_NN_IMPL", None)
for i in range(REPS):
    res, _ = ctx.mods_pair(A, B)
    print("  mods_pair run %d: total %.1f ms  dd %.1f match %.1f dup %.1f ransac %.1f | regions %d %d tent %d uniq %d inl %d ver %d"
          % (i, res.ms_total, res.ms_detect_describe, res.ms_match, res.ms_duplicate, res.ms_ransac, res.regions1, res.regions2,
             res.tentatives, res.unique_tentatives, res.ransac_inliers, res.verified))
