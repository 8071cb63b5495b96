import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mods_b200 as mb
from oracle.pyoracle import Oracle
import synth
os.system("grep 'model name' /proc/cpuinfo | head -1; grep -o -w 'fma' /proc/cpuinfo | head -1")
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
O = Oracle(); ctx = mb.Context(0)
im = synth.blob_image(320, 240, seed=11)
for raw in (True, False):
    g = ctx.hessaff_detect(im, as_regions=not raw); o = O.hessaff_detect(im, raw=raw); gold = G["s_raw"] if raw else G["s_reg"]
    for name, a, b in (("gpu-vs-oracle", g, o), ("oracle-vs-golden", o, gold), ("gpu-vs-golden", g, gold)):
        if a.shape != b.shape: print(raw, name, "shape", a.shape, b.shape); continue
        bad = np.argwhere(a != b)
        print(raw, name, "nbad", len(bad))
        for i in bad[:6]: print("    ", tuple(i), repr(a[tuple(i)]), repr(b[tuple(i)]))
o = O.view_pipeline(im)
for nm, a, b in (("det", o[0], G["s_det"]), ("rep", o[1], G["s_rep"]), ("desc", o[2], G["s_desc"].astype(np.float32))):
    print("oracle view vs golden", nm, a.shape, b.shape, (a.shape == b.shape) and int((a != b).sum()))
