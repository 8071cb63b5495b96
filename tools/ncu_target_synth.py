"""Workload for the ncu capture of the view-synthesis kernels: one tilt-2 rotated view of a 4096x3072 image."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mods_b200 as mb
from mods_b200 import synth
W, H = 4096, 3072
A = synth.blob_image(W, H, seed=1, n_blobs=20000)
ctx = mb.Context(0)
d = torch.from_numpy(A).cuda()
ow, oh = __import__("ctypes").c_int(), __import__("ctypes").c_int()
import ctypes as C
vp = mb.ViewParams(2.0, 0.5, 1.0, 0.5, 1); Hm = np.zeros(9)
rc = mb.lib().mb2_synth_view(ctx.h, C.c_void_p(d.data_ptr()), C.c_int(W), C.c_int(H), C.byref(vp), None, C.c_int(0), C.byref(ow), C.byref(oh), Hm.ctypes.data_as(C.c_void_p))
print("synth view", rc, ow.value, oh.value)
