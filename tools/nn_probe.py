"""Matcher probe: per-kernel times (CUDA events inside the library) of one N x N FGINN match of SIFT-like u8 descriptors, for the
epilogue-warp settings of k_nn_tc (MB2_NN_EPI_WARPS) -- a probe, not the bench.  Usage: python tools/nn_probe.py [N] [dup_frac]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mods_b200 as mb
import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
t_desc, _ = synth.random_descriptors(N, 1)
q_desc, src = synth.random_descriptors(N, 2, dup_of=t_desc, dup_frac=frac, jitter=4)
txy = np.random.default_rng(3).uniform(0, 4096, size=(N, 2))
ctx = mb.Context(0)
ref = None
for ew in (os.environ.get("NN_PROBE_EW", "8,16").split(",")):
    os.environ["MB2_NN_EPI_WARPS"] = ew
    for rep in range(3):
        ctx.profile_begin()
        a = ctx.match_fginn(q_desc, t_desc, txy)
        prof = ctx.profile_end()
    if ref is None: ref = a
    tc = {k: v for k, v in prof.items() if "k_nn" in k}
    tot = sum(v[1] for v in tc.values())
    flop = 2.0 * 2 * N * N * 128
    print("EW=%s  tentatives %d  same=%s  %s  | nn kernels %.3f ms  %.0f TFLOP/s over both passes"
          % (ew, len(a), np.array_equal(a, ref), "  ".join("%s %.3f" % (k, v[1]) for k, v in sorted(tc.items())), tot, flop / tot / 1e9))
