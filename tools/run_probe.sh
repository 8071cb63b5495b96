MB2_RANSAC_TRACE=1 python - <<'P' 2>&1 | tail -12
import sys, time
sys.path.insert(0, '.')
import torch, bench
import mods_b200 as mb
w, h = 4096, 3072
pairs = bench.make_pairs(w, h, 1, seed0=1)
ctx = mb.Context(0)
cfg = mb.PairConfig.default(); cfg.use_mser = 1
a, b = (torch.from_numpy(x).cuda() for x in pairs[0])
for it in range(3):
    t0 = time.perf_counter()
    res, _ = ctx.mods_pair(a, b, cfg, shape1=(h, w), shape2=(h, w))
    print("pair %.1f ms: detect %.1f match %.1f dup %.1f ransac %.1f" % (1e3 * (time.perf_counter() - t0), res.ms_detect_describe, res.ms_match, res.ms_duplicate, res.ms_ransac), flush=True)
P
