python -m pytest tests -m gpu -x -q -k "mods_pairs or mods_pair_with_mser" 2>&1 | tail -2
PROBE_SIZE=1920x1080 python - <<'P' 2>&1 | tail -3
import os, sys, time
sys.path.insert(0, '.')
import torch, bench
import mods_b200 as mb
w, h = 1920, 1080
pairs = bench.make_pairs(w, h, 3, seed0=1)
ctx = mb.Context(0); cfg = mb.PairConfig.default(); cfg.use_mser = 1
dev = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in pairs]
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res, _ = ctx.mods_pairs([dev[s % 3] for s in range(24)], cfg, shapes=[((h, w), (h, w))] * 24)
    torch.cuda.synchronize(); print("default lanes, 1080p: %.2f ms per pair, verified %s" % ((time.perf_counter() - t0) * 1e3 / 24, [r.verified for r in res[:3]]))
P
