"""Repeats the two timed arms of bench.py (device-resident images / pinned host images through mb2_mods_pairs) several times and prints
every repetition: shows whether the end-to-end arm is stable.  Usage: python tools/e2e_probe.py [steps] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mods_b200 as mb
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w, h = (int(v) for v in os.environ.get("PROBE_SIZE", "4096x3072").split("x"))
pairs = bench.make_pairs(w, h, bench.N_PAIRS, seed0=1)
ctx = mb.Context(0)
cfg = mb.PairConfig.default(); cfg.use_mser = 1
dev = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in pairs]
pin = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in pairs]


def timed(bufs, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res, _ = ctx.mods_pairs([bufs[s % len(bufs)] for s in range(n)], cfg, shapes=[((h, w), (h, w))] * n)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n, res


timed(dev, 3)
configs = [dict(MB2_LANES="1"), dict(MB2_LANES="2"), dict(MB2_LANES="1", MB2_MSER_AHEAD="1")] if not os.environ.get("PROBE_LANES_ONLY") else [dict(MB2_LANES="1"), dict(MB2_LANES="2")]
if os.environ.get("PROBE_SAMPLERS"):   # what samples the clocks while the arms run (bench.ClockSampler)
    for mode in ("none", "nvidia-smi", "nvml"):
        sampler = None
        if mode != "none":
            sampler = bench.ClockSampler(0, force_smi=(mode == "nvidia-smi"))
            sampler.start()
            time.sleep(1.0 if mode == "nvidia-smi" else 0.1)   # nvidia-smi takes a moment to come up
        for r in range(reps):
            for name, bufs in (("dev", dev), ("pin", pin)):
                ms, res = timed(bufs, steps)
                print("[sampler %s] %s rep %d: %.2f ms per pair (%.1f pairs/s)  verified %d" % (mode, name, r, ms, 1e3 / ms, res[0].verified), flush=True)
        if sampler:
            print("   ", sampler.stop(), flush=True)
for env in configs:   # schedule toggles of mb2_mods_pairs (read per call)
    for k in ("MB2_LANES", "MB2_MSER_AHEAD"):
        os.environ.pop(k, None)
    os.environ.update(env)
    timed(dev, 4); timed(pin, 4)
    for r in range(reps):
        for name, bufs in (("dev", dev), ("pin", pin)):
            ms, res = timed(bufs, steps)
            print("%s %s rep %d: %.2f ms per pair (%.1f pairs/s)  verified %s" % (env, name, r, ms, 1e3 / ms, [x.verified for x in res[:3]]), flush=True)
