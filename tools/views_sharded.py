#!/usr/bin/env python
"""View-sharded pair on N GPUs (BASELINE config C4 shape): the synthesised views of both images are dealt out over the ranks,
one NCCL all-gather of the packed regions, matching split by query rows, verification on rank 0.

    python tools/views_sharded.py [--size WxH] [--tier hess4|hess4+mser2|full]            (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/views_sharded.py ...     (N GPUs)

Prints one JSON line on rank 0 with a digest of the result (identical for every N) and the device time."""
import argparse, hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="2048x1536", type=lambda s: tuple(int(v) for v in s.lower().split("x")))
    ap.add_argument("--tier", default="hess4+mser2")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch, torch.distributed as dist
    import mods_b200 as mb
    from mods_b200 import sharding, synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    D = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local)); D = dist
    w, h = args.size
    A = synth.blob_image(w, h, seed=1, n_blobs=int(1.5e-3 * w * h))
    B = synth.warp_image(A, synth.gt_homography(w, h), seed=2)
    H = mb.host_lib()
    import ctypes as C

    def vs(scales, tilts, phi, prev):
        scales = np.asarray(scales, np.float64); tilts = np.asarray(tilts, np.float64); prev = np.ascontiguousarray(np.asarray(prev, np.float64).reshape(-1, 3))
        out = np.zeros((512, 3))
        n = H.mb2_host_set_vs_pars(scales.ctypes.data_as(C.c_void_p), C.c_int(len(scales)), tilts.ctypes.data_as(C.c_void_p), C.c_int(len(tilts)), C.c_double(phi),
                                   prev.ctypes.data_as(C.c_void_p), C.c_int(len(prev)), out.ctypes.data_as(C.c_void_p), C.c_int(512))
        return out[:n]
    h4 = vs([1], [1, 2, 4, 6, 8], 360, [])
    m2 = vs([1, 0.25, 0.125], [1], 360, [])
    tiers = {"HessianAffine": [tuple(r) + (0.2,) for r in h4]}                       # [HessianAffine4]: initSigma 0.2
    if "mser2" in args.tier or args.tier == "full":
        tiers["MSER"] = [tuple(r) + (0.8,) for r in m2]                               # [MSER2]: initSigma 0.8
    if args.tier == "full":
        h5 = vs([1], [1, 2, 4, 6, 8], 120, h4); h6 = vs([1], [1, 2, 4, 6, 8], 60, np.concatenate([h4, h5]))
        m3 = vs([1, 0.25, 0.125], [1, 3, 6, 9], 360, m2)
        tiers["HessianAffine"] += [tuple(r) + (0.2,) for r in np.concatenate([h5, h6])]
        tiers["MSER"] += [tuple(r) + (0.8,) for r in m3]
    units, costs = sharding.iters_units(w, h, tiers)
    ctx = mb.Context(local)
    cfg = mb.PairConfig.default(); cfg.use_mser = 1
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    compute, match = sharding.gpu_workers(ctx, dA, dB, cfg, shape=(h, w))
    dev = torch.device("cuda", local) if world > 1 else "cpu"
    times = []
    for rep in range(args.reps):
        torch.cuda.synchronize()
        if D: D.barrier()
        t0 = time.perf_counter()
        groups = sharding.pair_views_sharded(compute, match, units, costs, dist=D, device=dev)
        res = None
        if rank == 0:
            frames, keys = sharding.frames_and_keys(groups)
            res, ver = ctx.verify(frames, keys, cfg, capacity=len(keys) + 1)
        torch.cuda.synchronize()
        if D: D.barrier()
        times.append(time.perf_counter() - t0)
    if rank == 0:
        dig = hashlib.sha256()
        for det in sorted(groups):
            for a in groups[det]:
                dig.update(np.ascontiguousarray(a).tobytes())
        dig.update(np.ascontiguousarray(ver).tobytes())
        print(json.dumps({"mode": "views-sharded", "n_gpus": world, "size": "%dx%d" % (w, h), "views_per_image": {k: len(v) for k, v in tiers.items()},
                          "regions": {k: [len(v[0]), len(v[2])] for k, v in groups.items()}, "tentatives": int(res.tentatives), "verified": int(res.verified),
                          "s_per_pair": min(times), "pairs_per_s": 1.0 / min(times), "digest": dig.hexdigest()[:16]}))
    if D: D.destroy_process_group()


if __name__ == "__main__":
    main()
