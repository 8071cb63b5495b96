"""View-sharded pair through the C driver (mb2_views_sharded_pair, libmods_host.so): the [HessianAffine4..6] + [MSER2..3] tiers of
build/iters_mods_cviu.ini (61 + 27 views per image, BASELINE config C4) or a smaller tier set, on 1..N GPUs of one node.
    python tools/views_sharded_c.py [WxH] [c4|small]                         (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/views_sharded_c.py [WxH] [c4|small]
Prints one JSON line on rank 0: ms per pair (max over ranks, CUDA-synchronised), result counts, the order-sensitive digest (identical for every N)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mods_b200 as mb
from mods_b200 import synth


def tiers(which):
    H = mb.host_lib()

    def vs(scales, tilts, phi, prev):
        scales = np.asarray(scales, np.float64); tilts = np.asarray(tilts, np.float64)
        prev = np.ascontiguousarray(np.asarray(prev, np.float64).reshape(-1, 3)); out = np.zeros((512, 3))
        n = H.mb2_host_set_vs_pars(scales.ctypes.data_as(C.c_void_p), C.c_int(len(scales)), tilts.ctypes.data_as(C.c_void_p), C.c_int(len(tilts)),
                                   C.c_double(phi), prev.ctypes.data_as(C.c_void_p), C.c_int(len(prev)), out.ctypes.data_as(C.c_void_p), C.c_int(512))
        return out[:n].copy()
    m2 = vs([1, 0.25, 0.125], [1], 360, [])
    h4 = vs([1], [1, 2, 4, 6, 8], 360, [])
    if which == "small":
        hess, mser = h4, m2
    else:   # every SIFT tier of iters_mods_cviu.ini: [MSER2] 3 + [MSER3] 24, [HessianAffine4] 11 + [5] 20 + [6] 30
        m3 = vs([1, 0.25, 0.125], [1, 3, 6, 9], 360, m2)
        h5 = vs([1], [1, 2, 4, 6, 8], 120, h4); h6 = vs([1], [1, 2, 4, 6, 8], 60, np.concatenate([h4, h5]))
        hess, mser = np.concatenate([h4, h5, h6]), np.concatenate([m2, m3])
    # rows are (zoom, tilt, phi); set_views wants (tilt, phi, zoom, InitSigma)
    return [(r[1], r[2], r[0], 0.2) for r in hess], [(r[1], r[2], r[0], 0.8) for r in mser]


def main():
    w, h = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096x3072").split("x"))
    which = sys.argv[2] if len(sys.argv) > 2 else "c4"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    A = synth.blob_image(w, h, seed=1, n_blobs=int(1.5e-3 * w * h)); B = synth.warp_image(A, synth.gt_homography(w, h), seed=2)
    ctx = mb.Context(lr)
    cfg = mb.PairConfig.default(); cfg.use_mser = 1; cfg.mserMatchRatio = 0.85
    hess, mser = tiers(which)
    cfg.set_views(hess, mser)
    comm = ctx.dist_comm_create(rank, world) if world > 1 else None
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    times, last = [], None
    for it in range(steps + 1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        last = ctx.views_sharded_pair(dA, dB, cfg, comm, rank, world, shape1=(h, w), shape2=(h, w), capacity=1 << 17)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        if it > 0:
            times.append(dt)
    res, ver, dig, st = last
    digs = [dig]
    if world > 1:
        g = [None] * world; dist.all_gather_object(g, dig); digs = g
    if rank == 0:
        print(json.dumps({"size": [w, h], "tiers": which, "views_per_image": [len(hess), len(mser)], "n_gpus": world, "ms_per_pair": 1e3 * float(np.median(times)),
                          "regions": [res.regions1, res.regions2], "tentatives": res.tentatives, "unique": res.unique_tentatives, "inliers": res.ransac_inliers,
                          "verified": res.verified, "digest": ["%016x" % d for d in dig], "digest_same_on_all_ranks": all(d == dig for d in digs), "rank0_stats": st}))
    if world > 1:
        dist.barrier()
        ctx.dist_comm_destroy(comm)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
