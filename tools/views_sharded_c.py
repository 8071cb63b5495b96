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
    return mb.iters_mods_cviu_views(which)


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
                          "verified": res.verified, "digest": ["%016x" % d for d in dig], "digest_same_on_all_ranks": all(d == dig for d in digs), "rank0_stats": st, "ms_duplicate": res.ms_duplicate, "ms_ransac": res.ms_ransac}))
    if world > 1:
        dist.barrier()
        ctx.dist_comm_destroy(comm)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
