"""N > 1 host logic on CPU: two gloo ranks shard independent units (views), describe them with the CPU
oracle, exchange with one all-gather and must reproduce the single-process region list bit for bit,
in the reference's order (view-index order)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from mods_b200 import sharding, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _units():
    return [synth.blob_image(160 + 16 * i, 120, seed=40 + i) for i in range(5)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import Oracle
    O = Oracle()
    units = _units()
    mine = sharding.assign_units([sharding.view_cost(u.shape[1], u.shape[0]) for u in units], world)[rank]
    local = {}
    for i in mine:
        det, rep, desc = O.view_pipeline(units[i])
        local[i] = (det, rep, desc.astype(np.uint8))
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    det, rep, desc, offs = sharding.merge_in_unit_order(gathered)
    if rank == 0:
        q.put((det, rep, desc, offs, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_serial_order():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    det, rep, desc, offs, mine0 = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    from oracle.pyoracle import Oracle
    O = Oracle()
    serial = [O.view_pipeline(u) for u in _units()]
    assert np.array_equal(det, np.concatenate([s[0] for s in serial]))
    assert np.array_equal(rep, np.concatenate([s[1] for s in serial]))
    assert np.array_equal(desc, np.concatenate([s[2] for s in serial]).astype(np.uint8))
    assert 0 < len(mine0) < 5 and offs[0] == 0


def test_assignment_is_balanced_and_complete():
    costs = [sharding.view_cost(4096, 3072, t) for t in (1, 2, 2, 4, 4, 4, 6, 6, 8, 8, 8)]
    for world in (1, 2, 4, 8):
        a = sharding.assign_units(costs, world)
        assert sorted(i for r in a for i in r) == list(range(len(costs)))
        loads = [sum(costs[i] for i in r) for r in a]
        assert max(loads) <= sum(costs) / world + max(costs)


# ---- view-sharded pair (BASELINE C4): two gloo ranks, oracle as the per-view worker --------------------------------
def _pair_units():
    A = synth.blob_image(200, 150, seed=71, n_blobs=120)
    B = synth.warp_image(A, synth.gt_homography(200, 150), seed=72)
    views = [(1.0, 0.0, 1.0), (2.0, 0.0, 1.0), (2.0, 1.5707963, 1.0)]
    units = [(im, det, vi, v) for im in (0, 1) for det in ("HessianAffine", "MSER") for vi, v in enumerate(views if det == "HessianAffine" else views[:1])]
    costs = [sharding.view_cost(200, 150, abs(u[3][0]), u[3][2]) for u in units]
    return (A, B), units, costs


def _oracle_workers():
    from oracle.pyoracle import Oracle
    O = Oracle()
    (A, B), units, costs = _pair_units()

    def compute(u):
        img = (A, B)[u[0]]
        det, rep, desc = O.view_pipeline_synth(img, *u[3], detector=0 if u[1] == "HessianAffine" else 3)
        return det, rep, desc.astype(np.uint8)

    def match(det_name, q_rep, q_desc, t_rep, t_desc, lo, hi):
        rows = O.match_fginn(q_desc[lo:hi].astype(np.float32), t_desc.astype(np.float32), np.ascontiguousarray(t_rep[:, :2]))
        rows[:, 0] += lo
        return rows
    return compute, match, units, costs


def _pair_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    compute, match, units, costs = _oracle_workers()
    out = sharding.pair_views_sharded(compute, match, units, costs, dist=dist, device="cpu")
    if rank == 1:   # every rank ends with the same data: check on the non-zero one
        q.put({k: tuple(np.array(a) for a in v) for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharded_pair_equals_serial():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pair_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = q.get(timeout=300)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    compute, match, units, costs = _oracle_workers()
    want = sharding.pair_views_sharded(compute, match, units, costs)
    assert sorted(got) == ["HessianAffine", "MSER"]
    for det in want:
        assert all(np.array_equal(a, b) for a, b in zip(got[det], want[det])), det
        assert len(want[det][0]) > 10
    assert len(want["HessianAffine"][4]) > 5
    frames, keys = sharding.frames_and_keys(got)
    assert frames.shape[1] == 14 and len(frames) == len(keys) == sum(len(v[4]) for v in got.values())


# ---- the C driver's plan and wire layout (mods_b200/host/mods_sharded.cpp) with two gloo ranks ---------------------------------------
def _layout_worker(rank, world, port, q):
    """What mb2_views_sharded_pair does between detection and matching, on the CPU: every rank packs the 184-byte records of its units
    in ascending unit order, the per-unit counts and the padded record blocks are all-gathered, and mb2_shard_layout says where every
    unit sits in the gathered buffer."""
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    compute, match, units, costs = _oracle_workers()
    owner = sharding.owners_c(costs, world)
    counts = np.zeros(len(units), np.int32); blocks = []
    for u in range(len(units)):
        if owner[u] != rank:
            continue
        det, rep, desc = compute(units[u])
        counts[u] = len(rep); blocks.append(sharding.pack_regions(rep, desc))
    allc = [torch.zeros(len(units), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(allc, torch.from_numpy(counts))
    counts_all = np.array([int(allc[owner[u]][u]) for u in range(len(units))], np.int32)
    stride, off = sharding.layout_c(owner, counts_all, world)
    mine = np.concatenate(blocks) if blocks else np.zeros((0, sharding.REC), np.uint8)
    pad = np.zeros((stride, sharding.REC), np.uint8); pad[:len(mine)] = mine
    got = [torch.zeros((stride, sharding.REC), dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(got, torch.from_numpy(pad))
    buf = np.concatenate([g.numpy() for g in got])
    per_unit = [buf[off[u]: off[u] + counts_all[u]].copy() for u in range(len(units))]
    if rank == 1:
        q.put((owner.tolist(), counts_all.tolist(), int(stride), per_unit))
    dist.barrier()
    dist.destroy_process_group()


def test_c_plan_and_record_layout_with_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_layout_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    owner, counts, stride, per_unit = q.get(timeout=300)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    compute, match, units, costs = _oracle_workers()
    assert sorted(set(owner)) == [0, 1] and stride == max(sum(c for c, o in zip(counts, owner) if o == r) for r in (0, 1))
    for u, unit in enumerate(units):   # every unit's records arrive intact, wherever it was computed
        det, rep, desc = compute(unit)
        assert counts[u] == len(rep) and np.array_equal(per_unit[u], sharding.pack_regions(rep, desc)), u
    assert sharding.REC == 184
