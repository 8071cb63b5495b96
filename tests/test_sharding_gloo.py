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
