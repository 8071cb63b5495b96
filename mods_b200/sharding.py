"""Host-side sharding of independent units (synthesised views of an image, or whole pairs) over ranks.

SURVEY.md 8e: every (image, detector, view) is independent through detection -> description; the reference's
own OpenMP unit (imagerepresentation.cpp:621).  Units are dealt out statically, heaviest first, and the
per-unit results are put back in unit order after the exchange, because region identity in MODS is the
position in RegionVectorMap[det][desc] (views appended in view-index order, imagerepresentation.cpp:2044-2045).
"""
import numpy as np


def _host():
    import mods_b200 as mb
    return mb.host_lib()


def owners_c(costs, world):
    """The shipped plan (mb2_shard_assign, mods_b200/host/mods_sharded.cpp): owner rank of every unit."""
    import ctypes as C
    costs = np.ascontiguousarray(costs, np.float64); owner = np.zeros(max(1, len(costs)), np.int32)
    _host().mb2_shard_assign(costs.ctypes.data_as(C.c_void_p), C.c_int(len(costs)), C.c_int(world), owner.ctypes.data_as(C.c_void_p))
    return owner[:len(costs)]


def layout_c(owner, counts, world):
    """mb2_shard_layout: (stride in records, offset of every unit's records in the all-gathered buffer of world x stride records)."""
    import ctypes as C
    owner = np.ascontiguousarray(owner, np.int32); counts = np.ascontiguousarray(counts, np.int32); off = np.zeros(max(1, len(owner)), np.int32)
    stride = _host().mb2_shard_layout(owner.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), C.c_int(len(owner)), C.c_int(world),
                                      off.ctypes.data_as(C.c_void_p))
    return stride, off[:len(owner)]


def assign_units(costs, world):
    """Longest-processing-time-first (the C plan): returns a list (per rank) of unit indices, deterministic."""
    owner = owners_c(costs, world)
    return [[int(i) for i in np.flatnonzero(owner == r)] for r in range(world)]


def view_cost(w, h, tilt=1.0, zoom=1.0):
    """Pixel count of a synthesised view: tilt t shrinks one side by 1/t, zoom z both by z (synth-detection.cpp:301-342)."""
    import ctypes as C
    return float(_host().mb2_shard_view_cost(C.c_int(int(w)), C.c_int(int(h)), C.c_double(tilt), C.c_double(zoom)))


def merge_in_unit_order(gathered):
    """gathered: list over ranks of {unit: (det_kp, reproj_kp, desc)} -> concatenation in unit order."""
    merged = {}
    for per_rank in gathered:
        merged.update(per_rank)
    units = sorted(merged)
    det = np.concatenate([merged[u][0] for u in units]) if units else np.zeros((0, 9))
    rep = np.concatenate([merged[u][1] for u in units]) if units else np.zeros((0, 9))
    desc = np.concatenate([merged[u][2] for u in units]) if units else np.zeros((0, 128), np.uint8)
    offsets = np.cumsum([0] + [len(merged[u][0]) for u in units])
    return det, rep, desc, dict(zip(units, offsets[:-1]))


# --------------------------------------------------------------------------------------------------------------------
# View-sharded pair (BASELINE config C4): the synthesised views of both images are dealt out over the ranks, every rank
# detects / describes its views, ONE all-gather (padded, byte-packed: 128 B descriptor + 7 doubles of the reprojected frame
# per region) gives every rank all regions in the reference's order, the N1 x N2 matching is split by query rows, and the
# tentatives are gathered for verification on rank 0.
# --------------------------------------------------------------------------------------------------------------------
REC = 128 + 7 * 8   # bytes per region on the wire: u8[128] descriptor + reproj_kp (x y a11 a12 a21 a22 s)


def pack_regions(rep, desc):
    """(n x 9 reproj_kp, n x 128 u8) -> n x REC bytes."""
    n = len(rep)
    out = np.zeros((n, REC), np.uint8)
    if n:
        out[:, :128] = desc
        out[:, 128:] = np.ascontiguousarray(rep[:, :7], np.float64).view(np.uint8).reshape(n, 56)
    return out


def unpack_regions(buf):
    n = len(buf)
    desc = np.ascontiguousarray(buf[:, :128])
    rep7 = np.ascontiguousarray(buf[:, 128:]).view(np.float64).reshape(n, 7) if n else np.zeros((0, 7))
    return rep7, desc


def _all_gather_padded(local, dist, device):
    """local: uint8 array (m x width).  One all-gather of a count, one of the padded payload.  Returns the list of per-rank arrays."""
    import torch
    world = dist.get_world_size()
    width = local.shape[1]
    cnt = torch.tensor([len(local)], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    mx = max(max(cnts), 1)
    pad = np.zeros((mx, width), np.uint8)
    pad[: len(local)] = local
    t = torch.from_numpy(pad).to(device)
    out = torch.empty((world * mx, width), dtype=torch.uint8, device=device)   # concatenation along dim 0 (gloo and nccl agree on this form)
    dist.all_gather_into_tensor(out, t)
    out = out.cpu().numpy().reshape(world, mx, width)
    return [out[r, : cnts[r]] for r in range(world)]


def pair_views_sharded(compute_unit, match, units, costs, dist=None, device="cpu"):
    """units: list of (image 0/1, detector name, view index, view params); compute_unit(unit) -> (det_kp, reproj_kp, desc_u8);
    match(det_name, q_rep7, q_desc, t_rep7, t_desc, q_lo, q_hi) -> tentative rows (n x 7: q idx0 idxJ idx1 d0 dJ d1) for queries
    [q_lo, q_hi).  Returns {det_name: (rep7 image 0, desc image 0, rep7 image 1, desc image 1, tentative rows)} -- identical on every rank."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    mine = assign_units(costs, world)[rank]
    # wire layout of this rank: [unit index (8 B) | count (8 B)] headers travel in the same payload as rows of width REC
    rows, header = [], []
    for i in mine:
        det, rep, desc = compute_unit(units[i])
        header.append((i, len(rep)))
        rows.append(pack_regions(rep, desc))
    head = np.zeros((len(mine), REC), np.uint8)
    if mine:
        head[:, :16] = np.asarray(header, np.int64).view(np.uint8).reshape(len(mine), 16)
    n_units_local = np.zeros((1, REC), np.uint8); n_units_local[0, :8] = np.asarray([len(mine)], np.int64).view(np.uint8)
    payload = np.concatenate([n_units_local, head] + rows) if rows else np.concatenate([n_units_local, head])
    per_rank = _all_gather_padded(payload, dist, device) if dist is not None else [payload]
    blocks = {}
    for buf in per_rank:
        k = int(buf[0, :8].view(np.int64)[0])
        hd = buf[1: 1 + k, :16].copy().view(np.int64).reshape(k, 2)
        off = 1 + k
        for u, n in hd:
            blocks[int(u)] = buf[off: off + int(n)]
            off += int(n)
    out = {}
    for det_name in sorted({u[1] for u in units}):
        sets = []
        for image in (0, 1):
            idx = [i for i, u in enumerate(units) if u[0] == image and u[1] == det_name]   # view-index order (imagerepresentation.cpp:2044)
            buf = np.concatenate([blocks[i] for i in idx]) if idx else np.zeros((0, REC), np.uint8)
            sets.append(unpack_regions(buf))
        (q_rep, q_desc), (t_rep, t_desc) = sets
        nq = len(q_rep)
        lo, hi = (nq * rank) // world, (nq * (rank + 1)) // world
        local = match(det_name, q_rep, q_desc, t_rep, t_desc, lo, hi) if hi > lo and len(t_rep) else np.zeros((0, 7))
        local = np.ascontiguousarray(local, np.float64).reshape(-1, 7)
        if dist is not None:
            parts = _all_gather_padded(local.view(np.uint8).reshape(len(local), 56), dist, device)
            tents = np.concatenate([p.copy().view(np.float64).reshape(-1, 7) for p in parts])
        else:
            tents = local
        out[det_name] = (q_rep, q_desc, t_rep, t_desc, tents)
    return out


def frames_and_keys(groups):
    """Concatenates the per-detector tentatives the way GetCorresponcesVector("All") does (detector names in map order) into the
    14-double frames + ratio keys mb2_host_verify takes."""
    frames, keys = [], []
    for det_name in sorted(groups):
        q_rep, _, t_rep, _, rows = groups[det_name]
        if len(rows) == 0:
            continue
        qi, ti = rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64)
        frames.append(np.concatenate([q_rep[qi], t_rep[ti]], axis=1))
        keys.append(np.abs(np.sqrt((rows[:, 4].astype(np.float32) / rows[:, 5].astype(np.float32)).astype(np.float64))))
    if not frames:
        return np.zeros((0, 14)), np.zeros(0)
    return np.ascontiguousarray(np.concatenate(frames)), np.ascontiguousarray(np.concatenate(keys))


def gpu_workers(ctx, A, B, cfg, shape=None):
    """compute_unit / match closures over one mods_b200.Context (the CUDA path through the C ABI) for pair_views_sharded."""
    import mods_b200 as mb

    def compute(u):
        img = (A, B)[u[0]]
        det = cfg.mser if u[1] == "MSER" else cfg.det
        tilt, phi, zoom = u[3][:3]
        sigma = u[3][3] if len(u[3]) > 3 else 0.5
        return ctx.detect_describe_synth_view(img, tilt, phi, zoom, det=det, ori=cfg.ori, desc=cfg.desc, slot=7, InitSigma=sigma,
                                              shape=shape, capacity=max(4096, (shape[0] * shape[1] if shape else img.size) // 8))

    def match(det_name, q_rep, q_desc, t_rep, t_desc, lo, hi):
        ratio = cfg.mserMatchRatio if det_name == "MSER" else cfg.matchRatio
        rows = ctx.match_fginn(np.ascontiguousarray(q_desc[lo:hi]), t_desc, np.ascontiguousarray(t_rep[:, :2]), ratio=ratio, contradDist=cfg.contradDist)
        rows[:, 0] += lo
        return rows
    return compute, match


def iters_units(w, h, tiers):
    """tiers: {detector name: [(zoom, tilt, phi, InitSigma), ...]} (rows of SetVSPars) -> units of both images + their costs."""
    units = [(im, det, vi, (v[1], v[2], v[0], v[3] if len(v) > 3 else 0.5)) for im in (0, 1) for det in sorted(tiers) for vi, v in enumerate(tiers[det])]
    costs = [view_cost(w, h, abs(u[3][0]), u[3][2]) for u in units]
    return units, costs
