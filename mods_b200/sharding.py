"""Host-side sharding of independent units (synthesised views of an image, or whole pairs) over ranks.

SURVEY.md 8e: every (image, detector, view) is independent through detection -> description; the reference's
own OpenMP unit (imagerepresentation.cpp:621).  Units are dealt out statically, heaviest first, and the
per-unit results are put back in unit order after the exchange, because region identity in MODS is the
position in RegionVectorMap[det][desc] (views appended in view-index order, imagerepresentation.cpp:2044-2045).
"""
import numpy as np


def assign_units(costs, world):
    """Longest-processing-time-first: returns a list (per rank) of unit indices, deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(u) for u in out]


def view_cost(w, h, tilt=1.0, zoom=1.0):
    """Pixel count of a synthesised view: tilt t shrinks one side by 1/t, zoom z both by z (synth-detection.cpp:301-342)."""
    return (w * zoom) * (h * zoom) / max(tilt, 1e-9)


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
