"""Time-sharding of one capture over the ranks of a box.

The scan path shards with no data exchange: every block decode starts from a fresh reset, so a reel can be
cut in inter-block gaps and each piece scanned by its own GPU from its own copy of the rows (this divides the
host->device traffic by the number of ranks).  This module holds the host-side logic of that split:

  quiet_gaps()   all-track quiet stretches, from the same 32-row granule min/max map the device builds (k_ingest.cu)
  plan()         one contiguous row range per rank, cut at the centre of the gap nearest to an even split
  sub_desc()     the tape descriptor of a shard: same stream, first row = `a` (event times stay identical because
                 timenow is derived from integer nanoseconds, readtape.c:1423-1424)
  owner()        which shard a block decode that starts at `row` belongs to
  gather()       the only collective of the path: per-rank results to rank 0 (torch.distributed; NCCL or gloo)

Nothing here touches samples on the device; ranks call the C-ABI (readtape_b200.abi) on their own shard.
"""
from __future__ import annotations

import copy

import numpy as np

GRAN = 32


def quiet_gaps(rows: np.ndarray, thr_lsb: int, min_gap_rows: int):
    """-> list of (first_row, end_row) of stretches where every track's 32-row granules span <= thr_lsb."""
    n = rows.shape[0] // GRAN * GRAN
    if n == 0:
        return []
    g = rows[:n].reshape(n // GRAN, GRAN, rows.shape[1])
    quiet = ((g.max(axis=1).astype(np.int32) - g.min(axis=1).astype(np.int32)) <= thr_lsb).all(axis=1)
    edges = np.flatnonzero(np.diff(np.concatenate(([0], quiet.view(np.int8), [0]))))
    out = []
    for s, e in zip(edges[0::2], edges[1::2]):
        if (e - s) * GRAN >= min_gap_rows:
            out.append((int(s) * GRAN, int(e) * GRAN))
    return out


def plan(nrows: int, gaps, world: int):
    """-> [(a, b)] * world: contiguous, covering [0, nrows); every cut lies at the centre of a quiet gap
    (granule aligned).  With fewer usable gaps than cuts, trailing ranks get empty ranges."""
    centres = sorted({(s + e) // 2 // GRAN * GRAN for s, e in gaps if 0 < (s + e) // 2 < nrows})
    cuts = []
    for k in range(1, world):
        want = nrows * k // world
        cand = [c for c in centres if c > (cuts[-1] if cuts else 0)]
        if not cand:
            cuts.append(nrows)
            continue
        cuts.append(min(cand, key=lambda c: abs(c - want)))
    bounds = [0] + cuts + [nrows]
    return [(bounds[i], max(bounds[i], bounds[i + 1])) for i in range(world)]


def owner(shards, row: int) -> int:
    for r, (a, b) in enumerate(shards):
        if a <= row < b:
            return r
    return len(shards) - 1


def sub_desc(desc, a: int):
    """descriptor of the stream that starts at row `a` of `desc`'s stream"""
    d = copy.copy(desc)
    d.head_to_trk = type(desc.head_to_trk)(*desc.head_to_trk)
    d.tstart_ns = desc.tstart_ns + a * desc.tdelta_ns
    return d


def gather(obj, dist=None, dst: int = 0):
    """per-rank python objects -> list on rank `dst` (None elsewhere); single process: [obj]"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def assign_units(nparmsets: int, nshards: int, world: int):
    """(parameter set, time shard) scan units -> ranks (BASELINE config 4: tracks x parmsets sharded over the GPUs of a box).

    A parameter-set retry is a full independent re-scan of the same rows (readtape.c:1755-1795), so the unit of work is
    (parmset p, time shard s).  Units are dealt round-robin in shard-major order: consecutive ranks take the parameter sets
    of the SAME shard, so that a rank needs as few different shards (= host->device copies) as possible, and the load
    differs by at most one unit.  -> list over ranks of [(p, s), ...]"""
    units = [(p, s) for s in range(nshards) for p in range(nparmsets)]
    out = [[] for _ in range(world)]
    for i, u in enumerate(units):
        out[i * world // len(units) if len(units) >= world else i].append(u)
    return out
