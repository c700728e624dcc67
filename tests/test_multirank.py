"""N>1 host logic on CPU: two gloo ranks time-shard one capture at an inter-block gap, each decodes the block
segments it owns on ITS OWN copy of ITS OWN rows (oracle backend -- test infrastructure; on the GPU box the same
code runs against the CUDA library), and rank 0 gathers the results.  Every block decode of the reference must be
reproduced exactly once, bit-exactly (committed reference digests), with rows and event times re-based."""
import dataclasses
import os
import socket

import numpy as np
import pytest

from conftest import load_capture
from readtape_b200 import abi, evlog, shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, name, backend_lib, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        doc, segs, heads, rows = load_capture(name)
        desc = evlog.desc_from_heads(heads)
        end = np.flatnonzero(rows[:, 0] == -32768)
        nrows = int(end[0]) if len(end) else rows.shape[0]
        lsb = heads["maxvolts"] / 32767.0
        gaps = shard.quiet_gaps(rows[:nrows], thr_lsb=int(0.15 / lsb), min_gap_rows=2000)
        shards = shard.plan(nrows, gaps, world)
        a, b = shards[rank]
        mine = [(i, s) for i, s in enumerate(segs) if s.reset_kind == abi.RT_RESET_FULL and shard.owner(shards, s.row) == rank]
        need_end = max([b] + [(s.end_row if s.end_row >= 0 else nrows) for _, s in mine])
        lib = abi.load(backend_lib)
        tape = lib.open(shard.sub_desc(desc, a))
        tape.upload(rows[a:min(nrows, need_end)])                       # this rank's rows only
        local = []
        rebased = [dataclasses.replace(s, row=s.row - a, end_row=(s.end_row - a if s.end_row >= 0 else -1),
                                       stop_row=(s.stop_row - a if s.stop_row >= 0 else -1)) for _, s in mine]
        for (i, s), (_, got) in zip(mine, evlog.replay(tape, rebased)):
            got = got.copy()
            got["row"] += a                                             # back to reel coordinates
            local.append((i, bool(evlog.matches_fixture(s, got)), int(len(got))))
        tape.close()
        allres = shard.gather({"rank": rank, "range": (a, b), "segments": local}, dist)
        if rank == 0:
            q.put((shards, allres, len([s for s in segs if s.reset_kind == abi.RT_RESET_FULL])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("Microdata_20blks.nm_tap", 2), ("LJS009_part1_39blks", 2)])
def test_two_ranks_shard_one_capture(name, world, oracle_lib):
    import torch.multiprocessing as mp
    load_capture(name)                                                  # skip early if the capture is not staged
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, name, abi.ORACLE_LIB, q)) for r in range(world)]
    for p in procs:
        p.start()
    shards, allres, nfull = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(allres) == world
    assert all(a < b for a, b in shards), f"a rank got no rows: {shards}"
    assert shards[0][0] == 0 and all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
    seen = {}
    for res in allres:
        assert res["segments"], f"rank {res['rank']} decoded nothing"
        for i, ok, n in res["segments"]:
            assert i not in seen, f"segment {i} decoded twice"
            seen[i] = ok
            assert ok, f"segment {i} (rank {res['rank']}) differs from the reference"
    assert len(seen) == nfull


def test_plan_cuts_only_in_gaps():
    rows = np.zeros((64000, 9), dtype="<i2")
    for start in (5000, 25000, 45000):
        rows[start:start + 8000] = (np.sin(np.arange(8000) / 3.0) * 20000).astype("<i2")[:, None]
    gaps = shard.quiet_gaps(rows, thr_lsb=100, min_gap_rows=1000)
    assert [(s // 1000, e // 1000) for s, e in gaps] == [(0, 4), (13, 24), (33, 44), (53, 64)]
    for world in (1, 2, 3, 4, 8):
        sh = shard.plan(rows.shape[0], gaps, world)
        assert sh[0][0] == 0 and sh[-1][1] == rows.shape[0] and len(sh) == world
        for a, b in sh[:-1]:
            if b < rows.shape[0]:
                assert any(s < b < e for s, e in gaps), (world, sh)
                assert b % shard.GRAN == 0


@pytest.mark.parametrize("nparm,nshards,world", [(5, 8, 8), (5, 1, 8), (8, 2, 4), (1, 8, 8), (5, 3, 2), (2, 1, 1)])
def test_parmset_by_shard_units_are_dealt_evenly(nparm, nshards, world):
    """config 4: 5 GCR parameter sets x time shards over 8 GPUs -- every unit exactly once, loads within one unit,
    and a rank touches as few different shards as possible"""
    per_rank = shard.assign_units(nparm, nshards, world)
    flat = [u for r in per_rank for u in r]
    assert sorted(flat) == sorted((p, s) for p in range(nparm) for s in range(nshards))
    loads = [len(r) for r in per_rank]
    if nparm * nshards >= world:
        assert max(loads) - min(loads) <= 1, loads
    for r in per_rank:
        shards_touched = {s for _, s in r}
        assert len(shards_touched) <= (len(r) + nparm - 1) // nparm + 1
