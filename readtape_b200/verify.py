"""Full-scale verification of a whole-tape scan of a PERIODIC tape (bench.py, tests).  CHECKER SIDE: this module drives the CPU
oracle (oracle/scan_oracle.c through the same C-ABI); only tests/, smoke() and bench.py may use it.

The benchmark tape repeats one super-tile.  Every block decode starts from a fresh reset, so the events of tile k must be the
events of tile 1 shifted by (k-1) periods: rt_bulk_tile_digest() gives one order-independent digest and one count per tile,
computed on the device over ALL events of the scan, and the oracle provides the same two numbers for one tile from its own exact
scans (reset at the first row of every unit that owns rows of that tile, as the speculative scan does).
"""
from __future__ import annotations

import numpy as np

from . import abi, evlog


def oracle_tile_digest(oracle_lib, desc, cfg, tile_rows: np.ndarray, period: int, units, tile_index: int = 1):
    """(count, digest) of tile `tile_index` from the CPU oracle.  `units` = [(row0, next_row0)] of the units that own rows of that
    tile (from the product's unit table); `tile_rows` = one period of the tape (host array)."""
    span = tile_index + 2
    host = np.concatenate([tile_rows] * span)
    tape = oracle_lib.open(desc)
    tape.upload(host)
    sc = tape.scan(cfg)
    got = []
    for row0, nxt in units:
        sc.reset(abi.RT_RESET_FULL, row0)
        ev, _ = sc.run(min(nxt, host.shape[0]) - row0)
        got.append(ev)
    sc.end(); tape.close()
    ev = np.concatenate(got) if got else np.zeros(0, dtype=abi.EVENT_DTYPE)
    return evlog.tile_digest(ev, tile_index, period, desc.tstart_ns, desc.tdelta_ns)


def owners_of_tile(bulk, period: int, tile_index: int = 1):
    """[(row0, next_row0)] of the units owning rows of the tile, from the (fetched) unit table"""
    lo, hi = tile_index * period, (tile_index + 1) * period
    first = bulk.unit_info(0, lo)
    out, i = [], first["unit_index"]
    while True:
        ui = bulk.unit_at(0, i)
        if ui is None or ui["row0"] >= hi:
            return out
        nxt = bulk.unit_at(0, i + 1)
        out.append((ui["row0"], nxt["row0"] if nxt is not None else ui["row_end"]))
        i += 1


def verify_periodic(oracle_lib, bulk, desc, cfg, tile_rows: np.ndarray, nrows: int) -> dict:
    """Checks a finished (not yet fetched) rt_bulk_scan of a tape made of repeated `tile_rows`.  Returns a report; report["ok"]
    is False if any complete tile differs from tile 1, any event time is not reproduced exactly, or tile 1 differs from the oracle."""
    period = int(tile_rows.shape[0])
    ntiles = (nrows + period - 1) // period
    whole = nrows // period
    counts, digests, bad = bulk.tile_digest(0, period, ntiles)
    ref = 1 if whole > 1 else 0
    same = [int(counts[i] == counts[ref] and digests[i] == digests[ref]) for i in range(whole)]
    report = {"tiles": int(ntiles), "whole_tiles": int(whole), "tiles_equal_to_tile1": int(sum(same)),
              "events_per_tile": int(counts[ref]), "events_digested": int(counts.sum()), "bad_event_times": int(bad),
              "first_differing_tile": next((i for i, s in enumerate(same) if not s), None)}
    units = owners_of_tile(bulk, period, ref)                   # fetches the results
    n, d = oracle_tile_digest(oracle_lib, desc, cfg, tile_rows, period, units, ref)
    report["oracle_tile"] = {"tile": ref, "units": len(units), "events": n, "equal": bool(n == int(counts[ref]) and d == int(digests[ref]))}
    report["ok"] = bool(sum(same) == whole and bad == 0 and report["oracle_tile"]["equal"])
    return report
