"""The CPU oracle against the reference's own events (committed digests), on every staged capture.

The digests in tests/golden/*.segments.json were produced by the UNMODIFIED reference
(readtape 3.18, oracle/evdump_shim.c) -- see oracle/make_golden.py.  This is what pins the oracle.
"""
import numpy as np
import pytest

from conftest import ALL_FIXTURES, load_capture
from readtape_b200 import abi, evlog


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_oracle_reproduces_reference_events(name, oracle_lib):
    doc, segs, heads, rows = load_capture(name)
    tape = oracle_lib.open(evlog.desc_from_heads(heads))
    tape.upload(rows)
    nseg = 0
    for seg, got in evlog.replay(tape, segs):
        assert evlog.matches_fixture(seg, got), \
            f"segment at row {seg.row} parmset {seg.parmset}: {len(got)} events, reference {seg.nevents}, digest differs"
        nseg += 1
    tape.close()
    assert nseg == len(segs)


def test_rewind_and_chunked_runs_are_consistent(oracle_lib):
    """rt_scan_run in pieces == in one go; rt_scan_rewind restores the state exactly (Whirlwind protocol)."""
    doc, segs, heads, rows = load_capture("Microdata_20blks.nm_tap")
    tape = oracle_lib.open(evlog.desc_from_heads(heads))
    tape.upload(rows)
    seg = [s for s in segs if s.nevents > 1000][0]
    cfg = evlog.cfg_for(seg)
    a = tape.scan(cfg); a.reset(abi.RT_RESET_FULL, seg.row)
    whole, _ = a.run(20000)
    b = tape.scan(cfg); b.reset(abi.RT_RESET_FULL, seg.row)
    parts = [b.run(7000)[0]]
    b.rewind(seg.row + 3000)
    assert b.pos == seg.row + 3000
    parts = [parts[0][parts[0]["row"] < seg.row + 3000], b.run(17000)[0]]
    assert np.array_equal(np.concatenate(parts), whole)
    a.end(); b.end(); tape.close()
