"""Hit rate of the speculative whole-tape scan on the WHOLE zero-crossing (GCR) captures, modelled on the host.

The GPU test tests/test_host_shim.py::test_readtape_b200_whole_capture_matches_reference_golden requires >= 95 % of the
block decodes to be served by the speculative scan.  Whether a decode is served is decided by host logic
(readtape_b200/csrc/lookup_rules.h, called by rt_bulk_lookup; the host build compiles the same header) from three ingredients that all exist
without a GPU: the unit finder (k_units.cu: quiet granules, restated here in numpy), the proof data the scan kernel
leaves behind (the host build of scan_zc.cuh: the same __host__ __device__ code) and the reference's own reset rows
(the instrumented unmodified reference, oracle/_ref/readtape_evdump, run on the whole capture).  So a change of the
proof rule -- like the end-of-round-2 tightening for the zero-crossing detectors ("the unit must have been quiet since
its own first row") -- can be checked for its cost in hits here, before a GPU is at hand.  On both bundled GCR captures
every reset row of the reference is served (49 / 49 and 15 / 15), which is also what the GPU run recorded
(profiles/hitrates_r02.json).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, capture_path
from oracle import captures
from readtape_b200 import abi, evlog, tbin
from test_fast_host import fast_host, make_planes  # noqa: F401  (fixture)
import test_proof_host as proof

EVDUMP = os.path.join(ROOT, "oracle", "_ref", "readtape_evdump")
GRAN = 32                     # RT_GRAN (rt_dev.h)
ZEROCROSS_PEAK = 0.2          # RT_ZEROCROSS_PEAK (rt_dev.h), decoder.h:138


def find_units(planes, ntrks, nrows, desc, cfg):
    """k_units.cu for RT_DET_ZC with the parameters of plan_scan (rt_api.cu): -> [(row0, row_end)]"""
    lsb = desc.maxvolts / 32767.0
    thr = int(0.9 * ZEROCROSS_PEAK / lsb)
    dt = desc.tdelta_ns * 1e-9
    rows_per_bit = 1.0 / (cfg.bpi * cfg.ips * dt)
    min_gap = max(2, (int(6.0 * rows_per_bit) + 1 + GRAN - 1) // GRAN + 1)
    tail_rows = int(16.0 * rows_per_bit) + int(200e-6 / dt) + 1 + 64 + 50 + 50
    ngran = nrows // GRAN
    g = planes[:, :ngran * GRAN].reshape(ntrks, ngran, GRAN)
    quiet = ((g.max(axis=2) <= thr) & (g.min(axis=2) >= -thr)).all(axis=0)
    starts = np.flatnonzero(quiet & ~np.concatenate([[False], quiet[:-1]]))
    row0s = [0] + [int(s) * GRAN for s in starts if s != 0 and s + min_gap <= ngran and quiet[s:s + min_gap].all()]
    return [(r, min(nrows, row0s[i + 1] + tail_rows) if i + 1 < len(row0s) else nrows) for i, r in enumerate(row0s)]


@pytest.mark.parametrize("name", ["sf93_8blks", "1kblks_43blks"])
def test_whole_gcr_capture_is_served_by_the_speculative_scan(name, fast_host, tmp_path):
    if not os.path.exists(EVDUMP):
        pytest.skip("oracle/_ref/readtape_evdump was not built (python -c 'import __graft_entry__ as g; g.build()' in the build container)")
    L = proof._bind(fast_host)
    src = capture_path(name, full=True)
    evfile = str(tmp_path / "ev")
    env = dict(os.environ); env["RT_EVDUMP"] = evfile
    subprocess.run([EVDUMP] + captures.BY_NAME[name][2].split() + [f"-outf={tmp_path / 'out'}", src], stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT,
                   env=env, cwd=str(tmp_path), check=True)
    heads = evlog.parse_heads(evfile)
    resets = [s for s in evlog.parse(evfile) if s.reset_kind == abi.RT_RESET_FULL]
    _, rows = tbin.read_tbin(src, nheads=heads["nheads"]); rows = np.asarray(rows)
    desc = evlog.desc_from_heads(heads)
    nrows = rows.shape[0]
    end = np.flatnonzero(rows[:, 0] == -32768)
    if len(end): nrows = int(end[0])
    planes, stride = make_planes(rows[:nrows], desc)
    units = find_units(planes, desc.ntrks, nrows, desc, evlog.cfg_for(resets[0]))
    metas = {}
    def meta_of(i, cfg):
        key = (i, bytes(cfg))
        if key not in metas:
            metas[key] = proof.scan_unit(L, "zc", planes, stride, nrows, desc, cfg, units[i][0], units[i][1], 0.25)
        return metas[key]
    hits = restarts = chained = 0; missed = []
    for seg in resets:
        cfg = evlog.cfg_for(seg); s = seg.row
        lo = max(i for i, u in enumerate(units) if u[0] <= s)
        at = None
        for i in (lo, lo + 1):                                    # rt_bulk_lookup: this unit, else the next one, else the tail rule
            if i >= len(units) or at is not None: continue
            m = meta_of(i, cfg)[1]
            if m is not None and proof.covers(L, desc, cfg, m, units[i][0], units[i][1], s)[0]: at = i
        if at is None:
            m = meta_of(lo, cfg)[1]
            if m is not None and proof.tail_covers(L, desc, m, units[lo][0], units[lo][1], s): hits += 1
            else: missed.append(s)
            continue
        hits += 1
        # chaining of event-free units, then: do the events on offer reach the row where the reference's block ended?  If not, the
        # product carries on with the exact scan from there (a "restart": a hit that does not count towards the 95 %)
        start = s
        while at + 1 < len(units) and meta_of(at + 1, cfg)[1] is not None and proof.chains(L, desc, cfg, meta_of(at, cfg)[1], units[at][0], units[at][1], meta_of(at + 1, cfg)[1], start):
            at += 1; start = units[at][0]; chained += 1
        need = seg.stop_row if seg.stop_row >= 0 else (seg.end_row if seg.end_row >= 0 else nrows) - 1
        if units[at][1] <= need: restarts += 1
        # and the events on offer are the reference's (the GPU tests check this for the product; here for the model itself)
        ev = evlog.to_canon(meta_of(at, cfg)[0])
        ev = ev[ev["row"] <= need] if len(ev) else ev
        msg = evlog.compare(seg, ev) if units[at][1] > need else None
        assert msg is None, f"{name}: reset at row {s}, unit {units[at]}: {msg}"
    print(f"[hit-rate model] {name}: {len(units)} units, {hits} of {len(resets)} reset rows of the reference served ({chained} chain steps, {restarts} continued exactly); missed: {missed[:8]}")
    assert len(resets) >= 10 and hits - restarts >= 0.95 * len(resets), (name, hits, restarts, len(resets), missed[:8])
