"""GPU parity: the CUDA library (through the C-ABI) against the reference's events and the oracle.

Bit-exact: rows, polarity, event times (f64), voltages (f32) and AGC gains (f32) are compared as
bytes (SHA-256 of the canonical records) -- integer/byte work allows no tolerance, and the float
results are required to be identical too because .tap parity depends on them.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ALL_FIXTURES, load_capture
from readtape_b200 import abi, evlog

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_exact_scan_reproduces_reference_events(name, cuda_lib):
    """rt_scan_* driven through the reference's own reset sequence (all modes, all reset kinds)."""
    doc, segs, heads, rows = load_capture(name)
    tape = cuda_lib.open(evlog.desc_from_heads(heads))
    tape.upload(rows)
    bad = []
    for seg, got in evlog.replay(tape, segs):
        if not evlog.matches_fixture(seg, got):
            bad.append((seg.row, seg.parmset, len(got), seg.nevents))
    tape.close()
    assert not bad, f"{len(bad)}/{len(segs)} segments differ from the reference, first: {bad[:3]}"


@pytest.mark.parametrize("name", ["Microdata_20blks.nm_tap", "1600bpi_ukn_6s", "sf93_8blks", "analog", "tss_4secs"])
def test_cuda_equals_oracle_event_for_event(name, cuda_lib, oracle_lib):
    """Same seeded inputs through both libraries; on a mismatch the first differing event is shown."""
    doc, segs, heads, rows = load_capture(name)
    desc = evlog.desc_from_heads(heads)
    tg, to = cuda_lib.open(desc), oracle_lib.open(desc)
    tg.upload(rows); to.upload(rows)
    assert tg.nrows == to.nrows
    for (seg, a), (_, b) in zip(evlog.replay(tg, segs), evlog.replay(to, segs)):
        if a.tobytes() != b.tobytes():
            k = evlog._first_diff(a, b)
            pytest.fail(f"segment row {seg.row} parmset {seg.parmset}: event #{k}: cuda {a[k] if k < len(a) else None} "
                        f"oracle {b[k] if k < len(b) else None} ({len(a)} vs {len(b)} events)")
    tg.close(); to.close()


def test_ingest_tma_equals_plain_kernel(cuda_lib):
    """K1 with TMA-staged tiles vs the plain-load kernel: identical planes => identical scans."""
    doc, segs, heads, rows = load_capture("Microdata_20blks.nm_tap")
    desc = evlog.desc_from_heads(heads)
    seg = [s for s in segs if s.nevents > 1000][0]
    out = []
    for mode in ("tma", "simple"):
        os.environ["RT_INGEST"] = mode
        tape = cuda_lib.open(desc)
        tape.upload(rows)
        sc = tape.scan(evlog.cfg_for(seg)); sc.reset(abi.RT_RESET_FULL, 0)
        ev, done = sc.run(tape.nrows)
        out.append((tape.nrows, done, ev.tobytes()))
        sc.end(); tape.close()
    os.environ.pop("RT_INGEST", None)
    assert out[0] == out[1]


def test_chunked_upload_and_end_marker(cuda_lib, oracle_lib):
    """Appending in odd-sized pieces gives the same tape; the -32768 end marker in head 0 ends it."""
    doc, segs, heads, rows = load_capture("Microdata_20blks.nm_tap")
    desc = evlog.desc_from_heads(heads)
    rows = rows[:150001].copy()
    rows[140000, 0] = -32768                      # a marker in the middle: everything after it is ignored
    seg = segs[1]
    res = []
    for lib, pieces in ((cuda_lib, [150001]), (cuda_lib, [4096, 10000, 77, 135828]), (oracle_lib, [150001])):
        tape = lib.open(desc)
        at = 0
        for n in pieces:
            tape.upload(rows[at:at + n]); at += n
        assert tape.nrows == 140000
        sc = tape.scan(evlog.cfg_for(seg)); sc.reset(abi.RT_RESET_FULL, 0)
        ev, done = sc.run(10**9)
        assert done == 140000
        res.append(evlog.to_canon(ev).tobytes())
        sc.end(); tape.close()
    assert res[0] == res[1] == res[2]


@pytest.mark.parametrize("name", ["Microdata_20blks.nm_tap", "PLAGO_beginning.nm_tap", "LJS009_part1_39blks", "1kblks_43blks", "tss_4secs"])
def test_bulk_scan_lookup_is_exact(name, cuda_lib):
    """The speculative whole-tape scan: every lookup that HITS must return exactly the events of a
    fresh reset at the reference's real block start; and it must hit for the bulk of the blocks."""
    doc, segs, heads, rows = load_capture(name)
    tape = cuda_lib.open(evlog.desc_from_heads(heads))
    tape.upload(rows)
    full = [s for s in segs if s.reset_kind == abi.RT_RESET_FULL and not (s.flags & abi.RT_F_DENSITY_DETECT)
            and not (s.flags & abi.RT_F_DESKEWING)]
    parmsets_used = sorted({s.parmset for s in full})
    # the deskew pre-pass changes the skew for the main pass: group by (parmset, skew)
    keys = sorted({(s.parmset, tuple(s.skew)) for s in full})
    hits = misses = 0
    why = []
    for key in keys:
        group = [s for s in full if (s.parmset, tuple(s.skew)) == key]
        bulk = tape.bulk_scan([evlog.cfg_for(group[0])])
        st = bulk.stats()
        assert st.units >= 1 and st.rows == tape.nrows
        for seg in group:
            r = bulk.lookup(0, seg.row)
            if r is None:
                misses += 1
                ui = bulk.unit_info(0, seg.row)
                def okpair(s_, l_, need):
                    return s_ is not None and s_ >= need and (l_ is None or l_ < seg.row)
                bad = [k for k in range(len(ui["sync_row"]))
                       if not okpair(ui["sync_row"][k], ui["last_loud_row"][k], ui["need_sync_row"][k])
                       and not okpair(ui["sync_early"][k], ui["loud_early"][k], ui["need_sync_row"][k])]
                why.append(f"row {seg.row}: unit [{ui['row0']},{ui['row_end']}) bad tracks {bad}: " + "; ".join(
                    f"trk{k} first_ev {ui['first_event_row'][k]} sync {ui['sync_row'][k]}/{ui['sync_early'][k]} need {ui['need_sync_row'][k]} "
                    f"loud {ui['last_loud_row'][k]}/{ui['loud_early'][k]}" for k in bad[:3]))
                continue
            ev, valid = r
            end = seg.end_row if seg.end_row >= 0 else tape.nrows
            if seg.row + valid < end:
                misses += 1          # the unit ends before the reference's block does: caller must use the exact scan
                why.append(f"row {seg.row}: unit ends at {seg.row + valid} before the block does ({end})")
                continue
            canon = evlog.to_canon(ev)
            canon = canon[canon["row"] < end]
            if seg.stop_row >= 0:
                canon = canon[canon["row"] <= seg.stop_row]
            assert evlog.matches_fixture(seg, canon), f"bulk lookup at row {seg.row} parmset {seg.parmset} is not exact"
            hits += 1
        bulk.free()
    tape.close()
    print(f"{name}: bulk hits {hits}, misses {misses}")
    for w in why[:8]:
        print("   miss:", w)
    assert hits > 0 and hits >= 0.7 * (hits + misses), \
        f"only {hits} of {hits + misses} block starts were served by the bulk scan; " + " | ".join(why[:4])


def _all_units(bulk):
    out, i = [], 0
    while True:
        ui = bulk.unit_at(0, i)
        if ui is None:
            return out
        out.append(ui); i += 1


@pytest.mark.parametrize("name", ["Microdata_20blks.nm_tap", "PLAGO_beginning", "LJS009_part1_39blks", "1600bpi_ukn_6s", "tss_4secs", "sf93_8blks", "1kblks_43blks"])
def test_fast_kernels_equal_generic_kernel_on_every_unit(name, cuda_lib):
    """K3c (two passes: candidate masks + sparse scan, the default) and K3b (one-pass int16 fast path, RT_SCAN=fast) against
    K3a (exact generic scan, RT_SCAN=generic): same unit table, and for EVERY unit the same events (bit-exact) and the same
    unit-equivalence proof data, for every parameter set / skew the reference used.  K3c also with a mask threshold so high
    that it keeps falling back to its row-by-row mode (RT_SPARSE_T0=1.0) and without the granule map (RT_FAST_SKIP=0)."""
    doc, segs, heads, rows = load_capture(name)
    tape = cuda_lib.open(evlog.desc_from_heads(heads))
    tape.upload(rows)
    full = [s for s in segs if s.reset_kind == abi.RT_RESET_FULL and not (s.flags & abi.RT_F_DENSITY_DETECT)]
    keys = sorted({(s.parmset, tuple(s.skew)) for s in full})
    nunits = 0
    variants = (("generic", "1", None), ("fast", "0", None), ("fast", "1", None), (None, "1", None), (None, "0", None), (None, "1", "1.0"))
    for key in keys:
        cfg = evlog.cfg_for([s for s in full if (s.parmset, tuple(s.skew)) == key][0])
        res = []
        for force, skip, t0 in variants:
            os.environ["RT_FAST_SKIP"] = skip
            for k, v in (("RT_SCAN", force), ("RT_SPARSE_T0", t0)):
                if v:
                    os.environ[k] = v
                else:
                    os.environ.pop(k, None)
            bulk = tape.bulk_scan([cfg])
            units = _all_units(bulk)
            evs = []
            for ui in units:
                r = bulk.lookup(0, ui["row0"])
                evs.append(None if r is None else (r[0].tobytes(), r[1]))
            res.append((units, evs, bulk.stats().events))
            bulk.free()
        for k in ("RT_SCAN", "RT_FAST_SKIP", "RT_SPARSE_T0"):
            os.environ.pop(k, None)
        (ug, eg, ng) = res[0]
        for vi in (1, 3, 4, 5):                                    # walking / deriving every row: identical in everything
            (uf, ef, nf) = res[vi]
            assert len(ug) == len(uf) and ng == nf, (variants[vi], len(ug), len(uf), ng, nf)
            for a, b, x, y in zip(ug, uf, eg, ef):
                assert a == b, f"{variants[vi]}: proof data differ for unit {a['unit_index']} [{a['row0']},{a['row_end']}): generic {a} other {b}"
                assert x == y, f"{variants[vi]}: events differ for unit {a['unit_index']} [{a['row0']},{a['row_end']})"
        (us, es, ns) = res[2]
        assert len(ug) == len(us) and ng == ns
        for a, b in zip(ug, us):                                   # K3b jumping over quiet stretches: same events, sound proof data
            assert not any(b["failed"]), f"unit {b['unit_index']}: failed flags {b['failed']}"
            for key in ("row0", "row_end", "first_event_row", "nevents", "quiet_from"):
                assert a[key] == b[key], (key, a, b)
        n_same = sum(1 for x, y in zip(eg, es) if x == y)
        assert n_same >= 0.9 * len(eg), f"only {n_same} of {len(eg)} unit lookups agree with gap skipping on"
        for x, y in zip(eg, es):
            assert y is None or x is None or x == y, "a lookup that hits must return the same events"
        nunits += len(ug)
    tape.close()
    assert nunits > 0


@pytest.mark.parametrize("name", ["Microdata_20blks.nm_tap", "PLAGO_beginning.nm_tap", "LJS009_part1_39blks", "1600bpi_ukn_6s", "tss_4secs"])
def test_data_driven_mask_thresholds_keep_the_scan_sparse(name, cuda_lib):
    """the per-track mask thresholds chosen from the span histogram: on real captures (NRZI and PE, 0.5 .. 4 V peak to peak) the
    two-pass scan must stay out of its row-by-row mode for all but a few percent of the rows (a fixed fraction of the default-state
    threshold puts 38 % of LJS009 there), whereas a threshold above every bound (RT_SPARSE_T0=8) walks nearly everything"""
    doc, segs, heads, rows = load_capture(name)
    tape = cuda_lib.open(evlog.desc_from_heads(heads))
    tape.upload(rows)
    seg = [s for s in segs if s.reset_kind == abi.RT_RESET_FULL and not (s.flags & abi.RT_F_DENSITY_DETECT)][0]
    cfg = evlog.cfg_for(seg)
    bulk = tape.bulk_scan([cfg]); st = bulk.stats(); bulk.free()
    assert st.ms_masks > 0, "the two-pass scan was not used"
    dense = st.rows_scanned / float(st.track_samples)
    os.environ["RT_SPARSE_T0"] = "8"
    try:
        bulk = tape.bulk_scan([cfg]); st2 = bulk.stats(); bulk.free()
    finally:
        os.environ.pop("RT_SPARSE_T0", None)
    tape.close()
    print(f"{name}: dense-mode rows {100 * dense:.2f} %, with an unreachable threshold {100 * st2.rows_scanned / float(st2.track_samples):.1f} %")
    assert st2.events == st.events
    assert dense < 0.03, f"{100 * dense:.1f} % of the track-rows were walked one by one"
    assert st2.rows_scanned > 10 * max(st.rows_scanned, 1) or st2.rows_scanned > 0.3 * st2.track_samples


@pytest.mark.parametrize("name", ["Microdata_20blks", "LJS009_part1_39blks", "tss_4secs"])
def test_peak_mask_kernel_equals_definition(name, cuda_lib, oracle_lib):
    """phase A of the two-pass scan (int16x2 SIMD in registers) against the oracle's brute-force definition of the two bit planes,
    on real captures, for the window widths of every built-in parameter set and a few odd ones"""
    from readtape_b200 import parmsets
    doc, segs, heads, rows = load_capture(name)
    desc = evlog.desc_from_heads(heads)
    tg, to = cuda_lib.open(desc), oracle_lib.open(desc)
    tg.upload(rows); to.upload(rows)
    base = [s for s in segs if s.reset_kind == abi.RT_RESET_FULL and not (s.flags & abi.RT_F_DENSITY_DETECT)][0]
    table = parmsets.BUILTIN[base.mode]
    widths = set()
    for pi, scale in [(p, 1.0) for p in range(len(table))] + [(0, 0.3), (0, 0.45), (0, 2.2), (0, 3.9)]:
        cfg = abi.make_cfg(base.mode, table[pi], base.bpi * scale, base.ips, flags=base.flags, skew=base.skew)
        w = cuda_lib.L.rt_pkww_width(C.byref(cfg), desc.tdelta_ns)
        if w < 3 or (w, table[pi]["pkww_rise"]) in widths:
            continue
        widths.add((w, table[pi]["pkww_rise"]))
        for frac in (0.25, 1.0):
            cg, ag, t0g = tg.peak_masks(cfg, frac)
            co, ao, t0o = to.peak_masks(cfg, frac)
            assert t0g == t0o and t0g >= 16
            assert np.array_equal(cg, co), f"cand differs: width {w} T0 {t0g}"
            assert np.array_equal(ag, ao), f"acan differs: width {w} T0 {t0g}"
            assert ao.any() and (frac > 0.5 or scale != 1.0 or co.any())
    tg.close(); to.close()
    assert len(widths) >= 4, widths


def test_fast_kernel_on_synthetic_tape_equals_oracle(cuda_lib, oracle_lib):
    """size-independent check on the benchmark's own generator: every unit of a 24-block synthetic tape, each
    compared with the oracle's exact scan from a fresh reset at the unit's first row"""
    from readtape_b200 import parmsets, synth, tbin
    hdr, rows = synth.nrzi_tape(nblocks=24, seed=3)
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    tg, to = cuda_lib.open(desc), oracle_lib.open(desc)
    tg.upload(rows); to.upload(rows)
    for pi, bpi, mode, skew in ((0, hdr.bpi, tbin.MODE_NRZI, None), (5, hdr.bpi, tbin.MODE_NRZI, None), (1, 556.0, tbin.MODE_NRZI, [1, 0, 5, 2, 0, 9, 3, 0, 17]),
                                (0, 200.0, tbin.MODE_NRZI, None), (2, 400.0, tbin.MODE_PE, [0, 3, 0, 0, 7, 0, 0, 1, 0]), (0, 1600.0, tbin.MODE_PE, None)):
        table = parmsets.NRZI if mode == tbin.MODE_NRZI else parmsets.PE
        cfg = abi.make_cfg(mode, table[pi], bpi, hdr.ips, skew=skew)       # other densities: window widths 6..50, ring sizes 64 / 128
        bulk = tg.bulk_scan([cfg])
        units = _all_units(bulk)
        assert len(units) >= 1
        sc = to.scan(cfg)
        for ui in units:
            r = bulk.lookup(0, ui["row0"])
            assert r is not None
            ev, valid = r
            sc.reset(abi.RT_RESET_FULL, ui["row0"])
            want, _ = sc.run(valid)
            a, b = evlog.to_canon(ev), evlog.to_canon(want)
            # a lookup may chain into the next unit when this one is empty; compare the rows both cover
            assert a.tobytes() == b.tobytes(), f"unit {ui['unit_index']} parmset {pi}: {len(a)} vs {len(b)} events, first diff {evlog._first_diff(a, b)}"
        sc.end(); bulk.free()
    tg.close(); to.close()


def test_streamed_host_scan_equals_plain_sequence(cuda_lib):
    """rt_bulk_scan_host(): the first call on a tape object has no capacity history and runs upload + scan + fetch one after
    the other; the second one streams (scans segment by segment while the copy is still running).  Both must give the
    same units, proof data and events as the separate calls."""
    import ctypes
    from readtape_b200 import parmsets, synth, tbin
    tile = synth.nrzi_tile()
    reps = 14                                                       # 17.5 M rows: above the streaming threshold
    n = tile.shape[0] * reps
    hdr = synth.nrzi_header()
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[0], hdr.bpi, hdr.ips)
    L = cuda_lib.L
    hptr = L.rt_host_alloc(n * 18)
    assert hptr
    hbuf = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_int16)), shape=(n, 9))
    for i in range(reps):
        hbuf[i * tile.shape[0]:(i + 1) * tile.shape[0]] = tile
    tape = cuda_lib.open(desc)

    def snapshot(bulk):
        units = _all_units(bulk)
        sample = units[::max(1, len(units) // 40)]
        evs = [bulk.lookup(0, u["row0"]) for u in sample]
        return units, [(e[0].tobytes(), e[1]) for e in evs], bulk.stats()

    tape.upload_ptr(hptr, n)
    ref = tape.bulk_scan([cfg]); ref.fetch()
    u0, e0, s0 = snapshot(ref); ref.free()
    b1 = tape.bulk_scan_host(hptr, n, cfg)
    u1, e1, s1 = snapshot(b1); b1.free()
    b2 = tape.bulk_scan_host(hptr, n, cfg)
    u2, e2, s2 = snapshot(b2); b2.free()
    tape.close(); L.rt_host_free(hptr)
    assert s2.pad >= 2, f"the second call did not stream (segments {s2.pad})"
    assert s0.events == s1.events == s2.events and s0.units == s1.units == s2.units
    assert u0 == u1 == u2
    assert e0 == e1 == e2


@pytest.mark.parametrize("name", ["1600bpi_ukn_6s", "LJS009_part1_39blks"])
def test_pe_parmset_fanout_on_one_gpu(name, cuda_lib, oracle_lib):
    """BASELINE config 3: all 8 built-in PE parameter sets (parmsets.c:80-87) scanned in ONE rt_bulk_scan call; every lookup
    at a block start of the reference must equal the oracle's exact scan from a fresh reset there, for every parameter set."""
    from readtape_b200 import parmsets
    doc, segs, heads, rows = load_capture(name)
    desc = evlog.desc_from_heads(heads)
    tg, to = cuda_lib.open(desc), oracle_lib.open(desc)
    tg.upload(rows); to.upload(rows)
    base = [s for s in segs if s.reset_kind == abi.RT_RESET_FULL and not (s.flags & abi.RT_F_DENSITY_DETECT) and s.parmset == 0]
    assert base
    cfgs = [abi.make_cfg(base[0].mode, parmsets.PE[p], base[0].bpi, base[0].ips, flags=base[0].flags, skew=base[0].skew) for p in range(8)]
    bulk = tg.bulk_scan(cfgs)
    st = bulk.stats()
    assert st.track_samples == tg.nrows * desc.ntrks * 8
    hits = total = 0
    for p, cfg in enumerate(cfgs):
        sc = to.scan(cfg)
        for seg in base[:12]:
            total += 1
            r = bulk.lookup(p, seg.row)
            if r is None:
                continue
            ev, valid = r
            end = seg.end_row if seg.end_row >= 0 else tg.nrows
            span = min(valid, end - seg.row)
            sc.reset(abi.RT_RESET_FULL, seg.row)
            want, _ = sc.run(span)
            got = evlog.to_canon(ev); got = got[got["row"] < seg.row + span]
            assert got.tobytes() == evlog.to_canon(want).tobytes(), f"parmset {p} block at row {seg.row}: lookup differs from the oracle"
            hits += 1
        sc.end()
    bulk.free(); tg.close(); to.close()
    assert hits >= 0.7 * total, f"only {hits} of {total} (parmset, block) lookups were served by the bulk scan"


def test_tile_digest_at_scale_equals_oracle(cuda_lib, oracle_lib):
    """rt_bulk_tile_digest(): on a periodic tape every tile must carry the same events (count + order-independent digest, computed
    on the device over ALL events), every double-precision event time must be reproduced by the reference's expression, and the
    digest of one tile must equal the one computed from the CPU oracle's events -- the check bench.py applies to the full-size tape"""
    from readtape_b200 import parmsets, synth, tbin, verify
    tile = synth.nrzi_tile()
    reps = 5
    hdr = synth.nrzi_header()
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[0], hdr.bpi, hdr.ips)
    tape = cuda_lib.open(desc)
    for _ in range(reps):
        tape.upload(tile)
    nrows = tile.shape[0] * reps
    bulk = tape.bulk_scan([cfg])
    rep = verify.verify_periodic(oracle_lib, bulk, desc, cfg, tile, nrows)
    st = bulk.stats()
    bulk.free()
    print(rep)
    assert rep["ok"], rep
    assert rep["events_digested"] == st.events or rep["events_digested"] <= st.events       # overlap events are counted once
    # a corrupted tape must be noticed: one sample of tile 3 changed
    bad = tile.copy(); bad[700_000, 4] = 20000
    tape.clear()
    for i in range(reps):
        tape.upload(bad if i == 3 else tile)
    bulk = tape.bulk_scan([cfg])
    rep2 = verify.verify_periodic(oracle_lib, bulk, desc, cfg, tile, nrows)
    bulk.free(); tape.close()
    assert not rep2["ok"] and rep2["first_differing_tile"] == 3, rep2


def test_decodable_gcr_tape_cuda_equals_oracle_and_digest(cuda_lib, oracle_lib):
    """BASELINE config 4's generator (synth.gcr_tile: decodable 6250 BPI GCR blocks, zero-crossing detector): the whole-tape scan of all 5
    GCR parameter sets in one call; every unit equals the oracle's exact scan for two of them, and the tile digest check holds"""
    from readtape_b200 import parmsets, synth, tbin, verify
    tile = synth.gcr_tile(nblocks=3)
    hdr = synth.gcr_header()
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    cfgs = [abi.make_cfg(tbin.MODE_GCR, p, hdr.bpi, hdr.ips, flags=abi.RT_F_FIND_ZEROS) for p in parmsets.GCR]
    tg, to = cuda_lib.open(desc), oracle_lib.open(desc)
    rows = np.concatenate([tile] * 3)
    tg.upload(rows); to.upload(rows)
    bulk = tg.bulk_scan(cfgs)
    st = bulk.stats()
    assert st.track_samples == tg.nrows * 9 * 5 and st.events > 5 * 9 * 9 * 3000
    for ci in (0, 3):
        sc = to.scan(cfgs[ci])
        i = 0
        while True:
            ui = bulk.unit_at(ci, i)
            if ui is None:
                break
            r = bulk.lookup(ci, ui["row0"])
            assert r is not None
            ev, valid = r
            sc.reset(abi.RT_RESET_FULL, ui["row0"])
            want, _ = sc.run(valid)
            assert evlog.to_canon(ev).tobytes() == evlog.to_canon(want).tobytes(), f"parameter set {ci} unit {i}"
            i += 1
        sc.end()
        assert i >= 9
    bulk.free()
    bulk = tg.bulk_scan(cfgs[:1])
    rep = verify.verify_periodic(oracle_lib, bulk, desc, cfgs[0], tile, tg.nrows)
    bulk.free(); tg.close(); to.close()
    assert rep["ok"], rep


def test_fused_ingest_masks_equal_the_separate_pass(cuda_lib, capfd):
    """rt_prepare(): the ingest kernel writes the candidate / canonical bit planes while a tile is on chip.  The planes must equal the
    separate mask pass word for word (RT_PREMASK_CHECK compares them on every adopting scan) and the scan results must not change:
    synthetic tape attached in one piece and uploaded in odd chunks, and the 9-track NRZI captures."""
    from readtape_b200 import parmsets, synth, tbin
    import torch
    tile = synth.nrzi_tile()
    hdr = synth.nrzi_header()
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[0], hdr.bpi, hdr.ips)
    rows = np.concatenate([tile] * 8)[:9_700_123]                    # not a whole number of ingest tiles
    os.environ["RT_PREMASK_CHECK"] = "1"
    os.environ["RT_FUSED_MASKS"] = "1"
    try:
        def snapshot(tape):
            bulk = tape.bulk_scan([cfg]); st = bulk.stats()
            units = _all_units(bulk)
            evs = [bulk.lookup(0, u["row0"]) for u in units[::7]]
            bulk.free()
            return st.events, st.units, units, [(e[0].tobytes(), e[1]) for e in evs]
        plain = cuda_lib.open(desc); plain.upload(rows); want = snapshot(plain); plain.close()
        # (a) device-resident rows attached in one call, twice (the second attach has the thresholds from the start)
        t1 = cuda_lib.open(desc); t1.prepare(cfg)
        dev = torch.from_numpy(rows).cuda()
        for _ in range(2):
            t1.clear(); t1.attach_device(dev.data_ptr(), rows.shape[0])
            assert snapshot(t1) == want
        t1.close()
        # (b) host rows in chunks of odd sizes
        t2 = cuda_lib.open(desc); t2.prepare(cfg)
        at = 0
        for n in (3_000_000, 2048 * 700, 1_234_567, 10**9):
            t2.upload(rows[at:at + n]); at += n
        assert snapshot(t2) == want
        t2.close()
        err = capfd.readouterr().err
        assert err.count("fused mask planes identical") >= 2, err[-800:]    # (b) regrows the planes, which restarts the mask planes: no adoption there
        fused = [int(x.split(",")[1].split()[0]) for x in err.split("identical to the separate pass (")[1:]]
        assert max(fused) > 4_000_000, fused
        # (c) real captures (NRZI 800 BPI, window 13)
        for name in ("Microdata_20blks.nm_tap", "PLAGO_beginning.nm_tap"):
            doc, segs, heads, crow = load_capture(name)
            d2 = evlog.desc_from_heads(heads)
            seg = [s for s in segs if s.reset_kind == abi.RT_RESET_FULL and not (s.flags & abi.RT_F_DENSITY_DETECT)][0]
            c2 = evlog.cfg_for(seg)
            res = []
            for prep in (False, True):
                tp = cuda_lib.open(d2)
                if prep:
                    tp.prepare(c2)
                tp.upload(crow)
                bulk = tp.bulk_scan([c2]); st = bulk.stats(); units = _all_units(bulk)
                res.append((st.events, units, [bulk.lookup(0, u["row0"])[0].tobytes() for u in units]))
                bulk.free(); tp.close()
            assert res[0] == res[1], name
    finally:
        os.environ.pop("RT_PREMASK_CHECK", None); os.environ.pop("RT_FUSED_MASKS", None)


@pytest.mark.parametrize("name", ["Microdata_20blks.nm_tap", "LJS009_part1_39blks", "sf93_8blks"])
def test_invert_is_served_by_the_fast_kernels_and_is_exact(name, cuda_lib, oracle_lib):
    """-invert (readtape.c:1421): the whole-tape scan runs its fast kernels on a negated copy of the planes.  (1) Same events as the
    scan of a tape that was uploaded negated, without the flag (volts(-x) == -volts(x) exactly); (2) every lookup equals the CPU
    oracle's exact inverted scan from a fresh reset at that row."""
    doc, segs, heads, rows = load_capture(name)
    desc = evlog.desc_from_heads(heads)
    full = [s for s in segs if s.reset_kind == abi.RT_RESET_FULL and not (s.flags & (abi.RT_F_DENSITY_DETECT | abi.RT_F_DESKEWING))]
    seg0 = full[0]
    plain = evlog.cfg_for(seg0)
    inv = evlog.cfg_for(seg0); inv.flags |= abi.RT_F_INVERT
    tape = cuda_lib.open(desc); tape.upload(rows)
    bulk = tape.bulk_scan([inv])
    st = bulk.stats()
    neg = cuda_lib.open(desc); neg.upload((-rows.astype(np.int32)).clip(-32767, 32767).astype(np.int16) if (rows != -32768).all() else rows)
    bulk_neg = neg.bulk_scan([plain])
    st_neg = bulk_neg.stats()
    assert (st.units, st.events) == (st_neg.units, st_neg.events) and st.events > 10000
    assert st.two_pass == st_neg.two_pass                       # the same kernels served both
    ev_a, dg_a, bad_a = bulk.tile_digest(0, tape.nrows, 1)
    ev_b, dg_b, bad_b = bulk_neg.tile_digest(0, tape.nrows, 1)
    assert (int(ev_a[0]), int(dg_a[0]), bad_a) == (int(ev_b[0]), int(dg_b[0]), bad_b) and bad_a == 0
    otape = oracle_lib.open(desc); otape.upload(rows)
    osc = otape.scan(inv)
    hits = 0
    for seg in [s for s in full if (s.parmset, tuple(s.skew)) == (seg0.parmset, tuple(seg0.skew))][:6]:
        r = bulk.lookup(0, seg.row)
        if r is None:
            continue
        ev, valid = r
        n = min(valid, 200000)
        osc.reset(abi.RT_RESET_FULL, seg.row)
        want, _ = osc.run(n)
        got = evlog.to_canon(ev); got = got[got["row"] < seg.row + n]
        assert got.tobytes() == evlog.to_canon(want).tobytes(), f"inverted lookup at row {seg.row} differs from the oracle"
        hits += 1
    assert hits >= 1
    osc.end(); otape.close(); bulk.free(); bulk_neg.free(); tape.close(); neg.close()


def test_ingest_and_mask_kernels_side_by_side_give_the_same_planes(cuda_lib, capfd):
    """RT_FUSED_MASKS=2 (rt_prepare): the mask kernel of chunk i runs on a second stream while the ingest kernel of chunk i+1 runs.
    The planes must equal the separate pass word for word (RT_PREMASK_CHECK) and the scan results must not change."""
    from readtape_b200 import parmsets, synth, tbin
    import torch
    tile = synth.nrzi_tile()
    hdr = synth.nrzi_header()
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[0], hdr.bpi, hdr.ips)
    rows = np.concatenate([tile] * 8)[:9_700_123]
    os.environ["RT_PREMASK_CHECK"] = "1"
    os.environ["RT_FUSED_MASKS"] = "2"
    os.environ["RT_OVERLAP_CHUNK"] = str(1 << 20)                  # 1 Mi-row chunks: several hand-overs between the two streams
    try:
        def snapshot(tape):
            bulk = tape.bulk_scan([cfg]); st = bulk.stats()
            units = _all_units(bulk)
            evs = [bulk.lookup(0, u["row0"]) for u in units[::7]]
            fusedflag = st.masks_fused
            bulk.free()
            return (st.events, st.units, units, [(e[0].tobytes(), e[1]) for e in evs]), fusedflag
        plain = cuda_lib.open(desc); plain.upload(rows); want, _ = snapshot(plain); plain.close()
        t1 = cuda_lib.open(desc); t1.prepare(cfg)
        dev = torch.from_numpy(rows).cuda()
        for _ in range(2):
            t1.clear(); t1.attach_device(dev.data_ptr(), rows.shape[0])
            got, fusedflag = snapshot(t1)
            assert got == want and fusedflag == 1
        t1.close()
        err = capfd.readouterr().err
        assert err.count("fused mask planes identical") >= 2, err[-800:]
    finally:
        for k in ("RT_PREMASK_CHECK", "RT_FUSED_MASKS", "RT_OVERLAP_CHUNK"):
            os.environ.pop(k, None)
