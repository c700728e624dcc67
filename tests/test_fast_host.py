"""The int16 fast-path scan (readtape_b200/csrc/scan_fast.cuh) built for the HOST, against the reference.

The CUDA fast kernel and this host build instantiate the same __host__ __device__ code, so the algorithm
(van Herk sliding max/min on packed int16, lazy-minimum recurrence, integer pre-filter, candidate handling,
filling-window and deskew corner cases) is checked here without a GPU: every eligible decode segment of the
reference (fresh init_trackstate at the reference's own block-start row) must give exactly the reference's
events (committed digests, tests/golden/*.segments.json).  The GPU tests then check kernel == this.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_capture
from readtape_b200 import abi, evlog, parmsets, synth, tbin

HOST_LIB = os.path.join(ROOT, "tests", "host_fast", "_build", "libfast_host.so")


@pytest.fixture(scope="session")
def fast_host():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host_fast")], stdout=subprocess.DEVNULL)
    L = C.CDLL(HOST_LIB)
    L.fast_host_scan_unit.restype = C.c_int
    L.fast_host_scan_unit.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(abi.TapeDesc), C.POINTER(abi.ScanCfg),
                                      C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    return L


def make_planes(rows, desc):
    """track-major planes with the head->track permutation applied and slack behind the last row"""
    n = rows.shape[0]
    stride = (n + 2048 + 63) // 64 * 64
    planes = np.zeros((desc.ntrks, stride), dtype="<i2")
    for h in range(desc.nheads):
        k = desc.head_to_trk[h]
        if 0 <= k < desc.ntrks:
            planes[k, :n] = rows[:, h]
    return planes, stride


def fast_scan(L, planes, stride, nrows, desc, cfg, row0, row_end, cap=1 << 16, skip=1):
    nt = desc.ntrks
    out = np.zeros((nt, cap), dtype=abi.EVENT_DTYPE)
    counts = np.zeros(nt, dtype=np.uint32)
    meta = np.zeros((nt, L.fast_host_meta_size()), dtype=np.uint8)
    rc = L.fast_host_scan_unit(planes.ctypes.data, stride, nrows, C.byref(desc), C.byref(cfg), row0, row_end,
                               out.ctypes.data, cap, counts.ctypes.data, meta.ctypes.data, skip)
    if rc != 0:
        return None
    assert counts.max(initial=0) <= cap
    ev = np.concatenate([out[k, :counts[k]] for k in range(nt)])
    order = np.lexsort((ev["trk"], ev["row"]))
    fast_scan.failed = [int(np.frombuffer(meta[k].tobytes(), dtype='<u4')[-2]) for k in range(nt)]
    fast_scan.skipped = [int(np.frombuffer(meta[k].tobytes(), dtype='<u4')[-1]) for k in range(nt)]
    return ev[order]


ELIGIBLE = ["sf93_8blks", "1kblks_43blks", "Microdata_20blks.nm_tap", "Microdata_20blks", "PLAGO_beginning.nm_tap", "PLAGO_beginning", "1600bpi_ukn_6s",
            "LJS009_part1_39blks", "SRI_SDS_102715028_4secs", "tss_4secs"]


@pytest.mark.parametrize("name", ELIGIBLE)
def test_fast_path_reproduces_reference_events(name, fast_host):
    doc, segs, heads, rows = load_capture(name)
    desc = evlog.desc_from_heads(heads)
    nrows = rows.shape[0]
    end = np.nonzero(rows[:, 0] == -32768)[0]
    if len(end):
        nrows = int(end[0])
    planes, stride = make_planes(rows[:nrows], desc)
    done = skipped = walked = 0
    for seg in segs:
        if seg.reset_kind != abi.RT_RESET_FULL:
            continue
        cfg = evlog.cfg_for(seg)
        stop = seg.end_row if seg.end_row >= 0 else nrows
        ev = fast_scan(fast_host, planes, stride, nrows, desc, cfg, seg.row, min(stop, nrows))
        if ev is None:
            continue
        canon = evlog.to_canon(ev)
        if seg.stop_row >= 0:
            canon = canon[canon["row"] <= seg.stop_row]
        assert evlog.matches_fixture(seg, canon), \
            f"{name}: segment at row {seg.row} parmset {seg.parmset}: {len(canon)} events, reference {seg.nevents}"
        assert not any(fast_scan.failed), f"{name}: segment at row {seg.row}: failed flags {fast_scan.failed}"
        skipped += sum(fast_scan.skipped); walked += (min(stop, nrows) - seg.row) * desc.ntrks
        # and with gap skipping off: the same events
        ev2 = fast_scan(fast_host, planes, stride, nrows, desc, cfg, seg.row, min(stop, nrows), skip=0)
        assert ev2.tobytes() == ev.tobytes() and not any(fast_scan.skipped)
        done += 1
    assert done > 0
    print(f"{name}: {done} segments, {skipped} of {walked} track-rows jumped over ({100.0 * skipped / max(walked, 1):.1f} %)")


def test_fast_path_equals_oracle_on_synthetic(fast_host, oracle_lib):
    """whole synthetic tape (8 blocks) as ONE unit, and with a per-track skew: event-for-event against the oracle"""
    hdr, rows = synth.nrzi_tape(nblocks=8, seed=11)
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    planes, stride = make_planes(rows, desc)
    for skew in (None, [0, 3, 1, 7, 2, 0, 12, 5, 50]):
        for pi in (0, 4):
            cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[pi], hdr.bpi, hdr.ips, skew=skew)
            tape = oracle_lib.open(desc); tape.upload(rows)
            for row0 in (0, 4096, 64 * 137):
                sc = tape.scan(cfg); sc.reset(abi.RT_RESET_FULL, row0)
                want, _ = sc.run(rows.shape[0]); sc.end()
                got = fast_scan(fast_host, planes, stride, rows.shape[0], desc, cfg, row0, rows.shape[0])
                assert not any(fast_scan.failed) and sum(fast_scan.skipped) > 0.3 * rows.shape[0] * 9, (fast_scan.failed, fast_scan.skipped)
                a, b = evlog.to_canon(got), evlog.to_canon(want)
                if a.tobytes() != b.tobytes():
                    k = evlog._first_diff(a, b)
                    pytest.fail(f"skew {skew} parmset {pi} row0 {row0}: event #{k}: fast {a[k] if k < len(a) else None} oracle {b[k] if k < len(b) else None} "
                                f"({len(a)} vs {len(b)})")
            tape.close()


@pytest.mark.parametrize("bpi,pi,mode", [(1600, 0, "nrzi"), (556, 1, "nrzi"), (300, 4, "nrzi"), (200, 0, "nrzi"), (800, 0, "pe"), (400, 2, "pe"), (1100, 6, "pe")])
def test_fast_path_other_window_widths(bpi, pi, mode, fast_host, oracle_lib):
    """window widths 3..50 (ring sizes 64 and 128, windows longer than a batch of rows), PE feedback, odd skews: the same
    synthetic samples scanned under other densities -- not a meaningful decode, but every event must match the oracle"""
    hdr, rows = synth.nrzi_tape(nblocks=4, seed=5)
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    planes, stride = make_planes(rows, desc)
    m = tbin.MODE_NRZI if mode == "nrzi" else tbin.MODE_PE
    table = parmsets.NRZI if mode == "nrzi" else parmsets.PE
    cfg = abi.make_cfg(m, table[pi], float(bpi), hdr.ips, skew=[1, 0, 5, 2, 0, 9, 3, 0, 17])
    tape = oracle_lib.open(desc); tape.upload(rows)
    w = oracle_lib.L.rt_pkww_width(C.byref(cfg), hdr.tdelta_ns)
    assert 3 <= w <= 50
    for row0 in (0, 32 * 211):
        sc = tape.scan(cfg); sc.reset(abi.RT_RESET_FULL, row0)
        want, _ = sc.run(rows.shape[0]); sc.end()
        for skip in (1, 0):
            got = fast_scan(fast_host, planes, stride, rows.shape[0], desc, cfg, row0, rows.shape[0], skip=skip)
            assert got is not None and not any(fast_scan.failed), fast_scan.failed
            a, b = evlog.to_canon(got), evlog.to_canon(want)
            if a.tobytes() != b.tobytes():
                k = evlog._first_diff(a, b)
                pytest.fail(f"bpi {bpi} width {w} parmset {pi} row0 {row0} skip {skip}: event #{k}: fast {a[k] if k < len(a) else None} "
                            f"oracle {b[k] if k < len(b) else None} ({len(a)} vs {len(b)})")
    tape.close()


def test_zero_crossing_fast_path_all_gcr_parmsets(fast_host, oracle_lib):
    """the 5 built-in GCR parameter sets (EWMA and moving-window clocks, resync) with -zeros on the GCR-density synthetic tile,
    with and without skew: event-for-event against the oracle"""
    rows = synth.gcr_like_tile(nblocks=2, bits_per_block=6000, gap_rows=9000)
    hdr = synth.gcr_header()
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    planes, stride = make_planes(rows, desc)
    tape = oracle_lib.open(desc); tape.upload(rows)
    for skew in (None, [0, 2, 1, 0, 11, 3, 0, 5, 1]):
        for pi in range(5):
            cfg = abi.make_cfg(tbin.MODE_GCR, parmsets.GCR[pi], hdr.bpi, hdr.ips, flags=abi.RT_F_FIND_ZEROS, skew=skew)
            for row0 in (0, 32 * 100):
                sc = tape.scan(cfg); sc.reset(abi.RT_RESET_FULL, row0)
                want, _ = sc.run(rows.shape[0]); sc.end()
                got = fast_scan(fast_host, planes, stride, rows.shape[0], desc, cfg, row0, rows.shape[0])
                assert got is not None and len(got) > 10000
                a, b = evlog.to_canon(got), evlog.to_canon(want)
                if a.tobytes() != b.tobytes():
                    k = evlog._first_diff(a, b)
                    pytest.fail(f"skew {skew} parmset {pi} row0 {row0}: event #{k}: fast {a[k] if k < len(a) else None} oracle {b[k] if k < len(b) else None} ({len(a)} vs {len(b)})")
    tape.close()
