"""Soundness of the unit-equivalence proof on the HOST build of the scan kernels.

rt_bulk_lookup() (readtape_b200/csrc/lookup_rules.h: unit_covers, unit_tail_covers, chains_into_next) decides from the per-track
proof data a scan kernel leaves behind (TrkMeta: canonical rows, last loud rows, quiet tail) whether the events of a unit scanned
from row0 may stand in for a fresh reset at ANOTHER row.  The product's own rule code (the header is compiled into the host build,
lookup_host_* in tests/host_fast/fast_host.cu) is attacked: units cut ANYWHERE (also inside blocks --
the unit finder is only a heuristic, soundness must not depend on it), reset rows in front of, inside and at the tail of the unit,
synthetic NRZI tapes with noise and adversarial burst signals for the peak detector (two-pass scan) and the GCR zero-crossing
path, random parameter sets and skews.  Whenever the rule ACCEPTS, the unit's events must equal the oracle's fresh-reset scan from
that row, event for event.  Modelled: the quiet rule (late / early pair), the tail rule, the bridge scan (bridge_holds) and the
chaining of event-free units (units overlap: a unit is scanned past the start of the next).  The CUDA kernels instantiate the same
__host__ __device__ code.
"""
import ctypes as C

import numpy as np
import pytest

from readtape_b200 import abi, evlog, parmsets, synth, tbin
from test_fast_host import fast_host, make_planes  # noqa: F401  (fixture)

NOROW = 2**64 - 1
PRESCAN = 256                     # RT_PRESCAN_MIN (rt_dev.h)


def _bind(L):
    L.sparse_host_scan_unit.restype = C.c_int
    L.sparse_host_scan_unit.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(abi.TapeDesc), C.POINTER(abi.ScanCfg), C.c_uint64, C.c_uint64,
                                        C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int]
    L.fast_host_meta_size.restype = C.c_int
    L.lookup_host_covers.restype = C.c_int
    L.lookup_host_covers.argtypes = [C.POINTER(abi.TapeDesc), C.POINTER(abi.ScanCfg), C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.lookup_host_tail.restype = C.c_int
    L.lookup_host_tail.argtypes = [C.POINTER(abi.TapeDesc), C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
    L.lookup_host_chain.restype = C.c_int
    L.lookup_host_chain.argtypes = [C.POINTER(abi.TapeDesc), C.POINTER(abi.ScanCfg), C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
    return L


class Metas(list):
    """the proof data of one unit: parsed per track (for messages and the test's own bookkeeping) + the raw TrkMeta[] the rules take"""
    raw = None


def scan_unit(L, kind, planes, stride, n, desc, cfg, row0, row_end, frac):
    nt = desc.ntrks; cap = 1 << 17
    out = np.zeros((nt, cap), dtype=abi.EVENT_DTYPE); counts = np.zeros(nt, dtype=np.uint32); ms = L.fast_host_meta_size(); meta = np.zeros((nt, ms), dtype=np.uint8)
    if kind == 'sparse': rc = L.sparse_host_scan_unit(planes.ctypes.data, stride, n, C.byref(desc), C.byref(cfg), row0, row_end, out.ctypes.data, cap, counts.ctypes.data, meta.ctypes.data, frac, 1, 0)
    else: rc = L.fast_host_scan_unit(planes.ctypes.data, stride, n, C.byref(desc), C.byref(cfg), row0, row_end, out.ctypes.data, cap, counts.ctypes.data, meta.ctypes.data, 1)
    if rc != 0: return None, None
    ev = np.concatenate([out[k, :counts[k]] for k in range(nt)]); ev = ev[np.lexsort((ev['trk'], ev['row']))]
    m64 = meta[:, :72].copy().view('<u8'); m32 = meta[:, 72:88].copy().view('<u4')
    metas = Metas(dict(first_event_row=int(r[0]), sync_row=int(r[1]), last_loud_row=int(r[2]), sync_early=int(r[3]), loud_early=int(r[4]), sync_first=int(r[5]),
                  quiet_from=int(r[6]), last_event_row=int(r[7]), quiet_tail_from=int(r[8]), nevents=int(w[1]), failed=int(w[2])) for r, w in zip(m64, m32))
    metas.raw = np.ascontiguousarray(meta)
    return ev, metas


def covers(L, desc, cfg, metas, row0, row_end, start_row):
    """rtlookup::unit_covers -> (covered, bridge_to or None)"""
    br = C.c_uint64(NOROW)
    ok = L.lookup_host_covers(C.byref(desc), C.byref(cfg), row0, row_end, metas.raw.ctypes.data, start_row, C.byref(br))
    return bool(ok), (None if br.value == NOROW else int(br.value))


def tail_covers(L, desc, metas, row0, row_end, start_row):
    return bool(L.lookup_host_tail(C.byref(desc), row0, row_end, metas.raw.ctypes.data, start_row))


def chains(L, desc, cfg, metas, row0, row_end, metas_next, start_row):
    return bool(L.lookup_host_chain(C.byref(desc), C.byref(cfg), row0, row_end, metas.raw.ctypes.data, metas_next.raw.ctypes.data, start_row))


@pytest.mark.parametrize("seeds", [range(1, 5), range(5, 9), range(9, 13), [188, 4, 7]])     # 188: an event-free but LOUD zero-crossing unit was chained into the next one
def test_accepted_resets_reproduce_the_fresh_scan(seeds, fast_host, oracle_lib):
    L, ora = _bind(fast_host), oracle_lib
    failures = []; checked = accepted = bridged = chained = 0
    for seed in seeds:
        rng = np.random.default_rng(seed)
        style = seed % 3
        if style == 0:
            hdr, rows = synth.nrzi_tape(nblocks=int(rng.integers(2, 5)), seed=seed, data_bytes=int(rng.integers(8, 120)), noise_mv=float(rng.choice([2.0, 5.0, 40.0, 150.0])), wobble=0.01)
            rows = rows.copy(); n = len(rows)
            desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns if rng.random() < 0.8 else 0)
            mode, ps = (tbin.MODE_NRZI, parmsets.NRZI); bpi = 800.0
        else:
            n = 64 * int(rng.integers(150, 500)); t = np.arange(n); rows = np.zeros((n, 9), dtype=np.int64)
            for k in range(9):
                period = rng.uniform(14, 60)
                amp = rng.uniform(1500, 30000) * (1 + 0.8 * np.sign(np.sin(2 * np.pi * t / rng.uniform(3000, 9000))))
                gate = (rng.random(n).cumsum() % 2000 > rng.uniform(300, 1500))
                sig = amp * np.sin(2 * np.pi * t / period + rng.uniform(0, 6)) * gate + rng.uniform(0, 800) * np.sin(2 * np.pi * t / 5000.0) + rng.normal(0, rng.uniform(3, 60), n)
                q = int(rng.choice([1, 1, 64, 512])); rows[:, k] = np.clip(np.round(sig / q) * q, -32767, 32767)
            rows = rows.astype('<i2')
            desc = abi.make_desc(9, 4.4 if style == 1 else 1.5, 1280 if style == 1 else 160, 1_000_000_000)
            mode, ps = ((tbin.MODE_NRZI, parmsets.NRZI) if rng.random() < 0.5 else (tbin.MODE_PE, parmsets.PE)) if style == 1 else (tbin.MODE_GCR, parmsets.GCR)
            bpi = float(rng.choice([556, 800, 1600])) if style == 1 else 9042.0
        skew = [0] * 9 if rng.random() < 0.5 else [int(x) for x in rng.integers(0, 12, 9)]
        flags = abi.RT_F_FIND_ZEROS if mode == tbin.MODE_GCR else 0
        cfg = abi.make_cfg(mode, ps[int(rng.integers(0, len(ps)))], bpi, 50.0, flags=flags, skew=skew)
        det_peak = mode != tbin.MODE_GCR
        width = ora.L.rt_pkww_width(C.byref(cfg), desc.tdelta_ns) if det_peak else 0
        planes, stride = make_planes(rows, desc)
        tape = ora.open(desc); tape.upload(rows)
        try:
            sc = tape.scan(cfg)
        except abi.RtError:
            tape.close(); continue
        for _ in range(4):
            row0 = int(rng.integers(0, n - 3000)); row_end = min(n, row0 + int(rng.integers(2000, 40000)))
            ev, metas = scan_unit(L, 'sparse' if det_peak else 'zc', planes, stride, n, desc, cfg, row0, row_end, float(rng.choice([0.25, 0.7, 0.06])))
            if ev is None: continue
            cands = sorted(set([row0] + [max(0, row0 - int(d)) for d in rng.integers(1, 400, 12)] + [row0 + int(d) for d in rng.integers(1, 3000, 12)] + [row_end - int(d) for d in rng.integers(1, 3000, 8)]))
            for s in cands:
                if s >= row_end: continue
                c, upto = covers(L, desc, cfg, metas, row0, row_end, s); tl = (not c) and tail_covers(L, desc, metas, row0, row_end, s)
                if not c and not tl and upto is not None and upto < row_end:
                    # the bridge (rt_api.cu: bridge_holds): an exact scan from s must stay event-free up to every track's canonical row
                    sc.reset(abi.RT_RESET_FULL, s); evb, done_ = sc.run(upto - s + 1)
                    if done_ == upto - s + 1 and not any(int(e['row']) <= metas[int(e['trk'])]['sync_row'] for e in evb):
                        c = True; bridged += 1
                checked += 1
                if not (c or tl): continue
                accepted += 1
                sc.reset(abi.RT_RESET_FULL, s); want, _ = sc.run(row_end - s)
                a = evlog.to_canon(ev if c else ev[:0]); b = evlog.to_canon(want)
                if a.tobytes() != b.tobytes():
                    k = evlog._first_diff(a, b)
                    failures.append(f"seed {seed} style {style}: unit [{row0}, {row_end}) accepted for a reset at row {s} ({'covers / bridge' if c else 'tail'} rule), "
                                    f"but event #{k} differs: unit {a[k] if k < len(a) else None} / fresh scan {b[k] if k < len(b) else None} ({len(a)} vs {len(b)} events)")
        for _ in range(6):
            # aim at a quiet stretch: A = an event-free unit inside it, B = the unit behind it (reaching into the signal)
            amp = np.abs(rows.astype(np.int32)).max(axis=1); blk = amp[: n // 256 * 256].reshape(-1, 256).max(axis=1)
            quiet = np.flatnonzero(blk < 400)
            if len(quiet) == 0: break
            r0 = int(quiet[int(rng.integers(0, len(quiet)))]) * 256 + int(rng.integers(0, 200))
            if r0 > n - 3000: continue
            r1 = r0 + int(rng.integers(150, 1500)); r2 = min(n, r1 + int(rng.integers(1500, 20000)))
            frac = float(rng.choice([0.25, 0.7, 0.06]))
            r1e = min(r2, r1 + int(rng.integers(40, 700)))          # units overlap: A is scanned past the start of B
            evA, mA = scan_unit(L, 'sparse' if det_peak else 'zc', planes, stride, n, desc, cfg, r0, r1e, frac)
            evB, mB = scan_unit(L, 'sparse' if det_peak else 'zc', planes, stride, n, desc, cfg, r1, r2, frac)
            if evA is None or evB is None or len(evA): continue
            for s_ in sorted(set([r0] + [max(0, r0 - int(d)) for d in rng.integers(1, 300, 6)] + [r0 + int(d) for d in rng.integers(1, 200, 4)])):
                if s_ >= r1: continue
                if not covers(L, desc, cfg, mA, r0, r1e, s_)[0] or not chains(L, desc, cfg, mA, r0, r1e, mB, s_): continue
                chained += 1
                sc.reset(abi.RT_RESET_FULL, s_); want, _ = sc.run(r2 - s_)
                a = evlog.to_canon(evB); b = evlog.to_canon(want)
                if a.tobytes() != b.tobytes():
                    k = evlog._first_diff(a, b)
                    failures.append(f"seed {seed} style {style}: chain [{r0}, {r1e}) -> [{r1}, {r2}) accepted for a reset at row {s_}, but event #{k} differs: "
                                    f"unit {a[k] if k < len(a) else None} / fresh scan {b[k] if k < len(b) else None} ({len(a)} vs {len(b)} events)")
        sc.end(); tape.close()
    assert not failures, "\n".join(failures[:5])
    assert accepted >= 10 and bridged + chained >= 1, (checked, accepted, bridged, chained)
