"""The unit-equivalence rules on the GENERIC speculative unit scan (k_units_scan, readtape_b200/csrc/k_scan.cu), host side.

k_units_scan serves the configurations the fast kernels do not take: `-differentiate` (the zero-crossing detector of the
differentiated signal, decoder.c:654-683, and the peak detector fed with it) and the forced generic path (RT_SCAN=generic).  Its
detector code and its quiet tracker are the __host__ __device__ code of scan_generic.cuh; the bookkeeping of the proof data around
them is mirrored in tests/host_fast/fast_host.cu (generic_host_scan_unit).  Two things are checked on adversarial burst signals:
  1. a unit scanned from row0 gives the oracle's fresh-reset events from row0 (the detector code, differentiator included);
  2. whenever the product's rules (lookup_rules.h) accept a reset at ANOTHER row -- in front of the unit, inside it, at its tail,
     through a chain of event-free units -- the unit's events equal the oracle's fresh scan from that row.
"""
import ctypes as C

import numpy as np
import pytest

from readtape_b200 import abi, evlog, parmsets, tbin
from test_fast_host import fast_host, make_planes  # noqa: F401  (fixture)
import test_proof_host as proof

NOROW = proof.NOROW


def scan_unit(L, planes, stride, n, desc, cfg, row0, row_end):
    L.generic_host_scan_unit.restype = C.c_int
    L.generic_host_scan_unit.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(abi.TapeDesc), C.POINTER(abi.ScanCfg), C.c_uint64, C.c_uint64,
                                         C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    nt = desc.ntrks; cap = 1 << 17
    out = np.zeros((nt, cap), dtype=abi.EVENT_DTYPE); counts = np.zeros(nt, dtype=np.uint32); ms = L.fast_host_meta_size(); meta = np.zeros((nt, ms), dtype=np.uint8)
    rc = L.generic_host_scan_unit(planes.ctypes.data, stride, n, C.byref(desc), C.byref(cfg), row0, row_end, out.ctypes.data, cap, counts.ctypes.data, meta.ctypes.data)
    assert rc == 0 and counts.max(initial=0) <= cap
    ev = np.concatenate([out[k, :counts[k]] for k in range(nt)]); ev = ev[np.lexsort((ev['trk'], ev['row']))]
    m64 = meta[:, :72].copy().view('<u8'); m32 = meta[:, 72:88].copy().view('<u4')
    metas = proof.Metas(dict(first_event_row=int(r[0]), sync_row=int(r[1]), last_loud_row=int(r[2]), sync_early=int(r[3]), loud_early=int(r[4]), sync_first=int(r[5]),
                             quiet_from=int(r[6]), last_event_row=int(r[7]), quiet_tail_from=int(r[8]), nevents=int(w[1]), failed=int(w[2])) for r, w in zip(m64, m32))
    metas.raw = np.ascontiguousarray(meta)
    return ev, metas


def make_case(seed):
    rng = np.random.default_rng(seed)
    style = seed % 3          # 0: GCR -zeros -differentiate, 1: NRZI / PE peaks on the differentiated signal, 2: plain peaks / zero crossings on the generic path
    n = 64 * int(rng.integers(150, 450)); t = np.arange(n); rows = np.zeros((n, 9), dtype=np.int64)
    for k in range(9):
        period = rng.uniform(14, 60)
        amp = rng.uniform(1500, 30000) * (1 + 0.8 * np.sign(np.sin(2 * np.pi * t / rng.uniform(3000, 9000))))
        gate = (rng.random(n).cumsum() % 2000 > rng.uniform(300, 1500))
        sig = amp * np.sin(2 * np.pi * t / period + rng.uniform(0, 6)) * gate + rng.uniform(0, 800) * np.sin(2 * np.pi * t / 5000.0) + rng.normal(0, rng.uniform(3, 60), n)
        if rng.random() < 0.3:          # isolated spikes inside the gaps: loud rows that fire nothing
            at = rng.integers(0, n, 6); sig[at] += rng.uniform(2000, 9000) * rng.choice([-1, 1], 6)
        q = int(rng.choice([1, 1, 64, 512])); rows[:, k] = np.clip(np.round(sig / q) * q, -32767, 32767)
    rows = rows.astype('<i2')
    skew = [0] * 9 if rng.random() < 0.5 else [int(x) for x in rng.integers(0, 12, 9)]
    if style == 0:
        desc = abi.make_desc(9, 1.5, 160, 1_000_000_000)
        cfg = abi.make_cfg(tbin.MODE_GCR, parmsets.GCR[int(rng.integers(0, len(parmsets.GCR)))], 9042.0, float(rng.choice([50, 125])),
                           flags=abi.RT_F_FIND_ZEROS | abi.RT_F_DIFFERENTIATE, skew=skew)
    elif style == 1:
        desc = abi.make_desc(9, 4.4, 1280, 1_000_000_000)
        mode, ps = (tbin.MODE_NRZI, parmsets.NRZI) if rng.random() < 0.5 else (tbin.MODE_PE, parmsets.PE)
        cfg = abi.make_cfg(mode, ps[int(rng.integers(0, len(ps)))], float(rng.choice([556, 800, 1600])), 50.0, flags=abi.RT_F_DIFFERENTIATE, skew=skew)
    else:
        if rng.random() < 0.5:
            desc = abi.make_desc(9, 4.4, 1280, 1_000_000_000 if rng.random() < 0.8 else 0)
            cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[int(rng.integers(0, len(parmsets.NRZI)))], 800.0, 50.0, flags=abi.RT_F_INVERT if rng.random() < 0.3 else 0, skew=skew)
        else:
            desc = abi.make_desc(9, 1.5, 160, 1_000_000_000)
            cfg = abi.make_cfg(tbin.MODE_GCR, parmsets.GCR[int(rng.integers(0, len(parmsets.GCR)))], 9042.0, 50.0, flags=abi.RT_F_FIND_ZEROS, skew=skew)
    return rng, style, rows, n, desc, cfg


@pytest.mark.parametrize("seeds", [range(0, 6), range(6, 12)])
def test_generic_unit_scan_and_its_proof_data(seeds, fast_host, oracle_lib):
    L, ora = proof._bind(fast_host), oracle_lib
    failures = []; checked = accepted = bridged = chained = 0
    for seed in seeds:
        rng, style, rows, n, desc, cfg = make_case(seed)
        planes, stride = make_planes(rows, desc)
        tape = ora.open(desc); tape.upload(rows)
        try: sc = tape.scan(cfg)
        except abi.RtError: tape.close(); continue
        def fresh(s, upto):
            sc.reset(abi.RT_RESET_FULL, s); want, _ = sc.run(upto - s); return evlog.to_canon(want)
        def differs(tag, ev, s, upto):
            a = evlog.to_canon(ev); b = fresh(s, upto)
            if a.tobytes() == b.tobytes(): return
            k = evlog._first_diff(a, b)
            failures.append(f"seed {seed} style {style}: {tag}: event #{k} differs: unit {a[k] if k < len(a) else None} / fresh scan {b[k] if k < len(b) else None} ({len(a)} vs {len(b)} events)")
        for _ in range(4):
            row0 = int(rng.integers(0, n - 3000)); row_end = min(n, row0 + int(rng.integers(2000, 30000)))
            ev, metas = scan_unit(L, planes, stride, n, desc, cfg, row0, row_end)
            differs(f"unit [{row0}, {row_end}) against a fresh scan from its own first row", ev, row0, row_end)
            cands = sorted(set([max(0, row0 - int(d)) for d in rng.integers(1, 400, 12)] + [row0 + int(d) for d in rng.integers(1, 3000, 12)] + [row_end - int(d) for d in rng.integers(1, 3000, 8)]))
            for s in cands:
                if s >= row_end or s == row0: continue
                c, upto = proof.covers(L, desc, cfg, metas, row0, row_end, s); tl = (not c) and proof.tail_covers(L, desc, metas, row0, row_end, s)
                if not c and not tl and upto is not None and upto < row_end:
                    sc.reset(abi.RT_RESET_FULL, s); evb, done_ = sc.run(upto - s + 1)
                    if done_ == upto - s + 1 and not any(int(e['row']) <= metas[int(e['trk'])]['sync_row'] for e in evb):
                        c = True; bridged += 1
                checked += 1
                if not (c or tl): continue
                accepted += 1
                differs(f"unit [{row0}, {row_end}) accepted for a reset at row {s} ({'covers / bridge' if c else 'tail'} rule)", ev if c else ev[:0], s, row_end)
        for _ in range(6):                                        # chains: an event-free unit in a quiet stretch, and the unit behind it
            amp = np.abs(rows.astype(np.int32)).max(axis=1); blk = amp[: n // 256 * 256].reshape(-1, 256).max(axis=1)
            quiet = np.flatnonzero(blk < 400)
            if len(quiet) == 0: break
            r0 = int(quiet[int(rng.integers(0, len(quiet)))]) * 256 + int(rng.integers(0, 200))
            if r0 > n - 3000: continue
            r1 = r0 + int(rng.integers(150, 1500)); r2 = min(n, r1 + int(rng.integers(1500, 20000))); r1e = min(r2, r1 + int(rng.integers(40, 700)))
            evA, mA = scan_unit(L, planes, stride, n, desc, cfg, r0, r1e)
            evB, mB = scan_unit(L, planes, stride, n, desc, cfg, r1, r2)
            if len(evA): continue
            for s_ in sorted(set([r0] + [max(0, r0 - int(d)) for d in rng.integers(1, 300, 6)] + [r0 + int(d) for d in rng.integers(1, 200, 4)])):
                if s_ >= r1: continue
                if not proof.covers(L, desc, cfg, mA, r0, r1e, s_)[0] or not proof.chains(L, desc, cfg, mA, r0, r1e, mB, s_): continue
                chained += 1
                differs(f"chain [{r0}, {r1e}) -> [{r1}, {r2}) accepted for a reset at row {s_}", evB, s_, r2)
        sc.end(); tape.close()
    assert not failures, "\n".join(failures[:5])
    assert accepted >= 10, (checked, accepted, bridged, chained)
