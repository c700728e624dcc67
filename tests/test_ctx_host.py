"""The exact stateful scan (readtape_b200/csrc/scan_generic.cuh -- what k_ctx_scan runs per track: Whirlwind, retries, bridge scans,
everything the speculative scan cannot prove) built for the HOST, with its skip-ahead over the mask planes, against the oracle.

ctx_scan_rows() is the kernel's own loop (k_scan.cu calls the same function): walk rows, or jump from candidate row to candidate row
and rebuild the window, the deskew FIFO, the blind countdown and the lazily kept minimum from the plane.  Fuzzed here: noisy synthetic
NRZI tapes, adversarial burst signals for NRZI / PE / Whirlwind (the AGC state persists, thresholds cross the masks' in both
directions), random parameter sets, skews, reset rows, span lengths (one span = one kernel launch) and mask thresholds -- the events
must be the oracle's whatever the spans and the masks.
"""
import ctypes as C

import numpy as np
import pytest

from readtape_b200 import abi, evlog, parmsets, synth, tbin
from test_fast_host import fast_host, make_planes  # noqa: F401  (fixture)


@pytest.mark.parametrize("seeds", [range(0, 4), range(4, 8), range(8, 12)])
def test_exact_scan_with_skip_ahead_equals_oracle(seeds, fast_host, oracle_lib):
    L, ora = fast_host, oracle_lib
    L.generic_host_ctx_scan.restype = C.c_int
    L.generic_host_ctx_scan.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(abi.TapeDesc), C.POINTER(abi.ScanCfg), C.c_uint64, C.c_uint64, C.c_uint64,
                                        C.c_int, C.c_float, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    failures = []; jumps = runs = 0
    for seed in seeds:
        rng = np.random.default_rng(seed)
        style = seed % 4
        if style == 0:
            hdr, rows = synth.nrzi_tape(nblocks=int(rng.integers(2, 5)), seed=seed, data_bytes=int(rng.integers(8, 120)), noise_mv=float(rng.choice([2.0, 5.0, 40.0, 150.0])), wobble=0.01)
            rows = rows.copy(); n = len(rows); nt = 9
            desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns if rng.random() < 0.8 else 0)
            cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[int(rng.integers(0, len(parmsets.NRZI)))], 800.0, 50.0, skew=None if rng.random() < 0.5 else [int(x) for x in rng.integers(0, 12, 9)])
        else:
            nt = 6 if style == 3 else 9
            n = 64 * int(rng.integers(150, 500)); t = np.arange(n); rows = np.zeros((n, nt), dtype=np.int64)
            for k in range(nt):
                period = rng.uniform(14, 60) if style != 3 else rng.uniform(30, 120)
                amp = rng.uniform(1500, 30000) * (1 + 0.8 * np.sign(np.sin(2 * np.pi * t / rng.uniform(3000, 9000))))
                gate = (rng.random(n).cumsum() % 2000 > rng.uniform(300, 1500))
                sig = amp * np.sin(2 * np.pi * t / period + rng.uniform(0, 6)) * gate + rng.uniform(0, 800) * np.sin(2 * np.pi * t / 5000.0) + rng.normal(0, rng.uniform(3, 60), n)
                q = int(rng.choice([1, 1, 64, 512])); rows[:, k] = np.clip(np.round(sig / q) * q, -32767, 32767)
            rows = rows.astype('<i2')
            skew = None if rng.random() < 0.5 else [int(x) for x in rng.integers(0, 12, nt)]
            if style == 3:
                desc = abi.make_desc(6, 4.4, 3840, 1_000_000_000)
                cfg = abi.make_cfg(tbin.MODE_WW, parmsets.WW[int(rng.integers(0, 2))], 100.0, 50.0, skew=skew)
            else:
                desc = abi.make_desc(9, 4.4, 1280, 1_000_000_000)
                mode, ps = (tbin.MODE_NRZI, parmsets.NRZI) if style == 1 else (tbin.MODE_PE, parmsets.PE)
                cfg = abi.make_cfg(mode, ps[int(rng.integers(0, len(ps)))], float(rng.choice([556, 800, 1600])), 50.0, skew=skew)
        planes, stride = make_planes(rows, desc)
        tape = ora.open(desc); tape.upload(rows)
        try: sc = tape.scan(cfg)
        except abi.RtError: tape.close(); continue
        for _ in range(3):
            r0 = int(rng.integers(0, n // 2)); r1 = min(n, r0 + int(rng.integers(3000, 60000)))
            sc.reset(abi.RT_RESET_FULL, r0); want, _ = sc.run(r1 - r0); b = evlog.to_canon(want)
            for use_masks, span, frac in ((1, int(rng.choice([4096, 1000, 131072, 333])), float(rng.choice([0.25, 0.1, 0.6]))), (0, 5000, 0.25)):
                cap = 1 << 17
                out = np.zeros((nt, cap), dtype=abi.EVENT_DTYPE); counts = np.zeros(nt, dtype=np.uint32); st = np.zeros(5, dtype=np.uint64)
                rc = L.generic_host_ctx_scan(planes.ctypes.data, stride, n, C.byref(desc), C.byref(cfg), r0, r1, span, use_masks, frac, out.ctypes.data, cap, counts.ctypes.data, st.ctypes.data)
                ev = np.concatenate([out[k, :counts[k]] for k in range(nt)]); ev = ev[np.lexsort((ev['trk'], ev['row']))]
                a = evlog.to_canon(ev); runs += 1; jumps += int(st[1])
                if a.tobytes() != b.tobytes():
                    k = evlog._first_diff(a, b)
                    failures.append(f"seed {seed} style {style} rows [{r0}, {r1}) masks {use_masks} span {span} T0 fraction {frac} (rc {rc}): event #{k}: "
                                    f"host build {a[k] if k < len(a) else None} / oracle {b[k] if k < len(b) else None} ({len(a)} vs {len(b)} events)")
        sc.end(); tape.close()
    assert not failures, "\n".join(failures[:5])
    assert runs >= 12 and jumps > 1000, (runs, jumps)
