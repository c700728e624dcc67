"""The CPU oracle against the UNMODIFIED reference on inputs no capture holds.

The instrumented reference (oracle/_ref/readtape_evdump: the reference's own objects + the --wrap event-dump shim, test
infrastructure that travels with the repo) decodes adversarial captures in every mode -- NRZI, PE, GCR -zeros, GCR -zeros
-differentiate, Whirlwind; bursts and gaps, amplitude steps (AGC swings), frequency jitter, DC drift, coarse quantisation (runs of
equal samples), noise; -deskew, -nm, -invert at random -- and logs every reset and every event the mode handlers see.  The oracle
(oracle/scan_oracle.c) is driven through the same reset sequence and must reproduce every event: row, track, polarity, f64 time,
f32 top / bottom volts, f32 AGC gain, bit for bit.  (End of round 2: 350 captures, 7 571 decode segments, 22.8 M events, no
deviation.)  Everything else in the repo is compared with this oracle.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from readtape_b200 import abi, evlog, tbin

EVD = os.path.join(ROOT, "oracle", "_ref", "readtape_evdump")


@pytest.mark.parametrize("seeds", [range(0, 5), range(5, 10)])
def test_oracle_reproduces_the_reference_events_on_adversarial_captures(seeds, oracle_lib, tmp_path):
    if not os.path.exists(EVD):
        pytest.skip("oracle/_ref/readtape_evdump not built (make -C oracle ref in the build container)")
    ora, wd = oracle_lib, str(tmp_path)
    failures = []; nseg = nev = 0
    for seed in seeds:
        rng = np.random.default_rng(seed)
        style = seed % 5       # 0 NRZI, 1 PE, 2 GCR zeros, 3 GCR zeros differentiate, 4 Whirlwind
        nt = 6 if style == 4 else 9
        n = 64 * int(rng.integers(200, 900)); t = np.arange(n); rows = np.zeros((n, nt), dtype=np.int64)
        base_period = {0: 19.5, 1: 9.8, 2: 13.8, 3: 13.8, 4: 52.0}[style]
        for k in range(nt):
            period = base_period * rng.uniform(1.6, 2.4)
            amp = rng.uniform(3000, 28000) * (1 + 0.5 * np.sign(np.sin(2 * np.pi * t / rng.uniform(3000, 9000))))
            gate = (rng.random(n).cumsum() % 4000 > rng.uniform(800, 2500))
            phase = np.cumsum(2 * np.pi / period * (1 + 0.5 * (rng.random(n).cumsum() % 37 > 18)))
            sig = amp * np.sin(phase + rng.uniform(0, 6)) * gate + rng.uniform(0, 600) * np.sin(2 * np.pi * t / 5000.0) + rng.normal(0, rng.uniform(3, 80), n)
            q = int(rng.choice([1, 1, 64, 512])); rows[:, k] = np.clip(np.round(sig / q) * q, -32767, 32767)
        rows = rows.astype('<i2')
        mode = {0: tbin.MODE_NRZI, 1: tbin.MODE_PE, 2: tbin.MODE_GCR, 3: tbin.MODE_GCR, 4: tbin.MODE_WW}[style]
        hdr = tbin.TbinHeader(descr='fuzz', flags=tbin.TBIN_NO_REORDER if style != 4 else tbin.TBIN_NO_REORDER | tbin.TBIN_TRKORDER_INCLUDED, ntrks=nt,
                              tdelta_ns={0: 1280, 1: 1280, 2: 160, 3: 160, 4: 3840}[style], maxvolts=4.4 if style in (0, 1, 4) else 1.5, mode=mode,
                              bpi={0: 800.0, 1: 1600.0, 2: 9042.0, 3: 9042.0, 4: 100.0}[style], ips=50.0, tstart_ns=1_000_000_000,
                              trkorder='a0b1c2' if style == 4 else None)
        path = os.path.join(wd, 't.tbin')
        with open(path, 'wb') as fh:
            fh.write(tbin.build_header(hdr)); rows.tofile(fh); fh.write(np.array([-32768], dtype='<i2').tobytes())
        opts = {0: ['-nrzi'], 1: ['-pe'], 2: ['-gcr', '-zeros'], 3: ['-gcr', '-zeros', '-differentiate'], 4: ['-whirlwind', '-fluxdir=auto']}[style] + ['-tap']
        if rng.random() < 0.3 and style != 4: opts.append('-deskew')
        if rng.random() < 0.3: opts.append('-nm')
        if rng.random() < 0.2: opts.append('-invert')
        ev = os.path.join(wd, 't.ev')
        if os.path.exists(ev): os.remove(ev)
        p = subprocess.run([EVD] + opts + ['-outf=' + os.path.join(wd, 'o'), path], env=dict(os.environ, RT_EVDUMP=ev), capture_output=True, text=True, errors='replace', timeout=600)
        if not os.path.exists(ev): continue
        try:
            heads = evlog.parse_heads(ev); segs = evlog.parse(ev)
        except Exception as e:
            failures.append(f"seed {seed}: event log of the reference does not parse: {e}"); continue
        tape = ora.open(evlog.desc_from_heads(heads)); tape.upload(rows)
        try:
            for seg, got in evlog.replay(tape, segs):
                nseg += 1; nev += len(seg.events)
                msg = evlog.compare(seg, got)
                if msg:
                    failures.append(f"seed {seed} style {style} options {' '.join(opts)}: {msg[:300]}"); break
        except abi.RtError as e:
            failures.append(f"seed {seed} style {style} options {' '.join(opts)}: oracle error {str(e)[:200]} (reference rc {p.returncode})")
        tape.close()
    assert not failures, "\n".join(failures[:5])
    assert nseg >= 20 and nev >= 100000, (nseg, nev)
