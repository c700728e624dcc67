"""Randomised drop-in check: readtape with the event-driven readblock() (readtape_b200/host/readblock_b200.c) beside the unmodified
reference binary (oracle/_ref/readtape_ref, test infrastructure that travels with the repo) on inputs no fixture covers:

  * synthetic 9-track NRZI tapes with noise from 2 mV to 400 mV, speed wobble, dropouts on single tracks, tapes cut inside a block;
  * random windows of the bundled captures (NRZI 9/7-track, PE, GCR, Whirlwind) starting anywhere (also inside a block), with added
    noise and dropouts;
  * random options: mode / density given or auto-detected, -deskew, -nm, -invert, -subsample, -skip, -zeros, -v, -q.

Return code, every .tap/.bin/.csv and the log (minus the lines with the program name, the command line and the elapsed time) must
be identical.  CPU: the shim on the oracle backend; GPU: the product binary (speculative scan, bridge scans, exact continuation).
Known, documented difference: a .tbin WITHOUT its end marker makes the reference abort with fatal(); the shim ends
the tape there -- such files are not generated here.
"""
import glob
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from readtape_b200 import synth, tbin

REF = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
ORACLE_SHIM = os.path.join(ROOT, "oracle", "_ref", "readtape_shim_oracle")
CUDA_SHIM = os.path.join(ROOT, "readtape_b200", "bin", "readtape_b200")
VOLATILE = re.compile(r"^this is readtape version|command line:|samples were processed in")


def synthetic_case(rng, wd):
    nb = int(rng.integers(1, 5)); noise = float(rng.choice([2.0, 5.0, 40.0, 150.0, 400.0])); wob = float(rng.choice([0.0, 0.01, 0.04]))
    hdr, rows = synth.nrzi_tape(nblocks=nb, seed=int(rng.integers(1, 1 << 30)), data_bytes=int(rng.integers(8, 300)), noise_mv=noise, wobble=wob)
    rows = rows.copy()
    for _ in range(int(rng.integers(0, 3))):                                     # dropouts / dead stretches on one track
        a = int(rng.integers(0, len(rows) - 10)); b = a + int(rng.integers(10, 4000)); k = int(rng.integers(0, 9))
        rows[a:b, k] = (rows[a:b, k] * float(rng.choice([0.0, 0.1, 0.3]))).astype(np.int16)
    if rng.random() < 0.2:
        rows = rows[: int(rng.integers(len(rows) // 2, len(rows)))]              # the tape ends inside a block
    with open(os.path.join(wd, "t.tbin"), "wb") as fh:
        fh.write(tbin.build_header(hdr)); rows.tofile(fh); fh.write(np.array([tbin.END_MARK], dtype="<i2").tobytes())
    opts = ["-tap"]
    if rng.random() < 0.5: opts.append("-nrzi")
    r = rng.random()
    if r < 0.3: opts.append("-bpi=800")
    elif r < 0.4: opts.append("-bpi=1600")
    for p, o in ((0.3, "-deskew"), (0.3, "-nm"), (0.2, "-invert"), (0.15, "-subsample=2"), (0.3, "-v"), (0.15, "-q"),
                 (0.2, f"-skip={int(rng.integers(1, 20000))}"), (0.2, "-zeros")):
        if rng.random() < p:
            opts.append(o)
    return opts, f"synthetic NRZI, {nb} blocks, noise {noise} mV, {len(rows)} rows"


def capture_case(rng, wd):
    from oracle import captures
    docs = sorted(glob.glob(os.path.join(GOLDEN, "*.segments.json")))
    doc = json.load(open(docs[int(rng.integers(0, len(docs)))]))
    name = doc["capture"][:-5] if doc["capture"].endswith(".tbin") else doc["capture"]
    path = captures.staged_path(name)
    if path is None:
        return None, None
    raw = open(path, "rb").read(); hdr = tbin.parse_header(raw)
    nh = doc["heads"]["nheads"]
    rows = np.frombuffer(raw[hdr.payload_offset:], dtype="<i2"); rows = rows[: len(rows) // nh * nh].reshape(-1, nh)
    end = np.flatnonzero(rows[:, 0] == -32768); rows = rows[: end[0]] if len(end) else rows
    a = int(rng.integers(0, max(1, len(rows) - 60000))); b = a + int(rng.integers(20000, 250000))
    win = rows[a:b].copy()
    if rng.random() < 0.5:
        win = np.clip(win.astype(np.int32) + rng.normal(0, float(rng.choice([30, 300, 1500])), win.shape).astype(np.int32), -32767, 32767).astype("<i2")
    for _ in range(int(rng.integers(0, 3))):
        x = int(rng.integers(0, len(win) - 10)); y = x + int(rng.integers(10, 5000)); k = int(rng.integers(0, nh))
        win[x:y, k] = (win[x:y, k] * float(rng.choice([0.0, 0.1, 0.3]))).astype(np.int16)
    with open(os.path.join(wd, "t.tbin"), "wb") as fh:
        fh.write(raw[:hdr.payload_offset]); win.tofile(fh); fh.write(np.array([-32768], dtype="<i2").tobytes())
    opts = [o for o in doc["options"].split() if not o.startswith("-v")]
    for p, o in ((0.25, "-nm"), (0.25, "-v"), (0.15, "-q"), (0.15, f"-skip={int(rng.integers(1, 15000))}")):
        if rng.random() < p and o not in opts:
            opts.append(o)
    return opts, f"{name} rows [{a}, {b})"


def run_both(exe, opts, wd):
    res = []
    for binary, nm in ((REF, "ref"), (exe, "new")):
        sub = os.path.join(wd, nm); os.makedirs(sub, exist_ok=True)
        for f in os.listdir(sub):
            os.remove(os.path.join(sub, f))
        env = {k: v for k, v in os.environ.items() if k != "RT_STATS"}
        p = subprocess.run([binary] + opts + ["-outf=o", "../t.tbin"], capture_output=True, text=True, errors="replace", timeout=900, cwd=sub, env=env)
        outs = {f: open(os.path.join(sub, f), "rb").read() for f in sorted(os.listdir(sub)) if f.endswith((".tap", ".bin", ".csv"))}
        logf = os.path.join(sub, "o.log")
        log = [ln for ln in open(logf, errors="replace").read().splitlines() if not VOLATILE.search(ln)] if os.path.exists(logf) else None
        res.append((p.returncode, outs, log))
    return res


def fuzz(exe, make_case, seed, ncases, tmp_path):
    if not (os.path.exists(REF) and os.path.exists(exe)):
        pytest.skip("readtape_ref / the binary under test not built")
    rng = np.random.default_rng(seed)
    wd = str(tmp_path)
    failures = []
    for it in range(ncases):
        opts, what = make_case(rng, wd)
        if opts is None:
            continue
        ref, new = run_both(exe, opts, wd)
        if ref != new:
            why = "return code" if ref[0] != new[0] else "output files" if ref[1] != new[1] else "log"
            line = ""
            if why == "log" and ref[2] and new[2]:
                k = next((i for i, (x, y) in enumerate(zip(ref[2], new[2])) if x != y), min(len(ref[2]), len(new[2])))
                line = f" (line {k + 1}: ref {ref[2][k][:120] if k < len(ref[2]) else None!r} / new {new[2][k][:120] if k < len(new[2]) else None!r})"
            failures.append(f"case {it} of seed {seed}: {what}, options {' '.join(opts)}: {why} differs{line}")
    assert not failures, "\n".join(failures)


def test_shim_on_oracle_backend_random_synthetic_tapes(tmp_path):
    fuzz(ORACLE_SHIM, synthetic_case, 101, 16, tmp_path)


def test_shim_on_oracle_backend_random_capture_windows(tmp_path):
    fuzz(ORACLE_SHIM, capture_case, 102, 12, tmp_path)


@pytest.mark.gpu
def test_readtape_b200_random_synthetic_tapes(tmp_path):
    fuzz(CUDA_SHIM, synthetic_case, 201, 12, tmp_path)


@pytest.mark.gpu
def test_readtape_b200_random_capture_windows(tmp_path):
    fuzz(CUDA_SHIM, capture_case, 202, 12, tmp_path)
