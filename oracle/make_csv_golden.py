#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- generates tests/golden/csv_golden.json: what the UNMODIFIED reference converter
(oracle/_ref/csvtbin_ref, compiled by oracle/Makefile from /root/reference/src/csvtbin.c) writes for a set of CSV inputs
and option lines.  Run in the build container (the reference sources exist only there):

    make -C oracle ref && python oracle/make_csv_golden.py

Each case names its input (a seeded synthetic capture, re-made by tests/test_csv.py with the same call, or literal text kept
in the JSON) and records, from the reference's .tbin: the header fields the conversion decides (flags, ntrks, tdelta,
maxvolts, mode, bpi, ips, tstart, track-order extension), the number of rows, and the SHA-256 of the payload (rows + end
marker).  Small cases also keep the rows themselves.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from readtape_b200 import synth, tbin  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "csvtbin_ref")

EDGE_TEXT = {
    "forms": ("'title\nTime, a, b, c, d, e\n"
              "0.000001, 1, -2, 3.999999999999, 0.000001, -0.000001\n"
              "0.000002, .5, -.5, 4., 007.25, 1.23456789012345678\n"
              "0.000003,0.1,0.2,0.3,0.4,0.5\n"
              "0.000004,   -0.7 ,  0.7, -0, 0, 9.87654321\n"
              "0.000005, 2.5, -2.5, 1e3, 7, 7\n"           # scanfast stops at 'e': the rest of the line reads as zeros
              "0.000006, 3.3\n"                             # missing columns are zeros
              "\n"                                          # a blank line is a row of zeros
              "0.000008, abc, 1, 2, 3, 4\n"                 # garbage: the scanner does not advance, everything is zero
              "0.000009, -3.25, 3.25, -1.125, 1.125, 0.0625"),   # no newline at the end
    "crlf": "'title\r\nTime, a, b, c, d, e\r\n" + "".join(f"{i * 1e-6:.6f}, {i * 0.25:.3f}, {-i * 0.125:.4f}, 0.5, -0.5, {i}\r\n" for i in range(1, 40)),
}


def synthetic(seed: int, nrows: int, ntrks: int, maxvolts: float, tdelta: int = 1280, decimals: int = 5):
    rng = np.random.default_rng(seed)
    rows = np.clip(rng.normal(0, 9000, (nrows, ntrks)), -32767, 32767).astype(np.int16)
    rows[::97] = 0
    return synth.csv_from_rows(rows, maxvolts, 7452570880, tdelta, decimals)


CASES = [
    {"name": "forms", "text": "forms", "opts": "-ntrks=5 -nrzi"},
    {"name": "forms_inv_scale", "text": "forms", "opts": "-ntrks=5 -order=3p210 -pe -invert -scale=0.75 -bpi=1600 -ips=50"},
    {"name": "crlf", "text": "crlf", "opts": "-ntrks=5 -gcr -subsample=4 -skip=3"},
    {"name": "syn9", "synthetic": [11, 30000, 9, 4.4], "opts": "-ntrks=9 -order=01234567p -nrzi -bpi=800 -ips=50 -descr=golden"},
    {"name": "syn9_sub", "synthetic": [12, 30001, 9, 6.1], "opts": "-ntrks=9 -order=p76543210 -pe -subsample=3 -skip=5 -stopaft=7000 -invert -maxvolts=9"},
    {"name": "syn7_ww", "synthetic": [13, 20000, 7, 2.0], "opts": "-whirlwind -order=a0b1c2x -scale=2.5 -reverse"},
    {"name": "syn6_times", "synthetic": [14, 40000, 6, 3.0, 2500, 3], "opts": "-ntrks=6 -starttime=7.46 -endtime=7.5"},
    # more than PREREAD_COUNT lines with the large excursions after them: too-big / too-small samples and the -redo rule
    {"name": "long_redo", "long": [15, 1000400, 5], "opts": "-ntrks=5 -nrzi -redo"},
    {"name": "long_noredo", "long": [15, 1000400, 5], "opts": "-ntrks=5 -nrzi -maxvolts=1.5"},
]


def long_text(seed: int, nrows: int, ntrks: int):
    rng = np.random.default_rng(seed)
    rows = rng.integers(-3000, 3000, (nrows, ntrks)).astype(np.int16)
    rows[1000100:1000200] *= 9
    return synth.csv_from_rows(rows, 10.0, 0, 1000, 3)


def case_text(case) -> np.ndarray:
    if "text" in case:
        return np.frombuffer(EDGE_TEXT[case["text"]].encode(), dtype=np.uint8)
    if "long" in case:
        return long_text(*case["long"])
    return synthetic(*case["synthetic"])


def run_reference(case, workdir):
    base = os.path.join(workdir, case["name"])
    case_text(case).tofile(base + ".csv")
    # a relative name: csvtbin takes anything that starts with '/' for an option
    r = subprocess.run([REF] + case["opts"].split() + [case["name"]], capture_output=True, text=True, cwd=workdir)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(base + ".tbin", "rb").read()
    hdr = tbin.parse_header(raw)
    payload = raw[hdr.payload_offset:]
    assert payload[-2:] == b"\x00\x80"
    nrows = (len(payload) - 2) // (2 * hdr.ntrks)
    out = {"name": case["name"], "opts": case["opts"], "flags": hdr.flags, "ntrks": hdr.ntrks, "tdelta_ns": hdr.tdelta_ns,
           "maxvolts": float(np.float32(hdr.maxvolts)), "mode": hdr.mode, "bpi": hdr.bpi, "ips": hdr.ips, "tstart_ns": hdr.tstart_ns,
           "trkorder": hdr.trkorder, "nrows": nrows, "payload_sha256": hashlib.sha256(payload).hexdigest(),
           "log_tail": [ln for ln in r.stdout.splitlines() if "WARNING" in ln or "redoing" in ln or "you should" in ln or ln.startswith("done;")]}
    for k in ("text", "synthetic", "long"):
        if k in case:
            out[k] = case[k]
    if nrows <= 64:
        out["rows"] = np.frombuffer(payload[:-2], dtype="<i2").reshape(nrows, hdr.ntrks).tolist()
    return out


def main():
    assert os.path.exists(REF), "build the reference converter first: make -C oracle ref"
    with tempfile.TemporaryDirectory() as wd:
        docs = [run_reference(c, wd) for c in CASES]
    doc = {"generator": "oracle/make_csv_golden.py", "reference": "csvtbin 1.12 (src/csvtbin.c), unmodified, gcc -O2",
           "edge_text": EDGE_TEXT, "cases": docs}
    path = os.path.join(ROOT, "tests", "golden", "csv_golden.json")
    json.dump(doc, open(path, "w"), indent=1)
    print("wrote", path, [(d["name"], d["nrows"], d["maxvolts"], d["log_tail"]) for d in docs])


if __name__ == "__main__":
    main()
