"""CSV ingest: the .csv -> .tbin conversion of the reference's csvtbin tool (src/csvtbin.c:619-747) behind
include/rt_csv.h.

Parity chain:  reference tool (csvtbin_ref, unmodified)  ->  tests/golden/csv_golden.json (oracle/make_csv_golden.py)
               CPU oracle (oracle/csv_oracle.c)   == golden           [not gpu]
               host tool on the oracle backend    == golden           [not gpu]
               CUDA library (k_csv.cu)            == golden, == oracle, through the C-ABI   [gpu]
               csvtbin_b200 binary                == golden / == csvtbin_ref run beside it   [gpu]
Bit-exact everywhere: every int16 sample, the header fields, the clamp counters.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, capture_path
from readtape_b200 import abi, csvtbin, synth, tbin

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_csv_golden as gen  # noqa: E402   (the input generators only; nothing is run from the reference here)

DOC = json.load(open(os.path.join(GOLDEN, "csv_golden.json")))
CASES = {c["name"]: c for c in DOC["cases"]}
REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "csvtbin_ref")
ORACLE_TOOL = os.path.join(ROOT, "oracle", "_ref", "csvtbin_shim_oracle")
CUDA_TOOL = os.path.join(ROOT, "readtape_b200", "bin", "csvtbin_b200")
MODES = {"-nrzi": tbin.MODE_NRZI, "-pe": tbin.MODE_PE, "-gcr": tbin.MODE_GCR, "-whirlwind": tbin.MODE_WW}
_TEXT_CACHE = {}


def text_of(case) -> np.ndarray:
    key = json.dumps([case.get(k) for k in ("text", "synthetic", "long")])
    if key not in _TEXT_CACHE:
        _TEXT_CACHE[key] = gen.case_text(case)
    return _TEXT_CACHE[key]


def parse_opts(opts: str):
    """the option line of a golden case as keyword arguments of readtape_b200.csvtbin (what the C tool's main() does)"""
    o = {"ntrks": 9, "order": None, "scalefactor": 1.0, "invert": False, "subsample": 1, "maxvolts": 0.0, "skip": 0, "stopaft": None,
         "mode": 0, "redo": False, "starttime": 0.0, "endtime": None, "ww": None}
    for w in opts.split():
        k, _, v = w.partition("=")
        k = k.lower()
        if k in MODES: o["mode"] = MODES[k]
        elif k == "-ntrks": o["ntrks"] = int(v)
        elif k == "-order":
            if o["mode"] == tbin.MODE_WW: o["ww"] = v; o["ntrks"] = len(v)
            else: o["order"] = csvtbin.order_to_permutation(v, o["ntrks"])
        elif k == "-invert": o["invert"] = True
        elif k == "-scale": o["scalefactor"] = float(v)
        elif k == "-subsample": o["subsample"] = int(v)
        elif k == "-maxvolts": o["maxvolts"] = float(v)
        elif k == "-skip": o["skip"] = int(v)
        elif k == "-stopaft": o["stopaft"] = int(v)
        elif k == "-redo": o["redo"] = True
        elif k == "-starttime": o["starttime"] = float(v)
        elif k == "-endtime": o["endtime"] = float(v)
    return o


def convert_like_tool(lib, case):
    """csv_preread + write_tbin's row selection + the -redo rule, over the C-ABI of `lib`; -> (Preread, rows, stats list)"""
    o = parse_opts(case["opts"])
    with lib.csv_open(text_of(case)) as c:
        pre = csvtbin.preread(c, o["ntrks"], o["scalefactor"], o["subsample"], o["maxvolts"])
        first, t = csvtbin.HEADER_LINES, pre.tstart_ns
        start_ns = int(float(np.float32(o["starttime"])) * 1e9)
        skip = o["skip"]
        if skip > 0 or start_ns > 0:                       # csvtbin.c:668-678
            while True:
                first += 1; t += pre.tdelta_ns; skip = max(0, skip - 1)
                if not (t < start_ns or skip > 0):
                    break
        nrows = (c.nlines - first) // o["subsample"]
        if o["stopaft"] is not None: nrows = min(nrows, o["stopaft"])
        if o["endtime"] is not None:
            end_ns = int(float(np.float32(o["endtime"])) * 1e9)
            nrows = min(nrows, (end_ns - t) // pre.tdelta_ns + 1 if end_ns >= t else 1)
        maxvolts, stats = pre.maxvolts, []
        for _ in range(2):
            cfg = abi.make_csv_cfg(o["ntrks"], float(maxvolts), o["order"], o["scalefactor"], o["invert"], o["subsample"])
            rows, st = c.convert(cfg, first, nrows)
            stats.append(st)
            if not (st.too_big or st.too_small) or not o["redo"]:
                break
            newmax = max(st.maxvolts, -st.minvolts)
            maxvolts = np.float32(np.float32(int((np.float64(newmax) + 0.15) * np.float32(10.0))) / np.float32(10.0))   # csvtbin.c:737
    return pre, maxvolts, rows, stats


def check_against_golden(lib, name):
    case = CASES[name]
    pre, maxvolts, rows, stats = convert_like_tool(lib, case)
    assert (pre.tdelta_ns, pre.tstart_ns) == (case["tdelta_ns"], case["tstart_ns"])
    assert float(np.float32(maxvolts)) == case["maxvolts"]
    assert rows.shape == (case["nrows"], case["ntrks"])
    payload = rows.astype("<i2").tobytes() + b"\x00\x80"
    if "rows" in case:
        assert rows.tolist() == case["rows"]
    assert hashlib.sha256(payload).hexdigest() == case["payload_sha256"]
    warn = [ln for ln in case["log_tail"] if "too big" in ln or "too small" in ln]
    if warn:                                               # the counters of the first pass, as the reference logged them
        assert f"{stats[0].too_big} samples were too big" in warn[0].replace(",", "")
        assert f"{stats[0].too_small} samples were too small" in warn[1].replace(",", "")
    else:
        assert stats[0].too_big == 0 and stats[0].too_small == 0
    done = [ln for ln in case["log_tail"] if ln.startswith("done;")][0]
    assert done == "done; minimum voltage was %.1fV, maximum voltage was %.1fV" % (stats[0].minvolts, stats[0].maxvolts)


FAST = [n for n in CASES if not n.startswith("long")]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name, oracle_lib):
    check_against_golden(oracle_lib, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_matches_reference_golden(name, cuda_lib):
    check_against_golden(cuda_lib, name)


def masked(raw: bytes) -> bytes:
    """a .tbin without the 36 bytes of time_converted (csvtbin.h:60; the reference leaves tm_yday uninitialised)"""
    return raw[:168] + raw[204:]


def run_tool(tool, case, wd):
    if not os.path.exists(tool):
        pytest.skip(f"{tool} not built")
    os.makedirs(wd, exist_ok=True)
    text_of(case).tofile(os.path.join(wd, case["name"] + ".csv"))
    r = subprocess.run([tool] + case["opts"].split() + [case["name"]], capture_output=True, text=True, cwd=wd, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return open(os.path.join(wd, case["name"] + ".tbin"), "rb").read(), r.stdout


def check_tool(tool, name, tmp_path):
    case = CASES[name]
    raw, out = run_tool(tool, case, str(tmp_path / "new"))
    hdr = tbin.parse_header(raw)
    assert (hdr.flags, hdr.ntrks, hdr.tdelta_ns, hdr.mode, hdr.bpi, hdr.ips, hdr.tstart_ns, hdr.trkorder) == \
           (case["flags"], case["ntrks"], case["tdelta_ns"], case["mode"], case["bpi"], case["ips"], case["tstart_ns"], case["trkorder"])
    assert float(np.float32(hdr.maxvolts)) == case["maxvolts"]
    assert hashlib.sha256(raw[hdr.payload_offset:]).hexdigest() == case["payload_sha256"]
    for ln in case["log_tail"]:
        assert ln in out, ln
    if os.path.exists(REF_TOOL):                           # the reference tool travels with the repo: whole-file identity beside it
        want, _ = run_tool(REF_TOOL, case, str(tmp_path / "ref"))
        assert masked(raw) == masked(want)


@pytest.mark.parametrize("name", list(CASES))
def test_host_tool_on_oracle_backend_writes_the_reference_file(name, tmp_path):
    check_tool(ORACLE_TOOL, name, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_csvtbin_b200_writes_the_reference_file(name, tmp_path):
    check_tool(CUDA_TOOL, name, tmp_path)


def test_line_index_and_errors(oracle_lib):
    _check_line_index(oracle_lib)


@pytest.mark.gpu
def test_line_index_and_errors_cuda(cuda_lib):
    _check_line_index(cuda_lib)


def _check_line_index(lib):
    text = DOC["edge_text"]["forms"].encode()
    with lib.csv_open(np.frombuffer(text, dtype=np.uint8)) as c:
        want = text.split(b"\n")
        assert c.nlines == len(want) == 11
        for i, ln in enumerate(want):
            assert c.line(i) == ln
        with pytest.raises(abi.RtError):
            c.line(11)
        cfg = abi.make_csv_cfg(5, 4.4)
        with pytest.raises(abi.RtError):
            c.convert(cfg, 2, 10)                          # only 9 data lines
        bad = abi.make_csv_cfg(5, 4.4, order=[0, 1, 1, 2, 3])
        with pytest.raises(abi.RtError):
            c.convert(bad, 2, 1)
        want_max = np.float32(9)                                # scanfast_float("9.87654321"), csvtbin.c:411-415, in float32
        div = np.float32(10)
        for d in "87654321":
            want_max = np.float32(want_max + np.float32(np.float32(int(d)) / div)); div = np.float32(div * np.float32(10))
        assert c.max_abs(2, 9, 5) == want_max
    with lib.csv_open(np.frombuffer(b"", dtype=np.uint8)) as c:
        assert c.nlines == 0
    with lib.csv_open(np.frombuffer(b"\n\n1,2\n", dtype=np.uint8)) as c:
        assert c.nlines == 3 and c.line(0) == b"" and c.line(2) == b"1,2"


@pytest.mark.gpu
def test_cuda_equals_oracle_on_random_text(cuda_lib, oracle_lib):
    """ragged, hostile text: random field widths, signs, digit counts, blanks, missing columns; several geometries"""
    rng = np.random.default_rng(5)
    for ntrks, nlines, sub in [(5, 3000, 1), (9, 70000, 1), (12, 20011, 7), (19, 5000, 2)]:
        lines = ["title", "Time, ..."]
        for i in range(nlines):
            cols = [f"{i * 1.28e-6:.8f}"]
            for _ in range(ntrks if rng.random() > 0.01 else rng.integers(0, ntrks)):
                v = rng.normal(0, 3) * (10.0 ** rng.integers(-3, 3))
                nd = int(rng.integers(0, 14))
                s = f"{v:.{nd}f}"
                cols.append(" " * int(rng.integers(0, 4)) + s)
            lines.append(",".join(cols) if rng.random() > 0.5 else ", ".join(cols))
        text = np.frombuffer(("\n".join(lines) + ("\n" if ntrks % 2 else "")).encode(), dtype=np.uint8)
        order = list(rng.permutation(ntrks))
        cfg = abi.make_csv_cfg(ntrks, 7.3, order, 1.25, bool(ntrks & 1), sub)
        with oracle_lib.csv_open(text) as a, cuda_lib.csv_open(text) as b:
            assert a.nlines == b.nlines
            n = (a.nlines - 2) // sub
            ra, sa = a.convert(cfg, 2, n)
            rb, sb = b.convert(cfg, 2, n)
            assert np.array_equal(ra, rb)
            assert (sa.too_big, sa.too_small, sa.minvolts, sa.maxvolts) == (sb.too_big, sb.too_small, sb.minvolts, sb.maxvolts)
            assert sa.too_big > 0 and sa.too_small > 0
            assert a.max_abs(2, a.nlines - 2, ntrks, 1.25) == b.max_abs(2, b.nlines - 2, ntrks, 1.25)
            for i in (0, 1, 2, a.nlines // 2, a.nlines - 1):
                assert a.line(i) == b.line(i)


@pytest.mark.gpu
def test_csv_straight_into_the_tape_scans_like_the_tbin(cuda_lib):
    """rt_csv_convert(tape=...) leaves the rows on the device: the scan of a tape filled from CSV text gives the events of the
    same tape filled from the converted rows with rt_upload"""
    from readtape_b200 import parmsets
    path = capture_path("Microdata_20blks")
    hdr, rows = tbin.read_tbin(path)
    rows = np.asarray(rows)[:400000]
    text = synth.csv_from_rows(rows, hdr.maxvolts, hdr.tstart_ns, hdr.tdelta_ns)
    cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[0], 800, 50)
    with cuda_lib.csv_open(text) as c:
        pre = csvtbin.preread(c, 9)
        desc = abi.make_desc(9, float(pre.maxvolts), pre.tdelta_ns, pre.tstart_ns)
        t1 = cuda_lib.open(desc)
        _, conv, st = csvtbin.convert(c, 9, tape=t1)
        assert t1.nrows == len(rows) == st.rows
    t2 = cuda_lib.open(desc); t2.upload(conv)
    seen = []
    for t in (t1, t2):
        b = t.bulk_scan([cfg])
        ev, dg, bad = b.tile_digest(0, len(rows), 1)        # one tile = the whole tape: event count and order-free digest
        seen.append((int(ev[0]), int(dg[0]), bad))
        b.free()
    assert seen[0][0] > 10000 and seen[0] == seen[1] and seen[0][2] == 0


def _hostile_number(rng):
    k = int(rng.integers(0, 12))
    v = rng.normal(0, 3) * (10.0 ** rng.integers(-4, 3))
    return ["", "-", ".", f"{int(v)}", f"{v:.0f}.", ("%.3f" % v).replace("0.", "."), f"{v:.3e}", f"{v:.25f}", "--1.5", "1.2.3",
            f"{v:.{int(rng.integers(0, 9))}f}", f"{v:.{int(rng.integers(0, 9))}f}"][k]


def _hostile_case(rng):
    ntrks = int(rng.integers(5, 13)); nl = int(rng.integers(3, 300))
    lines = ["title", "Time" + "".join(f", c{i}" for i in range(ntrks))]
    t = 0.0
    for _ in range(nl):
        t += 1e-6
        sep = [", ", ",", " ", " , ", ",,"][int(rng.integers(0, 5))]
        cols = [f"{t:.8f}"] + [_hostile_number(rng) for _ in range(ntrks if rng.random() > 0.05 else int(rng.integers(0, ntrks + 3)))]
        ln = sep.join(cols)
        if rng.random() < 0.03: ln = ""
        if rng.random() < 0.03: ln += "\r"
        if rng.random() < 0.02: ln = "x" + ln                  # a time stamp that does not scan: wild header values, still identical
        lines.append(ln)
    text = "\n".join(lines) + ("\n" if rng.random() > 0.3 else "")
    opts = [f"-ntrks={ntrks}"]
    for p, o in ((0.3, "-invert"), (0.3, f"-scale={rng.uniform(0.5, 2):.3f}"), (0.3, f"-subsample={int(rng.integers(1, 4))}"),
                 (0.3, f"-skip={int(rng.integers(0, 5))}"), (0.3, "-redo"), (0.3, f"-maxvolts={rng.uniform(0.5, 14):.1f}")):
        if rng.random() < p:
            opts.append(o)
    return text, opts


def _fuzz_tool(tool, tmp_path, seed, ncases):
    """hostile text (garbage, exponents, doubled signs and points, blank lines, CR, missing and surplus columns, every separator the
    scanner skips) and random options through the unmodified reference tool and ours: the .tbin files must be identical"""
    if not (os.path.exists(REF_TOOL) and os.path.exists(tool)):
        pytest.skip("csvtbin_ref / the tool under test not built")
    rng = np.random.default_rng(seed)
    for it in range(ncases):
        text, opts = _hostile_case(rng)
        got = []
        for exe, nm in ((REF_TOOL, "a"), (tool, "b")):
            with open(tmp_path / f"{nm}.csv", "w", newline="") as fh:
                fh.write(text)
            out = tmp_path / f"{nm}.tbin"
            if out.exists():
                out.unlink()
            r = subprocess.run([exe] + opts + [nm], capture_output=True, text=True, cwd=str(tmp_path), timeout=120)
            got.append((r.returncode, masked(out.read_bytes()) if out.exists() and r.returncode == 0 else b""))
        assert got[0] == got[1], f"case {it} of seed {seed}: options {opts}, {len(got[0][1])} vs {len(got[1][1])} bytes"


def test_fuzz_host_tool_on_oracle_backend_against_reference(tmp_path):
    _fuzz_tool(ORACLE_TOOL, tmp_path, 11, 40)


@pytest.mark.gpu
def test_fuzz_csvtbin_b200_against_reference(tmp_path):
    _fuzz_tool(CUDA_TOOL, tmp_path, 12, 25)
