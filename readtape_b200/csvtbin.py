"""Host side of the CSV ingest, the Python mirror of what csvtbin's main()/csv_preread()/write_tbin() do around
the conversion loop (src/csvtbin.c:619-747): it decides the sample period, the start time and the full-scale voltage the way
the reference does, then hands the text to the C-ABI of include/rt_csv.h (CUDA library, or the oracle library in tests).
The compiled tool readtape_b200/host/csvtbin_b200.c does the same in C.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import abi

PREREAD_COUNT = 1_000_000          # csvtbin.c:96
HEADER_LINES = 2                   # "first two lines in the input file are headers from Saleae" (csvtbin.c:664, :626)


def scanfast_double(s: bytes) -> float:
    """csvtbin.c:419-433 on the first number of `s` (float64 arithmetic in the reference's order)"""
    i, n, neg = 0, 0.0, False
    while i < len(s) and s[i] in b" ,":
        i += 1
    if i < len(s) and s[i:i + 1] == b"-":
        i += 1; neg = True
    while i < len(s) and 48 <= s[i] <= 57:
        n = n * 10 + (s[i] - 48); i += 1
    if i < len(s) and s[i:i + 1] == b".":
        i += 1; div = 10.0
        while i < len(s) and 48 <= s[i] <= 57:
            n += (s[i] - 48) / div; div *= 10; i += 1
    return -n if neg else n


@dataclasses.dataclass
class Preread:
    tstart_ns: int
    tdelta_ns: int
    maxvolts: np.float32            # rounded up as csvtbin.c:649 does
    lines: int                      # data lines looked at


def order_to_permutation(order: str, ntrks: int):
    """parse_track_order for PE/NRZI/GCR (csvtbin.c:329-340): 'P' = the last position, digits = themselves"""
    if len(order) != ntrks:
        raise ValueError("bad track order")
    perm = []
    for ch in order:
        if ch.upper() == "P":
            perm.append(ntrks - 1)
        elif ch.isdigit() and int(ch) <= ntrks - 2:
            perm.append(int(ch))
        else:
            raise ValueError("bad track order")
    if sorted(perm) != list(range(ntrks)):
        raise ValueError("bad track order")
    return perm


def preread(csv: "abi.Csv", ntrks: int, scalefactor: float = 1.0, subsample: int = 1, maxvolts_given: float = 0.0) -> Preread:
    """csv_preread (csvtbin.c:619-657): lines 1 .. PREREAD_COUNT-1 of the data give the period and the maximum"""
    ndata = csv.nlines - HEADER_LINES
    if ndata < 2:
        raise ValueError("not enough data lines")
    n = min(ndata, PREREAD_COUNT - 1)
    first = scanfast_double(csv.line(HEADER_LINES))
    if first < 0:
        raise ValueError("negative first time stamp: csv_preread's period estimate is undefined for it")
    last = scanfast_double(csv.line(HEADER_LINES + n - 1))
    tstart = int((first + 0.5e-9) * 1e9)
    tdelta = int(((last - first) / (n - 1) + 0.5e-9) * 1e9) & 0xFFFFFFFF
    m = np.float32(csv.max_abs(HEADER_LINES, n, ntrks, scalefactor))
    mv = np.float32(np.float32(int(np.float32(np.float32(m + np.float32(0.55)) * np.float32(10.0)))) / np.float32(10.0))
    if subsample > 1:
        tstart += ((subsample - 1) * tdelta) & 0xFFFFFFFF          # csvtbin.c:653-654: unsigned 32-bit products, wrap included
        tdelta = (tdelta * subsample) & 0xFFFFFFFF
    given = np.float32(maxvolts_given)
    if given == 0 or given < mv:
        given = mv
    return Preread(tstart, tdelta, given, n)


def convert(csv: "abi.Csv", ntrks: int, order=None, scalefactor: float = 1.0, invert: bool = False, subsample: int = 1,
            maxvolts: float = 0.0, skip: int = 0, stopaft: int | None = None, tape=None, want_rows: bool = True):
    """(Preread, rows, CsvStats): csv_preread + one pass of write_tbin's loop (no -redo)"""
    pre = preread(csv, ntrks, scalefactor, subsample, maxvolts)
    first_line = HEADER_LINES + skip
    nrows = max(0, (csv.nlines - first_line) // subsample)
    if stopaft is not None:
        nrows = min(nrows, stopaft)
    cfg = abi.make_csv_cfg(ntrks, float(pre.maxvolts), order, scalefactor, invert, subsample)
    rows, st = csv.convert(cfg, first_line, nrows, tape=tape, want_rows=want_rows)
    return pre, rows, st
