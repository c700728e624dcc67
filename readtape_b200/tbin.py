"""TBIN container reader/writer (host side of the sample ingest).

Layout follows the reference's on-disk format, src/csvtbin.h:50-105:
  240-byte `tbin_hdr_t`  (tag "TBINHDR", descr[80], 4-byte little-endian fields)
  optional 28-byte `tbin_hdrext_trkorder_t` (tag "TBINORD") when flags & TBIN_TRKORDER_INCLUDED
  16-byte `tbin_dat_t`   (tag "DAT", sample_bits, tstart ns)
  rows of `nheads` little-endian int16, terminated by the single value -32768.
The header is parsed the way read_tbin_header() does (src/readtape.c:1319-1376).
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

HDR_TAG = b"TBINHDR\0"
ORD_TAG = b"TBINORD\0"
DAT_TAG = b"DAT\0"
TBIN_NO_REORDER = 0x01
TBIN_TRKORDER_INCLUDED = 0x02
TBIN_INVERTED = 0x04
TBIN_REVERSED = 0x08
END_MARK = -32768

MODE_UNKNOWN, MODE_PE, MODE_NRZI, MODE_GCR, MODE_WW = 0, 1, 2, 4, 8
MODE_NAMES = {MODE_PE: "PE", MODE_NRZI: "NRZI", MODE_GCR: "GCR", MODE_WW: "Whirlwind"}

_HDR = struct.Struct("<8s80sII27iIIIfIIIff")   # 240 bytes
assert _HDR.size == 240
_DAT = struct.Struct("<4sBBBBQ")
assert _DAT.size == 16


@dataclasses.dataclass
class TbinHeader:
    descr: str = ""
    flags: int = 0
    ntrks: int = 9
    tdelta_ns: int = 1280
    maxvolts: float = 4.4
    mode: int = MODE_UNKNOWN
    bpi: float = 0.0
    ips: float = 0.0
    trkorder: str | None = None
    tstart_ns: int = 0
    payload_offset: int = 256
    times: tuple = (0,) * 27


def parse_header(buf: bytes | memoryview) -> TbinHeader:
    f = _HDR.unpack_from(buf, 0)
    tag, descr, hdrsize, fmt = f[0], f[1], f[2], f[3]
    if tag != HDR_TAG:
        raise ValueError(".tbin file missing TBINHDR tag")
    if fmt != 1 or hdrsize != 240:
        raise ValueError(f"bad .tbin header: format {fmt} size {hdrsize}")
    times = f[4:31]
    flags, ntrks, tdelta, maxvolts, _r1, _r2, mode, bpi, ips = f[31:40]
    off = 240
    trkorder = None
    if flags & TBIN_TRKORDER_INCLUDED:
        if bytes(buf[off:off + 8]) != ORD_TAG:
            raise ValueError(".tbin file missing TBINORD tag")
        trkorder = bytes(buf[off + 8:off + 28]).split(b"\0")[0].decode("ascii")
        off += 28
    dtag, _opt, bits, _a, _b, tstart = _DAT.unpack_from(buf, off)
    if dtag != DAT_TAG:
        raise ValueError(".tbin file missing DAT tag")
    if bits != 16:
        raise ValueError(f"only 16-bit samples are supported, not {bits}")
    off += 16
    return TbinHeader(descr=descr.split(b"\0")[0].decode("latin1"), flags=flags, ntrks=ntrks,
                      tdelta_ns=tdelta, maxvolts=maxvolts, mode=mode, bpi=bpi, ips=ips,
                      trkorder=trkorder, tstart_ns=tstart, payload_offset=off, times=tuple(times))


def read_tbin(path: str, nheads: int | None = None, mmap: bool = True):
    """Return (header, rows) with rows an int16 array of shape (nrows_in_file, nheads).

    All complete rows stored in the file are returned, including anything after the end marker;
    locating the marker (first row whose head 0 is -32768, readtape.c:1410) is the library's job."""
    with open(path, "rb") as fh:
        head = fh.read(240 + 28 + 16)
    hdr = parse_header(head)
    if nheads is None:
        nheads = len(hdr.trkorder) if hdr.trkorder else hdr.ntrks
    if mmap:
        raw = np.memmap(path, dtype="<i2", mode="r", offset=hdr.payload_offset)
    else:
        raw = np.fromfile(path, dtype="<i2", offset=hdr.payload_offset)
    nrows = raw.shape[0] // nheads
    rows = raw[: nrows * nheads].reshape(nrows, nheads)
    return hdr, rows


def build_header(hdr: TbinHeader) -> bytes:
    out = _HDR.pack(HDR_TAG, hdr.descr.encode("latin1")[:79], 240, 1, *hdr.times,
                    hdr.flags, hdr.ntrks, hdr.tdelta_ns, hdr.maxvolts, 0, 0, hdr.mode, hdr.bpi, hdr.ips)
    if hdr.flags & TBIN_TRKORDER_INCLUDED:
        out += ORD_TAG + (hdr.trkorder or "").encode("ascii")[:19].ljust(20, b"\0")
    out += _DAT.pack(DAT_TAG, 0, 16, 0, 0, hdr.tstart_ns)
    return out


def write_tbin(path: str, hdr: TbinHeader, rows: np.ndarray) -> None:
    """Write rows (nrows, nheads) int16 plus the end marker (csvtbin.c write_tbin, :661-747)."""
    rows = np.ascontiguousarray(rows, dtype="<i2")
    with open(path, "wb") as fh:
        fh.write(build_header(hdr))
        rows.tofile(fh)
        fh.write(struct.pack("<h", END_MARK))
