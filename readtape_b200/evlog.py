"""Reference event logs: parse, replay through an rt_scan implementation, digest.

TEST/BENCH HARNESS.  `oracle/evdump_shim.c` makes the unmodified reference dump one 64-byte
record per reset / flux transition / block end.  This module turns such a log into a list of
*decode segments* (one per `readblock()` call of the reference: where the per-track state was
reset, with which parameter set and globals, and which events the reference produced before
its end-of-block logic stopped looking), and can drive any implementation of the C-ABI
(`include/rt_scan.h`) through exactly the same reset sequence to compare event streams.
"""
from __future__ import annotations

import dataclasses
import hashlib
import json

import numpy as np

from . import abi, parmsets
from .tbin import MODE_WW

REC = np.dtype([("type", "u1"), ("trk", "u1"), ("kind", "u1"), ("flags", "u1"), ("parmset", "<i4"),
                ("row", "<u8"), ("t", "<f8"), ("v_top", "<f4"), ("v_bot", "<f4"), ("agc_pre", "<f4"),
                ("agc_post", "<f4"), ("avgh", "<f4"), ("clk", "<f4"), ("peakcount", "<i4"), ("w", "<i4"),
                ("timenow", "<f8")])
CFGREC = np.dtype([("type", "u1"), ("ntrks", "u1"), ("find_zeros", "u1"), ("differentiate", "u1"),
                   ("parmset", "<i4"), ("row", "<u8"), ("bpi", "<f4"), ("ips", "<f4"), ("mode", "<i4"),
                   ("skew", "i1", (19,)), ("invert", "u1"), ("pad", "u1", (16,))])
TRKFREC = np.dtype([("type", "u1"), ("ntrks", "u1"), ("pad0", "u1", (2,)), ("parmset", "<i4"), ("row", "<u8"),
                    ("avg_height", "<f4", (10,)), ("pad", "u1", (8,))])
HEADREC = np.dtype([("type", "u1"), ("nheads", "u1"), ("ntrks", "u1"), ("pad0", "u1"), ("subsample", "<i4"),
                    ("tstart_ns", "<u8"), ("tdelta_ns", "<u8"), ("maxvolts", "<f4"), ("head_to_trk", "i1", (19,)),
                    ("pad", "u1", (17,))])
assert REC.itemsize == CFGREC.itemsize == TRKFREC.itemsize == HEADREC.itemsize == 64

T_RESET, T_EVENT, T_BLKEND, T_DENS, T_CFG, T_TRKF, T_IBG, T_HEADS = 1, 2, 3, 4, 5, 6, 7, 8

# canonical event bytes used for digests: what both sides must agree on bit-for-bit
CANON = np.dtype([("row", "<u8"), ("t_event", "<f8"), ("v_top", "<f4"), ("v_bot", "<f4"), ("agc_gain", "<f4"),
                  ("trk", "u1"), ("kind", "u1")])


@dataclasses.dataclass
class Segment:
    reset_kind: int          # abi.RT_RESET_*
    row: int                 # first row read after the reset
    parmset: int
    mode: int
    flags: int               # abi.RT_F_*
    bpi: float
    ips: float
    skew: list
    avg_height: list | None  # per-track v_avg_height in force after the reset (Whirlwind carry-in)
    set_avg_height: bool     # push avg_height into the scan state (after compute_avg_height)
    end_row: int             # rows [row, end_row) were read by this readblock(); -1 = to end of tape
    stop_row: int            # end-of-block fired on this row: later rows are skipped; -1 = none
    events: np.ndarray       # CANON records from the reference

    def meta(self) -> dict:
        d = dataclasses.asdict(self)
        d.pop("events")
        d["nevents"] = int(len(self.events))
        d["sha256"] = digest(self.events)
        return d


def digest(canon: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(canon).tobytes()).hexdigest()


def to_canon(ev: np.ndarray, agc_field: str = "agc_gain", t_field: str = "t_event") -> np.ndarray:
    out = np.zeros(len(ev), dtype=CANON)
    out["row"] = ev["row"]
    out["t_event"] = ev[t_field]
    out["v_top"] = ev["v_top"]
    out["v_bot"] = ev["v_bot"]
    out["agc_gain"] = ev[agc_field]
    out["trk"] = ev["trk"]
    out["kind"] = ev["kind"]
    return out


def parse_heads(path: str) -> dict:
    """The sample-stream description the reference derived (first record of the log)."""
    raw = np.fromfile(path, dtype=HEADREC, count=1)
    h = raw[0]
    assert h["type"] == T_HEADS
    nh = int(h["nheads"])
    return {"nheads": nh, "ntrks": int(h["ntrks"]), "subsample": int(h["subsample"]),
            "tstart_ns": int(h["tstart_ns"]), "tdelta_ns": int(h["tdelta_ns"]), "maxvolts": float(h["maxvolts"]),
            "head_to_trk": [int(x) for x in h["head_to_trk"][:nh]]}


def desc_from_heads(h: dict) -> "abi.TapeDesc":
    return abi.make_desc(h["ntrks"], h["maxvolts"], h["tdelta_ns"], h["tstart_ns"], nheads=h["nheads"],
                         head_to_trk=h["head_to_trk"])


def parse(path: str) -> list[Segment]:
    raw = np.fromfile(path, dtype=REC)
    types = raw["type"]
    segs: list[Segment] = []
    cur = None
    cur_events: list[int] = []
    pending_avg = None       # TRKF seen after compute_avg_height (not directly after a reset)
    last_was_reset = False

    def close(end_row):
        nonlocal cur, cur_events
        if cur is None:
            return
        evs = raw[cur_events] if cur_events else raw[:0]
        cur.events = to_canon(evs, agc_field="agc_post", t_field="t")
        if cur.end_row < 0 and end_row is not None:
            cur.end_row = end_row
        segs.append(cur)
        cur, cur_events = None, []

    i = 0
    n = len(raw)
    while i < n:
        t = types[i]
        if t == T_RESET:
            r = raw[i]
            kind = {0: abi.RT_RESET_FULL, 1: abi.RT_RESET_WW_PARTIAL, 2: abi.RT_RESET_PEAKSTATE}[int(r["kind"])]
            cfg = raw[i + 1:i + 2].view(CFGREC)[0]
            trkf = raw[i + 2:i + 3].view(TRKFREC)[0]
            assert cfg["type"] == T_CFG and trkf["type"] == T_TRKF
            if cur is not None:
                # a reset with no BLKEND in between: a queued Whirlwind block mark (no readblock() call
                # at all, readtape.c:1764-1767) or end of file reached (then the file was rewound)
                close(int(r["row"]) if (int(r["row"]) == cur.row and not cur_events) else None)
            flags = 0
            if cfg["find_zeros"]: flags |= abi.RT_F_FIND_ZEROS
            if cfg["differentiate"]: flags |= abi.RT_F_DIFFERENTIATE
            if cfg["invert"]: flags |= abi.RT_F_INVERT
            if r["flags"] & 1: flags |= abi.RT_F_DENSITY_DETECT
            if r["flags"] & 2: flags |= abi.RT_F_DESKEWING
            nt = int(cfg["ntrks"])
            cur = Segment(reset_kind=kind, row=int(r["row"]), parmset=int(r["parmset"]), mode=int(cfg["mode"]),
                          flags=flags, bpi=float(cfg["bpi"]), ips=float(cfg["ips"]),
                          skew=[int(x) for x in cfg["skew"][:nt]],
                          avg_height=[float(x) for x in trkf["avg_height"][:min(nt, 10)]],
                          set_avg_height=pending_avg is not None, end_row=-1, stop_row=-1,
                          events=None)
            if pending_avg is not None:
                cur.avg_height = pending_avg
                pending_avg = None
            i += 3
            continue
        if t == T_TRKF:        # compute_avg_height() ran (Whirlwind, after the deskew pre-pass)
            f = raw[i:i + 1].view(TRKFREC)[0]
            pending_avg = [float(x) for x in f["avg_height"][:min(int(f["ntrks"]), 10)]]
            if cur is not None:
                cur.avg_height = pending_avg
                cur.set_avg_height = True
                pending_avg = None
        elif t in (T_EVENT, T_DENS):
            if cur is not None:
                cur_events.append(i)
        elif t == T_IBG:
            if cur is not None and cur.stop_row < 0:
                cur.stop_row = int(raw[i]["row"])
        elif t == T_BLKEND:
            if cur is not None:
                cur.end_row = int(raw[i]["row"]) + 1
                close(cur.end_row)
        i += 1
    close(None)
    return segs


def cfg_for(seg: Segment, parmtable=None):
    table = parmtable or parmsets.BUILTIN[seg.mode]
    return abi.make_cfg(seg.mode, table[seg.parmset], seg.bpi, seg.ips, flags=seg.flags, skew=seg.skew)


def replay(tape: "abi.Tape", segs: list[Segment], parmtable=None, chunk: int | None = None):
    """Drive an rt_scan implementation through the reference's reset sequence.

    Yields (segment, canonical events produced for rows [seg.row, seg.end_row), clipped at
    seg.stop_row the way the reference's interblock skip does)."""
    nrows = tape.nrows
    scan = None
    scan_key = None
    for seg in segs:
        key = (seg.mode, seg.parmset, seg.flags, seg.bpi, seg.ips, tuple(seg.skew))
        persistent = seg.reset_kind != abi.RT_RESET_FULL
        if scan is None or (key != scan_key and not persistent):
            if scan is not None:
                scan.end()
            scan = tape.scan(cfg_for(seg, parmtable))
            scan_key = key
        elif key != scan_key:       # persistent state (Whirlwind), new globals: e.g. skew set after the pre-pass
            scan.set_cfg(cfg_for(seg, parmtable))
            scan_key = key
        if seg.set_avg_height and seg.avg_height:
            for k, v in enumerate(seg.avg_height):
                scan.set_avg_height(k, v)
        scan.reset(seg.reset_kind, seg.row)
        end = seg.end_row if seg.end_row >= 0 else nrows
        todo = max(0, end - seg.row)
        parts = []
        while True:
            step = todo if chunk is None else min(todo, chunk)
            ev, done = scan.run(step)
            parts.append(ev)
            todo -= done
            if todo <= 0 or done == 0:
                break
        ev = np.concatenate(parts) if len(parts) > 1 else parts[0]
        canon = to_canon(ev)
        if seg.stop_row >= 0:
            canon = canon[canon["row"] <= seg.stop_row]
        yield seg, canon
    if scan is not None:
        scan.end()


def compare(seg: Segment, got: np.ndarray) -> str | None:
    """None if `got` matches the reference's events of this segment, else a description."""
    ref = seg.events
    if seg.stop_row >= 0 and len(got) != len(ref):
        # the reference's end-of-block can fire in the middle of the track loop of row stop_row
        # (decoder.c:886-888 `goto exit`, :876): tracks after it are not looked at on that row
        keep = np.ones(len(got), dtype=bool)
        at = np.nonzero(got["row"] == seg.stop_row)[0]
        refat = ref[ref["row"] == seg.stop_row]
        for j in at:
            if not np.any(refat["trk"] == got["trk"][j]):
                keep[j] = False
        got = got[keep]
    if len(got) != len(ref):
        k = min(len(got), len(ref))
        bad = np.nonzero(got[:k].tobytes() != ref[:k].tobytes())[0] if False else None
        first = _first_diff(got[:k], ref[:k])
        return (f"event count {len(got)} != reference {len(ref)} (segment row {seg.row}, parmset {seg.parmset}); "
                f"first difference at #{first}: got {got[first] if first < len(got) else None} "
                f"ref {ref[first] if first < len(ref) else None}")
    first = _first_diff(got, ref)
    if first < len(ref):
        return (f"event #{first} differs (segment row {seg.row}, parmset {seg.parmset}): "
                f"got {got[first]} ref {ref[first]}")
    return None


def matches_fixture(seg: Segment, got: np.ndarray) -> bool:
    """`got` (canonical events of a replayed segment) against a fixture segment (digest only).

    When the reference's end-of-block fires inside the track loop of row stop_row (decoder.c:876,
    :886-888 `goto exit`) the tracks after it are not looked at on that row, so a free-running
    scan may legitimately hold extra events AT stop_row: they are dropped from the end."""
    if len(got) == seg.nevents:
        return digest(got) == seg.sha256
    if seg.stop_row < 0 or len(got) < seg.nevents:
        return False
    extra = got[seg.nevents:]
    if not np.all(extra["row"] == seg.stop_row):
        return False
    return digest(got[: seg.nevents]) == seg.sha256


def _first_diff(a: np.ndarray, b: np.ndarray) -> int:
    k = min(len(a), len(b))
    if k == 0:
        return 0
    av = np.ascontiguousarray(a[:k]).view(np.uint8).reshape(k, -1)
    bv = np.ascontiguousarray(b[:k]).view(np.uint8).reshape(k, -1)
    ne = np.nonzero((av != bv).any(axis=1))[0]
    return int(ne[0]) if len(ne) else k


def save_fixture(path: str, segs: list[Segment], extra: dict | None = None) -> None:
    doc = {"segments": [s.meta() for s in segs]}
    if extra:
        doc.update(extra)
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=0, separators=(",", ":"))
        fh.write("\n")


def load_fixture(path: str):
    with open(path) as fh:
        doc = json.load(fh)
    segs = []
    for m in doc["segments"]:
        m = dict(m)
        sha, nev = m.pop("sha256"), m.pop("nevents")
        s = Segment(events=None, **m)
        s.sha256, s.nevents = sha, nev
        segs.append(s)
    return doc, segs


# ---- the tile digest of rt_bulk_tile_digest (include/rt_scan.h), restated with numpy for the checker side ----------------
def _mix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30); x *= np.uint64(0xbf58476d1ce4e5b9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94d049bb133111eb)
        x ^= x >> np.uint64(31)
    return x


def tile_digest(ev: np.ndarray, tile_index: int, period_rows: int, tstart_ns: int, tdelta_ns: int):
    """(count, digest) of the events `ev` (abi.EVENT_DTYPE or CANON records) that lie in tile `tile_index`"""
    lo = tile_index * period_rows
    e = ev[(ev["row"] >= lo) & (ev["row"] < lo + period_rows)]
    if len(e) == 0:
        return 0, 0
    ns = (np.uint64(tstart_ns) + e["row"].astype(np.uint64) * np.uint64(tdelta_ns)).astype(np.int64)
    now = ns.astype(np.float64) / 1e9
    sd = np.float64(np.float32(np.float32(tdelta_ns) / np.float32(1e9)))          # sample_deltat as the reference holds it (a float)
    hsd = (now - e["t_event"]) / (sd * 0.5)
    hs = np.where(hsd < 0, hsd - 0.5, hsd + 0.5).astype(np.int64)
    k0 = (e["row"].astype(np.uint64) - np.uint64(lo)) | (e["trk"].astype(np.uint64) << np.uint64(40)) \
        | (e["kind"].astype(np.uint64) << np.uint64(48)) | ((hs.astype(np.uint64) & np.uint64(0xfff)) << np.uint64(52))
    k1 = e["v_top"].view("<u4").astype(np.uint64) | (e["v_bot"].view("<u4").astype(np.uint64) << np.uint64(32))
    k2 = e["agc_gain"].view("<u4").astype(np.uint64)
    h = _mix64(_mix64(_mix64(k0) ^ k1) ^ k2)
    with np.errstate(over="ignore"):
        return int(len(e)), int(np.sum(h, dtype=np.uint64))
