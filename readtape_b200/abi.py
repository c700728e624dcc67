"""ctypes binding of the C-ABI declared in include/rt_scan.h.

Two shared libraries export that ABI:
  * readtape_b200/lib/librt_scan_b200.so  -- the product (CUDA, sm_100a)
  * oracle/_ref/libscan_oracle.so          -- the test-only CPU oracle
`load_product()` fails loudly when the CUDA library is missing: there is no CPU fallback.
`load_oracle()` may only be used by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import tbin as _tbin

RT_MAXTRKS = 19
RT_HEAD_IGNORE = RT_MAXTRKS - 1
RT_OK, RT_MISS = 0, 1
RT_ERR_UNSUPPORTED = -4

RT_F_FIND_ZEROS, RT_F_DIFFERENTIATE, RT_F_INVERT, RT_F_DENSITY_DETECT, RT_F_DESKEWING = 1, 2, 4, 8, 16
RT_RESET_NONE, RT_RESET_FULL, RT_RESET_WW_PARTIAL, RT_RESET_PEAKSTATE = 0, 1, 2, 3

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_LIB = os.path.join(_ROOT, "readtape_b200", "lib", "librt_scan_b200.so")
ORACLE_LIB = os.path.join(_ROOT, "oracle", "_ref", "libscan_oracle.so")


class TapeDesc(C.Structure):
    _fields_ = [("ntrks", C.c_uint32), ("nheads", C.c_uint32), ("head_to_trk", C.c_int32 * RT_MAXTRKS),
                ("maxvolts", C.c_float), ("tdelta_ns", C.c_uint64), ("tstart_ns", C.c_uint64)]


class Parms(C.Structure):
    _fields_ = [("clk_window", C.c_int32), ("clk_alpha", C.c_float), ("agc_window", C.c_int32),
                ("agc_alpha", C.c_float), ("min_peak", C.c_float), ("clk_factor", C.c_float),
                ("pulse_adj", C.c_float), ("pkww_bitfrac", C.c_float), ("pkww_rise", C.c_float),
                ("z1pt", C.c_float), ("z2pt", C.c_float)]


class ScanCfg(C.Structure):
    _fields_ = [("mode", C.c_int32), ("flags", C.c_uint32), ("bpi", C.c_float), ("ips", C.c_float),
                ("parms", Parms), ("skew_delaycnt", C.c_int32 * RT_MAXTRKS)]


class BulkStats(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("units", C.c_uint64), ("events", C.c_uint64),
                ("rows_scanned", C.c_uint64), ("track_samples", C.c_uint64),
                ("ms_preprocess", C.c_double), ("ms_units", C.c_double), ("ms_scan", C.c_double),
                ("launches", C.c_uint32), ("pad", C.c_uint32), ("d2h_bytes", C.c_uint64), ("ms_masks", C.c_double),
                ("ms_records", C.c_double), ("masks_fused", C.c_uint32), ("two_pass", C.c_uint32),
                ("launches_ingest", C.c_uint32), ("pad2", C.c_uint32)]


class UnitInfo(C.Structure):
    _fields_ = [("unit_index", C.c_uint64), ("nunits", C.c_uint64), ("row0", C.c_uint64), ("row_end", C.c_uint64),
                ("ntrks", C.c_uint32), ("pad", C.c_uint32),
                ("first_event_row", C.c_uint64 * RT_MAXTRKS), ("sync_row", C.c_uint64 * RT_MAXTRKS),
                ("last_loud_row", C.c_uint64 * RT_MAXTRKS), ("need_sync_row", C.c_uint64 * RT_MAXTRKS),
                ("sync_early", C.c_uint64 * RT_MAXTRKS), ("loud_early", C.c_uint64 * RT_MAXTRKS),
                ("sync_first", C.c_uint64 * RT_MAXTRKS), ("quiet_from", C.c_uint64 * RT_MAXTRKS),
                ("nevents", C.c_uint32 * RT_MAXTRKS), ("failed", C.c_uint32 * RT_MAXTRKS)]


EVENT_DTYPE = np.dtype([("row", "<u8"), ("t_event", "<f8"), ("v_top", "<f4"), ("v_bot", "<f4"),
                        ("agc_gain", "<f4"), ("trk", "u1"), ("kind", "u1"), ("pad", "u1", (2,))])
assert EVENT_DTYPE.itemsize == 32

EXPORTS = ["rt_last_error", "rt_abi_version", "rt_backend", "rt_open", "rt_upload", "rt_upload_fd", "rt_attach_device", "rt_prepare", "rt_clear",
           "rt_nrows", "rt_close", "rt_host_alloc", "rt_host_free", "rt_scan_begin", "rt_scan_reset",
           "rt_scan_run", "rt_scan_rewind", "rt_scan_set_avg_height", "rt_scan_set_cfg", "rt_scan_pos", "rt_scan_end",
           "rt_bulk_scan", "rt_bulk_scan_host", "rt_bulk_fetch", "rt_bulk_results_size", "rt_bulk_results_to_device", "rt_bulk_fetch_to", "rt_host_register", "rt_host_unregister", "rt_bulk_lookup", "rt_bulk_unit_info", "rt_bulk_unit_at", "rt_bulk_get_stats", "rt_bulk_free", "rt_bulk_tile_digest", "rt_bulk_last_unit", "rt_set_option", "rt_peak_masks", "rt_pkww_width",
           "rt_row_time",
           "rt_csv_open", "rt_csv_close", "rt_csv_nlines", "rt_csv_line", "rt_csv_max_abs", "rt_csv_convert"]       # include/rt_csv.h


class CsvCfg(C.Structure):
    """rt_csv_cfg (include/rt_csv.h)"""
    _fields_ = [("ntrks", C.c_uint32), ("track_permutation", C.c_uint32 * RT_MAXTRKS), ("maxvolts", C.c_float),
                ("scalefactor", C.c_float), ("invert", C.c_uint32), ("subsample", C.c_uint32)]


class CsvStats(C.Structure):
    """rt_csv_stats (include/rt_csv.h)"""
    _fields_ = [("rows", C.c_uint64), ("too_big", C.c_uint64), ("too_small", C.c_uint64), ("minvolts", C.c_float),
                ("maxvolts", C.c_float), ("ms_convert", C.c_double)]


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rt_scan error {code}: {msg}")
        self.code = code


class Lib:
    """A loaded implementation of the rt_scan C-ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing -- build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "readtape_b200 has no CPU fallback.")
        self.path = path
        L = self.L = C.CDLL(path)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        P = C.POINTER
        L.rt_last_error.restype = C.c_char_p
        L.rt_backend.restype = C.c_char_p
        L.rt_abi_version.restype = i32
        L.rt_open.argtypes = [P(TapeDesc), i32, P(vp)]
        L.rt_upload.argtypes = [vp, vp, u64]
        L.rt_attach_device.argtypes = [vp, vp, u64]
        L.rt_clear.argtypes = [vp]
        L.rt_prepare.argtypes = [vp, P(ScanCfg)]; L.rt_prepare.restype = i32
        L.rt_bulk_fetch.argtypes = [vp]
        L.rt_nrows.argtypes = [vp]; L.rt_nrows.restype = u64
        L.rt_close.argtypes = [vp]; L.rt_close.restype = None
        L.rt_host_alloc.argtypes = [C.c_size_t]; L.rt_host_alloc.restype = vp
        L.rt_host_free.argtypes = [vp]; L.rt_host_free.restype = None
        L.rt_scan_begin.argtypes = [vp, P(ScanCfg), P(vp)]
        L.rt_scan_reset.argtypes = [vp, i32, u64]
        L.rt_scan_run.argtypes = [vp, u64, P(vp), P(u64), P(u64)]
        L.rt_scan_rewind.argtypes = [vp, u64]
        L.rt_scan_set_avg_height.argtypes = [vp, u32, C.c_float]
        L.rt_scan_set_cfg.argtypes = [vp, P(ScanCfg)]
        L.rt_scan_pos.argtypes = [vp]; L.rt_scan_pos.restype = u64
        L.rt_scan_end.argtypes = [vp]; L.rt_scan_end.restype = None
        L.rt_bulk_scan.argtypes = [vp, P(ScanCfg), u32, P(vp)]
        L.rt_bulk_scan_host.argtypes = [vp, vp, u64, P(ScanCfg), P(vp)]
        L.rt_bulk_lookup.argtypes = [vp, u32, u64, P(vp), P(u64), P(u64)]
        L.rt_bulk_get_stats.argtypes = [vp, P(BulkStats)]
        L.rt_bulk_unit_info.argtypes = [vp, u32, u64, P(UnitInfo)]
        L.rt_bulk_unit_at.argtypes = [vp, u32, u64, P(UnitInfo)]
        L.rt_bulk_free.argtypes = [vp]; L.rt_bulk_free.restype = None
        L.rt_bulk_results_size.argtypes = [vp, P(u64)]; L.rt_bulk_results_size.restype = i32
        L.rt_bulk_results_to_device.argtypes = [vp, vp, u64]; L.rt_bulk_results_to_device.restype = i32
        L.rt_bulk_tile_digest.argtypes = [vp, u32, u64, u64, vp, vp, P(u64)]; L.rt_bulk_tile_digest.restype = i32
        L.rt_peak_masks.argtypes = [vp, P(ScanCfg), C.c_float, vp, vp, u64, P(C.c_int32)]
        L.rt_pkww_width.argtypes = [P(ScanCfg), u64]
        L.rt_row_time.argtypes = [P(TapeDesc), u64]; L.rt_row_time.restype = C.c_double
        L.rt_csv_open.argtypes = [i32, vp, u64, P(vp)]; L.rt_csv_open.restype = i32
        L.rt_csv_close.argtypes = [vp]; L.rt_csv_close.restype = None
        L.rt_csv_nlines.argtypes = [vp]; L.rt_csv_nlines.restype = u64
        L.rt_csv_line.argtypes = [vp, u64, P(u64), P(u64)]; L.rt_csv_line.restype = i32
        L.rt_csv_max_abs.argtypes = [vp, u64, u64, u32, C.c_float, P(C.c_float)]; L.rt_csv_max_abs.restype = i32
        L.rt_csv_convert.argtypes = [vp, P(CsvCfg), u64, u64, vp, vp, P(CsvStats)]; L.rt_csv_convert.restype = i32
        for fn in ("rt_open", "rt_upload", "rt_upload_fd", "rt_attach_device", "rt_prepare", "rt_clear", "rt_bulk_fetch", "rt_scan_begin", "rt_scan_reset", "rt_scan_run",
                   "rt_scan_rewind", "rt_scan_set_avg_height", "rt_scan_set_cfg", "rt_bulk_scan", "rt_bulk_scan_host", "rt_bulk_lookup",
                   "rt_bulk_get_stats", "rt_bulk_unit_info", "rt_bulk_unit_at", "rt_pkww_width", "rt_peak_masks"):
            getattr(L, fn).restype = i32

    @property
    def backend(self) -> str:
        return self.L.rt_backend().decode()

    def check(self, rc: int) -> int:
        if rc < 0:
            raise RtError(rc, self.L.rt_last_error().decode(errors="replace"))
        return rc

    def open(self, desc: TapeDesc, device: int = 0) -> "Tape":
        h = C.c_void_p()
        self.check(self.L.rt_open(C.byref(desc), device, C.byref(h)))
        return Tape(self, h, desc)


    def csv_open(self, text, device: int = 0) -> "Csv":
        """CSV text (bytes / bytearray / uint8 array / mmap) onto the device, lines indexed (rt_csv_open)"""
        buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
        h = C.c_void_p()
        self.check(self.L.rt_csv_open(device, buf.ctypes.data, buf.size, C.byref(h)))
        return Csv(self, h, buf)


class Csv:
    """A CSV capture on the device (include/rt_csv.h).  `text` keeps the host copy alive for line()."""

    def __init__(self, lib: "Lib", handle, text: np.ndarray):
        self.lib, self.h, self.text = lib, handle, text

    def close(self) -> None:
        if self.h:
            self.lib.L.rt_csv_close(self.h); self.h = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()

    @property
    def nlines(self) -> int:
        return int(self.lib.L.rt_csv_nlines(self.h))

    def line(self, i: int) -> bytes:
        off, ln = C.c_uint64(), C.c_uint64()
        self.lib.check(self.lib.L.rt_csv_line(self.h, i, C.byref(off), C.byref(ln)))
        return bytes(self.text[off.value:off.value + ln.value])

    def max_abs(self, first_line: int, nlines: int, ntrks: int, scalefactor: float = 1.0) -> np.float32:
        out = C.c_float()
        self.lib.check(self.lib.L.rt_csv_max_abs(self.h, first_line, nlines, ntrks, scalefactor, C.byref(out)))
        return np.float32(out.value)

    def convert(self, cfg: CsvCfg, first_line: int, nrows: int, tape: "Tape | None" = None, want_rows: bool = True):
        """(rows int16 [nrows, ntrks] or None, CsvStats): the conversion loop of write_tbin (rt_csv_convert)"""
        rows = np.empty((nrows, cfg.ntrks), dtype="<i2") if want_rows else None
        st = CsvStats()
        self.lib.check(self.lib.L.rt_csv_convert(self.h, C.byref(cfg), first_line, nrows, rows.ctypes.data if want_rows and nrows else None,
                                                 tape.h if tape is not None else None, C.byref(st)))
        return rows, st


def make_csv_cfg(ntrks: int, maxvolts: float, order=None, scalefactor: float = 1.0, invert: bool = False, subsample: int = 1) -> CsvCfg:
    c = CsvCfg()
    c.ntrks = ntrks; c.maxvolts = maxvolts; c.scalefactor = scalefactor; c.invert = int(invert); c.subsample = subsample
    for i in range(ntrks):
        c.track_permutation[i] = i if order is None else order[i]
    return c


def _events_from(ptr: C.c_void_p, n: int) -> np.ndarray:
    if n == 0 or not ptr.value:
        return np.zeros(0, dtype=EVENT_DTYPE)
    buf = (C.c_char * (n * EVENT_DTYPE.itemsize)).from_address(ptr.value)
    return np.frombuffer(buf, dtype=EVENT_DTYPE, count=n).copy()


class Tape:
    def __init__(self, lib: Lib, handle, desc: TapeDesc):
        self.lib, self.h, self.desc = lib, handle, desc

    def upload(self, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype="<i2")
        assert rows.ndim == 2 and rows.shape[1] == self.desc.nheads
        self.lib.check(self.lib.L.rt_upload(self.h, rows.ctypes.data, rows.shape[0]))

    def upload_ptr(self, host_ptr: int, nrows: int) -> None:
        """rows already laid out in (pinned) host memory at `host_ptr`"""
        self.lib.check(self.lib.L.rt_upload(self.h, host_ptr, nrows))

    def attach_device(self, dev_ptr: int, nrows: int) -> None:
        self.lib.check(self.lib.L.rt_attach_device(self.h, dev_ptr, nrows))

    def clear(self) -> None:
        self.lib.check(self.lib.L.rt_clear(self.h))

    def prepare(self, cfg) -> None:
        """announce the configuration of the next whole-tape scan (rt_prepare): uploads then also build its mask planes"""
        self.lib.check(self.lib.L.rt_prepare(self.h, C.byref(cfg) if cfg is not None else None))

    @property
    def nrows(self) -> int:
        return int(self.lib.L.rt_nrows(self.h))

    def scan(self, cfg: ScanCfg) -> "Scan":
        h = C.c_void_p()
        self.lib.check(self.lib.L.rt_scan_begin(self.h, C.byref(cfg), C.byref(h)))
        return Scan(self, h)

    def bulk_scan(self, cfgs) -> "Bulk":
        arr = (ScanCfg * len(cfgs))(*cfgs)
        h = C.c_void_p()
        self.lib.check(self.lib.L.rt_bulk_scan(self.h, arr, len(cfgs), C.byref(h)))
        return Bulk(self, h)

    def peak_masks(self, cfg: ScanCfg, t0_frac: float = 0.25):
        """(cand, acan, T0): the bit planes of the two-pass peak scan, uint32 [ntrks][ceil(nrows/32)] (rt_peak_masks)"""
        import numpy as np
        nt = self.desc.ntrks
        wpt = (self.nrows + 31) // 32 + 1
        cand = np.zeros((nt, wpt), dtype=np.uint32); acan = np.zeros((nt, wpt), dtype=np.uint32)
        t0 = C.c_int32(0)
        self.lib.check(self.lib.L.rt_peak_masks(self.h, C.byref(cfg), t0_frac, cand.ctypes.data, acan.ctypes.data, wpt, C.byref(t0)))
        return cand, acan, int(t0.value)

    def bulk_scan_host(self, host_ptr: int, nrows: int, cfg: ScanCfg) -> "Bulk":
        """clear + upload + whole-tape scan + fetch, overlapped (rows at `host_ptr`, ideally pinned)"""
        h = C.c_void_p()
        self.lib.check(self.lib.L.rt_bulk_scan_host(self.h, host_ptr, nrows, C.byref(cfg), C.byref(h)))
        return Bulk(self, h)

    def close(self) -> None:
        if self.h:
            self.lib.L.rt_close(self.h)
            self.h = None


class Scan:
    def __init__(self, tape: Tape, handle):
        self.tape, self.lib, self.h = tape, tape.lib, handle

    def reset(self, kind: int, row: int) -> None:
        self.lib.check(self.lib.L.rt_scan_reset(self.h, kind, row))

    def run(self, nrows: int):
        ev, n, done = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self.lib.check(self.lib.L.rt_scan_run(self.h, nrows, C.byref(ev), C.byref(n), C.byref(done)))
        return _events_from(ev, n.value), int(done.value)

    def rewind(self, row: int) -> None:
        self.lib.check(self.lib.L.rt_scan_rewind(self.h, row))

    def set_cfg(self, cfg: ScanCfg) -> None:
        self.lib.check(self.lib.L.rt_scan_set_cfg(self.h, C.byref(cfg)))

    def set_avg_height(self, trk: int, v: float) -> None:
        self.lib.check(self.lib.L.rt_scan_set_avg_height(self.h, trk, v))

    @property
    def pos(self) -> int:
        return int(self.lib.L.rt_scan_pos(self.h))

    def end(self) -> None:
        if self.h:
            self.lib.L.rt_scan_end(self.h)
            self.h = None


class Bulk:
    def __init__(self, tape: Tape, handle):
        self.tape, self.lib, self.h = tape, tape.lib, handle

    def lookup(self, cfg_index: int, start_row: int):
        """-> (events, valid_rows) or None on RT_MISS."""
        ev, n, valid = C.c_void_p(), C.c_uint64(), C.c_uint64()
        rc = self.lib.check(self.lib.L.rt_bulk_lookup(self.h, cfg_index, start_row, C.byref(ev), C.byref(n),
                                                      C.byref(valid)))
        if rc == RT_MISS:
            return None
        return _events_from(ev, n.value), int(valid.value)

    def fetch(self) -> None:
        self.lib.check(self.lib.L.rt_bulk_fetch(self.h))

    @staticmethod
    def _unit_dict(ui: UnitInfo) -> dict:
        n = ui.ntrks
        none = (1 << 64) - 1
        f = lambda a: [None if a[i] == none else int(a[i]) for i in range(n)]
        return {"unit_index": int(ui.unit_index), "nunits": int(ui.nunits), "row0": int(ui.row0), "row_end": int(ui.row_end),
                "first_event_row": f(ui.first_event_row), "sync_row": f(ui.sync_row), "last_loud_row": f(ui.last_loud_row),
                "need_sync_row": f(ui.need_sync_row), "sync_early": f(ui.sync_early), "loud_early": f(ui.loud_early),
                "sync_first": f(ui.sync_first), "quiet_from": f(ui.quiet_from),
                "nevents": [int(ui.nevents[i]) for i in range(n)], "failed": [int(ui.failed[i]) for i in range(n)]}

    def unit_info(self, cfg_index: int, start_row: int) -> dict:
        ui = UnitInfo()
        self.lib.check(self.lib.L.rt_bulk_unit_info(self.h, cfg_index, start_row, C.byref(ui)))
        return self._unit_dict(ui)

    def unit_at(self, cfg_index: int, unit_index: int):
        """proof data of unit number `unit_index`, or None past the last unit"""
        ui = UnitInfo()
        rc = self.lib.check(self.lib.L.rt_bulk_unit_at(self.h, cfg_index, unit_index, C.byref(ui)))
        return None if rc == RT_MISS else self._unit_dict(ui)

    def results_size(self) -> int:
        n = C.c_uint64(0)
        self.lib.check(self.lib.L.rt_bulk_results_size(self.h, C.byref(n)))
        return int(n.value)

    def results_to_device(self, dst_dev_ptr: int, nbytes: int) -> None:
        """the scan's results (unit tables, proof data, chunk links, events) as one image, device to device (rt_bulk_results_to_device)"""
        self.lib.check(self.lib.L.rt_bulk_results_to_device(self.h, dst_dev_ptr, nbytes))

    def tile_digest(self, cfg_index: int, period_rows: int, ntiles: int):
        """(events[ntiles], digest[ntiles], bad_times): per-tile event counts and order-independent digests, computed on the
        device (rt_bulk_tile_digest); call before fetch()/lookup()"""
        ev = np.zeros(ntiles, dtype=np.uint64); dg = np.zeros(ntiles, dtype=np.uint64); bad = C.c_uint64(0)
        self.lib.check(self.lib.L.rt_bulk_tile_digest(self.h, cfg_index, period_rows, ntiles, ev.ctypes.data, dg.ctypes.data, C.byref(bad)))
        return ev, dg, int(bad.value)

    def stats(self) -> BulkStats:
        st = BulkStats()
        self.lib.check(self.lib.L.rt_bulk_get_stats(self.h, C.byref(st)))
        return st

    def free(self) -> None:
        if self.h:
            self.lib.L.rt_bulk_free(self.h)
            self.h = None


_cache: dict[str, Lib] = {}


def load(path: str) -> Lib:
    if path not in _cache:
        _cache[path] = Lib(path)
    return _cache[path]


def load_product() -> Lib:
    """The CUDA library.  Raises if it has not been built -- there is no fallback."""
    return load(PRODUCT_LIB)


def load_oracle() -> Lib:
    """TEST INFRASTRUCTURE: the CPU oracle (see oracle/scan_oracle.c)."""
    return load(ORACLE_LIB)


# ---- helpers to build the PODs -------------------------------------------------------------
def make_desc(ntrks: int, maxvolts: float, tdelta_ns: int, tstart_ns: int, nheads: int | None = None,
              head_to_trk=None) -> TapeDesc:
    d = TapeDesc()
    d.ntrks = ntrks
    d.nheads = nheads if nheads is not None else ntrks
    h2t = list(head_to_trk) if head_to_trk is not None else list(range(d.nheads))
    for i in range(RT_MAXTRKS):
        d.head_to_trk[i] = h2t[i] if i < len(h2t) else RT_HEAD_IGNORE
    d.maxvolts = maxvolts
    d.tdelta_ns = tdelta_ns
    d.tstart_ns = tstart_ns
    return d


def desc_from_header(hdr: _tbin.TbinHeader, ntrks: int | None = None, order: str | None = None) -> TapeDesc:
    """Mirror of process_file()'s head/track bookkeeping (readtape.c:1646-1648, parse_track_order :877)."""
    n = ntrks or hdr.ntrks
    if hdr.mode == _tbin.MODE_WW or (order and any(c in "CLMclmx" for c in order) and not order[0].isdigit()):
        order = order or hdr.trkorder
        h2t, nt = [], 0
        for ch in order:
            if ch == "x":
                h2t.append(RT_HEAD_IGNORE)
            else:
                h2t.append(nt)
                nt += 1
        return make_desc(nt, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns, nheads=len(order), head_to_trk=h2t)
    h2t = list(range(n))
    if order and (hdr.flags & _tbin.TBIN_NO_REORDER):   # a permutation given with -order= is only honoured
        h2t = []                                        # when csvtbin did not already reorder (readtape.c:1646-1648)
        for ch in order:
            h2t.append(n - 1 if ch in "pP" else int(ch))
    return make_desc(n, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns, head_to_trk=h2t)


def make_cfg(mode: int, parms: dict, bpi: float, ips: float, flags: int = 0, skew=None) -> ScanCfg:
    c = ScanCfg()
    c.mode, c.flags, c.bpi, c.ips = mode, flags, bpi, ips
    for name, _ in Parms._fields_:
        setattr(c.parms, name, parms.get(name, 0))
    for i in range(RT_MAXTRKS):
        c.skew_delaycnt[i] = int(skew[i]) if skew is not None and i < len(skew) else 0
    return c
