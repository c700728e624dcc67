"""Synthetic tape captures (TBIN payloads) for benchmarks and size-independent parity tests.

`nrzi_tape()` builds the BASELINE.json config-2 shape specified in SURVEY.md 8(d): 9-track
800 BPI NRZI at 50 IPS sampled at 781.25 kHz (tdelta 1280 ns, maxvolts 4.4), 512 random data
bytes per block with odd parity, the IBM 9-track CRC character 4 character times after the data
and the LRC character 4 after that (the arithmetic the reference checks in nrzi_postprocess(),
src/decode_nrzi.c:35-75, so that the reference decodes every block "ok"), alternating-polarity
raised-cosine flux pulses of half-width 0.35 bit, per-track amplitude 3.0 + 0.1*trk volts,
Gaussian noise sigma 5 mV, DC offset -15 mV, static head skew (trk mod 4) rows, a +-1 % 5 Hz
speed wobble and ~9375-row inter-block gaps.  One "super-tile" of 64 blocks is exactly
1 250 000 rows (8 wobble periods), so tiles can be concatenated seamlessly to any length.
"""
from __future__ import annotations

import hashlib

import numpy as np

from .tbin import MODE_NRZI, TbinHeader

TILE_ROWS = 1_250_000
TILE_BLOCKS = 64
ROWS_PER_BIT = 1.0 / (800 * 50 * 1.28e-6)      # 19.53125


class Xoshiro256ss:
    """xoshiro256** (Blackman/Vigna), seeded with splitmix64 -- the generator SURVEY 8(d) names."""
    M = (1 << 64) - 1

    def __init__(self, seed: int):
        s = seed & self.M
        st = []
        for _ in range(4):
            s = (s + 0x9E3779B97F4A7C15) & self.M
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.M
            st.append(z ^ (z >> 31))
        self.s = st

    @staticmethod
    def _rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & Xoshiro256ss.M

    def next(self) -> int:
        s = self.s
        r = (self._rotl((s[1] * 5) & self.M, 7) * 9) & self.M
        t = (s[1] << 17) & self.M
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]
        s[2] ^= t
        s[3] = self._rotl(s[3], 45)
        return r

    def bytes(self, n: int) -> np.ndarray:
        out = np.empty((n + 7) // 8, dtype="<u8")
        for i in range(len(out)):
            out[i] = self.next()
        return out.view(np.uint8)[:n].copy()


def _parity9(w: np.ndarray) -> np.ndarray:
    p = np.zeros_like(w)
    for b in range(9):
        p ^= (w >> b) & 1
    return p


def nrzi_block_words(data: np.ndarray) -> np.ndarray:
    """9-bit characters (tracks 0..7 = data bits MSB..LSB in bits 8..1, parity track in bit 0) of one
    800 BPI block: data with odd parity, then 00 00 00 CRC 00 00 00 LRC (decode_nrzi.c:41-49)."""
    w = data.astype(np.uint16) << 1
    w |= (_parity9(w) ^ 1)                       # odd parity over all 9 bits
    crc = 0
    lrc = 0
    for x in w.tolist():                         # decode_nrzi.c:56-67
        lrc ^= x
        crc ^= x
        if crc & 2:
            crc ^= 0xF0
        lsb = crc & 1
        crc >>= 1
        if lsb:
            crc |= 0x100
    crc ^= 0x1AF
    lrc ^= crc
    tail = np.array([0, 0, 0, crc, 0, 0, 0, lrc], dtype=np.uint16)
    return np.concatenate([w, tail])


def nrzi_tile(seed: int = 0x9E3779B97F4A7C15, noise_seed: int = 0x1234ABCD, nblocks: int = TILE_BLOCKS,
              tile_rows: int | None = TILE_ROWS, data_bytes: int = 512, ntrks: int = 9, noise_mv: float = 5.0,
              wobble: float = 0.01) -> np.ndarray:
    """One super-tile: int16 array (tile_rows, 9).  Deterministic for given seeds."""
    rng = Xoshiro256ss(seed)
    period = (tile_rows / nblocks) if tile_rows else (data_bytes + 8) * ROWS_PER_BIT + 9375
    nrows = tile_rows if tile_rows else int(period * nblocks)
    maxvolts = 4.4
    volts = np.zeros((ntrks, nrows + 64), dtype=np.float64)
    hw = 0.35 * ROWS_PER_BIT
    k = np.arange(-8, 9)
    for b in range(nblocks):
        words = nrzi_block_words(rng.bytes(data_bytes))
        start = b * period + 4700.0                               # the block sits in the middle of its period
        # character times with the +-1 % 5 Hz speed wobble (5 Hz * 1.28 us/row)
        n = len(words)
        t = np.empty(n)
        pos = start
        for i in range(n):
            t[i] = pos
            pos += ROWS_PER_BIT * (1.0 + wobble * np.sin(2 * np.pi * 5.0 * pos * 1.28e-6))
        for trk in range(ntrks):
            bits = (words >> (ntrks - 1 - trk)) & 1
            times = t[bits == 1] + (trk % 4)                      # static head skew
            sign = np.where(np.arange(len(times)) % 2 == 0, -1.0, 1.0)   # first flux change reads negative
            amp = 3.0 + 0.1 * trk
            centre = np.floor(times).astype(np.int64)
            idx = centre[:, None] + k[None, :]
            dt = idx - times[:, None]
            pulse = np.where(np.abs(dt) < hw, 0.5 * (1.0 + np.cos(np.pi * dt / hw)), 0.0)
            np.add.at(volts[trk], idx.ravel(), (pulse * (amp * sign)[:, None]).ravel())
    g = np.random.Generator(np.random.PCG64(noise_seed))
    volts = volts[:, :nrows]
    volts += g.normal(0.0, noise_mv * 1e-3, size=volts.shape) - 0.015
    q = np.rint(volts / maxvolts * 32767.0)
    np.clip(q, -32767, 32767, out=q)                              # never the -32768 end marker
    return np.ascontiguousarray(q.T.astype("<i2"))


def nrzi_header(tstart_ns: int = 1_000_000_000) -> TbinHeader:
    return TbinHeader(descr="synthetic 9-track 800 BPI NRZI (readtape_b200.synth)", flags=0, ntrks=9, tdelta_ns=1280,
                      maxvolts=4.4, mode=MODE_NRZI, bpi=800.0, ips=50.0, tstart_ns=tstart_ns)


def nrzi_tape(nblocks: int = 8, seed: int = 7, **kw):
    """A small stand-alone capture (not tile-aligned): -> (header, rows[int16, (n, 9)])."""
    rows = nrzi_tile(seed=seed, noise_seed=seed ^ 0x5555, nblocks=nblocks, tile_rows=None, **kw)
    return nrzi_header(), rows


def tile_sha256(tile: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(tile).tobytes()).hexdigest()


# ---- GCR-density workload (BASELINE config 4) ----------------------------------------------------------------------
GCR_TDELTA_NS = 160                                  # 6.25 MHz
GCR_ROWS_PER_BIT = 1.0 / (9042 * 50 * 160e-9)        # 13.82


def gcr_like_tile(seed: int = 0xC0FFEE, nblocks: int = 4, bits_per_block: int = 45_000, gap_rows: int = 37_500,
                  ntrks: int = 9, noise_mv: float = 4.0) -> np.ndarray:
    """A tape tile at GCR 6250 density (9042 flux cells per inch, 50 IPS, 6.25 MHz sampling, maxvolts 3.2): every track carries
    an independent random run-length-limited bit stream (at most two 0 cells between 1 cells, as 5-bit GCR groups guarantee),
    one flux reversal per 1 cell, alternating polarity raised-cosine pulses, 0.3-inch gaps.  It has the transition density,
    amplitudes and block/gap proportions of a 6250 BPI reel, which is what the scan cost depends on; it is NOT a decodable
    GCR block (no sync marks / ECC), so it is a throughput workload only -- parity for GCR uses the reference's captures."""
    rng = np.random.Generator(np.random.PCG64(seed))
    period = int(bits_per_block * GCR_ROWS_PER_BIT) + gap_rows
    period = (period + 2047) // 2048 * 2048                       # whole ingest tiles per block period
    nrows = period * nblocks
    volts = np.zeros((ntrks, nrows + 64), dtype=np.float64)
    hw = 0.55 * GCR_ROWS_PER_BIT
    k = np.arange(-9, 10)
    for b in range(nblocks):
        start = b * period + gap_rows // 2
        for trk in range(ntrks):
            bits = rng.random(bits_per_block) < 0.62
            run = 0
            for i in range(bits_per_block):                        # enforce the (0,2) run-length limit
                if bits[i]:
                    run = 0
                else:
                    run += 1
                    if run > 2:
                        bits[i] = True; run = 0
            cells = np.flatnonzero(bits)
            times = start + (cells + 0.5) * GCR_ROWS_PER_BIT * (1.0 + 0.004 * np.sin(2 * np.pi * cells / 9000.0)) + 0.37 * trk
            sign = np.where(np.arange(len(times)) % 2 == 0, 1.0, -1.0)
            amp = 1.6 + 0.05 * trk
            centre = np.floor(times).astype(np.int64)
            idx = centre[:, None] + k[None, :]
            dt = idx - times[:, None]
            pulse = np.where(np.abs(dt) < hw, 0.5 * (1.0 + np.cos(np.pi * dt / hw)), 0.0)
            np.add.at(volts[trk], idx.ravel(), (pulse * (amp * sign)[:, None]).ravel())
    volts = volts[:, :nrows]
    volts += rng.normal(0.0, noise_mv * 1e-3, size=volts.shape)
    q = np.rint(volts / 3.2 * 32767.0)
    np.clip(q, -32767, 32767, out=q)
    return np.ascontiguousarray(q.T.astype("<i2"))


def gcr_header(tstart_ns: int = 1_000_000_000) -> TbinHeader:
    from .tbin import MODE_GCR
    return TbinHeader(descr="synthetic GCR-density flux pattern (readtape_b200.synth)", flags=0, ntrks=9, tdelta_ns=GCR_TDELTA_NS,
                      maxvolts=3.2, mode=MODE_GCR, bpi=9042.0, ips=50.0, tstart_ns=tstart_ns)


# ---- decodable 6250 BPI GCR blocks (BASELINE config 4) ---------------------------------------------------------------
# Block format: reference A_documentation.txt:300-328; what the reference checks: gcr_postprocess (decode_gcr.c:503-674): the
# 5-bit storage codes (gcr_datamap :428), odd parity of every 9-bit character (:472-480) and the ECC character of every group
# of 7 data bytes (gcr_compute_ecc :127-144).
GCR_MARK1, GCR_MARK2, GCR_SYNC = 0b00111, 0b11100, 0b11111
_GCR_DECODE = [16 + 10, 16 + 9, 16 + 2, 16 + 3, 16 + 5, 16 + 5, 16 + 6, 16 + 7, 16 + 10, 9, 10, 11, 16 + 13, 13, 14, 15,
               16 + 2, 16 + 5, 2, 3, 16 + 5, 5, 6, 7, 16 + 0, 0, 8, 1, 16 + 12, 4, 12, 16 + 15]          # gcr_datamap, decode_gcr.c:428
GCR_ENCODE = {v: code for code, v in enumerate(_GCR_DECODE) if v < 16}                                   # 4-bit value -> 5-bit code
_GCR_ECC_A = [0x0f6a71994c5230, 0x70110840108004, 0x5a701108401080, 0x372be95d5a7011,
              0xe95d5a70110840, 0x4c523001884412, 0x2be95d5a701108, 0x5d5a7011084010]                  # decode_gcr.c:128-136


def gcr_ecc(seven: bytes) -> int:
    """the ECC character of 7 data bytes: bit i = <dblock, A[i]> mod 2 over the 56-bit big-endian block (decode_gcr.c:137-144)"""
    d = int.from_bytes(bytes(seven), "big")
    return sum((bin(d & a).count("1") & 1) << i for i, a in enumerate(_GCR_ECC_A))


def _word9(b: int) -> int:
    """9-bit character (data bits msb..lsb, parity last) with odd parity"""
    return (b << 1) | ((bin(b).count("1") & 1) ^ 1)


def _subgroup_codes(words4) -> list:
    """the 5-bit storage codes of the 9 tracks for 4 characters (track t carries bit 8-t of every character, decode_gcr.c:454-470)"""
    out = []
    for t in range(9):
        nib = 0
        for w in words4:
            nib = (nib << 1) | ((w >> (8 - t)) & 1)
        out.append(GCR_ENCODE[nib])
    return out


def gcr_block_cells(data: bytes) -> np.ndarray:
    """flux cells of one block, bool (ncells, 9): preamble, Mark 1, data groups (7 bytes + ECC -> two 5-bit subgroups per track)
    with a resync burst after every 158 storage groups, End Mark, residual and CRC groups (no residual bytes), Mark 2, postamble.
    len(data) must be a multiple of 7."""
    assert len(data) % 7 == 0
    same = lambda code: [code] * 9
    sg = [same(0b10101), same(0b01111)] + [same(GCR_SYNC)] * 14 + [same(GCR_MARK1)]
    ngroups = 0
    for at in range(0, len(data), 7):
        seven = data[at:at + 7]
        words = [_word9(b) for b in seven] + [_word9(gcr_ecc(seven))]
        sg.append(_subgroup_codes(words[:4])); sg.append(_subgroup_codes(words[4:]))
        ngroups += 1
        if ngroups % 158 == 0 and at + 7 < len(data):
            sg += [same(GCR_MARK2), same(GCR_SYNC), same(GCR_SYNC), same(GCR_MARK1)]
    zero4 = _subgroup_codes([_word9(0)] * 4)
    sg += [same(GCR_SYNC), zero4, zero4, zero4, zero4, same(GCR_MARK2)]
    sg += [same(GCR_SYNC)] * 14 + [same(0b11110), same(0b10101)]
    codes = np.array(sg, dtype=np.uint8)                                   # (nsubgroups, 9)
    cells = ((codes[:, None, :] >> np.arange(4, -1, -1)[None, :, None]) & 1).astype(bool)
    return cells.reshape(-1, 9)


def gcr_tile(seed: int = 0xC0FFEE, nblocks: int = 8, data_bytes: int = 4095, gap_rows: int = 37_500, noise_mv: float = 4.0,
             amplitude: float = 1.9) -> np.ndarray:
    """A tape tile of decodable 6250 BPI GCR blocks (9042 flux cells per inch, 50 IPS, 6.25 MHz sampling, maxvolts 3.2): NRZI flux,
    i.e. the read signal of the reference's GCR captures (decoded with -zeros) -- a level that changes sign at every 1 cell,
    band-limited to about a cell -- so that zero crossings mark the 1 cells.  `data_bytes` random bytes per block (a multiple of 7),
    0.3-inch gaps; the tile is a whole number of ingest tiles long and can be repeated seamlessly."""
    rng = np.random.Generator(np.random.PCG64(seed))
    per_block = None
    blocks = []
    for _ in range(nblocks):
        cells = gcr_block_cells(rng.integers(0, 256, size=data_bytes, dtype=np.uint8).tobytes())
        blocks.append(cells)
        per_block = cells.shape[0]
    period = int(per_block * GCR_ROWS_PER_BIT) + gap_rows
    period = (period + 2047) // 2048 * 2048
    nrows = period * nblocks
    volts = np.zeros((9, nrows), dtype=np.float64)
    kern_half = int(GCR_ROWS_PER_BIT)                                      # raised-cosine smoothing over ~2 cells
    k = np.arange(-kern_half, kern_half + 1)
    kern = 0.5 * (1.0 + np.cos(np.pi * k / (kern_half + 1))); kern /= kern.sum()
    for b, cells in enumerate(blocks):
        start = b * period + gap_rows // 2
        n = cells.shape[0]
        edges = start + np.arange(n + 1) * GCR_ROWS_PER_BIT
        for trk in range(9):
            level = np.zeros(period, dtype=np.float64)
            ones = np.flatnonzero(cells[:, trk])
            t = edges[ones] + 0.5 * GCR_ROWS_PER_BIT + 0.37 * trk - b * period          # the flux reversal sits in the middle of its cell
            sign = 1.0
            # the level is 0 in the gap, rises with the first reversal and returns to 0 behind the last one
            pos = np.round(t).astype(np.int64)
            for i in range(len(pos)):
                hi = pos[i + 1] if i + 1 < len(pos) else min(period, pos[i] + int(1.5 * GCR_ROWS_PER_BIT))
                level[pos[i]:hi] = sign
                sign = -sign
            first = pos[0]
            level[max(0, first - int(1.5 * GCR_ROWS_PER_BIT)):first] = -1.0                # so that the first reversal is a crossing
            smooth = np.convolve(level, kern, mode="same") * (amplitude + 0.03 * trk)
            volts[trk, b * period:(b + 1) * period] += smooth
    volts += rng.normal(0.0, noise_mv * 1e-3, size=volts.shape)
    q = np.rint(volts / 3.2 * 32767.0)
    np.clip(q, -32767, 32767, out=q)
    return np.ascontiguousarray(q.T.astype("<i2"))


def csv_from_rows(rows: np.ndarray, maxvolts: float, tstart_ns: int = 0, tdelta_ns: int = 1280, decimals: int = 5,
                  title: str = "'synthetic capture") -> np.ndarray:
    """The text a logic analyser would export for these int16 rows (uint8 array): two title lines, then per sample
    "%12.8f, " of the time and "%9.5f, " per track of row/32767*maxvolts -- the layout the reference's own TBIN -> CSV
    direction writes (csvtbin.c:579-590), produced with integer arithmetic on whole arrays.  Used as INPUT of the CSV ingest
    (tests, bench); what matters is that it is realistic text, not that it round-trips."""
    rows = np.asarray(rows)
    n, nt = rows.shape
    head = (title + "\n" + "Time, " + ", ".join(f"Track {i}" for i in range(nt)) + "\n").encode()
    tw, vw = 12 + 2, 9 + 2                                   # field widths with the ", " separator
    line = np.full((n, tw + nt * vw + 1), ord(" "), dtype=np.uint8)
    line[:, -1] = ord("\n")
    # time: seconds with 8 decimals, right-aligned in 12
    t = (tstart_ns + tdelta_ns * np.arange(n, dtype=np.int64) + 5) // 10          # units of 1e-8 s
    for k in range(8):
        line[:, 11 - k] = ord("0") + (t % 10); t //= 10
    line[:, 3] = ord(".")
    for k in range(3):                                        # up to 999 s
        d = t % 10; t //= 10
        line[:, 2 - k] = np.where((d > 0) | (t > 0) | (k == 0), ord("0") + d, ord(" "))
    line[:, 12] = ord(",")
    scale = 10 ** decimals
    v = np.rint(rows.astype(np.float64) / 32767.0 * maxvolts * scale).astype(np.int64)
    neg = v < 0
    a = np.abs(v)
    base = tw
    for c in range(nt):
        col = a[:, c].copy()
        o = base + c * vw
        for k in range(decimals):
            line[:, o + 8 - k] = ord("0") + (col % 10); col //= 10
        line[:, o + 8 - decimals] = ord(".")
        line[:, o + 7 - decimals] = ord("0") + (col % 10); col //= 10
        line[:, o + 6 - decimals] = np.where(col > 0, ord("0") + (col % 10), np.where(neg[:, c], ord("-"), ord(" ")))
        line[:, o + 5 - decimals] = np.where((col > 0) & neg[:, c], ord("-"), ord(" "))
        line[:, o + 9] = ord(",")
    return np.concatenate([np.frombuffer(head, dtype=np.uint8), line.reshape(-1)])
