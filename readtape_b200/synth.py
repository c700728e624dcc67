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
