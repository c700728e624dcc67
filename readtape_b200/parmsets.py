"""Built-in parameter sets, as data.

These are the numeric defaults of the reference's `parmcmds_*` tables (src/parmsets.c:78-118),
restated as plain tuples so that Python-side tests and the benchmark can build `rt_scan_cfg`
blocks without the C host.  The C host shim does not use this file: it passes readtape's own
`PARM` (struct parms_t, src/decoder.h:290) through the C-ABI.
Field order: clk_window, clk_alpha, agc_window, agc_alpha, min_peak, clk_factor, pulse_adj,
             pkww_bitfrac, pkww_rise, midbit, z1pt, z2pt
"""
from __future__ import annotations

from .tbin import MODE_GCR, MODE_NRZI, MODE_PE, MODE_WW

FIELDS = ("clk_window", "clk_alpha", "agc_window", "agc_alpha", "min_peak", "clk_factor",
          "pulse_adj", "pkww_bitfrac", "pkww_rise", "midbit", "z1pt", "z2pt")


def _p(**kw):
    d = dict.fromkeys(FIELDS, 0.0)
    d["clk_window"] = 0
    d["agc_window"] = 0
    d.update(kw)
    return d


PE = [
    _p(clk_window=0, clk_alpha=0.2, agc_window=5, min_peak=0.0, clk_factor=1.50, pulse_adj=0.4, pkww_bitfrac=0.7, pkww_rise=0.10),
    _p(clk_window=0, clk_alpha=0.2, agc_window=5, min_peak=0.1, clk_factor=1.50, pulse_adj=0.4, pkww_bitfrac=0.7, pkww_rise=0.10),
    _p(clk_window=3, clk_alpha=0.0, agc_window=5, min_peak=0.0, clk_factor=1.40, pulse_adj=0.0, pkww_bitfrac=0.7, pkww_rise=0.10),
    _p(clk_window=3, clk_alpha=0.0, agc_window=5, min_peak=0.0, clk_factor=1.40, pulse_adj=0.2, pkww_bitfrac=0.7, pkww_rise=0.10),
    _p(clk_window=5, clk_alpha=0.0, agc_window=5, min_peak=0.0, clk_factor=1.40, pulse_adj=0.0, pkww_bitfrac=0.7, pkww_rise=0.10),
    _p(clk_window=5, clk_alpha=0.0, agc_window=5, min_peak=0.0, clk_factor=1.50, pulse_adj=0.2, pkww_bitfrac=0.7, pkww_rise=0.10),
    _p(clk_window=5, clk_alpha=0.0, agc_window=5, min_peak=0.0, clk_factor=1.40, pulse_adj=0.4, pkww_bitfrac=0.7, pkww_rise=0.10),
    _p(clk_window=3, clk_alpha=0.0, agc_window=5, min_peak=0.0, clk_factor=1.40, pulse_adj=0.2, pkww_bitfrac=0.7, pkww_rise=0.10),
]

NRZI = [
    _p(clk_window=0, clk_alpha=0.2, agc_window=0, agc_alpha=0.3, min_peak=1.0, pulse_adj=0.3, pkww_bitfrac=0.7, pkww_rise=0.20, midbit=0.5),
    _p(clk_window=0, clk_alpha=0.3, agc_window=0, agc_alpha=0.3, min_peak=1.0, pulse_adj=0.4, pkww_bitfrac=0.6, pkww_rise=0.20, midbit=0.5),
    _p(clk_window=2, clk_alpha=0.0, agc_window=0, agc_alpha=0.3, min_peak=1.0, pulse_adj=0.4, pkww_bitfrac=0.7, pkww_rise=0.20, midbit=0.5),
    _p(clk_window=0, clk_alpha=0.6, agc_window=0, agc_alpha=0.3, min_peak=1.0, pulse_adj=0.4, pkww_bitfrac=0.6, pkww_rise=0.20, midbit=0.5),
    _p(clk_window=2, clk_alpha=0.0, agc_window=1, agc_alpha=0.0, min_peak=0.5, pulse_adj=0.5, pkww_bitfrac=0.9, pkww_rise=0.05, midbit=0.5),
    _p(clk_window=0, clk_alpha=0.2, agc_window=1, agc_alpha=0.0, min_peak=1.0, pulse_adj=0.5, pkww_bitfrac=0.7, pkww_rise=0.05, midbit=0.5),
    _p(clk_window=2, clk_alpha=0.0, agc_window=1, agc_alpha=0.0, min_peak=0.5, pulse_adj=0.5, pkww_bitfrac=0.7, pkww_rise=0.05, midbit=0.5),
    _p(clk_window=0, clk_alpha=0.6, agc_window=1, agc_alpha=0.0, min_peak=0.5, pulse_adj=0.5, pkww_bitfrac=0.6, pkww_rise=0.05, midbit=0.5),
]

GCR = [
    _p(clk_window=0, clk_alpha=0.015, agc_alpha=0.5, min_peak=0.2, pulse_adj=0.3, pkww_bitfrac=1.5, pkww_rise=0.20, z1pt=1.45, z2pt=2.35),
    _p(clk_window=0, clk_alpha=0.020, agc_alpha=0.5, min_peak=0.2, pulse_adj=0.3, pkww_bitfrac=1.5, pkww_rise=0.20, z1pt=1.45, z2pt=2.35),
    _p(clk_window=0, clk_alpha=0.010, agc_alpha=0.5, min_peak=0.2, pulse_adj=0.3, pkww_bitfrac=1.5, pkww_rise=0.20, z1pt=1.45, z2pt=2.35),
    _p(clk_window=10, clk_alpha=0.000, agc_alpha=0.5, min_peak=0.0, pulse_adj=0.6, pkww_bitfrac=1.5, pkww_rise=0.14, z1pt=1.40, z2pt=2.30),
    _p(clk_window=0, clk_alpha=0.020, agc_alpha=0.5, min_peak=0.2, pulse_adj=0.3, pkww_bitfrac=1.5, pkww_rise=0.20, z1pt=1.48, z2pt=2.35),
]

WW = [
    _p(clk_window=0, clk_alpha=0.05, agc_alpha=0.5, min_peak=1.00, pkww_bitfrac=0.4, pkww_rise=0.20),
    _p(clk_window=0, clk_alpha=0.02, agc_alpha=0.5, min_peak=0.05, pkww_bitfrac=0.2, pkww_rise=0.20),
]

BUILTIN = {MODE_PE: PE, MODE_NRZI: NRZI, MODE_GCR: GCR, MODE_WW: WW}
