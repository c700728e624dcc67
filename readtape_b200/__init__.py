"""readtape_b200 -- Blackwell-native per-track analog scan for readtape.

The product is the C-ABI library `readtape_b200/lib/librt_scan_b200.so` (hand-written CUDA for
sm_100a, built from `readtape_b200/csrc/`) plus the C host shim in `readtape_b200/host/`.
This Python package is only the test/benchmark harness around it: a ctypes binding (`abi`),
the TBIN container (`tbin`), the built-in parameter sets as data (`parmsets`) and a synthetic
tape generator (`synth`).
"""
__version__ = "0.1.0"
