"""The C-ABI boundary: both libraries export every symbol include/*.h declare."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from readtape_b200 import abi


def declared_symbols():
    text = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include"))) if h.endswith(".h"))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(abi.EXPORTS)


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_library_exports_every_declared_symbol(which, oracle_lib):
    path = abi.PRODUCT_LIB if which == "product" else abi.ORACLE_LIB
    assert os.path.exists(path), f"{path} not built"
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{path} does not export {name}"
    lib.rt_backend.restype = ctypes.c_char_p
    assert lib.rt_backend() == (b"cuda-sm100a" if which == "product" else b"oracle-cpu")
    assert lib.rt_abi_version() == 1


def test_pod_sizes_match_header():
    assert ctypes.sizeof(abi.TapeDesc) == 4 + 4 + 19 * 4 + 4 + 8 + 8 or ctypes.sizeof(abi.TapeDesc) == 104
    assert ctypes.sizeof(abi.Parms) == 44
    assert ctypes.sizeof(abi.ScanCfg) == 16 + 44 + 19 * 4
    assert abi.EVENT_DTYPE.itemsize == 32


def test_product_has_no_cpu_fallback(cuda_lib):
    """Without a GPU the product library must refuse to open a tape (and say why)."""
    from conftest import have_gpu
    if have_gpu():
        pytest.skip("a GPU is present")
    d = abi.make_desc(9, 4.4, 1280, 0)
    with pytest.raises(abi.RtError) as ei:
        cuda_lib.open(d)
    assert ei.value.code == -7 and "no CPU fallback" in str(ei.value)


def test_width_helper_matches_reference_values(oracle_lib, cuda_lib):
    """pkww_width = min(50, (int)(bitfrac/(bpi*ips*deltat))), readtape.c:1453-1457; SURVEY 8 shapes."""
    from readtape_b200 import parmsets, tbin
    cases = [(tbin.MODE_NRZI, parmsets.NRZI[0], 800, 50, 1280, 13), (tbin.MODE_PE, parmsets.PE[0], 1600, 50, 1280, 6),
             (tbin.MODE_PE, parmsets.PE[0], 1600, 50, 640, 13), (tbin.MODE_WW, parmsets.WW[0], 100, 50, 3840, 20),
             (tbin.MODE_NRZI, parmsets.NRZI[4], 800, 50, 1280, 17), (tbin.MODE_GCR, parmsets.GCR[0], 9042, 50, 160, 20)]
    for mode, p, bpi, ips, dt, want in cases:
        cfg = abi.make_cfg(mode, p, bpi, ips)
        for lib in (oracle_lib, cuda_lib):
            assert lib.L.rt_pkww_width(ctypes.byref(cfg), dt) == want, (mode, bpi, dt)
