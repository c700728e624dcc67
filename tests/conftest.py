"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here without a GPU: the oracle against the committed reference goldens, the
host-side logic, and the C-ABI export check.  `-m gpu` runs on a B200 and compares the CUDA
library with the reference goldens and with the CPU oracle through the same C-ABI.
Nothing under tests/ reads /root/reference: reference-derived inputs are the committed fixtures
in tests/golden/ and the staged captures (oracle/captures.py: xz files under oracle/_ref/examples_full/
that travel to the GPU box; the .tbin files the tests read are derived from them on first use).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EXAMPLES = os.path.join(ROOT, "oracle", "_ref", "examples")

ALL_FIXTURES = ["Microdata_20blks", "Microdata_20blks.nm_tap", "PLAGO_beginning", "PLAGO_beginning.nm_tap",
                "1600bpi_ukn_6s", "LJS009_part1_39blks", "1kblks_43blks", "sf93_8blks", "analog",
                "SRI_SDS_102715028_4secs", "tss_4secs", "132_pt1"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def capture_path(capture_file, full=False):
    """path of a staged capture (`<name>.tbin`): the prefix the segment fixtures were made from, or the whole file"""
    from oracle import captures
    name = capture_file[:-5] if capture_file.endswith(".tbin") else capture_file
    path = captures.full_path(name) if full else captures.staged_path(name)
    if path is None:
        pytest.skip(f"capture {name} was never staged (run `python oracle/make_golden.py --stage-only` in the build container)")
    return path


def load_capture(fixture_name):
    """-> (doc, segments, heads, rows) for a committed fixture; skips if the staged capture is absent."""
    from readtape_b200 import evlog, tbin
    doc, segs = evlog.load_fixture(os.path.join(GOLDEN, fixture_name + ".segments.json"))
    path = capture_path(doc["capture"])
    heads = doc["heads"]
    _, rows = tbin.read_tbin(path, nheads=heads["nheads"])
    return doc, segs, heads, np.asarray(rows)


@pytest.fixture(scope="session")
def oracle_lib():
    from readtape_b200 import abi
    if not os.path.exists(abi.ORACLE_LIB):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return abi.load_oracle()


@pytest.fixture(scope="session")
def cuda_lib():
    from readtape_b200 import abi
    return abi.load_product()      # raises if the CUDA library was not built: no fallback
