"""The C-ABI parity tests of tests/test_gpu_parity.py, run against the CPU simulation of the library where there is no GPU.

tests/test_gpu_parity.py calls the product through the Python mirror of the C-ABI (readtape_b200/abi.py) and compares bulk scans,
lookups, exact scans, resets, rewinds and fan-outs with the oracle and the committed reference digests.  Everything there that does
not look at a kernel as such (mask planes, the tile digest, the streamed upload, the fused ingest) is a statement about the ABI's
SEMANTICS -- and holds for tests/host_fast/hostsim.cu too, which builds rt_bulk_scan / rt_bulk_lookup from the host build of the
kernels' code and the product's lookup rules.  Those tests are run here with `cuda_lib` = that library: a change of the lookup rules
or of the scan templates that would break a GPU test shows up on the CPU first.  TEST INFRASTRUCTURE; the GPU run uses the product.
"""
import inspect
import itertools
import os
import subprocess

import pytest

from conftest import ROOT
from readtape_b200 import abi
import test_gpu_parity as gp

SIM_LIB = os.path.join(ROOT, "tests", "host_fast", "_build", "libhostsim.so")
# kernel-level tests: they inspect what only the CUDA library has (mask planes and thresholds, tile digest, streamed upload, fused ingest)
KERNEL_ONLY = {"test_data_driven_mask_thresholds_keep_the_scan_sparse", "test_streamed_host_scan_equals_plain_sequence", "test_tile_digest_at_scale_equals_oracle",
               "test_decodable_gcr_tape_cuda_equals_oracle_and_digest", "test_fused_ingest_masks_equal_the_separate_pass",
               "test_invert_is_served_by_the_fast_kernels_and_is_exact", "test_ingest_and_mask_kernels_side_by_side_give_the_same_planes"}


def _cases():
    out = []
    for name, fn in sorted(vars(gp).items()):
        if not name.startswith("test_") or not callable(fn) or name in KERNEL_ONLY: continue
        marks = getattr(fn, "pytestmark", [])            # the gpu mark is the module's (pytestmark = pytest.mark.gpu): every test there is a GPU test
        axes = []
        for m in marks:
            if m.name != "parametrize": continue
            names = [n.strip() for n in m.args[0].split(",")] if isinstance(m.args[0], str) else list(m.args[0])
            axes.append([dict(zip(names, v if len(names) > 1 else (v,))) for v in m.args[1]])
        for combo in itertools.product(*axes) if axes else [()]:
            kw = {}
            for d in combo: kw.update(d)
            out.append(pytest.param(name, kw, id=name[5:] + ("-" + "-".join(str(v) for v in kw.values()) if kw else "")))
    return out


@pytest.fixture(scope="session")
def sim_lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host_fast"), "sim"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return abi.load(SIM_LIB)


@pytest.mark.parametrize("name,kw", _cases())
def test_abi_semantics_on_the_cpu_simulation(name, kw, sim_lib, oracle_lib, tmp_path, request):
    fn = getattr(gp, name)
    args = dict(kw)
    for p in inspect.signature(fn).parameters:
        if p in args: continue
        if p == "cuda_lib": args[p] = sim_lib
        elif p == "oracle_lib": args[p] = oracle_lib
        elif p == "tmp_path": args[p] = tmp_path
        else: args[p] = request.getfixturevalue(p)
    fn(**args)
