"""The drop-in must not make readtape's .log / .peakstats.csv drift: the shim prints the same execution-time
configuration block as the reference (readtape.c:1460-1499) and replays every peak exactly once, so the log and
the peak statistics of a run are line for line the reference's, apart from the lines that carry the program name,
the wall clock and the elapsed time.

The unmodified reference binary (oracle/_ref/readtape_ref, test infrastructure) runs beside the shim on the same
capture; CPU: shim on the oracle backend, GPU: the product binary.
"""
import json
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT, capture_path

REF = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
ORACLE_SHIM = os.path.join(ROOT, "oracle", "_ref", "readtape_shim_oracle")
CUDA_SHIM = os.path.join(ROOT, "readtape_b200", "bin", "readtape_b200")
NAMES = ["Microdata_20blks", "PLAGO_beginning", "1600bpi_ukn_6s", "sf93_8blks", "analog", "tss_4secs", "132_pt1"]
VOLATILE = re.compile(r"^this is readtape version|^  command line:|samples were processed in \d+ seconds")


def run(exe, name, outdir):
    doc = json.load(open(os.path.join(GOLDEN, name + ".segments.json")))
    capture = capture_path(doc["capture"])
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built")
    os.makedirs(outdir, exist_ok=True)
    env = {k: v for k, v in os.environ.items() if k != "RT_STATS"}
    r = subprocess.run([exe] + doc["options"].split() + [f"-outf={outdir}/{name}", capture], capture_output=True, text=True, cwd=outdir, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = {}
    for f in sorted(os.listdir(outdir)):
        if f.endswith(".log") or f.endswith(".csv"):
            lines = open(os.path.join(outdir, f), errors="replace").read().replace(outdir + "/", "").splitlines()
            out[f] = [ln for ln in lines if not VOLATILE.search(ln)]
    return out


def compare(exe, name, tmp_path):
    want = run(REF, name, os.path.join(str(tmp_path), "ref"))
    got = run(exe, name, os.path.join(str(tmp_path), "new"))
    assert sorted(want) == sorted(got) and any(f.endswith(".log") for f in want)
    for f in want:
        if want[f] != got[f]:
            k = next((i for i, (a, b) in enumerate(zip(want[f], got[f])) if a != b), min(len(want[f]), len(got[f])))
            pytest.fail(f"{f} differs from the reference's at line {k + 1}:\n  ref: {want[f][k] if k < len(want[f]) else None}\n  new: {got[f][k] if k < len(got[f]) else None}")


@pytest.mark.parametrize("name", NAMES)
def test_shim_on_oracle_backend_keeps_log_and_peakstats(name, tmp_path):
    compare(ORACLE_SHIM, name, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_readtape_b200_keeps_log_and_peakstats(name, tmp_path):
    compare(CUDA_SHIM, name, tmp_path)
