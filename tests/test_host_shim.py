"""End-to-end drop-in test: readtape's own host code with readblock() replaced by the event-driven
shim (readtape_b200/host/readblock_b200.c) must write byte-identical .tap/.bin files.

  * CPU (here): the shim linked against the oracle library -- checks the replay logic itself
    (oracle/_ref/readtape_shim_oracle, test infrastructure).
  * GPU: the shim linked against the CUDA library (readtape_b200/bin/readtape_b200): the product,
    with the speculative whole-tape scan and with the exact scan only (RT_NO_BULK=1).
Expected outputs are the reference's own (tests/golden/*.tap, and SHA-256 of its .bin files recorded in
the fixtures by oracle/make_golden.py).  Both binaries are built in the build container from the
reference sources where they lie; the tests skip if they were not built.
"""
import hashlib
import json
import os
import subprocess

import pytest

from conftest import ALL_FIXTURES, EXAMPLES, GOLDEN, ROOT

ORACLE_SHIM = os.path.join(ROOT, "oracle", "_ref", "readtape_shim_oracle")
CUDA_SHIM = os.path.join(ROOT, "readtape_b200", "bin", "readtape_b200")


def run_shim(exe, name, tmp_path, env_extra=None):
    doc = json.load(open(os.path.join(GOLDEN, name + ".segments.json")))
    capture = os.path.join(EXAMPLES, doc["capture"])
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (make -C readtape_b200/host in the build container)")
    if not os.path.exists(capture):
        pytest.skip(f"staged capture {capture} missing")
    out = os.path.join(str(tmp_path), name)
    env = dict(os.environ, RT_STATS="1")
    env.update(env_extra or {})
    opts = [o for o in doc["options"].split() if o not in ("-v", "-v3")] + ["-v"]
    r = subprocess.run([exe] + opts + [f"-outf={out}", capture], capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FATAL" not in r.stdout
    for fname, sha in doc["reference_outputs"].items():
        made = os.path.join(str(tmp_path), fname)
        assert os.path.exists(made), f"{fname} was not written"
        data = open(made, "rb").read()
        assert hashlib.sha256(data).hexdigest() == sha, f"{fname} differs from the reference's output"
        gold = os.path.join(GOLDEN, fname)
        if os.path.exists(gold):
            assert data == open(gold, "rb").read()
    return r.stdout


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_shim_on_oracle_backend_writes_reference_output(name, tmp_path):
    run_shim(ORACLE_SHIM, name, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_readtape_b200_writes_reference_output(name, tmp_path):
    out = run_shim(CUDA_SHIM, name, tmp_path)
    assert "cuda-sm100a" in out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Microdata_20blks.nm_tap", "LJS009_part1_39blks", "sf93_8blks"])
def test_readtape_b200_exact_scan_only(name, tmp_path):
    run_shim(CUDA_SHIM, name, tmp_path, {"RT_NO_BULK": "1"})


@pytest.mark.gpu
def test_readtape_b200_uses_the_speculative_scan(tmp_path):
    out = run_shim(CUDA_SHIM, "Microdata_20blks.nm_tap", tmp_path)
    line = [l for l in out.splitlines() if "speculative hits" in l]
    assert line, out[-1500:]
    hits = int(line[-1].split("events,")[1].split("speculative hits")[0])
    assert hits >= 15, line[-1]
