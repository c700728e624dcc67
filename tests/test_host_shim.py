"""End-to-end drop-in test: readtape's own host code with readblock() replaced by the event-driven
shim (readtape_b200/host/readblock_b200.c) must write byte-identical .tap/.bin files.

  * CPU (here): the shim linked against the oracle library -- checks the replay logic itself
    (oracle/_ref/readtape_shim_oracle, test infrastructure).
  * GPU: the shim linked against the CUDA library (readtape_b200/bin/readtape_b200): the product,
    with the speculative whole-tape scan and with the exact scan only (RT_NO_BULK=1).
Expected outputs:
  * WHOLE captures: the reference-held goldens of examples/*/expected_results (every .tap/.bin, SHA-256 in
    tests/golden/full_outputs.json, verified byte for byte against the reference's files when that table was
    generated), plus the reference's own output for BASELINE config 1's "-nm -nrzi -tap" command line;
  * prefix captures (the per-segment event fixtures): the reference run on the prefix (tests/golden/*.tap and the
    SHA-256 of its .bin files recorded in the fixtures by oracle/make_golden.py).
Both binaries are built in the build container from the reference sources where they lie; the tests skip if they
were not built.
"""
import hashlib
import json
import os
import re
import subprocess

import pytest

from conftest import ALL_FIXTURES, GOLDEN, ROOT, capture_path

ORACLE_SHIM = os.path.join(ROOT, "oracle", "_ref", "readtape_shim_oracle")
CUDA_SHIM = os.path.join(ROOT, "readtape_b200", "bin", "readtape_b200")
FULL = json.load(open(os.path.join(GOLDEN, "full_outputs.json")))
MIN_HIT_RATE = 0.95          # block decodes served by the speculative whole-tape scan ((hits - restarts) / (hits + misses))


def shim_stats(stdout):
    """the RT_STATS=1 line of the shim -> dict"""
    m = re.search(r"B200 scan: (\d+) events, (\d+) speculative hits, (\d+) misses, (\d+) restarts, (\d+) exact spans", stdout)
    if not m:
        return None
    ev, hits, miss, restarts, spans = map(int, m.groups())
    return {"events": ev, "hits": hits, "misses": miss, "restarts": restarts, "exact_spans": spans}


def record_stats(label, st):
    """hit / miss / restart counts per fixture: printed (pytest -s / -rP) and kept in gpurun_out/hitrates.json"""
    print(f"[hit-rate] {label}: {st}")
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    path = os.path.join(out, "hitrates.json")
    try:
        doc = json.load(open(path))
    except Exception:
        doc = {}
    doc[label] = st
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=1, sort_keys=True)


def run_binary(exe, opts, capture, outbase, tmp_path, env_extra=None):
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (make -C readtape_b200/host in the build container)")
    env = dict(os.environ, RT_STATS="1")
    env.update(env_extra or {})
    opts = [o for o in opts.split() if o not in ("-v", "-v3")] + ["-v"]
    r = subprocess.run([exe] + opts + [f"-outf={outbase}", capture], capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FATAL" not in r.stdout
    return r.stdout


def run_shim(exe, name, tmp_path, env_extra=None):
    doc = json.load(open(os.path.join(GOLDEN, name + ".segments.json")))
    capture = capture_path(doc["capture"])
    out = os.path.join(str(tmp_path), name)
    stdout = run_binary(exe, doc["options"], capture, out, tmp_path, env_extra)
    for fname, sha in doc["reference_outputs"].items():
        made = os.path.join(str(tmp_path), fname)
        assert os.path.exists(made), f"{fname} was not written"
        data = open(made, "rb").read()
        assert hashlib.sha256(data).hexdigest() == sha, f"{fname} differs from the reference's output"
        gold = os.path.join(GOLDEN, fname)
        if os.path.exists(gold):
            assert data == open(gold, "rb").read()
    return stdout


def run_full(exe, label, tmp_path, env_extra=None):
    """the WHOLE capture with the command line of full_outputs.json[label]; every .tap/.bin must have the recorded SHA-256"""
    doc = FULL[label]
    capture = capture_path(doc["capture"], full=True)
    out = os.path.join(str(tmp_path), doc["capture"])
    stdout = run_binary(exe, doc["options"], capture, out, tmp_path, env_extra)
    assert doc["outputs"]
    for fname, want in doc["outputs"].items():
        made = os.path.join(str(tmp_path), fname)
        assert os.path.exists(made), f"{fname} was not written"
        data = open(made, "rb").read()
        assert len(data) == want["bytes"] and hashlib.sha256(data).hexdigest() == want["sha256"], \
            f"{fname} differs from the reference{'-held golden' if want['reference_held_golden'] else ''} ({len(data)} vs {want['bytes']} bytes)"
    return stdout


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_shim_on_oracle_backend_writes_reference_output(name, tmp_path):
    run_shim(ORACLE_SHIM, name, tmp_path)


@pytest.mark.parametrize("label", sorted(FULL))
def test_shim_on_oracle_backend_whole_capture_matches_reference_golden(label, tmp_path):
    """the replay logic on the WHOLE captures (later blocks, later AGC states, the real end of tape), CPU oracle behind the C-ABI"""
    run_full(ORACLE_SHIM, label, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_readtape_b200_writes_reference_output(name, tmp_path):
    out = run_shim(CUDA_SHIM, name, tmp_path)
    assert "cuda-sm100a" in out


@pytest.mark.gpu
@pytest.mark.parametrize("label", sorted(FULL))
def test_readtape_b200_whole_capture_matches_reference_golden(label, tmp_path):
    """the product on the WHOLE bundled captures against the reference-held goldens (all 12 .tap/.bin of
    examples/*/expected_results) and BASELINE config 1's command line; also records how many block decodes the
    speculative whole-tape scan served"""
    out = run_full(CUDA_SHIM, label, tmp_path)
    assert "cuda-sm100a" in out
    st = shim_stats(out)
    assert st is not None, out[-1500:]
    record_stats(label, st)
    doc = FULL[label]
    if "-whirlwind" in doc["options"]:
        return                               # Whirlwind: the detector state persists across blocks, no speculative units
    if "-differentiate" in doc["options"]:
        return                               # 3 block decodes on the exact generic kernel; its unit finder is a heuristic for this detector
    decodes = st["hits"] + st["misses"]          # a restart is a hit whose unit ended before the block did
    assert decodes > 0 and st["hits"] - st["restarts"] >= MIN_HIT_RATE * decodes, f"{label}: {st}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Microdata_20blks.nm_tap", "LJS009_part1_39blks", "sf93_8blks"])
def test_readtape_b200_exact_scan_only(name, tmp_path):
    run_shim(CUDA_SHIM, name, tmp_path, {"RT_NO_BULK": "1"})


@pytest.mark.gpu
def test_readtape_b200_uses_the_speculative_scan(tmp_path):
    out = run_shim(CUDA_SHIM, "Microdata_20blks.nm_tap", tmp_path)
    st = shim_stats(out)
    assert st is not None and st["hits"] >= 15, out[-1500:]


@pytest.mark.gpu
def test_readtape_b200_product_on_64_super_tiles_equals_reference(tmp_path):
    """whole program, file -> .tap, on a synthetic reel of 64 super-tiles (80 M rows, 1.44 GB, 4096 blocks): readtape with the B200
    scan against the unmodified reference binary; the two .tap files must be identical.  Exercises what only a long tape has:
    32-bit offsets inside units, event-pool regrowth, thousands of speculative lookups in sequence."""
    import shutil
    import tempfile
    import time
    import numpy as np
    from readtape_b200 import synth, tbin
    ref = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
    if not (os.path.exists(ref) and os.path.exists(CUDA_SHIM)):
        pytest.skip("reference / product binaries not built")
    reps = 64
    d = tempfile.mkdtemp(prefix="rt64_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        tile = synth.nrzi_tile()
        path = os.path.join(d, "reel.tbin")
        with open(path, "wb") as fh:
            fh.write(tbin.build_header(synth.nrzi_header()))
            for _ in range(reps):
                tile.tofile(fh)
            fh.write(np.array([tbin.END_MARK], dtype="<i2").tobytes())
        opts = ["-q", "-nm", "-nrzi", "-bpi=800", "-ips=50", "-tap", "-nolog", "-nolabels"]
        pr = subprocess.Popen([ref] + opts + [f"-outf={d}/ref", path], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        t0 = time.time()
        r = subprocess.run([CUDA_SHIM] + opts + [f"-outf={d}/new", path], capture_output=True, text=True, env=dict(os.environ, RT_STATS="1"), timeout=900)
        dt = time.time() - t0
        ref_out, _ = pr.communicate(timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert pr.returncode == 0, ref_out[-2000:]
        st = shim_stats(r.stdout)
        record_stats("synthetic_64_super_tiles", dict(st or {}, seconds=round(dt, 2), track_samples_per_s=reps * tile.shape[0] * 9 / dt))
        a = open(f"{d}/new.tap", "rb").read(); b = open(f"{d}/ref.tap", "rb").read()
        assert len(b) > reps * 64 * 512 and a == b, f".tap differs ({len(a)} vs {len(b)} bytes)"
        assert st["misses"] == 0 and st["restarts"] == 0 and st["hits"] >= reps * 64, st
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _reel(d, reps):
    import numpy as np
    from readtape_b200 import synth, tbin
    tile = synth.nrzi_tile()
    path = os.path.join(d, "reel.tbin")
    with open(path, "wb") as fh:
        fh.write(tbin.build_header(synth.nrzi_header()))
        for _ in range(reps):
            tile.tofile(fh)
        fh.write(np.array([tbin.END_MARK], dtype="<i2").tobytes())
    return path, reps * tile.shape[0]


@pytest.mark.gpu
def test_readtape_b200_split_between_worker_processes_equals_reference(tmp_path):
    """RT_WORKERS: one reel split between worker processes at inter-block gaps (each worker scans its share on the GPU and replays it
    through the reference's handlers; the hand-over between neighbours is PROVEN by rt_bulk_lookup or the reel is decoded unsplit).
    The concatenated .tap must be byte-identical to the unmodified reference's: 40 super-tiles in 5 workers, and real captures
    (NRZI 9-track, NRZI 7-track, PE, GCR) split in 3 with small shares forced."""
    import shutil
    import tempfile
    import time
    ref = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
    if not (os.path.exists(ref) and os.path.exists(CUDA_SHIM)):
        pytest.skip("reference / product binaries not built")
    d = tempfile.mkdtemp(prefix="rtw_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        path, nrows = _reel(d, 40)
        jobs = [("synthetic_40_tiles_5_workers", path, "-q -nm -nrzi -bpi=800 -ips=50 -tap -nolog -nolabels", {"RT_WORKERS": "5"}, 5)]
        for name, opts in (("PLAGO_beginning", "-q -nm -nrzi -ips=50 -tap -nolog"), ("tss_4secs", "-q -m -nrzi -ntrks=7 -tap -nolog"),
                           ("LJS009_part1_39blks", "-q -m -ntrks=9 -pe -bpi=1600 -ips=50 -tap -nolog"),
                           ("1kblks_43blks", "-q -m -gcr -ips=50 -order=76543210p -zeros -correct -tap -nolog")):
            jobs.append((name + "_3_workers", capture_path(name, full=True), opts,
                         {"RT_WORKERS": "3", "RT_WORKER_MIN_ROWS": "300000", "RT_WORKER_MARGIN_ROWS": "700000"}, 3))
        for label, cap, opts, env, nw in jobs:
            r0 = subprocess.run([ref] + opts.split() + [f"-outf={d}/ref_{label}", cap], capture_output=True, text=True, timeout=900)
            assert r0.returncode == 0, r0.stdout[-1500:]
            t0 = time.time()
            r = subprocess.run([CUDA_SHIM] + opts.split() + [f"-outf={d}/new_{label}", cap], capture_output=True, text=True,
                               env=dict(os.environ, RT_STATS="1", **env), timeout=900)
            dt = time.time() - t0
            assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-1500:]
            a = open(f"{d}/new_{label}.tap", "rb").read(); b = open(f"{d}/ref_{label}.tap", "rb").read()
            workers = len([l for l in r.stdout.splitlines() if "B200 scan: worker" in l and "speculative hits" in l]) + 1     # the last one reports as usual
            unsplit = "decoding the reel unsplit" in r.stdout
            record_stats(label, {"seconds": round(dt, 2), "tap_bytes": len(a), "workers_reported": workers, "unsplit_fallback": unsplit})
            assert a == b, f"{label}: .tap differs ({len(a)} vs {len(b)} bytes)\n" + r.stdout[-1500:]
            assert r.stdout.strip().splitlines()[-1] == r0.stdout.strip().splitlines()[-1], (r.stdout[-300:], r0.stdout[-300:])
            assert not [f for f in os.listdir(d) if ".part" in f], "part files left behind"
            if label.startswith("synthetic"):
                assert workers == nw and not unsplit, r.stdout[-1500:]
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _gcr_reel(d, nblocks=6):
    from readtape_b200 import synth, tbin
    path = os.path.join(d, "gcr.tbin")
    tbin.write_tbin(path, synth.gcr_header(), synth.gcr_tile(nblocks=nblocks))
    return path


def test_reference_decodes_the_synthetic_gcr_blocks(tmp_path):
    """BASELINE config 4 asks for synthetic GCR blocks the reference's gcr_postprocess validates: the unmodified reference must decode
    synth.gcr_tile() without errors or warnings and return exactly the generated bytes"""
    import numpy as np
    ref = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
    if not os.path.exists(ref):
        pytest.skip("reference binary not built")
    path = _gcr_reel(str(tmp_path))
    r = subprocess.run([ref, "-v", "-nm", "-gcr", "-ips=50", "-zeros", "-tap", "-nolabels", f"-outf={tmp_path}/ref", path], capture_output=True, text=True)
    assert r.returncode == 0 and "0 blocks had errors, 0 had warnings" in r.stdout, r.stdout[-1500:]
    tap = open(f"{tmp_path}/ref.tap", "rb").read()
    rng = np.random.Generator(np.random.PCG64(0xC0FFEE))
    off = 0
    for _ in range(6):
        want = rng.integers(0, 256, size=4095, dtype=np.uint8).tobytes()
        n = int.from_bytes(tap[off:off + 4], "little")
        assert n == 4095 and tap[off + 4:off + 4 + n] == want
        off += 4 + n + (n & 1) + 4
    assert tap[off:] == b"\xff" * 4


def test_shim_on_oracle_backend_decodes_the_synthetic_gcr_blocks(tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
    if not (os.path.exists(ref) and os.path.exists(ORACLE_SHIM)):
        pytest.skip("binaries not built")
    path = _gcr_reel(str(tmp_path), nblocks=3)
    for exe, tag in ((ref, "ref"), (ORACLE_SHIM, "new")):
        r = subprocess.run([exe, "-q", "-m", "-gcr", "-ips=50", "-zeros", "-tap", "-nolabels", "-nolog", f"-outf={tmp_path}/{tag}", path], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1500:]
    assert open(f"{tmp_path}/new.tap", "rb").read() == open(f"{tmp_path}/ref.tap", "rb").read()


@pytest.mark.gpu
def test_readtape_b200_decodes_the_synthetic_gcr_blocks(tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
    if not (os.path.exists(ref) and os.path.exists(CUDA_SHIM)):
        pytest.skip("binaries not built")
    path = _gcr_reel(str(tmp_path), nblocks=8)
    for exe, tag in ((ref, "ref"), (CUDA_SHIM, "new")):
        r = subprocess.run([exe, "-q", "-m", "-gcr", "-ips=50", "-zeros", "-tap", "-nolabels", "-nolog", f"-outf={tmp_path}/{tag}", path], capture_output=True, text=True,
                           env=dict(os.environ, RT_STATS="1"))
        assert r.returncode == 0, r.stdout[-1500:]
        if tag == "new":
            st = shim_stats(r.stdout)
            record_stats("synthetic_gcr_8_blocks", st)
            assert st["misses"] == 0 and st["hits"] >= 8, st
    assert open(f"{tmp_path}/new.tap", "rb").read() == open(f"{tmp_path}/ref.tap", "rb").read()


# ---- -subsample=n (readtape.c:1407-1413): of every n rows the last is used, the sample period is NOT rescaled -------------
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
SUBSAMPLE_CASES = ["-nrzi -bpi=1600 -ips=50 -subsample=2 -tap", "-nrzi -bpi=2400 -ips=50 -subsample=3 -tap -nm",
                   # -invert (readtape.c:1421): served by the fast kernels on a negated copy of the planes
                   "-nrzi -bpi=800 -ips=50 -invert -tap"]


def _subsample_case(exe, opts, tmp_path):
    """the unmodified reference runs beside the shim (it travels with the repo): same .tap, same log"""
    if not os.path.exists(REF_EXE):
        pytest.skip("oracle/_ref/readtape_ref not built")
    capture = capture_path("Microdata_20blks")
    got = {}
    for tag, binary in (("ref", REF_EXE), ("new", exe)):
        d = tmp_path / tag
        d.mkdir()
        if not os.path.exists(binary):
            pytest.skip(f"{binary} not built")
        env = {k: v for k, v in os.environ.items() if k != "RT_STATS"}
        r = subprocess.run([binary] + opts.split() + ["-outf=m", capture], capture_output=True, text=True, cwd=str(d), env=env, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        log = [ln for ln in open(d / "m.log", errors="replace").read().splitlines()
               if not re.search(r"^this is readtape version|command line:|samples were processed in", ln)]
        got[tag] = (open(d / "m.tap", "rb").read(), log)
    assert len(got["ref"][0]) > 5000
    assert got["new"][0] == got["ref"][0], ".tap differs from the reference's"
    assert got["new"][1] == got["ref"][1], "log differs from the reference's"


@pytest.mark.parametrize("opts", SUBSAMPLE_CASES)
def test_shim_on_oracle_backend_subsample(opts, tmp_path):
    _subsample_case(ORACLE_SHIM, opts, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("opts", SUBSAMPLE_CASES)
def test_readtape_b200_subsample(opts, tmp_path):
    _subsample_case(CUDA_SHIM, opts, tmp_path)
