"""The host shim's SPECULATIVE paths without a GPU: readblock_b200.c on the CPU simulation of the whole C-ABI.

The CPU oracle implements only the exact scan, so `readtape_shim_oracle` never takes the paths the product lives on: speculative
hits, units that end inside a block (the exact scan continues the replay), bridge scans, chained units, the parameter-set fan-out,
the hand-over proofs between worker processes.  tests/host_fast/hostsim.cu implements rt_bulk_scan / rt_bulk_lookup on the CPU from
the HOST BUILD of the product's scan code (the same __host__ __device__ templates the kernels instantiate, the same proof data), the
product's own lookup rules (readtape_b200/csrc/lookup_rules.h) and a re-statement of the unit finder; `readtape_shim_hostsim` is
the reference's host code + readblock_b200.c linked against it.  Everything below runs that binary beside the unmodified reference:

  * the WHOLE bundled captures against the reference-held goldens, with the hit / miss / restart counts the GPU run recorded
    (profiles/hitrates_r02.json: the simulation reproduces them number for number);
  * units cut every n rows, INSIDE blocks (HOSTSIM_UNIT_ROWS): outputs must not depend on the unit finder;
  * one reel split between worker processes (RT_WORKERS), the parameter-set fan-out (RT_FANOUT=1);
  * the randomised runs of tests/test_fuzz_shim.py (synthetic tapes, capture windows, random options), with random unit cuts.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, capture_path
import test_fuzz_shim as fz
import test_host_shim as hs

SIM = os.path.join(ROOT, "oracle", "_ref", "readtape_shim_hostsim")
REF = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")


@pytest.fixture(scope="session")
def sim():
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "readtape_b200", "host"), "hostsim"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if not (os.path.exists(SIM) and os.path.exists(REF)):
        pytest.skip("readtape_shim_hostsim / readtape_ref not built (make -C readtape_b200/host hostsim in the build container)")
    return SIM


@pytest.mark.parametrize("label", sorted(hs.FULL))
def test_hostsim_whole_capture_matches_reference_golden(label, sim, tmp_path):
    out = hs.run_full(sim, label, tmp_path)
    assert "hostsim-cpu" in out
    st = hs.shim_stats(out)
    assert st is not None, out[-1500:]
    print(f"[hostsim] {label}: {st}")
    doc = hs.FULL[label]
    if "-whirlwind" in doc["options"] or "-differentiate" in doc["options"]:
        return
    decodes = st["hits"] + st["misses"]
    assert decodes > 0 and st["hits"] - st["restarts"] >= hs.MIN_HIT_RATE * decodes, f"{label}: {st}"


@pytest.mark.parametrize("name,cut", [("Microdata_20blks.nm_tap", 4096), ("Microdata_20blks", 20000), ("LJS009_part1_39blks", 9000), ("sf93_8blks", 30000),
                                      ("1600bpi_ukn_6s", 2048), ("SRI_SDS_102715028_4secs", 5000)])
def test_hostsim_units_cut_inside_blocks(name, cut, sim, tmp_path):
    """the unit finder is a heuristic: with units cut every `cut` rows (in the middle of blocks) the proof has to reject what it cannot
    prove and the exact scan has to carry the replay on -- same outputs, many more misses / restarts"""
    out = hs.run_shim(sim, name, tmp_path, {"HOSTSIM_UNIT_ROWS": str(cut)})
    st = hs.shim_stats(out)
    print(f"[hostsim] {name}, units of {cut} rows: {st}")
    assert st["misses"] + st["restarts"] > 0, st


def test_hostsim_parameter_set_fanout(sim, tmp_path):
    """all active parameter sets scanned by ONE rt_bulk_scan call from the start (RT_FANOUT=1), BASELINE config 3"""
    out = hs.run_shim(sim, "LJS009_part1_39blks", tmp_path, {"RT_FANOUT": "1"})
    assert hs.shim_stats(out)["hits"] > 0


def test_hostsim_worker_processes(sim, tmp_path):
    """RT_WORKERS: one reel split between worker processes, the hand-overs proven by rt_bulk_lookup / rt_bulk_last_unit"""
    d = str(tmp_path)
    path, nrows = hs._reel(d, 6)
    jobs = [("synthetic_6_tiles_3_workers", path, "-q -nm -nrzi -bpi=800 -ips=50 -tap -nolog -nolabels", {"RT_WORKERS": "3", "RT_WORKER_MIN_ROWS": "100000", "RT_WORKER_MARGIN_ROWS": "100000"}, 3)]
    for name, opts in (("LJS009_part1_39blks", "-q -m -ntrks=9 -pe -bpi=1600 -ips=50 -tap -nolog"),
                       ("1kblks_43blks", "-q -m -gcr -ips=50 -order=76543210p -zeros -correct -tap -nolog"),
                       ("tss_4secs", "-q -m -nrzi -ntrks=7 -tap -nolog")):
        jobs.append((name + "_3_workers", capture_path(name, full=True), opts, {"RT_WORKERS": "3", "RT_WORKER_MIN_ROWS": "300000", "RT_WORKER_MARGIN_ROWS": "700000"}, 3))
    for label, cap, opts, env, nw in jobs:
        r0 = subprocess.run([REF] + opts.split() + [f"-outf={d}/ref_{label}", cap], capture_output=True, text=True, timeout=900)
        assert r0.returncode == 0, r0.stdout[-1500:]
        r = subprocess.run([sim] + opts.split() + [f"-outf={d}/new_{label}", cap], capture_output=True, text=True, env=dict(os.environ, RT_STATS="1", **env), timeout=900)
        assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-1500:]
        a = open(f"{d}/new_{label}.tap", "rb").read(); b = open(f"{d}/ref_{label}.tap", "rb").read()
        workers = len([l for l in r.stdout.splitlines() if "B200 scan: worker" in l and "speculative hits" in l]) + 1
        unsplit = "decoding the reel unsplit" in r.stdout
        print(f"[hostsim] {label}: {workers} workers reported, unsplit fallback {unsplit}")
        assert a == b, f"{label}: .tap differs ({len(a)} vs {len(b)} bytes)\n" + r.stdout[-1500:]
        assert r.stdout.strip().splitlines()[-1] == r0.stdout.strip().splitlines()[-1], (r.stdout[-300:], r0.stdout[-300:])
        assert not [f for f in os.listdir(d) if ".part" in f], "part files left behind"
        if label.startswith("synthetic"):
            assert workers == nw and not unsplit, r.stdout[-1500:]


SIM_KNOBS = ("HOSTSIM_UNIT_ROWS", "RT_FANOUT", "RT_BRIDGE")


def _with_random_cuts(make_case):
    """the case generator of test_fuzz_shim + the knobs of the speculative paths: units cut at random, all parameter sets scanned
    at once or one by one, bridge scans off"""
    def f(rng, wd):
        opts, what = make_case(rng, wd)
        for k in SIM_KNOBS: os.environ.pop(k, None)
        if rng.random() < 0.5: os.environ["HOSTSIM_UNIT_ROWS"] = str(int(rng.choice([1024, 4096, 20000, 100000])))
        r = rng.random()
        if r < 0.25: os.environ["RT_FANOUT"] = "1"
        elif r < 0.35: os.environ["RT_FANOUT"] = "0"
        if rng.random() < 0.2: os.environ["RT_BRIDGE"] = "0"
        return opts, (what or "") + ", " + " ".join(f"{k}={os.environ[k]}" for k in SIM_KNOBS if k in os.environ)
    return f


def _punch_all_track_dropouts(rng, wd):
    """all tracks attenuated for 40..600 rows at 1..4 places of t.tbin: the REAL unit finder then cuts units inside blocks (a quiet
    stretch of all tracks is what it looks for), whatever the reference makes of the damaged block"""
    path = os.path.join(wd, "t.tbin")
    raw = np.fromfile(path, dtype="<i2", offset=256)
    rows = raw[:-1].reshape(-1, 9).copy()
    for _ in range(int(rng.integers(1, 5))):
        a = int(rng.integers(0, max(1, len(rows) - 700))); b = a + int(rng.integers(40, 600))
        rows[a:b] = (rows[a:b] * float(rng.choice([0.0, 0.01, 0.03]))).astype(np.int16)
    head = open(path, "rb").read(256)
    with open(path, "wb") as fh:
        fh.write(head); rows.tofile(fh); fh.write(raw[-1:].tobytes())


def dropout_case(rng, wd):
    opts, what = fz.synthetic_case(rng, wd)
    _punch_all_track_dropouts(rng, wd)
    return opts, what + ", all-track drop-outs"


def test_hostsim_random_tapes_with_all_track_dropouts(sim, tmp_path):
    try: fz.fuzz(sim, _with_random_cuts(dropout_case), 304, 10, tmp_path)
    finally: [os.environ.pop(k, None) for k in SIM_KNOBS]


def test_hostsim_random_synthetic_tapes(sim, tmp_path):
    try: fz.fuzz(sim, _with_random_cuts(fz.synthetic_case), 301, 12, tmp_path)
    finally: [os.environ.pop(k, None) for k in SIM_KNOBS]


def test_hostsim_random_capture_windows(sim, tmp_path):
    try: fz.fuzz(sim, _with_random_cuts(fz.capture_case), 302, 10, tmp_path)
    finally: [os.environ.pop(k, None) for k in SIM_KNOBS]


def worker_case(rng, wd):
    """a noisy synthetic reel with dropouts (blocks the first parameter set fails on), decoded by 2..5 worker processes with small
    shares and margins, units cut at random: the .tap and the summary line must be the reference's"""
    from readtape_b200 import synth, tbin
    clean = rng.random() < 0.4            # a reel the split should succeed on: quiet gaps, units cut where the unit finder cuts them
    nb = int(rng.integers(6, 24)); noise = float(rng.choice([2.0, 5.0] if clean else [2.0, 5.0, 40.0, 150.0])); wob = float(rng.choice([0.0, 0.01, 0.04]))
    hdr, rows = synth.nrzi_tape(nblocks=nb, seed=int(rng.integers(1, 1 << 30)), data_bytes=int(rng.integers(8, 300)), noise_mv=noise, wobble=wob)
    rows = rows.copy()
    for _ in range(int(rng.integers(0, 5))):
        a = int(rng.integers(0, len(rows) - 10)); b = a + int(rng.integers(10, 6000)); k = int(rng.integers(0, 9))
        rows[a:b, k] = (rows[a:b, k] * float(rng.choice([0.0, 0.1, 0.3]))).astype(np.int16)
    if rng.random() < 0.2: rows = rows[: int(rng.integers(len(rows) // 2, len(rows)))]
    with open(os.path.join(wd, "t.tbin"), "wb") as fh:
        fh.write(tbin.build_header(hdr)); rows.tofile(fh); fh.write(np.array([tbin.END_MARK], dtype="<i2").tobytes())
    if not clean and rng.random() < 0.5: _punch_all_track_dropouts(rng, wd)
    opts = ["-q", "-tap", "-nolog", "-nrzi", "-bpi=800", "-ips=50", str(rng.choice(["-nm", "-m"]))]
    if rng.random() < 0.3: opts.append("-nolabels")
    env = {"RT_WORKERS": str(int(rng.integers(2, 6))), "RT_WORKER_MIN_ROWS": str(int(rng.choice([2000, 20000, 60000]))),
           "RT_WORKER_MARGIN_ROWS": str(int(rng.choice([1000, 30000, 200000])))}
    if not clean and rng.random() < 0.5: env["HOSTSIM_UNIT_ROWS"] = str(int(rng.choice([1024, 4096, 20000, 100000])))
    return opts, env, f"synthetic NRZI, {nb} blocks, noise {noise} mV, {len(rows)} rows, {env}"


def worker_capture_case(rng, wd):
    """a random window of a bundled capture (PE, GCR, NRZI; noise and drop-outs added) with a command line the worker split accepts"""
    opts, what = fz.capture_case(rng, wd)
    if opts is None: return None, None, None
    keep = [o for o in opts if o.split("=")[0] in ("-ntrks", "-order", "-pe", "-nrzi", "-gcr", "-bpi", "-ips", "-zeros", "-correct", "-nm", "-m", "-whirlwind", "-differentiate")]
    if "-nrzi" in keep and not any(o.startswith("-bpi") for o in keep): keep.append(str(rng.choice(["-bpi=800", "-bpi=556", "-bpi=200"])))
    opts = ["-q", "-tap", "-nolog"] + keep
    env = {"RT_WORKERS": str(int(rng.integers(2, 6))), "RT_WORKER_MIN_ROWS": str(int(rng.choice([5000, 20000, 60000]))),
           "RT_WORKER_MARGIN_ROWS": str(int(rng.choice([1000, 30000, 200000])))}
    if rng.random() < 0.3: env["HOSTSIM_UNIT_ROWS"] = str(int(rng.choice([4096, 20000, 100000])))
    return opts, env, what + f", {env}"


def worker_fuzz(sim, seed, ncases, tmp_path, make_case=None):
    rng = np.random.default_rng(seed); wd = str(tmp_path); failures = []; split = 0
    for it in range(ncases):
        opts, env, what = (make_case or worker_case)(rng, wd)
        if opts is None: continue
        res = []
        for exe, tag, e in ((REF, "ref", {}), (sim, "new", dict(env, RT_STATS="1"))):
            for f in os.listdir(wd):
                if f.startswith(tag): os.remove(os.path.join(wd, f))
            p = subprocess.run([exe] + opts + [f"-outf={wd}/{tag}", f"{wd}/t.tbin"], capture_output=True, text=True, errors="replace", timeout=900, env=dict(os.environ, **e))
            tapf = f"{wd}/{tag}.tap"
            lines = [l for l in p.stdout.strip().splitlines() if "B200 scan" not in l]
            res.append((p.returncode, open(tapf, "rb").read() if os.path.exists(tapf) else None, lines[-1:]))
            if tag == "new":
                tail = (p.stdout[-700:] + p.stderr[-300:]).replace("\n", " | ")
                if "B200 scan: worker" in p.stdout: split += 1
        if res[0] != res[1]:
            why = "return code" if res[0][0] != res[1][0] else ".tap" if res[0][1] != res[1][1] else "summary line"
            failures.append(f"case {it} of seed {seed}: {what}: {why} differs ({res[0][2]} / {res[1][2]}); rc {res[0][0]} / {res[1][0]}; output of the run under test: {tail}")
        if [f for f in os.listdir(wd) if ".part" in f]: failures.append(f"case {it} of seed {seed}: {what}: part files left behind")
    assert not failures, "\n".join(failures)
    return split


def test_hostsim_random_worker_splits(sim, tmp_path):
    assert worker_fuzz(sim, 303, 10, tmp_path) >= 3


def test_hostsim_random_worker_splits_of_capture_windows(sim, tmp_path):
    worker_fuzz(sim, 305, 8, tmp_path, worker_capture_case)
