#!/usr/bin/env python
"""oracle/make_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/ from the REAL reference.

Runs only in the build container (needs /root/reference and oracle/_ref built by oracle/Makefile):

  1. stages the example captures the GPU-box tests use into oracle/_ref/examples/ (whole files
     for the small ones, a prefix cut to --rows rows + end marker for the big ones),
  2. runs the unmodified reference (oracle/_ref/readtape_ref) with the example's own Makefile
     command line on the ORIGINAL capture and checks every golden .tap/.bin of the reference
     byte-for-byte (tests/golden/reference_goldens.json records the outcome),
  3. runs the instrumented reference (oracle/_ref/readtape_evdump) on each STAGED capture and
     commits, per capture, the reset/segment table and a SHA-256 of the reference's events of
     every decode segment (tests/golden/<name>.segments.json) plus the decoded .tap
     (tests/golden/<name>.tap) -- these pin both the CPU oracle and the CUDA path,
  4. replays every segment through the CPU oracle and fails if any event differs.

Usage: python oracle/make_golden.py [--skip-oracle-check] [--stage-only]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from readtape_b200 import abi, evlog, tbin  # noqa: E402

REF = os.environ.get("RT_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")

# name, directory, Makefile options (examples/*/Makefile), rows to stage (None = whole file)
EXAMPLES = [
    ("Microdata_20blks", "9trk_NRZI", "-v -m -nrzi -hex -ascii", None),
    ("PLAGO_beginning", "9trk_NRZI", "-v -m -nrzi -ips=50 -deskew -ebcdic -linefeed", 700_000),
    ("1600bpi_ukn_6s", "9trk_PE", "-v -m -ntrks=9 -pe -bpi=1600 -ips=50 -order=01234576p -tap -ascii -linefeed", 600_000),
    ("LJS009_part1_39blks", "9trk_PE", "-v -m -ntrks=9 -pe -bpi=1600 -ips=50 -tap -ebcdic -linesize=137", 600_000),
    ("1kblks_43blks", "9trk_GCR", "-v -m -gcr -ips=50 -order=76543210p -zeros -correct -tap -ascii -linefeed", 600_000),
    ("sf93_8blks", "9trk_GCR", "-v -m -gcr -ips=50 -zeros -correct -tap -ascii -linefeed", 600_000),
    ("analog", "9trk_GCR", "-v -gcr -ips=125 -differentiate -zeros -tap -ascii", None),
    ("SRI_SDS_102715028_4secs", "7trk_NRZI", "-v -m -nrzi -ntrks=7 -order=543210p -tap -SDS -linesize=144", 700_000),
    ("tss_4secs", "7trk_NRZI", "-v -m -nrzi -ntrks=7 -tap", 700_000),
    ("132_pt1", "6trk_Whirlwind", "-whirlwind -v3 -fluxdir=auto -tap -deskew -octal2 -flexo", None),
]
# extra command lines on the staged captures: BASELINE.json config 1 ("-nm -nrzi -tap", one parmset)
EXTRA = [
    ("Microdata_20blks", "nm_tap", "-nm -nrzi -tap"),
    ("PLAGO_beginning", "nm_tap", "-nm -nrzi -tap"),
]


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def run(binary, opts, src, outbase, evdump=None):
    env = dict(os.environ)
    if evdump:
        env["RT_EVDUMP"] = evdump
    cmd = [os.path.join(OUT, binary)] + opts.split() + [f"-outf={outbase}", src]
    with open(outbase + ".stdout", "w") as so:
        rc = subprocess.run(cmd, stdout=so, stderr=subprocess.STDOUT, env=env, cwd=os.path.dirname(outbase)).returncode
    if rc != 0:
        raise SystemExit(f"{' '.join(cmd)} exited {rc}")


def stage(name, directory, rows):
    src = os.path.join(REF, "examples", directory, name + ".tbin")
    dst = os.path.join(OUT, "examples", name + ".tbin")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if rows is None:
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
        return dst
    hdr, allrows = tbin.read_tbin(src)
    with open(src, "rb") as fh:
        head = fh.read(hdr.payload_offset)
    keep = np.array(allrows[:rows])
    with open(dst, "wb") as fh:
        fh.write(head)
        keep.tofile(fh)
        fh.write(np.array([tbin.END_MARK], dtype="<i2").tobytes())
    return dst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-oracle-check", action="store_true")
    ap.add_argument("--only", default=None)
    ap.add_argument("--stage-only", action="store_true",
                    help="only (re)create oracle/_ref/examples/ (what build() does in a fresh container)")
    args = ap.parse_args()
    if args.stage_only:
        for name, directory, _opts, rows in EXAMPLES:
            dst = os.path.join(OUT, "examples", name + ".tbin")
            if not os.path.exists(dst):
                stage(name, directory, rows)
        print("staged", len(EXAMPLES), "captures under", os.path.join(OUT, "examples"))
        return
    os.makedirs(GOLD, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="rtgold_")
    ref_status = {}
    failures = 0
    oracle = None if args.skip_oracle_check else abi.load_oracle()
    for name, directory, opts, rows in EXAMPLES:
        if args.only and args.only != name:
            continue
        # (2) the reference against its own goldens, on the original capture
        src = os.path.join(REF, "examples", directory, name + ".tbin")
        base = os.path.join(tmp, name)
        run("readtape_ref", opts, src, base)
        exp_dir = os.path.join(REF, "examples", directory, "expected_results")
        for g in sorted(os.listdir(exp_dir)):
            if not (g.endswith(".tap") or g.endswith(".bin")) or not g.startswith(name):
                continue
            made = os.path.join(tmp, g)
            ok = os.path.exists(made) and open(made, "rb").read() == open(os.path.join(exp_dir, g), "rb").read()
            ref_status[g] = {"matches_reference_golden": bool(ok), "sha256": sha256_file(os.path.join(exp_dir, g))}
            print(f"  reference vs golden {g}: {'ok' if ok else 'MISMATCH'}")
            failures += 0 if ok else 1
        # (1)+(3) staged capture through the instrumented reference
        staged = stage(name, directory, rows)
        variants = [("", opts)] + [(tag, o) for (n, tag, o) in EXTRA if n == name]
        for tag, o in variants:
            label = name + ("." + tag if tag else "")
            sbase = os.path.join(tmp, "staged_" + label)
            evfile = sbase + ".ev"
            run("readtape_evdump", o, staged, sbase, evdump=evfile)
            heads = evlog.parse_heads(evfile)
            segs = evlog.parse(evfile)
            nev = sum(len(s.events) for s in segs)
            outputs = {}
            for ext in (".tap",):
                if os.path.exists(sbase + ext):
                    shutil.copyfile(sbase + ext, os.path.join(GOLD, label + ext))
                    outputs[label + ext] = sha256_file(sbase + ext)
            for f in sorted(os.listdir(tmp)):
                if f.startswith("staged_" + label) and f.endswith(".bin"):
                    outputs[f[len("staged_"):]] = sha256_file(os.path.join(tmp, f))
            extra = {"capture": name + ".tbin", "capture_sha256": sha256_file(staged), "options": o,
                     "staged_rows": rows, "heads": heads, "reference_outputs": outputs,
                     "reference_version": "readtape 3.18 (LenShustek/readtape @ 85d8d62), gcc -O2"}
            evlog.save_fixture(os.path.join(GOLD, label + ".segments.json"), segs, extra)
            print(f"  {label}: {len(segs)} decode segments, {nev} reference events")
            # (4) the oracle restatement against the reference's events
            if oracle is not None:
                hdr, rws = tbin.read_tbin(staged, nheads=heads["nheads"])
                tape = oracle.open(evlog.desc_from_heads(heads))
                tape.upload(np.asarray(rws))
                bad = 0
                for seg, got in evlog.replay(tape, segs):
                    msg = evlog.compare(seg, got)
                    if msg:
                        bad += 1
                        if bad <= 3:
                            print("    ORACLE MISMATCH:", msg)
                tape.close()
                print(f"    oracle vs reference: {len(segs) - bad}/{len(segs)} segments identical")
                failures += bad
    with open(os.path.join(GOLD, "reference_goldens.json"), "w") as fh:
        json.dump(ref_status, fh, indent=1, sort_keys=True)
        fh.write("\n")
    shutil.rmtree(tmp, ignore_errors=True)
    if failures:
        raise SystemExit(f"{failures} failure(s)")
    print("all golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
