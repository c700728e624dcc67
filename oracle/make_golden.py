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

from oracle.captures import EXAMPLES, stage_xz, staged_path, full_path  # noqa: E402  (the capture table lives there)

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
    """-> path of the capture the segment fixtures are generated from (whole file or prefix + end marker)"""
    stage_xz(name)
    return staged_path(name)


def full_outputs():
    """The reference on the WHOLE captures, with the example's own Makefile command line and the EXTRA ones: SHA-256 (and size) of
    every .tap/.bin it writes -> tests/golden/full_outputs.json.  For the Makefile command lines these are the reference-held
    goldens of examples/*/expected_results (checked here byte for byte)."""
    tmp = tempfile.mkdtemp(prefix="rtfull_")
    doc = {}
    for name, directory, opts, _rows in EXAMPLES:
        src = os.path.join(REF, "examples", directory, name + ".tbin")
        exp_dir = os.path.join(REF, "examples", directory, "expected_results")
        for tag, o in [("", opts)] + [(tag, o) for (n, tag, o) in EXTRA if n == name]:
            label = name + ("." + tag if tag else "")
            wd = os.path.join(tmp, label); os.makedirs(wd)
            run("readtape_ref", o, src, os.path.join(wd, name))
            outs = {}
            for f in sorted(os.listdir(wd)):
                if f.endswith(".tap") or f.endswith(".bin"):
                    made = os.path.join(wd, f)
                    held = os.path.join(exp_dir, f)
                    is_held = (not tag) and os.path.exists(held)
                    if is_held and open(made, "rb").read() != open(held, "rb").read():
                        raise SystemExit(f"reference output {f} differs from the reference-held golden")
                    outs[f] = {"sha256": sha256_file(made), "bytes": os.path.getsize(made), "reference_held_golden": bool(is_held)}
            doc[label] = {"capture": name, "options": o, "outputs": outs}
            print(f"  {label}: " + ", ".join(f"{k} ({v['bytes']} B{', golden' if v['reference_held_golden'] else ''})" for k, v in outs.items()))
    with open(os.path.join(GOLD, "full_outputs.json"), "w") as fh:
        json.dump(doc, fh, indent=1, sort_keys=True); fh.write("\n")
    shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-oracle-check", action="store_true")
    ap.add_argument("--only", default=None)
    ap.add_argument("--stage-only", action="store_true",
                    help="only (re)create oracle/_ref/examples/ (what build() does in a fresh container)")
    ap.add_argument("--full-only", action="store_true",
                    help="only regenerate tests/golden/full_outputs.json (the reference's outputs on the WHOLE captures)")
    args = ap.parse_args()
    if args.full_only:
        full_outputs()
        return
    if args.stage_only:
        for name, directory, _opts, rows in EXAMPLES:
            stage(name, directory, rows)
            full_path(name)
        print("staged", len(EXAMPLES), "captures under", os.path.join(OUT, "examples_full"), "(xz) and", os.path.join(OUT, "examples"))
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
