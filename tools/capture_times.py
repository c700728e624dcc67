#!/usr/bin/env python
"""Whole-program wall time on the bundled captures (B200 box): readtape_b200 beside the unmodified reference binary, same command
lines as tests/golden/full_outputs.json.  The captures are 20-80 MB, so this table shows the fixed costs (CUDA start-up, upload)
and the exact-scan paths, not throughput.  Writes gpurun_out/capture_times.json."""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import captures  # noqa: E402

FULL = json.load(open(os.path.join(ROOT, "tests", "golden", "full_outputs.json")))
REF = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
NEW = os.path.join(ROOT, "readtape_b200", "bin", "readtape_b200")


def run(exe, doc, wd):
    cap = captures.full_path(doc["capture"])
    env = dict(os.environ, RT_STATS="2")
    t0 = time.perf_counter()
    r = subprocess.run([exe] + doc["options"].split() + [f"-outf={wd}/o", cap], capture_output=True, text=True, env=env, cwd=wd)
    dt = time.perf_counter() - t0
    out = {"seconds": round(dt, 3), "rc": r.returncode}
    m = re.search(r"B200 scan: ([\d.]+) s opening \+ upload, ([\d.]+) s in the scan library, ([\d.]+) s replaying", r.stdout)
    if m:
        out.update(open_upload_s=float(m.group(1)), scan_library_s=float(m.group(2)), replay_s=float(m.group(3)))
    m = re.search(r"B200 scan: ([\d.]+) s of the replay time were exact-scan spans", r.stdout)
    if m:
        out["exact_spans_s"] = float(m.group(1))
    m = re.search(r"B200 scan: (\d+) events, (\d+) speculative hits, (\d+) misses, (\d+) restarts, (\d+) exact spans", r.stdout)
    if m:
        out.update(events=int(m.group(1)), hits=int(m.group(2)), misses=int(m.group(3)), restarts=int(m.group(4)), exact_spans=int(m.group(5)))
    return out


def main():
    table = {}
    for label in sorted(FULL):
        doc = FULL[label]
        rows = os.path.getsize(captures.full_path(doc["capture"]))
        with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b:
            run(NEW, doc, b)                                   # warm the page cache and the driver
            table[label] = {"options": doc["options"], "bytes": rows, "reference_1core": run(REF, doc, a), "readtape_b200": run(NEW, doc, b)}
        t = table[label]
        print(f"{label:28s} {rows / 1e6:6.1f} MB  ref {t['reference_1core']['seconds']:6.2f} s   b200 {t['readtape_b200']['seconds']:6.2f} s   {t['readtape_b200']}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(table, open(os.path.join(ROOT, "gpurun_out", "capture_times.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
