"""Whole-program comparison: readtape with the B200 scan (readtape_b200/bin/readtape_b200) against the unmodified reference
(oracle/_ref/readtape_ref) on a synthetic tape of N super-tiles; checks that the two .tap files are identical.
Usage (GPU box): python tools/e2e_product.py [super_tiles]"""
import os, sys, time, subprocess, tempfile, numpy as np
sys.path.insert(0, os.getcwd())
from readtape_b200 import synth, tbin
tile = synth.nrzi_tile()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
d = tempfile.mkdtemp(dir="/dev/shm")
path = os.path.join(d, "big.tbin")
tbin.write_tbin(path, synth.nrzi_header(), np.concatenate([tile] * reps))
nrows = tile.shape[0] * reps
for exe, env in (("readtape_b200/bin/readtape_b200", {"RT_STATS": "1"}), ("readtape_b200/bin/readtape_b200", {"RT_STATS": "1"}), ("oracle/_ref/readtape_ref", {})):
    t0 = time.time()
    r = subprocess.run([exe, "-q", "-nm", "-nrzi", "-bpi=800", "-ips=50", "-tap", "-nolog", "-nolabels", f"-outf={d}/out_{os.path.basename(exe)}", path],
                       capture_output=True, text=True, env=dict(os.environ, **env))
    dt = time.time() - t0
    print(os.path.basename(exe), f"rc={r.returncode} {dt:.2f}s  {nrows*9/dt/1e6:.1f} M track-samples/s")
    for l in r.stdout.splitlines():
        if "B200 scan:" in l:
            print("   ", l.strip())
a = open(f"{d}/out_readtape_b200.tap", "rb").read(); b = open(f"{d}/out_readtape_ref.tap", "rb").read()
print("tap identical:", a == b, len(a))
