import os, sys, time, subprocess, tempfile, numpy as np, shutil
sys.path.insert(0, os.getcwd())
from readtape_b200 import synth, tbin
tile = synth.nrzi_tile()
reps = int(sys.argv[1]); workers = sys.argv[2]
d = tempfile.mkdtemp(dir="/dev/shm")
path = os.path.join(d, "reel.tbin")
with open(path, "wb") as fh:
    fh.write(tbin.build_header(synth.nrzi_header()))
    for _ in range(reps):
        tile.tofile(fh)
    fh.write(np.array([tbin.END_MARK], dtype="<i2").tobytes())
for w in workers.split(","):
    t0 = time.time()
    r = subprocess.run(["readtape_b200/bin/readtape_b200", "-q", "-nm", "-nrzi", "-bpi=800", "-ips=50", "-tap", "-nolog", "-nolabels", f"-outf={d}/o", path],
                       capture_output=True, text=True, env=dict(os.environ, RT_STATS="2", RT_WORKERS=w))
    dt = time.time() - t0
    print(f"== workers {w}: rc {r.returncode} {dt:.2f} s  {reps*tile.shape[0]*9/dt/1e9:.2f} G ts/s")
    for l in r.stdout.splitlines():
        if "B200 scan" in l and ("rt_open" in l or "opening" in l or "whole-tape" in l) and " worker " not in l:
            print("   ", l.strip())
shutil.rmtree(d)
