#!/bin/bash
# tools/missinfo.sh <capture-name> -- run the product on a whole bundled capture with RT_STATS=2 and show why block decodes
# were not served by the speculative whole-tape scan (misses / restarts).  GPU box.
set -e
name=$1
python - "$name" <<'PY'
import sys, os, subprocess, tempfile
sys.path.insert(0, os.getcwd())
from oracle import captures
name = sys.argv[1]
e = captures.BY_NAME[name]
cap = captures.full_path(name)
d = tempfile.mkdtemp()
r = subprocess.run(["readtape_b200/bin/readtape_b200"] + [o for o in e[2].split() if o not in ("-v", "-v3")] + [f"-outf={d}/{name}", cap],
                   capture_output=True, text=True, env=dict(os.environ, RT_STATS="2"))
print("\n".join(l for l in r.stdout.splitlines() if "B200 scan" in l or l.lstrip().startswith(("trk", "unit"))))
PY
