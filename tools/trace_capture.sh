#!/bin/bash
# RT_TRACE laps of one whole-program run on a bundled capture: tools/trace_capture.sh <label of tests/golden/full_outputs.json>
python - "$1" <<'PY'
import json, os, subprocess, sys, tempfile
sys.path.insert(0, '.')
from oracle import captures
doc = json.load(open('tests/golden/full_outputs.json'))[sys.argv[1]]
cap = captures.full_path(doc['capture'])
with tempfile.TemporaryDirectory() as d:
    for i in range(2):
        r = subprocess.run(['readtape_b200/bin/readtape_b200'] + doc['options'].split() + [f'-outf={d}/o', cap], capture_output=True, text=True,
                           env=dict(os.environ, RT_STATS='2', RT_TRACE='1'))
    print(r.stderr[-6000:])
    print('\n'.join(l for l in r.stdout.splitlines() if 'B200 scan' in l or l.startswith('[')))
PY
