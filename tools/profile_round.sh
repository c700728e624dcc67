#!/bin/bash
# tools/profile_round.sh -- the commands behind profiles/ (run on a B200 box from the repo root, e.g. through gpurun).
# A number printed by a run under ncu is never a bench value; bench lines come from plain `python bench.py`.
set -e
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_full.json                                   # value / e2e / roofline / cpu_baseline
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_arm.json
# launch list of the same command (per-launch times are cold-cache and serialised: compare SHARES with the live timing)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
# one full capture of each kernel of a step, with source correlation
ncu --set full --import-source on --clock-control none -k regex:"k_peak_masks|k_units_sparse|k_ingest_tma" -c 3 -o gpurun_out/prof_full \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
# read back here:
#   ncu -i gpurun_out/prof_full.ncu-rep --page raw --csv > raw.csv
#   ncu -i gpurun_out/prof_full.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_units_sparse > src.csv
#   python tools/ncu_lines.py src.csv 40
