#!/bin/bash
# tools/profile_round.sh -- the commands behind profiles/ (run on a B200 box from the repo root, e.g. through gpurun).
# A number printed by a run under ncu is never a bench value; bench lines come from plain `python bench.py`.
set -x
mkdir -p gpurun_out/prof
O=gpurun_out/prof
python bench.py > $O/bench_full.json 2> $O/bench_full.err                       # value / e2e / roofline / cpu_baseline / product / verified
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2>/dev/null
# launch list of the same command (per-launch times are cold-cache and serialised: compare SHARES with the live timing)
# (since rt_prepare runs the ingest and mask kernels chunk by chunk there are ~140 launches per step: -c 1200 covers 2 + 3 steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-product --no-verify > /dev/null 2>&1
# one full capture of each kernel of a step, with source correlation
ncu --set full --import-source on --clock-control none -k regex:"k_peak_masks|k_units_sparse|k_ingest_tma" -c 3 -o $O/prof_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-product --no-verify > /dev/null 2>&1
# config 4: the zero-crossing kernel at full per-GPU size; config 5: the exact stateful scan
ncu --set full --import-source on --clock-control none -k regex:"k_units_zc" -c 1 -o $O/prof_zc -f \
    python bench.py --workload gcr --steps 1 --warmup 1 --no-verify > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_ctx_scan" -s 40 -c 2 -o $O/prof_ctx -f \
    python bench.py --workload ww --steps 1 --warmup 0 > /dev/null 2>&1
for f in prof_full prof_zc prof_ctx; do
   ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
done
ncu -i $O/prof_full.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_units_sparse > $O/src_sparse.csv 2>/dev/null
python tools/ncu_lines.py $O/src_sparse.csv 60 > $O/k_units_sparse_lines.txt 2>&1
rm -f $O/src_sparse.csv
ls -la $O
