#!/bin/bash
# tools/sanitize_shim.sh -- build the host shim (readtape_b200/host/readblock_b200.c, product code) and the CPU oracle with
# AddressSanitizer + UndefinedBehaviorSanitizer, linked against the CPU simulation of the C-ABI (tests/host_fast/hostsim.cu), into
# $OUT (default /tmp/rt_asan).  Build container only (needs /root/reference for the reference's objects: make -C readtape_b200/host).
# Then run any of the shim tests with that binary, e.g.
#    ASAN_OPTIONS=detect_leaks=0 HOSTSIM_NO_DEEPBIND=1 HOSTSIM_ORACLE=$OUT/libscan_oracle_sym.so $OUT/shim_asan -nrzi -tap -outf=/tmp/x capture.tbin
# (the sanitizer runtimes refuse RTLD_DEEPBIND, so the oracle is built with -Bsymbolic and hostsim told not to use the flag).
# End of round 2: the 12 whole-capture command lines, the unit-cut runs and ~500 randomised cases (synthetic tapes, all-track
# drop-outs, worker splits) ran clean.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT=${OUT:-/tmp/rt_asan}; mkdir -p "$OUT"
CC=${SAN_CC:-/usr/bin/gcc-13}          # a gcc with its sanitizer runtimes installed (the image's /opt/gcc has none)
OBJ=$ROOT/readtape_b200/bin/obj
make -C "$ROOT/readtape_b200/host" hostsim > /dev/null
$CC -O1 -g -std=c99 -ffp-contract=off -fno-fast-math -Wall -fPIC -shared -Wl,-Bsymbolic -fsanitize=address,undefined -I"$ROOT/include" \
    -o "$OUT/libscan_oracle_sym.so" "$ROOT/oracle/scan_oracle.c" "$ROOT/oracle/csv_oracle.c" -lm
$CC -O1 -g --std=c99 -D_DEFAULT_SOURCE -Wall -Wno-unused-function -ffp-contract=off -fsanitize=address,undefined -fno-omit-frame-pointer \
    -I/root/reference/src -I"$ROOT/include" -c -o "$OUT/readblock_asan.o" "$ROOT/readtape_b200/host/readblock_b200.c"
$CC -fsanitize=address,undefined -o "$OUT/shim_asan" $OBJ/decoder.o $OBJ/decode_pe.o $OBJ/decode_nrzi.o $OBJ/decode_gcr.o $OBJ/decode_ww.o $OBJ/parmsets.o \
    $OBJ/textfile.o $OBJ/ibmlabels.o $OBJ/tapread.o $OBJ/trace.o $OBJ/readtape_weak.o "$OUT/readblock_asan.o" \
    -Wl,--wrap=init_trackstate -Wl,--wrap=ww_init_blockstate -Wl,--wrap=init_trackpeak_state -Wl,--wrap=compute_avg_height \
    -L"$ROOT/tests/host_fast/_build" -lhostsim -Wl,-rpath,"$ROOT/tests/host_fast/_build" -lm -lpthread
echo "built $OUT/shim_asan"
