#!/bin/bash
# tools/quick_gpu_check.sh -- a seconds-long hardware check of the product binary without Python: two bundled captures
# (staged by hand under gpurun_in/ as xz) through readtape_b200, outputs compared with the SHA-256 of tests/golden/full_outputs.json.
# Microdata_20blks exercises the speculative scan + exact spans (k_ctx_scan with the skip-ahead), 132_pt1 (Whirlwind) runs
# entirely on k_ctx_scan.  Used at the end of round 2 when the GPU budget allowed one call of about a minute.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/quick && cd gpurun_out/quick
B=${QUICK_BIN:-../../readtape_b200/bin/readtape_b200}
check() {  # name, file, want-sha, options...
  local name=$1 f=$2 want=$3; shift 3
  xz -dkc ../../gpurun_in/$name.tbin.xz > $name.tbin || return 1
  local t0=$(date +%s%N)
  RT_STATS=1 $B "$@" -outf=$PWD/$name $PWD/$name.tbin > $name.stdout 2>&1; local rc=$?
  local t1=$(date +%s%N)
  local got=$(sha256sum $f | cut -d' ' -f1)
  grep -h "B200 scan:" $name.stdout | tail -3
  if [ "$rc" = 0 ] && [ "$got" = "$want" ]; then echo "QUICK $name: IDENTICAL rc=$rc $(( (t1 - t0) / 1000000 )) ms"; else echo "QUICK $name: DIFFERENT rc=$rc got=$got"; tail -5 $name.stdout; fi
  rm -f $name.tbin
}
nvidia-smi --query-gpu=name --format=csv,noheader
if [ "${QUICK_SET:-a}" = a ]; then
check Microdata_20blks Microdata_20blks.001.bin 21af9d8842f39722fcd15452662f1709cf1c35b7450810f21470ae2605ce9260 -m -nrzi -hex -ascii -v
check 132_pt1 132_pt1.tap d861b6751eef576400d546d0e7a97428621197fc1b915ef75b1a6dd6781a4e89 -whirlwind -fluxdir=auto -tap -deskew -octal2 -flexo -v
elif [ "$QUICK_SET" = c ]; then   # host-side lookup rules after a rebuild: one peak-detector capture, one zero-crossing capture
check Microdata_20blks Microdata_20blks.001.bin 21af9d8842f39722fcd15452662f1709cf1c35b7450810f21470ae2605ce9260 -m -nrzi -hex -ascii -v
check 1kblks_43blks 1kblks_43blks.tap 25c703cb5dab51110e8b37091233f98a328e67414497174ca1d8233be04738bd -m -gcr -ips=50 -order=76543210p -zeros -correct -tap -ascii -linefeed -v
else   # the two GCR -zeros captures: the zero-crossing fast path and its (tightened) unit-equivalence rule
check sf93_8blks sf93_8blks.tap 452a3e2496df04846539524ae1d5a1a80d6cdcd5d1150c7c19b4efa580c575cf -m -gcr -ips=50 -zeros -correct -tap -ascii -linefeed -v
check 1kblks_43blks 1kblks_43blks.tap 25c703cb5dab51110e8b37091233f98a328e67414497174ca1d8233be04738bd -m -gcr -ips=50 -order=76543210p -zeros -correct -tap -ascii -linefeed -v
fi
