#!/bin/bash
# tools/memcheck.sh -- compute-sanitizer memcheck over the smoke test and a cross-section of the GPU parity tests (B200 box).
# Round 1: found one out-of-bounds write (k_quiet_bitmap wrote up to 3 words past the quiet bitmap when the granule count was
# not a multiple of 256; fixed), 0 errors since.  Round 2 adds the exact scan's skip-ahead, the tile digest, the candidate records
# and the fused ingest kernel to the cross-section.
set -e
compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
   -k "bulk_scan_lookup and Microdata or data_driven and LJS009 or peak_mask and tss or fanout and 1600 or exact_scan and 132_pt1 or exact_scan and Microdata_20blks-"
RT_SPARSE_RECORDS=1 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "bulk_scan_lookup and Microdata"
# round 2, CSV ingest: line index, csv_preread's maximum, the staged conversion kernel and its two fall-back scanners
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_csv.py -x -q -m gpu -k "cuda_matches_reference_golden and not long or line_index or random_text"
