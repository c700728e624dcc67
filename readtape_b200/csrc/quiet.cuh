/* readtape_b200/csrc/quiet.cuh -- int16-domain loudness tracker of the unit-equivalence proof.
 *
 * Used by both scan kernels for the moving-window peak detector.  Row j is "loud"
 * if the raw int16 samples of rows [j-L+1, j] (L = window width + skew delay: every window any
 * scan can hold at row j only contains samples of that span) have max-min >= thr LSBs, with thr
 * chosen below pkww_rise (cfg_host.h: quiet_thr_lsb).  "Not loud" proves that no scan in default
 * state can fire at row j; "loud" may be a false alarm.  Invert does not change a range, and
 * int16 -> volts is strictly monotone (readtape.c:1420), so the test is done on the raw samples.
 * Rows must be fed consecutively.
 */
#pragma once
#include "rt_dev.h"

#ifndef RT_HD
#define RT_HD __host__ __device__ inline
#endif

struct QuietInt {
   int runmin, runmax, thr, L; uint64_t last_loud; bool primed;
   RT_HD void init(int L_, int thr_) { L = L_; thr = thr_; last_loud = RT_NOROW; primed = false; runmin = runmax = 0; }
   RT_HD void feed(const int16_t *plane, uint64_t j, int x) {
      if (!primed) { runmin = runmax = x; primed = true; }
      if (x < runmin) runmin = x;
      if (x > runmax) runmax = x;
      if (runmax - runmin >= thr) {              /* re-anchor on the exact span; loud only if IT is */
         int mx = x, mn = x;
         uint64_t from = j + 1 >= (uint64_t)L ? j + 1 - (uint64_t)L : 0;
         for (uint64_t i = from; i < j; ++i) { int y = plane[i]; if (y > mx) mx = y; if (y < mn) mn = y; }
         runmin = mn; runmax = mx;
         if (mx - mn >= thr) last_loud = j; } } };
