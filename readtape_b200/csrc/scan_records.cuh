/* readtape_b200/csrc/scan_records.cuh -- K3c phase B1: one record per candidate row, computed row-parallel.
 *
 * Phase B of the two-pass peak scan used to do, at every candidate row and inside its sequential per-track walk, a scan of the
 * window's samples and -- for bottom candidates -- the reconstruction of the lazily refreshed minimum (decoder.c:765) by hopping
 * from refresh to refresh.  None of that depends on the sequential state: the window is the plane, and the lazy minimum at a row
 * is a PURE function of the samples once the scan has passed one canonical row (it only changes at refresh rows, and a refresh
 * happens when the window maximum leaves -- an `acan` row -- or when the leftmost sample carrying the minimum leaves).  So it is
 * computed here for every candidate row of a plane at once, by as many threads as there are candidates, and written as a 24-byte
 * record; the sequential walk (scan_sparse.cuh, record mode) then only streams through records: threshold tests with the current
 * AGC state, peak-time refinement from the neighbours stored in the record, feedback, event.
 *
 * Records of one track are grouped by 2048-row tile of the plane; within a tile they are in row order, tiles are placed in the
 * record pool by one atomic add each (rec_tile_base / rec_tile_cnt give the place).
 * Host + device code: tests/host_fast builds the records on the CPU with the same function.
 */
#pragma once
#include <stdint.h>
#include "rt_dev.h"

#ifndef RT_FHD
#define RT_FHD __host__ __device__ __forceinline__
#endif

#define RT_REC_TILE      2048          /* plane rows per record tile (64 mask words) */
#define RT_REC_NOPOS     255           /* posm: the lazy minimum could not be derived here (no canonical row within reach) */
#define RT_REC_REACH     4096          /* rows searched backwards for the last canonical row */

struct CandRec {                       /* 24 bytes */
   uint32_t row;                       /* plane row of the candidate (the window is plane[row-w+1 .. row]) */
   int16_t  S, xl, xr, m;              /* window maximum, left edge, right edge, lazy minimum at this row */
   uint8_t  posS, posm;                /* window position of the leftmost maximum / of the leftmost sample carrying m */
   int16_t  sprev, snext;              /* neighbours of the maximum inside the window (0 outside) */
   int16_t  mprev, mnext;              /* neighbours of the carrier of m inside the window (0 outside) */
   uint16_t pad;
};

namespace rtrec {

/* the record of plane row p (p >= w, p < rows of the plane); T0: the mask threshold of the track (the lazy minimum is only derived
   where the bottom pre-filter can pass at all: min(l, r) - Wmin >= T0) */
RT_FHD CandRec make_record(const int16_t *plane, const uint32_t *acan, int w, int T0, uint64_t p) {
   CandRec r;
   const int16_t *win = plane + (p - (uint64_t)w + 1u);
   uint32_t kmax = 0u, kmin = 0xffffffffu;
   for (int i = 0; i < w; ++i) {
      const uint32_t kb = ((uint32_t)((int)win[i] + 32768) << 6) + (uint32_t)i;
      if (kb < kmin) kmin = kb;
      const uint32_t kt = kb + (uint32_t)(63 - 2 * i);
      if (kt > kmax) kmax = kt; }
   const int S = (int)(kmax >> 6) - 32768, posS = 63 - (int)(kmax & 63u);
   int m = (int)(kmin >> 6) - 32768, posm = (int)(kmin & 63u);
   const int xl = win[0], xr = win[w - 1];
   r.row = (uint32_t)p; r.S = (int16_t)S; r.xl = (int16_t)xl; r.xr = (int16_t)xr; r.posS = (uint8_t)posS;
   r.sprev = posS > 0 ? win[posS - 1] : (int16_t)0; r.snext = posS < w - 1 ? win[posS + 1] : (int16_t)0;
   r.pad = 0;
   if ((xl < xr ? xl : xr) - m >= T0) {
      /* the bottom test can pass: the exact lazy minimum.  A refresh at or after the row at which the window's (leftmost) minimum
         entered makes it that minimum ... */
      const uint64_t ws = p - (uint64_t)w + 1u;
      bool fresh = false;
      for (uint64_t q = ws + (uint64_t)posm; q <= p && !fresh; ++q) fresh = (acan[q >> 5] >> (q & 31)) & 1u;
      if (!fresh) {
         /* ... otherwise it is what the last canonical row left, carried forward from refresh to refresh */
         uint64_t a = ws + (uint64_t)posm; bool have = false;                /* an acan row in [ws + posm, p] would have been "fresh" */
         const uint64_t lo = p > RT_REC_REACH ? p - RT_REC_REACH : 0;
         while (a > lo) { --a; if ((acan[a >> 5] >> (a & 31)) & 1u) { have = true; break; } }
         if (!have) posm = RT_REC_NOPOS;
         else {
            uint64_t rr = a;
            for (;;) {
               const int16_t *ww = plane + (rr - (uint64_t)w + 1u);
               uint32_t km = 0xffffffffu;
               for (int i = 0; i < w; ++i) { const uint32_t kb = ((uint32_t)((int)ww[i] + 32768) << 6) + (uint32_t)i; if (kb < km) km = kb; }
               m = (int)(km >> 6) - 32768;
               const uint64_t at = rr - (uint64_t)w + 1u + (km & 63u);
               if (at + (uint64_t)w > p) { posm = (int)(at - ws); break; }
               rr = at + (uint64_t)w; } } } }
   else posm = RT_REC_NOPOS;     /* m is the window's minimum here, a lower bound of the lazily kept one (decoder.c:765 never lowers the running
                                    minimum when a sample enters): enough to rule the bottom test out, but NOT the detector's state -- the walk
                                    must not adopt it (found by the host-build fuzz: a later dense-mode stretch started from a wrong minimum) */
   r.m = (int16_t)m; r.posm = (uint8_t)posm;
   if (posm != RT_REC_NOPOS) { r.mprev = posm > 0 ? win[posm - 1] : (int16_t)0; r.mnext = posm < w - 1 ? win[posm + 1] : (int16_t)0; }
   else r.mprev = r.mnext = 0;
   return r; }

}  // namespace rtrec
