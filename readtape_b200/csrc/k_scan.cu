/* readtape_b200/csrc/k_scan.cu -- scan kernels built on the exact generic per-track scan.
 *
 *  k_ctx_reset / k_ctx_scan   the stateful exact scan behind rt_scan_* (one thread per track,
 *                             state lives in device memory between launches; Whirlwind, retries,
 *                             anything the speculative scan cannot prove)
 *  k_units_scan               the speculative whole-tape scan: one thread per (unit, track),
 *                             fresh RT_RESET_FULL at the unit start, events into the chunk pool,
 *                             plus the per-track data rt_bulk_lookup() needs for its proof
 */
#include <stdio.h>
#include <stdlib.h>
#include "scan_generic.cuh"
#include "kernels.h"
#include "emit.cuh"

using namespace rtgen;

/* ---- stateful context ---------------------------------------------------------------------- */
__global__ void k_ctx_reset(DevCfg c, TrkState *st, SkewState *sk, int kind, uint64_t row, int time_is_zero) {
   int trk = threadIdx.x;
   if (trk >= c.ntrks) return;
   TrkState &t = st[trk]; SkewState &s = sk[trk];
   if (kind == RT_RESET_FULL) reset_full(c, t, s, trk, row, time_is_zero != 0);
   else if (kind == RT_RESET_WW_PARTIAL) {          /* decode_ww.c:42: t_lastpeak = t_prevlastpeak = 0 */
      t.t_lastpeak = 0;
      t.init_row = row + (uint64_t)trk + (time_is_zero ? 1u : 0u);
      t.pure_from = RT_NOROW; }                     /* until the re-initialisation row has passed (track_row sets it) */
   else if (kind == RT_RESET_PEAKSTATE) {           /* decoder.c:413-423 */
      for (int i = 0; i < RT_MAXSKEWSAMP; ++i) s.vdelayed[i] = 0;
      s.ndx_next = s.slots_filled = 0;
      t.left = t.right = 0; t.minv = t.maxv = 0; t.countdown = 0;
      t.pure_from = row + (uint64_t)(2 * c.width + c.skew[trk] + 2); }
   else if (kind == RT_RESET_NONE) t.pure_from = row + (uint64_t)(2 * c.width + c.skew[trk] + 2); }   /* repositioned, or the skew changed: the ring and the FIFO refill first */

__global__ void k_ctx_set_avg_height(TrkState *st, int trk, float v) {
   st[trk].avg_height = v; st[trk].avg_height_count = 0; st[trk].avg_height_sum = 0; }

__global__ void k_ctx_scan(DevCfg c, TrkState *st, SkewState *sk, uint64_t row_from, uint64_t row_to,
                           rt_event *evbuf, uint32_t cap, uint32_t *counts, uint32_t *failed, int trace) {
   /* One block of one thread per track.  (Measured: a second warp pulling the samples and mask words into L1 ahead of the walker
      changes nothing -- 1.085 s against 1.084 s on the Whirlwind reel; the walk is bound by its own dependent instructions, ~2.5 us
      per walked row or jump, not by misses.) */
   const int trk = blockIdx.x;
   if (trk >= c.ntrks) return;
   const int16_t *plane = c.planes + (size_t)trk * c.plane_stride;
   TrkState t = st[trk]; SkewState s = sk[trk];
   FlatEmit em{evbuf + (size_t)trk * cap, cap, 0, (uint8_t)trk};
   CtxStats cs{0, 0, 0, 0, 0};
   ctx_scan_rows(c, t, s, trk, plane, row_from, row_to, em, cs);
   if (trace) printf("[k_ctx_scan] trk %d rows %llu: walked %llu, %llu jumps over %llu rows, threshold below the mask's at %llu rows, next candidate too near at %llu, T0 %d gain %.3f avg_height %.3f\n",
                     trk, (unsigned long long)(row_to - row_from), cs.walked, cs.jumps, cs.jumped, cs.nothr, cs.near, c.T0[trk], t.agc_gain, t.avg_height);
   st[trk] = t; sk[trk] = s;
   counts[trk] = em.n;
   if (t.failed) atomicMax(failed, (uint32_t)t.failed); }

/* ---- speculative whole-tape scan (generic detector code) -------------------------------------- */
__global__ void __launch_bounds__(128)
k_units_scan(DevCfg c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta,
             rt_event *pool, uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks,
             float quiet_thr, int quiet_thr_lsb, unsigned long long *rows_scanned) {
   const uint64_t total = (uint64_t)nunits * (uint64_t)c.ntrks;
   for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (uint64_t)gridDim.x * blockDim.x) {
      const uint32_t u = (uint32_t)(f / c.ntrks); const int trk = (int)(f % c.ntrks);
      const UnitDesc ud = units[u];
      TrkState t; SkewState s;
      reset_full(c, t, s, trk, ud.row0, row_time(c, ud.row0) == 0.0);
      const int16_t *plane = c.planes + (size_t)trk * c.plane_stride;
      PoolEmit em{pool, chunk_next, cursor, cap_chunks, RT_NOCHUNK, RT_NOCHUNK, 0, RT_NOROW, (uint8_t)trk, RT_NOROW};
      /* proof data, collected only before the first event (DESIGN.md "unit equivalence"):
           last_loud     last loud row (QuietTracker) seen so far, starting DevCfg::prescan_rows before the unit
           sync_row      LAST row before the first event at which this scan's state is canonical and the
                         row is not loud; loud_at_sync = last_loud at that moment
           sync_first    FIRST such row with nothing loud since the unit start (used to chain units)
         canonical = a pure function of the samples: for the peak detector the running maximum has just
         left a FULL window (both scans rescan there, which also refreshes the lazily kept minimum);
         for the zero-crossing detectors any row once v_prev and the deskew FIFO are warm. */
      QuietTracker qt; qt.init(c, trk, quiet_thr, quiet_thr_lsb);
      const uint64_t pre0 = ud.row0 > (uint64_t)c.prescan_rows ? ud.row0 - (uint64_t)c.prescan_rows : 0;
      for (uint64_t j = pre0; j < ud.row0; ++j) qt.feed(c, plane, j, raw_at(c, plane, j));
      const uint64_t quiet_from = qt.last_loud == RT_NOROW ? pre0 : qt.last_loud + 1;
      uint64_t sync_row = RT_NOROW, loud_at_sync = RT_NOROW, sync_first = RT_NOROW, sync_early = RT_NOROW, loud_early = RT_NOROW;
      bool early_frozen = false;
      const int lead = max(trk + (row_time(c, ud.row0) == 0.0 ? 1 : 0), c.skew[trk]);
      const uint64_t own_fill = ud.row0 + (uint64_t)(c.det == RT_DET_PEAK ? lead + c.width + 1 : lead + 2);
      for (uint64_t row = ud.row0; row < ud.row_end; ++row) {
         float v_raw;
         unsigned probe = track_row(c, t, s, trk, plane, row, em, &v_raw);
         if (em.n == 0) {                              /* still before the first event: keep the proof data */
            qt.feed(c, plane, row, v_raw);
            const bool canonical = c.det == RT_DET_PEAK ? (probe & 2) != 0 : row >= own_fill;
            if (sync_early != RT_NOROW && qt.last_loud != loud_early) early_frozen = true;   /* the first quiet stretch is over */
            if (canonical && (qt.last_loud == RT_NOROW || qt.last_loud < row)) {
               sync_row = row; loud_at_sync = qt.last_loud;
               if (!early_frozen) { sync_early = row; loud_early = qt.last_loud; }
               if (sync_first == RT_NOROW && row >= own_fill && (qt.last_loud == RT_NOROW || qt.last_loud < ud.row0)) sync_first = row; } } }
      TrkMeta m;
      m.first_event_row = em.first_row; m.sync_row = sync_row; m.last_loud_row = loud_at_sync; m.sync_first = sync_first; m.quiet_from = quiet_from; m.sync_early = sync_early; m.loud_early = loud_early;
      m.first_chunk = em.first_chunk; m.nevents = em.n; m.failed = t.failed; m.pad = 0;
      m.last_event_row = em.last_row; m.quiet_tail_from = RT_NOROW;
      meta[f] = m;
      atomicAdd(&rows_scanned[0], (unsigned long long)(ud.row_end - ud.row0));
      if (em.n) atomicAdd(&rows_scanned[1], (unsigned long long)em.n); } }

/* ---- launchers ------------------------------------------------------------------------------ */
void launch_ctx_reset(const DevCfg &c, TrkState *st, SkewState *sk, int kind, uint64_t row, int tz, cudaStream_t s) {
   k_ctx_reset<<<1, 32, 0, s>>>(c, st, sk, kind, row, tz); }
void launch_ctx_set_avg_height(TrkState *st, int trk, float v, cudaStream_t s) {
   k_ctx_set_avg_height<<<1, 1, 0, s>>>(st, trk, v); }
void launch_ctx_scan(const DevCfg &c, TrkState *st, SkewState *sk, uint64_t from, uint64_t to, rt_event *ev,
                     uint32_t cap, uint32_t *counts, uint32_t *failed, cudaStream_t s) {
   static const int trace = getenv("RT_TRACE_CTX") != nullptr;
   /* one warp per track would waste 31 lanes; tracks are independent, so spread them over blocks of 1 thread
      each to get them on different SMs (each is a long serial walk) */
   k_ctx_scan<<<c.ntrks, 1, 0, s>>>(c, st, sk, from, to, ev, cap, counts, failed, trace); }
void launch_units_scan(const DevCfg &c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta, rt_event *pool,
                       uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks, float quiet_thr, int quiet_thr_lsb,
                       unsigned long long *rows_scanned, int grid, cudaStream_t s) {
   k_units_scan<<<grid, 128, 0, s>>>(c, units, nunits, meta, pool, chunk_next, cursor, cap_chunks, quiet_thr, quiet_thr_lsb, rows_scanned); }
