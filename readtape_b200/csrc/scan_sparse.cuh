/* readtape_b200/csrc/scan_sparse.cuh -- K3c phase B: the sequential part of the moving-window peak detector,
 * visiting candidate rows only.
 *
 * Same contract as UnitScan (scan_fast.cuh) and the (unit, track) scan of k_units_scan (k_scan.cu): one lane scans
 * one track of one unit from a fresh RT_RESET_FULL and produces the identical events and proof data.  Phase A
 * (scan_masks.cuh) has already evaluated the pure, row-parallel part of lookfor_peak (decoder.c:751-810) for every
 * row of the plane and left two bit planes: `cand` (rows that can pass the shape tests for any threshold >= T0) and
 * `acan` (rows where the window maximum left the window, i.e. where the reference rescans: decoder.c:767).  What is
 * left is sequential but sparse -- the AGC-dependent threshold (decoder.c:785), the blind countdown (:778), the lazily
 * refreshed minimum (:765) -- and is only needed AT candidate rows:
 *
 *   sparse mode   next set bit of `cand` at or after the end of the blind stretch -> window maximum, edges and the
 *                 position of the peak straight from the plane (w samples) -> exact float tests of decoder.c:790-803.
 *                 Only if the top test fails is the lazy minimum needed: m(p) = Wmin(a) at the last `acan` row a <= p,
 *                 then the recurrence  m <- Wmin(q) iff raw[q-w] == m  over (a, p]  (scan_fast.cuh, note 3).
 *   dense mode    a plain row-by-row restatement (window rescanned on every row).  Used for the first rows of a unit
 *                 -- while the window fills, while the deskew FIFO fills, and until the first `acan` row makes the
 *                 lazy minimum a pure function of the samples -- and whenever the current threshold bound T drops
 *                 below T0 (AGC gain high / low average height), until it is back.  Correctness therefore never
 *                 depends on T0; only speed does.
 *
 * The proof data of the unit-equivalence test concerns the rows before the first event: loud rows and
 * canonical (= acan) non-loud rows.  In sparse mode they are derived granule-wise when the first event is found
 * (advance_pre): 32 rows whose granule min/max keep the running span below the threshold cannot contain a loud row,
 * and their canonical rows are the bits of `acan`.
 *
 * Reference semantics (file:line in /root/reference/src): lookfor_peak decoder.c:751-810, refine_peak :700-749,
 * first-sample init :855-861, deskew FIFO :819-831, process_*_transition :560-609; feedback: feedback.cuh.
 * Host + device code: tests/host_fast runs the same functions on the CPU against the oracle.
 */
#pragma once
#include "scan_fast.cuh"
#include "scan_masks.cuh"
#include "scan_records.cuh"

namespace rtsparse {

#if defined(RT_SPARSE_STATS)
struct Stats { unsigned long long steps, cands, evals, tops, bcalls, shortcut, lazy, hops, acanfound, events, dense; };
inline Stats &stats() { static Stats s{}; return s; }
#endif
#if defined(RT_SPARSE_STATS) && !defined(__CUDA_ARCH__)
#define SP_STAT(f, n) (stats().f += (n))
#else
#define SP_STAT(f, n) ((void)0)
#endif

using rtfast::FastState; using rtfast::OFF_NONE; using rtfast::minmax; using rtfast::span_minmax;
using rtfast::row_time; using rtfast::volts;

enum { SP_DENSE = 0, SP_SPARSE = 1, SP_DONE = 2 };
constexpr uint32_t NO_ROW32 = 0xffffffffu;
#define SPARSE_SEARCH_WORDS 4          /* mask words examined per step while looking for the next candidate */
#define SPARSE_TRIES 1                 /* candidates a lane may reject with the integer pre-filter within one step (measured: retrying
                                          inside the step costs more in divergence than it saves in expensive-part utilisation) */

/* 32 mask bits starting at bit position p (any alignment); the arrays have slack words behind the last row */
RT_FHD uint32_t bits_at(const uint32_t *mk, uint64_t p) {
   const uint64_t wi = p >> 5; const int sh = (int)(p & 31);
   const uint32_t lo = mk[wi];
   if (sh == 0) return lo;
   return (lo >> sh) | (mk[wi + 1] << (32 - sh)); }
/* 64 mask bits starting at bit position p */
RT_FHD uint64_t bits64_at(const uint32_t *mk, uint64_t p) {
   const uint64_t wi = p >> 5; const int sh = (int)(p & 31);
   const uint32_t w0 = mk[wi], w1 = mk[wi + 1], w2 = mk[wi + 2];
   const uint32_t lo = sh ? (w0 >> sh) | (w1 << (32 - sh)) : w0, hi = sh ? (w1 >> sh) | (w2 << (32 - sh)) : w1;
   return (uint64_t)lo | ((uint64_t)hi << 32); }
RT_FHD int ctz32(uint32_t v) {
#ifdef __CUDA_ARCH__
   return __ffs((int)v) - 1;
#else
   return __builtin_ctz(v);
#endif
}
RT_FHD int clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
   return __clz((int)v);
#else
   return __builtin_clz(v);
#endif
}

/* REC: the candidate rows come as records (scan_records.cuh, phase B1) instead of being derived from the bit planes and the samples */
template <int STRIDE, class Emit, bool REC = false>
struct SparseScan {
   const DevCfg &c; const int16_t *plane; const uint32_t *mc, *mc_lo, *mc_hi, *ma; const uint32_t *gm;
   uint64_t row0; uint32_t end; int trk, w, delay; uint32_t io, o_pure;
   Emit em; FastState<STRIDE> t; uint32_t *ht;
   /* detector state: o = next row; tests are off for rows < resume (blind countdown, decoder.c:778); m = the lazy minimum,
      exact for row mq (dense mode: for row o - 1) */
   uint32_t o, resume, mq; int m, T, T0t, T1t, st; float inv_lsb, rise, reqmin;
   uint32_t ndense;                                             /* rows walked in dense mode (diagnostics) */
   uint64_t cb; uint32_t cb_o;                                  /* the 64 candidate bits of rows [cb_o, cb_o + 64) (cb_o = NO_ROW32: none held) */
   /* record mode: the cursor -- tile, index inside it, the tile's place in the record pool */
   const CandRec *recs; const uint32_t *tbase, *tcnt; uint32_t rt_tile, rt_idx, rt_cnt, rt_base, rt_last;
   /* proof data (offsets relative to row0; OFF_NONE = none), as in UnitScan */
   bool pre; uint32_t pre_pos, pre_end;                         /* pre_end: the row of the first event once it is known */
   int qmin, qmax, qthr, qL; int32_t ll, last_canon;
   int32_t sync_row, loud_at_sync, sync_first, sync_early, loud_early; bool early_frozen; uint32_t sf_from; uint64_t quiet_from;

   RT_FHD SparseScan(const DevCfg &c_, uint32_t *heights) : c(c_), w(c_.width), ht(heights) { st = SP_DONE; o = end = 0; }

   /* the sample the detector sees at stream offset o (deskew FIFO, decoder.c:819-831) */
   RT_FHD int sample(uint32_t oo) const { return (int)plane[row0 + (oo >= (uint32_t)delay ? oo - (uint32_t)delay : oo)]; }
   /* plane row of stream offset o, for o >= delay */
   RT_FHD uint64_t prow(uint32_t oo) const { return row0 + oo - (uint32_t)delay; }

   RT_FHD void thresholds() {                                   /* decoder.c:785-786; T: see UnitScan::thresholds */
      rise = c.p.pkww_rise * (t.avg_height / RT_PKWW_PEAKHEIGHT) / t.agc_gain;
      reqmin = c.p.min_peak * (t.avg_height / RT_PKWW_PEAKHEIGHT) / t.agc_gain;
      float q = rise * inv_lsb * 0.999f - 2.0f;
      T = !(q > 0) ? 0 : (q > 70000.0f ? 70000 : (int)q);
      /* follow the candidate plane with the highest threshold that T still covers */
      if (!REC) {
         const uint32_t *want = T1t > 0 && T >= T1t ? mc_hi : mc_lo;
         if (want != mc) { mc = want; cb_o = NO_ROW32; } } }

   /* ---- proof data: identical bookkeeping to UnitScan (scan_fast.cuh) ---- */
   RT_FHD void commit() {
      if (last_canon != OFF_NONE && last_canon != sync_row) {
         sync_row = last_canon; loud_at_sync = ll;
         if (!early_frozen) { sync_early = last_canon; loud_early = ll; } } }
   RT_FHD void loud_row(int32_t oo) {
      commit();
      if (sync_early != OFF_NONE) early_frozen = true;
      ll = oo;
      if (oo >= 0) sf_from = NO_ROW32; }
   RT_FHD void feed(int32_t oo, int raw) {
      if (raw < qmin) qmin = raw;
      if (raw > qmax) qmax = raw;
      if (qmax - qmin >= qthr) {
         const int64_t to = (int64_t)row0 + oo, from = to - qL + 1;
         const minmax r = span_minmax(plane, from < 0 ? 0 : from, to, raw);
         qmin = r.mn; qmax = r.mx;
         if (r.mx - r.mn >= qthr) loud_row(oo); } }
   RT_FHD void track(uint32_t oo, bool canonical) {
      feed((int32_t)oo, (int)plane[row0 + oo]);
      if (canonical && ll != (int32_t)oo) {
         last_canon = (int32_t)oo;
         if (oo >= sf_from) { sync_first = (int32_t)oo; sf_from = NO_ROW32; } } }
   /* rows [pre_pos, upto) of the sparse stretch: none of them had an event; canonical rows are the acan bits */
   RT_FHD void advance_pre(uint32_t upto) {
      while (pre_pos < upto) {
         if (gm && ((row0 + pre_pos) & (RT_GRAN - 1)) == 0 && pre_pos + RT_GRAN <= upto) {
            const uint32_t g = gm[(row0 + pre_pos) / RT_GRAN];
            const int gmn = (int)(int16_t)(uint16_t)(g & 0xffffu), gmx = (int)(int16_t)(uint16_t)(g >> 16);
            const int nmn = gmn < qmin ? gmn : qmin, nmx = gmx > qmax ? gmx : qmax;
            if (nmx - nmn < qthr) {                              /* no row of this granule can be loud */
               qmin = nmn; qmax = nmx;
               const uint32_t bits = bits_at(ma, prow(pre_pos));
               if (bits) {
                  if (sf_from != NO_ROW32) {
                     const uint32_t b2 = sf_from <= pre_pos ? bits : (sf_from - pre_pos < 32 ? bits >> (sf_from - pre_pos) << (sf_from - pre_pos) : 0u);
                     if (b2) { sync_first = (int32_t)(pre_pos + (uint32_t)ctz32(b2)); sf_from = NO_ROW32; } }
                  last_canon = (int32_t)(pre_pos + 31u - (uint32_t)clz32(bits)); }
               pre_pos += RT_GRAN; continue; } }
         const uint64_t p = prow(pre_pos);
         track(pre_pos, (ma[p >> 5] >> (p & 31)) & 1u);
         ++pre_pos; } }

   /* The quiet pre-scan of rows [-npre, -1] (UnitScan feeds them one by one).  Only two things survive it: the last loud row, and a
      (min, max) pair that covers at least the last qL rows.  A unit starts right behind the previous block, so walking the rows
      forwards means an exact span for nearly every one of them; instead the rows are visited BACKWARDS until the first loud one:
      for a row whose whole span [row - qL + 1, row] has been fed, "loud" is exactly "that span >= qthr" (quiet.cuh), whatever the
      tracker's running pair was, and a pair over a SUPERSET of the tracker's rows decides every later row the same way.  Whole
      quiet granules are skipped with the granule map.  Only if no such row is loud are the first qL - 1 rows -- whose verdict
      depends on where feeding started -- fed the original way. */
   RT_FHD void prescan(int32_t npre) {
      const int32_t pure_lo = -npre + qL - 1;
      int smn = 32767, smx = -32768;
      int32_t oo = -1;
      while (oo >= pure_lo) {
         const int64_t to = (int64_t)row0 + oo;
         if (gm && (to & (RT_GRAN - 1)) == RT_GRAN - 1 && oo - (RT_GRAN - 1) >= pure_lo) {
            int64_t gfirst = (to - (RT_GRAN - 1) - qL + 1); gfirst = gfirst < 0 ? 0 : gfirst / RT_GRAN;
            const int64_t glast = to / RT_GRAN;
            int gmn = 32767, gmx = -32768, lmn = 0, lmx = 0;
            for (int64_t g = gfirst; g <= glast; ++g) {
               const uint32_t v = gm[g];
               lmn = (int)(int16_t)(uint16_t)(v & 0xffffu); lmx = (int)(int16_t)(uint16_t)(v >> 16);
               if (lmn < gmn) gmn = lmn; if (lmx > gmx) gmx = lmx; }
            if (gmx - gmn < qthr) {                              /* no row of granule glast can be loud */
               if (lmn < smn) smn = lmn; if (lmx > smx) smx = lmx;
               oo -= RT_GRAN; continue; } }
         const int64_t from = to - qL + 1;
         const int x = (int)plane[to];
         const minmax r = span_minmax(plane, from < 0 ? 0 : from, to, x);
         if (r.mx - r.mn >= qthr) {
            ll = oo; qmin = r.mn < smn ? r.mn : smn; qmax = r.mx > smx ? r.mx : smx;
            return; }
         if (x < smn) smn = x; if (x > smx) smx = x;
         --oo; }
      for (int32_t q = -npre; q < pure_lo && q < 0; ++q) feed(q, (int)plane[(int64_t)row0 + q]);
      if (smn < qmin) qmin = smn; if (smx > qmax) qmax = smx; }

   /* The tail rule's proof datum (TrkMeta::quiet_tail_from): the earliest row q >= row0 + lo such that no row of [q, row_end) is
      loud, with loud(j) := the raw samples of rows [j - qL + 1, j] span >= qthr (the definition itself, independent of where any
      tracker started).  Walks backwards from the unit's end, whole quiet granules at a time, and stops at the first loud row. */
   RT_FHD uint64_t quiet_tail(uint32_t lo) const {
      int64_t oo = (int64_t)end - 1;
      while (oo >= (int64_t)lo) {
         const int64_t to = (int64_t)row0 + oo;
         if (gm && (to & (RT_GRAN - 1)) == RT_GRAN - 1 && oo - (RT_GRAN - 1) >= (int64_t)lo) {
            int64_t gfirst = (to - (RT_GRAN - 1) - qL + 1); gfirst = gfirst < 0 ? 0 : gfirst / RT_GRAN;
            const int64_t glast = to / RT_GRAN;
            int gmn = 32767, gmx = -32768;
            for (int64_t g = gfirst; g <= glast; ++g) {
               const uint32_t v = gm[g];
               const int lmn = (int)(int16_t)(uint16_t)(v & 0xffffu), lmx = (int)(int16_t)(uint16_t)(v >> 16);
               if (lmn < gmn) gmn = lmn; if (lmx > gmx) gmx = lmx; }
            if (gmx - gmn < qthr) { oo -= RT_GRAN; continue; } }      /* no row of granule glast can be loud */
         const int64_t from = to - qL + 1;
         const minmax r = span_minmax(plane, from < 0 ? 0 : from, to, (int)plane[to]);
         if (r.mx - r.mn >= qthr) return row0 + (uint64_t)oo + 1u;
         --oo; }
      return row0 + lo; }

   /* process_*_transition, decoder.c:560-609 */
   RT_FHD void transition(bool top, uint32_t oo) {
      const double t_ev = top ? t.t_top : t.t_bot;
      const float v_top_seen = t.v_top, v_bot_seen = t.v_bot;
      ++t.peakcount;
      if (c.density) { /* doing_density_detection: the mode handlers are bypassed (decoder.c:578), AGC and average height stay put */ }
      else if (c.mode == RT_MODE_NRZI) rtfb::nrzi_feedback(c, t, top);
      else if (c.mode == RT_MODE_PE) rtfb::pe_feedback(c, t, top, t_ev);
      else rtfb::agc_adjust(c, t);
      if (top) t.v_lasttop = t.v_top; else t.v_lastbot = t.v_bot;
      t.t_lastpeak = t_ev;
      em.emit(row0 + oo, t_ev, v_top_seen, v_bot_seen, t.agc_gain, top);
      thresholds(); }

   /* the event at row oo: peak value `val` (int16 domain), float value v; the peak is the sample at window position
      `pos` (stream offset / plane row arithmetic is the caller's), ld = its 1-based distance from the left edge, prev / next
      its neighbours.  refine_peak, decoder.c:700-749, one code path for both polarities (see UnitScan::refine). */
   RT_FHD void fire(bool top, float v, int ld, bool found, int xprev, int xnext, uint32_t oo) {
      if (pre) { pre = false; pre_end = oo; }                    /* the proof data of rows < oo is completed in finish() */
      if (!found || ld >= w || ld <= 1) { t.failed = 2; resume = oo + 1; return; }        /* the reference would fatal() */
      const float sg = top ? 1.0f : -1.0f;
      const float vprev = sg * volts(c, xprev), vnext = sg * volts(c, xnext);
      const float edge = sg * v - RT_PEAK_THRESHOLD / t.agc_gain;
      float adj = 0;
      if (vprev > edge && vnext < edge) adj = -0.5f;
      else if (vnext > edge && vprev < edge) adj = +0.5f;
      const double timenow = row_time(c, row0 + oo);
      const double tp = timenow - (double)(((float)(w - ld) - adj) * c.sample_deltat);
      resume = oo + (uint32_t)ld + 1u;                           /* pkww_countdown = left_distance */
      if (top) { t.v_top = v; t.t_top = tp; } else { t.v_bot = v; t.t_bot = tp; }
      transition(top, oo); }

   /* start the scan of unit rows [row0_, row_end) of track trk_ from a fresh RT_RESET_FULL */
   RT_FHD void begin(const int16_t *plane_, uint64_t row0_, uint64_t row_end, int trk_, Emit em_, int quiet_thr_lsb) {
      plane = plane_; row0 = row0_; end = (uint32_t)(row_end - row0_); trk = trk_; delay = c.skew[trk_]; em = em_;
      mc_lo = c.m_cand + (size_t)trk_ * c.mask_stride; mc_hi = c.m_cand2 + (size_t)trk_ * c.mask_stride; mc = mc_lo;
      ma = c.m_acan + (size_t)trk_ * c.mask_stride; T1t = c.T1[trk_]; cb = 0; cb_o = NO_ROW32;
      gm = c.gmm ? c.gmm + (size_t)trk_ * c.ngran_cap : nullptr; T0t = c.T0[trk_];
      if (REC) {
         recs = c.recs; tbase = c.rec_tile_base + (size_t)trk_ * c.rec_tiles; tcnt = c.rec_tile_cnt + (size_t)trk_ * c.rec_tiles;
         rt_tile = NO_ROW32; rt_idx = rt_cnt = rt_base = 0;
         rt_last = (uint32_t)((row_end - 1) / RT_REC_TILE); }
      const bool tz = row_time(c, row0) == 0.0;
      io = (uint32_t)trk + (tz ? 1u : 0u);
      o_pure = ((uint32_t)delay > io ? (uint32_t)delay : io) + (uint32_t)w;
      t.t_top = t.t_bot = t.t_lastpeak = 0; t.v_top = t.v_bot = t.v_lasttop = t.v_lastbot = 0; t.avg_height_sum = 0;
      t.avg_height_count = t.heightndx = t.peakcount = 0; t.datablock = t.bit1_up = t.failed = 0;
      t.heights.p = ht;
      for (int i = 0; i < RT_AGC_MAX_WINDOW; ++i) t.heights[i] = 0.0f;
      t.agc_gain = 1.0f; t.avg_height = RT_PKWW_PEAKHEIGHT;
      t.t_clkwindow = c.clk_init / 2 * c.p.clk_factor;
      inv_lsb = 32767.0f / c.maxvolts;
      thresholds(); resume = 0; m = 0; mq = 0; ndense = 0; cb = 0; cb_o = NO_ROW32;
      qthr = quiet_thr_lsb; qL = w + delay; qmin = 32767; qmax = -32768; ll = last_canon = OFF_NONE;
      sync_row = loud_at_sync = sync_first = sync_early = loud_early = OFF_NONE; early_frozen = false; pre = true; pre_end = 0;
      const int lead = (int)io > delay ? (int)io : delay;
      sf_from = (uint32_t)(lead + w + 1);
      const int32_t npre = row0 > (uint64_t)c.prescan_rows ? c.prescan_rows : (int32_t)row0;
      prescan(npre);
      quiet_from = ll == OFF_NONE ? row0 - (uint64_t)npre : (uint64_t)((int64_t)row0 + ll + 1);
      for (uint32_t oo = 0; oo <= io && oo < end; ++oo) track(oo, false);              /* decoder.c:855-861: not looked at yet */
      o = io + 1; pre_pos = o;
      st = o >= end ? SP_DONE : SP_DENSE;
      if (st == SP_DENSE) { m = sample(io); t.t_lastpeak = row_time(c, row0 + io); } }

   /* ---- dense mode: one row, by definition ---- */
   RT_FHD void dense_step() {
      const bool had_w = o >= io + (uint32_t)w;                 /* the window was full: its oldest sample leaves (decoder.c:755) */
      const uint32_t ws = had_w ? o - (uint32_t)w + 1u : io;
      int S = -32768, mn = 32767;
      for (uint32_t i = ws; i <= o; ++i) { const int v = sample(i); if (v > S) S = v; if (v < mn) mn = v; }
      const int lv = had_w ? sample(o - (uint32_t)w) : 0;        /* decoder.c:754: old_left stays 0 until the window is full */
      const bool A = had_w ? lv >= S : S == 0;
      if (A || lv == m) m = mn;                                  /* the rescan of decoder.c:767-775 */
      const bool canonical = had_w && A;
      ++ndense; SP_STAT(dense, 1);
      bool fired = false;
      if (o >= resume) {
         const int xl = sample(ws), xr = sample(o);
         const float vl = volts(c, xl), vr = volts(c, xr), maxv = volts(c, S), minv = volts(c, m);
         const bool top = maxv > vl + rise && maxv > vr + rise && (reqmin == 0 || maxv > reqmin);
         const bool bot = !top && minv < vl - rise && minv < vr - rise && (reqmin == 0 || minv < -reqmin);
         if (top || bot) {
            const int val = top ? S : m;
            uint32_t pos = NO_ROW32;
            for (uint32_t i = o + 1; i-- > ws;) if (sample(i) == val) pos = i;         /* leftmost match */
            const bool found = pos != NO_ROW32;
            const int ld = found ? (int)(pos - ws) + 1 : 0;
            const int xprev = found && pos > ws ? sample(pos - 1) : 0, xnext = found && pos < o ? sample(pos + 1) : 0;
            fire(top, top ? maxv : minv, ld, found, xprev, xnext, o);
            fired = true; } }
      if (pre && !fired) track(o, canonical);
      const uint32_t cur = o;
      ++o; if (pre) pre_pos = o;
      if (o >= end) { st = SP_DONE; return; }
      /* from a canonical row of the pure regime on, the state is what the masks describe */
      if (canonical && cur >= o_pure && T >= T0t && T0t > 0) { st = SP_SPARSE; mq = cur; } }

   /* ---- sparse mode ---- */
   /* One pass over the w samples at win[0 .. w): packed keys carry value and position through a single min / max,
         kmax = max_i ((v_i + 32768) << 6 | 63 - i)  ->  maximum and its LEFTMOST position
         kmin = min_i ((v_i + 32768) << 6 | i)       ->  minimum and its LEFTMOST position
         keq  = min_i (v_i == val ? i : 64)          ->  leftmost position of a sample equal to val (64: none)       (EQ only)
      (w <= 50 < 64).  Windows of up to 16 samples -- every built-in parameter set at the usual 10..20 samples per bit -- are
      read with 16 independent predicated loads, so that the lane waits for memory once, not once per sample. */
   struct WinKeys { uint32_t kmax, kmin, keq; };
   static RT_FHD int key_val(uint32_t k) { return (int)(k >> 6) - 32768; }
   static RT_FHD int kmax_pos(uint32_t k) { return 63 - (int)(k & 63u); }
   static RT_FHD int kmin_pos(uint32_t k) { return (int)(k & 63u); }
   template <bool MAXK, bool EQ>
   RT_FHD WinKeys scan_window(const int16_t *win, int val) const {
      WinKeys r{0u, 0xffffffffu, 64u};
      if (w <= 16) {
         int v[16];
#pragma unroll
         for (int i = 0; i < 16; ++i) v[i] = i < w ? (int)win[i] : 0;
#pragma unroll
         for (int i = 0; i < 16; ++i) {
            if (i < w) {
               const uint32_t kb = ((uint32_t)(v[i] + 32768) << 6) + (uint32_t)i;
               if (kb < r.kmin) r.kmin = kb;
               if (MAXK) { const uint32_t kt = kb + (uint32_t)(63 - 2 * i); if (kt > r.kmax) r.kmax = kt; }
               if (EQ) { if (v[i] == val && (uint32_t)i < r.keq) r.keq = (uint32_t)i; } } } }
      else {
         for (int i = 0; i < w; ++i) {
            const int vi = (int)win[i];
            const uint32_t kb = ((uint32_t)(vi + 32768) << 6) + (uint32_t)i;
            if (kb < r.kmin) r.kmin = kb;
            if (MAXK) { const uint32_t kt = kb + (uint32_t)(63 - 2 * i); if (kt > r.kmax) r.kmax = kt; }
            if (EQ) { if (vi == val && (uint32_t)i < r.keq) r.keq = (uint32_t)i; } } }
      return r; }
   /* is there an acan row in plane rows [lo, hi]  (hi - lo < 64) */
   RT_FHD bool acan_in(uint64_t lo, uint64_t hi) const {
      const uint32_t n = (uint32_t)(hi - lo) + 1u;
      uint32_t b = bits_at(ma, lo);
      if (n < 32) return (b & ((1u << n) - 1u)) != 0;
      if (b) return true;
      b = bits_at(ma, lo + 32);
      return (n >= 64 ? b : (b & ((1u << (n - 32)) - 1u))) != 0; }
   /* The lazy minimum at stream offset oo (plane row p), from its last exactly known value (row mq).  m only changes at
      refresh rows: acan rows, and rows where the sample that leaves equals m (decoder.c:767).  After a refresh at row r the value
      is Wmin(r) and the next refresh of the second kind happens when the LEFTMOST sample with that value leaves, at row
      (its position) + w -- so the walk hops from refresh to refresh instead of visiting every row. */
   /* hint: the acan bits of the rows of the window that ends at oo, if the caller holds them (hint_rows = w, else 0): the last
      acan row is then found without touching the bit plane again, unless it lies in front of the window */
   /* returns the plane row of the LEFTMOST sample of the window that ends at oo which carries m (~0: not determined): the carrier
      found by the walk has not left yet (else there would have been another refresh), and no equal sample lies left of it */
   RT_FHD uint64_t lazy_min(uint32_t oo, uint32_t hint_bits = 0, int hint_rows = 0) {
      if (mq == oo) return ~0ull;
      const uint64_t p = prow(oo), pm = prow(mq);
      /* last acan row in (pm, p] */
      uint64_t a = pm; bool have = false;
      uint64_t top = p;                                           /* rows above `top` are known to hold no acan row */
      bool search = true;
      if (hint_rows) {
         const uint64_t ws = p - (uint32_t)hint_rows + 1u;
         const uint32_t hb = hint_rows >= 32 ? hint_bits : hint_bits & ((1u << hint_rows) - 1u);
         if (hb) { const uint64_t ah = ws + 31u - (uint32_t)clz32(hb); search = false; if (ah > pm) { a = ah; have = true; } }
         else if (ws <= pm + 1) search = false;                   /* the window covers (pm, p]: none */
         else top = ws - 1; }
      if (search) {
         uint64_t wi = top >> 5;
         uint32_t bits = ma[wi] & (0xffffffffu >> (31 - (int)(top & 31)));
         for (;;) {
            const uint64_t base = wi << 5;
            if (base + 31 <= pm) break;                                                   /* the whole word is at or before pm */
            if (base <= pm) bits &= (pm - base) >= 31 ? 0u : (0xffffffffu << ((int)(pm - base) + 1));
            if (bits) { a = base + 31u - (uint32_t)clz32(bits); have = true; break; }
            if (base <= pm || wi == 0) break;
            --wi; bits = ma[wi]; } }
      SP_STAT(lazy, 1); SP_STAT(acanfound, have ? 1 : 0);
      /* one loop for both starts -- from the acan row a (m = its window minimum) or from row pm (m as it is: find the sample that
         carries it) -- and for the hops, so that the lanes of a warp that need this run the same code */
      uint64_t r = have ? a : pm;
      bool keep = !have;
      uint64_t at;
      for (;;) {
         const uint64_t ws = r - (uint32_t)w + 1u;
         const WinKeys k = scan_window<false, true>(plane + ws, m);
         if (keep && k.keq < 64u) at = ws + k.keq;
         else { m = key_val(k.kmin); at = ws + (uint32_t)kmin_pos(k.kmin); }
         keep = false;
         if (at + (uint32_t)w > p) break;
         r = at + (uint32_t)w; SP_STAT(hops, 1); }
      mq = oo;
      return at; }

   RT_FHD void sparse_step() {
      /* The cheap part -- next candidate row at or after the end of the blind stretch, its window, the integer pre-filter for the
         CURRENT threshold bound T (the top test can only pass if S - max(l, r) >= T, the bottom test only if min(l, r) - m >= T, and
         m >= Wmin) -- is repeated up to SPARSE_TRIES times, so that when the warp goes on to the expensive part most of its lanes
         bring a row that has a real chance of being an event. */
      uint64_t p = 0; uint32_t oc = 0; const int16_t *win = plane;
      int S = 0, mn = 0, posm = 0, pos = 0, xl = 0, xr = 0; bool tcand = false, bcand = false; uint32_t abits = 0;
      const uint64_t pend = prow(end);
#pragma unroll 1
      for (int tries = 0; tries < SPARSE_TRIES; ++tries) {
         const uint32_t from = o > resume ? o : resume;
         if (from >= end) { o = end; st = SP_DONE; return; }
         p = prow(from);
         /* the candidate bits of 64 rows stay in registers from step to step: successive candidates mostly share them, and the second
            half has arrived long before it is needed */
         uint64_t bits; uint32_t nv = 64;
         if (cb_o != NO_ROW32 && from >= cb_o && from - cb_o < 64u) { bits = cb >> (from - cb_o); nv = 64u - (from - cb_o); }
         else { bits = bits64_at(mc, p); cb = bits; cb_o = from; }
         int nw = 1;
         while (!bits && nw < SPARSE_SEARCH_WORDS / 2 && p + nv < pend) {
            p += nv; nv = 64; bits = bits64_at(mc, p); cb = bits; cb_o = (uint32_t)(p - row0) + (uint32_t)delay; ++nw; }
         if (!bits) { p += nv; o = p >= pend ? end : (uint32_t)(p - row0) + (uint32_t)delay; if (o >= end) st = SP_DONE; return; }
         { const uint32_t blo = (uint32_t)bits; p += blo ? (uint32_t)ctz32(blo) : 32u + (uint32_t)ctz32((uint32_t)(bits >> 32)); }
         if (p >= pend) { o = end; st = SP_DONE; return; }
         oc = (uint32_t)(p - row0) + (uint32_t)delay;
         win = plane + (p - (uint32_t)w + 1u);
         /* acan bits of the window's rows: the two words are requested together with the samples and only combined afterwards (an
            instruction that needs them in front of the sample loads would hold those back) */
         const uint64_t aw = (p - (uint32_t)w + 1u) >> 5; const int ash = (int)((p - (uint32_t)w + 1u) & 31);
         uint32_t a_lo = 0, a_hi = 0;
         if (w <= 32) { a_lo = ma[aw]; a_hi = ma[aw + 1]; }
         xr = win[w - 1];
         const WinKeys wk = scan_window<true, false>(win, 0);
         abits = ash ? (a_lo >> ash) | (a_hi << (32 - ash)) : a_lo;
         S = key_val(wk.kmax); mn = key_val(wk.kmin); posm = kmin_pos(wk.kmin); pos = kmax_pos(wk.kmax);
         xl = win[0];
         tcand = S - (xl > xr ? xl : xr) >= T; bcand = (xl < xr ? xl : xr) - mn >= T;
         o = oc + 1;
         SP_STAT(cands, 1);
         if (tcand || bcand) break;
         if (o >= end) { st = SP_DONE; return; } }
      if (!tcand && !bcand) return;
      SP_STAT(evals, 1);
      const float vl = volts(c, xl), vr = volts(c, xr), maxv = volts(c, S);
      const bool top = tcand && maxv > vl + rise && maxv > vr + rise && (reqmin == 0 || maxv > reqmin);
      bool bot = false; float minv = 0;
      if (!top && bcand) {
         /* a refresh at or after the row at which the window's (leftmost) minimum entered makes m that minimum */
         SP_STAT(bcalls, 1);
         const bool fresh = w <= 32 ? ((abits >> posm) & (w >= 32 ? 0xffffffffu : ((1u << (w - posm)) - 1u))) != 0
                                    : acan_in(p - (uint32_t)w + 1u + (uint32_t)posm, p);
         uint64_t m_at = p - (uint32_t)w + 1u + (uint32_t)posm;
         if (fresh) { m = mn; mq = oc; SP_STAT(shortcut, 1); }
         else m_at = lazy_min(oc, abits, w <= 32 ? w : 0);
         minv = volts(c, m);
         bot = minv < vl - rise && minv < vr - rise && (reqmin == 0 || minv < -reqmin);
         if (bot) {
            pos = m_at != ~0ull ? (int)(m_at - (p - (uint32_t)w + 1u)) : -1;
            if (pos < 0 || pos >= w || (int)win[pos] != m) { pos = -1; for (int i = w; i-- > 0;) if ((int)win[i] == m) pos = i; } } }   /* (never taken in practice) */
      if (top || bot) {
         SP_STAT(events, 1); SP_STAT(tops, top ? 1 : 0);
         const bool found = pos >= 0;
         const int xprev = found && pos > 0 ? win[pos - 1] : 0, xnext = found && pos < w - 1 ? win[pos + 1] : 0;
         fire(top, top ? maxv : minv, pos + 1, found, xprev, xnext, oc);
         if (T < T0t) { lazy_min(oc); st = SP_DENSE; } }           /* the masks no longer cover the threshold: walk rows */
      if (o >= end) st = SP_DONE; }

   /* ---- sparse mode on records ---- */
   RT_FHD void rec_tile(uint32_t tile) { rt_tile = tile; rt_base = tbase[tile]; rt_cnt = tcnt[tile]; rt_idx = 0; }
   /* the cursor at the first record with plane row >= p; false: none before the end of the unit */
   RT_FHD bool rec_at(uint64_t p) {
      const uint32_t tile = (uint32_t)(p / RT_REC_TILE);
      if (rt_tile == NO_ROW32 || tile != rt_tile) { if (tile > rt_last) return false; rec_tile(tile); }
      for (;;) {
         while (rt_idx < rt_cnt && (uint64_t)recs[rt_base + rt_idx].row < p) ++rt_idx;
         if (rt_idx < rt_cnt) return true;
         if (rt_tile >= rt_last) return false;
         rec_tile(rt_tile + 1); } }

   RT_FHD void sparse_step_rec() {
      const uint32_t from = o > resume ? o : resume;
      if (from >= end) { o = end; st = SP_DONE; return; }
      const uint64_t pend = prow(end);
      if (!rec_at(prow(from))) { o = end; st = SP_DONE; return; }
      const CandRec r = recs[rt_base + rt_idx];
      if ((uint64_t)r.row >= pend) { o = end; st = SP_DONE; return; }
      ++rt_idx;
      const uint32_t oc = (uint32_t)((uint64_t)r.row - row0) + (uint32_t)delay;
      o = oc + 1;
      SP_STAT(cands, 1);
      const int xl = r.xl, xr = r.xr, S = r.S;
      const bool tcand = S - (xl > xr ? xl : xr) >= T;
      bool bcand = (xl < xr ? xl : xr) - (int)r.m >= T;
      if (r.posm != RT_REC_NOPOS) { m = r.m; mq = oc; }          /* the lazy minimum is known exactly here */
      if (tcand || bcand) {
         SP_STAT(evals, 1);
         const float vl = volts(c, xl), vr = volts(c, xr), maxv = volts(c, S);
         const bool top = tcand && maxv > vl + rise && maxv > vr + rise && (reqmin == 0 || maxv > reqmin);
         bool bot = false; float minv = 0; int posb = -1, bprev = 0, bnext = 0;
         if (!top && bcand) {
            if (r.posm != RT_REC_NOPOS) { posb = r.posm; bprev = r.mprev; bnext = r.mnext; }
            else {                                                 /* no canonical row within the record's reach: walk from the last known value */
               const uint64_t m_at = lazy_min(oc);
               const uint64_t ws = (uint64_t)r.row - (uint32_t)w + 1u;
               if (m_at != ~0ull && m_at >= ws && m_at <= (uint64_t)r.row) posb = (int)(m_at - ws);
               else { const int16_t *win = plane + ws; for (int i = w; i-- > 0;) if ((int)win[i] == m) posb = i; }
               if (posb >= 0) { const int16_t *win = plane + ws; bprev = posb > 0 ? win[posb - 1] : 0; bnext = posb < w - 1 ? win[posb + 1] : 0; }
               bcand = (xl < xr ? xl : xr) - m >= T; }
            minv = volts(c, m);
            bot = bcand && minv < vl - rise && minv < vr - rise && (reqmin == 0 || minv < -reqmin); }
         if (top) fire(true, maxv, (int)r.posS + 1, true, r.sprev, r.snext, oc);
         else if (bot) fire(false, minv, posb + 1, posb >= 0, bprev, bnext, oc);
         if ((top || bot) && T < T0t) { if (mq != oc) lazy_min(oc); st = SP_DENSE; } }   /* the records no longer cover the threshold: walk rows */
      if (o >= end) st = SP_DONE; }

   RT_FHD void step() { if (st == SP_DENSE) dense_step(); else if (st == SP_SPARSE) { if (REC) sparse_step_rec(); else sparse_step(); } }

   RT_FHD void finish(TrkMeta &meta) {
      /* rows before the first event (or the whole unit) that the sparse mode has not tracked yet: done here, for all lanes of
         the warp together (dense mode keeps pre_pos == o: nothing is left then) */
      advance_pre(pre ? end : pre_end); commit();
      meta.first_event_row = em.first_row;
      meta.sync_row = sync_row == OFF_NONE ? RT_NOROW : row0 + (uint64_t)sync_row;
      meta.last_loud_row = loud_at_sync == OFF_NONE ? RT_NOROW : (uint64_t)((int64_t)row0 + loud_at_sync);
      meta.sync_first = sync_first == OFF_NONE ? RT_NOROW : row0 + (uint64_t)sync_first;
      meta.quiet_from = quiet_from;
      meta.sync_early = sync_early == OFF_NONE ? RT_NOROW : row0 + (uint64_t)sync_early;
      meta.loud_early = loud_early == OFF_NONE ? RT_NOROW : (uint64_t)((int64_t)row0 + loud_early);
      meta.last_event_row = em.last_row;
      /* only the unit that ends the tape needs it: behind every other unit the next one's quiet pre-scan examines the same rows */
      meta.quiet_tail_from = qthr > 0 && row0 + end >= c.nrows ? quiet_tail(end > (uint32_t)c.prescan_rows ? end - (uint32_t)c.prescan_rows : 0u) : RT_NOROW;
      meta.first_chunk = em.first_chunk; meta.nevents = em.n; meta.failed = t.failed; meta.pad = end > ndense ? end - ndense : 0; } };   /* pad: rows not walked (diagnostics) */

/* Drive one lane (host) or the 32 lanes of a warp (device) through (unit, track) jobs.  `Jobs` provides
 *   bool next(Scan&)   start the lane's next job, false if there is none (called by all lanes of the warp together)
 *   bool exhausted     set by next() when the job list has run out (the same for all lanes)
 *   void done(Scan&)   the lane's job is finished (store its TrkMeta)
 * `any(pred)` is the warp vote (identity on the host).  The warp takes 32 consecutive jobs (the tracks of 3-4 neighbouring units)
 * at a time, so that the per-job work -- quiet pre-scan, window fill, the walk to the first event, proof data -- runs for all lanes
 * together instead of one lane at a time while the other 31 wait. */
template <class Scan, class Jobs, class Vote>
RT_FHD void drive_sparse(Scan &us, Jobs &jobs, Vote any) {
   for (;;) {
      const bool have = jobs.next(us);
      if (jobs.exhausted) return;
      while (any(have && us.st != SP_DONE)) {
         if (have && us.st != SP_DONE) {
#pragma unroll 1
            for (int k = 0; k < 4 && us.st != SP_DONE; ++k) us.step(); } }
      if (have) jobs.done(us); } }

}  // namespace rtsparse
