/* readtape_b200/csrc/scan_generic.cuh -- the exact, fully general per-track scan (device code).
 *
 * One CUDA thread owns one track and walks its sample plane row by row, carrying the complete
 * per-track state (struct TrkState).  This is the always-correct formulation: every detector
 * (moving-window peaks, zero crossings, differentiated zero crossings), every mode's feedback
 * fragment, deskew, invert, differentiate, persistent state (Whirlwind) and all reset kinds.
 * The speculative whole-tape scan uses the same code for (unit, track) threads; the integer
 * fast path in scan_fast.cuh is an optimisation of the NRZI/PE/GCR common cases and is tested
 * against this one.
 *
 * Reference semantics followed (file:line in /root/reference/src), type-for-type:
 *   sample conversion / invert / differentiate   readtape.c:1418-1422, 1383-1388
 *   deskew FIFO                                  decoder.c:819-831
 *   first-sample init                            decoder.c:855-861
 *   lookfor_peak / refine_peak                   decoder.c:751-810 / 700-749
 *   lookfor_zerocrossing / differentiated        decoder.c:617-649 / 654-683
 *   process_*_transition glue                    decoder.c:560-609
 *   adjust_agc, adjust_clock, force_clock        decoder.c:500-558
 *   NRZI / PE / GCR / WW feedback fragments      decode_nrzi.c:184-230, decode_pe.c:127-201,
 *                                                decode_gcr.c:731-865, decode_ww.c:167-191
 *   GCR idle test                                decoder.c:879-882
 * Compiled with -fmad=false: no FMA contraction, IEEE division, float/double exactly as written
 * in the reference (built there with FLT_EVAL_METHOD 0).
 */
#pragma once
#include "rt_dev.h"
#include "feedback.cuh"
#include "quiet.cuh"

/* host + device: tests/host_fast builds this file for the CPU too (the exact scan and its skip-ahead are fuzzed there against the oracle) */
#define RT_GEN __host__ __device__

namespace rtgen {

RT_GEN __forceinline__ int gen_clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
   return __clz((int)v);
#else
   return v ? __builtin_clz(v) : 32;
#endif
}
RT_GEN __forceinline__ int gen_ffs32(uint32_t v) {
#ifdef __CUDA_ARCH__
   return __ffs((int)v);
#else
   return __builtin_ffs((int)v);
#endif
}

RT_GEN __forceinline__ double row_time(const DevCfg &c, uint64_t row) {
   long long ns = (long long)(c.tstart_ns + row * c.tdelta_ns);
   return (double)ns / 1e9; }

/* lazily evaluated timenow of the current row */
struct RowClock {
   const DevCfg &c; uint64_t row; double t; bool have;
   RT_GEN RowClock(const DevCfg &c_, uint64_t r) : c(c_), row(r), t(0), have(false) {}
   RT_GEN __forceinline__ double now() { if (!have) { t = row_time(c, row); have = true; } return t; } };

/* ---- clock averaging, decoder.c:533-558 ---------------------------------------------------- */
RT_GEN inline void clk_adjust(const DevCfg &c, TrkState &t, float delta) {
   int win = c.p.clk_window; float alpha = c.p.clk_alpha;
   if (win > 0) {
      float old = t.clk_spacing[t.clk_ndx];
      t.clk_spacing[t.clk_ndx] = delta;
      if (++t.clk_ndx >= win) t.clk_ndx = 0;
      t.clk_avg += (delta - old) / win; }
   else if (alpha > 0) t.clk_avg = alpha * delta + (1 - alpha) * t.clk_avg;
   else t.clk_avg = (c.mode & (RT_MODE_PE + RT_MODE_WW)) ? 1 / (c.bpi * c.ips) : 0.0f; }

RT_GEN inline void clk_force(TrkState &t, float v) {
   for (int i = 0; i < RT_CLKRATE_WINDOW; ++i) t.clk_spacing[i] = v;
   t.clk_avg = v; }

/* ---- AGC and the NRZI / PE feedback fragments: shared with the fast path (feedback.cuh) ------ */
using rtfb::agc_adjust;
using rtfb::baseline_accumulate;
using rtfb::nrzi_feedback;
using rtfb::pe_feedback;

RT_GEN inline void gcr_addbit(TrkState &t, int bit) {
   t.datablock = 1;
   if (t.datacount < RT_MAXBLOCK) { t.bit_m2 = t.bit_m1; t.bit_m1 = (uint8_t)bit; ++t.datacount; }
   t.lastbits = (uint8_t)((t.lastbits << 1) | bit);
   if (t.datacount % 5 == 0) {
      if ((t.lastbits & 0x1f) == RT_GCR_MARK2) t.resync_bitcount = 1;
      if ((t.lastbits & 0x1f) == RT_GCR_MARK1 && t.resync_bitcount > 0) t.resync_bitcount = 0; }
   if (t.resync_bitcount > 0) {
      if (t.resync_bitcount == 5) clk_force(t, t.t_peakdelta);
      ++t.resync_bitcount; } }

RT_GEN inline void gcr_feedback(const DevCfg &c, TrkState &t, bool top, double t_ev) {
   float delta = (float)(t_ev - t.t_lastpeak);
   int numbits = 1;
   if (t.datablock) {
      t.t_peakdeltaprev = t.t_peakdelta;
      t.t_peakdelta = delta;
      if (delta - t.t_pulse_adj > c.p.z1pt * t.clk_avg) {
         ++numbits; gcr_addbit(t, 0);
         if (delta - t.t_pulse_adj > c.p.z2pt * t.clk_avg) { ++numbits; gcr_addbit(t, 0); } }
      if (t.datacount > 3 && numbits == 1 && t.bit_m2) clk_adjust(c, t, t.t_peakdeltaprev);
      t.t_pulse_adj = c.p.pulse_adj * (numbits * t.clk_avg - delta); }
   gcr_addbit(t, 1);
   nrzi_feedback(c, t, top); }

/* ---- per-event glue, decoder.c:560-609 ------------------------------------------------------ */
template <class Emit>
RT_GEN inline void transition(const DevCfg &c, TrkState &t, bool top, uint64_t row, Emit &em) {
   double t_ev = top ? t.t_top : t.t_bot;
   float v_top_seen = t.v_top, v_bot_seen = t.v_bot;
   ++t.peakcount;
   if (!c.density) {
      if (c.mode == RT_MODE_NRZI) nrzi_feedback(c, t, top);
      else if (c.mode == RT_MODE_PE) pe_feedback(c, t, top, t_ev);
      else if (c.mode == RT_MODE_GCR) gcr_feedback(c, t, top, t_ev);
      else agc_adjust(c, t); }
   if (top) t.v_lasttop = t.v_top; else t.v_lastbot = t.v_bot;
   t.t_lastpeak = t_ev;
   em.emit(row, t_ev, v_top_seen, v_bot_seen, t.agc_gain, top); }

/* ---- moving-window peak detector, decoder.c:700-810 ------------------------------------------ */
RT_GEN inline double refine(const DevCfg &c, TrkState &t, float val, bool top, double timenow) {
   const int w = c.width;
   int left_distance = 1, prev = -1;
   float adj = 0;
   for (int ndx = t.left;;) {
      if (t.win[ndx] == val) {
         if (left_distance >= w || prev == -1) { t.failed = 2; return 0; }
         int next = ndx + 1; if (next >= w) next = 0;
         if (top) {
            float edge = val - RT_PEAK_THRESHOLD / t.agc_gain;
            if (t.win[prev] > edge && t.win[next] < edge) adj = -0.5f;
            else if (t.win[next] > edge && t.win[prev] < edge) adj = +0.5f; }
         else {
            float edge = val + RT_PEAK_THRESHOLD / t.agc_gain;
            if (t.win[prev] < edge && t.win[next] > edge) adj = -0.5f;
            else if (t.win[next] < edge && t.win[prev] > edge) adj = +0.5f; }
         double time = timenow - (double)(((float)(w - left_distance) - adj) * c.sample_deltat);
         t.countdown = left_distance;
         return time; }
      ++left_distance;
      if (ndx == t.right) break;
      prev = ndx;
      if (++ndx >= w) ndx = 0; }
   t.failed = 2;
   return 0; }

/* `probe` (out): bit 0 = the window was full (a sample left it), bit 1 = the rescan was triggered by
 * the running maximum leaving a full window -- the state-independent event at which two scans of
 * the same samples that were reset at different rows provably become identical. */
template <class Emit>
RT_GEN inline void peak_step(const DevCfg &c, TrkState &t, float v_now, RowClock &clk, Emit &em, unsigned &probe) {
   const int w = c.width;
   float leaving = 0;
   probe = 0;
   if (++t.right >= w) t.right = 0;
   if (t.right == t.left) {
      leaving = t.win[t.left];
      probe = 1;
      if (++t.left >= w) t.left = 0; }
   t.win[t.right] = v_now;
   if (v_now > t.maxv) t.maxv = v_now;
   if (probe && leaving == t.maxv) probe |= 2;
   if (leaving == t.maxv || leaving == t.minv) {      /* (Q1) the minimum is only refreshed here */
      float mx = -100, mn = +100;
      for (int ndx = t.left;;) {
         float v = t.win[ndx];
         if (v > mx) mx = v;
         if (v < mn) mn = v;
         if (ndx == t.right) break;
         if (++ndx >= w) ndx = 0; }
      t.maxv = mx; t.minv = mn; }
   if (t.countdown) { --t.countdown; return; }
   float rise = c.p.pkww_rise * (t.avg_height / RT_PKWW_PEAKHEIGHT) / t.agc_gain;
   float reqmin = c.p.min_peak * (t.avg_height / RT_PKWW_PEAKHEIGHT) / t.agc_gain;
   float vl = t.win[t.left], vr = t.win[t.right];
   if (t.maxv > vl + rise && t.maxv > vr + rise && (reqmin == 0 || t.maxv > reqmin)) {
      t.v_top = t.maxv;
      t.t_top = refine(c, t, t.maxv, true, clk.now());
      transition(c, t, true, clk.row, em); }
   else if (t.minv < vl - rise && t.minv < vr - rise && (reqmin == 0 || t.minv < -reqmin)) {
      t.v_bot = t.minv;
      t.t_bot = refine(c, t, t.minv, false, clk.now());
      transition(c, t, false, clk.row, em); } }

/* ---- zero-crossing detectors, decoder.c:617-683 ---------------------------------------------- */
template <class Emit>
RT_GEN inline void zc_step(const DevCfg &c, TrkState &t, float v_now, RowClock &clk, Emit &em) {
   if (v_now > 0) {
      t.dn_pending = 0;
      if (t.v_top < v_now) {
         t.v_top = v_now;
         if (t.up_pending && t.v_top > RT_ZEROCROSS_PEAK) {
            if (t.t_top == 0) t.t_top = clk.now();
            t.up_pending = 0;
            t.v_bot = 0;
            if (clk.now() - t.t_top <= (double)(t.clk_avg * RT_ZEROCROSS_SLOPE)) transition(c, t, true, clk.row, em); } }
      if (t.v_prev < 0 && t.v_bot < -RT_ZEROCROSS_PEAK) { t.t_top = clk.now(); t.up_pending = 1; } }
   else if (v_now < 0) {
      t.up_pending = 0;
      if (t.v_bot > v_now) {
         t.v_bot = v_now;
         if (t.dn_pending && t.v_bot < -RT_ZEROCROSS_PEAK) {
            if (t.t_bot == 0) t.t_bot = clk.now();
            t.dn_pending = 0;
            t.v_top = 0;
            if (clk.now() - t.t_bot <= (double)(t.clk_avg * RT_ZEROCROSS_SLOPE)) transition(c, t, false, clk.row, em); } }
      if (t.v_prev > 0 && t.v_top > RT_ZEROCROSS_PEAK) { t.t_bot = clk.now(); t.dn_pending = 1; } }
   t.v_prev = v_now; }

template <class Emit>
RT_GEN inline void dzc_step(const DevCfg &c, TrkState &t, float v_now, RowClock &clk, Emit &em) {
   if (v_now > 0) {
      if (t.v_top < v_now) t.v_top = v_now;
      if (t.up_pending) {
         t.t_top = t.t_firstzero > 0 ? (t.t_firstzero + t.t_lastzero) / 2 : clk.now() - (double)(c.sample_deltat / 2);
         t.up_pending = 0;
         t.t_firstzero = 0;
         transition(c, t, true, clk.row, em); }
      if (v_now > RT_ZEROCROSS_PEAK) { t.dn_pending = 1; t.t_firstzero = 0; t.v_bot = 0; } }
   else if (v_now < 0) {
      if (t.v_bot > v_now) t.v_bot = v_now;
      if (t.dn_pending) {
         t.t_bot = t.t_firstzero > 0 ? (t.t_firstzero + t.t_lastzero) / 2 : clk.now() - (double)(c.sample_deltat / 2);
         t.dn_pending = 0;
         t.t_firstzero = 0;
         transition(c, t, false, clk.row, em); }
      if (v_now < -RT_ZEROCROSS_PEAK) { t.up_pending = 1; t.t_firstzero = 0; t.v_top = 0; } }
   else {
      t.t_lastzero = clk.now();
      if (t.t_firstzero == 0) t.t_firstzero = clk.now(); } }

/* ---- resets (decoder.c:413-455, decode_ww.c:33-49) ------------------------------------------ */
/* The skew FIFO / differentiator state of the generic path. */
struct SkewState { float vdelayed[RT_MAXSKEWSAMP]; int32_t ndx_next, slots_filled; float v_last_raw; int32_t pad; };

RT_GEN inline void reset_full(const DevCfg &c, TrkState &t, SkewState &s, int trk, uint64_t row, bool time_is_zero) {
   /* memset(trkstate,0) + the non-zero members, decoder.c:437-449 */
   char *p = (char *)&t; for (unsigned i = 0; i < sizeof(TrkState); ++i) p[i] = 0;
   p = (char *)&s; for (unsigned i = 0; i < sizeof(SkewState); ++i) p[i] = 0;
   t.agc_gain = 1.0f;
   t.avg_height = RT_PKWW_PEAKHEIGHT;
   if (!c.density) { t.clk_avg = c.clk_init; for (int i = 0; i < RT_CLKRATE_WINDOW; ++i) t.clk_spacing[i] = c.clk_init; }
   t.t_clkwindow = t.clk_avg / 2 * c.p.clk_factor;
   t.init_row = row + (uint64_t)trk + (time_is_zero ? 1u : 0u);
   t.pure_from = t.init_row + (uint64_t)(c.width + c.skew[trk] + 2); }

/* ---- quiet tracking: the proof data of the speculative scan (DESIGN.md "unit equivalence") ---------
 * raw_at(j) is the sample of row j as it ENTERS the deskew FIFO.  Any detector window a scan can hold at
 * row j -- whatever row it was reset at, full or still filling, delayed or not -- only contains raw
 * samples of rows [j-L+1, j] with L = width + skew delay (peak detector) or 1 + skew delay (zero
 * crossing).  Row j is "loud" if that span could make ANY such scan in default state fire or arm:
 *   peak detector : max - min over the span >= 0.999 * pkww_rise   (required_rise at AGC 1, height 4); differentiated: additionally
 *                   an undifferentiated |v| >= DIFFERENTIATE_THRESHOLD at row j itself (see below)
 *   zero crossing : some |v| > ZEROCROSS_PEAK in the span; differentiated: additionally some
 *                   undifferentiated |v| >= DIFFERENTIATE_THRESHOLD (a reset zeroes v_last_raw, so the
 *                   first delta after it is the sample itself)
 * Conservative: "not loud" is a proof, "loud" may be a false alarm. */
RT_GEN __forceinline__ float volts_at(const DevCfg &c, const int16_t *plane, uint64_t j) {
   float v = (float)plane[j] / 32767 * c.maxvolts;
   return c.invert ? -v : v; }

RT_GEN inline float raw_at(const DevCfg &c, const int16_t *plane, uint64_t j) {
   float v = volts_at(c, plane, j);
   if (c.differentiate) {
      float prev = j ? volts_at(c, plane, j - 1) : 0.0f;
      float delta = v - prev;
      if (delta < RT_DIFF_THRESHOLD && delta > -RT_DIFF_THRESHOLD) delta = 0;
      v = delta * RT_DIFF_SCALE * c.samples_per_bit; }
   return v; }

struct QuietTracker {
   float runmin, runmax, thr; int L; uint64_t last_loud; bool primed, use_int; QuietInt qi;
   RT_GEN void init(const DevCfg &c, int trk, float quiet_thr, int quiet_thr_lsb) {
      L = (c.det == RT_DET_PEAK ? c.width : 1) + c.skew[trk];
      thr = quiet_thr; last_loud = RT_NOROW; primed = false; runmin = runmax = 0;
      use_int = c.det == RT_DET_PEAK && !c.differentiate;          /* quiet.cuh: the int16-domain test both scan kernels share */
      qi.init(L, quiet_thr_lsb); }
   /* feed row j (rows must be fed consecutively); v = raw_at(j) */
   RT_GEN void feed(const DevCfg &c, const int16_t *plane, uint64_t j, float v) {
      if (use_int) { qi.feed(plane, j, (int)plane[j]); last_loud = qi.last_loud; return; }
      if (c.det == RT_DET_PEAK) {
         if (!primed) { runmin = runmax = v; primed = true; }
         if (v < runmin) runmin = v;
         if (v > runmax) runmax = v;
         if (runmax - runmin >= thr) {              /* re-anchor on the exact span; loud only if IT is */
            float mx = v, mn = v;
            uint64_t from = j + 1 >= (uint64_t)L ? j + 1 - L : 0;
            for (uint64_t i = from; i < j; ++i) { float x = raw_at(c, plane, i); if (x > mx) mx = x; if (x < mn) mn = x; }
            runmin = mn; runmax = mx;
            if (mx - mn >= thr) last_loud = j; }
         /* -differentiate: a scan reset AT row j sees the sample itself as its first delta (v_last_raw is zeroed, readtape.c:1383-1394,
            decoder.c:437), a spike no other scan has, unless the dead-band swallows it -- so no reset row may carry |v| beyond it
            (found by the fuzz of tests/test_proof_generic_host.py, end of round 2) */
         if (c.differentiate) { const float u = volts_at(c, plane, j); if (u >= RT_DIFF_THRESHOLD || u <= -RT_DIFF_THRESHOLD) last_loud = j; } }
      else {
         bool loud = v > RT_ZEROCROSS_PEAK || v < -RT_ZEROCROSS_PEAK;
         if (c.differentiate) { float u = volts_at(c, plane, j); loud = loud || u >= RT_DIFF_THRESHOLD || u <= -RT_DIFF_THRESHOLD; }
         if (loud) last_loud = j + (uint64_t)(L - 1); } } };   /* it stays inside every span for L rows */

/* ---- skip-ahead of the exact stateful scan (k_ctx_scan) ----------------------------------------------------------------------
 * The moving-window peak detector does per-row work on every sample, but it can only FIRE at rows whose window passes the shape
 * tests, and phase A of the two-pass scan (scan_masks.cuh) has marked those rows for every threshold >= T0 (`cand`).  Between two
 * candidate rows the detector state evolves as a pure function of the samples: the ring and the deskew FIFO hold the last samples,
 * the running maximum is the exact window maximum, the blind countdown counts down, and the lazily refreshed minimum (quirk Q1,
 * decoder.c:765) follows from its last value by hopping from refresh to refresh (at an `acan` row the minimum is the window's; after
 * a refresh the next one comes when the leftmost sample carrying the minimum leaves).  So instead of walking those rows, the state
 * at the row in front of the next candidate is rebuilt from the plane -- bit-identical to having walked there.  Only for the plain
 * peak detector (no -invert / -differentiate), only while the threshold bound covers T0, and only once the window and the FIFO are
 * free of the perturbations a reset leaves behind (TrkState::pure_from). */
/* x / 32767.0f without the division subroutine: q0 = x*r, e = fma(-32767, q0, x), q = fma(e, r, q0) with r = RN(1/32767) is
   bit-identical to the IEEE quotient for every int16 x (exhaustive check in tests/test_sparse_host.py, the same routine as
   rtfast::div32767 in scan_fast.cuh) */
RT_GEN __forceinline__ float exact_div32767(float xf) {
   const float r = 1.0f / 32767.0f;
#ifdef __CUDA_ARCH__
   const float q0 = __fmul_rn(xf, r);
   return __fmaf_rn(__fmaf_rn(-32767.0f, q0, xf), r, q0);
#else
   const float q0 = xf * r;
   return fmaf(fmaf(-32767.0f, q0, xf), r, q0);
#endif
}
RT_GEN __forceinline__ float gvolts(const DevCfg &c, int x) { return exact_div32767((float)x) * c.maxvolts; }

/* the lazy minimum (int16 domain) at plane row pr, exact value m at plane row pr0 < pr; see SparseScan::lazy_min (scan_sparse.cuh) */
RT_GEN inline int lazy_min_hop(const int16_t *plane, const uint32_t *acan, int w, int64_t pr0, int64_t pr, int m) {
   int64_t a = pr0; bool have = false;
   {  int64_t wi = pr >> 5;                                       /* last acan row in (pr0, pr] */
      uint32_t bits = acan[wi] & (0xffffffffu >> (31 - (int)(pr & 31)));
      for (;;) {
         const int64_t base = wi << 5;
         if (base + 31 <= pr0) break;
         if (base <= pr0) bits &= (pr0 - base) >= 31 ? 0u : (0xffffffffu << ((int)(pr0 - base) + 1));
         if (bits) { a = base + 31 - gen_clz32(bits); have = true; break; }
         if (base <= pr0 || wi == 0) break;
         --wi; bits = acan[wi]; } }
   int64_t r = have ? a : pr0;
   bool keep = !have;
   for (;;) {
      const int64_t ws = r - w + 1;
      int mn = 32767, pmn = 0, peq = -1;
      for (int i = 0; i < w; ++i) { const int v = plane[ws + i]; if (v < mn) { mn = v; pmn = i; } if (peq < 0 && v == m) peq = i; }
      int64_t at;
      if (keep && peq >= 0) at = ws + peq; else { m = mn; at = ws + pmn; }
      keep = false;
      if (at + w > pr) break;
      r = at + w; }
   return m; }

/* Rebuild the state as it is after row `r` has been processed, coming from the state after row `cur` (cur < r), given that no row in
   (cur, r] can fire.  Window full and pure on entry.  Returns false (state untouched) if the lazy minimum's carrier cannot be found. */
RT_GEN inline bool skip_to(const DevCfg &c, TrkState &t, SkewState &s, int trk, const int16_t *plane, uint64_t cur, uint64_t r) {
   const int w = c.width, delay = c.skew[trk];
   const int64_t pr0 = (int64_t)cur - delay, pr = (int64_t)r - delay;
   /* the int16 sample that carries the lazy minimum now */
   int m = 32768;
   {  /* gvolts is strictly monotone, so the int16 value whose voltage is minv is found by inverting once and checking the neighbours;
         the window is then searched with integer compares (first occurrence, as before) */
      int guess = (int)rintf(t.minv / c.maxvolts * 32767.0f), mi = 32768;
      for (int d = -2; d <= 2; ++d) { const int x = guess + d; if (x >= -32768 && x <= 32767 && gvolts(c, x) == t.minv) { mi = x; break; } }
      if (mi == 32768) return false;
      for (int i = 0; i < w; ++i) if ((int)plane[pr0 - w + 1 + i] == mi) { m = mi; break; } }
   if (m == 32768) return false;
   m = lazy_min_hop(plane, c.m_acan + (size_t)trk * c.mask_stride, w, pr0, pr, m);
   const uint64_t n = r - cur;
   t.right = (int)(((uint64_t)t.right + n) % (uint64_t)w);
   t.left = t.right + 1 >= w ? 0 : t.right + 1;
   float mx = -100;
   for (int i = 0; i < w; ++i) {
      const float v = gvolts(c, plane[pr - w + 1 + i]);
      int ndx = t.left + i; if (ndx >= w) ndx -= w;
      t.win[ndx] = v;
      if (v > mx) mx = v; }
   t.maxv = mx; t.minv = gvolts(c, m);
   t.countdown = (uint64_t)t.countdown > n ? t.countdown - (int)n : 0;
   if (delay) {                                                  /* the FIFO holds the last `delay` raw samples */
      s.ndx_next = (int)(((uint64_t)s.ndx_next + n) % (uint64_t)delay);
      for (int i = 0; i < delay; ++i) {
         int ndx = s.ndx_next + i; if (ndx >= delay) ndx -= delay;
         s.vdelayed[ndx] = gvolts(c, plane[(int64_t)r - delay + 1 + i]); }
      s.slots_filled = delay; }
   return true; }

/* first stream row >= `from` (and < `to`) whose candidate bit is set; `to` if there is none */
RT_GEN inline uint64_t next_candidate(const DevCfg &c, int trk, uint64_t from, uint64_t to) {
   const int delay = c.skew[trk];
   const uint32_t *mc = c.m_cand + (size_t)trk * c.mask_stride;
   uint64_t p = from - (uint64_t)delay; const uint64_t pend = to - (uint64_t)delay;
   uint64_t wi = p >> 5;
   uint32_t bits = mc[wi] & (0xffffffffu << (int)(p & 31));
   for (;;) {
      if (bits) { const uint64_t q = (wi << 5) + (uint64_t)(gen_ffs32(bits) - 1); return q < pend ? q + (uint64_t)delay : to; }
      ++wi;
      if ((wi << 5) >= pend) return to;
      bits = mc[wi]; } }

/* ---- one row of one track ------------------------------------------------------------------- */
/* returns the probe bits of peak_step (0 for the other detectors / skipped rows);
   *v_out = the sample before the deskew FIFO (after invert/differentiate) */
template <class Emit>
RT_GEN inline unsigned track_row(const DevCfg &c, TrkState &t, SkewState &s, int trk, const int16_t *plane,
                                     uint64_t row, Emit &em, float *v_out) {
   /* int16 -> volts, invert, differentiate */
   float v = gvolts(c, (int)plane[row]);
   if (c.invert) v = -v;
   if (c.differentiate) {
      float delta = v - s.v_last_raw;
      if (delta < RT_DIFF_THRESHOLD && delta > -RT_DIFF_THRESHOLD) delta = 0;
      s.v_last_raw = v;
      v = delta * RT_DIFF_SCALE * c.samples_per_bit; }
   /* deskew FIFO */
   float v_now;
   int delay = c.skew[trk];
   if (delay == 0) v_now = v;
   else {
      if (s.slots_filled < delay) { v_now = v; ++s.slots_filled; }
      else v_now = s.vdelayed[s.ndx_next];
      s.vdelayed[s.ndx_next] = v;
      if (++s.ndx_next >= delay) s.ndx_next = 0; }
   *v_out = v;
   /* (Q2) rows before this track's (re)initialisation row are not looked at */
   if (t.init_row != RT_NOROW) {
      if (row < t.init_row) return 0;
      if (row == t.init_row) {
         t.win[0] = v_now;
         t.maxv = t.minv = v_now;
         t.t_lastpeak = row_time(c, row);
         t.init_row = RT_NOROW;
         t.pure_from = row + (uint64_t)(c.width + c.skew[trk] + 2);   /* slot 0 of the ring was overwritten, this row's sample not pushed */
         return 0; } }
   RowClock clk(c, row);
   unsigned probe = 0;
   if (c.det == RT_DET_PEAK) peak_step(c, t, v_now, clk, em, probe);
   else if (c.det == RT_DET_ZC) zc_step(c, t, v_now, clk, em);
   else dzc_step(c, t, v_now, clk, em);
   if (c.mode == RT_MODE_GCR && t.datablock
         && clk.now() > t.t_lastpeak + RT_GCR_IDLE_THRESH * (double)t.clk_avg)
      t.datablock = 0;
   return probe; }

/* ---- a span of rows of one track of the stateful exact scan (k_ctx_scan), with the skip-ahead ---------------------------------- */
struct CtxStats { unsigned long long walked, jumps, jumped, nothr, near; };

template <class Emit>
RT_GEN inline void ctx_scan_rows(const DevCfg &c, TrkState &t, SkewState &s, int trk, const int16_t *plane, uint64_t row_from, uint64_t row_to,
                                 Emit &em, CtxStats &cs) {
   float v;
   /* skip-ahead: jump from candidate row to candidate row where the masks of phase A are at hand */
   const bool can_skip = c.m_cand && c.m_acan && c.det == RT_DET_PEAK && !c.invert && !c.differentiate && c.T0[trk] > 0;
   const float inv_lsb = 32767.0f / c.maxvolts;
   const uint64_t min_jump = 4;                      /* rebuilding the state costs about as much as walking three rows */
   for (uint64_t row = row_from; row < row_to; ++row) {
      if (can_skip && t.init_row == RT_NOROW && t.pure_from != RT_NOROW && row > t.pure_from && row > row_from) {
         /* the integer bound of required_rise (decoder.c:785) as the two-pass scan derives it (SparseScan::thresholds) */
         const float rise = c.p.pkww_rise * (t.avg_height / RT_PKWW_PEAKHEIGHT) / t.agc_gain;
         const float q = rise * inv_lsb * 0.999f - 2.0f;
         if (q > 0 && (q > 70000.0f ? 70000 : (int)q) >= c.T0[trk]) {
            const uint64_t from = row + (uint64_t)t.countdown;          /* blind until then anyway (decoder.c:778) */
            const uint64_t nc = next_candidate(c, trk, from < row_to ? from : row_to, row_to);
            if (nc >= row + min_jump && skip_to(c, t, s, trk, plane, row - 1, nc - 1)) {
               ++cs.jumps; cs.jumped += nc - row;
               row = nc;
               if (row >= row_to) break; }
            else ++cs.near; }
         else ++cs.nothr; }
      ++cs.walked;
      track_row(c, t, s, trk, plane, row, em, &v); } }

}  // namespace rtgen
