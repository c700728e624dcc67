/* readtape_b200/csrc/scan_fast.cuh -- int16-domain fast path of the moving-window peak detector.
 *
 * Same contract as the (unit, track) scan of k_units_scan (k_scan.cu): one thread scans one track of
 * one unit from a fresh RT_RESET_FULL, produces the identical events and the identical proof data.
 * It exists because the exact generic code spends ~230 instruction slots per track-sample; this
 * formulation needs ~25.  NRZI and PE with the peak detector, no -invert / -differentiate; anything
 * else stays on the generic kernel.  Host + device code: the CPU tests run the very same functions
 * (tests/host_fast) against the oracle.
 *
 * Reference semantics reproduced (file:line in /root/reference/src):
 *   lookfor_peak   decoder.c:751-810      refine_peak  decoder.c:700-749
 *   first-sample init / staggered start    decoder.c:855-861      deskew FIFO  decoder.c:819-831
 *   process_*_transition glue              decoder.c:560-609      feedback: feedback.cuh
 *
 * How it differs from a transliteration:
 *  1. int16 domain.  v = (float)i/32767*maxvolts (readtape.c:1420) is strictly monotone in i, so the
 *     max / min / == of lookfor_peak are evaluated on the raw samples; floats are formed only at
 *     candidate rows, with the reference's exact expressions.
 *  2. O(1) sliding max and min (van Herk / Gil-Werman): rows are cut into blocks of `width` rows from
 *     the track's first sample; g = running max inside the current block, H[k] = suffix max of the
 *     previous block from position k; window max = max(g, H[j+1]).  Samples are held as packed int16x2
 *     words (x, ~x): one signed 16x2 max yields the max of x and (complemented) the min of x.
 *  3. The reference's lazily refreshed minimum (decoder.c:765 never updates pkww_minv; only the rescan at
 *     :767-775 does) is the recurrence  m <- Wmin(r)  iff  leaving >= S(r)  or  leaving == m,
 *     where S / Wmin are the exact window max / min: `leaving == max(S(r-1), v_now)` <=> leaving >= S(r).
 *     While the window is still filling, `leaving` is the constant 0.0f of decoder.c:754.
 *  4. Integer pre-filter: a row can only fire if S - max(left,right) >= T or min(left,right) - m >= T with
 *     T a conservative integer bound of required_rise (recomputed after every event, the only place
 *     AGC gain / average height change).  The exact float tests of decoder.c:790-803 run at those rows.
 *  5. Stop-and-handle loop: lanes search for their next candidate row in a tight loop and handle events
 *     together afterwards, so that a warp is not serialised on 32 different event times.
 *
 * Per-lane scratch (device: lane-interleaved shared memory, bank = lane for every access):
 *   X[RING]  ring of packed samples indexed by stream offset,  RING = pow2 >= 2*width+6
 *   H[width] suffix maxima of the previous block
 */
#pragma once
#include <string.h>
#include <math.h>
#include "rt_dev.h"
#include "feedback.cuh"
#include "quiet.cuh"

#define RT_FHD __host__ __device__ __forceinline__

namespace rtfast {

constexpr uint32_t PK_NEG = 0x80008000u;                 /* identity of vmax2 */
constexpr int32_t  OFF_NONE = INT32_MIN;                 /* "no row" for offsets relative to the unit start */

RT_FHD uint32_t pk(int x) { return ((uint32_t)x << 16) | ((uint32_t)(~x) & 0xffffu); }
RT_FHD int pk_hi(uint32_t w) { return (int)w >> 16; }                               /* x, or a max of x          */
RT_FHD int pk_min(uint32_t w) { return ~(int)(int16_t)(uint16_t)(w & 0xffffu); }     /* min of x (from max of ~x) */
RT_FHD uint32_t vmax2(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
   return __vmaxs2(a, b);
#else
   int ah = (int16_t)(a >> 16), bh = (int16_t)(b >> 16), al = (int16_t)(a & 0xffff), bl = (int16_t)(b & 0xffff);
   int h = ah > bh ? ah : bh, l = al > bl ? al : bl;
   return ((uint32_t)h << 16) | ((uint32_t)l & 0xffffu);
#endif
}

RT_FHD double row_time(const DevCfg &c, uint64_t row) {                              /* readtape.c:1423 */
   long long ns = (long long)(c.tstart_ns + row * c.tdelta_ns);
   return (double)ns / 1e9; }

/* readtape.c:1420:  v = (float)i16 / 32767 * maxvolts  (float division, then float multiplication).  The IEEE division by the
   constant is replaced by its exact equivalent  q0 = x*r, e = fma(-32767, q0, x), q = fma(e, r, q0)  with r = RN(1/32767): for
   every int16 x the result is bit-identical to x / 32767.0f (checked exhaustively by tests/test_sparse_host.py), at 3 instructions
   instead of a division subroutine. */
RT_FHD float div32767(float xf) {
   const float r = 1.0f / 32767.0f;
#ifdef __CUDA_ARCH__
   const float q0 = __fmul_rn(xf, r);
   return __fmaf_rn(__fmaf_rn(-32767.0f, q0, xf), r, q0);
#else
   const float q0 = xf * r;
   return fmaf(fmaf(-32767.0f, q0, xf), r, q0);
#endif
}
RT_FHD float volts(const DevCfg &c, int x) { return div32767((float)x) * c.maxvolts; }

/* Lanes run FAST_K rows between two maintenance points (loader, job switch); a lane that meets a candidate row waits
   for the handler, which runs when FAST_NPEND lanes of the warp are waiting or at the maintenance point. */
#define FAST_K      16
/* gap skipping (try_skip): granules examined per attempt, and the quiet rows left in front of the first loud granule */
#define FAST_SKIP_LOOKAHEAD 512
#define FAST_M_UNKNOWN      (1 << 20)

/* ring size for a window width: holds offsets [o-w+1, o+FAST_K+w+7] */
RT_FHD uint32_t ring_size(int w) { uint32_t r = 32; while (r < (uint32_t)(2 * w + FAST_K + 8)) r <<= 1; return r; }

/* lane-private scratch; STRIDE = distance (in words) between consecutive entries of one lane.
 *   X[RING]      : ring of packed samples indexed by stream offset (the loader's target)
 *   H[2][w+1]    : suffix maxima of the previous block (read by the forward step: row j reads H[prev][j+1], entry w is
 *                  the identity) and of the current block (written by the backward step, one entry per row)
 *   HT[10]       : AGC height history (v_heights[], decoder.h:226) -- indexed dynamically, so not in registers */
template <int STRIDE>
struct LaneMem {
   uint32_t *x, *h, *ht; uint32_t mask;
   RT_FHD uint32_t &X(uint32_t o) const { return x[(size_t)(o & mask) * STRIDE]; }
   RT_FHD uint32_t &H(int k) const { return h[(size_t)k * STRIDE]; } };

/* exact min / max of plane[from .. to) and x: the rare re-anchoring step of the loudness test.  Kept out of line so that
   its address arithmetic is not hoisted into the row loop. */
struct minmax { int mn, mx; };
__host__ __device__ __noinline__ static minmax span_minmax(const int16_t *plane, int64_t from, int64_t to, int x) {
   minmax r{x, x};
   for (int64_t i = from; i < to; ++i) { const int y = plane[i]; if (y > r.mx) r.mx = y; if (y < r.mn) r.mn = y; }
   return r; }

template <int STRIDE>
struct HeightsRef {
   uint32_t *p;
   RT_FHD float &operator[](int i) const { return reinterpret_cast<float *>(p)[(size_t)i * STRIDE]; } };

/* The feedback state of one track: the members of TrkState (rt_dev.h) the peak detector path uses. */
template <int STRIDE>
struct FastState {
   double  t_top, t_bot, t_lastpeak;
   float   v_top, v_bot, v_lasttop, v_lastbot;
   float   avg_height, avg_height_sum, agc_gain;
   HeightsRef<STRIDE> heights;
   int32_t avg_height_count, heightndx, peakcount;
   float   t_clkwindow;
   uint8_t datablock, bit1_up, failed;
};

/* words of scratch one lane needs: ring, the two H arrays, the heights */
RT_FHD uint32_t scratch_words(int w) { return ring_size(w) + 2u * (uint32_t)(w + 1) + RT_AGC_MAX_WINDOW; }
template <int STRIDE>
RT_FHD LaneMem<STRIDE> lane_mem(uint32_t *lane_base /* first word of this lane in a [scratch_words][STRIDE] array */, int w) {
   LaneMem<STRIDE> m;
   m.x = lane_base; m.mask = ring_size(w) - 1;
   m.h = m.x + (size_t)ring_size(w) * STRIDE; m.ht = m.h + (size_t)2 * (w + 1) * STRIDE;
   return m; }
/* One lane's scan of one (unit, track) as a resumable state machine, so that the 32 lanes of a warp can be driven
 * together (drive()):  begin() -> { step() | handle() }* -> finish().   step() processes exactly one row: forward step of the
 * sliding max/min, one backward step of the current block's suffix maxima, lazy-minimum update, pre-filter.  A row that
 * passes the pre-filter leaves the lane waiting (pend) for handle(), which runs the exact tests for all waiting lanes at once.
 * A "block" is `width` consecutive rows counted from the track's first sample (the van Herk blocks). */
enum { ST_RUN = 0, ST_PEND = 1, ST_DONE = 2 };          /* scanning | stopped on a candidate row, waiting for handle() | unit finished */
template <int STRIDE, class Emit>
struct UnitScan {
   const DevCfg &c; const int16_t *plane; uint64_t row0; uint32_t end; int trk, w, delay; uint32_t io;
   LaneMem<STRIDE> mem; Emit em; FastState<STRIDE> t;
   /* detector state */
   int m, T, blind; uint32_t lvw; float inv_lsb, rise, reqmin;
   /* position: next row o = b + j; g = running packed max of block [b, b+w) up to j, hbk = running max of its backward pass;
      hp / hc = first entries of the previous / current block's H array; fillblk = the window is still filling */
   uint32_t o, b, g, hbk; int j, hp, hc, st, fillblk, candA, pre;
   /* loader: offsets < ld are in the ring */
   uint32_t ld;
   /* gap skipping: this track's granule min/max map (null = off), next offset at which an attempt is worthwhile, and the
      offset by which the lazily kept minimum must have been re-established after a jump (0 = not pending) */
   const uint32_t *gm; uint32_t next_try, munk_until, nskipped;
   /* proof data, kept lazily (offsets relative to row0; OFF_NONE = none): see track() */
   int qmin, qmax, qthr, qL; int32_t ll, last_canon;
   int32_t sync_row, loud_at_sync, sync_first, sync_early, loud_early; bool early_frozen; uint32_t sf_from; uint64_t quiet_from;

   RT_FHD UnitScan(const DevCfg &c_, LaneMem<STRIDE> mem_) : c(c_), w(c_.width), mem(mem_) { st = ST_DONE; j = 0; o = end = 0; ld = 0; pf_at = 0xffffffffu; }

   /* the sample the detector sees at stream offset o (deskew FIFO, decoder.c:819-831) */
   RT_FHD int sample(uint32_t o) const { return (int)plane[row0 + (o >= (uint32_t)delay ? o - (uint32_t)delay : o)]; }

   /* ---- loader: 16-byte chunks of 8 samples -> packed ring entries --------------------------------------------------
      The chunks the next batch of rows will need are requested one batch ahead (prefetch) and stored when they are due
      (ensure), so that no lane ever waits for DRAM inside the row loop. */
   struct chunk8 { uint32_t w[4]; };
   RT_FHD chunk8 load_chunk(uint32_t at) const {                  /* samples of stream offsets [at, at+8), at >= delay */
      const int16_t *p = plane + row0 + (at - (uint32_t)delay);    /* 16-byte aligned: row0 % 32 == 0, (at-delay) % 8 == 0 */
      chunk8 q;
#ifdef __CUDA_ARCH__
      const uint4 v = *reinterpret_cast<const uint4 *>(p);
      q.w[0] = v.x; q.w[1] = v.y; q.w[2] = v.z; q.w[3] = v.w;
#else
      memcpy(q.w, p, 16);
#endif
      return q; }
   RT_FHD void store_chunk(uint32_t at, const chunk8 &q) const {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
         const uint32_t v = q.w[i], nv = ~v;
         mem.X(at + 2 * i) = (v << 16) | (nv & 0xffffu);
         mem.X(at + 2 * i + 1) = (v & 0xffff0000u) | (nv >> 16); } }
   chunk8 pf0, pf1; uint32_t pf_at;                               /* two chunks in flight for offsets pf_at, pf_at+8; pf_at = ~0: none */
   RT_FHD void prefetch() {
      pf_at = 0xffffffffu;
      if (ld < (uint32_t)delay || ld >= end + 16u) return;
      pf_at = ld; pf0 = load_chunk(ld); pf1 = load_chunk(ld + 8); }
   /* make ring entries of offsets < upto available */
   RT_FHD void ensure(uint32_t upto) {
      if (pf_at == ld) {                                          /* the prefetched chunks are the next ones due */
         if (ld < upto) { store_chunk(ld, pf0); ld += 8; }
         if (ld < upto && ld == pf_at + 8) { store_chunk(ld, pf1); ld += 8; } }
      pf_at = 0xffffffffu;
      while (ld < upto) {
         if (ld < (uint32_t)delay) { mem.X(ld) = pk(sample(ld)); ++ld; }
         else { store_chunk(ld, load_chunk(ld)); ld += 8; } } }

   /* required_rise / required_min (decoder.c:785-786) only change at events: cached, with the integer bound T */
   RT_FHD void thresholds() {
      rise = c.p.pkww_rise * (t.avg_height / RT_PKWW_PEAKHEIGHT) / t.agc_gain;
      reqmin = c.p.min_peak * (t.avg_height / RT_PKWW_PEAKHEIGHT) / t.agc_gain;
      /* a top needs S_f > X_f + rise and |volts(i) - i*lsb| <= 2^-22 * 32768 * lsb, so S - X >= rise/lsb*(1-1e-3) - 2 */
      float q = rise * inv_lsb * 0.999f - 2.0f;
      T = !(q > 0) ? 0 : (q > 70000.0f ? 70000 : (int)q); }

   /* ---- proof data of the unit-equivalence test, collected before the first event ------------------
    * k_units_scan (k_scan.cu) evaluates, on every row: loudness (quiet.cuh), then
    *     if (sync_early set && last_loud != loud_early) early_frozen = true;
    *     if (canonical && last_loud < row) { sync_row = row; loud_at_sync = last_loud; if (!early_frozen) {sync_early = row; loud_early = last_loud;}
    *                                         if (sync_first unset && row >= own_fill && last_loud < row0) sync_first = row; }
    * last_loud only changes on loud rows, which are rare, so the same values are obtained by remembering the last
    * canonical quiet row (one predicated move per row) and committing it whenever last_loud is about to change. */
   RT_FHD void commit() {
      if (last_canon != OFF_NONE && last_canon != sync_row) {        /* a canonical row newer than the committed one */
         sync_row = last_canon; loud_at_sync = ll;
         if (!early_frozen) { sync_early = last_canon; loud_early = ll; } } }
   RT_FHD void loud_row(int32_t o) {
      commit();
      if (sync_early != OFF_NONE) early_frozen = true;
      ll = o;
      if (o >= 0) sf_from = 0xffffffffu; }
   /* loudness of the row at offset o (negative: the pre-scan) whose raw (undelayed) sample is `raw` */
   RT_FHD void feed(int32_t o, int raw) {
      if (raw < qmin) qmin = raw;
      if (raw > qmax) qmax = raw;
      if (qmax - qmin >= qthr) {                   /* re-anchor on the exact span; loud only if IT is */
         const int64_t to = (int64_t)row0 + o, from = to - qL + 1;
         const minmax r = span_minmax(plane, from < 0 ? 0 : from, to, raw);
         qmin = r.mn; qmax = r.mx;
         if (r.mx - r.mn >= qthr) loud_row(o); } }
   RT_FHD void track(uint32_t o, int xraw_stream, bool canonical) {
      feed((int32_t)o, delay ? (int)plane[row0 + o] : xraw_stream);
      if (canonical && ll != (int32_t)o) {
         last_canon = (int32_t)o;
         if (o >= sf_from) { sync_first = (int32_t)o; sf_from = 0xffffffffu; } } }

   /* refine_peak, decoder.c:700-749: window = stream offsets [wstart, o]; v = volts(val).
      Written for lanes in lock step: a fixed-trip search for the FIRST sample equal to val (no early exit), and one code
      path for both polarities -- a bottom peak is a top peak of the negated signal (float negation is exact and rounding
      is symmetric, so  a < b + e  <=>  -a > -b - e  bit for bit). */
   RT_FHD double refine(int val, float v, bool top, uint32_t wstart, uint32_t o) {
      uint32_t pos = 0xffffffffu;
      for (uint32_t i = o + 1; i-- > wstart;) if (pk_hi(mem.X(i)) == val) pos = i;          /* ends on the leftmost match */
      const int left_distance = (int)(pos - wstart) + 1;
      if (pos == 0xffffffffu || left_distance >= w || pos == wstart) { t.failed = 2; return 0; }   /* the reference would fatal() */
      const float sg = top ? 1.0f : -1.0f;
      const float vprev = sg * volts(c, pk_hi(mem.X(pos - 1)));
      const float vnext = pos < o ? sg * volts(c, pk_hi(mem.X(pos + 1))) : 0.0f;     /* pos == o: only in a filling window, the unwritten slot */
      const float edge = sg * v - RT_PEAK_THRESHOLD / t.agc_gain;
      float adj = 0;
      if (vprev > edge && vnext < edge) adj = -0.5f;
      else if (vnext > edge && vprev < edge) adj = +0.5f;
      const double timenow = row_time(c, row0 + o);
      blind = left_distance;
      return timenow - (double)(((float)(w - left_distance) - adj) * c.sample_deltat); }

   /* process_*_transition, decoder.c:560-609 */
   RT_FHD void transition(bool top, uint32_t o) {
      const double t_ev = top ? t.t_top : t.t_bot;
      const float v_top_seen = t.v_top, v_bot_seen = t.v_bot;
      ++t.peakcount;
      if (c.density) { /* doing_density_detection: the mode handlers are bypassed (decoder.c:578), AGC and average height stay put */ }
      else if (c.mode == RT_MODE_NRZI) rtfb::nrzi_feedback(c, t, top);
      else if (c.mode == RT_MODE_PE) rtfb::pe_feedback(c, t, top, t_ev);
      else rtfb::agc_adjust(c, t);
      if (top) t.v_lasttop = t.v_top; else t.v_lastbot = t.v_bot;
      t.t_lastpeak = t_ev;
      em.emit(row0 + o, t_ev, v_top_seen, v_bot_seen, t.agc_gain, top);
      thresholds(); }

   /* start the scan of unit rows [row0_, row_end) of track trk_ from a fresh RT_RESET_FULL */
   RT_FHD void begin(const int16_t *plane_, uint64_t row0_, uint64_t row_end, int trk_, Emit em_, int quiet_thr_lsb) {
      plane = plane_; row0 = row0_; end = (uint32_t)(row_end - row0_); trk = trk_; delay = c.skew[trk_]; em = em_;
      gm = c.gmm ? c.gmm + (size_t)trk_ * c.ngran_cap : nullptr; next_try = 0; munk_until = 0; nskipped = 0;
      const bool tz = row_time(c, row0) == 0.0;
      io = (uint32_t)trk + (tz ? 1u : 0u);
      /* reset_full (scan_generic.cuh) restricted to the members of FastState */
      t.t_top = t.t_bot = t.t_lastpeak = 0; t.v_top = t.v_bot = t.v_lasttop = t.v_lastbot = 0; t.avg_height_sum = 0;
      t.avg_height_count = t.heightndx = t.peakcount = 0; t.datablock = t.bit1_up = t.failed = 0;
      t.heights.p = mem.ht;
      for (int i = 0; i < RT_AGC_MAX_WINDOW; ++i) t.heights[i] = 0.0f;
      t.agc_gain = 1.0f; t.avg_height = RT_PKWW_PEAKHEIGHT;
      t.t_clkwindow = c.clk_init / 2 * c.p.clk_factor;
      inv_lsb = 32767.0f / c.maxvolts;
      thresholds(); blind = 0; m = 0; lvw = 0;
      /* proof data: quiet pre-scan of the rows before the unit, then the rows up to the track's first sample */
      qthr = quiet_thr_lsb; qL = w + delay; qmin = 32767; qmax = -32768; ll = last_canon = OFF_NONE;
      sync_row = loud_at_sync = sync_first = sync_early = loud_early = OFF_NONE; early_frozen = false;
      const int lead = (int)io > delay ? (int)io : delay;
      sf_from = (uint32_t)(lead + w + 1);
      const int32_t npre = row0 > (uint64_t)c.prescan_rows ? c.prescan_rows : (int32_t)row0;
      for (int32_t o = -npre; o < 0; ++o) feed(o, (int)plane[(int64_t)row0 + o]);
      quiet_from = ll == OFF_NONE ? row0 - (uint64_t)npre : (uint64_t)((int64_t)row0 + ll + 1);
      for (uint32_t o = 0; o <= io && o < end; ++o) track(o, sample(o), false);      /* decoder.c:855-861: not looked at yet */
      j = 0; g = hbk = PK_NEG; fillblk = true; pre = true; o = b = io + 1; hp = 0; hc = w + 1;
      st = b >= end ? ST_DONE : ST_RUN;
      if (st == ST_RUN) {
         const int x0 = sample(io);
         m = x0; lvw = pk(x0);
         t.t_lastpeak = row_time(c, row0 + io);
         ld = b < (uint32_t)delay ? b : (b - (uint32_t)delay) / 8 * 8 + (uint32_t)delay; pf_at = 0xffffffffu;
         ensure(o + FAST_K + (uint32_t)w);
         mem.X(io) = pk(x0);                                       /* the loader starts at b (or at the chunk that holds it) */
         /* the window of the filling phase always starts at the first sample: with that sample as the "previous block"
            the steady-state expression yields max(x0, g); entry w is the identity (row w-1's window is the block itself) */
         for (int k = 1; k < w; ++k) mem.H(hp + k) = pk(x0);
         mem.H(hp + w) = PK_NEG; mem.H(hc + w) = PK_NEG;
         /* the first block (window still filling) is walked here, once per job, so that step() does not carry its cases */
         while (fillblk && st != ST_DONE) { step_fill(); if (st == ST_PEND) handle(); }
         if (st == ST_RUN) ensure(o + FAST_K + (uint32_t)w); } }

   /* move to the next row; at a block boundary the block just finished becomes the "previous block" */
   RT_FHD void step_pos() {
      ++o;
      if (++j == w) { j = 0; b += (uint32_t)w; g = hbk = PK_NEG; const int x = hp; hp = hc; hc = x; fillblk = false; }
      if (o >= end) st = ST_DONE; }

   /* one row.  decoder.c:753-775 (window update, exact max S, lazily refreshed minimum m), :778 (blind countdown), and the
      integer pre-filter of the shape tests :790-803 */
   RT_FHD void step() {
      const uint32_t xw = mem.X(o);
      const uint32_t xlw = mem.X(o - (uint32_t)w + 1u);           /* the window's left edge */
      g = vmax2(g, xw);
      const uint32_t sw = vmax2(g, mem.H(hp + j + 1));
      const int S = pk_hi(sw);
      const int lv = pk_hi(lvw);
      const bool A = lv >= S;                                     /* leaving == max(maxv, v_now)  <=>  leaving >= S */
      if (A || lv == m) m = pk_min(sw);                           /* the rescan of decoder.c:767-775 */
      lvw = xlw;
      const int kb = w - 1 - j;                                   /* backward pass of this block, one position per row */
      if (kb >= 1) { hbk = vmax2(hbk, mem.X(o + (uint32_t)(kb - j))); mem.H(hc + kb) = hbk; }   /* block start + kb */
      if (blind) { --blind; step_pos(); return; }
      const uint32_t xy = vmax2(xlw, xw);                         /* (max(left,right), ~min(left,right)) */
      if (S - pk_hi(xy) >= T || pk_min(xy) - m >= T) { st = ST_PEND; candA = A; return; }
      if (pre) track(o, pk_hi(xw), A);
      step_pos(); }

   /* the same for the block right after the track's first sample, whose window is still filling for j < w-1 */
   RT_FHD void step_fill() {
      const uint32_t xw = mem.X(o);
      const bool full = j == w - 1;                               /* the window holds `width` samples */
      const uint32_t xlw = mem.X(full ? o - (uint32_t)w + 1u : io);
      g = vmax2(g, xw);
      const uint32_t sw = vmax2(g, mem.H(hp + j + 1));
      const int S = pk_hi(sw);
      const int lv = full ? pk_hi(lvw) : 0;                       /* decoder.c:754: old_left stays 0 until the window is full */
      const bool A = full ? lv >= S : S == 0;
      if (A || lv == m) m = pk_min(sw);
      lvw = xlw;
      const int kb = w - 1 - j;
      if (kb >= 1) { hbk = vmax2(hbk, mem.X(b + (uint32_t)kb)); mem.H(hc + kb) = hbk; }
      if (blind) { --blind; step_pos(); return; }
      const uint32_t xy = vmax2(xlw, xw);
      if (S - pk_hi(xy) >= T || pk_min(xy) - m >= T) { st = ST_PEND; candA = full && A; return; }
      if (pre) track(o, pk_hi(xw), full && A);
      step_pos(); }

   /* the exact tests of decoder.c:790-803 at the candidate row o, then move on */
   RT_FHD void handle() {
      const uint32_t xw = mem.X(o);
      const bool full = !fillblk || j == w - 1;
      const uint32_t wstart = full ? o - (uint32_t)w + 1u : io;
      const uint32_t xlw = mem.X(wstart);
      const int S = pk_hi(vmax2(g, mem.H(hp + j + 1)));
      const float vl = volts(c, pk_hi(xlw)), vr = volts(c, pk_hi(xw));
      const float maxv = volts(c, S), minv = volts(c, m);
      const bool top = maxv > vl + rise && maxv > vr + rise && (reqmin == 0 || maxv > reqmin);
      const bool bot = !top && minv < vl - rise && minv < vr - rise && (reqmin == 0 || minv < -reqmin);
      if (top || bot) {
         if (pre) { commit(); pre = false; }
         const double tp = refine(top ? S : m, top ? maxv : minv, top, wstart, o);
         if (top) { t.v_top = maxv; t.t_top = tp; } else { t.v_bot = minv; t.t_bot = tp; }
         transition(top, o); }
      else if (pre) track(o, pk_hi(xw), candA);
      st = ST_RUN;
      step_pos(); }

   /* Gap skipping, called between two batches.  Let thr be the integer bound T of required_rise (and, before the first
      event, also the loudness threshold of the proof).  If the raw samples of rows [A, B) span less than thr, no row whose
      window lies inside [A, B) can pass the pre-filter:  S - max(l,r) <= max - min  and  min(l,r) - m <= max - min  because
      the lazily kept minimum m is always the value of a sample inside the window.  Such rows change nothing but the window
      state, so the lane jumps over them: it restarts its block structure `width` rows before the target (one warm-up block
      with the tests off rebuilds ring, g and the suffix maxima), and marks m unknown; m is exact again at the first row
      where the window maximum leaves (A-row) -- the jump stops early enough for that to happen inside the quiet stretch,
      and a unit whose m is still unknown after the margin is flagged failed (the host then uses the exact scan).
      The span test uses the 32-row granule min/max map of k_ingest.cu. */
   RT_FHD void try_skip() {
      if (!gm || o < next_try || o < (uint32_t)(delay + w) || fillblk) return;
      if (pre && sf_from != 0xffffffffu) return;                  /* sync_first (unit chaining) wants the FIRST canonical row: find it first */
      if (munk_until) {
         if (m != FAST_M_UNKNOWN) munk_until = 0;
         else if (o >= munk_until) { t.failed = 4; munk_until = 0; }
         else return; }
      const int thr = pre && qthr < T ? qthr : T;
      const int margin = 12 * w + 32;
      next_try = o + 64;
      if (thr < 8) return;
      const uint64_t first_raw = row0 + o - (uint32_t)delay - (uint32_t)(w - 1);    /* earliest raw row in the window of row o */
      const uint64_t gi = first_raw / RT_GRAN, gend = (row0 + end - (uint32_t)delay + RT_GRAN - 1) / RT_GRAN;
      const uint64_t gmax = gi + FAST_SKIP_LOOKAHEAD < gend ? gi + FAST_SKIP_LOOKAHEAD : gend;
      int mn = 32767, mx = -32768;
      uint64_t gq = gi;
      bool open_run = true;
      while (open_run && gq < gmax) {
         /* eight granules per round trip (two 16-byte loads) where alignment and range allow, else one */
         uint32_t v[8]; int nv = 1;
         if ((gq & 3) == 0 && gq + 8 <= gmax) {
#ifdef __CUDA_ARCH__
            const uint4 q0 = *reinterpret_cast<const uint4 *>(gm + gq), q1 = *reinterpret_cast<const uint4 *>(gm + gq + 4);
            v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
#else
            memcpy(v, gm + gq, 32);
#endif
            nv = 8; }
         else v[0] = gm[gq];
#pragma unroll
         for (int i = 0; i < 8; ++i) {
            if (i < nv && open_run) {
               const int a = (int)(int16_t)(uint16_t)(v[i] & 0xffffu), bb = (int)(int16_t)(uint16_t)(v[i] >> 16);
               const int nmn = a < mn ? a : mn, nmx = bb > mx ? bb : mx;
               if (nmx - nmn >= thr) open_run = false;
               else { mn = nmn; mx = nmx; ++gq; } } } }
      /* raw rows [32*gi, 32*gq) are quiet: stream rows < oq cannot fire */
      int64_t oq = (int64_t)(gq * RT_GRAN) - (int64_t)row0 + delay;
      const bool to_end = oq >= (int64_t)end;
      if (to_end) oq = end;
      const int64_t target = to_end ? oq : oq - margin;
      if (target < (int64_t)o + w + 32) return;
      const uint32_t jump = (uint32_t)target - o;
      const int blind_after = (uint32_t)blind > jump ? blind - (int)jump : 0;
      if (pre) { if (mn < qmin) qmin = mn; if (mx > qmax) qmax = mx; }      /* the loudness tracker stays conservative */
      nskipped += jump - (to_end ? 0u : (uint32_t)w);
      if (to_end) { o = end; st = ST_DONE; return; }
      o = b = (uint32_t)target - (uint32_t)w; j = 0; g = hbk = PK_NEG;
      ld = (o - (uint32_t)delay) / 8 * 8 + (uint32_t)delay; pf_at = 0xffffffffu;
      ensure(o + FAST_K + 2u * (uint32_t)w);
      blind = w;
      for (int i = 0; i < w; ++i) step();                         /* warm-up block: tests off */
      blind = blind_after; m = FAST_M_UNKNOWN; munk_until = (uint32_t)target + (uint32_t)margin; next_try = o; }

   RT_FHD void finish(TrkMeta &meta) {
      if (pre) commit();
      meta.first_event_row = em.first_row;
      meta.sync_row = sync_row == OFF_NONE ? RT_NOROW : row0 + (uint64_t)sync_row;
      meta.last_loud_row = loud_at_sync == OFF_NONE ? RT_NOROW : (uint64_t)((int64_t)row0 + loud_at_sync);
      meta.sync_first = sync_first == OFF_NONE ? RT_NOROW : row0 + (uint64_t)sync_first;
      meta.quiet_from = quiet_from;
      meta.sync_early = sync_early == OFF_NONE ? RT_NOROW : row0 + (uint64_t)sync_early;
      meta.loud_early = loud_early == OFF_NONE ? RT_NOROW : (uint64_t)((int64_t)row0 + loud_early);
      meta.first_chunk = em.first_chunk; meta.nevents = em.n; meta.failed = t.failed; meta.pad = nskipped;
      meta.last_event_row = em.last_row; meta.quiet_tail_from = RT_NOROW; } };   /* pad: rows jumped over (diagnostics) */

/* Drive one lane (host) or the 32 lanes of a warp (device) through a list of (unit, track) jobs.  `Jobs` provides
 *   bool next(UnitScan&)   start the lane's next job, false if there is none (called by all lanes of the warp together)
 *   bool exhausted         set by next() when the job list has run out (the same for all lanes)
 *   void done(UnitScan&)   the lane's job is finished (store its TrkMeta)
 * and `count(pred)` is the warp vote popc(ballot(pred)) (0/1 on the host).  Between two votes every running lane walks up
 * to FAST_K rows of ITS job; a lane that stops on a candidate row waits for the end of the batch, then all waiting lanes run
 * the exact tests TOGETHER (tests + refine_peak + AGC + event store cost ~10 row steps, so they must not run for one lane
 * at a time); then finished lanes pick their next job and every lane refills its sample ring. */
template <class Scan, class Jobs, class Vote>
RT_FHD void drive(Scan &us, Jobs &jobs, Vote count) {
   for (;;) {
      /* the warp takes its next GROUP of jobs (32 consecutive (unit, track) pairs = the tracks of 3-4 neighbouring units) only when
         all its lanes are idle: the lanes then meet blocks and gaps together -- large handler batches, converged jumps */
      bool active = jobs.next(us);
      if (jobs.exhausted) return;
      while (count(active)) {
         if (active && us.st == ST_RUN) {
            us.prefetch();                                         /* the chunks due at the end of this batch: requested now */
            for (int k = 0; k < FAST_K; ++k) { us.step(); if (us.st != ST_RUN) break; } }
         if (count(active && us.st == ST_PEND)) { if (active && us.st == ST_PEND) us.handle(); }
         if (active && us.st == ST_RUN) us.try_skip();
         if (active && us.st == ST_DONE) { jobs.done(us); active = false; }
         if (active && us.st == ST_RUN) us.ensure(us.o + FAST_K + (uint32_t)us.w); } } }

}  // namespace rtfast
