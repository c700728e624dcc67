/* readtape_b200/csrc/scan_zc.cuh -- int16-domain fast path of the zero-crossing detector for GCR (-zeros).
 *
 * Same contract and the same warp driver (rtfast::drive, scan_fast.cuh) as the peak-detector fast path: one lane scans one
 * track of one unit from a fresh RT_RESET_FULL and produces the identical events and proof data as k_units_scan.  All three
 * GCR examples of the reference and BASELINE config 4 use this detector.  GCR blocks are long (a 4 KB block is ~600 000
 * rows at 6.25 MHz), so a reel offers few (unit, track) jobs and the scan is bound by the LATENCY of one row, which is what
 * this formulation attacks: the generic code needs ~3 us per row (state in local memory, a double-precision row time on
 * every row), this one a few dozen integer instructions on registers.
 *
 * Reference semantics reproduced (file:line in /root/reference/src):
 *   lookfor_zerocrossing            decoder.c:617-649          first-sample init / staggered start  decoder.c:855-861
 *   process_*_transition glue       decoder.c:560-609          deskew FIFO                          decoder.c:819-831
 *   gcr_top/bot -> gcr_checkzeros, gcr_addbit (per-track clock, resync)   decode_gcr.c:731-865
 *   adjust_clock / force_clock      decoder.c:533-558          GCR idle test                        decoder.c:879-882
 *
 * How it differs from a transliteration:
 *  1. int16 domain: v > 0, v_top < v, v_top > ZEROCROSS_PEAK ... are evaluated on the raw samples (volts() is strictly
 *     monotone and odd); I_ZC is the smallest sample whose voltage exceeds ZEROCROSS_PEAK.  AGC is off with -zeros
 *     (decoder.c:501), so the gain in every event is the constant 1.
 *  2. Times are kept as row numbers; the two doubles the reference compares are formed only at candidate rows.
 *  3. The per-row GCR idle test only clears `datablock`, which nothing reads before the next transition: it is evaluated
 *     there, for the last row the reference would have tested (row_time is monotone), not on every row.
 *  4. The NRZI-style AGC baseline bookkeeping the GCR handler also runs (decode_gcr.c:843-864) only feeds adjust_agc, which
 *     is disabled: dropped.
 */
#pragma once
#include "scan_fast.cuh"

namespace rtfast {

template <int STRIDE>
struct ZcMem {
   uint32_t *x, *clk; uint32_t mask;                              /* ring of samples (as int), clock-spacing window */
   RT_FHD int &X(uint32_t o) const { return reinterpret_cast<int *>(x)[(size_t)(o & mask) * STRIDE]; }
   RT_FHD float &CLK(int i) const { return reinterpret_cast<float *>(clk)[(size_t)i * STRIDE]; } };

#define ZC_RING 64u
RT_FHD uint32_t zc_scratch_words(const DevCfg &c) { return ZC_RING + (uint32_t)(c.p.clk_window > 0 ? c.p.clk_window : 1); }
template <int STRIDE>
RT_FHD ZcMem<STRIDE> zc_mem(uint32_t *lane_base) {
   ZcMem<STRIDE> m; m.x = lane_base; m.mask = ZC_RING - 1; m.clk = lane_base + (size_t)ZC_RING * STRIDE; return m; }

RT_FHD bool zc_scan_eligible(const DevCfg &c) {
   return c.det == RT_DET_ZC && c.mode == RT_MODE_GCR && !c.invert && !c.differentiate && !c.density; }

template <int STRIDE, class Emit>
struct ZcScan {
   const DevCfg &c; const int16_t *plane; uint64_t row0; uint32_t end; int trk, w, delay; uint32_t io;
   ZcMem<STRIDE> mem; Emit em;
   /* detector state (decoder.h:194-255): extremes since the last crossing, previous sample, pending crossings and the rows they were armed at */
   int vtop, vbot, vprev, izc; int upp, dnp; uint32_t ttop, tbot;
   uint32_t o; int st, pre;
   /* per-track GCR clock and bit history (decode_gcr.c) */
   double t_lastpeak, idle_after;                                 /* idle_after: t_lastpeak + GCR_IDLE_THRESH * clk_avg as of the last transition */
   float clk_avg, t_peakdelta, t_peakdeltaprev, t_pulse_adj;
   int clk_ndx, datacount, resync_bitcount, datablock, lastbits, bit_m1, bit_m2;
   /* loader */
   uint32_t ld, pf_at; struct chunk8 { uint32_t w[4]; } pf0, pf1;
   /* proof data (lazy, as in UnitScan) */
   int qL; int32_t ll, last_canon, sync_row, loud_at_sync, sync_first, sync_early, loud_early; bool early_frozen; uint32_t sf_from, own_fill; uint64_t quiet_from;

   RT_FHD ZcScan(const DevCfg &c_, ZcMem<STRIDE> mem_) : c(c_), w(0), mem(mem_) { st = ST_DONE; o = end = 0; ld = 0; pf_at = 0xffffffffu; }

   RT_FHD int sample(uint32_t off) const { return (int)plane[row0 + (off >= (uint32_t)delay ? off - (uint32_t)delay : off)]; }

   /* ---- loader (same scheme as UnitScan: chunks requested one batch ahead) */
   RT_FHD chunk8 load_chunk(uint32_t at) const {
      const int16_t *p = plane + row0 + (at - (uint32_t)delay);
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
         mem.X(at + 2 * i) = (int)(int16_t)(uint16_t)(q.w[i] & 0xffffu);
         mem.X(at + 2 * i + 1) = (int)q.w[i] >> 16; } }
   RT_FHD void prefetch() {
      pf_at = 0xffffffffu;
      if (ld < (uint32_t)delay || ld >= end + 16u) return;
      pf_at = ld; pf0 = load_chunk(ld); pf1 = load_chunk(ld + 8); }
   RT_FHD void ensure(uint32_t upto) {
      if (pf_at == ld) {
         if (ld < upto) { store_chunk(ld, pf0); ld += 8; }
         if (ld < upto && ld == pf_at + 8) { store_chunk(ld, pf1); ld += 8; } }
      pf_at = 0xffffffffu;
      while (ld < upto) {
         if (ld < (uint32_t)delay) { mem.X(ld) = sample(ld); ++ld; }
         else { store_chunk(ld, load_chunk(ld)); ld += 8; } } }
   RT_FHD void try_skip() {}                                      /* GCR reels are ~94 % signal: nothing to jump over */

   /* ---- proof data: k_units_scan's zero-crossing rules (quiet tracking of scan_generic.cuh), kept lazily like UnitScan's */
   RT_FHD void commit() {
      if (last_canon != OFF_NONE && last_canon != sync_row) {
         sync_row = last_canon; loud_at_sync = ll;
         if (!early_frozen) { sync_early = last_canon; loud_early = ll; } } }
   RT_FHD void loud_until(int32_t upto) {                         /* a sample beyond ZEROCROSS_PEAK stays inside every span for L rows */
      if (upto == ll) return;
      commit();
      if (sync_early != OFF_NONE) early_frozen = true;
      ll = upto;
      if (upto >= 0) sf_from = 0xffffffffu; }
   RT_FHD void track(int32_t off, int raw) {
      if (raw >= izc || raw <= -izc) loud_until(off + qL - 1);
      if (off >= (int32_t)own_fill && ll < off) {
         last_canon = off;
         if ((uint32_t)off >= sf_from) { sync_first = off; sf_from = 0xffffffffu; } } }

   /* ---- the per-track GCR clock, decoder.c:533-558 */
   RT_FHD void clk_adjust(float delta) {
      const int win = c.p.clk_window; const float alpha = c.p.clk_alpha;
      if (win > 0) {
         const float old = mem.CLK(clk_ndx);
         mem.CLK(clk_ndx) = delta;
         if (++clk_ndx >= win) clk_ndx = 0;
         clk_avg += (delta - old) / win; }
      else if (alpha > 0) clk_avg = alpha * delta + (1 - alpha) * clk_avg;
      else clk_avg = 0.0f; }
   RT_FHD void clk_force(float v) {
      for (int i = 0; i < c.p.clk_window; ++i) mem.CLK(i) = v;     /* entries beyond clk_window are never read */
      clk_avg = v; }
   RT_FHD void addbit(int bit) {                                  /* gcr_addbit, decode_gcr.c:731-787 (the per-track part) */
      datablock = 1;
      if (datacount < RT_MAXBLOCK) { bit_m2 = bit_m1; bit_m1 = bit; ++datacount; }
      lastbits = ((lastbits << 1) | bit) & 0xff;
      if (datacount % 5 == 0) {
         if ((lastbits & 0x1f) == RT_GCR_MARK2) resync_bitcount = 1;
         if ((lastbits & 0x1f) == RT_GCR_MARK1 && resync_bitcount > 0) resync_bitcount = 0; }
      if (resync_bitcount > 0) {
         if (resync_bitcount == 5) clk_force(t_peakdelta);
         ++resync_bitcount; } }

   /* process_*_transition + gcr_top/bot: the crossing armed at row `armed` is confirmed at row o */
   RT_FHD void transition(bool top, uint32_t armed) {
      const double t_ev = row_time(c, row0 + armed);
      /* the rows after the previous transition were tested for idleness (decoder.c:879-882) up to the row before this one */
      if (datablock && row_time(c, row0 + o - 1) > idle_after) datablock = 0;
      const float delta = (float)(t_ev - t_lastpeak);
      int numbits = 1;
      if (datablock) {                                            /* gcr_checkzeros, decode_gcr.c:789-834 */
         t_peakdeltaprev = t_peakdelta;
         t_peakdelta = delta;
         if (delta - t_pulse_adj > c.p.z1pt * clk_avg) {
            ++numbits; addbit(0);
            if (delta - t_pulse_adj > c.p.z2pt * clk_avg) { ++numbits; addbit(0); } }
         if (datacount > 3 && numbits == 1 && bit_m2) clk_adjust(t_peakdeltaprev);
         t_pulse_adj = c.p.pulse_adj * (numbits * clk_avg - delta); }
      addbit(1);
      t_lastpeak = t_ev;
      idle_after = t_lastpeak + RT_GCR_IDLE_THRESH * (double)clk_avg;
      if (pre) { commit(); pre = 0; }
      em.emit(row0 + o, t_ev, volts(c, vtop), volts(c, vbot), 1.0f, top); }

   RT_FHD void begin(const int16_t *plane_, uint64_t row0_, uint64_t row_end, int trk_, Emit em_, int /*quiet_thr_lsb*/) {
      plane = plane_; row0 = row0_; end = (uint32_t)(row_end - row0_); trk = trk_; delay = c.skew[trk_]; em = em_;
      const bool tz = row_time(c, row0) == 0.0;
      io = (uint32_t)trk + (tz ? 1u : 0u);
      /* I_ZC: the smallest sample whose voltage exceeds ZEROCROSS_PEAK (volts() is monotone and odd) */
      { int lo = 0, hi = 32768; while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (volts(c, mid) > RT_ZEROCROSS_PEAK) hi = mid; else lo = mid; } izc = hi; }
      vtop = vbot = vprev = 0; upp = dnp = 0; ttop = tbot = 0;
      clk_avg = c.clk_init; for (int i = 0; i < c.p.clk_window; ++i) mem.CLK(i) = c.clk_init;
      t_peakdelta = t_peakdeltaprev = t_pulse_adj = 0; clk_ndx = datacount = resync_bitcount = datablock = lastbits = bit_m1 = bit_m2 = 0;
      t_lastpeak = 0; idle_after = 0; pre = 1;
      /* proof data */
      qL = 1 + delay; ll = last_canon = OFF_NONE;
      sync_row = loud_at_sync = sync_first = sync_early = loud_early = OFF_NONE; early_frozen = false;
      const int lead = (int)io > delay ? (int)io : delay;
      own_fill = (uint32_t)(lead + 2); sf_from = own_fill;
      const int32_t npre = row0 > (uint64_t)c.prescan_rows ? c.prescan_rows : (int32_t)row0;
      for (int32_t off = -npre; off < 0; ++off) { const int raw = (int)plane[(int64_t)row0 + off]; if (raw >= izc || raw <= -izc) loud_until(off + qL - 1); }
      quiet_from = ll == OFF_NONE ? row0 - (uint64_t)npre : (uint64_t)((int64_t)row0 + ll + 1);
      for (uint32_t off = 0; off <= io && off < end; ++off) track((int32_t)off, (int)plane[row0 + off]);     /* not looked at yet (decoder.c:855-861) */
      o = io + 1;
      st = o >= end ? ST_DONE : ST_RUN;
      if (st == ST_RUN) {
         t_lastpeak = row_time(c, row0 + io);                    /* decoder.c:858 */
         ld = o < (uint32_t)delay ? o : (o - (uint32_t)delay) / 8 * 8 + (uint32_t)delay; pf_at = 0xffffffffu;
         ensure(o + FAST_K); } }

   RT_FHD void next_row(int x) { vprev = x; ++o; if (o >= end) st = ST_DONE; }

   /* one row of lookfor_zerocrossing, decoder.c:617-649 */
   RT_FHD void step() {
      const int x = mem.X(o);
      if (x > 0) {
         dnp = 0;
         if (vtop < x) {
            vtop = x;
            if (upp && x >= izc) { st = ST_PEND; return; } }
         if (vprev < 0 && vbot <= -izc) { ttop = o; upp = 1; } }
      else if (x < 0) {
         upp = 0;
         if (vbot > x) {
            vbot = x;
            if (dnp && x <= -izc) { st = ST_PEND; return; } }
         if (vprev > 0 && vtop >= izc) { tbot = o; dnp = 1; } }
      if (pre) track((int32_t)o, delay ? (int)plane[row0 + o] : x);
      next_row(x); }

   /* a pending crossing has been confirmed at row o (the extreme is already updated): the slope test of decoder.c:629/:643
      decides whether it is reported; then the rest of the row */
   RT_FHD void handle() {
      const int x = mem.X(o);
      const bool top = x > 0;
      const uint32_t armed = top ? ttop : tbot;
      if (top) { upp = 0; vbot = 0; } else { dnp = 0; vtop = 0; }
      const bool fire = row_time(c, row0 + o) - row_time(c, row0 + armed) <= (double)(clk_avg * RT_ZEROCROSS_SLOPE);
      if (fire) transition(top, armed);
      /* the arming test of the same row: the opposite extreme has just been cleared, so it cannot arm (decoder.c:634/:648) */
      if (!fire && pre) track((int32_t)o, delay ? (int)plane[row0 + o] : x);
      st = ST_RUN;
      next_row(x); }

   RT_FHD void finish(TrkMeta &meta) {
      if (pre) commit();
      meta.first_event_row = em.first_row;
      meta.sync_row = sync_row == OFF_NONE ? RT_NOROW : row0 + (uint64_t)sync_row;
      meta.last_loud_row = loud_at_sync == OFF_NONE ? RT_NOROW : (uint64_t)((int64_t)row0 + loud_at_sync);
      meta.sync_first = sync_first == OFF_NONE ? RT_NOROW : row0 + (uint64_t)sync_first;
      meta.quiet_from = quiet_from;
      meta.sync_early = sync_early == OFF_NONE ? RT_NOROW : row0 + (uint64_t)sync_early;
      meta.loud_early = loud_early == OFF_NONE ? RT_NOROW : (uint64_t)((int64_t)row0 + loud_early);
      meta.last_event_row = em.last_row;
      meta.quiet_tail_from = RT_NOROW;
      if (row0 + end >= c.nrows) {  /* the tail rule (the unit that ends the tape): the last sample beyond ZEROCROSS_PEAK keeps the rows up to qL - 1 behind it loud */
         const uint32_t lo = end > (uint32_t)c.prescan_rows ? end - (uint32_t)c.prescan_rows : 0u;
         int64_t scan_lo = (int64_t)lo - (qL - 1); if ((int64_t)row0 + scan_lo < 0) scan_lo = -(int64_t)row0;
         int64_t r = (int64_t)end - 1;
         for (; r >= scan_lo; --r) { const int x = (int)plane[(int64_t)row0 + r]; if (x >= izc || x <= -izc) break; }
         const uint64_t q = r < scan_lo ? row0 + lo : (uint64_t)((int64_t)row0 + r + qL);
         meta.quiet_tail_from = q > row0 + end ? row0 + end : q; }
      meta.first_chunk = em.first_chunk; meta.nevents = em.n; meta.failed = 0; meta.pad = 0; } };

}  // namespace rtfast
