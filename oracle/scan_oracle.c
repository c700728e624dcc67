/* oracle/scan_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into the product path.
 *
 * A plain-C, single-threaded CPU restatement of the per-sample, per-track part of readtape's
 * decoder (reference readtape 3.18 @ /root/reference/src), exported behind the same C-ABI as
 * the CUDA library (include/rt_scan.h) so that tests can diff the two event streams 1:1.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * PARITY PINS: this restatement is checked event-for-event (row, track, polarity, t_event,
 * v_top, v_bot, agc_gain -- bit-exact) against logs dumped from the unmodified reference by
 * oracle/evdump_shim.c on the reference's bundled examples (tests/test_oracle_golden.py, and the
 * committed digests in tests/golden/).  The reference itself reproduces all of its golden
 * .tap/.bin files here (tests/golden/reference_goldens.json, tests/golden/full_outputs.json).
 *
 * What is restated, and where it lives in the reference:
 *   int16 -> volts, invert, differentiate     readtape.c:1418-1422, 1383-1394
 *   deskew FIFO                               decoder.c:819-831 (struct skew_t :227-231)
 *   first-sample initialisation + `break`     decoder.c:855-861
 *   moving-window peak detector               decoder.c:751-810  (lookfor_peak)
 *   peak time refinement                      decoder.c:700-749  (refine_peak)
 *   zero-crossing detectors                   decoder.c:617-649, 654-683
 *   per-event glue                            decoder.c:560-609  (process_*_transition)
 *   AGC                                       decoder.c:500-531  (adjust_agc)
 *   clock averaging                           decoder.c:533-558  (adjust_clock/force_clock)
 *   per-track feedback fragments of the mode handlers:
 *      NRZI  decode_nrzi.c:184-230     PE   decode_pe.c:127-201
 *      GCR   decode_gcr.c:731-865      WW   decode_ww.c:167-191
 *   GCR per-row idle test                     decoder.c:879-882
 *   resets                                    decoder.c:413-455, decode_ww.c:33-49
 *   window width / samples-per-bit            readtape.c:1453-1457, 1402
 *
 * What is NOT here (host side, cross-track): bit assembly, nrzi_zerocheck, end-of-block
 * detection, interblock_counter, parity/CRC/ECC, file output.
 *
 * Floating point: compiled with -O2 -ffp-contract=off on x86-64 (FLT_EVAL_METHOD 0), i.e.
 * float expressions are evaluated in float, double in double, exactly like the reference
 * build; the float/double type of every sub-expression follows the reference source.
 */
#define _XOPEN_SOURCE 700
#include <stdlib.h>
#include <unistd.h>
#include <sys/types.h>
#include <string.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include "rt_scan.h"

#define PKWW_PEAKHEIGHT  4.0f     /* decoder.h:133 */
#define DIFF_THRESHOLD   0.05f    /* decoder.h:135 */
#define DIFF_SCALE       0.4f     /* decoder.h:136 */
#define ZEROCROSS_PEAK   0.2f     /* decoder.h:138 */
#define ZEROCROSS_SLOPE  1.5f     /* decoder.h:139 */
#define PEAK_THRESHOLD   0.005f   /* decoder.h:141 */
#define AGC_MAX_VALUE    2.0f     /* decoder.h:153 */
#define AGC_STARTBASE    5        /* decoder.h:154 */
#define AGC_ENDBASE      15       /* decoder.h:155 */
#define GCR_IDLE_THRESH  6.00     /* decoder.h:111 (double) */
#define PE_MIN_PREBITS   70       /* decoder.h:118 */
#define GCR_MARK1        0x07     /* decode_gcr.c:422 */
#define GCR_MARK2        0x1c     /* decode_gcr.c:423 */

static char g_err[512];
static int set_err(int code, const char *fmt, ...) {
   va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
   return code; }
const char *rt_last_error(void) { return g_err; }
__attribute__((visibility("hidden"))) int oracle_fail(int code, const char *fmt, ...) {      /* for csv_oracle.c */
   va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
   return code; }
int rt_abi_version(void) { return RT_ABI_VERSION; }
const char *rt_backend(void) { return "oracle-cpu"; }

/* ------------------------------------------------------------------------------------- */
struct rt_tape {
   rt_tape_desc desc;
   int16_t *rows;        /* interleaved, nheads per row */
   uint64_t nrows;       /* rows before the end marker */
   uint64_t cap; };

struct clk { float spacing[RT_CLKRATE_WINDOW]; int ndx; float avg; };   /* struct clkavg_t, decoder.h:185 */

struct trk {             /* the part of struct trkstate_t (+ struct skew_t) that feeds the detectors */
   /* deskew FIFO */
   float vdelayed[RT_MAXSKEWSAMP]; int ndx_next, slots_filled;
   float v_last_raw;
   float v_now, v_prev;
   float v_top, v_bot, v_lasttop, v_lastbot;
   double t_top, t_bot, t_lastpeak, t_prevlastpeak;
   unsigned char up_pending, dn_pending;
   double t_firstzero, t_lastzero;
   /* window */
   float win[RT_PKWW_MAX_WIDTH]; float minv, maxv; int left, right, countdown;
   /* AGC */
   float avg_height, avg_height_sum; int avg_height_count;
   float agc_gain; float heights[RT_AGC_MAX_WINDOW]; int heightndx;
   int peakcount;
   /* PE */
   unsigned char datablock, bit1_up; float t_clkwindow;
   /* GCR */
   struct clk clk; float t_peakdelta, t_peakdeltaprev, t_pulse_adj;
   int datacount; unsigned char lastbits, bit_m1, bit_m2; int resync_bitcount; };

struct rt_scan {
   rt_tape *tape; rt_scan_cfg cfg;
   int width, samples_per_bit; float sample_deltat;
   uint64_t pos; int positioned;
   struct trk trk[RT_MAXTRKS];
   struct trk ckpt[RT_MAXTRKS]; uint64_t ckpt_pos, ckpt_end; int have_ckpt;
   rt_event *ev; uint64_t nev, capev;
   int failed; };

double rt_row_time(const rt_tape_desc *d, uint64_t row) {
   int64_t ns = (int64_t)(d->tstart_ns + row * d->tdelta_ns);
   return (double)ns / 1e9; }

int rt_pkww_width(const rt_scan_cfg *cfg, uint64_t tdelta_ns) {
   float sample_deltat = (float)(int64_t)tdelta_ns / 1e9f;             /* readtape.c:1345 */
   if (cfg->bpi == 0 || (cfg->flags & RT_F_DENSITY_DETECT)) return 8;   /* readtape.c:1457 */
   int w = (int)(cfg->parms.pkww_bitfrac / (cfg->bpi * cfg->ips * sample_deltat));
   return w < RT_PKWW_MAX_WIDTH ? w : RT_PKWW_MAX_WIDTH; }

/* ---- tape ------------------------------------------------------------------------------ */
int rt_open(const rt_tape_desc *desc, int device, rt_tape **out) {
   (void)device;
   if (!desc || !out) return set_err(RT_ERR_ARG, "rt_open: null argument");
   if (desc->ntrks < 1 || desc->ntrks > RT_MAXTRKS || desc->nheads < desc->ntrks || desc->nheads > RT_MAXTRKS)
      return set_err(RT_ERR_ARG, "rt_open: bad ntrks/nheads %u/%u", desc->ntrks, desc->nheads);
   rt_tape *t = calloc(1, sizeof *t);
   if (!t) return set_err(RT_ERR_NOMEM, "rt_open: out of memory");
   t->desc = *desc; *out = t; return RT_OK; }

int rt_upload(rt_tape *t, const int16_t *rows, uint64_t nrows) {
   if (!t || (!rows && nrows)) return set_err(RT_ERR_ARG, "rt_upload: null argument");
   uint64_t nh = t->desc.nheads, have = t->cap;
   int16_t *p = realloc(t->rows, (size_t)((have + nrows) * nh * 2 + 2));
   if (!p) return set_err(RT_ERR_NOMEM, "rt_upload: out of memory");
   t->rows = p;
   memcpy(p + have * nh, rows, (size_t)(nrows * nh * 2));
   t->cap = have + nrows;
   if (t->nrows == have) {           /* no end marker seen so far: look for one in the new rows */
      uint64_t r = have;
      while (r < t->cap && p[r * nh] != RT_TBIN_END_MARK) ++r;
      t->nrows = r; }
   return RT_OK; }

int rt_upload_fd(rt_tape *t, int fd, uint64_t offset, uint64_t n) {
   if (!t || fd < 0) return RT_ERR_ARG;
   int16_t *buf = malloc((size_t)n * t->desc.nheads * 2 + 2);
   if (!buf) return RT_ERR_NOMEM;
   size_t want = (size_t)n * t->desc.nheads * 2, got = 0;
   while (got < want) { ssize_t k = pread(fd, (char *)buf + got, want - got, (off_t)(offset + got)); if (k <= 0) { free(buf); return RT_ERR_ARG; } got += (size_t)k; }
   int rc = rt_upload(t, buf, n);
   free(buf);
   return rc; }
int rt_attach_device(rt_tape *t, const void *p, uint64_t n) {
   (void)t; (void)p; (void)n; return set_err(RT_ERR_UNSUPPORTED, "oracle has no device memory"); }
uint64_t rt_nrows(const rt_tape *t) { return t ? t->nrows : 0; }
void rt_close(rt_tape *t) { if (t) { free(t->rows); free(t); } }
void *rt_host_alloc(size_t n) { return malloc(n); }
void rt_host_free(void *p) { free(p); }

/* ---- clock averaging, decoder.c:407-411, 533-558 ----------------------------------------- */
static void clk_init(struct clk *c, float v) {
   c->avg = v; c->ndx = 0;
   for (int i = 0; i < RT_CLKRATE_WINDOW; ++i) c->spacing[i] = v; }

static void clk_adjust(const rt_scan *s, struct clk *c, float delta) {
   int win = s->cfg.parms.clk_window; float alpha = s->cfg.parms.clk_alpha;
   if (win > 0) {                     /* incremental moving average, roundoff and all (Q4) */
      float old = c->spacing[c->ndx];
      c->spacing[c->ndx] = delta;
      if (++c->ndx >= win) c->ndx = 0;
      c->avg += (delta - old) / win; }
   else if (alpha > 0) c->avg = alpha * delta + (1 - alpha) * c->avg;
   else c->avg = (s->cfg.mode & (RT_MODE_PE + RT_MODE_WW)) ? 1 / (s->cfg.bpi * s->cfg.ips)
                 : 0.0f; /* nrzi.clkavg.t_bitspaceavg: never initialised outside NRZI mode (decoder.c:450-452) */ }

static void clk_force(struct clk *c, float v) {
   for (int i = 0; i < RT_CLKRATE_WINDOW; ++i) c->spacing[i] = v;
   c->avg = v; }

/* ---- AGC, decoder.c:500-531 -------------------------------------------------------------- */
static void agc_adjust(const rt_scan *s, struct trk *t) {
   const rt_parms *p = &s->cfg.parms;
   if (s->cfg.flags & RT_F_FIND_ZEROS) return;
   float gain, lastheight;
   if (p->agc_alpha) {
      lastheight = t->v_lasttop - t->v_lastbot;
      if (lastheight > 0) {
         gain = t->avg_height / lastheight;
         gain = p->agc_alpha * gain + (1 - p->agc_alpha) * t->agc_gain;
         if (gain > AGC_MAX_VALUE) gain = AGC_MAX_VALUE;
         t->agc_gain = gain; } }
   if (p->agc_window) {
      lastheight = t->v_lasttop - t->v_lastbot;
      if (lastheight > 0) {
         t->heights[t->heightndx] = lastheight;
         if (++t->heightndx >= p->agc_window) t->heightndx = 0;
         float minheight = 99;
         for (int i = 0; i < p->agc_window; ++i) if (t->heights[i] < minheight) minheight = t->heights[i];
         gain = t->avg_height / minheight;
         if (gain > AGC_MAX_VALUE) gain = AGC_MAX_VALUE;
         t->agc_gain = gain; } } }

static void baseline_accumulate(const rt_scan *s, struct trk *t) {   /* peaks 5..15 */
   t->avg_height_sum += t->v_top - t->v_bot;
   ++t->avg_height_count;
   t->heights[t->heightndx] = t->v_top - t->v_bot;
   if (++t->heightndx >= s->cfg.parms.agc_window) t->heightndx = 0; }

/* ---- per-track feedback fragments of the mode handlers ------------------------------------ */
static void nrzi_feedback(const rt_scan *s, struct trk *t, int top) {      /* decode_nrzi.c:184-230 */
   if (top) {
      if (t->peakcount >= AGC_STARTBASE && t->peakcount <= AGC_ENDBASE) baseline_accumulate(s, t);
      else if (t->peakcount > AGC_ENDBASE) {
         if (t->avg_height_count) {
            t->avg_height = t->avg_height_sum / t->avg_height_count;
            t->avg_height_count = 0; }
         else agc_adjust(s, t); } }
   else if (t->peakcount > AGC_ENDBASE && t->avg_height_count == 0) agc_adjust(s, t); }

static void pe_feedback(const rt_scan *s, struct trk *t, int top, double t_ev) {  /* decode_pe.c:127-201 */
   if (t->datablock) { agc_adjust(s, t); return; }
   if (t->peakcount == 1) t->bit1_up = !top;
   if (t->peakcount > PE_MIN_PREBITS && t->bit1_up == top && t_ev - t->t_lastpeak > t->t_clkwindow) {
      t->datablock = 1;
      t->avg_height = t->avg_height_sum / t->avg_height_count; }
   else if (t->peakcount >= AGC_STARTBASE && t->peakcount <= AGC_ENDBASE && t->v_top > t->v_bot)
      baseline_accumulate(s, t); }

static void gcr_addbit(struct trk *t, int bit) {                           /* decode_gcr.c:731-787 */
   t->datablock = 1;
   if (t->datacount < RT_MAXBLOCK) { t->bit_m2 = t->bit_m1; t->bit_m1 = (unsigned char)bit; ++t->datacount; }
   t->lastbits = (unsigned char)((t->lastbits << 1) | bit);
   if (t->datacount % 5 == 0) {
      if ((t->lastbits & 0x1f) == GCR_MARK2) t->resync_bitcount = 1;
      if ((t->lastbits & 0x1f) == GCR_MARK1 && t->resync_bitcount > 0) t->resync_bitcount = 0; }
   if (t->resync_bitcount > 0) {
      if (t->resync_bitcount == 5) clk_force(&t->clk, t->t_peakdelta);
      ++t->resync_bitcount; } }

static void gcr_feedback(const rt_scan *s, struct trk *t, int top, double t_ev) { /* decode_gcr.c:789-865 */
   const rt_parms *p = &s->cfg.parms;
   float delta = (float)(t_ev - t->t_lastpeak);
   int numbits = 1;
   if (t->datablock) {                                                     /* gcr_checkzeros */
      t->t_peakdeltaprev = t->t_peakdelta;
      t->t_peakdelta = delta;
      if (delta - t->t_pulse_adj > p->z1pt * t->clk.avg) {
         ++numbits; gcr_addbit(t, 0);
         if (delta - t->t_pulse_adj > p->z2pt * t->clk.avg) { ++numbits; gcr_addbit(t, 0); } }
      if (t->datacount > 3 && numbits == 1 && t->bit_m2) clk_adjust(s, &t->clk, t->t_peakdeltaprev);
      t->t_pulse_adj = p->pulse_adj * (numbits * t->clk.avg - delta); }
   gcr_addbit(t, 1);
   nrzi_feedback(s, t, top); /* the AGC part of gcr_top/gcr_bot has the same form as NRZI's */ }

/* ---- per-event glue, decoder.c:560-609 ------------------------------------------------- */
static void push_event(rt_scan *s, uint64_t row, int trknum, int top, double t_ev, float v_top, float v_bot, float agc) {
   if (s->nev == s->capev) {
      uint64_t nc = s->capev ? s->capev * 2 : 4096;
      rt_event *p = realloc(s->ev, (size_t)(nc * sizeof *p));
      if (!p) { s->failed = 1; return; }
      s->ev = p; s->capev = nc; }
   rt_event *e = &s->ev[s->nev++];
   memset(e, 0, sizeof *e);
   e->row = row; e->t_event = t_ev;
   e->v_top = v_top; e->v_bot = v_bot; e->agc_gain = agc;
   e->trk = (uint8_t)trknum; e->kind = (uint8_t)(top ? RT_EV_TOP : RT_EV_BOT); }

static void transition(rt_scan *s, struct trk *t, int trknum, int top, uint64_t row) {
   double t_ev = top ? t->t_top : t->t_bot;
   float v_top_seen = t->v_top, v_bot_seen = t->v_bot;
   ++t->peakcount;
   if (!(s->cfg.flags & RT_F_DENSITY_DETECT))
      switch (s->cfg.mode) {
      case RT_MODE_NRZI: nrzi_feedback(s, t, top); break;
      case RT_MODE_PE:   pe_feedback(s, t, top, t_ev); break;
      case RT_MODE_GCR:  gcr_feedback(s, t, top, t_ev); break;
      case RT_MODE_WW:   agc_adjust(s, t); break; }      /* ww_pulse_start/end both call it, decode_ww.c:171,190 */
   if (top) t->v_lasttop = t->v_top; else t->v_lastbot = t->v_bot;
   t->t_prevlastpeak = t->t_lastpeak;
   t->t_lastpeak = t_ev;
   /* the event carries what the handler saw, and the gain after it ran */
   push_event(s, row, trknum, top, t_ev, v_top_seen, v_bot_seen, t->agc_gain); }

/* ---- moving-window peak detector, decoder.c:700-810 -------------------------------------- */
static double refine(rt_scan *s, struct trk *t, float val, int top, double timenow) {
   int w = s->width, left_distance = 1, prev = -1;
   float adj = 0;
   for (int ndx = t->left;;) {
      if (t->win[ndx] == val) {
         if (left_distance >= w || prev == -1) { s->failed = 2; return 0; }   /* reference: fatal() */
         int next = ndx + 1; if (next >= w) next = 0;
         if (top) {
            float edge = val - PEAK_THRESHOLD / t->agc_gain;
            if (t->win[prev] > edge && t->win[next] < edge) adj = -0.5;
            else if (t->win[next] > edge && t->win[prev] < edge) adj = +0.5; }
         else {
            float edge = val + PEAK_THRESHOLD / t->agc_gain;
            if (t->win[prev] < edge && t->win[next] > edge) adj = -0.5;
            else if (t->win[next] < edge && t->win[prev] > edge) adj = +0.5; }
         double time = timenow - ((float)(w - left_distance) - adj) * s->sample_deltat;
         t->countdown = left_distance;
         return time; }
      ++left_distance;
      if (ndx == t->right) break;
      prev = ndx;
      if (++ndx >= w) ndx = 0; }
   s->failed = 2;                                                             /* reference: fatal() */
   return 0; }

static void peak_step(rt_scan *s, struct trk *t, int trknum, double timenow, uint64_t row) {
   const rt_parms *p = &s->cfg.parms;
   int w = s->width;
   float leaving = 0;
   if (++t->right >= w) t->right = 0;
   if (t->right == t->left) {
      leaving = t->win[t->left];
      if (++t->left >= w) t->left = 0; }
   t->win[t->right] = t->v_now;
   if (t->v_now > t->maxv) t->maxv = t->v_now;
   /* (Q1) decoder.c:765 compares pkww_minv with itself: the running minimum is only ever
      refreshed by the rescan below. */
   if (leaving == t->maxv || leaving == t->minv) {
      float mx = -100, mn = +100;
      for (int ndx = t->left;;) {
         if (t->win[ndx] > mx) mx = t->win[ndx];
         if (t->win[ndx] < mn) mn = t->win[ndx];
         if (ndx == t->right) break;
         if (++ndx >= w) ndx = 0; }
      t->maxv = mx; t->minv = mn; }
   if (t->countdown) { --t->countdown; return; }
   float rise = p->pkww_rise * (t->avg_height / PKWW_PEAKHEIGHT) / t->agc_gain;
   float reqmin = p->min_peak * (t->avg_height / PKWW_PEAKHEIGHT) / t->agc_gain;
   float vl = t->win[t->left], vr = t->win[t->right];
   if (t->maxv > vl + rise && t->maxv > vr + rise && (reqmin == 0 || t->maxv > reqmin)) {
      t->v_top = t->maxv;
      t->t_top = refine(s, t, t->maxv, 1, timenow);
      transition(s, t, trknum, 1, row); }
   else if (t->minv < vl - rise && t->minv < vr - rise && (reqmin == 0 || t->minv < -reqmin)) {
      t->v_bot = t->minv;
      t->t_bot = refine(s, t, t->minv, 0, timenow);
      transition(s, t, trknum, 0, row); } }

/* ---- zero-crossing detectors, decoder.c:617-683 ------------------------------------------ */
static void zc_step(rt_scan *s, struct trk *t, int trknum, double timenow, uint64_t row) {
   if (t->v_now > 0) {
      t->dn_pending = 0;
      if (t->v_top < t->v_now) {
         t->v_top = t->v_now;
         if (t->up_pending && t->v_top > ZEROCROSS_PEAK) {
            if (t->t_top == 0) t->t_top = timenow;
            t->up_pending = 0;
            t->v_bot = 0;
            if (timenow - t->t_top <= t->clk.avg * ZEROCROSS_SLOPE) transition(s, t, trknum, 1, row); } }
      if (t->v_prev < 0 && t->v_bot < -ZEROCROSS_PEAK) { t->t_top = timenow; t->up_pending = 1; } }
   else if (t->v_now < 0) {
      t->up_pending = 0;
      if (t->v_bot > t->v_now) {
         t->v_bot = t->v_now;
         if (t->dn_pending && t->v_bot < -ZEROCROSS_PEAK) {
            if (t->t_bot == 0) t->t_bot = timenow;
            t->dn_pending = 0;
            t->v_top = 0;
            if (timenow - t->t_bot <= t->clk.avg * ZEROCROSS_SLOPE) transition(s, t, trknum, 0, row); } }
      if (t->v_prev > 0 && t->v_top > ZEROCROSS_PEAK) { t->t_bot = timenow; t->dn_pending = 1; } }
   t->v_prev = t->v_now; }

static void dzc_step(rt_scan *s, struct trk *t, int trknum, double timenow, uint64_t row) {
   if (t->v_now > 0) {
      if (t->v_top < t->v_now) t->v_top = t->v_now;
      if (t->up_pending) {
         t->t_top = t->t_firstzero > 0 ? (t->t_firstzero + t->t_lastzero) / 2 : timenow - s->sample_deltat / 2;
         t->up_pending = 0;
         t->t_firstzero = 0;
         transition(s, t, trknum, 1, row); }
      if (t->v_now > ZEROCROSS_PEAK) { t->dn_pending = 1; t->t_firstzero = 0; t->v_bot = 0; } }
   else if (t->v_now < 0) {
      if (t->v_bot > t->v_now) t->v_bot = t->v_now;
      if (t->dn_pending) {
         t->t_bot = t->t_firstzero > 0 ? (t->t_firstzero + t->t_lastzero) / 2 : timenow - s->sample_deltat / 2;
         t->dn_pending = 0;
         t->t_firstzero = 0;
         transition(s, t, trknum, 0, row); }
      if (t->v_now < -ZEROCROSS_PEAK) { t->up_pending = 1; t->t_firstzero = 0; t->v_top = 0; } }
   else {
      t->t_lastzero = timenow;
      if (t->t_firstzero == 0) t->t_firstzero = timenow; } }

/* ---- resets, decoder.c:413-455, decode_ww.c:33-49 ---------------------------------------- */
static void reset_peakstate(rt_scan *s) {
   for (uint32_t k = 0; k < s->tape->desc.ntrks; ++k) {
      struct trk *t = &s->trk[k];
      memset(t->vdelayed, 0, sizeof t->vdelayed); t->ndx_next = t->slots_filled = 0;
      t->left = t->right = 0; t->minv = t->maxv = 0; t->countdown = 0; } }

static void reset_full(rt_scan *s) {
   memset(s->trk, 0, sizeof s->trk);
   for (uint32_t k = 0; k < s->tape->desc.ntrks; ++k) {
      struct trk *t = &s->trk[k];
      t->agc_gain = 1.0;
      t->avg_height = PKWW_PEAKHEIGHT;
      if (!(s->cfg.flags & RT_F_DENSITY_DETECT)) clk_init(&t->clk, 1 / (s->cfg.bpi * s->cfg.ips));
      t->t_clkwindow = t->clk.avg / 2 * s->cfg.parms.clk_factor; } }

/* ---- scan context ------------------------------------------------------------------------ */
static int cfg_check(const rt_tape *tape, const rt_scan_cfg *cfg) {
   if (cfg->mode != RT_MODE_PE && cfg->mode != RT_MODE_NRZI && cfg->mode != RT_MODE_GCR && cfg->mode != RT_MODE_WW)
      return set_err(RT_ERR_ARG, "rt_scan_begin: bad mode %d", cfg->mode);
   if ((cfg->flags & RT_F_FIND_ZEROS) && cfg->mode == RT_MODE_PE)
      return set_err(RT_ERR_UNSUPPORTED, "-zeros with PE needs the PE bit clock in the scan; not supported");
   if (!(cfg->flags & RT_F_DENSITY_DETECT) && !(cfg->bpi > 0 && cfg->ips > 0))
      return set_err(RT_ERR_ARG, "rt_scan_begin: bpi/ips must be positive");
   for (uint32_t k = 0; k < tape->desc.ntrks; ++k)
      if (cfg->skew_delaycnt[k] < 0 || cfg->skew_delaycnt[k] > RT_MAXSKEWSAMP)
         return set_err(RT_ERR_ARG, "rt_scan_begin: bad skew delay for track %u", k);
   if (!(cfg->flags & RT_F_FIND_ZEROS) && rt_pkww_width(cfg, tape->desc.tdelta_ns) < 3)
      return set_err(RT_ERR_UNSUPPORTED, "peak window narrower than 3 samples (%d)", rt_pkww_width(cfg, tape->desc.tdelta_ns));
   return RT_OK; }

static void cfg_apply(rt_scan *s, const rt_scan_cfg *cfg) {
   s->cfg = *cfg;
   s->sample_deltat = (float)(int64_t)s->tape->desc.tdelta_ns / 1e9f;
   s->width = rt_pkww_width(cfg, s->tape->desc.tdelta_ns);
   s->samples_per_bit = cfg->bpi > 0 ? (int)(1 / (cfg->bpi * cfg->ips * s->sample_deltat)) : 20; } /* readtape.c:1402 */

int rt_scan_begin(rt_tape *tape, const rt_scan_cfg *cfg, rt_scan **out) {
   if (!tape || !cfg || !out) return set_err(RT_ERR_ARG, "rt_scan_begin: null argument");
   int rc = cfg_check(tape, cfg);
   if (rc) return rc;
   rt_scan *s = calloc(1, sizeof *s);
   if (!s) return set_err(RT_ERR_NOMEM, "rt_scan_begin: out of memory");
   s->tape = tape;
   cfg_apply(s, cfg);
   *out = s; return RT_OK; }

int rt_scan_set_cfg(rt_scan *s, const rt_scan_cfg *cfg) {
   if (!s || !cfg) return set_err(RT_ERR_ARG, "rt_scan_set_cfg: null argument");
   if (cfg->mode != s->cfg.mode) return set_err(RT_ERR_ARG, "rt_scan_set_cfg: the mode cannot change");
   int rc = cfg_check(s->tape, cfg);
   if (rc) return rc;
   cfg_apply(s, cfg);
   s->have_ckpt = 0;
   return RT_OK; }

int rt_scan_reset(rt_scan *s, int kind, uint64_t row) {
   if (!s) return set_err(RT_ERR_ARG, "rt_scan_reset: null");
   switch (kind) {
   case RT_RESET_NONE: break;
   case RT_RESET_FULL: reset_full(s); break;
   case RT_RESET_WW_PARTIAL:
      for (uint32_t k = 0; k < s->tape->desc.ntrks; ++k) s->trk[k].t_lastpeak = s->trk[k].t_prevlastpeak = 0;
      break;
   case RT_RESET_PEAKSTATE: reset_peakstate(s); break;
   default: return set_err(RT_ERR_ARG, "rt_scan_reset: bad kind %d", kind); }
   s->pos = row; s->positioned = 1; s->have_ckpt = 0;
   return RT_OK; }

static void scan_rows(rt_scan *s, uint64_t from, uint64_t to) {
   const rt_tape_desc *d = &s->tape->desc;
   const uint32_t nh = d->nheads, nt = d->ntrks;
   const int fz = (s->cfg.flags & RT_F_FIND_ZEROS) != 0, diff = (s->cfg.flags & RT_F_DIFFERENTIATE) != 0;
   const int inv = (s->cfg.flags & RT_F_INVERT) != 0;
   float volts[RT_MAXTRKS];
   for (uint64_t row = from; row < to; ++row) {
      const int16_t *r = s->tape->rows + row * nh;
      for (uint32_t h = 0; h < nh; ++h) {                      /* readtape.c:1418-1422 */
         int k = d->head_to_trk[h];
         if (k < 0 || k >= (int)nt) continue;                    /* unused Whirlwind head */
         float v = (float)r[h] / 32767 * d->maxvolts;
         if (inv) v = -v;
         if (diff) {                                             /* readtape.c:1383-1388 */
            struct trk *t = &s->trk[k];
            float delta = v - t->v_last_raw;
            if (delta < DIFF_THRESHOLD && delta > -DIFF_THRESHOLD) delta = 0;
            t->v_last_raw = v;
            v = delta * DIFF_SCALE * s->samples_per_bit; }
         volts[k] = v; }
      double timenow = rt_row_time(d, row);
      for (uint32_t k = 0; k < nt; ++k) {                        /* deskew, decoder.c:820-830 */
         struct trk *t = &s->trk[k];
         int delay = s->cfg.skew_delaycnt[k];
         if (delay == 0) t->v_now = volts[k];
         else {
            if (t->slots_filled < delay) { t->v_now = volts[k]; ++t->slots_filled; }
            else t->v_now = t->vdelayed[t->ndx_next];
            t->vdelayed[t->ndx_next] = volts[k];
            if (++t->ndx_next >= delay) t->ndx_next = 0; } }
      for (uint32_t k = 0; k < nt; ++k) {                        /* decoder.c:847-890 */
         struct trk *t = &s->trk[k];
         if (t->t_lastpeak == 0) {                               /* (Q2) initialise ONE track, then leave the loop */
            t->win[0] = t->v_now;
            t->maxv = t->minv = t->v_now;
            t->t_lastpeak = timenow;
            break; }
         if (fz) { if (diff) dzc_step(s, t, (int)k, timenow, row); else zc_step(s, t, (int)k, timenow, row); }
         else peak_step(s, t, (int)k, timenow, row);
         if (s->cfg.mode == RT_MODE_GCR && t->datablock
               && timenow > t->t_lastpeak + GCR_IDLE_THRESH * t->clk.avg)     /* decoder.c:879-882 */
            t->datablock = 0; } } }

int rt_scan_run(rt_scan *s, uint64_t nrows, const rt_event **events, uint64_t *nevents, uint64_t *rows_done) {
   if (!s) return set_err(RT_ERR_ARG, "rt_scan_run: null");
   if (!s->positioned) return set_err(RT_ERR_STATE, "rt_scan_run before rt_scan_reset");
   uint64_t end = s->pos + nrows;
   if (end > s->tape->nrows || end < s->pos) end = s->tape->nrows;
   if (end < s->pos) end = s->pos;
   memcpy(s->ckpt, s->trk, sizeof s->trk); s->ckpt_pos = s->pos; s->ckpt_end = end; s->have_ckpt = 1;
   s->nev = 0; s->failed = 0;
   scan_rows(s, s->pos, end);
   if (s->failed == 1) return set_err(RT_ERR_NOMEM, "rt_scan_run: out of memory for events");
   if (s->failed == 2) return set_err(RT_ERR_STATE, "peak not found in window: the reference would call fatal() here");
   if (rows_done) *rows_done = end - s->pos;
   s->pos = end;
   if (events) *events = s->ev;
   if (nevents) *nevents = s->nev;
   return RT_OK; }

int rt_scan_rewind(rt_scan *s, uint64_t row) {
   if (!s) return set_err(RT_ERR_ARG, "rt_scan_rewind: null");
   if (!s->have_ckpt || row < s->ckpt_pos || row > s->ckpt_end)
      return set_err(RT_ERR_STATE, "rt_scan_rewind: row outside the last scanned span");
   memcpy(s->trk, s->ckpt, sizeof s->trk);
   s->nev = 0; s->failed = 0;
   scan_rows(s, s->ckpt_pos, row);
   s->nev = 0;
   s->pos = row;
   return RT_OK; }

int rt_scan_set_avg_height(rt_scan *s, uint32_t trk, float v) {
   if (!s || trk >= s->tape->desc.ntrks) return set_err(RT_ERR_ARG, "rt_scan_set_avg_height: bad argument");
   s->trk[trk].avg_height = v; s->trk[trk].avg_height_count = 0; s->trk[trk].avg_height_sum = 0;
   return RT_OK; }

uint64_t rt_scan_pos(const rt_scan *s) { return s ? s->pos : 0; }
void rt_scan_end(rt_scan *s) { if (s) { free(s->ev); free(s); } }

/* ---- the speculative whole-tape scan is a property of the product library only ----------- */
int rt_bulk_scan(rt_tape *t, const rt_scan_cfg *c, uint32_t n, rt_bulk **o) {
   (void)t; (void)c; (void)n; (void)o;
   return set_err(RT_ERR_UNSUPPORTED, "the oracle only implements the exact scan"); }
int rt_bulk_lookup(rt_bulk *b, uint32_t c, uint64_t r, const rt_event **e, uint64_t *n, uint64_t *v) {
   (void)b; (void)c; (void)r; (void)e; (void)n; (void)v;
   return set_err(RT_ERR_UNSUPPORTED, "the oracle only implements the exact scan"); }
int rt_clear(rt_tape *t) { if (!t) return RT_ERR_ARG; t->nrows = 0; t->cap = 0; return RT_OK; }
int rt_bulk_fetch(rt_bulk *b) { (void)b; return RT_ERR_UNSUPPORTED; }
int rt_bulk_results_size(const rt_bulk *b, uint64_t *n) { (void)b; (void)n; return RT_ERR_UNSUPPORTED; }
int rt_bulk_results_to_device(const rt_bulk *b, void *d, uint64_t n) { (void)b; (void)d; (void)n; return RT_ERR_UNSUPPORTED; }
int rt_bulk_fetch_to(rt_bulk *b, void *p, size_t n) { (void)b; (void)p; (void)n; return RT_ERR_UNSUPPORTED; }
int rt_host_register(rt_tape *t, void *p, size_t n) { (void)t; (void)p; (void)n; return RT_OK; }
int rt_host_unregister(rt_tape *t, void *p) { (void)t; (void)p; return RT_OK; }

/* The two bit planes of the product's two-pass peak scan, by definition (include/rt_scan.h: rt_peak_masks).  The window of
   lookfor_peak (decoder.c:751-775) at row p holds samples p-w+1 .. p of the track; int16 -> volts (readtape.c:1420) is strictly
   monotone, so max / min / comparisons are taken on the raw samples. */
int rt_peak_masks(rt_tape *t, const rt_scan_cfg *cfg, float t0_frac, uint32_t *cand, uint32_t *acan, uint64_t wpt, int32_t *t0) {
   if (!t || !cfg || !cand || !acan) return set_err(RT_ERR_ARG, "rt_peak_masks: null argument");
   if ((cfg->flags & (RT_F_FIND_ZEROS | RT_F_DENSITY_DETECT | RT_F_INVERT | RT_F_DIFFERENTIATE)) || wpt * 32 < t->nrows)
      return set_err(RT_ERR_UNSUPPORTED, "rt_peak_masks: not the plain peak detector, or buffer too small");
   const int w = rt_pkww_width(cfg, t->desc.tdelta_ns);
   const float inv_lsb = 32767.0f / t->desc.maxvolts;
   const float q = cfg->parms.pkww_rise * inv_lsb * 0.999f - 2.0f;
   if (!(q > 0) || w < 3) return set_err(RT_ERR_UNSUPPORTED, "rt_peak_masks: no usable threshold");
   const int T = q > 70000.0f ? 70000 : (int)q;
   int T0 = (int)((float)T * t0_frac);
   if (T0 > 65535) T0 = 65535;
   if (T0 < 16) return set_err(RT_ERR_UNSUPPORTED, "rt_peak_masks: T0 < 16");
   if (t0) *t0 = T0;
   const uint32_t nh = t->desc.nheads;
   for (uint32_t k = 0; k < t->desc.ntrks; ++k) {
      memset(cand + (size_t)k * wpt, 0, (size_t)wpt * 4); memset(acan + (size_t)k * wpt, 0, (size_t)wpt * 4); }
   for (uint32_t h = 0; h < nh; ++h) {
      const int k = t->desc.head_to_trk[h];
      if (k < 0 || k >= (int)t->desc.ntrks) continue;
      uint32_t *c = cand + (size_t)k * wpt, *a = acan + (size_t)k * wpt;
      for (uint64_t p = (uint64_t)w; p < t->nrows; ++p) {
         int S = -32768, mn = 32767;
         for (uint64_t r = p - w + 1; r <= p; ++r) { int v = t->rows[r * nh + h]; if (v > S) S = v; if (v < mn) mn = v; }
         const int l = t->rows[(p - w + 1) * nh + h], r_ = t->rows[p * nh + h], lv = t->rows[(p - w) * nh + h];
         const int mx = l > r_ ? l : r_, mi = l < r_ ? l : r_;
         if (S - mx >= T0 || mi - mn >= T0) c[p >> 5] |= 1u << (p & 31);
         if (lv >= S) a[p >> 5] |= 1u << (p & 31); } }
   return RT_OK; }
int rt_bulk_unit_info(const rt_bulk *b, uint32_t c, uint64_t r, rt_unit_info *o) { (void)b; (void)c; (void)r; (void)o; return RT_ERR_UNSUPPORTED; }
int rt_bulk_scan_host(rt_tape *t, const int16_t *r, uint64_t n, const rt_scan_cfg *c, rt_bulk **o) { (void)t; (void)r; (void)n; (void)c; (void)o; return RT_ERR_UNSUPPORTED; }
int rt_bulk_unit_at(const rt_bulk *b, uint32_t c, uint64_t r, rt_unit_info *o) { (void)b; (void)c; (void)r; (void)o; return RT_ERR_UNSUPPORTED; }
int rt_bulk_get_stats(const rt_bulk *b, rt_bulk_stats *o) { (void)b; (void)o; return RT_ERR_UNSUPPORTED; }
void rt_bulk_free(rt_bulk *b) { (void)b; }
int rt_prepare(rt_tape *t, const rt_scan_cfg *c) { (void)t; (void)c; return RT_OK; }
int rt_set_option(int o, int v) { (void)o; (void)v; return RT_OK; }
int rt_bulk_last_unit(const rt_bulk *b, uint32_t c, uint64_t *r0, uint64_t *r1) { (void)b; (void)c; (void)r0; (void)r1; return RT_ERR_UNSUPPORTED; }
int rt_bulk_tile_digest(rt_bulk *b, uint32_t c, uint64_t p, uint64_t n, uint64_t *e, uint64_t *d, uint64_t *bt) {
   (void)b; (void)c; (void)p; (void)n; (void)e; (void)d; (void)bt; return RT_ERR_UNSUPPORTED; }
