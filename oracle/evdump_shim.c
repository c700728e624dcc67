/* oracle/evdump_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Link-time instrumentation of the UNMODIFIED reference (readtape 3.18) that dumps a
 * fine-grained per-event log.  Nothing in the reference sources is changed: the shim is
 * linked in with  -Wl,--wrap=process_sample,--wrap=init_trackstate,...  so that cross-TU
 * calls go through the __wrap_ functions below (see oracle/Makefile).
 *
 * What is recorded (binary, fixed 64-byte records, file named by $RT_EVDUMP):
 *   RESET   : init_trackstate()/ww_init_blockstate() was called; next row to be read, parmset
 *   EVENT   : a flux transition handed to a mode handler ({nrzi,pe,gcr,ww}_{top,bot}),
 *             with the values the handler sees and the AGC/clock state after it returns
 *   DENS    : (density-detection / any mode) a transition seen only via peakcount change
 *   BLKEND  : process_sample() returned != BS_NONE (readblock() returns after this row)
 *   CFG     : (after every RESET) the globals the scan depends on: mode, bpi, ips, flags, skew
 *   TRKF    : (after every RESET, and after compute_avg_height) per-track v_avg_height
 *   IBG     : interblock_counter became non-zero on this row (end-of-block fired here)
 *
 * Reference entry points intercepted (file:line in /root/reference/src):
 *   process_sample            decoder.c:817   (called from readtape.c:1504)
 *   init_trackstate           decoder.c:425   (called from readtape.c:1665,1693,1760,1859)
 *   ww_init_blockstate        decode_ww.c:33  (called from readtape.c:1692,1759)
 *   init_trackpeak_state      decoder.c:413   (called from readtape.c:1707)
 *   compute_avg_height        decoder.c:491   (called from readtape.c:1713)
 *   nrzi_top/bot              decode_nrzi.c:199/184   (called from decoder.c:584,602)
 *   pe_top/bot                decode_pe.c:157/180
 *   gcr_top/bot               decode_gcr.c:846/836
 *   ww_top/bot                decode_ww.c:263/253
 */
#include "decoder.h"

extern long long numsamples;          /* readtape.c:498 */
extern bool doing_density_detection, doing_deskew;

struct evrec {                 /* 64 bytes, little-endian, packed by construction */
   uint8_t  type;              /* 1=RESET 2=EVENT 3=BLKEND 4=DENS */
   uint8_t  trk;
   uint8_t  kind;              /* EVENT/DENS: 1=top/up 0=bot/down ; RESET: 0=full 1=ww-partial ; BLKEND: bstate */
   uint8_t  flags;             /* bit0 doing_density_detection, bit1 doing_deskew */
   int32_t  parmset;
   uint64_t row;               /* row index of the sample being processed (0-based, after -skip) */
   double   t_event;           /* t_top or t_bot as seen by the handler */
   float    v_top, v_bot;      /* as seen by the handler */
   float    agc_pre, agc_post; /* t->agc_gain before / after the handler */
   float    avg_height_post;   /* t->v_avg_height after the handler */
   float    clkavg_post;       /* t->clkavg.t_bitspaceavg after the handler (PE/GCR per-track clock) */
   int32_t  peakcount;         /* t->peakcount as seen by the handler (already incremented) */
   int32_t  pkww_width;        /* window width in force */
   double   timenow;           /* global timenow at the moment of the record */
};

struct cfgrec {                /* type 5, 64 bytes */
   uint8_t  type, ntrks, find_zeros, differentiate;
   int32_t  parmset;
   uint64_t row;
   float    bpi, ips;
   int32_t  mode;
   int8_t   skew[MAXTRKS];     /* 19 */
   uint8_t  invert;
   uint8_t  pad[16]; };
struct trkfrec {               /* type 6, 64 bytes: v_avg_height of tracks 0..9 */
   uint8_t  type, ntrks, pad0[2];
   int32_t  parmset;
   uint64_t row;
   float    avg_height[10];
   uint8_t  pad[8]; };
struct headrec {               /* type 8, 64 bytes: the sample-stream description, once per file */
   uint8_t  type, nheads, ntrks, pad0;
   int32_t  subsample;
   uint64_t tstart_ns;
   uint64_t tdelta_ns;
   float    maxvolts;
   int8_t   head_to_trk[MAXTRKS];  /* 19 */
   uint8_t  pad[17]; };
typedef char headrec_is_64[sizeof(struct headrec) == 64 ? 1 : -1];
extern int head_to_trk[MAXTRKS], nheads, subsample;
extern struct tbin_hdr_t tbin_hdr;
extern struct tbin_dat_t tbin_dat;
typedef char cfgrec_is_64[sizeof(struct cfgrec) == 64 ? 1 : -1];
typedef char trkfrec_is_64[sizeof(struct trkfrec) == 64 ? 1 : -1];
typedef char evrec_is_64[sizeof(struct evrec) == 64 ? 1 : -1];

extern bool invert_data, find_zeros, do_differentiate;
static FILE *evf;
static int ev_inited;

static void ev_open(void) {
   if (ev_inited) return;
   ev_inited = 1;
   const char *name = getenv("RT_EVDUMP");
   if (name && name[0]) evf = fopen(name, "wb"); }

static void ev_write(struct evrec *r) {
   ev_open();
   if (evf) fwrite(r, sizeof *r, 1, evf); }

static uint8_t ev_flags(void) {
   return (uint8_t)((doing_density_detection ? 1 : 0) | (doing_deskew ? 2 : 0)); }

/* ---- resets -------------------------------------------------------------------------- */
static void ev_trkf(void) {
   struct trkfrec f; memset(&f, 0, sizeof f);
   f.type = 6; f.ntrks = (uint8_t)ntrks; f.parmset = block.parmset; f.row = (uint64_t)numsamples;
   for (int i = 0; i < ntrks && i < 10; ++i) f.avg_height[i] = trkstate[i].v_avg_height;
   ev_open(); if (evf) fwrite(&f, sizeof f, 1, evf); }

static void ev_reset(int kind) {
   static int said_heads;
   ev_open();
   if (!said_heads && evf) {
      struct headrec h; memset(&h, 0, sizeof h);
      h.type = 8; h.nheads = (uint8_t)nheads; h.ntrks = (uint8_t)ntrks; h.subsample = subsample;
      h.tstart_ns = tbin_dat.tstart; h.tdelta_ns = (uint64_t)sample_deltat_ns; h.maxvolts = tbin_hdr.u.s.maxvolts;
      for (int i = 0; i < MAXTRKS; ++i) h.head_to_trk[i] = (int8_t)(i < nheads ? head_to_trk[i] : MAXTRKS - 1);
      fwrite(&h, sizeof h, 1, evf);
      said_heads = 1; }
   struct evrec r; memset(&r, 0, sizeof r);
   r.type = 1; r.kind = (uint8_t)kind; r.flags = ev_flags(); r.parmset = block.parmset;
   r.row = (uint64_t)numsamples; r.timenow = timenow; r.pkww_width = pkww_width;
   ev_write(&r);
   struct cfgrec c; memset(&c, 0, sizeof c);
   c.type = 5; c.ntrks = (uint8_t)ntrks; c.find_zeros = find_zeros; c.differentiate = do_differentiate;
   c.parmset = block.parmset; c.row = (uint64_t)numsamples; c.bpi = bpi; c.ips = ips; c.mode = (int32_t)mode;
   for (int i = 0; i < MAXTRKS; ++i) c.skew[i] = (int8_t)skew_delaycnt[i];
   c.invert = invert_data;
   if (evf) fwrite(&c, sizeof c, 1, evf);
   ev_trkf(); }

void __real_init_trackstate(void);
void __wrap_init_trackstate(void) { __real_init_trackstate(); ev_reset(0); }

void __real_ww_init_blockstate(void);
void __wrap_ww_init_blockstate(void) { __real_ww_init_blockstate(); ev_reset(1); }

void __real_init_trackpeak_state(void);
void __wrap_init_trackpeak_state(void) { __real_init_trackpeak_state(); ev_reset(2); }

void __real_compute_avg_height(struct trkstate_t *t);
void __wrap_compute_avg_height(struct trkstate_t *t) {
   __real_compute_avg_height(t);
   if (t->trknum == ntrks - 1) ev_trkf(); }

/* ---- handlers ------------------------------------------------------------------------ */
static void ev_handler(struct trkstate_t *t, int is_top, void (*real)(struct trkstate_t *)) {
   struct evrec r; memset(&r, 0, sizeof r);
   r.type = 2; r.trk = (uint8_t)t->trknum; r.kind = (uint8_t)is_top; r.flags = ev_flags();
   r.parmset = block.parmset; r.row = (uint64_t)(numsamples - 1);
   r.t_event = is_top ? t->t_top : t->t_bot;
   r.v_top = t->v_top; r.v_bot = t->v_bot;
   r.agc_pre = t->agc_gain; r.peakcount = t->peakcount; r.pkww_width = pkww_width;
   r.timenow = timenow;
   real(t);
   r.agc_post = t->agc_gain; r.avg_height_post = t->v_avg_height;
   r.clkavg_post = t->clkavg.t_bitspaceavg;
   ev_write(&r); }

#define WRAP_HANDLER(name, is_top) \
   void __real_##name(struct trkstate_t *t); \
   void __wrap_##name(struct trkstate_t *t) { ev_handler(t, is_top, __real_##name); }
WRAP_HANDLER(nrzi_top, 1) WRAP_HANDLER(nrzi_bot, 0)
WRAP_HANDLER(pe_top, 1)   WRAP_HANDLER(pe_bot, 0)
WRAP_HANDLER(gcr_top, 1)  WRAP_HANDLER(gcr_bot, 0)
WRAP_HANDLER(ww_top, 1)   WRAP_HANDLER(ww_bot, 0)

/* ---- the per-row entry point ----------------------------------------------------------- */
enum bstate_t __real_process_sample(struct sample_t *sample);
enum bstate_t __wrap_process_sample(struct sample_t *sample) {
   int before[MAXTRKS];
   int dens = doing_density_detection;
   int ibg_before = interblock_counter;
   if (dens) for (int i = 0; i < ntrks; ++i) before[i] = trkstate[i].peakcount;
   enum bstate_t bs = __real_process_sample(sample);
   if (dens) /* handlers are bypassed (decoder.c:578-580): recover events from peakcount changes */
      for (int i = 0; i < ntrks; ++i) if (trkstate[i].peakcount != before[i]) {
            struct trkstate_t *t = &trkstate[i];
            struct evrec r; memset(&r, 0, sizeof r);
            int is_top = (t->t_lastpeak == t->t_top);
            r.type = 4; r.trk = (uint8_t)i; r.kind = (uint8_t)is_top; r.flags = ev_flags();
            r.parmset = block.parmset; r.row = (uint64_t)(numsamples - 1);
            r.t_event = t->t_lastpeak; r.v_top = t->v_top; r.v_bot = t->v_bot;
            r.agc_pre = r.agc_post = t->agc_gain; r.avg_height_post = t->v_avg_height;
            r.clkavg_post = t->clkavg.t_bitspaceavg;
            r.peakcount = t->peakcount; r.pkww_width = pkww_width; r.timenow = timenow;
            ev_write(&r); }
   if (ibg_before == 0 && (interblock_counter != 0 || (bs != BS_NONE && mode != WW))) {
      /* end-of-block processing fired on this row: rows after it are skipped (decoder.c:841,900-903) */
      struct evrec r; memset(&r, 0, sizeof r);
      r.type = 7; r.flags = ev_flags(); r.parmset = block.parmset;
      r.row = (uint64_t)(numsamples - 1); r.timenow = timenow; r.pkww_width = interblock_counter;
      ev_write(&r); }
   if (bs != BS_NONE) {
      struct evrec r; memset(&r, 0, sizeof r);
      r.type = 3; r.kind = (uint8_t)bs; r.flags = ev_flags(); r.parmset = block.parmset;
      r.row = (uint64_t)(numsamples - 1); r.timenow = timenow; r.pkww_width = pkww_width;
      ev_write(&r);
      if (evf) fflush(evf); }
   return bs; }
