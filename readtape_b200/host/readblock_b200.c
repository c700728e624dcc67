/* readtape_b200/host/readblock_b200.c -- the host side of the drop-in: a replacement for
 *
 *        bool readblock(bool retry)                 reference src/readtape.c:1396-1517
 *
 * that obtains flux-transition EVENTS from the rt_scan C-ABI (include/rt_scan.h: the CUDA library)
 * instead of reading rows and calling process_sample() (src/decoder.c:817) once per sample.
 * Everything above it (process_file, options, parmsets.c, got_datablock, the .tap/.bin/.log
 * writers) and everything in decode_*.c stays the reference's own unchanged C.
 *
 * It is written against the reference's own header (decoder.h) and is meant to be compiled
 * into readtape by a maintainer.  Nothing of the reference is copied here:
 * the row loop and process_sample()'s per-row dispatcher are re-formulated event-driven:
 *
 *   per block decode:   events = scan(start_row, PARM, globals)          <- GPU
 *   then, in row order: * the rows at which process_sample() would do something without a
 *                         detection are computed exactly from the same inequalities it
 *                         evaluates every row:
 *                           NRZI  timenow > nrzi.t_lastclock + 2*bitspaceavg  -> nrzi_zerocheck()  (decoder.c:844)
 *                           PE    timenow - t_lastpeak > bitspaceavg*2.5f      -> idle / pe_end_of_block (:868-877)
 *                           GCR   timenow > t_lastpeak + 6.00*bitspaceavg      -> idle / gcr_end_of_block (:879-888)
 *                           WW    timenow - t_lastclkpulseend > avg*1.5f       -> ww_end_of_block (:892-894)
 *                           the per-track first-sample initialisation rows      (:855-861)
 *                           interblock_counter expiry                           (:841, :900-903)
 *                       * each event is handed to the reference's process_up_transition() /
 *                         process_down_transition() (decoder.c:574/592) exactly as lookfor_peak /
 *                         lookfor_*zerocrossing would: t->v_top/t_top (or v_bot/t_bot) and the
 *                         global timenow are set first.
 * The handlers re-compute AGC themselves; the gain the GPU mirrored is compared after every
 * event (a free parity self-check): a mismatch is a fatal() error, never silently ignored.
 *
 * Link-time glue (no reference source is modified; see readtape_b200/host/Makefile):
 *   objcopy --weaken-symbol=readblock readtape.o          our readblock() wins
 *   -Wl,--wrap=init_trackstate,--wrap=ww_init_blockstate,--wrap=init_trackpeak_state,
 *       --wrap=compute_avg_height                         so we know which reset preceded a call
 */
#include "decoder.h"
#include "rt_scan.h"
#include <time.h>
#include <unistd.h>
#include <sys/wait.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <poll.h>
#include <pthread.h>

/* ---- reference globals we read (all non-static in readtape.c / decoder.c) ------------------------ */
extern FILE *inf;
extern char indatafilename[];
extern bool tbin_file, invert_data, do_differentiate, find_zeros, doing_density_detection, doing_deskew;
extern struct tbin_hdr_t tbin_hdr;
extern struct tbin_dat_t tbin_dat;
extern int nheads, samples_per_bit, subsample, head_to_trk[MAXTRKS], trk_to_head[MAXTRKS];
extern bool set_ntrks_from_order;
extern char track_order_string[MAXTRKS + 1];
extern char *wwtracktype_names[WWTRK_NUMTYPES];
extern long long lines_in, numsamples, numoutbytes;
extern double torigin;
extern bool tap_format, do_txtfile, deskew, skew_given;
extern FILE *outf;
extern int numblks, numblks_limit;
extern char baseoutfilename[], baseinfilename[];
void create_datafile(const char *name);                /* readtape.c:1092 */
void output_tap_marker(uint32_t num);                  /* readtape.c:1077 */
void close_file(void);                                 /* readtape.c:1085 */
void force_end_of_block(void);                         /* readtape.c:1378 */
void process_up_transition(struct trkstate_t *t);      /* decoder.c:574 */
void process_down_transition(struct trkstate_t *t);    /* decoder.c:592 */

/* ---- state ------------------------------------------------------------------------------------- */
static struct {
   int opened;
   rt_tape *tape;
   rt_tape_desc desc;
   long long base_pos;              /* file offset of row 0 */
   uint64_t nrows;                  /* rows before the end marker */
   size_t rows_bytes;               /* payload bytes uploaded */
   long long group_bytes;           /* file bytes per row of the tape: nheads * 2 * subsample (readtape.c:1407: of every `subsample` rows the last is used) */
   long long endfile_extra;         /* file bytes the reference has read past the last whole group when it meets the end marker */
   /* the pending reset, noted by the wrapped reset functions */
   int pending_reset;               /* RT_RESET_*; RT_RESET_NONE if none since the last readblock() */
   /* stateful exact context (Whirlwind, and the fallback for everything else) */
   rt_scan *ctx; rt_scan_cfg ctx_cfg; int ctx_valid;
   /* speculative whole-tape scans, one per parameter set on demand */
   struct { rt_bulk *bulk; rt_scan_cfg cfg; int valid; uint32_t ci; int shared; } bulk[MAXPARMSETS];   /* ci: configuration index inside `bulk` */
   int use_bulk;
   /* statistics */
   long long n_bulk_hits, n_bulk_miss, n_exact_spans, n_events, n_restarts;
   double s_scan, s_replay, s_open;  /* wall seconds: inside the rt_scan library / replaying events into the handlers / opening */
   double s_exact;                   /* of s_replay: inside exact-scan spans asked for in the middle of a block decode */
   int said_config;
   /* one reel split between worker processes (RT_WORKERS, see run_workers) */
   int par_checked, nworkers, worker;   /* worker: 0 .. nworkers-1, or -1 for the classic single process */
   uint64_t start_row;                  /* worker: the unit boundary this worker starts at */
   uint64_t stop_row;                   /* the unit boundary where the next worker starts; UINT64_MAX: run to the end */
   int must_seek_start;                 /* worker > 0: the first readblock() call only positions the file at the worker's first row */
   int remote_fd; rt_event *remote_buf; /* worker: exact scans are done by the parent (the only process with a CUDA context) */
   volatile int *handover_ps;           /* worker: shared slot for the parameter set its NEXT block would have been tried with first */
} S = { .worker = -1, .stop_row = UINT64_MAX, .remote_fd = -1 };

static double wall(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* the exact scan runs in spans: the first one of a block decode is short (Whirlwind blocks are a few thousand rows, and what was
   scanned beyond the block's end is scanned again after the next reset), the following ones double */
#define EXACT_SPAN_FIRST (1u << 12)
#define EXACT_SPAN_ROWS  (1u << 17)
/* the smallest share a reel is split into; the environment override exists for tests on the small bundled captures */
#define WORKER_MIN_ROWS     (getenv("RT_WORKER_MIN_ROWS") ? strtoull(getenv("RT_WORKER_MIN_ROWS"), NULLP, 10) : (uint64_t)(8u << 20))
#define WORKER_UNPROVEN     98              /* exit code of a worker whose hand-over could not be proven: the reel is decoded unsplit */

static void rtfatal(const char *what, int rc) {
   fatal("B200 scan: %s failed (%d): %s", what, rc, rt_last_error()); }

static double rowtime(uint64_t row) {                  /* readtape.c:1423 */
   return (double)(int64_t)(S.desc.tstart_ns + row * S.desc.tdelta_ns) / 1e9; }

/* ---- resets: remember what the host did before calling readblock() ------------------------------- */
void __real_init_trackstate(void);
void __wrap_init_trackstate(void) { __real_init_trackstate(); S.pending_reset = RT_RESET_FULL; }
void __real_ww_init_blockstate(void);
void __wrap_ww_init_blockstate(void) {
   __real_ww_init_blockstate();
   if (S.pending_reset != RT_RESET_FULL && S.pending_reset != RT_RESET_PEAKSTATE) S.pending_reset = RT_RESET_WW_PARTIAL;
   else S.pending_reset |= 0x100; /* a partial reset stacked on a pending full/peakstate one */ }
void __real_init_trackpeak_state(void);
void __wrap_init_trackpeak_state(void) { __real_init_trackpeak_state(); S.pending_reset = RT_RESET_PEAKSTATE; }
void __real_compute_avg_height(struct trkstate_t *t);
void __wrap_compute_avg_height(struct trkstate_t *t) {
   __real_compute_avg_height(t);
   if (S.ctx) { int rc = rt_scan_set_avg_height(S.ctx, (uint32_t)t->trknum, t->v_avg_height); if (rc) rtfatal("rt_scan_set_avg_height", rc); } }

/* ---- configuration ------------------------------------------------------------------------------ */
static void make_cfg(rt_scan_cfg *c) {
   memset(c, 0, sizeof *c);
   c->mode = (int32_t)mode;
   c->flags = (find_zeros ? RT_F_FIND_ZEROS : 0) | (do_differentiate ? RT_F_DIFFERENTIATE : 0) | (invert_data ? RT_F_INVERT : 0)
              | (doing_density_detection ? RT_F_DENSITY_DETECT : 0) | (doing_deskew ? RT_F_DESKEWING : 0);
   c->bpi = bpi; c->ips = ips;
   c->parms.clk_window = PARM.clk_window; c->parms.clk_alpha = PARM.clk_alpha;
   c->parms.agc_window = PARM.agc_window; c->parms.agc_alpha = PARM.agc_alpha;
   c->parms.min_peak = PARM.min_peak; c->parms.clk_factor = PARM.clk_factor; c->parms.pulse_adj = PARM.pulse_adj;
   c->parms.pkww_bitfrac = PARM.pkww_bitfrac; c->parms.pkww_rise = PARM.pkww_rise;
   c->parms.z1pt = PARM.z1pt; c->parms.z2pt = PARM.z2pt;
   for (int k = 0; k < MAXTRKS; ++k) c->skew_delaycnt[k] = skew_delaycnt[k]; }

static void open_tape(void) {
   assert(tbin_file, "the B200 scan reads .tbin captures only (convert CSV with csvtbin first)");
   assert(sizeof(int16_t) == 2 && nheads >= ntrks && nheads <= MAXTRKS, "bad head count %d", nheads);
   memset(&S.desc, 0, sizeof S.desc);
   S.desc.ntrks = (uint32_t)ntrks; S.desc.nheads = (uint32_t)nheads;
   for (int h = 0; h < MAXTRKS; ++h) S.desc.head_to_trk[h] = h < nheads ? head_to_trk[h] : RT_HEAD_IGNORE;
   S.desc.maxvolts = tbin_hdr.u.s.maxvolts;
   S.desc.tdelta_ns = (uint64_t)sample_deltat_ns;
   S.desc.tstart_ns = (uint64_t)timenow_ns;              /* time of the row at the current file position */
   S.base_pos = ftello(inf);
   assert(S.base_pos >= 0, "ftell failed");
   assert(fseeko(inf, 0, SEEK_END) == 0, "fseek failed");
   long long end = ftello(inf);
   assert(fseeko(inf, S.base_pos, SEEK_SET) == 0, "fseek failed");
   uint64_t rowbytes = (uint64_t)nheads * 2;
   uint64_t nrows_file = (uint64_t)(end - S.base_pos) / rowbytes;
   S.group_bytes = (long long)rowbytes * subsample; S.endfile_extra = 2;
   int dev = getenv("RT_DEVICE") ? atoi(getenv("RT_DEVICE")) : 0;
   double w1 = wall();
   int rc = rt_open(&S.desc, dev, &S.tape);
   if (rc) rtfatal("rt_open", rc);
   double w2 = wall();
   /* the payload goes to the GPU straight from the page cache: the library reads the file with a few threads into its pinned
      staging ring (no whole-file buffer, pinned or not) */
   S.rows_bytes = (size_t)(nrows_file * rowbytes);
   if (subsample > 1) {
      /* -subsample=n (readtape.c:1407-1413): the reference reads n rows per sample and uses the last; the end marker ends the file
         at whichever of them carries it.  The rows are thinned on the host (a rarely used option), an end-marker row is appended, and
         the tape the GPU sees is the thinned one; file positions map to its rows in groups of n. */
      int16_t *raw = malloc(S.rows_bytes + rowbytes);
      assert(raw != NULLP, "no memory for the .tbin payload");
      size_t got = 0;
      while (got < S.rows_bytes) { ssize_t k = pread(fileno(inf), (char *)raw + got, S.rows_bytes - got, (off_t)(S.base_pos + (long long)got)); assert(k > 0, "can't read .tbin data"); got += (size_t)k; }
      uint64_t m = 0;
      while (m < nrows_file && raw[m * nheads] != -32768) ++m;
      uint64_t kept = m / (uint64_t)subsample;
      for (uint64_t r = 0; r < kept; ++r) memmove(raw + r * nheads, raw + ((r + 1) * subsample - 1) * nheads, rowbytes);
      int have_marker = m < nrows_file;
      if (have_marker) { raw[kept * nheads] = -32768; S.endfile_extra = (long long)((m - kept * subsample) * rowbytes) + 2; }
      rc = rt_upload(S.tape, raw, kept + (uint64_t)have_marker);
      free(raw); }
   else if (getenv("RT_UPLOAD_MMAP") && atoi(getenv("RT_UPLOAD_MMAP"))) {   /* experiment: hand the library a mapping of the file instead */
      const long pg = sysconf(_SC_PAGESIZE);
      const off_t map_off = (off_t)(S.base_pos / pg * pg);
      const size_t map_len = S.rows_bytes + (size_t)(S.base_pos - map_off);
      void *map = mmap(NULLP, map_len, PROT_READ, MAP_SHARED, fileno(inf), map_off);
      assert(map != MAP_FAILED, "cannot map the .tbin payload");
      rc = rt_upload(S.tape, (const int16_t *)((const char *)map + (S.base_pos - map_off)), nrows_file);
      munmap(map, map_len); }
   else rc = rt_upload_fd(S.tape, fileno(inf), (uint64_t)S.base_pos, nrows_file);
   if (rc) rtfatal("rt_upload_fd", rc);
   if (getenv("RT_STATS") && atoi(getenv("RT_STATS")) >= 2)
      rlog("  B200 scan: rt_open %.3f s, upload of %.2f GB %.3f s\n", w2 - w1, S.rows_bytes / 1e9, wall() - w2);
   S.nrows = rt_nrows(S.tape);
   S.use_bulk = !(getenv("RT_NO_BULK") && atoi(getenv("RT_NO_BULK")));
   S.opened = 1;
   /* extra log lines only on request: without RT_STATS the .log is line for line the reference's */
   if (!quiet && getenv("RT_STATS")) rlog("  B200 scan: %s rows x %d heads resident on the GPU (%s)\n", longlongcommas((long long)S.nrows), nheads, rt_backend()); }

/* ---- event sources ------------------------------------------------------------------------------ */
struct evsrc {
   const rt_event *ev; uint64_t n, at;      /* current batch */
   uint64_t valid_end;                      /* rows < valid_end are covered by what we hold */
   uint64_t unit_end;                       /* speculative source: the end row of the unit the lookup arrived at (after chaining) */
   int exact;                               /* 1: S.ctx is running and can be continued */
   /* what is needed to carry on with the exact scan when a speculative unit ends inside the block */
   uint64_t row0; const rt_scan_cfg *cfg;
   uint64_t taken; uint64_t last_row; int last_trk;   /* events handed out so far, and the last of them */
   uint64_t span;                                     /* rows of the next exact span */
};
#define TAKE(src) do { (src)->last_row = (src)->ev[(src)->at].row; (src)->last_trk = (src)->ev[(src)->at].trk; ++(src)->at; ++(src)->taken; } while (0)

static void ctx_prepare(const rt_scan_cfg *cfg) {
   int rc;
   if (!S.ctx) {
      rc = rt_scan_begin(S.tape, cfg, &S.ctx); if (rc) rtfatal("rt_scan_begin", rc);
      S.ctx_cfg = *cfg; }
   else if (memcmp(&S.ctx_cfg, cfg, sizeof *cfg) != 0) {
      rc = rt_scan_set_cfg(S.ctx, cfg); if (rc) rtfatal("rt_scan_set_cfg", rc);
      S.ctx_cfg = *cfg; } }

/* worker processes have no CUDA context: their exact scans (misses, continuations) are served by the parent over a socket pair,
   the events come back through a shared buffer */
struct wreq { int op; int reset_kind; uint64_t row; rt_scan_cfg cfg; };
struct wrsp { int rc; uint64_t n, pos, done; char err[160]; };
enum { WOP_START = 1, WOP_MORE = 2 };
static void xfer(int fd, void *p, size_t n, int wr) {
   char *c = p;
   while (n) { ssize_t k = wr ? write(fd, c, n) : read(fd, c, n); if (k <= 0) fatal("B200 scan: lost the connection to the scanning process"); c += k; n -= (size_t)k; } }
static void remote_exact(struct evsrc *src, struct wreq *rq) {
   struct wrsp rs;
   xfer(S.remote_fd, rq, sizeof *rq, 1); xfer(S.remote_fd, &rs, sizeof rs, 0);
   if (rs.rc) fatal("B200 scan: exact scan failed in the scanning process (%d): %s", rs.rc, rs.err);
   src->ev = S.remote_buf; src->n = rs.n; src->at = 0; ++S.n_exact_spans;
   src->valid_end = rs.done ? rs.pos : UINT64_MAX; }

static void exact_more(struct evsrc *src) {  /* continue the exact scan by one span */
   if (S.remote_fd >= 0) { struct wreq rq; memset(&rq, 0, sizeof rq); rq.op = WOP_MORE; remote_exact(src, &rq); return; }
   uint64_t done = 0;
   if (!src->span) src->span = EXACT_SPAN_FIRST;
   const double x0 = wall();
   int rc = rt_scan_run(S.ctx, src->span, &src->ev, &src->n, &done);
   S.s_exact += wall() - x0;
   if (src->span < EXACT_SPAN_ROWS) src->span *= 2;
   if (rc) rtfatal("rt_scan_run", rc);
   src->at = 0; src->valid_end = rt_scan_pos(S.ctx); ++S.n_exact_spans;
   if (done == 0) src->valid_end = UINT64_MAX; /* end of tape: nothing more will ever come */ }

static void exact_start(struct evsrc *src, const rt_scan_cfg *cfg, int reset_kind, uint64_t row) {
   if (S.remote_fd >= 0) {
      struct wreq rq; memset(&rq, 0, sizeof rq); rq.op = WOP_START; rq.reset_kind = reset_kind; rq.row = row; rq.cfg = *cfg;
      memset(src, 0, sizeof *src); src->exact = 1; src->row0 = row; src->cfg = cfg;
      remote_exact(src, &rq);
      return; }
   ctx_prepare(cfg);
   int kind = reset_kind & 0xff;
   int rc = rt_scan_reset(S.ctx, kind, row); if (rc) rtfatal("rt_scan_reset", rc);
   if (reset_kind & 0x100) { rc = rt_scan_reset(S.ctx, RT_RESET_WW_PARTIAL, row); if (rc) rtfatal("rt_scan_reset", rc); }
   memset(src, 0, sizeof *src); src->exact = 1; src->row0 = row; src->cfg = cfg;
   exact_more(src); }

/* The speculative unit ended before the block did.  Its events ARE the events of a fresh reset at the block's first row (that is
   what rt_bulk_lookup proved), so the exact scan from that row reproduces them one for one: it is started, the events already
   replayed are skipped, and the replay carries on where it was -- nothing is replayed twice (peak statistics, log lines and the
   handlers' one-shot warnings stay exactly the reference's). */
static void continue_exact(struct evsrc *src) {
   const uint64_t taken = src->taken, last_row = src->last_row; const int last_trk = src->last_trk;
   const uint64_t row0 = src->row0; const rt_scan_cfg *cfg = src->cfg;
   ++S.n_restarts;
   if (getenv("RT_STATS") && atoi(getenv("RT_STATS")) >= 2)
      rlog("  B200 scan: the unit found for row %llu ends at row %llu, inside the block: exact scan from there on\n",
           (unsigned long long)row0, (unsigned long long)src->valid_end);
   exact_start(src, cfg, RT_RESET_FULL, row0);
   uint64_t skip = taken;
   while (skip) {
      const uint64_t have = src->n - src->at;
      if (have >= skip) {
         src->at += skip; skip = 0;
         const rt_event *e = &src->ev[src->at - 1];
         if (e->row != last_row || e->trk != last_trk)
            fatal("B200 scan: exact scan disagrees with the speculative unit at row %llu", (unsigned long long)last_row); }
      else {
         skip -= have; src->at = src->n;
         if (src->valid_end == UINT64_MAX) fatal("B200 scan: exact scan ended before the speculative unit did");
         exact_more(src); } }
   src->taken = taken; src->last_row = last_row; src->last_trk = last_trk; }

/* diagnostics (RT_STATS=2): why no speculative unit could be proven equivalent to a fresh reset at `row` */
/* the rule of unit_covers (rt_api.cu): no loud row since the reset row -- and, for the zero-crossing detectors, none since the unit's own first row */
static int quiet_since(const rt_unit_info *ui, uint64_t loud, uint64_t row) {
   return loud == UINT64_MAX || (loud < row && (!find_zeros || loud < ui->row0)); }
static void say_unit(const rt_unit_info *ui, uint64_t row, int all) {
   rlog("     unit %llu of %llu [%llu, %llu)\n", (unsigned long long)ui->unit_index, (unsigned long long)ui->nunits,
        (unsigned long long)ui->row0, (unsigned long long)ui->row_end);
   for (uint32_t k = 0; k < ui->ntrks; ++k) {
      const int late = ui->sync_row[k] != UINT64_MAX && ui->sync_row[k] >= ui->need_sync_row[k] && quiet_since(ui, ui->last_loud_row[k], row);
      const int early = ui->sync_early[k] != UINT64_MAX && ui->sync_early[k] >= ui->need_sync_row[k] && quiet_since(ui, ui->loud_early[k], row);
      if ((late || early) && !all) continue;
      rlog("       trk %u%s: first event %lld, sync %lld (loud %lld), early sync %lld (loud %lld), need %lld, failed %u, events %u\n", k, late || early ? " ok" : "",
           (long long)ui->first_event_row[k], (long long)ui->sync_row[k], (long long)ui->last_loud_row[k], (long long)ui->sync_early[k],
           (long long)ui->loud_early[k], (long long)ui->need_sync_row[k], ui->failed[k], ui->nevents[k]); } }
static void say_miss(rt_bulk *bulk, uint32_t ci, uint64_t row) {
   static rt_unit_info ui, un;
   if (rt_bulk_unit_info(bulk, ci, row, &ui) != RT_OK) { rlog("  B200 scan: miss at row %llu: no unit\n", (unsigned long long)row); return; }
   rlog("  B200 scan: miss at row %llu, parmset %d\n", (unsigned long long)row, block.parmset);
   say_unit(&ui, row, 0);
   if (rt_bulk_unit_at(bulk, ci, ui.unit_index + 1, &un) == RT_OK) {
      for (uint32_t k = 0; k < un.ntrks; ++k) un.need_sync_row[k] = ui.need_sync_row[k];       /* as seen from `row` */
      say_unit(&un, row, 0); } }

/* BASELINE config 3: every active parameter set scanned by ONE rt_bulk_scan() call (their kernels share the sample planes and run
   side by side); the reference then picks per block (readtape.c:1755-1795).  Returns 0 if the fan-out does not apply. */
static int scan_parmsets(void) {
   static rt_scan_cfg cfgs[MAXPARMSETS]; int which[MAXPARMSETS], n = 0;
   const int keep = block.parmset;
   for (int i = 0; i < MAXPARMSETS; ++i) {
      if (!(i == keep || (multiple_tries && parmsetsptr[i].active))) continue;
      block.parmset = i; make_cfg(&cfgs[n]);
      if (i != keep && S.bulk[i].valid && memcmp(&S.bulk[i].cfg, &cfgs[n], sizeof cfgs[n]) == 0) continue;   /* already scanned */
      which[n++] = i; }
   block.parmset = keep;
   rt_bulk *bulk = NULLP;
   int rc = rt_bulk_scan(S.tape, cfgs, (uint32_t)n, &bulk);
   if (rc == RT_ERR_UNSUPPORTED) return 0;
   if (rc) rtfatal("rt_bulk_scan", rc);
   for (int k = 0; k < n; ++k) {
      const int ps = which[k];
      if (S.bulk[ps].valid && !S.bulk[ps].shared) rt_bulk_free(S.bulk[ps].bulk);
      S.bulk[ps].bulk = bulk; S.bulk[ps].cfg = cfgs[k]; S.bulk[ps].valid = 1; S.bulk[ps].ci = (uint32_t)k; S.bulk[ps].shared = 1; }
   return n; }

static int bulk_start(struct evsrc *src, const rt_scan_cfg *cfg, uint64_t row) {
   int ps = block.parmset;
   if (!S.use_bulk || mode == WW) return 0;         /* Whirlwind: the detector state persists from block to block */
   /* the density and deskew pre-passes (readtape.c:1656-1717) reset per block like the main pass: the same speculative scan
      serves them, with their own configuration (handlers bypassed and width 8 / skew delays still zero) */
   if (S.bulk[ps].valid && memcmp(&S.bulk[ps].cfg, cfg, sizeof *cfg) != 0) {    /* e.g. the skew changed after the pre-pass */
      if (!S.bulk[ps].shared && S.remote_fd < 0) rt_bulk_free(S.bulk[ps].bulk);
      S.bulk[ps].valid = 0; }
   if (!S.bulk[ps].valid) {
      if (S.remote_fd >= 0) return 0;                 /* a worker cannot scan: the parent serves this decode with the exact scan */
      /* Most blocks decode with the first parameter set, so that one is scanned alone.  The first block that needs another try
         usually is not the last: on a tape of moderate size all remaining active sets are then scanned by ONE call, side by side,
         instead of one whole-tape scan per set (RT_FANOUT=1: all of them from the start; RT_FANOUT=0: never). */
      const char *fo = getenv("RT_FANOUT");
      int any_valid = 0;
      for (int i = 0; i < MAXPARMSETS; ++i) any_valid |= S.bulk[i].valid;
      const int fan = fo ? (atoi(fo) != 0) : (any_valid && multiple_tries && (double)S.nrows * ntrks < 2e9);
      if (fan && !doing_density_detection && !doing_deskew && scan_parmsets()) { /* all at once */ }
      else {
         int rc = rt_bulk_scan(S.tape, cfg, 1, &S.bulk[ps].bulk);
         if (rc == RT_ERR_UNSUPPORTED) { S.use_bulk = 0; return 0; }
         if (rc) rtfatal("rt_bulk_scan", rc);
         S.bulk[ps].cfg = *cfg; S.bulk[ps].valid = 1; S.bulk[ps].ci = 0; S.bulk[ps].shared = 0; } }
   if (!S.bulk[ps].valid) return 0;
   uint64_t valid = 0;
   memset(src, 0, sizeof *src); src->row0 = row; src->cfg = cfg;
   int rc = rt_bulk_lookup(S.bulk[ps].bulk, S.bulk[ps].ci, row, &src->ev, &src->n, &valid);
   if (rc == RT_MISS) {
      ++S.n_bulk_miss;
      if (getenv("RT_STATS") && atoi(getenv("RT_STATS")) >= 2) say_miss(S.bulk[ps].bulk, S.bulk[ps].ci, row);
      return 0; }
   if (rc) rtfatal("rt_bulk_lookup", rc);
   src->valid_end = src->unit_end = row + valid;
   if (src->valid_end >= S.nrows) src->valid_end = UINT64_MAX;
   ++S.n_bulk_hits;
   return 1; }

/* the next event, or NULL if none is known below row `limit` (then rows < limit hold no event) */
static const rt_event *peek_event(struct evsrc *src, uint64_t limit) {
   for (;;) {
      if (src->at < src->n) return src->ev[src->at].row < limit ? &src->ev[src->at] : NULLP;
      if (src->valid_end == UINT64_MAX || src->valid_end >= limit) return NULLP;
      if (src->exact) exact_more(src);
      else continue_exact(src); } }

/* ---- exact "first row at which process_sample() sees the condition" searches ----------------------- */
/* smallest row r >= lo with  rowtime(r) > x  (monotone in r) */
static uint64_t first_row_time_gt(double x, uint64_t lo) {
   if (!(x == x) || lo >= S.nrows) return UINT64_MAX;                        /* NaN, or past the end */
   double est = (x * 1e9 - (double)S.desc.tstart_ns) / (double)S.desc.tdelta_ns;
   if (est >= (double)S.nrows + 4) return UINT64_MAX;                         /* never before the end marker */
   uint64_t r = est <= (double)lo ? lo : (uint64_t)est;
   if (r > lo + 2) r -= 2; else r = lo;
   while (r > lo && rowtime(r - 1) > x) --r;
   while (r < S.nrows && !(rowtime(r) > x)) ++r;
   return r < S.nrows ? r : UINT64_MAX; }
/* smallest row r >= lo with  rowtime(r) - t0 > thr   (the subtraction is done in double, as in the reference) */
static uint64_t first_row_delta_gt(double t0, double thr, uint64_t lo) {
   if (lo >= S.nrows) return UINT64_MAX;
   uint64_t r = first_row_time_gt(t0 + thr, lo);
   if (r == UINT64_MAX) r = S.nrows - 1;
   if (r > lo + 2) r -= 2; else r = lo;
   while (r > lo && rowtime(r - 1) - t0 > thr) --r;
   while (r < S.nrows && !(rowtime(r) - t0 > thr)) ++r;
   return r < S.nrows ? r : UINT64_MAX; }

/* ---- the block decode ------------------------------------------------------------------------------ */
struct rowstate { uint64_t init_row[MAXTRKS]; };

/* the "execution-time configuration" block the reference logs before its first sample (readtape.c:1460-1499): the log must not
   drift, so every line it prints is printed here too, from the same globals, in the same order */
static void say_configuration(void) {
   if (quiet || S.said_config) return;
   S.said_config = 1;
   const int spb = bpi != 0 ? (int)(1 / (bpi * ips * sample_deltat)) : 0;
   rlog("\nexecution-time configuration:\n");
   if (set_ntrks_from_order) rlog("  we set ntrks=%d as implied by the -order string \"%s\"\n", ntrks, track_order_string);
   rlog("  %d track %s encoding, %s parity, %d BPI at %d IPS", ntrks, modename(),
        mode == WW ? "no" : expected_parity ? "odd" : "even", (int)bpi, (int)ips);
   if (bpi != 0) rlog(" (%.2f usec/bit)", 1e6f / (bpi * ips));
   rlog("\n  first sample is at time %.8lf seconds on the tape\n", timenow);
   if (subsample > 1) rlog("  subsampling every %d samples\n", subsample);
   if (invert_data) rlog("  inverting the data polarity\n");
   if (reverse_tape) rlog("  reversing the bit pairs in each word, and the words in each block\n");
   rlog("  sampling rate is %s Hz (%.2f usec)", intcommas((int)(1.0 / sample_deltat)), sample_deltat * 1e6);
   if (bpi != 0) rlog(", or about %d samples per bit", spb);
   rlog("\n");
   if (bpi != 0 && spb > 100) rlog("  ---> Warning: excessive samples per bit; consider using the -subsample option\n");
   if (find_zeros) rlog("  will look for zero crossings, not peaks\n");
   else rlog("  peak detection window width is %d samples (%.2f usec)\n", pkww_width, pkww_width * sample_deltat * 1e6);
   if (mode == WW) {
      rlog("  Whirlwind data has %d tracks from %d data heads assigned as follows:\n", ntrks, nheads);
      for (int ty = 0; ty < WWTRK_NUMTYPES; ++ty) {
         const int trk = ww_type_to_trk[ty];
         if (trk == -1) rlog("              there is no  ");
         else rlog("    track %d, head %d is the ", trk, trk_to_head[trk]);
         rlog(" %s, '%c'\n", wwtracktype_names[ty], WWTRKTYPE_SYMBOLS[ty]); }
      for (int h = 0; h < nheads; ++h) if (head_to_trk[h] == WWHEAD_IGNORE) rlog("             head %d is unused\n", h);
      const char *dir = flux_direction_requested == FLUX_AUTO ? "will be automatically determined for each block"
                        : flux_direction_requested == FLUX_POS ? "is expected to be positive"
                        : flux_direction_requested == FLUX_NEG ? "is expected to be negative" : "--- internal error  ---";
      rlog("  the initial peak polarity for each flux change %s\n", dir); }
   else {
      rlog("  input data order: ");
      for (int i = 0; i < ntrks; ++i) {
         const int k = head_to_trk[i];
         if (k == ntrks - 1) rlog("p"); else rlog("%d", k);
         if (k == 0) rlog("(msb)");
         if (k == ntrks - 2) rlog("(lsb)"); }
      rlog("\n"); }
   rlog("\n"); }

/* returns the last row consumed (the row after which the reference's readblock() returns); *endfile set at EOF */
/* exit logic of process_sample() once a row is done, decoder.c:900-904; returns 1 if the block decode ends here */
static inline int row_exit(uint64_t row, uint64_t nrows, uint64_t *last_row, bool *endfile) {
   if (interblock_counter) {
      /* rows row .. row+interblock_counter-1 are swallowed; the block is returned after the last of them */
      uint64_t ret = row + (uint64_t)interblock_counter - 1;
      if (ret >= nrows) {               /* the end marker comes first */
         timenow = rowtime(nrows - 1);
         force_end_of_block();
         *endfile = true; *last_row = nrows;
         return 1; }
      interblock_counter = 0;
      timenow = rowtime(ret);
      *last_row = ret + 1;
      return 1; }
   if (block.results[block.parmset].blktype != BS_NONE) { *last_row = row + 1; return 1; }
   return 0; }

/* hand one event to the reference's handler exactly as lookfor_peak / lookfor_*zerocrossing would (decoder.c:574/592) */
static inline void apply_event(struct trkstate_t *t, const rt_event *e, uint64_t row) {
   ++S.n_events;
   t->v_top = e->v_top; t->v_bot = e->v_bot;
   if (e->kind == RT_EV_TOP) { t->t_top = e->t_event; process_up_transition(t); }
   else { t->t_bot = e->t_event; process_down_transition(t); }
   if (t->agc_gain != e->agc_gain)
      fatal("B200 scan diverged from the host on track %d at row %llu: AGC %.9g (scan) vs %.9g (host)",
            (int)e->trk, (unsigned long long)row, e->agc_gain, t->agc_gain); }

/* Is any of process_sample()'s per-row conditions (decoder.c:844,868,879,892) true at a row whose time is `te`?  They are
   monotone in the time, so "false at te" means false at every earlier row since the state last changed. */
static inline int timer_due_at(double te) {
   if (mode == NRZI) return nrzi.datablock && te > nrzi.t_lastclock + 2 * nrzi.clkavg.t_bitspaceavg;
   if (mode == PE) {
      for (int k = 0; k < ntrks; ++k) {
         const struct trkstate_t *t = &trkstate[k];
         if (!t->idle && t->t_lastpeak != 0 && te - t->t_lastpeak > t->clkavg.t_bitspaceavg * PE_IDLE_FACTOR) return 1; }
      return 0; }
   if (mode == GCR) {
      for (int k = 0; k < ntrks; ++k) {
         const struct trkstate_t *t = &trkstate[k];
         if (t->datablock && te > t->t_lastpeak + GCR_IDLE_THRESH * t->clkavg.t_bitspaceavg) return 1; }
      return 0; }
   return ww.datablock && ww.t_lastclkpulseend > 0 && te - ww.t_lastclkpulseend > ww.clkavg.t_bitspaceavg * WW_CLKSTOP_BITS; }

static void decode_from(uint64_t row0, int reset_kind, const rt_scan_cfg *cfg, struct evsrc *src, uint64_t *last_row, bool *endfile) {
   struct rowstate rs;
   const uint64_t nrows = S.nrows;
   const int tz = rowtime(row0) == 0.0;
   /* (Q2) decoder.c:855-861: after a reset the tracks initialise one per row */
   for (int k = 0; k < ntrks; ++k) rs.init_row[k] = UINT64_MAX;
   if ((reset_kind & 0xff) == RT_RESET_FULL || (reset_kind & 0xff) == RT_RESET_WW_PARTIAL || (reset_kind & 0x100))
      for (int k = 0; k < ntrks; ++k) rs.init_row[k] = row0 + (uint64_t)k + (tz ? 1u : 0u);
   else for (int k = 0; k < ntrks; ++k) if (trkstate[k].t_lastpeak == 0) rs.init_row[k] = row0 + (uint64_t)k;   /* defensive */

   uint64_t row = row0;                 /* next row to be "processed" */
   uint64_t pe_idle_row[MAXTRKS], gcr_idle_row[MAXTRKS];
   for (int k = 0; k < ntrks; ++k) pe_idle_row[k] = gcr_idle_row[k] = UINT64_MAX;
   uint64_t zerocheck_row = UINT64_MAX, ww_stop_row = UINT64_MAX;
   int dirty = 1;                       /* timers must be recomputed */
   *endfile = false;

   int init_pending = 0;
   for (int k = 0; k < ntrks; ++k) if (rs.init_row[k] != UINT64_MAX) ++init_pending;

   for (;;) {
      /* ---- the common case, one event at a time: every track is initialised, the next event is at hand, and none of the per-row
              conditions becomes true at or before its row -- then process_sample() does nothing on the rows in between, and on the
              event's row only what the events of that row make it do ---- */
      while (!init_pending && src->at < src->n) {
         const rt_event *e = &src->ev[src->at];
         const double te = rowtime(e->row);
         if (timer_due_at(te)) break;
         row = e->row; timenow = te;
         for (;;) {                                             /* the events of this row, in track order */
            struct trkstate_t *t = &trkstate[e->trk];
            TAKE(src);
            apply_event(t, e, row);
            /* the per-track tests behind the detector call, decoder.c:868-888 (the other tracks: not due, see above) */
            if (mode == PE && !t->idle && t->t_lastpeak != 0 && timenow - t->t_lastpeak > t->clkavg.t_bitspaceavg * PE_IDLE_FACTOR) {
               t->v_lastpeak = t->v_now; t->idle = true;
               if (++num_trks_idle >= ntrks) pe_end_of_block(); }
            if (mode == GCR && t->datablock && timenow > t->t_lastpeak + GCR_IDLE_THRESH * t->clkavg.t_bitspaceavg) {
               t->datablock = false; t->idle = true;
               if (++num_trks_idle >= ntrks) { gcr_end_of_block(); break; } }
            if (src->at >= src->n || src->ev[src->at].row != row) break;
            e = &src->ev[src->at]; }
         dirty = 1;
         if (mode == WW && timer_due_at(timenow)) ww_end_of_block();
         while (src->at < src->n && src->ev[src->at].row == row) TAKE(src);      /* behind a `goto exit`: not reached */
         if (row_exit(row, nrows, last_row, endfile)) return;
         ++row; }

      /* ---- timers: the first row >= `row` at which each per-row condition of process_sample() holds ---- */
      if (dirty) {
         zerocheck_row = ww_stop_row = UINT64_MAX;
         if (mode == NRZI && nrzi.datablock)
            zerocheck_row = first_row_time_gt(nrzi.t_lastclock + 2 * nrzi.clkavg.t_bitspaceavg, row);
         for (int k = 0; k < ntrks; ++k) {
            struct trkstate_t *t = &trkstate[k];
            pe_idle_row[k] = gcr_idle_row[k] = UINT64_MAX;
            uint64_t lo = row;
            if (rs.init_row[k] != UINT64_MAX && rs.init_row[k] + 1 > lo) lo = rs.init_row[k] + 1;   /* not looked at before that */
            if (rs.init_row[k] != UINT64_MAX) continue;             /* t_lastpeak == 0: the reference does not look at it yet */
            if (mode == PE && !t->idle && t->t_lastpeak != 0)
               pe_idle_row[k] = first_row_delta_gt(t->t_lastpeak, t->clkavg.t_bitspaceavg * PE_IDLE_FACTOR, lo);
            if (mode == GCR && t->datablock)
               gcr_idle_row[k] = first_row_time_gt(t->t_lastpeak + GCR_IDLE_THRESH * t->clkavg.t_bitspaceavg, lo); }
         if (mode == WW && ww.datablock && ww.t_lastclkpulseend > 0)
            ww_stop_row = first_row_delta_gt(ww.t_lastclkpulseend, ww.clkavg.t_bitspaceavg * WW_CLKSTOP_BITS, row);
         dirty = 0; }

      /* ---- the next row at which anything happens ---- */
      uint64_t next = nrows;               /* the end marker */
      if (zerocheck_row < next) next = zerocheck_row;
      if (ww_stop_row < next) next = ww_stop_row;
      for (int k = 0; k < ntrks; ++k) {
         if (rs.init_row[k] != UINT64_MAX && rs.init_row[k] >= row && rs.init_row[k] < next) next = rs.init_row[k];
         if (pe_idle_row[k] < next) next = pe_idle_row[k];
         if (gcr_idle_row[k] < next) next = gcr_idle_row[k]; }
      const rt_event *e = peek_event(src, next + 1);
      if (e && e->row < next) next = e->row;
      if (next >= nrows) {                /* readtape.c:1410-1413: the end marker */
         timenow = nrows ? rowtime(nrows - 1) : timenow;
         if (nrows > row0) force_end_of_block();
         *endfile = true;
         *last_row = nrows;               /* file position: at the marker */
         return; }

      /* ---- process row `next` in the order process_sample() does ---- */
      row = next;
      timenow = rowtime(row);
      if (mode == NRZI && nrzi.datablock && zerocheck_row == row) { nrzi_zerocheck(); dirty = 1; }
      int stop_tracks = 0;
      for (int k = 0; k < ntrks && !stop_tracks; ++k) {
         struct trkstate_t *t = &trkstate[k];
         if (rs.init_row[k] != UINT64_MAX) {
            if (row < rs.init_row[k]) break;                       /* the reference's `break`: later tracks are not reached either */
            if (row == rs.init_row[k]) {                           /* decoder.c:855-861 */
               t->t_lastpeak = timenow;
               rs.init_row[k] = UINT64_MAX; --init_pending;
               dirty = 1;
               break; } }
         e = peek_event(src, row + 1);
         if (e && e->row == row && e->trk == k) {
            TAKE(src);
            apply_event(t, e, row);
            dirty = 1; }
         else if (e && e->row == row && e->trk < k)
            fatal("B200 scan: event order violated at row %llu", (unsigned long long)row);
         if (mode == PE && !t->idle && t->t_lastpeak != 0 && timenow - t->t_lastpeak > t->clkavg.t_bitspaceavg * PE_IDLE_FACTOR) {
            t->v_lastpeak = t->v_now;                              /* decoder.c:868-877 */
            t->idle = true;
            dirty = 1;
            if (++num_trks_idle >= ntrks) pe_end_of_block(); }
         if (mode == GCR && t->datablock && timenow > t->t_lastpeak + GCR_IDLE_THRESH * t->clkavg.t_bitspaceavg) {
            t->datablock = false;                                  /* decoder.c:879-888 */
            t->idle = true;
            dirty = 1;
            if (++num_trks_idle >= ntrks) { gcr_end_of_block(); stop_tracks = 1; } } }
      if (!stop_tracks && mode == WW && ww.datablock && ww.t_lastclkpulseend > 0
            && timenow - ww.t_lastclkpulseend > ww.clkavg.t_bitspaceavg * WW_CLKSTOP_BITS) {
         ww_end_of_block(); dirty = 1; }
      /* events of this row on tracks the reference did not reach (its `break` / `goto exit`) are dropped */
      for (;;) {
         e = peek_event(src, row + 1);
         if (!e || e->row != row) break;
         TAKE(src); }

      if (row_exit(row, nrows, last_row, endfile)) return;
      ++row; } }

/* ---- one reel, several worker processes (opt-in: RT_WORKERS=P) ------------------------------------------------------------
 * The replay of the events through the reference's handlers is single-threaded host work (the reference is not re-entrant), and once
 * the scan runs on the GPU it is all that is left: ~150 ns per flux transition.  Every block decode starts from init_trackstate(),
 * so blocks are independent, and the reference's per-block outputs in .tap format simply concatenate.  With RT_WORKERS=P the first
 * readblock() call
 *   1. opens the tape and scans it for every parameter set the run may use (one rt_bulk_scan call), results fetched into memory
 *      that forked children can read (RT_OPT_SHARED_RESULTS);
 *   2. cuts the reel into P parts at unit boundaries (a unit boundary is the start of an all-track quiet stretch, k_units.cu) and
 *      forks P workers.  A worker never touches CUDA (a context does not survive fork(), and creating one costs seconds): it runs
 *      the reference's unchanged process_file() loop on its part, looking its events up in the shared results and writing
 *      <out>.partNNN.tap.  The rare decodes the speculative scan cannot serve (misses, units that end inside a block) are scanned by
 *      the parent, which answers over a socket pair;
 *   3. worker i > 0 starts with a fresh init_trackstate() at its boundary; worker i < P-1 stops at the first block start for which
 *      rt_bulk_lookup() PROVES that a fresh reset there is equivalent to a fresh reset at a unit at or behind the next worker's
 *      boundary.  If that proof fails the worker exits with WORKER_UNPROVEN and the parent decodes the reel unsplit;
 *   4. the parent concatenates the parts (each without its end-of-medium marker), writes the marker, and reports like readtape's
 *      main() does in quiet mode.
 * Only what is safe to split is split: .tap output, quiet mode (per-block log lines carry running block numbers), no text file, no
 * Whirlwind (state persists), no density / deskew pre-pass pending. */
#define WORKER_BUF_EVENTS (1u << 19)
struct chan { int fd; pid_t pid; int alive; rt_event *buf; rt_scan *ctx; rt_scan_cfg cfg; uint64_t span; };

static void part_name(char *buf, size_t n, int i, const char *ext) { snprintf(buf, n, "%s.part%03d%s", baseoutfilename, i, ext); }
static void worker_exit(int status, void *arg) { (void)arg; fflush(NULLP); _exit(status); }   /* no CUDA teardown in a forked child */

static void serve_one(struct chan *ch) {                       /* parent: one exact-scan request of a worker */
   struct wreq rq; struct wrsp rs;
   memset(&rs, 0, sizeof rs);
   size_t got = 0;
   while (got < sizeof rq) { ssize_t k = read(ch->fd, (char *)&rq + got, sizeof rq - got); if (k <= 0) { ch->alive = 0; return; } got += (size_t)k; }
   int rc = RT_OK;
   if (rq.op == WOP_START) {
      if (!ch->ctx) rc = rt_scan_begin(S.tape, &rq.cfg, &ch->ctx);
      else if (memcmp(&ch->cfg, &rq.cfg, sizeof rq.cfg) != 0) rc = rt_scan_set_cfg(ch->ctx, &rq.cfg);
      ch->cfg = rq.cfg;
      if (!rc) rc = rt_scan_reset(ch->ctx, rq.reset_kind & 0xff, rq.row);
      ch->span = EXACT_SPAN_FIRST; }
   const rt_event *ev = NULLP; uint64_t n = 0, done = 0;
   uint64_t span0 = ch->span ? ch->span : EXACT_SPAN_FIRST;
   if (ch->span < EXACT_SPAN_ROWS) ch->span = span0 * 2;
   for (uint64_t span = span0; !rc; span /= 2) {                /* a span whose events fit the shared buffer */
      rc = rt_scan_run(ch->ctx, span, &ev, &n, &done);
      if (rc || n <= WORKER_BUF_EVENTS) break;
      rc = rt_scan_rewind(ch->ctx, rt_scan_pos(ch->ctx) - done);
      if (!rc && span <= 1) rc = RT_ERR_OVERFLOW; }
   if (!rc) { if (n) memcpy(ch->buf, ev, (size_t)n * sizeof *ev); rs.n = n; rs.pos = rt_scan_pos(ch->ctx); rs.done = done; }
   else { rs.rc = rc; strncpy(rs.err, rt_last_error(), sizeof rs.err - 1); }
   xfer(ch->fd, &rs, sizeof rs, 1); }

/* the memory the workers will read the events from is prepared while CUDA starts up and the file travels to the GPU: an anonymous
   MAP_SHARED mapping, its pages touched by a few threads (4-5 GB of page faults are a second of single-thread work otherwise) */
static struct { pthread_t th; char *buf; size_t bytes; int started; volatile int stop; int registered; } PREP;
struct touch { char *p; size_t n; };
static void *prep_touch(void *arg) { struct touch *r = arg; for (size_t i = 0; i < r->n; i += 4096) r->p[i] = 0; return NULLP; }
static void *prep_main(void *arg) {
   (void)arg;
   void *p = mmap(NULLP, PREP.bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
   if (p == MAP_FAILED) { PREP.buf = NULLP; return NULLP; }
   PREP.buf = p;
   enum { NT = 8 };
   pthread_t th[NT]; struct touch r[NT];
   const size_t per = (PREP.bytes / NT + 4095) / 4096 * 4096;
   for (int k = 0; k < NT; ++k) {
      size_t lo = (size_t)k * per; if (lo > PREP.bytes) lo = PREP.bytes;
      r[k].p = PREP.buf + lo; r[k].n = PREP.bytes - lo < per ? PREP.bytes - lo : per;
      pthread_create(&th[k], NULLP, prep_touch, &r[k]); }
   for (int k = 0; k < NT; ++k) pthread_join(th[k], NULLP);
   /* pin it for the device-to-host copy as soon as there is a CUDA context (rt_open done), while the main thread uploads the file */
   while (!S.tape && !PREP.stop) usleep(2000);
   if (S.tape && !PREP.stop) PREP.registered = rt_host_register(S.tape, PREP.buf, PREP.bytes) == RT_OK;
   return NULLP; }

#define WRET do { if (PREP.buf) { if (PREP.registered) rt_host_unregister(S.tape, PREP.buf); munmap(PREP.buf, PREP.bytes); PREP.buf = NULLP; } return; } while (0)
static void run_workers(void) {
   const char *env = getenv("RT_WORKERS");
   int P = env ? atoi(env) : 1;
   if (P <= 1) return;
   if (!(tbin_file && tap_format && quiet && !do_txtfile && mode != WW && subsample == 1 && bpi != 0 && !doing_density_detection
         && !doing_deskew && (!deskew || skew_given) && numblks == 0 && numblks_limit == INT_MAX && outf == NULLP)) return;
   double w0 = wall();
   {  /* events to expect: one per ~48 track-samples per parameter set (800 BPI NRZI at 20 samples per bit: one per ~76) */
      long long pos = ftello(inf); fseeko(inf, 0, SEEK_END); long long end = ftello(inf); fseeko(inf, pos, SEEK_SET);
      int nps = 0; for (int i = 0; i < MAXPARMSETS; ++i) if (i == block.parmset || (multiple_tries && parmsetsptr[i].active)) ++nps;
      PREP.bytes = ((size_t)((end - pos) / 2 / 48) * 32 * (size_t)nps + (64u << 20)) / 4096 * 4096;
      PREP.started = pthread_create(&PREP.th, NULLP, prep_main, NULLP) == 0; }
   open_tape();
   PREP.stop = 1;
   if (PREP.started) pthread_join(PREP.th, NULLP);
   if (!S.use_bulk) WRET;
   if (P > 256) P = 256;
   if (S.nrows / (uint64_t)P < WORKER_MIN_ROWS) P = (int)(S.nrows / WORKER_MIN_ROWS);
   if (P <= 1) WRET;
   /* 1. every parameter set the run may use, scanned at once, results where children can read them */
   rt_set_option(RT_OPT_SHARED_RESULTS, 1);
   const int ps0 = block.parmset;
   double w1 = wall();
   if (!scan_parmsets()) { rt_set_option(RT_OPT_SHARED_RESULTS, 0); WRET; }
   double w2 = wall();
   int rc = RT_ERR_OVERFLOW;
   if (PREP.started && PREP.buf) {                              /* straight into the prepared mapping, pinned for the copy */
      if (!PREP.registered) PREP.registered = rt_host_register(S.tape, PREP.buf, PREP.bytes) == RT_OK;
      rc = rt_bulk_fetch_to(S.bulk[ps0].bulk, PREP.buf, PREP.bytes);
      if (rc == RT_ERR_OVERFLOW) { if (PREP.registered) rt_host_unregister(S.tape, PREP.buf); munmap(PREP.buf, PREP.bytes); PREP.buf = NULLP; }
      else PREP.buf = NULLP; }                                  /* the results live there from now on */
   if (rc == RT_ERR_OVERFLOW) rc = rt_bulk_fetch(S.bulk[ps0].bulk);   /* more events than expected: the library maps what is needed */
   rt_set_option(RT_OPT_SHARED_RESULTS, 0);
   if (rc) rtfatal("rt_bulk_fetch", rc);
   if (getenv("RT_STATS") && atoi(getenv("RT_STATS")) >= 2) printf("  B200 scan: whole-tape scan %.3f s, results to the host %.3f s\n", w2 - w1, wall() - w2);
   /* 2. the cut points: the first unit boundary behind each nominal share */
   static uint64_t cut[257]; static rt_unit_info ui;
   int parts = 1; cut[0] = 0;
   for (int i = 1; i < P; ++i) {
      rc = rt_bulk_unit_info(S.bulk[ps0].bulk, S.bulk[ps0].ci, (uint64_t)i * S.nrows / (uint64_t)P, &ui);
      if (rc == RT_OK) rc = rt_bulk_unit_at(S.bulk[ps0].bulk, S.bulk[ps0].ci, ui.unit_index + 1, &ui);
      if (rc != RT_OK) break;                                   /* no boundary behind this row: the last part takes the rest */
      if (ui.row0 > cut[parts - 1]) cut[parts++] = ui.row0; }
   P = parts; cut[P] = UINT64_MAX;
   if (P <= 1) WRET;
   S.s_open += wall() - w0;
   static struct chan ch[256];
   static char base[MAXPATH + 50];
   strcpy(base, baseoutfilename);
   rt_event *bufs = mmap(NULLP, (size_t)P * WORKER_BUF_EVENTS * sizeof(rt_event), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
   assert(bufs != MAP_FAILED, "cannot map the workers' event buffers");
   /* the parameter set each worker hands over with (see the hand-over in readblock): -1 = none reported */
   volatile int *final_ps = mmap(NULLP, 4096 + (size_t)P * sizeof(int), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
   assert(final_ps != MAP_FAILED, "cannot map the workers' hand-over slots");
   for (int i = 0; i < P; ++i) final_ps[i] = -1;
   const int ps_start = block.parmset;                          /* what every worker starts trying with */
   fflush(NULLP);
   const long long inf_pos0 = ftello(inf);                      /* where process_file() is in the input file */
   for (int i = 0; i < P; ++i) {
      int sv[2];
      assert(socketpair(AF_UNIX, SOCK_STREAM, 0, sv) == 0, "socketpair failed");
      pid_t pid = fork();
      assert(pid >= 0, "fork failed");
      if (pid == 0) {                                           /* the worker: carries on in readblock() with its part */
         char name[MAXPATH + 80];
         for (int k = 0; k < i; ++k) close(ch[k].fd);
         close(sv[0]);
         on_exit(worker_exit, NULLP);
         {  /* An input stream of its own.  fork() leaves all processes on ONE open file description, hence on one file offset: the
               reference keeps its place in the capture with ftello / fseeko on `inf` (save_file_position, readtape.c:1127-1140),
               and a seek of one worker moved the others (rare, timing dependent: "unexpected file position", or a block decoded
               from the wrong row; found by the worker fuzz on the CPU simulation at the end of round 2). */
            char pth[64]; snprintf(pth, sizeof pth, "/proc/self/fd/%d", fileno(inf));
            FILE *own = fopen(pth, "rb");
            if (!own) own = fopen(indatafilename, "rb");
            assert(own != NULLP, "a worker cannot reopen the input file");
            assert(fseeko(own, inf_pos0, SEEK_SET) == 0, "fseek failed");
            inf = own; }                                        /* the inherited stream is left alone (closing it would seek the shared offset) */
         setenv("RT_BRIDGE", "0", 1);                           /* lookups must stay on the host: what a bridge would prove, the parent scans */
         S.worker = i; S.nworkers = P; S.remote_fd = sv[1]; S.remote_buf = bufs + (size_t)i * WORKER_BUF_EVENTS; S.handover_ps = &final_ps[i];
         S.start_row = cut[i]; S.stop_row = cut[i + 1]; S.must_seek_start = i > 0;
         S.ctx = NULLP; S.ctx_valid = 0;
         S.n_events = S.n_bulk_hits = S.n_bulk_miss = S.n_restarts = S.n_exact_spans = 0; S.s_scan = S.s_replay = 0;
         if (i > 0) numblks = 1 << 20;                          /* "wrote block 1" (readtape.c:1271) is the first worker's line */
         part_name(name, sizeof name, i, ".out");
         assert(freopen(name, "w", stdout) != NULLP, "cannot create %s", name);
         part_name(name, sizeof name, i, "");
         assert(strlen(name) < MAXPATH, "output name too long");
         strcpy(baseoutfilename, name);                         /* create_datafile() -> <out>.partNNN.tap */
         return; }
      close(sv[1]);
      ch[i].fd = sv[0]; ch[i].pid = pid; ch[i].alive = 1; ch[i].buf = bufs + (size_t)i * WORKER_BUF_EVENTS; ch[i].ctx = NULLP; }
   /* the parent: serve exact scans until every worker has hung up */
   for (int live = P; live > 0;) {
      static struct pollfd pf[256];
      for (int i = 0; i < P; ++i) { pf[i].fd = ch[i].alive ? ch[i].fd : -1; pf[i].events = POLLIN; pf[i].revents = 0; }
      if (poll(pf, (nfds_t)P, -1) < 0) continue;
      for (int i = 0; i < P; ++i)
         if (ch[i].alive && (pf[i].revents & (POLLIN | POLLHUP | POLLERR))) {
            serve_one(&ch[i]);
            if (!ch[i].alive) { close(ch[i].fd); --live; } } }
   int worst = 0;
   for (int i = 0; i < P; ++i) {
      int st = 0;
      assert(waitpid(ch[i].pid, &st, 0) == ch[i].pid, "waitpid failed");
      int code = WIFEXITED(st) ? WEXITSTATUS(st) : 99;
      if (code != 0 && (worst == 0 || worst == WORKER_UNPROVEN)) worst = code;
      if (ch[i].ctx) rt_scan_end(ch[i].ctx); }
   munmap(bufs, (size_t)P * WORKER_BUF_EVENTS * sizeof(rt_event));
   /* Every block is tried with `starting_parmset` first (readtape.c:1722), which stays 0 unless the reference is compiled with
      USE_ALL_PARMSETS (then it rotates from block to block, readtape.c:1875-1878, and the try order -- hence which clean decoding a
      block stops at -- would depend on how many blocks came before).  A cheap guard for that build: every worker must have ended on
      the set its neighbour started trying with, else no split. */
   if (worst == 0 && multiple_tries)
      for (int i = 0; i + 1 < P; ++i) if (final_ps[i] != ps_start) worst = WORKER_UNPROVEN;
   munmap((void *)final_ps, 4096 + (size_t)P * sizeof(int));
   bool all_ok = true;
   char name[MAXPATH + 80], line[MAXLINE];
   if (worst == 0) {
      for (int i = 0; i < P; ++i) {                             /* the parts, in order, without their end-of-medium markers */
         part_name(name, sizeof name, i, ".tap");
         FILE *f = fopen(name, "rb");
         if (!f) continue;
         if (!outf) create_datafile(NULLP);
         assert(fseeko(f, 0, SEEK_END) == 0, "fseek failed");
         long long len = ftello(f);
         unsigned char tail[4] = {0, 0, 0, 0};
         if (len >= 4) { assert(fseeko(f, len - 4, SEEK_SET) == 0 && fread(tail, 1, 4, f) == 4, "cannot read %s", name);
                         if (tail[0] == 0xff && tail[1] == 0xff && tail[2] == 0xff && tail[3] == 0xff) len -= 4; }
         assert(fseeko(f, 0, SEEK_SET) == 0, "fseek failed");
         static char buf[1 << 20];
         for (long long left = len; left > 0;) {
            size_t want = left > (long long)sizeof buf ? sizeof buf : (size_t)left;
            assert(fread(buf, 1, want, f) == want && fwrite(buf, 1, want, outf) == want, "cannot copy %s", name);
            left -= (long long)want; }
         numoutbytes += len;
         fclose(f); } }
   for (int i = 0; i < P; ++i) {                                /* what the workers printed; their verdicts */
      part_name(name, sizeof name, i, ".out");
      FILE *f = fopen(name, "r");
      if (f) {
         size_t tag = strlen(baseinfilename);
         while (fgets(line, MAXLINE, f)) {
            if (strncmp(line, baseinfilename, tag) == 0 && line[tag] == ':') { if (strstr(line + tag, "bad")) all_ok = false; continue; }
            if (worst != WORKER_UNPROVEN) fputs(line, stdout); }
         fclose(f); }
      remove(name);
      part_name(name, sizeof name, i, ".tap"); remove(name); }
   if (worst == WORKER_UNPROVEN) {                              /* decode the reel in one piece after all: the tape and its scans are here */
      if (getenv("RT_STATS")) printf("  B200 scan: a worker could not prove its hand-over; decoding the reel unsplit\n");
      assert(fseeko(inf, inf_pos0, SEEK_SET) == 0, "fseek failed");
      return; }
   if (worst != 0) exit(worst);
   if (getenv("RT_STATS")) printf("  B200 scan: %d workers; parent: %.3f s opening + upload + scan\n", P, S.s_open);
   /* what process_file() does at the end of the file (readtape.c:1862-1867) and main() in quiet mode (:2014) */
   if (tap_format && outf) output_tap_marker(0xffffffffl);
   close_file();
   printf("%s: %s\n", baseinfilename, all_ok ? "ok" : "bad");
   fflush(NULLP);
   exit(0); }

bool readblock(bool retry) {
   double w0 = wall();
   if (!S.par_checked) { S.par_checked = 1; run_workers(); }      /* RT_WORKERS: the parent does not come back from there */
   if (!S.opened) { open_tape(); S.s_open += wall() - w0; w0 = wall(); }
   if (S.worker > 0 && S.must_seek_start) {
      /* a worker's first call (other than the first worker's): move to its first row and report "noise", so that process_file()
         takes its next block start (blockstart, readtape.c:1722) from there */
      S.must_seek_start = 0;
      timenow_ns = (int64_t)(S.desc.tstart_ns + S.start_row * S.desc.tdelta_ns);
      timenow = rowtime(S.start_row);
      assert(fseeko(inf, S.base_pos + (long long)S.start_row * S.group_bytes, SEEK_SET) == 0, "fseek failed");
      block.results[block.parmset].blktype = BS_NOISE;
      S.pending_reset = RT_RESET_NONE;
      return true; }
   long long pos = ftello(inf);
   assert(pos >= S.base_pos && (pos - S.base_pos) % S.group_bytes == 0, "B200 scan: unexpected file position %lld", pos);
   uint64_t row0 = (uint64_t)((pos - S.base_pos) / S.group_bytes);
   int reset_kind = S.pending_reset; S.pending_reset = RT_RESET_NONE;
   samples_per_bit = bpi > 0 ? (int)(1 / (bpi * ips * sample_deltat)) : 20;     /* readtape.c:1402 */
   rt_scan_cfg cfg; make_cfg(&cfg);
   if (!block.window_set) {                                                   /* readtape.c:1453-1501 */
      pkww_width = rt_pkww_width(&cfg, (uint64_t)sample_deltat_ns);
      if (row0 < S.nrows) timenow = rowtime(row0);
      if (torigin == 0) torigin = timenow;
      say_configuration();
      block.window_set = true; }

   struct evsrc src;
   uint64_t last_row = row0; bool endfile = false;
   int persistent = mode == WW || reset_kind != RT_RESET_FULL;   /* Whirlwind: the scan state carries over from block to block */
   int from_bulk = !persistent && bulk_start(&src, &cfg, row0), exact_started = 0;
   if (S.stop_row != UINT64_MAX && !retry) {                     /* a worker that is not the last: is this block the next worker's? */
      /* The next worker starts with a fresh reset at S.stop_row, the first row of a unit.  This worker may stop in front of a block
         only if that reset PROVABLY sees what a fresh reset here sees: both lookups arrive at the same unit (same end row), so both
         get the same events.  The block in front of us must not be decoded unless it begins on this side of the boundary, or it
         would come out twice -- the lookup may have been served by a unit BEHIND the boundary (directly, or by chaining through
         event-free units), and a miss leaves us with an exact scan that runs on as far as it needs.  Whatever cannot be decided
         ends the split: WORKER_UNPROVEN, the parent decodes the reel in one piece.  (Before the end of round 2 the test was "the
         lookup resolved to a unit at or behind the boundary", which let a block that STRADDLES the boundary -- units can begin
         inside blocks, at all-track drop-outs -- be followed by a worker starting in the middle of it, and a block served through
         a chain across the boundary be decoded by both workers: found by the worker fuzz on the CPU simulation, tests/test_hostsim.py.) */
      if (from_bulk && src.unit_end > S.stop_row) {
         /* "the same unit": the same end row AND the same events (count, first, last) -- the end row alone does not identify a unit
            near the end of the tape, where every unit is clamped to the last row */
         const uint64_t n1 = src.n;
         const uint64_t f1 = n1 ? src.ev[0].row : 0, l1 = n1 ? src.ev[n1 - 1].row : 0;
         const int ft1 = n1 ? src.ev[0].trk : 0, lt1 = n1 ? src.ev[n1 - 1].trk : 0;
         const rt_event *e2 = NULLP; uint64_t n2 = 0, valid2 = 0;
         int rc2 = rt_bulk_lookup(S.bulk[block.parmset].bulk, S.bulk[block.parmset].ci, S.stop_row, &e2, &n2, &valid2);
         if (rc2 != RT_OK && rc2 != RT_MISS) rtfatal("rt_bulk_lookup", rc2);
         if (rc2 == RT_OK && S.stop_row + valid2 == src.unit_end && n2 == n1
             && (n1 == 0 || (e2[0].row == f1 && e2[0].trk == ft1 && e2[n1 - 1].row == l1 && e2[n1 - 1].trk == lt1))) {
            if (S.handover_ps) *S.handover_ps = block.parmset;   /* the set the block in front of us would be tried with first */
            S.s_scan += wall() - w0;
            if (getenv("RT_STATS")) {
               rlog("  B200 scan: worker %d of %d: %lld events, %lld speculative hits, %lld misses, %lld restarts, %lld exact spans\n", S.worker, S.nworkers,
                    S.n_events, S.n_bulk_hits, S.n_bulk_miss, S.n_restarts, S.n_exact_spans);
               rlog("  B200 scan: worker %d: %.3f s opening + upload, %.3f s in the scan library, %.3f s replaying events into the handlers\n", S.worker, S.s_open, S.s_scan, S.s_replay); }
            return false; }                                       /* end of this worker's part: block type BS_NONE, "end of file" */
         /* not the next worker's view: the lookup above has overwritten the library's result buffer, fetch ours again */
         uint64_t valid = 0;
         int rc1 = rt_bulk_lookup(S.bulk[block.parmset].bulk, S.bulk[block.parmset].ci, row0, &src.ev, &src.n, &valid);
         if (rc1 != RT_OK || row0 + valid != src.unit_end) fatal("B200 scan: a repeated lookup at row %llu gave another answer", (unsigned long long)row0); }
      if (row0 >= S.stop_row) { fflush(NULL); _exit(WORKER_UNPROVEN); }
      if (!from_bulk) { exact_start(&src, &cfg, reset_kind, row0); exact_started = 1; }
      /* the block must begin on this side of the boundary */
      if (peek_event(&src, S.stop_row) == NULLP) { fflush(NULL); _exit(WORKER_UNPROVEN); } }
   if (!from_bulk && !exact_started) exact_start(&src, &cfg, reset_kind, row0);
   S.s_scan += wall() - w0; w0 = wall();
   decode_from(row0, reset_kind, &cfg, &src, &last_row, &endfile);
   if (endfile && S.stop_row != UINT64_MAX) {
      /* a worker that is not the last has run into the end of the tape: its block began on its side of the boundary and never
         ended, so the next worker started in the middle of it.  No hand-over was proven: the reel is decoded in one piece. */
      fflush(NULL); _exit(WORKER_UNPROVEN); }
   if (src.exact && persistent && !endfile) {   /* Whirlwind continues from here: leave the scan state exactly where the host stopped */
      int rc = rt_scan_rewind(S.ctx, last_row); if (rc) rtfatal("rt_scan_rewind", rc); }

   /* bookkeeping the reference's loop does per row (readtape.c:1404,1424,1449-1451) */
   uint64_t consumed = last_row - row0;
   numsamples += (long long)consumed;
   if (!retry) lines_in += (long long)consumed + (endfile ? 1 : 0);
   timenow_ns = (int64_t)(S.desc.tstart_ns + last_row * S.desc.tdelta_ns);
   assert(fseeko(inf, S.base_pos + (long long)last_row * S.group_bytes + (endfile ? S.endfile_extra : 0), SEEK_SET) == 0, "fseek failed");

   struct results_t *result = &block.results[block.parmset];                 /* readtape.c:1509-1515 */
   if (S.worker > 0 && mode == GCR && (result->blktype == BS_BLOCK || result->blktype == BS_BADBLOCK)) {
      /* The reference's bit arrays (data[], data_time[]) are never cleared between blocks, and gcr_end_of_block lets a block whose
         tracks differ by up to two bits through to gcr_postprocess (decode_gcr.c:684-729), which then reads what EARLIER blocks
         left behind the end of a short track.  A worker other than the first does not have that history -- its arrays hold only
         its own blocks -- so such a block cannot be proven to come out as in one piece: the split ends.  (NRZI and PE write and
         check a block only up to its shortest track: decode_nrzi.c:35-75, decode_pe.c:96-102, readtape.c:1183,1214.)  Found by the
         worker fuzz on windows of the GCR captures: the last byte of a block cut off by the end of the tape. */
      int lo = INT_MAX, hi = 0;
      for (int k = 0; k < ntrks; ++k) { if (trkstate[k].datacount < lo) lo = trkstate[k].datacount; if (trkstate[k].datacount > hi) hi = trkstate[k].datacount; }
      if (lo != hi) { fflush(NULL); _exit(WORKER_UNPROVEN); } }
   result->errcount = result->track_mismatch + result->vparity_errs + result->ecc_errs + result->crc_errs + result->lrc_errs
                      + result->gcr_bad_sequence + result->ww_bad_length + result->ww_speed_err;
   result->warncount = result->missed_midbits + result->corrected_bits + result->gcr_bad_dgroups
                       + result->ww_leading_clock + result->ww_missing_onebit + result->ww_missing_clock;
   S.s_replay += wall() - w0;
   if (endfile && getenv("RT_STATS")) {
      rlog("  B200 scan: %lld events, %lld speculative hits, %lld misses, %lld restarts, %lld exact spans\n",
           S.n_events, S.n_bulk_hits, S.n_bulk_miss, S.n_restarts, S.n_exact_spans);
      rlog("  B200 scan: %.3f s opening + upload, %.3f s in the scan library, %.3f s replaying events into the handlers\n", S.s_open, S.s_scan, S.s_replay);
      if (S.s_exact > 0) rlog("  B200 scan: %.3f s of the replay time were exact-scan spans continued inside block decodes\n", S.s_exact); }
   return !endfile; }
