/* include/rt_scan.h -- C-ABI of the B200 per-track analog scan ("readtape_b200").
 *
 * This is the drop-in boundary for readtape's hot path.  The reference has no plugin/FFI
 * interface; the seam is the internal call
 *       bool readblock(bool retry)                    src/readtape.c:1396
 *   ->  enum bstate_t process_sample(struct sample_t*) src/decoder.c:817  (called at readtape.c:1504)
 * i.e. "read one row of int16 samples, convert to volts, run the per-track detectors".
 * A replacement readblock() (see INTEGRATION.md and readtape_b200/host/readblock_b200.c)
 * obtains flux-transition EVENTS from this library instead of rows, and replays them into
 * readtape's unchanged mode handlers (process_up/down_transition, decoder.c:574/592).
 *
 * Everything here is plain C: pointers, sizes, PODs.  No C++/torch types cross the ABI.
 * Two libraries export exactly this ABI:
 *   readtape_b200/lib/librt_scan_b200.so   the product: hand-written sm_100a CUDA kernels
 *   oracle/_ref/libscan_oracle.so          test-only CPU restatement (oracle/scan_oracle.c)
 *
 * Error convention: every function returns 0 (RT_OK) or a negative RT_ERR_* code; a message
 * is available from rt_last_error().  No exceptions, no callbacks into host code.  The host
 * shim maps a non-zero return to readtape's fatal() (readtape.c:596).
 * Threading: one host thread per rt_tape (the reference host is not re-entrant); all
 * concurrency (streams, multi-GPU, parameter-set fan-out) lives inside the library.
 */
#ifndef RT_SCAN_H
#define RT_SCAN_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 1

#define RT_MAXTRKS        19   /* MAXTRKS              src/csvtbin.h:29   */
#define RT_PKWW_MAX_WIDTH 50   /* PKWW_MAX_WIDTH       src/decoder.h:132  */
#define RT_MAXSKEWSAMP    50   /* MAXSKEWSAMP          src/decoder.h:97   */
#define RT_AGC_MAX_WINDOW 10   /* AGC_MAX_WINDOW       src/decoder.h:152  */
#define RT_CLKRATE_WINDOW 50   /* CLKRATE_WINDOW       src/decoder.h:142  */
#define RT_MAXBLOCK       131072 /* MAXBLOCK           src/decoder.h:91   */
#define RT_HEAD_IGNORE    (RT_MAXTRKS - 1)  /* WWHEAD_IGNORE  src/decoder.h:126 */
#define RT_TBIN_END_MARK  (-32768)          /* end-of-data marker in head 0, src/csvtbin.h:103, readtape.c:1410 */

/* error codes */
#define RT_OK               0
#define RT_ERR_ARG         -1   /* bad argument                                           */
#define RT_ERR_NOMEM       -2   /* host or device allocation failed                       */
#define RT_ERR_CUDA        -3   /* CUDA runtime error (message has the cudaError string)  */
#define RT_ERR_UNSUPPORTED -4   /* a mode/flag combination this library does not scan     */
#define RT_ERR_STATE       -5   /* call sequence error (e.g. run before reset)            */
#define RT_ERR_OVERFLOW    -6   /* internal event pool exhausted even after regrowth      */
#define RT_ERR_NODEVICE    -7   /* no usable CUDA device / kernel image (product lib only)*/
#define RT_MISS             1   /* rt_bulk_lookup: no verified unit; use the exact scan   */

/* enum mode_t values, src/csvtbin.h:46-48 */
#define RT_MODE_PE    0x01
#define RT_MODE_NRZI  0x02
#define RT_MODE_GCR   0x04
#define RT_MODE_WW    0x08

/* rt_scan_cfg.flags : the reference globals the scan reads (src/decoder.h:433-436) */
#define RT_F_FIND_ZEROS     0x01  /* find_zeros      (-zeros)          decoder.c:863         */
#define RT_F_DIFFERENTIATE  0x02  /* do_differentiate(-differentiate)  readtape.c:1383,1422  */
#define RT_F_INVERT         0x04  /* invert_data     (-invert)         readtape.c:1421       */
#define RT_F_DENSITY_DETECT 0x08  /* doing_density_detection: handlers bypassed, decoder.c:578; width 8, clkavg 0 */
#define RT_F_DESKEWING      0x10  /* doing_deskew: skew pre-pass (no effect on the scan itself; kept for logging)  */

/* reset kinds for rt_scan_reset(): the state a block decode starts from */
#define RT_RESET_NONE       0  /* reposition only: keep every bit of per-track state                         */
#define RT_RESET_FULL       1  /* init_trackstate()       decoder.c:425                                       */
#define RT_RESET_WW_PARTIAL 2  /* ww_init_blockstate()    decode_ww.c:33 (t_lastpeak=t_prevlastpeak=0 only)   */
#define RT_RESET_PEAKSTATE  3  /* init_trackpeak_state()  decoder.c:413 (window + skew FIFO cleared)          */

/* event kinds */
#define RT_EV_BOT 0            /* bottom peak / downward zero crossing -> process_down_transition */
#define RT_EV_TOP 1            /* top peak    / upward zero crossing   -> process_up_transition   */

/* Describes the sample stream: what read_tbin_header() (readtape.c:1319) and process_file()
 * (readtape.c:1646-1648) establish before the first readblock(). */
typedef struct rt_tape_desc {
   uint32_t ntrks;                       /* tracks decoded                                   */
   uint32_t nheads;                      /* int16 columns per row (== ntrks except Whirlwind) */
   int32_t  head_to_trk[RT_MAXTRKS];     /* head h feeds track head_to_trk[h]; RT_HEAD_IGNORE = unused head */
   float    maxvolts;                    /* tbin_hdr.u.s.maxvolts                            */
   uint64_t tdelta_ns;                   /* sample_deltat_ns  (tbin_hdr.u.s.tdelta)          */
   uint64_t tstart_ns;                   /* timenow_ns of row 0 (tbin_dat.tstart)            */
} rt_tape_desc;

/* The members of struct parms_t (src/decoder.h:290-310) that the per-sample path reads. */
typedef struct rt_parms {
   int32_t clk_window;    /* moving-window clock average width, 0 = use clk_alpha             */
   float   clk_alpha;     /* EWMA weight for the clock, 0 = constant                          */
   int32_t agc_window;    /* AGC = min of last n peak heights, 0 = use agc_alpha              */
   float   agc_alpha;     /* AGC EWMA weight, 0 = no AGC                                      */
   float   min_peak;      /* minimum absolute peak height, volts                             */
   float   clk_factor;    /* PE clock window factor                                          */
   float   pulse_adj;     /* PE/GCR pulse position adjustment fraction                       */
   float   pkww_bitfrac;  /* peak window width as a fraction of a bit                        */
   float   pkww_rise;     /* required rise, volts                                            */
   float   z1pt, z2pt;    /* GCR zero-bit thresholds                                         */
} rt_parms;

/* Everything one decode attempt of one block needs ("PARM" + the globals), SURVEY 8(b). */
typedef struct rt_scan_cfg {
   int32_t  mode;                          /* RT_MODE_*                                       */
   uint32_t flags;                         /* RT_F_*                                          */
   float    bpi, ips;                      /* globals bpi, ips (bpi==0 only with DENSITY_DETECT) */
   rt_parms parms;
   int32_t  skew_delaycnt[RT_MAXTRKS];     /* skew_delaycnt[], decoder.c:232 (0..50 rows)     */
} rt_scan_cfg;

/* One detected flux transition, exactly what the mode handler of the reference sees when
 * process_up_transition()/process_down_transition() is entered (decoder.c:574/592).
 * Events of one scan are ordered by (row, trk) == the order the reference generates them
 * (row loop readtape.c:1403, track loop decoder.c:847). */
typedef struct rt_event {
   uint64_t row;        /* sample row at which the detector fired (timenow of the reference) */
   double   t_event;    /* t->t_top (kind TOP) or t->t_bot (kind BOT), seconds               */
   float    v_top;      /* t->v_top as the handler sees it                                  */
   float    v_bot;      /* t->v_bot as the handler sees it                                  */
   float    agc_gain;   /* t->agc_gain AFTER the handler (parity self-check for the host)   */
   uint8_t  trk;        /* track number                                                     */
   uint8_t  kind;       /* RT_EV_TOP / RT_EV_BOT                                            */
   uint8_t  pad[2];
} rt_event;             /* 32 bytes */

typedef struct rt_tape rt_tape;   /* a sample stream resident in device memory               */
typedef struct rt_scan rt_scan;   /* a stateful scan context (per-track detector state)       */
typedef struct rt_bulk rt_bulk;   /* result of a whole-tape speculative scan                  */

const char *rt_last_error(void);
/* Process-wide options.  RT_OPT_SHARED_RESULTS (value 0/1): rt_bulk_fetch() places the events in an anonymous MAP_SHARED mapping
 * (pinned only while the copy runs) instead of cudaHostAlloc memory, so that worker processes forked AFTER the fetch can read them.
 * Such children must not call anything of this library that touches the device: only rt_bulk_lookup() with bridge scans off
 * (environment RT_BRIDGE=0), rt_bulk_last_unit(), rt_bulk_unit_info() / rt_bulk_unit_at(). */
#define RT_OPT_SHARED_RESULTS 1
int  rt_set_option(int option, int value);
int  rt_abi_version(void);
/* Name of the implementation behind the ABI: "cuda-sm100a" or "oracle-cpu". */
const char *rt_backend(void);

/* ---- sample stream ------------------------------------------------------------------- */
/* Create a tape on CUDA device `device` (ignored by the oracle).  Replaces fopen()+
 * read_tbin_header() state, readtape.c:1594-1601. */
int  rt_open(const rt_tape_desc *desc, int device, rt_tape **out);
/* Copy `nrows` rows (nheads little-endian int16 each; TBIN payload layout, csvtbin.h:98-105)
 * from HOST memory to the device and pre-process them (de-interleave into track-major planes,
 * quiet-gap map).  May be called repeatedly to append.  `rows` need not be pinned; pinned
 * memory (rt_host_alloc) makes the copy asynchronous.  Replaces the fread()s of
 * readtape.c:1408,1414. */
int  rt_upload(rt_tape *tape, const int16_t *rows, uint64_t nrows);
/* Same, but the rows are read from file descriptor `fd` at byte offset `offset` (pread): the library reads the file (page cache)
 * with a few threads straight into its pinned staging ring, so the caller needs neither a whole-file buffer nor a mapping and the
 * copy engine is kept busy.  Replaces the fread()s of readtape.c:1408,1414 for a whole capture at once. */
int  rt_upload_fd(rt_tape *tape, int fd, uint64_t offset, uint64_t nrows);
/* Same, but `rows_dev` is already a DEVICE pointer (product library only). */
int  rt_attach_device(rt_tape *tape, const void *rows_dev, uint64_t nrows);
/* Optional, speed only.  Announce the configuration the next whole-tape scan (rt_bulk_scan) will use: rows uploaded / attached from
 * now on get the candidate / canonical bit planes of that configuration's peak-detector window (phase A of the two-pass scan)
 * computed by the ingest kernel itself, while their tile is on chip, instead of by a separate pass that re-reads the planes.
 * The per-track mask thresholds are chosen from the first rows of the tape (or kept from the previous tape of this object).
 * Only plain NRZI / PE peak detection on 9-head captures with a window of 6..20 samples is fused; anything else is ignored.
 * cfg == NULL cancels.  Results never depend on it.  EXPERIMENTAL: measured slower than the two separate kernels on a B200 (the mask
 * arithmetic is ALU-bound and gets fewer warps inside the persistent ingest CTA), so it only takes effect with RT_FUSED_MASKS=1. */
int  rt_prepare(rt_tape *tape, const rt_scan_cfg *cfg);
/* Forget the samples but keep the device buffers (re-use the tape for the next capture of similar size). */
int  rt_clear(rt_tape *tape);
uint64_t rt_nrows(const rt_tape *tape);
void rt_close(rt_tape *tape);
/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost; malloc/free in the oracle) */
void *rt_host_alloc(size_t bytes);
void  rt_host_free(void *p);

/* ---- exact, stateful scan: the always-correct primitive -------------------------------- */
/* The context holds, for every track, the state process_sample() keeps in trkstate[]/skew[]
 * (decoder.h:194-255, decoder.c:227-231) that feeds back into the detectors. */
int  rt_scan_begin(rt_tape *tape, const rt_scan_cfg *cfg, rt_scan **out);
/* Position the context at `row` and apply a reset there (RT_RESET_*). */
int  rt_scan_reset(rt_scan *scan, int reset_kind, uint64_t row);
/* Scan rows [pos, pos+nrows) (clamped to the end of the tape; the end marker row is never
 * scanned), advance pos, and return the events in (row,trk) order.  The returned array is
 * owned by the context and valid until the next call on it.  *rows_done = rows scanned. */
int  rt_scan_run(rt_scan *scan, uint64_t nrows, const rt_event **events, uint64_t *nevents,
                 uint64_t *rows_done);
/* Roll the context back to `row`, which must lie inside the span of the last rt_scan_run()
 * (the state at its entry is checkpointed; the span is silently re-scanned up to `row`). */
int  rt_scan_rewind(rt_scan *scan, uint64_t row);
/* Whirlwind: compute_avg_height() (decoder.c:491, readtape.c:1713) runs on the host after the
 * deskew pre-pass; push the result into the scan state. */
int  rt_scan_set_avg_height(rt_scan *scan, uint32_t trk, float v_avg_height);
/* Replace the configuration but KEEP all per-track state (same mode required).  Whirlwind: the
 * skew delays are set after the deskew pre-pass (skew_compute_deskew, decoder.c:243) while
 * trkstate[] carries over into the main pass (readtape.c:1674,1701-1716). */
int  rt_scan_set_cfg(rt_scan *scan, const rt_scan_cfg *cfg);
uint64_t rt_scan_pos(const rt_scan *scan);
void rt_scan_end(rt_scan *scan);

/* ---- whole-tape speculative scan: the fast path ------------------------------------------ */
/* Scans the whole tape for each of `ncfgs` configurations (parameter sets) at once: the tape
 * is cut at all-track quiet gaps into independent units, every (unit, track, cfg) is scanned
 * by its own GPU thread from a fresh RT_RESET_FULL state.  For each unit the kernel also
 * records what is needed to PROVE that a fresh reset at any later row r (the reference's real
 * block start) yields bit-identical events; rt_bulk_lookup() applies that proof.
 * Not for Whirlwind (its state persists across blocks): returns RT_ERR_UNSUPPORTED. */
int  rt_bulk_scan(rt_tape *tape, const rt_scan_cfg *cfgs, uint32_t ncfgs, rt_bulk **out);
/* rt_bulk_scan() leaves its results in device memory; rt_bulk_fetch() copies them to the host (pinned
 * memory).  The first rt_bulk_lookup()/rt_bulk_unit_info() call does it implicitly. */
int  rt_bulk_fetch(rt_bulk *bulk);
/* The results of rt_bulk_scan() as one relocatable image, copied DEVICE to DEVICE (before rt_bulk_fetch): for a caller that moves
 * results between GPUs itself -- the result gather of a reel sharded over several GPUs (NCCL), SURVEY 8(e).  Image layout, all
 * configurations in turn: { u64 nunits, u64 ntrks, UnitDesc[nunits], TrkMeta[nunits*ntrks] } ..., then u64 chunks, u32 chunk_next[chunks]
 * (padded to 8 bytes), rt_event pool[chunks*32].  rt_bulk_results_size() gives its size in bytes. */
int  rt_bulk_results_size(const rt_bulk *bulk, uint64_t *bytes);
int  rt_bulk_results_to_device(const rt_bulk *bulk, void *dst_dev, uint64_t bytes);
/* rt_bulk_fetch() with the events placed in the CALLER's buffer (RT_ERR_OVERFLOW if it is too small: nothing is fetched then and
 * rt_bulk_fetch() can still be called).  For a caller that wants the results in memory of its own choosing -- e.g. a MAP_SHARED
 * mapping prepared in the background that worker processes forked later can read; rt_host_register() pins such memory for the copy
 * (cudaHostRegister), rt_host_unregister() unpins it. */
int  rt_bulk_fetch_to(rt_bulk *bulk, void *events_buf, size_t bytes);
int  rt_host_register(rt_tape *tape, void *p, size_t bytes);
int  rt_host_unregister(rt_tape *tape, void *p);
/* rt_clear() + rt_upload(rows) + rt_bulk_scan(cfg) + rt_bulk_fetch() in one call with the three stages overlapped: the tape is
 * scanned segment by segment while it is still being copied (PCIe is the slow stage), and the events of each segment travel
 * back while the next one arrives.  `rows` should be pinned (rt_host_alloc).  Same results as the separate calls. */
int  rt_bulk_scan_host(rt_tape *tape, const int16_t *rows, uint64_t nrows, const rt_scan_cfg *cfg, rt_bulk **out);
/* Events a fresh RT_RESET_FULL scan of configuration `cfg_index` starting at `start_row`
 * would produce, for rows [start_row, start_row + *valid_rows).  RT_MISS if no unit can be
 * proven equivalent (the caller then uses rt_scan_*). */
int  rt_bulk_lookup(rt_bulk *bulk, uint32_t cfg_index, uint64_t start_row,
                    const rt_event **events, uint64_t *nevents, uint64_t *valid_rows);

/* The unit the last successful rt_bulk_lookup() of configuration `cfg_index` resolved to (before chaining through event-free
 * units): its first row and end.  Diagnostics.  (A caller that splits one tape between workers proves a hand-over with two
 * lookups instead: start_row + *valid_rows is the end row of the unit a lookup ARRIVED at, after chaining; two rows whose lookups
 * arrive at the same unit see the same events -- readblock_b200.c.) */
int  rt_bulk_last_unit(const rt_bulk *bulk, uint32_t cfg_index, uint64_t *row0, uint64_t *row_end);

/* Diagnostics: the unit rt_bulk_lookup() would consult for `start_row` and the per-track proof data
 * (~0 means "none").  A unit covers start_row iff start_row == row0, or start_row >= row0 - P (P = the quiet pre-scan
 * length, 256 .. 4096 rows depending on the sample rate and density) and for every
 * track  sync_row >= need_sync_row  and  (last_loud_row is none or < start_row), or the same with
 * (sync_early, loud_early).  rt_bulk_lookup() tries this unit and the next one. */
typedef struct rt_unit_info {
   uint64_t unit_index, nunits, row0, row_end;
   uint32_t ntrks, pad;
   uint64_t first_event_row[RT_MAXTRKS], sync_row[RT_MAXTRKS], last_loud_row[RT_MAXTRKS], need_sync_row[RT_MAXTRKS];
   uint64_t sync_early[RT_MAXTRKS], loud_early[RT_MAXTRKS];
   uint64_t sync_first[RT_MAXTRKS], quiet_from[RT_MAXTRKS];
   uint32_t nevents[RT_MAXTRKS];
   uint32_t failed[RT_MAXTRKS];
} rt_unit_info;
int  rt_bulk_unit_info(const rt_bulk *bulk, uint32_t cfg_index, uint64_t start_row, rt_unit_info *out);
/* The same for unit number `unit_index` (0 .. nunits-1), with start_row = its own first row; RT_MISS past the end. */
int  rt_bulk_unit_at(const rt_bulk *bulk, uint32_t cfg_index, uint64_t unit_index, rt_unit_info *out);

typedef struct rt_bulk_stats {
   uint64_t rows;            /* rows on the tape                                             */
   uint64_t units;           /* units per cfg                                                */
   uint64_t events;          /* total events over all cfgs                                   */
   uint64_t rows_scanned;    /* sum over (unit,track,cfg) of the rows a lane walked one by one: all of them for the one-pass kernels,
                                only the dense-mode rows (window fill, threshold below the mask threshold) for the two-pass scan */
   uint64_t track_samples;   /* rows * ntrks * ncfgs: the metric's numerator                 */
   double   ms_preprocess;   /* device time: de-interleave + gap map (0 if done at upload)   */
   double   ms_units;        /* device time: unit table construction                         */
   double   ms_scan;         /* device time: scan kernel                                     */
   uint32_t launches;        /* kernels launched by the call                                 */
   uint32_t pad;             /* rt_bulk_scan_host: number of segments streamed, 0 = plain sequence    */
   uint64_t d2h_bytes;       /* bytes rt_bulk_fetch() copied to the host                     */
   double   ms_masks;        /* device time: candidate-mask pass of the two-pass peak scan (a part of ms_scan; 0 if not used) */
   double   ms_records;      /* device time: candidate-record pass (phase B1) of the two-pass peak scan (a part of ms_scan; 0 if not used) */
   uint32_t masks_fused;     /* 1: the mask planes came from the ingest kernel (rt_prepare), ms_masks is 0 and their time is in ms_preprocess */
   uint32_t two_pass;        /* 1: the two-pass peak scan (K3c) was used */
   uint32_t launches_ingest; /* kernels launched on this tape between rt_clear() / rt_open() and this scan: the ingest kernels, and the
                                mask kernels that run beside them after rt_prepare() */
   uint32_t pad2;
} rt_bulk_stats;
int  rt_bulk_get_stats(const rt_bulk *bulk, rt_bulk_stats *out);

/* Verification at scale (product library only; must be called BEFORE rt_bulk_fetch()/rt_bulk_lookup(), while the results are
 * still in device memory).  For a tape that repeats every `period_rows` rows, every block decode starts from a fresh reset, so
 * the events of tile k are those of tile 0 shifted by k * period_rows rows.  For tile i < ntiles: events[i] = number of events
 * with i * period_rows <= row < (i+1) * period_rows (each counted once, by the unit that owns its row), digest[i] = the sum
 * (mod 2^64) over those events of
 *      mix(mix(mix(k0) ^ k1) ^ k2),   mix = the splitmix64 finaliser,
 *      k0 = (row - i*period_rows) | trk << 40 | kind << 48 | (hs & 0xfff) << 52,  k1 = bits(v_top) | bits(v_bot) << 32,  k2 = bits(agc_gain),
 *      hs = round((row time - t_event) / (sample period / 2)).
 * *bad_times = events whose double-precision t_event is not bit-identical to the reference's expression for (row, hs):
 * decoder.c:732 for the peak detector, the row time of row - hs/2 for the zero-crossing detector. */
int  rt_bulk_tile_digest(rt_bulk *bulk, uint32_t cfg_index, uint64_t period_rows, uint64_t ntiles,
                         uint64_t *events, uint64_t *digest, uint64_t *bad_times);
void rt_bulk_free(rt_bulk *bulk);

/* Diagnostics / tests: the two bit planes the two-pass peak scan derives from the samples (pure functions of the window of
 * lookfor_peak, decoder.c:751-775, width = rt_pkww_width(cfg)):  bit (row % 32) of word (row / 32) of track k, for rows
 * [0, rt_nrows), in cand[k * words_per_track ...] / acan[...]:
 *    cand: max(window) - max(left edge, right edge) >= T0  or  min(left edge, right edge) - min(window) >= T0
 *    acan: the sample that left the window at this row was >= max(window)      (the rescan condition of decoder.c:767)
 * T0 = (int)(t0_frac * T), T = the integer bound of required_rise (decoder.c:785) with AGC gain 1 and average height 4 V;
 * it is returned in *t0.  Rows whose window would reach before row 0 have both bits clear.  The oracle computes the
 * definition; the product library runs its mask kernel.  RT_ERR_UNSUPPORTED if the configuration does not use the peak
 * detector or T0 < 16. */
int  rt_peak_masks(rt_tape *tape, const rt_scan_cfg *cfg, float t0_frac, uint32_t *cand, uint32_t *acan,
                   uint64_t words_per_track, int32_t *t0);

/* ---- helpers shared by both libraries ------------------------------------------------- */
/* pkww_width exactly as readtape.c:1453-1457 computes it (float arithmetic, truncation). */
int  rt_pkww_width(const rt_scan_cfg *cfg, uint64_t tdelta_ns);
/* timenow of a row: (double)(tstart_ns + row*tdelta_ns)/1e9, readtape.c:1423-1424. */
double rt_row_time(const rt_tape_desc *desc, uint64_t row);

#ifdef __cplusplus
}
#endif
#endif /* RT_SCAN_H */
