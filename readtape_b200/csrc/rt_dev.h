/* readtape_b200/csrc/rt_dev.h -- structures shared by the CUDA kernels and the host side of
 * librt_scan_b200.so.  See DESIGN.md for the data layout in HBM.
 *
 *  planes   int16 [ntrks][plane_stride]   track-major copy of the TBIN payload, head->track
 *                                         permutation applied (k_ingest.cu)
 *  gmm      int16 [ntrks][ngran][2]       min / max of every 32-row granule (k_ingest.cu)
 *  units    UnitDesc [nunits]             independent scan units cut at all-track quiet gaps
 *  umeta    TrkMeta  [nunits][ntrks]      per (unit, track) results + equivalence proof data
 *  evpool   rt_event [pool_chunks][EVC]   events, in fixed-size chunks chained per (unit, track)
 */
#ifndef RT_DEV_H
#define RT_DEV_H

#include <stdint.h>
#include "rt_scan.h"

#define RT_GRAN            32      /* rows per granule of the quiet map                         */
#define RT_EVC             32      /* events per pool chunk                                     */
#define RT_NOCHUNK         0xffffffffu
#define RT_NOROW           0xffffffffffffffffull
#define RT_PRESCAN_MIN      256     /* DevCfg::prescan_rows: rows before a unit start examined for quietness (at least / at most) */
#define RT_PRESCAN_MAX      4096

/* compile-time constants of the reference the scan needs (src/decoder.h) */
#define RT_PKWW_PEAKHEIGHT 4.0f    /* :133 */
#define RT_DIFF_THRESHOLD  0.05f   /* :135 */
#define RT_DIFF_SCALE      0.4f    /* :136 */
#define RT_ZEROCROSS_PEAK  0.2f    /* :138 */
#define RT_ZEROCROSS_SLOPE 1.5f    /* :139 */
#define RT_PEAK_THRESHOLD  0.005f  /* :141 */
#define RT_AGC_MAX_VALUE   2.0f    /* :153 */
#define RT_AGC_STARTBASE   5       /* :154 */
#define RT_AGC_ENDBASE     15      /* :155 */
#define RT_GCR_IDLE_THRESH 6.00    /* :111, a double in the reference */
#define RT_PE_MIN_PREBITS  70      /* :118 */
#define RT_GCR_MARK1       0x07    /* decode_gcr.c:422 */
#define RT_GCR_MARK2       0x1c    /* decode_gcr.c:423 */

/* detector selected by the flags */
enum { RT_DET_PEAK = 0, RT_DET_ZC = 1, RT_DET_DZC = 2 };

/* Everything a scan kernel needs to know about the tape and the decode configuration. */
struct DevCfg {
   const int16_t *planes;         /* [ntrks][plane_stride] */
   uint64_t plane_stride;         /* elements */
   uint64_t nrows;                /* valid rows (before the end marker) */
   uint64_t tstart_ns, tdelta_ns;
   float    maxvolts;
   float    sample_deltat;        /* (float)tdelta_ns / 1e9f, readtape.c:1345 */
   float    bpi, ips;
   float    clk_init;             /* 1 / (bpi*ips), or 0 while detecting the density */
   int32_t  ntrks;
   int32_t  mode;                 /* RT_MODE_* */
   int32_t  det;                  /* RT_DET_* */
   int32_t  width;                /* pkww_width */
   int32_t  samples_per_bit;      /* readtape.c:1402 */
   int32_t  invert, differentiate, density, find_zeros;
   rt_parms p;
   int32_t  skew[RT_MAXTRKS];
   const uint32_t *gmm;           /* [ntrks][ngran_cap] packed (min | max<<16) of every 32-row granule; null: no gap skipping */
   uint64_t ngran_cap;
   /* K3c (scan_masks.cuh / scan_sparse.cuh): bit planes of phase A, [ntrks][mask_stride] words, bit p%32 of word p/32 = plane row p;
      T0[k] = the integer threshold the planes of track k were built for (0: not built) */
   const uint32_t *m_cand, *m_acan;
   uint64_t mask_stride;
   int32_t  T0[RT_MAXTRKS];
   /* a second candidate plane for a higher threshold T1[k] >= T0[k] (0: none): once the AGC has settled, required_rise is usually
      well above its default-state value, and the sparse scan then follows the plane with fewer candidates */
   const uint32_t *m_cand2;
   int32_t  T1[RT_MAXTRKS];
   /* rows in front of a unit's first row that every scan examines for quietness: the reference starts looking for the next block
      where the previous one ended plus its inter-block skip, which on real tapes lies up to an inter-block-gap time in front of
      the first all-quiet granule the unit finder cuts at (decaying noise behind a block) */
   int32_t  prescan_rows;
   /* K3c phase B1 (scan_records.cuh): one record per candidate row of the T0 plane; records of track k, tile t (2048 plane rows) are
      recs[rec_tile_base[k * rec_tiles + t] ...], rec_tile_cnt[...] of them, in row order.  Null: the sparse scan reads the planes. */
   const struct CandRec *recs;
   const uint32_t *rec_tile_base, *rec_tile_cnt;
   uint64_t rec_tiles;
};

/* Per-track detector + feedback state: the device mirror of the parts of struct trkstate_t
 * (decoder.h:194-255) and struct skew_t (decoder.c:227) that feed back into the detectors. */
struct TrkState {
   /* replaces the `t_lastpeak == 0` test + `break` of decoder.c:855-861 (quirk Q2): the row at which
      this track (re)initialises after a reset; rows before it are skipped; RT_NOROW = done */
   uint64_t init_row;
   /* the exact stateful scan may jump over rows that cannot fire (k_scan.cu: skip-ahead) once the window and the deskew FIFO hold
      nothing but the last consecutive samples of the plane: true from this row on (set by every reset / re-initialisation) */
   uint64_t pure_from;
   /* times */
   double   t_top, t_bot, t_lastpeak, t_firstzero, t_lastzero;
   /* voltages */
   float    v_prev, v_top, v_bot, v_lasttop, v_lastbot;
   /* peak window */
   float    win[RT_PKWW_MAX_WIDTH];
   float    minv, maxv;
   int32_t  left, right, countdown;
   /* AGC */
   float    avg_height, avg_height_sum, agc_gain;
   float    heights[RT_AGC_MAX_WINDOW];
   int32_t  avg_height_count, heightndx, peakcount;
   /* PE */
   float    t_clkwindow;
   /* GCR clock */
   float    clk_spacing[RT_CLKRATE_WINDOW];
   float    clk_avg, t_peakdelta, t_peakdeltaprev, t_pulse_adj;
   int32_t  clk_ndx, datacount, resync_bitcount;
   uint8_t  up_pending, dn_pending, datablock, bit1_up, lastbits, bit_m1, bit_m2, failed;
};

/* One independent scan unit of the speculative whole-tape scan. */
struct UnitDesc {
   uint64_t row0;                 /* fresh RT_RESET_FULL happens here                        */
   uint64_t row_end;              /* scan rows [row0, row_end)                               */
};

/* Per (unit, track) result. */
struct TrkMeta {
   uint64_t first_event_row;      /* RT_NOROW if the track had no event                      */
   uint64_t sync_row;             /* LAST quiet row before the first event at which the state became canonical, RT_NOROW if none */
   uint64_t last_loud_row;        /* last row before sync_row that violates the quiet rule, RT_NOROW if none */
   uint64_t sync_early;           /* last canonical quiet row of the FIRST quiet stretch of the unit (frozen at the first loud row after it) */
   uint64_t loud_early;           /* last loud row before sync_early (from the pre-scan), RT_NOROW if none */
   uint64_t sync_first;           /* FIRST such row, with the unit quiet from its start up to it (used to chain units), RT_NOROW if none */
   uint64_t quiet_from;           /* earliest row q <= row0 such that no row in [q, row0] is loud: a reset at any row in [q, row0) is covered too */
   /* the tail rule (rt_bulk_lookup): a fresh reset at a row s behind the unit's last event, with no loud row in [s, row_end),
      finds nothing up to row_end either (a scan without events stays in default state, and "not loud" proves it cannot fire) */
   uint64_t last_event_row;       /* RT_NOROW if the track had no event                      */
   uint64_t quiet_tail_from;      /* earliest row q such that no row in [q, row_end) is loud; RT_NOROW: not examined (then the rule is off) */
   uint32_t first_chunk;          /* head of the chunk chain in the event pool               */
   uint32_t nevents;
   uint32_t failed;               /* 2: the reference would have called fatal() (peak not found) */
   uint32_t pad;
};

#endif
