/* readtape_b200/csrc/kernels.h -- host-callable launchers of the CUDA kernels. */
#ifndef RT_KERNELS_H
#define RT_KERNELS_H
#include <cuda_runtime.h>
#include "rt_dev.h"

namespace rtgen { struct SkewState; }
struct TrkState;

/* k_ingest.cu */
/* the bit planes of K3c phase A, to be written by the ingest kernel itself (fused) for the tiles it handles */
struct IngestMasks { uint32_t *cand, *cand2, *acan; uint64_t mask_stride; int ntrks, width; int32_t T0[RT_MAXTRKS], T1[RT_MAXTRKS]; };
bool ingest_masks_supported(int nheads, int ntrks, int width);
cudaError_t launch_ingest(const int16_t *src, uint64_t nrows, uint64_t row_base, int nheads, const int32_t *trk_of_head,
                          int16_t *planes, uint64_t plane_stride, int16_t *gmm, uint64_t ngran_cap,
                          unsigned long long *first_end_row, int sms, int force_simple, cudaStream_t st, int *launches,
                          const IngestMasks *mask = nullptr, uint64_t *masked_rows = nullptr);

/* k_units.cu */
struct UnitParams {
   int32_t det;            /* RT_DET_* */
   int32_t thr;            /* peak/dzc: a granule is quiet if max-min <= thr ; zc: if max <= thr and min >= -thr (int16 LSBs) */
   uint32_t min_gap_gran;  /* a gap is >= this many consecutive all-track-quiet granules */
   uint64_t tail_rows;     /* a unit keeps scanning this many rows into the next unit */
};
cudaError_t launch_find_units(const int16_t *gmm, uint64_t ngran_cap, int ntrks, uint64_t nrows, const UnitParams &up,
                              uint32_t *d_bitmap, uint32_t *d_flags, uint32_t *d_blockcount, UnitDesc *d_units,
                              uint32_t units_cap, uint32_t *d_nunits, cudaStream_t st, int *launches);
size_t units_bitmap_words(uint64_t nrows);
size_t units_blocks(uint64_t nrows);

/* k_scan.cu */
void launch_ctx_reset(const DevCfg &c, TrkState *st, rtgen::SkewState *sk, int kind, uint64_t row, int tz, cudaStream_t s);
void launch_ctx_set_avg_height(TrkState *st, int trk, float v, cudaStream_t s);
void launch_ctx_scan(const DevCfg &c, TrkState *st, rtgen::SkewState *sk, uint64_t from, uint64_t to, rt_event *ev,
                     uint32_t cap, uint32_t *counts, uint32_t *failed, cudaStream_t s);
void launch_units_scan(const DevCfg &c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta, rt_event *pool,
                       uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks, float quiet_thr, int quiet_thr_lsb,
                       unsigned long long *rows_scanned, int grid, cudaStream_t s);
/* k_fast.cu: the int16-domain fast path of the moving-window peak detector (same contract as launch_units_scan) */
bool fast_scan_eligible(const DevCfg &c);
cudaError_t launch_units_fast(const DevCfg &c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta, rt_event *pool,
                              uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks, int quiet_thr_lsb,
                              unsigned long long *rows_scanned, int sms, int max_ctas_per_sm, cudaStream_t s);
/* k_sparse.cu: the two-pass scan of the moving-window peak detector (K3c): phase A writes the candidate / canonical bit planes
   DevCfg::m_cand / m_acan for plane rows [row_lo, row_hi), phase B scans the units (same contract as launch_units_fast) */
uint64_t peak_mask_stride(uint64_t plane_stride);
int peak_mask_T0(const DevCfg &c, float frac);
bool sparse_scan_eligible(const DevCfg &c);
uint32_t span_hist_words(int ntrks);
cudaError_t launch_span_hist(const uint32_t *gmm, uint64_t ngran_cap, uint64_t nrows, int ntrks, uint32_t *d_hist, cudaStream_t s);
bool peak_mask_auto_T0(DevCfg &c, const uint32_t *hist);
cudaError_t launch_peak_masks(const DevCfg &c, uint64_t row_lo, uint64_t row_hi, cudaStream_t s);
/* phase B1 (scan_records.cuh): the candidate records of plane rows [row_lo, row_hi) (whole 2048-row tiles) */
struct CandRec;
uint64_t cand_rec_tiles(uint64_t plane_stride);
cudaError_t launch_cand_records(const DevCfg &c, uint64_t row_lo, uint64_t row_hi, CandRec *recs, uint32_t rec_cap, uint32_t *tile_base, uint32_t *tile_cnt,
                                unsigned int *cursor, cudaStream_t s);
cudaError_t launch_units_sparse(const DevCfg &c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta, rt_event *pool,
                                uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks, int quiet_thr_lsb,
                                unsigned long long *counters, int sms, int max_ctas_per_sm, cudaStream_t s);
/* k_units.cu: negated copy of the planes / granule map for -invert */
cudaError_t launch_negate(const int16_t *planes, int16_t *planes_inv, uint64_t plane_stride, const int16_t *gmm, int16_t *gmm_inv, uint64_t ngran_cap,
                          uint64_t row_lo, uint64_t row_hi, int ntrks, cudaStream_t s);
/* k_csv.cu: CSV ingest (line index, csv_preread's maximum, parse + quantise) */
uint32_t csv_index_warps(uint64_t nbytes);      /* scratch entries launch_csv_count / launch_csv_index need: round up to a multiple of 8 */
cudaError_t launch_csv_count(const char *txt, uint64_t n, uint32_t *warp_counts, uint64_t *warp_offsets, uint64_t *total, cudaStream_t s);
cudaError_t launch_csv_index(const char *txt, uint64_t n, const uint64_t *warp_offsets, uint64_t *line_start, uint64_t cap, cudaStream_t s);
cudaError_t launch_csv_maxabs(const char *txt, uint64_t n, const uint64_t *line_start, uint64_t nlines_total, uint64_t first_line, uint64_t nlines,
                              int ntrks, float scalefactor, float *out, cudaStream_t s);
cudaError_t launch_csv_parse(const char *txt, uint64_t n, const uint64_t *line_start, uint64_t nlines_total, int ntrks, const uint32_t *perm,
                             float maxvolts, float scalefactor, int invert, uint32_t subsample, uint64_t first_line, uint64_t nrows,
                             int16_t *rows, unsigned long long *stats, float *fstats, cudaStream_t s);
/* k_digest.cu: per-tile event counts and order-independent digests of a finished whole-tape scan (verification at scale) */
cudaError_t launch_tile_digest(const DevCfg &c, const UnitDesc *units, uint32_t nunits, const TrkMeta *meta, const rt_event *pool,
                               const uint32_t *chunk_next, uint64_t period, uint64_t ntiles, unsigned long long *counts,
                               unsigned long long *digests, unsigned long long *bad, int sms, cudaStream_t s);
#endif
