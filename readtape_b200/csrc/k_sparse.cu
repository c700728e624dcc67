/* readtape_b200/csrc/k_sparse.cu -- K3c: the speculative whole-tape scan of the moving-window peak detector in two passes.
 *
 *  k_peak_masks<W>   phase A (scan_masks.cuh): one thread per run of 64 rows of one track; int16x2 SIMD sliding max / min by
 *                    doubling in registers; writes the `cand` and `acan` bit planes (2 bits per track-sample).  HBM-bound:
 *                    2 B read + 0.25 B written per track-sample, no divergence, no shared memory.
 *  k_units_sparse    phase B (scan_sparse.cuh): one lane per (unit, track) job, fetched dynamically; visits candidate rows only.
 *                    Same outputs as k_units_fast / k_units_scan: events into the chunk pool, TrkMeta with the proof data.
 */
#include <stdlib.h>
#include "scan_sparse.cuh"
#include "kernels.h"
#include "emit.cuh"

#define MASK_THREADS   128
#define SPARSE_THREADS 128

using namespace rtsparse;

/* grid: x = groups of MASK_THREADS runs, y = track.  Runs whose halo would reach in front of the plane use the scalar path. */
struct TrackT0 { int32_t v[RT_MAXTRKS]; };                      /* per-track mask threshold, passed by value */

template <int W>
__global__ void __launch_bounds__(MASK_THREADS)
k_peak_masks(const int16_t *planes, uint64_t plane_stride, uint64_t run_lo, uint64_t nruns, const __grid_constant__ TrackT0 t0s,
             const __grid_constant__ TrackT0 t1s, uint32_t *cand, uint32_t *cand2, uint32_t *acan, uint64_t mask_stride) {
   const uint64_t r = (uint64_t)blockIdx.x * MASK_THREADS + threadIdx.x;
   if (r >= nruns) return;
   const int trk = blockIdx.y;
   const uint32_t T0 = (uint32_t)t0s.v[trk], T1 = (uint32_t)t1s.v[trk];
   const int16_t *plane = planes + (size_t)trk * plane_stride;
   const int64_t p0 = (int64_t)(run_lo + r) * rtmask::MASK_RUN;
   uint32_t cw[2], dw[2], aw[2];
   if (p0 >= rtmask::RunMasks<W>::HALO) rtmask::RunMasks<W>::run(plane, p0, T0, T1, cw, dw, aw);
   else {
      rtmask::word_masks_scalar(plane, p0 / 32, W, (int)T0, (int)T1, &cw[0], &dw[0], &aw[0]);
      rtmask::word_masks_scalar(plane, p0 / 32 + 1, W, (int)T0, (int)T1, &cw[1], &dw[1], &aw[1]); }
   const size_t wi = (size_t)trk * mask_stride + (size_t)(p0 / 32);
   *reinterpret_cast<uint2 *>(cand + wi) = make_uint2(cw[0], cw[1]);
   *reinterpret_cast<uint2 *>(cand2 + wi) = make_uint2(dw[0], dw[1]);
   *reinterpret_cast<uint2 *>(acan + wi) = make_uint2(aw[0], aw[1]);
   if (r == nruns - 1) {                                       /* the slack words behind the last run: readers fetch up to 64 bits past a row */
      for (int k = 2; k < 6 && (size_t)(p0 / 32) + k < mask_stride; ++k) cand[wi + k] = cand2[wi + k] = acan[wi + k] = 0u; } }

template <int W>
static cudaError_t launch_masks_w(const DevCfg &c, uint64_t run_lo, uint64_t nruns, uint32_t *cand, uint32_t *cand2, uint32_t *acan, cudaStream_t s) {
   dim3 grid((unsigned)((nruns + MASK_THREADS - 1) / MASK_THREADS), (unsigned)c.ntrks);
   TrackT0 t0s, t1s;
   for (int k = 0; k < RT_MAXTRKS; ++k) { t0s.v[k] = c.T0[k] > 0 ? c.T0[k] : 65535; t1s.v[k] = c.T1[k] > 0 ? c.T1[k] : 65535; }
   k_peak_masks<W><<<grid, MASK_THREADS, 0, s>>>(c.planes, c.plane_stride, run_lo, nruns, t0s, t1s, cand, cand2, acan, c.mask_stride);
   return cudaGetLastError(); }

/* phase A over plane rows [row_lo, row_hi) (rounded outwards to whole runs; row_hi <= plane_stride) */
cudaError_t launch_peak_masks(const DevCfg &c, uint64_t row_lo, uint64_t row_hi, cudaStream_t s) {
   if (row_hi <= row_lo) return cudaSuccess;
   const uint64_t run_lo = row_lo / rtmask::MASK_RUN, run_hi = (row_hi + rtmask::MASK_RUN - 1) / rtmask::MASK_RUN;
   const uint64_t nruns = run_hi - run_lo;
   uint32_t *cand = const_cast<uint32_t *>(c.m_cand), *cand2 = const_cast<uint32_t *>(c.m_cand2), *acan = const_cast<uint32_t *>(c.m_acan);
   switch (c.width) {
#define MW(W) case W: return launch_masks_w<W>(c, run_lo, nruns, cand, cand2, acan, s);
      MW(3) MW(4) MW(5) MW(6) MW(7) MW(8) MW(9) MW(10) MW(11) MW(12) MW(13) MW(14) MW(15) MW(16) MW(17) MW(18) MW(19) MW(20)
      MW(21) MW(22) MW(23) MW(24) MW(25) MW(26) MW(27) MW(28) MW(29) MW(30) MW(31) MW(32) MW(33) MW(34) MW(35) MW(36) MW(37) MW(38)
      MW(39) MW(40) MW(41) MW(42) MW(43) MW(44) MW(45) MW(46) MW(47) MW(48) MW(49) MW(50)
#undef MW
      default: return cudaErrorInvalidValue; } }

/* words per track of a mask plane for a plane of `plane_stride` rows (+ slack for unaligned 32-bit reads) */
uint64_t peak_mask_stride(uint64_t plane_stride) { return (plane_stride + 31) / 32 + 4; }

/* the integer threshold the masks are built for: a fraction of the default-state bound (AGC gain 1, average height 4 V).
   0 = the two-pass scan is not worthwhile / not applicable for this configuration */
int peak_mask_T0(const DevCfg &c, float frac) {
   const float inv_lsb = 32767.0f / c.maxvolts;
   const float q = c.p.pkww_rise * inv_lsb * 0.999f - 2.0f;
   if (!(q > 0)) return 0;
   const int T = q > 70000.0f ? 70000 : (int)q;
   const int T0 = (int)((float)T * frac);
   return T0 < 16 ? 0 : (T0 > 65535 ? 65535 : T0); }

/* ---- phase B1: the candidate records (scan_records.cuh) ---------------------------------------------------------------------
 * One warp per (track, 2048-row tile): lane l owns mask words 2l and 2l+1 of the tile.  The tile's records are placed in the pool
 * by ONE atomic add (their order across tiles does not matter, rec_tile_base finds them); inside the tile a warp prefix sum over
 * the lanes' candidate counts keeps them in row order.  Each lane then builds the records of its own candidates. */
#define REC_WARPS 4
__global__ void __launch_bounds__(32 * REC_WARPS)
k_cand_records(const int16_t *planes, uint64_t plane_stride, const uint32_t *cand, const uint32_t *acan, uint64_t mask_stride, int w,
               const __grid_constant__ TrackT0 t0s, uint32_t tile_lo, uint32_t ntiles, uint64_t rec_tiles, uint64_t nrows, CandRec *recs, uint32_t rec_cap,
               uint32_t *tile_base, uint32_t *tile_cnt, unsigned int *cursor) {
   const uint32_t tl = blockIdx.x * REC_WARPS + (threadIdx.x >> 5);
   if (tl >= ntiles) return;
   const uint32_t tile = tile_lo + tl;
   const int trk = blockIdx.y, lane = threadIdx.x & 31;
   const int16_t *plane = planes + (size_t)trk * plane_stride;
   const uint32_t *mc = cand + (size_t)trk * mask_stride, *ma = acan + (size_t)trk * mask_stride;
   const uint64_t w0 = (uint64_t)tile * (RT_REC_TILE / 32) + 2u * (uint32_t)lane;
   const uint64_t nwords = (nrows + 31) / 32;
   uint32_t c0 = w0 < nwords ? mc[w0] : 0u, c1 = w0 + 1 < nwords ? mc[w0 + 1] : 0u;
   if (w0 * 32 + 31 >= nrows) c0 &= w0 * 32 < nrows ? (0xffffffffu >> (31 - (int)((nrows - 1) & 31))) : 0u;       /* rows past the data */
   if ((w0 + 1) * 32 + 31 >= nrows) c1 &= (w0 + 1) * 32 < nrows ? (0xffffffffu >> (31 - (int)((nrows - 1) & 31))) : 0u;
   const uint32_t n = (uint32_t)__popc(c0) + (uint32_t)__popc(c1);
   uint32_t incl = n;
#pragma unroll
   for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
   const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
   uint32_t base = 0;
   if (lane == 0 && total) base = atomicAdd(cursor, total);
   base = __shfl_sync(0xffffffffu, base, 0);
   if (lane == 0) {
      const bool fits = (uint64_t)base + total <= rec_cap;
      tile_base[(size_t)trk * rec_tiles + tile] = base;
      tile_cnt[(size_t)trk * rec_tiles + tile] = fits ? total : 0u; }     /* on overflow the host regrows the pool and reruns */
   if ((uint64_t)base + total > rec_cap) return;
   uint32_t at = base + incl - n;
   const int T0 = t0s.v[trk];
   for (int half = 0; half < 2; ++half) {
      uint32_t bits = half ? c1 : c0;
      while (bits) {
         const int b = __ffs((int)bits) - 1; bits &= bits - 1;
         recs[at++] = rtrec::make_record(plane, ma, w, T0, (w0 + (uint64_t)half) * 32 + (uint64_t)b); } } }

cudaError_t launch_cand_records(const DevCfg &c, uint64_t row_lo, uint64_t row_hi, CandRec *recs, uint32_t rec_cap, uint32_t *tile_base, uint32_t *tile_cnt,
                                unsigned int *cursor, cudaStream_t s) {
   if (row_hi <= row_lo) return cudaSuccess;
   const uint32_t tile_lo = (uint32_t)(row_lo / RT_REC_TILE), tile_hi = (uint32_t)((row_hi + RT_REC_TILE - 1) / RT_REC_TILE);
   const uint32_t ntiles = tile_hi - tile_lo;
   TrackT0 t0s;
   for (int k = 0; k < RT_MAXTRKS; ++k) t0s.v[k] = c.T0[k] > 0 ? c.T0[k] : 65535;
   dim3 grid((ntiles + REC_WARPS - 1) / REC_WARPS, (unsigned)c.ntrks);
   k_cand_records<<<grid, 32 * REC_WARPS, 0, s>>>(c.planes, c.plane_stride, c.m_cand, c.m_acan, c.mask_stride, c.width, t0s, tile_lo, ntiles, c.rec_tiles, row_hi,
                                                  recs, rec_cap, tile_base, tile_cnt, cursor);
   return cudaGetLastError(); }
uint64_t cand_rec_tiles(uint64_t plane_stride) { return plane_stride / RT_REC_TILE + 2; }

struct SparseJobs {
   const DevCfg &c; const UnitDesc *units; TrkMeta *meta; rt_event *pool; uint32_t *chunk_next; unsigned int *cursor; uint32_t cap_chunks;
   int quiet_thr_lsb; unsigned long long *counters /* [0] rows, [1] events, [2] next job group */; uint64_t total; uint64_t cur; bool exhausted;
   template <class Scan>
   __device__ bool next(Scan &us) {                           /* warp-collective: groups of 32 consecutive jobs, fetched dynamically */
      const int lane = threadIdx.x & 31;
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&counters[2], 32ull);
      base = __shfl_sync(0xffffffffu, base, 0);
      exhausted = base >= total;
      cur = base + (unsigned)lane;
      if (cur >= total) return false;
      const uint32_t u = (uint32_t)(cur / c.ntrks); const int trk = (int)(cur % c.ntrks);
      const UnitDesc ud = units[u];
      if (ud.row_end - ud.row0 > (1ull << 30)) {            /* offsets are 32-bit: leave such a unit to the exact scan */
         TrkMeta m;
         m.first_event_row = m.sync_row = m.last_loud_row = m.sync_early = m.loud_early = m.sync_first = RT_NOROW;
         m.quiet_from = ud.row0; m.first_chunk = RT_NOCHUNK; m.nevents = 0; m.failed = 3; m.pad = 0;
         m.last_event_row = m.quiet_tail_from = RT_NOROW;
         meta[cur] = m;
         return false; }
      PoolEmit em{pool, chunk_next, cursor, cap_chunks, RT_NOCHUNK, RT_NOCHUNK, 0, RT_NOROW, (uint8_t)trk, RT_NOROW};
      us.begin(c.planes + (size_t)trk * c.plane_stride, ud.row0, ud.row_end, trk, em, quiet_thr_lsb);
      return true; }
   template <class Scan>
   __device__ void done(Scan &us) {
      TrkMeta m; us.finish(m); meta[cur] = m;
      if (us.ndense) atomicAdd(&counters[0], (unsigned long long)us.ndense);      /* rows walked one by one (dense mode) */
      if (us.em.n) atomicAdd(&counters[1], (unsigned long long)us.em.n); } };

struct WarpAny { __device__ bool operator()(bool p) const { return __any_sync(0xffffffffu, p); } };

/* MINB = CTAs per SM the register allocation aims at (the lane state is large: 4 -> 128 registers, 6 -> 80 with some cold
   state spilled); which one is faster is a latency-hiding question, decided by measurement (RT_SPARSE_OCC) */
template <int MINB, bool REC>
__global__ void __launch_bounds__(SPARSE_THREADS, MINB)
k_units_sparse(DevCfg c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta,
               rt_event *pool, uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks,
               int quiet_thr_lsb, unsigned long long *counters) {
   __shared__ uint32_t heights[RT_AGC_MAX_WINDOW * SPARSE_THREADS];      /* v_heights[] of every lane, [entry][thread] */
   const uint64_t total = (uint64_t)nunits * (uint64_t)c.ntrks;
   SparseJobs jobs{c, units, meta, pool, chunk_next, cursor, cap_chunks, quiet_thr_lsb, counters, total, 0, false};
   SparseScan<SPARSE_THREADS, PoolEmit, REC> us(c, heights + threadIdx.x);
   drive_sparse(us, jobs, WarpAny()); }

bool sparse_scan_eligible(const DevCfg &c) {
   for (int k = 0; k < c.ntrks; ++k) if (c.T0[k] <= 0) return false;
   return c.det == RT_DET_PEAK && (c.mode == RT_MODE_NRZI || c.mode == RT_MODE_PE || c.density) && !c.invert && !c.differentiate
          && c.width >= 3 && c.width <= RT_PKWW_MAX_WIDTH && c.m_cand && c.m_cand2 && c.m_acan; }

/* ---- choosing T0 from the data ----------------------------------------------------------------------------------------------
 * required_rise (decoder.c:785) follows the signal: pkww_rise * (average peak-to-peak height / 4 V) / AGC gain.  A mask threshold
 * above the bound T of the moment sends the sparse scan into its slow row-by-row mode; one far below it floods it with candidates.
 * k_span_hist builds, per track, a histogram of the peak-to-peak span (max - min) of a sample of 128-row stretches (4 granules of
 * the quiet map: several bit cells, so that both polarities are inside); the host (peak_mask_auto_T0) takes the full-scale height A
 * (99.8th percentile), calls everything above A/4 signal and uses the 10th percentile of that as the smallest height to expect,
 * halved once more because the AGC gain may reach 2.  Only speed depends on the choice, never a result. */
#define HIST_BINS 1024                                          /* 64 LSB per bin */
__global__ void __launch_bounds__(256)
k_span_hist(const uint32_t *gmm, uint64_t ngran_cap, uint64_t ngroups, uint32_t nsamples, uint32_t *hist) {
   __shared__ uint32_t h[HIST_BINS];
   for (int i = threadIdx.x; i < HIST_BINS; i += 256) h[i] = 0;
   __syncthreads();
   const uint32_t *g = gmm + (size_t)blockIdx.x * ngran_cap;
   const uint64_t stride = ngroups / nsamples > 0 ? ngroups / nsamples : 1;
   for (uint32_t i = blockIdx.y * 256 + threadIdx.x; i < nsamples; i += gridDim.y * 256) {
      const uint64_t gi = (uint64_t)i * stride;
      if (gi >= ngroups) break;
      const uint4 q = *reinterpret_cast<const uint4 *>(g + 4 * gi);            /* 4 granules: 16-byte aligned */
      const uint32_t v[4] = {q.x, q.y, q.z, q.w};
      int mn = 32767, mx = -32768;
      for (int k = 0; k < 4; ++k) {
         const int a = (int)(int16_t)(uint16_t)(v[k] & 0xffffu), b = (int)(int16_t)(uint16_t)(v[k] >> 16);
         if (a < mn) mn = a; if (b > mx) mx = b; }
      atomicAdd(&h[min(HIST_BINS - 1, max(mx - mn, 0) >> 6)], 1u); }
   __syncthreads();
   for (int i = threadIdx.x; i < HIST_BINS; i += 256) if (h[i]) atomicAdd(&hist[(size_t)blockIdx.x * HIST_BINS + i], h[i]); }

uint32_t span_hist_words(int ntrks) { return (uint32_t)ntrks * HIST_BINS; }
cudaError_t launch_span_hist(const uint32_t *gmm, uint64_t ngran_cap, uint64_t nrows, int ntrks, uint32_t *d_hist, cudaStream_t s) {
   const uint64_t ngroups = nrows / (4 * RT_GRAN);
   if (!ngroups) return cudaMemsetAsync(d_hist, 0, (size_t)span_hist_words(ntrks) * 4, s);
   const uint32_t nsamples = (uint32_t)(ngroups < 65536 ? ngroups : 65536);
   cudaError_t e = cudaMemsetAsync(d_hist, 0, (size_t)span_hist_words(ntrks) * 4, s);
   if (e != cudaSuccess) return e;
   k_span_hist<<<dim3((unsigned)ntrks, 16), 256, 0, s>>>(gmm, ngran_cap, ngroups, nsamples, d_hist);
   return cudaGetLastError(); }

/* T0 of every track from the host copy of the histograms: min(0.75 * default-state bound, 0.5 * bound for the smallest signal height
   to expect), at least 5 % of the default-state bound.  Returns false if the two-pass scan does not apply. */
bool peak_mask_auto_T0(DevCfg &c, const uint32_t *hist) {
   const int Tdef = peak_mask_T0(c, 1.0f);
   if (Tdef <= 0) return false;
   for (int k = 0; k < c.ntrks; ++k) {
      const uint32_t *h = hist + (size_t)k * HIST_BINS;
      unsigned long long total = 0, acc = 0;
      for (int b = 0; b < HIST_BINS; ++b) total += h[b];
      int T0 = (int)(0.75f * (float)Tdef), T1 = 0;
      if (total >= 64) {
         int bA = 0;                                            /* full-scale height: 99.8th percentile of all stretches */
         for (; bA < HIST_BINS; ++bA) { acc += h[bA]; if (acc * 1000 >= total * 998) break; }
         const int b0 = bA / 4 > 1 ? bA / 4 : 1;               /* signal: at least a quarter of that */
         unsigned long long sig = 0;
         for (int b = b0; b < HIST_BINS; ++b) sig += h[b];
         if (sig >= 16) {
            acc = 0; int b = b0;
            for (; b < HIST_BINS; ++b) { acc += h[b]; if (acc * 10 >= sig) break; }
            const float alow = (float)(b << 6);
            const int Test = (int)(0.5f * (c.p.pkww_rise * alow / RT_PKWW_PEAKHEIGHT * 0.999f - 2.0f));
            if (Test < T0) T0 = Test;
            for (; b < HIST_BINS; ++b) { if (acc * 2 >= sig) break; acc += h[b + 1 < HIST_BINS ? b + 1 : b]; }   /* median signal height */
            T1 = (int)(0.9f * (c.p.pkww_rise * (float)(b << 6) / RT_PKWW_PEAKHEIGHT * 0.999f - 2.0f)); } }   /* below it the scan simply follows the T0 plane */
      const int floor_ = Tdef / 20 > 16 ? Tdef / 20 : 16;
      c.T0[k] = T0 < floor_ ? floor_ : (T0 > 65535 ? 65535 : T0);
      c.T1[k] = T1 > c.T0[k] + c.T0[k] / 8 ? (T1 > 65535 ? 65535 : T1) : 0; }
   return true; }
/* phase B only: the masks of rows [0, max row_end) must have been built (launch_peak_masks) on the same stream */
cudaError_t launch_units_sparse(const DevCfg &c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta, rt_event *pool,
                                uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks, int quiet_thr_lsb,
                                unsigned long long *counters, int sms, int max_ctas_per_sm, cudaStream_t s) {
   int per_sm = 0;                                                  /* per launch: the occupancy belongs to the device of the launch */
   cudaError_t e;
   const bool rec = c.recs != nullptr;
   e = rec ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_units_sparse<4, true>, SPARSE_THREADS, 0)
           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_units_sparse<4, false>, SPARSE_THREADS, 0);
   if (e != cudaSuccess) return e;
   if (per_sm < 1) per_sm = 1;
   if (max_ctas_per_sm > 0 && per_sm > max_ctas_per_sm) per_sm = max_ctas_per_sm;
   const uint64_t jobs = (uint64_t)nunits * (uint64_t)c.ntrks;
   uint64_t grid = (jobs + SPARSE_THREADS - 1) / SPARSE_THREADS;
   if (grid > (uint64_t)sms * (uint64_t)per_sm) grid = (uint64_t)sms * (uint64_t)per_sm;
   if (grid < 1) grid = 1;
   e = cudaMemsetAsync(counters + 2, 0, sizeof(unsigned long long), s);
   if (e != cudaSuccess) return e;
   if (rec) k_units_sparse<4, true><<<(unsigned)grid, SPARSE_THREADS, 0, s>>>(c, units, nunits, meta, pool, chunk_next, cursor, cap_chunks, quiet_thr_lsb, counters);
   else k_units_sparse<4, false><<<(unsigned)grid, SPARSE_THREADS, 0, s>>>(c, units, nunits, meta, pool, chunk_next, cursor, cap_chunks, quiet_thr_lsb, counters);
   return cudaGetLastError(); }
