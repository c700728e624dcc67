/* readtape_b200/csrc/k_units.cu -- K1b: cut the tape into independent scan units.
 *
 * In the reference every block decode starts from init_trackstate() (src/decoder.c:425) at a
 * row decided by the host's end-of-block logic, always inside an inter-block gap.  A fresh
 * detector started anywhere in a quiet gap converges to the same state ("unit
 * equivalence"), so the tape can be cut at all-track quiet gaps and every piece scanned
 * independently.  This file only PROPOSES the cuts from the 32-row granule min/max map of
 * k_ingest.cu; the scan kernel records the exact per-row data that rt_bulk_lookup() needs to
 * prove a proposal right for the reference's real reset row, so the heuristics here affect
 * speed, never results.
 *
 *   k_quiet_bitmap   1 bit per granule: every track quiet in that granule
 *   k_gap_flags      bit set for the first granule of each run of >= min_gap quiet granules
 *                    (and granule 0: the tape start is always a unit start), + per-block counts
 *   k_scan_counts    exclusive prefix over the per-block counts (one CTA)
 *   k_write_units    ordered compaction of the flags into UnitDesc.row0
 *   k_close_units    row_end of every unit
 */
#include <cuda_runtime.h>
#include "kernels.h"

#define UB_THREADS 256                      /* one bitmap word (32 granules) per thread */

size_t units_bitmap_words(uint64_t nrows) { return (size_t)(((nrows + RT_GRAN - 1) / RT_GRAN + 31) / 32) + 4; }
size_t units_blocks(uint64_t nrows) { return (units_bitmap_words(nrows) + UB_THREADS - 1) / UB_THREADS; }

__global__ void __launch_bounds__(256)
k_quiet_bitmap(const uint32_t *gmm, uint64_t ngran_cap, int ntrks, uint64_t ngran, UnitParams up, uint32_t *bitmap) {
   const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
   bool quiet = false;
   if (g < ngran) {
      quiet = true;
      for (int k = 0; k < ntrks; ++k) {
         uint32_t mm = gmm[(size_t)k * ngran_cap + g];
         int mn = (int)(int16_t)(mm & 0xffff), mx = (int)(int16_t)(mm >> 16);
         if (up.det == RT_DET_ZC) { if (mx > up.thr || mn < -up.thr) quiet = false; }
         else if (mx - mn > up.thr) quiet = false; } }
   uint32_t word = __ballot_sync(0xffffffffu, quiet);
   if ((threadIdx.x & 31) == 0 && (g >> 5) < (ngran + 31) / 32) bitmap[g >> 5] = word; }     /* the grid is rounded up to whole CTAs */

__device__ __forceinline__ bool quiet_at(const uint32_t *bitmap, uint64_t g) { return (bitmap[g >> 5] >> (g & 31)) & 1u; }

__global__ void __launch_bounds__(UB_THREADS)
k_gap_flags(const uint32_t *bitmap, uint64_t ngran, uint32_t min_gap, uint32_t *flags, uint32_t *blockcount) {
   const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   /* bitmap word */
   const uint64_t nwords = (ngran + 31) / 32;
   uint32_t out = 0;
   if (w < nwords) {
      uint32_t cur = bitmap[w];
      uint32_t prevbit = w ? (bitmap[w - 1] >> 31) : 0u;
      uint32_t starts = cur & ~((cur << 1) | prevbit);                    /* quiet and the granule before is not */
      while (starts) {
         int b = __ffs(starts) - 1; starts &= starts - 1;
         uint64_t g = w * 32 + b;
         bool ok = g + min_gap <= ngran;
         for (uint32_t j = 1; ok && j < min_gap; ++j) ok = quiet_at(bitmap, g + j);
         if (ok) out |= 1u << b; }
      if (w == 0) out |= 1u;                                                /* the tape start */
      flags[w] = out; }
   __shared__ uint32_t red[UB_THREADS / 32];
   uint32_t c = __popc(out);
   for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
   if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
   __syncthreads();
   if (threadIdx.x == 0) { uint32_t s = 0; for (int i = 0; i < UB_THREADS / 32; ++i) s += red[i]; blockcount[blockIdx.x] = s; } }

__global__ void k_scan_counts(uint32_t *blockcount, uint32_t nblocks, uint32_t *nunits) {
   /* single thread: nblocks is a few thousand at most (one per 8192 granules = 262144 rows) */
   if (threadIdx.x == 0 && blockIdx.x == 0) {
      uint32_t run = 0;
      for (uint32_t i = 0; i < nblocks; ++i) { uint32_t c = blockcount[i]; blockcount[i] = run; run += c; }
      *nunits = run; } }

__global__ void __launch_bounds__(UB_THREADS)
k_write_units(const uint32_t *flags, uint64_t ngran, const uint32_t *blockoffset, UnitDesc *units, uint32_t cap) {
   const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const uint64_t nwords = (ngran + 31) / 32;
   uint32_t f = w < nwords ? flags[w] : 0u;
   /* exclusive prefix of popc(f) inside the block */
   __shared__ uint32_t wsum[UB_THREADS / 32];
   uint32_t c = __popc(f), incl = c;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += n; }
   if (lane == 31) wsum[wid] = incl;
   __syncthreads();
   uint32_t base = blockoffset[blockIdx.x];
   for (int i = 0; i < wid; ++i) base += wsum[i];
   uint32_t pos = base + incl - c;
   while (f) {
      int b = __ffs(f) - 1; f &= f - 1;
      if (pos < cap) { units[pos].row0 = (w * 32 + b) * RT_GRAN; units[pos].row_end = 0; }
      ++pos; } }

__global__ void k_close_units(UnitDesc *units, const uint32_t *nunits_p, uint32_t cap, uint64_t nrows, uint64_t tail_rows) {
   const uint32_t n = min(*nunits_p, cap);
   for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) {
      uint64_t end = nrows;
      if (u + 1 < n) { end = units[u + 1].row0 + tail_rows; if (end > nrows) end = nrows; }
      units[u].row_end = end; } }

cudaError_t launch_find_units(const int16_t *gmm, uint64_t ngran_cap, int ntrks, uint64_t nrows, const UnitParams &up,
                              uint32_t *d_bitmap, uint32_t *d_flags, uint32_t *d_blockcount, UnitDesc *d_units,
                              uint32_t units_cap, uint32_t *d_nunits, cudaStream_t st, int *launches) {
   /* only whole granules vote; a trailing partial granule is scanned as part of the last unit */
   const uint64_t ngran = nrows / RT_GRAN;
   const uint64_t nwords = (ngran + 31) / 32;
   const uint32_t nblocks = (uint32_t)((nwords + UB_THREADS - 1) / UB_THREADS);
   if (ngran == 0 || nblocks == 0) {
      UnitDesc u{0, nrows}; uint32_t one = nrows ? 1u : 0u;
      cudaMemcpyAsync(d_units, &u, sizeof u, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(d_nunits, &one, sizeof one, cudaMemcpyHostToDevice, st);
      return cudaStreamSynchronize(st); }
   const uint64_t gthreads = nwords * 32;
   k_quiet_bitmap<<<(unsigned)((gthreads + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint32_t *>(gmm), ngran_cap, ntrks, ngran, up, d_bitmap);
   k_gap_flags<<<nblocks, UB_THREADS, 0, st>>>(d_bitmap, ngran, up.min_gap_gran, d_flags, d_blockcount);
   k_scan_counts<<<1, 32, 0, st>>>(d_blockcount, nblocks, d_nunits);
   k_write_units<<<nblocks, UB_THREADS, 0, st>>>(d_flags, ngran, d_blockcount, d_units, units_cap);
   k_close_units<<<64, 256, 0, st>>>(d_units, d_nunits, units_cap, nrows, up.tail_rows);
   *launches += 5;
   return cudaGetLastError(); }


/* ---- -invert (readtape.c:1421): a negated copy of the planes and of the granule map ------------------------------------------------
 * The reference negates every sample as it is read.  volts(-x) == -volts(x) exactly (rounding is symmetric, -32768 never occurs:
 * it is the end marker), so a scan of the negated planes with the invert flag cleared IS the inverted scan, and every fast kernel
 * serves it unchanged.  One thread per 8 samples; granule (min, max) becomes (-max, -min). */
__global__ void __launch_bounds__(256)
k_negate_planes(const int16_t *src, int16_t *dst, uint64_t plane_stride, uint64_t row_lo, uint64_t row_hi, int ntrks) {
   const uint64_t per = (row_hi - row_lo + 7) / 8;                 /* row_lo is a multiple of 8 */
   const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= per * (uint64_t)ntrks) return;
   const uint64_t k = i / per, r = row_lo + (i - k * per) * 8;
   const uint4 v = *reinterpret_cast<const uint4 *>(src + k * plane_stride + r);     /* the planes are padded beyond the last row */
   uint4 o; o.x = __vneg2(v.x); o.y = __vneg2(v.y); o.z = __vneg2(v.z); o.w = __vneg2(v.w);
   *reinterpret_cast<uint4 *>(dst + k * plane_stride + r) = o; }

__global__ void __launch_bounds__(256)
k_negate_gmm(const uint32_t *src, uint32_t *dst, uint64_t ngran_cap, uint64_t g_lo, uint64_t g_hi, int ntrks) {
   const uint64_t per = g_hi - g_lo;
   const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= per * (uint64_t)ntrks) return;
   const uint64_t k = i / per, g = g_lo + (i - k * per);
   const uint32_t w = src[k * ngran_cap + g];
   const int mn = (int)(int16_t)(uint16_t)(w & 0xffffu), mx = (int)(int16_t)(uint16_t)(w >> 16);
   dst[k * ngran_cap + g] = ((uint32_t)(-mx) & 0xffffu) | ((uint32_t)(-mn) << 16); }

cudaError_t launch_negate(const int16_t *planes, int16_t *planes_inv, uint64_t plane_stride, const int16_t *gmm, int16_t *gmm_inv, uint64_t ngran_cap,
                          uint64_t row_lo, uint64_t row_hi, int ntrks, cudaStream_t s) {
   row_lo = row_lo / RT_GRAN * RT_GRAN;                              /* whole granules (and 16-byte aligned vectors) */
   if (row_hi <= row_lo) return cudaSuccess;
   const uint64_t n = (row_hi - row_lo + 7) / 8 * (uint64_t)ntrks;
   k_negate_planes<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(planes, planes_inv, plane_stride, row_lo, row_hi, ntrks);
   const uint64_t g_lo = row_lo / RT_GRAN, g_hi = (row_hi + RT_GRAN - 1) / RT_GRAN;
   const uint64_t m = (g_hi - g_lo) * (uint64_t)ntrks;
   k_negate_gmm<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint32_t *>(gmm), reinterpret_cast<uint32_t *>(gmm_inv), ngran_cap, g_lo, g_hi, ntrks);
   return cudaGetLastError(); }
