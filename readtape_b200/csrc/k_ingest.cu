/* readtape_b200/csrc/k_ingest.cu -- K1: TBIN rows -> track-major planes + quiet map.
 *
 * Replaces the per-row fread()+convert loop of readblock() (reference src/readtape.c:1405-1425)
 * for the whole capture at once.  The TBIN payload is row-interleaved: nheads little-endian
 * int16 per row (18-byte pitch for 9 tracks, csvtbin.h:98-105).  Every later kernel walks ONE
 * track sequentially, so the first thing done with the bytes is a transpose into track-major
 * int16 planes (head->track permutation applied, unused Whirlwind heads dropped), together
 * with the min/max of every 32-row granule per track (the map the unit finder reads) and the
 * position of the end-of-data marker (-32768 in head 0, readtape.c:1410).
 *
 * This is the bandwidth-bound kernel of the pipeline: 2 B read + 2 B written per track-sample.
 * Layout of the work (sm_100a):
 *   - persistent CTAs (grid = #SMs x CTAs/SM), each loops over row tiles of TILE_ROWS rows;
 *   - tiles are staged into shared memory by the TMA engine with 1-D bulk async copies
 *     (cp.async.bulk ... mbarrier::complete_tx), NSTAGE deep, full/empty mbarrier pairs;
 *   - each thread owns 8 consecutive rows of the tile = NH x 16 bytes of contiguous shared
 *     memory, read with NH conflict-free LDS.128, de-interleaved in registers (static
 *     indexing, NH is a template parameter) and written as one coalesced STG.128 per track;
 *   - granule min/max: per-thread partials to shared memory, reduced 4:1 after a CTA barrier.
 * A plain-load kernel (k_ingest_simple) covers odd head counts, unaligned sources and tails.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "rt_dev.h"
#include "kernels.h"
#include "scan_masks.cuh"

#define ING_THREADS   256
#define ING_TROWS     (ING_THREADS * 8)      /* 2048 rows per tile */
#define ING_NSTAGE    4

struct IngestArgs {
   const int16_t *src;        /* interleaved rows of this upload chunk (device) */
   uint64_t nrows;            /* rows in this chunk */
   uint64_t row_base;         /* tape row of src row 0 (multiple of RT_GRAN unless it is the final append) */
   int16_t *planes; uint64_t plane_stride;
   int16_t *gmm;              /* [ntrks][ngran_cap][2] */
   uint64_t ngran_cap;
   unsigned long long *first_end_row;
   int32_t trk_of_head[RT_MAXTRKS];   /* -1: dropped */
};

#include "tma_1d.cuh"

/* ---- per-thread de-interleave of 8 rows -------------------------------------------------------- */
/* fused mask pass: the de-interleaved tile also goes to shared memory, one 144-byte slot per 64-row run and track (128 bytes of
   samples + 16 bytes of padding: the mask pass reads 16-byte chunks with one run per lane, and a 144-byte lane stride spreads a
   quarter-warp over all 32 banks); slot 0 holds the rows in front of the tile (the halo) */
#define PT_SLOT   144
#define PT_TRK    (33 * PT_SLOT)
__device__ __forceinline__ uint32_t pt_off(int k, int r /* row relative to the tile start, >= -64 */) {
   return (uint32_t)(k * PT_TRK + ((r + 64) >> 6) * PT_SLOT + ((r + 64) & 63) * 2); }

template <int NH>
__device__ __forceinline__ void deinterleave8(const uint32_t (&w)[NH * 4], const IngestArgs &a, uint64_t row0 /* tape row */,
                                              uint32_t *part /* smem [NH][ING_THREADS] packed min|max<<16 or null */, int tid,
                                              unsigned char *ptile = nullptr) {
   /* w holds 8 rows x NH halfwords, row-major; element (r,h) is halfword r*NH+h */
#pragma unroll
   for (int h = 0; h < NH; ++h) {
      int k = a.trk_of_head[h];
      uint32_t o[4];
      int mn = 32767, mx = -32768;
#pragma unroll
      for (int r = 0; r < 8; r += 2) {
         const int e0 = r * NH + h, e1 = (r + 1) * NH + h;
         uint32_t lo = (e0 & 1) ? (w[e0 >> 1] >> 16) : (w[e0 >> 1] & 0xffffu);
         uint32_t hi = (e1 & 1) ? (w[e1 >> 1] >> 16) : (w[e1 >> 1] & 0xffffu);
         o[r >> 1] = lo | (hi << 16);
         int s0 = (int)(int16_t)lo, s1 = (int)(int16_t)hi;
         mn = min(mn, min(s0, s1)); mx = max(mx, max(s0, s1)); }
      if (h == 0) {                        /* end-of-data marker, readtape.c:1410 */
#pragma unroll
         for (int r = 0; r < 8; ++r) {
            const int e = r * NH;
            uint32_t v = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu);
            if (v == 0x8000u) { atomicMin(a.first_end_row, (unsigned long long)(row0 + r)); break; } } }
      if (k >= 0) {
         uint4 q = make_uint4(o[0], o[1], o[2], o[3]);
         *reinterpret_cast<uint4 *>(a.planes + (size_t)k * a.plane_stride + row0) = q;
         if (ptile) *reinterpret_cast<uint4 *>(ptile + pt_off(k, 8 * tid)) = q;
         if (part) part[h * ING_THREADS + tid] = ((uint32_t)mn & 0xffffu) | ((uint32_t)mx << 16); } } }

template <int NH>
__global__ void __launch_bounds__(ING_THREADS)
k_ingest_tma(IngestArgs a, uint64_t ntiles) {
   extern __shared__ __align__(128) unsigned char smem_raw[];
   constexpr uint32_t TILE_BYTES = ING_TROWS * NH * 2;
   unsigned char *stage = smem_raw;                                         /* NSTAGE * TILE_BYTES */
   uint32_t *part = reinterpret_cast<uint32_t *>(smem_raw + ING_NSTAGE * TILE_BYTES);  /* [NH][ING_THREADS] */
   uint64_t *full = reinterpret_cast<uint64_t *>(part + NH * ING_THREADS);
   uint64_t *empty = full + ING_NSTAGE;
   const int tid = threadIdx.x;
   if (tid == 0) {
      for (int s = 0; s < ING_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ING_THREADS / 32); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
   __syncthreads();

   const uint64_t first = blockIdx.x, step = gridDim.x;
   const uint64_t mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;   /* tiles this CTA processes */
   /* prologue: fill the pipeline */
   if (tid == 0) {
      for (uint64_t i = 0; i < mine && i < ING_NSTAGE; ++i) {
         mbar_expect_tx(&full[i], TILE_BYTES);
         tma_load_1d(stage + i * TILE_BYTES, reinterpret_cast<const unsigned char *>(a.src) + (first + i * step) * TILE_BYTES,
                     TILE_BYTES, &full[i]); } }
   for (uint64_t i = 0; i < mine; ++i) {
      const int s = (int)(i % ING_NSTAGE);
      const uint32_t par = (uint32_t)((i / ING_NSTAGE) & 1);
      mbar_wait(&full[s], par);
      /* 8 consecutive rows = NH*16 contiguous bytes; lane stride NH*16 B: conflict-free LDS.128 for odd NH */
      const uint4 *p = reinterpret_cast<const uint4 *>(stage + s * TILE_BYTES + (size_t)tid * NH * 16);
      uint32_t w[NH * 4];
#pragma unroll
      for (int j = 0; j < NH; ++j) { uint4 q = p[j]; w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w; }
      /* this thread's copy of the stage is in registers: release the slot to the producer early */
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);
      if (tid == 0 && i + ING_NSTAGE < mine) {
         mbar_wait(&empty[s], par);                                       /* all warps have drained stage s */
         mbar_expect_tx(&full[s], TILE_BYTES);
         tma_load_1d(stage + s * TILE_BYTES,
                     reinterpret_cast<const unsigned char *>(a.src) + (first + (i + ING_NSTAGE) * step) * TILE_BYTES,
                     TILE_BYTES, &full[s]); }
      const uint64_t tile = first + i * step;
      const uint64_t row0 = a.row_base + tile * ING_TROWS + (uint64_t)tid * 8;
      deinterleave8<NH>(w, a, row0, part, tid);
      __syncthreads();
      /* granule = 32 rows = 4 threads: reduce the partials, one (head, granule) per thread */
      for (int o = tid; o < NH * (ING_THREADS / 4); o += ING_THREADS) {
         const int h = o / (ING_THREADS / 4), g = o % (ING_THREADS / 4);
         const int k = a.trk_of_head[h];
         if (k < 0) continue;
         const uint4 q = *reinterpret_cast<const uint4 *>(&part[h * ING_THREADS + g * 4]);
         int mn = min(min((int)(int16_t)(q.x & 0xffff), (int)(int16_t)(q.y & 0xffff)),
                      min((int)(int16_t)(q.z & 0xffff), (int)(int16_t)(q.w & 0xffff)));
         int mx = max(max((int)(int16_t)(q.x >> 16), (int)(int16_t)(q.y >> 16)),
                      max((int)(int16_t)(q.z >> 16), (int)(int16_t)(q.w >> 16)));
         const uint64_t gran = (a.row_base + tile * ING_TROWS) / RT_GRAN + (uint64_t)g;
         reinterpret_cast<uint32_t *>(a.gmm)[(size_t)k * a.ngran_cap + gran] = ((uint32_t)mn & 0xffffu) | ((uint32_t)mx << 16); }
      __syncthreads(); } }

/* ---- K1 + K3c phase A fused: the tile that the TMA engine has just brought on chip is de-interleaved AND turned into the candidate /
 * canonical bit planes (scan_masks.cuh) before it leaves -- the separate mask pass re-read the 20 GB of planes this kernel has just
 * written.  One more warp than the plain kernel: 9 warps = 9 tracks x 32 runs of 64 rows; warp k computes the masks of track k from
 * the shared-memory copy of the tile (the rows in front of the tile come from the source rows / the planes of the previous chunk).
 * Only for nheads == ntrks == NH; the window width is a template parameter like in k_peak_masks. */
struct MaskArgs { uint32_t *cand, *cand2, *acan; uint64_t mask_stride; int32_t T0[RT_MAXTRKS], T1[RT_MAXTRKS]; };
#define INGM_THREADS (ING_THREADS + 32)

template <int NH, int W>
__global__ void __launch_bounds__(INGM_THREADS)
k_ingest_masks_tma(IngestArgs a, const __grid_constant__ MaskArgs m, uint64_t ntiles) {
   static_assert(NH == 9, "one warp per track: 9 warps");
   using RM = rtmask::RunMasks<W>;
   extern __shared__ __align__(128) unsigned char smem_raw[];
   constexpr uint32_t TILE_BYTES = ING_TROWS * NH * 2;
   unsigned char *stage = smem_raw;
   uint32_t *part = reinterpret_cast<uint32_t *>(smem_raw + ING_NSTAGE * TILE_BYTES);
   unsigned char *ptile = reinterpret_cast<unsigned char *>(part + NH * ING_THREADS);
   uint64_t *full = reinterpret_cast<uint64_t *>(ptile + NH * PT_TRK);
   uint64_t *empty = full + ING_NSTAGE;
   const int tid = threadIdx.x;
   if (tid == 0) {
      for (int s = 0; s < ING_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ING_THREADS / 32); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
   __syncthreads();
   const uint64_t first = blockIdx.x, step = gridDim.x;
   const uint64_t mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;
   if (tid == 0) {
      for (uint64_t i = 0; i < mine && i < ING_NSTAGE; ++i) {
         mbar_expect_tx(&full[i], TILE_BYTES);
         tma_load_1d(stage + i * TILE_BYTES, reinterpret_cast<const unsigned char *>(a.src) + (first + i * step) * TILE_BYTES, TILE_BYTES, &full[i]); } }
   for (uint64_t i = 0; i < mine; ++i) {
      const int s = (int)(i % ING_NSTAGE);
      const uint32_t par = (uint32_t)((i / ING_NSTAGE) & 1);
      const uint64_t tile = first + i * step;
      const uint64_t trow0 = a.row_base + tile * ING_TROWS;          /* tape row of the tile's first row */
      if (tid < ING_THREADS) {
         mbar_wait(&full[s], par);
         const uint4 *p = reinterpret_cast<const uint4 *>(stage + s * TILE_BYTES + (size_t)tid * NH * 16);
         uint32_t w[NH * 4];
#pragma unroll
         for (int j = 0; j < NH; ++j) { uint4 q = p[j]; w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w; }
         __syncwarp();
         if ((tid & 31) == 0) mbar_arrive(&empty[s]);
         if (tid == 0 && i + ING_NSTAGE < mine) {
            mbar_wait(&empty[s], par);
            mbar_expect_tx(&full[s], TILE_BYTES);
            tma_load_1d(stage + s * TILE_BYTES, reinterpret_cast<const unsigned char *>(a.src) + (first + (i + ING_NSTAGE) * step) * TILE_BYTES, TILE_BYTES, &full[s]); }
         deinterleave8<NH>(w, a, trow0 + (uint64_t)tid * 8, part, tid, ptile); }
      /* the halo: the RM::HALO rows in front of the tile, per track -- from the source rows of this chunk, from the planes the
         previous chunk left, or zeros at the very start of the tape */
      for (int e = tid; e < RM::HALO * NH; e += INGM_THREADS) {
         const int h = e / RM::HALO, r = e % RM::HALO - RM::HALO;    /* r in [-HALO, -1] */
         const int k = a.trk_of_head[h];
         int16_t v = 0;
         if (tile > 0) v = a.src[(tile * ING_TROWS + (uint64_t)(int64_t)r) * NH + h];
         else if (trow0 >= (uint64_t)RM::HALO) v = a.planes[(size_t)k * a.plane_stride + trow0 + (uint64_t)(int64_t)r];
         *reinterpret_cast<int16_t *>(ptile + pt_off(k, r)) = v; }
      __syncthreads();
      for (int o = tid; o < NH * (ING_THREADS / 4); o += INGM_THREADS) {      /* granule min/max, as in the plain kernel */
         const int h = o / (ING_THREADS / 4), g = o % (ING_THREADS / 4);
         const int k = a.trk_of_head[h];
         const uint4 q = *reinterpret_cast<const uint4 *>(&part[h * ING_THREADS + g * 4]);
         int mn = min(min((int)(int16_t)(q.x & 0xffff), (int)(int16_t)(q.y & 0xffff)), min((int)(int16_t)(q.z & 0xffff), (int)(int16_t)(q.w & 0xffff)));
         int mx = max(max((int)(int16_t)(q.x >> 16), (int)(int16_t)(q.y >> 16)), max((int)(int16_t)(q.z >> 16), (int)(int16_t)(q.w >> 16)));
         reinterpret_cast<uint32_t *>(a.gmm)[(size_t)k * a.ngran_cap + trow0 / RT_GRAN + (uint64_t)g] = ((uint32_t)mn & 0xffffu) | ((uint32_t)mx << 16); }
      {  /* phase A of the two-pass scan for this tile: warp = track, lane = run of 64 rows */
         const int k = tid >> 5, run = tid & 31;
         uint32_t x[RM::NW];
#pragma unroll
         for (int cch = 0; cch < RM::NW / 4; ++cch) {
            const uint4 v = *reinterpret_cast<const uint4 *>(ptile + pt_off(k, run * 64 - RM::HALO + 8 * cch));
            x[4 * cch] = v.x ^ rtmask::BIAS2; x[4 * cch + 1] = v.y ^ rtmask::BIAS2; x[4 * cch + 2] = v.z ^ rtmask::BIAS2; x[4 * cch + 3] = v.w ^ rtmask::BIAS2; }
         uint32_t cw[2], dw[2], aw[2];
         const uint32_t T0 = m.T0[k] > 0 ? (uint32_t)m.T0[k] : 65535u, T1 = m.T1[k] > 0 ? (uint32_t)m.T1[k] : 65535u;
         RM::core(x, T0, T1, cw, dw, aw);
         const uint64_t p0 = trow0 + (uint64_t)run * 64;
         if (p0 < (uint64_t)W) {                                   /* rows whose window would reach in front of row 0: no bits */
            const uint32_t keep0 = W >= 32 ? 0u : (0xffffffffu << W), keep1 = W > 32 ? (0xffffffffu << (W - 32)) : 0xffffffffu;
            cw[0] &= keep0; dw[0] &= keep0; aw[0] &= keep0; cw[1] &= keep1; dw[1] &= keep1; aw[1] &= keep1; }
         const size_t wi = (size_t)k * m.mask_stride + (size_t)(p0 / 32);
         *reinterpret_cast<uint2 *>(m.cand + wi) = make_uint2(cw[0], cw[1]);
         *reinterpret_cast<uint2 *>(m.cand2 + wi) = make_uint2(dw[0], dw[1]);
         *reinterpret_cast<uint2 *>(m.acan + wi) = make_uint2(aw[0], aw[1]); }
      __syncthreads(); } }

/* ---- plain-load kernel: any head count, any alignment, partial granules -------------------------- */
__global__ void __launch_bounds__(256)
k_ingest_simple(IngestArgs a, uint64_t row_from, uint64_t row_to, int nheads) {
   /* one warp per (granule, all heads): lane = row within the 32-row granule */
   const int lane = threadIdx.x & 31;
   const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
   const uint64_t g_from = (a.row_base + row_from) / RT_GRAN, g_to = (a.row_base + row_to + RT_GRAN - 1) / RT_GRAN;
   for (uint64_t g = g_from + warp; g < g_to; g += nwarps) {
      const uint64_t trow = g * RT_GRAN + lane;                 /* tape row */
      const bool ok = trow >= a.row_base + row_from && trow < a.row_base + row_to;
      const int16_t *r = a.src + (trow - a.row_base) * (uint64_t)nheads;
      for (int h = 0; h < nheads; ++h) {
         const int k = a.trk_of_head[h];
         int v = ok ? (int)r[h] : 0;
         if (h == 0 && ok && v == -32768) atomicMin(a.first_end_row, (unsigned long long)trow);
         if (k < 0) continue;
         if (ok) a.planes[(size_t)k * a.plane_stride + trow] = (int16_t)v;
         int mn = ok ? v : 32767, mx = ok ? v : -32768;
#pragma unroll
         for (int d = 16; d > 0; d >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d)); }
         if (lane == 0) {
            uint32_t *slot = &reinterpret_cast<uint32_t *>(a.gmm)[(size_t)k * a.ngran_cap + g];
            if (trow < a.row_base + row_from) {                   /* granule partly filled by an earlier append: merge */
               uint32_t old = *slot;
               mn = min(mn, (int)(int16_t)(old & 0xffff)); mx = max(mx, (int)(int16_t)(old >> 16)); }
            *slot = ((uint32_t)mn & 0xffffu) | ((uint32_t)mx << 16); } } } }

/* ---- launcher ------------------------------------------------------------------------------------ */
template <int NH>
static cudaError_t launch_tma(const IngestArgs &a, uint64_t ntiles, int sms, cudaStream_t st) {
   const size_t smem = (size_t)ING_NSTAGE * ING_TROWS * NH * 2 + (size_t)NH * ING_THREADS * 4 + 2 * ING_NSTAGE * 8;
   /* function attributes belong to the device / context of the launch: set on every launch (a host-side table look-up), never
      cached in a process-wide static -- a second device, or a second thread with another tape, must see them too */
   cudaError_t e = cudaFuncSetAttribute(k_ingest_tma<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
   if (e != cudaSuccess) return e;
   int grid = (int)(ntiles < (uint64_t)sms ? ntiles : (uint64_t)sms);
   k_ingest_tma<NH><<<grid, ING_THREADS, smem, st>>>(a, ntiles);
   return cudaGetLastError(); }

template <int W>
static cudaError_t launch_tma_masks(const IngestArgs &a, const MaskArgs &m, uint64_t ntiles, int sms, cudaStream_t st) {
   constexpr int NH = 9;
   const size_t smem = (size_t)ING_NSTAGE * ING_TROWS * NH * 2 + (size_t)NH * ING_THREADS * 4 + (size_t)NH * PT_TRK + 2 * ING_NSTAGE * 8;
   cudaError_t e = cudaFuncSetAttribute(k_ingest_masks_tma<NH, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
   if (e != cudaSuccess) return e;
   int grid = (int)(ntiles < (uint64_t)sms ? ntiles : (uint64_t)sms);
   k_ingest_masks_tma<NH, W><<<grid, INGM_THREADS, smem, st>>>(a, m, ntiles);
   return cudaGetLastError(); }

bool ingest_masks_supported(int nheads, int ntrks, int width) { return nheads == 9 && ntrks == 9 && width >= 6 && width <= 20; }

/* mask != null: the fused kernel also writes the candidate / canonical bit planes of the whole tiles it ingests (the caller runs the
   separate mask pass over what is left: the tail rows behind the last whole tile).  *masked_rows = rows (from row_base) whose masks
   were written. */
cudaError_t launch_ingest(const int16_t *src, uint64_t nrows, uint64_t row_base, int nheads, const int32_t *trk_of_head,
                          int16_t *planes, uint64_t plane_stride, int16_t *gmm, uint64_t ngran_cap,
                          unsigned long long *first_end_row, int sms, int force_simple, cudaStream_t st, int *launches,
                          const IngestMasks *mask, uint64_t *masked_rows) {
   if (masked_rows) *masked_rows = 0;
   IngestArgs a;
   a.src = src; a.nrows = nrows; a.row_base = row_base; a.planes = planes; a.plane_stride = plane_stride;
   a.gmm = gmm; a.ngran_cap = ngran_cap; a.first_end_row = first_end_row;
   for (int h = 0; h < RT_MAXTRKS; ++h) a.trk_of_head[h] = h < nheads ? trk_of_head[h] : -1;
   uint64_t done = 0;
   const bool aligned = ((uintptr_t)src % 16 == 0) && (row_base % ING_TROWS == 0) && ((uintptr_t)planes % 16 == 0) && (plane_stride % 8 == 0);
   if (!force_simple && aligned && (nheads == 9 || nheads == 7 || nheads == 6)) {
      uint64_t ntiles = nrows / ING_TROWS;
      bool fused = false;
      if (ntiles && mask && ingest_masks_supported(nheads, mask->ntrks, mask->width)) {
         bool identity = true;                                     /* one warp per track: every head feeds a track */
         for (int h = 0; h < nheads; ++h) identity = identity && a.trk_of_head[h] >= 0;
         if (identity) {
            MaskArgs m; m.cand = mask->cand; m.cand2 = mask->cand2; m.acan = mask->acan; m.mask_stride = mask->mask_stride;
            for (int k = 0; k < RT_MAXTRKS; ++k) { m.T0[k] = mask->T0[k]; m.T1[k] = mask->T1[k]; }
            cudaError_t e = cudaErrorInvalidValue;
            switch (mask->width) {
#define MW(W) case W: e = launch_tma_masks<W>(a, m, ntiles, sms, st); break;
               MW(6) MW(7) MW(8) MW(9) MW(10) MW(11) MW(12) MW(13) MW(14) MW(15) MW(16) MW(17) MW(18) MW(19) MW(20)
#undef MW
            }
            if (e != cudaSuccess) return e;
            fused = true; ++*launches; done = ntiles * ING_TROWS;
            if (masked_rows) *masked_rows = done; } }
      if (ntiles && !fused) {
         cudaError_t e = nheads == 9 ? launch_tma<9>(a, ntiles, sms, st) : nheads == 7 ? launch_tma<7>(a, ntiles, sms, st)
                         : launch_tma<6>(a, ntiles, sms, st);
         if (e != cudaSuccess) return e;
         ++*launches;
         done = ntiles * ING_TROWS; } }
   if (done < nrows) {
      uint64_t grans = (nrows - done + RT_GRAN - 1) / RT_GRAN + 1;
      uint64_t blocks = (grans + 7) / 8;
      if (blocks > (uint64_t)sms * 8) blocks = (uint64_t)sms * 8;
      k_ingest_simple<<<(int)blocks, 256, 0, st>>>(a, done, nrows, nheads);
      ++*launches; }
   return cudaGetLastError(); }
