/* readtape_b200/csrc/k_fast.cu -- K3b: the speculative whole-tape scan, int16 fast path (scan_fast.cuh).
 *
 * One thread per (unit, track), grid-stride over the unit table; every thread owns a lane-private
 * scratch area in shared memory laid out [entry][thread] so that every access of a warp hits 32
 * different banks whatever the lanes' ring positions are.  The 32 lanes of a warp are kept converged by
 * warp votes (rtfast::drive): search together, handle candidate rows together, change blocks together.
 * Same outputs as k_units_scan (k_scan.cu): events into the chunk pool, TrkMeta with the
 * unit-equivalence proof data.
 */
#include "scan_fast.cuh"
#include "scan_zc.cuh"
#include "kernels.h"
#include "emit.cuh"

#define FAST_THREADS 128

using namespace rtfast;

struct DevJobs {
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
      atomicAdd(&counters[0], (unsigned long long)us.end);
      if (us.em.n) atomicAdd(&counters[1], (unsigned long long)us.em.n); } };

struct WarpCount { __device__ int operator()(bool p) const { return __popc(__ballot_sync(0xffffffffu, p)); } };

__global__ void __launch_bounds__(FAST_THREADS, 4)
k_units_fast(DevCfg c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta,
             rt_event *pool, uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks,
             int quiet_thr_lsb, unsigned long long *rows_scanned, uint32_t ring) {
   extern __shared__ __align__(16) uint32_t fast_smem[];
   LaneMem<FAST_THREADS> mem = lane_mem<FAST_THREADS>(fast_smem + threadIdx.x, c.width);      /* [entry][thread] layout */
   const uint64_t total = (uint64_t)nunits * (uint64_t)c.ntrks;
   DevJobs jobs{c, units, meta, pool, chunk_next, cursor, cap_chunks, quiet_thr_lsb, rows_scanned, total, 0, false};
   UnitScan<FAST_THREADS, PoolEmit> us(c, mem);
   drive(us, jobs, WarpCount()); }

/* the zero-crossing fast path for GCR (scan_zc.cuh): same driver, same job groups */
__global__ void __launch_bounds__(FAST_THREADS, 4)
k_units_zc(DevCfg c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta,
           rt_event *pool, uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks, unsigned long long *rows_scanned) {
   extern __shared__ __align__(16) uint32_t fast_smem[];
   ZcMem<FAST_THREADS> mem = zc_mem<FAST_THREADS>(fast_smem + threadIdx.x);
   const uint64_t total = (uint64_t)nunits * (uint64_t)c.ntrks;
   DevJobs jobs{c, units, meta, pool, chunk_next, cursor, cap_chunks, 0, rows_scanned, total, 0, false};
   ZcScan<FAST_THREADS, PoolEmit> us(c, mem);
   drive(us, jobs, WarpCount()); }

bool fast_scan_eligible(const DevCfg &c) {
   if (zc_scan_eligible(c)) return true;
   return c.det == RT_DET_PEAK && (c.mode == RT_MODE_NRZI || c.mode == RT_MODE_PE || c.density) && !c.invert && !c.differentiate
          && c.width >= 3 && c.width <= RT_PKWW_MAX_WIDTH; }

cudaError_t launch_units_fast(const DevCfg &c, const UnitDesc *units, uint32_t nunits, TrkMeta *meta, rt_event *pool,
                              uint32_t *chunk_next, unsigned int *cursor, uint32_t cap_chunks, int quiet_thr_lsb,
                              unsigned long long *rows_scanned, int sms, int max_ctas_per_sm, cudaStream_t s) {
   if (zc_scan_eligible(c)) {
      const size_t zsmem = (size_t)zc_scratch_words(c) * FAST_THREADS * sizeof(uint32_t);
      int zcfg_per_sm = 0;                                            /* per launch: attributes are per device, not per process */
      cudaError_t ze;
      ze = cudaFuncSetAttribute(k_units_zc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem); if (ze != cudaSuccess) return ze;
      ze = cudaFuncSetAttribute(k_units_zc, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); if (ze != cudaSuccess) return ze;
      ze = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&zcfg_per_sm, k_units_zc, FAST_THREADS, zsmem); if (ze != cudaSuccess) return ze;
      int zper = zcfg_per_sm < 1 ? 1 : zcfg_per_sm;
      if (max_ctas_per_sm > 0 && zper > max_ctas_per_sm) zper = max_ctas_per_sm;
      uint64_t zgrid = ((uint64_t)nunits * (uint64_t)c.ntrks + FAST_THREADS - 1) / FAST_THREADS;
      if (zgrid > (uint64_t)sms * (uint64_t)zper) zgrid = (uint64_t)sms * (uint64_t)zper;
      if (zgrid < 1) zgrid = 1;
      ze = cudaMemsetAsync(rows_scanned + 2, 0, sizeof(unsigned long long), s); if (ze != cudaSuccess) return ze;
      k_units_zc<<<(unsigned)zgrid, FAST_THREADS, zsmem, s>>>(c, units, nunits, meta, pool, chunk_next, cursor, cap_chunks, rows_scanned);
      return cudaGetLastError(); }
   const uint32_t ring = ring_size(c.width);
   const size_t smem = (size_t)scratch_words(c.width) * FAST_THREADS * sizeof(uint32_t);
   int cfg_per_sm = 0;                                              /* per launch: attributes are per device, not per process */
   cudaError_t e;
   e = cudaFuncSetAttribute(k_units_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
   if (e != cudaSuccess) return e;
   e = cudaFuncSetAttribute(k_units_fast, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
   if (e != cudaSuccess) return e;
   e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg_per_sm, k_units_fast, FAST_THREADS, smem);
   if (e != cudaSuccess) return e;
   int per_sm = cfg_per_sm < 1 ? 1 : cfg_per_sm;
   if (max_ctas_per_sm > 0 && per_sm > max_ctas_per_sm) per_sm = max_ctas_per_sm;       /* leave room for concurrent ingest kernels */
   const uint64_t threads = (uint64_t)nunits * (uint64_t)c.ntrks;
   uint64_t grid = (threads + FAST_THREADS - 1) / FAST_THREADS;
   const uint64_t resident = (uint64_t)sms * (uint64_t)per_sm;
   if (grid > resident) grid = resident;
   if (grid < 1) grid = 1;
   e = cudaMemsetAsync(rows_scanned + 2, 0, sizeof(unsigned long long), s);       /* the job-group counter of this launch */
   if (e != cudaSuccess) return e;
   k_units_fast<<<(unsigned)grid, FAST_THREADS, smem, s>>>(c, units, nunits, meta, pool, chunk_next, cursor, cap_chunks, quiet_thr_lsb, rows_scanned, ring);
   return cudaGetLastError(); }
