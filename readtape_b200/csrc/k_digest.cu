/* readtape_b200/csrc/k_digest.cu -- verification of a whole-tape scan AT SCALE, on the device (rt_bulk_tile_digest).
 *
 * A benchmark tape is periodic (a super-tile of `period` rows repeated), and every block decode starts from a fresh reset,
 * so the events of tile k are the events of tile 0 shifted by k * period rows.  This kernel folds every event of a finished
 * rt_bulk_scan() into one order-independent 64-bit digest and one count PER TILE, so that a 131-million-event result can be
 * compared with the CPU oracle's events of a single tile -- on every step of a benchmark if wanted -- for the price of
 * reading the event pool once.
 *
 * What enters an event's hash: row relative to its tile, track, polarity, v_top, v_bot, AGC gain (bit patterns) and the
 * event time as the number of HALF sample periods between the detection row and the event time.  The double-precision
 * event time itself cannot be shifted by k * period bit for bit (rounding depends on the magnitude of the time), so it is
 * checked separately: it must be reproduced exactly by the reference's own expression for the row and half-sample count
 * (refine_peak, decoder.c:732, or the row time of the arming row for the zero-crossing detector); events for which it is
 * not are counted in *bad_times.
 *
 * Units overlap (a unit keeps scanning tail_rows into its successor): an event is counted by the unit that owns its row,
 * i.e. the last unit whose first row is <= the event's row.
 */
#include <cuda_runtime.h>
#include "kernels.h"

__device__ __forceinline__ uint64_t mix64(uint64_t x) {                 /* splitmix64 finaliser */
   x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
   return x; }

__device__ __forceinline__ double dg_row_time(const DevCfg &c, uint64_t row) {
   long long ns = (long long)(c.tstart_ns + row * c.tdelta_ns);
   return (double)ns / 1e9; }

__global__ void __launch_bounds__(128)
k_tile_digest(DevCfg c, const UnitDesc *units, uint32_t nunits, const TrkMeta *meta, const rt_event *pool, const uint32_t *chunk_next,
              uint64_t period, uint64_t ntiles, unsigned long long *counts, unsigned long long *digests, unsigned long long *bad) {
   const uint64_t total = (uint64_t)nunits * (uint64_t)c.ntrks;
   for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (uint64_t)gridDim.x * blockDim.x) {
      const uint32_t u = (uint32_t)(f / c.ntrks);
      const TrkMeta m = meta[f];
      if (!m.nevents || m.first_chunk == RT_NOCHUNK) continue;
      const uint64_t own_end = u + 1 < nunits ? units[u + 1].row0 : ~0ull;
      uint64_t tile = ~0ull, acc = 0, n = 0, nbad = 0;
      uint32_t chunk = m.first_chunk, slot = 0;
      for (uint32_t i = 0; i < m.nevents; ++i) {
         const rt_event e = pool[(size_t)chunk * RT_EVC + slot];
         if (++slot == RT_EVC) { slot = 0; chunk = chunk_next[chunk]; if (chunk == RT_NOCHUNK) i = m.nevents; }
         if (e.row >= own_end) break;                                   /* the successor owns the rest */
         const uint64_t tl = e.row / period;
         if (tl != tile) {
            if (n && tile < ntiles) { atomicAdd(&counts[tile], (unsigned long long)n); atomicAdd(&digests[tile], (unsigned long long)acc); }
            tile = tl; acc = 0; n = 0; }
         const double now = dg_row_time(c, e.row);
         const double hsd = (now - e.t_event) / ((double)c.sample_deltat * 0.5);
         const long long hs = (long long)(hsd < 0 ? hsd - 0.5 : hsd + 0.5);
         /* the reference's own expressions for that row and half-sample count */
         const double t_peak = now - (double)(((float)hs * 0.5f) * c.sample_deltat);
         const double t_row = (hs & 1) == 0 && (long long)e.row >= hs / 2 ? dg_row_time(c, e.row - (uint64_t)(hs / 2)) : -1.0;
         if (!(c.det == RT_DET_PEAK ? e.t_event == t_peak : e.t_event == t_row)) ++nbad;
         const uint64_t k0 = (e.row - tl * period) | ((uint64_t)e.trk << 40) | ((uint64_t)e.kind << 48) | ((uint64_t)(hs & 0xfff) << 52);
         const uint64_t k1 = (uint64_t)__float_as_uint(e.v_top) | ((uint64_t)__float_as_uint(e.v_bot) << 32);
         const uint64_t k2 = (uint64_t)__float_as_uint(e.agc_gain);
         acc += mix64(mix64(mix64(k0) ^ k1) ^ k2);
         ++n; }
      if (n && tile < ntiles) { atomicAdd(&counts[tile], (unsigned long long)n); atomicAdd(&digests[tile], (unsigned long long)acc); }
      if (nbad) atomicAdd(bad, (unsigned long long)nbad); } }

cudaError_t launch_tile_digest(const DevCfg &c, const UnitDesc *units, uint32_t nunits, const TrkMeta *meta, const rt_event *pool,
                               const uint32_t *chunk_next, uint64_t period, uint64_t ntiles, unsigned long long *counts,
                               unsigned long long *digests, unsigned long long *bad, int sms, cudaStream_t s) {
   const uint64_t total = (uint64_t)nunits * (uint64_t)c.ntrks;
   uint64_t grid = (total + 127) / 128;
   if (grid > (uint64_t)sms * 16) grid = (uint64_t)sms * 16;
   if (grid < 1) grid = 1;
   k_tile_digest<<<(unsigned)grid, 128, 0, s>>>(c, units, nunits, meta, pool, chunk_next, period, ntiles, counts, digests, bad);
   return cudaGetLastError(); }
