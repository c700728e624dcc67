/* tests/host_fast/fast_host.cu -- TEST INFRASTRUCTURE: a HOST build of the product's fast-path scan code.
 *
 * readtape_b200/csrc/scan_fast.cuh is written as __host__ __device__ code; this file instantiates the very
 * same UnitScan<> on the CPU (lane stride 1, flat event buffer) so that the algorithm can be compared with
 * the oracle and the reference's events where there is no GPU.  It is never linked into the product.
 * Build: make -C tests/host_fast   (nvcc as a host compiler, -ffp-contract=off like every host unit).
 */
#include <vector>
#include <cstdint>
#include "scan_fast.cuh"
#include "scan_zc.cuh"
#include "cfg_host.h"

struct HostEmit {
   rt_event *buf; uint32_t cap, n; uint64_t first_row; uint32_t first_chunk; uint8_t trk;
   void emit(uint64_t row, double t_ev, float v_top, float v_bot, float agc, bool top) {
      if (n == 0) first_row = row;
      if (n < cap) {
         rt_event e;
         e.row = row; e.t_event = t_ev; e.v_top = v_top; e.v_bot = v_bot; e.agc_gain = agc;
         e.trk = trk; e.kind = top ? RT_EV_TOP : RT_EV_BOT; e.pad[0] = e.pad[1] = 0;
         buf[n] = e; }
      ++n; } };

/* the job list of one lane: the tracks of one unit, one after the other */
struct HostJobs {
   const DevCfg &dc; const int16_t *planes; uint64_t plane_stride, row0, row_end; rt_event *out; uint32_t cap; uint32_t *counts; TrkMeta *meta;
   int thr, k; bool exhausted;
   template <class Scan> bool next(Scan &us) {
      exhausted = k >= dc.ntrks;
      if (exhausted) return false;
      HostEmit em{out + (size_t)k * cap, cap, 0, RT_NOROW, RT_NOCHUNK, (uint8_t)k};
      us.begin(planes + (size_t)k * plane_stride, row0, row_end, k, em, thr);
      return true; }
   template <class Scan> void done(Scan &us) { us.finish(meta[k]); counts[k] = us.em.n; ++k; } };
struct HostCount { int operator()(bool p) const { return p ? 1 : 0; } };

/* planes: [ntrks][plane_stride] int16 (track-major, >= 16 readable rows past row_end).  Scans every track of the
 * unit [row0, row_end) from a fresh RT_RESET_FULL; events of track k go to out[k*cap ...], counts[k] = events produced
 * (may exceed cap), meta[k] = proof data.  Returns 0, or -4 if the configuration is not eligible for the fast path. */
extern "C" int fast_host_scan_unit(const int16_t *planes, uint64_t plane_stride, uint64_t nrows, const rt_tape_desc *desc,
                                   const rt_scan_cfg *cfg, uint64_t row0, uint64_t row_end, rt_event *out, uint32_t cap,
                                   uint32_t *counts, TrkMeta *meta, int skip_gaps) {
   DevCfg dc;
   rtcfg::to_dev(*desc, planes, plane_stride, nrows, cfg, &dc);
   std::vector<uint32_t> gmm;                                      /* the granule min/max map k_ingest.cu builds on the device */
   if (skip_gaps) {
      const uint64_t ngran = (nrows + RT_GRAN - 1) / RT_GRAN + 1;
      gmm.assign((size_t)ngran * dc.ntrks, 0);
      for (int k = 0; k < dc.ntrks; ++k)
         for (uint64_t g = 0; g * RT_GRAN < nrows; ++g) {
            int mn = 32767, mx = -32768;
            for (uint64_t r = g * RT_GRAN; r < (g + 1) * RT_GRAN && r < nrows; ++r) { int v = planes[(size_t)k * plane_stride + r]; if (v < mn) mn = v; if (v > mx) mx = v; }
            gmm[(size_t)k * ngran + g] = ((uint32_t)mn & 0xffffu) | ((uint32_t)mx << 16); }
      dc.gmm = gmm.data(); dc.ngran_cap = ngran; }
   if (rtfast::zc_scan_eligible(dc)) {                             /* the zero-crossing fast path (GCR -zeros) */
      std::vector<uint32_t> zs(rtfast::zc_scratch_words(dc), 0x7fff8000u);
      rtfast::ZcMem<1> zm = rtfast::zc_mem<1>(zs.data());
      HostJobs zjobs{dc, planes, plane_stride, row0, row_end, out, cap, counts, meta, 0, 0, false};
      rtfast::ZcScan<1, HostEmit> zus(dc, zm);
      rtfast::drive(zus, zjobs, HostCount());
      return RT_OK; }
   const bool eligible = dc.det == RT_DET_PEAK && (dc.mode == RT_MODE_NRZI || dc.mode == RT_MODE_PE) && !dc.invert && !dc.differentiate
                         && !dc.density && dc.width >= 3 && dc.width <= RT_PKWW_MAX_WIDTH;
   if (!eligible) return RT_ERR_UNSUPPORTED;
   std::vector<uint32_t> scratch(rtfast::scratch_words(dc.width), 0x7fff8000u);   /* poison: the largest sample and the smallest complement */
   rtfast::LaneMem<1> mem = rtfast::lane_mem<1>(scratch.data(), dc.width);
   HostJobs jobs{dc, planes, plane_stride, row0, row_end, out, cap, counts, meta, rtcfg::quiet_thr_lsb(dc), 0, false};
   rtfast::UnitScan<1, HostEmit> us(dc, mem);
   rtfast::drive(us, jobs, HostCount());
   return RT_OK; }

extern "C" int fast_host_meta_size(void) { return (int)sizeof(TrkMeta); }
