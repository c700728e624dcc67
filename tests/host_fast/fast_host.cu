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
#include "scan_sparse.cuh"
#include <map>
#include <tuple>
#include "cfg_host.h"
#include "lookup_rules.h"

struct HostEmit {
   rt_event *buf; uint32_t cap, n; uint64_t first_row; uint32_t first_chunk; uint8_t trk; uint64_t last_row;
   void emit(uint64_t row, double t_ev, float v_top, float v_bot, float agc, bool top) {
      if (n == 0) first_row = row;
      last_row = row;
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
      HostEmit em{out + (size_t)k * cap, cap, 0, RT_NOROW, RT_NOCHUNK, (uint8_t)k, RT_NOROW};
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
   const bool eligible = dc.det == RT_DET_PEAK && (dc.mode == RT_MODE_NRZI || dc.mode == RT_MODE_PE || dc.density) && !dc.invert && !dc.differentiate
                         && dc.width >= 3 && dc.width <= RT_PKWW_MAX_WIDTH;
   if (!eligible) return RT_ERR_UNSUPPORTED;
   std::vector<uint32_t> scratch(rtfast::scratch_words(dc.width), 0x7fff8000u);   /* poison: the largest sample and the smallest complement */
   rtfast::LaneMem<1> mem = rtfast::lane_mem<1>(scratch.data(), dc.width);
   HostJobs jobs{dc, planes, plane_stride, row0, row_end, out, cap, counts, meta, rtcfg::quiet_thr_lsb(dc), 0, false};
   rtfast::UnitScan<1, HostEmit> us(dc, mem);
   rtfast::drive(us, jobs, HostCount());
   return RT_OK; }

/* ---- K3c: phase A masks + sparse scan (scan_masks.cuh / scan_sparse.cuh), host build ---------------------------------- */
template <int W>
static void host_masks_w(const int16_t *plane, uint64_t nruns, uint32_t T0, uint32_t T1, uint32_t *cand, uint32_t *cand2, uint32_t *acan) {
   for (uint64_t r = 0; r < nruns; ++r) {
      const int64_t p0 = (int64_t)r * rtmask::MASK_RUN;
      uint32_t cw[2], dw[2], aw[2];
      if (p0 >= rtmask::RunMasks<W>::HALO) rtmask::RunMasks<W>::run(plane, p0, T0, T1, cw, dw, aw);
      else { rtmask::word_masks_scalar(plane, p0 / 32, W, (int)T0, (int)T1, &cw[0], &dw[0], &aw[0]); rtmask::word_masks_scalar(plane, p0 / 32 + 1, W, (int)T0, (int)T1, &cw[1], &dw[1], &aw[1]); }
      cand[2 * r] = cw[0]; cand[2 * r + 1] = cw[1]; cand2[2 * r] = dw[0]; cand2[2 * r + 1] = dw[1]; acan[2 * r] = aw[0]; acan[2 * r + 1] = aw[1]; } }

/* masks of rows [0, nruns*64) of every track (plane_stride >= nruns*64); simd=0: the brute-force definition instead */
extern "C" int masks_host_build(const int16_t *planes, uint64_t plane_stride, int ntrks, uint64_t nruns, int w, int T0, int T1,
                                uint32_t *cand, uint32_t *cand2, uint32_t *acan, uint64_t mask_stride, int simd) {
   if (w < 3 || w > RT_PKWW_MAX_WIDTH || T0 < 1 || T0 > 65535 || T1 < T0 || T1 > 65535) return RT_ERR_ARG;
   for (int k = 0; k < ntrks; ++k) {
      const int16_t *plane = planes + (size_t)k * plane_stride;
      uint32_t *c = cand + (size_t)k * mask_stride, *d = cand2 + (size_t)k * mask_stride, *a = acan + (size_t)k * mask_stride;
      if (!simd) { for (uint64_t wi = 0; wi < 2 * nruns; ++wi) rtmask::word_masks_scalar(plane, (int64_t)wi, w, T0, T1, &c[wi], &d[wi], &a[wi]); continue; }
      switch (w) {
#define MW(W) case W: host_masks_w<W>(plane, nruns, (uint32_t)T0, (uint32_t)T1, c, d, a); break;
         MW(3) MW(4) MW(5) MW(6) MW(7) MW(8) MW(9) MW(10) MW(11) MW(12) MW(13) MW(14) MW(15) MW(16) MW(17) MW(18) MW(19) MW(20)
         MW(21) MW(22) MW(23) MW(24) MW(25) MW(26) MW(27) MW(28) MW(29) MW(30) MW(31) MW(32) MW(33) MW(34) MW(35) MW(36) MW(37) MW(38)
         MW(39) MW(40) MW(41) MW(42) MW(43) MW(44) MW(45) MW(46) MW(47) MW(48) MW(49) MW(50)
#undef MW
      } }
   return RT_OK; }

struct HostSparseJobs {
   const DevCfg &dc; const int16_t *planes; uint64_t plane_stride, row0, row_end; rt_event *out; uint32_t cap; uint32_t *counts; TrkMeta *meta;
   int thr, k; bool exhausted;
   template <class Scan> bool next(Scan &us) {
      exhausted = k >= dc.ntrks;
      if (exhausted) return false;
      HostEmit em{out + (size_t)k * cap, cap, 0, RT_NOROW, RT_NOCHUNK, (uint8_t)k, RT_NOROW};
      us.begin(planes + (size_t)k * plane_stride, row0, row_end, k, em, thr);
      return true; }
   template <class Scan> void done(Scan &us) { us.finish(meta[k]); counts[k] = us.em.n; ++k; } };
struct HostAny { bool operator()(bool p) const { return p; } };

/* Same contract as fast_host_scan_unit, through the two-pass scan.  t0_frac: T0 as a fraction of the default-state threshold bound
   (1.0 exercises the dense fall-back whenever the AGC lowers the threshold; tiny values make every row a candidate).
   The masks and the granule map of a capture are cached between calls (keyed by plane pointer, width, T0, rows). */
static uint64_t g_planes_id = 0;
extern "C" int sparse_host_scan_unit(const int16_t *planes, uint64_t plane_stride, uint64_t nrows, const rt_tape_desc *desc,
                                     const rt_scan_cfg *cfg, uint64_t row0, uint64_t row_end, rt_event *out, uint32_t cap,
                                     uint32_t *counts, TrkMeta *meta, float t0_frac, int use_gmm, int use_records) {
   DevCfg dc;
   rtcfg::to_dev(*desc, planes, plane_stride, nrows, cfg, &dc);
   const bool eligible = dc.det == RT_DET_PEAK && (dc.mode == RT_MODE_NRZI || dc.mode == RT_MODE_PE || dc.density) && !dc.invert && !dc.differentiate
                         && dc.width >= 3 && dc.width <= RT_PKWW_MAX_WIDTH;
   if (!eligible) return RT_ERR_UNSUPPORTED;
   const float inv_lsb = 32767.0f / dc.maxvolts;
   const float q = dc.p.pkww_rise * inv_lsb * 0.999f - 2.0f;
   int T0 = q > 0 ? (int)((q > 70000.0f ? 70000.0f : (float)(int)q) * t0_frac) : 0;
   if (T0 > 65535) T0 = 65535;
   if (T0 < 1) return RT_ERR_UNSUPPORTED;
   struct Cache { std::vector<uint32_t> cand, cand2, acan, gmm; uint64_t mask_stride, ngran;
                  std::vector<CandRec> recs; std::vector<uint32_t> tbase, tcnt; uint64_t rec_tiles; };
   const int T1 = T0 * 8 / 5 <= 65535 ? T0 * 8 / 5 : 65535;      /* a second plane at 1.6 x T0, as the library does for a fixed RT_SPARSE_T0 */
   /* keyed by the CONTENT of the planes too: a test's next array may well land at the address of the last one */
   uint64_t fp = 1469598103934665603ull;
   if (g_planes_id) fp = g_planes_id;                             /* hostsim.cu: the caller vouches for the content (one id per build of the planes) */
   else for (int k = 0; k < dc.ntrks; ++k) {
      const int16_t *pl = planes + (size_t)k * plane_stride;
      for (uint64_t r = 0; r < nrows; ++r) fp = (fp ^ (uint16_t)pl[r]) * 1099511628211ull; }
   static std::map<std::tuple<const int16_t *, uint64_t, int, int, int, uint64_t>, Cache> cache;
   auto key = std::make_tuple(planes, nrows, dc.ntrks, dc.width, T0, fp);
   auto it = cache.find(key);
   if (it == cache.end()) {
      if (cache.size() > 8) cache.clear();
      Cache cc;
      const uint64_t nruns = (nrows + rtmask::MASK_RUN - 1) / rtmask::MASK_RUN;
      if (nruns * rtmask::MASK_RUN > plane_stride) return RT_ERR_ARG;
      cc.mask_stride = 2 * nruns + 4;
      cc.cand.assign((size_t)cc.mask_stride * dc.ntrks, 0); cc.cand2.assign((size_t)cc.mask_stride * dc.ntrks, 0); cc.acan.assign((size_t)cc.mask_stride * dc.ntrks, 0);
      masks_host_build(planes, plane_stride, dc.ntrks, nruns, dc.width, T0, T1, cc.cand.data(), cc.cand2.data(), cc.acan.data(), cc.mask_stride, 1);
      cc.ngran = (nrows + RT_GRAN - 1) / RT_GRAN + 1;
      cc.gmm.assign((size_t)cc.ngran * dc.ntrks, 0);
      for (int k = 0; k < dc.ntrks; ++k)
         for (uint64_t g = 0; g * RT_GRAN < nrows; ++g) {
            int mn = 32767, mx = -32768;
            for (uint64_t r = g * RT_GRAN; r < (g + 1) * RT_GRAN && r < nrows; ++r) { int v = planes[(size_t)k * plane_stride + r]; if (v < mn) mn = v; if (v > mx) mx = v; }
            cc.gmm[(size_t)k * cc.ngran + g] = ((uint32_t)mn & 0xffffu) | ((uint32_t)mx << 16); }
      /* phase B1 on the host: the records of every candidate row, tile by tile (the device places the tiles in arbitrary order) */
      cc.rec_tiles = plane_stride / RT_REC_TILE + 2;
      cc.tbase.assign((size_t)cc.rec_tiles * dc.ntrks, 0); cc.tcnt.assign((size_t)cc.rec_tiles * dc.ntrks, 0);
      for (int k = 0; k < dc.ntrks; ++k) {
         const int16_t *plane = planes + (size_t)k * plane_stride;
         const uint32_t *mc = cc.cand.data() + (size_t)k * cc.mask_stride, *ma = cc.acan.data() + (size_t)k * cc.mask_stride;
         for (uint64_t tile = 0; tile * RT_REC_TILE < nrows; ++tile) {
            cc.tbase[(size_t)k * cc.rec_tiles + tile] = (uint32_t)cc.recs.size();
            uint32_t n = 0;
            for (uint64_t p = tile * RT_REC_TILE; p < (tile + 1) * RT_REC_TILE && p < nrows; ++p)
               if ((mc[p >> 5] >> (p & 31)) & 1u) { cc.recs.push_back(rtrec::make_record(plane, ma, dc.width, T0, p)); ++n; }
            cc.tcnt[(size_t)k * cc.rec_tiles + tile] = n; } }
      it = cache.emplace(key, std::move(cc)).first; }
   Cache &cc = it->second;
   dc.m_cand = cc.cand.data(); dc.m_cand2 = cc.cand2.data(); dc.m_acan = cc.acan.data(); dc.mask_stride = cc.mask_stride;
   for (int k = 0; k < RT_MAXTRKS; ++k) { dc.T0[k] = T0; dc.T1[k] = T1; }
   if (use_gmm) { dc.gmm = cc.gmm.data(); dc.ngran_cap = cc.ngran; }
   uint32_t heights[RT_AGC_MAX_WINDOW];
   HostSparseJobs jobs{dc, planes, plane_stride, row0, row_end, out, cap, counts, meta, rtcfg::quiet_thr_lsb(dc), 0, false};
   if (use_records) {
      dc.recs = cc.recs.data(); dc.rec_tile_base = cc.tbase.data(); dc.rec_tile_cnt = cc.tcnt.data(); dc.rec_tiles = cc.rec_tiles;
      rtsparse::SparseScan<1, HostEmit, true> us(dc, heights);
      rtsparse::drive_sparse(us, jobs, HostAny()); }
   else {
      rtsparse::SparseScan<1, HostEmit> us(dc, heights);
      rtsparse::drive_sparse(us, jobs, HostAny()); }
   return RT_OK; }

#ifdef RT_SPARSE_STATS
extern "C" void sparse_host_stats(unsigned long long *out) { memcpy(out, &rtsparse::stats(), sizeof(rtsparse::Stats)); memset(&rtsparse::stats(), 0, sizeof(rtsparse::Stats)); }
#endif
/* volts() of the scan code for every int16 value (exhaustive check of the division-free formulation) */
extern "C" void volts_host_all(float maxvolts, float *out /* [65536], index x + 32768 */) {
   DevCfg dc; memset(&dc, 0, sizeof dc); dc.maxvolts = maxvolts;
   for (int x = -32768; x <= 32767; ++x) out[x + 32768] = rtfast::volts(dc, x); }

extern "C" int fast_host_meta_size(void) { return (int)sizeof(TrkMeta); }

/* debugging aid for tests: the candidate record of plane row p of track trk, with the masks built as sparse_host_scan_unit does */
extern "C" int rec_host_debug(const int16_t *planes, uint64_t plane_stride, uint64_t nrows, const rt_tape_desc *desc, const rt_scan_cfg *cfg,
                              float t0_frac, int trk, uint64_t p, CandRec *out, uint32_t *acan_bits /* 64 rows ending at p */, int *T0_out) {
   DevCfg dc;
   rtcfg::to_dev(*desc, planes, plane_stride, nrows, cfg, &dc);
   const float inv_lsb = 32767.0f / dc.maxvolts;
   const float q = dc.p.pkww_rise * inv_lsb * 0.999f - 2.0f;
   int T0 = q > 0 ? (int)((q > 70000.0f ? 70000.0f : (float)(int)q) * t0_frac) : 0;
   const int T1 = T0 * 8 / 5 <= 65535 ? T0 * 8 / 5 : 65535;
   const uint64_t nruns = (nrows + rtmask::MASK_RUN - 1) / rtmask::MASK_RUN;
   const uint64_t ms = 2 * nruns + 4;
   std::vector<uint32_t> cand(ms * dc.ntrks), cand2(ms * dc.ntrks), acan(ms * dc.ntrks);
   masks_host_build(planes, plane_stride, dc.ntrks, nruns, dc.width, T0, T1, cand.data(), cand2.data(), acan.data(), ms, 1);
   const uint32_t *ma = acan.data() + (size_t)trk * ms;
   *out = rtrec::make_record(planes + (size_t)trk * plane_stride, ma, dc.width, T0, p);
   for (int i = 0; i < 64; ++i) { const uint64_t r = p - 63 + i; acan_bits[i] = (ma[r >> 5] >> (r & 31)) & 1u; }
   *T0_out = T0;
   return (int)((cand[(size_t)trk * ms + (p >> 5)] >> (p & 31)) & 1u); }

/* ---- the exact stateful scan (scan_generic.cuh: what k_ctx_scan runs per track) on the host -----------------------------------------
 * Fresh RT_RESET_FULL at reset_row, then rows [reset_row, row_to) in spans of span_rows (one span = one kernel launch of the
 * library), with the skip-ahead over the mask planes (use_masks; T0 = t0_frac of the default-state bound, as scan_prepare_masks
 * chooses it) or walking every row.  stats[5]: rows walked, jumps, rows jumped over, rows with the threshold below the masks',
 * candidates too near.  The events must equal the oracle's whatever the spans and the masks. */
#include "scan_generic.cuh"
struct GenEmit {
   rt_event *buf; uint32_t cap, n; uint8_t trk;
   void emit(uint64_t row, double t_ev, float v_top, float v_bot, float agc, bool top) {
      if (n < cap) {
         rt_event e;
         e.row = row; e.t_event = t_ev; e.v_top = v_top; e.v_bot = v_bot; e.agc_gain = agc;
         e.trk = trk; e.kind = top ? RT_EV_TOP : RT_EV_BOT; e.pad[0] = e.pad[1] = 0;
         buf[n] = e; }
      ++n; } };

extern "C" int generic_host_ctx_scan(const int16_t *planes, uint64_t plane_stride, uint64_t nrows, const rt_tape_desc *desc, const rt_scan_cfg *cfg,
                                     uint64_t reset_row, uint64_t row_to, uint64_t span_rows, int use_masks, float t0_frac,
                                     rt_event *out, uint32_t cap, uint32_t *counts, unsigned long long *stats) {
   DevCfg dc;
   rtcfg::to_dev(*desc, planes, plane_stride, nrows, cfg, &dc);
   if (row_to > nrows) row_to = nrows;
   std::vector<uint32_t> cand, cand2, acan;
   if (use_masks && dc.det == RT_DET_PEAK && !dc.invert && !dc.differentiate && dc.width >= 3 && dc.width <= RT_PKWW_MAX_WIDTH) {
      const float inv_lsb = 32767.0f / dc.maxvolts;
      const float q = dc.p.pkww_rise * inv_lsb * 0.999f - 2.0f;
      int T0 = q > 0 ? (int)((float)(q > 70000.0f ? 70000 : (int)q) * t0_frac) : 0;
      if (T0 > 65535) T0 = 65535;
      if (T0 >= 16) {
         const int T1 = T0 * 8 / 5 <= 65535 ? T0 * 8 / 5 : 65535;
         const uint64_t nruns = (nrows + rtmask::MASK_RUN - 1) / rtmask::MASK_RUN;
         if (nruns * rtmask::MASK_RUN > plane_stride) return RT_ERR_ARG;
         const uint64_t ms = 2 * nruns + 4;
         cand.assign((size_t)ms * dc.ntrks, 0); cand2.assign((size_t)ms * dc.ntrks, 0); acan.assign((size_t)ms * dc.ntrks, 0);
         masks_host_build(planes, plane_stride, dc.ntrks, nruns, dc.width, T0, T1, cand.data(), cand2.data(), acan.data(), ms, 1);
         dc.m_cand = cand.data(); dc.m_cand2 = cand2.data(); dc.m_acan = acan.data(); dc.mask_stride = ms;
         for (int k = 0; k < RT_MAXTRKS; ++k) { dc.T0[k] = T0; dc.T1[k] = T1; } } }
   rtgen::CtxStats cs{0, 0, 0, 0, 0};
   const bool tz = rtgen::row_time(dc, reset_row) == 0.0;
   int failed = 0;
   for (int k = 0; k < dc.ntrks; ++k) {
      TrkState t; rtgen::SkewState s;
      rtgen::reset_full(dc, t, s, k, reset_row, tz);
      GenEmit em{out + (size_t)k * cap, cap, 0, (uint8_t)k};
      const int16_t *plane = planes + (size_t)k * plane_stride;
      for (uint64_t from = reset_row; from < row_to; from += span_rows)
         rtgen::ctx_scan_rows(dc, t, s, k, plane, from, from + span_rows < row_to ? from + span_rows : row_to, em, cs);
      counts[k] = em.n;
      if (t.failed) failed = t.failed; }
   if (stats) { stats[0] = cs.walked; stats[1] = cs.jumps; stats[2] = cs.jumped; stats[3] = cs.nothr; stats[4] = cs.near; }
   return failed ? -100 - failed : RT_OK; }

/* ---- the unit-equivalence rules of rt_bulk_lookup (lookup_rules.h: the product's own code) ------------------------------------------
 * meta = the TrkMeta[ntrks] a *_host_scan_unit call returned for the unit [row0, row_end).  lookup_host_covers: 1 = a fresh reset at
 * start_row is covered (quiet rule), 0 = not; *bridge_to as unit_covers sets it.  lookup_host_tail / lookup_host_chain: the tail rule
 * and the chaining of an event-free unit into the next one. */
extern "C" int lookup_host_covers(const rt_tape_desc *desc, const rt_scan_cfg *cfg, uint64_t row0, uint64_t row_end, const TrkMeta *meta, uint64_t start_row, uint64_t *bridge_to) {
   DevCfg dc; rtcfg::to_dev(*desc, nullptr, 0, 0, cfg, &dc);
   const UnitDesc u{row0, row_end};
   const bool tz = (long long)(dc.tstart_ns + start_row * dc.tdelta_ns) == 0;
   return rtlookup::unit_covers(dc, u, meta, (uint32_t)dc.ntrks, start_row, tz, bridge_to) ? 1 : 0; }
extern "C" int lookup_host_tail(const rt_tape_desc *desc, uint64_t row0, uint64_t row_end, const TrkMeta *meta, uint64_t start_row) {
   const UnitDesc u{row0, row_end};
   return rtlookup::unit_tail_covers(u, meta, desc->ntrks, start_row) ? 1 : 0; }
extern "C" int lookup_host_chain(const rt_tape_desc *desc, const rt_scan_cfg *cfg, uint64_t row0, uint64_t row_end, const TrkMeta *meta, const TrkMeta *meta_next, uint64_t start_row) {
   DevCfg dc; rtcfg::to_dev(*desc, nullptr, 0, 0, cfg, &dc);
   const UnitDesc u{row0, row_end};
   return rtlookup::chains_into_next(dc, u, meta, meta_next, desc->ntrks, start_row) ? 1 : 0; }

/* ---- the speculative unit scan on the exact generic detector code (k_units_scan, k_scan.cu) on the host ------------------------------
 * k_units_scan serves what the fast kernels do not: -differentiate (zero-crossing of the differentiated signal, and the peak detector
 * on it), widths outside the fast path.  Its per-(unit, track) body -- track_row() + QuietTracker + the proof-data bookkeeping -- is
 * MIRRORED here statement for statement (k_scan.cu:61-100; the detector and the tracker themselves are the shared __host__ __device__
 * code of scan_generic.cuh), so that the unit-equivalence rules can be attacked for those detectors too (tests/test_proof_host.py).
 * Keep the two in step; the kernel's body becomes a call of a shared function the next time it can be re-verified on hardware. */
extern "C" int generic_host_scan_unit(const int16_t *planes, uint64_t plane_stride, uint64_t nrows, const rt_tape_desc *desc, const rt_scan_cfg *cfg,
                                      uint64_t row0, uint64_t row_end, rt_event *out, uint32_t cap, uint32_t *counts, TrkMeta *meta) {
   using namespace rtgen;
   DevCfg c;
   rtcfg::to_dev(*desc, planes, plane_stride, nrows, cfg, &c);
   if (row_end > nrows) row_end = nrows;
   float quiet_thr = 0;                                           /* make_plan (rt_api.cu) */
   if (c.det == RT_DET_PEAK) { quiet_thr = c.p.pkww_rise * 0.999f; if (c.p.pkww_rise < 1e-3f) quiet_thr = 0; }
   const int quiet_thr_lsb = rtcfg::quiet_thr_lsb(c);
   const UnitDesc ud{row0, row_end};
   for (int trk = 0; trk < c.ntrks; ++trk) {
      TrkState t; SkewState s;
      reset_full(c, t, s, trk, ud.row0, row_time(c, ud.row0) == 0.0);
      const int16_t *plane = c.planes + (size_t)trk * c.plane_stride;
      HostEmit em{out + (size_t)trk * cap, cap, 0, RT_NOROW, RT_NOCHUNK, (uint8_t)trk, RT_NOROW};
      QuietTracker qt; qt.init(c, trk, quiet_thr, quiet_thr_lsb);
      const uint64_t pre0 = ud.row0 > (uint64_t)c.prescan_rows ? ud.row0 - (uint64_t)c.prescan_rows : 0;
      for (uint64_t j = pre0; j < ud.row0; ++j) qt.feed(c, plane, j, raw_at(c, plane, j));
      const uint64_t quiet_from = qt.last_loud == RT_NOROW ? pre0 : qt.last_loud + 1;
      uint64_t sync_row = RT_NOROW, loud_at_sync = RT_NOROW, sync_first = RT_NOROW, sync_early = RT_NOROW, loud_early = RT_NOROW;
      bool early_frozen = false;
      const int lead = std::max(trk + (row_time(c, ud.row0) == 0.0 ? 1 : 0), c.skew[trk]);
      const uint64_t own_fill = ud.row0 + (uint64_t)(c.det == RT_DET_PEAK ? lead + c.width + 1 : lead + 2);
      for (uint64_t row = ud.row0; row < ud.row_end; ++row) {
         float v_raw;
         unsigned probe = track_row(c, t, s, trk, plane, row, em, &v_raw);
         if (em.n == 0) {
            qt.feed(c, plane, row, v_raw);
            const bool canonical = c.det == RT_DET_PEAK ? (probe & 2) != 0 : row >= own_fill;
            if (sync_early != RT_NOROW && qt.last_loud != loud_early) early_frozen = true;
            if (canonical && (qt.last_loud == RT_NOROW || qt.last_loud < row)) {
               sync_row = row; loud_at_sync = qt.last_loud;
               if (!early_frozen) { sync_early = row; loud_early = qt.last_loud; }
               if (sync_first == RT_NOROW && row >= own_fill && (qt.last_loud == RT_NOROW || qt.last_loud < ud.row0)) sync_first = row; } } }
      TrkMeta m;
      m.first_event_row = em.first_row; m.sync_row = sync_row; m.last_loud_row = loud_at_sync; m.sync_first = sync_first; m.quiet_from = quiet_from; m.sync_early = sync_early; m.loud_early = loud_early;
      m.first_chunk = em.first_chunk; m.nevents = em.n; m.failed = t.failed; m.pad = 0;
      m.last_event_row = em.last_row; m.quiet_tail_from = RT_NOROW;
      meta[trk] = m; counts[trk] = em.n; }
   return RT_OK; }
