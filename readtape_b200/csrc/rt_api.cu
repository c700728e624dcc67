/* readtape_b200/csrc/rt_api.cu -- host side of librt_scan_b200.so: the C-ABI of include/rt_scan.h.
 *
 * Plain CUDA runtime, no torch, no CPU compute path: every scan goes through the kernels in
 * k_ingest.cu / k_units.cu / k_scan.cu / k_fast.cu.  If no CUDA device (or no sm_100a kernel
 * image) is usable, rt_open() fails with RT_ERR_NODEVICE -- there is no fallback.
 */
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <unistd.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <chrono>
#include "scan_generic.cuh"
#include "scan_records.cuh"
#include "kernels.h"
#include "cfg_host.h"
#include "lookup_rules.h"
#include "rt_internal.h"

using rtgen::SkewState;

/* ---- errors ------------------------------------------------------------------------------------ */
static thread_local char g_err[512];
static int set_err(int code, const char *fmt, ...) {
   va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
   return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
      return set_err(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" const char *rt_last_error(void) { return g_err; }
int rt_fail(int code, const char *fmt, ...) {                     /* rt_internal.h: set_err for the other translation units */
   va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
   return code; }

/* process-wide options (rt_set_option) */
static int g_opt_shared_results = 0;
extern "C" int rt_set_option(int option, int value) {
   if (option == RT_OPT_SHARED_RESULTS) { g_opt_shared_results = value != 0; return RT_OK; }
   return set_err(RT_ERR_ARG, "rt_set_option: unknown option %d", option); }
extern "C" int rt_abi_version(void) { return RT_ABI_VERSION; }
extern "C" const char *rt_backend(void) { return "cuda-sm100a"; }

extern "C" double rt_row_time(const rt_tape_desc *d, uint64_t row) {
   long long ns = (long long)(d->tstart_ns + row * d->tdelta_ns);
   return (double)ns / 1e9; }

extern "C" int rt_pkww_width(const rt_scan_cfg *cfg, uint64_t tdelta_ns) { return rtcfg::pkww_width(cfg, tdelta_ns); }

/* ---- tape ---------------------------------------------------------------------------------------- */
#define RT_STAGE_BUFS 6
#define RT_RING_SLOTS 16      /* one reader thread fills a slot at ~2 GB/s from the page cache: the slots in flight set the file -> GPU rate */
struct rt_tape {
   rt_tape_desc desc{};
   int device = 0, sms = 148;
   cudaStream_t stream = nullptr;
   int32_t trk_of_head[RT_MAXTRKS];
   uint64_t nrows = 0;            /* rows ingested */
   uint64_t cap_rows = 0;         /* capacity of the planes, rows */
   int16_t *planes = nullptr; uint64_t plane_stride = 0;
   int16_t *gmm = nullptr; uint64_t ngran_cap = 0;
   /* -invert: negated copies (same strides), made on demand by rt_bulk_scan for rows [0, inv_rows) */
   int16_t *planes_inv = nullptr, *gmm_inv = nullptr; uint64_t inv_rows = 0;
   unsigned long long *d_first_end = nullptr;
   uint64_t nrows_valid = 0; bool valid_known = false;
   /* device staging buffers of the chunked host->device copy: two by default.  RT_STAGE=n (up to 6) lets the copy engine run further
      ahead of the ingest kernel when scan kernels of earlier segments hold the SMs (rt_bulk_scan_host); measured on a B200: no
      difference (460 / 482 / 556 / 516 / 572 ms end to end for n = 2 / 6 / 2 / 6 / 4 on one box -- the spread is the host's) */
   int nstage = 2;                /* buffers in use (RT_STAGE=n, 2..RT_STAGE_BUFS) */
   int16_t *d_stage[RT_STAGE_BUFS] = {}; size_t stage_bytes = 0; cudaEvent_t stage_done[RT_STAGE_BUFS] = {};
   int force_simple_ingest = 0;
   int launches = 0;
   float ms_ingest = 0;
   std::vector<cudaEvent_t> ingest_events;
   uint64_t h2d_bytes = 0;
   /* grow-only caches re-used by successive rt_bulk_scan()/rt_bulk_fetch() calls (cudaMalloc/cudaHostAlloc of
      many GB cost far more than the scan itself) */
   rt_event *pool_cache = nullptr; uint32_t *next_cache = nullptr; uint32_t pool_cache_chunks = 0; bool pool_cache_busy = false;
   rt_event *pin_cache = nullptr; size_t pin_cache_events = 0; bool pin_cache_busy = false;
   /* rt_bulk_scan_host(): extra streams, and the sizes the last whole-tape scan needed (capacity planning of the streamed scan) */
   std::vector<cudaStream_t> s_par;                               /* rt_bulk_scan with several configurations: their kernels run side by side */
   cudaStream_t s_scan = nullptr, s_out = nullptr, s_copy = nullptr; cudaEvent_t stage_copied[RT_STAGE_BUFS] = {};
   uint64_t hist_rows = 0; uint32_t hist_units = 0, hist_chunks = 0;
   /* phase B1: candidate records (scan_records.cuh), grow-only like the event pool; one pool shared by the mask sets of a scan */
   CandRec *rec_cache = nullptr; uint32_t rec_cache_cap = 0; uint32_t rec_hist = 0;
   /* rt_prepare(): the bit planes of K3c phase A written by the ingest kernel itself for an announced configuration */
   struct PreMask {
      bool active = false, have_thr = false; rt_scan_cfg cfg{}; int width = 0; float rise = 0;
      int32_t T0[RT_MAXTRKS] = {}, T1[RT_MAXTRKS] = {};
      uint32_t *mc = nullptr, *md = nullptr, *ma = nullptr; uint64_t stride = 0;
      uint64_t rows = 0;                                             /* the planes are complete for tape rows [0, rows) */
      uint64_t fused_rows = 0;                                       /* of which written by the fused kernel (diagnostics) */
      bool overlap = false;                                          /* RT_FUSED_MASKS=2: separate mask kernel, one chunk behind the ingest kernel, on its own stream */
   } pm;
   cudaStream_t s_mask = nullptr; std::vector<cudaEvent_t> mask_events;
   int16_t *h_ring = nullptr; cudaEvent_t ring_done[RT_RING_SLOTS] = {};   /* pinned ring for uploads from pageable memory / files */
   int launches_at_clear = 0;     /* value of `launches` at the last rt_clear() */
   uint32_t chunks_hist = 0;      /* event chunks per configuration the last whole-tape scan of this tape used (first guess of the next) */
   int ring_slots = 0;            /* slots h_ring holds (pinning costs ~0.4 ms per MB: a small capture gets a small ring) */
};

static int tape_reserve(rt_tape *t, uint64_t rows) {
   if (rows <= t->cap_rows) return RT_OK;
   cudaSetDevice(t->device);
   uint64_t newcap = t->cap_rows ? std::max(rows, t->cap_rows * 2) : rows;
   newcap = (newcap + 2047) / 2048 * 2048 + 2048;                 /* tile multiple + slack for vector reads */
   uint64_t ngran = newcap / RT_GRAN + 64;
   int16_t *np = nullptr, *ng = nullptr;
   const uint32_t nt = t->desc.ntrks;
   CU(cudaMalloc(&np, (size_t)newcap * nt * sizeof(int16_t)));
   cudaError_t e = cudaMalloc(&ng, (size_t)ngran * nt * 4);
   if (e != cudaSuccess) { cudaFree(np); return set_err(RT_ERR_CUDA, "cudaMalloc(gmm) failed: %s", cudaGetErrorString(e)); }
   if (t->planes) {
      for (uint32_t k = 0; k < nt; ++k) {
         CU(cudaMemcpyAsync(np + (size_t)k * newcap, t->planes + (size_t)k * t->plane_stride, (size_t)t->nrows * 2, cudaMemcpyDeviceToDevice, t->stream));
         CU(cudaMemcpyAsync(reinterpret_cast<char *>(ng) + (size_t)k * ngran * 4, reinterpret_cast<char *>(t->gmm) + (size_t)k * t->ngran_cap * 4,
                            (size_t)((t->nrows + RT_GRAN - 1) / RT_GRAN) * 4, cudaMemcpyDeviceToDevice, t->stream)); }
      CU(cudaStreamSynchronize(t->stream));
      cudaFree(t->planes); cudaFree(t->gmm); }
   cudaFree(t->planes_inv); cudaFree(t->gmm_inv); t->planes_inv = t->gmm_inv = nullptr; t->inv_rows = 0;   /* other strides now: rebuilt on demand */
   t->planes = np; t->plane_stride = newcap; t->gmm = ng; t->ngran_cap = ngran; t->cap_rows = newcap;
   return RT_OK; }

extern "C" int rt_open(const rt_tape_desc *desc, int device, rt_tape **out) {
   if (!desc || !out) return set_err(RT_ERR_ARG, "rt_open: null argument");
   if (desc->ntrks < 1 || desc->ntrks > RT_MAXTRKS || desc->nheads < desc->ntrks || desc->nheads > RT_MAXTRKS)
      return set_err(RT_ERR_ARG, "rt_open: bad ntrks/nheads %u/%u", desc->ntrks, desc->nheads);
   if (!(desc->maxvolts > 0 && desc->maxvolts < 100.0f)) return set_err(RT_ERR_ARG, "rt_open: maxvolts out of range");
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0)
      return set_err(RT_ERR_NODEVICE, "no CUDA device available (%s); readtape_b200 has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
   if (device < 0 || device >= ndev) return set_err(RT_ERR_ARG, "rt_open: bad device %d (have %d)", device, ndev);
   CU(cudaSetDevice(device));
   cudaDeviceProp prop;
   CU(cudaGetDeviceProperties(&prop, device));
   if (prop.major != 10)
      return set_err(RT_ERR_NODEVICE, "device %d is sm_%d%d; this library only carries sm_100a kernels", device, prop.major, prop.minor);
   rt_tape *t = new (std::nothrow) rt_tape();
   if (!t) return set_err(RT_ERR_NOMEM, "rt_open: out of memory");
   t->desc = *desc; t->device = device; t->sms = prop.multiProcessorCount;
   std::vector<bool> seen(desc->ntrks, false);
   for (uint32_t h = 0; h < RT_MAXTRKS; ++h) {
      int k = h < desc->nheads ? desc->head_to_trk[h] : -1;
      if (k < 0 || k >= (int)desc->ntrks) k = -1;
      else if (seen[k]) { delete t; return set_err(RT_ERR_ARG, "rt_open: track %d is fed by two heads", k); }
      else seen[k] = true;
      t->trk_of_head[h] = k; }
   for (uint32_t k = 0; k < desc->ntrks; ++k) if (!seen[k]) { delete t; return set_err(RT_ERR_ARG, "rt_open: track %u has no head", k); }
   const char *fs = getenv("RT_INGEST");
   t->force_simple_ingest = fs && strcmp(fs, "simple") == 0;
   CU(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
   {  /* the bulk scan allocates its scratch and result tables with the stream-ordered allocator: keep freed memory cached */
      cudaMemPool_t pool; unsigned long long keep = ~0ull;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
   CU(cudaMalloc(&t->d_first_end, sizeof(unsigned long long)));
   unsigned long long none = ~0ull;
   CU(cudaMemcpy(t->d_first_end, &none, sizeof none, cudaMemcpyHostToDevice));
   *out = t; return RT_OK; }

/* enqueue the ingest of `nrows` rows at d_src (device) on the tape's stream; does not wait */
static void cfg_to_dev(const rt_tape *t, const rt_scan_cfg *cfg, DevCfg *d);
static cudaError_t plan_thresholds_from_rows(rt_tape *t, uint64_t rows_avail);

/* the mask planes of an announced configuration (rt_prepare) follow the sample planes' capacity */
static int premask_reserve(rt_tape *t) {
   rt_tape::PreMask &pm = t->pm;
   const uint64_t ms = peak_mask_stride(t->plane_stride);
   if (pm.stride == ms && pm.mc) return RT_OK;
   cudaFree(pm.mc); cudaFree(pm.md); cudaFree(pm.ma); pm.mc = pm.md = pm.ma = nullptr; pm.stride = 0; pm.rows = 0;
   const size_t bytes = (size_t)ms * t->desc.ntrks * 4;
   CU(cudaMalloc(&pm.mc, bytes)); CU(cudaMalloc(&pm.md, bytes)); CU(cudaMalloc(&pm.ma, bytes));
   CU(cudaMemsetAsync(pm.mc, 0, bytes, t->stream)); CU(cudaMemsetAsync(pm.md, 0, bytes, t->stream)); CU(cudaMemsetAsync(pm.ma, 0, bytes, t->stream));
   pm.stride = ms;
   return RT_OK; }

static void premask_devcfg(const rt_tape *t, DevCfg *dc) {
   const rt_tape::PreMask &pm = t->pm;
   cfg_to_dev(t, &pm.cfg, dc);
   dc->m_cand = pm.mc; dc->m_cand2 = pm.md; dc->m_acan = pm.ma; dc->mask_stride = pm.stride;
   for (int k = 0; k < RT_MAXTRKS; ++k) { dc->T0[k] = pm.T0[k]; dc->T1[k] = pm.T1[k]; } }

static int tape_ingest(rt_tape *t, const int16_t *d_src, uint64_t nrows) {
   cudaEvent_t e0, e1;
   CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
   rt_tape::PreMask &pm = t->pm;
   IngestMasks im{}; const IngestMasks *imp = nullptr;
   if (pm.active) { int rc = premask_reserve(t); if (rc) return rc; }
   if (pm.active && pm.have_thr && pm.rows == t->nrows && pm.overlap) {
      /* The ingest kernel is bound by HBM (86 % of the peak, 1 CTA per SM), the mask kernel by the integer pipe (93 %, half the
         bandwidth): they are run side by side -- the rows go through in chunks, the mask kernel of chunk i on a second stream while
         the ingest kernel of chunk i+1 runs (the window only looks back, so a chunk needs nothing from the next one). */
      if (!t->s_mask) CU(cudaStreamCreateWithFlags(&t->s_mask, cudaStreamNonBlocking));
      const char *ce = getenv("RT_OVERLAP_CHUNK");
      const uint64_t chunk = (ce && atoll(ce) > 0 ? (uint64_t)atoll(ce) : (uint64_t)16 << 20) / 2048 * 2048;   /* 4 Mi .. 128 Mi rows measured alike */
      DevCfg dc; premask_devcfg(t, &dc);
      CU(cudaEventRecord(e0, t->stream));
      size_t nev = 0;
      for (uint64_t at = 0; at < nrows; at += chunk) {
         const uint64_t n = std::min(chunk, nrows - at);
         uint64_t masked = 0;
         cudaError_t e = launch_ingest(d_src + at * t->desc.nheads, n, t->nrows + at, (int)t->desc.nheads, t->trk_of_head, t->planes, t->plane_stride,
                                       t->gmm, t->ngran_cap, t->d_first_end, t->sms, t->force_simple_ingest, t->stream, &t->launches, nullptr, &masked);
         if (e != cudaSuccess) return set_err(RT_ERR_CUDA, "ingest kernel launch failed: %s", cudaGetErrorString(e));
         if (nev == t->mask_events.size()) { cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); t->mask_events.push_back(ev); }
         CU(cudaEventRecord(t->mask_events[nev], t->stream));
         CU(cudaStreamWaitEvent(t->s_mask, t->mask_events[nev], 0)); ++nev;
         e = launch_peak_masks(dc, t->nrows + at, t->nrows + at + n, t->s_mask); ++t->launches;
         if (e != cudaSuccess) return set_err(RT_ERR_CUDA, "mask kernel launch failed: %s", cudaGetErrorString(e)); }
      if (nev == t->mask_events.size()) { cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); t->mask_events.push_back(ev); }
      CU(cudaEventRecord(t->mask_events[nev], t->s_mask));
      CU(cudaStreamWaitEvent(t->stream, t->mask_events[nev], 0));   /* whatever follows on the tape's stream sees complete planes */
      CU(cudaEventRecord(e1, t->stream));
      t->ingest_events.push_back(e0); t->ingest_events.push_back(e1);
      pm.rows = t->nrows + nrows;
      t->nrows += nrows; t->valid_known = false;
      return RT_OK; }
   if (pm.active && pm.have_thr && pm.rows == t->nrows) {           /* the planes are complete so far: this chunk's tiles are done on chip */
      im.cand = pm.mc; im.cand2 = pm.md; im.acan = pm.ma; im.mask_stride = pm.stride; im.ntrks = (int)t->desc.ntrks; im.width = pm.width;
      memcpy(im.T0, pm.T0, sizeof im.T0); memcpy(im.T1, pm.T1, sizeof im.T1);
      imp = &im; }
   CU(cudaEventRecord(e0, t->stream));
   uint64_t masked = 0;
   cudaError_t e = launch_ingest(d_src, nrows, t->nrows, (int)t->desc.nheads, t->trk_of_head, t->planes, t->plane_stride,
                                 t->gmm, t->ngran_cap, t->d_first_end, t->sms, t->force_simple_ingest, t->stream, &t->launches, imp, &masked);
   if (e != cudaSuccess) return set_err(RT_ERR_CUDA, "ingest kernel launch failed: %s", cudaGetErrorString(e));
   if (imp) {                                                     /* what the fused kernel left: the tail behind the last whole tile (or everything) */
      if (masked < nrows) {
         DevCfg dc; premask_devcfg(t, &dc);
         e = launch_peak_masks(dc, t->nrows + masked, t->nrows + nrows, t->stream); ++t->launches;
         if (e != cudaSuccess) return set_err(RT_ERR_CUDA, "mask kernel launch failed: %s", cudaGetErrorString(e)); }
      pm.rows = t->nrows + nrows; pm.fused_rows += masked; }
   CU(cudaEventRecord(e1, t->stream));
   t->ingest_events.push_back(e0); t->ingest_events.push_back(e1);
   t->nrows += nrows; t->valid_known = false;
   if (pm.active && !pm.have_thr && t->nrows >= (1u << 20)) {      /* enough rows to choose the thresholds from: the rest of the tape is fused */
      CU(cudaStreamSynchronize(t->stream));
      e = plan_thresholds_from_rows(t, t->nrows);
      if (e != cudaSuccess) return set_err(RT_ERR_CUDA, "threshold selection failed: %s", cudaGetErrorString(e));
      if (pm.have_thr) {
         DevCfg dc; premask_devcfg(t, &dc);
         e = launch_peak_masks(dc, 0, t->nrows, t->stream); ++t->launches;
         if (e != cudaSuccess) return set_err(RT_ERR_CUDA, "mask kernel launch failed: %s", cudaGetErrorString(e));
         pm.rows = t->nrows; } }
   return RT_OK; }

/* wait for everything enqueued on the tape's stream; fold the ingest kernel times into ms_ingest */
static int tape_drain(rt_tape *t) {
   CU(cudaStreamSynchronize(t->stream));
   for (size_t i = 0; i + 1 < t->ingest_events.size(); i += 2) {
      float ms = 0; cudaEventElapsedTime(&ms, t->ingest_events[i], t->ingest_events[i + 1]); t->ms_ingest += ms;
      cudaEventDestroy(t->ingest_events[i]); cudaEventDestroy(t->ingest_events[i + 1]); }
   t->ingest_events.clear();
   return RT_OK; }

/* two device staging buffers for the chunked host->device copy; *stage_rows = rows per chunk */
static int stage_prepare(rt_tape *t, uint64_t nrows, uint64_t *stage_rows) {
   const uint64_t nh = t->desc.nheads;
   /* 2 Mi rows per chunk (36 MiB for 9 heads); a small upload is cut into ~8 chunks so that reading, copying and ingesting overlap
      and the pinned ring stays small (pinning costs ~0.4 ms per MB) */
   const uint64_t chunk_rows = std::min<uint64_t>((uint64_t)2048 * 1024, std::max<uint64_t>(128 * 1024, (nrows / 8 + 2047) / 2048 * 2048));
   const size_t need = std::max((size_t)std::min(chunk_rows, nrows + 2048) * nh * 2 + 256, (size_t)1 << 20);
   if (t->d_stage[0] && need > t->stage_bytes) {
      /* a small first upload sized the staging buffers: regrow them for this one (a 1.1 G-row tape went through in 57 k-row
         chunks otherwise); the pinned ring of the pageable path follows */
      CU(cudaStreamSynchronize(t->stream)); CU(cudaStreamSynchronize(t->s_copy));
      for (int i = 0; i < RT_STAGE_BUFS; ++i) { cudaFree(t->d_stage[i]); t->d_stage[i] = nullptr; }
      if (t->h_ring) { cudaFreeHost(t->h_ring); t->h_ring = nullptr; } }
   if (!t->d_stage[0]) {
      t->stage_bytes = need;
      if (!t->s_copy) CU(cudaStreamCreateWithFlags(&t->s_copy, cudaStreamNonBlocking));
      { const char *se = getenv("RT_STAGE"); const int want = se ? atoi(se) : 2; t->nstage = want < 2 ? 2 : want > RT_STAGE_BUFS ? RT_STAGE_BUFS : want; }
      for (int i = 0; i < t->nstage; ++i) {
         CU(cudaMalloc(&t->d_stage[i], t->stage_bytes));
         if (!t->stage_done[i]) { CU(cudaEventCreateWithFlags(&t->stage_done[i], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&t->stage_copied[i], cudaEventDisableTiming)); } } }
   *stage_rows = (t->stage_bytes - 256) / (nh * 2) / 2048 * 2048;
   return RT_OK; }

/* one chunk: host->device copy on the copy stream, ingest kernel on the tape's stream; the copy engine never waits for a
   kernel launch (the copy into a staging buffer only waits for the ingest that last read that buffer, two chunks earlier) */
static int enqueue_chunk(rt_tape *t, const int16_t *src, uint64_t n, int buf) {
   CU(cudaStreamWaitEvent(t->s_copy, t->stage_done[buf], 0));
   CU(cudaMemcpyAsync(t->d_stage[buf], src, (size_t)n * t->desc.nheads * 2, cudaMemcpyHostToDevice, t->s_copy));
   CU(cudaEventRecord(t->stage_copied[buf], t->s_copy));
   CU(cudaStreamWaitEvent(t->stream, t->stage_copied[buf], 0));
   int rc = tape_ingest(t, t->d_stage[buf], n);
   if (rc) return rc;
   CU(cudaEventRecord(t->stage_done[buf], t->stream));
   return RT_OK; }

/* Upload from PAGEABLE host memory (a file mapping, a malloc'd buffer): cudaMemcpyAsync would stage such a copy through the driver's
   own bounce buffer with one thread (~10 GB/s).  Instead a few threads copy chunks into a ring of pinned buffers while the copy
   engine drains them, which keeps PCIe busy (RT_UPLOAD_THREADS, default 4). */
static int upload_pageable(rt_tape *t, const int16_t *rows, int fd, uint64_t fd_offset, uint64_t nrows, uint64_t stage_rows) {
   const uint64_t nh = t->desc.nheads;
   const size_t slot_bytes = (size_t)stage_rows * nh * 2;
   const uint64_t nchunks = (nrows + stage_rows - 1) / stage_rows;
   const int NB = (int)std::min<uint64_t>(RT_RING_SLOTS, std::max<uint64_t>(3, nchunks));   /* >= 3: the loop below keeps NB - 2 copies in flight */
   if (t->h_ring && t->ring_slots < NB) {                          /* a later, larger upload: grow the ring */
      CU(cudaStreamSynchronize(t->stream)); if (t->s_copy) CU(cudaStreamSynchronize(t->s_copy));
      cudaFreeHost(t->h_ring); t->h_ring = nullptr; }
   if (!t->h_ring) {
      CU(cudaHostAlloc(&t->h_ring, slot_bytes * NB, cudaHostAllocDefault));
      t->ring_slots = NB;
      for (int i = 0; i < NB; ++i) if (!t->ring_done[i]) CU(cudaEventCreateWithFlags(&t->ring_done[i], cudaEventDisableTiming)); }
   const char *env = getenv("RT_UPLOAD_THREADS");
   const int hw = (int)std::thread::hardware_concurrency();
   int nthreads = env && atoi(env) > 0 ? atoi(env) : std::max(4, std::min(RT_RING_SLOTS - 2, hw - 2));
   std::atomic<int> io_error{0};
   nthreads = std::max(1, std::min<int>(nthreads, (int)std::min<uint64_t>(nchunks, 16)));
   std::mutex mu; std::condition_variable cv;
   std::vector<char> filled(nchunks, 0);
   uint64_t released = 0;                                         /* chunks whose ring slot may be overwritten again */
   std::atomic<uint64_t> next{0};
   auto producer = [&]() {
      for (;;) {
         const uint64_t i = next.fetch_add(1);
         if (i >= nchunks) return;
         { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return i < released + NB; }); }
         const uint64_t r0 = i * stage_rows, n = std::min(stage_rows, nrows - r0);
         char *dst = reinterpret_cast<char *>(t->h_ring) + (size_t)(i % NB) * slot_bytes;
         if (rows) memcpy(dst, rows + r0 * nh, (size_t)n * nh * 2);
         else {                                                   /* straight from the file (page cache) into pinned memory */
            size_t got = 0; const size_t want = (size_t)n * nh * 2;
            while (got < want) {
               const ssize_t k = pread(fd, dst + got, want - got, (off_t)(fd_offset + r0 * nh * 2 + got));
               if (k <= 0) { io_error = 1; memset(dst + got, 0, want - got); break; }
               got += (size_t)k; } }
         { std::lock_guard<std::mutex> lk(mu); filled[i] = 1; }
         cv.notify_all(); } };
   std::vector<std::thread> th;
   for (int k = 0; k < nthreads; ++k) th.emplace_back(producer);
   int rc = RT_OK;
   for (uint64_t i = 0; i < nchunks; ++i) {
      { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return filled[i] != 0; }); }
      const uint64_t r0 = i * stage_rows, n = std::min(stage_rows, nrows - r0);
      if (rc == RT_OK) {
         rc = enqueue_chunk(t, reinterpret_cast<const int16_t *>(reinterpret_cast<char *>(t->h_ring) + (size_t)(i % NB) * slot_bytes), n, (int)(i % (uint64_t)t->nstage));
         if (rc == RT_OK && cudaEventRecord(t->ring_done[i % NB], t->s_copy) != cudaSuccess) rc = set_err(RT_ERR_CUDA, "cudaEventRecord failed"); }
      if (i + 2 >= (uint64_t)NB) {                                /* keep NB - 2 copies in flight, then free the oldest slot */
         const uint64_t k = i + 2 - NB;
         if (rc == RT_OK) cudaEventSynchronize(t->ring_done[k % NB]);
         { std::lock_guard<std::mutex> lk(mu); released = k + 1; }
         cv.notify_all(); } }
   { std::lock_guard<std::mutex> lk(mu); released = nchunks; }
   cv.notify_all();
   for (auto &x : th) x.join();
   if (rc == RT_OK && io_error) rc = set_err(RT_ERR_ARG, "rt_upload_fd: short read");
   return rc; }

extern "C" int rt_upload(rt_tape *t, const int16_t *rows, uint64_t nrows) {
   if (!t || (!rows && nrows)) return set_err(RT_ERR_ARG, "rt_upload: null argument");
   if (nrows == 0) return RT_OK;
   CU(cudaSetDevice(t->device));
   int rc = tape_reserve(t, t->nrows + nrows);
   if (rc) return rc;
   const uint64_t nh = t->desc.nheads;
   uint64_t stage_rows = 0;
   rc = stage_prepare(t, nrows, &stage_rows); if (rc) return rc;
   if (nrows >= 4 * stage_rows && stage_rows) {                   /* a large upload from pageable memory: staged by our own threads */
      cudaPointerAttributes pa{};
      const bool pageable = cudaPointerGetAttributes(&pa, rows) != cudaSuccess || pa.type == cudaMemoryTypeUnregistered;
      cudaGetLastError();
      const char *env = getenv("RT_UPLOAD_THREADS");
      const char *reg = getenv("RT_UPLOAD_REGISTER");
      if (pageable && reg && atoi(reg)) {
         /* experiment (RT_UPLOAD_REGISTER=1): pin the caller's pages in place (e.g. a mapping of a file in the page cache) and let
            the copy engine read them directly -- no CPU copy at all */
         const uintptr_t pg = 4096, lo = reinterpret_cast<uintptr_t>(rows) / pg * pg;
         const uintptr_t hi = (reinterpret_cast<uintptr_t>(rows) + nrows * nh * 2 + pg - 1) / pg * pg;
         auto w0 = std::chrono::steady_clock::now();
         if (cudaHostRegister(reinterpret_cast<void *>(lo), hi - lo, cudaHostRegisterReadOnly) == cudaSuccess) {
            auto w1 = std::chrono::steady_clock::now();
            uint64_t done = 0; int buf = 0;
            while (done < nrows) {
               const uint64_t n = std::min(stage_rows, nrows - done);
               rc = enqueue_chunk(t, rows + done * nh, n, buf);
               if (rc) { cudaDeviceSynchronize(); return rc; }
               done += n; buf = (buf + 1) % t->nstage; }
            rc = tape_drain(t);
            auto w2 = std::chrono::steady_clock::now();
            cudaHostUnregister(reinterpret_cast<void *>(lo));
            auto w3 = std::chrono::steady_clock::now();
            if (getenv("RT_TRACE")) fprintf(stderr, "[rt_upload] register %.3f s, copy+ingest %.3f s, unregister %.3f s\n", std::chrono::duration<double>(w1 - w0).count(),
                                            std::chrono::duration<double>(w2 - w1).count(), std::chrono::duration<double>(w3 - w2).count());
            t->h2d_bytes += nrows * nh * 2;
            return rc; }
         cudaGetLastError();
         if (getenv("RT_TRACE")) fprintf(stderr, "[rt_upload] cudaHostRegister of the caller's memory failed: staging instead\n"); }
      if (pageable && !(env && atoi(env) == 0)) {
         rc = upload_pageable(t, rows, -1, 0, nrows, stage_rows);
         if (rc) { cudaDeviceSynchronize(); return rc; }
         t->h2d_bytes += nrows * nh * 2;
         return tape_drain(t); } }
   if (t->nrows % 2048 != 0 && !t->force_simple_ingest) { /* appended onto a partial tile: fine, the plain kernel handles it */ }
   /* copies on the copy stream, ingest kernels on the tape's stream, two staging buffers; one synchronisation at the end */
   uint64_t done = 0; int buf = 0;
   while (done < nrows) {
      uint64_t n = std::min(stage_rows ? stage_rows : nrows - done, nrows - done);
      rc = enqueue_chunk(t, rows + done * nh, n, buf);
      if (rc) return rc;
      done += n; buf = (buf + 1) % t->nstage; }
   t->h2d_bytes += nrows * nh * 2;
   return tape_drain(t); }

extern "C" int rt_upload_fd(rt_tape *t, int fd, uint64_t offset, uint64_t nrows) {
   if (!t || fd < 0) return set_err(RT_ERR_ARG, "rt_upload_fd: bad argument");
   if (nrows == 0) return RT_OK;
   CU(cudaSetDevice(t->device));
   int rc = tape_reserve(t, t->nrows + nrows);
   if (rc) return rc;
   uint64_t stage_rows = 0;
   rc = stage_prepare(t, nrows, &stage_rows); if (rc) return rc;
   rc = upload_pageable(t, nullptr, fd, offset, nrows, stage_rows);
   if (rc) { cudaDeviceSynchronize(); return rc; }
   t->h2d_bytes += nrows * t->desc.nheads * 2;
   return tape_drain(t); }

extern "C" int rt_attach_device(rt_tape *t, const void *rows_dev, uint64_t nrows) {
   if (!t || (!rows_dev && nrows)) return set_err(RT_ERR_ARG, "rt_attach_device: null argument");
   if (nrows == 0) return RT_OK;
   CU(cudaSetDevice(t->device));
   int rc = tape_reserve(t, t->nrows + nrows);
   if (rc) return rc;
   const int16_t *src = static_cast<const int16_t *>(rows_dev);
   if (t->pm.active && !t->pm.have_thr && t->nrows == 0 && nrows > (8u << 20)) {   /* thresholds from the first 4 Mi rows, the rest fused */
      const uint64_t head = 4u << 20;
      rc = tape_ingest(t, src, head); if (rc) return rc;
      src += head * t->desc.nheads; nrows -= head; }
   rc = tape_ingest(t, src, nrows);
   if (rc) return rc;
   return tape_drain(t); }

extern "C" int rt_clear(rt_tape *t) {
   if (!t) return set_err(RT_ERR_ARG, "rt_clear: null");
   CU(cudaSetDevice(t->device));
   int rc = tape_drain(t); if (rc) return rc;
   unsigned long long none = ~0ull;
   CU(cudaMemcpy(t->d_first_end, &none, sizeof none, cudaMemcpyHostToDevice));
   t->nrows = 0; t->nrows_valid = 0; t->valid_known = false; t->ms_ingest = 0; t->h2d_bytes = 0;
   t->pm.rows = 0; t->pm.fused_rows = 0; t->inv_rows = 0; t->launches_at_clear = t->launches;
   return RT_OK; }

static int tape_sync_valid(rt_tape *t) {
   if (t->valid_known) return RT_OK;
   unsigned long long fe = ~0ull;
   CU(cudaSetDevice(t->device));
   { int rc = tape_drain(t); if (rc) return rc; }
   CU(cudaMemcpy(&fe, t->d_first_end, sizeof fe, cudaMemcpyDeviceToHost));
   t->nrows_valid = std::min<uint64_t>(t->nrows, fe);
   t->valid_known = true;
   return RT_OK; }

extern "C" uint64_t rt_nrows(const rt_tape *t) {
   if (!t) return 0;
   tape_sync_valid(const_cast<rt_tape *>(t));
   return t->nrows_valid; }

extern "C" void rt_close(rt_tape *t) {
   if (!t) return;
   cudaSetDevice(t->device);
   if (t->stream) cudaStreamSynchronize(t->stream);
   cudaFree(t->planes); cudaFree(t->gmm); cudaFree(t->d_first_end); cudaFree(t->planes_inv); cudaFree(t->gmm_inv);
   cudaFree(t->pool_cache); cudaFree(t->next_cache); if (t->pin_cache) cudaFreeHost(t->pin_cache);
   cudaFree(t->rec_cache); cudaFree(t->pm.mc); cudaFree(t->pm.md); cudaFree(t->pm.ma);
   for (int i = 0; i < RT_STAGE_BUFS; ++i) { cudaFree(t->d_stage[i]); if (t->stage_done[i]) cudaEventDestroy(t->stage_done[i]); if (t->stage_copied[i]) cudaEventDestroy(t->stage_copied[i]); }
   if (t->h_ring) { cudaFreeHost(t->h_ring); for (auto e : t->ring_done) if (e) cudaEventDestroy(e); }
   if (t->s_copy) cudaStreamDestroy(t->s_copy);
   if (t->s_mask) cudaStreamDestroy(t->s_mask);
   for (auto e : t->mask_events) cudaEventDestroy(e);
   if (t->stream) cudaStreamDestroy(t->stream);
   for (auto st : t->s_par) cudaStreamDestroy(st);
   if (t->s_scan) cudaStreamDestroy(t->s_scan);
   if (t->s_out) cudaStreamDestroy(t->s_out);
   delete t; }

extern "C" void *rt_host_alloc(size_t bytes) {
   void *p = nullptr;
   if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { set_err(RT_ERR_NOMEM, "cudaHostAlloc(%zu) failed", bytes); return nullptr; }
   return p; }
extern "C" void rt_host_free(void *p) { if (p) cudaFreeHost(p); }

/* ---- configuration -> device form ------------------------------------------------------------------ */
static int cfg_check(const rt_tape *t, const rt_scan_cfg *cfg) {
   if (cfg->mode != RT_MODE_PE && cfg->mode != RT_MODE_NRZI && cfg->mode != RT_MODE_GCR && cfg->mode != RT_MODE_WW)
      return set_err(RT_ERR_ARG, "bad mode %d", cfg->mode);
   if ((cfg->flags & RT_F_FIND_ZEROS) && cfg->mode == RT_MODE_PE)
      return set_err(RT_ERR_UNSUPPORTED, "-zeros with PE needs the PE bit clock in the scan; not supported");
   if (!(cfg->flags & RT_F_DENSITY_DETECT) && !(cfg->bpi > 0 && cfg->ips > 0))
      return set_err(RT_ERR_ARG, "bpi/ips must be positive");
   for (uint32_t k = 0; k < t->desc.ntrks; ++k)
      if (cfg->skew_delaycnt[k] < 0 || cfg->skew_delaycnt[k] > RT_MAXSKEWSAMP)
         return set_err(RT_ERR_ARG, "bad skew delay for track %u", k);
   if (cfg->parms.agc_window < 0 || cfg->parms.agc_window > RT_AGC_MAX_WINDOW || cfg->parms.clk_window < 0 || cfg->parms.clk_window > RT_CLKRATE_WINDOW)
      return set_err(RT_ERR_ARG, "agc_window/clk_window out of range");
   if (!(cfg->flags & RT_F_FIND_ZEROS) && rt_pkww_width(cfg, t->desc.tdelta_ns) < 3)
      return set_err(RT_ERR_UNSUPPORTED, "peak window narrower than 3 samples (%d)", rt_pkww_width(cfg, t->desc.tdelta_ns));
   return RT_OK; }

static void cfg_to_dev(const rt_tape *t, const rt_scan_cfg *cfg, DevCfg *d) {
   rtcfg::to_dev(t->desc, t->planes, t->plane_stride, t->nrows_valid, cfg, d);
   const char *skip = getenv("RT_FAST_SKIP");                       /* RT_FAST_SKIP=0: the fast path walks every row (tests) */
   if (!(skip && skip[0] == '0')) { d->gmm = reinterpret_cast<const uint32_t *>(t->gmm); d->ngran_cap = t->ngran_cap; } }

/* k-way merge of per-track event streams into (row, trk) order */
struct Cursor { const rt_event *p, *end; };

/* ---- stateful exact scan ----------------------------------------------------------------------- */
struct rt_scan {
   rt_tape *tape = nullptr; rt_scan_cfg cfg{}; DevCfg dc{};
   TrkState *d_st = nullptr, *d_ckst = nullptr; SkewState *d_sk = nullptr, *d_cksk = nullptr;
   uint64_t pos = 0, ckpt_pos = 0, ckpt_end = 0; bool positioned = false, have_ckpt = false;
   rt_event *d_ev = nullptr; uint32_t cap = 0; uint32_t *d_counts = nullptr, *d_failed = nullptr;
   rt_event *h_ev = nullptr; size_t h_ev_cap = 0;       /* pinned staging */
   std::vector<rt_event> merged;
   /* skip-ahead (scan_generic.cuh): candidate / canonical bit planes of the whole tape for this configuration's window width */
   uint32_t *d_mc = nullptr, *d_md = nullptr, *d_ma = nullptr; uint64_t mask_rows = 0, mask_stride = 0; int mask_width = 0;
};

static int scan_alloc_events(rt_scan *s, uint32_t cap) {
   const uint32_t nt = s->tape->desc.ntrks;
   if (s->d_ev) cudaFree(s->d_ev);
   if (s->h_ev) cudaFreeHost(s->h_ev);
   s->d_ev = nullptr; s->h_ev = nullptr;
   CU(cudaMalloc(&s->d_ev, (size_t)cap * nt * sizeof(rt_event)));
   CU(cudaHostAlloc(&s->h_ev, (size_t)cap * nt * sizeof(rt_event), cudaHostAllocDefault));
   s->cap = cap; s->h_ev_cap = (size_t)cap * nt;
   return RT_OK; }

extern "C" int rt_scan_begin(rt_tape *t, const rt_scan_cfg *cfg, rt_scan **out) {
   if (!t || !cfg || !out) return set_err(RT_ERR_ARG, "rt_scan_begin: null argument");
   int rc = cfg_check(t, cfg); if (rc) return rc;
   rc = tape_sync_valid(t); if (rc) return rc;
   CU(cudaSetDevice(t->device));
   rt_scan *s = new (std::nothrow) rt_scan();
   if (!s) return set_err(RT_ERR_NOMEM, "rt_scan_begin: out of memory");
   s->tape = t; s->cfg = *cfg; cfg_to_dev(t, cfg, &s->dc);
   const uint32_t nt = t->desc.ntrks;
   cudaError_t e;
   if ((e = cudaMalloc(&s->d_st, nt * sizeof(TrkState))) != cudaSuccess || (e = cudaMalloc(&s->d_ckst, nt * sizeof(TrkState))) != cudaSuccess ||
       (e = cudaMalloc(&s->d_sk, nt * sizeof(SkewState))) != cudaSuccess || (e = cudaMalloc(&s->d_cksk, nt * sizeof(SkewState))) != cudaSuccess ||
       (e = cudaMalloc(&s->d_counts, (nt + 1) * sizeof(uint32_t))) != cudaSuccess) {
      rt_scan_end(s); return set_err(RT_ERR_CUDA, "rt_scan_begin: cudaMalloc failed: %s", cudaGetErrorString(e)); }
   s->d_failed = s->d_counts + nt;
   cudaMemsetAsync(s->d_st, 0, nt * sizeof(TrkState), t->stream);
   cudaMemsetAsync(s->d_sk, 0, nt * sizeof(SkewState), t->stream);
   rc = scan_alloc_events(s, 16384);
   if (rc) { rt_scan_end(s); return rc; }
   *out = s; return RT_OK; }

extern "C" int rt_scan_set_cfg(rt_scan *s, const rt_scan_cfg *cfg) {
   if (!s || !cfg) return set_err(RT_ERR_ARG, "rt_scan_set_cfg: null argument");
   if (cfg->mode != s->cfg.mode) return set_err(RT_ERR_ARG, "rt_scan_set_cfg: the mode cannot change");
   int rc = cfg_check(s->tape, cfg); if (rc) return rc;
   s->cfg = *cfg; cfg_to_dev(s->tape, cfg, &s->dc); s->have_ckpt = false;
   if (s->positioned) {                                          /* e.g. new skew delays: the deskew FIFO refills before rows are skipped again */
      CU(cudaSetDevice(s->tape->device));
      launch_ctx_reset(s->dc, s->d_st, s->d_sk, RT_RESET_NONE, s->pos, 0, s->tape->stream);
      CU(cudaGetLastError()); ++s->tape->launches; }
   return RT_OK; }

extern "C" int rt_scan_reset(rt_scan *s, int kind, uint64_t row) {
   if (!s) return set_err(RT_ERR_ARG, "rt_scan_reset: null");
   if (kind < RT_RESET_NONE || kind > RT_RESET_PEAKSTATE) return set_err(RT_ERR_ARG, "rt_scan_reset: bad kind %d", kind);
   rt_tape *t = s->tape;
   CU(cudaSetDevice(t->device));
   int rc = tape_sync_valid(t); if (rc) return rc;
   s->dc.planes = t->planes; s->dc.plane_stride = t->plane_stride; s->dc.nrows = t->nrows_valid;
   if (kind != RT_RESET_NONE || (s->positioned && row != s->pos)) {
      int tz = rt_row_time(&t->desc, row) == 0.0;
      launch_ctx_reset(s->dc, s->d_st, s->d_sk, kind, row, tz, t->stream);
      CU(cudaGetLastError()); ++t->launches; }
   s->pos = row; s->positioned = true; s->have_ckpt = false;
   return RT_OK; }

/* the bit planes the exact scan's skip-ahead reads: built once per (tape contents, window width), thresholds = a quarter of the
   default-state bound (the scan only trusts them while its own bound is above that; never a result depends on them) */
static int scan_prepare_masks(rt_scan *s) {
   rt_tape *t = s->tape; DevCfg &dc = s->dc;
   const char *env = getenv("RT_EXACT_SKIP");
   const bool want = !(env && env[0] == '0') && dc.det == RT_DET_PEAK && !dc.invert && !dc.differentiate && dc.width >= 3
                     && dc.width <= RT_PKWW_MAX_WIDTH && peak_mask_T0(dc, 0.25f) > 0 && t->nrows_valid > 4096;
   dc.m_cand = dc.m_cand2 = dc.m_acan = nullptr; dc.mask_stride = 0;
   for (int k = 0; k < RT_MAXTRKS; ++k) dc.T0[k] = dc.T1[k] = 0;
   if (!want) return RT_OK;
   const uint32_t nt = t->desc.ntrks;
   const uint64_t ms = peak_mask_stride(t->plane_stride);
   if (ms != s->mask_stride) {
      cudaFree(s->d_mc); cudaFree(s->d_md); cudaFree(s->d_ma); s->d_mc = s->d_md = s->d_ma = nullptr; s->mask_stride = 0; s->mask_rows = 0;
      CU(cudaMalloc(&s->d_mc, (size_t)ms * nt * 4)); CU(cudaMalloc(&s->d_md, (size_t)ms * nt * 4)); CU(cudaMalloc(&s->d_ma, (size_t)ms * nt * 4));
      s->mask_stride = ms; }
   const int T0 = peak_mask_T0(dc, 0.25f);
   for (uint32_t k = 0; k < nt; ++k) dc.T0[k] = T0;
   dc.m_cand = s->d_mc; dc.m_cand2 = s->d_md; dc.m_acan = s->d_ma; dc.mask_stride = ms;
   if (s->mask_rows != t->nrows_valid || s->mask_width != dc.width) {
      cudaError_t e = launch_peak_masks(dc, 0, t->nrows_valid, t->stream); ++t->launches;
      if (e != cudaSuccess) return set_err(RT_ERR_CUDA, "mask kernel launch failed: %s", cudaGetErrorString(e));
      s->mask_rows = t->nrows_valid; s->mask_width = dc.width; }
   return RT_OK; }

static int scan_span(rt_scan *s, uint64_t from, uint64_t to, bool want_events) {
   rt_tape *t = s->tape; const uint32_t nt = t->desc.ntrks;
   static const bool trace = getenv("RT_TRACE") != nullptr;
   const auto w0 = std::chrono::steady_clock::now();
   struct Lap { const std::chrono::steady_clock::time_point w0; uint64_t from, to; bool ev, on; ~Lap() { if (on) fprintf(stderr, "[scan_span] rows [%llu, %llu) = %llu, %s: %.3f ms\n",
      (unsigned long long)from, (unsigned long long)to, (unsigned long long)(to - from), ev ? "events" : "rewind", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count()); } } lap{w0, from, to, want_events, trace};
   { int rc = scan_prepare_masks(s); if (rc) return rc; }
   for (;;) {
      uint32_t zero = 0;
      CU(cudaMemcpyAsync(s->d_failed, &zero, sizeof zero, cudaMemcpyHostToDevice, t->stream));
      launch_ctx_scan(s->dc, s->d_st, s->d_sk, from, to, s->d_ev, want_events ? s->cap : 0, s->d_counts, s->d_failed, t->stream);
      CU(cudaGetLastError()); ++t->launches;
      uint32_t counts[RT_MAXTRKS + 1];
      CU(cudaMemcpyAsync(counts, s->d_counts, (nt + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, t->stream));
      CU(cudaStreamSynchronize(t->stream));
      if (counts[nt] == 2) return set_err(RT_ERR_STATE, "peak not found in window: the reference would call fatal() here");
      if (!want_events) return RT_OK;
      uint32_t mx = 0; for (uint32_t k = 0; k < nt; ++k) mx = std::max(mx, counts[k]);
      if (mx > s->cap) {           /* event buffers too small: regrow, restore the entry state, rerun */
         int rc = scan_alloc_events(s, mx + mx / 4 + 1024); if (rc) return rc;
         CU(cudaMemcpyAsync(s->d_st, s->d_ckst, nt * sizeof(TrkState), cudaMemcpyDeviceToDevice, t->stream));
         CU(cudaMemcpyAsync(s->d_sk, s->d_cksk, nt * sizeof(SkewState), cudaMemcpyDeviceToDevice, t->stream));
         continue; }
      /* fetch and merge */
      size_t total = 0;
      for (uint32_t k = 0; k < nt; ++k) {
         if (counts[k]) CU(cudaMemcpyAsync(s->h_ev + (size_t)k * s->cap, s->d_ev + (size_t)k * s->cap, (size_t)counts[k] * sizeof(rt_event), cudaMemcpyDeviceToHost, t->stream));
         total += counts[k]; }
      CU(cudaStreamSynchronize(t->stream));
      s->merged.clear(); s->merged.reserve(total);
      Cursor cur[RT_MAXTRKS];
      for (uint32_t k = 0; k < nt; ++k) { cur[k].p = s->h_ev + (size_t)k * s->cap; cur[k].end = cur[k].p + counts[k]; }
      for (size_t n = 0; n < total; ++n) {
         int best = -1; uint64_t brow = ~0ull;
         for (uint32_t k = 0; k < nt; ++k) if (cur[k].p < cur[k].end && cur[k].p->row < brow) { brow = cur[k].p->row; best = (int)k; }
         s->merged.push_back(*cur[best].p++); }
      return RT_OK; } }

extern "C" int rt_scan_run(rt_scan *s, uint64_t nrows, const rt_event **events, uint64_t *nevents, uint64_t *rows_done) {
   if (!s) return set_err(RT_ERR_ARG, "rt_scan_run: null");
   if (!s->positioned) return set_err(RT_ERR_STATE, "rt_scan_run before rt_scan_reset");
   rt_tape *t = s->tape; const uint32_t nt = t->desc.ntrks;
   CU(cudaSetDevice(t->device));
   uint64_t end = s->pos + nrows;
   if (end > t->nrows_valid || end < s->pos) end = t->nrows_valid;
   if (end < s->pos) end = s->pos;
   CU(cudaMemcpyAsync(s->d_ckst, s->d_st, nt * sizeof(TrkState), cudaMemcpyDeviceToDevice, t->stream));
   CU(cudaMemcpyAsync(s->d_cksk, s->d_sk, nt * sizeof(SkewState), cudaMemcpyDeviceToDevice, t->stream));
   s->ckpt_pos = s->pos; s->ckpt_end = end; s->have_ckpt = true;
   s->merged.clear();
   if (end > s->pos) { int rc = scan_span(s, s->pos, end, true); if (rc) return rc; }
   if (rows_done) *rows_done = end - s->pos;
   s->pos = end;
   if (events) *events = s->merged.data();
   if (nevents) *nevents = s->merged.size();
   return RT_OK; }

extern "C" int rt_scan_rewind(rt_scan *s, uint64_t row) {
   if (!s) return set_err(RT_ERR_ARG, "rt_scan_rewind: null");
   if (!s->have_ckpt || row < s->ckpt_pos || row > s->ckpt_end) return set_err(RT_ERR_STATE, "rt_scan_rewind: row outside the last scanned span");
   rt_tape *t = s->tape; const uint32_t nt = t->desc.ntrks;
   CU(cudaSetDevice(t->device));
   CU(cudaMemcpyAsync(s->d_st, s->d_ckst, nt * sizeof(TrkState), cudaMemcpyDeviceToDevice, t->stream));
   CU(cudaMemcpyAsync(s->d_sk, s->d_cksk, nt * sizeof(SkewState), cudaMemcpyDeviceToDevice, t->stream));
   if (row > s->ckpt_pos) { int rc = scan_span(s, s->ckpt_pos, row, false); if (rc) return rc; }
   s->pos = row; s->merged.clear();
   return RT_OK; }

extern "C" int rt_scan_set_avg_height(rt_scan *s, uint32_t trk, float v) {
   if (!s || trk >= s->tape->desc.ntrks) return set_err(RT_ERR_ARG, "rt_scan_set_avg_height: bad argument");
   CU(cudaSetDevice(s->tape->device));
   launch_ctx_set_avg_height(s->d_st, (int)trk, v, s->tape->stream);
   CU(cudaGetLastError()); ++s->tape->launches;
   return RT_OK; }

extern "C" uint64_t rt_scan_pos(const rt_scan *s) { return s ? s->pos : 0; }

extern "C" void rt_scan_end(rt_scan *s) {
   if (!s) return;
   cudaSetDevice(s->tape->device);
   cudaStreamSynchronize(s->tape->stream);
   cudaFree(s->d_st); cudaFree(s->d_ckst); cudaFree(s->d_sk); cudaFree(s->d_cksk); cudaFree(s->d_counts); cudaFree(s->d_ev);
   cudaFree(s->d_mc); cudaFree(s->d_md); cudaFree(s->d_ma);
   if (s->h_ev) cudaFreeHost(s->h_ev);
   delete s; }

/* ---- speculative whole-tape scan ---------------------------------------------------------------- */
struct BulkCfg {
   rt_scan_cfg cfg{}; DevCfg dc{};
   UnitDesc *d_units = nullptr; TrkMeta *d_meta = nullptr;       /* device-resident results of the scan */
   uint32_t nunits = 0;
   bool fast = false;
   size_t last_unit = ~(size_t)0;                                 /* rt_bulk_last_unit() */
   std::vector<UnitDesc> units; std::vector<TrkMeta> meta;        /* host copies, filled by rt_bulk_fetch() */
};
struct rt_bulk {
   rt_tape *tape = nullptr; std::vector<BulkCfg> cfgs; rt_bulk_stats stats{};
   bool fetched = false;
   /* ONE event pool for all configurations of a scan: their kernels run concurrently and take chunks from the same cursor */
   rt_event *d_pool = nullptr; uint32_t *d_chunk_next = nullptr; uint32_t pool_chunks = 0, chunks_used = 0;
   bool pool_from_cache = false, pin_from_cache = false;
   rt_event *h_pool = nullptr; size_t h_pool_events = 0;          /* pinned */
   rt_event *user_pool = nullptr; size_t user_pool_bytes = 0;     /* rt_bulk_fetch_to(): the caller's buffer */
   size_t h_pool_shared_bytes = 0;                                /* != 0: h_pool is an anonymous MAP_SHARED mapping of this size (RT_OPT_SHARED_RESULTS) */
   std::vector<uint32_t> chunk_next;
   std::vector<rt_event> result;
   rt_scan *bridge = nullptr; uint32_t bridge_cfg = ~0u;          /* exact context for bridge scans (rt_bulk_lookup) */
   uint64_t n_bridged = 0;
};

using rtlookup::fill_of;                                      /* lookup_rules.h */

/* Everything derived from one configuration that the unit finder and the scan kernels need. */
struct ScanPlan { DevCfg dc; UnitParams up; float quiet_thr; int quiet_thr_lsb; bool use_fast; bool use_sparse; bool t0_auto; };
/* -invert with the fast kernels: negated planes (rows [0, nrows)), made once per tape and extended as it grows */
static int tape_ensure_inverted(rt_tape *t, uint64_t nrows) {
   if (t->inv_rows >= nrows) return RT_OK;
   const uint32_t nt = t->desc.ntrks;
   if (!t->planes_inv) {
      CU(cudaMalloc(&t->planes_inv, (size_t)t->plane_stride * nt * sizeof(int16_t)));
      cudaError_t e = cudaMalloc(&t->gmm_inv, (size_t)t->ngran_cap * nt * 4);
      if (e != cudaSuccess) { cudaFree(t->planes_inv); t->planes_inv = nullptr; return set_err(RT_ERR_CUDA, "cudaMalloc(inverted granule map) failed: %s", cudaGetErrorString(e)); }
      t->inv_rows = 0; }
   CU(launch_negate(t->planes, t->planes_inv, t->plane_stride, t->gmm, t->gmm_inv, t->ngran_cap, t->inv_rows, nrows, (int)nt, t->stream));
   t->launches += 2; t->inv_rows = nrows;
   return RT_OK; }

static void make_plan(const rt_tape *t, const rt_scan_cfg *cfg, ScanPlan *pl, bool use_inverted = false) {
   cfg_to_dev(t, cfg, &pl->dc);
   if (use_inverted && pl->dc.invert && !pl->dc.differentiate && t->planes_inv) {   /* scan the negated planes, flag cleared (k_units.cu) */
      pl->dc.planes = t->planes_inv; pl->dc.invert = 0;
      if (pl->dc.gmm) pl->dc.gmm = reinterpret_cast<const uint32_t *>(t->gmm_inv); }
   const DevCfg &dc = pl->dc;
   /* proposal thresholds (heuristic) and the exact quiet threshold the scan kernel applies */
   const double lsb = (double)t->desc.maxvolts / 32767.0;
   /* density detection (bpi == 0): the bit length is what is being looked for; assume a long one (200 BPI at 50 IPS): fewer and
      longer units, never a cut inside a block */
   const double rows_per_bit = dc.bpi > 0 && dc.ips > 0 ? 1.0 / ((double)dc.bpi * dc.ips * dc.sample_deltat) : 1.0 / (200.0 * 50.0 * dc.sample_deltat);
   pl->up = UnitParams{};
   pl->up.det = dc.det;
   pl->quiet_thr = 0;
   pl->quiet_thr_lsb = rtcfg::quiet_thr_lsb(dc);
   /* K3b (int16 fast path) for the peak detector of NRZI / PE; RT_SCAN=generic forces the exact generic kernel (tests) */
   const char *force = getenv("RT_SCAN");
   pl->use_fast = fast_scan_eligible(dc) && !(force && strcmp(force, "generic") == 0);
   /* K3c (two passes: candidate masks + sparse scan) where the peak-detector fast path applies; RT_SCAN=fast keeps the one-pass
      kernel K3b (tests).  The mask threshold T0 is chosen per track from the data (plan_auto_t0) unless RT_SPARSE_T0 fixes it as a
      fraction of the default-state bound (tests). */
   const char *t0env = getenv("RT_SPARSE_T0");
   for (int k = 0; k < RT_MAXTRKS; ++k) pl->dc.T0[k] = pl->dc.T1[k] = 0;
   pl->dc.m_cand = pl->dc.m_cand2 = pl->dc.m_acan = nullptr; pl->dc.mask_stride = 0;
   pl->use_sparse = pl->t0_auto = false;
   if (pl->use_fast && dc.det == RT_DET_PEAK && !(force && strcmp(force, "fast") == 0)) {
      if (t0env) {
         const int T0 = peak_mask_T0(dc, (float)atof(t0env));
         for (int k = 0; k < dc.ntrks; ++k) { pl->dc.T0[k] = T0; pl->dc.T1[k] = T0 > 0 && T0 * 8 / 5 <= 65535 ? T0 * 8 / 5 : 0; }
         pl->use_sparse = T0 > 0; }
      else { pl->use_sparse = peak_mask_T0(dc, 1.0f) > 0; pl->t0_auto = pl->use_sparse; } }
   if (dc.det == RT_DET_PEAK) {
      pl->quiet_thr = dc.p.pkww_rise * 0.999f;
      if (dc.p.pkww_rise < 1e-3f) pl->quiet_thr = 0;                  /* nothing can be proven quiet: every lookup misses */
      pl->up.thr = (int)(0.75 * dc.p.pkww_rise / lsb); }
   else if (dc.det == RT_DET_ZC) pl->up.thr = (int)(0.9 * RT_ZEROCROSS_PEAK / lsb);
   else pl->up.thr = (int)(0.9 * std::max(0.05, 0.5 / std::max(1, dc.samples_per_bit)) / lsb);
   uint64_t gap_rows = (uint64_t)(6.0 * rows_per_bit) + 1;
   pl->up.min_gap_gran = (uint32_t)std::max<uint64_t>(2, (gap_rows + RT_GRAN - 1) / RT_GRAN + 1);
   const uint64_t ibg_rows = (uint64_t)(200e-6 / dc.sample_deltat) + 1;      /* *_IBG_SECS, decoder.h:105,113,116 */
   pl->up.tail_rows = (uint64_t)(16.0 * rows_per_bit) + ibg_rows + 64 + RT_PKWW_MAX_WIDTH + RT_MAXSKEWSAMP; }

static cudaError_t launch_scan(const rt_tape *t, const ScanPlan &pl, const UnitDesc *d_units, uint32_t nunits, TrkMeta *d_meta, rt_event *d_pool,
                               uint32_t *d_chunk_next, unsigned int *d_cursor, uint32_t pool_chunks, unsigned long long *d_counters,
                               int max_ctas_per_sm, cudaStream_t st) {
   if (pl.use_sparse) return launch_units_sparse(pl.dc, d_units, nunits, d_meta, d_pool, d_chunk_next, d_cursor, pool_chunks, pl.quiet_thr_lsb, d_counters,
                                                 t->sms, max_ctas_per_sm, st);
   if (pl.use_fast) return launch_units_fast(pl.dc, d_units, nunits, d_meta, d_pool, d_chunk_next, d_cursor, pool_chunks, pl.quiet_thr_lsb, d_counters,
                                             t->sms, max_ctas_per_sm, st);
   const uint64_t threads = (uint64_t)nunits * t->desc.ntrks;
   int grid = (int)std::min<uint64_t>((threads + 127) / 128, (uint64_t)t->sms * (max_ctas_per_sm > 0 ? 4 : 16));
   launch_units_scan(pl.dc, d_units, nunits, d_meta, d_pool, d_chunk_next, d_cursor, pool_chunks, pl.quiet_thr, pl.quiet_thr_lsb, d_counters, grid, st);
   return cudaGetLastError(); }

/* per-track mask thresholds of the two-pass scan from the span histogram of the rows ingested so far (one small kernel + 36 KB D2H) */
static cudaError_t plan_auto_t0(const rt_tape *t, ScanPlan *pl, uint64_t rows_avail, uint32_t *d_hist, std::vector<uint32_t> &h_hist, bool &have_hist, cudaStream_t st) {
   if (!pl->t0_auto) return cudaSuccess;
   const int nt = (int)t->desc.ntrks;
   if (!have_hist) {
      cudaError_t e = launch_span_hist(reinterpret_cast<const uint32_t *>(t->gmm), t->ngran_cap, rows_avail, nt, d_hist, st);
      if (e != cudaSuccess) return e;
      h_hist.resize(span_hist_words(nt));
      e = cudaMemcpyAsync(h_hist.data(), d_hist, h_hist.size() * 4, cudaMemcpyDeviceToHost, st); if (e != cudaSuccess) return e;
      e = cudaStreamSynchronize(st); if (e != cudaSuccess) return e;
      have_hist = true; }
   if (!peak_mask_auto_T0(pl->dc, h_hist.data())) pl->use_sparse = false;
   if (getenv("RT_TRACE")) { fprintf(stderr, "[two-pass scan] T0/T1 per track:"); for (int k = 0; k < nt; ++k) fprintf(stderr, " %d/%d", pl->dc.T0[k], pl->dc.T1[k]); fprintf(stderr, "\n"); }
   return cudaSuccess; }

/* rt_prepare(): the per-track mask thresholds of the announced configuration from the rows ingested so far */
static cudaError_t plan_thresholds_from_rows(rt_tape *t, uint64_t rows_avail) {
   rt_tape::PreMask &pm = t->pm;
   ScanPlan pl; make_plan(t, &pm.cfg, &pl);
   if (!pl.use_sparse) { pm.active = false; return cudaSuccess; }
   if (pl.t0_auto) {
      uint32_t *d_hist = nullptr; std::vector<uint32_t> h_hist; bool have = false;
      cudaError_t e = cudaMalloc(&d_hist, (size_t)span_hist_words((int)t->desc.ntrks) * 4);
      if (e != cudaSuccess) return e;
      e = plan_auto_t0(t, &pl, rows_avail, d_hist, h_hist, have, t->stream);
      cudaFree(d_hist);
      if (e != cudaSuccess) return e;
      if (!pl.use_sparse) { pm.active = false; return cudaSuccess; } }
   memcpy(pm.T0, pl.dc.T0, sizeof pm.T0); memcpy(pm.T1, pl.dc.T1, sizeof pm.T1);
   pm.have_thr = true;
   return cudaSuccess; }

extern "C" int rt_prepare(rt_tape *t, const rt_scan_cfg *cfg) {
   if (!t) return set_err(RT_ERR_ARG, "rt_prepare: null");
   rt_tape::PreMask &pm = t->pm;
   if (!cfg) { pm.active = false; return RT_OK; }
   int rc = cfg_check(t, cfg); if (rc) return rc;
   DevCfg dc; cfg_to_dev(t, cfg, &dc);
   /* Measured on a B200 (config 2): the fused kernel takes 17.4 ms where the TMA ingest (7.2 ms) and the separate mask pass (7.2 ms)
      take 14.4 ms together -- the mask arithmetic is ALU-bound and gets 9 warps per SM inside the persistent ingest CTA instead of ~21 --
      so the fusion saves the 20 GB re-read but loses time.  It stays available for experiments (RT_FUSED_MASKS=1); DESIGN.md 6b. */
   /* What does pay is running the two kernels side by side (mode 2, the default): the ingest kernel is bound by HBM, the mask kernel by
      the integer pipe; chunk by chunk on two streams they take 12.0 ms together instead of 14.3 ms one after the other.
      RT_FUSED_MASKS=0: rt_prepare has no effect (the mask pass runs inside rt_bulk_scan); =1: the fused kernel; =2: side by side. */
   const char *env = getenv("RT_FUSED_MASKS");
   if (!env) env = "2";
   const bool ok = (env[0] == '1' || env[0] == '2') && dc.det == RT_DET_PEAK && !dc.invert && !dc.differentiate && !dc.density
                   && (cfg->mode == RT_MODE_NRZI || cfg->mode == RT_MODE_PE)
                   && (env[0] == '2' || (ingest_masks_supported((int)t->desc.nheads, (int)t->desc.ntrks, dc.width) && !t->force_simple_ingest));
   if (!ok) { pm.active = false; return RT_OK; }
   if (pm.active && pm.width == dc.width && pm.rise == cfg->parms.pkww_rise && memcmp(&pm.cfg, cfg, sizeof *cfg) == 0) return RT_OK;   /* unchanged: thresholds and planes stay */
   const bool same_masks = pm.width == dc.width && pm.rise == cfg->parms.pkww_rise && pm.cfg.bpi == cfg->bpi && pm.cfg.ips == cfg->ips && pm.cfg.mode == cfg->mode;
   pm.active = true; pm.cfg = *cfg; pm.width = dc.width; pm.rise = cfg->parms.pkww_rise; pm.overlap = env[0] == '2';
   if (!same_masks) { pm.have_thr = false; pm.rows = 0; }
   return RT_OK; }

extern "C" void rt_bulk_free(rt_bulk *b) {
   if (!b) return;
   rt_tape *t = b->tape;
   cudaSetDevice(t->device);
   if (b->bridge) rt_scan_end(b->bridge);
   if (b->user_pool) { /* the caller's memory */ }
   else if (b->h_pool_shared_bytes) munmap(b->h_pool, b->h_pool_shared_bytes);
   else if (b->pin_from_cache) t->pin_cache_busy = false; else if (b->h_pool) cudaFreeHost(b->h_pool);
   if (b->pool_from_cache) t->pool_cache_busy = false; else { cudaFree(b->d_pool); cudaFree(b->d_chunk_next); }
   for (auto &c : b->cfgs) { if (c.d_units) cudaFreeAsync(c.d_units, t->stream); if (c.d_meta) cudaFreeAsync(c.d_meta, t->stream); }
   delete b; }

extern "C" int rt_bulk_scan(rt_tape *t, const rt_scan_cfg *cfgs, uint32_t ncfgs, rt_bulk **out) {
   if (!t || !cfgs || !ncfgs || !out) return set_err(RT_ERR_ARG, "rt_bulk_scan: null argument");
   int rc = tape_sync_valid(t); if (rc) return rc;
   CU(cudaSetDevice(t->device));
   for (uint32_t i = 0; i < ncfgs; ++i) {
      rc = cfg_check(t, &cfgs[i]); if (rc) return rc;
      if (cfgs[i].mode == RT_MODE_WW) return set_err(RT_ERR_UNSUPPORTED, "Whirlwind state persists across blocks: use rt_scan_*");
      if ((cfgs[i].flags & RT_F_DENSITY_DETECT) && (cfgs[i].flags & RT_F_FIND_ZEROS))
         return set_err(RT_ERR_UNSUPPORTED, "density detection with the zero-crossing detector: use rt_scan_*"); }
   const bool trace = getenv("RT_TRACE") != nullptr;
   auto wall0 = std::chrono::steady_clock::now();
   auto lap = [&](const char *what) {
      if (!trace) return;
      auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[rt_bulk_scan] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - wall0).count());
      wall0 = now; };
   rt_bulk *b = new (std::nothrow) rt_bulk();
   if (!b) return set_err(RT_ERR_NOMEM, "rt_bulk_scan: out of memory");
   b->tape = t; b->cfgs.resize(ncfgs);
   const uint32_t nt = t->desc.ntrks; const uint64_t nrows = t->nrows_valid;
   int launches0 = t->launches;
   /* scratch for the unit finder; one set of counters per configuration */
   const size_t words = units_bitmap_words(nrows), nblocks = units_blocks(nrows) + 1;
   const uint32_t units_cap = (uint32_t)std::min<uint64_t>(nrows / RT_GRAN + 2, 0x7fffffffu);
   uint32_t *d_bitmap = nullptr, *d_flags = nullptr, *d_blockcount = nullptr, *d_nunits = nullptr; UnitDesc *d_units_tmp = nullptr;
   unsigned long long *d_counters = nullptr; unsigned int *d_cursor = nullptr;
   cudaEvent_t ev[5]; for (auto &e : ev) cudaEventCreate(&e);
   cudaEvent_t ev_rec = nullptr; cudaEventCreate(&ev_rec);
   std::vector<cudaEvent_t> done_ev(ncfgs, nullptr);
   /* K3c: candidate / canonical bit planes, one set per distinct (window width, T0) -- parameter sets that only differ in
      clock / AGC constants share them */
   struct MaskSet { int width; int32_t T0[RT_MAXTRKS], T1[RT_MAXTRKS]; uint32_t *mc, *md, *ma; uint32_t first_cfg; uint32_t *tb, *tc; bool borrowed; };
   unsigned int *d_rec_cursor = nullptr;                            /* phase B1: records of all mask sets share one pool and one cursor */
   /* phase B1 records are opt-in (RT_SPARSE_RECORDS=1): bit-exact, but measured slower than deriving the candidates inside the walk
      (B200, config 2: records 69 ms + walk 46 ms against 19 ms) */
   const char *recenv = getenv("RT_SPARSE_RECORDS");
   bool use_records = recenv && recenv[0] == '1';
   const uint64_t rec_tiles = cand_rec_tiles(t->plane_stride);
   uint32_t *d_hist = nullptr; std::vector<uint32_t> h_hist; bool have_hist = false;
   std::vector<MaskSet> msets;
   auto cleanup = [&]() {
      void *scr[] = {d_bitmap, d_flags, d_blockcount, d_nunits, d_units_tmp, d_counters, d_cursor, d_hist};
      for (void *p : scr) if (p) cudaFreeAsync(p, t->stream);
      for (auto &m : msets) { if (!m.borrowed) { if (m.mc) cudaFreeAsync(m.mc, t->stream); if (m.md) cudaFreeAsync(m.md, t->stream); if (m.ma) cudaFreeAsync(m.ma, t->stream); }
                              if (m.tb) cudaFreeAsync(m.tb, t->stream); if (m.tc) cudaFreeAsync(m.tc, t->stream); }
      if (d_rec_cursor) cudaFreeAsync(d_rec_cursor, t->stream);
      for (auto &e : ev) cudaEventDestroy(e);
      if (ev_rec) cudaEventDestroy(ev_rec);
      for (auto e : done_ev) if (e) cudaEventDestroy(e); };
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaDeviceSynchronize(); cleanup(); rt_bulk_free(b); \
      return set_err(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
   CUB(cudaMallocAsync(&d_bitmap, words * 4, t->stream)); CUB(cudaMallocAsync(&d_flags, words * 4, t->stream)); CUB(cudaMallocAsync(&d_blockcount, nblocks * 4, t->stream));
   CUB(cudaMallocAsync(&d_nunits, 4, t->stream)); CUB(cudaMallocAsync(&d_units_tmp, (size_t)units_cap * sizeof(UnitDesc), t->stream));
   CUB(cudaMallocAsync(&d_counters, 32 * (size_t)ncfgs, t->stream)); CUB(cudaMallocAsync(&d_cursor, 4, t->stream));
   if (ncfgs > 1 && t->s_par.empty()) {                           /* streams for the configurations' scan kernels */
      t->s_par.resize(4);
      for (auto &st : t->s_par) CUB(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); }
   lap("scratch alloc");
   b->stats.rows = nrows; b->stats.ms_preprocess = t->ms_ingest;

   /* 1. unit tables, one per configuration (the proposal thresholds depend on it) */
   std::vector<ScanPlan> plans(ncfgs);
   uint64_t total_units = 0;
   for (uint32_t ci = 0; ci < ncfgs; ++ci) {
      BulkCfg &bc = b->cfgs[ci];
      if ((cfgs[ci].flags & RT_F_INVERT) && !(cfgs[ci].flags & RT_F_DIFFERENTIATE)) { int rc_ = tape_ensure_inverted(t, nrows); if (rc_) { cleanup(); rt_bulk_free(b); return rc_; } }
      ScanPlan &pl = plans[ci]; make_plan(t, &cfgs[ci], &pl, true);
      bc.cfg = cfgs[ci]; bc.dc = pl.dc; bc.fast = pl.use_fast;
      CUB(cudaEventRecord(ev[0], t->stream));
      CUB(launch_find_units(t->gmm, t->ngran_cap, (int)nt, nrows, pl.up, d_bitmap, d_flags, d_blockcount, d_units_tmp, units_cap, d_nunits, t->stream, &t->launches));
      CUB(cudaEventRecord(ev[1], t->stream));
      uint32_t nunits = 0;
      CUB(cudaMemcpyAsync(&nunits, d_nunits, 4, cudaMemcpyDeviceToHost, t->stream));
      CUB(cudaStreamSynchronize(t->stream));
      if (nunits > units_cap) { cleanup(); rt_bulk_free(b); return set_err(RT_ERR_OVERFLOW, "unit table overflow (%u > %u)", nunits, units_cap); }
      bc.nunits = nunits; total_units += nunits;
      float ms_units = 0; cudaEventElapsedTime(&ms_units, ev[0], ev[1]); b->stats.ms_units += ms_units;
      if (nunits) {
         CUB(cudaMallocAsync(&bc.d_units, (size_t)nunits * sizeof(UnitDesc), t->stream));
         CUB(cudaMemcpyAsync(bc.d_units, d_units_tmp, (size_t)nunits * sizeof(UnitDesc), cudaMemcpyDeviceToDevice, t->stream));
         CUB(cudaMallocAsync(&bc.d_meta, (size_t)nunits * nt * sizeof(TrkMeta), t->stream)); } }
   b->stats.units = b->cfgs[ncfgs - 1].nunits;
   lap("unit finder + sync");
   {  const uint64_t mstride = peak_mask_stride(t->plane_stride);
      for (uint32_t ci = 0; ci < ncfgs; ++ci) {
         ScanPlan &pl = plans[ci];
         if (!pl.use_sparse || !b->cfgs[ci].nunits) continue;
         const rt_tape::PreMask &pm = t->pm;
         const bool adopt = pm.active && pm.have_thr && pm.mc && pm.stride == mstride && pm.rows >= nrows && pm.width == pl.dc.width
                            && pm.rise == pl.dc.p.pkww_rise && pm.cfg.bpi == cfgs[ci].bpi && pm.cfg.ips == cfgs[ci].ips && pm.cfg.mode == cfgs[ci].mode;
         if (adopt) {                                              /* the ingest kernel has already written these planes (rt_prepare) */
            memcpy(pl.dc.T0, pm.T0, sizeof pl.dc.T0); memcpy(pl.dc.T1, pm.T1, sizeof pl.dc.T1);
            size_t k = 0;
            while (k < msets.size() && msets[k].mc != pm.mc) ++k;
            if (k == msets.size()) {
               MaskSet m{}; m.width = pl.dc.width; memcpy(m.T0, pl.dc.T0, sizeof m.T0); memcpy(m.T1, pl.dc.T1, sizeof m.T1); m.first_cfg = ci;
               m.mc = pm.mc; m.md = pm.md; m.ma = pm.ma; m.borrowed = true;
               msets.push_back(m);
               if (use_records) {
                  CUB(cudaMallocAsync(&msets[k].tb, (size_t)rec_tiles * nt * 4, t->stream));
                  CUB(cudaMallocAsync(&msets[k].tc, (size_t)rec_tiles * nt * 4, t->stream)); } }
            pl.dc.m_cand = pm.mc; pl.dc.m_cand2 = pm.md; pl.dc.m_acan = pm.ma; pl.dc.mask_stride = mstride;
            if (getenv("RT_PREMASK_CHECK")) {                       /* tests: the fused planes against the separate mask pass, word for word */
               uint32_t *tc_ = nullptr, *td_ = nullptr, *ta_ = nullptr;
               const size_t bytes = (size_t)mstride * nt * 4;
               CUB(cudaMalloc(&tc_, bytes)); CUB(cudaMalloc(&td_, bytes)); CUB(cudaMalloc(&ta_, bytes));
               DevCfg dchk = pl.dc; dchk.m_cand = tc_; dchk.m_cand2 = td_; dchk.m_acan = ta_;
               CUB(launch_peak_masks(dchk, 0, nrows, t->stream));
               std::vector<uint32_t> h0(mstride * nt), h1(mstride * nt);
               const uint32_t *pairs[3][2] = {{pm.mc, tc_}, {pm.md, td_}, {pm.ma, ta_}};
               uint64_t bad = 0; const uint64_t nw = nrows / 32;       /* whole words of valid rows */
               for (auto &pr : pairs) {
                  CUB(cudaMemcpyAsync(h0.data(), pr[0], bytes, cudaMemcpyDeviceToHost, t->stream)); CUB(cudaMemcpyAsync(h1.data(), pr[1], bytes, cudaMemcpyDeviceToHost, t->stream));
                  CUB(cudaStreamSynchronize(t->stream));
                  for (uint32_t k2 = 0; k2 < nt; ++k2) for (uint64_t wi = 0; wi < nw; ++wi) if (h0[k2 * mstride + wi] != h1[k2 * mstride + wi]) ++bad; }
               cudaFree(tc_); cudaFree(td_); cudaFree(ta_);
               if (bad) { cleanup(); rt_bulk_free(b); return set_err(RT_ERR_STATE, "fused mask planes differ from the separate pass in %llu words", (unsigned long long)bad); }
               fprintf(stderr, "[rt_bulk_scan] fused mask planes identical to the separate pass (%llu rows, %llu fused)\n", (unsigned long long)nrows, (unsigned long long)pm.fused_rows); }
            continue; }
         if (pl.t0_auto) {
            if (!d_hist) { CUB(cudaMallocAsync(&d_hist, (size_t)span_hist_words((int)nt) * 4, t->stream)); ++t->launches; }
            CUB(plan_auto_t0(t, &pl, nrows, d_hist, h_hist, have_hist, t->stream));
            if (!pl.use_sparse) continue; }
         size_t k = 0;
         while (k < msets.size() && !(msets[k].width == pl.dc.width && memcmp(msets[k].T0, pl.dc.T0, sizeof pl.dc.T0) == 0
                                      && memcmp(msets[k].T1, pl.dc.T1, sizeof pl.dc.T1) == 0)) ++k;
         if (k == msets.size()) {
            MaskSet m{}; m.width = pl.dc.width; memcpy(m.T0, pl.dc.T0, sizeof m.T0); memcpy(m.T1, pl.dc.T1, sizeof m.T1); m.first_cfg = ci;
            msets.push_back(m);
            CUB(cudaMallocAsync(&msets[k].mc, (size_t)mstride * nt * 4, t->stream));
            CUB(cudaMallocAsync(&msets[k].md, (size_t)mstride * nt * 4, t->stream));
            CUB(cudaMallocAsync(&msets[k].ma, (size_t)mstride * nt * 4, t->stream));
            if (use_records) {
               CUB(cudaMallocAsync(&msets[k].tb, (size_t)rec_tiles * nt * 4, t->stream));
               CUB(cudaMallocAsync(&msets[k].tc, (size_t)rec_tiles * nt * 4, t->stream)); } }
         pl.dc.m_cand = msets[k].mc; pl.dc.m_cand2 = msets[k].md; pl.dc.m_acan = msets[k].ma; pl.dc.mask_stride = mstride; } }
   lap("mask thresholds");

   /* 2. all scan kernels, concurrently, into one event pool: first guess one event per 16 track-samples, regrown on overflow */
   if (total_units) {
      /* first guess: one event per 64 track-samples (0.5 B per track-sample; a block of 800 BPI NRZI at 20 samples per bit has one
         per ~40, half of a reel is gap) plus one partly filled chunk per (unit, track); an overflow costs one regrowth + rescan, after
         which the tape's cached pool fits */
      uint64_t want_chunks = std::max<uint64_t>(4096, (uint64_t)ncfgs * (nrows * nt / 64 / RT_EVC) + total_units * nt);
      if (t->chunks_hist) want_chunks = std::max<uint64_t>(want_chunks, (uint64_t)ncfgs * ((uint64_t)t->chunks_hist + t->chunks_hist / 8 + 64));
      /* phase B1's record pool: what the last scan of this tape needed, else one record per 28 track-samples per mask set */
      uint64_t want_recs = 0;
      if (use_records && !msets.empty()) {
         want_recs = t->rec_hist ? (uint64_t)t->rec_hist + t->rec_hist / 16 + 4096 : (uint64_t)msets.size() * (nrows * nt / 28) + 65536;
         if (!d_rec_cursor) CUB(cudaMallocAsync(&d_rec_cursor, 4, t->stream)); }
      bool recs_stale = true;
      for (int attempt = 0; attempt < 3; ++attempt) {
         if (use_records && !msets.empty() && want_recs > t->rec_cache_cap) {
            if (want_recs > 0xfffffff0ull) use_records = false;
            else {
               cudaFree(t->rec_cache); cudaFree(t->pm.mc); cudaFree(t->pm.md); cudaFree(t->pm.ma); t->rec_cache = nullptr; t->rec_cache_cap = 0;
               if (cudaMalloc(&t->rec_cache, (size_t)want_recs * sizeof(CandRec)) != cudaSuccess) { cudaGetLastError(); use_records = false; }
               else { t->rec_cache_cap = (uint32_t)want_recs; recs_stale = true; } } }
         if (want_chunks > b->pool_chunks) {
            if (want_chunks > 0xfffffff0ull) { cleanup(); rt_bulk_free(b); return set_err(RT_ERR_OVERFLOW, "event pool too large"); }
            if (b->pool_from_cache || (!b->d_pool && !t->pool_cache_busy)) {          /* use / grow the tape's cached pool */
               if (want_chunks > t->pool_cache_chunks) {
                  cudaFree(t->pool_cache); cudaFree(t->next_cache); t->pool_cache = nullptr; t->next_cache = nullptr; t->pool_cache_chunks = 0;
                  CUB(cudaMalloc(&t->pool_cache, (size_t)want_chunks * RT_EVC * sizeof(rt_event)));
                  CUB(cudaMalloc(&t->next_cache, (size_t)want_chunks * 4));
                  t->pool_cache_chunks = (uint32_t)want_chunks; }
               b->d_pool = t->pool_cache; b->d_chunk_next = t->next_cache; b->pool_chunks = t->pool_cache_chunks;
               b->pool_from_cache = true; t->pool_cache_busy = true; }
            else {
               cudaFree(b->d_pool); cudaFree(b->d_chunk_next); b->d_pool = nullptr; b->d_chunk_next = nullptr;
               b->pool_chunks = (uint32_t)want_chunks;
               CUB(cudaMalloc(&b->d_pool, (size_t)b->pool_chunks * RT_EVC * sizeof(rt_event)));
               CUB(cudaMalloc(&b->d_chunk_next, (size_t)b->pool_chunks * 4)); } }
         CUB(cudaMemsetAsync(d_cursor, 0, 4, t->stream));
         CUB(cudaMemsetAsync(d_counters, 0, 32 * (size_t)ncfgs, t->stream));
         CUB(cudaEventRecord(ev[2], t->stream));
         if (attempt == 0) {                                      /* phase A of the two-pass scan: once per mask set */
            for (auto &m : msets) if (!m.borrowed) { CUB(launch_peak_masks(plans[m.first_cfg].dc, 0, nrows, t->stream)); ++t->launches; } }
         CUB(cudaEventRecord(ev[4], t->stream));
         if (use_records && !msets.empty() && recs_stale) {       /* phase B1: the candidate records of every mask set */
            CUB(cudaMemsetAsync(d_rec_cursor, 0, 4, t->stream));
            for (auto &m : msets) {
               DevCfg dcm = plans[m.first_cfg].dc; dcm.rec_tiles = rec_tiles;
               CUB(launch_cand_records(dcm, 0, nrows, t->rec_cache, t->rec_cache_cap, m.tb, m.tc, d_rec_cursor, t->stream)); ++t->launches; }
            recs_stale = false; }
         for (uint32_t ci = 0; ci < ncfgs; ++ci) {                 /* hand the records to the configurations of each mask set */
            ScanPlan &pl = plans[ci];
            pl.dc.recs = nullptr;
            if (!use_records || !pl.use_sparse) continue;
            for (auto &m : msets)
               if (m.mc == pl.dc.m_cand) { pl.dc.recs = t->rec_cache; pl.dc.rec_tile_base = m.tb; pl.dc.rec_tile_cnt = m.tc; pl.dc.rec_tiles = rec_tiles; } }
         CUB(cudaEventRecord(ev_rec, t->stream));
         for (uint32_t ci = 0; ci < ncfgs; ++ci) {
            BulkCfg &bc = b->cfgs[ci];
            if (!bc.nunits) continue;
            cudaStream_t st = ncfgs > 1 ? t->s_par[ci % t->s_par.size()] : t->stream;
            if (st != t->stream) CUB(cudaStreamWaitEvent(st, ev[4], 0));
            CUB(launch_scan(t, plans[ci], bc.d_units, bc.nunits, bc.d_meta, b->d_pool, b->d_chunk_next, d_cursor, b->pool_chunks, d_counters + 4 * ci, 0, st));
            ++t->launches;
            if (st != t->stream) {
               if (!done_ev[ci]) CUB(cudaEventCreateWithFlags(&done_ev[ci], cudaEventDisableTiming));
               CUB(cudaEventRecord(done_ev[ci], st)); CUB(cudaStreamWaitEvent(t->stream, done_ev[ci], 0)); } }
         CUB(cudaEventRecord(ev[3], t->stream));
         unsigned int used = 0, used_recs = 0;
         CUB(cudaMemcpyAsync(&used, d_cursor, 4, cudaMemcpyDeviceToHost, t->stream));
         if (use_records && !msets.empty()) CUB(cudaMemcpyAsync(&used_recs, d_rec_cursor, 4, cudaMemcpyDeviceToHost, t->stream));
         CUB(cudaStreamSynchronize(t->stream));
         float ms = 0; cudaEventElapsedTime(&ms, ev[2], ev[3]); b->stats.ms_scan = attempt == 0 ? ms : b->stats.ms_scan + ms;
         if (attempt == 0 && !msets.empty()) { cudaEventElapsedTime(&ms, ev[2], ev[4]); b->stats.ms_masks = ms;
                                               cudaEventElapsedTime(&ms, ev[4], ev_rec); b->stats.ms_records = ms; }
         lap("scan kernel(s) + sync");
         bool again = false;
         if (use_records && !msets.empty()) {
            t->rec_hist = used_recs;
            if (used_recs > t->rec_cache_cap) { want_recs = (uint64_t)used_recs + used_recs / 16 + 4096; again = true; } }   /* some tiles had no room: redo */
         if (used > b->pool_chunks) { want_chunks = (uint64_t)used + used / 8 + 1024; again = true; }
         if (!again) { b->chunks_used = used; t->chunks_hist = (uint32_t)((used + ncfgs - 1) / ncfgs); break; }
         if (attempt == 2) { cleanup(); rt_bulk_free(b); return set_err(RT_ERR_OVERFLOW, "event pool overflow after regrowth"); } }
      std::vector<unsigned long long> counters(4 * (size_t)ncfgs, 0);
      CUB(cudaMemcpy(counters.data(), d_counters, 32 * (size_t)ncfgs, cudaMemcpyDeviceToHost));
      for (uint32_t ci = 0; ci < ncfgs; ++ci) { b->stats.rows_scanned += counters[4 * ci]; b->stats.events += counters[4 * ci + 1]; } }
   lap("counters");
   b->stats.track_samples = nrows * nt * ncfgs;
   b->stats.launches = (uint32_t)(t->launches - launches0);
   b->stats.launches_ingest = (uint32_t)(launches0 - t->launches_at_clear);
   for (auto &m : msets) { b->stats.two_pass = 1; if (m.borrowed) b->stats.masks_fused = 1; }
   if (ncfgs == 1) { t->hist_rows = nrows; t->hist_units = b->cfgs[0].nunits; t->hist_chunks = b->chunks_used; }   /* sizes the streamed scan */
   cleanup();
#undef CUB
   *out = b; return RT_OK; }

/* Bring the results of rt_bulk_scan() to the host (unit tables, proof data, events).  Called by the first
 * rt_bulk_lookup(); separate so that a caller can overlap it or time it. */
extern "C" int rt_bulk_fetch(rt_bulk *b) {
   if (!b) return set_err(RT_ERR_ARG, "rt_bulk_fetch: null");
   if (b->fetched) return RT_OK;
   rt_tape *t = b->tape; const uint32_t nt = t->desc.ntrks;
   CU(cudaSetDevice(t->device));
   for (BulkCfg &bc : b->cfgs) {
      const uint32_t nunits = bc.nunits;
      bc.units.resize(nunits); bc.meta.resize((size_t)nunits * nt);
      if (!nunits) continue;
      CU(cudaMemcpyAsync(bc.units.data(), bc.d_units, (size_t)nunits * sizeof(UnitDesc), cudaMemcpyDeviceToHost, t->stream));
      CU(cudaMemcpyAsync(bc.meta.data(), bc.d_meta, (size_t)nunits * nt * sizeof(TrkMeta), cudaMemcpyDeviceToHost, t->stream));
      b->stats.d2h_bytes += (uint64_t)nunits * (sizeof(UnitDesc) + nt * sizeof(TrkMeta)); }
   b->chunk_next.resize(b->chunks_used);
   if (b->chunks_used) {
      CU(cudaMemcpyAsync(b->chunk_next.data(), b->d_chunk_next, (size_t)b->chunks_used * 4, cudaMemcpyDeviceToHost, t->stream));
      b->h_pool_events = (size_t)b->chunks_used * RT_EVC;
      if (b->user_pool) {
         if (b->h_pool_events * sizeof(rt_event) > b->user_pool_bytes)
            return set_err(RT_ERR_OVERFLOW, "rt_bulk_fetch_to: the buffer holds %zu bytes, the events need %zu", b->user_pool_bytes, b->h_pool_events * sizeof(rt_event));
         b->h_pool = b->user_pool;
         CU(cudaMemcpyAsync(b->h_pool, b->d_pool, b->h_pool_events * sizeof(rt_event), cudaMemcpyDeviceToHost, t->stream));
         b->stats.d2h_bytes += b->h_pool_events * sizeof(rt_event) + (uint64_t)b->chunks_used * 4; }
      else if (g_opt_shared_results) {
         /* memory that worker processes forked after this call can read: shared anonymous pages, pinned only while the copy runs */
         const size_t bytes = (b->h_pool_events * sizeof(rt_event) + 4095) / 4096 * 4096;
         void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_POPULATE, -1, 0);
         if (p == MAP_FAILED) return set_err(RT_ERR_NOMEM, "rt_bulk_fetch: cannot map %zu bytes of shared memory", bytes);
         b->h_pool = static_cast<rt_event *>(p); b->h_pool_shared_bytes = bytes;
         const bool reg = cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess;
         if (!reg) cudaGetLastError();
         CU(cudaMemcpyAsync(b->h_pool, b->d_pool, b->h_pool_events * sizeof(rt_event), cudaMemcpyDeviceToHost, t->stream));
         CU(cudaStreamSynchronize(t->stream));
         if (reg) cudaHostUnregister(p);
         b->stats.d2h_bytes += b->h_pool_events * sizeof(rt_event) + (uint64_t)b->chunks_used * 4; }
      else {
      if (!t->pin_cache_busy) {                                  /* use / grow the tape's cached pinned buffer */
         if (b->h_pool_events > t->pin_cache_events) {
            if (t->pin_cache) cudaFreeHost(t->pin_cache);
            t->pin_cache = nullptr; t->pin_cache_events = 0;
            size_t want = b->h_pool_events + b->h_pool_events / 16;
            CU(cudaHostAlloc(&t->pin_cache, want * sizeof(rt_event), cudaHostAllocDefault));
            t->pin_cache_events = want; }
         b->h_pool = t->pin_cache; b->pin_from_cache = true; t->pin_cache_busy = true; }
      else CU(cudaHostAlloc(&b->h_pool, b->h_pool_events * sizeof(rt_event), cudaHostAllocDefault));
      CU(cudaMemcpyAsync(b->h_pool, b->d_pool, b->h_pool_events * sizeof(rt_event), cudaMemcpyDeviceToHost, t->stream));
      b->stats.d2h_bytes += b->h_pool_events * sizeof(rt_event) + (uint64_t)b->chunks_used * 4; } }
   CU(cudaStreamSynchronize(t->stream));
   /* the device copies are no longer needed */
   for (BulkCfg &bc : b->cfgs) {
      if (bc.d_units) cudaFreeAsync(bc.d_units, t->stream); if (bc.d_meta) cudaFreeAsync(bc.d_meta, t->stream);
      bc.d_units = nullptr; bc.d_meta = nullptr; }
   if (b->pool_from_cache) { t->pool_cache_busy = false; b->pool_from_cache = false; } else { cudaFree(b->d_pool); cudaFree(b->d_chunk_next); }
   b->d_pool = nullptr; b->d_chunk_next = nullptr;
   b->fetched = true;
   return RT_OK; }

extern "C" int rt_bulk_results_size(const rt_bulk *b, uint64_t *bytes) {
   if (!b || !bytes) return set_err(RT_ERR_ARG, "rt_bulk_results_size: null argument");
   const uint32_t nt = b->tape->desc.ntrks;
   uint64_t n = 0;
   for (const BulkCfg &bc : b->cfgs) n += 16 + (uint64_t)bc.nunits * (sizeof(UnitDesc) + nt * sizeof(TrkMeta));
   n += 8 + ((uint64_t)b->chunks_used * 4 + 7) / 8 * 8 + (uint64_t)b->chunks_used * RT_EVC * sizeof(rt_event);
   *bytes = n; return RT_OK; }

extern "C" int rt_bulk_results_to_device(const rt_bulk *b, void *dst_dev, uint64_t bytes) {
   uint64_t need = 0;
   int rc = rt_bulk_results_size(b, &need); if (rc) return rc;
   if (!dst_dev || bytes < need) return set_err(RT_ERR_ARG, "rt_bulk_results_to_device: buffer too small (%llu < %llu)", (unsigned long long)bytes, (unsigned long long)need);
   if (b->fetched || (b->chunks_used && !b->d_pool)) return set_err(RT_ERR_STATE, "rt_bulk_results_to_device: the results have left the device");
   rt_tape *t = b->tape; const uint32_t nt = t->desc.ntrks;
   CU(cudaSetDevice(t->device));
   char *p = static_cast<char *>(dst_dev);
   for (const BulkCfg &bc : b->cfgs) {
      const uint64_t hdr[2] = {bc.nunits, nt};
      CU(cudaMemcpyAsync(p, hdr, 16, cudaMemcpyHostToDevice, t->stream)); p += 16;
      if (bc.nunits) {
         CU(cudaMemcpyAsync(p, bc.d_units, (size_t)bc.nunits * sizeof(UnitDesc), cudaMemcpyDeviceToDevice, t->stream)); p += (size_t)bc.nunits * sizeof(UnitDesc);
         CU(cudaMemcpyAsync(p, bc.d_meta, (size_t)bc.nunits * nt * sizeof(TrkMeta), cudaMemcpyDeviceToDevice, t->stream)); p += (size_t)bc.nunits * nt * sizeof(TrkMeta); } }
   const uint64_t chunks = b->chunks_used;
   CU(cudaMemcpyAsync(p, &chunks, 8, cudaMemcpyHostToDevice, t->stream)); p += 8;
   if (chunks) {
      CU(cudaMemcpyAsync(p, b->d_chunk_next, (size_t)chunks * 4, cudaMemcpyDeviceToDevice, t->stream)); p += ((size_t)chunks * 4 + 7) / 8 * 8;
      CU(cudaMemcpyAsync(p, b->d_pool, (size_t)chunks * RT_EVC * sizeof(rt_event), cudaMemcpyDeviceToDevice, t->stream)); }
   CU(cudaStreamSynchronize(t->stream));
   return RT_OK; }

extern "C" int rt_bulk_fetch_to(rt_bulk *b, void *events_buf, size_t bytes) {
   if (!b || !events_buf) return set_err(RT_ERR_ARG, "rt_bulk_fetch_to: null argument");
   if (b->fetched) return set_err(RT_ERR_STATE, "rt_bulk_fetch_to: already fetched");
   b->user_pool = static_cast<rt_event *>(events_buf); b->user_pool_bytes = bytes;
   int rc = rt_bulk_fetch(b);
   if (rc) { b->user_pool = nullptr; b->user_pool_bytes = 0; }
   return rc; }

extern "C" int rt_host_register(rt_tape *t, void *p, size_t bytes) {
   if (!t || !p) return set_err(RT_ERR_ARG, "rt_host_register: null argument");
   CU(cudaSetDevice(t->device));
   CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
   return RT_OK; }
extern "C" int rt_host_unregister(rt_tape *t, void *p) {
   if (!t || !p) return set_err(RT_ERR_ARG, "rt_host_unregister: null argument");
   CU(cudaSetDevice(t->device));
   CU(cudaHostUnregister(p));
   return RT_OK; }

/* Upload + whole-tape scan + fetch in one call, overlapped.  Replaces  rt_clear(); rt_upload(); rt_bulk_scan(cfg); rt_bulk_fetch().
 * The host->device copy of a 20 GB capture takes ~7x longer than scanning it, so the tape is processed in segments while it
 * arrives: all copies + ingest kernels are enqueued on the tape's stream; after each segment the unit finder runs on the
 * prefix ingested so far (stream s_scan), the units that are already complete are scanned (with at most 2 CTAs per SM, so
 * that the ingest kernels of the following chunks still find room), and their events, proof data and chunk links go back
 * to the host on a third stream while the next segment is being copied.  Buffer capacities come from the previous
 * whole-tape scan of this tape object; without such history (or when they do not suffice) the plain sequence is used. */
extern "C" int rt_bulk_scan_host(rt_tape *t, const int16_t *rows, uint64_t nrows, const rt_scan_cfg *cfg, rt_bulk **out) {
   if (!t || !rows || !cfg || !out) return set_err(RT_ERR_ARG, "rt_bulk_scan_host: null argument");
   CU(cudaSetDevice(t->device));
   int rc = rt_clear(t); if (rc) return rc;
   const uint32_t nt = t->desc.ntrks; const uint64_t nh = t->desc.nheads;
   const char *env = getenv("RT_STREAM");
   const bool trace = getenv("RT_TRACE") != nullptr;
   bool stream_ok = !(env && env[0] == '0') && t->hist_rows && nrows <= t->hist_rows + t->hist_rows / 20 && nrows >= (16u << 20)
                    && cfg->mode != RT_MODE_WW && !(cfg->flags & RT_F_DENSITY_DETECT) && !t->pool_cache_busy && !t->pin_cache_busy
                    && t->pool_cache && t->pin_cache && t->pool_cache_chunks >= t->hist_chunks + t->hist_chunks / 32
                    && t->pin_cache_events / RT_EVC >= (size_t)t->hist_chunks + t->hist_chunks / 32;
   if (stream_ok) { rc = tape_reserve(t, nrows); if (rc) return rc; rc = cfg_check(t, cfg); if (rc) return rc; }
   if (!stream_ok) {
      rc = rt_upload(t, rows, nrows); if (rc) return rc;
      rc = rt_bulk_scan(t, cfg, 1, out); if (rc) return rc;
      return rt_bulk_fetch(*out); }

   uint64_t stage_rows = 0;
   rc = stage_prepare(t, nrows, &stage_rows); if (rc) return rc;
   if (!t->s_scan) { CU(cudaStreamCreateWithFlags(&t->s_scan, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&t->s_out, cudaStreamNonBlocking)); }
   ScanPlan pl; make_plan(t, cfg, &pl); pl.dc.nrows = nrows;
   const uint32_t cap_units = t->hist_units + t->hist_units / 8 + 1024;
   const uint32_t cap_chunks = (uint32_t)std::min<size_t>(t->pool_cache_chunks, t->pin_cache_events / RT_EVC);
   const size_t words = units_bitmap_words(nrows), nblocks = units_blocks(nrows) + 1;
   const uint32_t units_cap_tmp = (uint32_t)std::min<uint64_t>(nrows / RT_GRAN + 2, 0x7fffffffu);
   uint32_t *d_bitmap = nullptr, *d_flags = nullptr, *d_blockcount = nullptr, *d_nunits = nullptr; UnitDesc *d_units_tmp = nullptr;
   unsigned long long *d_counters = nullptr; unsigned int *d_cursor = nullptr; TrkMeta *d_meta = nullptr;
   uint32_t *d_mc = nullptr, *d_md = nullptr, *d_ma = nullptr; uint64_t masks_done = 0; double ms_masks = 0;   /* K3c bit planes, built segment by segment */
   uint32_t *d_hist = nullptr; std::vector<uint32_t> h_hist; bool have_hist = false;
   std::vector<cudaEvent_t> seg_ev; std::vector<uint64_t> seg_rows_done;
   cudaEvent_t ev_a = nullptr, ev_b = nullptr;
   rt_bulk *b = new (std::nothrow) rt_bulk();
   if (!b) return set_err(RT_ERR_NOMEM, "rt_bulk_scan_host: out of memory");
   b->tape = t; b->cfgs.resize(1);
   BulkCfg &bc = b->cfgs[0];
   bc.cfg = *cfg; bc.dc = pl.dc; bc.fast = pl.use_fast;
   bool failed = false; const char *why = "";
   auto release = [&]() {
      void *scr[] = {d_bitmap, d_flags, d_blockcount, d_nunits, d_units_tmp, d_counters, d_cursor, d_meta, d_mc, d_md, d_ma, d_hist};
      for (void *p : scr) if (p) cudaFreeAsync(p, t->s_scan);
      for (auto e : seg_ev) cudaEventDestroy(e);
      if (ev_a) cudaEventDestroy(ev_a); if (ev_b) cudaEventDestroy(ev_b); };
   struct EvGuard { cudaEvent_t *e; ~EvGuard() { if (*e) cudaEventDestroy(*e); } };
#define CUS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaDeviceSynchronize(); release(); t->pool_cache_busy = t->pin_cache_busy = false; delete b; \
      return set_err(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
   CUS(cudaMallocAsync(&d_bitmap, words * 4, t->s_scan)); CUS(cudaMallocAsync(&d_flags, words * 4, t->s_scan)); CUS(cudaMallocAsync(&d_blockcount, nblocks * 4, t->s_scan));
   CUS(cudaMallocAsync(&d_nunits, 4, t->s_scan)); CUS(cudaMallocAsync(&d_units_tmp, (size_t)units_cap_tmp * sizeof(UnitDesc), t->s_scan));
   CUS(cudaMallocAsync(&d_counters, 32, t->s_scan)); CUS(cudaMallocAsync(&d_cursor, 4, t->s_scan));
   CUS(cudaMallocAsync(&d_meta, (size_t)cap_units * nt * sizeof(TrkMeta), t->s_scan));
   if (pl.use_sparse) {
      const uint64_t mstride = peak_mask_stride(t->plane_stride);
      CUS(cudaMallocAsync(&d_mc, (size_t)mstride * nt * 4, t->s_scan)); CUS(cudaMallocAsync(&d_md, (size_t)mstride * nt * 4, t->s_scan));
      CUS(cudaMallocAsync(&d_ma, (size_t)mstride * nt * 4, t->s_scan));
      if (pl.t0_auto) CUS(cudaMallocAsync(&d_hist, (size_t)span_hist_words((int)nt) * 4, t->s_scan));
      pl.dc.m_cand = d_mc; pl.dc.m_cand2 = d_md; pl.dc.m_acan = d_ma; pl.dc.mask_stride = mstride; }
   cudaEvent_t ev_m = nullptr; EvGuard ev_m_guard{&ev_m};
   CUS(cudaMemsetAsync(d_cursor, 0, 4, t->s_scan)); CUS(cudaMemsetAsync(d_counters, 0, 32, t->s_scan));
   CUS(cudaEventCreate(&ev_a)); CUS(cudaEventCreate(&ev_b));
   t->pool_cache_busy = t->pin_cache_busy = true;
   rt_event *d_pool = t->pool_cache; uint32_t *d_chunk_next = t->next_cache; rt_event *h_pool = t->pin_cache;
   const int launches0 = t->launches;

   /* 1. everything that moves samples: enqueued once, runs by itself */
   {  const uint64_t seg_target = std::max<uint64_t>(stage_rows * 4, (nrows / 16 + stage_rows - 1) / stage_rows * stage_rows);
      uint64_t done = 0, next_mark = seg_target; int buf = 0;
      while (done < nrows) {
         const uint64_t n = std::min(stage_rows, nrows - done);
         rc = enqueue_chunk(t, rows + done * nh, n, buf);
         if (rc) { cudaDeviceSynchronize(); release(); t->pool_cache_busy = t->pin_cache_busy = false; delete b; return rc; }
         done += n; buf = (buf + 1) % t->nstage;
         if (done >= next_mark || done == nrows) {
            cudaEvent_t e; CUS(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); CUS(cudaEventRecord(e, t->stream));
            seg_ev.push_back(e); seg_rows_done.push_back(done); next_mark = done + seg_target; } }
      t->h2d_bytes += nrows * nh * 2; }

   /* the ingest kernels enqueued above may already be writing the mask planes of this configuration (rt_prepare) */
   bool masks_fused = false;
   {  const rt_tape::PreMask &pm = t->pm;
      if (pl.use_sparse && pm.active && pm.have_thr && pm.mc && pm.stride == peak_mask_stride(t->plane_stride) && pm.rows >= nrows && pm.width == pl.dc.width
          && pm.rise == pl.dc.p.pkww_rise && pm.cfg.bpi == cfg->bpi && pm.cfg.ips == cfg->ips && pm.cfg.mode == cfg->mode) {
         memcpy(pl.dc.T0, pm.T0, sizeof pl.dc.T0); memcpy(pl.dc.T1, pm.T1, sizeof pl.dc.T1);
         pl.dc.m_cand = pm.mc; pl.dc.m_cand2 = pm.md; pl.dc.m_acan = pm.ma; pl.t0_auto = false; have_hist = true;
         bc.dc = pl.dc; masks_fused = true; } }

   /* 2. per segment: units of the prefix, scan the complete ones, send their results home */
   uint32_t u_done = 0, c_done = 0, nun = 0; double ms_scan = 0, ms_units = 0;
   for (size_t k = 0; k < seg_ev.size() && !failed; ++k) {
      const bool last = k + 1 == seg_ev.size();
      const uint64_t R = seg_rows_done[k];
      CUS(cudaStreamWaitEvent(t->s_scan, seg_ev[k], 0));
      if (pl.use_sparse && pl.t0_auto && !have_hist) {            /* mask thresholds from the first segment of the tape */
         CUS(plan_auto_t0(t, &pl, R, d_hist, h_hist, have_hist, t->s_scan)); ++t->launches;
         bc.dc = pl.dc; }
      CUS(cudaEventRecord(ev_a, t->s_scan));
      CUS(launch_find_units(t->gmm, t->ngran_cap, (int)nt, R, pl.up, d_bitmap, d_flags, d_blockcount, d_units_tmp, units_cap_tmp, d_nunits, t->s_scan, &t->launches));
      CUS(cudaEventRecord(ev_b, t->s_scan));
      uint32_t nunits = 0;
      CUS(cudaMemcpyAsync(&nunits, d_nunits, 4, cudaMemcpyDeviceToHost, t->s_scan));
      CUS(cudaStreamSynchronize(t->s_scan));
      { float ms = 0; cudaEventElapsedTime(&ms, ev_a, ev_b); ms_units += ms; }
      if (nunits > cap_units || nunits > units_cap_tmp) { failed = true; why = "more units than the previous scan"; break; }
      nun = nunits;
      bc.units.resize(nun);
      if (nun > u_done) CUS(cudaMemcpy(bc.units.data() + u_done, d_units_tmp + u_done, (size_t)(nun - u_done) * sizeof(UnitDesc), cudaMemcpyDeviceToHost));
      uint32_t u_final = u_done;
      if (last) {
         rc = tape_sync_valid(t);
         if (rc || t->nrows_valid != nrows) { failed = true; why = "end-of-data marker inside the capture"; break; }
         u_final = nun; }
      else while (u_final + 1 < nun && bc.units[u_final + 1].row0 + pl.up.tail_rows + 4096 <= R) ++u_final;    /* row_end can no longer change */
      if (u_final == u_done) continue;
      CUS(cudaEventRecord(ev_a, t->s_scan));
      if (pl.use_sparse && !masks_fused) {                       /* phase A for the rows that arrived since the last segment */
         if (!ev_m) CUS(cudaEventCreate(&ev_m));
         CUS(launch_peak_masks(pl.dc, masks_done, R, t->s_scan)); ++t->launches;
         masks_done = R / 64 * 64;                               /* the run that holds row R is redone when it is complete */
         CUS(cudaEventRecord(ev_m, t->s_scan)); }
      CUS(launch_scan(t, pl, d_units_tmp + u_done, u_final - u_done, d_meta + (size_t)u_done * nt, d_pool, d_chunk_next, d_cursor, cap_chunks, d_counters,
                      last ? 0 : 2, t->s_scan));
      ++t->launches;
      CUS(cudaEventRecord(ev_b, t->s_scan));
      unsigned int used = 0;
      CUS(cudaMemcpyAsync(&used, d_cursor, 4, cudaMemcpyDeviceToHost, t->s_scan));
      CUS(cudaStreamSynchronize(t->s_scan));
      { float ms = 0; cudaEventElapsedTime(&ms, ev_a, ev_b); ms_scan += ms; if (ev_m) { cudaEventElapsedTime(&ms, ev_a, ev_m); ms_masks += ms; } }
      if (trace) fprintf(stderr, "[rt_bulk_scan_host] segment %zu: rows %llu, units %u..%u, chunks %u..%u\n", k, (unsigned long long)R, u_done, u_final, c_done, used);
      if (used > cap_chunks) { failed = true; why = "more events than the previous scan"; break; }
      bc.meta.resize((size_t)u_final * nt); b->chunk_next.resize(used);
      CUS(cudaMemcpy(bc.meta.data() + (size_t)u_done * nt, d_meta + (size_t)u_done * nt, (size_t)(u_final - u_done) * nt * sizeof(TrkMeta), cudaMemcpyDeviceToHost));
      if (used > c_done) {
         CUS(cudaMemcpy(b->chunk_next.data() + c_done, d_chunk_next + c_done, (size_t)(used - c_done) * 4, cudaMemcpyDeviceToHost));
         CUS(cudaMemcpyAsync(h_pool + (size_t)c_done * RT_EVC, d_pool + (size_t)c_done * RT_EVC, (size_t)(used - c_done) * RT_EVC * sizeof(rt_event),
                             cudaMemcpyDeviceToHost, t->s_out)); }
      b->stats.d2h_bytes += (uint64_t)(used - c_done) * (RT_EVC * sizeof(rt_event) + 4) + (uint64_t)(u_final - u_done) * (sizeof(UnitDesc) + nt * sizeof(TrkMeta));
      u_done = u_final; c_done = used; }
   CUS(cudaStreamSynchronize(t->s_out));
   rc = tape_drain(t);
   if (failed || rc) {                                           /* the samples are resident by now: plain scan + fetch */
      cudaDeviceSynchronize(); release(); t->pool_cache_busy = t->pin_cache_busy = false; delete b;
      if (rc) return rc;
      if (trace) fprintf(stderr, "[rt_bulk_scan_host] streamed scan abandoned (%s): plain scan\n", why);
      rc = tape_sync_valid(t); if (rc) return rc;
      rc = rt_bulk_scan(t, cfg, 1, out); if (rc) return rc;
      return rt_bulk_fetch(*out); }
   unsigned long long counters[2] = {0, 0};
   CUS(cudaMemcpy(counters, d_counters, 16, cudaMemcpyDeviceToHost));
   bc.nunits = nun; b->chunks_used = c_done; b->pool_chunks = cap_chunks;
   b->h_pool = h_pool; b->h_pool_events = (size_t)c_done * RT_EVC; b->pin_from_cache = true;
   t->pool_cache_busy = false;                                    /* the device pool is free again; the pinned copy belongs to the bulk */
   b->fetched = true;
   b->stats.rows = nrows; b->stats.units = nun; b->stats.events = counters[1]; b->stats.rows_scanned = counters[0];
   b->stats.track_samples = nrows * nt; b->stats.ms_preprocess = t->ms_ingest; b->stats.ms_units = ms_units; b->stats.ms_scan = ms_scan; b->stats.ms_masks = ms_masks;
   b->stats.launches = (uint32_t)(t->launches - launches0);
   b->stats.launches_ingest = (uint32_t)(launches0 - t->launches_at_clear);
   b->stats.pad = (uint32_t)seg_ev.size();                        /* segments streamed (0: the plain sequence was used) */
   b->stats.two_pass = pl.use_sparse; b->stats.masks_fused = masks_fused;
   t->hist_rows = nrows; t->hist_units = nun; t->hist_chunks = c_done;
   release();
#undef CUS
   *out = b; return RT_OK; }

/* diagnostics / tests: run phase A of the two-pass peak scan and hand the bit planes to the host */
extern "C" int rt_peak_masks(rt_tape *t, const rt_scan_cfg *cfg, float t0_frac, uint32_t *cand, uint32_t *acan, uint64_t wpt, int32_t *t0) {
   if (!t || !cfg || !cand || !acan) return set_err(RT_ERR_ARG, "rt_peak_masks: null argument");
   int rc = tape_sync_valid(t); if (rc) return rc;
   rc = cfg_check(t, cfg); if (rc) return rc;
   CU(cudaSetDevice(t->device));
   DevCfg dc; cfg_to_dev(t, cfg, &dc);
   const uint64_t nrows = t->nrows_valid;
   if (dc.det != RT_DET_PEAK || dc.invert || dc.differentiate || dc.density || wpt * 32 < nrows)
      return set_err(RT_ERR_UNSUPPORTED, "rt_peak_masks: not the plain peak detector, or buffer too small");
   const int T0 = peak_mask_T0(dc, t0_frac);
   if (T0 <= 0) return set_err(RT_ERR_UNSUPPORTED, "rt_peak_masks: T0 < 16");
   for (int k = 0; k < RT_MAXTRKS; ++k) dc.T0[k] = T0;
   if (t0) *t0 = T0;
   const uint64_t ms = peak_mask_stride(t->plane_stride); const uint32_t nt = t->desc.ntrks;
   uint32_t *d_mc = nullptr, *d_md = nullptr, *d_ma = nullptr;
   CU(cudaMalloc(&d_mc, (size_t)ms * nt * 4));
   cudaError_t e = cudaMalloc(&d_ma, (size_t)ms * nt * 4);
   if (e == cudaSuccess) e = cudaMalloc(&d_md, (size_t)ms * nt * 4);
   if (e == cudaSuccess) {
      dc.m_cand = d_mc; dc.m_cand2 = d_md; dc.m_acan = d_ma; dc.mask_stride = ms;
      e = launch_peak_masks(dc, 0, nrows, t->stream); ++t->launches;
      const uint64_t nw = (nrows + 31) / 32;
      for (uint32_t k = 0; k < nt && e == cudaSuccess; ++k) {
         memset(cand + (size_t)k * wpt, 0, (size_t)wpt * 4); memset(acan + (size_t)k * wpt, 0, (size_t)wpt * 4);
         e = cudaMemcpyAsync(cand + (size_t)k * wpt, d_mc + (size_t)k * ms, (size_t)nw * 4, cudaMemcpyDeviceToHost, t->stream);
         if (e == cudaSuccess) e = cudaMemcpyAsync(acan + (size_t)k * wpt, d_ma + (size_t)k * ms, (size_t)nw * 4, cudaMemcpyDeviceToHost, t->stream); }
      if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
      if (e == cudaSuccess && (nrows & 31)) {                    /* rows past the end of the data: not defined, cleared */
         const uint32_t keep = (1u << (nrows & 31)) - 1u;
         for (uint32_t k = 0; k < nt; ++k) { cand[(size_t)k * wpt + nw - 1] &= keep; acan[(size_t)k * wpt + nw - 1] &= keep; } } }
   cudaFree(d_mc); cudaFree(d_md); cudaFree(d_ma);
   if (e != cudaSuccess) return set_err(RT_ERR_CUDA, "rt_peak_masks: %s", cudaGetErrorString(e));
   return RT_OK; }

extern "C" int rt_bulk_last_unit(const rt_bulk *b, uint32_t ci, uint64_t *row0, uint64_t *row_end) {
   if (!b || ci >= b->cfgs.size()) return set_err(RT_ERR_ARG, "rt_bulk_last_unit: bad argument");
   const BulkCfg &bc = b->cfgs[ci];
   if (bc.last_unit >= bc.units.size()) return set_err(RT_ERR_STATE, "rt_bulk_last_unit: no successful lookup yet");
   if (row0) *row0 = bc.units[bc.last_unit].row0;
   if (row_end) *row_end = bc.units[bc.last_unit].row_end;
   return RT_OK; }

extern "C" int rt_bulk_tile_digest(rt_bulk *b, uint32_t ci, uint64_t period, uint64_t ntiles, uint64_t *events, uint64_t *digest, uint64_t *bad_times) {
   if (!b || ci >= b->cfgs.size() || !period || !ntiles || !events || !digest) return set_err(RT_ERR_ARG, "rt_bulk_tile_digest: bad argument");
   if (b->fetched || !b->d_pool) return set_err(RT_ERR_STATE, "rt_bulk_tile_digest: the results have left the device (call it before rt_bulk_fetch / rt_bulk_lookup)");
   rt_tape *t = b->tape; BulkCfg &bc = b->cfgs[ci];
   CU(cudaSetDevice(t->device));
   memset(events, 0, ntiles * 8); memset(digest, 0, ntiles * 8); if (bad_times) *bad_times = 0;
   if (!bc.nunits) return RT_OK;
   unsigned long long *d = nullptr;
   CU(cudaMallocAsync(&d, (2 * ntiles + 1) * 8, t->stream));
   cudaError_t e = cudaMemsetAsync(d, 0, (2 * ntiles + 1) * 8, t->stream);
   if (e == cudaSuccess) e = launch_tile_digest(bc.dc, bc.d_units, bc.nunits, bc.d_meta, b->d_pool, b->d_chunk_next, period, ntiles, d, d + ntiles, d + 2 * ntiles, t->sms, t->stream);
   ++t->launches;
   unsigned long long bad = 0;
   if (e == cudaSuccess) e = cudaMemcpyAsync(events, d, ntiles * 8, cudaMemcpyDeviceToHost, t->stream);
   if (e == cudaSuccess) e = cudaMemcpyAsync(digest, d + ntiles, ntiles * 8, cudaMemcpyDeviceToHost, t->stream);
   if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d + 2 * ntiles, 8, cudaMemcpyDeviceToHost, t->stream);
   if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
   cudaFreeAsync(d, t->stream);
   if (e != cudaSuccess) return set_err(RT_ERR_CUDA, "rt_bulk_tile_digest: %s", cudaGetErrorString(e));
   if (bad_times) *bad_times = bad;
   return RT_OK; }

extern "C" int rt_bulk_get_stats(const rt_bulk *b, rt_bulk_stats *out) {
   if (!b || !out) return set_err(RT_ERR_ARG, "rt_bulk_get_stats: null argument");
   *out = b->stats; return RT_OK; }

static void fill_unit_info(const rt_bulk *b, const BulkCfg &bc, size_t lo, uint64_t start_row, rt_unit_info *out) {
   const uint32_t nt = b->tape->desc.ntrks;
   memset(out, 0, sizeof *out);
   const UnitDesc &u = bc.units[lo];
   out->unit_index = lo; out->nunits = bc.units.size(); out->row0 = u.row0; out->row_end = u.row_end; out->ntrks = nt;
   const bool tz = rt_row_time(&b->tape->desc, start_row) == 0.0;
   for (uint32_t k = 0; k < nt; ++k) {
      const TrkMeta &m = bc.meta[lo * nt + k];
      out->first_event_row[k] = m.first_event_row; out->sync_row[k] = m.sync_row; out->last_loud_row[k] = m.last_loud_row;
      out->sync_early[k] = m.sync_early; out->loud_early[k] = m.loud_early;
      out->sync_first[k] = m.sync_first; out->quiet_from[k] = m.quiet_from; out->failed[k] = m.failed;
      out->nevents[k] = m.nevents; out->need_sync_row[k] = start_row + (uint64_t)fill_of(bc.dc, k, tz); } }

extern "C" int rt_bulk_unit_info(const rt_bulk *b, uint32_t ci, uint64_t start_row, rt_unit_info *out) {
   if (!b || ci >= b->cfgs.size() || !out) return set_err(RT_ERR_ARG, "rt_bulk_unit_info: bad argument");
   if (!b->fetched) { int rc = rt_bulk_fetch(const_cast<rt_bulk *>(b)); if (rc) return rc; }
   const BulkCfg &bc = b->cfgs[ci];
   memset(out, 0, sizeof *out);
   if (bc.units.empty()) return RT_MISS;
   size_t lo = 0, hi = bc.units.size();
   while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (bc.units[mid].row0 <= start_row) lo = mid; else hi = mid; }
   fill_unit_info(b, bc, lo, start_row, out);
   return RT_OK; }

extern "C" int rt_bulk_unit_at(const rt_bulk *b, uint32_t ci, uint64_t unit_index, rt_unit_info *out) {
   if (!b || ci >= b->cfgs.size() || !out) return set_err(RT_ERR_ARG, "rt_bulk_unit_at: bad argument");
   if (!b->fetched) { int rc = rt_bulk_fetch(const_cast<rt_bulk *>(b)); if (rc) return rc; }
   const BulkCfg &bc = b->cfgs[ci];
   memset(out, 0, sizeof *out);
   if (unit_index >= bc.units.size()) return RT_MISS;
   fill_unit_info(b, bc, (size_t)unit_index, bc.units[unit_index].row0, out);
   return RT_OK; }

/* Can unit `ui` stand in for a fresh RT_RESET_FULL at `start_row`?  The rules live in lookup_rules.h (DESIGN.md "unit equivalence"),
   which the CPU test harness compiles too. */
static bool unit_covers(const BulkCfg &bc, const rt_tape_desc &desc, uint32_t nt, size_t ui, uint64_t start_row, uint64_t *bridge_to = nullptr) {
   return rtlookup::unit_covers(bc.dc, bc.units[ui], &bc.meta[ui * nt], nt, start_row, rt_row_time(&desc, start_row) == 0.0, bridge_to); }

/* The bridge: unit `ui` has, on every track, a canonical row c_k (TrkMeta::sync_row: the window maximum has just left a full window,
   so the detector state is a pure function of the samples there) in front of its first event and far enough behind start_row for a
   scan reset at start_row to have a full window too -- but the rows in between are not provably quiet.  Then the exact scan is run
   over [start_row, max c_k]: if it finds no event on track k up to c_k, both scans reach c_k in default state (no event, no feedback)
   with the same window, and are identical from there on.  A few hundred rows of the exact scan instead of a whole block. */
static int bridge_holds(rt_bulk *b, uint32_t ci, size_t ui, uint64_t start_row, uint64_t upto, bool *ok) {
   BulkCfg &bc = b->cfgs[ci];
   const uint32_t nt = b->tape->desc.ntrks;
   const TrkMeta *m = &bc.meta[ui * nt];
   *ok = false;
   int rc;
   if (b->bridge && b->bridge_cfg != ci) { rt_scan_end(b->bridge); b->bridge = nullptr; }
   if (!b->bridge) { rc = rt_scan_begin(b->tape, &bc.cfg, &b->bridge); if (rc) return rc; b->bridge_cfg = ci; }
   rc = rt_scan_reset(b->bridge, RT_RESET_FULL, start_row); if (rc) return rc;
   const rt_event *ev = nullptr; uint64_t n = 0, done = 0;
   rc = rt_scan_run(b->bridge, upto - start_row + 1, &ev, &n, &done); if (rc) return rc;
   if (done != upto - start_row + 1) return RT_OK;
   for (uint64_t i = 0; i < n; ++i) if (ev[i].row <= m[ev[i].trk].sync_row) return RT_OK;       /* an event in front of the canonical row */
   *ok = true; ++b->n_bridged;
   return RT_OK; }

static bool unit_tail_covers(const BulkCfg &bc, uint32_t nt, size_t ui, uint64_t start_row) {      /* the tail rule, lookup_rules.h */
   return rtlookup::unit_tail_covers(bc.units[ui], &bc.meta[ui * nt], nt, start_row); }

extern "C" int rt_bulk_lookup(rt_bulk *b, uint32_t ci, uint64_t start_row, const rt_event **events, uint64_t *nevents, uint64_t *valid_rows) {
   if (!b || ci >= b->cfgs.size()) return set_err(RT_ERR_ARG, "rt_bulk_lookup: bad argument");
   if (!b->fetched) { int rc = rt_bulk_fetch(b); if (rc) return rc; }
   BulkCfg &bc = b->cfgs[ci];
   const uint32_t nt = b->tape->desc.ntrks;
   if (bc.units.empty()) return RT_MISS;
   /* candidates: the last unit that starts at or before start_row, and the one after it (the reference's
      reset row may lie a little BEFORE the row the unit finder picked) */
   size_t lo = 0, hi = bc.units.size();
   while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (bc.units[mid].row0 <= start_row) lo = mid; else hi = mid; }
   uint64_t br0 = RT_NOROW, br1 = RT_NOROW;
   if (!unit_covers(bc, b->tape->desc, nt, lo, start_row, &br0)) {
      bool bridged = false;
      if (lo + 1 < bc.units.size() && unit_covers(bc, b->tape->desc, nt, lo + 1, start_row, &br1)) { ++lo; bridged = true; }   /* covered by the next unit */
      else if (br0 != RT_NOROW || br1 != RT_NOROW) {             /* not provably quiet: a short exact scan can still prove the equivalence */
         const char *env = getenv("RT_BRIDGE");
         if (!(env && env[0] == '0')) {
            if (br0 != RT_NOROW) { int rc = bridge_holds(b, ci, lo, start_row, br0, &bridged); if (rc) return rc; }
            if (!bridged && br1 != RT_NOROW) { int rc = bridge_holds(b, ci, lo + 1, start_row, br1, &bridged); if (rc) return rc; if (bridged) ++lo; } } }
      if (bridged) { /* covered */ }
      else if (unit_tail_covers(bc, nt, lo, start_row)) {        /* nothing up to the end of this unit (typically: the end of the tape) */
         b->result.clear(); bc.last_unit = lo;
         if (events) *events = b->result.data();
         if (nevents) *nevents = 0;
         if (valid_rows) *valid_rows = bc.units[lo].row_end - start_row;
         return RT_OK; }
      else return RT_MISS; }
   const TrkMeta *m = &bc.meta[lo * nt];
   bc.last_unit = lo;
   /* Chaining: while the unit holds no event at all, the reference's scan passes through it unchanged; it is
      then identical to the NEXT unit's fresh scan from that unit's first canonical row on, provided that row
      lies inside the stretch where this unit has already shown the scan to be event-free (the overlap). */
   for (uint64_t from = start_row; lo + 1 < bc.units.size() && rtlookup::chains_into_next(bc.dc, bc.units[lo], m, &bc.meta[(lo + 1) * nt], nt, from); ) {
      ++lo; m = &bc.meta[lo * nt]; from = bc.units[lo].row0; }      /* the passing scan equals this unit's fresh scan from its first canonical row on */
   const UnitDesc &ue = bc.units[lo];
   /* merge the per-track chunk chains into (row, trk) order */
   struct CC { uint32_t chunk, left, slot; } cc[RT_MAXTRKS];
   size_t total = 0;
   for (uint32_t k = 0; k < nt; ++k) { cc[k].chunk = m[k].first_chunk; cc[k].left = m[k].nevents; cc[k].slot = 0; total += m[k].nevents; }
   b->result.clear(); b->result.reserve(total);
   for (size_t n = 0; n < total; ++n) {
      int best = -1; uint64_t brow = ~0ull;
      for (uint32_t k = 0; k < nt; ++k) if (cc[k].left) {
            const rt_event &e = b->h_pool[(size_t)cc[k].chunk * RT_EVC + cc[k].slot];
            if (e.row < brow) { brow = e.row; best = (int)k; } }
      CC &c = cc[best];
      b->result.push_back(b->h_pool[(size_t)c.chunk * RT_EVC + c.slot]);
      --c.left;
      if (++c.slot == RT_EVC) { c.slot = 0; c.chunk = c.left ? b->chunk_next[c.chunk] : RT_NOCHUNK; } }
   if (events) *events = b->result.data();
   if (nevents) *nevents = b->result.size();
   if (valid_rows) *valid_rows = ue.row_end - start_row;
   return RT_OK; }
