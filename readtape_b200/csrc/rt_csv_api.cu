/* readtape_b200/csrc/rt_csv_api.cu -- the C-ABI of include/rt_csv.h on the device: text upload, line index, csv_preread's
 * maximum and the conversion loop (kernels in k_csv.cu).  No CPU fallback: without a device rt_csv_open fails. */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/rt_csv.h"
#include "kernels.h"
#include "rt_internal.h"

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
      return rt_fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct rt_csv {
   int device = 0;
   cudaStream_t stream = nullptr;
   char *d_text = nullptr; uint64_t nbytes = 0;
   uint64_t *d_line_start = nullptr; uint64_t nlines = 0;      /* nlines + 1 entries: the last one is nbytes (+1 when the text ends without a newline) */
   bool open_tail = false;                                      /* the text does not end in '\n': its last line has no terminator */
   ~rt_csv() {
      if (d_text) cudaFree(d_text);
      if (d_line_start) cudaFree(d_line_start);
      if (stream) cudaStreamDestroy(stream); } };

/* pageable host memory -> device through a ring of pinned slots filled by a few threads (the same scheme as rt_upload).  Pinning
   192 MB costs tens of milliseconds, so the ring is allocated once per process and kept (one upload at a time uses it). */
static std::mutex g_ring_mu;
static char *g_ring = nullptr;
static int upload_text(rt_csv *c, const char *text, uint64_t nbytes) {
   const size_t slot = 16u << 20; const int NB = 12;
   if (nbytes <= slot) { CU(cudaMemcpyAsync(c->d_text, text, nbytes, cudaMemcpyHostToDevice, c->stream)); CU(cudaStreamSynchronize(c->stream)); return RT_OK; }
   std::lock_guard<std::mutex> ring_lock(g_ring_mu);
   if (!g_ring) CU(cudaHostAlloc(&g_ring, slot * NB, cudaHostAllocPortable));
   char *ring = g_ring; cudaEvent_t done[NB];
   for (int i = 0; i < NB; ++i) cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
   const uint64_t nchunks = (nbytes + slot - 1) / slot;
   const int hw = (int)std::thread::hardware_concurrency();
   const int nthreads = std::max(1, std::min({10, hw - 2, (int)nchunks}));
   std::mutex mu; std::condition_variable cv;
   std::vector<char> filled(nchunks, 0);
   uint64_t released = 0; std::atomic<uint64_t> next{0};
   auto producer = [&]() {
      for (;;) {
         const uint64_t i = next.fetch_add(1);
         if (i >= nchunks) return;
         { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return i < released + NB; }); }
         memcpy(ring + (i % NB) * slot, text + i * slot, (size_t)std::min<uint64_t>(slot, nbytes - i * slot));
         { std::lock_guard<std::mutex> lk(mu); filled[i] = 1; }
         cv.notify_all(); } };
   std::vector<std::thread> th;
   for (int k = 0; k < nthreads; ++k) th.emplace_back(producer);
   cudaError_t e = cudaSuccess;
   for (uint64_t i = 0; i < nchunks; ++i) {
      { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return filled[i] != 0; }); }
      if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_text + i * slot, ring + (i % NB) * slot, (size_t)std::min<uint64_t>(slot, nbytes - i * slot), cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) e = cudaEventRecord(done[i % NB], c->stream);
      if (i + 2 >= (uint64_t)NB) {
         const uint64_t k = i + 2 - NB;
         if (e == cudaSuccess) e = cudaEventSynchronize(done[k % NB]);
         { std::lock_guard<std::mutex> lk(mu); released = k + 1; }
         cv.notify_all(); } }
   { std::lock_guard<std::mutex> lk(mu); released = nchunks; }
   cv.notify_all();
   for (auto &x : th) x.join();
   if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
   for (int i = 0; i < NB; ++i) cudaEventDestroy(done[i]);
   if (e != cudaSuccess) return rt_fail(RT_ERR_CUDA, "rt_csv_open: text upload failed: %s", cudaGetErrorString(e));
   return RT_OK; }

extern "C" int rt_csv_open(int device, const char *text, uint64_t nbytes, rt_csv **out) {
   if (!out || (!text && nbytes)) return rt_fail(RT_ERR_ARG, "rt_csv_open: null argument");
   *out = nullptr;
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return rt_fail(RT_ERR_NODEVICE, "rt_csv_open: no CUDA device (this library has no CPU fallback)"); }
   if (device < 0 || device >= ndev) return rt_fail(RT_ERR_ARG, "rt_csv_open: device %d of %d", device, ndev);
   CU(cudaSetDevice(device));
   rt_csv *c = new rt_csv; c->device = device; c->nbytes = nbytes;
   int rc = RT_OK;
   auto bail = [&](int code) { delete c; return code; };
   if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(rt_fail(RT_ERR_CUDA, "rt_csv_open: no stream"));
   if (cudaMalloc(&c->d_text, nbytes + 32) != cudaSuccess) { cudaGetLastError(); return bail(rt_fail(RT_ERR_NOMEM, "rt_csv_open: %llu bytes of device memory for the text", (unsigned long long)nbytes)); }
   /* the 16..32 bytes after the text are newlines: the indexing kernels read whole 16-byte words, and a last line without a
      terminator ends at one for the parser */
   if (cudaMemsetAsync(c->d_text + nbytes, '\n', 32, c->stream) != cudaSuccess) return bail(rt_fail(RT_ERR_CUDA, "rt_csv_open: memset"));
   rc = upload_text(c, text, nbytes); if (rc) return bail(rc);
   const uint32_t nb = (csv_index_warps(nbytes) + 7) / 8 * 8;          /* scratch entries: one per warp of the indexing kernels */
   uint32_t *d_cnt = nullptr; uint64_t *d_off = nullptr, *d_total = nullptr;
   uint64_t newlines = 0;
   if (nb) {
      if (cudaMalloc(&d_cnt, nb * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&d_off, (nb + 1) * sizeof(uint64_t)) != cudaSuccess) { cudaGetLastError(); cudaFree(d_cnt); return bail(rt_fail(RT_ERR_NOMEM, "rt_csv_open: index scratch")); }
      d_total = d_off + nb;
      cudaError_t e = launch_csv_count(c->d_text, nbytes, d_cnt, d_off, d_total, c->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(&newlines, d_total, sizeof newlines, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) { cudaFree(d_cnt); cudaFree(d_off); return bail(rt_fail(RT_ERR_CUDA, "rt_csv_open: line count: %s", cudaGetErrorString(e))); } }
   c->open_tail = nbytes > 0 && text[nbytes - 1] != '\n';
   c->nlines = newlines + (c->open_tail ? 1 : 0);
   if (cudaMalloc(&c->d_line_start, (c->nlines + 2) * sizeof(uint64_t)) != cudaSuccess) { cudaGetLastError(); cudaFree(d_cnt); cudaFree(d_off); return bail(rt_fail(RT_ERR_NOMEM, "rt_csv_open: line index")); }
   {
      const uint64_t zero = 0;
      cudaError_t e = cudaMemcpyAsync(c->d_line_start, &zero, sizeof zero, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess && nb) e = launch_csv_index(c->d_text, nbytes, d_off, c->d_line_start, c->nlines + 2, c->stream);
      /* the entry after the last line: where a following line would start (one past a missing final newline) */
      const uint64_t end = nbytes + (c->open_tail ? 1 : 0);
      if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_line_start + c->nlines, &end, sizeof end, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      cudaFree(d_cnt); cudaFree(d_off);
      if (e != cudaSuccess) return bail(rt_fail(RT_ERR_CUDA, "rt_csv_open: line index: %s", cudaGetErrorString(e))); }
   *out = c;
   return RT_OK; }

extern "C" void rt_csv_close(rt_csv *c) { if (c) { cudaSetDevice(c->device); delete c; } }
extern "C" uint64_t rt_csv_nlines(const rt_csv *c) { return c ? c->nlines : 0; }

extern "C" int rt_csv_line(const rt_csv *c, uint64_t line, uint64_t *offset, uint64_t *length) {
   if (!c || !offset || !length) return rt_fail(RT_ERR_ARG, "rt_csv_line: null argument");
   if (line >= c->nlines) return rt_fail(RT_ERR_ARG, "rt_csv_line: line %llu of %llu", (unsigned long long)line, (unsigned long long)c->nlines);
   uint64_t se[2];
   CU(cudaSetDevice(c->device));
   CU(cudaMemcpy(se, c->d_line_start + line, sizeof se, cudaMemcpyDeviceToHost));
   *offset = se[0]; *length = se[1] - se[0] - 1;                  /* without the '\n' (the virtual one of an unterminated last line) */
   return RT_OK; }

extern "C" int rt_csv_max_abs(rt_csv *c, uint64_t first_line, uint64_t nlines, uint32_t ntrks, float scalefactor, float *max_abs) {
   if (!c || !max_abs || ntrks < 1 || ntrks > RT_MAXTRKS) return rt_fail(RT_ERR_ARG, "rt_csv_max_abs: bad argument");
   if (first_line > c->nlines || nlines > c->nlines - first_line) return rt_fail(RT_ERR_ARG, "rt_csv_max_abs: lines [%llu, +%llu) of %llu", (unsigned long long)first_line, (unsigned long long)nlines, (unsigned long long)c->nlines);
   CU(cudaSetDevice(c->device));
   float *d = nullptr;
   CU(cudaMalloc(&d, sizeof(float)));
   cudaError_t e = cudaMemsetAsync(d, 0, sizeof(float), c->stream);
   if (e == cudaSuccess) e = launch_csv_maxabs(c->d_text, c->nbytes, c->d_line_start, c->nlines, first_line, nlines, (int)ntrks, scalefactor, d, c->stream);
   if (e == cudaSuccess) e = cudaMemcpyAsync(max_abs, d, sizeof(float), cudaMemcpyDeviceToHost, c->stream);
   if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
   cudaFree(d);
   if (e != cudaSuccess) return rt_fail(RT_ERR_CUDA, "rt_csv_max_abs: %s", cudaGetErrorString(e));
   return RT_OK; }

extern "C" int rt_csv_convert(rt_csv *c, const rt_csv_cfg *cfg, uint64_t first_line, uint64_t nrows, int16_t *rows_out, rt_tape *tape, rt_csv_stats *stats) {
   if (!c || !cfg) return rt_fail(RT_ERR_ARG, "rt_csv_convert: null argument");
   if (cfg->ntrks < 1 || cfg->ntrks > RT_MAXTRKS || cfg->subsample < 1 || !(cfg->maxvolts > 0)) return rt_fail(RT_ERR_ARG, "rt_csv_convert: bad configuration");
   uint32_t seen = 0;
   for (uint32_t k = 0; k < cfg->ntrks; ++k) { if (cfg->track_permutation[k] >= cfg->ntrks) return rt_fail(RT_ERR_ARG, "rt_csv_convert: track_permutation[%u] = %u", k, cfg->track_permutation[k]); seen |= 1u << cfg->track_permutation[k]; }
   if (seen + 1 != 1u << cfg->ntrks) return rt_fail(RT_ERR_ARG, "rt_csv_convert: track_permutation is not a permutation");
   if (first_line > c->nlines || nrows > (c->nlines - first_line) / cfg->subsample) return rt_fail(RT_ERR_ARG, "rt_csv_convert: %llu rows from line %llu need more than the %llu lines of the text", (unsigned long long)nrows, (unsigned long long)first_line, (unsigned long long)c->nlines);
   if (stats) memset(stats, 0, sizeof *stats);
   if (!nrows) return RT_OK;
   CU(cudaSetDevice(c->device));
   int16_t *d_rows = nullptr; unsigned long long *d_stats = nullptr;
   const size_t row_bytes = (size_t)nrows * cfg->ntrks * 2;
   if (cudaMalloc(&d_rows, row_bytes + 16) != cudaSuccess) { cudaGetLastError(); return rt_fail(RT_ERR_NOMEM, "rt_csv_convert: %zu bytes for the rows", row_bytes); }
   if (cudaMalloc(&d_stats, 32) != cudaSuccess) { cudaGetLastError(); cudaFree(d_rows); return rt_fail(RT_ERR_NOMEM, "rt_csv_convert: stats"); }
   float *d_f = reinterpret_cast<float *>(d_stats + 2);
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   cudaError_t e = cudaMemsetAsync(d_stats, 0, 32, c->stream);
   if (e == cudaSuccess) e = cudaEventRecord(e0, c->stream);
   if (e == cudaSuccess) e = launch_csv_parse(c->d_text, c->nbytes, c->d_line_start, c->nlines, (int)cfg->ntrks, cfg->track_permutation, cfg->maxvolts, cfg->scalefactor,
                                              cfg->invert != 0, cfg->subsample, first_line, nrows, d_rows, d_stats, d_f, c->stream);
   if (e == cudaSuccess) e = cudaEventRecord(e1, c->stream);
   unsigned long long h_stats[4] = {0, 0, 0, 0};
   if (e == cudaSuccess) e = cudaMemcpyAsync(h_stats, d_stats, 32, cudaMemcpyDeviceToHost, c->stream);
   if (e == cudaSuccess && rows_out) e = cudaMemcpyAsync(rows_out, d_rows, row_bytes, cudaMemcpyDeviceToHost, c->stream);
   if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
   float ms = 0; if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   int rc = RT_OK;
   if (e != cudaSuccess) rc = rt_fail(RT_ERR_CUDA, "rt_csv_convert: %s", cudaGetErrorString(e));
   if (rc == RT_OK && tape) rc = rt_attach_device(tape, d_rows, nrows);        /* ingests on the tape's stream and drains it */
   cudaFree(d_rows); cudaFree(d_stats);
   if (rc == RT_OK && stats) {
      float f[2]; memcpy(f, &h_stats[2], sizeof f);
      stats->rows = nrows; stats->too_big = h_stats[0]; stats->too_small = h_stats[1];
      stats->maxvolts = f[0]; stats->minvolts = -f[1]; stats->ms_convert = ms; }
   return rc; }
