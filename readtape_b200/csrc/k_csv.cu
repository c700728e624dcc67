/* readtape_b200/csrc/k_csv.cu -- CSV ingest on the device: the conversion the reference's csvtbin tool does line by
 * line on the CPU (write_tbin, src/csvtbin.c:661-747; csv_preread :619-657; scanfast_float / scanfast_double :403-433).
 *
 * A Saleae-style capture is text: two title lines, then one line per sample, "time, v0, v1, ... v(n-1)".  It is ~10x the size of the
 * TBIN it becomes, and csvtbin parses it with one fgets + n scanfast_float calls per line.  Here the text is put on the device once:
 *   k_csv_count / k_csv_index   where every line starts (each warp owns 128 KB of text and reads it 512 coalesced bytes at a time; the
 *                               newline counts are prefix-summed across warps, then the chunk is read again to write the offsets)
 *   k_csv_maxabs                csv_preread's maximum |voltage * scalefactor| over the first lines (the TBIN's maxvolts comes from it)
 *   k_csv_parse                 one thread per kept line: skip the time stamp, parse n voltages, scale, permute, invert, quantise to
 *                               int16 exactly as csvtbin.c:703-715 does, write the TBIN payload row
 * The number parser is the reference's, operation for operation, in float (no FMA: the library is built with -fmad=false): the
 * digits accumulate as n = n*10 + d, the fraction as n += d / divisor with divisor *= 10 -- so every parsed voltage, and therefore
 * every int16 sample, is bit-identical to csvtbin's (tests/test_csv.py compares whole payloads).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.h"

#define CSV_THREADS 256
#define CSV_WARP_BYTES (128u << 10)                         /* text per warp in the indexing kernels: 256 coalesced 512-byte reads */
#define CSV_WARPS (CSV_THREADS / 32)

/* newlines among the 16 bytes at p (16-byte aligned), as a 16-bit mask; bytes at or beyond `n` do not count */
__device__ __forceinline__ uint32_t newline_mask16(const char *txt, uint64_t at, uint64_t n) {
   if (at >= n) return 0;
   const uint4 v = *reinterpret_cast<const uint4 *>(txt + at);               /* the buffer is padded to a multiple of 16 bytes */
   const uint32_t w[4] = {v.x, v.y, v.z, v.w};
   uint32_t m = 0;
#pragma unroll
   for (int k = 0; k < 4; ++k) {
      const uint32_t e = __vcmpeq4(w[k], 0x0a0a0a0au) & 0x01010101u;          /* bit 0 of each byte that is a newline */
      m |= ((e | e >> 7 | e >> 14 | e >> 21) & 0xfu) << (4 * k); }
   if (n - at < 16) m &= (1u << (n - at)) - 1;
   return m; }

/* one warp per 128 KB of text: warp_counts[w] = newlines in its chunk */
__global__ void __launch_bounds__(CSV_THREADS)
k_csv_count(const char *txt, uint64_t n, uint32_t *warp_counts) {
   const uint64_t w = (uint64_t)blockIdx.x * CSV_WARPS + (threadIdx.x >> 5);
   const uint64_t lo = w * CSV_WARP_BYTES;
   const int lane = threadIdx.x & 31;
   uint32_t c = 0;
   if (lo < n)
      for (uint32_t i = 0; i < CSV_WARP_BYTES; i += 512) c += __popc(newline_mask16(txt, lo + i + lane * 16, n));
   for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
   if (lane == 0) warp_counts[w] = c; }

__global__ void k_csv_scan(const uint32_t *warp_counts, uint32_t nwarps, uint64_t *warp_offsets, uint64_t *total) {
   /* one warp: chunks of 32 counts, shuffle prefix, running base */
   const int lane = threadIdx.x;
   uint64_t run = 0;
   for (uint32_t i = 0; i < nwarps; i += 32) {
      const uint32_t c = i + lane < nwarps ? warp_counts[i + lane] : 0;
      uint32_t incl = c;
      for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
      if (i + lane < nwarps) warp_offsets[i + lane] = run + incl - c;
      run += __shfl_sync(0xffffffffu, incl, 31); }
   if (lane == 0) *total = run; }

/* line_start[k] = offset of the first character of line k (line 0 starts at 0); every newline starts the next line */
__global__ void __launch_bounds__(CSV_THREADS)
k_csv_index(const char *txt, uint64_t n, const uint64_t *warp_offsets, uint64_t *line_start, uint64_t cap) {
   const uint64_t w = (uint64_t)blockIdx.x * CSV_WARPS + (threadIdx.x >> 5);
   const uint64_t lo = w * CSV_WARP_BYTES;
   const int lane = threadIdx.x & 31;
   if (lo >= n) return;
   uint64_t at = warp_offsets[w] + 1;                         /* index of the line the next newline starts */
   for (uint32_t i = 0; i < CSV_WARP_BYTES; i += 512) {
      const uint64_t pos = lo + i + lane * 16;
      uint32_t m = newline_mask16(txt, pos, n);
      const uint32_t c = __popc(m);
      uint32_t incl = c;
      for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
      uint64_t k = at + incl - c;
      while (m) { const int b = __ffs(m) - 1; m &= m - 1; if (k < cap) line_start[k] = pos + b + 1; ++k; }
      at += __shfl_sync(0xffffffffu, incl, 31); } }

/* ---- the reference's number scanner, csvtbin.c:403-417, on a bounded line ------------------------------------------------------ */
struct Cursor { const char *p, *end; __device__ int ch() const { return p < end ? (unsigned char)*p : 0; } };
__device__ __forceinline__ bool is_digit(int c) { return c >= '0' && c <= '9'; }

__device__ inline float scan_float(Cursor &c) {
   float n = 0;
   bool negative = false;
   while (c.ch() == ' ' || c.ch() == ',') ++c.p;
   if (c.ch() == '-') { ++c.p; negative = true; }
   while (is_digit(c.ch())) { n = n * 10 + (float)(c.ch() - '0'); ++c.p; }
   if (c.ch() == '.') {
      float divisor = 10;
      ++c.p;
      while (is_digit(c.ch())) { n += (float)(c.ch() - '0') / divisor; divisor *= 10; ++c.p; } }
   return negative ? -n : n; }

__device__ inline void skip_number(Cursor &c) {                /* scanfast_double with the value discarded (csvtbin.c:691) */
   while (c.ch() == ' ' || c.ch() == ',') ++c.p;
   if (c.ch() == '-') ++c.p;
   while (is_digit(c.ch())) ++c.p;
   if (c.ch() == '.') { ++c.p; while (is_digit(c.ch())) ++c.p; } }

__device__ __forceinline__ Cursor line_cursor(const char *txt, uint64_t n, const uint64_t *line_start, uint64_t nlines_total, uint64_t line) {
   const uint64_t lo = line_start[line];
   uint64_t hi = line + 1 <= nlines_total ? line_start[line + 1] : n;
   if (hi > n) hi = n;
   if (hi - lo > 399) hi = lo + 399;                           /* fgets(line, MAXLINE = 400): a longer line is cut there */
   return Cursor{txt + lo, txt + hi}; }

__device__ __forceinline__ void atomic_max_float_pos(float *addr, float v) { atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v)); }   /* v >= 0 */

/* ---- the same scanner without per-character bounds checks and without divisions --------------------------------------------------
 * A line of at most 399 characters that ends in '\n' (rt_csv_open puts newlines after the text, so the last line does too) needs
 * no end pointer: '\n' is neither blank, comma, sign, digit nor point, so every loop of the scanner stops at it, and a scan that
 * starts at it returns 0 without moving -- what the reference's scanner does at the '\n' / NUL of its fgets buffer.
 * The fraction term  digit / divisor  takes one of 10 values per decimal place; frac[10 * k + d] holds (float)d / divisor_k with
 * divisor_0 = 10, divisor_(k+1) = divisor_k * 10 computed in float as the reference's loop does (from 10^39 on the divisor is
 * +inf and the term 0: rows 38 and 39 are zeros and the index stays on row 39).  Same IEEE operations, done once per block. */
#define CSV_FRAC_ROWS 40

__device__ inline void fill_frac_table(float *frac) {
   for (int i = threadIdx.x; i < CSV_FRAC_ROWS * 10; i += blockDim.x) {
      const int k = i / 10, d = i - 10 * k;
      float divisor = 10;
      for (int j = 0; j < k; ++j) divisor *= 10;
      frac[i] = (float)d / divisor; }
   __syncthreads(); }

__device__ __forceinline__ float digit_value(unsigned d) { return __int_as_float(0x4B000000u | d) - 8388608.0f; }   /* (float)d, exactly, on the FP32 pipe */

__device__ __forceinline__ float scan_float_fast(const unsigned char *&p, const float *frac) {
   float n = 0;
   bool negative = false;
   unsigned c = *p, d;
   while (c == ' ' || c == ',') c = *++p;
   if (c == '-') { negative = true; c = *++p; }
   while ((d = c - '0') < 10u) { n = n * 10 + digit_value(d); c = *++p; }
   if (c == '.') {
      int k = 0;
      c = *++p;
      while ((d = c - '0') < 10u) { n += frac[k + d]; k = k < 10 * (CSV_FRAC_ROWS - 1) ? k + 10 : k; c = *++p; } }
   return negative ? -n : n; }

__device__ __forceinline__ void skip_number_fast(const unsigned char *&p) {
   unsigned c = *p;
   while (c == ' ' || c == ',') c = *++p;
   if (c == '-') c = *++p;
   while (c - '0' < 10u) c = *++p;
   if (c == '.') { c = *++p; while (c - '0' < 10u) c = *++p; } }

/* the line's first character, and whether the unchecked scanner may be used on it */
__device__ __forceinline__ bool line_is_plain(const uint64_t *line_start, uint64_t line, uint64_t *lo) {
   *lo = line_start[line];
   return line_start[line + 1] - *lo <= 399; }                   /* with its '\n' (a real one, or the padding's) */

__global__ void __launch_bounds__(CSV_THREADS)
k_csv_maxabs(const char *txt, uint64_t n, const uint64_t *line_start, uint64_t nlines_total, uint64_t first_line, uint64_t nlines,
             int ntrks, float scalefactor, float *out) {
   __shared__ float frac[CSV_FRAC_ROWS * 10];
   fill_frac_table(frac);
   const uint64_t i = (uint64_t)blockIdx.x * CSV_THREADS + threadIdx.x;
   float m = 0;
   if (i < nlines) {
      uint64_t lo;
      if (line_is_plain(line_start, first_line + i, &lo)) {
         const unsigned char *p = reinterpret_cast<const unsigned char *>(txt) + lo;
         skip_number_fast(p);
         for (int k = 0; k < ntrks; ++k) { float v = scan_float_fast(p, frac) * scalefactor; if (v < 0) v = -v; if (m < v) m = v; } }
      else {
         Cursor c = line_cursor(txt, n, line_start, nlines_total, first_line + i);
         skip_number(c);
         for (int k = 0; k < ntrks; ++k) { float v = scan_float(c) * scalefactor; if (v < 0) v = -v; if (m < v) m = v; } } }
   for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
   if ((threadIdx.x & 31) == 0 && m > 0) atomic_max_float_pos(out, m); }

struct CsvParseArgs {
   int ntrks; uint32_t perm[RT_MAXTRKS]; float maxvolts, scalefactor; int invert; uint32_t subsample; uint64_t first_line; };

/* csvtbin.c:702-715 for one sample */
__device__ __forceinline__ int16_t quantise(float fsample, const CsvParseArgs &a, unsigned &big, unsigned &small_, float &vmin, float &vmax) {
   if (a.invert) fsample = -fsample;
   const float round = fsample < 0 ? -0.5f : 0.5f;
   int sample = (int)((fsample / a.maxvolts * 32767) + round);
   if (fsample < vmin) vmin = fsample;
   if (fsample > vmax) vmax = fsample;
   if (sample <= -32767) { sample = -32767; ++small_; }
   if (sample >= 32767) { sample = 32767; ++big; }
   return (int16_t)sample; }

/* stats: [0] too big, [1] too small (u64); fstats: [0] max volts (>= 0), [1] -min volts (>= 0) */
template <bool IDENTITY>
__global__ void __launch_bounds__(CSV_THREADS)
k_csv_parse(const char *txt, uint64_t n, const uint64_t *line_start, uint64_t nlines_total, const __grid_constant__ CsvParseArgs a, uint64_t nrows,
            int16_t *rows, unsigned long long *stats, float *fstats) {
   __shared__ float frac[CSV_FRAC_ROWS * 10];
   fill_frac_table(frac);
   const uint64_t r = (uint64_t)blockIdx.x * CSV_THREADS + threadIdx.x;
   unsigned big = 0, small_ = 0; float vmax = 0, vmin = 0;
   if (r < nrows) {
      /* csvtbin.c:686: of every `subsample` lines the LAST one is used */
      const uint64_t line = a.first_line + (r + 1) * a.subsample - 1;
      int16_t *out = rows + r * (uint64_t)a.ntrks;
      uint64_t lo;
      if (line_is_plain(line_start, line, &lo)) {
         const unsigned char *p = reinterpret_cast<const unsigned char *>(txt) + lo;
         skip_number_fast(p);
         if (IDENTITY) {                                        /* no -order: column k is position k, nothing to hold back */
#pragma unroll 1
            for (int k = 0; k < a.ntrks; ++k) out[k] = quantise(scan_float_fast(p, frac) * a.scalefactor, a, big, small_, vmin, vmax); }
         else {
            float samples[RT_MAXTRKS];
#pragma unroll 1
            for (int k = 0; k < a.ntrks; ++k) samples[a.perm[k]] = scan_float_fast(p, frac) * a.scalefactor;
#pragma unroll 1
            for (int k = 0; k < a.ntrks; ++k) out[k] = quantise(samples[k], a, big, small_, vmin, vmax); } }
      else {
         Cursor c = line_cursor(txt, n, line_start, nlines_total, line);
         skip_number(c);
         float samples[RT_MAXTRKS];
#pragma unroll 1
         for (int k = 0; k < a.ntrks; ++k) samples[a.perm[k]] = scan_float(c) * a.scalefactor;
#pragma unroll 1
         for (int k = 0; k < a.ntrks; ++k) out[k] = quantise(samples[k], a, big, small_, vmin, vmax); } }
   for (int d = 16; d > 0; d >>= 1) {
      big += __shfl_xor_sync(0xffffffffu, big, d); small_ += __shfl_xor_sync(0xffffffffu, small_, d);
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d)); vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d)); }
   if ((threadIdx.x & 31) == 0) {
      if (big) atomicAdd(&stats[0], (unsigned long long)big);
      if (small_) atomicAdd(&stats[1], (unsigned long long)small_);
      if (vmax > 0) atomic_max_float_pos(&fstats[0], vmax);
      if (vmin < 0) atomic_max_float_pos(&fstats[1], -vmin); } }

/* ---- launchers ------------------------------------------------------------------------------------------------------------------ */
uint32_t csv_index_warps(uint64_t nbytes) { return (uint32_t)((nbytes + CSV_WARP_BYTES - 1) / CSV_WARP_BYTES); }
static uint32_t csv_index_blocks(uint64_t nbytes) { return (csv_index_warps(nbytes) + CSV_WARPS - 1) / CSV_WARPS; }

/* warp_counts / warp_offsets: csv_index_blocks * CSV_WARPS entries (round csv_index_warps up to a multiple of 8) */
cudaError_t launch_csv_count(const char *txt, uint64_t n, uint32_t *warp_counts, uint64_t *warp_offsets, uint64_t *total, cudaStream_t s) {
   const uint32_t nb = csv_index_blocks(n);
   if (!nb) return cudaSuccess;
   k_csv_count<<<nb, CSV_THREADS, 0, s>>>(txt, n, warp_counts);
   k_csv_scan<<<1, 32, 0, s>>>(warp_counts, nb * CSV_WARPS, warp_offsets, total);
   return cudaGetLastError(); }

cudaError_t launch_csv_index(const char *txt, uint64_t n, const uint64_t *warp_offsets, uint64_t *line_start, uint64_t cap, cudaStream_t s) {
   const uint32_t nb = csv_index_blocks(n);
   if (!nb) return cudaSuccess;
   k_csv_index<<<nb, CSV_THREADS, 0, s>>>(txt, n, warp_offsets, line_start, cap);
   return cudaGetLastError(); }

cudaError_t launch_csv_maxabs(const char *txt, uint64_t n, const uint64_t *line_start, uint64_t nlines_total, uint64_t first_line, uint64_t nlines,
                              int ntrks, float scalefactor, float *out, cudaStream_t s) {
   if (!nlines) return cudaSuccess;
   k_csv_maxabs<<<(unsigned)((nlines + CSV_THREADS - 1) / CSV_THREADS), CSV_THREADS, 0, s>>>(txt, n, line_start, nlines_total, first_line, nlines, ntrks, scalefactor, out);
   return cudaGetLastError(); }

cudaError_t launch_csv_parse(const char *txt, uint64_t n, const uint64_t *line_start, uint64_t nlines_total, int ntrks, const uint32_t *perm,
                             float maxvolts, float scalefactor, int invert, uint32_t subsample, uint64_t first_line, uint64_t nrows,
                             int16_t *rows, unsigned long long *stats, float *fstats, cudaStream_t s) {
   if (!nrows) return cudaSuccess;
   CsvParseArgs a; a.ntrks = ntrks; a.maxvolts = maxvolts; a.scalefactor = scalefactor; a.invert = invert; a.subsample = subsample; a.first_line = first_line;
   for (int k = 0; k < RT_MAXTRKS; ++k) a.perm[k] = k < ntrks ? perm[k] : 0;
   bool identity = true;
   for (int k = 0; k < ntrks; ++k) identity = identity && perm[k] == (uint32_t)k;
   const unsigned nb = (unsigned)((nrows + CSV_THREADS - 1) / CSV_THREADS);
   if (identity) k_csv_parse<true><<<nb, CSV_THREADS, 0, s>>>(txt, n, line_start, nlines_total, a, nrows, rows, stats, fstats);
   else k_csv_parse<false><<<nb, CSV_THREADS, 0, s>>>(txt, n, line_start, nlines_total, a, nrows, rows, stats, fstats);
   return cudaGetLastError(); }
