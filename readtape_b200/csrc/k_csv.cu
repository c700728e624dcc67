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
#include "tma_1d.cuh"

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

/* ---- the conversion kernel ------------------------------------------------------------------------------------------------------
 * One block = CSV_PB consecutive rows = (without -subsample) one contiguous piece of text, ~14 KB for a 9-track export.  The bulk
 * copy engine brings that piece into shared memory in one cp.async.bulk (coalesced, no L1 thrash from 128 threads walking 128
 * different cache lines a byte at a time); each thread then reads ITS line through a 4-byte window: one aligned 32-bit shared-memory
 * load per 4 characters, issued a whole word ahead of its use, so the scanner's character-by-character dependency chain is register
 * shifts, not memory latency.  The finished rows go to shared memory and leave as one coalesced store per block.
 * A block whose text does not fit the stage (long lines, -subsample) falls back to the scanner that reads global memory. */
#define CSV_PB 128
#define CSV_STAGE_BYTES (CSV_PB * 160)                      /* 20 KB: lines of up to 160 characters on average */

/* One line through the reference's scanner, written as a state machine that takes one character per turn so that the threads of a
 * warp stay together: the calls  scanfast_double(&p); scanfast_float(&p) x ntrks  (csvtbin.c:691-693) see the line as
 *      field := [ ,]* -? digit* ( . digit* )?
 * repeated, each field starting at the character that ended the one before.  State 0 = in the leading blanks / commas of a field,
 * 1 = after the sign or in the integer digits, 2 = in the fraction digits.  A character that can start nothing ('\n', a letter)
 * leaves this field and every later one 0, as the reference's scanner returns 0 without advancing.
 * The characters come four at a time: `cur` = the next 4 characters of the line whatever its alignment (funnel shift of two aligned
 * shared-memory words, the following word already loaded), consumed by four unrolled copies of the step; a digit -- three
 * characters in four -- takes the short branch.  vals[f * CSV_PB] receives field f (the time stamp, field -1, is dropped). */
struct LineScan {
   int field, state, k, ntrks;
   float n;
   bool negative;
   float *vals;
   const float *frac;

   /* a character that is not a digit; false = the line is finished (all fields done, or nothing more can be scanned) */
   __device__ __forceinline__ bool other(unsigned c) {
      if (state == 1 && c == '.') { state = 2; return true; }
      if (state != 0) {                                       /* the field ends here; c is looked at again as the start of the next */
         if (field >= 0) vals[field * CSV_PB] = negative ? -n : n;
         if (++field == ntrks) return false;
         n = 0; negative = false; k = 0; state = 0; }
      if (c == ' ' || c == ',') return true;
      if (c == '-') { negative = true; state = 1; return true; }
      if (c == '.') { state = 2; return true; }
      return false; }

   __device__ __forceinline__ bool step(unsigned c) {
      const unsigned d = c - '0';
      if (d < 10u) {
         if (state == 2) { n += frac[k + d]; k = min(k + 10, 10 * (CSV_FRAC_ROWS - 1)); }
         else { n = n * 10 + digit_value(d); state = 1; }
         return true; }
      return other(c); }

   __device__ __forceinline__ void run(const unsigned char *stage, uint32_t off) {
      const uint32_t *w = reinterpret_cast<const uint32_t *>(stage) + (off >> 2);
      const uint32_t sh = 8 * (off & 3u);
      uint32_t w0 = w[0], last = w[1]; w += 2;
      uint32_t cur = __funnelshift_r(w0, last, sh);
      field = -1; state = 0; k = 0; n = 0; negative = false;
      for (;;) {
         const uint32_t v = *w++;                             /* the word after next: its latency hides behind four characters */
         const uint32_t nxt = __funnelshift_r(last, v, sh);
         last = v;
         if (!step(cur & 0xffu)) break;
         if (!step(cur >> 8 & 0xffu)) break;
         if (!step(cur >> 16 & 0xffu)) break;
         if (!step(cur >> 24)) break;
         cur = nxt; }
      for (int f = field < 0 ? 0 : field; f < ntrks; ++f) vals[f * CSV_PB] = 0.0f; } };

/* stats: [0] too big, [1] too small (u64); fstats: [0] max volts (>= 0), [1] -min volts (>= 0)
   dynamic shared memory: stage[CSV_STAGE_BYTES + 32] | frac[CSV_FRAC_ROWS * 10] | mbarrier | vals[ntrks][CSV_PB] floats;
   the finished rows reuse the stage */
static size_t csv_parse_smem(int ntrks) { return CSV_STAGE_BYTES + 32 + (size_t)ntrks * CSV_PB * 4 + CSV_FRAC_ROWS * 10 * 4 + 16; }

__global__ void __launch_bounds__(CSV_PB, 8)
k_csv_parse(const char *txt, uint64_t n, const uint64_t *line_start, uint64_t nlines_total, const __grid_constant__ CsvParseArgs a, uint64_t nrows,
            int16_t *rows, unsigned long long *stats, float *fstats) {
   extern __shared__ __align__(128) unsigned char smem[];
   unsigned char *stage = smem;
   float *frac = reinterpret_cast<float *>(smem + CSV_STAGE_BYTES + 32);
   uint64_t *bar = reinterpret_cast<uint64_t *>(frac + CSV_FRAC_ROWS * 10);
   float *vals = reinterpret_cast<float *>(bar + 2);
   int16_t *srows = reinterpret_cast<int16_t *>(stage);
   const uint64_t r0 = (uint64_t)blockIdx.x * CSV_PB;
   const uint32_t nr = (uint32_t)(nrows - r0 < CSV_PB ? nrows - r0 : CSV_PB);
   /* the text of this block's rows: from the first character of its first line to the '\n' of its last (csvtbin.c:686: of every
      `subsample` lines the LAST one is used) */
   const uint64_t g_lo = line_start[a.first_line + (r0 + 1) * a.subsample - 1];
   const uint64_t g_hi = line_start[a.first_line + (r0 + nr) * a.subsample];
   const uint64_t base = g_lo & ~15ull;
   const uint64_t bytes = (g_hi - base + 15) & ~15ull;      /* the text buffer is padded: reading up to 31 bytes past its end is fine */
   const bool staged = bytes <= CSV_STAGE_BYTES;
   if (threadIdx.x == 0 && staged) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(bar, (uint32_t)bytes);
      tma_load_1d(stage, txt + base, (uint32_t)bytes, bar); }
   fill_frac_table(frac);                                    /* ends with __syncthreads(): the barrier is initialised for everybody */
   const uint32_t t = threadIdx.x;
   float *my = vals + t;
   if (staged) mbar_wait(bar, 0);
   if (t < nr) {
      const uint64_t line = a.first_line + (r0 + t + 1) * a.subsample - 1;
      uint64_t lo;
      const bool plain = line_is_plain(line_start, line, &lo);
      if (staged && plain) { LineScan s; s.ntrks = a.ntrks; s.vals = my; s.frac = frac; s.run(stage, (uint32_t)(lo - base)); }
      else if (plain) {
         const unsigned char *p = reinterpret_cast<const unsigned char *>(txt) + lo;
         skip_number_fast(p);
#pragma unroll 1
         for (int k = 0; k < a.ntrks; ++k) my[k * CSV_PB] = scan_float_fast(p, frac); }
      else {
         Cursor c = line_cursor(txt, n, line_start, nlines_total, line);
         skip_number(c);
#pragma unroll 1
         for (int k = 0; k < a.ntrks; ++k) my[k * CSV_PB] = scan_float(c); } }
   __syncthreads();                                          /* everybody is done with the text: the stage becomes the rows */
   unsigned big = 0, small_ = 0; float vmax = 0, vmin = 0;
   if (t < nr) {
#pragma unroll 1
      for (int k = 0; k < a.ntrks; ++k)                       /* csvtbin.c:693: column k goes to position track_permutation[k] */
         srows[t * a.ntrks + a.perm[k]] = quantise(my[k * CSV_PB] * a.scalefactor, a, big, small_, vmin, vmax); }
   __syncthreads();
   {  /* the block's rows are contiguous in the output and start 16-byte aligned (128 rows x ntrks x 2 bytes per block) */
      const uint32_t nbytes = nr * (uint32_t)a.ntrks * 2u;
      unsigned char *dst = reinterpret_cast<unsigned char *>(rows + r0 * (uint64_t)a.ntrks);
      const uint32_t n16 = nbytes / 16;
      for (uint32_t i = t; i < n16; i += CSV_PB) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(srows)[i];
      for (uint32_t i = n16 * 8 + t; i < nbytes / 2; i += CSV_PB) reinterpret_cast<int16_t *>(dst)[i] = srows[i]; }
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
   const unsigned nb = (unsigned)((nrows + CSV_PB - 1) / CSV_PB);
   const size_t smem = csv_parse_smem(ntrks);
   cudaError_t e = cudaFuncSetAttribute(k_csv_parse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csv_parse_smem(RT_MAXTRKS));
   if (e != cudaSuccess) return e;
   k_csv_parse<<<nb, CSV_PB, smem, s>>>(txt, n, line_start, nlines_total, a, nrows, rows, stats, fstats);
   return cudaGetLastError(); }
