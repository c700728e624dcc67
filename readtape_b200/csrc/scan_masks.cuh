/* readtape_b200/csrc/scan_masks.cuh -- K3c phase A: the row-parallel part of the moving-window peak detector.
 *
 * lookfor_peak (decoder.c:751-810) evaluates, on every row of every track, a window maximum / minimum and the
 * shape tests of :790-803.  Everything in those tests except the AGC-dependent threshold, the blind countdown
 * and the lazily refreshed minimum (decoder.c:765) is a PURE function of the last `width`+1 samples of the track.
 * This pass computes those pure parts for every row of a plane, 2 rows per 32-bit lane (int16x2 SIMD), and
 * boils them down to two bit planes (1 bit per track-sample each):
 *
 *    cand[p]  =  S(p) - max(l, r) >= T0   or   min(l, r) - Wmin(p) >= T0
 *    acan[p]  =  raw[p - w] >= S(p)                       (the leaving sample was the window maximum)
 *
 * with window = raw[p-w+1 .. p], S / Wmin its max / min, l / r its edges.  `cand` is a conservative pre-filter:
 * a row can only pass the exact tests if the integer bound T of required_rise (scan_fast.cuh: thresholds())
 * satisfies T >= T0 and the bit is set -- for the bottom test because the lazy minimum m is always a sample of
 * the window, so m >= Wmin.  `acan` marks the rows at which the reference rescans its window because the
 * maximum left it (decoder.c:767): there the detector state is a pure function of the samples ("canonical",
 * DESIGN.md 4), and the sparse scan (scan_sparse.cuh) re-derives the lazy minimum from the last such row.
 *
 * Sliding max / min by doubling: P1 = x, P2[p] = max(P1[p], P1[p-1]), P4[p] = max(P2[p], P2[p-2]) ... and
 * window(w) = max(Pk[p], Pk[p-(w-k)]), k = largest power of two <= w: log2(w)+1 SIMD max per 2 rows, all in
 * registers, shifts by whole words free and by odd row counts one PRMT.  The width is a template parameter.
 * A thread owns a run of MASK_RUN rows of one track (+ a halo of `width` rows it re-reads).
 *
 * Host + device code: the CPU tests run the same functions (tests/host_fast) against a brute-force restatement.
 */
#pragma once
#include <stdint.h>
#include <string.h>
#include "rt_dev.h"

#ifndef RT_FHD
#define RT_FHD __host__ __device__ __forceinline__
#endif

namespace rtmask {

constexpr int MASK_RUN = 64;                                  /* rows per thread: two 32-row mask words */

/* ---- uint16x2 SIMD primitives (device: one VIMNMX.U16x2 each; host: emulation for the CPU tests).  Samples are biased
   (x ^ 0x8000) so that the signed order becomes the unsigned one: a per-half difference a - b with a >= b is then ONE
   32-bit subtraction (no borrow crosses the halves). ---- */
constexpr uint32_t BIAS2 = 0x80008000u;
RT_FHD uint32_t v_maxu2(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
   return __vmaxu2(a, b);
#else
   uint32_t ah = a >> 16, bh = b >> 16, al = a & 0xffffu, bl = b & 0xffffu;
   return ((ah > bh ? ah : bh) << 16) | (al > bl ? al : bl);
#endif
}
RT_FHD uint32_t v_minu2(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
   return __vminu2(a, b);
#else
   uint32_t ah = a >> 16, bh = b >> 16, al = a & 0xffffu, bl = b & 0xffffu;
   return ((ah < bh ? ah : bh) << 16) | (al < bl ? al : bl);
#endif
}
/* (hi half of lo_word, lo half of hi_word): the word one row earlier than hi_word */
RT_FHD uint32_t v_mid(uint32_t lo_word, uint32_t hi_word) {
#ifdef __CUDA_ARCH__
   return __byte_perm(lo_word, hi_word, 0x5432);
#else
   return (lo_word >> 16) | (hi_word << 16);
#endif
}

/* word i of array `a` shifted back by S rows: halves = rows (2i - S, 2i + 1 - S) */
template <int S, int N>
RT_FHD uint32_t shifted(const uint32_t (&a)[N], int i) {
   if constexpr (S % 2 == 0) return a[i - S / 2];
   else return v_mid(a[i - (S + 1) / 2], a[i - (S - 1) / 2]); }

template <int W> struct Pow2 { static constexpr int value = W >= 32 ? 32 : W >= 16 ? 16 : W >= 8 ? 8 : W >= 4 ? 4 : W >= 2 ? 2 : 1; };

/* One run: rows [p0, p0 + MASK_RUN) of `plane`, p0 a multiple of 32 and p0 >= HALO (the caller uses the scalar path
   below near the start of a plane).  1 <= T0 <= T1 <= 65535 (T1 = 65535: second plane practically empty).  Results: two mask words each. */
template <int W>
struct RunMasks {
   static constexpr int HALO = (W + 7) / 8 * 8;               /* rows read in front of the run (needs raw[p - W]) */
   static constexpr int NW = (MASK_RUN + HALO) / 2;           /* packed words held */
   static constexpr int K = Pow2<W>::value;
   static constexpr int H2 = HALO / 2;

   template <bool MAX>
   static RT_FHD uint32_t op(uint32_t a, uint32_t b) { return MAX ? v_maxu2(a, b) : v_minu2(a, b); }

   /* P <- sliding max (min) of the last K rows, in place, highest word first */
   template <bool MAX>
   static RT_FHD void doubling(uint32_t (&P)[NW]) {
      if constexpr (K >= 2) {
#pragma unroll
         for (int i = NW - 1; i >= 1; --i) P[i] = op<MAX>(P[i], v_mid(P[i - 1], P[i])); }
      if constexpr (K >= 4) {
#pragma unroll
         for (int i = NW - 1; i >= 2; --i) P[i] = op<MAX>(P[i], P[i - 1]); }
      if constexpr (K >= 8) {
#pragma unroll
         for (int i = NW - 1; i >= 4; --i) P[i] = op<MAX>(P[i], P[i - 2]); }
      if constexpr (K >= 16) {
#pragma unroll
         for (int i = NW - 1; i >= 8; --i) P[i] = op<MAX>(P[i], P[i - 4]); }
      if constexpr (K >= 32) {
#pragma unroll
         for (int i = NW - 1; i >= 16; --i) P[i] = op<MAX>(P[i], P[i - 8]); } }

   /* Flags are 0/1 at bit 0 (even row) and bit 16 (odd row) of a word.  Word j of a 16-row group is deposited at bit 2*(j&7):
      even rows land at their natural position in the low half, odd rows one bit too low in the high half. */
   static RT_FHD uint32_t fold(uint32_t g_lo, uint32_t g_hi) {                 /* two 16-row groups -> one row-ordered 32-row word */
      const uint32_t n0 = (g_lo & 0x5555u) | ((g_lo >> 15) & 0xaaaau);
      const uint32_t n1 = (g_hi & 0x5555u) | ((g_hi >> 15) & 0xaaaau);
      return n0 | (n1 << 16); }

   static RT_FHD void run(const int16_t *plane, int64_t p0, uint32_t T0, uint32_t T1, uint32_t (&cand)[2], uint32_t (&cand2)[2], uint32_t (&acan)[2]) {
      uint32_t x[NW];
      const int16_t *src = plane + (p0 - HALO);                /* 16-byte aligned: p0 % 32 == 0, HALO % 8 == 0, planes 16-byte aligned */
#pragma unroll
      for (int cch = 0; cch < NW / 4; ++cch) {
#ifdef __CUDA_ARCH__
         const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src) + cch);
         x[4 * cch] = v.x ^ BIAS2; x[4 * cch + 1] = v.y ^ BIAS2; x[4 * cch + 2] = v.z ^ BIAS2; x[4 * cch + 3] = v.w ^ BIAS2;
#else
         memcpy(&x[4 * cch], src + 8 * cch, 16);
         for (int k = 0; k < 4; ++k) x[4 * cch + k] ^= BIAS2;
#endif
      }
      core(x, T0, T1, cand, cand2, acan); }

   /* the same from packed samples already held: x[i] = rows (p0 - HALO + 2i, p0 - HALO + 2i + 1), biased (^ BIAS2).  Used by the
      ingest kernel, which has the de-interleaved tile in shared memory (k_ingest.cu) */
   static RT_FHD void core(const uint32_t (&x)[NW], uint32_t T0, uint32_t T1, uint32_t (&cand)[2], uint32_t (&cand2)[2], uint32_t (&acan)[2]) {
      uint32_t P[NW], Dt[MASK_RUN / 2];
      const uint32_t one2 = 0x00010001u, t0m1 = (T0 - 1u) | ((T0 - 1u) << 16), t1m1 = (T1 - 1u) | ((T1 - 1u) << 16);
      uint32_t gc[4] = {0, 0, 0, 0}, gd[4] = {0, 0, 0, 0}, gn[4] = {0, 0, 0, 0};
      /* pass 1: window maximum -> top-side margin S - max(l, r), and "the leaving sample is below the maximum" (not acan) */
#pragma unroll
      for (int i = 0; i < NW; ++i) P[i] = x[i];
      doubling<true>(P);
#pragma unroll
      for (int j = MASK_RUN / 2 - 1; j >= 0; --j) {
         const int i = H2 + j;
         const uint32_t S = K == W ? P[i] : v_maxu2(P[i], shifted<W - K>(P, i));
         const uint32_t l = shifted<W - 1>(x, i), lv = shifted<W>(x, i), r = x[i];
         Dt[j] = S - v_maxu2(l, r);
         const uint32_t na = v_minu2(S - v_minu2(S, lv), one2);               /* 1 where S > lv */
         gn[j >> 3] += na << (2 * (j & 7)); }
      /* pass 2: window minimum -> bottom-side margin min(l, r) - Wmin; a row is a candidate if the larger margin reaches T0 */
#pragma unroll
      for (int i = 0; i < NW; ++i) P[i] = x[i];
      doubling<false>(P);
#pragma unroll
      for (int j = MASK_RUN / 2 - 1; j >= 0; --j) {
         const int i = H2 + j;
         const uint32_t Wm = K == W ? P[i] : v_minu2(P[i], shifted<W - K>(P, i));
         const uint32_t l = shifted<W - 1>(x, i), r = x[i];
         const uint32_t Dm = v_maxu2(Dt[j], v_minu2(l, r) - Wm);
         const uint32_t hit = v_minu2(v_maxu2(Dm, t0m1) - t0m1, one2);        /* 1 where the margin >= T0 */
         const uint32_t hit2 = v_minu2(v_maxu2(Dm, t1m1) - t1m1, one2);       /* 1 where the margin >= T1 */
         gc[j >> 3] += hit << (2 * (j & 7));
         gd[j >> 3] += hit2 << (2 * (j & 7)); }
      cand[0] = fold(gc[0], gc[1]); cand[1] = fold(gc[2], gc[3]);
      cand2[0] = fold(gd[0], gd[1]); cand2[1] = fold(gd[2], gd[3]);
      acan[0] = ~fold(gn[0], gn[1]); acan[1] = ~fold(gn[2], gn[3]); } };

/* The same bits by definition, one row at a time: rows [p_from, p_to) (any alignment inside one mask word range is the
   caller's business: this returns the bits of ONE 32-row word `wi`).  Rows whose window would reach before row 0 get
   cand = acan = 0 (the sparse scan never looks at them).  Used near the start of a plane and by the tests. */
__host__ __device__ inline void word_masks_scalar(const int16_t *plane, int64_t wi, int w, int T0, int T1, uint32_t *cand, uint32_t *cand2, uint32_t *acan) {
   uint32_t cb = 0, db = 0, ab = 0;
   for (int b = 0; b < 32; ++b) {
      const int64_t p = wi * 32 + b;
      if (p < w) continue;
      int S = -32768, mn = 32767;
      for (int64_t q = p - w + 1; q <= p; ++q) { const int v = plane[q]; if (v > S) S = v; if (v < mn) mn = v; }
      const int l = plane[p - w + 1], r = plane[p], lv = plane[p - w];
      const int mxlr = l > r ? l : r, mnlr = l < r ? l : r;
      if (S - mxlr >= T0 || mnlr - mn >= T0) cb |= 1u << b;
      if (S - mxlr >= T1 || mnlr - mn >= T1) db |= 1u << b;
      if (lv >= S) ab |= 1u << b; }
   *cand = cb; *cand2 = db; *acan = ab; }

}  // namespace rtmask
