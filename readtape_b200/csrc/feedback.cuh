/* readtape_b200/csrc/feedback.cuh -- the per-event feedback the mode handlers apply to the detector.
 *
 * Shared by the exact generic scan (scan_generic.cuh, state = TrkState) and the int16 fast path
 * (scan_fast.cuh, state = FastState): templates over the state type, which only has to carry the
 * members used.  Reference (file:line in /root/reference/src):
 *   adjust_agc                          decoder.c:500-531
 *   AGC baseline, peaks 5..15           decode_nrzi.c:218-229 (top), :196-197 (bot)
 *   PE preamble / data-block start      decode_pe.c:127-155, :157-201
 * Compiled with -fmad=false; every expression keeps the reference's float type.
 */
#pragma once
#include "rt_dev.h"

#ifndef RT_HD
#define RT_HD __host__ __device__ inline
#endif

namespace rtfb {

template <class S>
RT_HD void agc_adjust(const DevCfg &c, S &t) {
   if (c.find_zeros) return;
   float gain, lastheight;
   if (c.p.agc_alpha != 0) {
      lastheight = t.v_lasttop - t.v_lastbot;
      if (lastheight > 0) {
         gain = t.avg_height / lastheight;
         gain = c.p.agc_alpha * gain + (1 - c.p.agc_alpha) * t.agc_gain;
         if (gain > RT_AGC_MAX_VALUE) gain = RT_AGC_MAX_VALUE;
         t.agc_gain = gain; } }
   if (c.p.agc_window != 0) {
      lastheight = t.v_lasttop - t.v_lastbot;
      if (lastheight > 0) {
         t.heights[t.heightndx] = lastheight;
         if (++t.heightndx >= c.p.agc_window) t.heightndx = 0;
         float minheight = 99;
         for (int i = 0; i < c.p.agc_window; ++i) if (t.heights[i] < minheight) minheight = t.heights[i];
         gain = t.avg_height / minheight;
         if (gain > RT_AGC_MAX_VALUE) gain = RT_AGC_MAX_VALUE;
         t.agc_gain = gain; } } }

template <class S>
RT_HD void baseline_accumulate(const DevCfg &c, S &t) {
   t.avg_height_sum += t.v_top - t.v_bot;
   ++t.avg_height_count;
   t.heights[t.heightndx] = t.v_top - t.v_bot;
   if (++t.heightndx >= c.p.agc_window) t.heightndx = 0; }

template <class S>
RT_HD void nrzi_feedback(const DevCfg &c, S &t, bool top) {
   if (top) {
      if (t.peakcount >= RT_AGC_STARTBASE && t.peakcount <= RT_AGC_ENDBASE) baseline_accumulate(c, t);
      else if (t.peakcount > RT_AGC_ENDBASE) {
         if (t.avg_height_count) {
            t.avg_height = t.avg_height_sum / t.avg_height_count;
            t.avg_height_count = 0; }
         else agc_adjust(c, t); } }
   else if (t.peakcount > RT_AGC_ENDBASE && t.avg_height_count == 0) agc_adjust(c, t); }

template <class S>
RT_HD void pe_feedback(const DevCfg &c, S &t, bool top, double t_ev) {
   if (t.datablock) { agc_adjust(c, t); return; }
   if (t.peakcount == 1) t.bit1_up = !top;
   if (t.peakcount > RT_PE_MIN_PREBITS && (t.bit1_up != 0) == top && t_ev - t.t_lastpeak > t.t_clkwindow) {
      t.datablock = 1;
      t.avg_height = t.avg_height_sum / t.avg_height_count; }
   else if (t.peakcount >= RT_AGC_STARTBASE && t.peakcount <= RT_AGC_ENDBASE && t.v_top > t.v_bot)
      baseline_accumulate(c, t); }

}  // namespace rtfb
