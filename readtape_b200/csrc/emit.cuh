/* readtape_b200/csrc/emit.cuh -- event sinks of the scan kernels (device code).
 *
 * An event is what the reference's mode handler sees on entry to process_up/down_transition
 * (decoder.c:574/592), see rt_event in include/rt_scan.h.
 */
#pragma once
#include "rt_dev.h"

/* ---- emitters -------------------------------------------------------------------------------- */
struct FlatEmit {                 /* per-track flat buffer; counts past capacity so the host can regrow */
   rt_event *buf; uint32_t cap; uint32_t n; uint8_t trk;
   __device__ void emit(uint64_t row, double t_ev, float v_top, float v_bot, float agc, bool top) {
      if (n < cap) {
         rt_event e;
         e.row = row; e.t_event = t_ev; e.v_top = v_top; e.v_bot = v_bot; e.agc_gain = agc;
         e.trk = trk; e.kind = top ? RT_EV_TOP : RT_EV_BOT; e.pad[0] = e.pad[1] = 0;
         buf[n] = e; }
      ++n; } };

struct PoolEmit {                 /* chained fixed-size chunks from a global pool */
   rt_event *pool; uint32_t *chunk_next; unsigned int *cursor; uint32_t cap_chunks;
   uint32_t first_chunk, cur_chunk, n; uint64_t first_row; uint8_t trk; uint64_t last_row;
   __device__ void emit(uint64_t row, double t_ev, float v_top, float v_bot, float agc, bool top) {
      if (n == 0) first_row = row;
      last_row = row;
      uint32_t slot = n % RT_EVC;
      if (slot == 0) {
         uint32_t c = atomicAdd(cursor, 1u);          /* counts past capacity: the host regrows and reruns */
         if (c < cap_chunks) {
            chunk_next[c] = RT_NOCHUNK;
            if (n == 0) first_chunk = c; else if (cur_chunk != RT_NOCHUNK) chunk_next[cur_chunk] = c;
            cur_chunk = c; }
         else cur_chunk = RT_NOCHUNK; }
      if (cur_chunk != RT_NOCHUNK) {
         rt_event e;
         e.row = row; e.t_event = t_ev; e.v_top = v_top; e.v_bot = v_bot; e.agc_gain = agc;
         e.trk = trk; e.kind = top ? RT_EV_TOP : RT_EV_BOT; e.pad[0] = e.pad[1] = 0;
         pool[(size_t)cur_chunk * RT_EVC + slot] = e; }
      ++n; } };

