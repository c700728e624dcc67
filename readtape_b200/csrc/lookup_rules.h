/* readtape_b200/csrc/lookup_rules.h -- the unit-equivalence rules of rt_bulk_lookup(), host code.
 *
 * The speculative whole-tape scan resets every unit at a row the unit finder PROPOSED; the reference resets at the row its
 * end-of-block logic decides (src/readtape.c:1759, init_trackstate() src/decoder.c:425).  Whether the events of a unit may stand in for
 * a fresh reset at the reference's row is decided here, from the per-track proof data the scan kernel left behind (TrkMeta, rt_dev.h).
 * Plain functions of (DevCfg, UnitDesc, TrkMeta[]) with no CUDA and no library state, in a header of their own so that the CPU test
 * harness (tests/host_fast) compiles THIS code: tests/test_proof_host.py attacks it with units cut anywhere and compares every accepted
 * reset with the oracle's fresh scan, tests/test_hitrate_model.py counts what it serves on the whole bundled GCR captures.
 */
#ifndef RT_LOOKUP_RULES_H
#define RT_LOOKUP_RULES_H
#include <algorithm>
#include <cstdint>
#include "rt_dev.h"

namespace rtlookup {

/* rows a scan reset at some row needs before track k's detector state is a pure function of the samples: the rows not looked at
   after a reset (decoder.c:855-861: track k starts at row k, one later when the time stamp is 0), the deskew delay, the window */
inline int fill_of(const DevCfg &dc, uint32_t k, bool tz) {
   int lead = std::max<int>((int)k + (tz ? 1 : 0), dc.skew[k]);
   return dc.det == RT_DET_PEAK ? lead + dc.width + 1 : lead + 2; }

/* Can unit `u` (proof data m[0..nt)) stand in for a fresh RT_RESET_FULL at `start_row`?  tz: the time stamp of start_row is 0.
   *bridge_to (if given) is set when the only thing missing is the quietness of the rows between start_row and the unit's canonical
   rows: the row up to which an exact scan from start_row has to stay event-free for the equivalence to hold all the same (rt_api.cu:
   bridge_holds); RT_NOROW if that cannot help. */
inline bool unit_covers(const DevCfg &dc, const UnitDesc &u, const TrkMeta *m, uint32_t nt, uint64_t start_row, bool tz, uint64_t *bridge_to = nullptr) {
   if (bridge_to) *bridge_to = RT_NOROW;
   if (start_row >= u.row_end) return false;
   for (uint32_t k = 0; k < nt; ++k) if (m[k].failed) return false;
   if (start_row == u.row0) return true;                        /* the very same reset: trivially identical */
   const uint64_t pre0 = u.row0 > (uint64_t)dc.prescan_rows ? u.row0 - (uint64_t)dc.prescan_rows : 0;
   const bool examined = start_row >= pre0;                     /* quietness before pre0 was never examined */
   bool all = true, bridgeable = dc.det == RT_DET_PEAK; uint64_t upto = 0;
   /* The zero-crossing detectors keep their extremes and their armed flags through quiet rows (decoder.c:617-683: v_top / v_bot and
      zerocross_*_pending only change at crossings), so "not loud since start_row" is not enough there: a loud excursion of the UNIT's
      own scan in [row0, start_row) leaves it armed where the fresh scan is not.  For them the unit itself must have been quiet from its
      first row up to the canonical row (found by the proof-soundness fuzz of tests/test_proof_host.py, end of round 2). */
   const bool zc = dc.det != RT_DET_PEAK;
   auto quiet_since = [&](uint64_t loud) { return loud == RT_NOROW || (loud < start_row && (!zc || loud < u.row0)); };
   for (uint32_t k = 0; k < nt; ++k) {
      const uint64_t need = start_row + (uint64_t)fill_of(dc, k, tz);
      /* two recorded (canonical row, last loud row before it) pairs: the end of the unit's first quiet stretch,
         and the last one before its first event; either proves the equivalence */
      const bool late = examined && m[k].sync_row != RT_NOROW && m[k].sync_row >= need && quiet_since(m[k].last_loud_row);
      const bool early = examined && m[k].sync_early != RT_NOROW && m[k].sync_early >= need && quiet_since(m[k].loud_early);
      if (!late && !early) {
         all = false;
         if (m[k].sync_row != RT_NOROW && m[k].sync_row >= need) upto = std::max(upto, m[k].sync_row); else bridgeable = false; } }
   if (!all && bridgeable && bridge_to && upto - start_row <= 65536) *bridge_to = upto;
   return all; }

/* The tail rule: start_row lies behind every event of the unit, and no row of [start_row, row_end) is loud on any track: a fresh
   scan from start_row stays in default state (no event, so no feedback) and cannot fire before row_end either. */
inline bool unit_tail_covers(const UnitDesc &u, const TrkMeta *m, uint32_t nt, uint64_t start_row) {
   if (start_row < u.row0 || start_row >= u.row_end) return false;
   for (uint32_t k = 0; k < nt; ++k) {
      if (m[k].failed) return false;
      if (m[k].nevents && m[k].last_event_row >= start_row) return false;
      if (m[k].quiet_tail_from == RT_NOROW || m[k].quiet_tail_from > start_row) return false; }
   return true; }

/* Chaining: a unit that holds no event at all is passed through unchanged by the reference's scan, which is then identical to the
   NEXT unit's fresh scan from that unit's first canonical row on -- provided that row lies inside the stretch where this unit has
   already shown the scan to be event-free (the overlap of the two units).  `start_row` is the reset row that unit_covers() accepted
   for `u`.
   Zero-crossing detectors: event-free is not enough.  Their extremes and armed flags survive quiet rows (see unit_covers), so a loud
   excursion inside `u` that fires nothing still leaves the passing scan in a state the next unit's fresh scan does not have.  There
   the passing scan must be QUIET, not just event-free, up to the next unit's canonical row: this unit's last canonical row lies at
   or behind it, with no loud row since the reset (found by tools/fuzz_campaign.py proof, seed 188, at the end of round 2). */
inline bool chains_into_next(const DevCfg &dc, const UnitDesc &u, const TrkMeta *m, const TrkMeta *mn, uint32_t nt, uint64_t start_row) {
   const bool zc = dc.det != RT_DET_PEAK;
   const uint64_t quiet_from = std::min(start_row, u.row0);
   for (uint32_t k = 0; k < nt; ++k) if (m[k].nevents) return false;
   for (uint32_t k = 0; k < nt; ++k) {
      if (mn[k].failed || mn[k].sync_first == RT_NOROW || mn[k].sync_first >= u.row_end) return false;
      if (zc && !(m[k].sync_row != RT_NOROW && m[k].sync_row >= mn[k].sync_first && (m[k].last_loud_row == RT_NOROW || m[k].last_loud_row < quiet_from)))
         return false; }
   return true; }

}  // namespace rtlookup
#endif
