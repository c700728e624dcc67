/* readtape_b200/csrc/cfg_host.h -- host-side derivation of the device configuration block.
 *
 * Shared by rt_api.cu and the host build of the fast-path code used by the CPU tests.
 * Host code is compiled without FMA contraction (-Xcompiler -ffp-contract=off); `volatile`
 * pins the float rounding of every intermediate exactly as the reference build does
 * (FLT_EVAL_METHOD 0).
 */
#ifndef RT_CFG_HOST_H
#define RT_CFG_HOST_H
#include <string.h>
#include "rt_dev.h"

namespace rtcfg {

/* pkww_width, readtape.c:1453-1457 */
inline int pkww_width(const rt_scan_cfg *cfg, uint64_t tdelta_ns) {
   volatile float sample_deltat = (float)(long long)tdelta_ns / 1e9f;
   if (cfg->bpi == 0 || (cfg->flags & RT_F_DENSITY_DETECT)) return 8;
   volatile float a = cfg->bpi * cfg->ips;
   volatile float b = a * sample_deltat;
   int w = (int)(cfg->parms.pkww_bitfrac / b);
   return w < RT_PKWW_MAX_WIDTH ? w : RT_PKWW_MAX_WIDTH; }

inline void to_dev(const rt_tape_desc &desc, const int16_t *planes, uint64_t plane_stride, uint64_t nrows_valid,
                   const rt_scan_cfg *cfg, DevCfg *d) {
   memset(d, 0, sizeof *d);
   d->planes = planes; d->plane_stride = plane_stride; d->nrows = nrows_valid;
   d->tstart_ns = desc.tstart_ns; d->tdelta_ns = desc.tdelta_ns; d->maxvolts = desc.maxvolts;
   volatile float sd = (float)(long long)desc.tdelta_ns / 1e9f;              /* readtape.c:1345 */
   d->sample_deltat = sd;
   d->bpi = cfg->bpi; d->ips = cfg->ips;
   d->density = (cfg->flags & RT_F_DENSITY_DETECT) != 0;
   volatile float bi = cfg->bpi * cfg->ips;
   d->clk_init = d->density ? 0.0f : 1 / bi;                                 /* decoder.c:407-411, :448 */
   d->ntrks = (int)desc.ntrks; d->mode = cfg->mode;
   d->find_zeros = (cfg->flags & RT_F_FIND_ZEROS) != 0;
   d->differentiate = (cfg->flags & RT_F_DIFFERENTIATE) != 0;
   d->invert = (cfg->flags & RT_F_INVERT) != 0;
   d->det = d->find_zeros ? (d->differentiate ? RT_DET_DZC : RT_DET_ZC) : RT_DET_PEAK;
   d->width = pkww_width(cfg, desc.tdelta_ns);
   volatile float bid = bi * sd;
   d->samples_per_bit = cfg->bpi > 0 ? (int)(1 / bid) : 20;                  /* readtape.c:1402 */
   /* rows in front of a unit examined for quietness.  Measured: a longer pre-scan (an inter-block-gap time) costs every job of the
      whole-tape scan ~5 % and serves a handful of block starts per reel; those are proven by a bridge scan instead (rt_api.cu) */
   d->prescan_rows = RT_PRESCAN_MIN;
   d->p = cfg->parms;
   for (int k = 0; k < RT_MAXTRKS; ++k) d->skew[k] = cfg->skew_delaycnt[k]; }

/* Integer loudness threshold of the unit-equivalence proof for the moving-window peak detector
 * (DESIGN.md 4): a span of raw int16 samples whose max-min is below this cannot make any scan in
 * default state (required_rise == pkww_rise, decoder.c:785 with AGC 1 and height 4) fire.
 * 0 = nothing can be proven quiet. */
inline int quiet_thr_lsb(const DevCfg &dc) {
   if (dc.det != RT_DET_PEAK || dc.differentiate || !(dc.p.pkww_rise >= 1e-3f)) return 0;
   double lsb = (double)dc.maxvolts / 32767.0;
   double q = 0.998 * (double)dc.p.pkww_rise / lsb;
   if (q < 4) return 0;
   return q > 65535.0 ? 65535 : (int)q; }

}  // namespace rtcfg
#endif
