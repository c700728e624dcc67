/* tests/host_fast/hostsim.cu -- TEST INFRASTRUCTURE: the whole C-ABI of include/rt_scan.h on the CPU, speculative scan included.
 *
 * The CPU oracle (oracle/scan_oracle.c) implements only the exact scan, so the host shim linked against it never sees a speculative
 * hit, a unit that ends inside a block, a bridge scan, a chained unit, a parameter-set fan-out or a worker hand-over -- the paths
 * of readtape_b200/host/readblock_b200.c that only ran on a GPU.  This library closes that gap where there is no GPU:
 *   rt_scan_*, rt_upload*, ...   forwarded to the oracle (dlopen'ed with RTLD_DEEPBIND, its own symbols stay its own);
 *   rt_bulk_scan                 the unit finder of k_units.cu re-stated for the host (quiet granules, min_gap, tail_rows with the
 *                                thresholds of make_plan, rt_api.cu), every (unit, track) scanned by the HOST BUILD of the product's
 *                                scan code (fast_host.cu: two-pass peak scan, zero-crossing fast path, generic unit scan) --
 *                                the same __host__ __device__ code the kernels instantiate, with the same proof data;
 *   rt_bulk_lookup               the product's own rules (readtape_b200/csrc/lookup_rules.h), the bridge scan through the oracle's
 *                                exact scan, chaining, the (row, track) merge.
 * `readtape_shim_hostsim` (readtape_b200/host/Makefile) = the reference's host code + readblock_b200.c + this library; the tests run
 * it beside the unmodified reference (tests/test_hostsim.py, tools/fuzz_campaign.py shim-sim-*).  Never linked into the product.
 * HOSTSIM_UNIT_ROWS=n additionally cuts every unit after n rows (a cut INSIDE blocks: restarts and continued exact scans).
 */
#include "fast_host.cu"
#include "rt_csv.h"
#include <dlfcn.h>
#include <unistd.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace {
struct Ora {
   void *h = nullptr;
#define ORA_FN(name) decltype(&::name) name = nullptr;
   ORA_FN(rt_last_error) ORA_FN(rt_open) ORA_FN(rt_upload) ORA_FN(rt_upload_fd) ORA_FN(rt_clear) ORA_FN(rt_nrows) ORA_FN(rt_close) ORA_FN(rt_host_alloc) ORA_FN(rt_host_free)
   ORA_FN(rt_scan_begin) ORA_FN(rt_scan_reset) ORA_FN(rt_scan_run) ORA_FN(rt_scan_rewind) ORA_FN(rt_scan_set_avg_height) ORA_FN(rt_scan_set_cfg) ORA_FN(rt_scan_pos)
   ORA_FN(rt_scan_end) ORA_FN(rt_peak_masks) ORA_FN(rt_pkww_width) ORA_FN(rt_row_time)
#undef ORA_FN
};
Ora &ora() {
   static Ora o;
   if (!o.h) {
      const char *p = getenv("HOSTSIM_ORACLE");
      std::string path = p ? p : "";
      if (path.empty()) {                                          /* beside this library's own directory: ../../../oracle/_ref */
         Dl_info di; dladdr((void *)&ora, &di);
         std::string me = di.dli_fname; me = me.substr(0, me.rfind('/'));
         path = me + "/../../../oracle/_ref/libscan_oracle.so"; }
      /* RTLD_DEEPBIND keeps the oracle's internal calls of its own rt_* functions its own (this library exports the same names).  The
         sanitizer runtimes refuse that flag: HOSTSIM_NO_DEEPBIND=1 drops it, for an oracle built with -Wl,-Bsymbolic (HOSTSIM_ORACLE) */
      o.h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL | (getenv("HOSTSIM_NO_DEEPBIND") ? 0 : RTLD_DEEPBIND));
      if (!o.h) { fprintf(stderr, "hostsim: cannot load the oracle (%s): %s\n", path.c_str(), dlerror()); abort(); }
#define ORA_FN(name) o.name = (decltype(o.name))dlsym(o.h, #name); if (!o.name) { fprintf(stderr, "hostsim: oracle lacks %s\n", #name); abort(); }
      ORA_FN(rt_last_error) ORA_FN(rt_open) ORA_FN(rt_upload) ORA_FN(rt_upload_fd) ORA_FN(rt_clear) ORA_FN(rt_nrows) ORA_FN(rt_close) ORA_FN(rt_host_alloc) ORA_FN(rt_host_free)
      ORA_FN(rt_scan_begin) ORA_FN(rt_scan_reset) ORA_FN(rt_scan_run) ORA_FN(rt_scan_rewind) ORA_FN(rt_scan_set_avg_height) ORA_FN(rt_scan_set_cfg) ORA_FN(rt_scan_pos)
      ORA_FN(rt_scan_end) ORA_FN(rt_peak_masks) ORA_FN(rt_pkww_width) ORA_FN(rt_row_time)
#undef ORA_FN
   }
   return o; }

char g_err[512]; bool g_err_own = false;
int sim_err(int code, const char *fmt, ...) {
   va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); g_err_own = true; return code; }
int fwd(int rc) { if (rc) g_err_own = false; return rc; }
}  // namespace

/* the tape: the oracle's object for the exact scans + our own track-major planes for the host build of the scan kernels */
struct rt_tape {
   rt_tape *ot = nullptr;                                          /* the oracle's rt_tape (an opaque type of ITS library) */
   rt_tape_desc desc{};
   std::vector<int16_t> rows;                                      /* everything uploaded so far, row-major as in the file */
   std::vector<int16_t> planes, planes_inv; uint64_t stride = 0, planes_rows = 0, planes_id = 0; bool inv_ready = false;
};
struct rt_scan { rt_scan *os = nullptr; rt_tape *tape = nullptr; };

struct SimCfg {
   rt_scan_cfg cfg{}; DevCfg dc{};
   std::vector<UnitDesc> units; std::vector<TrkMeta> meta;         /* meta[unit * ntrks + trk] */
   std::vector<std::vector<rt_event>> ev;                          /* ev[unit * ntrks + trk] */
   size_t last_unit = ~(size_t)0;
};
struct rt_bulk {
   rt_tape *tape = nullptr; std::vector<SimCfg> cfgs; std::vector<rt_event> result;
   rt_scan *bridge = nullptr; uint32_t bridge_cfg = ~0u; rt_bulk_stats stats{};
};

extern "C" {

const char *rt_last_error(void) { return g_err_own ? g_err : ora().rt_last_error(); }
int rt_set_option(int, int) { return RT_OK; }
int rt_abi_version(void) { return RT_ABI_VERSION; }
const char *rt_backend(void) { return "hostsim-cpu"; }

int rt_open(const rt_tape_desc *desc, int device, rt_tape **out) {
   if (!desc || !out) return sim_err(RT_ERR_ARG, "rt_open: null argument");
   rt_tape *t = new rt_tape(); t->desc = *desc;
   int rc = ora().rt_open(desc, device, &t->ot);
   if (rc) { delete t; return fwd(rc); }
   *out = t; return RT_OK; }
int rt_upload(rt_tape *t, const int16_t *rows, uint64_t nrows) {
   if (!t) return sim_err(RT_ERR_ARG, "rt_upload: null tape");
   int rc = ora().rt_upload(t->ot, rows, nrows); if (rc) return fwd(rc);
   t->rows.insert(t->rows.end(), rows, rows + nrows * t->desc.nheads);
   return RT_OK; }
int rt_upload_fd(rt_tape *t, int fd, uint64_t offset, uint64_t nrows) {
   if (!t) return sim_err(RT_ERR_ARG, "rt_upload_fd: null tape");
   std::vector<int16_t> buf((size_t)nrows * t->desc.nheads);
   size_t got = 0, want = buf.size() * 2;
   while (got < want) { ssize_t k = pread(fd, (char *)buf.data() + got, want - got, (off_t)(offset + got)); if (k <= 0) break; got += (size_t)k; }
   if (got < want) memset((char *)buf.data() + got, 0, want - got);
   return rt_upload(t, buf.data(), nrows); }
int rt_attach_device(rt_tape *, const void *, uint64_t) { return sim_err(RT_ERR_UNSUPPORTED, "hostsim has no device memory"); }
int rt_prepare(rt_tape *, const rt_scan_cfg *) { return RT_OK; }
int rt_clear(rt_tape *t) { if (!t) return RT_ERR_ARG; t->rows.clear(); t->planes_rows = 0; t->inv_ready = false; return fwd(ora().rt_clear(t->ot)); }
uint64_t rt_nrows(const rt_tape *t) { return t ? ora().rt_nrows(t->ot) : 0; }
void rt_close(rt_tape *t) { if (!t) return; ora().rt_close(t->ot); delete t; }
void *rt_host_alloc(size_t n) { return ora().rt_host_alloc(n); }
void rt_host_free(void *p) { ora().rt_host_free(p); }
int rt_host_register(rt_tape *, void *, size_t) { return RT_OK; }
int rt_host_unregister(rt_tape *, void *) { return RT_OK; }

int rt_scan_begin(rt_tape *t, const rt_scan_cfg *cfg, rt_scan **out) {
   if (!t || !out) return sim_err(RT_ERR_ARG, "rt_scan_begin: null argument");
   rt_scan *s = new rt_scan(); s->tape = t;
   int rc = ora().rt_scan_begin(t->ot, cfg, &s->os);
   if (rc) { delete s; return fwd(rc); }
   *out = s; return RT_OK; }
int rt_scan_reset(rt_scan *s, int kind, uint64_t row) { return fwd(ora().rt_scan_reset(s->os, kind, row)); }
int rt_scan_run(rt_scan *s, uint64_t n, const rt_event **ev, uint64_t *nev, uint64_t *done) { return fwd(ora().rt_scan_run(s->os, n, ev, nev, done)); }
int rt_scan_rewind(rt_scan *s, uint64_t row) { return fwd(ora().rt_scan_rewind(s->os, row)); }
int rt_scan_set_avg_height(rt_scan *s, uint32_t trk, float v) { return fwd(ora().rt_scan_set_avg_height(s->os, trk, v)); }
int rt_scan_set_cfg(rt_scan *s, const rt_scan_cfg *cfg) { return fwd(ora().rt_scan_set_cfg(s->os, cfg)); }
uint64_t rt_scan_pos(const rt_scan *s) { return ora().rt_scan_pos(s->os); }
void rt_scan_end(rt_scan *s) { if (!s) return; ora().rt_scan_end(s->os); delete s; }
int rt_peak_masks(rt_tape *t, const rt_scan_cfg *cfg, float f, uint32_t *c, uint32_t *a, uint64_t w, int32_t *t0) { return fwd(ora().rt_peak_masks(t->ot, cfg, f, c, a, w, t0)); }
int rt_pkww_width(const rt_scan_cfg *cfg, uint64_t tdelta_ns) { return ora().rt_pkww_width(cfg, tdelta_ns); }
double rt_row_time(const rt_tape_desc *d, uint64_t row) { return ora().rt_row_time(d, row); }

/* ---- the speculative whole-tape scan -------------------------------------------------------------------------------------------- */
static void build_planes(rt_tape *t, uint64_t nrows) {
   if (t->planes_rows == nrows && t->stride) return;
   const uint32_t nt = t->desc.ntrks, nh = t->desc.nheads;
   t->stride = (nrows + 4096 + 63) / 64 * 64;
   t->planes.assign((size_t)t->stride * nt, 0);
   for (uint32_t h = 0; h < nh; ++h) {
      const int k = t->desc.head_to_trk[h];
      if (k < 0 || k >= (int)nt) continue;
      int16_t *pl = t->planes.data() + (size_t)k * t->stride;
      for (uint64_t r = 0; r < nrows; ++r) pl[r] = t->rows[(size_t)r * nh + h]; }
   t->planes_rows = nrows; t->inv_ready = false;
   static uint64_t next_id = 0x9e3779b97f4a7c15ull; t->planes_id = next_id += 2; }

static void find_units(const rt_tape *t, const int16_t *planes, uint64_t nrows, const DevCfg &dc, std::vector<UnitDesc> &units) {
   /* make_plan (rt_api.cu) + k_units.cu */
   const uint32_t nt = t->desc.ntrks;
   const double lsb = (double)t->desc.maxvolts / 32767.0;
   const double rows_per_bit = dc.bpi > 0 && dc.ips > 0 ? 1.0 / ((double)dc.bpi * dc.ips * dc.sample_deltat) : 1.0 / (200.0 * 50.0 * dc.sample_deltat);
   int thr;
   if (dc.det == RT_DET_PEAK) thr = (int)(0.75 * dc.p.pkww_rise / lsb);
   else if (dc.det == RT_DET_ZC) thr = (int)(0.9 * RT_ZEROCROSS_PEAK / lsb);
   else thr = (int)(0.9 * std::max(0.05, 0.5 / std::max(1, dc.samples_per_bit)) / lsb);
   const uint64_t gap_rows = (uint64_t)(6.0 * rows_per_bit) + 1;
   const uint64_t min_gap = std::max<uint64_t>(2, (gap_rows + RT_GRAN - 1) / RT_GRAN + 1);
   const uint64_t ibg_rows = (uint64_t)(200e-6 / dc.sample_deltat) + 1;
   const uint64_t tail_rows = (uint64_t)(16.0 * rows_per_bit) + ibg_rows + 64 + RT_PKWW_MAX_WIDTH + RT_MAXSKEWSAMP;
   const uint64_t ngran = nrows / RT_GRAN;
   std::vector<uint8_t> quiet(ngran + 1, 0);
   for (uint64_t g = 0; g < ngran; ++g) {
      bool q = true;
      for (uint32_t k = 0; k < nt && q; ++k) {
         const int16_t *pl = planes + (size_t)k * t->stride + g * RT_GRAN;
         int mn = 32767, mx = -32768;
         for (int i = 0; i < RT_GRAN; ++i) { if (pl[i] < mn) mn = pl[i]; if (pl[i] > mx) mx = pl[i]; }
         if (dc.det == RT_DET_ZC) { if (mx > thr || mn < -thr) q = false; }
         else if (mx - mn > thr) q = false; }
      quiet[g] = q; }
   std::vector<uint64_t> row0s; row0s.push_back(0);
   for (uint64_t g = 1; g < ngran; ++g) {
      if (!quiet[g] || quiet[g - 1] || g + min_gap > ngran) continue;
      bool ok = true;
      for (uint64_t j = 1; ok && j < min_gap; ++j) ok = quiet[g + j];
      if (ok) row0s.push_back(g * RT_GRAN); }
   /* HOSTSIM_UNIT_ROWS: extra cuts anywhere (the unit finder is only a heuristic: results must not depend on it) */
   const char *cut = getenv("HOSTSIM_UNIT_ROWS");
   if (cut && atoll(cut) >= 64) {
      const uint64_t step = (uint64_t)atoll(cut) / RT_GRAN * RT_GRAN;
      std::vector<uint64_t> more;
      for (size_t i = 0; i < row0s.size(); ++i) {
         const uint64_t hi = i + 1 < row0s.size() ? row0s[i + 1] : nrows;
         for (uint64_t r = row0s[i]; r < hi; r += step) more.push_back(r); }
      row0s.swap(more); }
   units.clear();
   for (size_t i = 0; i < row0s.size(); ++i) {
      uint64_t end = nrows;
      if (i + 1 < row0s.size()) end = std::min(nrows, row0s[i + 1] + tail_rows);
      units.push_back(UnitDesc{row0s[i], end}); }
   if (!nrows) units.clear(); }

int rt_bulk_scan(rt_tape *t, const rt_scan_cfg *cfgs, uint32_t ncfgs, rt_bulk **out) {
   if (!t || !cfgs || !ncfgs || !out) return sim_err(RT_ERR_ARG, "rt_bulk_scan: null argument");
   const uint64_t nrows = rt_nrows(t);
   const uint32_t nt = t->desc.ntrks;
   for (uint32_t i = 0; i < ncfgs; ++i) {
      rt_scan *probe = nullptr;                                    /* cfg_check: whatever the exact scan refuses, the bulk scan refuses too */
      int rc = rt_scan_begin(t, &cfgs[i], &probe); if (rc) return rc;
      rt_scan_end(probe);
      if (cfgs[i].mode == RT_MODE_WW) return sim_err(RT_ERR_UNSUPPORTED, "Whirlwind state persists across blocks: use rt_scan_*");
      if ((cfgs[i].flags & RT_F_DENSITY_DETECT) && (cfgs[i].flags & RT_F_FIND_ZEROS))
         return sim_err(RT_ERR_UNSUPPORTED, "density detection with the zero-crossing detector: use rt_scan_*"); }
   build_planes(t, nrows);
   rt_bulk *b = new rt_bulk(); b->tape = t; b->cfgs.resize(ncfgs);
   const uint32_t cap = 1u << 20;
   std::vector<rt_event> evbuf((size_t)cap * nt); std::vector<uint32_t> counts(nt); std::vector<TrkMeta> metas(nt);
   for (uint32_t ci = 0; ci < ncfgs; ++ci) {
      SimCfg &sc = b->cfgs[ci]; sc.cfg = cfgs[ci];
      rt_scan_cfg cfg = cfgs[ci];
      const int16_t *planes = t->planes.data();
      if ((cfg.flags & RT_F_INVERT) && !(cfg.flags & RT_F_DIFFERENTIATE)) {      /* the product scans a negated copy with the flag cleared */
         if (!t->inv_ready) { t->planes_inv.resize(t->planes.size()); for (size_t i = 0; i < t->planes.size(); ++i) t->planes_inv[i] = (int16_t)-t->planes[i]; t->inv_ready = true; }
         planes = t->planes_inv.data(); cfg.flags &= ~(uint32_t)RT_F_INVERT; }
      g_planes_id = t->planes_id + (planes == t->planes_inv.data() ? 1 : 0);     /* fast_host.cu: no content hash per unit */
      rtcfg::to_dev(t->desc, planes, t->stride, nrows, &cfg, &sc.dc);
      find_units(t, planes, nrows, sc.dc, sc.units);
      sc.meta.resize(sc.units.size() * nt); sc.ev.resize(sc.units.size() * nt);
      for (size_t u = 0; u < sc.units.size(); ++u) {
         const UnitDesc ud = sc.units[u];
         int rc = sparse_host_scan_unit(planes, t->stride, nrows, &t->desc, &cfg, ud.row0, ud.row_end, evbuf.data(), cap, counts.data(), metas.data(), 0.25f, 1, 0);
         if (rc == RT_ERR_UNSUPPORTED) rc = fast_host_scan_unit(planes, t->stride, nrows, &t->desc, &cfg, ud.row0, ud.row_end, evbuf.data(), cap, counts.data(), metas.data(), 0);
         if (rc == RT_ERR_UNSUPPORTED) rc = generic_host_scan_unit(planes, t->stride, nrows, &t->desc, &cfg, ud.row0, ud.row_end, evbuf.data(), cap, counts.data(), metas.data());
         if (rc) { delete b; return sim_err(rc, "hostsim: unit scan failed (%d)", rc); }
         for (uint32_t k = 0; k < nt; ++k) {
            if (counts[k] > cap) { delete b; return sim_err(RT_ERR_OVERFLOW, "hostsim: more than %u events in one (unit, track)", cap); }
            sc.meta[u * nt + k] = metas[k];
            sc.ev[u * nt + k].assign(evbuf.begin() + (size_t)k * cap, evbuf.begin() + (size_t)k * cap + counts[k]);
            b->stats.events += counts[k]; } } }
   g_planes_id = 0;
   b->stats.rows = nrows; b->stats.units = b->cfgs[ncfgs - 1].units.size(); b->stats.track_samples = nrows * nt * ncfgs;
   *out = b; return RT_OK; }

int rt_bulk_fetch(rt_bulk *b) { return b ? RT_OK : RT_ERR_ARG; }
int rt_bulk_fetch_to(rt_bulk *b, void *, size_t) { return b ? RT_OK : RT_ERR_ARG; }
int rt_bulk_results_size(const rt_bulk *, uint64_t *) { return sim_err(RT_ERR_UNSUPPORTED, "hostsim: no result image"); }
int rt_bulk_results_to_device(const rt_bulk *, void *, uint64_t) { return sim_err(RT_ERR_UNSUPPORTED, "hostsim: no device"); }
int rt_bulk_scan_host(rt_tape *t, const int16_t *rows, uint64_t nrows, const rt_scan_cfg *cfg, rt_bulk **out) {
   int rc = rt_clear(t); if (rc) return rc;
   rc = rt_upload(t, rows, nrows); if (rc) return rc;
   return rt_bulk_scan(t, cfg, 1, out); }
int rt_bulk_get_stats(const rt_bulk *b, rt_bulk_stats *out) { if (!b || !out) return RT_ERR_ARG; *out = b->stats; return RT_OK; }
int rt_bulk_tile_digest(rt_bulk *, uint32_t, uint64_t, uint64_t, uint64_t *, uint64_t *, uint64_t *) { return sim_err(RT_ERR_UNSUPPORTED, "hostsim: no tile digest"); }
void rt_bulk_free(rt_bulk *b) { if (!b) return; if (b->bridge) rt_scan_end(b->bridge); delete b; }

int rt_bulk_last_unit(const rt_bulk *b, uint32_t ci, uint64_t *row0, uint64_t *row_end) {
   if (!b || ci >= b->cfgs.size()) return sim_err(RT_ERR_ARG, "rt_bulk_last_unit: bad argument");
   const SimCfg &sc = b->cfgs[ci];
   if (sc.last_unit >= sc.units.size()) return sim_err(RT_ERR_STATE, "rt_bulk_last_unit: no successful lookup yet");
   if (row0) *row0 = sc.units[sc.last_unit].row0;
   if (row_end) *row_end = sc.units[sc.last_unit].row_end;
   return RT_OK; }

static void fill_unit_info(const rt_bulk *b, const SimCfg &sc, size_t lo, uint64_t start_row, rt_unit_info *out) {
   const uint32_t nt = b->tape->desc.ntrks;
   memset(out, 0, sizeof *out);
   out->unit_index = lo; out->nunits = sc.units.size(); out->row0 = sc.units[lo].row0; out->row_end = sc.units[lo].row_end; out->ntrks = nt;
   const bool tz = rt_row_time(&b->tape->desc, start_row) == 0.0;
   for (uint32_t k = 0; k < nt; ++k) {
      const TrkMeta &m = sc.meta[lo * nt + k];
      out->first_event_row[k] = m.first_event_row; out->sync_row[k] = m.sync_row; out->last_loud_row[k] = m.last_loud_row;
      out->sync_early[k] = m.sync_early; out->loud_early[k] = m.loud_early; out->sync_first[k] = m.sync_first; out->quiet_from[k] = m.quiet_from;
      out->failed[k] = m.failed; out->nevents[k] = m.nevents; out->need_sync_row[k] = start_row + (uint64_t)rtlookup::fill_of(sc.dc, k, tz); } }
static size_t unit_at_or_before(const SimCfg &sc, uint64_t start_row) {
   size_t lo = 0, hi = sc.units.size();
   while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (sc.units[mid].row0 <= start_row) lo = mid; else hi = mid; }
   return lo; }
int rt_bulk_unit_info(const rt_bulk *b, uint32_t ci, uint64_t start_row, rt_unit_info *out) {
   if (!b || ci >= b->cfgs.size() || !out) return sim_err(RT_ERR_ARG, "rt_bulk_unit_info: bad argument");
   const SimCfg &sc = b->cfgs[ci]; memset(out, 0, sizeof *out);
   if (sc.units.empty()) return RT_MISS;
   fill_unit_info(b, sc, unit_at_or_before(sc, start_row), start_row, out); return RT_OK; }
int rt_bulk_unit_at(const rt_bulk *b, uint32_t ci, uint64_t ui, rt_unit_info *out) {
   if (!b || ci >= b->cfgs.size() || !out) return sim_err(RT_ERR_ARG, "rt_bulk_unit_at: bad argument");
   const SimCfg &sc = b->cfgs[ci]; memset(out, 0, sizeof *out);
   if (ui >= sc.units.size()) return RT_MISS;
   fill_unit_info(b, sc, (size_t)ui, sc.units[ui].row0, out); return RT_OK; }

/* rt_api.cu: bridge_holds */
static int bridge_holds(rt_bulk *b, uint32_t ci, size_t ui, uint64_t start_row, uint64_t upto, bool *ok) {
   SimCfg &sc = b->cfgs[ci]; const uint32_t nt = b->tape->desc.ntrks; const TrkMeta *m = &sc.meta[ui * nt];
   *ok = false; int rc;
   if (b->bridge && b->bridge_cfg != ci) { rt_scan_end(b->bridge); b->bridge = nullptr; }
   if (!b->bridge) { rc = rt_scan_begin(b->tape, &sc.cfg, &b->bridge); if (rc) return rc; b->bridge_cfg = ci; }
   rc = rt_scan_reset(b->bridge, RT_RESET_FULL, start_row); if (rc) return rc;
   const rt_event *ev = nullptr; uint64_t n = 0, done = 0;
   rc = rt_scan_run(b->bridge, upto - start_row + 1, &ev, &n, &done); if (rc) return rc;
   if (done != upto - start_row + 1) return RT_OK;
   for (uint64_t i = 0; i < n; ++i) if (ev[i].row <= m[ev[i].trk].sync_row) return RT_OK;
   *ok = true; return RT_OK; }

int rt_bulk_lookup(rt_bulk *b, uint32_t ci, uint64_t start_row, const rt_event **events, uint64_t *nevents, uint64_t *valid_rows) {
   if (!b || ci >= b->cfgs.size()) return sim_err(RT_ERR_ARG, "rt_bulk_lookup: bad argument");
   SimCfg &sc = b->cfgs[ci]; const uint32_t nt = b->tape->desc.ntrks;
   if (sc.units.empty()) return RT_MISS;
   const bool tz = rt_row_time(&b->tape->desc, start_row) == 0.0;
   auto covers = [&](size_t ui, uint64_t *br) { return rtlookup::unit_covers(sc.dc, sc.units[ui], &sc.meta[ui * nt], nt, start_row, tz, br); };
   size_t lo = unit_at_or_before(sc, start_row);
   uint64_t br0 = RT_NOROW, br1 = RT_NOROW;
   if (!covers(lo, &br0)) {
      bool bridged = false;
      if (lo + 1 < sc.units.size() && covers(lo + 1, &br1)) { ++lo; bridged = true; }
      else if (br0 != RT_NOROW || br1 != RT_NOROW) {
         const char *env = getenv("RT_BRIDGE");
         if (!(env && env[0] == '0')) {
            if (br0 != RT_NOROW) { int rc = bridge_holds(b, ci, lo, start_row, br0, &bridged); if (rc) return rc; }
            if (!bridged && br1 != RT_NOROW) { int rc = bridge_holds(b, ci, lo + 1, start_row, br1, &bridged); if (rc) return rc; if (bridged) ++lo; } } }
      if (bridged) { }
      else if (rtlookup::unit_tail_covers(sc.units[lo], &sc.meta[lo * nt], nt, start_row)) {
         b->result.clear(); sc.last_unit = lo;
         if (events) *events = b->result.data();
         if (nevents) *nevents = 0;
         if (valid_rows) *valid_rows = sc.units[lo].row_end - start_row;
         return RT_OK; }
      else return RT_MISS; }
   sc.last_unit = lo;
   for (uint64_t from = start_row; lo + 1 < sc.units.size() && rtlookup::chains_into_next(sc.dc, sc.units[lo], &sc.meta[lo * nt], &sc.meta[(lo + 1) * nt], nt, from); ) {
      ++lo; from = sc.units[lo].row0; }
   /* merge the tracks into (row, trk) order */
   size_t at[RT_MAXTRKS] = {0}, total = 0;
   for (uint32_t k = 0; k < nt; ++k) total += sc.ev[lo * nt + k].size();
   b->result.clear(); b->result.reserve(total);
   for (size_t n = 0; n < total; ++n) {
      int best = -1; uint64_t brow = ~0ull;
      for (uint32_t k = 0; k < nt; ++k) { const auto &v = sc.ev[lo * nt + k]; if (at[k] < v.size() && v[at[k]].row < brow) { brow = v[at[k]].row; best = (int)k; } }
      b->result.push_back(sc.ev[lo * nt + best][at[best]++]); }
   if (getenv("HOSTSIM_TRACE")) fprintf(stderr, "[hostsim] lookup(%llu): first unit %zu [%llu, %llu), arrived at unit %zu [%llu, %llu), %zu events, first at row %lld\n", (unsigned long long)start_row,
                                        sc.last_unit, (unsigned long long)sc.units[sc.last_unit].row0, (unsigned long long)sc.units[sc.last_unit].row_end, lo, (unsigned long long)sc.units[lo].row0,
                                        (unsigned long long)sc.units[lo].row_end, b->result.size(), b->result.empty() ? -1ll : (long long)b->result[0].row);
   if (events) *events = b->result.data();
   if (nevents) *nevents = b->result.size();
   if (valid_rows) *valid_rows = sc.units[lo].row_end - start_row;
   return RT_OK; }

/* ---- include/rt_csv.h: forwarded to the oracle (so that the Python mirror of the ABI, readtape_b200/abi.py, can load this library) ---- */
#define CSV_FN(name) static decltype(&::name) f = (decltype(&::name))dlsym(ora().h, #name)
int rt_csv_open(int device, const char *text, uint64_t nbytes, rt_csv **out) { CSV_FN(rt_csv_open); return f ? f(device, text, nbytes, out) : RT_ERR_UNSUPPORTED; }
void rt_csv_close(rt_csv *csv) { CSV_FN(rt_csv_close); if (f) f(csv); }
uint64_t rt_csv_nlines(const rt_csv *csv) { CSV_FN(rt_csv_nlines); return f ? f(csv) : 0; }
int rt_csv_line(const rt_csv *csv, uint64_t line, uint64_t *offset, uint64_t *length) { CSV_FN(rt_csv_line); return f ? f(csv, line, offset, length) : RT_ERR_UNSUPPORTED; }
int rt_csv_max_abs(rt_csv *csv, uint64_t first_line, uint64_t nlines, uint32_t ntrks, float scalefactor, float *max_abs) {
   CSV_FN(rt_csv_max_abs); return f ? f(csv, first_line, nlines, ntrks, scalefactor, max_abs) : RT_ERR_UNSUPPORTED; }
int rt_csv_convert(rt_csv *csv, const rt_csv_cfg *cfg, uint64_t first_line, uint64_t nrows, int16_t *rows_out, rt_tape *tape, rt_csv_stats *stats) {
   if (tape) return sim_err(RT_ERR_UNSUPPORTED, "hostsim: rt_csv_convert into a tape");
   CSV_FN(rt_csv_convert); return f ? f(csv, cfg, first_line, nrows, rows_out, nullptr, stats) : RT_ERR_UNSUPPORTED; }
#undef CSV_FN

}  /* extern "C" */
