/* oracle/csv_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's CSV -> TBIN conversion behind the C-ABI of
 * include/rt_csv.h.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may use it; the product
 * (readtape_b200/csrc/k_csv.cu) never does.
 *
 * Restates, line by line:   scanfast_float  /root/reference/src/csvtbin.c:403-417
 *                           scanfast_double                          :419-433 (value discarded when converting, :691)
 *                           csv_preread's maximum                    :643-648
 *                           write_tbin's conversion loop             :685-717
 * Pinned: tests/test_csv.py compares it (and the CUDA library) with the payload the UNMODIFIED reference tool
 * (oracle/_ref/csvtbin_ref, compiled by oracle/Makefile from the sources where they lie) writes for the same CSV files, and
 * with the committed golden digests in tests/golden/csv_golden.json.
 * Built with -ffp-contract=off -fno-fast-math: float expressions evaluate as written, as the reference's do at its -O0/-O2.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/rt_csv.h"

int oracle_fail(int code, const char *fmt, ...);              /* scan_oracle.c: sets rt_last_error() */

struct rt_csv { const char *text; char *copy; uint64_t nbytes, nlines; uint64_t *line_start; };

#define MAXLINE 400

int rt_csv_open(int device, const char *text, uint64_t nbytes, rt_csv **out) {
   (void)device;
   if (!out || (!text && nbytes)) return oracle_fail(RT_ERR_ARG, "rt_csv_open: null argument");
   rt_csv *c = calloc(1, sizeof *c);
   c->copy = malloc(nbytes + 1); memcpy(c->copy, text, nbytes); c->copy[nbytes] = 0; c->text = c->copy; c->nbytes = nbytes;
   uint64_t nl = 0;
   for (uint64_t i = 0; i < nbytes; ++i) nl += text[i] == '\n';
   const int open_tail = nbytes > 0 && text[nbytes - 1] != '\n';
   c->nlines = nl + open_tail;
   c->line_start = malloc((c->nlines + 2) * sizeof(uint64_t));
   uint64_t k = 0; c->line_start[0] = 0;
   for (uint64_t i = 0; i < nbytes; ++i) if (text[i] == '\n') c->line_start[++k] = i + 1;
   c->line_start[c->nlines] = nbytes + open_tail;
   *out = c; return RT_OK; }
void rt_csv_close(rt_csv *c) { if (c) { free(c->copy); free(c->line_start); free(c); } }
uint64_t rt_csv_nlines(const rt_csv *c) { return c ? c->nlines : 0; }
int rt_csv_line(const rt_csv *c, uint64_t line, uint64_t *offset, uint64_t *length) {
   if (!c || !offset || !length || line >= c->nlines) return oracle_fail(RT_ERR_ARG, "rt_csv_line: bad argument");
   *offset = c->line_start[line]; *length = c->line_start[line + 1] - c->line_start[line] - 1; return RT_OK; }

/* what fgets(line, MAXLINE, inf) leaves in the buffer for this line: at most MAXLINE-1 characters, zero terminated */
static void get_line(const rt_csv *c, uint64_t line, char buf[MAXLINE + 1]) {
   uint64_t lo = c->line_start[line], hi = c->line_start[line + 1];
   if (hi > c->nbytes) hi = c->nbytes;
   uint64_t n = hi - lo; if (n > MAXLINE - 1) n = MAXLINE - 1;
   memcpy(buf, c->text + lo, n); buf[n] = 0; }

static float scan_float(char **p) {                           /* csvtbin.c:403 */
   float n = 0; int negative = 0;
   while (**p == ' ' || **p == ',') ++*p;
   if (**p == '-') { ++*p; negative = 1; }
   while (**p >= '0' && **p <= '9') n = n * 10 + (*(*p)++ - '0');
   if (**p == '.') {
      float divisor = 10; ++*p;
      while (**p >= '0' && **p <= '9') { n += (*(*p)++ - '0') / divisor; divisor *= 10; } }
   return negative ? -n : n; }
static double scan_double(char **p) {                         /* csvtbin.c:419 */
   double n = 0; int negative = 0;
   while (**p == ' ' || **p == ',') ++*p;
   if (**p == '-') { ++*p; negative = 1; }
   while (**p >= '0' && **p <= '9') n = n * 10 + (*(*p)++ - '0');
   if (**p == '.') {
      double divisor = 10; ++*p;
      while (**p >= '0' && **p <= '9') { n += (*(*p)++ - '0') / divisor; divisor *= 10; } }
   return negative ? -n : n; }

int rt_csv_max_abs(rt_csv *c, uint64_t first_line, uint64_t nlines, uint32_t ntrks, float scalefactor, float *max_abs) {
   if (!c || !max_abs || ntrks < 1 || ntrks > RT_MAXTRKS || first_line > c->nlines || nlines > c->nlines - first_line) return oracle_fail(RT_ERR_ARG, "rt_csv_max_abs: bad argument");
   float maxvolts = 0; char line[MAXLINE + 1];
   for (uint64_t i = 0; i < nlines; ++i) {
      get_line(c, first_line + i, line);
      char *linep = line; (void)scan_double(&linep);
      for (uint32_t trk = 0; trk < ntrks; ++trk) {
         float voltage = scan_float(&linep) * scalefactor;
         if (voltage < 0) voltage = -voltage;
         if (maxvolts < voltage) maxvolts = voltage; } }
   *max_abs = maxvolts; return RT_OK; }

int rt_csv_convert(rt_csv *c, const rt_csv_cfg *cfg, uint64_t first_line, uint64_t nrows, int16_t *rows_out, rt_tape *tape, rt_csv_stats *stats) {
   if (!c || !cfg || cfg->ntrks < 1 || cfg->ntrks > RT_MAXTRKS || cfg->subsample < 1 || !(cfg->maxvolts > 0)) return oracle_fail(RT_ERR_ARG, "rt_csv_convert: bad argument");
   uint32_t seen = 0;
   for (uint32_t k = 0; k < cfg->ntrks; ++k) { if (cfg->track_permutation[k] >= cfg->ntrks) return oracle_fail(RT_ERR_ARG, "rt_csv_convert: bad permutation"); seen |= 1u << cfg->track_permutation[k]; }
   if (seen + 1 != 1u << cfg->ntrks) return oracle_fail(RT_ERR_ARG, "rt_csv_convert: track_permutation is not a permutation");
   if (first_line > c->nlines || nrows > (c->nlines - first_line) / cfg->subsample) return oracle_fail(RT_ERR_ARG, "rt_csv_convert: not that many lines");
   int16_t *rows = rows_out ? rows_out : malloc((size_t)nrows * cfg->ntrks * 2 + 2);
   long long count_toosmall = 0, count_toobig = 0; float maxvolts = 0, minvolts = 0;
   char line[MAXLINE + 1]; float samples[RT_MAXTRKS];
   for (uint64_t r = 0; r < nrows; ++r) {
      get_line(c, first_line + (r + 1) * cfg->subsample - 1, line);
      char *linep = line; (void)scan_double(&linep);
      for (uint32_t trk = 0; trk < cfg->ntrks; ++trk) samples[cfg->track_permutation[trk]] = scan_float(&linep) * cfg->scalefactor;
      for (uint32_t trk = 0; trk < cfg->ntrks; ++trk) {
         float fsample = samples[trk], round;
         if (cfg->invert) fsample = -fsample;
         if (fsample < 0) round = -0.5; else round = 0.5;
         int32_t sample = (int)((fsample / cfg->maxvolts * 32767) + round);
         if (fsample < minvolts) minvolts = fsample;
         if (fsample > maxvolts) maxvolts = fsample;
         if (sample <= -32767) { sample = -32767; ++count_toosmall; }
         if (sample >= 32767) { sample = 32767; ++count_toobig; }
         rows[r * cfg->ntrks + trk] = (int16_t)sample; } }
   int rc = RT_OK;
   if (tape && nrows) rc = rt_upload(tape, rows, nrows);
   if (!rows_out) free(rows);
   if (stats) { memset(stats, 0, sizeof *stats); stats->rows = nrows; stats->too_big = count_toobig; stats->too_small = count_toosmall; stats->minvolts = minvolts; stats->maxvolts = maxvolts; }
   return rc; }
