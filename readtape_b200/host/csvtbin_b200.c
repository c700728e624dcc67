/* readtape_b200/host/csvtbin_b200.c -- "csvtbin with the B200 parser": converts a logic-analyser .csv capture into the .tbin
 * file readtape reads, as the reference's csvtbin tool does in its write direction (src/csvtbin.c: main :752-850, csv_preread
 * :619-657, write_tbin :661-747, write_tbin_hdr :596-617), with the text parsed on the GPU through the C-ABI of
 * include/rt_csv.h.  Same command line for that direction:
 *
 *     csvtbin_b200 -ntrks=9 -order=01234567p -nrzi -bpi=800 -ips=50 [-maxvolts=x] [-scale=x] [-invert] [-reverse]
 *                  [-subsample=n] [-skip=n] [-stopaft=n] [-starttime=s] [-endtime=s] [-descr=txt] [-redo]  basefilename
 *
 * reads basefilename.csv, writes basefilename.tbin and basefilename.csvtbin.log.  The .tbin is byte-identical to csvtbin's
 * apart from the 36 bytes of the header that hold the time of the conversion (tests/test_csv.py).  Not here: -read /
 * -showheader (TBIN -> CSV, a debugging aid), -graph, -datewritten/-dateread.
 *
 * This file is host glue (options, the two time stamps of csv_preread, the TBIN header, the -redo rule); it links against
 * librt_scan_b200.so and has no parser of voltages of its own: without the GPU it fails.
 */
#define _FILE_OFFSET_BITS 64
#include <ctype.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include "rt_csv.h"

#define PREREAD_COUNT 1000000          /* csvtbin.c:96 */
#define TITLE_LINES 2                  /* the two Saleae title lines, csvtbin.c:664-665 */
enum { M_UNKNOWN = 0, M_PE = 1, M_NRZI = 2, M_GCR = 4, M_WW = 8 };     /* enum mode_t of the reference (decoder.h) */
#define F_NO_REORDER 1
#define F_TRKORDER 2
#define F_INVERTED 4
#define F_REVERSED 8

static FILE *logf_;
static void say(const char *fmt, ...) {
   va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap);
   if (logf_) { va_start(ap, fmt); vfprintf(logf_, fmt, ap); va_end(ap); } }
static void die(const char *fmt, ...) {
   va_list ap; va_start(ap, fmt); fprintf(stderr, "csvtbin_b200: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n"); va_end(ap);
   exit(8); }
static const char *commas(unsigned long long n) {
   static char buf[4][40]; static int k; char tmp[32]; char *out = buf[k = (k + 1) & 3];
   int len = snprintf(tmp, sizeof tmp, "%llu", n), o = 0;
   for (int i = 0; i < len; ++i) { out[o++] = tmp[i]; if ((len - 1 - i) % 3 == 0 && i != len - 1) out[o++] = ','; }
   out[o] = 0; return out; }

/* option helpers: "-KEY=value", key matched without regard to case, '/' allowed as the switch character */
static const char *opt_val(const char *arg, const char *key) {
   size_t n = strlen(key);
   return strncasecmp(arg, key, n) == 0 ? arg + n : NULL; }

static double first_number(const char *s, size_t len) {                /* scanfast_double, csvtbin.c:419-433 */
   size_t i = 0; double n = 0; int neg = 0;
   while (i < len && (s[i] == ' ' || s[i] == ',')) ++i;
   if (i < len && s[i] == '-') { ++i; neg = 1; }
   while (i < len && isdigit((unsigned char)s[i])) n = n * 10 + (s[i++] - '0');
   if (i < len && s[i] == '.') {
      double divisor = 10; ++i;
      while (i < len && isdigit((unsigned char)s[i])) { n += (s[i++] - '0') / divisor; divisor *= 10; } }
   return neg ? -n : n; }

static void put4(FILE *f, uint32_t v) { unsigned char b[4] = {v & 0xff, v >> 8 & 0xff, v >> 16 & 0xff, v >> 24}; fwrite(b, 1, 4, f); }
static void putf(FILE *f, float x) { uint32_t v; memcpy(&v, &x, 4); put4(f, v); }

int main(int argc, char **argv) {
   unsigned ntrks = 9, subsample = 1, flags = 0, mode = M_UNKNOWN; int have_order = 0, redo = 0;
   uint32_t perm[RT_MAXTRKS]; char ww_order[RT_MAXTRKS + 1] = "";
   float bpi = 0, ips = 0, maxvolts_opt = 0, scale = 1.0f; char descr[80] = "";
   uint64_t skip = 0, stopaft = UINT64_MAX, starttime = 0, endtime = UINT64_MAX;
   const char *order_arg = NULL;
   printf("csvtbin_b200: convert .CSV to .TBIN with the parser on the GPU (%s)\n", rt_backend());
   int argn = 1;
   for (; argn < argc && (argv[argn][0] == '-' || argv[argn][0] == '/'); ++argn) {
      const char *a = argv[argn] + 1, *v;
      if ((v = opt_val(a, "NTRKS="))) { if (have_order) die("can't give -ntrks after -order"); ntrks = (unsigned)atoi(v); if (ntrks < 5 || ntrks > RT_MAXTRKS) die("bad -ntrks"); }
      else if ((v = opt_val(a, "ORDER="))) { order_arg = v; have_order = 1;
         if (mode == M_WW) {                                            /* csvtbin.c:322-328: Whirlwind keeps the string for readtape */
            if (strlen(v) > RT_MAXTRKS) die("Whirlwind -order string too long: %s", v);
            strcpy(ww_order, v); ntrks = (unsigned)strlen(v); flags |= F_TRKORDER | F_NO_REORDER; have_order = 2;
            printf("using Whirlwind -order=%s and ntrks=%d\n", v, ntrks); }
         else {                                                         /* csvtbin.c:329-340 */
            if (strlen(v) != ntrks) die("bad track order at %s", v);
            unsigned seen = 0;
            for (unsigned i = 0; i < ntrks; ++i) {
               unsigned ch = (unsigned char)v[i];
               if (toupper(ch) == 'P') ch = ntrks - 1;
               else { if (!isdigit(ch) || (ch -= '0') > ntrks - 2) die("bad track order at %s", v); }
               perm[i] = ch; seen |= 1u << ch; }
            if (seen + 1 != 1u << ntrks) die("bad track order at %s", v); } }
      else if (!strcasecmp(a, "NRZI")) mode = M_NRZI;
      else if (!strcasecmp(a, "PE")) mode = M_PE;
      else if (!strcasecmp(a, "GCR")) mode = M_GCR;
      else if (!strcasecmp(a, "WHIRLWIND")) mode = M_WW;
      else if (!strcasecmp(a, "INVERT")) flags |= F_INVERTED;
      else if (!strcasecmp(a, "REVERSE")) flags |= F_REVERSED;
      else if (!strcasecmp(a, "REDO")) redo = 1;
      else if ((v = opt_val(a, "BPI="))) bpi = (float)atof(v);
      else if ((v = opt_val(a, "IPS="))) ips = (float)atof(v);
      else if ((v = opt_val(a, "MAXVOLTS="))) { maxvolts_opt = (float)atof(v); if (maxvolts_opt < 0.1f || maxvolts_opt > 15.0f) die("bad -maxvolts"); }
      else if ((v = opt_val(a, "SCALE="))) { scale = (float)atof(v); if (scale < 1e-4f || scale > 1e4f) die("bad -scale"); }
      else if ((v = opt_val(a, "DESCR="))) { strncpy(descr, v, sizeof descr); descr[sizeof descr - 1] = 0; }
      else if ((v = opt_val(a, "SKIP="))) skip = strtoull(v, NULL, 10);
      else if ((v = opt_val(a, "STOPAFT="))) { stopaft = strtoull(v, NULL, 10); if (!stopaft) die("bad -stopaft"); }
      else if ((v = opt_val(a, "SUBSAMPLE="))) { subsample = (unsigned)atoi(v); if (subsample < 1) die("bad -subsample"); }
      else if ((v = opt_val(a, "STARTTIME="))) starttime = (uint64_t)((double)(float)atof(v) * 1e9);
      else if ((v = opt_val(a, "ENDTIME="))) endtime = (uint64_t)((double)(float)atof(v) * 1e9);
      else die("bad option: %s (the TBIN -> CSV direction, -graph and the date options are not in this tool)", argv[argn]); }
   if (argn != argc - 1) die("usage: csvtbin_b200 <options> <basefilename>");
   if (starttime >= endtime) die("starttime is after endtime");
   (void)order_arg;
   const char *base = argv[argn];
   char inname[4096], outname[4096], logname[4096];
   snprintf(inname, sizeof inname, "%s.csv", base); snprintf(outname, sizeof outname, "%s.tbin", base); snprintf(logname, sizeof logname, "%s.csvtbin.log", base);
   logf_ = fopen(logname, "w"); if (!logf_) die("file create failed for %s", logname);
   say("command line: "); for (int i = 0; i < argc; ++i) say("%s ", argv[i]); say("\n");
   say("opening  %s\n", inname);
   int fd = open(inname, O_RDONLY); if (fd < 0) die("unable to open input file %s", inname);
   struct stat sb; fstat(fd, &sb);
   const uint64_t nbytes = (uint64_t)sb.st_size;
   const char *text = nbytes ? mmap(NULL, nbytes, PROT_READ, MAP_PRIVATE, fd, 0) : "";
   if (text == MAP_FAILED) die("can't map %s", inname);
   say("creating %s\n", outname);
   if (have_order != 1) {                                               /* csvtbin.c:805-810 */
      if (!(flags & F_TRKORDER)) { say("WARNING: using the default track ordering, and marking the .tbin file to show it wasn't given\n"); flags |= F_NO_REORDER; }
      for (unsigned i = 0; i < ntrks; ++i) perm[i] = i; }
   say("input track order: ");
   if (flags & F_TRKORDER) say("%s", ww_order);
   else for (unsigned i = 0; i < ntrks; ++i) { if (perm[i] == ntrks - 1) say("p"); else say("%d", perm[i]); }
   say("\n");
   if (flags & F_INVERTED) say("the data will be inverted\n");
   if (flags & F_REVERSED) say("the tape might have been read or written backwards\n");
   if (scale != 1.0f) say("input voltages will be scaled by %f\n", scale);

   struct timespec w0, w1; clock_gettime(CLOCK_MONOTONIC, &w0);
   rt_csv *csv = NULL;
   if (rt_csv_open(0, text, nbytes, &csv) != RT_OK) die("%s", rt_last_error());
   const uint64_t nlines = rt_csv_nlines(csv);
   if (nlines < TITLE_LINES + 2) die("no data lines in %s", inname);
   const uint64_t ndata = nlines - TITLE_LINES;

   /* ---- csv_preread (csvtbin.c:619-657) ---- */
   uint64_t off, len;
   rt_csv_line(csv, 1, &off, &len);
   unsigned numcommas = 0; for (uint64_t i = 0; i < len && i < 399; ++i) numcommas += text[off + i] == ',';
   if (numcommas != ntrks) say("*** WARNING *** file has %d columns of data, but ntrks=%d\n", numcommas, ntrks);
   const uint64_t npre = ndata < PREREAD_COUNT - 1 ? ndata : PREREAD_COUNT - 1;
   rt_csv_line(csv, TITLE_LINES, &off, &len);
   const double first_timestamp = first_number(text + off, len < 399 ? len : 399);
   if (first_timestamp < 0) die("the first time stamp is negative: csvtbin's period estimate is undefined for it");
   rt_csv_line(csv, TITLE_LINES + npre - 1, &off, &len);
   const double timestamp = first_number(text + off, len < 399 ? len : 399);
   uint64_t tstart = (uint64_t)((first_timestamp + 0.5e-9) * 1e9);
   uint32_t tdelta = (uint32_t)(((timestamp - first_timestamp) / (double)((int)npre - 1) + 0.5e-9) * 1e9);
   float seen_max = 0;
   if (rt_csv_max_abs(csv, TITLE_LINES, npre, ntrks, scale, &seen_max) != RT_OK) die("%s", rt_last_error());
   float pre_maxvolts = ((float)(int)((seen_max + 0.55f) * 10.0f)) / 10.0f;
   say("after %s samples, the sample delta is %.2lf usec (%u nsec), samples start at %.6lf seconds, and the rounded-up maximum voltage is %.1fV\n",
       commas(npre), (double)tdelta / 1e3, tdelta, (double)tstart / 1e9, pre_maxvolts);
   if (subsample > 1) {
      tstart += (uint32_t)((subsample - 1) * tdelta); tdelta *= subsample;   /* csvtbin.c:653-654: unsigned 32-bit products, wrap included */
      say("for subsampling every %d samples, we adjusted the delta to %.2lf usec (%u nsec), and the sample start to %.6lf seconds\n", subsample, (double)tdelta / 1e3, tdelta, (double)tstart / 1e9); }
   float hdr_maxvolts = maxvolts_opt;
   if (hdr_maxvolts == 0) hdr_maxvolts = pre_maxvolts;
   else if (hdr_maxvolts < pre_maxvolts) { say("maxvolts was increased from %.1f to %.1f\n", hdr_maxvolts, pre_maxvolts); hdr_maxvolts = pre_maxvolts; }
   else say("we used maxvolts=%.1f\n", hdr_maxvolts);

   /* ---- which lines (csvtbin.c:668-678, :719-720) ---- */
   uint64_t first_line = TITLE_LINES, sample_time = tstart;
   if (skip > 0 || starttime > 0) {
      uint64_t skipped = 0;
      do { if (first_line >= nlines) die("endfile with samples left to skip"); ++first_line; sample_time += tdelta; ++skipped; if (skip > 0) --skip; }
      while (sample_time < starttime || skip > 0);
      say("skipped %s samples\n", commas(skipped)); }
   uint64_t nrows = (nlines - first_line) / subsample;
   if (nrows > stopaft) nrows = stopaft;
   if (endtime != UINT64_MAX && nrows > 0) {                            /* the loop stops after the row that brings sample_time past endtime */
      uint64_t n = endtime >= sample_time ? (endtime - sample_time) / tdelta + 1 : 1;
      if (n < nrows) nrows = n; }

   int16_t *rows = nrows ? rt_host_alloc((size_t)nrows * ntrks * 2) : NULL;
   if (nrows && !rows) die("no pinned memory for %s rows", commas(nrows));
   rt_csv_cfg cfg; memset(&cfg, 0, sizeof cfg);
   cfg.ntrks = ntrks; memcpy(cfg.track_permutation, perm, sizeof(uint32_t) * ntrks); cfg.scalefactor = scale; cfg.invert = (flags & F_INVERTED) != 0; cfg.subsample = subsample;
   for (int tries = 0; tries < 2; ++tries) {
      cfg.maxvolts = hdr_maxvolts;
      rt_csv_stats st;
      if (rt_csv_convert(csv, &cfg, first_line, nrows, rows, NULL, &st) != RT_OK) die("%s", rt_last_error());
      FILE *outf = fopen(outname, "wb"); if (!outf) die("file create failed for %s", outname);
      /* ---- write_tbin_hdr (csvtbin.c:596-617; layout src/csvtbin.h:50-96) ---- */
      char tag[8] = "TBINHDR"; fwrite(tag, 1, 8, outf); fwrite(descr, 1, 80, outf);
      const unsigned hdrsize = 8 + 80 + 4 * (2 + 27 + 9);
      put4(outf, hdrsize); put4(outf, 1);
      time_t now = time(NULL); struct tm *tm = localtime(&now);
      for (int i = 0; i < 18; ++i) put4(outf, 0);                      /* time_written, time_read: not given */
      put4(outf, tm->tm_sec); put4(outf, tm->tm_min); put4(outf, tm->tm_hour); put4(outf, tm->tm_mday); put4(outf, tm->tm_mon); put4(outf, tm->tm_year);
      put4(outf, tm->tm_wday); put4(outf, tm->tm_yday); put4(outf, tm->tm_isdst);
      put4(outf, flags); put4(outf, ntrks); put4(outf, tdelta); putf(outf, hdr_maxvolts); put4(outf, 0); put4(outf, 0); put4(outf, mode); putf(outf, bpi); putf(outf, ips);
      if (flags & F_TRKORDER) { char ext[8 + RT_MAXTRKS + 1]; memset(ext, 0, sizeof ext); strcpy(ext, "TBINORD"); strcpy(ext + 8, ww_order); fwrite(ext, 1, sizeof ext, outf); }
      unsigned char dat[8] = {'D', 'A', 'T', 0, 0, 16, 0, 0}; fwrite(dat, 1, 8, outf);
      put4(outf, (uint32_t)tstart); put4(outf, (uint32_t)(tstart >> 32));
      if (nrows && fwrite(rows, (size_t)ntrks * 2, nrows, outf) != nrows) die("can't write the samples");
      unsigned char endmark[2] = {0x00, 0x80}; fwrite(endmark, 1, 2, outf);
      fclose(outf);
      say("\ndone; minimum voltage was %.1fV, maximum voltage was %.1fV\n", st.minvolts, st.maxvolts);
      if (st.too_big) say("*** WARNING ***  %s samples were too big\n", commas(st.too_big));
      if (st.too_small) say("*** WARNING ***  %s samples were too small\n", commas(st.too_small));
      if (!(st.too_big || st.too_small)) break;
      float newmax = st.maxvolts > -st.minvolts ? st.maxvolts : -st.minvolts;
      if (!redo) { say("you should specify -maxvolts=%.1f\n", newmax + 0.1); break; }
      hdr_maxvolts = ((float)(int)((newmax + 0.15) * 10.0f)) / 10.0f;   /* csvtbin.c:737 */
      say("redoing the conversion with -maxvolts=%.1f\n", hdr_maxvolts); }
   clock_gettime(CLOCK_MONOTONIC, &w1);
   say("%s samples representing %.3lf tape seconds were processed in %.1f seconds\n", commas(nrows), (float)((double)nrows * tdelta) / 1e9,
       (w1.tv_sec - w0.tv_sec) + (w1.tv_nsec - w0.tv_nsec) / 1e9);
   if (rows) rt_host_free(rows);
   rt_csv_close(csv);
   fclose(logf_);
   return 0; }
