/* include/rt_csv.h -- C-ABI of the CSV ingest: the .csv -> .tbin conversion of the reference's csvtbin tool
 * (write_tbin, src/csvtbin.c:661-747, with csv_preread :619-657 and the scanfast_float / scanfast_double parsers :403-433),
 * with the text parsed on the GPU.  Same conventions as rt_scan.h: plain pointers and sizes, 0 = RT_OK, negative = RT_ERR_*
 * with the text in rt_last_error().  Two libraries export it: the product (librt_scan_b200.so, CUDA, no CPU fallback) and the
 * test oracle (oracle/_ref/libscan_oracle.so, a CPU restatement of the same reference lines).
 *
 * A capture exported by the logic analyser is text: title lines, then one line per sample,
 *        time, v0, v1, ... v(ntrks-1)
 * csvtbin reads it twice with fgets(line, 400): once over the first 999,999 data lines for the sample period and the
 * largest |voltage| (csv_preread), once to convert.  Here the text goes to the device once (rt_csv_open builds the index of
 * line starts there), rt_csv_max_abs() is csv_preread's maximum, and rt_csv_convert() is the conversion loop: every int16 it
 * produces is bit-identical to csvtbin's, because the number parser is the reference's float recurrence
 * (n = n*10 + d;  n += d / divisor, divisor *= 10) evaluated in the same order without FMA contraction.
 *
 * What stays with the caller (readtape_b200/host/csvtbin_b200.c does it the way csvtbin.c does): the options, the two time
 * stamps csv_preread needs (rt_csv_line() says where a line is in the caller's own text), the TBIN header, -skip / -starttime /
 * -stopaft / -endtime (as first_line / nrows), and the -redo rule.
 *
 * Lines are what fgets would return: they end at '\n' (a last line without one counts); only the first 399 characters of a
 * line are looked at (fgets(line, MAXLINE = 400)); csvtbin would see the rest of a longer line as further lines, which no
 * capture file has -- rt_csv_convert does not reproduce that.
 */
#ifndef RT_CSV_H
#define RT_CSV_H
#include "rt_scan.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_csv rt_csv;             /* a CSV text resident on the device, with the index of its lines */

typedef struct rt_csv_cfg {
   uint32_t ntrks;                        /* voltage columns after the time stamp (csvtbin -ntrks) */
   uint32_t track_permutation[RT_MAXTRKS];/* column c is written to position track_permutation[c] of the row (csvtbin.c:693, -order) */
   float    maxvolts;                     /* full scale of the int16 samples (hdr.u.s.maxvolts, csvtbin.c:706) */
   float    scalefactor;                  /* -scale (csvtbin.c:693); 1 = none */
   uint32_t invert;                       /* TBIN_INVERTED (csvtbin.c:704) */
   uint32_t subsample;                    /* of every `subsample` lines the last one is converted (csvtbin.c:686); >= 1 */
} rt_csv_cfg;

typedef struct rt_csv_stats {
   uint64_t rows;                         /* rows produced */
   uint64_t too_big, too_small;           /* samples clamped to +32767 / -32767 (csvtbin.c:712-715) */
   float    minvolts, maxvolts;           /* the extremes csvtbin logs (csvtbin.c:708-709; both start at 0) */
   double   ms_convert;                   /* device time of the conversion kernel (product library) */
} rt_csv_stats;

/* Put `nbytes` of CSV text (host memory) on `device` and index its lines. */
int      rt_csv_open(int device, const char *text, uint64_t nbytes, rt_csv **out);
void     rt_csv_close(rt_csv *csv);
/* lines in the text, title lines included */
uint64_t rt_csv_nlines(const rt_csv *csv);
/* where line `line` is in the text given to rt_csv_open: offset of its first character and its length without the '\n' */
int      rt_csv_line(const rt_csv *csv, uint64_t line, uint64_t *offset, uint64_t *length);
/* csv_preread's maximum (csvtbin.c:645-648): max over lines [first_line, first_line + nlines) and `ntrks` columns of
 * |voltage * scalefactor| -- BEFORE the reference rounds it up ((int)((m + 0.55f) * 10) / 10, csvtbin.c:649) */
int      rt_csv_max_abs(rt_csv *csv, uint64_t first_line, uint64_t nlines, uint32_t ntrks, float scalefactor, float *max_abs);
/* The conversion loop of write_tbin: row r (0 <= r < nrows) comes from line first_line + (r + 1) * subsample - 1.
 * rows_out (host memory, nrows * ntrks int16, row-major = the TBIN payload without its 0x8000 end marker) and `tape` are each
 * optional: with a tape (opened with nheads == ntrks) the rows are appended to it on the device as rt_upload() would, without
 * leaving the GPU.  first_line + nrows * subsample must not exceed rt_csv_nlines(). */
int      rt_csv_convert(rt_csv *csv, const rt_csv_cfg *cfg, uint64_t first_line, uint64_t nrows,
                        int16_t *rows_out, rt_tape *tape, rt_csv_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
