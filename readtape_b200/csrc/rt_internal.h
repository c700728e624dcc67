/* readtape_b200/csrc/rt_internal.h -- shared by the translation units that implement the C-ABI (not exported) */
#ifndef RT_INTERNAL_H
#define RT_INTERNAL_H
/* sets the text rt_last_error() returns on this thread and returns `code` */
__attribute__((visibility("hidden"), format(printf, 2, 3))) int rt_fail(int code, const char *fmt, ...);
#endif
