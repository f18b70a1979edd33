/* Stand-in for htslib/hts.h (oracle build only; SAM/BAM read input is unsupported in the oracle:
 * hts_open returns NULL and the reference then reports "Cannot determine format", Input.h:146-151). */
#ifndef LRA_B200_ORACLE_SHIM_HTS_H
#define LRA_B200_ORACLE_SHIM_HTS_H
#include <stddef.h>
#include <stdint.h>
#include <zlib.h>
#ifndef KSTRING_T
#define KSTRING_T kstring_t
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
#endif
enum htsExactFormat { unknown_format, binary_format, text_format, sam, bam, bai, cram };
typedef struct htsFormat { enum htsExactFormat format; } htsFormat;
typedef struct htsFile { int dummy; } htsFile;
static inline htsFile *hts_open(const char *, const char *) { return NULL; }
static inline int hts_close(htsFile *) { return 0; }
static inline const htsFormat *hts_get_format(htsFile *) { return NULL; }
static inline const char *hts_format_file_extension(const htsFormat *) { return "?"; }
static const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
#endif
