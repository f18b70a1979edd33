/* Stand-in for htslib/kseq.h, written for the oracle build only (TEST INFRASTRUCTURE).
 * htslib (pinned 1.11 in the reference's subprojects/htslib.wrap) is not present in this image and
 * supplies no arithmetic of the MapRead hot path: the reference uses it to parse FASTA/FASTQ
 * (Genome.h:122-137, MMIndex.h:248-253,296-312, Input.h:21). This header provides the small
 * subset of the kseq API those call sites need: KSEQ_INIT, kseq_t{name,comment,seq,qual},
 * kseq_init / kseq_read / kseq_destroy.  Semantics kept: name = first whitespace-delimited token
 * of the header line, sequence returned verbatim with line breaks removed. */
#ifndef LRA_B200_ORACLE_SHIM_KSEQ_H
#define LRA_B200_ORACLE_SHIM_KSEQ_H
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#ifndef KSTRING_T
#define KSTRING_T kstring_t
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
#endif

#define KSEQ_SHIM_BUFSZ 65536

#define KSEQ_INIT(type_t, read_fn)                                                               \
  typedef struct kseq_t {                                                                        \
    kstring_t name, comment, seq, qual;                                                          \
    type_t f; unsigned char *buf; int beg, end, eof, last;                                       \
  } kseq_t;                                                                                      \
  static inline kseq_t *kseq_init(type_t fd) {                                                   \
    kseq_t *ks = (kseq_t*)calloc(1, sizeof(kseq_t));                                             \
    ks->f = fd; ks->buf = (unsigned char*)malloc(KSEQ_SHIM_BUFSZ); ks->last = 0; return ks; }    \
  static inline void kseq_destroy(kseq_t *ks) {                                                  \
    if (!ks) return; free(ks->name.s); free(ks->comment.s); free(ks->seq.s); free(ks->qual.s);   \
    free(ks->buf); free(ks); }                                                                   \
  static inline int kseq_shim_getc(kseq_t *ks) {                                                 \
    if (ks->beg >= ks->end) {                                                                    \
      if (ks->eof) return -1;                                                                    \
      ks->beg = 0; ks->end = read_fn(ks->f, ks->buf, KSEQ_SHIM_BUFSZ);                           \
      if (ks->end <= 0) { ks->eof = 1; ks->end = 0; return -1; } }                               \
    return ks->buf[ks->beg++]; }                                                                 \
  static inline void kseq_shim_push(kstring_t *s, int c) {                                       \
    if (s->l + 2 > s->m) { s->m = s->m ? s->m * 2 : 256; s->s = (char*)realloc(s->s, s->m); }    \
    s->s[s->l++] = (char)c; s->s[s->l] = 0; }                                                    \
  static inline int kseq_read(kseq_t *ks) {                                                      \
    int c;                                                                                       \
    if (ks->last == 0) {                                                                         \
      while ((c = kseq_shim_getc(ks)) != -1 && c != '>' && c != '@') {}                          \
      if (c == -1) return -1;                                                                    \
      ks->last = c; }                                                                            \
    ks->name.l = ks->comment.l = ks->seq.l = ks->qual.l = 0;                                     \
    kseq_shim_push(&ks->name, 0); ks->name.l = 0; kseq_shim_push(&ks->seq, 0); ks->seq.l = 0;    \
    while ((c = kseq_shim_getc(ks)) != -1 && !isspace(c)) kseq_shim_push(&ks->name, c);          \
    if (c != '\n' && c != -1)                                                                    \
      while ((c = kseq_shim_getc(ks)) != -1 && c != '\n') kseq_shim_push(&ks->comment, c);       \
    while ((c = kseq_shim_getc(ks)) != -1 && c != '>' && c != '+' && c != '@') {                 \
      if (c == '\n' || c == '\r') continue;                                                      \
      kseq_shim_push(&ks->seq, c); }                                                             \
    if (c == '>' || c == '@') ks->last = c;                                                      \
    if (c != '+') { if (c == -1) ks->last = 0, ks->eof = 1; return (int)ks->seq.l; }             \
    while ((c = kseq_shim_getc(ks)) != -1 && c != '\n') {}                                       \
    while (ks->qual.l < ks->seq.l && (c = kseq_shim_getc(ks)) != -1) {                           \
      if (c == '\n' || c == '\r') continue;                                                      \
      kseq_shim_push(&ks->qual, c); }                                                            \
    ks->last = 0;                                                                                \
    return (int)ks->seq.l; }

#endif
