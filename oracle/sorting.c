/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the anchor sorts of /root/reference/Sorting.h:
 *   mode 0  DiagonalSort        key ((long) q - (long) t, q)                  Sorting.h:33-75
 *   mode 1  AntiDiagonalSort    key ((GenomePos)(q + t) -- 32-bit wrap --, q) Sorting.h:77-139
 *   mode 2  CartesianSort       key (q, t)                                    Sorting.h:141-169
 *   mode 3  CartesianTargetSort key (t, q)                                    Sorting.h:182-209
 * Each comparator is a total order on the position pair (equal keys <=> identical anchors), so the result does not depend on the sorting
 * algorithm; perm is made unique by breaking such ties on the source index.
 * Pinned by tests/test_sorting.py against the unmodified reference (oracle/ref_wrap.cpp: ref_sort_matches). */
#include <stdint.h>
#include <stdlib.h>

typedef struct { uint64_t p; uint32_t s, idx, q, t; } srec;
static int cmp_srec(const void *a, const void *b) {
  const srec *x = (const srec *)a, *y = (const srec *)b;
  if (x->p != y->p) return x->p < y->p ? -1 : 1;
  if (x->s != y->s) return x->s < y->s ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}
void lra_oracle_sort_matches(int mode, uint32_t *q, uint32_t *t, long n, uint32_t *perm) {
  srec *v = (srec *)malloc(sizeof(srec) * (size_t)(n > 0 ? n : 1));
  for (long i = 0; i < n; i++) {
    v[i].q = q[i]; v[i].t = t[i]; v[i].idx = (uint32_t)i;
    if (mode == 0) { v[i].p = (uint64_t)((int64_t)q[i] - (int64_t)t[i] + ((int64_t)1 << 32)); v[i].s = q[i]; }
    else if (mode == 1) { v[i].p = (uint32_t)(q[i] + t[i]); v[i].s = q[i]; }
    else if (mode == 2) { v[i].p = q[i]; v[i].s = t[i]; }
    else { v[i].p = t[i]; v[i].s = q[i]; }
  }
  qsort(v, (size_t)n, sizeof(srec), cmp_srec);
  for (long i = 0; i < n; i++) { q[i] = v[i].q; t[i] = v[i].t; if (perm) perm[i] = v[i].idx; }
  free(v);
}
