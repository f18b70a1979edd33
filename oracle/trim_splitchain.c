/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of TrimSplitChainDiagonal (/root/reference/ChainRefine.h:189-331; Map_lowacc.h:331): after
 * Refine_splitchain, the refined anchors of a split chain are put in Cartesian order and those whose diagonal (second.pos - first.pos, GenomePos
 * arithmetic) lies more than 100 off the smaller diagonal of the two chain anchors around them are dropped.  Forward chains walk all chain anchors;
 * reverse chains only use the first two (the reference's loop body is a plain block there) and walk the anchors from the end.  Chains of one anchor are
 * left untouched (not even sorted).  The trailing "inline" pass of the reference only writes its local flag vector: no effect.
 * One split chain: cq/ct = qStart / tStart of its chain anchors in sptc order; q/t = the refined anchors (sorted in place unless n_chain == 1);
 * keep[i] = 1 for the anchors that stay.  Returns nRemoved.
 * Pinned by tests/test_trim_splitchain.py against the unmodified reference (oracle/ref_wrap.cpp: ref_trim_splitchain). */
#include <stdint.h>
#include <stdlib.h>

typedef struct { uint32_t q, t; } tp;
static int cart_cmp(const void *a, const void *b) {
  const tp *x = (const tp *)a, *y = (const tp *)b;
  if (x->q != y->q) return x->q < y->q ? -1 : 1;
  return x->t < y->t ? -1 : (x->t > y->t ? 1 : 0);
}

long lra_oracle_trim_splitchain(const uint32_t *cq, const uint32_t *ct, int n_chain, int strand, uint32_t *q, uint32_t *t, int n, uint8_t *keep) {
  for (int i = 0; i < n; i++) keep[i] = 1;
  if (n_chain == 1) return 0;
  if (n > 1) {
    tp *v = (tp *)malloc((size_t)n * sizeof(tp));
    for (int i = 0; i < n; i++) { v[i].q = q[i]; v[i].t = t[i]; }
    qsort(v, (size_t)n, sizeof(tp), cart_cmp);
    for (int i = 0; i < n; i++) { q[i] = v[i].q; t[i] = v[i].t; }
    free(v);
  }
  long removed = 0;
  const long offset = 100;
#define CD(i) ((uint32_t)(ct[i] - cq[i]))
  if (strand == 0) {
    int ci = 0, mi = 0;
    while (ci < n_chain - 1) {
      const uint32_t m = CD(ci) < CD(ci + 1) ? CD(ci) : CD(ci + 1);
      const long minDiag = (long)m - offset, maxDiag = (long)m + offset;
      while (mi < n && q[mi] < cq[ci + 1]) {
        const long d = (long)(uint32_t)(t[mi] - q[mi]);
        if (d < minDiag || d > maxDiag) { keep[mi] = 0; removed++; }
        mi++;
      }
      ci++;
    }
  } else {
    int mi = n - 1;
    const uint32_t m = CD(0) < CD(1) ? CD(0) : CD(1);
    const long minDiag = (long)m - offset, maxDiag = (long)m + offset;
    while (mi >= 0 && q[mi] > cq[1]) {
      const long d = (long)(uint32_t)(t[mi] - q[mi]);
      if (d < minDiag || d > maxDiag) { keep[mi] = 0; removed++; }
      mi--;
    }
  }
#undef CD
  return removed;
}
