/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the chain filters of /root/reference/Chain.h:
 *   mode 0  RemoveSmallPairedIndels<Tup>(chain)                      Chain.h:546-606
 *   mode 1  RemovePairedIndels<Tup>(chain, refineEnds = true)        Chain.h:611-748   (float mean / sd of the anchor distances; note the mixed
 *   mode 2  RemovePairedIndels<Tup>(chain, refineEnds = false)                          axes of qDist, :631 / :696, and the int truncation of dist, :703)
 *   mode 3  RemovePairedIndels(matches, chain, lengths)              Chain.h:754-822   (no strands; q/t/len are those of matches[chain[i]])
 *   mode 4  RemoveSpuriousAnchors<Tup>(chain)                        Chain.h:828-890
 *   mode 5  RemoveSpuriousJump<Tup>(chain)                           Chain.h:896-960
 * A chain is its anchors in chain order: q = qStart, t = tStart, len = length, strand.  The result is the keep mask; the reference then
 * compacts chain.chain / ClusterIndex (and link: link[m-1] = link[i-1] for every kept anchor i with m >= 1 anchors kept before it).
 * Pinned by tests/test_chain_filters.py against the unmodified templates (oracle/ref_wrap.cpp: ref_chain_filter). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline int sgn(int v) { return v >= 0; }
static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int imax2(int a, int b) { return a > b ? a : b; }

void lra_oracle_chain_filter(int mode, const uint32_t *q, const uint32_t *t, const uint32_t *len, const uint8_t *strand, long n, uint8_t *keep) {
  for (long i = 0; i < n; i++) keep[i] = 1;
  if (n < 2) return;
  int *SV = (int *)malloc(sizeof(int) * (size_t)n), *SVpos = (int *)malloc(sizeof(int) * (size_t)n);
  long *SVg = (long *)malloc(sizeof(long) * (size_t)n);
  long ns = 0;
  const int thr = mode == 0 ? 5 : (mode == 4 ? 499 : (mode == 5 ? 100 : 30));      /* |Gap| > thr  (mode 4: >= 500) */
  /* `long` in the reference: dist * dist overflows for out-of-order anchor pairs and the stock build wraps; unsigned makes the wrap defined */
  unsigned long totalDist = 0, totDistSq = 0;
#define QE(i) (q[i] + len[i])
#define TE(i) (t[i] + len[i])
  for (long c = 1; c < n; c++) {
    if (mode == 1) {
      long tDist, qDist;
      if (t[c] > TE(c - 1)) tDist = (uint32_t)(t[c] - TE(c - 1)); else tDist = (uint32_t)(t[c - 1] - TE(c));
      if (q[c] > QE(c - 1)) qDist = (uint32_t)(q[c] - TE(c - 1)); else qDist = (uint32_t)(q[c - 1] - QE(c));
      long dist = tDist < qDist ? tDist : qDist;
      totDistSq += (unsigned long)dist * (unsigned long)dist; totalDist += (unsigned long)dist;
    }
    if (mode == 3) {
      int Gap = (int)(((long)t[c] - (long)q[c]) - ((long)t[c - 1] - (long)q[c - 1]));
      if (iabs(Gap) > 30) { SV[ns] = Gap; SVg[ns] = (int)t[c]; SVpos[ns] = (int)c; ns++; }
      continue;
    }
    if (strand[c] == strand[c - 1]) {
      int Gap;
      if (strand[c] == 0) Gap = (int)(((long)t[c] - (long)q[c]) - ((long)t[c - 1] - (long)q[c - 1]));
      else Gap = (int)((long)(uint32_t)(QE(c) + t[c]) - (long)(uint32_t)(QE(c - 1) + t[c - 1]));
      const int take = mode == 0 ? (iabs(Gap) > 5 && iabs(Gap) <= 50) : (iabs(Gap) > thr);
      if (take) { SV[ns] = Gap; SVg[ns] = t[c]; SVpos[ns] = (int)c; ns++; }
    } else { SVg[ns] = t[c]; SVpos[ns] = (int)c; SV[ns] = 0; ns++; }
  }
  if (mode == 0) {
    for (long c = 1; c < ns; c++)
      if (sgn(SV[c]) != sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0 && iabs(SV[c] + SV[c - 1]) <= 20 && SVpos[c] - SVpos[c - 1] < 3)
        for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if (len[i] <= 50) keep[i] = 0;
  } else if (mode == 1 || mode == 2) {
    const float nDist = (float)(n - 1);
    const float meanDist = (float)(long)totalDist / nDist;
    const float varDist = (float)(long)totDistSq / (float)nDist - meanDist * meanDist;
    const float sdDist = sqrtf(varDist);
    int firstValidDist = -1, lastValidDist = -1;
    for (long c = 1; c < ns; c++) {
      if (sgn(SV[c]) != sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0 && iabs(SV[c]) >= 300 && iabs(SV[c - 1]) >= 300 && SVpos[c] - SVpos[c - 1] < 3)
        for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if (len[i] < 100) keep[i] = 0;
      if (sgn(SV[c]) != sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0 && iabs(SV[c] + SV[c - 1]) < 100 && SVpos[c] - SVpos[c - 1] < 3)
        for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if (len[i] < 100) keep[i] = 0;
    }
    if (mode == 1) {
      for (long c = 1; c < n; c++) {
        long tDist, qDist;
        if (t[c] > TE(c - 1)) tDist = (uint32_t)(t[c] - TE(c - 1)); else tDist = (uint32_t)(t[c - 1] - TE(c));
        if (q[c] > QE(c - 1)) qDist = (uint32_t)(q[c] - TE(c - 1)); else qDist = (uint32_t)(q[c - 1] - QE(c));
        int dist = (int)(tDist < qDist ? tDist : qDist);
        if ((float)dist < meanDist + 4 * sdDist) { if (firstValidDist == -1) firstValidDist = (int)c - 1; lastValidDist = (int)c; }
      }
      if (lastValidDist == -1 || firstValidDist == -1) for (long i = 0; i < n; i++) if (len[i] < 100) keep[i] = 0;
      if (firstValidDist > 0 && firstValidDist < 3) for (int i = 0; i < firstValidDist; i++) if (len[i] < 100) keep[i] = 0;
      if (lastValidDist + 1 <= n && n - lastValidDist < 3) for (long i = lastValidDist + 1; i < n; i++) if (len[i] < 100) keep[i] = 0;
    }
  } else if (mode == 3) {
    for (long c = 1; c < ns; c++) {
      const int blink = imax2(iabs(SV[c]), iabs(SV[c - 1]));
      const int g = (int)SVg[c], gp = (int)SVg[c - 1];
      int hit = 0;
      if (sgn(SV[c]) != sgn(SV[c - 1]) && iabs(SV[c] + SV[c - 1]) < 600 && iabs(SV[c]) != 0 && SV[c - 1] != 0) {
        if ((sgn(SV[c]) == 1 && iabs(g - gp) < imax2(2 * blink, 1000)) || (sgn(SV[c]) == 0 && iabs(g - SV[c] - gp) < imax2(2 * blink, 1000))) hit = 1;
      } else if (sgn(SV[c]) != sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0 &&
                 ((sgn(SV[c]) == 1 && iabs(g - gp) < 500) || (sgn(SV[c]) == 0 && iabs(g - SV[c] - gp) < 500))) hit = 1;
      else if (sgn(SV[c]) == sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0) {
        if ((sgn(SV[c]) == 1 && iabs(g - gp) < imax2(2 * blink, 1000)) || (sgn(SV[c]) == 0 && iabs(g - SV[c] - gp) < imax2(2 * blink, 1000))) hit = 1;
      }
      if (hit) for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if ((int)len[i] < 100) keep[i] = 0;
    }
  } else if (mode == 4) {
    for (long c = 1; c < ns; c++)
      if (SV[c] != 0 && SV[c - 1] != 0 && SVpos[c] - SVpos[c - 1] <= 10) {
        int check = 0;
        for (int b = SVpos[c - 1]; b < SVpos[c]; b++) if (len[b] >= 50) { check = 1; break; }
        if (!check) for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if (len[i] < 50) keep[i] = 0;
      }
  } else {
    for (long c = 1; c < ns; c++)
      if (keep[SVpos[c - 1]] == 1 && sgn(SV[c]) != sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0 && SVpos[c] - SVpos[c - 1] == 1)
        for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if (len[i] < 50) keep[i] = 0;
  }
  free(SV); free(SVpos); free(SVg);
#undef QE
#undef TE
}
