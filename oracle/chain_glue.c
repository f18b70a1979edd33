/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of two small chain-glue routines of SURVEY.md 8(a) row a11:
 *   MergeChain     /root/reference/ChainRefine.h:767-802       group adjacent clusters of a split chain that lie within 500 bases on both axes
 *   switchindex    /root/reference/Mapping_ultility.h:39-161   map a chain over split clusters back to the clusters they came from
 * Pinned by tests/test_chain_glue.py against the unmodified reference (oracle/ref_wrap.cpp: ref_merge_chain, ref_switchindex). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* One split chain: sp[0..n) = cluster indices; per cluster chrom, strand, box[4c..] = qStart, qEnd, tStart, tEnd.
 * head[t] = 1 iff entry t starts a new Merge_SplitChain (mergeinfo element); returns the number of groups. */
long lra_oracle_merge_chain(const int32_t *sp, int n, const int32_t *chrom, const uint8_t *strand, const uint32_t *box, uint8_t *head) {
  long groups = 0;
  for (int t = 0; t < n; t++) {
    if (t == 0) { head[0] = 1; groups++; continue; }
    const int cur = sp[t], prev = sp[t - 1];
    int qdist = 9999, tdist = 9999;
    if (chrom[prev] == chrom[cur] && strand[prev] == strand[cur]) {
      const uint32_t pqs = box[4 * prev], pts = box[4 * prev + 2], pte = box[4 * prev + 3];
      const uint32_t cqe = box[4 * cur + 1], cts = box[4 * cur + 2], cte = box[4 * cur + 3];
      qdist = (pqs > cqe) ? (int)(pqs - cqe) : 0;
      if (strand[prev] == 0) tdist = (pts >= cte) ? (int)(pts - cte) : 9999;
      else if (strand[prev] == 1) tdist = (pte <= cts) ? (int)(cts - pte) : 9999;
    }
    if (qdist <= 500 && tdist <= 500) head[t] = 0;
    else { head[t] = 1; groups++; }
  }
  return groups;
}

/* switchindex for one chain.  ch[0..n) = split-cluster indices, link[0..n-1) (n_link = 0 when the chain carries no links), coarse[] maps a split
 * cluster to its cluster, cq[2c], cq[2c+1] = qStart, qEnd of cluster c.  The chain and its links are rewritten in place; returns the new chain
 * length, *n_link_out the new number of links.  (The last pass writes link[sc - 1] = link[c - 1] and resizes link to sc - 1 whatever its size
 * was; a chain without links stays without.) */
long lra_oracle_switchindex(int32_t *ch, int n, uint8_t *link, int n_link, const int32_t *coarse, const uint32_t *cq, int32_t *n_link_out) {
  for (int c = 0; c < n; c++) ch[c] = coarse[ch[c]];
  int nl = n_link;
  if (nl > 0) {
    uint8_t *rm = (uint8_t *)calloc((size_t)nl + 1, 1);
    for (int c = 1; c < n; c++) if (ch[c] == ch[c - 1]) rm[c - 1] = 1;
    int sm = 0;
    for (int c = 0; c < nl; c++) if (rm[c] == 0) link[sm++] = link[c];
    nl = sm;
    free(rm);
  }
  { int m = 0;                                     /* std::unique */
    for (int c = 0; c < n; c++) if (c == 0 || ch[c] != ch[m - 1]) ch[m++] = ch[c];
    n = m; }
  if (n > 0) {
    /* clusters that appear more than once: cut out everything between the first and the last appearance */
    int ns = 0;                                    /* (start, end) of every value with end > start + 1 */
    int *ss = (int *)malloc((size_t)n * sizeof(int)), *se = (int *)malloc((size_t)n * sizeof(int));
    for (int c = 0; c < n; c++) {
      int f = -1;
      for (int d = 0; d < c; d++) if (ch[d] == ch[c]) { f = d; break; }
      if (f >= 0) continue;                        /* not the first appearance */
      int e = c + 1;
      for (int d = c + 1; d < n; d++) if (ch[d] == ch[c]) e = d + 1;
      if (e > c + 1) { ss[ns] = c; se[ns] = e; ns++; }
    }
    /* sort(start_end): tuples (start, end); starts are distinct */
    for (int i = 1; i < ns; i++) { int a = ss[i], b = se[i], j = i; while (j > 0 && ss[j - 1] > a) { ss[j] = ss[j - 1]; se[j] = se[j - 1]; j--; } ss[j] = a; se[j] = b; }
    int32_t *newch = (int32_t *)malloc((size_t)n * sizeof(int32_t)); uint8_t *newlink = (uint8_t *)malloc((size_t)n + 1);
    int nn = 0, nnl = 0, ste = 0, nc = 0;
    while (ste < ns) {
      while (nc <= ss[ste]) {
        newch[nn++] = ch[nc];
        if (nn > 1) newlink[nnl++] = (nc - 1 >= 0 && nc - 1 < nl) ? link[nc - 1] : 0;
        nc++;
      }
      nc = se[ste];
      ste++;
    }
    while (nc < n) {
      newch[nn++] = ch[nc];
      if (nn > 1) newlink[nnl++] = (nc - 1 >= 0 && nc - 1 < nl) ? link[nc - 1] : 0;
      nc++;
    }
    memcpy(ch, newch, (size_t)nn * sizeof(int32_t)); memcpy(link, newlink, (size_t)nnl);
    n = nn; nl = nnl;
    free(ss); free(se); free(newch); free(newlink);
  }
  { /* clusters covered on the read by their predecessor */
    uint8_t *cr = (uint8_t *)calloc((size_t)n + 1, 1);
    for (int c = 1; c < n; c++) {
      const int r = ch[c], p = ch[c - 1];
      if (cr[c - 1] == 0 && cq[2 * r] >= cq[2 * p] && cq[2 * r + 1] <= cq[2 * p + 1]) cr[c] = 1;
    }
    int sc = 0;
    for (int c = 0; c < n; c++) if (cr[c] == 0) { ch[sc] = ch[c]; if (sc >= 1) link[sc - 1] = link[c - 1]; sc++; }
    n = sc; nl = sc - 1 > 0 ? sc - 1 : 0;
    free(cr);
  }
  *n_link_out = nl;
  return n;
}
