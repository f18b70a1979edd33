/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the cluster splitting of the high-accuracy pipeline:
 *   IntervalSet (the line through a cluster's box, and the comparator that projects q / t coordinates through it)   SplitClusters.h:17-61
 *   SplitClusters                  SplitClusters.h:63-171   every cluster is cut where any cluster of the read starts or ends, in q or in t
 *   DecideSplitClustersValue       SplitClusters.h:174-248  Val (covered bases x length ratio, binary32) and NumofAnchors0 of every piece
 * The mixed q/t coordinate list is sorted with std::sort under a comparator that is NOT a strict weak order in general (a q coordinate is
 * compared to a t coordinate through a double line): the result is whatever libstdc++'s introsort does with these
 * answers, so that algorithm is restated (GCC 13.3 bits/stl_algo.h) and driven by the same comparator.
 * (GenomePos) ceil(x) is the x86-64 conversion: truncate to int64, keep the low 32 bits.
 * Pinned by tests/test_split_clusters.py against the unmodified reference (oracle/ref_wrap.cpp: ref_split_clusters). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t first; uint8_t second; } sc_pt;
typedef struct { double slope, intercept; int strand; } sc_line;

static int sc_less(const sc_line *L, sc_pt a, sc_pt b) {
  if (a.second == b.second && a.second == 0) return a.first < b.first;
  else if (a.second == b.second && a.second == 1) return L->strand == 0 ? a.first < b.first : a.first > b.first;
  else if (a.second == 0 && b.second == 1) return L->strand == 0 ? (a.first * L->slope + L->intercept < (double)b.first) : (a.first * L->slope + L->intercept > (double)b.first);
  else return L->strand == 0 ? ((double)a.first < b.first * L->slope + L->intercept) : ((double)a.first > b.first * L->slope + L->intercept);
}
#define LESS(a, b) sc_less(L, (a), (b))
static void sc_unguarded_linear_insert(const sc_line *L, sc_pt *last) {
  sc_pt val = *last; sc_pt *next = last - 1;
  while (LESS(val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void sc_insertion_sort(const sc_line *L, sc_pt *first, sc_pt *last) {
  if (first == last) return;
  for (sc_pt *i = first + 1; i != last; ++i) {
    if (LESS(*i, *first)) { sc_pt val = *i; memmove(first + 1, first, (size_t)(i - first) * sizeof(sc_pt)); *first = val; }
    else sc_unguarded_linear_insert(L, i);
  }
}
static void sc_adjust_heap(const sc_line *L, sc_pt *first, long holeIndex, long len, sc_pt value) {
  const long topIndex = holeIndex;
  long secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (LESS(first[secondChild], first[secondChild - 1])) secondChild--;
    first[holeIndex] = first[secondChild]; holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) { secondChild = 2 * (secondChild + 1); first[holeIndex] = first[secondChild - 1]; holeIndex = secondChild - 1; }
  long parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && LESS(first[parent], value)) { first[holeIndex] = first[parent]; holeIndex = parent; parent = (holeIndex - 1) / 2; }
  first[holeIndex] = value;
}
static void sc_heap_sort(const sc_line *L, sc_pt *first, sc_pt *last) {
  long len = last - first;
  if (len >= 2) for (long parent = (len - 2) / 2;; parent--) { sc_pt v = first[parent]; sc_adjust_heap(L, first, parent, len, v); if (parent == 0) break; }
  while (last - first > 1) { --last; sc_pt v = *last; *last = *first; sc_adjust_heap(L, first, 0, last - first, v); }
}
static void sc_introsort_loop(const sc_line *L, sc_pt *first, sc_pt *last, long depth_limit) {
  while (last - first > 16) {
    if (depth_limit == 0) { sc_heap_sort(L, first, last); return; }
    --depth_limit;
    sc_pt *mid = first + (last - first) / 2, *a = first + 1, *b = mid, *c = last - 1, t;
#define SWP(x, y) do { t = *(x); *(x) = *(y); *(y) = t; } while (0)
    if (LESS(*a, *b)) { if (LESS(*b, *c)) SWP(first, b); else if (LESS(*a, *c)) SWP(first, c); else SWP(first, a); }
    else if (LESS(*a, *c)) SWP(first, a);
    else if (LESS(*b, *c)) SWP(first, c);
    else SWP(first, b);
    sc_pt *lo = first + 1, *hi = last;
    for (;;) {
      while (LESS(*lo, *first)) ++lo;
      --hi;
      while (LESS(*first, *hi)) --hi;
      if (!(lo < hi)) break;
      SWP(lo, hi);
      ++lo;
    }
    sc_introsort_loop(L, lo, last, depth_limit);
    last = lo;
  }
}
static void sc_sort(const sc_line *L, sc_pt *v, long n) {
  if (n <= 1) return;
  long lg = 0; { unsigned long x = (unsigned long)n; while (x > 1) { x >>= 1; lg++; } }
  sc_introsort_loop(L, v, v + n, lg * 2);
  if (n > 16) { sc_insertion_sort(L, v, v + 16); for (sc_pt *i = v + 16; i != v + n; ++i) sc_unguarded_linear_insert(L, i); }
  else sc_insertion_sort(L, v, v + n);
}
static int cmp_u32(const void *a, const void *b) { uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return x < y ? -1 : (x > y ? 1 : 0); }
static uint32_t to_gp(double x) { return (uint32_t)(uint64_t)(int64_t)x; }       /* (GenomePos) of a double as x86-64 compiles it */

/* box[4m..] = qStart, qEnd, tStart, tEnd; strand[m]; freq[m] = anchorfreq; contig = (opts.readType == Options::contig).
 * Anchors of cluster m (CartesianSort order): mq[m_off[m] .. m_off[m+1]) (read positions), for Val / NumofAnchors0.
 * Out: split[m], val_cluster[m]; pieces sp[6k..] = qStart, qEnd, tStart, tEnd, strand, coarse; sp_val[k], sp_n0[k].  Returns the number of pieces
 * (all counted; the first `cap` stored). */
long lra_oracle_split_clusters(const uint32_t *box, const uint8_t *strand, const float *freq, long n, int contig, const uint32_t *mq, const uint64_t *m_off, int globalK,
                               uint8_t *split, int32_t *val_cluster, uint32_t *sp, int32_t *sp_val, int32_t *sp_n0, long cap) {
  long ns = 0;
#define PUSH(qs, qe, ts, te, st, co) do { if (ns < cap) { sp[6 * ns] = (qs); sp[6 * ns + 1] = (qe); sp[6 * ns + 2] = (ts); sp[6 * ns + 3] = (te); sp[6 * ns + 4] = (uint32_t)(st); sp[6 * ns + 5] = (uint32_t)(co); } ns++; } while (0)
  uint32_t *qSet = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(2 * n + 1)), *tSet = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(2 * n + 1));
  long nq = 0, nt = 0;
  for (long m = 0; m < n; m++) {
    const uint32_t qS = box[4 * m], qE = box[4 * m + 1], tS = box[4 * m + 2], tE = box[4 * m + 3];
    const uint32_t big = (tE - tS) > (qE - qS) ? (tE - tS) : (qE - qS);
    if (contig && (freq[m] <= 3.0f || (freq[m] <= 5.0f && big <= 2000))) { split[m] = 1; qSet[nq++] = qS; qSet[nq++] = qE; tSet[nt++] = tS; tSet[nt++] = tE; }
    else if (contig) { split[m] = 0; PUSH(qS, qE, tS, tE, strand[m], m); }
    else { split[m] = 1; qSet[nq++] = qS; qSet[nq++] = qE; tSet[nt++] = tS; tSet[nt++] = tE; }
  }
  qsort(qSet, (size_t)nq, 4, cmp_u32); qsort(tSet, (size_t)nt, 4, cmp_u32);
  { long k = 0; for (long i = 0; i < nq; i++) if (i == 0 || qSet[i] != qSet[i - 1]) qSet[k++] = qSet[i]; nq = k; }
  { long k = 0; for (long i = 0; i < nt; i++) if (i == 0 || tSet[i] != tSet[i - 1]) tSet[k++] = tSet[i]; nt = k; }
  sc_pt *S = (sc_pt *)malloc(sizeof(sc_pt) * (size_t)(nq + nt + 1));
  for (long m = 0; m < n; m++) {
    if (!split[m]) continue;
    const uint32_t qS = box[4 * m], qE = box[4 * m + 1], tS = box[4 * m + 2], tE = box[4 * m + 3];
    const int st = strand[m];
    sc_line Ln;
    Ln.slope = (double)((int64_t)tE - (int64_t)tS) / (double)((int64_t)qE - (int64_t)qS);
    if (st == 0) Ln.intercept = ((double)((int64_t)qE * tS - (int64_t)qS * tE)) / (double)((int64_t)qE - (int64_t)qS);
    else { Ln.slope = -1 * Ln.slope; Ln.intercept = (double)((int64_t)qS * tS - (int64_t)qE * tE) / (double)((int64_t)qS - (int64_t)qE); }
    Ln.strand = st;
    long k = 0;
    for (long i = 0; i < nq; i++) if (qSet[i] > qS && qSet[i] < qE) { S[k].first = qSet[i]; S[k].second = 0; k++; }     /* (upper_bound(qStart), lower_bound(qEnd)) */
    for (long i = 0; i < nt; i++) if (tSet[i] > tS && tSet[i] < tE) { S[k].first = tSet[i]; S[k].second = 1; k++; }
    sc_sort(&Ln, S, k);
    uint32_t pf = qS, ps = st == 0 ? tS : tE;
    for (long i = 0; i < k; i++) {
      if (S[i].second == 0) {
        const uint32_t t = to_gp(ceil(Ln.slope * S[i].first + Ln.intercept));
        if (pf < S[i].first) {
          if (st == 0 && S[i].first >= pf + 3 && t >= ps + 3) PUSH(pf, S[i].first, ps, t, st, m);
          else if (st == 1 && S[i].first >= pf + 3 && ps >= t + 3) PUSH(pf, S[i].first, t, ps, st, m);
        } else continue;
        pf = S[i].first; ps = t;
      } else {
        const uint32_t q = to_gp(ceil((S[i].first - Ln.intercept) / Ln.slope));
        if (pf < q) {
          if (st == 0 && q >= pf + 3 && S[i].first >= ps + 3) PUSH(pf, q, ps, S[i].first, st, m);
          else if (st == 1 && q >= pf + 3 && ps >= S[i].first + 3) PUSH(pf, q, S[i].first, ps, st, m);
        } else continue;
        pf = q; ps = S[i].first;
      }
    }
    if (pf < qE) {
      if (st == 0 && qE >= pf + 3 && tE >= ps + 3) PUSH(pf, qE, ps, tE, st, m);
      else if (st == 1 && qE >= pf + 3 && ps >= tS + 3) PUSH(pf, qE, tS, ps, st, m);
    }
  }
  free(qSet); free(tSet); free(S);
  /* DecideSplitClustersValue */
  for (long m = 0; m < n; m++) val_cluster[m] = 0;
  if (ns == 0 || ns > cap) return ns;
  for (long m = 0; m < n; m++) {
    const long a = (long)m_off[m], b = (long)m_off[m + 1];
    if (b == a) continue;
    uint32_t cur_len = mq[a], MatNum = 0;
    for (long i = a; i < b; i++) {
      if (cur_len > mq[i]) MatNum += mq[i] + (uint32_t)globalK - cur_len; else MatNum += (uint32_t)globalK;
      cur_len = mq[i] + (uint32_t)globalK;
    }
    val_cluster[m] = (int32_t)MatNum;
  }
  for (long k = 0; k < ns; k++) {
    const long ic = sp[6 * k + 5];
    const uint32_t a = (sp[6 * k + 1] - sp[6 * k]) < (sp[6 * k + 3] - sp[6 * k + 2]) ? (sp[6 * k + 1] - sp[6 * k]) : (sp[6 * k + 3] - sp[6 * k + 2]);
    const uint32_t b = (box[4 * ic + 1] - box[4 * ic]) < (box[4 * ic + 3] - box[4 * ic + 2]) ? (box[4 * ic + 1] - box[4 * ic]) : (box[4 * ic + 3] - box[4 * ic + 2]);
    const float pika = (float)a / (float)b;
    sp_val[k] = (int32_t)((float)(int)val_cluster[ic] * pika);
    sp_n0[k] = 0;
  }
  long m = 0, nn = 1, matchS = 0, matchE = 0;
  long ic_m = sp[5], ic_n = ns > 1 ? (long)sp[6 + 5] : 0;
  while (nn < ns) {
    if (ic_m == ic_n) {
      const long a = (long)m_off[ic_n], b = (long)m_off[ic_n + 1];
      long lo = 0, len = b - a;          /* CartesianLowerBound(q = qStart of piece nn, t = 0): first anchor with q >= query */
      while (len > 0) { long half = len >> 1; if (mq[a + lo + half] < sp[6 * nn]) { lo += half + 1; len -= half + 1; } else len = half; }
      matchE = lo;
      sp_n0[m] = (int32_t)(matchE - matchS);
      matchS = matchE;
    } else {
      matchE = (long)(m_off[ic_m + 1] - m_off[ic_m]);
      sp_n0[m] = (int32_t)(matchE - matchS);
      matchS = 0;
    }
    m = nn; ic_m = ic_n; nn++;
    if (nn < ns) ic_n = sp[6 * nn + 5];
  }
  sp_n0[nn - 1] = (int32_t)((long)(m_off[ic_m + 1] - m_off[ic_m]) - matchS);
  return ns;
}
