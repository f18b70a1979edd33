/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the diagonal cleaning of the anchors of one read strand:
 *   CleanOffDiagonal             /root/reference/Clustering.h:565-800   (called with diagOrigin = diagDrift = -1, Clustering.h:1567-1862)
 *   SecondRoundCleanOffDiagonal  Clustering.h:801-868
 *   AVGfreq                      Clustering.h:549-563   (anchors of the run / distinct read tuples among them, binary32)
 *   DiagonalDifference           Clustering.h:501-514
 * The anchors come sorted by DiagonalSort (strand 0) / AntiDiagonalSort (strand 1).  The repeat-aware thresholds (MinDiagCluster) mix int,
 * float and double arithmetic (std::floor of a float is a float, of an int a double) and are truncated to int; that arithmetic is kept.
 * Outputs: keep[i] (Second_onDiag), freq[i] (matches_freq, valid where kept), cnt[i] (the run counter, valid where kept) per INPUT anchor, and
 * with opts.ExtractDiagonalFromClean the clusters over the COMPACTED anchors: cl[7k..] = start, end, qStart, qEnd, tStart, tEnd, chromIndex
 * (chromIndex only with bypassClustering, else 0), cl_freq[k] = anchorfreq.  Returns the number of clusters.
 * Pinned by tests/test_clean_off_diagonal.py against the unmodified reference (oracle/ref_wrap.cpp: ref_clean_off_diagonal). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { int cleanMaxDiag, minDiagCluster, bypassClustering, cleanClustersize, SecondCleanMinDiagCluster, punish_anchorfreq, anchorPerlength,
                 SecondCleanMaxDiag, ExtractDiagonalFromClean, globalK; } cod_opts;

static long diag_diff(const uint32_t *q, const uint32_t *t, long a, long b, int strand) {
  if (strand == 0) return ((long)t[a] - (long)q[a]) - ((long)t[b] - (long)q[b]);
  return (long)(uint32_t)(q[a] + t[a]) - (long)(uint32_t)(q[b] + t[b]);
}
static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : (x > y ? 1 : 0); }
static float avg_freq(const uint64_t *qt, long as, long ae) {
  uint64_t *v = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(ae - as));
  for (long i = as; i < ae; i++) v[i - as] = qt[i];
  qsort(v, (size_t)(ae - as), sizeof(uint64_t), cmp_u64);
  long distinct = 1;
  for (long i = 1; i < ae - as; i++) if (v[i] != v[i - 1]) distinct++;
  free(v);
  return (float)(ae - as) / distinct;
}
static void second_round(int32_t *count, int out_counter, const uint32_t *q, const uint32_t *t, int MinDiagCluster, int CleanMaxDiag, uint8_t *orig, long os, long oe, int strand) {
  if (MinDiagCluster >= oe - os) return;
  if (MinDiagCluster <= 0) { for (long i = os; i < oe; i++) { orig[i] = 1; count[i] = out_counter; } return; }
  if (oe - os <= 1) return;
  uint8_t *fw = (uint8_t *)calloc((size_t)(oe - os), 1), *rv = (uint8_t *)calloc((size_t)(oe - os), 1);
  for (long i = os + 1; i < oe; i++) if (labs(diag_diff(q, t, i, i - 1, strand)) < CleanMaxDiag) fw[i - 1 - os] = 1;
  int prev = 0; long diagStart = 0;
  for (long i = os; i < oe; i++) {
    if (!prev && fw[i - os]) diagStart = i;
    if (prev && !fw[i - os]) {
      if (i - diagStart + 1 < MinDiagCluster) for (long j = diagStart; j <= i; j++) fw[j - os] = 0;
      else fw[i - os] = 1;
    }
    prev = fw[i - os];
  }
  for (long i = oe - 2; i >= os; i--) if (labs(diag_diff(q, t, i, i + 1, strand)) < CleanMaxDiag) rv[i + 1 - os] = 1;
  prev = 0;
  for (long i = oe - 1; i >= os; i--) {
    if (!prev && rv[i - os]) diagStart = i;
    if (prev && !rv[i - os]) {
      if (diagStart - i + 1 < MinDiagCluster) for (long j = i; j <= diagStart; j++) rv[j - os] = 0;
      else rv[i - os] = 1;
    }
    prev = rv[i - os];
  }
  for (long i = os; i < oe; i++) { if (fw[i - os] && rv[i - os]) { orig[i] = 1; count[i] = out_counter; } else orig[i] = 0; }
  free(fw); free(rv);
}
static int hdr_find2(const uint64_t *pos, int n, uint64_t query) {
  if (n > 0 && query == pos[0]) return 0;
  int lo = 0, len = n;
  while (len > 0) { int half = len >> 1; if (pos[lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
  if (lo < n && query == pos[lo]) return lo;
  return lo - 1;
}

long lra_oracle_clean_off_diagonal(const uint32_t *q, const uint32_t *t, const uint64_t *qt, long n, int strand, const cod_opts *o, const uint64_t *hdr_pos, int n_hdr,
                                   uint8_t *keep, float *freq, int32_t *cnt, int32_t *cl, float *cl_freq) {
  for (long i = 0; i < n; i++) { keep[i] = 0; freq[i] = 0; cnt[i] = -1; }
  if (n == 0) return 0;
  uint8_t *onDiag = (uint8_t *)calloc((size_t)n, 1);
  if (n > 1 && labs(diag_diff(q, t, 0, 1, strand)) < o->cleanMaxDiag) onDiag[0] = 1;
  for (long i = 1; i < n; i++) if (labs(diag_diff(q, t, i, i - 1, strand)) < o->cleanMaxDiag) onDiag[i - 1] = 1;
  int prev = 0, set = 0, Largest = 0; long diagStart = 0;
  for (long i = 0; i < n; i++) {
    if (!prev && onDiag[i]) { diagStart = i; set = 1; }
    if (prev && !onDiag[i]) { if (i - diagStart + 1 > Largest) Largest = (int)(i - diagStart + 1); }
    prev = onDiag[i];
  }
  if (!set) { free(onDiag); return 0; }
  if (n - diagStart > Largest) Largest = (int)(n - diagStart);
  int minDiagCluster = (int)floor((double)(Largest / 10));
  if (minDiagCluster >= o->minDiagCluster) minDiagCluster = o->minDiagCluster;
  int counter = 0;
  prev = 0;
  const int ccs = o->cleanClustersize, S = o->SecondCleanMinDiagCluster, pa = o->punish_anchorfreq, apl = o->anchorPerlength;
  for (long i = 0; i < n; i++) {
    if (!prev && onDiag[i]) diagStart = i;
    if (prev && !onDiag[i]) {
      const int size = (int)(i - diagStart + 1);
      if (size >= minDiagCluster) {
        const float avgfreq = avg_freq(qt, diagStart, i + 1);
        for (long j = diagStart; j <= i; j++) freq[j] = avgfreq;
        int M = 0, second = 0, all = 0;
        if (o->bypassClustering) {
          if (avgfreq >= 3.0f && size < 10) { }
          else if (avgfreq >= 2.0f && size >= ccs) { M = (int)(S + floorf((avgfreq - 1.5f) / 1.0f) * pa + floor((double)((size - ccs) / ccs)) * apl); second = 1; }
          else if (avgfreq >= 1.5f && size >= ccs) { M = (int)(S + floorf((avgfreq - 1.5f) / 1.5f) * pa + floor((double)((size - ccs) / ccs)) * apl); second = 1; }
          else all = 1;
        } else {
          if (avgfreq >= 3.0f && size < 10) { }
          else if (avgfreq >= 4.0f && size >= ccs) { M = (int)(S + floorf((avgfreq - 1.5f) / 1.0f) * pa + floor((double)((size - ccs) / ccs)) * apl); second = 1; }
          else if (avgfreq >= 1.5f && size >= ccs) { M = (int)(S + floorf((avgfreq - 1.5f) / 1.5f) * pa + floor((double)((size - ccs) / ccs)) * apl); second = 1; }
          else if (avgfreq > 1.0f && size >= ccs) { M = (int)(S - (5 - floorf((avgfreq - 1.0f) / 0.1f)) * (pa / 2) + floor((double)((size - ccs) / ccs)) * (apl / 2)); second = 1; }
          else if (avgfreq > 1.0f) { M = (int)(S - (5 - floorf((avgfreq - 1.0f) / 0.1f)) * (pa / 2) - floor((double)((ccs - (int)i + (int)diagStart - 1) / 15)) * (apl / 2)); second = 1; }
          else all = 1;
        }
        if (second) second_round(cnt, counter, q, t, M, o->SecondCleanMaxDiag, keep, diagStart, i + 1, strand);
        if (all) for (long j = diagStart; j <= i; j++) { keep[j] = 1; cnt[j] = counter; }
      }
      counter++;
    }
    prev = onDiag[i];
  }
  free(onDiag);
  long ncl = 0;
  if (o->ExtractDiagonalFromClean) {
    /* the compacted arrays */
    long m = 0;
    for (long i = 0; i < n; i++) if (keep[i]) m++;
    uint32_t *cq = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(m + 1)), *ct = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(m + 1));
    int32_t *cc = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m + 1)); float *cf = (float *)malloc(sizeof(float) * (size_t)(m + 1));
    m = 0;
    for (long i = 0; i < n; i++) if (keep[i]) { cq[m] = q[i]; ct[m] = t[i]; cc[m] = cnt[i]; cf[m] = freq[i]; m++; }
    long count_s = 0, c = 1;
    const uint32_t K = (uint32_t)o->globalK;
#define EMIT_CLUSTER(cs, ce) do { \
      uint32_t qS = cq[cs], qE = cq[cs] + K, tS = ct[cs], tE = ct[cs] + K; \
      for (long b = (cs); b < (ce); b++) { if (cq[b] < qS) qS = cq[b]; if (cq[b] + K > qE) qE = cq[b] + K; if (ct[b] < tS) tS = ct[b]; if (ct[b] + K > tE) tE = ct[b] + K; } \
      cl[7 * ncl] = (int32_t)(cs); cl[7 * ncl + 1] = (int32_t)(ce); cl[7 * ncl + 2] = (int32_t)qS; cl[7 * ncl + 3] = (int32_t)qE; cl[7 * ncl + 4] = (int32_t)tS; cl[7 * ncl + 5] = (int32_t)tE; \
      cl[7 * ncl + 6] = o->bypassClustering ? hdr_find2(hdr_pos, n_hdr, (uint64_t)tS) : 0; cl_freq[ncl] = cf[cs]; ncl++; } while (0)
    while (c < m) {
      if (cc[c] == cc[c - 1]) { c++; continue; }
      EMIT_CLUSTER(count_s, c);
      count_s = c; c++;
    }
    if (c == m && count_s < c) EMIT_CLUSTER(count_s, c);
    free(cq); free(ct); free(cc); free(cf);
  }
  return ncl;
}
