// a7: CleanOffDiagonal + SecondRoundCleanOffDiagonal + AVGfreq (reference Clustering.h:549-868), batched over anchor lists (the matches
// of one read strand, sorted by DiagonalSort / AntiDiagonalSort; the reference always calls it with diagOrigin = diagDrift = -1).
// The function is a handful of sequential scans over the list whose state (run starts, the run counter, flags that later passes re-read)
// carries from anchor to anchor, so one list is one thread; the lists of a batch (two per read) are the parallelism.
//   * on-diagonal flags from consecutive diagonal differences, runs of flagged anchors, the largest run -> minDiagCluster;
//   * per run: AVGfreq = anchors / distinct read tuples (open-addressing table in a per-list scratch slot) and the repeat-aware
//     threshold MinDiagCluster, whose int / float / double mix (std::floor of a float is a float, of an int a double; the sum is truncated
//     to int) is evaluated with explicitly rounded operations in the reference's order;
//   * SecondRoundCleanOffDiagonal: forward and backward re-flagging with the tighter SecondCleanMaxDiag;
//   * opts.ExtractDiagonalFromClean: clusters = maximal stretches of equal run counter over the surviving anchors.
#pragma once
#include "lra_common.cuh"

namespace lra {

struct CodOpts { int cleanMaxDiag, minDiagCluster, bypassClustering, cleanClustersize, SecondCleanMinDiagCluster, punish_anchorfreq, anchorPerlength,
                 SecondCleanMaxDiag, ExtractDiagonalFromClean, globalK; };

struct CodBatch {
  int n_lists;
  const unsigned long long *off;        // [n_lists + 1]
  const uint32_t *q, *t;
  const unsigned long long *qt;         // read tuple of every anchor (GenomePair::first.t)
  const uint8_t *strand;                // [n_lists]
  CodOpts o;
  const unsigned long long *hdr_pos; int n_hdr;
  // per input anchor
  uint8_t *keep; float *freq; int32_t *cnt;
  // clusters of list s at cl[7 * off[s] ..] / cl_freq[off[s] ..]: start, end (in the compacted list), qStart, qEnd, tStart, tEnd, chromIndex
  int32_t *cl; float *cl_freq; int32_t *n_cl;
  // scratch
  uint8_t *flags;                       // [3 * total]: onDiag, forward, reverse
  unsigned long long *hkeys;            // [4 * total]
  uint8_t *hused;                       // [4 * total]
};

__device__ __forceinline__ long long cod_diag_diff(const uint32_t *q, const uint32_t *t, int a, int b, int strand) {
  if (strand == 0) return ((long long)t[a] - (long long)q[a]) - ((long long)t[b] - (long long)q[b]);
  return (long long)(uint32_t)(q[a] + t[a]) - (long long)(uint32_t)(q[b] + t[b]);
}
__device__ __forceinline__ long long cod_abs(long long v) { return v < 0 ? -v : v; }

__device__ __forceinline__ float cod_avgfreq(const unsigned long long *qt, int as, int ae, unsigned long long *hk, uint8_t *hu) {
  const int n = ae - as;
  int cap = 2;
  while (cap < 2 * n) cap <<= 1;
  for (int i = 0; i < cap; i++) hu[i] = 0;
  int distinct = 0;
  for (int i = as; i < ae; i++) {
    const unsigned long long k = qt[i];
    unsigned h = (unsigned)((k * 0x9E3779B97F4A7C15ull) >> 40) & (unsigned)(cap - 1);
    for (;;) {
      if (!hu[h]) { hu[h] = 1; hk[h] = k; distinct++; break; }
      if (hk[h] == k) break;
      h = (h + 1) & (unsigned)(cap - 1);
    }
  }
  return __fdiv_rn((float)n, (float)distinct);
}

__device__ __forceinline__ void cod_second_round(int32_t *count, int out_counter, const uint32_t *q, const uint32_t *t, int M, int CleanMaxDiag, uint8_t *orig, int os, int oe,
                                                 int strand, uint8_t *fw, uint8_t *rv) {
  if (M >= oe - os) return;
  if (M <= 0) { for (int i = os; i < oe; i++) { orig[i] = 1; count[i] = out_counter; } return; }
  if (oe - os <= 1) return;
  for (int i = os; i < oe; i++) { fw[i] = 0; rv[i] = 0; }
  for (int i = os + 1; i < oe; i++) if (cod_abs(cod_diag_diff(q, t, i, i - 1, strand)) < CleanMaxDiag) fw[i - 1] = 1;
  bool prev = false; int diagStart = 0;
  for (int i = os; i < oe; i++) {
    if (!prev && fw[i]) diagStart = i;
    if (prev && !fw[i]) {
      if (i - diagStart + 1 < M) for (int j = diagStart; j <= i; j++) fw[j] = 0;
      else fw[i] = 1;
    }
    prev = fw[i] != 0;
  }
  for (int i = oe - 2; i >= os; i--) if (cod_abs(cod_diag_diff(q, t, i, i + 1, strand)) < CleanMaxDiag) rv[i + 1] = 1;
  prev = false;
  for (int i = oe - 1; i >= os; i--) {
    if (!prev && rv[i]) diagStart = i;
    if (prev && !rv[i]) {
      if (diagStart - i + 1 < M) for (int j = i; j <= diagStart; j++) rv[j] = 0;
      else rv[i] = 1;
    }
    prev = rv[i] != 0;
  }
  for (int i = os; i < oe; i++) { if (fw[i] && rv[i]) { orig[i] = 1; count[i] = out_counter; } else orig[i] = 0; }
}

__device__ __noinline__ void cod_one(const CodBatch &b, const int s) {
  const unsigned long long o = b.off[s];
  const int n = (int)(b.off[s + 1] - o);
  const uint32_t *q = b.q + o, *t = b.t + o;
  const unsigned long long *qt = b.qt + o;
  uint8_t *keep = b.keep + o; float *freq = b.freq + o; int32_t *cnt = b.cnt + o;
  uint8_t *onDiag = b.flags + 3 * o, *fw = onDiag + n, *rv = fw + n;
  unsigned long long *hk = b.hkeys + 4 * o; uint8_t *hu = b.hused + 4 * o;
  const int strand = b.strand[s];
  const CodOpts &O = b.o;
  b.n_cl[s] = 0;
  for (int i = 0; i < n; i++) { keep[i] = 0; freq[i] = 0.0f; cnt[i] = -1; onDiag[i] = 0; }
  if (n == 0) return;
  if (n > 1 && cod_abs(cod_diag_diff(q, t, 0, 1, strand)) < O.cleanMaxDiag) onDiag[0] = 1;
  for (int i = 1; i < n; i++) if (cod_abs(cod_diag_diff(q, t, i, i - 1, strand)) < O.cleanMaxDiag) onDiag[i - 1] = 1;
  bool prev = false, set = false; int Largest = 0, diagStart = 0;
  for (int i = 0; i < n; i++) {
    if (!prev && onDiag[i]) { diagStart = i; set = true; }
    if (prev && !onDiag[i]) Largest = imax(Largest, i - diagStart + 1);
    prev = onDiag[i] != 0;
  }
  if (!set) return;
  Largest = imax(Largest, n - diagStart);
  int minDiagCluster = Largest / 10;
  if (minDiagCluster >= O.minDiagCluster) minDiagCluster = O.minDiagCluster;
  const int ccs = O.cleanClustersize, S = O.SecondCleanMinDiagCluster, pa = O.punish_anchorfreq, apl = O.anchorPerlength;
  int counter = 0;
  prev = false;
  for (int i = 0; i < n; i++) {
    if (!prev && onDiag[i]) diagStart = i;
    if (prev && !onDiag[i]) {
      const int size = i - diagStart + 1;
      if (size >= minDiagCluster) {
        const float avgfreq = cod_avgfreq(qt, diagStart, i + 1, hk, hu);
        for (int j = diagStart; j <= i; j++) freq[j] = avgfreq;
        int M = 0; bool second = false, all = false;
        // S + floor((avgfreq - a) / d) * pa   [float]   + floor((size - ccs) / ccs) * apl   [double]
        auto f1 = [&](float a, float d) { return __fadd_rn((float)S, __fmul_rn(floorf(__fdiv_rn(__fsub_rn(avgfreq, a), d)), (float)pa)); };
        const double grow = __dmul_rn((double)((size - ccs) / ccs), (double)apl);
        // S - (5 - floor((avgfreq - 1) / 0.1f)) * (pa / 2)   [float]
        auto f2 = [&]() { return __fsub_rn((float)S, __fmul_rn(__fsub_rn(5.0f, floorf(__fdiv_rn(__fsub_rn(avgfreq, 1.0f), 0.1f))), (float)(pa / 2))); };
        if (O.bypassClustering) {
          if (avgfreq >= 3.0f && size < 10) { }
          else if (avgfreq >= 2.0f && size >= ccs) { M = (int)__dadd_rn((double)f1(1.5f, 1.0f), grow); second = true; }
          else if (avgfreq >= 1.5f && size >= ccs) { M = (int)__dadd_rn((double)f1(1.5f, 1.5f), grow); second = true; }
          else all = true;
        } else {
          if (avgfreq >= 3.0f && size < 10) { }
          else if (avgfreq >= 4.0f && size >= ccs) { M = (int)__dadd_rn((double)f1(1.5f, 1.0f), grow); second = true; }
          else if (avgfreq >= 1.5f && size >= ccs) { M = (int)__dadd_rn((double)f1(1.5f, 1.5f), grow); second = true; }
          else if (avgfreq > 1.0f && size >= ccs) { M = (int)__dadd_rn((double)f2(), __dmul_rn((double)((size - ccs) / ccs), (double)(apl / 2))); second = true; }
          else if (avgfreq > 1.0f) { M = (int)__dsub_rn((double)f2(), __dmul_rn((double)((ccs - i + diagStart - 1) / 15), (double)(apl / 2))); second = true; }
          else all = true;
        }
        if (second) cod_second_round(cnt, counter, q, t, M, O.SecondCleanMaxDiag, keep, diagStart, i + 1, strand, fw, rv);
        if (all) for (int j = diagStart; j <= i; j++) { keep[j] = 1; cnt[j] = counter; }
      }
      counter++;
    }
    prev = onDiag[i] != 0;
  }
  if (!O.ExtractDiagonalFromClean) return;
  // clusters over the compacted list: maximal stretches of equal run counter
  int32_t *cl = b.cl + 7 * o; float *clf = b.cl_freq + o;
  const uint32_t K = (uint32_t)O.globalK;
  int ncl = 0, m = 0, cs = 0, prevCnt = 0;
  uint32_t qS = 0, qE = 0, tS = 0, tE = 0; float f0 = 0.0f;
  auto emit = [&](int ce) {
    cl[7 * ncl] = cs; cl[7 * ncl + 1] = ce; cl[7 * ncl + 2] = (int32_t)qS; cl[7 * ncl + 3] = (int32_t)qE; cl[7 * ncl + 4] = (int32_t)tS; cl[7 * ncl + 5] = (int32_t)tE;
    int chrom = 0;
    if (O.bypassClustering) {      // Header::Find(tStart)
      const unsigned long long query = tS;
      if (b.n_hdr > 0 && query == b.hdr_pos[0]) chrom = 0;
      else { int lo = 0, len = b.n_hdr; while (len > 0) { const int half = len >> 1; if (b.hdr_pos[lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
             chrom = (lo < b.n_hdr && query == b.hdr_pos[lo]) ? lo : lo - 1; }
    }
    cl[7 * ncl + 6] = chrom; clf[ncl] = f0; ncl++;
  };
  for (int i = 0; i < n; i++) {
    if (!keep[i]) continue;
    if (m > 0 && cnt[i] != prevCnt) { emit(m); cs = m; }
    if (m == cs) { qS = q[i]; qE = q[i] + K; tS = t[i]; tE = t[i] + K; f0 = freq[i]; }
    else { qS = q[i] < qS ? q[i] : qS; qE = q[i] + K > qE ? q[i] + K : qE; tS = t[i] < tS ? t[i] : tS; tE = t[i] + K > tE ? t[i] + K : tE; }
    prevCnt = cnt[i];
    m++;
  }
  if (m > 0) emit(m);
  b.n_cl[s] = ncl;
}

__global__ void __launch_bounds__(64) cod_kernel(CodBatch b) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= b.n_lists) return;
  cod_one(b, s);
}

}  // namespace lra
