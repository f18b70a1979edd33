// a9: SplitClusters + DecideSplitClustersValue (reference SplitClusters.h:17-248), batched over reads.
// Every cluster of a read is cut where any cluster of the read starts or ends, in q or in t: the q and t coordinates inside a cluster's box
// are merged into one list, ordered by projecting them through the line of the box (binary64), and walked once.  That comparator is not a
// strict weak order in general, so the order is whatever libstdc++'s std::sort produces with its answers: the sort
// is replayed (introsort.cuh) with the same comparator, evaluated with explicitly rounded, never fused binary64 operations.  (GenomePos) of a
// double follows the x86-64 conversion (cvttsd2si to 64 bits, low 32 bits kept).
// One read per thread, two passes (count the pieces -> scan -> emit and value them); the reads of a batch are the parallelism.
#pragma once
#include "lra_common.cuh"
#include "introsort.cuh"

namespace lra {

struct ScPoint { uint32_t first; uint32_t second; };      // second: 0 = a q coordinate, 1 = a t coordinate

struct ScLess {
  double slope, intercept; int strand;
  __device__ __forceinline__ bool operator()(const ScPoint &a, const ScPoint &b) const {
    if (a.second == b.second && a.second == 0) return a.first < b.first;
    if (a.second == b.second) return strand == 0 ? a.first < b.first : a.first > b.first;
    if (a.second == 0) { const double p = __dadd_rn(__dmul_rn((double)a.first, slope), intercept); return strand == 0 ? p < (double)b.first : p > (double)b.first; }
    const double p = __dadd_rn(__dmul_rn((double)b.first, slope), intercept);
    return strand == 0 ? (double)a.first < p : (double)a.first > p;
  }
};

__device__ __forceinline__ uint32_t sc_to_gp(double x) {     // (GenomePos) x as x86-64 compiles it
  long long v;
  if (!(x > -9.2233720368547758e18 && x < 9.2233720368547758e18)) v = (long long)0x8000000000000000ull;
  else v = (long long)x;
  return (uint32_t)(unsigned long long)v;
}

struct SplitBatch {
  int n_reads, contig, globalK;
  const unsigned long long *cl_off;     // [n_reads + 1] clusters of every read
  const uint32_t *box;                  // [clusters][4] qStart, qEnd, tStart, tEnd
  const uint8_t *strand;
  const float *freq;                    // anchorfreq
  const unsigned long long *m_off;      // [clusters + 1] anchors of every cluster (CartesianSort order)
  const uint32_t *mq;                   // read positions of the anchors
  uint8_t *split;                       // [clusters]
  int32_t *val_cluster;                 // [clusters]
  unsigned long long *sp_off;           // [n_reads + 1] pieces per read, then offsets
  uint32_t *sp;                         // [pieces][6] qStart, qEnd, tStart, tEnd, strand, coarse (index of the cluster within its read)
  int32_t *sp_val, *sp_n0;
  unsigned long long sp_cap;
  uint32_t *sets;                       // scratch [4 * clusters]: qSet, tSet of the read
  ScPoint *pts;                         // scratch [4 * clusters]
};

template <bool EMIT>
__device__ __noinline__ void split_one(const SplitBatch &b, const int r) {
  const unsigned long long c0 = b.cl_off[r];
  const int n = (int)(b.cl_off[r + 1] - c0);
  const uint32_t *box = b.box + 4 * c0;
  const uint8_t *strand = b.strand + c0;
  const float *freq = b.freq + c0;
  uint8_t *split = b.split + c0;
  uint32_t *qSet = b.sets + 4 * c0, *tSet = qSet + 2 * n;
  ScPoint *S = b.pts + 4 * c0;
  const unsigned long long obase = EMIT ? b.sp_off[r] : 0ull;
  unsigned long long ns = 0;
  auto push = [&](uint32_t qs, uint32_t qe, uint32_t ts, uint32_t te, int st, int co) {
    if (EMIT) {
      const unsigned long long o = obase + ns;
      if (o < b.sp_cap) { uint32_t *p = b.sp + 6 * o; p[0] = qs; p[1] = qe; p[2] = ts; p[3] = te; p[4] = (uint32_t)st; p[5] = (uint32_t)co; }
    }
    ns++;
  };
  int nq = 0, nt = 0;
  for (int m = 0; m < n; m++) {
    const uint32_t qS = box[4 * m], qE = box[4 * m + 1], tS = box[4 * m + 2], tE = box[4 * m + 3];
    const uint32_t big = (tE - tS) > (qE - qS) ? (tE - tS) : (qE - qS);
    bool sp;
    if (b.contig && (freq[m] <= 3.0f || (freq[m] <= 5.0f && big <= 2000u))) sp = true;
    else if (b.contig) { sp = false; push(qS, qE, tS, tE, strand[m], m); }
    else sp = true;
    split[m] = sp ? 1 : 0;
    if (sp) { qSet[nq++] = qS; qSet[nq++] = qE; tSet[nt++] = tS; tSet[nt++] = tE; }
  }
  // std::set: sorted, unique (insertion sort: a read has tens of clusters)
  auto sort_unique = [](uint32_t *v, int k) {
    for (int i = 1; i < k; i++) { const uint32_t x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; j--; } v[j + 1] = x; }
    int u = 0;
    for (int i = 0; i < k; i++) if (i == 0 || v[i] != v[i - 1]) v[u++] = v[i];
    return u;
  };
  nq = sort_unique(qSet, nq); nt = sort_unique(tSet, nt);
  for (int m = 0; m < n; m++) {
    if (!split[m]) continue;
    const uint32_t qS = box[4 * m], qE = box[4 * m + 1], tS = box[4 * m + 2], tE = box[4 * m + 3];
    const int st = strand[m];
    ScLess L;
    L.strand = st;
    L.slope = __ddiv_rn((double)((long long)tE - (long long)tS), (double)((long long)qE - (long long)qS));
    if (st == 0) L.intercept = __ddiv_rn((double)((long long)qE * (long long)tS - (long long)qS * (long long)tE), (double)((long long)qE - (long long)qS));
    else { L.slope = __dmul_rn(-1.0, L.slope); L.intercept = __ddiv_rn((double)((long long)qS * (long long)tS - (long long)qE * (long long)tE), (double)((long long)qS - (long long)qE)); }
    int k = 0;
    for (int i = 0; i < nq; i++) if (qSet[i] > qS && qSet[i] < qE) S[k++] = ScPoint{qSet[i], 0u};
    for (int i = 0; i < nt; i++) if (tSet[i] > tS && tSet[i] < tE) S[k++] = ScPoint{tSet[i], 1u};
    std_sort_replay(S, k, L);
    uint32_t pf = qS, ps = st == 0 ? tS : tE;
    for (int i = 0; i < k; i++) {
      const uint32_t c = S[i].first;
      if (S[i].second == 0) {
        const uint32_t t = sc_to_gp(ceil(__dadd_rn(__dmul_rn(L.slope, (double)c), L.intercept)));
        if (pf < c) {
          if (st == 0 && c >= pf + 3 && t >= ps + 3) push(pf, c, ps, t, st, m);
          else if (st == 1 && c >= pf + 3 && ps >= t + 3) push(pf, c, t, ps, st, m);
        } else continue;
        pf = c; ps = t;
      } else {
        const uint32_t q = sc_to_gp(ceil(__ddiv_rn(__dsub_rn((double)c, L.intercept), L.slope)));
        if (pf < q) {
          if (st == 0 && q >= pf + 3 && c >= ps + 3) push(pf, q, ps, c, st, m);
          else if (st == 1 && q >= pf + 3 && ps >= c + 3) push(pf, q, c, ps, st, m);
        } else continue;
        pf = q; ps = c;
      }
    }
    if (pf < qE) {
      if (st == 0 && qE >= pf + 3 && tE >= ps + 3) push(pf, qE, ps, tE, st, m);
      else if (st == 1 && qE >= pf + 3 && ps >= tS + 3) push(pf, qE, tS, ps, st, m);
    }
  }
  if (!EMIT) { b.sp_off[r] = ns; return; }
  // DecideSplitClustersValue
  int32_t *vc = b.val_cluster + c0;
  for (int m = 0; m < n; m++) vc[m] = 0;
  if (ns == 0 || obase + ns > b.sp_cap) return;
  const unsigned long long *moff = b.m_off + c0;
  const uint32_t K = (uint32_t)b.globalK;
  for (int m = 0; m < n; m++) {
    const unsigned long long a = moff[m], e = moff[m + 1];
    if (e == a) continue;
    uint32_t cur_len = b.mq[a], MatNum = 0;
    for (unsigned long long i = a; i < e; i++) {
      const uint32_t p = b.mq[i];
      MatNum += cur_len > p ? p + K - cur_len : K;
      cur_len = p + K;
    }
    vc[m] = (int32_t)MatNum;
  }
  const uint32_t *sp = b.sp + 6 * obase;
  int32_t *sv = b.sp_val + obase, *s0 = b.sp_n0 + obase;
  for (unsigned long long k = 0; k < ns; k++) {
    const uint32_t *p = sp + 6 * k;
    const int ic = (int)p[5];
    const uint32_t ua = (p[1] - p[0]) < (p[3] - p[2]) ? (p[1] - p[0]) : (p[3] - p[2]);
    const uint32_t ub = (box[4 * ic + 1] - box[4 * ic]) < (box[4 * ic + 3] - box[4 * ic + 2]) ? (box[4 * ic + 1] - box[4 * ic]) : (box[4 * ic + 3] - box[4 * ic + 2]);
    const float pika = __fdiv_rn((float)ua, (float)ub);
    sv[k] = (int32_t)__fmul_rn((float)vc[ic], pika);
    s0[k] = 0;
  }
  unsigned long long m = 0, nn = 1;
  long long matchS = 0, matchE = 0;
  int ic_m = (int)sp[5], ic_n = ns > 1 ? (int)sp[6 + 5] : 0;
  while (nn < ns) {
    if (ic_m == ic_n) {
      const unsigned long long a = moff[ic_n];
      long long lo = 0, len = (long long)(moff[ic_n + 1] - a);
      const uint32_t query = sp[6 * nn];
      while (len > 0) { const long long half = len >> 1; if (b.mq[a + lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
      matchE = lo;
      s0[m] = (int32_t)(matchE - matchS);
      matchS = matchE;
    } else {
      matchE = (long long)(moff[ic_m + 1] - moff[ic_m]);
      s0[m] = (int32_t)(matchE - matchS);
      matchS = 0;
    }
    m = nn; ic_m = ic_n; nn++;
    if (nn < ns) ic_n = (int)sp[6 * nn + 5];
  }
  s0[nn - 1] = (int32_t)((long long)(moff[ic_m + 1] - moff[ic_m]) - matchS);
}

template <bool EMIT>
__global__ void __launch_bounds__(64) split_kernel(SplitBatch b) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= b.n_reads) return;
  split_one<EMIT>(b, r);
}

}  // namespace lra
