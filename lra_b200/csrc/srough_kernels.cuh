// a8 (first half): SplitRoughClustersWithGaps (reference Clustering.h:1358-1430, with CloseToPreviousCluster / MergeTwoClusters / UpdateCluster
// :1332-1356) over all rough clusters of an anchor list, batched over lists (MatchesToFineClusters, Clustering.h:1578-1581 and :1632-1635).
// The split clusters of a list go into one vector whose size and last element (box, strand, chromIndex) decide what happens to the next piece --
// also across rough clusters -- so one thread replays one list literally; the parallelism is the lists of a batch (two per read).
// A split cluster's splitmatchindex is a list of anchor ranges (pieces): the first range and every range MergeTwoClusters appends.
#pragma once
#include "lra_common.cuh"

namespace lra {

struct SplitRoughBatch {
  int n_lists;
  int globalK, maxGap, minClusterSize, maxDiag;
  const unsigned long long *l_off;        // [n_lists + 1] anchors of each list
  const unsigned long long *lr_off;       // [n_lists + 1] rough clusters of each list
  const uint32_t *q, *t;                  // anchors, Cartesian-sorted inside every rough cluster
  const int32_t *r_start, *r_end;         // per rough cluster: anchor range, relative to its list
  const uint32_t *r_box;                  // [n_rough * 4]
  const uint8_t *r_strand;
  const float *r_freq;
  const int32_t *r_chrom;
  // out, slot layout: split cluster i of list l at l_off[l] + lr_off[l] + i, piece j at the same base + j
  int32_t *n_split, *n_piece;             // [n_lists]
  int32_t *s_start, *s_end, *s_coarse, *s_chrom;
  uint32_t *s_box;
  uint8_t *s_strand;
  float *s_freq;
  int32_t *p_cluster, *p_start, *p_end;
};

__device__ __noinline__ void split_rough_one(const SplitRoughBatch &b, const int l) {
  const unsigned long long a0 = b.l_off[l], c0 = b.lr_off[l], base = a0 + c0;
  const int n_rough = (int)(b.lr_off[l + 1] - c0);
  const uint32_t *q = b.q + a0, *t = b.t + a0;
  int32_t *s_start = b.s_start + base, *s_end = b.s_end + base, *s_coarse = b.s_coarse + base, *s_chrom = b.s_chrom + base;
  uint32_t *s_box = b.s_box + 4 * base;
  uint8_t *s_strand = b.s_strand + base;
  float *s_freq = b.s_freq + base;
  int32_t *p_cluster = b.p_cluster + base, *p_start = b.p_start + base, *p_end = b.p_end + base;
  const uint32_t K = (uint32_t)b.globalK;
  int ns = 0, np = 0;
  auto piece = [&](int cl, int a, int e) { p_cluster[np] = cl; p_start[np] = a; p_end[np] = e; np++; };
  auto newcl = [&](int st, int en, uint32_t qS, uint32_t qE, uint32_t tS, uint32_t tE, int strand, int coarse, float freq, int chrom) {
    s_start[ns] = st; s_end[ns] = en; s_box[4 * ns] = qS; s_box[4 * ns + 1] = qE; s_box[4 * ns + 2] = tS; s_box[4 * ns + 3] = tE;
    s_strand[ns] = (uint8_t)strand; s_coarse[ns] = coarse; s_freq[ns] = freq; s_chrom[ns] = chrom;
    ns++;
    piece(ns - 1, st, en);
  };
  auto labs_ll = [](long long x) { return x < 0 ? -x : x; };
  auto close_to_prev = [&](uint32_t qS, uint32_t tS, uint32_t tE) {       // CloseToPreviousCluster(split.back(), ...)
    const uint32_t *bx = s_box + 4 * (ns - 1);
    const int st = s_strand[ns - 1];
    const long long aDiff = labs_ll((long long)qS - (long long)bx[1]);
    const long long bDiff = st == 0 ? labs_ll((long long)tS - (long long)bx[3]) : labs_ll((long long)bx[2] - (long long)tE);
    long long aDiag, bDiag;
    if (st == 0) { aDiag = (long long)bx[3] - (long long)bx[1]; bDiag = (long long)tS - (long long)qS; }
    else { aDiag = (long long)bx[1] + (long long)bx[2]; bDiag = (long long)qS + (long long)tE; }
    return (aDiff < bDiff ? aDiff : bDiff) <= (long long)b.maxGap && labs_ll(aDiag - bDiag) < (long long)b.maxDiag;
  };
  auto merge = [&](uint32_t qS, uint32_t qE, uint32_t tS, uint32_t tE, int st, int en) {   // MergeTwoClusters(split.back(), ...)
    uint32_t *bx = s_box + 4 * (ns - 1);
    if (qS < bx[0]) bx[0] = qS; if (qE > bx[1]) bx[1] = qE; if (tS < bx[2]) bx[2] = tS; if (tE > bx[3]) bx[3] = tE;
    piece(ns - 1, st, en);
    s_end[ns - 1] = en;
  };
  for (int c = 0; c < n_rough; c++) {
    const unsigned long long rc = c0 + c;
    const int os = b.r_start[rc], oe = b.r_end[rc];
    if (oe - os == 0) continue;
    const int rstrand = b.r_strand[rc], rchrom = b.r_chrom[rc];
    const float rfreq = b.r_freq[rc];
    if (rfreq >= 10.0f) { newcl(os, oe, b.r_box[4 * rc], b.r_box[4 * rc + 1], b.r_box[4 * rc + 2], b.r_box[4 * rc + 3], rstrand, c, rfreq, -1); continue; }
    const int cur_s = ns;
    int split_cs = os;
    uint32_t sqS = q[os], stS = t[os], sqE = sqS + K, stE = stS + K;
    for (int m = os + 1; m < oe; m++) {
      const long long ad = labs_ll((long long)q[m - 1] - (long long)q[m]), bd = labs_ll((long long)t[m - 1] - (long long)t[m]);
      const int gap = (int)(ad < bd ? ad : bd);
      if (gap > b.maxGap || (ns > 1 && s_chrom[ns - 1] != rchrom)) {
        if (m - split_cs >= b.minClusterSize) {
          if (ns > cur_s && s_chrom[ns - 1] == rchrom && close_to_prev(sqS, stS, stE)) merge(sqS, sqE, stS, stE, split_cs, m);
          else newcl(split_cs, m, sqS, sqE, stS, stE, rstrand, c, rfreq, rchrom);
        }
        sqS = q[m]; stS = t[m]; sqE = sqS + K; stE = stS + K; split_cs = m;
      } else {
        if (q[m] < sqS) sqS = q[m]; if (t[m] < stS) stS = t[m];
        if (q[m] + K > sqE) sqE = q[m] + K; if (t[m] + K > stE) stE = t[m] + K;
      }
    }
    if (oe - split_cs >= b.minClusterSize) {
      if (ns > cur_s && close_to_prev(sqS, stS, stE)) merge(sqS, sqE, stS, stE, split_cs, oe);
      else newcl(split_cs, oe, sqS, sqE, stS, stE, rstrand, c, rfreq, -1);
    }
  }
  b.n_split[l] = ns; b.n_piece[l] = np;
}

__global__ void __launch_bounds__(64) split_rough_kernel(SplitRoughBatch b) {
  const int l = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (l >= b.n_lists) return;
  split_rough_one(b, l);
}

// StoreDiagonalClusters (reference Clustering.h:1442-1487, RemoveSuperRepetitiveClusters :1432-1439): the clusters CleanMatches makes from the cleaned,
// diagonal-sorted anchors of one strand when opts.ExtractDiagonalFromClean is off.  A run ends where the diagonal jumps by maxDiag or more; anchorfreq
// sums matches_freq in binary32 in anchor order and is only reset when a run is KEPT, so a rejected run's frequencies leak into the next cluster: one
// thread per anchor list, literal.
struct StoreDiagBatch {
  int n_lists;
  int globalK, maxDiag, minClusterSize, minClusterLength, bypass;
  const unsigned long long *l_off;        // [n_lists + 1]
  const uint32_t *q, *t;
  const unsigned long long *qt;           // first.t of every anchor
  const float *freq;                      // matches_freq
  const uint8_t *strand;                  // [n_lists]
  const unsigned long long *hdr_pos;
  int n_hdr;
  int32_t *n_cl;                          // [n_lists] out
  int32_t *c_start, *c_end, *c_chrom;     // out, slot layout: cluster i of list l at l_off[l] + i
  uint32_t *c_box;
  float *c_freq;
};

__device__ __forceinline__ int sdc_hdr_find(const unsigned long long *pos, int n, unsigned long long query) {   // Header::Find
  if (n > 0 && query == pos[0]) return 0;
  int lo = 0, len = n;
  while (len > 0) { const int half = len >> 1; if (pos[lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
  if (lo < n && query == pos[lo]) return lo;
  return lo - 1;
}

__global__ void __launch_bounds__(64) store_diagonal_kernel(StoreDiagBatch b) {
  const int l = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (l >= b.n_lists) return;
  const unsigned long long a0 = b.l_off[l];
  const int n = (int)(b.l_off[l + 1] - a0);
  const uint32_t *q = b.q + a0, *t = b.t + a0;
  const unsigned long long *qt = b.qt + a0;
  const float *freq = b.freq + a0;
  const int strand = b.strand[l];
  const uint32_t K = (uint32_t)b.globalK;
  int nc = 0, cs = 0;
  float totalfreq = 0.0f;
  while (cs < n) {
    int ce = cs + 1;
    uint32_t qS = q[cs], qE = q[cs] + K, tS = t[cs], tE = t[cs] + K;
    totalfreq = __fadd_rn(totalfreq, freq[cs]);
    const int cI = sdc_hdr_find(b.hdr_pos, b.n_hdr, (unsigned long long)tS);
    bool rep = true;
    while (ce < n) {
      long long d;
      if (strand == 0) d = ((long long)t[ce] - (long long)q[ce]) - ((long long)t[ce - 1] - (long long)q[ce - 1]);
      else d = (long long)(uint32_t)(q[ce] + t[ce]) - (long long)(uint32_t)(q[ce - 1] + t[ce - 1]);
      if (!((d < 0 ? -d : d) < (long long)b.maxDiag)) break;
      if (q[ce] < qS) qS = q[ce]; if (q[ce] + K > qE) qE = q[ce] + K; if (t[ce] < tS) tS = t[ce]; if (t[ce] + K > tE) tE = t[ce] + K;
      totalfreq = __fadd_rn(totalfreq, freq[ce]);
      if (qt[ce] != qt[cs]) rep = false;
      ce++;
    }
    if (ce - cs >= b.minClusterSize && qE - qS >= (uint32_t)b.minClusterLength && tE - tS >= (uint32_t)b.minClusterLength && !rep) {
      const unsigned long long o = a0 + nc;
      b.c_start[o] = cs; b.c_end[o] = ce; b.c_box[4 * o] = qS; b.c_box[4 * o + 1] = qE; b.c_box[4 * o + 2] = tS; b.c_box[4 * o + 3] = tE;
      b.c_freq[o] = __fdiv_rn(totalfreq, (float)(ce - cs));
      b.c_chrom[o] = b.bypass ? sdc_hdr_find(b.hdr_pos, b.n_hdr, (unsigned long long)tS) : cI;
      totalfreq = 0.0f;
      nc++;
    }
    cs = ce;
  }
  b.n_cl[l] = nc;
}

}  // namespace lra
