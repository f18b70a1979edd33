// a16: the chain filters of the reference (Chain.h:546-960), batched over chains.
//   mode 0  RemoveSmallPairedIndels<Tup>(chain)                 :546-606
//   mode 1  RemovePairedIndels<Tup>(chain, refineEnds = true)   :611-748   float mean / sd of the anchor distances, in the reference's operation
//   mode 2  RemovePairedIndels<Tup>(chain, refineEnds = false)             order and rounding (binary32, no contraction); mixed axes of qDist kept
//   mode 3  RemovePairedIndels(matches, chain, lengths)         :754-822
//   mode 4  RemoveSpuriousAnchors<Tup>(chain)                   :828-890
//   mode 5  RemoveSpuriousJump<Tup>(chain)                      :896-960
// A chain is its anchors in chain order (q = qStart, t = tStart, len = length, strand).  Every filter is two sequential scans: collect the
// "SV" events between consecutive anchors, then look at consecutive events (mode 5 even reads the removals it has just made).  One chain per
// thread; the events of a chain live in a global scratch slot the size of the chain.  The output is the keep mask.
#pragma once
#include "lra_common.cuh"

namespace lra {

struct ChainfBatch {
  int n_chains, mode;
  const unsigned long long *off;       // [n_chains + 1]
  const uint32_t *q, *t, *len;
  const uint8_t *strand;               // unused by mode 3
  uint8_t *keep;
  int32_t *sv, *svpos, *svg;           // scratch, one slot per anchor (svg: SVgenome, only mode 3 reads it)
};

__device__ __forceinline__ int cf_sgn(int v) { return v >= 0; }
__device__ __forceinline__ int cf_abs(int v) { return v < 0 ? -v : v; }

__device__ __noinline__ void chainf_one(const ChainfBatch &b, const int ch) {
  const unsigned long long o = b.off[ch];
  const int n = (int)(b.off[ch + 1] - o);
  const uint32_t *q = b.q + o, *t = b.t + o, *len = b.len + o;
  const uint8_t *strand = b.strand + o;
  uint8_t *keep = b.keep + o;
  int32_t *SV = b.sv + o, *SVpos = b.svpos + o, *SVg = b.svg + o;
  const int mode = b.mode;
  for (int i = 0; i < n; i++) keep[i] = 1;
  if (n < 2) return;
  int ns = 0;
  const int thr = mode == 0 ? 5 : (mode == 4 ? 499 : (mode == 5 ? 100 : 30));
  // `long` sums in the reference; dist * dist overflows 64 bits when an anchor pair is out of order (dist is then an unsigned 32-bit
  // wrap-around near 2^32), and the stock build simply wraps (the mean / sd then turn negative / NaN, which is observable: no distance
  // is "valid" and every short anchor goes).  Unsigned arithmetic makes that wrap well defined here.
  unsigned long long totalDist = 0, totDistSq = 0;
  auto dists = [&](int c, long long &tDist, long long &qDist) {       // unsigned 32-bit differences widened, as in the reference
    const uint32_t te1 = t[c - 1] + len[c - 1], qe1 = q[c - 1] + len[c - 1], tec = t[c] + len[c], qec = q[c] + len[c];
    tDist = t[c] > te1 ? (long long)(uint32_t)(t[c] - te1) : (long long)(uint32_t)(t[c - 1] - tec);
    qDist = q[c] > qe1 ? (long long)(uint32_t)(q[c] - te1) : (long long)(uint32_t)(q[c - 1] - qec);
  };
  for (int c = 1; c < n; c++) {
    if (mode == 1) {
      long long tDist, qDist;
      dists(c, tDist, qDist);
      const long long dist = tDist < qDist ? tDist : qDist;
      totDistSq += (unsigned long long)dist * (unsigned long long)dist; totalDist += (unsigned long long)dist;
    }
    if (mode == 3) {
      const int Gap = (int)(((long long)t[c] - (long long)q[c]) - ((long long)t[c - 1] - (long long)q[c - 1]));
      if (cf_abs(Gap) > 30) { SV[ns] = Gap; SVg[ns] = (int)t[c]; SVpos[ns] = c; ns++; }
      continue;
    }
    if (strand[c] == strand[c - 1]) {
      int Gap;
      if (strand[c] == 0) Gap = (int)(((long long)t[c] - (long long)q[c]) - ((long long)t[c - 1] - (long long)q[c - 1]));
      else Gap = (int)((long long)(uint32_t)(q[c] + len[c] + t[c]) - (long long)(uint32_t)(q[c - 1] + len[c - 1] + t[c - 1]));
      const bool take = mode == 0 ? (cf_abs(Gap) > 5 && cf_abs(Gap) <= 50) : (cf_abs(Gap) > thr);
      if (take) { SV[ns] = Gap; SVpos[ns] = c; ns++; }
    } else { SVpos[ns] = c; SV[ns] = 0; ns++; }
  }
  if (mode == 0) {
    for (int c = 1; c < ns; c++)
      if (cf_sgn(SV[c]) != cf_sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0 && cf_abs(SV[c] + SV[c - 1]) <= 20 && SVpos[c] - SVpos[c - 1] < 3)
        for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if (len[i] <= 50) keep[i] = 0;
  } else if (mode == 1 || mode == 2) {
    for (int c = 1; c < ns; c++) {
      const bool opp = cf_sgn(SV[c]) != cf_sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0 && SVpos[c] - SVpos[c - 1] < 3;
      if (opp && ((cf_abs(SV[c]) >= 300 && cf_abs(SV[c - 1]) >= 300) || cf_abs(SV[c] + SV[c - 1]) < 100))
        for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if (len[i] < 100) keep[i] = 0;
    }
    if (mode == 1) {
      const float nDist = (float)(n - 1);
      const float meanDist = __fdiv_rn(__ll2float_rn((long long)totalDist), nDist);
      const float varDist = __fsub_rn(__fdiv_rn(__ll2float_rn((long long)totDistSq), nDist), __fmul_rn(meanDist, meanDist));
      const float sdDist = __fsqrt_rn(varDist);
      const float bound = __fadd_rn(meanDist, __fmul_rn(4.0f, sdDist));
      int firstValidDist = -1, lastValidDist = -1;
      for (int c = 1; c < n; c++) {
        long long tDist, qDist;
        dists(c, tDist, qDist);
        const int dist = (int)(tDist < qDist ? tDist : qDist);
        if ((float)dist < bound) { if (firstValidDist == -1) firstValidDist = c - 1; lastValidDist = c; }
      }
      if (lastValidDist == -1 || firstValidDist == -1) for (int i = 0; i < n; i++) if (len[i] < 100) keep[i] = 0;
      if (firstValidDist > 0 && firstValidDist < 3) for (int i = 0; i < firstValidDist; i++) if (len[i] < 100) keep[i] = 0;
      if (lastValidDist + 1 <= n && n - lastValidDist < 3) for (int i = lastValidDist + 1; i < n; i++) if (len[i] < 100) keep[i] = 0;
    }
  } else if (mode == 3) {
    for (int c = 1; c < ns; c++) {
      const int blink = imax(cf_abs(SV[c]), cf_abs(SV[c - 1]));
      const int g = SVg[c], gp = SVg[c - 1];
      const bool near = cf_sgn(SV[c]) ? (cf_abs(g - gp) < imax(2 * blink, 1000)) : (cf_abs(g - SV[c] - gp) < imax(2 * blink, 1000));
      const bool near500 = cf_sgn(SV[c]) ? (cf_abs(g - gp) < 500) : (cf_abs(g - SV[c] - gp) < 500);
      bool hit = false;
      if (cf_sgn(SV[c]) != cf_sgn(SV[c - 1]) && cf_abs(SV[c] + SV[c - 1]) < 600 && SV[c] != 0 && SV[c - 1] != 0) hit = near;
      else if (cf_sgn(SV[c]) != cf_sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0 && near500) hit = true;
      else if (cf_sgn(SV[c]) == cf_sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0) hit = near;
      if (hit) for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if ((int)len[i] < 100) keep[i] = 0;
    }
  } else if (mode == 4) {
    for (int c = 1; c < ns; c++)
      if (SV[c] != 0 && SV[c - 1] != 0 && SVpos[c] - SVpos[c - 1] <= 10) {
        bool check = false;
        for (int k = SVpos[c - 1]; k < SVpos[c]; k++) if (len[k] >= 50) { check = true; break; }
        if (!check) for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if (len[i] < 50) keep[i] = 0;
      }
  } else {
    for (int c = 1; c < ns; c++)
      if (keep[SVpos[c - 1]] == 1 && cf_sgn(SV[c]) != cf_sgn(SV[c - 1]) && SV[c] != 0 && SV[c - 1] != 0 && SVpos[c] - SVpos[c - 1] == 1)
        for (int i = SVpos[c - 1]; i < SVpos[c]; i++) if (len[i] < 50) keep[i] = 0;
  }
}

__global__ void __launch_bounds__(128) chainf_kernel(ChainfBatch b) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= b.n_chains) return;
  chainf_one(b, ch);
}

}  // namespace lra
