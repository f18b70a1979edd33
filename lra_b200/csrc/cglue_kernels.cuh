// a11: two small chain-glue routines, batched --
//   MergeChain   (reference ChainRefine.h:767-802): whether a cluster of a split chain joins the group of its predecessor depends on that pair
//                alone (same contig and strand, within 500 bases on the read and on the genome), so one thread per chain entry writes a head flag;
//                mergeinfo[r].merged_clusterIndex = the entries between two heads.
//   switchindex  (reference Mapping_ultility.h:39-161): a chain over split clusters is mapped to the clusters they were cut from, then its
//                repeats are squeezed in four passes that each read what the previous one left (links included), so one thread replays one chain.
#pragma once
#include "lra_common.cuh"

namespace lra {

struct MergeChainBatch {
  unsigned long long n_entries;
  const int32_t *sp;             // [n_entries] cluster index of every split-chain entry
  const uint8_t *first;          // [n_entries] 1 = first entry of its split chain
  const int32_t *chrom;          // per cluster
  const uint8_t *strand;
  const uint32_t *box;           // [n_clusters * 4] qStart, qEnd, tStart, tEnd
  uint8_t *head;                 // [n_entries] out
};

__global__ void __launch_bounds__(256) merge_chain_kernel(MergeChainBatch b) {
  const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= b.n_entries) return;
  if (b.first[e]) { b.head[e] = 1; return; }
  const int cur = b.sp[e], prev = b.sp[e - 1];
  int qdist = 9999, tdist = 9999;
  if (b.chrom[prev] == b.chrom[cur] && b.strand[prev] == b.strand[cur]) {
    const uint32_t pqs = b.box[4 * (size_t)prev], pts = b.box[4 * (size_t)prev + 2], pte = b.box[4 * (size_t)prev + 3];
    const uint32_t cqe = b.box[4 * (size_t)cur + 1], cts = b.box[4 * (size_t)cur + 2], cte = b.box[4 * (size_t)cur + 3];
    qdist = (pqs > cqe) ? (int)(pqs - cqe) : 0;
    if (b.strand[prev] == 0) tdist = (pts >= cte) ? (int)(pts - cte) : 9999;
    else if (b.strand[prev] == 1) tdist = (pte <= cts) ? (int)(cts - pte) : 9999;
  }
  b.head[e] = (qdist <= 500 && tdist <= 500) ? 0 : 1;
}

struct SwitchIndexBatch {
  int n_chains;
  const unsigned long long *c_off;   // [n_chains + 1]
  int32_t *ch;                       // in: split-cluster indices; out: cluster indices, n_out[k] of them, in place
  uint8_t *link;                     // link[c_off[k] + i] between entries i and i + 1 (n - 1 of them); out: nl_out[k] of them, in place
  const int32_t *coarse;             // per split cluster
  const uint32_t *cq;                // [n_clusters * 2] qStart, qEnd
  int32_t *ss, *se, *newch;          // scratch, one int per entry
  uint8_t *newlink, *flag;           // scratch, one byte per entry
  int32_t *n_out, *nl_out;           // [n_chains]
};

__device__ __noinline__ void switchindex_one(const SwitchIndexBatch &b, const int k) {
  const unsigned long long a0 = b.c_off[k];
  int n = (int)(b.c_off[k + 1] - a0);
  int32_t *ch = b.ch + a0, *ss = b.ss + a0, *se = b.se + a0, *newch = b.newch + a0;
  uint8_t *link = b.link + a0, *newlink = b.newlink + a0, *flag = b.flag + a0;
  int nl = n > 0 ? n - 1 : 0;
  for (int c = 0; c < n; c++) ch[c] = b.coarse[ch[c]];
  if (nl > 0) {                                   // drop the link between two entries of one cluster
    int sm = 0;
    for (int c = 0; c < nl; c++) if (!(ch[c + 1] == ch[c])) link[sm++] = link[c];
    nl = sm;
  }
  { int m = 0;                                    // std::unique
    for (int c = 0; c < n; c++) if (c == 0 || ch[c] != ch[m - 1]) ch[m++] = ch[c];
    n = m; }
  if (n > 0) {                                    // clusters that come back later: cut from the first appearance to the last
    int ns = 0;
    for (int c = 0; c < n; c++) {
      bool seen = false;
      for (int d = 0; d < c; d++) if (ch[d] == ch[c]) { seen = true; break; }
      if (seen) continue;
      int e = c + 1;
      for (int d = c + 1; d < n; d++) if (ch[d] == ch[c]) e = d + 1;
      if (e > c + 1) { ss[ns] = c; se[ns] = e; ns++; }      // first appearances come in ascending order: already sorted by start
    }
    int nn = 0, nnl = 0, nc = 0;
    for (int ste = 0; ste < ns; ste++) {
      while (nc <= ss[ste]) {
        newch[nn++] = ch[nc];
        if (nn > 1) newlink[nnl++] = (nc - 1 >= 0 && nc - 1 < nl) ? link[nc - 1] : 0;
        nc++;
      }
      nc = se[ste];
    }
    while (nc < n) {
      newch[nn++] = ch[nc];
      if (nn > 1) newlink[nnl++] = (nc - 1 >= 0 && nc - 1 < nl) ? link[nc - 1] : 0;
      nc++;
    }
    for (int c = 0; c < nn; c++) ch[c] = newch[c];
    for (int c = 0; c < nnl; c++) link[c] = newlink[c];
    n = nn; nl = nnl;
  }
  {                                               // clusters whose read range lies inside their predecessor's
    for (int c = 0; c < n; c++) flag[c] = 0;
    for (int c = 1; c < n; c++) {
      const int r = ch[c], p = ch[c - 1];
      if (flag[c - 1] == 0 && b.cq[2 * (size_t)r] >= b.cq[2 * (size_t)p] && b.cq[2 * (size_t)r + 1] <= b.cq[2 * (size_t)p + 1]) flag[c] = 1;
    }
    int sc = 0;
    for (int c = 0; c < n; c++) if (flag[c] == 0) { ch[sc] = ch[c]; if (sc >= 1) link[sc - 1] = link[c - 1]; sc++; }
    n = sc; nl = sc - 1 > 0 ? sc - 1 : 0;
  }
  b.n_out[k] = n; b.nl_out[k] = nl;
}

__global__ void __launch_bounds__(64) switchindex_kernel(SwitchIndexBatch b) {
  const int k = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (k >= b.n_chains) return;
  switchindex_one(b, k);
}

// a17: SwitchToOriginalAnchors (reference LocalRefineAlignment.h:187-198).  Entry i of a FinalChain names run k = chain[i] of the same-diagonal runs
// (MergeMatchesSameDiag) of extended cluster c = ClusterNum(i); the UltimateChain gets the run's anchors end[k]-1 .. start[k] (descending) with
// ClusterIndex = the cluster's `coarse`.  The entries' output ranges are an exclusive scan of end - start; one thread per entry writes its range.
struct SwitchOrigBatch {
  unsigned long long n_entries;
  const int32_t *run_start, *run_end;    // [n_entries] ExtendClusters[c]->start[k], ->end[k]
  const int32_t *coarse;                 // [n_entries] ExtendClusters[c]->coarse
  unsigned long long *off;               // [n_entries + 1] counts, then their exclusive scan
  uint32_t *chain;                       // out
  int32_t *cluster_index;                // out
};

__global__ void __launch_bounds__(256) switch_orig_count_kernel(SwitchOrigBatch b) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n_entries) return;
  const int n = b.run_end[i] - b.run_start[i];
  b.off[i] = n > 0 ? (unsigned long long)n : 0ull;
}

__global__ void __launch_bounds__(256) switch_orig_emit_kernel(SwitchOrigBatch b) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n_entries) return;
  unsigned long long o = b.off[i];
  for (int j = b.run_end[i] - 1; j >= b.run_start[i]; j--, o++) { b.chain[o] = (uint32_t)j; b.cluster_index[o] = b.coarse[i]; }
}

// TrimSplitChainDiagonal (reference ChainRefine.h:189-331; Map_lowacc.h:331): the refined anchors of a split chain, Cartesian-sorted by the a6 kernel first
// (chains of one anchor are skipped there: mode 255), are kept or dropped by the diagonal window of the two chain anchors around them.  The reference walks
// chain anchors and refined anchors with two monotone cursors (forward chains) or only looks at the first two chain anchors from the far end (reverse
// chains): one thread per chain replays it.  Diagonals are GenomePos differences (uint32 wrap), widened to 64 bits for the +-100 window.
struct TrimChainBatch {
  int n_chains;
  const unsigned long long *c_off;      // [n_chains + 1] chain anchors
  const uint32_t *cq, *ct;              // qStart / tStart of the chain anchors, sptc order
  const uint8_t *strand;                // [n_chains] SplitChain::Strand
  const unsigned long long *m_off;      // [n_chains + 1] refined anchors
  const uint32_t *q, *t;                // sorted
  uint8_t *keep;                        // out
  int32_t *removed;                     // [n_chains] out
};

__global__ void __launch_bounds__(64) trim_splitchain_kernel(TrimChainBatch b) {
  const int c = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (c >= b.n_chains) return;
  const unsigned long long a0 = b.c_off[c], m0 = b.m_off[c];
  const int nch = (int)(b.c_off[c + 1] - a0), n = (int)(b.m_off[c + 1] - m0);
  const uint32_t *cq = b.cq + a0, *ct = b.ct + a0, *q = b.q + m0, *t = b.t + m0;
  uint8_t *keep = b.keep + m0;
  for (int i = 0; i < n; i++) keep[i] = 1;
  int removed = 0;
  if (nch != 1) {
    auto cd = [&](int i) { return (uint32_t)(ct[i] - cq[i]); };
    if (b.strand[c] == 0) {
      int mi = 0;
      for (int ci = 0; ci < nch - 1; ci++) {
        const uint32_t m = cd(ci) < cd(ci + 1) ? cd(ci) : cd(ci + 1);
        const long long minDiag = (long long)m - 100, maxDiag = (long long)m + 100;
        while (mi < n && q[mi] < cq[ci + 1]) {
          const long long d = (long long)(uint32_t)(t[mi] - q[mi]);
          if (d < minDiag || d > maxDiag) { keep[mi] = 0; removed++; }
          mi++;
        }
      }
    } else if (nch >= 2) {
      const uint32_t m = cd(0) < cd(1) ? cd(0) : cd(1);
      const long long minDiag = (long long)m - 100, maxDiag = (long long)m + 100;
      for (int mi = n - 1; mi >= 0 && q[mi] > cq[1]; mi--) {
        const long long d = (long long)(uint32_t)(t[mi] - q[mi]);
        if (d < minDiag || d > maxDiag) { keep[mi] = 0; removed++; }
      }
    }
  }
  b.removed[c] = removed;
}

}  // namespace lra
