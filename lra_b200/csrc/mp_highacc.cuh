// Mapper worker, MapRead_highacc (Map_highacc.h:37-798; -CCS / -CONTIG): MatchesToFineClusters (Clustering.h:1555-1676) with StoreFineClusters
// (:892-1331), SplitClusters + DecideSplitClustersValue (SplitClusters.h), the first SparseDP over the split clusters (SparseDP.h:1956), switchindex
// (Mapping_ultility.h:39-161), RefineBtwnClusters_chain + RefineBtwnSpace (ClusterRefine.h:331-614), LinearExtend_chain + MergeMatchesSameDiag
// (LinearExtend.h:134-350, 782-823), SPLITChain on Cluster_SameDiag + MergeSplitchainINS (Mapping_ultility.h:163-353), and LocalRefineAlignment
// (LocalRefineAlignment.h:553-768) with the second SparseDP over same-diagonal runs (SparseDP.h:1766).
// The serial, state-carrying walks are replayed by lane 0 with the pinned routines of the stage kernels (split_kernels.cuh, srough_kernels.cuh,
// cglue_kernels.cuh, lext_kernels.cuh, chainf_kernels.cuh); sorts, the sparse DPs, RefineSpace and the alignment use the whole warp.
#pragma once
#include "mp_align.cuh"
#include "split_kernels.cuh"
#include "srough_kernels.cuh"
#include "cglue_kernels.cuh"
#include "lext_kernels.cuh"

namespace lra {
namespace mp {


__device__ __forceinline__ long long ha_labs(long long x) { return x < 0 ? -x : x; }
// DiagonalDifference / minGapDifference (Clustering.h:503-537) on anchors a, b of a list
__device__ __forceinline__ long long ha_diagdiff(const uint32_t *q, const uint32_t *t, int a, int b, int strand) {
  if (strand == 0) return ((long long)t[a] - (long long)q[a]) - ((long long)t[b] - (long long)q[b]);
  return (long long)(uint32_t)(q[a] + t[a]) - (long long)(uint32_t)(q[b] + t[b]);
}
__device__ __forceinline__ long long ha_mingap(const uint32_t *q, const uint32_t *t, int a, int b) {
  const long long x = ha_labs((long long)q[b] - (long long)q[a]), y = ha_labs((long long)t[b] - (long long)t[a]);
  return x < y ? x : y;
}

// ---- growing list of fine clusters (lane 0) -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fc_new(ClusterSet &F, int strand) { const int c = F.ncl; F.off[c + 1] = F.off[c]; F.strand[c] = strand; F.chrom[c] = -1; F.freq[c] = 0.0f; F.ncl++; }
__device__ __forceinline__ void fc_push(ClusterSet &F, uint32_t q, uint32_t t) { const int c = F.ncl - 1; const int a = F.off[c + 1]++; F.q[a] = q; F.t[a] = t; }
__device__ __forceinline__ int fc_size(const ClusterSet &F, int c) { return F.off[c + 1] - F.off[c]; }
__device__ inline void fc_set_bounds(ClusterSet &F, int c, int K) {                  // Cluster::SetClusterBoundariesFromMatches(opts, false)
  const int a0 = F.off[c], a1 = F.off[c + 1];
  uint32_t qS = F.q[a0], qE = qS + (uint32_t)K, tS = F.t[a0], tE = tS + (uint32_t)K;
  for (int i = a0 + 1; i < a1; i++) {
    tE = F.t[i] + (uint32_t)K > tE ? F.t[i] + (uint32_t)K : tE; tS = F.t[i] < tS ? F.t[i] : tS;
    qE = F.q[i] + (uint32_t)K > qE ? F.q[i] + (uint32_t)K : qE; qS = F.q[i] < qS ? F.q[i] : qS;
  }
  F.qS[c] = qS; F.qE[c] = qE; F.tS[c] = tS; F.tE[c] = tE;
}
__device__ inline int fc_chrom_index(const MpCtx &C, ClusterSet &F, int c) {       // Cluster::CHROMIndex (Clustering.h:326-336)
  if (fc_size(F, c) == 0) return 1;
  const int first = lref_hdr_find(C.ix.hdr_pos, C.ix.n_hdr, (unsigned long long)F.tS[c] + 1ull);
  const int last = lref_hdr_find(C.ix.hdr_pos, C.ix.n_hdr, (unsigned long long)F.tE[c]);
  if (first != last) return 1;
  F.chrom[c] = first;
  return 0;
}

// StoreFineClusters (Clustering.h:892-1331) for one split rough cluster: mq / mt the strand's cleaned matches, smi[0..n) its splitmatchindex.
// scr: 6 * n + 16 ints.  Lane 0.
__device__ __noinline__ void mp_store_fine(const MpCtx &C, const uint32_t *mq, const uint32_t *mt, const int *smi, int n, float anchorfreq, int ri, int strand,
                                           ClusterSet &F, int *scr) {
  const MpOpts &O = C.o;
  const int K = O.globalK;
  if (n == 1) return;
  if ((double)fabsf(__fsub_rn(anchorfreq, 1.0f)) <= 0.005) {
    fc_new(F, strand);
    for (int i = 0; i < n; i++) fc_push(F, mq[smi[i]], mt[smi[i]]);
    const int c = F.ncl - 1;
    fc_set_bounds(F, c, K); F.chrom[c] = ri; F.freq[c] = 1.0f;
    if (fc_chrom_index(C, F, c) == 1) F.ncl--;
    return;
  }
  int *match_num = scr, *pos_start = scr + n + 2, *Start = scr + 2 * n + 4, *End = scr + 3 * n + 6, *AddOrNot = scr + 4 * n + 8, *cidx = scr + 5 * n + 10;
  int M = 0;
  { int oc = 1, us = 0;
    for (int i = 1; i < n; i++) {
      if (mq[smi[i]] == mq[smi[i - 1]]) oc++;
      else { match_num[M] = oc; pos_start[M] = us; M++; us = i; oc = 1; }
      if (i == n - 1) { match_num[M] = oc; pos_start[M] = us; M++; }
    } }
  auto DD = [&](int a, int b) { return ha_diagdiff(mq, mt, smi[a], smi[b], strand); };
  auto MG = [&](int a, int b) { return ha_mingap(mq, mt, smi[a], smi[b]); };
  int u_start = 0, u_end = 0, u_maxstart = 0, u_maxend = 0, max_pos = 0, NS = 0;
  if (M == 1) { u_maxstart = 0; u_maxend = 1; Start[NS] = 0; End[NS] = 1; NS++; }
  else {
    int k = 0;
    while (k < M - 1) {
      while (k < M - 1 && match_num[k] != 1) k++;
      u_start = k; u_end = k + 1;
      while (k < M - 1 && match_num[k + 1] == match_num[k] && ha_labs(DD(pos_start[k + 1], pos_start[k])) < (long long)O.maxDiag &&
             MG(pos_start[k + 1], pos_start[k]) <= (long long)O.maxGap) { u_end = k + 2; k++; }
      Start[NS] = u_start; End[NS] = u_end; NS++;
      k++;
      if ((u_maxstart == 0 && u_maxend == 0) || (u_maxend - u_maxstart < u_end - u_start)) { u_maxstart = u_start; u_maxend = u_end; max_pos = NS - 1; }
    }
  }
  if (u_maxstart == 0 && u_maxend == 0) return;
  int c_s = pos_start[u_maxstart], c_e = pos_start[u_maxend - 1] + 1;
  if (!(c_e - c_s >= O.minUniqueStretchNum && (uint32_t)(mq[smi[c_e - 1]] + (uint32_t)K - mq[smi[c_s]]) >= (uint32_t)O.minUniqueStretchDist)) return;
  fc_new(F, strand);
  for (int i = 0; i < NS; i++) AddOrNot[i] = 0;
  if (c_e - c_s == n) {
    for (int i = c_s; i < c_e; i++) fc_push(F, mq[smi[i]], mt[smi[i]]);
    F.freq[F.ncl - 1] = anchorfreq;
    AddOrNot[0] = 1;
  } else {
    auto joins = [&](int i_m, int prev_anchor) {
      const long long mg = MG(i_m, prev_anchor);
      return (ha_labs(DD(i_m, prev_anchor)) <= (long long)O.maxDiag && mg <= (long long)O.maxGap) || mg <= (long long)(O.maxGap / 2);
    };
    int prev_anchor = c_s;
    AddOrNot[max_pos] = 1;
    for (int i = max_pos - 1; i >= 0; i--) {
      const int i_m = pos_start[End[i] - 1];
      if (joins(i_m, prev_anchor)) { AddOrNot[i] = 1; prev_anchor = pos_start[Start[i]]; }
    }
    prev_anchor = c_e - 1;
    for (int i = max_pos + 1; i < NS; i++) {
      const int i_m = pos_start[Start[i]];
      if (joins(i_m, prev_anchor)) { AddOrNot[i] = 1; prev_anchor = pos_start[End[i] - 1]; }
    }
    // StretchOfOne walked from its back: the added stretches in ascending order
    int first_it = -1, last_it = -1;
    for (int i = 0; i < NS; i++) if (AddOrNot[i]) { if (first_it < 0) first_it = i; last_it = i; }
    int prev_stretch = -1, p_s = 0, p_e = 0;
    for (int it = 0; it < NS; it++) {
      if (!AddOrNot[it]) continue;
      c_s = pos_start[Start[it]]; c_e = pos_start[End[it] - 1] + 1;
      if (it == first_it) { p_s = it == 0 ? 0 : pos_start[End[it - 1]]; p_e = pos_start[Start[it]]; }
      else { p_s = pos_start[End[prev_stretch]]; p_e = pos_start[Start[it]]; }
      prev_stretch = it;
      int prev_match = c_s, nci = 0;
      for (int si = p_e - 1; si >= p_s; si--)
        if (ha_labs(DD(si, prev_match)) < (long long)O.maxDiag) { cidx[nci++] = si; prev_match = si; }
      for (int ci = nci - 1; ci >= 0; ci--) fc_push(F, mq[smi[cidx[ci]]], mt[smi[cidx[ci]]]);
      for (int si = c_s; si < c_e; si++) fc_push(F, mq[smi[si]], mt[smi[si]]);
      if (it == last_it) {
        p_s = pos_start[End[it] - 1] + 1;
        p_e = it == NS - 1 ? n : pos_start[Start[it + 1]];
        prev_match = c_e - 1;
        for (int si = p_s; si < p_e; si++)
          if (ha_labs(DD(si, prev_match)) < (long long)O.maxDiag) { fc_push(F, mq[smi[si]], mt[smi[si]]); prev_match = si; }
      }
    }
    F.freq[F.ncl - 1] = anchorfreq;
  }
  { const int c = F.ncl - 1; fc_set_bounds(F, c, K); F.chrom[c] = ri; }
  if (F.ncl > 0) {
    const int c = F.ncl - 1;
    if (fc_chrom_index(C, F, c) == 1) F.ncl--;
    else if (fc_size(F, c) <= O.minClusterSize) F.ncl--;
    else if (F.qE[c] == F.qS[c]) F.ncl--;
    else if (((long long)F.tE[c] - (long long)F.tS[c]) >= 5ll * ((long long)F.qE[c] - (long long)F.qS[c])) F.ncl--;
  }
  for (int ar_ = 0; ar_ < NS; ar_++) {
    if (!AddOrNot[ar_] && End[ar_] - Start[ar_] >= 15) {
      fc_new(F, strand);
      for (int i = pos_start[Start[ar_]]; i < pos_start[End[ar_] - 1] + 1; i++) fc_push(F, mq[smi[i]], mt[smi[i]]);
      const int c = F.ncl - 1;
      fc_set_bounds(F, c, K); F.chrom[c] = ri; F.freq[c] = anchorfreq;
      if (fc_chrom_index(C, F, c) == 1) F.ncl--;
      if (F.ncl > 0) {        // (the reference tests clusters.back() again, whichever cluster that now is)
        const int b = F.ncl - 1;
        if (((long long)F.tE[b] - (long long)F.tS[b]) / ((long long)F.qE[b] - (long long)F.qS[b]) >= 5) F.ncl--;
      }
    }
  }
}

// MatchesToFineClusters for both strands: raw = cleaned matches + rough clusters of mp_seed_clean -> F (fine clusters, global t)
__device__ __noinline__ int mp_fine_clusters(const MpCtx &C, Arena &ar, ClusterSet &raw, int n_raw_a, ClusterSet &F) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  if (!mp_alloc_clusterset(F, ar, n_raw_a + 2, n_raw_a + 2, false)) return MP_ERR_ARENA;
  if (lane == 0) F.off[0] = 0;
  wsync();
  int c0 = 0;
  for (int s = 0; s < 2; s++) {
    int c1 = c0;
    while (c1 < raw.ncl && raw.strand[c1] == s) c1++;
    const int nr = c1 - c0;
    if (nr == 0) continue;
    const int A0 = raw.off[c0], nA = raw.off[c1] - A0;
    const unsigned long long mk = ar.mark();
    // CartesianSort of every rough cluster's range
    for (int c = c0; c < c1; c++) {
      const int a0 = raw.off[c], n = raw.off[c + 1] - a0;
      if (n < 2) continue;
      const unsigned long long mk2 = ar.mark();
      MpKey *keys = ar.alloc<MpKey>((unsigned long long)next_pow2(n));
      if (ar.overflow) return MP_ERR_ARENA;
      for (int i = lane; i < n; i += kLanes) { keys[i].k = ((unsigned long long)raw.q[a0 + i] << 32) | raw.t[a0 + i]; keys[i].q = 0; keys[i].idx = (uint32_t)i; }
      wsync();
      mp_sort_keys(keys, n);
      for (int i = lane; i < n; i += kLanes) { raw.q[a0 + i] = (uint32_t)(keys[i].k >> 32); raw.t[a0 + i] = (uint32_t)keys[i].k; }
      wsync();
      ar.release(mk2);
    }
    // SplitRoughClustersWithGaps over the strand's rough clusters, then StoreFineClusters per split cluster
    const unsigned long long slots = (unsigned long long)nA + nr + 2;
    unsigned long long *l_off = ar.alloc<unsigned long long>(2), *lr_off = ar.alloc<unsigned long long>(2);
    int32_t *r_start = ar.alloc<int32_t>(nr), *r_end = ar.alloc<int32_t>(nr), *r_chrom = ar.alloc<int32_t>(nr);
    uint32_t *r_box = ar.alloc<uint32_t>(4ull * nr);
    uint8_t *r_strand = ar.alloc<uint8_t>(nr);
    float *r_freq = ar.alloc<float>(nr);
    int32_t *n_out = ar.alloc<int32_t>(2);
    int32_t *s_start = ar.alloc<int32_t>(slots), *s_end = ar.alloc<int32_t>(slots), *s_coarse = ar.alloc<int32_t>(slots), *s_chrom = ar.alloc<int32_t>(slots);
    uint32_t *s_box = ar.alloc<uint32_t>(4ull * slots);
    uint8_t *s_strand = ar.alloc<uint8_t>(slots);
    float *s_freq = ar.alloc<float>(slots);
    int32_t *p_cluster = ar.alloc<int32_t>(slots), *p_start = ar.alloc<int32_t>(slots), *p_end = ar.alloc<int32_t>(slots);
    int *smi = ar.alloc<int>(nA + 2), *scr = ar.alloc<int>(6ull * nA + 32);
    if (ar.overflow) return MP_ERR_ARENA;
    for (int c = lane; c < nr; c += kLanes) {
      r_start[c] = raw.off[c0 + c] - A0; r_end[c] = raw.off[c0 + c + 1] - A0; r_chrom[c] = -1; r_strand[c] = (uint8_t)s; r_freq[c] = raw.freq[c0 + c];
      r_box[4 * c] = raw.qS[c0 + c]; r_box[4 * c + 1] = raw.qE[c0 + c]; r_box[4 * c + 2] = raw.tS[c0 + c]; r_box[4 * c + 3] = raw.tE[c0 + c];
    }
    wsync();
    if (lane == 0) {
      l_off[0] = 0; l_off[1] = (unsigned long long)nA; lr_off[0] = 0; lr_off[1] = (unsigned long long)nr;
      SplitRoughBatch b;
      b.n_lists = 1; b.globalK = O.globalK; b.maxGap = O.RoughClustermaxGap; b.minClusterSize = O.minClusterSize; b.maxDiag = O.maxDiag;
      b.l_off = l_off; b.lr_off = lr_off; b.q = raw.q + A0; b.t = raw.t + A0; b.r_start = r_start; b.r_end = r_end; b.r_box = r_box; b.r_strand = r_strand;
      b.r_freq = r_freq; b.r_chrom = r_chrom; b.n_split = n_out; b.n_piece = n_out + 1; b.s_start = s_start; b.s_end = s_end; b.s_coarse = s_coarse;
      b.s_chrom = s_chrom; b.s_box = s_box; b.s_strand = s_strand; b.s_freq = s_freq; b.p_cluster = p_cluster; b.p_start = p_start; b.p_end = p_end;
      split_rough_one(b, 0);
      const int ns = n_out[0], np = n_out[1];
      int pi = 0;
      for (int c = 0; c < ns; c++) {
        int n = 0;
        while (pi < np && p_cluster[pi] == c) { for (int i = p_start[pi]; i < p_end[pi]; i++) smi[n++] = i; pi++; }
        const int rci = lref_hdr_find(C.ix.hdr_pos, C.ix.n_hdr, (unsigned long long)s_box[4 * c + 2]);
        mp_store_fine(C, raw.q + A0, raw.t + A0, smi, n, s_freq[c], rci, s, F, scr);
      }
    }
    wsync();
    ar.release(mk);
    c0 = c1;
  }
  wsync();
  F.ncl = bcast(F.ncl, 0);
  return MP_OK;
}

// the chains of Primary_chains[0] over the clusters that survive the first SparseDP
struct HaChains {
  int nch;
  int *ch[kMaxChains]; uint8_t *link[kMaxChains]; int n[kMaxChains], nl[kMaxChains];
  float value[kMaxChains]; int n0[kMaxChains];
};

// SplitClusters + DecideSplitClustersValue + SparseDP (SparseDP.h:1956) + switchindex + the removal of clusters off the chains (Map_highacc.h:157-330).
// On return the chains index `keep` (clusters of F on a chain, ascending), nkeep of them.
__device__ __noinline__ int mp_first_sdp_highacc(const MpCtx &C, int r, Arena &ar, ClusterSet &F, HaChains &H, int *&keep, int &nkeep) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  const int ncl = F.ncl;
  H.nch = 0; nkeep = 0;
  // ---- SplitClusters, DecideSplitClustersValue
  unsigned long long *cl_off = ar.alloc<unsigned long long>(2), *sp_off = ar.alloc<unsigned long long>(2), *m_off = ar.alloc<unsigned long long>(ncl + 1);
  uint32_t *box = ar.alloc<uint32_t>(4ull * ncl), *sets = ar.alloc<uint32_t>(4ull * ncl + 4);
  uint8_t *strand = ar.alloc<uint8_t>(ncl), *split = ar.alloc<uint8_t>(ncl);
  int32_t *val_cluster = ar.alloc<int32_t>(ncl);
  ScPoint *pts = ar.alloc<ScPoint>(4ull * ncl + 4);
  int *ns_p = ar.alloc<int>(2);
  if (ar.overflow) return MP_ERR_ARENA;
  for (int c = lane; c < ncl; c += kLanes) {
    box[4 * c] = F.qS[c]; box[4 * c + 1] = F.qE[c]; box[4 * c + 2] = F.tS[c]; box[4 * c + 3] = F.tE[c]; strand[c] = (uint8_t)(F.strand[c] != 0);
    m_off[c] = (unsigned long long)F.off[c];
  }
  if (lane == 0) { m_off[ncl] = (unsigned long long)F.off[ncl]; cl_off[0] = 0; cl_off[1] = (unsigned long long)ncl; }
  wsync();
  SplitBatch sb;
  sb.n_reads = 1; sb.contig = O.readType == 3 ? 1 : 0; sb.globalK = O.globalK; sb.cl_off = cl_off; sb.box = box; sb.strand = strand; sb.freq = F.freq; sb.m_off = m_off;
  sb.mq = F.q; sb.split = split; sb.val_cluster = val_cluster; sb.sp_off = sp_off; sb.sp = 0; sb.sp_val = 0; sb.sp_n0 = 0; sb.sp_cap = 0; sb.sets = sets; sb.pts = pts;
  if (lane == 0) { split_one<false>(sb, 0); ns_p[0] = (int)sp_off[0]; }
  wsync();
  const int ns = ns_p[0];
  if (ns == 0) return MP_UNALIGNED;
  uint32_t *sp = ar.alloc<uint32_t>(6ull * ns);
  int32_t *sp_val = ar.alloc<int32_t>(ns), *sp_n0 = ar.alloc<int32_t>(ns);
  if (ar.overflow) return MP_ERR_ARENA;
  sb.sp = sp; sb.sp_val = sp_val; sb.sp_n0 = sp_n0; sb.sp_cap = (unsigned long long)ns;
  if (lane == 0) { sp_off[0] = 0; sp_off[1] = (unsigned long long)ns; split_one<true>(sb, 0); }
  wsync();
  // ---- SparseDP over the split clusters
  uint32_t *fq = ar.alloc<uint32_t>(ns), *ft = ar.alloc<uint32_t>(ns), *fqe = ar.alloc<uint32_t>(ns), *fte = ar.alloc<uint32_t>(ns);
  int32_t *flen = ar.alloc<int32_t>(ns);
  uint8_t *fstr = ar.alloc<uint8_t>(ns);
  float *fval = ar.alloc<float>(ns);
  const int NA = O.NumAln < kMaxChains ? O.NumAln : kMaxChains;
  SdpChain *sc = ar.alloc<SdpChain>(NA);
  int *n0 = ar.alloc<int>(NA);
  uint32_t *cbuf = ar.alloc<uint32_t>((unsigned long long)NA * (ns + 1));
  uint8_t *lbuf = ar.alloc<uint8_t>((unsigned long long)NA * (ns + 1));
  int32_t *chb = ar.alloc<int32_t>((unsigned long long)NA * (ns + 1));
  if (ar.overflow) return MP_ERR_ARENA;
  for (int i = lane; i < ns; i += kLanes) {
    fq[i] = sp[6 * i]; fqe[i] = sp[6 * i + 1]; ft[i] = sp[6 * i + 2]; fte[i] = sp[6 * i + 3]; fstr[i] = (uint8_t)sp[6 * i + 4]; fval[i] = (float)sp_val[i]; flen[i] = 0;
  }
  if (lane == 0) for (int c = 0; c < NA; c++) { sc[c].chain = cbuf + (unsigned long long)c * (ns + 1); sc[c].link = lbuf + (unsigned long long)c * (ns + 1); sc[c].n = 0; n0[c] = 0; }
  wsync();
  float rate = O.initial_anchorbonus;
  if ((unsigned)ns / (unsigned)ncl > 20u) rate = (float)((double)rate / 2.0);
  SdpAnchors A; A.q = fq; A.t = ft; A.len = flen; A.nfrag = ns; A.cl_off = 0; A.cl_strand = 0; A.ncl = 0; A.qe = fqe; A.te = fte; A.fstrand = fstr; A.fval = fval; A.fn0 = sp_n0;
  unsigned long long tk = mp_clock();
  const int nc = sdp_split_clusters(A, rate, O.alnthres, O.globalK, NA, (int)L, *C.pwl, ar, sc, n0);
  tk = mp_tick(C, PF_SDP1, tk);
  if (nc < 0) return MP_ERR_ARENA;
  wsync();
  if (nc == 0) return MP_UNALIGNED;
  // ---- switchindex: split-cluster indices -> clusters, repeats squeezed
  unsigned long long *c_off = ar.alloc<unsigned long long>(2);
  int32_t *coarse = ar.alloc<int32_t>(ns), *ss = ar.alloc<int32_t>(ns + 1), *se = ar.alloc<int32_t>(ns + 1), *newch = ar.alloc<int32_t>(ns + 1), *nout = ar.alloc<int32_t>(2);
  uint8_t *newlink = ar.alloc<uint8_t>(ns + 1), *flag = ar.alloc<uint8_t>(ns + 1), *used = ar.alloc<uint8_t>(ncl + 1);
  uint32_t *cq = ar.alloc<uint32_t>(2ull * ncl);
  int *renum = ar.alloc<int>(ncl + 1);
  keep = ar.alloc<int>(ncl + 1);
  int *res = ar.alloc<int>(2);
  if (ar.overflow) return MP_ERR_ARENA;
  for (int i = lane; i < ns; i += kLanes) coarse[i] = (int32_t)sp[6 * i + 5];
  for (int c = lane; c < ncl; c += kLanes) { cq[2 * c] = F.qS[c]; cq[2 * c + 1] = F.qE[c]; used[c] = 0; }
  wsync();
  for (int h = 0; h < nc; h++) {
    int32_t *ch = chb + (unsigned long long)h * (ns + 1);
    const int n = sc[h].n;
    for (int i = lane; i < n; i += kLanes) ch[i] = (int32_t)sc[h].chain[i];
    wsync();
    if (lane == 0) {
      c_off[0] = 0; c_off[1] = (unsigned long long)n;
      SwitchIndexBatch b;
      b.n_chains = 1; b.c_off = c_off; b.ch = ch; b.link = sc[h].link; b.coarse = coarse; b.cq = cq; b.ss = ss; b.se = se; b.newch = newch; b.newlink = newlink; b.flag = flag;
      b.n_out = nout; b.nl_out = nout + 1;
      switchindex_one(b, 0);
      for (int i = 0; i < nout[0]; i++) used[ch[i]] = 1;
    }
    wsync();
    H.ch[h] = (int *)ch; H.link[h] = sc[h].link; H.n[h] = nout[0]; H.nl[h] = nout[1]; H.value[h] = sc[h].value; H.n0[h] = n0[h];
    wsync();
  }
  H.nch = nc;
  // ---- clusters off every chain are dropped; the chains are renumbered (Map_highacc.h:283-327)
  if (lane == 0) {
    int lm = 0;
    for (int s = 0; s < ncl; s++) if (used[s]) { renum[s] = lm; keep[lm++] = s; }
    for (int h = 0; h < nc; h++) for (int i = 0; i < H.n[h]; i++) H.ch[h][i] = renum[H.ch[h][i]];
    res[0] = lm;
  }
  wsync();
  nkeep = res[0];
  if (nkeep == 0) return MP_UNALIGNED;
  return MP_OK;
}

// RefineBtwnSpace (ClusterRefine.h:331-431), both forms.  RevBtwnCluster only collects clusters nothing reads afterwards, so its branch leaves no trace.
__device__ __noinline__ bool mp_refine_btwn_space_ha(const MpCtx &C, int r, Arena &ar, int K, int W, bool twoblocks, RCluster &cl, uint32_t qe, uint32_t qs, uint32_t te, uint32_t ts,
                                                     int st, uint32_t lrts, uint32_t lrlength) {
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  if (st == 1) { const uint32_t t = qs; qs = L - qe; qe = L - t; }
  int refineSpaceDiag = 0;
  if (O.readType == 3 || O.readType == 2) { const float v = fmaxf(100.f, __fmul_rn(0.01f, (float)(uint32_t)(qe - qs))); const int f = (int)floorf(v); refineSpaceDiag = f < 100 ? f : 100; }
  else { const float v = fmaxf(100.f, __fmul_rn(0.15f, (float)(uint32_t)(qe - qs))); const int f = (int)floorf(v); refineSpaceDiag = f < 1000 ? f : 1000; }
  uint32_t *pq, *pt; float ident;
  RSeg *node = ar.alloc<RSeg>(1);
  if (ar.overflow) return false;
  const int np = mp_refine_space(C, r, ar, K, W, refineSpaceDiag, true, O.localMaxFreq, cl.chrom, qe, qs, te, ts, st, lrts, lrlength, &pq, &pt, &ident);
  if (np < 0) return false;
  const uint32_t mn = (qe - qs) < (te - ts) ? (qe - qs) : (te - ts);
  const float eff = __fdiv_rn((float)np, (float)mn);
  if ((np > 0 && twoblocks) || (np > 0 && eff >= __fmul_rn(O.anchorstoosparse, 2.0f))) {
    if (lane_id() == 0) { rc_append(cl, node, pq, pt, np); rc_set_boundaries(cl, K); cl.refinespace = 1; }
    wsync();
    return true;
  }
  if (twoblocks) return true;
  const int rst = st == 1 ? 0 : 1;
  { const uint32_t t = qs; qs = L - qe; qe = L - t; }
  uint32_t *rq, *rt;
  const int nrev = mp_refine_space(C, r, ar, K, W, refineSpaceDiag, true, O.localMaxFreq, cl.chrom, qe, qs, te, ts, rst, lrts, lrlength, &rq, &rt, &ident);
  if (nrev < 0) return false;
  const uint32_t mn2 = (qe - qs) < (te - ts) ? (qe - qs) : (te - ts);
  const float reff = __fdiv_rn((float)nrev, (float)mn2);
  if (eff >= reff) {
    if (lane_id() == 0) { if (np > 0) rc_append(cl, node, pq, pt, np); rc_set_boundaries(cl, K); cl.refinespace = 1; cl.freq = 1.0f; }
    wsync();
  }
  return true;
}

// RefineBtwnClusters_chain (ClusterRefine.h:433-614) for one chain over the refined clusters RC
__device__ __noinline__ bool mp_refine_btwn_clusters_chain(const MpCtx &C, int r, Arena &ar, int K, int W, const int *ch, int n, RCluster *RC) {
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  const bool contig = O.readType == 3;
  const uint32_t low_b = contig ? 1000u : 20u, upper = contig ? 100000u : 50000u;
  bool twoblocks = false;
  int st1 = 0, st2 = 0;
  for (int c = 1; c < n; c++) {
    wsync();
    RCluster &cur = RC[ch[c]], &prev = RC[ch[c - 1]];
    const uint32_t qs = cur.qE, qe = prev.qS;
    uint32_t te1 = 0, ts1 = 0, te2 = 0, ts2 = 0;
    if (qe <= qs || cur.chrom != prev.chrom) continue;
    if (contig) twoblocks = false;
    const uint32_t clen = contig_len(C.ix, cur.chrom);
    const uint32_t d = qe - qs;
    if (cur.strand == prev.strand) {
      twoblocks = false; st1 = cur.strand;
      if (cur.tE <= prev.tS) { ts1 = cur.tE; te1 = prev.tS; }
      else if (cur.tS > prev.tE) { ts1 = prev.tE; te1 = cur.tS; }
      else continue;
    } else if (!contig) {
      st1 = cur.strand; st2 = prev.strand; twoblocks = true;
      if (cur.tE <= prev.tS) {
        if (st1 == 0) { ts1 = cur.tE; te1 = clen < ts1 + d ? clen : ts1 + d; ts2 = prev.tE; te2 = clen < ts2 + d ? clen : ts2 + d; }
        else { te1 = cur.tS; ts1 = te1 > d ? te1 - d : 0; te2 = prev.tS; ts2 = te2 > d ? te2 - d : 0; }
      } else if (cur.tS > prev.tE) {
        if (st1 == 0) { ts1 = cur.tE; te1 = clen < ts1 + d ? clen : ts1 + d; te2 = cur.tS; ts2 = te2 > d ? te2 - d : 0; }
        else { te1 = cur.tS; ts1 = te1 > d ? te1 - d : 0; te2 = prev.tS; ts2 = te2 > d ? te2 - d : 0; }
      } else continue;
    }
    if (te1 <= ts1) continue;
    const uint32_t sl1 = d > (te1 - ts1) ? d : (te1 - ts1);
    if (sl1 >= low_b && sl1 <= upper) { if (!mp_refine_btwn_space_ha(C, r, ar, K, W, twoblocks, cur, qe, qs, te1, ts1, st1, 0, 0)) return false; }
    if (te2 <= ts2) continue;
    const uint32_t sl2 = d > (te2 - ts2) ? d : (te2 - ts2);
    if (sl2 >= low_b && sl2 <= upper) { if (!mp_refine_btwn_space_ha(C, r, ar, K, W, twoblocks, prev, qe, qs, te2, ts2, st2, 0, 0)) return false; }
  }
  wsync();
  {   // right end of the read
    RCluster &rh = RC[ch[0]];
    const int st = rh.strand;
    const uint32_t qs = rh.qE, qe = L;
    uint32_t ts = 0, te = 0; bool ok = true;
    if (st == 0) { ts = rh.tE; te = ts + qe - qs; }
    else { te = rh.tS; if (te > qe - qs) ts = te - (qe - qs); else { te = 0; ok = false; } }
    if (ok && qe > qs && te > ts) {
      const uint32_t slen = (qe - qs) > (te - ts) ? (qe - qs) : (te - ts);
      if (slen >= low_b && slen < upper && te + 500u < contig_len(C.ix, rh.chrom)) {
        uint32_t lrts = 0, lrlength = 0;
        if (st == 0) { lrts = 0; lrlength = 500; } else { if (ts > 500) lrts = 500; lrlength = lrts; }
        if (!mp_refine_btwn_space_ha(C, r, ar, K, W, true, rh, qe, qs, te, ts, st, lrts, lrlength)) return false;
      }
    }
  }
  wsync();
  {   // left end
    RCluster &lh = RC[ch[n - 1]];
    const uint32_t qs = 0, qe = lh.qS;
    const int st = lh.strand;
    uint32_t ts, te;
    if (st == 0) { te = lh.tS; ts = te > qe - qs ? te - (qe - qs) : 0; } else { ts = lh.tE; te = ts + (qe - qs); }
    if (qe > qs && te > ts) {
      const uint32_t slen = (qe - qs) > (te - ts) ? (qe - qs) : (te - ts);
      if (slen >= low_b && slen < upper && te + 500u < contig_len(C.ix, lh.chrom)) {
        uint32_t lrts = 0, lrlength = 0;
        if (st == 0) { if (ts > 500) lrts = 500; lrlength = lrts; } else { lrts = 0; lrlength = 500; }
        if (!mp_refine_btwn_space_ha(C, r, ar, K, W, true, lh, qe, qs, te, ts, st, lrts, lrlength)) return false;
      }
    }
  }
  wsync();
  return true;
}

// REFINEclusters (ClusterRefine.h:45-240) for cluster c of F -> refined cluster R (anchors of the LocalIndex k-mers, t chromosome-relative).  The window
// walk is stateless per genome window (a12 / a13 form, lref_kernels.cuh); the window pairs are compared one per lane (count, scan, emit).
__device__ __noinline__ bool mp_refine_cluster(const MpCtx &C, int r, const ClusterSet &F, int c, Arena &ar, RCluster &R, RSeg *node) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  const int a0 = F.off[c], nm = F.off[c + 1] - a0;
  const int Strand = F.strand[c] != 0 ? 1 : 0;
  const int first = lref_hdr_find(C.ix.hdr_pos, C.ix.n_hdr, (unsigned long long)F.tS[c] + 1ull), last = lref_hdr_find(C.ix.hdr_pos, C.ix.n_hdr, (unsigned long long)F.tE[c]);
  if (lane == 0) { R.head = R.tail = 0; R.n = 0; R.qS = 0xffffffffu; R.qE = 0; R.tS = 0xffffffffu; R.tE = 0; R.strand = Strand; R.chrom = first; R.refinespace = 0; R.freq = F.freq[c]; }
  wsync();
  if (nm == 0 || first != last) return true;
  const int chrom = first;
  const uint32_t chromOffset = (uint32_t)C.ix.hdr_pos[chrom], chromEndOffset = (uint32_t)C.ix.hdr_pos[last + 1];
  const unsigned long long mk = ar.mark();
  uint32_t *mq = ar.alloc<uint32_t>(nm), *mt = ar.alloc<uint32_t>(nm);
  MpKey *keys = ar.alloc<MpKey>((unsigned long long)next_pow2(nm));
  if (ar.overflow) return false;
  long long maxDN = -(1ll << 62), minDN = (1ll << 62);
  for (int i = lane; i < nm; i += kLanes) {
    uint32_t q = F.q[a0 + i];
    if (Strand == 1) q = L - (q + (uint32_t)O.globalK);            // SwapStrand(read, opts, clusters[ph], opts.globalK)
    const uint32_t t = F.t[a0 + i] - chromOffset;
    const long long d = (long long)t - (long long)q;
    maxDN = d > maxDN ? d : maxDN; minDN = d < minDN ? d : minDN;
    keys[i].k = ((unsigned long long)t << 32) | q; keys[i].q = 0; keys[i].idx = (uint32_t)i;
  }
  maxDN = wmax(maxDN) + 100; minDN = wmin(minDN) - 100;
  wsync();
  mp_sort_keys(keys, nm);                                          // CartesianTargetSort
  for (int i = lane; i < nm; i += kLanes) { mt[i] = (uint32_t)(keys[i].k >> 32); mq[i] = (uint32_t)keys[i].k; }
  wsync();
  uint32_t bqs = F.qS[c], bqe = F.qE[c];
  if (Strand == 1) { const uint32_t t = bqs; bqs = L - bqe; bqe = L - t; }
  const uint32_t tStart = F.tS[c], tEnd = F.tE[c];
  const uint32_t bts = tStart - chromOffset, bte = tEnd - chromOffset;
  const uint32_t W = (uint32_t)O.window;
  const uint32_t wts = chromOffset + W > tStart ? chromOffset : tStart - W;
  const uint32_t wte = tEnd + W > chromEndOffset ? chromEndOffset - 1u : tEnd + W;
  const LidxView &gl = C.ix.gl;
  const unsigned long long gend = gl.win_off[gl.n_win];
  const int ls = lref_lookup(gl.win_off, gl.n_win, 0ull, gend, wts), le = lref_lookup(gl.win_off, gl.n_win, 0ull, gend, wte);
  const LidxView &rdx = C.rd.rd[Strand];
  const int rsq = C.rd.lidx_slot ? C.rd.lidx_slot[r] : r;
  const int wf = (int)rdx.win_first[rsq], nw = (int)rdx.win_first[rsq + 1] - wf;
  const unsigned long long rbase = rdx.seq_start[rsq], rend = rbase + L;
  RsTask *tasks = 0;
  int n_tasks = 0;
  for (int pass = 0; pass < 2; pass++) {
    int nt = 0;
    if (lane == 0) {
      for (int lsi = ls; lsi <= le; lsi++) {
        if (lsi >= gl.n_win) continue;
        const unsigned long long o0 = gl.win_off[lsi], o1 = gl.win_off[lsi + 1];
        if (o0 < chromOffset || o1 < chromOffset) continue;
        const uint32_t gStart = (uint32_t)(o0 - chromOffset), gEnd = (uint32_t)(o1 - 1 - chromOffset);
        if (gStart >= gEnd) continue;
        int matchStart, matchEnd;
        { int lo = 0, len = nm;
          while (len > 0) { const int half = len >> 1; if (mt[lo + half] < gStart) { lo += half + 1; len -= half + 1; } else len = half; }
          matchStart = lo; }
        { int lo = matchStart, len = nm - matchStart;
          while (len > 0) {
            const int half = len >> 1, mid = lo + half;
            const bool less = (gEnd != mt[mid]) ? (gEnd < mt[mid]) : (0u < mq[mid]);
            if (less) len = half; else { lo = mid + 1; len -= half + 1; }
          }
          matchEnd = lo; }
        if (matchEnd == nm) matchEnd--;
        if (matchStart >= nm) continue;
        uint32_t readStart = mq[matchStart], readEnd = mq[matchEnd];
        if (readStart == readEnd) { if (lsi > ls && readStart > 0u) readStart = 0u; }     // prev_readEnd is 0 at its only read (ClusterRefine.h:152,162)
        if (lsi == ls) readStart = readStart < W ? 0u : readStart - W;
        if (lsi == le) readEnd = readEnd + W > L ? L : readEnd + W;
        if (readStart > readEnd) continue;
        const int qis = lref_lookup(rdx.win_off + wf, nw, rbase, rend, readStart);
        const int qie = lref_lookup(rdx.win_off + wf, nw, rbase, rend, readEnd < L - 1 ? readEnd : L - 1);
        for (int qi = qis; qi <= qie; qi++) {
          if (pass == 1) {
            RsTask tk; tk.lsi = lsi; tk.qw = wf + qi; tk.gStart = gStart; tk.rsStart = (uint32_t)(rdx.win_off[wf + qi] - rbase); tk.bmin = minDN; tk.bmax = maxDN;
            tasks[nt] = tk;
          }
          nt++;
        }
      }
    }
    wsync();
    nt = bcast(nt, 0);
    if (pass == 0) {
      n_tasks = nt;
      if (n_tasks == 0) break;
      tasks = ar.alloc<RsTask>(n_tasks);
      if (ar.overflow) return false;
    }
  }
  if (n_tasks == 0) { ar.release(mk); return true; }
  const long maxFreq = (long)O.localMaxFreq;
  int *cnt = ar.alloc<int>(n_tasks + 1);
  if (ar.overflow) return false;
  for (int k = lane; k < n_tasks; k += kLanes) {
    const RsTask tk = tasks[k];
    const uint32_t *q = rdx.mins + rdx.bnd[tk.qw]; const long nq = (long)(rdx.bnd[tk.qw + 1] - rdx.bnd[tk.qw]);
    const uint32_t *t = gl.mins + gl.bnd[tk.lsi]; const long nt = (long)(gl.bnd[tk.lsi + 1] - gl.bnd[tk.lsi]);
    int cc = 0;
    lt_compare(q, nq, t, nt, maxFreq, [&](long qi, long ti) {
      const uint32_t qp = (q[qi] >> 20) + tk.rsStart, tp = (t[ti] >> 20) + tk.gStart;
      const long long d = (long long)tp - (long long)qp;
      if (d >= tk.bmin && d <= tk.bmax && qp >= bqs && qp < bqe && tp >= bts && tp < bte) cc++;
    });
    cnt[k] = cc;
  }
  wsync();
  int total = 0;
  if (lane == 0) { for (int k = 0; k < n_tasks; k++) { const int cc = cnt[k]; cnt[k] = total; total += cc; } cnt[n_tasks] = total; }
  wsync();
  total = bcast(total, 0);
  if (total == 0) { ar.release(mk); return true; }
  uint32_t *rq = ar.alloc<uint32_t>(total), *rt = ar.alloc<uint32_t>(total);
  if (ar.overflow) return false;
  for (int k = lane; k < n_tasks; k += kLanes) {
    const RsTask tk = tasks[k];
    if (cnt[k + 1] == cnt[k]) continue;
    const uint32_t *q = rdx.mins + rdx.bnd[tk.qw]; const long nq = (long)(rdx.bnd[tk.qw + 1] - rdx.bnd[tk.qw]);
    const uint32_t *t = gl.mins + gl.bnd[tk.lsi]; const long nt = (long)(gl.bnd[tk.lsi + 1] - gl.bnd[tk.lsi]);
    int o = cnt[k];
    lt_compare(q, nq, t, nt, maxFreq, [&](long qi, long ti) {
      const uint32_t qp = (q[qi] >> 20) + tk.rsStart, tp = (t[ti] >> 20) + tk.gStart;
      const long long d = (long long)tp - (long long)qp;
      if (d >= tk.bmin && d <= tk.bmax && qp >= bqs && qp < bqe && tp >= bts && tp < bte) { rq[o] = qp; rt[o] = tp; o++; }
    });
  }
  wsync();
  if (Strand == 1) { for (int i = lane; i < total; i += kLanes) rq[i] = L - (rq[i] + (uint32_t)O.smallK); wsync(); }
  if (lane == 0) { rc_append(R, node, rq, rt, total); rc_set_boundaries(R, O.smallK); }
  wsync();
  return true;
}

// state of a read between the stages
struct HaState {
  HaChains H;
  ClusterSet xs;                          // extend_clusters of all chains, in chain order (t chromosome-relative)
  uint8_t *ovp;                           // Cluster::overlap per extended anchor
  int *run_off;                           // [xs.ncl + 1] same-diagonal runs of every extended cluster (MergeMatchesSameDiag)
  int *run_s, *run_e;                     // start / end of every run (anchor indices inside the cluster)
  int chain_cl0[kMaxChains + 1];          // first extended cluster of every chain (cur_cluster)
};

// everything up to MergeMatchesSameDiag (Map_highacc.h:40-650)
__device__ __noinline__ int mp_stage1_highacc(const MpCtx &C, int r, Arena &ar, HaState &S) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  ClusterSet raw; int n_raw_a = 0;
  { const int rc = mp_seed_clean(C, r, ar, raw, n_raw_a); if (rc != MP_OK) return rc; }
  unsigned long long tk = mp_clock();
  ClusterSet F;
  { const int rc = mp_fine_clusters(C, ar, raw, n_raw_a, F); if (rc != MP_OK) return rc; }
  tk = mp_tick(C, PF_STRAND_CLEAN, tk);
  if (F.ncl == 0) return MP_UNALIGNED;
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  int *keep = 0, nkeep = 0;
  { const int rc = mp_first_sdp_highacc(C, r, ar, F, S.H, keep, nkeep); if (rc != MP_OK) return rc; }
  HaChains &H = S.H;
  // ---- sparse clusters take REFINEclusters (Map_highacc.h:425-447)
  int sparse = 0;
  for (int p = 0; p < nkeep; p++) {
    const int c = keep[p];
    if (__fdiv_rn((float)fc_size(F, c), (float)(uint32_t)(F.qE[c] - F.qS[c])) <= 0.01f && L <= 50000u) sparse = 1;
  }
  RCluster *RC = ar.alloc<RCluster>(nkeep);
  RSeg *nodes = ar.alloc<RSeg>(nkeep);
  if (ar.overflow) return MP_ERR_ARENA;
  if (sparse && (C.rd.rd[0].win_off == nullptr || (C.rd.lidx_slot && C.rd.lidx_slot[r] < 0))) return MP_NEED_LIDX;     // the host indexes the read and maps it again
  if (sparse) {
    // ---- REFINEclusters with the LocalIndex k-mers (smallOpts.globalK / globalW = the index's k / w, :427-441)
    for (int p = 0; p < nkeep; p++) if (!mp_refine_cluster(C, r, F, keep[p], ar, RC[p], nodes + p)) return MP_ERR_ARENA;
    // chain entries whose refined cluster came back empty are dropped (:475-486); the links stay as they are
    int *ncp = ar.alloc<int>(kMaxChains);
    if (ar.overflow) return MP_ERR_ARENA;
    if (lane == 0) for (int h = 0; h < H.nch; h++) { int cp = 0; for (int i = 0; i < H.n[h]; i++) if (RC[H.ch[h][i]].n != 0) H.ch[h][cp++] = H.ch[h][i]; ncp[h] = cp; }
    wsync();
    for (int h = 0; h < H.nch; h++) H.n[h] = ncp[h];
    if (H.nch == 0 || H.n[0] == 0) return MP_UNALIGNED;
  } else {
  // ---- RefinedClusters = the clusters themselves, t chromosome-relative (:449-461)
  for (int p = 0; p < nkeep; p++) {
    const int c = keep[p];
    const uint32_t coff = (uint32_t)C.ix.hdr_pos[F.chrom[c]];
    const int a0 = F.off[c], n = fc_size(F, c);
    for (int m = lane; m < n; m += kLanes) F.t[a0 + m] -= coff;
    if (lane == 0) {
      RCluster &R = RC[p];
      R.head = R.tail = 0; R.n = 0; R.qS = F.qS[c]; R.qE = F.qE[c]; R.tS = F.tS[c] - coff; R.tE = F.tE[c] - coff; R.strand = F.strand[c]; R.chrom = F.chrom[c];
      R.refinespace = 0; R.freq = F.freq[c];
      rc_append(R, nodes + p, F.q + a0, F.t + a0, n);
    }
  }
  }
  wsync();
  const int K = sparse ? O.smallK : O.globalK, W = sparse ? O.smallW : O.globalW;
  for (int h = 0; h < H.nch; h++) {
    if (H.n[h] == 0) continue;
    if (!mp_refine_btwn_clusters_chain(C, r, ar, K, W, H.ch[h], H.n[h], RC)) return MP_ERR_ARENA;
  }
  tk = mp_tick(C, PF_REFINE_BTWN, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  // ---- LinearExtend_chain of every chain (:536-546), MergeMatchesSameDiag (:595)
  int n_units = 0; long long tot = 0;
  for (int h = 0; h < H.nch; h++) { n_units += H.n[h]; for (int i = 0; i < H.n[h]; i++) tot += RC[H.ch[h][i]].n; }
  {
    long long all = 0;
    for (int p = 0; p < nkeep; p++) all += RC[p].n;
    if (all == 0) return MP_UNALIGNED;           // SizeRefinedClusters == 0
  }
  ClusterSet &xs = S.xs;
  if (!mp_alloc_clusterset(xs, ar, n_units + 1, (int)tot + 1, true)) return MP_ERR_ARENA;
  S.ovp = ar.alloc<uint8_t>((unsigned long long)tot + 1);
  S.run_off = ar.alloc<int>(n_units + 2); S.run_s = ar.alloc<int>((unsigned long long)tot + 1); S.run_e = ar.alloc<int>((unsigned long long)tot + 1);
  if (ar.overflow) return MP_ERR_ARENA;
  if (lane == 0) { xs.off[0] = 0; S.run_off[0] = 0; }
  wsync();
  int xc = 0, xo = 0;
  for (int h = 0; h < H.nch; h++) {
    S.chain_cl0[h] = xc;
    const int n = H.n[h];
    if (n == 0) continue;
    const unsigned long long mk = ar.mark();
    long long ct = 0;
    for (int i = 0; i < n; i++) ct += RC[H.ch[h][i]].n;
    // the clusters of the chain, sorted (DiagonalSort / AntiDiagonalSort, LinearExtend.h:201-210), flattened in chain order
    uint32_t *cq = ar.alloc<uint32_t>((unsigned long long)ct + 1), *ctt = ar.alloc<uint32_t>((unsigned long long)ct + 1), *tmpq = ar.alloc<uint32_t>((unsigned long long)ct + 1),
             *tmpt = ar.alloc<uint32_t>((unsigned long long)ct + 1);
    unsigned long long *cl_off = ar.alloc<unsigned long long>(n + 1), *slot_off = ar.alloc<unsigned long long>(n + 1), *chrom_off = ar.alloc<unsigned long long>(n),
                       *read_off = ar.alloc<unsigned long long>(n), *cnt = ar.alloc<unsigned long long>(n + 1);
    uint32_t *unit_cl = ar.alloc<uint32_t>(n), *cl_box = ar.alloc<uint32_t>(4ull * n), *chrom_len = ar.alloc<uint32_t>(n), *read_len = ar.alloc<uint32_t>(n);
    uint8_t *unit_edge = ar.alloc<uint8_t>(n), *cl_strand = ar.alloc<uint8_t>(n);
    float *cl_freq = ar.alloc<float>(n);
    int32_t *u_overlap = ar.alloc<int32_t>(n);
    uint32_t *sq = ar.alloc<uint32_t>((unsigned long long)ct + 1), *stt = ar.alloc<uint32_t>((unsigned long long)ct + 1);
    int32_t *sl = ar.alloc<int32_t>((unsigned long long)ct + 1);
    uint8_t *so = ar.alloc<uint8_t>((unsigned long long)ct + 1);
    int *tidx = ar.alloc<int>((unsigned long long)ct + 1);
    if (ar.overflow) return MP_ERR_ARENA;
    {
      int o = 0;
      for (int i = 0; i < n; i++) {
        RCluster &R = RC[H.ch[h][i]];
        const int m = R.n;
        { int k = 0;
          for (RSeg *s = R.head; s; s = s->next) { const int sn = s->n; for (int j = lane; j < sn; j += kLanes) { tmpq[o + k + j] = s->q[j]; tmpt[o + k + j] = s->t[j]; } k += sn; } }
        if (lane == 0) {
          cl_off[i] = (unsigned long long)o; slot_off[i] = (unsigned long long)o; unit_cl[i] = (uint32_t)i; unit_edge[i] = (uint8_t)((i == 0 ? 1 : 0) | (i == n - 1 ? 2 : 0));
          cl_box[4 * i] = R.qS; cl_box[4 * i + 1] = R.qE; cl_box[4 * i + 2] = R.tS; cl_box[4 * i + 3] = R.tE; cl_strand[i] = (uint8_t)(R.strand != 0); cl_freq[i] = R.freq;
          chrom_off[i] = C.ix.hdr_pos[R.chrom]; chrom_len[i] = contig_len(C.ix, R.chrom); read_off[i] = C.rd.read_off[r]; read_len[i] = L;
        }
        wsync();
        if (m > 0) {
          const unsigned long long mk2 = ar.mark();
          MpKey *keys = ar.alloc<MpKey>((unsigned long long)next_pow2(m));
          if (ar.overflow) return MP_ERR_ARENA;
          const int st = R.strand != 0;
          for (int j = lane; j < m; j += kLanes) {
            const uint32_t q = tmpq[o + j], t = tmpt[o + j];
            keys[j].k = st == 0 ? (unsigned long long)((long long)q - (long long)t + (1ll << 33)) : (unsigned long long)(uint32_t)(q + t);
            keys[j].q = q; keys[j].idx = (uint32_t)j;
          }
          wsync();
          mp_sort_keys(keys, m);
          for (int j = lane; j < m; j += kLanes) { cq[o + j] = tmpq[o + keys[j].idx]; ctt[o + j] = tmpt[o + keys[j].idx]; }
          wsync();
          ar.release(mk2);
        }
        o += m;
      }
      if (lane == 0) { cl_off[n] = (unsigned long long)o; slot_off[n] = (unsigned long long)o; }
      wsync();
    }
    LextChainBatch b;
    b.n_units = n; b.K = K; b.skiprepetitive = 1; b.trim = 1; b.merge_dist = (long long)O.merge_dist; b.reads = C.rd.fwd; b.genome = C.ix.genome;
    b.unit_cl = unit_cl; b.unit_edge = unit_edge; b.slot_off = slot_off; b.cl_off = cl_off; b.cq = cq; b.ct = ctt; b.cl_box = cl_box; b.cl_strand = cl_strand; b.cl_freq = cl_freq;
    b.cl_chrom_off = chrom_off; b.cl_chrom_len = chrom_len; b.cl_read_off = read_off; b.cl_read_len = read_len; b.sq = sq; b.st = stt; b.sl = sl; b.so = so; b.cnt = cnt;
    b.u_overlap = u_overlap; b.lidx = 0; b.eq = 0; b.et = 0; b.elen = 0; b.eovp = 0; b.md_head = 0; b.box = 0;
    for (int u0 = 0; u0 < n; u0 += kLanes) { const int u = u0 + lane; if (u < n) lextc_walk_one(b, u); }
    wsync();
    // the extended clusters of the chain: copy, DecideCoordinates, TrimOverlappedAnchors(ExtendClusters, start), MergeMatchesSameDiag
    for (int i = 0; i < n; i++) {
      RCluster &R = RC[H.ch[h][i]];
      const int cn = (int)cnt[i], so0 = (int)slot_off[i];
      for (int j = lane; j < cn; j += kLanes) { xs.q[xo + j] = sq[so0 + j]; xs.t[xo + j] = stt[so0 + j]; xs.len[xo + j] = sl[so0 + j]; S.ovp[xo + j] = so[so0 + j]; }
      wsync();
      if (lane == 0) {
        xs.off[xc + 1] = xo + cn;
        xs.strand[xc] = -1; xs.chrom[xc] = R.chrom; xs.freq[xc] = R.freq; xs.qS[xc] = 0xffffffffu; xs.qE[xc] = 0; xs.tS[xc] = 0xffffffffu; xs.tE[xc] = 0;
        if (R.n > 0) mp_decide_coordinates(xs, xc, R.strand, R.chrom, R.freq);
      }
      wsync();
      if (cn > 0 && !mp_trim_overlapped_warp(xs.q + xo, xs.t + xo, xs.len + xo, cn, xs.strand[xc], 40, true, ar)) return MP_ERR_ARENA;
      wsync();
      {   // same-diagonal runs (MergeMatchesSameDiag): anchor j opens a run unless it continues the one of j - 1; one anchor per lane, heads compacted by ballot
        const int nr0 = S.run_off[xc];
        int nr = nr0;
        const int st = xs.strand[xc];
        const uint32_t *Q = xs.q + xo, *T = xs.t + xo; const int *Ln = xs.len + xo; const uint8_t *Ov = S.ovp + xo;
        for (int b0 = 0; b0 < cn; b0 += kLanes) {
          const int j = b0 + lane;
          bool head = j < cn;
          if (j < cn && j > 0) {
            const long long dp = st == 0 ? (long long)T[j - 1] - (long long)Q[j - 1] : (long long)Q[j - 1] + (long long)T[j - 1] + (long long)Ln[j - 1];
            const long long dc = st == 0 ? (long long)T[j] - (long long)Q[j] : (long long)Q[j] + (long long)T[j] + (long long)Ln[j];
            const uint32_t prev_qEnd = Q[j - 1] + (uint32_t)Ln[j - 1];
            const long long gd = ha_labs((long long)Q[j] - ((long long)Q[j - 1] + (long long)Ln[j - 1]));
            if (Ov[j - 1] == 0 && Ov[j] == 0 && dp == dc && prev_qEnd < Q[j] && gd <= (long long)O.merge_dist) head = false;
          }
          const unsigned mk3 = ballot(head);
          if (head) S.run_s[nr + __popc(mk3 & lanemask_lt())] = j;
          nr += __popc(mk3);
        }
        wsync();
        for (int k = nr0 + lane; k < nr; k += kLanes) S.run_e[k] = k + 1 < nr ? S.run_s[k + 1] : cn;
        if (lane == 0) S.run_off[xc + 1] = nr;
      }
      wsync();
      xo += cn; xc++;
    }
    ar.release(mk);
  }
  S.chain_cl0[H.nch] = xc;
  xs.ncl = xc;
  tk = mp_tick(C, PF_LEXT1, tk);
  return MP_OK;
}

// Cluster_SameDiag accessors over (xs, runs): run k of extended cluster c
__device__ __forceinline__ uint32_t sd_len(const HaState &S, int c, int k) {
  const int a = S.xs.off[c], s = S.run_s[S.run_off[c] + k], e = S.run_e[S.run_off[c] + k];
  const uint32_t last = S.xs.q[a + e - 1] + (uint32_t)S.xs.len[a + e - 1], first = S.xs.q[a + s];
  return last >= first ? last - first : 0u;
}
__device__ __forceinline__ uint32_t sd_qstart(const HaState &S, int c, int k) { return S.xs.q[S.xs.off[c] + S.run_s[S.run_off[c] + k]]; }
__device__ __forceinline__ uint32_t sd_qlast(const HaState &S, int c, int k) { return S.xs.q[S.xs.off[c] + S.run_e[S.run_off[c] + k] - 1]; }
__device__ __forceinline__ uint32_t sd_tstart(const HaState &S, int c, int k) {
  const int a = S.xs.off[c];
  return S.xs.strand[c] == 0 ? S.xs.t[a + S.run_s[S.run_off[c] + k]] : S.xs.t[a + S.run_e[S.run_off[c] + k] - 1];
}
__device__ __forceinline__ float sd_overlap_rate(const ClusterSet &X, int a, int b) {     // a->OverlaprateOnGenome(b)
  if (X.tE[a] <= X.tS[b] || X.tE[b] <= X.tS[a]) return 0.0f;
  const int ovp = (int)((X.tE[a] < X.tE[b] ? X.tE[a] : X.tE[b]) - (X.tS[a] > X.tS[b] ? X.tS[a] : X.tS[b]));
  const float denomA = (float)(uint32_t)(X.tE[a] - X.tS[a]);
  return __fdiv_rn((float)ovp, denomA);
}

// one chain h: SPLITChain (Cluster_SameDiag, Mapping_ultility.h:267-353) + MergeSplitchainINS (:171-264), LargestSplitChain_dist (Chain.h:974),
// LocalRefineAlignment (LocalRefineAlignment.h:553-768) -> segments.  (Included from mp_map.cuh after MapOut / mp_segbuild_alloc / mp_emit_segments.)
// Returns 0 ok, < 0 error.
__device__ __noinline__ int mp_map_chain_highacc(const MpCtx &C, int r, Arena &ar, HaState &S, int h, const MapOut &out, int &nseg_out, int &seg0_out) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  const ClusterSet &X = S.xs;
  const int n = S.H.n[h], cl0 = S.chain_cl0[h];
  nseg_out = 0; seg0_out = 0;
  unsigned long long tk = mp_clock();
  // ---- split chains as lists of pieces [pa, pb) of the chain's clusters (MergeSplitchainINS appends lists)
  int *pa = ar.alloc<int>(n + 1), *pb = ar.alloc<int>(n + 1), *pnext = ar.alloc<int>(n + 1), *head = ar.alloc<int>(n + 1), *tail = ar.alloc<int>(n + 1), *size = ar.alloc<int>(n + 1);
  int *chrom = ar.alloc<int>(n + 1), *cur_ind = ar.alloc<int>(n + 1), *order = ar.alloc<int>(n + 1), *res = ar.alloc<int>(4);
  uint32_t *QS = ar.alloc<uint32_t>(n + 1), *QE = ar.alloc<uint32_t>(n + 1), *TS = ar.alloc<uint32_t>(n + 1), *TE = ar.alloc<uint32_t>(n + 1);
  uint8_t *type = ar.alloc<uint8_t>(n + 1), *pstrand = ar.alloc<uint8_t>(n + 1), *keepf = ar.alloc<uint8_t>(n + 1);
  if (ar.overflow) return -MP_ERR_ARENA;
  if (lane == 0) {
    const uint8_t *link = S.H.link[h]; const int nl = S.H.nl[h];
    int ns = 0, a = 0;
    auto close = [&](int b_, char ty, int st, int ci) { pa[ns] = a; pb[ns] = b_; pnext[ns] = -1; head[ns] = ns; tail[ns] = ns; size[ns] = b_ - a; type[ns] = (uint8_t)ty; pstrand[ns] = (uint8_t)st; chrom[ns] = ci; ns++; a = b_; };
    for (int im = 0; im + 1 < n; im++) {
      const int cur = cl0 + im + 1, prev = cl0 + im;
      const int lk = im < nl ? link[im] : 0;
      bool rep_map = false;
      if (((lk == 1 && X.strand[cur] == 0 && X.strand[prev] == 0) || (lk == 0 && X.strand[cur] == 1 && X.strand[prev] == 1)) &&
          sd_overlap_rate(X, prev, cur) >= 0.6f && sd_overlap_rate(X, cur, prev) >= 0.6f) rep_map = true;
      if (X.tS[cur] > X.tE[prev] + (uint32_t)O.splitdist || X.tE[cur] + (uint32_t)O.splitdist < X.tS[prev] || X.chrom[cur] != X.chrom[prev]) close(im + 1, 'T', X.strand[prev] != 0, X.chrom[cur]);
      else if (rep_map) close(im + 1, 'D', X.strand[prev] != 0, X.chrom[cur]);
      else if ((X.strand[cur] == 0 && X.strand[prev] == 1) || (X.strand[cur] == 1 && X.strand[prev] == 0)) close(im + 1, 'I', X.strand[prev] != 0, X.chrom[cur]);
    }
    close(n, 'N', X.strand[cl0 + n - 1] != 0, -2);      // (chromIndex of the last split chain is never assigned in the reference: compares unequal here)
    for (int m = 0; m < ns; m++) {
      QS[m] = X.qS[cl0 + pa[m]]; QE[m] = X.qE[cl0 + pa[m]]; TS[m] = X.tS[cl0 + pa[m]]; TE[m] = X.tE[cl0 + pa[m]];
      for (int v = pa[m] + 1; v < pb[m]; v++) {
        QS[m] = X.qS[cl0 + v] < QS[m] ? X.qS[cl0 + v] : QS[m]; QE[m] = X.qE[cl0 + v] > QE[m] ? X.qE[cl0 + v] : QE[m];
        TS[m] = X.tS[cl0 + v] < TS[m] ? X.tS[cl0 + v] : TS[m]; TE[m] = X.tE[cl0 + v] > TE[m] ? X.tE[cl0 + v] : TE[m];
      }
      cur_ind[m] = m; keepf[m] = 1;
    }
    // MergeSplitchainINS
    if (ns >= 3) {
      int im = 0;
      while (im <= ns - 3) {
        const int c = cur_ind[im];
        if (type[c] != 'T') { im++; continue; }
        int nn = cur_ind[im + 2];
        while (nn < ns) {
          const long long tdist = TS[c] > TE[nn] ? (long long)TS[c] - (long long)TE[nn] : (long long)TE[nn] - (long long)TS[c];
          if (tdist > 1500) { nn++; continue; }
          if (pstrand[c] != pstrand[nn]) { nn++; continue; }
          if (chrom[c] != chrom[nn] || chrom[nn] == -2) { nn++; continue; }
          pnext[tail[c]] = head[nn]; tail[c] = tail[nn]; size[c] += size[nn];
          QS[c] = QS[nn] < QS[c] ? QS[nn] : QS[c]; TS[c] = TS[nn] < TS[c] ? TS[nn] : TS[c]; QE[c] = QE[nn] > QE[c] ? QE[nn] : QE[c]; TE[c] = TE[nn] > TE[c] ? TE[nn] : TE[c];
          type[c] = type[nn];
          cur_ind[nn] = cur_ind[c]; keepf[nn] = 0;
          break;
        }
        im = nn;
      }
    }
    int no = 0;
    for (int m = 0; m < ns; m++) if (keepf[m]) order[no++] = m;
    // LargestSplitChain_dist
    int maxi = 0, maxi_d = no > 0 ? (QE[order[0]] > QS[order[0]] ? (int)(QE[order[0]] - QS[order[0]]) : 0) : 0;
    for (int mi = 1; mi < no; mi++) { const int m = order[mi]; const int d = QE[m] > QS[m] ? (int)(QE[m] - QS[m]) : 0; if (d > maxi_d) { maxi = mi; maxi_d = d; } }
    res[0] = no; res[1] = maxi;
  }
  wsync();
  const int nsp = res[0], LSC = res[1];
  tk = mp_tick(C, PF_SPLIT, tk);
  // ---- the second SparseDP of every split chain over its same-diagonal runs, the chain filters, SwitchToOriginalAnchors
  UChain *uc = ar.alloc<UChain>(nsp > 0 ? nsp : 1);
  if (ar.overflow) return -MP_ERR_ARENA;
  for (int st = 0; st < nsp; st++) {
    const int m = order[st];
    // runs of the split chain's clusters, concatenated in split-chain order
    int nclu = 0, nrun = 0, nanch = 0;
    for (int p = head[m]; p >= 0; p = pnext[p]) for (int v = pa[p]; v < pb[p]; v++) { nclu++; nrun += S.run_off[cl0 + v + 1] - S.run_off[cl0 + v]; nanch += X.off[cl0 + v + 1] - X.off[cl0 + v]; }
    UChain u; u.n = 0; u.nlink = 0; u.FirstSDPValue = S.H.value[h]; u.NumOfAnchors0 = S.H.n0[h]; u.NumOfAnchors1 = 0; u.QStart = u.QEnd = u.TStart = u.TEnd = 0;
    u.idx = ar.alloc<uint32_t>(nanch + 1); u.cl = ar.alloc<int>(nanch + 1); u.link = ar.alloc<uint8_t>(nanch + 1);
    const unsigned long long mk = ar.mark();
    int *clv = ar.alloc<int>(nclu + 1), *cl_off = ar.alloc<int>(nclu + 2);
    uint8_t *cl_strand = ar.alloc<uint8_t>(nclu + 1);
    uint32_t *rq = ar.alloc<uint32_t>(nrun + 1), *rt = ar.alloc<uint32_t>(nrun + 1), *rqf = ar.alloc<uint32_t>(nrun + 1);
    int32_t *rl = ar.alloc<int32_t>(nrun + 1);
    uint32_t *chain = ar.alloc<uint32_t>(nrun + 1);
    uint8_t *lk = ar.alloc<uint8_t>(nrun + 1);
    float *valp = ar.alloc<float>(1);
    int *np_ = ar.alloc<int>(2);
    if (ar.overflow) return -MP_ERR_ARENA;
    if (lane == 0) {
      int ci = 0, o = 0;
      for (int p = head[m]; p >= 0; p = pnext[p]) for (int v = pa[p]; v < pb[p]; v++) {
        const int c = cl0 + v, nr = S.run_off[c + 1] - S.run_off[c];
        clv[ci] = v; cl_off[ci] = o; cl_strand[ci] = (uint8_t)(X.strand[c] != 0);
        for (int k = 0; k < nr; k++) { rq[o] = sd_qstart(S, c, k); rt[o] = sd_tstart(S, c, k); rl[o] = (int32_t)sd_len(S, c, k); rqf[o] = X.strand[c] != 0 ? sd_qlast(S, c, k) : rq[o]; o++; }
        ci++;
      }
      cl_off[ci] = o; *valp = 0.0f;
    }
    wsync();
    SdpAnchors A; A.q = rq; A.t = rt; A.len = rl; A.nfrag = nrun; A.cl_off = cl_off; A.cl_strand = cl_strand; A.ncl = nclu; A.qe = 0; A.te = 0; A.fstrand = 0; A.fval = 0; A.fn0 = 0;
    unsigned long long t2 = mp_clock();
    int nch = sdp_samediag_chain(A, O.second_anchorbonus, *C.pwl, ar, chain, lk, valp);
    t2 = mp_tick(C, PF_SDP2, t2);
    if (nch < 0) return -MP_ERR_ARENA;
    wsync();
    // RemoveSmallPairedIndels, RemovePairedIndels(refineEnds = false), RemoveSpuriousAnchors on the FinalChain (qEnd(i) of a run is its LAST anchor's
    // start + the run length, Clustering.h:381-383: the reverse-strand gaps see that value)
    for (int pass = 0; pass < 3 && nch >= 2; pass++) {
      const int mode = pass == 0 ? 0 : (pass == 1 ? 2 : 4);
      const unsigned long long mk2 = ar.mark();
      uint32_t *fq = ar.alloc<uint32_t>(nch), *ft = ar.alloc<uint32_t>(nch), *fl = ar.alloc<uint32_t>(nch);
      uint8_t *fs = ar.alloc<uint8_t>(nch), *keep = ar.alloc<uint8_t>(nch);
      int32_t *sv = ar.alloc<int32_t>(nch), *svp = ar.alloc<int32_t>(nch), *svg = ar.alloc<int32_t>(nch);
      unsigned long long *off = ar.alloc<unsigned long long>(2);
      if (ar.overflow) return -MP_ERR_ARENA;
      for (int i = lane; i < nch; i += kLanes) {
        const int g = (int)chain[i];
        int ci = upper_bound_idx(cl_off, nclu + 1, g) - 1;
        fq[i] = rqf[g]; ft[i] = rt[g]; fl[i] = (uint32_t)rl[g]; fs[i] = cl_strand[ci];
      }
      if (lane == 0) { off[0] = 0; off[1] = (unsigned long long)nch; }
      wsync();
      int mkeep = 0;
      if (lane == 0) {
        ChainfBatch b; b.n_chains = 1; b.mode = mode; b.off = off; b.q = fq; b.t = ft; b.len = fl; b.strand = fs; b.keep = keep; b.sv = sv; b.svpos = svp; b.svg = svg;
        chainf_one(b, 0);
        for (int i = 0; i < nch; i++) if (keep[i]) chain[mkeep++] = chain[i];
      }
      wsync();
      nch = bcast(mkeep, 0);
      ar.release(mk2);
    }
    // SwitchToOriginalAnchors (LocalRefineAlignment.h:186-198)
    if (lane == 0) {
      int o = 0;
      for (int i = 0; i < nch; i++) {
        const int g = (int)chain[i];
        const int ci = upper_bound_idx(cl_off, nclu + 1, g) - 1, k = g - cl_off[ci], c = cl0 + clv[ci];
        const int s = S.run_s[S.run_off[c] + k], e = S.run_e[S.run_off[c] + k];
        for (int j = e - 1; j >= s; j--) { u.idx[o] = (uint32_t)j; u.cl[o] = c; o++; }
      }
      np_[0] = o;
    }
    wsync();
    u.n = np_[0]; u.nlink = 0; u.NumOfAnchors1 = u.n;
    ar.release(mk);
    if (lane == 0) uc[st] = u;
    wsync();
  }
  tk = mp_clock();
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  SegBuild B;
  if (!mp_segbuild_alloc(B, ar, uc, nsp, L)) return -MP_ERR_ARENA;
  if (!mp_local_refine_alignment(C, r, ar, B, X, uc, nsp, LSC, 1)) return ar.overflow ? -MP_ERR_ARENA : -MP_ERR_CAP;
  wsync();
  tk = mp_tick(C, PF_LOCAL_REFINE, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  const int rc_emit = mp_emit_segments(r, h, B, out, nseg_out, seg0_out);
  tk = mp_tick(C, PF_OUTPUT, tk);
  return rc_emit;
}

}  // namespace mp
}  // namespace lra
