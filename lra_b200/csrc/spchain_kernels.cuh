// a11 (low-accuracy pipeline): SPLITChain on an UltimateChain (reference Mapping_ultility.h:380-437) with push_new (:355-378) and
// SplitChain::CHROMIndex (Chain.h:388-396), MergeSplitchainINS (Mapping_ultility.h:163-264), the per-strand reversal, and
// RemoveSpuriousSplitChain (Map_lowacc.h:38-66), batched over chains (Map_lowacc.h:261-262).
// The splitter's state runs from anchor to anchor and from piece to piece (which pieces exist decides which ones merge, and the
// vector<bool> of piece links is resized, never re-derived), so one thread replays one chain literally; a batch has 10^4..10^5 chains.
// A piece is a contiguous range [pa, pb) of the chain (the reference's `onec` only ever grows by the next anchor); a split chain is a list of
// pieces (MergeSplitchainINS appends the piece two or more places further on), so nothing is copied until the result is written out:
// anchor indices (reversed on the forward strand), link bits (the chain's own inside a piece, 0 at a junction), the ClusterIndex run-length list.
#pragma once
#include "lra_common.cuh"

namespace lra {

struct SpChainBatch {
  int n_chains;
  int splitdist, bypass;
  const unsigned long long *c_off;        // [n_chains + 1] anchors of each chain
  const uint32_t *q, *t;                  // anchors in chain order (t: global genome coordinate)
  const int32_t *len;
  const uint8_t *strand;                  // strand of the cluster the anchor comes from
  const int32_t *cnum;                    // UltimateChain::ClusterNum
  const uint8_t *link;                    // link[c_off[k] + i] between anchors i and i + 1 of chain k
  const unsigned long long *hdr_pos;
  int n_hdr;
  // scratch, one int per anchor each (slot = anchor offset of the chain)
  int32_t *pa, *pb, *pnext, *tail, *size, *chrom, *cur_ind, *ord;
  uint32_t *QS, *QE, *TS, *TE;
  uint8_t *type, *pstrand, *keep, *SL;
  // out, in slot layout: piece s of chain k at c_off[k] + k + s (offset arrays) / c_off[k] + s (per-piece fields); entries at c_off[k] + ...
  int32_t *n_sp, *n_link;                 // [n_chains]
  int32_t *sp_off, *ci_off;               // [N + n_chains + 1]
  int32_t *sptc, *ci;                     // [N]
  uint8_t *sp_lk;                         // [N]
  uint32_t *sp_box;                       // [N * 4]
  int32_t *sp_chrom;                      // [N]
  uint8_t *sp_type, *sp_strand, *sp_link; // [N]
};

__device__ __forceinline__ int spc_hdr_find(const unsigned long long *pos, int n, unsigned long long query) {   // Header::Find
  if (n > 0 && query == pos[0]) return 0;
  int lo = 0, len = n;
  while (len > 0) { const int half = len >> 1; if (pos[lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
  if (lo < n && query == pos[lo]) return lo;
  return lo - 1;
}

__device__ __noinline__ void spchain_one(const SpChainBatch &b, const int k) {
  const unsigned long long a0 = b.c_off[k];
  const int n = (int)(b.c_off[k + 1] - a0);
  const uint32_t *q = b.q + a0, *t = b.t + a0;
  const int32_t *len = b.len + a0, *cnum = b.cnum + a0;
  const uint8_t *strand = b.strand + a0, *link = b.link + a0;
  int32_t *pa = b.pa + a0, *pb = b.pb + a0, *pnext = b.pnext + a0, *tail = b.tail + a0, *size = b.size + a0, *chrom = b.chrom + a0, *cur_ind = b.cur_ind + a0,
          *ord = b.ord + a0;
  uint32_t *QS = b.QS + a0, *QE = b.QE + a0, *TS = b.TS + a0, *TE = b.TE + a0;
  uint8_t *type = b.type + a0, *pstrand = b.pstrand + a0, *keep = b.keep + a0, *SL = b.SL + a0;
  int ns = 0, nsl = 0;
  if (n == 0) { b.n_sp[k] = 0; b.n_link[k] = 0; return; }
  auto qend = [&](int i) { return q[i] + (uint32_t)len[i]; };
  auto tend = [&](int i) { return t[i] + (uint32_t)len[i]; };
  auto diag = [&](int i) { return strand[i] == 1 ? (long long)qend(i) + (long long)t[i] : (long long)t[i] - (long long)q[i]; };
  // push_new for the piece [ps, pe): true if it is kept
  auto push_new = [&](int ps, int pe) {
    const int st = strand[ps];
    const uint32_t qs = q[pe - 1], qe = qend(ps);
    uint32_t ts, te;
    if (st == 0) { ts = t[pe - 1]; te = tend(ps); } else { ts = t[ps]; te = tend(pe - 1); }
    const int f = spc_hdr_find(b.hdr_pos, b.n_hdr, (unsigned long long)(uint32_t)(ts + 1u)), l = spc_hdr_find(b.hdr_pos, b.n_hdr, (unsigned long long)te);
    if (f != l) return false;
    pa[ns] = ps; pb[ns] = pe; pnext[ns] = -1; tail[ns] = ns; size[ns] = pe - ps; chrom[ns] = f; type[ns] = 'N'; pstrand[ns] = (uint8_t)st;
    QS[ns] = qs; QE[ns] = qe; TS[ns] = ts; TE[ns] = te;
    ns++;
    return true;
  };
  {
    int ps = 0, cur = 0;
    for (int im = 0; im < n - 1; im++) {
      cur = im + 1;
      const int prev = im;
      const int qdist = (int)(q[prev] - qend(cur));
      const int tdist = (t[prev] > tend(cur)) ? (int)(t[prev] - tend(cur)) : (int)(tend(cur) - t[prev]);
      const int dist = qdist < tdist ? qdist : tdist;
      long long dd = diag(cur) - diag(prev);
      if (dd < 0) dd = -dd;
      if (strand[cur] == strand[prev] && dist >= 1000 && (double)dd <= ceil(__dmul_rn(0.15, (double)dist))) {
        if (push_new(ps, cur)) { SL[nsl++] = 0; type[ns - 1] = 'N'; }
        ps = cur;
      } else if (t[cur] > tend(prev) + (uint32_t)b.splitdist || tend(cur) + (uint32_t)b.splitdist < t[prev]) {
        if (push_new(ps, cur)) { SL[nsl++] = 0; type[ns - 1] = 'T'; }
        ps = cur;
      } else if (strand[cur] != strand[prev]) {
        if (push_new(ps, cur)) { type[ns - 1] = 'I'; SL[nsl++] = 1; }
        ps = cur;
      }
    }
    push_new(ps, n);
  }
  // MergeSplitchainINS
  for (int i = 0; i < ns; i++) ord[i] = i;
  int nk = ns;                                   // split chains alive, in ord[]
  if (ns >= 3) {
    for (int i = 0; i < ns; i++) { cur_ind[i] = i; keep[i] = 1; }
    bool change = false;
    int im = 0;
    while (im <= ns - 3) {
      const int c = cur_ind[im];
      if (type[c] != 'T') { im++; continue; }
      int nn = cur_ind[im + 2];
      while (nn < ns) {
        const long long tdist = (TS[c] > TE[nn]) ? ((long long)TS[c] - (long long)TE[nn]) : ((long long)TE[nn] - (long long)TS[c]);
        if (tdist > 1500 || pstrand[c] != pstrand[nn] || chrom[c] != chrom[nn]) { nn++; continue; }
        change = true;
        pnext[tail[c]] = nn; tail[c] = tail[nn]; size[c] += size[nn];
        QS[c] = QS[nn] < QS[c] ? QS[nn] : QS[c]; TS[c] = TS[nn] < TS[c] ? TS[nn] : TS[c];
        QE[c] = QE[nn] > QE[c] ? QE[nn] : QE[c]; TE[c] = TE[nn] > TE[c] ? TE[nn] : TE[c];
        type[c] = type[nn];
        cur_ind[nn] = cur_ind[c];
        keep[nn] = 0;
        break;
      }
      im = nn;
    }
    if (change) {
      int r = 0;
      for (int s = 0; s < ns; s++) if (keep[s]) ord[r++] = s;
      nk = r;
      for (int i = nsl; i < r - 1; i++) SL[i] = 0;
      nsl = r - 1;
      if (b.bypass) for (int i = 1; i < nk; i++) SL[i - 1] = type[ord[i]] == 'I' ? 1 : 0;
    }
  }
  // RemoveSpuriousSplitChain
  {
    int total = 0;
    for (int i = 0; i < nk; i++) total += size[ord[i]];
    int filter = (int)floorf(__fmul_rn(0.02f, (float)total)); if (filter < 2) filter = 2;
    int filter2 = (int)floorf(__fmul_rn(0.03f, (float)total)); if (filter2 < 2) filter2 = 2;
    const int f1 = filter < 2 ? filter : 2, f2 = filter2 < 4 ? filter2 : 4;
    int c = 0;
    for (int i = 0; i < nk; i++) {
      bool rem = size[ord[i]] < f1;
      if (i > 0 && SL[i - 1] == 1 && size[ord[i]] < f2) rem = true;
      cur_ind[i] = rem ? 1 : 0;              // cur_ind re-used as the reference's remove[] vector: all decisions first, then the compaction
    }
    for (int i = 0; i < nk; i++) {
      if (!cur_ind[i]) {
        ord[c] = ord[i];
        if (c > 1) SL[c - 1] = SL[i - 1];
        c++;
      }
    }
    nk = c;
    if (c > 1) { for (int i = nsl; i < c - 1; i++) SL[i] = 0; nsl = c - 1; } else nsl = 0;
  }
  // write out
  int32_t *sp_off = b.sp_off + a0 + k, *ci_off = b.ci_off + a0 + k;
  int32_t *sptc = b.sptc + a0, *ci = b.ci + a0;
  uint8_t *sp_lk = b.sp_lk + a0;
  int o = 0, oc = 0;
  for (int i = 0; i < nk; i++) {
    const int s = ord[i];
    sp_off[i] = o; ci_off[i] = oc;
    const int sz = size[s];
    const bool rev = pstrand[s] == 0;
    int pos = 0;                                  // position in the un-reversed concatenation
    int prevc = -1; bool first_piece = true;
    for (int p = s; p != -1; p = pnext[p]) {
      for (int x = pa[p]; x < pb[p]; x++, pos++) {
        sptc[o + (rev ? sz - 1 - pos : pos)] = x;
        if (pos + 1 < sz) {                       // link `pos` of the un-reversed list: the chain's own inside a piece, 0 at a junction
          const uint8_t lv = (x + 1 < pb[p]) ? link[x] : 0;
          sp_lk[o + (rev ? sz - 2 - pos : pos)] = lv;
        }
        if (first_piece || b.bypass) {
          const int cn = cnum[x];
          if (oc == ci_off[i] || cn != prevc) { ci[oc++] = cn; prevc = cn; }
        }
      }
      first_piece = false;
    }
    sp_lk[o + sz - 1] = 0;
    o += sz;
    uint32_t *bx = b.sp_box + 4 * (a0 + i);
    bx[0] = QS[s]; bx[1] = QE[s]; bx[2] = TS[s]; bx[3] = TE[s];
    b.sp_chrom[a0 + i] = chrom[s]; b.sp_type[a0 + i] = type[s]; b.sp_strand[a0 + i] = pstrand[s];
  }
  sp_off[nk] = o; ci_off[nk] = oc;
  for (int i = 0; i < nsl; i++) b.sp_link[a0 + i] = SL[i];
  b.n_sp[k] = nk; b.n_link[k] = nsl;
}

__global__ void __launch_bounds__(64) spchain_kernel(SpChainBatch b) {
  const int k = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (k >= b.n_chains) return;
  spchain_one(b, k);
}

}  // namespace lra
