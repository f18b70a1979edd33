// The three SparseDP drivers the low-accuracy pipeline calls (reference SparseDP.h:2139-2282 pure matches + DecidePrimaryChains
// :1658-1765; :2287-2440 one cluster; SparseDP_Forward.h:312-490 forward only), on top of mp_sdp.cuh, and the batched
// stand-alone kernel behind lra_b200_sdp_batch (one problem per warp).
#pragma once
#include "mp_sdp.cuh"

namespace lra {
namespace mp {

struct SdpChain {            // one output chain
  uint32_t *chain; uint8_t *link; int n;
  float value;
  uint32_t QStart, QEnd, TStart, TEnd;
};

// UltimateChain::OverlapsOnT (Chain.h:262-276) of chain 0 against [tS,tE)
__device__ __forceinline__ bool sdp_overlaps_on_t(uint32_t TStart, uint32_t TEnd, uint32_t tS, uint32_t tE, float rate) {
  int ovp = 0;
  if (tS >= TStart && tS < TEnd) ovp = (int)((tE < TEnd ? tE : TEnd) - tS);
  else if (tE > TStart && tE <= TEnd) ovp = (int)(tE - (tS > TStart ? tS : TStart));
  else if (tS < TStart && tE > TEnd) ovp = (int)(TEnd - TStart);
  const float denomA = (float)(uint32_t)(TEnd - TStart);
  return __fdiv_rn((float)ovp, denomA) <= rate;
}

// mode 0.  chains[c].chain / link must have room for A.nfrag entries each.  Returns the number of chains (<= NumAln), -1 on arena overflow.
__device__ __noinline__ int sdp_pure_matches(const SdpAnchors &A, float rate, float alnthres, int NumAln, int read_len, const Pwl &P, Arena &ar,
                                       SdpChain *chains, int *cl_of_frag /* optional [nfrag] */) {
  const unsigned long long mk = ar.mark();
  SdpWork W;
  if (!sdp_build(W, A, 0, 0, rate, 0, ar)) { ar.release(mk); return -1; }
  const int n = W.nfrag;
  int *order = ar.alloc<int>(n > 0 ? n : 1);
  uint8_t *used = ar.alloc<uint8_t>(n > 0 ? n : 1);
  int *nch_p = ar.alloc<int>(1);
  if (ar.overflow || !sdp_open_dyn(W, ar)) { ar.release(mk); return -1; }
  { const unsigned long long tp_ = ar.now(); sdp_process(W, A, 0, 0, rate, 0, P); ar.tick(18, tp_); }
  { const unsigned long long e = W.dyn.base_off + *W.dyn.top; if (e > ar.peak) ar.peak = e; }
  if (*W.dyn.err) { ar.release(mk); return -1; }
  for (int i = lane_id(); i < n; i += kLanes) { order[i] = i; used[i] = 0; if (cl_of_frag) cl_of_frag[i] = W.val[i].cl; }
  wsync();
  if (lane_id() == 0) {
    int nch = 0;
    if (n > 0) {
      const SdpVal *val = W.val;
      std_sort_replay(order, n, [val](int a, int b) { return val[a].val > val[b].val; });   // Fragment_valueOrder (Fragment_Info.h:65-99)
      const float thres = __fmul_rn(alnthres, val[order[0]].val);
      int fv = 0;
      while (nch < NumAln && fv < n && val[order[fv]].val >= thres) {
        SdpChain &c = chains[nch];
        const int len = sdp_traceback_used(W, (uint32_t)order[fv], c.chain, c.link, used);
        if (len != 0) {
          int f = (int)c.chain[0], l = (int)c.chain[len - 1];
          uint32_t QEnd = A.q[f] + (uint32_t)A.len[f], QStart = A.q[l], TEnd = A.t[f] + (uint32_t)A.len[f], TStart = A.t[l];
          for (int x = 0; x < len; x++) {
            f = (int)c.chain[x];
            const uint32_t qe = A.q[f] + (uint32_t)A.len[f], te = A.t[f] + (uint32_t)A.len[f];
            QEnd = qe > QEnd ? qe : QEnd; QStart = A.q[f] < QStart ? A.q[f] : QStart;
            TStart = A.t[f] < TStart ? A.t[f] : TStart; TEnd = te < TEnd ? te : TEnd;       // sic: min (SparseDP.h:1693)
          }
          if (len >= 3 && QEnd > QStart && (double)__fdiv_rn((float)(uint32_t)(QEnd - QStart), (float)read_len) > 0.005 && QEnd - QStart >= 200) {
            bool take = false;
            if (nch == 0) take = true;
            else if (nch < NumAln) take = sdp_overlaps_on_t(chains[0].TStart, chains[0].TEnd, TStart, TEnd, 0.05f);
            if (take) { c.n = len; c.value = val[order[fv]].val; c.QStart = QStart; c.QEnd = QEnd; c.TStart = TStart; c.TEnd = TEnd; nch++; }
            else for (int x = 0; x < len; x++) { /* the anchors stay marked used, as in the reference */ }
          } else break;
        }
        fv++;
      }
    }
    *nch_p = nch;
  }
  wsync();
  const int nch = *nch_p;
  ar.release(mk);
  return nch;
}

// mode 1 (one cluster).  Returns the chain length (0 for an empty cluster), -1 on arena overflow; chain holds cluster-local indices.
__device__ __noinline__ int sdp_one_cluster(const SdpAnchors &A, int cl, float rate, const Pwl &P, Arena &ar, uint32_t *chain, uint8_t *link, float *value) {
  const int f0 = A.cl_off[cl], nf = A.cl_off[cl + 1] - f0;
  if (nf == 0) return 0;
  const unsigned long long mk = ar.mark();
  SdpWork W;
  if (!sdp_build(W, A, 1, cl, rate, 0, ar)) { ar.release(mk); return -1; }
  int *res = ar.alloc<int>(2);
  if (ar.overflow || !sdp_open_dyn(W, ar)) { ar.release(mk); return -1; }
  { const unsigned long long tp_ = ar.now(); sdp_process(W, A, f0, 1, rate, 0, P); ar.tick(18, tp_); }
  { const unsigned long long e = W.dyn.base_off + *W.dyn.top; if (e > ar.peak) ar.peak = e; }
  if (*W.dyn.err) { ar.release(mk); return -1; }
  if (lane_id() == 0) {
    float mx = 0.0f; uint32_t pos = 0;
    for (int l = 0; l < nf; l++) if (W.val[l].val > mx) { mx = W.val[l].val; pos = (uint32_t)l; }
    *value = mx;
    res[0] = sdp_traceback(W, pos, chain, link);
  }
  wsync();
  const int n = res[0];
  ar.release(mk);
  return n;
}

// mode 4 (SparseDP.h:1956 + DecidePrimaryChains :1586-1652): split clusters as fragments; up to NumAln chains of Primary_chains[0].  chains[c].value is
// the chain's value, QStart.. its box; n0[c] receives NumOfAnchors0 (sum of the fragments' NumofAnchors0).  Returns the number of chains.
__device__ __noinline__ int sdp_split_clusters(const SdpAnchors &A, float rate, float alnthres, int globalK, int NumAln, int read_len, const Pwl &P, Arena &ar,
                                               SdpChain *chains, int *n0) {
  if (A.nfrag == 0) return 0;
  const unsigned long long mk = ar.mark();
  SdpWork W;
  if (!sdp_build(W, A, 4, 0, rate, 0, ar)) { ar.release(mk); return -1; }
  const int n = W.nfrag;
  int *order = ar.alloc<int>(n);
  uint8_t *used = ar.alloc<uint8_t>(n);
  int *nch_p = ar.alloc<int>(1);
  if (ar.overflow || !sdp_open_dyn(W, ar)) { ar.release(mk); return -1; }
  { const unsigned long long tp_ = ar.now(); sdp_process(W, A, 0, 4, rate, 0, P); ar.tick(18, tp_); }
  { const unsigned long long e = W.dyn.base_off + *W.dyn.top; if (e > ar.peak) ar.peak = e; }
  if (*W.dyn.err) { ar.release(mk); return -1; }
  for (int i = lane_id(); i < n; i += kLanes) { order[i] = i; used[i] = 0; }
  wsync();
  if (lane_id() == 0) {
    int nch = 0;
    const SdpVal *val = W.val;
    std_sort_replay(order, n, [val](int a, int b) { return val[a].val > val[b].val; });   // Fragment_valueOrder (Fragment_Info.h:65-99)
    const float top = val[order[0]].val;
    const float t1 = __fmul_rn(alnthres, top), t2 = __fsub_rn(top, (float)(130 * globalK));
    const float thres = t1 > t2 ? t1 : t2;
    int fv = 0;
    while (fv < n && val[order[fv]].val >= thres) {
      if (nch >= NumAln && nch > 0) {
        // the reference traces the chain before it finds Primary_chains[0] full and breaks: the anchors of that chain stay marked, nothing reads them
        break;
      }
      SdpChain &c = chains[nch];
      const int len = sdp_traceback_used(W, (uint32_t)order[fv], c.chain, c.link, used);
      if (len != 0) {
        int f = (int)c.chain[0], l = (int)c.chain[len - 1];
        uint32_t QEnd = A.qe[f], TEnd = A.te[f], QStart = A.q[l], TStart = A.t[l];
        for (int x = 0; x < len; x++) {
          f = (int)c.chain[x];
          QEnd = A.qe[f] > QEnd ? A.qe[f] : QEnd; TEnd = A.te[f] > TEnd ? A.te[f] : TEnd;
          QStart = A.q[f] < QStart ? A.q[f] : QStart; TStart = A.t[f] < TStart ? A.t[f] : TStart;
        }
        if ((double)__fdiv_rn((float)(uint32_t)(QEnd - QStart), (float)read_len) > 0.005) {
          int na = 0;
          for (int x = 0; x < len; x++) na += A.fn0[c.chain[x]];
          c.n = len; c.value = val[order[fv]].val; c.QStart = QStart; c.QEnd = QEnd; c.TStart = TStart; c.TEnd = TEnd; n0[nch] = na; nch++;
        } else break;
      }
      fv++;
    }
    *nch_p = nch;
  }
  wsync();
  const int nch = *nch_p;
  ar.release(mk);
  return nch;
}

// mode 3 (SparseDP.h:1766): the same-diagonal anchors of the clusters of one split chain (A.cl_off / A.cl_strand in split-chain order), value =
// length x second_anchorbonus, one chain from the best anchor.  Returns the chain length; chain holds indices into the concatenation (the
// FinalChain's (ClusterIndex, chain) pair is (cluster of the index, index - cl_off[cluster])).
__device__ __noinline__ int sdp_samediag_chain(const SdpAnchors &A, float rate, const Pwl &P, Arena &ar, uint32_t *chain, uint8_t *link, float *value) {
  if (A.nfrag == 0) return 0;
  const unsigned long long mk = ar.mark();
  SdpWork W;
  if (!sdp_build(W, A, 3, 0, rate, 0, ar)) { ar.release(mk); return -1; }
  int *res = ar.alloc<int>(2);
  if (ar.overflow || !sdp_open_dyn(W, ar)) { ar.release(mk); return -1; }
  { const unsigned long long tp_ = ar.now(); sdp_process(W, A, 0, 3, rate, 0, P); ar.tick(18, tp_); }
  { const unsigned long long e = W.dyn.base_off + *W.dyn.top; if (e > ar.peak) ar.peak = e; }
  if (*W.dyn.err) { ar.release(mk); return -1; }
  if (lane_id() == 0) {
    float mx = 0.0f; uint32_t pos = 0;
    for (int l = 0; l < A.nfrag; l++) if (W.val[l].val > mx) { mx = W.val[l].val; pos = (uint32_t)l; }
    *value = mx;
    res[0] = sdp_traceback(W, pos, chain, link);
  }
  wsync();
  const int n = res[0];
  ar.release(mk);
  return n;
}

// mode 2 (forward only, SparseDP_ForwardOnly).  A.q/t/len are the anchors, no clusters.  Returns the chain length.
__device__ __noinline__ int sdp_forward_only(const SdpAnchors &A, int irate, const Pwl &P, Arena &ar, uint32_t *chain, float *value) {
  if (A.nfrag == 0) return 0;
  const unsigned long long mk = ar.mark();
  SdpWork W;
  if (!sdp_build(W, A, 2, 0, 0.0f, irate, ar)) { ar.release(mk); return -1; }
  int *res = ar.alloc<int>(2);
  uint8_t *link = ar.alloc<uint8_t>(A.nfrag + 1);
  if (ar.overflow || !sdp_open_dyn(W, ar)) { ar.release(mk); return -1; }
  { const unsigned long long tp_ = ar.now(); sdp_process(W, A, 0, 2, 0.0f, irate, P); ar.tick(18, tp_); }
  { const unsigned long long e = W.dyn.base_off + *W.dyn.top; if (e > ar.peak) ar.peak = e; }
  if (*W.dyn.err) { ar.release(mk); return -1; }
  if (lane_id() == 0) {
    float mx = 0.0f; uint32_t pos = 0;
    for (int l = 0; l < A.nfrag; l++) if (W.val[l].val > mx) { mx = W.val[l].val; pos = (uint32_t)l; }
    *value = mx;
    res[0] = sdp_traceback(W, pos, chain, link);
  }
  wsync();
  const int n = res[0];
  ar.release(mk);
  return n;
}

// ---- stand-alone batch (lra_b200_sdp_batch): one problem per warp ------------------------------------------------------
struct SdpBatch {
  int n_prob, max_aln;
  const int *mode;                         // 0 pure matches, 1 one cluster, 2 forward only, 3 same-diagonal anchors of a split chain
  const unsigned long long *frag_off;      // [n_prob + 1] into q / t / len
  const uint32_t *q, *t; const int32_t *len;
  const unsigned long long *cl_off_off;    // [n_prob + 1] into cl_off (ncl + 1 entries per problem, problem-relative) and cl_strand (ncl + 1 slots)
  const int *cl_off; const uint8_t *cl_strand;
  const int *only_cl; const float *rate; const int *irate; const int *read_len;
  // mode 4 only (per fragment, at frag_off): box ends, strand, Cluster::Val, NumofAnchors0; out_n0 [n_prob * max_aln]
  const uint32_t *qe, *te; const uint8_t *fstrand; const float *fval; const int32_t *fn0; int *out_n0; int globalK;
  float alnthres; int NumAln;
  const Pwl *pwl;
  // out: chain c of problem p: chain / link at max_aln * frag_off[p] + c * nfrag_p; scalars at p * max_aln + c
  int *n_chains; int *chain_len; float *chain_val; uint32_t *bounds; uint32_t *chain; uint8_t *link; int *cl_of_frag;
  unsigned char *arena; unsigned long long arena_per_warp; int *err;
  unsigned long long *peak;                // optional: high-water mark of one problem's arena use (sizing aid)
};

__global__ void __launch_bounds__(128) sdp_batch_kernel(SdpBatch b) {
  const int warps_per_block = (int)blockDim.x / kLanes;
  const int wid = (int)blockIdx.x * warps_per_block + (int)threadIdx.x / kLanes;
  const int nw = (int)gridDim.x * warps_per_block;
  Arena ar; ar.init(b.arena + (unsigned long long)wid * b.arena_per_warp, b.arena_per_warp);
  for (int p = wid; p < b.n_prob; p += nw) {
    ar.top = 0; ar.overflow = 0; ar.peak = 0;
    const unsigned long long fo = b.frag_off[p]; const int nf = (int)(b.frag_off[p + 1] - fo);
    const unsigned long long co = b.cl_off_off[p]; const int ncl = (int)(b.cl_off_off[p + 1] - co) - 1;
    SdpAnchors A; A.q = b.q + fo; A.t = b.t + fo; A.len = b.len + fo; A.nfrag = nf; A.cl_off = b.cl_off + co; A.cl_strand = b.cl_strand + co; A.ncl = ncl;
    A.qe = b.qe ? b.qe + fo : nullptr; A.te = b.te ? b.te + fo : nullptr; A.fstrand = b.fstrand ? b.fstrand + fo : nullptr; A.fval = b.fval ? b.fval + fo : nullptr;
    A.fn0 = b.fn0 ? b.fn0 + fo : nullptr;
    const int mode = b.mode[p];
    uint32_t *cb = b.chain + (unsigned long long)b.max_aln * fo; uint8_t *lb = b.link + (unsigned long long)b.max_aln * fo;
    int nch = 0;
    if (mode == 0) {
      SdpChain ch[8];
      for (int c = 0; c < b.max_aln && c < 8; c++) { ch[c].chain = cb + (unsigned long long)c * nf; ch[c].link = lb + (unsigned long long)c * nf; ch[c].n = 0; }
      nch = sdp_pure_matches(A, b.rate[p], b.alnthres, b.NumAln < b.max_aln ? b.NumAln : b.max_aln, b.read_len[p], *b.pwl, ar, ch, b.cl_of_frag + fo);
      if (lane_id() == 0) for (int c = 0; c < nch; c++) {
        const int o = p * b.max_aln + c;
        b.chain_len[o] = ch[c].n; b.chain_val[o] = ch[c].value;
        b.bounds[4 * o] = ch[c].QStart; b.bounds[4 * o + 1] = ch[c].QEnd; b.bounds[4 * o + 2] = ch[c].TStart; b.bounds[4 * o + 3] = ch[c].TEnd;
      }
    } else if (mode == 1) {
      float v = 0.0f; float *vp = b.chain_val + p * b.max_aln;
      const int n = sdp_one_cluster(A, b.only_cl[p], b.rate[p], *b.pwl, ar, cb, lb, vp);
      (void)v;
      if (n < 0) nch = -1; else { nch = 1; if (lane_id() == 0) b.chain_len[p * b.max_aln] = n; }
    } else if (mode == 4) {
      SdpChain ch[8]; int n0[8];
      for (int c = 0; c < b.max_aln && c < 8; c++) { ch[c].chain = cb + (unsigned long long)c * nf; ch[c].link = lb + (unsigned long long)c * nf; ch[c].n = 0; n0[c] = 0; }
      nch = sdp_split_clusters(A, b.rate[p], b.alnthres, b.globalK, b.NumAln < b.max_aln ? b.NumAln : b.max_aln, b.read_len[p], *b.pwl, ar, ch, n0);
      if (lane_id() == 0) for (int c = 0; c < nch; c++) {
        const int o = p * b.max_aln + c;
        b.chain_len[o] = ch[c].n; b.chain_val[o] = ch[c].value; if (b.out_n0) b.out_n0[o] = n0[c];
        b.bounds[4 * o] = ch[c].QStart; b.bounds[4 * o + 1] = ch[c].QEnd; b.bounds[4 * o + 2] = ch[c].TStart; b.bounds[4 * o + 3] = ch[c].TEnd;
      }
    } else if (mode == 3) {
      float *vp = b.chain_val + p * b.max_aln;
      const int n = sdp_samediag_chain(A, b.rate[p], *b.pwl, ar, cb, lb, vp);
      if (n < 0) nch = -1; else { nch = 1; if (lane_id() == 0) b.chain_len[p * b.max_aln] = n; }
    } else {
      float *vp = b.chain_val + p * b.max_aln;
      const int n = sdp_forward_only(A, b.irate[p], *b.pwl, ar, cb, vp);
      if (n < 0) nch = -1; else { nch = 1; if (lane_id() == 0) b.chain_len[p * b.max_aln] = n; }
    }
    if (lane_id() == 0) { b.n_chains[p] = nch; if (nch < 0) atomicOr(b.err, 1); if (b.peak) atomicMax(b.peak, ar.peak); }
    wsync();
  }
}

}  // namespace mp
}  // namespace lra
