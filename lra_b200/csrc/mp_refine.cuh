// Mapper worker, refinement of one chain of MapRead_lowacc: Refine_splitchain (ChainRefine.h:383-576), RefineSpace (ClusterRefine.h:242-327),
// RefineBtwnSpace (:331-431, the twoblocks form the low-accuracy pipeline uses), RefineBtwnSpace_AppendCloseCluster + append_to_closetcluster
// (ChainRefine.h:23-187) and Refine_Btwnsplitchain (:578-761).
// Control flow is warp-uniform; lane 0 performs the memory mutations of the serial parts; the window-pair comparisons of Refine_splitchain run one
// pair per lane (count, scan, emit: anchors land in the reference's order) and AffineOneGapAlign uses the whole warp.
#pragma once
#include "mp_stage1.cuh"
#include "mp_aog_band.cuh"
#include "spchain_kernels.cuh"

namespace lra {
namespace mp {

// a refined cluster: a multiset of raw smallK-mers kept as a list of array segments (appends never copy), its box and flags
struct RSeg { uint32_t *q, *t; int n; RSeg *next; };
struct RCluster {
  RSeg *head, *tail; int n;               // matches
  uint32_t qS, qE, tS, tE;                // qStart..tEnd (t chromosome-relative)
  int strand, chrom, refinespace;
  float freq;
};

__device__ inline void rc_append(RCluster &c, RSeg *node, uint32_t *q, uint32_t *t, int n) {   // lane 0
  node->q = q; node->t = t; node->n = n; node->next = 0;
  if (c.tail) c.tail->next = node; else c.head = node;
  c.tail = node; c.n += n;
}
// Cluster::SetClusterBoundariesFromMatches (Clustering.h:308-322), K = opts.globalK of the caller's option set
__device__ __noinline__ void rc_set_boundaries(RCluster &c, int K) {           // lane 0
  bool first = true;
  for (RSeg *s = c.head; s; s = s->next)
    for (int i = 0; i < s->n; i++) {
      const uint32_t q = s->q[i], t = s->t[i];
      if (first) { c.qS = q; c.qE = q + (uint32_t)K; c.tS = t; c.tE = t + (uint32_t)K; first = false; }
      else { c.tE = t + (uint32_t)K > c.tE ? t + (uint32_t)K : c.tE; c.tS = t < c.tS ? t : c.tS; c.qE = q + (uint32_t)K > c.qE ? q + (uint32_t)K : c.qE; c.qS = q < c.qS ? q : c.qS; }
    }
}

// CompareLists<LocalTuple, SmallTuple>(Global = false) (CompareLists.h:8-146) on two sorted LocalTuple lists; push(qi, ti) in the reference's order
template <class Push>
__device__ inline void lt_compare(const uint32_t *q, long nq, const uint32_t *t, long nt, long maxFreq, Push push) {
#define QK(i) lt_t(q[i])
#define TK(i) lt_t(t[i])
  if (nq > 0 && nt > 0) {
    long qs = 0, qe = nq - 1, ts = 0, te = nt;
    do {
      while (qs <= qe && QK(qs) < TK(ts)) qs++;
      if (qs >= qe) break;
      const uint32_t startGap = (QK(qs) - TK(ts)) & 0xFFFFFu;
      while (qe > qs && te > ts && QK(qe) > TK(te - 1)) qe--;
      const uint32_t endGap = (TK(te - 1) - QK(qe)) & 0xFFFFFu;
      if (startGap == 0 || startGap > endGap) {
        const long tsOrig = ts, qsOrig = qs;
        { long lo = ts, len = te - ts;
          const uint32_t key = QK(qs);
          while (len > 0) { const long half = len >> 1, mid = lo + half; if (TK(mid) < key) { lo = mid + 1; len = len - half - 1; } else len = half; }
          ts = lo; }
        if (ts < nt && TK(ts) == QK(qs)) {
          const long tsStart = ts;
          long tsi = ts;
          while (tsi != te && QK(qs) == TK(tsi)) tsi++;
          const long qsStart = qs;
          while (qs < qe && QK(qs + 1) == QK(qs)) qs++;
          for (long ti = tsStart; ti != tsi; ti++)
            if (qs - qsStart < maxFreq)
              for (long qi = qsStart; qi <= qs; qi++) push(qi, ti);
        }
        { const uint32_t k0 = TK(tsOrig); while (ts < te && TK(ts) == k0) ts++; }
        { const uint32_t k0 = QK(qsOrig); while (qs < qe && QK(qs) == k0) qs++; }
      } else {
        if (te != nt && TK(te - 1) == QK(qe)) { /* pass */ }
        else {
          long lo = ts, len = te - ts;
          const uint32_t key = QK(qe);
          while (len > 0) { const long half = len >> 1, mid = lo + half; if (key < TK(mid)) len = half; else { lo = mid + 1; len = len - half - 1; } }
          te = lo;
        }
        const long teStart = te;
        long tei = te;
        while (tei > ts && TK(tei - 1) == QK(qe)) tei--;
        if (tei < teStart && teStart > 0) {
          const long qeStart = qe;
          while (qe > qs && QK(qe) == QK(qe - 1)) qe--;
          for (long ti = tei; ti < teStart; ti++)
            if (qeStart - qe < maxFreq)
              for (long qi = qe; qi <= qeStart; qi++) push(qi, ti);
        }
        te = tei;
      }
    } while (qs < qe && ts < te);
  }
#undef QK
#undef TK
}

struct RsTask { int lsi, qw; uint32_t gStart, rsStart; long long bmin, bmax; };

// the split chains of one chain (output of SPLITChain + RemoveSpuriousSplitChain, spchain_kernels.cuh) in worker form
struct SplitSet {
  int n;                                   // split chains
  int32_t *sp_off, *sptc;                  // anchors of split chain s: chain positions sptc[sp_off[s] .. sp_off[s+1])
  uint32_t *box;                           // [n][4] QStart, QEnd, TStart, TEnd (T global)
  int32_t *chrom; uint8_t *strand, *link;  // link[s]: spchain_link between s and s + 1
  int nlink;
};

// SPLITChain (Mapping_ultility.h:385-444) + RemoveSpuriousSplitChain (Map_lowacc.h:38-67) for one chain
__device__ __noinline__ bool mp_split_chain(const MpCtx &C, const ClusterSet &ext, const UChain &ch, Arena &ar, SplitSet &sp) {
  const int n = ch.n;
  uint32_t *q = ar.alloc<uint32_t>(n + 1), *t = ar.alloc<uint32_t>(n + 1);
  int32_t *len = ar.alloc<int32_t>(n + 1), *cnum = ar.alloc<int32_t>(n + 1);
  uint8_t *st = ar.alloc<uint8_t>(n + 1), *lk = ar.alloc<uint8_t>(n + 1);
  unsigned long long *off = ar.alloc<unsigned long long>(2);
  int32_t *ibuf = ar.alloc<int32_t>(8ull * (n + 1));
  uint32_t *ubuf = ar.alloc<uint32_t>(4ull * (n + 1));
  uint8_t *bbuf = ar.alloc<uint8_t>(4ull * (n + 1));
  int32_t *n_sp = ar.alloc<int32_t>(2);
  sp.sp_off = ar.alloc<int32_t>(n + 3); int32_t *ci_off = ar.alloc<int32_t>(n + 3);
  sp.sptc = ar.alloc<int32_t>(n + 1); int32_t *ci = ar.alloc<int32_t>(n + 1);
  uint8_t *sp_lk = ar.alloc<uint8_t>(n + 1);
  sp.box = ar.alloc<uint32_t>(4ull * (n + 1)); sp.chrom = ar.alloc<int32_t>(n + 1);
  uint8_t *sp_type = ar.alloc<uint8_t>(n + 1); sp.strand = ar.alloc<uint8_t>(n + 1); sp.link = ar.alloc<uint8_t>(n + 1);
  if (ar.overflow) return false;
  for (int i = lane_id(); i < n; i += kLanes) {
    const int k = ch.cl[i], a = ext.off[k] + (int)ch.idx[i];
    q[i] = ext.q[a]; t[i] = ext.t[a]; len[i] = ext.len[a]; cnum[i] = k; st[i] = (uint8_t)(ext.strand[k] != 0);
    lk[i] = (i < ch.nlink) ? ch.link[i] : 0;
  }
  if (lane_id() == 0) { off[0] = 0; off[1] = (unsigned long long)n; }
  wsync();
  if (lane_id() == 0) {
    SpChainBatch b;
    b.n_chains = 1; b.splitdist = C.o.splitdist; b.bypass = 1; b.c_off = off; b.q = q; b.t = t; b.len = len; b.strand = st; b.cnum = cnum; b.link = lk;
    b.hdr_pos = C.ix.hdr_pos; b.n_hdr = C.ix.n_hdr;
    b.pa = ibuf; b.pb = ibuf + (n + 1); b.pnext = ibuf + 2 * (n + 1); b.tail = ibuf + 3 * (n + 1); b.size = ibuf + 4 * (n + 1); b.chrom = ibuf + 5 * (n + 1);
    b.cur_ind = ibuf + 6 * (n + 1); b.ord = ibuf + 7 * (n + 1);
    b.QS = ubuf; b.QE = ubuf + (n + 1); b.TS = ubuf + 2 * (n + 1); b.TE = ubuf + 3 * (n + 1);
    b.type = bbuf; b.pstrand = bbuf + (n + 1); b.keep = bbuf + 2 * (n + 1); b.SL = bbuf + 3 * (n + 1);
    b.n_sp = n_sp; b.n_link = n_sp + 1; b.sp_off = sp.sp_off; b.ci_off = ci_off; b.sptc = sp.sptc; b.ci = ci; b.sp_lk = sp_lk; b.sp_box = sp.box;
    b.sp_chrom = sp.chrom; b.sp_type = sp_type; b.sp_strand = sp.strand; b.sp_link = sp.link;
    spchain_one(b, 0);
  }
  wsync();
  sp.n = n_sp[0]; sp.nlink = n_sp[1];
  return true;
}

// Refine_splitchain for split chain ph -> refined cluster R (matches in one fresh segment)
__device__ __noinline__ bool mp_refine_splitchain(const MpCtx &C, int r, const ClusterSet &ext, const UChain &ch, const SplitSet &sp, int ph, Arena &ar, RCluster &R,
                                            RSeg *node) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  const int a0 = sp.sp_off[ph], nm = sp.sp_off[ph + 1] - a0;
  const int chrom = sp.chrom[ph], Strand = sp.strand[ph];
  const uint32_t chromOffset = (uint32_t)C.ix.hdr_pos[chrom];
  const uint32_t QStart = sp.box[4 * ph], QEnd = sp.box[4 * ph + 1], TStart = sp.box[4 * ph + 2], TEnd = sp.box[4 * ph + 3];
  if (lane == 0) { R.head = R.tail = 0; R.n = 0; R.qS = 0xffffffffu; R.qE = 0; R.tS = 0xffffffffu; R.tE = 0; R.strand = -1; R.chrom = chrom; R.refinespace = 0; R.freq = 0.0f; }
  wsync();
  if (nm == 0) return true;
  const unsigned long long mk = ar.mark();
  uint32_t *mq = ar.alloc<uint32_t>(nm), *mt = ar.alloc<uint32_t>(nm), *ml = ar.alloc<uint32_t>(nm);
  if (ar.overflow) return false;
  long long maxDN = -(1ll << 62), minDN = (1ll << 62);
  for (int i = lane; i < nm; i += kLanes) {
    const int x = sp.sptc[a0 + i], k = ch.cl[x], a = ext.off[k] + (int)ch.idx[x];
    uint32_t q = ext.q[a];
    if (ext.strand[k] == 1) q = L - (q + (uint32_t)O.globalK);      // SwapStrand(read, opts, clusters[cI], opts.globalK)
    const uint32_t t = ext.t[a] - chromOffset;
    mq[i] = q; mt[i] = t; ml[i] = (uint32_t)ext.len[a];
    const long long d = (long long)t - (long long)q;
    maxDN = d > maxDN ? d : maxDN; minDN = d < minDN ? d : minDN;
  }
  maxDN = wmax(maxDN) + 50; minDN = wmin(minDN) - 50;
  wsync();
  const uint32_t chromEndOffset = (uint32_t)C.ix.hdr_pos[lref_hdr_find(C.ix.hdr_pos, C.ix.n_hdr, (unsigned long long)TEnd) + 1];
  const uint32_t wts = (TStart >= chromOffset + (uint32_t)O.window) ? TStart - (uint32_t)O.window : chromOffset;
  const uint32_t wte = (TEnd + (uint32_t)O.window < chromEndOffset) ? TEnd + (uint32_t)O.window : chromEndOffset;
  const LidxView &gl = C.ix.gl;
  const unsigned long long gend = gl.win_off[gl.n_win];
  const int ls = lref_lookup(gl.win_off, gl.n_win, 0ull, gend, wts), le = lref_lookup(gl.win_off, gl.n_win, 0ull, gend, wte);
  const LidxView &rdx = C.rd.rd[Strand];
  const int wf = (int)rdx.win_first[r], nw = (int)rdx.win_first[r + 1] - wf;
  const unsigned long long rbase = rdx.seq_start[r], rend = rbase + L;
  // the window walk (ChainRefine.h:437-521): pass 0 counts the (genome window, read window) pairs, pass 1 records them
  RsTask *tasks = 0;
  int n_tasks = 0;
  for (int pass = 0; pass < 2; pass++) {
    int nt = 0;
    if (lane == 0) {
      int matchStart = 0, matchEnd = 0;
      for (int lsi = ls; lsi <= le; lsi++) {
        if (lsi >= gl.n_win) continue;
        const unsigned long long o0 = gl.win_off[lsi], o1 = gl.win_off[lsi + 1];
        if (o0 < chromOffset || o1 < chromOffset) continue;
        const uint32_t gStart = (uint32_t)(o0 - chromOffset), gEnd = (uint32_t)(o1 - 1 - chromOffset);
        if (gStart >= gEnd) continue;
        while (matchStart < nm && mt[matchStart] <= gStart) matchStart++;
        matchEnd = matchStart;
        while (matchEnd < nm && mt[matchEnd] < gEnd) matchEnd++;
        if (matchStart >= nm) continue;
        if (matchEnd == matchStart) continue;
        uint32_t readStart = mq[matchStart], readEnd = mq[matchEnd - 1];
        long long mn = (long long)mt[matchStart] - (long long)mq[matchStart];
        for (int mi = matchStart; mi < matchEnd; mi++) {
          const uint32_t q = mq[mi];
          if (q < readStart) readStart = q;
          if (q + ml[mi] > readEnd) readEnd = q + ml[mi];
          const long long d = (long long)mt[mi] - (long long)q;
          mn = d < mn ? d : mn;
        }
        if (readStart == readEnd) { if (lsi > ls && readStart > 0u) readStart = 0u; }
        long long bandMin = minDN, bandMax = maxDN;
        // opts.limitrefine: the upper bound is an uninitialised variable in the reference (ChainRefine.h:491-501) that never filters in the stock build
        if (O.limitrefine) { bandMin = mn - 100; bandMax = 0x7FFFFFFFFFFFFFFFll; }
        const uint32_t sow = 500;
        if (lsi == ls) readStart = (readStart < sow) ? 0 : readStart - sow;
        if (lsi == le) readEnd = (readEnd + sow > L) ? L : readEnd + sow;
        if (readStart > readEnd) continue;
        const int qis = lref_lookup(rdx.win_off + wf, nw, rbase, rend, readStart);
        const int qie = lref_lookup(rdx.win_off + wf, nw, rbase, rend, readEnd < L - 1 ? readEnd : L - 1);
        for (int qi = qis; qi <= qie; qi++) {
          if (pass == 1) {
            RsTask tk; tk.lsi = lsi; tk.qw = wf + qi; tk.gStart = gStart; tk.rsStart = (uint32_t)(rdx.win_off[wf + qi] - rbase); tk.bmin = bandMin; tk.bmax = bandMax;
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
  // filter box of AppendValues (ChainRefine.h:527-538)
  uint32_t bqs, bqe;
  if (Strand == 0) { bqs = QStart; bqe = QEnd; } else { bqs = L - QEnd; bqe = L - QStart; }
  const uint32_t bts = TStart - chromOffset, bte = TEnd - chromOffset;
  const long maxFreq = (long)O.localMaxFreq;
  int *cnt = ar.alloc<int>(n_tasks + 1);
  if (ar.overflow) return false;
  for (int k = lane; k < n_tasks; k += kLanes) {
    const RsTask tk = tasks[k];
    const uint32_t *q = rdx.mins + rdx.bnd[tk.qw]; const long nq = (long)(rdx.bnd[tk.qw + 1] - rdx.bnd[tk.qw]);
    const uint32_t *t = gl.mins + gl.bnd[tk.lsi]; const long nt = (long)(gl.bnd[tk.lsi + 1] - gl.bnd[tk.lsi]);
    int c = 0;
    lt_compare(q, nq, t, nt, maxFreq, [&](long qi, long ti) {
      const uint32_t qp = (q[qi] >> 20) + tk.rsStart, tp = (t[ti] >> 20) + tk.gStart;
      const long long d = (long long)tp - (long long)qp;
      if (d >= tk.bmin && d <= tk.bmax && qp >= bqs && qp < bqe && tp >= bts && tp < bte) c++;
    });
    cnt[k] = c;
  }
  wsync();
  int total = 0;
  if (lane == 0) { for (int k = 0; k < n_tasks; k++) { const int c = cnt[k]; cnt[k] = total; total += c; } cnt[n_tasks] = total; }
  wsync();
  total = bcast(total, 0);
  if (total == 0) { ar.release(mk); return true; }
  // the refined anchors outlive this call: allocate them below the scratch mark by releasing first (scratch is dead after the emit pass, which
  // therefore runs into a staging copy) -- simpler: keep everything and only give back nothing.  The arena is reset per chain.
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
  // finish (ChainRefine.h:559-573): swap back, boundaries with smallOpts.globalK
  if (Strand == 1) { for (int i = lane; i < total; i += kLanes) rq[i] = L - (rq[i] + (uint32_t)O.smallK); wsync(); }
  if (lane == 0) { rc_append(R, node, rq, rt, total); rc_set_boundaries(R, O.smallK); R.strand = Strand; R.refinespace = 0; }
  wsync();
  return true;
}

// RefineSpace (ClusterRefine.h:242-327).  Result pairs in (*pq, *pt) (arena, kept), count returned (-1: arena overflow); identity in *ident.
__device__ __noinline__ int mp_refine_space(const MpCtx &C, int r, Arena &ar, int K, int W, int refineSpaceDiag, bool consider_str, int maxFreq, int chrom, uint32_t qe,
                                      uint32_t qs, uint32_t te, uint32_t ts, int st, uint32_t lrts, uint32_t lrlength, uint32_t **pq, uint32_t **pt, float *ident) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  const unsigned long long roff = C.rd.read_off[r];
  const unsigned long long coff = C.ix.hdr_pos[chrom];
  const SeqView &rs = st ? C.rd.rc : C.rd.fwd;
  const uint32_t qlen = qe - qs, tlen = te - ts + lrlength;
  const uint32_t tshift = ts - lrts;
  const long long diag2 = (long long)(uint32_t)(te - (ts - lrts)) - (long long)(uint32_t)(qe - qs);
  const long long minDiagNum = (diag2 < 0 ? diag2 : 0) - (long long)refineSpaceDiag, maxDiagNum = (diag2 > 0 ? diag2 : 0) + (long long)refineSpaceDiag;
  float identity = -1.0f;
  int np = 0;
  uint32_t *oq = 0, *ot = 0;
  if (qlen < 1000 && tlen < 1000) {
    const int capb = (int)(qlen < tlen ? qlen : tlen) + 2;
    uint32_t *blk = ar.alloc<uint32_t>(3ull * capb);
    const int capp = (int)((qlen < tlen ? qlen : tlen) / (uint32_t)K) + 2;
    oq = ar.alloc<uint32_t>(capp); ot = ar.alloc<uint32_t>(capp);
    int *errp = ar.alloc<int>(1);
    if (ar.overflow) return -1;
    int nb = 0;
    mp_aog_any(rs, (uint32_t)(roff + qs), (int)qlen, C.ix.genome, (uint32_t)(coff + tshift), (int)tlen, O.localMatch, O.localMismatch, O.localIndel, 30, ar, blk, capb, &nb, errp);
    if (nb < 0) return -1;
    wsync();
    int nMatch = 0;
    if (lane == 0) {
      const unsigned long long q0 = roff + qs, t0 = coff + tshift;
      for (int i = 0; i < nb; i++) {
        const uint32_t bq = blk[3 * i], bt = blk[3 * i + 1], len = blk[3 * i + 2];
        for (uint32_t x = 0; x < len; x++) nMatch += seq_code(rs, q0 + bq + x) == seq_code(C.ix.genome, t0 + bt + x) ? 1 : 0;
        if (len > (uint32_t)K) {
          for (uint32_t bp = 0; bp + (uint32_t)K < len; bp += (uint32_t)K) {
            bool mis = false;
            for (uint32_t x = 0; x < (uint32_t)K; x++) if (seq_code(C.ix.genome, t0 + bt + bp + x) != seq_code(rs, q0 + bq + bp + x)) { mis = true; break; }
            if (!mis && np < capp) { oq[np] = bq + bp; ot[np] = bt + bp; np++; }
          }
        }
      }
      const uint32_t mn = qlen < tlen ? qlen : tlen;
      identity = (mn == 0 && nMatch == 0) ? __uint_as_float(0xFFC00000u) : __fdiv_rn((float)nMatch, (float)mn);
    }
    wsync();
    np = bcast(np, 0); identity = bcast(identity, 0);
  } else {
    // StoreMinimizers_noncanonical of both windows, std::sort, CompareLists(Global = false) inside the diagonal band
    unsigned long long *gt = ar.alloc<unsigned long long>((unsigned long long)tlen + 2), *qt = ar.alloc<unsigned long long>((unsigned long long)qlen + 2);
    uint32_t *gp = ar.alloc<uint32_t>((unsigned long long)tlen + 2), *qp = ar.alloc<uint32_t>((unsigned long long)qlen + 2);
    if (ar.overflow) return -1;
    const unsigned long long top0 = (ar.top + 15ull) & ~15ull;
    const unsigned long long room = ar.cap > top0 ? (ar.cap - top0) / 8ull : 0ull;       // (q, t) pairs interleaved in the open end
    uint32_t *open = (uint32_t *)(ar.base + top0);
    long long cntp = 0;
    if (lane == 0) {
      const uint32_t ng = mm_scan<false>(C.ix.genome, coff + tshift, tlen, K, W, gt, gp);
      mm_sort(MmRef{gt, gp}, (long)ng);
      const uint32_t nq = mm_scan<false>(rs, roff + qs, qlen, K, W, qt, qp);
      mm_sort(MmRef{qt, qp}, (long)nq);
      if (nq != 0 && ng != 0)
        mm_compare(qt, (long)nq, gt, (long)ng, (long long)maxFreq, [&](long qi, long ti) {
          if (maxDiagNum != 0 && minDiagNum != 0) {
            const long long D = (long long)gp[ti] - (long long)qp[qi];
            if (!(D <= maxDiagNum && D >= minDiagNum)) return;
          }
          if ((unsigned long long)cntp < room / 2) { open[2 * cntp] = qp[qi]; open[2 * cntp + 1] = gp[ti]; }
          cntp++;
        });
    }
    wsync();
    cntp = bcast(cntp, 0);
    if ((unsigned long long)cntp >= room / 2) return -1;
    np = (int)cntp;
    ar.alloc<uint32_t>(2ull * np);
    oq = ar.alloc<uint32_t>(np + 1); ot = ar.alloc<uint32_t>(np + 1);
    if (ar.overflow) return -1;
    for (int i = lane; i < np; i += kLanes) { oq[i] = open[2 * i]; ot[i] = open[2 * i + 1]; }
    wsync();
  }
  for (int i = lane; i < np; i += kLanes) {
    uint32_t fq = oq[i] + qs;
    if (consider_str && st == 1) fq = L - fq - (uint32_t)K;
    oq[i] = fq; ot[i] += tshift;
  }
  wsync();
  *pq = oq; *pt = ot; *ident = identity;
  return np;
}

// RefineBtwnSpace (ClusterRefine.h:331-431) with twoblocks == true, the only form MapRead_lowacc reaches
__device__ __noinline__ bool mp_refine_btwn_space(const MpCtx &C, int r, Arena &ar, RCluster &cl, uint32_t qe, uint32_t qs, uint32_t te, uint32_t ts, int st, uint32_t lrts,
                                            uint32_t lrlength) {
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  if (st == 1) { const uint32_t t = qs; qs = L - qe; qe = L - t; }
  int refineSpaceDiag = 0;
  if (O.readType == 3 || O.readType == 2) { const float v = fmaxf(100.f, __fmul_rn(0.01f, (float)(uint32_t)(qe - qs))); const int f = (int)floorf(v); refineSpaceDiag = f < 100 ? f : 100; }
  else { const float v = fmaxf(100.f, __fmul_rn(0.15f, (float)(uint32_t)(qe - qs))); const int f = (int)floorf(v); refineSpaceDiag = f < 1000 ? f : 1000; }
  uint32_t *pq, *pt; float ident;
  RSeg *node = ar.alloc<RSeg>(1);
  if (ar.overflow) return false;
  const int np = mp_refine_space(C, r, ar, O.smallK, O.smallW, refineSpaceDiag, true, O.localMaxFreq, cl.chrom, qe, qs, te, ts, st, lrts, lrlength, &pq, &pt, &ident);
  if (np < 0) return false;
  if (np > 0) {
    if (lane_id() == 0) { rc_append(cl, node, pq, pt, np); rc_set_boundaries(cl, O.smallK); cl.refinespace = 1; }
    wsync();
  }
  return true;
}

// append_to_closetcluster (ChainRefine.h:23-54): pairs [start, end) of a sorted list go to the nearer of the two clusters
__device__ __noinline__ void mp_append_to_closest(const uint32_t *pq, const uint32_t *pt, int start, int end, RCluster &cluster, RCluster &prev, RSeg *node, int st, int K) {   // lane 0
  uint32_t qStart = pq[start], qEnd = qStart + (uint32_t)K, tStart = pt[start], tEnd = tStart + (uint32_t)K;
  for (int i = start + 1; i < end; i++) {
    tEnd = pt[i] + (uint32_t)K > tEnd ? pt[i] + (uint32_t)K : tEnd; tStart = pt[i] < tStart ? pt[i] : tStart;
    qEnd = pq[i] + (uint32_t)K > qEnd ? pq[i] + (uint32_t)K : qEnd; qStart = pq[i] < qStart ? pq[i] : qStart;
  }
  int qdist = (qStart >= cluster.qE) ? (int)(qStart - cluster.qE) : 0, tdist;
  if (st == 0) tdist = (tStart >= cluster.tE) ? (int)(tStart - cluster.tE) : 0; else tdist = (cluster.tS >= tEnd) ? (int)(cluster.tS - tEnd) : 0;
  const int dist_cur = qdist > tdist ? qdist : tdist;
  qdist = (prev.qS >= qEnd) ? (int)(prev.qS - qEnd) : 0;
  if (st == 0) tdist = (prev.tS >= tEnd) ? (int)(prev.tS - tEnd) : 0; else tdist = (tStart >= prev.tE) ? (int)(tStart - prev.tE) : 0;
  const int dist_prev = qdist > tdist ? qdist : tdist;
  RCluster &dst = dist_cur <= dist_prev ? cluster : prev;
  rc_append(dst, node, const_cast<uint32_t *>(pq) + start, const_cast<uint32_t *>(pt) + start, end - start);
  rc_set_boundaries(dst, K);
}

// RefineBtwnSpace_AppendCloseCluster (ChainRefine.h:58-120)
__device__ __noinline__ bool mp_refine_btwn_append(const MpCtx &C, int r, Arena &ar, bool twoblocks, RCluster &cluster, RCluster &prev, uint32_t qe, uint32_t qs, uint32_t te,
                                             uint32_t ts, int st) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  const int K = O.smallK;
  if (st == 1) { const uint32_t t = qs; qs = L - qe; qe = L - t; }
  int refineSpaceDiag = 0;     // (uninitialised in the reference for other read types; the low-accuracy presets are clr / ont)
  { const float v = fmaxf(100.f, __fmul_rn(0.15f, (float)(uint32_t)(qe - qs))); const int f = (int)floorf(v); refineSpaceDiag = f < 1000 ? f : 1000; }
  uint32_t *pq, *pt; float ident;
  const int np = mp_refine_space(C, r, ar, K, O.smallW, refineSpaceDiag, true, O.localMaxFreq, cluster.chrom, qe, qs, te, ts, st, 0, 0, &pq, &pt, &ident);
  if (np < 0) return false;
  const uint32_t mn = (qe - qs) < (te - ts) ? (qe - qs) : (te - ts);
  const float eff = __fdiv_rn((float)np, (float)mn);
  if (np == 0) return true;
  const float thr = __fmul_rn(O.anchorstoosparse, 2.0f);
  if (eff >= thr) {
    RSeg *node = ar.alloc<RSeg>(1);
    if (ar.overflow) return false;
    if (lane == 0) { rc_append(cluster, node, pq, pt, np); rc_set_boundaries(cluster, K); cluster.refinespace = 1; }
    wsync();
    return true;
  }
  if (twoblocks) return true;
  // CartesianSort(EndPairs)
  const unsigned long long mk = ar.mark();
  MpKey *keys = ar.alloc<MpKey>((unsigned long long)next_pow2(np));
  uint32_t *sq = ar.alloc<uint32_t>(np), *stt = ar.alloc<uint32_t>(np);
  if (ar.overflow) return false;
  for (int i = lane; i < np; i += kLanes) { keys[i].k = ((unsigned long long)pq[i] << 32) | pt[i]; keys[i].q = 0; keys[i].idx = (uint32_t)i; }
  wsync();
  mp_sort_keys(keys, np);
  for (int i = lane; i < np; i += kLanes) { sq[i] = pq[keys[i].idx]; stt[i] = pt[keys[i].idx]; }
  wsync();
  for (int i = lane; i < np; i += kLanes) { pq[i] = sq[i]; pt[i] = stt[i]; }
  wsync();
  ar.release(mk);
  // (max_pairdist <= 100 and eff >= thr cannot hold here: eff < thr) -> chunks of consecutive pairs within 200 of each other, at least 4 long
  int start = 0;
  while (start < np) {
    int end = start + 1;
    while (end < np) {
      long long a = (long long)pq[end] - (long long)pq[end - 1]; if (a < 0) a = -a;
      long long b = (long long)pt[end] - (long long)pt[end - 1]; if (b < 0) b = -b;
      if ((a < b ? a : b) <= 200) end++; else break;
    }
    if (end - start >= 4) {
      RSeg *node = ar.alloc<RSeg>(1);
      if (ar.overflow) return false;
      if (lane == 0) mp_append_to_closest(pq, pt, start, end, cluster, prev, node, st, K);
      wsync();
    }
    start = end;
  }
  return true;
}

// Refine_Btwnsplitchain (ChainRefine.h:578-761) over the refined clusters RC[0..n) of the split chains
__device__ __noinline__ bool mp_refine_btwn_splitchain(const MpCtx &C, int r, Arena &ar, const SplitSet &sp, RCluster *RC) {
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  const int n = sp.n;
  for (int c = 1; c < n; c++) {
    wsync();
    RCluster &cur = RC[c], &prev = RC[c - 1];
    if (cur.n == 0 || prev.n == 0) continue;
    uint32_t qs = cur.qE, qe = prev.qS, ts1 = 0, te1 = 0, ts2 = 0, te2 = 0;
    if (qe <= qs) continue;
    int st1 = 0, st2 = 0; bool twoblocks = false;
    const int lk = sp.link[c - 1];
    if (cur.strand == prev.strand && lk == 0) {
      twoblocks = false; st1 = cur.strand;
      if (cur.tE <= prev.tS) { ts1 = cur.tE; te1 = prev.tS; }
      else if (cur.tS > prev.tE) { ts1 = prev.tE; te1 = cur.tS; }
      else continue;
    } else if (cur.strand != prev.strand && lk == 1) {
      st1 = cur.strand; st2 = prev.strand; twoblocks = true;
      const uint32_t d = qe - qs;
      if (cur.tE <= prev.tS) {
        if (st1 == 0) { ts1 = cur.tE; te1 = ts1 + d; ts2 = prev.tE; te2 = ts2 + d; }
        else { te1 = cur.tS; ts1 = te1 > d ? te1 - d : 0; te2 = prev.tS; ts2 = te2 > d ? te2 - d : 0; }
      } else if (cur.tS > prev.tE) {
        if (st1 == 0) { ts1 = cur.tE; te1 = ts1 + d; te2 = cur.tS; ts2 = te2 > d ? te2 - d : 0; }
        else { te1 = cur.tS; ts1 = te1 > d ? te1 - d : 0; te2 = prev.tS; ts2 = te2 > d ? te2 - d : 0; }
      } else continue;
    } else if (cur.strand == prev.strand && lk == 1) {
      st1 = cur.strand; st2 = st1; twoblocks = true;
      const uint32_t d = qe - qs;
      if (st1 == 0 && cur.tE > prev.tS) { ts1 = cur.tE; te1 = ts1 + d; te2 = prev.tS; ts2 = te2 > d ? te2 - d : 0; }
      else if (st1 == 1 && cur.tS < prev.tE) { te1 = cur.tS; ts1 = te1 > d ? te1 - d : 0; ts2 = prev.tE; te2 = ts2 + d; }
      else continue;
    }
    // (strands differ and link == 0: st1 / st2 / twoblocks keep the values of the previous iteration in the reference; with ts1 = te1 = 0 the
    //  next test leaves the iteration)
    if (te1 <= ts1) continue;
    const uint32_t clen = contig_len(C.ix, cur.chrom);
    if (te1 >= clen) continue;
    const uint32_t sl1 = (qe - qs) > (te1 - ts1) ? (qe - qs) : (te1 - ts1);
    if (sl1 >= 5u * (uint32_t)O.refineSpaceDist) continue;
    if (sl1 >= 20 && sl1 <= (uint32_t)O.refineSpaceDist && cur.chrom == prev.chrom) {
      if (!mp_refine_btwn_append(C, r, ar, twoblocks, cur, prev, qe, qs, te1, ts1, st1)) return false;
    }
    if (twoblocks) {
      if (te2 <= ts2) continue;
      if (te2 >= clen) continue;
      const uint32_t sl2 = (qe - qs) > (te2 - ts2) ? (qe - qs) : (te2 - ts2);
      if (sl2 >= 5u * (uint32_t)O.refineSpaceDist) continue;
      if (sl2 >= 20 && sl2 <= (uint32_t)O.refineSpaceDist && cur.chrom == prev.chrom) {
        if (!mp_refine_btwn_space(C, r, ar, prev, qe, qs, te2, ts2, st2, 0, 0)) return false;
      }
    }
  }
  wsync();
  // right end: splitchains[0].clusterIndex == 0
  {
    RCluster &rh = RC[0];
    if (rh.n > 0) {
      const int st = rh.strand;
      const uint32_t qs = rh.qE, qe = L;
      uint32_t ts = 0, te = 0; bool ts_set = true;
      if (st == 0) { ts = rh.tE; te = ts + qe - qs; }
      else { te = rh.tS; if (te > qe - qs) ts = te - (qe - qs); else { te = 0; ts_set = false; } }
      // (`ts` is uninitialised on that last path; te == 0 makes the guard below false for every unsigned ts)
      if (ts_set && qe > qs && te > ts) {
        const uint32_t slen = (qe - qs) > (te - ts) ? (qe - qs) : (te - ts);
        if (slen >= 20 && slen < (uint32_t)O.refineSpaceDist && te + 500u < contig_len(C.ix, rh.chrom)) {
          uint32_t lrts = 0, lrlength = 0;
          if (st == 0) { lrts = 0; lrlength = 500; } else { if (ts > 500) lrts = 500; lrlength = lrts; }
          if (!mp_refine_btwn_space(C, r, ar, rh, qe, qs, te, ts, st, lrts, lrlength)) return false;
        }
      }
    }
  }
  wsync();
  // left end
  {
    RCluster &lh = RC[n - 1];
    if (lh.n > 0) {
      const uint32_t qs = 0, qe = lh.qS;
      const int st = lh.strand;
      uint32_t ts, te;
      if (st == 0) { te = lh.tS; ts = te > qe - qs ? te - (qe - qs) : 0; } else { ts = lh.tE; te = ts + (qe - qs); }
      if (qe > qs && te > ts) {
        const uint32_t slen = (qe - qs) > (te - ts) ? (qe - qs) : (te - ts);
        if (slen >= 20 && slen < (uint32_t)O.refineSpaceDist && te + 500u < contig_len(C.ix, lh.chrom)) {
          uint32_t lrts = 0, lrlength = 0;
          if (st == 0) { if (ts > 500) lrts = 500; lrlength = lrts; } else { lrts = 0; lrlength = 500; }
          if (!mp_refine_btwn_space(C, r, ar, lh, qe, qs, te, ts, st, lrts, lrlength)) return false;
        }
      }
    }
  }
  wsync();
  return true;
}

}  // namespace mp
}  // namespace lra
