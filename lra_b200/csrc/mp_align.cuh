// Mapper worker, second half of one chain of MapRead_lowacc: MergeChain (ChainRefine.h:767-802), LinearExtend + TrimOverlappedAnchors on the merged
// clusters (Map_lowacc.h:444-475, LinearExtend.h:574-649, 722-780), the second SparseDP + RemovePairedIndels + RemoveSpuriousAnchors (Map_lowacc.h:529-541),
// LocalRefineAlignment (LocalRefineAlignment.h:884-1030) with RefinedAlignmentbtwnAnchors (:203-550) and RefineByLinearAlignment (:144-185).
#pragma once
#include "mp_refine.cuh"

namespace lra {
namespace mp {

// blocks of the segments of one chain, built sequentially (only the last segment ever receives blocks)
struct SegBuild {
  uint32_t *blk; int nblk, cap;           // (qPos, tPos, length) triples of all segments, concatenated
  int *seg_start;                         // first block of every segment
  int *seg_strand, *seg_chrom, *seg_n0, *seg_n1, *seg_supp; float *seg_val;
  int nseg, cap_seg;
  // RefineByLinearAlignment calls are independent of one another: they are queued (a placeholder block with length 0 holds the job index) and
  // aligned 32 at a time, one per lane, by mp_flush_aog; the placeholders are then replaced by the jobs' blocks
  struct AogJob *jobs; int njobs, cap_jobs;
};
struct AogJob { uint32_t qoff, toff, addq, addt; int qLen, tLen, band, str; };
__device__ __forceinline__ bool sb_push_block(SegBuild &B, uint32_t q, uint32_t t, uint32_t len) {
  if (B.nblk >= B.cap) return false;
  if (lane_id() == 0) { B.blk[3 * B.nblk] = q; B.blk[3 * B.nblk + 1] = t; B.blk[3 * B.nblk + 2] = len; }
  B.nblk++;
  return true;
}
__device__ __forceinline__ bool sb_new_segment(SegBuild &B, int strand, int chrom, int n0, int n1, int supp, float val) {
  if (B.nseg >= B.cap_seg) return false;
  if (lane_id() == 0) { B.seg_start[B.nseg] = B.nblk; B.seg_strand[B.nseg] = strand; B.seg_chrom[B.nseg] = chrom; B.seg_n0[B.nseg] = n0; B.seg_n1[B.nseg] = n1; B.seg_supp[B.nseg] = supp; B.seg_val[B.nseg] = val; }
  B.nseg++;
  return true;
}

// TrimOverlappedAnchors (LinearExtend.h:574-649 cluster form: thr 40, strand aware; :722-780 pair form: thr 50, strand 0); serial
__device__ __noinline__ void mp_trim_overlapped(uint32_t *q, uint32_t *t, int *len, int n, int strand, int thr, bool cluster_form, int *idx /* scratch n */) {
  int nl = 0;
  for (int i = 0; i < n; i++) if (len[i] >= thr) idx[nl++] = i;
  auto less = [&](int i, int j) {
    if (strand == 0) { if (q[i] != q[j]) return q[i] < q[j]; return t[i] < t[j]; }
    if (q[i] + (uint32_t)len[i] != q[j] + (uint32_t)len[j]) return q[i] + (uint32_t)len[i] > q[j] + (uint32_t)len[j];
    return t[i] < t[j];
  };
  std_sort_replay(idx, nl, less);
  for (int ln = 1; ln < nl; ln++) {
    const int prev = idx[ln - 1], cur = idx[ln];
    int overlap_r = 0, overlap_g = 0;
    if (strand == 0) {
      if (q[cur] < q[prev] + (uint32_t)len[prev] && q[cur] >= q[prev] + (uint32_t)len[prev] - 30u) overlap_r = (int)(q[prev] + (uint32_t)len[prev] - q[cur]);
    } else {
      if (q[cur] + (uint32_t)len[cur] > q[prev] && q[cur] + (uint32_t)len[cur] <= q[prev] + 30u) overlap_r = (int)(q[cur] + (uint32_t)len[cur] - q[prev]);
    }
    if (t[cur] < t[prev] + (uint32_t)len[prev] && t[cur] >= t[prev] + (uint32_t)len[prev] - 30u) overlap_g = (int)(t[prev] + (uint32_t)len[prev] - t[cur]);
    if (overlap_r > 0 || overlap_g > 0) {
      const int overlap = overlap_r > overlap_g ? overlap_r : overlap_g;
      if (cluster_form && strand == 1) q[prev] += (uint32_t)(overlap + 1);
      len[prev] -= overlap + 1;
    }
  }
}

// The same on the whole warp (cluster form: thr 40, strand aware).  The long anchors are sorted by the LongAnchors key with a warp bitonic sort; the reference's std::sort leaves
// equal keys (two long anchors with the same start, or the same end on the reverse strand, and the same t) in introsort's order, which decides which of them is `prev`, so a list
// with equal keys goes to the literal replay above.  Iteration ln of the trim loop reads anchors idx[ln - 1], idx[ln] and writes idx[ln - 1] only, and no earlier iteration writes
// either, so 32 iterations run at a time: all reads, then all writes (lext_kernels.cuh: lext_trim_warp).  scratch: next_pow2(n) MpKey + n ints from the arena.
__device__ __noinline__ bool mp_trim_overlapped_warp(uint32_t *q, uint32_t *t, int *len, int n, int strand, int thr, bool cluster_form, Arena &ar) {
  const int lane = lane_id();
  if (n <= 0) return true;
  const unsigned long long mk = ar.mark();
  int *idx = ar.alloc<int>(n + 1);
  if (ar.overflow) { ar.release(mk); return false; }
  int nl = 0;
  for (int base = 0; base < n; base += kLanes) {
    const int i = base + lane;
    const bool is_long = i < n && len[i] >= thr;
    const unsigned m = ballot(is_long);
    if (is_long) idx[nl + __popc(m & lanemask_lt())] = i;
    nl += __popc(m);
  }
  wsync();
  if (nl >= 2) {
    MpKey *keys = ar.alloc<MpKey>((unsigned long long)next_pow2(nl));
    if (ar.overflow) { ar.release(mk); return false; }
    for (int i = lane; i < nl; i += kLanes) {
      const int a = idx[i];
      const uint32_t k0 = strand == 0 ? q[a] : ~(q[a] + (uint32_t)len[a]);
      keys[i].k = ((unsigned long long)k0 << 32) | t[a]; keys[i].q = 0; keys[i].idx = (uint32_t)a;
    }
    wsync();
    mp_sort_keys(keys, nl);
    bool dup = false;
    for (int i = 1 + lane; i < nl; i += kLanes) dup = dup || keys[i].k == keys[i - 1].k;
    if (wany(dup)) {
      if (lane == 0) mp_trim_overlapped(q, t, len, n, strand, thr, cluster_form, idx);
      wsync();
      ar.release(mk);
      return true;
    }
    for (int i = lane; i < nl; i += kLanes) idx[i] = (int)keys[i].idx;
    wsync();
    for (int base = 1; base < nl; base += kLanes) {
      const int ln = base + lane;
      int prev = 0, cut = 0;
      if (ln < nl) {
        prev = idx[ln - 1];
        const int cur = idx[ln];
        int overlap_r = 0, overlap_g = 0;
        const uint32_t pend = q[prev] + (uint32_t)len[prev];
        if (strand == 0) { if (q[cur] < pend && q[cur] >= pend - 30u) overlap_r = (int)(pend - q[cur]); }
        else { const uint32_t cend = q[cur] + (uint32_t)len[cur]; if (cend > q[prev] && cend <= q[prev] + 30u) overlap_r = (int)(cend - q[prev]); }
        const uint32_t ptend = t[prev] + (uint32_t)len[prev];
        if (t[cur] < ptend && t[cur] >= ptend - 30u) overlap_g = (int)(ptend - t[cur]);
        if (overlap_r > 0 || overlap_g > 0) cut = (overlap_r > overlap_g ? overlap_r : overlap_g) + 1;
      }
      wsync();
      if (cut) { if (cluster_form && strand == 1) q[prev] += (uint32_t)cut; len[prev] -= cut; }
      wsync();
    }
  }
  ar.release(mk);
  return true;
}

// One AffineOneGapAlign job by one lane: the one-sided case with an even doubled half-width K <= 14 and at most kLaneAogRows rows, i.e. the body
// of aog_thread_kernel<K> (aog_kernels.cuh) with K at run time.  Blocks go to `out` in the reference's order; returns their number.
constexpr int kLaneAogRows = 96;
__device__ __forceinline__ bool mp_aog_lane_ok(int qLen, int tLen, int k_in) {
  const AogShape s = aog_shape(qLen, tLen, k_in, 0);
  return s.cls < kAogThreadClasses && s.rows <= kLaneAogRows;
}
__device__ __noinline__ int mp_aog_lane(const SeqView &q, uint32_t qoff, int qLen, const SeqView &t, uint32_t toff, int tLen, int m, int mm, int indel, int k_in, uint32_t *out) {
  const AogShape sh = aog_shape(qLen, tLen, k_in, 0);
  const int K = sh.k, W = 2 * K + 1;
  const int diag = sh.diag, qB = sh.qB, tB = sh.tB;
  const bool keep0 = !((qLen >= tLen && diag - K - 1 >= 0) || (qLen <= tLen && diag >= 2));
  unsigned long long tb[kLaneAogRows + 1];
  int prev[31];
  for (int c = 0; c <= W; c++) { const int i = c - K; prev[c] = (i < 0 || i > K) ? kMissing : indel * i; }
  SeqStream qs, ts;
  qs.init(q, qoff); ts.init(t, toff);
  unsigned long long qw = 0, qn = 0;
  int qnext = 1;
  for (int c = K; c <= 2 * K; c++) {
    int code = 0;
    if (qnext <= qLen) code = qs.next();
    qnext++;
    qw |= (unsigned long long)(code & 3) << (2 * c);
    qn |= (unsigned long long)(code == 4 ? 1 : 0) << (2 * c);
  }
  const unsigned long long kOdd = 0x5555555555555555ull;
  const int rows = tB - 1;
  for (int j = 1; j <= rows; j++) {
    const int tc = ts.next();
    unsigned long long e;
    if (tc == 4) e = qn;
    else { const unsigned long long x = qw ^ ((unsigned long long)tc * kOdd); e = ~(x | (x >> 1)) & kOdd & ~qn; }
    int run = (j == K + 1 && keep0) ? indel * (K + 1) : kMissing;
    unsigned long long bits = 0;
    for (int c = 0; c < W; c++) {
      const int sM = prev[c] + (((e >> (2 * c)) & 1ull) ? m : mm);
      const int sD = prev[c + 1] + indel;
      const int sI = run + indel;
      const int best = imax(sI, imax(sD, sM));
      const int arrow = (best == sI) ? AR_LEFT : ((best == sD) ? AR_DOWN : AR_DIAG);
      bits |= (unsigned long long)arrow << (2 * c);
      prev[c] = best;
      run = best;
    }
    tb[j] = bits;
    qw >>= 2; qn >>= 2;
    int code = 0;
    if (qnext <= qLen) code = qs.next();
    qnext++;
    qw |= (unsigned long long)(code & 3) << (4 * K);
    qn |= (unsigned long long)(code == 4 ? 1 : 0) << (4 * K);
  }
  int nb = 0;
  { int i = qB - 1, j = tB - 1, run = 0;
    while (i > 0 && j > 0) {
      const int a = (int)((tb[j] >> (2 * (i - j + K))) & 3ull);
      if (a == AR_DIAG) { run++; i--; j--; } else { if (run) { nb++; run = 0; } if (a == AR_LEFT) i--; else j--; }
    }
    if (run) nb++; }
  { int i = qB - 1, j = tB - 1, run = 0, r = nb - 1;
    while (i > 0 && j > 0) {
      const int a = (int)((tb[j] >> (2 * (i - j + K))) & 3ull);
      if (a == AR_DIAG) { run++; i--; j--; }
      else { if (run) { out[3 * r] = (uint32_t)i; out[3 * r + 1] = (uint32_t)j; out[3 * r + 2] = (uint32_t)run; r--; run = 0; } if (a == AR_LEFT) i--; else j--; }
    }
    if (run) { out[3 * r] = (uint32_t)i; out[3 * r + 1] = (uint32_t)j; out[3 * r + 2] = (uint32_t)run; } }
  return nb;
}

// Run the queued RefineByLinearAlignment jobs and splice their blocks into the segment's block list
__device__ __noinline__ bool mp_flush_aog(const MpCtx &C, int r, Arena &ar, SegBuild &B) {
  const int nj = B.njobs;
  if (nj == 0) return true;
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const unsigned long long tk = mp_clock();
  const unsigned long long mk = ar.mark();
  int *joff = ar.alloc<int>(nj + 1), *jnb = ar.alloc<int>(nj);
  int *newpos = ar.alloc<int>(B.nblk + 1);
  int *errp = ar.alloc<int>(1);
  if (ar.overflow) return false;
  // capacity of every job's block list: min(qLen, tLen) + 2 triples
  int carry = 0;
  for (int b0 = 0; b0 < nj; b0 += kLanes) {
    const int j = b0 + lane;
    const int c = j < nj ? imin(B.jobs[j].qLen, B.jobs[j].tLen) + 2 : 0;
    const int incl = wscan_incl(c);
    if (j < nj) joff[j] = carry + incl - c;
    carry += bcast(incl, kLanes - 1);
  }
  if (lane == 0) joff[nj] = carry;
  wsync();
  uint32_t *jblk = ar.alloc<uint32_t>(3ull * (unsigned long long)(carry > 0 ? carry : 1));
  if (ar.overflow) return false;
  // one job per lane where the thread form applies
  bool pending = false;
  for (int j = lane; j < nj; j += kLanes) {
    const AogJob &J = B.jobs[j];
    if (mp_aog_lane_ok(J.qLen, J.tLen, J.band)) {
      const SeqView &rs = J.str ? C.rd.rc : C.rd.fwd;
      jnb[j] = mp_aog_lane(rs, J.qoff, J.qLen, C.ix.genome, J.toff, J.tLen, O.localMatch, O.localMismatch, O.localIndel, J.band, jblk + 3ull * joff[j]);
    } else { jnb[j] = -2; pending = true; }
  }
  wsync();
  // the rest one after the other on the whole warp
  if (wany(pending)) {
    for (int j = 0; j < nj; j++) {
      if (jnb[j] != -2) continue;
      const AogJob J = B.jobs[j];
      const SeqView &rs = J.str ? C.rd.rc : C.rd.fwd;
      int nb = 0;
      mp_aog_any(rs, J.qoff, J.qLen, C.ix.genome, J.toff, J.tLen, O.localMatch, O.localMismatch, O.localIndel, J.band, ar, jblk + 3ull * joff[j], joff[j + 1] - joff[j], &nb, errp);
      if (nb < 0) return false;
      wsync();
      if (lane == 0) jnb[j] = nb;
      wsync();
    }
  }
  // new position of every block of the list (a placeholder becomes its job's blocks)
  carry = 0;
  for (int b0 = 0; b0 < B.nblk; b0 += kLanes) {
    const int i = b0 + lane;
    int c = 0;
    if (i < B.nblk) c = (B.blk[3 * i + 2] == 0u && B.blk[3 * i + 1] == 0xffffffffu) ? jnb[B.blk[3 * i]] : 1;
    const int incl = wscan_incl(c);
    if (i < B.nblk) newpos[i] = carry + incl - c;
    carry += bcast(incl, kLanes - 1);
  }
  if (lane == 0) newpos[B.nblk] = carry;
  wsync();
  const int total = carry;
  if (total > B.cap) return false;
  uint32_t *tmp = ar.alloc<uint32_t>(3ull * (unsigned long long)(total > 0 ? total : 1));
  if (ar.overflow) return false;
  for (int i = lane; i < B.nblk; i += kLanes) {
    const int at = newpos[i];
    if (!(B.blk[3 * i + 2] == 0u && B.blk[3 * i + 1] == 0xffffffffu)) { tmp[3 * at] = B.blk[3 * i]; tmp[3 * at + 1] = B.blk[3 * i + 1]; tmp[3 * at + 2] = B.blk[3 * i + 2]; }
    else {
      const int j = (int)B.blk[3 * i];
      const AogJob &J = B.jobs[j];
      const uint32_t *src = jblk + 3ull * joff[j];
      for (int x = 0; x < jnb[j]; x++) { tmp[3 * (at + x)] = src[3 * x] + J.addq; tmp[3 * (at + x) + 1] = src[3 * x + 1] + J.addt; tmp[3 * (at + x) + 2] = src[3 * x + 2]; }
    }
  }
  wsync();
  for (int sgi = lane; sgi < B.nseg; sgi += kLanes) B.seg_start[sgi] = newpos[B.seg_start[sgi]];
  for (int i = lane; i < 3 * total; i += kLanes) B.blk[i] = tmp[i];
  wsync();
  B.nblk = total; B.njobs = 0;
  ar.release(mk);
  mp_tick(C, PF_AOG, tk);
  return true;
}

// RefineByLinearAlignment (LocalRefineAlignment.h:144-185): AffineOneGapAlign between two anchors, blocks appended to the current segment
__device__ __noinline__ bool mp_refine_linear(const MpCtx &C, int r, Arena &ar, SegBuild &B, uint32_t curReadEnd, uint32_t curGenomeEnd, uint32_t nextReadStart,
                                        uint32_t nextGenomeStart, int str, int chrom) {
  const MpOpts &O = C.o;
  // SetMatchAndGaps / Matched in GenomePos arithmetic (:92-99)
  const uint32_t a = nextReadStart - curReadEnd + 1u, b = nextGenomeStart - curGenomeEnd + 1u;
  const int m = (int)(a < b ? a : b);
  if (m <= 0) return true;
  const int qLen = (int)(nextReadStart - curReadEnd), tLen = (int)(nextGenomeStart - curGenomeEnd);
  int drift = qLen - tLen; if (drift < 0) drift = -drift;
  const int band = (drift * 2 + 1) < O.localBand ? (drift * 2 + 1) : O.localBand;
  if (B.njobs >= B.cap_jobs && !mp_flush_aog(C, r, ar, B)) return false;
  if (B.nblk >= B.cap) return false;
  if (lane_id() == 0) {
    AogJob J;
    J.qoff = (uint32_t)(C.rd.read_off[r] + curReadEnd); J.toff = (uint32_t)(C.ix.hdr_pos[chrom] + curGenomeEnd); J.addq = curReadEnd; J.addt = curGenomeEnd;
    J.qLen = qLen; J.tLen = tLen; J.band = band; J.str = str;
    B.jobs[B.njobs] = J;
    B.blk[3 * B.nblk] = (uint32_t)B.njobs; B.blk[3 * B.nblk + 1] = 0xffffffffu; B.blk[3 * B.nblk + 2] = 0u;      // placeholder: length 0 at t = 2^32 - 1
  }
  B.njobs++; B.nblk++;
  wsync();
  return true;
}

struct GapState { bool inversion, breakalignment; };

// anchors of an UltimateChain over the extended clusters
__device__ __forceinline__ void uc_get(const ClusterSet &S, const UChain &ch, int i, uint32_t &q, uint32_t &t, int &len) {
  const int a = S.off[ch.cl[i]] + (int)ch.idx[i];
  q = S.q[a]; t = S.t[a]; len = S.len[a];
}

// RefinedAlignmentbtwnAnchors (LocalRefineAlignment.h:203-550)
__device__ __noinline__ bool mp_refined_alignment_btwn(const MpCtx &C, int r, Arena &ar, SegBuild &B, const ClusterSet &S, const UChain &ch, int cur, int next, int str,
                                                 int inv_str, int chrom, int n0, float first_sdp, GapState &gs) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  uint32_t cq, ct, nq, nt; int clen, nlen;
  uc_get(S, ch, cur, cq, ct, clen); uc_get(S, ch, next, nq, nt, nlen);
  if (str == 0) { if (!sb_push_block(B, cq, ct, (uint32_t)clen)) return false; }
  else { if (!sb_push_block(B, L - cq - (uint32_t)clen, ct, (uint32_t)clen)) return false; }
  uint32_t curGenomeEnd, curReadEnd, nextGenomeStart, nextReadStart;
  if (str == 0) { curReadEnd = cq + (uint32_t)clen; nextReadStart = nq; curGenomeEnd = ct + (uint32_t)clen; nextGenomeStart = nt; }
  else { curReadEnd = L - cq; nextReadStart = L - nq - (uint32_t)nlen; curGenomeEnd = ct + (uint32_t)clen; nextGenomeStart = nt; }
  if (!(curGenomeEnd <= nextGenomeStart)) return true;
  const long long read_dist = (long long)(uint32_t)(nextReadStart - curReadEnd), genome_dist = (long long)(uint32_t)(nextGenomeStart - curGenomeEnd);
  const long long mind = read_dist < genome_dist ? read_dist : genome_dist, maxd = read_dist > genome_dist ? read_dist : genome_dist;
  if (!(O.RefineBySDP && mind >= 300)) return mp_refine_linear(C, r, ar, B, curReadEnd, curGenomeEnd, nextReadStart, nextGenomeStart, str, chrom);
  // a band that is not too big, not too small
  const int sv_diag = (int)(maxd - mind);
  int refineSpaceDiag = 0;
  if (O.readType == 3 || O.readType == 2) { const int f = (int)floorf(fmaxf(80.f, __fmul_rn(0.01f, (float)read_dist))); refineSpaceDiag = f < 500 ? f : 500; }
  else { const int f = (int)floorf(fmaxf(100.f, __fmul_rn(0.15f, (float)read_dist))); refineSpaceDiag = f < 2000 ? f : 2000; }
  refineSpaceDiag = 2 * sv_diag > refineSpaceDiag ? 2 * sv_diag : refineSpaceDiag;
  int tK, tW, tMaxFreq = O.localMaxFreq; float minRatio;
  if (maxd < 100) { tK = 6; tW = 5; minRatio = (float)(0.5 / 29.5); }
  else if (maxd < 500) { tK = 9; tW = 7; tMaxFreq = 50; minRatio = (float)(0.5 / 69.1); }
  else { tK = 12; tW = 7; minRatio = (float)(0.5 / 140.2); }
  uint32_t *fq = 0, *ft = 0, *rq = 0, *rt = 0; float identity = 0.0f;
  unsigned long long tk = mp_clock();
  int nfor = mp_refine_space(C, r, ar, tK, tW, refineSpaceDiag, false, tMaxFreq, chrom, nextReadStart, curReadEnd, nextGenomeStart, curGenomeEnd, str, 0, 0, &fq, &ft, &identity);
  tk = mp_tick(C, PF_REFINE_SPACE, tk);
  if (nfor < 0) return false;
  const int minDist = (int)mind;
  uint32_t *bq = fq, *bt = ft; int nb = nfor;
  bool inversion = false;
  // (the number of blocks of the current segment enters the test: the queued linear alignments are run first when the other two conditions hold)
  int cur_seg_blocks = 0;
  if (__fdiv_rn((float)nfor, (float)minDist) < minRatio && identity < 0.8f) {
    if (!mp_flush_aog(C, r, ar, B)) return false;
    cur_seg_blocks = B.nblk - B.seg_start[B.nseg - 1];
  }
  if (__fdiv_rn((float)nfor, (float)minDist) < minRatio && cur_seg_blocks >= 5 && identity < 0.8f) {
    const uint32_t temp = curReadEnd;
    curReadEnd = L - nextReadStart; nextReadStart = L - temp;
    // (tinyOpts.globalW = tinyOpts.localW = opts.localW, Map_lowacc.h:241)
    int nrev = mp_refine_space(C, r, ar, tK, O.localW, refineSpaceDiag, false, tMaxFreq, chrom, nextReadStart, curReadEnd, nextGenomeStart, curGenomeEnd, inv_str, 0, 0, &rq, &rt, &identity);
    if (nrev < 0) return false;
    const double driftRate = (O.readType == 3 || O.readType == 2) ? (double)0.01f : (double)0.10f;
    const double lim = 50.0 > (double)minDist * driftRate ? 50.0 : (double)minDist * driftRate;
    if (nfor == 0 && nrev == 0 && minDist > 500 && (double)sv_diag <= lim) { gs.breakalignment = true; gs.inversion = false; return true; }
    if (identity < 0.8f && __fdiv_rn((float)nrev, (float)minDist) < minRatio) { gs.breakalignment = true; gs.inversion = false; return true; }
    if (nfor >= nrev) {
      bq = fq; bt = ft; nb = nfor; inversion = false;
      const uint32_t t2 = curReadEnd; curReadEnd = L - nextReadStart; nextReadStart = L - t2;
    } else { bq = rq; bt = rt; nb = nrev; inversion = true; }
  }
  gs.inversion = inversion;
  if (nb == 0) return mp_refine_linear(C, r, ar, B, curReadEnd, curGenomeEnd, nextReadStart, nextGenomeStart, str, chrom);
  // LinearExtend (sorted), the two flanking anchors, TrimOverlappedAnchors, SparseDP_ForwardOnly, RemovePairedIndels
  MpKey *keys = ar.alloc<MpKey>((unsigned long long)next_pow2(nb));
  uint32_t *sq = ar.alloc<uint32_t>(nb), *stt = ar.alloc<uint32_t>(nb);
  uint32_t *eq = ar.alloc<uint32_t>(nb + 3), *et = ar.alloc<uint32_t>(nb + 3); int *el = ar.alloc<int>(nb + 3), *tidx = ar.alloc<int>(nb + 3);
  int *ne_p = ar.alloc<int>(2);
  if (ar.overflow) return false;
  for (int i = lane; i < nb; i += kLanes) { keys[i].k = (unsigned long long)((long long)bq[i] - (long long)bt[i] + (1ll << 33)); keys[i].q = bq[i]; keys[i].idx = (uint32_t)i; }
  wsync();
  mp_sort_keys(keys, nb);
  for (int i = lane; i < nb; i += kLanes) { sq[i] = bq[keys[i].idx]; stt[i] = bt[keys[i].idx]; }
  wsync();
  if (lane == 0) {
    int ne = mp_linear_extend(C, C.rd.read_off[r], L, chrom, sq, stt, nb, 0, tK, eq, et, el, 0);
    if (!inversion) {
      eq[ne] = nextReadStart; et[ne] = nextGenomeStart; el[ne] = nlen; ne++;
      eq[ne] = curReadEnd - (uint32_t)clen; et[ne] = curGenomeEnd - (uint32_t)clen; el[ne] = clen; ne++;
    }
    mp_trim_overlapped(eq, et, el, ne, 0, 50, false, tidx);
    ne_p[0] = ne;
  }
  wsync();
  const int ne = ne_p[0];
  uint32_t *bchain = ar.alloc<uint32_t>(ne + 1);
  float *ivp = ar.alloc<float>(1);
  if (ar.overflow) return false;
  SdpAnchors A; A.q = eq; A.t = et; A.len = el; A.nfrag = ne; A.cl_off = 0; A.cl_strand = 0; A.ncl = 0;
  tk = mp_clock();
  int nbc = sdp_forward_only(A, 2, *C.pwl, ar, bchain, ivp);
  tk = mp_tick(C, PF_SDP3, tk);
  if (nbc < 0) return false;
  wsync();
  const float inv_value = *ivp;
  // RemovePairedIndels(ExtendBtwnPairs, BtwnChain, lengths) (Chain.h:754-822): chain filter mode 3 on the chained anchors
  if (nbc >= 2) {
    const unsigned long long mk = ar.mark();
    uint32_t *fq2 = ar.alloc<uint32_t>(nbc), *ft2 = ar.alloc<uint32_t>(nbc), *fl2 = ar.alloc<uint32_t>(nbc);
    uint8_t *fs = ar.alloc<uint8_t>(nbc), *keep = ar.alloc<uint8_t>(nbc);
    int32_t *sv = ar.alloc<int32_t>(nbc), *svp = ar.alloc<int32_t>(nbc), *svg = ar.alloc<int32_t>(nbc);
    unsigned long long *off = ar.alloc<unsigned long long>(2);
    if (ar.overflow) return false;
    for (int i = lane; i < nbc; i += kLanes) { const int f = (int)bchain[i]; fq2[i] = eq[f]; ft2[i] = et[f]; fl2[i] = (uint32_t)el[f]; fs[i] = 0; }
    if (lane == 0) { off[0] = 0; off[1] = (unsigned long long)nbc; }
    wsync();
    int mkeep = 0;
    if (lane == 0) {
      ChainfBatch b; b.n_chains = 1; b.mode = 3; b.off = off; b.q = fq2; b.t = ft2; b.len = fl2; b.strand = fs; b.keep = keep; b.sv = sv; b.svpos = svp; b.svg = svg;
      chainf_one(b, 0);
      for (int i = 0; i < nbc; i++) if (keep[i]) bchain[mkeep++] = bchain[i];
    }
    wsync();
    nbc = bcast(mkeep, 0);
    ar.release(mk);
  }
  // ligate the gaps of the local chain with linear alignments (:497-538)
  uint32_t btc_curReadEnd = curReadEnd, btc_curGenomeEnd = curGenomeEnd;
  int btc_end = nbc - 1, btc_start = 0;
  if (nbc > 0 && (int)bchain[nbc - 1] == ne - 1) btc_end = nbc - 2;
  if (nbc > 0 && (int)bchain[0] == ne - 2) btc_start = 1;
  if (inversion) {
    // the segment so far keeps strand `str`; a new supplementary segment on the inverted strand takes the anchors of the gap
    if (!sb_new_segment(B, inv_str, chrom, nbc, nbc, 1, inv_value)) return false;
  }
  const int seg_str = inversion ? inv_str : str;    // (the reference passes `str` to RefineByLinearAlignment here: see below)
  (void)seg_str; (void)n0; (void)first_sdp;
  for (int btc = btc_end; btc >= btc_start; btc--) {
    const int f = (int)bchain[btc];
    const uint32_t gq = eq[f], gt = et[f]; const int gl = el[f];
    // RefineByLinearAlignment(..., str, ...) reads strands[str] even when the anchors were found on the inverted strand (LocalRefineAlignment.h:518)
    if (!mp_refine_linear(C, r, ar, B, btc_curReadEnd, btc_curGenomeEnd, gq, gt, str, chrom)) return false;
    if (!sb_push_block(B, gq, gt, (uint32_t)gl)) return false;
    btc_curReadEnd = gq + (uint32_t)gl; btc_curGenomeEnd = gt + (uint32_t)gl;
  }
  if (nextGenomeStart > btc_curGenomeEnd && nextReadStart > btc_curReadEnd)
    if (!mp_refine_linear(C, r, ar, B, btc_curReadEnd, btc_curGenomeEnd, nextReadStart, nextGenomeStart, str, chrom)) return false;
  return true;
}

// LocalRefineAlignment, pure-matches form (LocalRefineAlignment.h:884-1030), for the ultimate chains uc[0..nuc) of one chain p
// min_n: 2 for the pure-matches form (`size() <= 1` is skipped), 1 for the high-accuracy form (:553-768, only an empty chain is skipped; the two bodies are
// otherwise the same)
__device__ __noinline__ bool mp_local_refine_alignment(const MpCtx &C, int r, Arena &ar, SegBuild &B, const ClusterSet &S, UChain *uc, int nuc, int LSC, int min_n) {
  const uint32_t L = C.rd.read_len[r];
  for (int st = 0; st < nuc; st++) {
    const UChain &ch = uc[st];
    if (ch.n < min_n) continue;
    const int start = 0, end = ch.n - 1;
    const int str = S.strand[ch.cl[start]] != 0 ? 1 : 0;
    const int chrom = S.chrom[ch.cl[start]];
    if (!sb_new_segment(B, str, chrom, ch.NumOfAnchors0, ch.NumOfAnchors1, st != LSC ? 1 : 0, ch.FirstSDPValue)) return false;
    GapState gs; gs.inversion = false; gs.breakalignment = false;
    uint32_t q, t; int len;
    if (str == 0) {
      int last = end, fl = end; const int inv_str = 1;
      while (fl > start) {
        if (!mp_refined_alignment_btwn(C, r, ar, B, S, ch, fl, fl - 1, str, inv_str, chrom, ch.NumOfAnchors0, ch.FirstSDPValue, gs)) return false;
        if (gs.inversion || gs.breakalignment) {
          // close the current segment (UpdateParameters: strand inv_str after an inversion) and open the next supplementary one
          if (lane_id() == 0) { B.seg_strand[B.nseg - 1] = gs.inversion ? inv_str : str; B.seg_n0[B.nseg - 1] = ch.NumOfAnchors0; B.seg_n1[B.nseg - 1] = last - fl; }
          if (!sb_new_segment(B, str, chrom, ch.NumOfAnchors0, 0, 1, ch.FirstSDPValue)) return false;
          last = fl; gs.inversion = false; gs.breakalignment = false;
        }
        fl--;
      }
      if (lane_id() == 0) { B.seg_n0[B.nseg - 1] = ch.NumOfAnchors0; B.seg_n1[B.nseg - 1] = last - fl; }
      uc_get(S, ch, start, q, t, len);
      if (!sb_push_block(B, q, t, (uint32_t)len)) return false;
    } else {
      int last = start, fl = start; const int inv_str = 0;
      while (fl < end) {
        if (!mp_refined_alignment_btwn(C, r, ar, B, S, ch, fl, fl + 1, str, inv_str, chrom, ch.NumOfAnchors0, ch.FirstSDPValue, gs)) return false;
        if (gs.inversion || gs.breakalignment) {
          if (lane_id() == 0) { B.seg_strand[B.nseg - 1] = gs.inversion ? inv_str : str; B.seg_n0[B.nseg - 1] = ch.NumOfAnchors0; B.seg_n1[B.nseg - 1] = fl - last; }
          if (!sb_new_segment(B, str, chrom, ch.NumOfAnchors0, 0, 1, ch.FirstSDPValue)) return false;
          last = fl; gs.inversion = false; gs.breakalignment = false;
        }
        fl++;
      }
      if (lane_id() == 0) { B.seg_n0[B.nseg - 1] = ch.NumOfAnchors0; B.seg_n1[B.nseg - 1] = fl - last; }
      uc_get(S, ch, end, q, t, len);
      if (!sb_push_block(B, L - q - (uint32_t)len, t, (uint32_t)len)) return false;
    }
    // the last segment: UpdateParameters(str, ...)
    if (lane_id() == 0) B.seg_strand[B.nseg - 1] = str;
    wsync();
  }
  return mp_flush_aog(C, r, ar, B);
}

}  // namespace mp
}  // namespace lra
