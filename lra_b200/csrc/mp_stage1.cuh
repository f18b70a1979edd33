// Mapper worker, first half of MapRead_lowacc: seeding (MapRead.h:169-203), CleanMatches (Clustering.h:1840-1908), the first LinearExtend
// (Map_lowacc.h:132-153, LinearExtend.h:658-716), the first SparseDP + RemoveSpuriousJump (Map_lowacc.h:184-191).
// Serial scans whose state runs from anchor to anchor (minimizer window, the CompareLists galloping, CleanOffDiagonal's run counters,
// LinearExtend's m / n cursors) are replayed by lane 0 with the pinned device routines of the stage kernels (seed_kernels.cuh, cod_kernels.cuh);
// sorts, strand tests, compactions and the sparse DP use all lanes.
#pragma once
#include "mm_range.cuh"
#include "mp_compare.cuh"
#include "mp_types.cuh"
#include "mp_sdp_driver.cuh"
#include "seed_kernels.cuh"
#include "cod_kernels.cuh"
#include "chainf_kernels.cuh"

namespace lra {
namespace mp {

struct MpCtx {
  MpOpts o;
  MpIndex ix;
  MpReads rd;
  const Pwl *pwl;
  unsigned long long *prof;               // optional [n_warps][kProfStages] clock64 totals per stage (lane 0)
};

// ---- stage profile of the worker kernel (clock64 deltas accumulated per warp; summed on the host)
enum { PF_MINIMIZERS = 0, PF_COMPARE, PF_STRAND_CLEAN, PF_LEXT1, PF_SDP1, PF_SPLIT, PF_REFINE_SPLIT, PF_REFINE_BTWN, PF_LEXT2, PF_SDP2, PF_LOCAL_REFINE, PF_AOG, PF_REFINE_SPACE,
       PF_SDP3, PF_OUTPUT, PF_BARRIER, PF_SDP_POINTS, PF_SDP_DIVIDE, PF_SDP_PROCESS, PF_SDP_TRACE, kProfStages = 24 };
__device__ __forceinline__ unsigned long long mp_clock() {
#ifdef LRA_EMU
  return 0ull;
#else
  return (unsigned long long)clock64();
#endif
}
__device__ __forceinline__ unsigned long long mp_tick(const MpCtx &C, int stage, unsigned long long t0) {
#ifdef LRA_EMU
  (void)C; (void)stage; (void)t0; return 0ull;
#else
  const unsigned long long t1 = (unsigned long long)clock64();
  if (C.prof && (threadIdx.x & 31u) == 0) {
    const unsigned wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    C.prof[(unsigned long long)wid * kProfStages + stage] += t1 - t0;
  }
  return t1;
#endif
}

struct MpMatch { uint32_t q, t; unsigned long long qt; };
struct MpKey { unsigned long long k; uint32_t q, idx; };

__device__ __forceinline__ uint32_t contig_len(const MpIndex &ix, int c) { return (uint32_t)(ix.hdr_pos[c + 1] - ix.hdr_pos[c]); }

// sort `n` records by (k, q, idx) and leave the permutation in keys[].idx; keys has room for next_pow2(n)
__device__ __noinline__ void mp_sort_keys(MpKey *keys, int n) {
  const int P = next_pow2(n > 0 ? n : 1);
  for (int i = n + lane_id(); i < P; i += kLanes) { keys[i].k = ~0ull; keys[i].q = 0xffffffffu; keys[i].idx = 0xffffffffu; }
  wsort_pow2(keys, P, [](const MpKey &a, const MpKey &b) { if (a.k != b.k) return a.k < b.k; if (a.q != b.q) return a.q < b.q; return a.idx < b.idx; });
}

// Checkbp (LinearExtend.h:50-85): extend after anchor `cur` by exact base comparison until `next`; t chromosome-relative
__device__ __noinline__ void mp_checkbp(const MpCtx &C, unsigned long long roff, uint32_t readLen, unsigned long long coff, uint32_t chromLen, uint32_t cq, uint32_t ct,
                                  uint32_t nq, uint32_t nt, int strand, int K, uint32_t &qe, uint32_t &te) {
  uint32_t curQ, curT, nextQ, nextT;
  if (strand == 0) {
    curQ = cq + K; curT = chromLen < ct + K ? chromLen : ct + K;
    nextQ = nq; nextT = chromLen < nt ? chromLen : nt;
    while (curQ < readLen && curT < chromLen && nextQ > curQ && nextT > curT && seq_code(C.ix.genome, coff + curT) == seq_code(C.rd.fwd, roff + curQ)) { curQ++; curT++; }
  } else {
    curQ = cq + K; curT = (chromLen - 1u) < (ct - 1u) ? (chromLen - 1u) : (ct - 1u);
    nextQ = nq; nextT = (chromLen - 1u) < (nt + K - 1u) ? (chromLen - 1u) : (nt + K - 1u);
    while (curQ < readLen && nextQ > curQ && nextT < curT && curT < chromLen && seq_code(C.ix.genome, coff + curT) == seq_code(C.rd.fwd, roff + curQ)) { curQ++; curT--; }
  }
  qe = curQ; te = curT;
}

// LinearExtend, GenomePairs overload (LinearExtend.h:658-716), sorting excluded; appends to (oq, ot, ol) from index o0; returns the new count.
// Serial (called by one lane).
__device__ __noinline__ int mp_linear_extend(const MpCtx &C, unsigned long long roff, uint32_t readLen, int chrom, const uint32_t *q, const uint32_t *t, int np, int strand,
                                       int K, uint32_t *oq, uint32_t *ot, int *ol, int o0) {
  const unsigned long long coff = C.ix.hdr_pos[chrom];
  const uint32_t chromLen = contig_len(C.ix, chrom);
  int n = 1, m = 0, o = o0;
  while (n < np) {
    long long curDiag, nextDiag;
    if (strand == 0) { curDiag = (long long)q[n - 1] - (long long)t[n - 1]; nextDiag = (long long)q[n] - (long long)t[n]; }
    else { curDiag = (long long)q[n - 1] + (long long)t[n - 1]; nextDiag = (long long)q[n] + (long long)t[n]; }
    if (curDiag == nextDiag) {
      if (q[n] < q[n - 1] + (uint32_t)K) n++;
      else {
        uint32_t qe, te;
        mp_checkbp(C, roff, readLen, coff, chromLen, q[n - 1], t[n - 1], q[n], t[n], strand, K, qe, te);
        if (strand == 0 && qe == q[n] && te == t[n]) n++;
        else if (strand == 1 && qe == q[n] && te == t[n] + (uint32_t)K - 1u) n++;
        else {
          oq[o] = q[m]; ot[o] = strand == 0 ? t[m] : te + 1u; ol[o] = (int)(qe - q[m]); o++;
          m = n; n++;
        }
      }
    } else {
      oq[o] = q[m]; ot[o] = strand == 0 ? t[m] : t[n - 1]; ol[o] = (int)(q[n - 1] + (uint32_t)K - q[m]); o++;
      m = n; n++;
    }
  }
  if (n == np) { oq[o] = q[m]; ot[o] = strand == 0 ? t[m] : t[n - 1]; ol[o] = (int)(q[n - 1] + (uint32_t)K - q[m]); o++; }
  return o;
}

// The same on the whole warp.  Whether the run of co-diagonal anchors ends between anchors n-1 and n depends on that pair alone (same diagonal?
// overlapping K-mers? does the exact extension from n-1 reach n?), so every lane decides one pair; the start m of the run a break closes is the
// previous break (a warp max-scan), the output slot the number of earlier breaks.
__device__ __noinline__ int mp_linear_extend_warp(const MpCtx &C, unsigned long long roff, uint32_t readLen, int chrom, const uint32_t *q, const uint32_t *t, int np, int strand,
                                            int K, uint32_t *oq, uint32_t *ot, int *ol, int o0) {
  if (np <= 0) return o0;
  const int lane = lane_id();
  const unsigned long long coff = C.ix.hdr_pos[chrom];
  const uint32_t chromLen = contig_len(C.ix, chrom);
  int o = o0, m_carry = 0;
  for (int b = 1; b < np; b += kLanes) {
    const int n = b + lane;
    bool brk = false, typeA = false; uint32_t qe = 0, te = 0;
    if (n < np) {
      long long curDiag, nextDiag;
      if (strand == 0) { curDiag = (long long)q[n - 1] - (long long)t[n - 1]; nextDiag = (long long)q[n] - (long long)t[n]; }
      else { curDiag = (long long)q[n - 1] + (long long)t[n - 1]; nextDiag = (long long)q[n] + (long long)t[n]; }
      if (curDiag != nextDiag) brk = true;
      else if (!(q[n] < q[n - 1] + (uint32_t)K)) {
        mp_checkbp(C, roff, readLen, coff, chromLen, q[n - 1], t[n - 1], q[n], t[n], strand, K, qe, te);
        const bool reach = strand == 0 ? (qe == q[n] && te == t[n]) : (qe == q[n] && te == t[n] + (uint32_t)K - 1u);
        if (!reach) { brk = true; typeA = true; }
      }
    }
    const unsigned mk = ballot(brk);
    // start of the run this break closes: the n of the previous break (in this chunk: highest lower lane with a break; else the carry)
    int m = m_carry;
    { const unsigned lower = mk & lanemask_lt(); if (lower) m = b + (31 - __clz((int)lower)); }
    if (brk) {
      const int at = o + __popc(mk & lanemask_lt());
      oq[at] = q[m];
      if (typeA) { ot[at] = strand == 0 ? t[m] : te + 1u; ol[at] = (int)(qe - q[m]); }
      else { ot[at] = strand == 0 ? t[m] : t[n - 1]; ol[at] = (int)(q[n - 1] + (uint32_t)K - q[m]); }
    }
    o += __popc(mk);
    if (mk) m_carry = b + (31 - __clz((int)mk));
  }
  if (lane == 0) { oq[o] = q[m_carry]; ot[o] = strand == 0 ? t[m_carry] : t[np - 1]; ol[o] = (int)(q[np - 1] + (uint32_t)K - q[m_carry]); }
  o++;
  wsync();
  return o;
}

// DecideCoordinates (LinearExtend.h:105-128) over anchors [a0, a1) of a cluster set entry
__device__ __noinline__ void mp_decide_coordinates(ClusterSet &S, int c, int strand, int chrom, float freq) {
  const int a0 = S.off[c], a1 = S.off[c + 1];
  if (a1 == a0) return;
  uint32_t qS = S.q[a0], qE = qS + (uint32_t)S.len[a0], tS = S.t[a0], tE = tS + (uint32_t)S.len[a0];
  for (int n = a0 + 1; n < a1; n++) {
    qS = S.q[n] < qS ? S.q[n] : qS; qE = S.q[n] + (uint32_t)S.len[n] > qE ? S.q[n] + (uint32_t)S.len[n] : qE;
    tS = S.t[n] < tS ? S.t[n] : tS; tE = S.t[n] + (uint32_t)S.len[n] > tE ? S.t[n] + (uint32_t)S.len[n] : tE;
  }
  S.qS[c] = qS; S.qE[c] = qE; S.tS[c] = tS; S.tE[c] = tE; S.strand[c] = strand; S.chrom[c] = chrom; S.freq[c] = freq;
}

__device__ __noinline__ bool mp_alloc_clusterset(ClusterSet &S, Arena &ar, int cap_cl, int cap_a, bool with_len) {
  S.cap_cl = cap_cl; S.cap_a = cap_a; S.ncl = 0;
  S.q = ar.alloc<uint32_t>(cap_a + 1); S.t = ar.alloc<uint32_t>(cap_a + 1); S.len = with_len ? ar.alloc<int>(cap_a + 1) : (int *)0;
  S.off = ar.alloc<int>(cap_cl + 2);
  S.qS = ar.alloc<uint32_t>(cap_cl + 1); S.qE = ar.alloc<uint32_t>(cap_cl + 1); S.tS = ar.alloc<uint32_t>(cap_cl + 1); S.tE = ar.alloc<uint32_t>(cap_cl + 1);
  S.strand = ar.alloc<int>(cap_cl + 1); S.chrom = ar.alloc<int>(cap_cl + 1); S.freq = ar.alloc<float>(cap_cl + 1);
  return !ar.overflow;
}

// the keep mask of a chain filter (Chain.h:546-960, chainf_kernels.cuh) over the anchors of a chain; serial.  scratch: 3 * n ints + n bytes + SoA copies
__device__ __noinline__ void mp_chain_filter(int mode, const ClusterSet &S, UChain &ch, Arena &ar, bool compact_link) {
  const int n = ch.n;
  if (n < 2) return;
  const unsigned long long mk = ar.mark();
  uint32_t *q = ar.alloc<uint32_t>(n), *t = ar.alloc<uint32_t>(n), *len = ar.alloc<uint32_t>(n);
  uint8_t *st = ar.alloc<uint8_t>(n), *keep = ar.alloc<uint8_t>(n);
  int32_t *sv = ar.alloc<int32_t>(n), *svp = ar.alloc<int32_t>(n), *svg = ar.alloc<int32_t>(n);
  unsigned long long *off = ar.alloc<unsigned long long>(2);
  if (ar.overflow) { ar.release(mk); return; }
  for (int i = lane_id(); i < n; i += kLanes) {
    const int a = S.off[ch.cl[i]] + (int)ch.idx[i];
    q[i] = S.q[a]; t[i] = S.t[a]; len[i] = (uint32_t)S.len[a]; st[i] = (uint8_t)(S.strand[ch.cl[i]] != 0);
  }
  if (lane_id() == 0) { off[0] = 0; off[1] = (unsigned long long)n; }
  wsync();
  int m = 0;
  if (lane_id() == 0) {
    ChainfBatch b; b.n_chains = 1; b.mode = mode; b.off = off; b.q = q; b.t = t; b.len = len; b.strand = st; b.keep = keep; b.sv = sv; b.svpos = svp; b.svg = svg;
    chainf_one(b, 0);
    for (int i = 0; i < n; i++) {
      if (keep[i]) {
        ch.idx[m] = ch.idx[i]; ch.cl[m] = ch.cl[i];
        if (compact_link && ch.nlink > 0 && m >= 1) ch.link[m - 1] = ch.link[i - 1];
        m++;
      }
    }
  }
  wsync();
  m = bcast(m, 0);
  ch.n = m;
  if (compact_link && ch.nlink > 0) ch.nlink = m - 1;     // (m == 0 would be resize(-1) in the reference: not reachable, at least two anchors survive)
  ar.release(mk);
}

// ---- a2 warp-parallel: StoreMinimizers of the read, one lane per chunk of 256 loop steps (mm_range.cuh), compacted in place.  Returns the number of
// minimizers in (mm_t, mm_p), or -1 when a chunk's entry state is not provably the sequential one (the caller falls back to the literal scan).
__device__ __noinline__ int mp_minimizers_warp(const SeqView &seq, unsigned long long roff, uint32_t L, int k, int w, unsigned long long *mm_t, uint32_t *mm_p) {
  const int lane = lane_id();
  if (L < (uint32_t)(w + k - 1)) return 0;
  const int CH = 256, WARM = 64;
  const int steps = (int)(L - (uint32_t)k + 1u);
  const int nch = (steps + CH - 1) / CH;
  int total = 0;
  for (int c0 = 0; c0 < nch; c0 += kLanes) {
    const int c = c0 + lane;
    uint32_t n = 0; bool unc = false;
    if (c < nch) {
      const uint32_t pb = (uint32_t)c * CH, pe = c == nch - 1 ? 0xffffffffu : pb + CH;
      const uint32_t s0 = pb > (uint32_t)WARM ? pb - WARM : 0u;
      bool certain;
      n = gidx_scan_range<true>(seq, roff, L, k, w, s0, pb, pe, mm_t + pb, mm_p + pb, 0u, certain);
      if (n == 0xffffffffu) { unc = true; n = 0; }
    }
    if (wany(unc)) return -1;
    wsync();
    // move the chunks of this round down to the end of the list, in order (destination never ahead of the source)
    for (int j = 0; j < kLanes && c0 + j < nch; j++) {
      const int nj = (int)bcast(n, j), src = (c0 + j) * CH;
      if (src != total) {
        for (int i0 = 0; i0 < nj; i0 += kLanes) {
          const int i = i0 + lane;
          unsigned long long tv = 0; uint32_t pv = 0;
          if (i < nj) { tv = mm_t[src + i]; pv = mm_p[src + i]; }
          wsync();
          if (i < nj) { mm_t[total + i] = tv; mm_p[total + i] = pv; }
          wsync();
        }
      }
      total += nj;
    }
  }
  return total;
}

// ---- a3 warp-parallel: std::sort(readmm) on the masked tuple (MapRead.h:185).  With all keys distinct every sort gives the same list; equal keys
// (the same k-mer twice in the read) leave introsort's order in the reference, so that case is sent back to the literal replay (returns false,
// the arrays untouched).
__device__ __noinline__ bool mp_sort_minimizers_warp(Arena &ar, unsigned long long *mm_t, uint32_t *mm_p, int n) {
  if (n < 2) return true;
  const int lane = lane_id();
  const unsigned long long mk = ar.mark();
  MpKey *keys = ar.alloc<MpKey>((unsigned long long)next_pow2(n));
  unsigned long long *tt = ar.alloc<unsigned long long>(n);
  if (ar.overflow) { ar.overflow = 0; ar.release(mk); return false; }
  for (int i = lane; i < n; i += kLanes) { keys[i].k = mm_t[i] & kForMask; keys[i].q = mm_p[i]; keys[i].idx = (uint32_t)i; tt[i] = mm_t[i]; }
  wsync();
  mp_sort_keys(keys, n);
  bool dup = false;
  for (int i = 1 + lane; i < n; i += kLanes) dup = dup || keys[i].k == keys[i - 1].k;
  if (wany(dup)) { ar.release(mk); return false; }
  for (int i = lane; i < n; i += kLanes) { mm_t[i] = tt[keys[i].idx]; mm_p[i] = keys[i].q; }
  wsync();
  ar.release(mk);
  return true;
}

// ---- seeding + the sort / CleanOffDiagonal of each strand's matches for one read (MapRead.h:169-203; the common head of CleanMatches, Clustering.h:1840-1908,
// and MatchesToFineClusters, :1555-1676).  Returns MP_OK with `raw`: the cleaned matches of strand 0 then strand 1 (global t) and the clusters
// CleanOffDiagonal cut them into (ranges, boxes, anchorfreq; chrom only under bypassClustering), n_raw_a anchors in all.
__device__ __noinline__ int mp_seed_clean(const MpCtx &C, int r, Arena &ar, ClusterSet &raw, int &n_raw_a_out) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const unsigned long long roff = C.rd.read_off[r];
  const uint32_t L = C.rd.read_len[r];
  // ---- a2 / a3: minimizers of the read, std::sort (MapRead.h:181-185)
  unsigned long long *mm_t = ar.alloc<unsigned long long>((unsigned long long)L + 2);
  uint32_t *mm_p = ar.alloc<uint32_t>((unsigned long long)L + 2);
  if (ar.overflow) return MP_ERR_ARENA;
  unsigned long long tk = mp_clock();
  int n_mm = 0;
  n_mm = mp_minimizers_warp(C.rd.fwd, roff, L, O.globalK, O.globalW, mm_t, mm_p);
  if (n_mm < 0) {           // low-complexity stretch: literal scan
    if (lane == 0) n_mm = (int)mm_scan<true>(C.rd.fwd, roff, L, O.globalK, O.globalW, mm_t, mm_p);
    wsync();
    n_mm = bcast(n_mm, 0);
  }
  if (!mp_sort_minimizers_warp(ar, mm_t, mm_p, n_mm)) {
    if (lane == 0) mm_sort(MmRef{mm_t, mm_p}, (long)n_mm);
    wsync();
  }
  tk = mp_tick(C, PF_MINIMIZERS, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  // ---- a4: CompareLists against the global index (MapRead.h:190); matches land in the open end of the arena
  const unsigned long long top0 = (ar.top + 15ull) & ~15ull;
  const unsigned long long room = ar.cap > top0 ? (ar.cap - top0) / sizeof(MpMatch) : 0ull;
  MpMatch *M = (MpMatch *)(ar.base + top0);
  long long n_match = 0;
  {
    // the plan's arrays live above the match list's start only until the matches are written: plan first (arena), then the matches behind it
    const unsigned long long mkp = ar.mark();
    CmpPlan *plan = ar.alloc<CmpPlan>((unsigned long long)n_mm + 2);
    if (ar.overflow) return MP_ERR_ARENA;
    const int np = mp_compare_plan(mm_t, n_mm, C.ix.idx_t, C.ix.n_idx, (long long)O.globalMaxFreq, ar, plan);
    if (np < 0) return MP_ERR_ARENA;
    const unsigned long long top1 = (ar.top + 15ull) & ~15ull;
    const unsigned long long room1 = ar.cap > top1 ? (ar.cap - top1) / sizeof(MpMatch) : 0ull;
    MpMatch *M1 = (MpMatch *)(ar.base + top1);
    const uint32_t *idx_pos = C.ix.idx_pos;
    n_match = mp_compare_expand(plan, np, [&](long long slot, int qi, uint32_t ti) {
      if ((unsigned long long)slot < room1) { MpMatch x; x.q = mm_p[qi]; x.t = idx_pos[ti]; x.qt = mm_t[qi]; M1[slot] = x; }
    });
    if ((unsigned long long)n_match > room1) return MP_ERR_ARENA;
    // slide the matches down to where the plan started (M == the open end of the arena before the plan)
    ar.release(mkp);
    if ((unsigned char *)M1 != (unsigned char *)M) {
      for (long long b0 = 0; b0 < n_match; b0 += kLanes) {
        const long long i = b0 + lane;
        MpMatch x; x.q = 0; x.t = 0; x.qt = 0;
        if (i < n_match) x = M1[i];
        wsync();
        if (i < n_match) M[i] = x;
        wsync();
      }
    }
  }
  wsync();
  tk = mp_tick(C, PF_COMPARE, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  if ((unsigned long long)n_match > room / 4) return MP_ERR_ARENA;        // leave room for the stages below
  ar.alloc<MpMatch>((unsigned long long)n_match);
  const int NM = (int)n_match;
  if (NM == 0) return MP_UNALIGNED;
  // ---- a5: SeparateMatchesByStrand (MapRead.h:109-150): strncmp(read + q, genome + t, K) == 0 -> forward
  uint8_t *mstr = ar.alloc<uint8_t>(NM);
  if (ar.overflow) return MP_ERR_ARENA;
  for (int i = lane; i < NM; i += kLanes) {
    int differ = 0;
    for (int j = 0; j < O.globalK && !differ; j++) differ = seq_code(C.rd.fwd, roff + M[i].q + j) != seq_code(C.ix.genome, (unsigned long long)M[i].t + j);
    mstr[i] = (uint8_t)differ;
  }
  wsync();
  // ---- CleanMatches per strand (Clustering.h:1840-1908): DiagonalSort / AntiDiagonalSort, CleanOffDiagonal, clusters with their matches
  if (!mp_alloc_clusterset(raw, ar, NM, NM, false)) return MP_ERR_ARENA;
  if (lane == 0) raw.off[0] = 0;
  int n_raw_a = 0;
  for (int s = 0; s < 2; s++) {
    const unsigned long long mk = ar.mark();
    // gather this strand's matches (order is irrelevant: the sort key below is total up to identical records)
    MpKey *keys = ar.alloc<MpKey>((unsigned long long)next_pow2(NM));
    if (ar.overflow) return MP_ERR_ARENA;
    int ns = 0;
    for (int b = 0; b < NM; b += kLanes) {
      const int i = b + lane;
      const bool take = i < NM && mstr[i] == s;
      const unsigned mk2 = ballot(take);
      if (take) {
        MpKey k;
        if (s == 0) k.k = (unsigned long long)((long long)M[i].q - (long long)M[i].t + (1ll << 33));
        else k.k = (unsigned long long)(uint32_t)(M[i].q + M[i].t);
        k.q = M[i].q; k.idx = (uint32_t)i;
        keys[ns + __popc(mk2 & lanemask_lt())] = k;
      }
      ns += __popc(mk2);
    }
    wsync();
    if (ns == 0) { ar.release(mk); continue; }
    mp_sort_keys(keys, ns);
    uint32_t *sq = ar.alloc<uint32_t>(ns), *stt = ar.alloc<uint32_t>(ns);
    unsigned long long *sqt = ar.alloc<unsigned long long>(ns);
    uint8_t *keep = ar.alloc<uint8_t>(ns), *flags = ar.alloc<uint8_t>(3ull * ns + 3), *hused = ar.alloc<uint8_t>(4ull * ns + 8);
    float *freq = ar.alloc<float>(ns), *clf = ar.alloc<float>(ns);
    int32_t *cnt = ar.alloc<int32_t>(ns), *cl = ar.alloc<int32_t>(7ull * ns), *ncl_p = ar.alloc<int32_t>(2);
    unsigned long long *hkeys = ar.alloc<unsigned long long>(4ull * ns + 8), *off = ar.alloc<unsigned long long>(2);
    uint8_t *sstr = ar.alloc<uint8_t>(2);
    if (ar.overflow) return MP_ERR_ARENA;
    for (int i = lane; i < ns; i += kLanes) { const MpMatch &x = M[keys[i].idx]; sq[i] = x.q; stt[i] = x.t; sqt[i] = x.qt; }
    if (lane == 0) { off[0] = 0; off[1] = (unsigned long long)ns; sstr[0] = (uint8_t)s; }
    wsync();
    if (lane == 0) {
      CodBatch b;
      b.n_lists = 1; b.off = off; b.q = sq; b.t = stt; b.qt = sqt; b.strand = sstr;
      b.o.cleanMaxDiag = O.cleanMaxDiag; b.o.minDiagCluster = O.minDiagCluster; b.o.bypassClustering = O.bypassClustering; b.o.cleanClustersize = O.cleanClustersize;
      b.o.SecondCleanMinDiagCluster = O.SecondCleanMinDiagCluster; b.o.punish_anchorfreq = O.punish_anchorfreq; b.o.anchorPerlength = O.anchorPerlength;
      b.o.SecondCleanMaxDiag = O.SecondCleanMaxDiag; b.o.ExtractDiagonalFromClean = 1; b.o.globalK = O.globalK;
      b.hdr_pos = C.ix.hdr_pos; b.n_hdr = C.ix.n_hdr;
      b.keep = keep; b.freq = freq; b.cnt = cnt; b.cl = cl; b.cl_freq = clf; b.n_cl = ncl_p; b.flags = flags; b.hkeys = hkeys; b.hused = hused;
      cod_one(b, 0);
      // the cleaned list is the concatenation of the clusters' matches (cluster ranges index the compacted list)
      int m = 0;
      for (int i = 0; i < ns; i++) if (keep[i]) { raw.q[n_raw_a + m] = sq[i]; raw.t[n_raw_a + m] = stt[i]; m++; }
      const int nc = ncl_p[0];
      for (int c = 0; c < nc; c++) {
        const int k = raw.ncl + c;
        raw.off[k + 1] = n_raw_a + cl[7 * c + 1];
        raw.qS[k] = (uint32_t)cl[7 * c + 2]; raw.qE[k] = (uint32_t)cl[7 * c + 3]; raw.tS[k] = (uint32_t)cl[7 * c + 4]; raw.tE[k] = (uint32_t)cl[7 * c + 5];
        raw.chrom[k] = cl[7 * c + 6]; raw.strand[k] = s; raw.freq[k] = clf[c];
      }
      ncl_p[1] = m;
    }
    wsync();
    raw.ncl += ncl_p[0]; n_raw_a += ncl_p[1];
    wsync();
    ar.release(mk);
  }
  tk = mp_tick(C, PF_STRAND_CLEAN, tk);
  n_raw_a_out = n_raw_a;
  return MP_OK;
}

// ---- seeding + CleanMatches + LinearExtend + first SparseDP for one read.  Returns MP_OK with `ext` (extended clusters, global t) and
// `chains` (nch of them), or MP_UNALIGNED / an error.
__device__ __noinline__ int mp_stage1(const MpCtx &C, int r, Arena &ar, ClusterSet &ext, UChain *&chains, int &nch) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const unsigned long long roff = C.rd.read_off[r];
  const uint32_t L = C.rd.read_len[r];
  nch = 0;
  ClusterSet raw; int n_raw_a = 0;
  { const int rc = mp_seed_clean(C, r, ar, raw, n_raw_a); if (rc != MP_OK) return rc; }
  unsigned long long tk = mp_clock();
  int repetitive = 0;
  if (raw.ncl == 0) return MP_UNALIGNED;
  for (int c = 0; c < raw.ncl; c++) {
    const float f = raw.freq[c];
    if (f > 1.0f && f <= 2.0f && raw.off[c + 1] - raw.off[c] >= 500) repetitive = 1;
  }
  // ---- LinearExtend on the raw K-mers of every cluster (Map_lowacc.h:118-153): t chromosome-relative inside, global again afterwards
  if (!mp_alloc_clusterset(ext, ar, raw.ncl, n_raw_a, true)) return MP_ERR_ARENA;
  {
    if (lane == 0) ext.off[0] = 0;
    int o = 0;
    for (int d = 0; d < raw.ncl; d++) {
      const int chrom = raw.chrom[d];
      const uint32_t coff = (uint32_t)C.ix.hdr_pos[chrom];
      const int a0 = raw.off[d], np = raw.off[d + 1] - a0;
      for (int m = lane; m < np; m += kLanes) raw.t[a0 + m] -= coff;
      wsync();
      const int o1 = mp_linear_extend_warp(C, roff, L, chrom, raw.q + a0, raw.t + a0, np, raw.strand[d], O.globalK, ext.q, ext.t, ext.len, o);
      if (lane == 0) {
        ext.off[d + 1] = o1;
        ext.strand[d] = -1; ext.chrom[d] = 0; ext.freq[d] = 0.0f; ext.qS[d] = 0xffffffffu; ext.qE[d] = 0; ext.tS[d] = 0xffffffffu; ext.tE[d] = 0;
        mp_decide_coordinates(ext, d, raw.strand[d], chrom, raw.freq[d]);
      }
      wsync();
      for (int m = o + lane; m < o1; m += kLanes) ext.t[m] += coff;
      if (lane == 0) { ext.tS[d] += coff; ext.tE[d] += coff; }
      wsync();
      o = o1;
    }
  }
  wsync();
  ext.ncl = raw.ncl;
  const int NE = ext.off[ext.ncl];
  tk = mp_tick(C, PF_LEXT1, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  // ---- first SparseDP on all anchors (Map_lowacc.h:184-188) + RemoveSpuriousJump
  const float match_rate = repetitive ? 3.0f : O.initial_anchorbonus;
  const int NA = O.NumAln < 8 ? O.NumAln : 8;
  chains = ar.alloc<UChain>(NA);
  SdpChain *sc = ar.alloc<SdpChain>(NA);
  int *cl_of = ar.alloc<int>(NE + 1);
  uint8_t *clst = ar.alloc<uint8_t>(ext.ncl + 1);
  uint32_t *cbuf = ar.alloc<uint32_t>((unsigned long long)NA * (NE + 1));
  uint8_t *lbuf = ar.alloc<uint8_t>((unsigned long long)NA * (NE + 1));
  int *clbuf = ar.alloc<int>((unsigned long long)NA * (NE + 1));
  if (ar.overflow) return MP_ERR_ARENA;
  for (int c = lane; c < ext.ncl; c += kLanes) clst[c] = (uint8_t)(ext.strand[c] != 0);
  if (lane == 0) for (int c = 0; c < NA; c++) { sc[c].chain = cbuf + (unsigned long long)c * (NE + 1); sc[c].link = lbuf + (unsigned long long)c * (NE + 1); sc[c].n = 0; }
  wsync();
  SdpAnchors A; A.q = ext.q; A.t = ext.t; A.len = ext.len; A.nfrag = NE; A.cl_off = ext.off; A.cl_strand = clst; A.ncl = ext.ncl;
  const int nc = sdp_pure_matches(A, match_rate, O.alnthres, NA, (int)L, *C.pwl, ar, sc, cl_of);
  if (nc < 0) return MP_ERR_ARENA;
  wsync();
  for (int c = 0; c < nc; c++) {
    UChain u;
    u.idx = sc[c].chain; u.link = sc[c].link; u.cl = clbuf + (unsigned long long)c * (NE + 1);
    u.n = sc[c].n; u.nlink = u.n - 1; u.FirstSDPValue = sc[c].value; u.NumOfAnchors0 = u.n; u.NumOfAnchors1 = 0;
    u.QStart = sc[c].QStart; u.QEnd = sc[c].QEnd; u.TStart = sc[c].TStart; u.TEnd = sc[c].TEnd;
    for (int i = lane; i < u.n; i += kLanes) { const int f = (int)u.idx[i]; const int k = cl_of[f]; u.cl[i] = k; u.idx[i] = (uint32_t)(f - ext.off[k]); }
    wsync();
    mp_chain_filter(5, ext, u, ar, true);       // RemoveSpuriousJump
    if (lane == 0) chains[c] = u;
    wsync();
  }
  nch = nc;
  tk = mp_tick(C, PF_SDP1, tk);
  if (nc == 0) return MP_UNALIGNED;
  return MP_OK;
}

}  // namespace mp
}  // namespace lra
