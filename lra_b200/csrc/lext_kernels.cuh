// a15: the low-accuracy pipeline's linear extension, batched over reads --
//   LinearExtend (GenomePairs overload, reference LinearExtend.h:658-716) with Checkbp (:50-84), DecideCoordinates (:104-128),
//   TrimOverlappedAnchors (vector<Cluster>&, :573-647) with the LongAnchors order (:11-47).
// The reference walks the diagonal-sorted anchors of a cluster with two cursors (m = first anchor of the run being built, n = next anchor).
// Whether anchor n joins the run of anchor n-1 depends on that pair alone -- same diagonal and (overlapping K-mers, or the exact-base
// extension after n-1 reaches n) -- and what a run emits depends on its first anchor, its last anchor and the extension end computed for
// the last pair.  So the walk is restated as: one thread per ADJACENT PAIR decides join / break and records where the run of n-1 ends
// (lext_link_kernel), a scan of the break flags numbers the runs, one thread per anchor writes the head / tail fields of its run
// (lext_emit_kernel), and one warp per extended cluster finishes lengths, reduces the bounding box and trims (lext_group_kernel).
// TrimOverlappedAnchors visits the long anchors in std::sort order; two long anchors can tie under its comparator and which of them is
// `prev` decides which one is trimmed, so libstdc++'s introsort is replayed on the index list (introsort.cuh).  Its loop iteration ln reads
// only values that no EARLIER iteration wrote (iteration ln writes anchor idx[ln-1], which only iterations ln-1 and ln read), so a warp
// runs 32 iterations at a time: all reads, then all writes.
// Bases are compared as 2-bit codes + N flag (both sequences were normalised at upload); the reference compares the bytes, on strand 1
// without complementing -- restated as is.  GenomePos arithmetic is uint32 arithmetic.
#pragma once
#include "lra_common.cuh"
#include "introsort.cuh"

namespace lra {

__device__ __forceinline__ uint32_t lext_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t lext_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

struct LextBatch {
  int n_groups;
  long long n_parts;
  unsigned long long N;                         // anchors of the batch
  int K, trim;
  SeqView reads, genome;
  const unsigned long long *g_off;              // [n_groups + 1] parts of every group (extended cluster)
  const unsigned long long *p_off;              // [n_parts + 1]  anchors of every part (input cluster)
  const uint8_t *p_strand;                      // [n_parts]
  const unsigned long long *chrom_off;          // [n_parts] contig of the part: position in the packed genome, length
  const uint32_t *chrom_len;
  const unsigned long long *read_off;           // [n_parts] the read of the part: position in the read arena, length
  const uint32_t *read_len;
  const uint32_t *q, *t;                        // [N] diagonal-sorted anchors (t relative to the contig)
  uint32_t *part_of;                            // [N] scratch: part of every anchor
  unsigned long long *run;                      // [N + 1] scratch: break flags, then their exclusive scan (run[N] = number of extended anchors)
  uint32_t *end_q, *end_t;                      // [N] scratch, valid at the last anchor of a run: read end of the run, strand-1 genome start
  int *lidx;                                    // [N] scratch: long-anchor index lists of the groups (at the group's first output)
  unsigned long long *e_off;                    // [n_groups + 1] out
  uint32_t *eq, *et;                            // [N] out
  int32_t *elen;                                // [N] out
  uint32_t *box;                                // [n_groups * 4] out: qStart, qEnd, tStart, tEnd (before trimming, as DecideCoordinates runs first)
};

// Checkbp.  Returns through qe / te where the extension after anchor (cq, ct) stopped.
__device__ __forceinline__ void lext_checkbp(const LextBatch &b, uint32_t cq, uint32_t ct, uint32_t nq, uint32_t nt, unsigned long long coff, uint32_t L,
                                             unsigned long long roff, uint32_t rlen, int strand, uint32_t &qe, uint32_t &te) {
  const uint32_t K = (uint32_t)b.K;
  uint32_t curQ = cq + K, curT, nextT;
  if (strand == 0) {
    curT = ct + K < L ? ct + K : L;
    nextT = nt < L ? nt : L;
    while (curQ < rlen && curT < L && nq > curQ && nextT > curT && seq_code(b.genome, coff + curT) == seq_code(b.reads, roff + curQ)) { curQ++; curT++; }
  } else {
    curT = ct - 1u < L - 1u ? ct - 1u : L - 1u;
    nextT = nt + K - 1u < L - 1u ? nt + K - 1u : L - 1u;
    while (curQ < rlen && nq > curQ && nextT < curT && seq_code(b.genome, coff + curT) == seq_code(b.reads, roff + curQ)) { curQ++; curT--; }
  }
  qe = curQ; te = curT;
}

__global__ void __launch_bounds__(256) lext_link_kernel(LextBatch b) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.N) return;
  long long lo = 0, hi = b.n_parts;             // the part with p_off[p] <= i < p_off[p + 1]
  while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (b.p_off[mid] <= i) lo = mid; else hi = mid; }
  const long long p = lo;
  b.part_of[i] = (uint32_t)p;
  const int strand = b.p_strand[p];
  const uint32_t K = (uint32_t)b.K;
  const uint32_t qi = b.q[i], ti = b.t[i];
  bool head = true;
  if (i != b.p_off[p]) {
    const uint32_t qp = b.q[i - 1], tp = b.t[i - 1];
    long long curDiag, nextDiag;
    if (strand == 0) { curDiag = (long long)qp - (long long)tp; nextDiag = (long long)qi - (long long)ti; }
    else { curDiag = (long long)qp + (long long)tp; nextDiag = (long long)qi + (long long)ti; }
    if (curDiag == nextDiag) {
      if (qi < qp + K) head = false;
      else {
        uint32_t qe, te;
        lext_checkbp(b, qp, tp, qi, ti, b.chrom_off[p], b.chrom_len[p], b.read_off[p], b.read_len[p], strand, qe, te);
        if (strand == 0 ? (qe == qi && te == ti) : (qe == qi && te == ti + K - 1u)) head = false;
        else { b.end_q[i - 1] = qe; b.end_t[i - 1] = te + 1u; }
      }
    } else { b.end_q[i - 1] = qp + K; b.end_t[i - 1] = tp; }
  }
  if (i + 1 == b.p_off[p + 1]) { b.end_q[i] = qi + K; b.end_t[i] = ti; }
  b.run[i] = head ? 1ull : 0ull;
}

// after the exclusive scan of run[]: anchor i belongs to output run[i + 1] - 1; it is a head if run[i + 1] != run[i], a tail if it is the last
// anchor of its part or the next anchor is a head
__global__ void __launch_bounds__(256) lext_emit_kernel(LextBatch b) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.N) return;
  const unsigned long long r0 = b.run[i], r1 = b.run[i + 1];
  const unsigned long long o = r1 - 1;
  const uint32_t p = b.part_of[i];
  const int strand = b.p_strand[p];
  if (r1 != r0) { b.eq[o] = b.q[i]; if (strand == 0) b.et[o] = b.t[i]; }
  const bool tail = (i + 1 == b.p_off[p + 1]) || (b.run[i + 2] != r1);
  if (tail) { b.elen[o] = (int32_t)b.end_q[i]; if (strand != 0) b.et[o] = b.end_t[i]; }     // elen: the run's read end for now, the group kernel subtracts eq
}

struct LextLongLess {
  const uint32_t *q, *t; const int32_t *len; int strand;
  __device__ __forceinline__ bool operator()(int i, int j) const {
    if (strand == 0) { if (q[i] != q[j]) return q[i] < q[j]; return t[i] < t[j]; }
    const uint32_t ei = q[i] + (uint32_t)len[i], ej = q[j] + (uint32_t)len[j];
    if (ei != ej) return ei > ej;
    return t[i] < t[j];
  }
};

// TrimOverlappedAnchors for one extended cluster, by one warp (see the header of this file)
__device__ __forceinline__ void lext_trim_warp(uint32_t *Q, uint32_t *T, int32_t *L, int cnt, int *idx, int strand, int thr, int lane) {
  int nl = 0;
  for (int base = 0; base < cnt; base += 32) {  // ordered compaction of the long anchors
    const int i = base + lane;
    const bool is_long = i < cnt && L[i] >= thr;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, is_long);
    if (is_long) idx[nl + __popc(m & ((1u << lane) - 1u))] = i;
    nl += __popc(m);
  }
  __syncwarp();
  if (lane == 0) std_sort_replay(idx, nl, LextLongLess{Q, T, L, strand});
  __syncwarp();
  for (int base = 1; base < nl; base += 32) {
    const int ln = base + lane;
    int prev = 0, cut = 0;
    if (ln < nl) {
      prev = idx[ln - 1];
      const int cur = idx[ln];
      int overlap_r = 0, overlap_g = 0;
      const uint32_t pend = Q[prev] + (uint32_t)L[prev];
      if (strand == 0) { if (Q[cur] < pend && Q[cur] >= pend - 30u) overlap_r = (int)(pend - Q[cur]); }
      else { const uint32_t cend = Q[cur] + (uint32_t)L[cur]; if (cend > Q[prev] && cend <= Q[prev] + 30u) overlap_r = (int)(cend - Q[prev]); }
      const uint32_t ptend = T[prev] + (uint32_t)L[prev];
      if (T[cur] < ptend && T[cur] >= ptend - 30u) overlap_g = (int)(ptend - T[cur]);
      if (overlap_r > 0 || overlap_g > 0) cut = (overlap_r > overlap_g ? overlap_r : overlap_g) + 1;
    }
    __syncwarp();
    if (cut) { if (strand == 1) Q[prev] += (uint32_t)cut; L[prev] -= cut; }
    __syncwarp();
  }
}

// one warp per group
__global__ void __launch_bounds__(128) lext_group_kernel(LextBatch b) {
  const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= b.n_groups) return;
  const unsigned long long pa = b.g_off[g], pb = b.g_off[g + 1];
  const unsigned long long e0 = b.run[b.p_off[pa]], e1 = b.run[b.p_off[pb]];
  if (lane == 0) { b.e_off[g] = e0; if (g == b.n_groups - 1) b.e_off[g + 1] = e1; }
  const int cnt = (int)(e1 - e0);
  uint32_t *Q = b.eq + e0, *T = b.et + e0;
  int32_t *L = b.elen + e0;
  uint32_t qs = 0xFFFFFFFFu, qe = 0, ts = 0xFFFFFFFFu, te = 0;
  for (int i = lane; i < cnt; i += 32) {
    const int32_t len = (int32_t)((uint32_t)L[i] - Q[i]);
    L[i] = len;
    qs = lext_umin(qs, Q[i]); qe = lext_umax(qe, Q[i] + (uint32_t)len); ts = lext_umin(ts, T[i]); te = lext_umax(te, T[i] + (uint32_t)len);
  }
  for (int d = 16; d; d >>= 1) {
    qs = lext_umin(qs, __shfl_xor_sync(0xFFFFFFFFu, qs, d)); qe = lext_umax(qe, __shfl_xor_sync(0xFFFFFFFFu, qe, d));
    ts = lext_umin(ts, __shfl_xor_sync(0xFFFFFFFFu, ts, d)); te = lext_umax(te, __shfl_xor_sync(0xFFFFFFFFu, te, d));
  }
  if (lane == 0) {
    uint32_t *bx = b.box + 4 * (size_t)g;
    if (cnt > 0) { bx[0] = qs; bx[1] = qe; bx[2] = ts; bx[3] = te; } else { bx[0] = bx[1] = bx[2] = bx[3] = 0; }
  }
  if (!b.trim || cnt == 0) return;
  __syncwarp();
  // trim 1: the vector<Cluster> overload (anchors >= 40, strand = the reference's `st` after its loop over the merged clusters: the last part's);
  // trim 2: the GenomePairs overload (LinearExtend.h:724-777: anchors >= 50, forward only)
  lext_trim_warp(Q, T, L, cnt, b.lidx + e0, b.trim == 2 ? 0 : (int)b.p_strand[pb - 1], b.trim == 2 ? 50 : 40, lane);
}

// ---- the high-accuracy overload: LinearExtend(vector<Cluster*> clusters, vector<Cluster> &extCluster, vector<Tup> &chain, ...) (LinearExtend.h:134-350),
// TrimOverlappedAnchors from `start` (LinearExtend_chain :782-792) and MergeMatchesSameDiag (:794-823).
// A unit = one chain entry (chain c of a read, position e): it extends cluster chain[e] with the overlap points of clusters chain[e-1] and
// chain[e+1] in its Set.  The walk carries state from anchor to anchor (m, n, chm; an anchor that covers a Set point consumes its successor's
// turn), so it is replayed literally by one thread per unit into the unit's slot (at most one output per input anchor); the counts are scanned,
// and one warp per unit compacts the slot, reduces the box, trims, and flags the heads of the same-diagonal runs.
struct LextChainBatch {
  int n_units;
  int K, skiprepetitive, trim;
  long long merge_dist;
  SeqView reads, genome;
  const uint32_t *unit_cl;                      // [n_units] cluster of the unit
  const uint8_t *unit_edge;                     // [n_units] bit 0: first entry of its chain, bit 1: last
  const unsigned long long *slot_off;           // [n_units + 1] exclusive scan of the units' cluster sizes
  const unsigned long long *cl_off;             // [n_clusters + 1]
  const uint32_t *cq, *ct;                      // cluster anchors, already sorted (DiagonalSort / AntiDiagonalSort by strand)
  const uint32_t *cl_box;                       // [n_clusters * 4] qStart, qEnd, tStart, tEnd
  const uint8_t *cl_strand;
  const float *cl_freq;
  const unsigned long long *cl_chrom_off;       // contig of the cluster in the packed genome
  const uint32_t *cl_chrom_len;
  const unsigned long long *cl_read_off;        // the read of the cluster in the read arena
  const uint32_t *cl_read_len;
  uint32_t *sq, *st;                            // [slot_off[n_units]] slot scratch
  int32_t *sl;
  uint8_t *so;
  unsigned long long *cnt;                      // [n_units + 1] outputs per unit, then their exclusive scan
  int32_t *u_overlap;                           // [n_units] out: increments of the reference's `overlap` counter
  int *lidx;                                    // [slot_off[n_units]] scratch
  uint32_t *eq, *et;                            // out, compacted
  int32_t *elen;
  uint8_t *eovp;                                // out: Cluster::overlap
  uint8_t *md_head;                             // out: 1 = the anchor starts a new same-diagonal run (MergeMatchesSameDiag's start[] entries)
  uint32_t *box;                                // [n_units * 4]
};

__device__ __forceinline__ bool lext_check_overlap(uint32_t q, uint32_t t, uint32_t K, const uint32_t *sp, const uint32_t sf, int ns) {   // CheckOverlap
  for (int i = 0; i < ns; i++) {
    const bool on_t = (sf >> i) & 1u;
    const uint32_t lo = on_t ? t : q;
    if (sp[i] >= lo && sp[i] < lo + K) return true;
  }
  return false;
}

__device__ __noinline__ void lextc_walk_one(const LextChainBatch &b, const int u) {
  const uint32_t cm = b.unit_cl[u];
  const unsigned long long a = b.cl_off[cm];
  const long size = (long)(b.cl_off[cm + 1] - a);
  unsigned long long no = 0;
  int ovl = 0;
  if (size > 0) {
    const uint32_t K = (uint32_t)b.K;
    const int strand = b.cl_strand[cm];
    uint32_t sp[8]; uint32_t sf = 0; int ns = 0;
    const uint32_t *bx = b.cl_box + 4 * (size_t)cm;
    const uint32_t qsb = bx[0], qeb = bx[1], tsb = bx[2], teb = bx[3];
    if (b.skiprepetitive && b.cl_freq[cm] <= 1.1f) {
      const int edge = b.unit_edge[u];
#pragma unroll
      for (int side = 0; side < 2; side++) {
        if (side == 0 ? (edge & 1) : (edge & 2)) continue;
        const uint32_t *ob = b.cl_box + 4 * (size_t)b.unit_cl[side == 0 ? u - 1 : u + 1];
        if (ob[0] > qsb && ob[0] < qeb) { sp[ns++] = ob[0]; }
        if (ob[1] > qsb && ob[1] < qeb) { sp[ns++] = ob[1]; }
        if (ob[2] > tsb && ob[2] < teb) { sf |= 1u << ns; sp[ns++] = ob[2]; }
        if (ob[3] > tsb && ob[3] < teb) { sf |= 1u << ns; sp[ns++] = ob[3]; }
      }
    }
    const uint32_t *pq = b.cq + a, *pt = b.ct + a;
    const unsigned long long so = b.slot_off[u];
    uint32_t *oq = b.sq + so, *ot = b.st + so; int32_t *ol = b.sl + so; uint8_t *oo = b.so + so;
    const unsigned long long coff = b.cl_chrom_off[cm], roff = b.cl_read_off[cm];
    const uint32_t clen = b.cl_chrom_len[cm], rlen = b.cl_read_len[cm];
    // the walk only needs the fields LextBatch's Checkbp reads
    LextBatch lb; lb.K = b.K; lb.reads = b.reads; lb.genome = b.genome;
    auto push = [&](uint32_t q, uint32_t t, uint32_t len, uint8_t o) { oq[no] = q; ot[no] = t; ol[no] = (int32_t)len; oo[no] = o; no++; };
    long n = 1, m = 0;
    bool chm = true;
    while (n < size) {
      if (chm) {
        if (lext_check_overlap(pq[m], pt[m], K, sp, sf, ns)) { push(pq[m], pt[m], K, 1); ovl++; m = n; n++; chm = true; continue; }
        chm = false;
      }
      if (lext_check_overlap(pq[n], pt[n], K, sp, sf, ns)) {
        push(pq[m], strand == 0 ? pt[m] : pt[n - 1], pq[n - 1] + K - pq[m], 0);
        push(pq[n], pt[n], K, 1); ovl++;
        m = n + 1; n = m + 1; chm = true;
        continue;
      }
      long long curDiag, nextDiag;
      if (strand == 0) { curDiag = (long long)pq[n - 1] - (long long)pt[n - 1]; nextDiag = (long long)pq[n] - (long long)pt[n]; }
      else { curDiag = (long long)pq[n - 1] + (long long)pt[n - 1]; nextDiag = (long long)pq[n] + (long long)pt[n]; }
      if (curDiag == nextDiag) {
        if (pq[n] < pq[n - 1] + K) n++;
        else {
          uint32_t qe, te;
          lext_checkbp(lb, pq[n - 1], pt[n - 1], pq[n], pt[n], coff, clen, roff, rlen, strand, qe, te);
          if (strand == 0 ? (qe == pq[n] && te == pt[n]) : (qe == pq[n] && te == pt[n] + K - 1u)) n++;
          else { push(pq[m], strand == 0 ? pt[m] : te + 1u, qe - pq[m], 0); m = n; n++; }
        }
      } else { push(pq[m], strand == 0 ? pt[m] : pt[n - 1], pq[n - 1] + K - pq[m], 0); m = n; n++; }
      chm = false;
    }
    if (n == size) push(pq[m], strand == 0 ? pt[m] : pt[n - 1], pq[n - 1] + K - pq[m], 0);
  }
  b.cnt[u] = no;
  b.u_overlap[u] = ovl;
}

__global__ void __launch_bounds__(128) lextc_walk_kernel(LextChainBatch b) {
  const int u = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (u >= b.n_units) return;
  lextc_walk_one(b, u);
}

// one warp per unit, after the scan of cnt[]
__global__ void __launch_bounds__(128) lextc_group_kernel(LextChainBatch b) {
  const int u = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (u >= b.n_units) return;
  const unsigned long long e0 = b.cnt[u], so = b.slot_off[u];
  const int cnt = (int)(b.cnt[u + 1] - e0);
  uint32_t *Q = b.eq + e0, *T = b.et + e0;
  int32_t *L = b.elen + e0;
  uint8_t *O = b.eovp + e0;
  uint32_t qs = 0xFFFFFFFFu, qe = 0, ts = 0xFFFFFFFFu, te = 0;
  for (int i = lane; i < cnt; i += 32) {
    const uint32_t q = b.sq[so + i], t = b.st[so + i];
    const int32_t len = b.sl[so + i];
    Q[i] = q; T[i] = t; L[i] = len; O[i] = b.so[so + i];
    qs = lext_umin(qs, q); qe = lext_umax(qe, q + (uint32_t)len); ts = lext_umin(ts, t); te = lext_umax(te, t + (uint32_t)len);
  }
  for (int d = 16; d; d >>= 1) {
    qs = lext_umin(qs, __shfl_xor_sync(0xFFFFFFFFu, qs, d)); qe = lext_umax(qe, __shfl_xor_sync(0xFFFFFFFFu, qe, d));
    ts = lext_umin(ts, __shfl_xor_sync(0xFFFFFFFFu, ts, d)); te = lext_umax(te, __shfl_xor_sync(0xFFFFFFFFu, te, d));
  }
  if (lane == 0) {
    uint32_t *bx = b.box + 4 * (size_t)u;
    if (cnt > 0) { bx[0] = qs; bx[1] = qe; bx[2] = ts; bx[3] = te; } else { bx[0] = bx[1] = bx[2] = bx[3] = 0; }
  }
  if (cnt == 0) return;
  __syncwarp();
  const int strand = b.cl_strand[b.unit_cl[u]];
  if (b.trim) lext_trim_warp(Q, T, L, cnt, b.lidx + so, strand, 40, lane);
  __syncwarp();
  // MergeMatchesSameDiag: anchor i continues the run of i-1 iff neither is an overlap anchor, they share GetDiag, i starts after i-1 ends on the
  // read and the read gap is <= merge_dist
  for (int i = lane; i < cnt; i += 32) {
    bool head = true;
    if (i > 0) {
      const long long dp = strand == 0 ? (long long)T[i - 1] - (long long)Q[i - 1] : (long long)Q[i - 1] + (long long)T[i - 1] + (long long)L[i - 1];
      const long long dc = strand == 0 ? (long long)T[i] - (long long)Q[i] : (long long)Q[i] + (long long)T[i] + (long long)L[i];
      const uint32_t prev_qEnd = Q[i - 1] + (uint32_t)L[i - 1];
      long long gd = (long long)Q[i] - ((long long)Q[i - 1] + (long long)L[i - 1]);
      if (gd < 0) gd = -gd;
      if (O[i - 1] == 0 && O[i] == 0 && dp == dc && prev_qEnd < Q[i] && gd <= b.merge_dist) head = false;
    }
    b.md_head[e0 + i] = head ? 1 : 0;
  }
}

}  // namespace lra
