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
  const int strand = b.p_strand[pb - 1];        // the reference's `st` after the loop over the merged clusters: the last part's strand
  int *idx = b.lidx + e0;
  int nl = 0;
  for (int base = 0; base < cnt; base += 32) {  // ordered compaction of the anchors >= 40
    const int i = base + lane;
    const bool is_long = i < cnt && L[i] >= 40;
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

}  // namespace lra
