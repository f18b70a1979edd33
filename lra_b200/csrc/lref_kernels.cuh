// a13: REFINEclusters (reference ClusterRefine.h:50-240), batched over clusters.
//   Cluster::CHROMIndex / Header::Find / GetNextOffset   Clustering.h:326-336, Genome.h:19-47
//   SwapStrand                                           ClusterRefine.h:24-31
//   CartesianTargetSort / LowerBound / UpperBound        Sorting.h:182-224      (a total order on (t, q): any correct sort)
//   LocalIndex::LookupIndex                              MMIndex.h:175-190
//   CompareLists<LocalTuple,SmallTuple>(Global = false)  CompareLists.h:8-146   (literal: front/back galloping, emission order kept)
//   AppendValues (diagonal band + cluster box)           TupleOps.h:159-195
//   Cluster::SetClusterBoundariesFromMatches             Clustering.h:308-322
//
// The reference's triple loop  cluster -> genome window lsi -> read window qi  is flattened into three launches:
//   lref_prep_kernel     one warp per cluster: contig, strand swap, diagonal band, sort by (t, q), window range [ls, le]
//   lref_unit_kernel     one thread per (cluster, lsi): anchors inside the window -> read interval -> read windows [qis, qie]
//   lref_task_literal_kernel<E>  one thread per (cluster, lsi, qi): CompareLists of two ~700-tuple lists + the band/box filter;
//                        count pass, scan, emit pass: refined anchors land in the reference's order
//                        (lref_task_kernel<E>: the same per warp with shared searches -- slower, kept as a cross-check)
//   lref_finish_kernel   one warp per cluster: strand swap back, boundaries, refineEffiency (binary32 division)
#pragma once
#include "lra_common.cuh"
#include "lidx_kernels.cuh"

namespace lra {

struct LidxView {                      // a LocalIndex image (lidx_kernels.cuh)
  const unsigned long long *win_off;   // [n_win + 1] arena position of each window; last entry = end of the last sequence
  const uint32_t *win_len;             // [n_win]
  const unsigned long long *bnd;       // [n_win + 1] tuple boundaries
  const uint32_t *mins;
  const uint32_t *win_first;           // [n_seq + 1] first window of each sequence
  const unsigned long long *seq_start; // [n_seq] arena position of each sequence
  const uint32_t *seq_len;             // [n_seq]
  int n_win, n_seq;
};

struct LrefBatch {
  int n_clusters;
  // clusters (input)
  const uint32_t *in_q, *in_t;         // anchors, concatenated
  const unsigned long long *m_off;     // [n_clusters + 1]
  const uint32_t *in_box;              // [n_clusters][4] qStart, qEnd, tStart, tEnd (tStart / tEnd global)
  const uint8_t *strand;               // [n_clusters]
  const uint32_t *read_id;             // [n_clusters] sequence index in the read LocalIndex images
  const unsigned long long *hdr_pos;   // genome.header.pos
  int n_hdr;
  LidxView gl, rd[2];                  // genome; reads forward / reverse complement
  int global_k, small_k, window;
  long long local_max_freq;
  // mode 1 = Refine_splitchain (ChainRefine.h:383-576): the units are split chains; anchors in chain order with their lengths and the
  // strand of the cluster each comes from, chromIndex given, per-window diagonal bands (opts.limitrefine)
  int mode, limitrefine;
  const uint32_t *in_len;
  const uint8_t *in_mstrand;
  const int32_t *in_chrom;
  // working copy of the clusters = what the reference leaves in clusters[ph] (output)
  uint32_t *m_q, *m_t;
  uint32_t *box;                       // [n_clusters][4]
  uint32_t *fbox;                      // [n_clusters][4] the box AppendValues filters with: read start/end on the unit's strand, target chromosome-relative
  unsigned long long *keys;            // sort scratch, key_off[c] .. (power-of-two sized slots)
  const unsigned long long *key_off;   // [n_clusters]
  // per cluster
  int32_t *status, *chrom;             // status: 0 refined, 1 no anchors, 2 spans two contigs
  long long *diag;                     // [n_clusters][2] minDiagNum, maxDiagNum
  uint32_t *chrom_off;                 // [n_clusters]
  int32_t *ls;                         // [n_clusters]
  unsigned long long *unit_off;        // [n_clusters + 1] number of windows le - ls + 1, then offsets
  // per unit (cluster, lsi)
  uint32_t *u_cluster, *u_qis, *u_gstart;
  long long *u_band;                   // [n_units][2] diagonal band of the unit's anchors
  unsigned long long *task_off;        // [n_units + 1]
  // per task
  unsigned long long *out_off;         // [n_tasks + 1]
  // refined anchors
  uint32_t *r_q, *r_t, *r_tup;
  unsigned long long out_cap;
  unsigned long long *r_off;           // [n_clusters + 1] first refined anchor of each cluster
  uint32_t *rbox;                      // [n_clusters][4]
  float *eff;                          // [n_clusters]
  // the count pass parks the first slot_cap anchors of every window pair here ((q, t, tuple) triples), so that the emit pass only re-runs the pairs
  // that produced more (nullptr: every non-empty pair is re-run)
  uint32_t *slot;
  int slot_cap;
};

__device__ __forceinline__ int lref_hdr_find(const unsigned long long *pos, int n, unsigned long long query) {   // Header::Find
  if (n > 0 && query == pos[0]) return 0;
  int lo = 0, len = n;
  while (len > 0) { const int half = len >> 1; if (pos[lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
  if (lo < n && query == pos[lo]) return lo;
  return lo - 1;
}

// LocalIndex::LookupIndex over the offsets {off[0], .., off[n_win - 1], end} minus `base`
__device__ __forceinline__ int lref_lookup(const unsigned long long *off, int n_win, unsigned long long base, unsigned long long end, unsigned long long query) {
  int lo = 0, len = n_win + 1;
  while (len > 0) {
    const int half = len >> 1, mid = lo + half;
    const unsigned long long o = (mid < n_win ? off[mid] : end) - base;
    if (o < query) { lo = mid + 1; len -= half + 1; } else len = half;
  }
  if (lo <= n_win && ((lo < n_win ? off[lo] : end) - base) == query) return lo;
  return lo - 1;
}

__global__ void __launch_bounds__(128) lref_prep_kernel(LrefBatch b) {
  const int lane = threadIdx.x & 31;
  const int c = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (c >= b.n_clusters) return;
  const unsigned long long m0 = b.m_off[c];
  const int nm = (int)(b.m_off[c + 1] - m0);
  uint32_t *mq = b.m_q + m0, *mt = b.m_t + m0;
  uint32_t *box = b.box + 4 * c;
  if (lane < 4) box[lane] = b.in_box[4 * c + lane];
  if (lane == 0) { b.unit_off[c] = 0; b.chrom[c] = 0; b.diag[2 * c] = 0; b.diag[2 * c + 1] = 0; b.chrom_off[c] = 0; b.ls[c] = 0; }
  __syncwarp();
  if (nm == 0) { if (lane == 0) b.status[c] = 1; return; }
  if (b.mode == 1) {
    // Refine_splitchain: flip the anchors for the duration of the call (ChainRefine.h:399-411), band +-50 over the whole chain (:417-428)
    const int chrom = b.in_chrom[c];
    const uint32_t chromOffset = (uint32_t)b.hdr_pos[chrom];
    const uint32_t QStart = box[0], QEnd = box[1], TStart = box[2], TEnd = box[3];
    const uint32_t chromEndOffset = (uint32_t)b.hdr_pos[lref_hdr_find(b.hdr_pos, b.n_hdr, (unsigned long long)TEnd) + 1];
    const uint32_t readLen = b.rd[0].seq_len[b.read_id[c]];
    long long maxDN = -(1ll << 62), minDN = (1ll << 62);
    for (int i = lane; i < nm; i += 32) {
      uint32_t q = b.in_q[m0 + i];
      const uint32_t t = b.in_t[m0 + i] - chromOffset;
      if (b.in_mstrand[m0 + i]) q = readLen - (q + (uint32_t)b.global_k);
      mq[i] = q; mt[i] = t;
      const long long d = (long long)t - (long long)q;
      maxDN = d > maxDN ? d : maxDN; minDN = d < minDN ? d : minDN;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const long long a = __shfl_xor_sync(0xffffffffu, maxDN, o), bb = __shfl_xor_sync(0xffffffffu, minDN, o);
      maxDN = a > maxDN ? a : maxDN; minDN = bb < minDN ? bb : minDN;
    }
    maxDN += 50; minDN -= 50;
    if (lane == 0) {
      const int strand = b.strand[c];
      uint32_t *fb = b.fbox + 4 * c;
      if (strand == 0) { fb[0] = QStart; fb[1] = QEnd; } else { fb[0] = readLen - QEnd; fb[1] = readLen - QStart; }
      fb[2] = TStart - chromOffset; fb[3] = TEnd - chromOffset;
      const uint32_t wts = (TStart >= chromOffset + (uint32_t)b.window) ? TStart - (uint32_t)b.window : chromOffset;
      const uint32_t wte = (TEnd + (uint32_t)b.window < chromEndOffset) ? TEnd + (uint32_t)b.window : chromEndOffset;
      const unsigned long long gend = b.gl.win_off[b.gl.n_win];
      const int ls = lref_lookup(b.gl.win_off, b.gl.n_win, 0ull, gend, wts), le = lref_lookup(b.gl.win_off, b.gl.n_win, 0ull, gend, wte);
      b.status[c] = 0; b.chrom[c] = chrom; b.diag[2 * c] = minDN; b.diag[2 * c + 1] = maxDN; b.chrom_off[c] = chromOffset; b.ls[c] = ls;
      b.unit_off[c] = le >= ls ? (unsigned long long)(le - ls + 1) : 0ull;
    }
    return;
  }
  const uint32_t tStart = box[2], tEnd = box[3];
  const int first = lref_hdr_find(b.hdr_pos, b.n_hdr, (unsigned long long)tStart + 1), last = lref_hdr_find(b.hdr_pos, b.n_hdr, (unsigned long long)tEnd);
  if (first != last) {
    if (lane == 0) b.status[c] = 2;
    for (int i = lane; i < nm; i += 32) { mq[i] = b.in_q[m0 + i]; mt[i] = b.in_t[m0 + i]; }
    return;
  }
  const uint32_t chromOffset = (uint32_t)b.hdr_pos[first];
  const uint32_t chromEndOffset = (uint32_t)b.hdr_pos[last + 1];
  const int strand = b.strand[c];
  const uint32_t readLen = b.rd[0].seq_len[b.read_id[c]];
  long long maxDN = -(1ll << 62), minDN = (1ll << 62);
  unsigned long long *keys = b.keys + b.key_off[c];
  int P = 1;
  while (P < nm) P <<= 1;
  for (int i = lane; i < P; i += 32) {
    if (i < nm) {
      uint32_t q = b.in_q[m0 + i];
      const uint32_t t = b.in_t[m0 + i] - chromOffset;
      if (strand == 1) q = readLen - (q + (uint32_t)b.global_k);
      const long long d = (long long)t - (long long)q;
      maxDN = d > maxDN ? d : maxDN; minDN = d < minDN ? d : minDN;
      keys[i] = ((unsigned long long)t << 32) | q;
    } else keys[i] = ~0ull;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, maxDN, o), bb = __shfl_xor_sync(0xffffffffu, minDN, o);
    maxDN = a > maxDN ? a : maxDN; minDN = bb < minDN ? bb : minDN;
  }
  maxDN += 100; minDN -= 100;
  __syncwarp();
  for (int kk = 2; kk <= P; kk <<= 1) {          // CartesianTargetSort
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], cc = keys[ixj];
          const bool up = (i & kk) == 0;
          if ((a > cc) == up) { keys[i] = cc; keys[ixj] = a; }
        }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < nm; i += 32) { const unsigned long long kx = keys[i]; mt[i] = (uint32_t)(kx >> 32); mq[i] = (uint32_t)kx; }
  if (lane == 0) {
    if (strand == 1) { const uint32_t r = box[0]; box[0] = readLen - box[1]; box[1] = readLen - r; }
    uint32_t *fb = b.fbox + 4 * c;
    fb[0] = box[0]; fb[1] = box[1]; fb[2] = box[2] - chromOffset; fb[3] = box[3] - chromOffset;
    uint32_t wts, wte;
    if (chromOffset + (uint32_t)b.window > tStart) wts = chromOffset; else wts = tStart - (uint32_t)b.window;
    if (tEnd + (uint32_t)b.window > chromEndOffset) wte = chromEndOffset - 1; else wte = tEnd + (uint32_t)b.window;
    const unsigned long long gend = b.gl.win_off[b.gl.n_win];
    const int ls = lref_lookup(b.gl.win_off, b.gl.n_win, 0ull, gend, wts), le = lref_lookup(b.gl.win_off, b.gl.n_win, 0ull, gend, wte);
    b.status[c] = 0; b.chrom[c] = first; b.diag[2 * c] = minDN; b.diag[2 * c + 1] = maxDN; b.chrom_off[c] = chromOffset; b.ls[c] = ls;
    b.unit_off[c] = le >= ls ? (unsigned long long)(le - ls + 1) : 0ull;
  }
}

// one thread per (cluster, lsi)
__global__ void __launch_bounds__(128) lref_unit_kernel(LrefBatch b, unsigned long long n_units) {
  const unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_units) return;
  int c;
  { int lo = 0, len = b.n_clusters;      // the cluster whose unit range contains u: last c with unit_off[c] <= u
    while (len > 0) { const int half = len >> 1; if (b.unit_off[lo + half] <= u) { lo += half + 1; len -= half + 1; } else len = half; }
    c = lo - 1; }
  const int ls = b.ls[c];
  const int lsi = ls + (int)(u - b.unit_off[c]);
  const int le = ls + (int)(b.unit_off[c + 1] - b.unit_off[c]) - 1;
  b.u_cluster[u] = (uint32_t)c; b.u_qis[u] = 0; b.u_gstart[u] = 0; b.task_off[u] = 0;
  const uint32_t chromOffset = b.chrom_off[c];
  const unsigned long long o0 = b.gl.win_off[lsi], o1 = b.gl.win_off[lsi + 1];
  if (o0 < chromOffset || o1 < chromOffset) return;
  const uint32_t gStart = (uint32_t)(o0 - chromOffset), gEnd = (uint32_t)(o1 - 1 - chromOffset);
  if (gStart >= gEnd) return;
  const unsigned long long m0 = b.m_off[c];
  const int nm = (int)(b.m_off[c + 1] - m0);
  const uint32_t *mq = b.m_q + m0, *mt = b.m_t + m0;
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
  if (matchStart >= nm) return;
  const int s = b.strand[c];
  const uint32_t rid = b.read_id[c];
  const uint32_t readLen = b.rd[s].seq_len[rid];
  uint32_t readStart = mq[matchStart], readEnd = mq[matchEnd];
  if (readStart == readEnd) { if (lsi > ls && readStart > 0u) readStart = 0u; }   // prev_readEnd is 0 at its only read (ClusterRefine.h:152,162)
  if (lsi == ls) { if (readStart < (uint32_t)b.window) readStart = 0; else readStart -= (uint32_t)b.window; }
  if (lsi == le) { if (readEnd + (uint32_t)b.window > readLen) readEnd = readLen; else readEnd += (uint32_t)b.window; }
  if (readStart > readEnd) return;
  const LidxView &rd = b.rd[s];
  const int wf = (int)rd.win_first[rid], nw = (int)rd.win_first[rid + 1] - wf;
  const unsigned long long base = rd.seq_start[rid], end = base + readLen;
  const int qis = lref_lookup(rd.win_off + wf, nw, base, end, readStart);
  const int qie = lref_lookup(rd.win_off + wf, nw, base, end, readEnd < readLen - 1 ? readEnd : readLen - 1);
  b.u_qis[u] = (uint32_t)qis; b.u_gstart[u] = gStart;
  b.u_band[2 * u] = b.diag[2 * c]; b.u_band[2 * u + 1] = b.diag[2 * c + 1];
  b.task_off[u] = qie >= qis ? (unsigned long long)(qie - qis + 1) : 0ull;
}

// Refine_splitchain: one thread per split chain walks its genome windows in order -- the anchor cursor (matchStart) is carried from
// window to window and the anchors are scanned sequentially, in chain order (ChainRefine.h:451-521)
__global__ void __launch_bounds__(128) lref_chain_unit_kernel(LrefBatch b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= b.n_clusters) return;
  const unsigned long long u0 = b.unit_off[c];
  const int nu = (int)(b.unit_off[c + 1] - u0);
  if (nu == 0) return;
  const int ls = b.ls[c], le = ls + nu - 1;
  const uint32_t chromOffset = b.chrom_off[c];
  const unsigned long long m0 = b.m_off[c];
  const int nm = (int)(b.m_off[c + 1] - m0);
  const uint32_t *mq = b.m_q + m0, *mt = b.m_t + m0, *ml = b.in_len + m0;
  const int s = b.strand[c];
  const uint32_t rid = b.read_id[c];
  const uint32_t readLen = b.rd[s].seq_len[rid];
  const LidxView &rd = b.rd[s];
  const int wf = (int)rd.win_first[rid], nw = (int)rd.win_first[rid + 1] - wf;
  const unsigned long long base = rd.seq_start[rid], end = base + readLen;
  int matchStart = 0, matchEnd = 0;
  for (int lsi = ls; lsi <= le; lsi++) {
    const unsigned long long u = u0 + (unsigned long long)(lsi - ls);
    b.u_cluster[u] = (uint32_t)c; b.u_qis[u] = 0; b.u_gstart[u] = 0; b.task_off[u] = 0; b.u_band[2 * u] = 0; b.u_band[2 * u + 1] = 0;
    if (lsi >= b.gl.n_win) continue;          // (one past the last window: the reference reads past seqOffsets there)
    const unsigned long long o0 = b.gl.win_off[lsi], o1 = b.gl.win_off[lsi + 1];
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
    long long bandMin = b.diag[2 * c], bandMax = b.diag[2 * c + 1];
    // opts.limitrefine: [min diagonal of the window's anchors - 100, +inf) -- the reference's upper bound is an uninitialised variable
    // that holds a stack address in the stock build (ChainRefine.h:491-501): it never filters
    if (b.limitrefine) { bandMin = mn - 100; bandMax = 0x7FFFFFFFFFFFFFFFll; }
    const uint32_t sow = 500;
    if (lsi == ls) readStart = (readStart < sow) ? 0 : readStart - sow;
    if (lsi == le) readEnd = (readEnd + sow > readLen) ? readLen : readEnd + sow;
    if (readStart > readEnd) continue;
    const int qis = lref_lookup(rd.win_off + wf, nw, base, end, readStart);
    const int qie = lref_lookup(rd.win_off + wf, nw, base, end, readEnd < readLen - 1 ? readEnd : readLen - 1);
    b.u_qis[u] = (uint32_t)qis; b.u_gstart[u] = gStart;
    b.u_band[2 * u] = bandMin; b.u_band[2 * u + 1] = bandMax;
    b.task_off[u] = qie >= qis ? (unsigned long long)(qie - qis + 1) : 0ull;
  }
}

// one THREAD per (cluster, lsi, qi): CompareLists<LocalTuple,SmallTuple>(Global = false) + AppendValues, statement by statement
// (the default: measured 4x faster than the warp form below, which LRA_B200_LREF_WARP=1 selects)
template <bool EMIT>
__global__ void __launch_bounds__(128) lref_task_literal_kernel(LrefBatch b, unsigned long long n_units, unsigned long long n_tasks) {
  const unsigned long long task = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= n_tasks) return;
  if (EMIT) {                                                      // the count pass found nothing for this window pair, or parked all of it
    const unsigned long long cnt = b.out_off[task + 1] - b.out_off[task];
    if (cnt == 0 || (b.slot != nullptr && cnt <= (unsigned long long)b.slot_cap)) return;
  }
  unsigned long long u;
  { unsigned long long lo = 0, len = n_units;
    while (len > 0) { const unsigned long long half = len >> 1; if (b.task_off[lo + half] <= task) { lo += half + 1; len -= half + 1; } else len = half; }
    u = lo - 1; }
  const int c = (int)b.u_cluster[u];
  const int lsi = b.ls[c] + (int)(u - b.unit_off[c]);
  const int s = b.strand[c];
  const uint32_t rid = b.read_id[c];
  const LidxView &rd = b.rd[s];
  const int qw = (int)rd.win_first[rid] + (int)b.u_qis[u] + (int)(task - b.task_off[u]);
  const uint32_t *q = rd.mins + rd.bnd[qw];
  const long nq = (long)(rd.bnd[qw + 1] - rd.bnd[qw]);
  const uint32_t *t = b.gl.mins + b.gl.bnd[lsi];
  const long nt = (long)(b.gl.bnd[lsi + 1] - b.gl.bnd[lsi]);
  const uint32_t readSegmentStart = (uint32_t)(rd.win_off[qw] - rd.seq_start[rid]);
  const uint32_t gStart = b.u_gstart[u];
  const long long minDN = b.u_band[2 * u], maxDN = b.u_band[2 * u + 1];
  const uint32_t *box = b.fbox + 4 * c;
  const uint32_t bqs = box[0], bqe = box[1], bts = box[2], bte = box[3];
  const long maxFreq = (long)b.local_max_freq;
  unsigned long long n_out = 0;
  const unsigned long long obase = EMIT ? b.out_off[task] : 0ull;
  auto push = [&](long qi, long ti) {
    const uint32_t qv = q[qi], tv = t[ti];
    const uint32_t qp = (qv >> 20) + readSegmentStart, tp = (tv >> 20) + gStart;
    const long long d = (long long)tp - (long long)qp;
    if (d >= minDN && d <= maxDN && qp >= bqs && qp < bqe && tp >= bts && tp < bte) {
      if (EMIT) {
        const unsigned long long o = obase + n_out;
        if (o < b.out_cap) { b.r_q[o] = qp; b.r_t[o] = tp; b.r_tup[o] = lt_t(qv); }
      } else if (b.slot != nullptr && n_out < (unsigned long long)b.slot_cap) {
        uint32_t *sl = b.slot + (task * (unsigned long long)b.slot_cap + n_out) * 3ull;
        sl[0] = qp; sl[1] = tp; sl[2] = lt_t(qv);
      }
      n_out++;
    }
  };
#define QK(i) lt_t(q[i])
#define TK(i) lt_t(t[i])
  if (nq > 0 && nt > 0) {
    long qs = 0, qe = nq - 1, ts = 0, te = nt;
    do {
      while (qs <= qe && QK(qs) < TK(ts)) qs++;
      if (qs >= qe) break;
      const uint32_t startGap = (QK(qs) - TK(ts)) & 0xFFFFFu;       // LocalTuple bit-field arithmetic: 20 bits
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
  if (!EMIT) b.out_off[task] = n_out;
}

// the parked anchors of every window pair that fitted its slot go to their place in the reference's order (8 lanes per pair)
__global__ void __launch_bounds__(256) lref_task_copy_kernel(LrefBatch b, unsigned long long n_tasks) {
  const unsigned long long task = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  if (task >= n_tasks) return;
  const unsigned long long o0 = b.out_off[task], cnt = b.out_off[task + 1] - o0;
  if (cnt == 0 || cnt > (unsigned long long)b.slot_cap) return;
  const uint32_t *sl = b.slot + task * (unsigned long long)b.slot_cap * 3ull;
  for (unsigned long long i = sub; i < cnt; i += 8) {
    const unsigned long long o = o0 + i;
    if (o < b.out_cap) { b.r_q[o] = sl[3 * i]; b.r_t[o] = sl[3 * i + 1]; b.r_tup[o] = sl[3 * i + 2]; }
  }
}

// ---- one WARP per (cluster, lsi, qi).  The control flow of CompareLists is replayed by all lanes in lock step on uniform state
// (no divergence, every load a broadcast out of L1); the lanes share the two searches that dominate it: the binary searches
// become 32-ary searches (one ballot per level), the "skip what cannot match" loops advance 32 tuples per ballot.
__device__ __forceinline__ long lref_warp_bound(const uint32_t *t, long lo, long hi, uint32_t key, bool upper, int lane) {
  // first index in [lo, hi) whose tuple is >= key (upper: > key); the list is sorted, so the result is that of std::lower/upper_bound
  while (hi - lo > 32) {
    const long stride = (hi - lo + 31) >> 5;
    const long idx = lo + (long)lane * stride;
    bool before = false;
    if (idx < hi) { const uint32_t x = lt_t(t[idx]); before = upper ? (x <= key) : (x < key); }
    const int c = __popc(__ballot_sync(0xffffffffu, before));
    if (c == 0) return lo;
    const long nlo = lo + (long)(c - 1) * stride + 1;
    const long nhi = lo + (long)c * stride;
    lo = nlo; hi = nhi < hi ? nhi : hi;
  }
  bool before = false;
  if (lo + lane < hi) { const uint32_t x = lt_t(t[lo + lane]); before = upper ? (x <= key) : (x < key); }
  return lo + __popc(__ballot_sync(0xffffffffu, before));
}

template <bool EMIT>
__global__ void __launch_bounds__(128) lref_task_kernel(LrefBatch b, unsigned long long n_units, unsigned long long n_tasks) {
  const int lane = threadIdx.x & 31;
  const unsigned long long task = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (task >= n_tasks) return;
  if (EMIT && b.out_off[task + 1] == b.out_off[task]) return;
  unsigned long long u;
  { unsigned long long lo = 0, len = n_units;
    while (len > 0) { const unsigned long long half = len >> 1; if (b.task_off[lo + half] <= task) { lo += half + 1; len -= half + 1; } else len = half; }
    u = lo - 1; }
  const int c = (int)b.u_cluster[u];
  const int lsi = b.ls[c] + (int)(u - b.unit_off[c]);
  const int s = b.strand[c];
  const uint32_t rid = b.read_id[c];
  const LidxView &rd = b.rd[s];
  const int qw = (int)rd.win_first[rid] + (int)b.u_qis[u] + (int)(task - b.task_off[u]);
  const uint32_t *q = rd.mins + rd.bnd[qw];
  const long nq = (long)(rd.bnd[qw + 1] - rd.bnd[qw]);
  const uint32_t *t = b.gl.mins + b.gl.bnd[lsi];
  const long nt = (long)(b.gl.bnd[lsi + 1] - b.gl.bnd[lsi]);
  const uint32_t readSegmentStart = (uint32_t)(rd.win_off[qw] - rd.seq_start[rid]);
  const uint32_t gStart = b.u_gstart[u];
  const long long minDN = b.u_band[2 * u], maxDN = b.u_band[2 * u + 1];
  const uint32_t *box = b.fbox + 4 * c;
  const uint32_t bqs = box[0], bqe = box[1], bts = box[2], bte = box[3];
  const long maxFreq = (long)b.local_max_freq;
  unsigned long long n_out = 0;
  const unsigned long long obase = EMIT ? b.out_off[task] : 0ull;
  auto push = [&](long qi, long ti) {       // uniform: every lane sees the same pair, lane 0 stores it
    const uint32_t qv = q[qi], tv = t[ti];
    const uint32_t qp = (qv >> 20) + readSegmentStart, tp = (tv >> 20) + gStart;
    const long long d = (long long)tp - (long long)qp;
    if (d >= minDN && d <= maxDN && qp >= bqs && qp < bqe && tp >= bts && tp < bte) {
      if (EMIT && lane == 0) {
        const unsigned long long o = obase + n_out;
        if (o < b.out_cap) { b.r_q[o] = qp; b.r_t[o] = tp; b.r_tup[o] = lt_t(qv); }
      }
      n_out++;
    }
  };
#define QK(i) lt_t(q[i])
#define TK(i) lt_t(t[i])
  if (nq > 0 && nt > 0) {
    long qs = 0, qe = nq - 1, ts = 0, te = nt;
    do {
      { // while (qs <= qe && QK(qs) < TK(ts)) qs++;
        const uint32_t k0 = TK(ts);
        for (;;) {
          const long i = qs + lane;
          const bool adv = i <= qe && QK(i) < k0;
          const unsigned m = __ballot_sync(0xffffffffu, adv);
          const int run = (m == 0xffffffffu) ? 32 : __ffs(~m) - 1;     // tuples are sorted: the predicate holds for a prefix
          qs += run;
          if (run < 32) break;
        }
      }
      if (qs >= qe) break;
      const uint32_t startGap = (QK(qs) - TK(ts)) & 0xFFFFFu;       // LocalTuple bit-field arithmetic: 20 bits
      { // while (qe > qs && te > ts && QK(qe) > TK(te - 1)) qe--;
        if (te > ts) {
          const uint32_t k1 = TK(te - 1);
          for (;;) {
            const long i = qe - lane;
            const bool adv = i > qs && QK(i) > k1;
            const unsigned m = __ballot_sync(0xffffffffu, adv);
            const int run = (m == 0xffffffffu) ? 32 : __ffs(~m) - 1;
            qe -= run;
            if (run < 32) break;
          }
        }
      }
      const uint32_t endGap = (TK(te - 1) - QK(qe)) & 0xFFFFFu;
      if (startGap == 0 || startGap > endGap) {
        const long tsOrig = ts, qsOrig = qs;
        ts = lref_warp_bound(t, ts, te, QK(qs), false, lane);
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
        else te = lref_warp_bound(t, ts, te, QK(qe), true, lane);
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
  if (!EMIT && lane == 0) b.out_off[task] = n_out;
}

// one warp per cluster: range of its refined anchors, strand swap back, boundaries, efficiency
__global__ void __launch_bounds__(128) lref_finish_kernel(LrefBatch b, unsigned long long n_units, unsigned long long n_tasks) {
  const int lane = threadIdx.x & 31;
  const int c = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (c >= b.n_clusters) return;
  // first task of the cluster's first unit / of the next cluster's first unit
  const unsigned long long u0 = b.unit_off[c], u1 = b.unit_off[c + 1];
  const unsigned long long t0 = u0 < n_units ? b.task_off[u0] : n_tasks, t1 = u1 < n_units ? b.task_off[u1] : n_tasks;
  const unsigned long long o0 = b.out_off[t0], o1 = b.out_off[t1];
  if (lane == 0) { b.r_off[c] = o0; if (c == b.n_clusters - 1) b.r_off[c + 1] = b.out_off[n_tasks]; }
  const unsigned long long n = o1 - o0;
  uint32_t *rb = b.rbox + 4 * c;
  if (n == 0 || o1 > b.out_cap) { if (lane < 4) rb[lane] = 0; if (lane == 0) b.eff[c] = 0.0f; return; }
  const int s = b.strand[c];
  const uint32_t readLen = b.rd[0].seq_len[b.read_id[c]];
  const uint32_t K = (uint32_t)b.small_k;
  uint32_t qS = 0xFFFFFFFFu, qE = 0, tS = 0xFFFFFFFFu, tE = 0;
  for (unsigned long long i = o0 + lane; i < o1; i += 32) {
    uint32_t qp = b.r_q[i];
    if (s == 1) { qp = readLen - (qp + K); b.r_q[i] = qp; }
    const uint32_t tp = b.r_t[i];
    qS = qp < qS ? qp : qS; qE = qp + K > qE ? qp + K : qE;
    tS = tp < tS ? tp : tS; tE = tp + K > tE ? tp + K : tE;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t a = __shfl_xor_sync(0xffffffffu, qS, o), bb = __shfl_xor_sync(0xffffffffu, qE, o);
    const uint32_t cc = __shfl_xor_sync(0xffffffffu, tS, o), d = __shfl_xor_sync(0xffffffffu, tE, o);
    qS = a < qS ? a : qS; qE = bb > qE ? bb : qE; tS = cc < tS ? cc : tS; tE = d > tE ? d : tE;
  }
  if (lane == 0) {
    rb[0] = qS; rb[1] = qE; rb[2] = tS; rb[3] = tE;
    const uint32_t den = (qE - qS) < (tE - tS) ? (qE - qS) : (tE - tS);
    b.eff[c] = __fdiv_rn((float)(unsigned long long)n, (float)den);      // ((float) matches.size()) / min(..): binary32, round to nearest
  }
}

}  // namespace lra
