// a20: RefineBreakpoint (reference RefineBreakpoint.h:212-462), batched over pairs of adjacent alignment segments of a read.
//   RSdp :150-197 (full DP, match 2 / mismatch -2 / indel -4, arrows diag > left > down), FindMax :199-210 (first maximum in row-major
//   order), StoreQScoreVect :120-147, TraceBack :92-118, PathToBlocks :50-83, PrependBlocks / AppendBlocks :6-47.
// For every pair the unaligned read span between the two segments (< 500 bases) is aligned against the genome flanking the left segment and
// against the genome flanking the right one; if the two extensions overlap in the read, the split that maximises the summed scores wins.
//
// Mapping: ONE WARP PER PAIR.  The two DP matrices ((span+1) x (tLen+1) <= 500 x 500 cells: int32 scores + 1-byte arrows, both needed in
// full by FindMax / StoreQScoreVect / TraceBack) live in a per-warp global slab that stays in L2; the warp sweeps anti-diagonals, lanes
// strided over the cells of one anti-diagonal (each cell depends on the two previous anti-diagonals only).  FindMax is a lane-strided scan
// with a (score, lowest index) reduction.  The path walks (<= 1000 steps) are replayed by lane 0.  This stage only runs for reads with
// several segments (< 1 % of the reference's time); the batch is the parallelism.
#pragma once
#include "lra_common.cuh"

namespace lra {

constexpr int kRbpMaxGap = 500;                       // MAX_GAP, RefineBreakpoint.h:257
constexpr int kRbpCells = kRbpMaxGap * kRbpMaxGap;    // (span + 1) * (tLen + 1) <= 500 * 500
constexpr int kRbpCap = 512;                          // blocks per side in the output (an extension has at most span blocks)

struct RbpBatch {
  int n_pairs;
  SeqView reads_fwd, reads_rc, genome;
  const uint32_t *lf, *ll, *rf, *rl;          // [n][3] first / last block (q, t, len) of the left / right alignment
  const uint8_t *lstrand, *rstrand;           // [n]
  const unsigned long long *read_off;         // [n] arena position of the read (same in both strand arenas)
  const uint32_t *read_len;                   // [n]
  const unsigned long long *lchrom_off, *rchrom_off;   // [n] arena position of the contig of each alignment
  const uint32_t *lchrom_len, *rchrom_len;    // [n]
  int32_t *score;                             // scratch [n_slabs][2][kRbpCells]
  uint8_t *path;                              // scratch [n_slabs][2][kRbpCells]
  int32_t *walk;                              // scratch [n_slabs][8][512]: qv / index of both sides, trace-backs
  int32_t *mode, *n_out;                      // [n][2]: 0 nothing, 1 append, 2 prepend; blocks to splice in
  uint32_t *bound;                            // [n][2][3] the alignment's boundary block after the splice
  uint32_t *out;                              // [n][2][kRbpCap][3]
  int32_t *refined;                           // [n] 1 if the pair was within MAX_GAP
};

struct RbpSide {            // one of the two DP problems of a pair
  const SeqView *qa, *ta;   // read strand arena, genome arena
  unsigned long long q0, t0;   // arena position of the first base of the query / target string (before reversal)
  int qs, ts;               // lengths
  bool rev;                 // strings reversed (backward extension)
};

__device__ __forceinline__ int rbp_q(const RbpSide &s, int j) { return seq_code(*s.qa, s.q0 + (unsigned long long)(s.rev ? s.qs - 1 - j : j)); }
__device__ __forceinline__ int rbp_t(const RbpSide &s, int i) { return seq_code(*s.ta, s.t0 + (unsigned long long)(s.rev ? s.ts - 1 - i : i)); }

// RSdp: fills score / path ((ts + 1) rows of qs + 1 cells); path codes 1 left, 2 down, 3 diag, 0 for the origin (-1 in the reference)
__device__ __forceinline__ void rbp_fill(const RbpSide &s, int32_t *score, uint8_t *path, uint8_t *qc, uint8_t *tc, int lane) {
  const int qs = s.qs, ts = s.ts, row = qs + 1;
  for (int j = lane; j < qs; j += 32) qc[j] = (uint8_t)rbp_q(s, j);
  for (int i = lane; i < ts; i += 32) tc[i] = (uint8_t)rbp_t(s, i);
  for (int j = lane; j <= qs; j += 32) { score[j] = -4 * j; path[j] = j ? 1 : 0; }
  for (int i = 1 + lane; i <= ts; i += 32) { score[i * row] = -4 * i; path[i * row] = 2; }
  __syncwarp();
  for (int d = 0; d <= qs + ts - 2; d++) {
    const int i0 = d - (qs - 1) > 0 ? d - (qs - 1) : 0, i1 = d < ts - 1 ? d : ts - 1;
    for (int i = i0 + lane; i <= i1; i += 32) {
      const int j = d - i;
      const int diagScore = score[i * row + j] + (qc[j] == tc[i] ? 2 : -2);
      const int leftScore = score[(i + 1) * row + j] - 4;
      const int downScore = score[i * row + (j + 1)] - 4;
      const int m = imax(diagScore, imax(leftScore, downScore));
      score[(i + 1) * row + (j + 1)] = m;
      path[(i + 1) * row + (j + 1)] = m == diagScore ? 3 : (m == leftScore ? 1 : 2);
    }
    __syncwarp();
  }
}

// FindMax: first maximum in row-major order -> (q, t) = (index % row - 1, index / row - 1)
__device__ __forceinline__ void rbp_find_max(const int32_t *score, int cells, int row, int lane, int &q, int &t) {
  int best = -(1 << 30), bi = 0x7FFFFFFF;
  for (int i = lane; i < cells; i += 32) { const int v = score[i]; if (v > best) { best = v; bi = i; } }
  for (int o = 16; o > 0; o >>= 1) {
    const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  t = bi / row - 1; q = bi % row - 1;
}

// lane 0: StoreQScoreVect
__device__ __forceinline__ void rbp_store_qscore(const int32_t *score, const uint8_t *path, int q, int t, int r, int32_t *qv, int32_t *index) {
  for (int i = 0; i < r - 1; i++) { qv[i] = 0; index[i] = 0; }
  int i = (t + 1) * r + q + 1;
  q++; t++;
  while (i > 0) {
    const int p = path[i];
    if (p == 3 || p == 1) { qv[q - 1] = score[i]; index[q - 1] = i; }
    if (p == 3) { q--; t--; }
    if (p == 1) q--;
    if (p == 2) t--;
    i = t * r + q;
  }
}
// lane 0: TraceBack (the ops in forward order in tb[0 .. n)), returns n
__device__ __forceinline__ int rbp_trace_back(const uint8_t *path, int q, int t, int r, int32_t *tb) {
  int n = 0;
  q++; t++;
  int i = t * r + q;
  while (q > 0 || t > 0) {
    const int p = path[i];
    if (p == 3) { q--; t--; tb[n++] = 3; }
    if (p == 1) { q--; tb[n++] = 1; }
    if (p == 2) { t--; tb[n++] = 2; }
    i = t * r + q;
  }
  for (int a = 0, b = n - 1; a < b; a++, b--) { const int x = tb[a]; tb[a] = tb[b]; tb[b] = x; }
  return n;
}
// lane 0: PathToBlocks + offsets + Prepend/AppendBlocks.  Returns the number of blocks written to out.
__device__ __forceinline__ int rbp_blocks(const int32_t *path, int n, int qOff, int tOff, int mode, uint32_t *bound, uint32_t *out) {
  int i = 0, q = 0, t = 0, nb = 0;
  while (i < n && path[i] != 3 && (path[i] == 1 || path[i] == 2)) { if (path[i] == 1) q++; if (path[i] == 2) t++; i++; }
  while (i < n) {
    const int qs = q, ts = t;
    while (i < n && path[i] == 3) { q++; t++; i++; }
    while (i < n && (path[i] == 1 || path[i] == 2)) { if (path[i] == 1) q++; if (path[i] == 2) t++; i++; }
    const int match = imin(q - qs, t - ts);
    if (match > 0 && nb < kRbpCap) { out[3 * nb] = (uint32_t)(qs + qOff); out[3 * nb + 1] = (uint32_t)(ts + tOff); out[3 * nb + 2] = (uint32_t)match; nb++; }
  }
  if (nb == 0) return 0;
  if (mode == 1) {          // AppendBlocks: a gapless continuation of the last block is merged into it
    if (bound[1] + bound[2] == out[1] && bound[0] + bound[2] == out[0]) {
      bound[2] += out[2];
      for (int k = 3; k < 3 * nb; k++) out[k - 3] = out[k];
      nb--;
    }
  } else {                  // PrependBlocks: a block ending exactly where the first block starts is merged into it
    const int last = nb - 1;
    if (out[3 * last + 1] + out[3 * last + 2] == bound[1] && out[3 * last] + out[3 * last + 2] == bound[0]) {
      bound[1] -= out[3 * last + 2]; bound[0] -= out[3 * last + 2]; bound[2] += out[3 * last + 2];
      nb--;
    }
  }
  return nb;
}

__global__ void __launch_bounds__(128) rbp_kernel(RbpBatch b, int first_pair, int n_here) {
  __shared__ uint8_t codes[4][2][512];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int slab = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (slab >= n_here) return;
  const int p = first_pair + slab;
  const uint32_t *lf = b.lf + 3 * p, *ll = b.ll + 3 * p, *rf = b.rf + 3 * p, *rl = b.rl + 3 * p;
  const int lstrand = b.lstrand[p], rstrand = b.rstrand[p];
  const int readLen = (int)b.read_len[p];
  const int lqs = (int)lf[0], lqe = (int)(ll[0] + ll[2]), lts = (int)lf[1], lte = (int)(ll[1] + ll[2]);
  const int rqs = (int)rf[0], rqe = (int)(rl[0] + rl[2]), rts = (int)rf[1], rte = (int)(rl[1] + rl[2]);
  const int flqe = lstrand == 0 ? lqe : readLen - lqs;
  const int frqs = rstrand == 0 ? rqs : readLen - rqe;
  if (lane == 0) { b.mode[2 * p] = 0; b.mode[2 * p + 1] = 0; b.n_out[2 * p] = 0; b.n_out[2 * p + 1] = 0; b.refined[p] = 0; }
  if (!(frqs > flqe && frqs - flqe < kRbpMaxGap)) return;
  const int span = frqs - flqe;
  const int lchromLen = (int)b.lchrom_len[p], rchromLen = (int)b.rchrom_len[p];
  RbpSide L, R;
  L.qa = lstrand == 0 ? &b.reads_fwd : &b.reads_rc; R.qa = rstrand == 0 ? &b.reads_fwd : &b.reads_rc;
  L.ta = &b.genome; R.ta = &b.genome;
  L.qs = span; R.qs = span;
  if (lstrand == 0) {
    L.q0 = b.read_off[p] + (unsigned long long)lqe; L.ts = imin(lchromLen - lte, span); L.t0 = b.lchrom_off[p] + (unsigned long long)lte; L.rev = false;
  } else {
    const int ltExtStart = imax(0, lts - span);
    L.q0 = b.read_off[p] + (unsigned long long)(lqs - span); L.ts = lts - ltExtStart; L.t0 = b.lchrom_off[p] + (unsigned long long)ltExtStart; L.rev = true;
  }
  if (rstrand == 0) {
    const int rtSpan = imin(rts, span);
    R.q0 = b.read_off[p] + (unsigned long long)(rqs - span); R.ts = rtSpan; R.t0 = b.rchrom_off[p] + (unsigned long long)(rts - rtSpan); R.rev = true;
  } else {
    int tSpan = span;
    if (rte + span >= rchromLen) tSpan = rchromLen - rte;
    R.q0 = b.read_off[p] + (unsigned long long)rqe; R.ts = tSpan; R.t0 = b.rchrom_off[p] + (unsigned long long)rte; R.rev = false;
  }
  if (L.ts < 0) L.ts = 0;
  if (R.ts < 0) R.ts = 0;
  int32_t *lScore = b.score + (size_t)slab * 2 * kRbpCells, *rScore = lScore + kRbpCells;
  uint8_t *lPath = b.path + (size_t)slab * 2 * kRbpCells, *rPath = lPath + kRbpCells;
  rbp_fill(L, lScore, lPath, codes[wib][0], codes[wib][1], lane);
  rbp_fill(R, rScore, rPath, codes[wib][0], codes[wib][1], lane);
  int mlq, mlt, mrq, mrt;
  rbp_find_max(lScore, (span + 1) * (L.ts + 1), span + 1, lane, mlq, mlt);
  rbp_find_max(rScore, (span + 1) * (R.ts + 1), span + 1, lane, mrq, mrt);
  if (lane == 0) {
    int32_t *W = b.walk + (size_t)slab * 8 * 512;
    if (!(mlq < span - mrq)) {
      int32_t *lqS = W, *rqS = W + 512, *lqI = W + 1024, *rqI = W + 1536;
      rbp_store_qscore(lScore, lPath, mlq, mlt, span + 1, lqS, lqI);
      rbp_store_qscore(rScore, rPath, mrq, mrt, span + 1, rqS, rqI);
      int maxScore = 0, maxL = 0, maxR = 0;
      for (int i = 0; i < span; i++)
        if (lqS[i] + rqS[span - i - 1] > maxScore) { maxScore = lqS[i] + rqS[span - i - 1]; maxL = i; maxR = span - i - 1; }
      mlq = maxL; mlt = lqI[maxL] / (span + 1) - 1; mrq = maxR; mrt = rqI[maxR] / (span + 1) - 1;
    }
    int32_t *ltb = W + 2048, *rtb = W + 3072;       // up to 2 * span <= 998 ops each
    const int nl = rbp_trace_back(lPath, mlq, mlt, span + 1, ltb), nr = rbp_trace_back(rPath, mrq, mrt, span + 1, rtb);
    int lqBlockStart, ltBlockStart, rqBlockStart, rtBlockStart;
    if (L.rev) { for (int a = 0, c = nl - 1; a < c; a++, c--) { const int x = ltb[a]; ltb[a] = ltb[c]; ltb[c] = x; } lqBlockStart = lqs - mlq - 1; ltBlockStart = lts - mlt - 1; }
    else { lqBlockStart = lqe; ltBlockStart = lte; }
    if (R.rev) { for (int a = 0, c = nr - 1; a < c; a++, c--) { const int x = rtb[a]; rtb[a] = rtb[c]; rtb[c] = x; } rqBlockStart = rqs - mrq - 1; rtBlockStart = rts - mrt - 1; }
    else { rqBlockStart = rqe; rtBlockStart = rte; }
    uint32_t *bound = b.bound + 6 * (size_t)p;
    uint32_t *out = b.out + (size_t)p * 2 * kRbpCap * 3;
    const int lmode = L.rev ? 2 : 1, rmode = R.rev ? 2 : 1;
    for (int k = 0; k < 3; k++) { bound[k] = L.rev ? lf[k] : ll[k]; bound[3 + k] = R.rev ? rf[k] : rl[k]; }
    b.n_out[2 * p] = rbp_blocks(ltb, nl, lqBlockStart, ltBlockStart, lmode, bound, out);
    b.n_out[2 * p + 1] = rbp_blocks(rtb, nr, rqBlockStart, rtBlockStart, rmode, bound + 3, out + kRbpCap * 3);
    b.mode[2 * p] = lmode; b.mode[2 * p + 1] = rmode;
    b.refined[p] = 1;
  }
}

}  // namespace lra
