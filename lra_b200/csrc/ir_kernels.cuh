// a19  IndelRefineAlignment -- banded 3-state (match / ins / del) DP over a ragged band, batched.
// Reference: IndelRefine.h:359-745 (matrix set-up :383-431, recurrence :438-622, traceback :626-674, path->blocks :702-745).
//
// A "group" is one banded window of a segment: rows t = 0..tLen-1 (target bases tStart+t), row t covers the query
// positions qS[t]..qE[t] (absolute read coordinates, both non-decreasing in t).  Scoring: gap = indel,
// gapOpen = 2*indel+1, gapExtend = 0 (IndelRefine.h:338-340).  Boundary cells: the first cell of every row >= 1 and
// the last cell of every row but the last are "bound" (never computed, never a valid predecessor); row 0 is a pure
// insertion ramp from its first cell.  Tie order M: diag, left, down, delClose, insClose; del/ins: open before extend.
//
// ir_dp_thread_kernel<WMAX>: ONE GROUP PER THREAD (long thin bands: ONT/CLR groups are 10^4 rows x 15..57 cells; CCS
// groups are 10^2 rows).  The previous row of M and D lives in shared memory ([cell][thread], conflict-free), updated in
// place; the query window is a packed shift register and the row's match mask is computed bit-parallel; arrows are 5 bits
// per cell (3 for M, 1 for del, 1 for ins), 6 cells per 32-bit word, streamed to HBM once per row.  This is the part
// of the hot path whose traffic is real: ~10^5..10^6 cells per group spill 0.67 B/cell of traceback.
// ir_dp_generic_kernel: any width, rows and byte arrows in a per-thread scratch slab (slow, exact, rare).
#pragma once
#include "aog_kernels.cuh"

namespace lra {

constexpr int kIrBad = -999999999;
enum IrArrow : int { IR_DIAG = 0, IR_LEFT = 1, IR_DOWN = 2, IR_DELCLOSE = 3, IR_INSCLOSE = 4 };

struct IrBatch {
  SeqView q, t;
  const uint32_t *q_base;   // arena offset of qSeq[0] (the read strand the blocks refer to)
  const uint32_t *t_base;   // arena offset of tSeq[0] (contig start)
  const int32_t *q_start, *t_start, *t_len, *q_seq_len, *t_seq_len;
  const uint32_t *band_off; // band[off .. off+tLen) = qS, band[off+tLen .. off+2 tLen) = qE
  const int32_t *band;
  int n_groups;
  int match, mismatch, gap;
  int32_t *n_blocks;
  unsigned long long *block_off;
  uint32_t *blocks;
  unsigned long long block_cap;
  unsigned long long *block_cursor;
  int *err;                 // bit0 block overflow, bit4 traceback failed, bit5 path end mismatch
  // scratch
  uint32_t *tb;             // arrows
  unsigned long long *tb_off;  // per group, in 32-bit words (assigned by the classify kernel)
  int32_t *max_width;       // per group
};

constexpr int kIrClsW24 = 0, kIrClsW64 = 1, kIrClsGeneric = 2, kIrClsWarp32 = 3, kIrClsPipe = 4, kIrClsWarp64 = 5, kIrNumCls = 6;
constexpr int kIrPipeSpan = 24;        // rows over which the band may advance at most kIrPipeMaxAdvance for the pipeline kernel
constexpr int kIrPipeMaxAdvance = 64;
constexpr int kIrWarpMinRows = 192;   // longer groups of width <= 32 go to the warp-per-group kernel
__host__ __device__ inline int ir_words(int wmax) { return (wmax + 5) / 6; }

// one warp per group: band maximum width -> class; reserves traceback storage; fills the planner histogram
__global__ void __launch_bounds__(128) ir_classify_kernel(IrBatch b, AogPlan *plan, uint32_t *bin_of_group, unsigned long long *tb_cursor,
                                                          unsigned long long *cells_total, int no_warp, int long_rows) {
  const int lane = threadIdx.x & 31;
  const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (g >= b.n_groups) return;
  const int rows = b.t_len[g];
  const int32_t *qS = b.band + b.band_off[g];
  const int32_t *qE = qS + rows;
  int mw = 0;
  long long cells = 0;
  int bad = 0;
  for (int r = lane; r < rows; r += 32) {
    const int w = qE[r] - qS[r] + 1;
    mw = imax(mw, w);
    cells += w;
    if (w < 2) bad = 1;
    if (r > 0 && (qS[r] < qS[r - 1] || qE[r] < qE[r - 1])) bad = 1;
    if (qE[imin(r + kIrPipeSpan, rows - 1)] - qS[r] > kIrPipeMaxAdvance) bad |= 2;   // too steep for the pipeline kernel's query ring
  }
  for (int o = 16; o > 0; o >>= 1) {
    mw = imax(mw, __shfl_down_sync(0xffffffffu, mw, o));
    cells += __shfl_down_sync(0xffffffffu, cells, o);
    bad |= __shfl_down_sync(0xffffffffu, bad, o);
  }
  if (lane == 0) {
    if ((bad & 1) || rows < 2) {
      atomicOr(b.err, 8);
      bin_of_group[g] = 0xFFFFFFFFu;
      b.n_blocks[g] = 0; b.block_off[g] = 0;
      return;
    }
    // no_warp: 1 = thread kernels only, 2 = long groups through the scan kernel only.  The longest groups (rows >= long_rows)
    // take the scan kernel, whose latency per row is lower; the row-pipeline kernel has the higher throughput for the rest.
    const bool longGroup = mw <= 32 && rows >= kIrWarpMinRows && no_warp != 1;
    // long groups of width 33..64 (CLR: refineBand 20): the two-cells-per-lane form of the warp kernel
    const bool longWide = mw > 32 && mw <= 64 && rows >= kIrWarpMinRows && no_warp != 1;
    const int cls = longGroup ? ((no_warp == 2 || (bad & 2) || rows >= long_rows) ? kIrClsWarp32 : kIrClsPipe) : longWide ? kIrClsWarp64 : mw <= 24 ? kIrClsW24 : (mw <= 64 ? kIrClsW64 : kIrClsGeneric);
    unsigned long long words = longGroup ? (unsigned long long)rows * 6ull + 8ull + 3ull * ((unsigned long long)rows + (unsigned long long)b.q_seq_len[g] + 4ull)
                               : longWide ? (unsigned long long)rows * 10ull + 8ull + 3ull * ((unsigned long long)rows + (unsigned long long)b.q_seq_len[g] + 4ull)
                               : cls == kIrClsGeneric ? ((unsigned long long)rows * (unsigned long long)mw + 3ull) / 4ull + 2ull * (unsigned long long)mw + 4ull
                                                           : (unsigned long long)rows * (unsigned long long)ir_words(cls == kIrClsW24 ? 24 : 64);
    words = (words + 1ull) & ~1ull;   // keep every group's arrows 8-byte aligned
    b.tb_off[g] = atomicAdd(tb_cursor, words);
    b.max_width[g] = mw;
    const int bucket = kAogBuckets - 1 - imin((longGroup || longWide) ? rows >> 8 : rows >> 3, kAogBuckets - 1);
    const uint32_t bin = (uint32_t)(cls * kAogBuckets + bucket);
    bin_of_group[g] = bin;
    atomicAdd(&plan->hist[bin], 1u);
    atomicAdd(&plan->cls_cells[cls], (unsigned long long)cells);
    // algorithmic bytes: 2-bit windows + 8 B per band row + 32 B descriptor + arrows spilled at 5 bit/cell
    atomicAdd(&plan->cls_bytes[cls], (unsigned long long)((b.q_seq_len[g] + 3) / 4 + (b.t_seq_len[g] + 3) / 4 + 8 * rows + 32) + (unsigned long long)((5 * cells + 7) / 8));
    atomicAdd(cells_total, (unsigned long long)cells);
  }
}

// Walks the stored arrows backwards.  WRITE == false counts the blocks the reference's path->blocks loop would emit,
// WRITE == true stores them (last block first) at out[nb-1], out[nb-2], ...  Returns the number of blocks or -1.
template <int WORDS, bool WRITE>
__device__ __forceinline__ int ir_walk(const uint32_t *tb, const int32_t *qS, const int32_t *qE, int rows, int tStart, uint32_t *out, int nb) {
  int t = rows - 1;
  int qs = qS[t];
  int x = qE[t] - qs;
  int mat = 0;            // 0 M, 1 del, 2 ins
  int run = 0;            // length of the diagonal run being walked (backwards)
  int lastOp = -1;        // type of the previous op in walk order (= the NEXT op in forward order)
  int count = 0;
  long guard = 0;
  const long guardMax = 4L * rows + 4L * (qE[rows - 1] - qS[0]) + 64;
  // forward semantics: blocks = one per diagonal run, plus a zero-length block for a gap run that directly follows
  // another gap run.  Walking backwards, an op `op` at a cell with forward start coordinates (qb, tb0) closes things:
  auto emit = [&](uint32_t qq, uint32_t tt, uint32_t ln) {
    if (WRITE) { const int r = nb - 1 - count; out[3 * r] = qq; out[3 * r + 1] = tt; out[3 * r + 2] = ln; }
    count++;
  };
  // coordinates of the forward-first cell of the current diagonal run are known when the run ends (see below)
  while (true) {
    if (++guard > guardMax) return -1;
    if (t == 0) {
      // row 0: `x` insertion ops back to the origin cell, then the origin diagonal (IndelRefine.h:420-424, :674)
      int op;
      if (x > 0) {
        if (mat != 0) return -1;
        // a left run of length x at row 0; forward start of its first op: q = qs+1, t = tStart+1
        if (run > 0) { emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), (uint32_t)run); run = 0; }
        else if (lastOp == IR_DOWN) emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), 0u);
        lastOp = IR_LEFT;
        x = 0;
      }
      op = IR_DIAG;  // origin
      if (lastOp != IR_DIAG && lastOp != -1 && run == 0) { /* gap run precedes: absorbed by this diagonal block */ }
      run += 1;
      emit((uint32_t)qs, (uint32_t)tStart, (uint32_t)run);
      (void)op;
      break;
    }
    const uint32_t word = tb[(unsigned long long)t * WORDS + (unsigned)(x / 6)];
    const int a5 = (int)((word >> (5 * (x % 6))) & 31u);
    int op;          // op pushed by the reference at this step (or -1 for a matrix hop)
    int nt = t, nx = x, nmat = mat;
    if (mat == 0) {
      const int a = a5 & 7;
      if (a == IR_DELCLOSE) { op = -1; nmat = 1; }
      else if (a == IR_INSCLOSE) { op = -1; nmat = 2; }
      else if (a == IR_DIAG) { op = IR_DIAG; nt = t - 1; }
      else if (a == IR_LEFT) { op = IR_LEFT; nx = x - 1; }
      else if (a == IR_DOWN) { op = IR_DOWN; nt = t - 1; }
      else return -1;
    } else if (mat == 1) {
      op = IR_DOWN; nmat = ((a5 >> 3) & 1) ? 1 : 0; nt = t - 1;
    } else {
      op = IR_LEFT; nmat = ((a5 >> 4) & 1) ? 2 : 0; nx = x - 1;
    }
    if (op >= 0) {
      const int q = qs + x;
      if (op == IR_DIAG) {
        run++;
      } else {
        // a gap op; forward start coordinates of THIS op: left: (q, tStart+t+1), down: (q+1, tStart+t)
        if (run > 0) {
          // the diagonal run that follows this gap op in forward order starts right after it
          const uint32_t qb = (uint32_t)(q + 1), tb0 = (uint32_t)(tStart + t + 1);
          emit(qb, tb0, (uint32_t)run);
          run = 0;
        } else if (lastOp != -1 && lastOp != op) {
          // forward order: this gap run, then a different gap run (lastOp) with no diagonal between -> the later run got a
          // zero-length block at its own start = the position after this op
          const uint32_t qb = (uint32_t)(q + 1), tb0 = (uint32_t)(tStart + t + 1);
          emit(qb, tb0, 0u);
        }
      }
      lastOp = op;
    }
    if (nt != t) {
      // moving up one row: diag -> (t-1, q-1), down -> (t-1, q)
      const int pqs = qS[t - 1];
      const int q = qs + x;
      nx = (op == IR_DIAG ? q - 1 : q) - pqs;
      qs = pqs;
    }
    t = nt; x = nx; mat = nmat;
    if (x < 0) return -1;
  }
  return count;
}

template <int WMAX> struct IrWin { typedef unsigned long long type; };
template <> struct IrWin<64> { typedef unsigned __int128 type; };

template <int WMAX>
__global__ void __launch_bounds__(64) ir_dp_thread_kernel(IrBatch b, AogPlan *plan, const uint32_t *sorted, int cls) {
  typedef typename IrWin<WMAX>::type BT;
  constexpr int WORDS = (WMAX + 5) / 6;
  __shared__ int sM[WMAX + 1][64];
  __shared__ int sD[WMAX + 1][64];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const uint32_t begin = plan->bin_start[cls * kAogBuckets];
  const uint32_t end = plan->bin_start[(cls + 1) * kAogBuckets];
  BT kOdd = 0;
#pragma unroll
  for (int i = 0; i < WMAX; i++) kOdd |= (BT)1 << (2 * i);
  const int match = b.match, mismatch = b.mismatch, gap = b.gap, gapOpen = 2 * b.gap + 1;

  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&plan->work[cls], 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (begin + base >= end) break;
    const uint32_t idx = begin + base + lane;
    const bool active = idx < end;
    int g = 0, nb = 0, rows = 0, tStart = 0;
    const int32_t *qS = nullptr, *qE = nullptr;
    const uint32_t *tbp = nullptr;
    if (active) {
      g = (int)sorted[idx];
      rows = b.t_len[g];
      tStart = b.t_start[g];
      qS = b.band + b.band_off[g];
      qE = qS + rows;
      uint32_t *tbw = b.tb + b.tb_off[g];
      tbp = tbw;
      // ---- row 0: M[0] = 0 (done), ramp, last cell bound; D = BAD
      int qsPrev = qS[0], qePrev = qE[0];
      int lenPrev = qePrev - qsPrev + 1;
      for (int x = 0; x < lenPrev; x++) { sM[x][tid] = x * gap; sD[x][tid] = kIrBad; }
      SeqStream qst, tst;
      qst.init(b.q, (uint64_t)b.q_base[g] + (uint64_t)qsPrev);
      tst.init(b.t, (uint64_t)(uint32_t)(b.t_base[g] + (uint32_t)(tStart + 1)));  // 32-bit wrap-around: t_base may be 'negative'
      BT qw = 0, qn = 0;
      for (int x = 0; x < lenPrev; x++) {
        const int code = qst.next();
        qw |= (BT)(code & 3) << (2 * x);
        qn |= (BT)(code == 4 ? 1 : 0) << (2 * x);
      }
      for (int t = 1; t < rows; t++) {
        const int qs = qS[t], qe = qE[t];
        const int len = qe - qs + 1;
        const int off = qs - qsPrev;
        const int rowEnd = (t == rows - 1) ? len : len - 1;
        // slide the query window
        if (off >= WMAX) { qw = 0; qn = 0; } else { qw >>= (2 * off); qn >>= (2 * off); }
        {
          int qnext = qePrev + 1;
          for (; qnext < qs; qnext++) (void)qst.next();  // (never in practice: bands overlap)
          for (; qnext <= qe; qnext++) {
            const int code = qst.next();
            const int xx = qnext - qs;
            qw |= (BT)(code & 3) << (2 * xx);
            qn |= (BT)(code == 4 ? 1 : 0) << (2 * xx);
          }
        }
        const int tc = tst.next();
        BT e;
        if (tc == 4) e = qn;
        else { BT xr = qw ^ ((BT)tc * kOdd); e = ~(xr | (xr >> 1)) & kOdd & ~qn; }
        // in-place row update, left to right
        int Mdiag = (off < lenPrev) ? sM[off][tid] : kIrBad;  // prev-row cell under x = 0, diagonal predecessor of x = 1
        int Mleft = kIrBad, Ileft = kIrBad;
        sM[0][tid] = kIrBad;             // first cell of rows >= 1 is bound
        uint32_t words[WORDS];
#pragma unroll
        for (int w = 0; w < WORDS; w++) words[w] = 0;
        const bool prevIsRow0 = (t == 1);
#pragma unroll
        for (int x = 1; x < WMAX; x++) {
          if (x >= rowEnd) break;
          const int xp = x + off;
          const bool upIn = xp <= lenPrev - 1;
          const bool upOk = xp < lenPrev - 1;     // in range and not the bound last cell of the (non-final) previous row
          const bool diagOk = upIn && !(xp - 1 == 0 && !prevIsRow0);
          int Mup = kIrBad, Dup = kIrBad;
          if (upIn) { Mup = sM[xp][tid]; Dup = sD[xp][tid]; }
          const int delOpen = upOk ? Mup + gapOpen : kIrBad;
          const int delExt = upOk ? Dup : kIrBad;
          const int D = imax(delOpen, delExt);
          const int dbit = (D == delOpen) ? 0 : 1;
          const int insOpen = Mleft + gapOpen;
          const int I = imax(insOpen, Ileft);
          const int ibit = (I == insOpen) ? 0 : 1;
          const int mS = diagOk ? Mdiag + (((e >> (2 * x)) & 1) ? match : mismatch) : kIrBad;
          const int iS = Mleft + gap;
          const int dS = upOk ? Mup + gap : kIrBad;
          const int mx = imax(imax(mS, iS), imax(dS, imax(D, I)));
          const int a = (mx == mS) ? IR_DIAG : (mx == iS) ? IR_LEFT : (mx == dS) ? IR_DOWN : (mx == D) ? IR_DELCLOSE : IR_INSCLOSE;
          words[x / 6] |= (uint32_t)(a | (dbit << 3) | (ibit << 4)) << (5 * (x % 6));
          sM[x][tid] = mx;
          sD[x][tid] = D;
          Mdiag = Mup;
          Mleft = mx;
          Ileft = I;
        }
#pragma unroll
        for (int w = 0; w < WORDS; w++)
          if (w * 6 < rowEnd) tbw[(unsigned long long)t * WORDS + w] = words[w];
        qsPrev = qs; qePrev = qe; lenPrev = len;
      }
      nb = ir_walk<WORDS, false>(tbp, qS, qE, rows, tStart, nullptr, 0);
      if (nb < 0) { atomicOr(b.err, 16); nb = 0; }
    }
    const unsigned long long slot = aog_reserve_blocks(AogBatch{b.q, b.t, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, nullptr,
                                                                nullptr, b.blocks, b.block_cap, b.block_cursor, b.err},
                                                       nb, lane, &plan->cls_blocks[cls]);
    if (active) {
      b.n_blocks[g] = nb;
      b.block_off[g] = slot;
      if (slot != ~0ull && nb > 0) ir_walk<WORDS, true>(tbp, qS, qE, rows, tStart, b.blocks + 3ull * slot, nb);
    }
  }
}


// ---------------------------------------------------------------------------------------------------- warp per group
// Long groups (ONT/CLR: 10^3..10^4 rows).  Lane x owns band cell x of the current row (width <= 32).  With
//   a[x] = max(match, down, delClose)            (candidates that come from the previous row),
// the in-row recurrences  I[x] = max(M[x-1]+gapOpen, I[x-1]),  M[x] = max(a[x], M[x-1]+gap, I[x])  have the closed form
//   M[x] = max( max_{y<=x} a[y] + (x-y)*gap ,  max_{y<x} a[y] + gapOpen ,  BAD ),   I[x] = max(BAD, max_{y<x} a[y] + gapOpen)
// (gapExtend = 0 and gapOpen < gap < 0), i.e. two independent 5-step prefix-max scans per row instead of a serial chain.
// Arrows are then decided per cell from the exact M[x-1], I[x].  The previous row is exchanged through shared memory;
// arrows are stored as 5 bit-planes (one ballot each) = 5 words per row; the traceback is walked by the whole warp with
// 32 rows of planes held in registers at a time.
constexpr int kIrNeg = -1073741824;

// MODE 0: count only; MODE 1: write forward (needs nb); MODE 2: write in walk order (last block first) at out[0..)
template <int MODE>
__device__ __forceinline__ int ir_walk_planes(const uint32_t *tb, const int32_t *qS, const int32_t *qE, int rows, int tStart, uint32_t *out,
                                              int nb, int lane) {
  int t = rows - 1;
  int qs = qS[t];
  int x = qE[t] - qs;
  int mat = 0, run = 0, lastOp = -1, count = 0;
  long guard = 0;
  const long guardMax = 4L * rows + 4L * (qE[rows - 1] - qS[0]) + 64;
  auto emit = [&](uint32_t qq, uint32_t tt, uint32_t ln) {
    if (MODE != 0 && lane == 0) { const int r = (MODE == 1) ? nb - 1 - count : count; out[3 * r] = qq; out[3 * r + 1] = tt; out[3 * r + 2] = ln; }
    count++;
  };
  int top = -1;                 // rows [top-31, top] are held: lane l has row top-l
  uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0;
  int rqs = 0;
  while (true) {
    if (++guard > guardMax) return -1;
    if (t == 0) {
      if (x > 0) {
        if (mat != 0) return -1;
        if (run > 0) { emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), (uint32_t)run); run = 0; }
        else if (lastOp == IR_DOWN) emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), 0u);
        lastOp = IR_LEFT; x = 0;
      }
      run += 1;
      emit((uint32_t)qs, (uint32_t)tStart, (uint32_t)run);
      break;
    }
    if (top < 0 || t > top || t < top - 31) {
      top = t;
      const int r = top - lane;
      if (r >= 1) {
        const uint32_t *w = tb + (unsigned long long)r * 5ull;
        p0 = w[0]; p1 = w[1]; p2 = w[2]; p3 = w[3]; p4 = w[4];
      }
      rqs = (r >= 1) ? qS[r - 1] : 0;   // qS of the row above row r
    }
    const int src = top - t;
    const uint32_t w0 = __shfl_sync(0xffffffffu, p0, src), w1 = __shfl_sync(0xffffffffu, p1, src), w2 = __shfl_sync(0xffffffffu, p2, src);
    const uint32_t w3 = __shfl_sync(0xffffffffu, p3, src), w4 = __shfl_sync(0xffffffffu, p4, src);
    const int pqs = __shfl_sync(0xffffffffu, rqs, src);
    const int a = (int)(((w0 >> x) & 1u) | (((w1 >> x) & 1u) << 1) | (((w2 >> x) & 1u) << 2));
    const int dbit = (int)((w3 >> x) & 1u), ibit = (int)((w4 >> x) & 1u);
    int op, nt = t, nx = x, nmat = mat;
    if (mat == 0) {
      if (a == IR_DELCLOSE) { op = -1; nmat = 1; }
      else if (a == IR_INSCLOSE) { op = -1; nmat = 2; }
      else if (a == IR_DIAG) { op = IR_DIAG; nt = t - 1; }
      else if (a == IR_LEFT) { op = IR_LEFT; nx = x - 1; }
      else if (a == IR_DOWN) { op = IR_DOWN; nt = t - 1; }
      else return -1;
    } else if (mat == 1) { op = IR_DOWN; nmat = dbit ? 1 : 0; nt = t - 1; }
    else { op = IR_LEFT; nmat = ibit ? 2 : 0; nx = x - 1; }
    if (op >= 0) {
      const int q = qs + x;
      if (op == IR_DIAG) run++;
      else {
        if (run > 0) { emit((uint32_t)(q + 1), (uint32_t)(tStart + t + 1), (uint32_t)run); run = 0; }
        else if (lastOp != -1 && lastOp != op) emit((uint32_t)(q + 1), (uint32_t)(tStart + t + 1), 0u);
      }
      lastOp = op;
    }
    if (nt != t) { const int q = qs + x; nx = (op == IR_DIAG ? q - 1 : q) - pqs; qs = pqs; }
    t = nt; x = nx; mat = nmat;
    if (x < 0 || x > 31) return -1;
  }
  return count;
}

__global__ void __launch_bounds__(128) ir_dp_warp_kernel(IrBatch b, AogPlan *plan, const uint32_t *sorted) {
  constexpr int cls = kIrClsWarp32;
  __shared__ int sM[4][33];
  __shared__ int sD[4][33];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  int *rowM = sM[wib], *rowD = sD[wib];
  const uint32_t begin = plan->bin_start[cls * kAogBuckets];
  const uint32_t end = plan->bin_start[(cls + 1) * kAogBuckets];
  const int match = b.match, mismatch = b.mismatch, gap = b.gap, gapOpen = 2 * b.gap + 1;
  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&plan->work[cls], 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (begin + w >= end) break;
    const int g = (int)sorted[begin + w];
    const int rows = b.t_len[g];
    const int tStart = b.t_start[g];
    const int32_t *qS = b.band + b.band_off[g];
    const int32_t *qE = qS + rows;
    uint32_t *tbw = b.tb + b.tb_off[g];
    const uint64_t qbase = (uint64_t)b.q_base[g];
    const uint32_t tbase = b.t_base[g];
    int qsPrev = qS[0];
    int lenPrev = qE[0] - qsPrev + 1;
    __syncwarp();
    rowM[lane] = lane * gap;
    rowD[lane] = kIrBad;
    __syncwarp();
    auto qcode_at = [&](int pos) -> int { const uint64_t p = qbase + (uint64_t)pos; return p < b.q.n ? seq_code(b.q, p) : 5; };
    for (int t0 = 1; t0 < rows; t0 += 32) {
      // band limits and target codes of the next 32 rows, one row per lane; query codes of the positions those rows
      // can touch, three per lane (the band advances about one position per row, so 96 positions cover 32 rows)
      const int rr = t0 + lane;
      const int pqs = rr < rows ? qS[rr] : 0, pqe = rr < rows ? qE[rr] : 0;
      const int ptc = rr < rows ? seq_code(b.t, (uint64_t)(uint32_t)(tbase + (uint32_t)(tStart + rr))) : 5;
      const int nrow = imin(32, rows - t0);
      const int p0 = __shfl_sync(0xffffffffu, pqs, 0);
      const int pend = __shfl_sync(0xffffffffu, pqe, nrow - 1);
      const bool inRegs = pend - p0 < 96;
      int qr0 = 5, qr1 = 5, qr2 = 5;
      if (inRegs) { qr0 = qcode_at(p0 + lane); qr1 = qcode_at(p0 + 32 + lane); qr2 = qcode_at(p0 + 64 + lane); }
      for (int l = 0; l < nrow; l++) {
        const int t = t0 + l;
        const int qs = __shfl_sync(0xffffffffu, pqs, l), qe = __shfl_sync(0xffffffffu, pqe, l), tc = __shfl_sync(0xffffffffu, ptc, l);
        const int len = qe - qs + 1;
        const int off = qs - qsPrev;
        const int rowEnd = (t == rows - 1) ? len : len - 1;
        int qc;
        if (inRegs) {
          const int idx = qs + lane - p0;          // 0 .. 95+31: lanes beyond the band read garbage that is never used
          const int src = idx & 31;
          const int c0 = __shfl_sync(0xffffffffu, qr0, src), c1 = __shfl_sync(0xffffffffu, qr1, src), c2 = __shfl_sync(0xffffffffu, qr2, src);
          qc = idx < 32 ? c0 : (idx < 64 ? c1 : (idx < 96 ? c2 : 5));
        } else {
          qc = qcode_at(qs + lane);
        }
        const int x = lane;
        const int xp = x + off;
        const bool upIn = xp <= lenPrev - 1;
        const bool upOk = xp < lenPrev - 1;
        const bool diagOk = upIn && !(xp - 1 == 0 && t != 1);
        const int Mup = upIn ? rowM[xp] : kIrBad;
        const int Dup = upIn ? rowD[xp] : kIrBad;
        const int Mdg = (upIn && xp >= 1) ? rowM[xp - 1] : kIrBad;
        const bool valid = x >= 1 && x < rowEnd;
        __syncwarp();
        const int delOpen = upOk ? Mup + gapOpen : kIrBad;
        const int delExt = upOk ? Dup : kIrBad;
        const int D = imax(delOpen, delExt);
        const int dbit = (D == delOpen) ? 0 : 1;
        const int mS = diagOk ? Mdg + (qc == tc ? match : mismatch) : kIrBad;
        const int dS = upOk ? Mup + gap : kIrBad;
        const int a = valid ? imax(imax(mS, dS), D) : (x == 0 ? kIrBad : kIrNeg);
        int u = a - x * gap;         // L[x] = x*gap + max_{y<=x}(a[y] - y*gap)
        int pm = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int uu = __shfl_up_sync(0xffffffffu, u, o);
          const int pp = __shfl_up_sync(0xffffffffu, pm, o);
          if (lane >= o) { u = imax(u, uu); pm = imax(pm, pp); }
        }
        int pmx = __shfl_up_sync(0xffffffffu, pm, 1);
        if (lane == 0) pmx = kIrNeg;
        const int L = u + x * gap;
        const int I = imax(kIrBad, pmx + gapOpen);
        const int M = (x == 0) ? kIrBad : imax(imax(L, pmx + gapOpen), kIrBad);
        int Mleft = __shfl_up_sync(0xffffffffu, M, 1);
        if (lane == 0) Mleft = kIrBad;
        const int iS = Mleft + gap;
        const int ibit = (I == Mleft + gapOpen) ? 0 : 1;
        const int arrow = (M == mS) ? IR_DIAG : (M == iS) ? IR_LEFT : (M == dS) ? IR_DOWN : (M == D) ? IR_DELCLOSE : IR_INSCLOSE;
        const uint32_t b0 = __ballot_sync(0xffffffffu, valid && (arrow & 1));
        const uint32_t b1 = __ballot_sync(0xffffffffu, valid && (arrow & 2));
        const uint32_t b2 = __ballot_sync(0xffffffffu, valid && (arrow & 4));
        const uint32_t b3 = __ballot_sync(0xffffffffu, valid && dbit);
        const uint32_t b4 = __ballot_sync(0xffffffffu, valid && ibit);
        if (lane == 0) { uint32_t *w5 = tbw + (unsigned)t * 5u; w5[0] = b0; w5[1] = b1; w5[2] = b2; w5[3] = b3; w5[4] = b4; }
        rowM[lane] = M;
        rowD[lane] = valid ? D : kIrBad;
        __syncwarp();
        qsPrev = qs; lenPrev = len;
      }
    }
    __syncwarp();
    // one traceback walk: blocks come out last-first into the group's scratch, then are copied in forward order
    uint32_t *rb = tbw + (unsigned long long)rows * 5ull + 8ull;
    int nb = ir_walk_planes<2>(tbw, qS, qE, rows, tStart, rb, 0, lane);
    if (nb < 0) { if (lane == 0) atomicOr(b.err, 16); nb = 0; }
    unsigned long long slot = aog_reserve_blocks(AogBatch{b.q, b.t, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, nullptr,
                                                          nullptr, b.blocks, b.block_cap, b.block_cursor, b.err},
                                                 lane == 0 ? nb : 0, lane, &plan->cls_blocks[cls]);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (lane == 0) { b.n_blocks[g] = nb; b.block_off[g] = slot; }
    __syncwarp();
    if (slot != ~0ull) {
      uint32_t *out = b.blocks + 3ull * slot;
      for (int i = lane; i < 3 * nb; i += 32) { const int r = i / 3, c = i - 3 * r; out[i] = rb[3 * (nb - 1 - r) + c]; }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------- warp per group, width <= 64
// The same kernel with TWO band cells per lane (cell x = lane + 32 c): CLR's refineBand 20 makes rows of 41..64 cells.  The two prefix-max
// scans of a row run over the low half, then over the high half seeded with the low half's totals; arrows are 5 bit-planes of 64 bits
// (two ballots each) = 10 words per row.
template <int MODE>
__device__ __forceinline__ int ir_walk_planes64(const uint32_t *tb, const int32_t *qS, const int32_t *qE, int rows, int tStart, uint32_t *out, int nb, int lane) {
  int t = rows - 1;
  int qs = qS[t];
  int x = qE[t] - qs;
  int mat = 0, run = 0, lastOp = -1, count = 0;
  long guard = 0;
  const long guardMax = 4L * rows + 4L * (qE[rows - 1] - qS[0]) + 64;
  auto emit = [&](uint32_t qq, uint32_t tt, uint32_t ln) {
    if (MODE != 0 && lane == 0) { const int r = (MODE == 1) ? nb - 1 - count : count; out[3 * r] = qq; out[3 * r + 1] = tt; out[3 * r + 2] = ln; }
    count++;
  };
  int top = -1;                 // rows [top-31, top] are held: lane l has row top-l
  unsigned long long p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0;
  int rqs = 0;
  while (true) {
    if (++guard > guardMax) return -1;
    if (t == 0) {
      if (x > 0) {
        if (mat != 0) return -1;
        if (run > 0) { emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), (uint32_t)run); run = 0; }
        else if (lastOp == IR_DOWN) emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), 0u);
        lastOp = IR_LEFT; x = 0;
      }
      run += 1;
      emit((uint32_t)qs, (uint32_t)tStart, (uint32_t)run);
      break;
    }
    if (top < 0 || t > top || t < top - 31) {
      top = t;
      const int r = top - lane;
      if (r >= 1) {
        const uint32_t *w = tb + (unsigned long long)r * 10ull;
        p0 = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32); p1 = (unsigned long long)w[2] | ((unsigned long long)w[3] << 32);
        p2 = (unsigned long long)w[4] | ((unsigned long long)w[5] << 32); p3 = (unsigned long long)w[6] | ((unsigned long long)w[7] << 32);
        p4 = (unsigned long long)w[8] | ((unsigned long long)w[9] << 32);
      }
      rqs = (r >= 1) ? qS[r - 1] : 0;
    }
    const int src = top - t;
    const unsigned long long w0 = __shfl_sync(0xffffffffu, p0, src), w1 = __shfl_sync(0xffffffffu, p1, src), w2 = __shfl_sync(0xffffffffu, p2, src);
    const unsigned long long w3 = __shfl_sync(0xffffffffu, p3, src), w4 = __shfl_sync(0xffffffffu, p4, src);
    const int pqs = __shfl_sync(0xffffffffu, rqs, src);
    const int a = (int)(((w0 >> x) & 1ull) | (((w1 >> x) & 1ull) << 1) | (((w2 >> x) & 1ull) << 2));
    const int dbit = (int)((w3 >> x) & 1ull), ibit = (int)((w4 >> x) & 1ull);
    int op, nt = t, nx = x, nmat = mat;
    if (mat == 0) {
      if (a == IR_DELCLOSE) { op = -1; nmat = 1; }
      else if (a == IR_INSCLOSE) { op = -1; nmat = 2; }
      else if (a == IR_DIAG) { op = IR_DIAG; nt = t - 1; }
      else if (a == IR_LEFT) { op = IR_LEFT; nx = x - 1; }
      else if (a == IR_DOWN) { op = IR_DOWN; nt = t - 1; }
      else return -1;
    } else if (mat == 1) { op = IR_DOWN; nmat = dbit ? 1 : 0; nt = t - 1; }
    else { op = IR_LEFT; nmat = ibit ? 2 : 0; nx = x - 1; }
    if (op >= 0) {
      const int q = qs + x;
      if (op == IR_DIAG) run++;
      else {
        if (run > 0) { emit((uint32_t)(q + 1), (uint32_t)(tStart + t + 1), (uint32_t)run); run = 0; }
        else if (lastOp != -1 && lastOp != op) emit((uint32_t)(q + 1), (uint32_t)(tStart + t + 1), 0u);
      }
      lastOp = op;
    }
    if (nt != t) { const int q = qs + x; nx = (op == IR_DIAG ? q - 1 : q) - pqs; qs = pqs; }
    t = nt; x = nx; mat = nmat;
    if (x < 0 || x > 63) return -1;
  }
  return count;
}

__global__ void __launch_bounds__(128) ir_dp_warp64_kernel(IrBatch b, AogPlan *plan, const uint32_t *sorted) {
  constexpr int cls = kIrClsWarp64;
  __shared__ int sM[4][66];
  __shared__ int sD[4][66];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  int *rowM = sM[wib], *rowD = sD[wib];
  const uint32_t begin = plan->bin_start[cls * kAogBuckets];
  const uint32_t end = plan->bin_start[(cls + 1) * kAogBuckets];
  const int match = b.match, mismatch = b.mismatch, gap = b.gap, gapOpen = 2 * b.gap + 1;
  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&plan->work[cls], 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (begin + w >= end) break;
    const int g = (int)sorted[begin + w];
    const int rows = b.t_len[g];
    const int tStart = b.t_start[g];
    const int32_t *qS = b.band + b.band_off[g];
    const int32_t *qE = qS + rows;
    uint32_t *tbw = b.tb + b.tb_off[g];
    const uint64_t qbase = (uint64_t)b.q_base[g];
    const uint32_t tbase = b.t_base[g];
    int qsPrev = qS[0];
    int lenPrev = qE[0] - qsPrev + 1;
    __syncwarp();
    rowM[lane] = lane * gap; rowM[lane + 32] = (lane + 32) * gap;
    rowD[lane] = kIrBad; rowD[lane + 32] = kIrBad;
    __syncwarp();
    auto qcode_at = [&](int pos) -> int { const uint64_t p = qbase + (uint64_t)pos; return p < b.q.n ? seq_code(b.q, p) : 5; };
    for (int t0 = 1; t0 < rows; t0 += 32) {
      const int rr = t0 + lane;
      const int pqs = rr < rows ? qS[rr] : 0, pqe = rr < rows ? qE[rr] : 0;
      const int ptc = rr < rows ? seq_code(b.t, (uint64_t)(uint32_t)(tbase + (uint32_t)(tStart + rr))) : 5;
      const int nrow = imin(32, rows - t0);
      for (int l = 0; l < nrow; l++) {
        const int t = t0 + l;
        const int qs = __shfl_sync(0xffffffffu, pqs, l), qe = __shfl_sync(0xffffffffu, pqe, l), tc = __shfl_sync(0xffffffffu, ptc, l);
        const int len = qe - qs + 1;
        const int off = qs - qsPrev;
        const int rowEnd = (t == rows - 1) ? len : len - 1;
        int a[2], D[2], dbit[2], mS[2], dS[2]; bool valid[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int x = lane + 32 * c;
          const int qc = qcode_at(qs + x);
          const int xp = x + off;
          const bool upIn = xp <= lenPrev - 1;
          const bool upOk = xp < lenPrev - 1;
          const bool diagOk = upIn && !(xp - 1 == 0 && t != 1);
          const int Mup = upIn ? rowM[xp] : kIrBad;
          const int Dup = upIn ? rowD[xp] : kIrBad;
          const int Mdg = (upIn && xp >= 1) ? rowM[xp - 1] : kIrBad;
          valid[c] = x >= 1 && x < rowEnd;
          const int delOpen = upOk ? Mup + gapOpen : kIrBad;
          const int delExt = upOk ? Dup : kIrBad;
          D[c] = imax(delOpen, delExt);
          dbit[c] = (D[c] == delOpen) ? 0 : 1;
          mS[c] = diagOk ? Mdg + (qc == tc ? match : mismatch) : kIrBad;
          dS[c] = upOk ? Mup + gap : kIrBad;
          a[c] = valid[c] ? imax(imax(mS[c], dS[c]), D[c]) : (x == 0 ? kIrBad : kIrNeg);
        }
        __syncwarp();
        // inclusive prefix maxima over the 64 cells: u = a - x gap (for L), pm = a (for I and the gap-open term)
        int u0 = a[0] - lane * gap, pm0 = a[0];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int uu = __shfl_up_sync(0xffffffffu, u0, o), pp = __shfl_up_sync(0xffffffffu, pm0, o);
          if (lane >= o) { u0 = imax(u0, uu); pm0 = imax(pm0, pp); }
        }
        const int cu = __shfl_sync(0xffffffffu, u0, 31), cp = __shfl_sync(0xffffffffu, pm0, 31);
        int u1 = a[1] - (lane + 32) * gap, pm1 = a[1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int uu = __shfl_up_sync(0xffffffffu, u1, o), pp = __shfl_up_sync(0xffffffffu, pm1, o);
          if (lane >= o) { u1 = imax(u1, uu); pm1 = imax(pm1, pp); }
        }
        u1 = imax(u1, cu); pm1 = imax(pm1, cp);
        int pmx0 = __shfl_up_sync(0xffffffffu, pm0, 1); if (lane == 0) pmx0 = kIrNeg;
        int pmx1 = __shfl_up_sync(0xffffffffu, pm1, 1); if (lane == 0) pmx1 = cp;
        int M[2], I[2];
        { const int L = u0 + lane * gap; I[0] = imax(kIrBad, pmx0 + gapOpen); M[0] = (lane == 0) ? kIrBad : imax(imax(L, pmx0 + gapOpen), kIrBad); }
        { const int L = u1 + (lane + 32) * gap; I[1] = imax(kIrBad, pmx1 + gapOpen); M[1] = imax(imax(L, pmx1 + gapOpen), kIrBad); }
        int Ml0 = __shfl_up_sync(0xffffffffu, M[0], 1); if (lane == 0) Ml0 = kIrBad;
        const int M0last = __shfl_sync(0xffffffffu, M[0], 31);
        int Ml1 = __shfl_up_sync(0xffffffffu, M[1], 1); if (lane == 0) Ml1 = M0last;
        const int Mleft[2] = {Ml0, Ml1};
        uint32_t pl[10];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int iS = Mleft[c] + gap;
          const int ibit = (I[c] == Mleft[c] + gapOpen) ? 0 : 1;
          const int arrow = (M[c] == mS[c]) ? IR_DIAG : (M[c] == iS) ? IR_LEFT : (M[c] == dS[c]) ? IR_DOWN : (M[c] == D[c]) ? IR_DELCLOSE : IR_INSCLOSE;
          pl[0 + c] = __ballot_sync(0xffffffffu, valid[c] && (arrow & 1));
          pl[2 + c] = __ballot_sync(0xffffffffu, valid[c] && (arrow & 2));
          pl[4 + c] = __ballot_sync(0xffffffffu, valid[c] && (arrow & 4));
          pl[6 + c] = __ballot_sync(0xffffffffu, valid[c] && dbit[c]);
          pl[8 + c] = __ballot_sync(0xffffffffu, valid[c] && ibit);
        }
        if (lane < 10) {
          uint32_t v = 0;
#pragma unroll
          for (int k = 0; k < 10; k++) if (lane == k) v = pl[k];
          tbw[(unsigned long long)t * 10ull + (unsigned)lane] = v;
        }
        rowM[lane] = M[0]; rowM[lane + 32] = M[1];
        rowD[lane] = valid[0] ? D[0] : kIrBad; rowD[lane + 32] = valid[1] ? D[1] : kIrBad;
        __syncwarp();
        qsPrev = qs; lenPrev = len;
      }
    }
    __syncwarp();
    uint32_t *rb = tbw + (unsigned long long)rows * 10ull + 8ull;
    int nb = ir_walk_planes64<2>(tbw, qS, qE, rows, tStart, rb, 0, lane);
    if (nb < 0) { if (lane == 0) atomicOr(b.err, 16); nb = 0; }
    unsigned long long slot = aog_reserve_blocks(AogBatch{b.q, b.t, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, nullptr,
                                                          nullptr, b.blocks, b.block_cap, b.block_cursor, b.err},
                                                 lane == 0 ? nb : 0, lane, &plan->cls_blocks[cls]);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (lane == 0) { b.n_blocks[g] = nb; b.block_off[g] = slot; }
    __syncwarp();
    if (slot != ~0ull) {
      uint32_t *out = b.blocks + 3ull * slot;
      for (int i = lane; i < 3 * nb; i += 32) { const int r = i / 3, c = i - 3 * r; out[i] = rb[3 * (nb - 1 - r) + c]; }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------- 8-lane row pipeline
// Long groups, default kernel.  EIGHT lanes own one group and four groups share a warp.  Sub-lane j computes the rows
// t = j+1 (mod 8) cell by cell -- the in-row chain (I, M) is a plain serial recurrence again, no scans -- one row behind its
// left neighbour: cell (t, x) may run once row t-1 has produced cell x+off (off = qS[t]-qS[t-1]).  Each lane does up to two
// cells per lock step; rows are 15..20 cells, so all eight rows in flight keep their lanes busy and a warp retires up to 64
// cells per step.  Rows are exchanged through shared memory (one M/D row buffer per lane, reused every 8 rows: a lane may
// overwrite cell x only after the right neighbour, which still reads the old row, is past it), progress (row id, last cell)
// through shuffles.  Row descriptors (band limits + the target base's packed words) and the packed query window are
// prefetched into shared-memory rings with cp.async, a whole row-round ahead, so no global load sits in the dependent chain.
// Arrows: 5 bits per cell (3 M, 1 del, 1 ins), 12 cells per 64-bit word, 3 words per row, cell x >= 1 of a row in word
// (x-1)/12 at bit 5*(11-(x-1)%12).  The traceback is walked by the 8 lanes with 8 rows of arrow words in registers and the
// next 8 rows prefetched behind them.
__device__ __forceinline__ uint32_t ir_pipe_code(unsigned long long a0, unsigned long long a1, unsigned long long a2, int x) {
  if (x <= 0) return 0u;
  const int kb = (x - 1) / 12, i = (x - 1) - 12 * kb;
  const unsigned long long wv = kb == 0 ? a0 : (kb == 1 ? a1 : a2);
  return (uint32_t)(wv >> (5 * (11 - i))) & 31u;
}

// MODE 2: write in walk order (last block first) at out[0..); MODE 0: count only
template <int MODE>
__device__ __forceinline__ int ir_walk_pipe(const uint32_t *tb, const int32_t *qS, const int32_t *qE, int rows, int tStart, uint32_t *out,
                                            int sl, int sbase, bool have) {
  int t = 0, qs = 0, x = 0, mat = 0, run = 0, lastOp = -1, count = 0;
  long guard = 0, guardMax = 0;
  bool done = !have;
  if (have) { t = rows - 1; qs = qS[t]; x = qE[t] - qs; guardMax = 4L * rows + 4L * (qE[rows - 1] - qS[0]) + 64; }
  auto emit = [&](uint32_t qq, uint32_t tt, uint32_t ln) {
    if (MODE != 0 && sl == 0) { out[3 * count] = qq; out[3 * count + 1] = tt; out[3 * count + 2] = ln; }
    count++;
  };
  const unsigned long long *tb64 = (const unsigned long long *)tb;
  int top = have ? rows - 1 : 0;
  unsigned long long c0 = 0, c1 = 0, c2 = 0, n0 = 0, n1 = 0, n2 = 0;
  int rqs = 0, nqs = 0;
  if (have) {
    int r = top - sl;
    if (r >= 1) { const unsigned long long *w = tb64 + (unsigned long long)r * 3ull; c0 = w[0]; c1 = w[1]; c2 = w[2]; rqs = qS[r - 1]; }
    r -= 8;
    if (r >= 1) { const unsigned long long *w = tb64 + (unsigned long long)r * 3ull; n0 = w[0]; n1 = w[1]; n2 = w[2]; nqs = qS[r - 1]; }
  }
  while (__any_sync(0xffffffffu, !done)) {
    bool step = !done;
    if (step && ++guard > guardMax) { count = -1; done = true; step = false; }
    if (step && t == 0) {
      if (x > 0) {
        if (mat != 0) { count = -1; }
        else {
          if (run > 0) { emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), (uint32_t)run); run = 0; }
          else if (lastOp == IR_DOWN) emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), 0u);
          lastOp = IR_LEFT; x = 0;
        }
      }
      if (count >= 0) { run += 1; emit((uint32_t)qs, (uint32_t)tStart, (uint32_t)run); }
      done = true; step = false;
    }
    if (step && t < top - 7) {     // the walk leaves a chunk one row at a time, so the next chunk is always top - 8
      top -= 8;
      c0 = n0; c1 = n1; c2 = n2; rqs = nqs;
      const int r = top - 8 - sl;
      if (r >= 1) { const unsigned long long *w = tb64 + (unsigned long long)r * 3ull; n0 = w[0]; n1 = w[1]; n2 = w[2]; nqs = qS[r - 1]; }
    }
    // every lane decodes the cell of ITS row at the walker's column; the walker's row owner is picked by shuffle
    const uint32_t mine = ir_pipe_code(c0, c1, c2, x);
    const int src = step ? (top - t) : 0;
    const uint32_t a5 = __shfl_sync(0xffffffffu, mine, sbase + src);
    const int pqs = __shfl_sync(0xffffffffu, rqs, sbase + src);
    if (step) {
      const int a = (int)(a5 & 7u);
      const int dbit = (int)((a5 >> 3) & 1u), ibit = (int)((a5 >> 4) & 1u);
      int op = -2, nt = t, nx = x, nmat = mat;
      if (mat == 0) {
        if (a == IR_DELCLOSE) { op = -1; nmat = 1; }
        else if (a == IR_INSCLOSE) { op = -1; nmat = 2; }
        else if (a == IR_DIAG) { op = IR_DIAG; nt = t - 1; }
        else if (a == IR_LEFT) { op = IR_LEFT; nx = x - 1; }
        else if (a == IR_DOWN) { op = IR_DOWN; nt = t - 1; }
      } else if (mat == 1) { op = IR_DOWN; nmat = dbit ? 1 : 0; nt = t - 1; }
      else { op = IR_LEFT; nmat = ibit ? 2 : 0; nx = x - 1; }
      if (op == -2) { count = -1; done = true; }
      else {
        if (op >= 0) {
          const int q = qs + x;
          if (op == IR_DIAG) run++;
          else {
            if (run > 0) { emit((uint32_t)(q + 1), (uint32_t)(tStart + t + 1), (uint32_t)run); run = 0; }
            else if (lastOp != -1 && lastOp != op) emit((uint32_t)(q + 1), (uint32_t)(tStart + t + 1), 0u);
          }
          lastOp = op;
        }
        if (nt != t) { const int q = qs + x; nx = (op == IR_DIAG ? q - 1 : q) - pqs; qs = pqs; }
        t = nt; x = nx; mat = nmat;
        if (x < 0 || x > 32) { count = -1; done = true; }
      }
    }
  }
  return count;
}

struct IrPipeSmem {          // per group (8 lanes)
  int M[8][33];
  int D[8][33];
  uint32_t row[32][4];       // ring of row descriptors {qS, qE, target b2 word, target mask word}, slot = row & 31
  uint32_t qb2[16];          // ring of packed query words (256 bases), slot = word index & 15
  uint32_t qnm[8];           // ring of query mask words (256 bases), slot = word index & 7
};
constexpr int kIrPipeAhead = 160;   // query bases kept prefetched beyond the row that sub-lane 0 is starting
constexpr int kIrPipeDone = 1 << 20;

template <int CPS>   // cells per lane per lock step
__global__ void __launch_bounds__(128) ir_dp_pipe_kernel(IrBatch b, AogPlan *plan, const uint32_t *sorted, int cls) {
  __shared__ IrPipeSmem sm[4][4];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int sg = lane >> 3, sl = lane & 7, sbase = sg << 3;
  const int prevLane = sbase + ((sl + 7) & 7), nextLane = sbase + ((sl + 1) & 7);
  IrPipeSmem &S = sm[wib][sg];
  const int ps = (sl + 7) & 7;
  const uint32_t begin = plan->bin_start[cls * kAogBuckets];
  const uint32_t end = plan->bin_start[(cls + 1) * kAogBuckets];
  const int match = b.match, mismatch = b.mismatch, gap = b.gap, gapOpen = 2 * b.gap + 1;
  const uint32_t qWordMax = (uint32_t)(b.q.n >> 4) + 3u, qMaskMax = (uint32_t)(b.q.n >> 5) + 3u;
  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&plan->work[cls], 4u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (begin + w >= end) break;
    const bool have = begin + w + (uint32_t)sg < end;
    int g = 0, rows = 0, tStart = 0;
    const int32_t *qS = nullptr, *qE = nullptr;
    uint32_t *tbw = nullptr;
    uint32_t qbase = 0, tbase = 0;     // arenas are < 2^32 bases (uint32 offsets in the ABI)
    if (have) {
      g = (int)sorted[begin + w + sg];
      rows = b.t_len[g]; tStart = b.t_start[g];
      qS = b.band + b.band_off[g]; qE = qS + rows;
      tbw = b.tb + b.tb_off[g];
      qbase = b.q_base[g]; tbase = b.t_base[g] + (uint32_t)tStart;
    }
    unsigned long long *tb64 = (unsigned long long *)tbw;
    __syncwarp();
    auto issue_row = [&](int r) {     // asynchronous fetch of the descriptor of row r
      uint32_t *e = S.row[r & 31];
      cp_async4(e, qS + r);
      cp_async4(e + 1, qE + r);
      const uint32_t tp = tbase + (uint32_t)r;
      cp_async4(e + 2, b.t.b2 + (tp >> 4));
      cp_async4(e + 3, b.t.nm + (tp >> 5));
    };
    uint32_t qwNext = 0, qmNext = 0;   // sub-lane 0: next query word / mask word to request
    auto issue_query = [&](int qs) {   // keep the query ring filled up to qs + kIrPipeAhead
      const uint32_t last = qbase + (uint32_t)qs + (uint32_t)kIrPipeAhead;
      for (; qwNext <= (last >> 4); qwNext++) cp_async4(&S.qb2[qwNext & 15], b.q.b2 + (qwNext < qWordMax ? qwNext : qWordMax));
      for (; qmNext <= (last >> 5); qmNext++) cp_async4(&S.qnm[qmNext & 7], b.q.nm + (qmNext < qMaskMax ? qmNext : qMaskMax));
    };
    int prow = -1, pdone = 0;     // progress of this lane: row it works on / last cell it finished (kIrPipeDone: row complete)
    int t = sl + 1;
    if (have) {
      const int q0 = qS[0], e0 = qE[0];
      for (int xx = sl; xx <= e0 - q0 && xx < 33; xx += 8) { S.M[7][xx] = xx * gap; S.D[7][xx] = kIrBad; }
      if (sl == 7) { prow = 0; pdone = kIrPipeDone; S.row[0][0] = (uint32_t)q0; S.row[0][1] = (uint32_t)e0; }
      if (t < rows) issue_row(t);
      if (t + 8 < rows) issue_row(t + 8);
      if (sl == 0) { qwNext = (qbase + (uint32_t)q0) >> 4; qmNext = (qbase + (uint32_t)q0) >> 5; issue_query(q0); }
      cp_async_wait_all();
    }
    __syncwarp();
    bool started = false;
    int x = 1, rowEnd = 0, off = 0, lenPrev = 0, tc = 5;
    uint32_t qp = 0;                  // arena position of band cell 0 of the current row
    int Mleft = kIrBad, Ileft = kIrBad, Mdiag = kIrBad;
    unsigned long long acc0 = 0, acc1 = 0, acc2 = 0;
    long guard = 0;
    const long guardMax = 64L * (long)__reduce_max_sync(0xffffffffu, rows) + 4096;
    while (__any_sync(0xffffffffu, have && t < rows)) {
      if (++guard > guardMax) { if (lane == 0) atomicOr(b.err, 32); break; }
      const int pr = __shfl_sync(0xffffffffu, prow, prevLane);
      const int pd = __shfl_sync(0xffffffffu, pdone, prevLane);
      // (one cell per step never overwrites a cell the right neighbour still needs: it reads strictly beyond its own column)
      const int nr = CPS > 1 ? __shfl_sync(0xffffffffu, prow, nextLane) : -1;
      const int nd = CPS > 1 ? __shfl_sync(0xffffffffu, pdone, nextLane) : 0;
      if (have && t < rows) {
        if (!started && pr == t - 1) {
          cp_async_wait_all();              // my own copies: rows t and t+8 (issued >= 8 rows ago), query words (sub-lane 0)
          const uint32_t *e = S.row[t & 31];
          const uint32_t *pe = S.row[(t - 1) & 31];
          const int qs = (int)e[0], qe = (int)e[1];
          const uint32_t tw = e[2], tm = e[3];
          const int pqs = (int)pe[0], pqe = (int)pe[1];
          const int len = qe - qs + 1;
          off = qs - pqs;
          lenPrev = pqe - pqs + 1;
          rowEnd = (t == rows - 1) ? len : len - 1;
          const uint32_t tp = tbase + (uint32_t)t;
          tc = ((tm >> (tp & 31)) & 1u) ? 4 : (int)((tw >> ((tp & 15) * 2)) & 3u);
          qp = qbase + (uint32_t)qs;
          x = 1; Mleft = kIrBad; Ileft = kIrBad;
          Mdiag = (off < lenPrev && off >= 0) ? kIrNeg : kIrBad;   // marker: load on the first cell
          acc0 = acc1 = acc2 = 0;
          started = true;
          prow = t; pdone = 0;
          if (t + 16 < rows) issue_row(t + 16);
          if (sl == 0) issue_query(qs);
        }
        if (started) {
          // cells of the previous row that are final / cells of my old row the right neighbour no longer reads
          const int avail = (pr > t - 1) ? kIrPipeDone : pd;
          const int safe = (CPS > 1 && nr == t - 7) ? nd : kIrPipeDone;
#pragma unroll
          for (int c = 0; c < CPS; c++) {
            const int xp = x + off;
            if (x < rowEnd && imin(xp, lenPrev - 2) <= avail && x <= safe) {
              const bool upIn = xp <= lenPrev - 1;
              const bool upOk = xp < lenPrev - 1;
              const bool diagOk = upIn && !(xp - 1 == 0 && t != 1);
              if (Mdiag == kIrNeg) Mdiag = S.M[ps][off];
              int Mup = kIrBad, Dup = kIrBad;
              if (upIn) { Mup = S.M[ps][xp]; Dup = S.D[ps][xp]; }
              const uint32_t p = qp + (uint32_t)x;
              const int qc = ((S.qnm[(p >> 5) & 7] >> (p & 31)) & 1u) ? 4 : (int)((S.qb2[(p >> 4) & 15] >> ((p & 15) * 2)) & 3u);
              const int delOpen = upOk ? Mup + gapOpen : kIrBad;
              const int delExt = upOk ? Dup : kIrBad;
              const int D = imax(delOpen, delExt);
              const uint32_t dbit = (D == delOpen) ? 0u : 8u;
              const int insOpen = Mleft + gapOpen;
              const int I = imax(insOpen, Ileft);
              const uint32_t ibit = (I == insOpen) ? 0u : 16u;
              const int mS = diagOk ? Mdiag + (qc == tc ? match : mismatch) : kIrBad;
              const int iS = Mleft + gap;
              const int dS = upOk ? Mup + gap : kIrBad;
              const int mx = imax(imax(mS, iS), imax(dS, imax(D, I)));
              const uint32_t a = (mx == mS) ? IR_DIAG : (mx == iS) ? IR_LEFT : (mx == dS) ? IR_DOWN : (mx == D) ? IR_DELCLOSE : IR_INSCLOSE;
              if (x == 13 || x == 25) { acc2 = acc1; acc1 = acc0; acc0 = 0; }
              acc0 = (acc0 << 5) | (unsigned long long)(a | dbit | ibit);
              S.M[sl][x] = mx; S.D[sl][x] = D;
              Mdiag = Mup; Mleft = mx; Ileft = I;
              pdone = x;
              x++;
            }
          }
          if (x >= rowEnd) {
            // row complete (possibly without any computed cell): publish its arrows, take the next row of this lane
            const int n = rowEnd - 1;
            unsigned long long v0 = 0, v1 = 0, v2 = 0;     // cells never computed read as 0, like untouched bit planes
            if (n >= 1) {
              const int kb = (n - 1) / 12, m = n - 12 * kb;
              const unsigned long long last = acc0 << (5 * (12 - m));
              if (kb == 0) v0 = last;
              else if (kb == 1) { v0 = acc1; v1 = last; }
              else { v0 = acc2; v1 = acc1; v2 = last; }
            }
            unsigned long long *w3 = tb64 + (unsigned)t * 3u;
            w3[0] = v0; w3[1] = v1; w3[2] = v2;
            pdone = kIrPipeDone;
            t += 8; started = false;
          }
        }
      }
      __syncwarp();
    }
    cp_async_wait_all();
    __syncwarp();
    // traceback, all four groups of the warp in lock step, each walked by its own 8 lanes
    uint32_t *rb = have ? tbw + (unsigned long long)rows * 6ull + 8ull : nullptr;
    int nb = ir_walk_pipe<2>(tbw, qS, qE, rows, tStart, rb, sl, sbase, have);
    if (have && nb < 0) { if (sl == 0) atomicOr(b.err, 16); nb = 0; }
    if (!have) nb = 0;
    unsigned long long slot = aog_reserve_blocks(AogBatch{b.q, b.t, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, nullptr,
                                                          nullptr, b.blocks, b.block_cap, b.block_cursor, b.err},
                                                 (have && sl == 0) ? nb : 0, lane, &plan->cls_blocks[cls]);
    slot = __shfl_sync(0xffffffffu, slot, sbase);
    if (have && sl == 0) { b.n_blocks[g] = nb; b.block_off[g] = slot; }
    __syncwarp();
    if (have && slot != ~0ull) {
      uint32_t *out = b.blocks + 3ull * slot;
      for (int i = sl; i < 3 * nb; i += 8) { const int r = i / 3, c = i - 3 * r; out[i] = rb[3 * (nb - 1 - r) + c]; }
    }
    __syncwarp();
  }
}

// ---- generic width: one group per thread, rows and byte arrows in the group's scratch region (exact, slow, rare)
__device__ __forceinline__ int ir_walk_bytes(const uint8_t *ar, int W, const int32_t *qS, const int32_t *qE, int rows, int tStart,
                                             uint32_t *out, int nb, bool write) {
  int t = rows - 1, qs = qS[t], x = qE[t] - qs, mat = 0, run = 0, lastOp = -1, count = 0;
  long guard = 0;
  const long guardMax = 4L * rows + 4L * (qE[rows - 1] - qS[0]) + 64;
  auto emit = [&](uint32_t qq, uint32_t tt, uint32_t ln) {
    if (write) { const int r = nb - 1 - count; out[3 * r] = qq; out[3 * r + 1] = tt; out[3 * r + 2] = ln; }
    count++;
  };
  while (true) {
    if (++guard > guardMax) return -1;
    if (t == 0) {
      if (x > 0) {
        if (mat != 0) return -1;
        if (run > 0) { emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), (uint32_t)run); run = 0; }
        else if (lastOp == IR_DOWN) emit((uint32_t)(qs + x + 1), (uint32_t)(tStart + 1), 0u);
        lastOp = IR_LEFT; x = 0;
      }
      run += 1;
      emit((uint32_t)qs, (uint32_t)tStart, (uint32_t)run);
      break;
    }
    const int a5 = ar[(unsigned long long)t * W + x];
    int op, nt = t, nx = x, nmat = mat;
    if (mat == 0) {
      const int a = a5 & 7;
      if (a == IR_DELCLOSE) { op = -1; nmat = 1; }
      else if (a == IR_INSCLOSE) { op = -1; nmat = 2; }
      else if (a == IR_DIAG) { op = IR_DIAG; nt = t - 1; }
      else if (a == IR_LEFT) { op = IR_LEFT; nx = x - 1; }
      else if (a == IR_DOWN) { op = IR_DOWN; nt = t - 1; }
      else return -1;
    } else if (mat == 1) { op = IR_DOWN; nmat = ((a5 >> 3) & 1) ? 1 : 0; nt = t - 1; }
    else { op = IR_LEFT; nmat = ((a5 >> 4) & 1) ? 2 : 0; nx = x - 1; }
    if (op >= 0) {
      const int q = qs + x;
      if (op == IR_DIAG) run++;
      else {
        if (run > 0) { emit((uint32_t)(q + 1), (uint32_t)(tStart + t + 1), (uint32_t)run); run = 0; }
        else if (lastOp != -1 && lastOp != op) emit((uint32_t)(q + 1), (uint32_t)(tStart + t + 1), 0u);
      }
      lastOp = op;
    }
    if (nt != t) { const int pqs = qS[t - 1]; const int q = qs + x; nx = (op == IR_DIAG ? q - 1 : q) - pqs; qs = pqs; }
    t = nt; x = nx; mat = nmat;
    if (x < 0) return -1;
  }
  return count;
}

__global__ void __launch_bounds__(64) ir_dp_generic_kernel(IrBatch b, AogPlan *plan, const uint32_t *sorted) {
  const int cls = kIrClsGeneric;
  const int lane = threadIdx.x & 31;
  const uint32_t begin = plan->bin_start[cls * kAogBuckets];
  const uint32_t end = plan->bin_start[(cls + 1) * kAogBuckets];
  const int match = b.match, mismatch = b.mismatch, gap = b.gap, gapOpen = 2 * b.gap + 1;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&plan->work[cls], 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (begin + base >= end) break;
    const uint32_t idx = begin + base + lane;
    const bool active = idx < end;
    int g = 0, nb = 0, rows = 0, tStart = 0, W = 0;
    const int32_t *qS = nullptr, *qE = nullptr;
    uint8_t *ar = nullptr;
    if (active) {
      g = (int)sorted[idx];
      rows = b.t_len[g]; tStart = b.t_start[g]; W = b.max_width[g];
      qS = b.band + b.band_off[g]; qE = qS + rows;
      uint32_t *region = b.tb + b.tb_off[g];
      int *M = (int *)region;                    // [W]
      int *D = M + W;                            // [W]
      ar = (uint8_t *)(D + W);                   // [rows][W]
      int qsPrev = qS[0], lenPrev = qE[0] - qS[0] + 1;
      for (int x = 0; x < lenPrev; x++) { M[x] = x * gap; D[x] = kIrBad; }
      for (int t = 1; t < rows; t++) {
        const int qs = qS[t], len = qE[t] - qs + 1, off = qs - qsPrev;
        const int rowEnd = (t == rows - 1) ? len : len - 1;
        const int tc = seq_code(b.t, (uint64_t)(uint32_t)(b.t_base[g] + (uint32_t)(tStart + t)));
        int Mdiag = off <= lenPrev - 1 ? M[off] : kIrBad;
        int Mleft = kIrBad, Ileft = kIrBad;
        M[0] = kIrBad;
        for (int x = 1; x < rowEnd; x++) {
          const int xp = x + off;
          const bool upIn = xp <= lenPrev - 1, upOk = xp < lenPrev - 1;
          const bool diagOk = upIn && !(xp - 1 == 0 && t != 1);
          int Mup = kIrBad, Dup = kIrBad;
          if (upIn) { Mup = M[xp]; Dup = D[xp]; }
          const int delOpen = upOk ? Mup + gapOpen : kIrBad, delExt = upOk ? Dup : kIrBad;
          const int Dv = imax(delOpen, delExt);
          const int dbit = (Dv == delOpen) ? 0 : 1;
          const int insOpen = Mleft + gapOpen;
          const int Iv = imax(insOpen, Ileft);
          const int ibit = (Iv == insOpen) ? 0 : 1;
          const int qc = seq_code(b.q, (uint64_t)b.q_base[g] + (uint64_t)(qs + x));
          const int mS = diagOk ? Mdiag + (qc == tc ? match : mismatch) : kIrBad;
          const int iS = Mleft + gap;
          const int dS = upOk ? Mup + gap : kIrBad;
          const int mx = imax(imax(mS, iS), imax(dS, imax(Dv, Iv)));
          const int a = (mx == mS) ? IR_DIAG : (mx == iS) ? IR_LEFT : (mx == dS) ? IR_DOWN : (mx == Dv) ? IR_DELCLOSE : IR_INSCLOSE;
          ar[(unsigned long long)t * W + x] = (uint8_t)(a | (dbit << 3) | (ibit << 4));
          M[x] = mx; D[x] = Dv;
          Mdiag = Mup; Mleft = mx; Ileft = Iv;
        }
        qsPrev = qs; lenPrev = len;
      }
      nb = ir_walk_bytes(ar, W, qS, qE, rows, tStart, nullptr, 0, false);
      if (nb < 0) { atomicOr(b.err, 16); nb = 0; }
    }
    const unsigned long long slot = aog_reserve_blocks(AogBatch{b.q, b.t, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, nullptr,
                                                                nullptr, b.blocks, b.block_cap, b.block_cursor, b.err},
                                                       nb, lane, &plan->cls_blocks[cls]);
    if (active) {
      b.n_blocks[g] = nb;
      b.block_off[g] = slot;
      if (slot != ~0ull && nb > 0) ir_walk_bytes(ar, W, qS, qE, rows, tStart, b.blocks + 3ull * slot, nb, true);
    }
  }
}

}  // namespace lra
