// a19  IndelRefineAlignment -- the segment-level part: end padding, grouping of blocks, band construction and the final
// assembly of the refined block list.  Reference: IndelRefine.h:53-784 (end padding :89-130, grouping :132-211, band
// :220-333, small-window fallback to AffineOneGapAlign :344-357, stitching :761-771).
//
//   ir_group_kernel     one thread per segment: replays the (sequential, cheap) grouping loop on a scratch copy of the
//                       segment's blocks and emits an ordered list of PIECES (literal block | AffineOneGapAlign job |
//                       banded DP group) plus the job / group descriptors.
//   ir_band_kernel      one warp per DP group: the reference's band construction, steps in order, the inner +-k
//                       neighbourhood update spread over the lanes; then the monotone fix-ups as warp scans.
//   ir_assemble_kernel  one thread per segment: concatenates the pieces' blocks in order into the output arena.
#pragma once
#include "ir_kernels.cuh"

namespace lra {

enum IrPiece : uint32_t { IR_PIECE_LITERAL = 0, IR_PIECE_AOG = 1, IR_PIECE_DP = 2 };

struct IrSegBatch {
  // input (device)
  const uint32_t *blocks_in;        // [total_in][3]
  const unsigned long long *blk_off;  // [n_seg] first triple of the segment (exclusive prefix of blk_cnt)
  const int32_t *blk_cnt;           // [n_seg]
  const uint32_t *q_base, *t_base;  // [n_seg] arena offsets of qSeq[0] / tSeq[0]
  const int32_t *read_len, *contig_len;
  int n_seg;
  int k, end_align;
  // scratch (device)
  uint32_t *work;                   // [total_in + 2 n_seg][3] mutable copy of the blocks (with end padding)
  uint32_t *pieces;                 // [2 total_in + 8 n_seg][4]
  int32_t *n_pieces;                // [n_seg]
  // AffineOneGapAlign fallback jobs (SoA, capacity total_in + n_seg)
  uint32_t *aog_q_off, *aog_t_off;
  int32_t *aog_q_len, *aog_t_len, *aog_k;
  // DP groups (SoA, capacity total_in + n_seg)
  uint32_t *g_q_base, *g_t_base;
  int32_t *g_q_start, *g_t_start, *g_t_len, *g_q_seq_len, *g_t_seq_len;
  uint32_t *g_band_off;
  int32_t *g_seg, *g_first_block, *g_last_block;   // indices into the segment's work copy
  uint32_t *g_first;                // [n][3] the (trimmed) first block as the band builder must see it
  uint32_t *g_last;                 // [n][3] the last block (q, t, trimmed length) as seen by the band builder
  unsigned long long *counters;     // [0] n_aog, [1] n_groups, [2] band ints, [3] status flags
};

__device__ __forceinline__ unsigned long long ir_work_off(const IrSegBatch &b, int s) { return b.blk_off[s] + 2ull * (unsigned long long)s; }
__device__ __forceinline__ unsigned long long ir_piece_off(const IrSegBatch &b, int s) { return 2ull * b.blk_off[s] + 8ull * (unsigned long long)s; }

// One WARP per segment.  The grouping loop is sequential, but it only ever modifies the block that currently starts a
// group (trimmed head / restored tail), so the loop runs on the read-only padded copy plus one override block held in
// registers; all lanes execute it uniformly, block fields come from a 32-block register chunk by shuffle, lane 0 writes.
__global__ void __launch_bounds__(128) ir_group_kernel(IrSegBatch b) {
  const int lane = threadIdx.x & 31;
  const int s = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (s >= b.n_seg) return;
  const int n_in = b.blk_cnt[s];
  const uint32_t *in = b.blocks_in + 3ull * b.blk_off[s];
  uint32_t *W = b.work + 3ull * ir_work_off(b, s);
  uint32_t *P = b.pieces + 4ull * ir_piece_off(b, s);
  int np = 0;
  auto piece = [&](uint32_t kind, uint32_t a, uint32_t c, uint32_t d) {
    if (lane == 0) { P[4 * np] = kind; P[4 * np + 1] = a; P[4 * np + 2] = c; P[4 * np + 3] = d; }
    np++;
  };
  if (n_in <= 1) {
    for (int i = 0; i < n_in; i++) piece(IR_PIECE_LITERAL, in[3 * i], in[3 * i + 1], in[3 * i + 2]);
    if (lane == 0) b.n_pieces[s] = np;
    return;
  }
  const int k = b.k, maxGap = b.k - 1;
  int nb = 0;
  {
    long qS0 = in[0], tS0 = in[1];
    const long qAlnEnd = (long)in[3 * (n_in - 1)] + in[3 * (n_in - 1) + 2];
    const long tAlnEnd = (long)in[3 * (n_in - 1) + 1] + in[3 * (n_in - 1) + 2];
    int addStart = 0, addEnd = 0, startMatch = 0, endMatch = 0;
    if (b.end_align) {
      const int minStart = (int)(qS0 < tS0 ? qS0 : tS0);
      if (minStart < 40) { tS0 -= minStart; qS0 -= minStart; startMatch = minStart; addStart = 1; }
      const long a = (long)b.read_len[s] - qAlnEnd, c = (long)b.contig_len[s] - tAlnEnd;
      const int minEnd = (int)(a < c ? a : c);
      if (minEnd < 40) { endMatch = minEnd; addEnd = 1; }
    }
    if (addStart && lane == 0) { W[0] = (uint32_t)qS0; W[1] = (uint32_t)tS0; W[2] = (uint32_t)startMatch; }
    for (int i = lane; i < 3 * n_in; i += 32) W[3 * addStart + i] = in[i];
    nb = n_in + addStart;
    if (addEnd) { if (lane == 0) { W[3 * nb] = (uint32_t)qAlnEnd; W[3 * nb + 1] = (uint32_t)tAlnEnd; W[3 * nb + 2] = (uint32_t)endMatch; } nb++; }
  }
  __syncwarp();
  // register chunk of 32 blocks + the override block
  int cbase = -1000;
  uint32_t cq = 0, ct = 0, cl = 0;
  int ovIdx = -1;
  uint32_t ovQ = 0, ovT = 0, ovL = 0;
  auto fetch = [&](int i, uint32_t &q, uint32_t &t, uint32_t &l) {
    if (i < cbase || i >= cbase + 32) {
      cbase = i;
      const int j = cbase + lane;
      if (j < nb) { cq = W[3 * j]; ct = W[3 * j + 1]; cl = W[3 * j + 2]; }
    }
    q = __shfl_sync(0xffffffffu, cq, i - cbase); t = __shfl_sync(0xffffffffu, ct, i - cbase); l = __shfl_sync(0xffffffffu, cl, i - cbase);
    if (i == ovIdx) { q = ovQ; t = ovT; l = ovL; }
  };
  int startBlock = 0, endBlock = 0;
  while (endBlock < nb) {
    uint32_t sQ, sT, sL;
    fetch(startBlock, sQ, sT, sL);
    long qStart = sQ, tStart = sT;
    long qPos = (long)sQ + (int)sL, tPos = (long)sT + (int)sL;
    int tGap = 0, qGap = 0;
    uint32_t eQ = sQ, eT = sT, eL = sL;      // block endBlock
    uint32_t nQ = 0, nT = 0, nL = 0;         // block endBlock + 1
    if (endBlock < nb - 1) { fetch(endBlock + 1, nQ, nT, nL); tGap = (int)(nT - (uint32_t)tPos); qGap = (int)(nQ - (uint32_t)qPos); }
    while (endBlock < nb - 1 && qGap < maxGap && tGap < maxGap && (startBlock == endBlock || eL < 100u)) {
      endBlock++;
      eQ = nQ; eT = nT; eL = nL;
      qPos = (long)eQ + (int)eL; tPos = (long)eT + (int)eL;
      if (endBlock + 1 < nb - 1) { fetch(endBlock + 1, nQ, nT, nL); tGap = (int)(nT - (uint32_t)tPos); qGap = (int)(nQ - (uint32_t)qPos); }
      else if (endBlock + 1 < nb) fetch(endBlock + 1, nQ, nT, nL);   // keep (n*) = block endBlock+1 for the next iteration; gaps stay stale as in the reference
    }
    uint32_t altQ = 0, altT = 0, altL = 0;
    bool usedAlt = false;
    if (endBlock == startBlock) {
      piece(IR_PIECE_LITERAL, sQ, sT, sL);
    } else {
      if ((long)sL > maxGap) {
        const int advanced = (int)sL - maxGap;
        piece(IR_PIECE_LITERAL, sQ, sT, sL - (uint32_t)maxGap);
        sQ += (uint32_t)advanced; sT += (uint32_t)advanced; sL = (uint32_t)maxGap;
        qStart += advanced; tStart += advanced;
      }
      if ((long)eL > maxGap) {
        usedAlt = true;
        altQ = eQ + (uint32_t)maxGap; altT = eT + (uint32_t)maxGap; altL = eL - (uint32_t)maxGap;
        eL = (uint32_t)maxGap;
        qPos = (long)eQ + maxGap; tPos = (long)eT + maxGap;
      }
      const long qEnd = (long)eQ + eL, tEnd = (long)eT + eL;
      const long tLen = tPos - tStart;
      const long tSeqLen = tEnd - tStart, qSeqLen = qEnd - qStart;
      if (tSeqLen < k || qSeqLen < k) {
        unsigned long long j = 0;
        if (lane == 0) {
          j = atomicAdd(&b.counters[0], 1ull);
          b.aog_q_off[j] = b.q_base[s] + (uint32_t)qStart;
          b.aog_t_off[j] = b.t_base[s] + (uint32_t)tStart;
          b.aog_q_len[j] = (int32_t)qSeqLen; b.aog_t_len[j] = (int32_t)tSeqLen; b.aog_k[j] = k;
        }
        j = __shfl_sync(0xffffffffu, j, 0);
        piece(IR_PIECE_AOG, (uint32_t)j, (uint32_t)qStart, (uint32_t)tStart);
      } else {
        unsigned long long g = 0;
        if (lane == 0) {
          g = atomicAdd(&b.counters[1], 1ull);
          b.g_q_base[g] = b.q_base[s]; b.g_t_base[g] = b.t_base[s];
          b.g_q_start[g] = (int32_t)qStart; b.g_t_start[g] = (int32_t)tStart; b.g_t_len[g] = (int32_t)tLen;
          b.g_q_seq_len[g] = (int32_t)qSeqLen; b.g_t_seq_len[g] = (int32_t)tSeqLen;
          b.g_band_off[g] = (uint32_t)atomicAdd(&b.counters[2], (unsigned long long)(2 * tLen));
          b.g_seg[g] = s; b.g_first_block[g] = startBlock; b.g_last_block[g] = endBlock;
          b.g_first[3 * g] = sQ; b.g_first[3 * g + 1] = sT; b.g_first[3 * g + 2] = sL;
          b.g_last[3 * g] = eQ; b.g_last[3 * g + 1] = eT; b.g_last[3 * g + 2] = eL;
        }
        g = __shfl_sync(0xffffffffu, g, 0);
        piece(IR_PIECE_DP, (uint32_t)g, 0u, 0u);
      }
    }
    if (!usedAlt) { endBlock++; ovIdx = -1; }
    else { ovIdx = endBlock; ovQ = altQ; ovT = altT; ovL = altL; }
    startBlock = endBlock;
  }
  if (lane == 0) b.n_pieces[s] = np;
}

// One warp per DP group.  The work copy of the segment's blocks may have moved on (later groups trim / restore their own
// boundary blocks), so the first block and the last block's length come from the group record; inner blocks are untouched.
__global__ void __launch_bounds__(128) ir_band_literal_kernel(IrSegBatch b, int n_groups, int32_t *band) {
  const int lane = threadIdx.x & 31;
  const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (g >= n_groups) return;
  const int s = b.g_seg[g];
  const uint32_t *W = b.work + 3ull * ir_work_off(b, s);
  const int startBlock = b.g_first_block[g], endBlock = b.g_last_block[g];
  const int k = b.k;
  const long tLen = b.g_t_len[g];
  const long qStart = b.g_q_start[g];
  const long qEnd = qStart + b.g_q_seq_len[g];
  int32_t *qS = band + b.g_band_off[g];
  int32_t *qE = qS + tLen;
  auto blkq = [&](int i) -> long { return i == startBlock ? (long)b.g_first[3 * g] : (i == endBlock ? (long)b.g_last[3 * g] : (long)W[3 * i]); };
  auto blkt = [&](int i) -> long { return i == startBlock ? (long)b.g_first[3 * g + 1] : (i == endBlock ? (long)b.g_last[3 * g + 1] : (long)W[3 * i + 1]); };
  auto blkl = [&](int i) -> long { return i == startBlock ? (long)b.g_first[3 * g + 2] : (i == endBlock ? (long)b.g_last[3 * g + 2] : (long)W[3 * i + 2]); };
  // Steps run in path order.  Within one step the lanes touch distinct rows (lane ki: rows tOff-ki and tOff+ki), and the
  // row's own update (IndelRefine.h:253-266) is done by lane 0 right before its ki = 0 neighbourhood update, so one warp
  // barrier per step orders everything.  Only the rows [tOff-k+1, tOff+k-1] (+ a target gap) are live at any time: they are
  // kept in a shared-memory ring (row & 255) and flushed to HBM once the path has moved k rows past them.
  __shared__ int ringS[4][256];
  __shared__ int ringE[4][256];
  const int tl = (int)tLen;
  const int qStartI = (int)qStart, qEndI = (int)qEnd;
  if (2 * k + (k - 1) <= 250) {
    int *rS = ringS[threadIdx.x >> 5], *rE = ringE[threadIdx.x >> 5];
    for (int x = lane; x < 256; x += 32) { rS[x] = -1; rE[x] = -1; }
    __syncwarp();
    int q = (int)blkq(startBlock);
    int tOff = 0;
    int nextFlush = 0;     // rows < nextFlush are final and in HBM
    auto flush_to = [&](int upto) {   // make rows < upto final (upto <= first row that can still change)
      if (upto > tl) upto = tl;
      if (upto > nextFlush) {
        for (int r = nextFlush + lane; r < upto; r += 32) { qS[r] = rS[r & 255]; qE[r] = rE[r & 255]; rS[r & 255] = -1; rE[r & 255] = -1; }
        nextFlush = upto;
        __syncwarp();
      }
    };
    for (int bb = startBlock; bb <= endBlock; bb++) {
      int qGap = 0, tGap = 0;
      int blockLength = (int)blkl(bb);
      if (bb < endBlock) {
        qGap = (int)(blkq(bb + 1) - (blkq(bb) + blockLength));
        tGap = (int)(blkt(bb + 1) - (blkt(bb) + blockLength));
        if (qGap > 0 && tGap > 0) { const int c = qGap < tGap ? qGap : tGap; qGap -= c; tGap -= c; blockLength += c; }
      }
      for (int bi = 0; bi < blockLength; tOff++, bi++, q++) {
        flush_to(tOff - k + 1);
        for (int ki = lane; ki < k; ki += 32) {
          const int im = tOff - ki, ip = tOff + ki;
          if (ki == 0 && tOff < tl) {
            const int lo = imax(q - k, qStartI);
            const int cs = rS[tOff & 255];
            if (cs == -1 || lo < cs) rS[tOff & 255] = lo;
            const int ce = rE[tOff & 255];
            if (ce == -1 || ce < q + k) rE[tOff & 255] = imin(qEndI - 1, q + k);
          }
          if (im >= 0 && im < tl) { if (rE[im & 255] < q) rE[im & 255] = q; }
          if (ip < tl) { const int v = rS[ip & 255]; if (v == -1 || v > q) rS[ip & 255] = q; }
        }
        __syncwarp();
      }
      if (qGap > tGap) {
        flush_to(tOff - k + 1);
        for (int qi = 0; qi < qGap; qi++, q++) {
          for (int ki = lane; ki < k; ki += 32) {
            const int im = tOff - ki, ip = tOff + ki;
            if (im >= 0 && im < tl) { if (rE[im & 255] < q) rE[im & 255] = q; }
            if (ip < tl) { const int v = rS[ip & 255]; if (v == 0 || v > q) rS[ip & 255] = q; }
          }
          __syncwarp();
        }
      }
      if (tGap > qGap) {
        flush_to(tOff - k + 1);
        const int lo = imax(q - k, qStartI);
        const int hi = imin(qEndI - 1, q + k);
        for (int ti = lane; ti < tGap; ti += 32) if (tOff + ti < tl) { rS[(tOff + ti) & 255] = lo; rE[(tOff + ti) & 255] = hi; }
        tOff += tGap;
        __syncwarp();
      }
    }
    flush_to(tl);
  } else {
  for (long x = lane; x < tLen; x += 32) { qS[x] = -1; qE[x] = -1; }
  __syncwarp();
  int q = (int)blkq(startBlock);
  int tOff = 0;
  for (int bb = startBlock; bb <= endBlock; bb++) {
    int qGap = 0, tGap = 0;
    int blockLength = (int)blkl(bb);
    if (bb < endBlock) {
      qGap = (int)(blkq(bb + 1) - (blkq(bb) + blockLength));
      tGap = (int)(blkt(bb + 1) - (blkt(bb) + blockLength));
      if (qGap > 0 && tGap > 0) { const int c = qGap < tGap ? qGap : tGap; qGap -= c; tGap -= c; blockLength += c; }
    }
    for (int bi = 0; bi < blockLength; tOff++, bi++, q++) {
      for (int ki = lane; ki < k; ki += 32) {
        const int im = tOff - ki, ip = tOff + ki;
        if (ki == 0 && tOff < tl) {
          const int lo = imax(q - k, qStartI);
          const int cs = qS[tOff];
          if (cs == -1 || lo < cs) qS[tOff] = lo;
          const int ce = qE[tOff];
          if (ce == -1 || ce < q + k) qE[tOff] = imin(qEndI - 1, q + k);
        }
        if (im >= 0 && im < tl) { if (qE[im] < q) qE[im] = q; }
        if (ip < tl) { const int v = qS[ip]; if (v == -1 || v > q) qS[ip] = q; }
      }
      __syncwarp();
    }
    if (qGap > tGap) {
      for (int qi = 0; qi < qGap; qi++, q++) {
        for (int ki = lane; ki < k; ki += 32) {
          const int im = tOff - ki, ip = tOff + ki;
          if (im >= 0 && im < tl) { if (qE[im] < q) qE[im] = q; }
          if (ip < tl) { const int v = qS[ip]; if (v == 0 || v > q) qS[ip] = q; }
        }
        __syncwarp();
      }
    }
    if (tGap > qGap) {
      const int lo = imax(q - k, qStartI);
      const int hi = imin(qEndI - 1, q + k);
      for (int ti = lane; ti < tGap; ti += 32) if (tOff + ti < tl) { qS[tOff + ti] = lo; qE[tOff + ti] = hi; }
      tOff += tGap;
      __syncwarp();
    }
  }
  }
  __syncwarp();
  // monotone fix-ups (IndelRefine.h:318-325): qS := suffix minimum, qE := prefix maximum
  {
    int carry = 0x7fffffff;
    for (long base = ((tLen - 1) / 32) * 32; base >= 0; base -= 32) {
      const long i = base + lane;
      int v = (i < tLen) ? qS[i] : 0x7fffffff;
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v = imin(v, u); }
      v = imin(v, carry);
      if (i < tLen) qS[i] = v;
      carry = __shfl_sync(0xffffffffu, v, 0);
    }
    carry = -0x7fffffff;
    for (long base = 0; base < tLen; base += 32) {
      const long i = base + lane;
      int v = (i < tLen) ? qE[i] : -0x7fffffff;
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = imax(v, u); }
      v = imax(v, carry);
      if (i < tLen) qE[i] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  }
}


// Closed form of the band construction (default).  The reference walks the old path step by step (IndelRefine.h:232-315);
// because q and the row index never decrease along the path, every update it makes is a running min / max whose winner is
// known in O(1) per row from two per-row arrays:
//   Q[r]  = query position of the path at row r   (diagonal rows: the step's q; target-gap rows: the q the gap sits at)
//   F[r]  = bit0: row r lies in a target gap; bit1: a query gap precedes the diagonal step of row r
// With R2 = min(r+k-1, tLen-1) and a = max(0, r-k+1):
//   qE[r] = max( min(qEnd-1, Q[r]+k),  Q[R2] - (F[R2]&1) )                      (last step at or before row r+k-1)
//   qS[r] = max(Q[r]-k, qStart)                                  on target-gap rows (the gap overwrites earlier touches)
//   qS[r] = min( v, max(Q[r]-k, qStart) ),  v = Q[a]             on diagonal rows r >= 1 (first step at or after row a);
//           if v == 0 and a query gap precedes a row R in (a, r], v becomes that gap's first q  (the `== 0` test of :298)
//           row 0: v does not exist, qS[0] = max(Q[0]-k, qStart)
// followed by the same monotone fix-ups.  Validated against the step-by-step kernel (ir_band_literal_kernel) and the oracle.
__global__ void __launch_bounds__(128) ir_band_kernel(IrSegBatch b, int n_groups, int32_t *band, int32_t *tmp) {
  const int lane = threadIdx.x & 31;
  const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (g >= n_groups) return;
  const int s = b.g_seg[g];
  const uint32_t *W = b.work + 3ull * ir_work_off(b, s);
  const int startBlock = b.g_first_block[g], endBlock = b.g_last_block[g];
  const int k = b.k;
  const int tl = b.g_t_len[g];
  const int qStart = b.g_q_start[g];
  const int qEnd = qStart + b.g_q_seq_len[g];
  const int tStart = b.g_t_start[g];
  int32_t *qS = band + b.g_band_off[g];
  int32_t *qE = qS + tl;
  int32_t *Q = tmp + b.g_band_off[g];
  int32_t *F = Q + tl;
  auto blkq = [&](int i) -> int { return i == startBlock ? (int)b.g_first[3 * g] : (i == endBlock ? (int)b.g_last[3 * g] : (int)W[3 * i]); };
  auto blkt = [&](int i) -> int { return i == startBlock ? (int)b.g_first[3 * g + 1] : (i == endBlock ? (int)b.g_last[3 * g + 1] : (int)W[3 * i + 1]); };
  auto blkl = [&](int i) -> int { return i == startBlock ? (int)b.g_first[3 * g + 2] : (i == endBlock ? (int)b.g_last[3 * g + 2] : (int)W[3 * i + 2]); };
  // phase 1: per-row path description, one block per lane
  for (int bb = startBlock + lane; bb <= endBlock; bb += 32) {
    int len = blkl(bb), qGap = 0, tGap = 0;
    const int q0 = blkq(bb), r0 = blkt(bb) - tStart;
    if (bb < endBlock) {
      qGap = blkq(bb + 1) - (q0 + len);
      tGap = blkt(bb + 1) - (blkt(bb) + len);
      if (qGap > 0 && tGap > 0) { const int c = qGap < tGap ? qGap : tGap; qGap -= c; tGap -= c; len += c; }
    }
    for (int i = 0; i < len; i++) if (r0 + i < tl) { Q[r0 + i] = q0 + i; F[r0 + i] = (i == 0 && bb > startBlock) ? (F[r0] & 2) : 0; }
    if (tGap > qGap) for (int i = 0; i < tGap; i++) if (r0 + len + i < tl) { Q[r0 + len + i] = q0 + len; F[r0 + len + i] = 1; }
  }
  __syncwarp();
  // query-gap markers (second pass so that they are not overwritten by the row fill above)
  for (int bb = startBlock + lane; bb < endBlock; bb += 32) {
    int len = blkl(bb);
    int qGap = blkq(bb + 1) - (blkq(bb) + len), tGap = blkt(bb + 1) - (blkt(bb) + len);
    if (qGap > 0 && tGap > 0) { const int c = qGap < tGap ? qGap : tGap; qGap -= c; tGap -= c; }
    const int rn = blkt(bb + 1) - tStart;
    if (qGap > tGap && rn < tl) F[rn] |= 2;
  }
  __syncwarp();
  // phase 2: every row independently
  for (int r = lane; r < tl; r += 32) {
    const int Qr = Q[r], Fr = F[r];
    const int R2 = imin(r + k - 1, tl - 1);
    const int lastQ = Q[R2] - (F[R2] & 1);
    const int e = imax(imin(qEnd - 1, Qr + k), lastQ);
    const int lo = imax(Qr - k, qStart);
    int v;
    if (Fr & 1) v = lo;
    else if (r == 0) v = lo;
    else {
      const int a = imax(0, r - k + 1);
      v = Q[a];
      if (v == 0) {
        for (int R = a + 1; R <= r; R++) if (F[R] & 2) { v = Q[R - 1] + 1; break; }
      }
      v = imin(v, lo);
    }
    qS[r] = v;
    qE[r] = e;
  }
  __syncwarp();
  // monotone fix-ups (IndelRefine.h:318-325): qS := suffix minimum, qE := prefix maximum
  {
    int carry = 0x7fffffff;
    for (int base = ((tl - 1) / 32) * 32; base >= 0; base -= 32) {
      const int i = base + lane;
      int v = (i < tl) ? qS[i] : 0x7fffffff;
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v = imin(v, u); }
      v = imin(v, carry);
      if (i < tl) qS[i] = v;
      carry = __shfl_sync(0xffffffffu, v, 0);
    }
    carry = -0x7fffffff;
    for (int base = 0; base < tl; base += 32) {
      const int i = base + lane;
      int v = (i < tl) ? qE[i] : -0x7fffffff;
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = imax(v, u); }
      v = imax(v, carry);
      if (i < tl) qE[i] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  }
}

struct IrAssemble {
  const int32_t *aog_n_blocks; const unsigned long long *aog_block_off; const uint32_t *aog_blocks;
  const int32_t *dp_n_blocks; const unsigned long long *dp_block_off; const uint32_t *dp_blocks;
  int32_t *out_n; unsigned long long *out_off; uint32_t *out_blocks;
  unsigned long long out_cap; unsigned long long *out_cursor; int *err;
};

// one WARP per segment: lanes sum the piece sizes, lane 0 reserves the output range, then every piece is copied cooperatively
__global__ void __launch_bounds__(128) ir_assemble_kernel(IrSegBatch b, IrAssemble a) {
  const int lane = threadIdx.x & 31;
  const int s = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (s >= b.n_seg) return;
  const int np = b.n_pieces[s];
  const uint32_t *P = b.pieces + 4ull * ir_piece_off(b, s);
  int total = 0;
  for (int p = lane; p < np; p += 32) {
    const uint32_t kind = P[4 * p];
    total += kind == IR_PIECE_LITERAL ? 1 : (kind == IR_PIECE_AOG ? a.aog_n_blocks[P[4 * p + 1]] : a.dp_n_blocks[P[4 * p + 1]]);
  }
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  unsigned long long slot = 0;
  if (lane == 0 && total > 0) slot = atomicAdd(a.out_cursor, (unsigned long long)total);
  slot = __shfl_sync(0xffffffffu, slot, 0);
  const bool over = slot + (unsigned long long)total > a.out_cap;
  if (lane == 0) {
    a.out_n[s] = total;
    a.out_off[s] = slot;
    if (over) atomicOr(a.err, 1);
  }
  if (over) return;
  uint32_t *out = a.out_blocks + 3ull * slot;
  int o = 0;
  for (int p = 0; p < np; p++) {
    const uint32_t kind = P[4 * p];
    if (kind == IR_PIECE_LITERAL) {
      if (lane < 3) out[3 * o + lane] = P[4 * p + 1 + lane];
      o++;
    } else if (kind == IR_PIECE_AOG) {
      const uint32_t j = P[4 * p + 1];
      const uint32_t *src = a.aog_blocks + 3ull * a.aog_block_off[j];
      const int n = a.aog_n_blocks[j];
      for (int i = lane; i < 3 * n; i += 32) { const int c = i % 3; out[3 * o + i] = src[i] + (c == 0 ? P[4 * p + 2] : (c == 1 ? P[4 * p + 3] : 0u)); }
      o += n;
    } else {
      const uint32_t g = P[4 * p + 1];
      const uint32_t *src = a.dp_blocks + 3ull * a.dp_block_off[g];
      const int n = a.dp_n_blocks[g];
      for (int i = lane; i < 3 * n; i += 32) out[3 * o + i] = src[i];
      o += n;
    }
  }
  __syncwarp();
  // "ERROR with alignment consistency" check of the reference (IndelRefine.h:772-782): reported as a status flag
  int bad = 0;
  for (int i = lane; i + 1 < total; i += 32)
    if ((unsigned long long)out[3 * i] + out[3 * i + 2] > out[3 * (i + 1)] || (unsigned long long)out[3 * i + 1] + out[3 * i + 2] > out[3 * (i + 1) + 1]) bad = 1;
  if (bad) atomicOr(a.err, 64);
}

}  // namespace lra
