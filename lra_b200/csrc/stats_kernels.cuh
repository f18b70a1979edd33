// a21  Alignment::CalculateStatistics (showmm == true), batched over segments.
// Reference: CreateAlignmentStrings Alignment.h:247-333, AlignStringsToCigar :414-504, CalculateStatistics :513-531.
// One segment per thread, two passes (count CIGAR ops -> exclusive scan -> emit).  Bases are compared through seqMap
// (non-ACGT -> 0), i.e. on the 2-bit plane of the packed arenas only.  NV (`value`) is accumulated in binary32 in CIGAR
// order with explicitly rounded, never-fused operations and the HOST-built logf table (LogLookUpTable.h:9-15), which
// is what makes it bit-identical to the reference.
#pragma once
#include "lra_common.cuh"

namespace lra {

struct StatsBatch {
  SeqView q, t;
  const uint32_t *blocks;
  const unsigned long long *blk_off;
  const int32_t *blk_cnt;
  const uint32_t *q_base, *t_base;
  const int32_t *read_len;
  int n_seg;
  const float *lut;               // 2001 floats
  int32_t *stats;                 // [n_seg][16]: nm,nmm,n_D,n_I,tdel,tins,nSmallDel,nMedDel,nLargeDel,nSmallIns,nMedIns,nLargeIns,refLen,preClip,sufClip,n_cigar
  float *value;                   // [n_seg]
  unsigned long long *cig_off;    // [n_seg+1] counts, then offsets
  uint32_t *cigar;                // BAM-style (len << 4 | op): '=' 7, 'X' 8, 'I' 1, 'D' 2
  unsigned long long cigar_cap;
  // scratch of the warp kernels
  uint32_t *pre;                  // [total blocks][3] column / read / target offsets of every block's region
  uint32_t *lane_info;            // [n_seg][64] per lane: first op index | skip-first flag, extra length of the last run
};

__device__ __forceinline__ int seq_code2(const SeqView &s, uint64_t p) { return (int)((s.b2[p >> 4] >> ((uint32_t)(p & 15) * 2)) & 3u); }

template <bool EMIT>
__global__ void __launch_bounds__(128) stats_kernel(StatsBatch b) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= b.n_seg) return;
  const int nb = b.blk_cnt[s];
  const uint32_t *B = b.blocks + 3ull * b.blk_off[s];
  const uint64_t qb = b.q_base[s];
  const uint32_t tb = b.t_base[s];
  unsigned long long nops = 0;
  const unsigned long long obase = EMIT ? b.cig_off[s] : 0ull;
  int nm = 0, nmm = 0, nD = 0, nI = 0, tdel = 0, tins = 0, sD = 0, mD = 0, lD = 0, sI = 0, mI = 0, lI = 0;
  float value = 0.0f;
  if (nb == 0) {
    if (EMIT) { for (int i = 0; i < 16; i++) b.stats[16 * s + i] = 0; b.value[s] = 0.0f; }
    else b.cig_off[s] = 0;
    return;
  }
  int cur = -1;
  long long run = 0;
  uint32_t q = B[0], t = B[1];
  auto flush = [&]() {
    if (run > 0) {
      if (EMIT) {
        const int op = cur == 0 ? 7 : cur == 1 ? 8 : cur == 2 ? 2 : 1;
        if (obase + nops < b.cigar_cap) b.cigar[obase + nops] = ((uint32_t)run << 4) | (uint32_t)op;
        if (cur == 0) { nm += (int)run; value = __fadd_rn(value, (float)(int)run); }
        else if (cur == 1) { nmm += (int)run; value = __fsub_rn(value, (float)(int)run); }
        else {
          if (cur == 2) { tdel += (int)run; nD++; if (run <= 10) sD++; if (run > 10 && run < 50) mD++; else if (run > 50) lD++; }
          else { tins += (int)run; nI++; if (run <= 10) sI++; if (run > 10 && run < 50) mI++; else if (run > 50) lI++; }
          if (run <= 20) { value = __fsub_rn(value, (float)(int)run); if (cur == 3) sI++; }
          else if (run <= 10001) { const int a = (int)((run - 1) / 5); value = __fadd_rn(value, __fsub_rn(__fmul_rn(-3.0f, b.lut[a]), 1.0f)); }
          else if (run <= 100001) value = __fadd_rn(value, -1000.0f);
          else value = __fadd_rn(value, -2000.0f);
        }
      }
      nops++;
      run = 0;
    }
  };
  auto col = [&](int c) { if (c != cur) { flush(); cur = c; } run++; };
  for (int bi = 0; bi < nb; bi++) {
    const uint32_t len = B[3 * bi + 2];
    for (uint32_t bl = 0; bl < len; bl++, q++, t++) col(seq_code2(b.q, qb + q) != seq_code2(b.t, (uint64_t)(uint32_t)(tb + t)) ? 1 : 0);
    if (bi == nb - 1) continue;
    int qg = (int)(B[3 * (bi + 1)] - B[3 * bi] - len);
    int tg = (int)(B[3 * (bi + 1) + 1] - B[3 * bi + 1] - len);
    if (qg > 0 || tg > 0) {
      const int common = qg > tg ? tg : qg;
      tg -= common; qg -= common;
      for (int g = 0; g < qg; g++, q++) col(3);
      for (int g = 0; g < tg; g++, t++) col(2);
      for (int g = 0; g < common; g++, q++, t++) col(seq_code2(b.q, qb + q) != seq_code2(b.t, (uint64_t)(uint32_t)(tb + t)) ? 1 : 0);
    }
  }
  flush();
  if (!EMIT) { b.cig_off[s] = nops; return; }
  int32_t *st = b.stats + 16 * s;
  st[0] = nm; st[1] = nmm; st[2] = nD; st[3] = nI; st[4] = tdel; st[5] = tins; st[6] = sD; st[7] = mD; st[8] = lD; st[9] = sI; st[10] = mI; st[11] = lI;
  st[12] = (int32_t)t; st[13] = (int32_t)B[0]; st[14] = b.read_len[s] - (int32_t)B[3 * (nb - 1)] - (int32_t)B[3 * (nb - 1) + 2]; st[15] = (int32_t)nops;
  b.value[s] = value;
}

// ---- one WARP per segment.
// The alignment is a stream of columns: block bi contributes `len` match/mismatch columns, then (gap to the next block) qg' 'I'
// columns, tg' 'D' columns and `common` match/mismatch columns (CreateAlignmentStrings, Alignment.h:283-327); the CIGAR is the run-length
// encoding of the column types.  Pass A (lanes over blocks) turns the per-block column counts and read/target advances into offsets
// with warp scans; pass B gives every lane a contiguous 1/32 of the columns, which it run-length encodes locally; a 32-step fix-up
// merges runs across lane boundaries (a run may span any number of lanes).  The count kernel stores, per lane, the index of its first
// op and the length its last run gains from the lanes after it; the emit kernel walks the columns once more and writes the ops.
// NV: without a gap longer than 20 every term is an integer and every partial sum is exact in binary32, so the sum is taken in integers;
// otherwise the warp replays the reference's float accumulation over the ops in CIGAR order (Alignment.h:466-500).
__device__ __forceinline__ void stats_region(const uint32_t *B, int bi, int nb, int &len, int &qgp, int &tgp, int &com) {
  len = (int)B[3 * bi + 2];
  qgp = tgp = com = 0;
  if (bi < nb - 1) {
    int qg = (int)(B[3 * (bi + 1)] - B[3 * bi] - (uint32_t)len);
    int tg = (int)(B[3 * (bi + 1) + 1] - B[3 * bi + 1] - (uint32_t)len);
    if (qg > 0 || tg > 0) {
      const int common = qg > tg ? tg : qg;
      tg -= common; qg -= common;
      qgp = qg > 0 ? qg : 0; tgp = tg > 0 ? tg : 0; com = common > 0 ? common : 0;
    }
  }
}

struct StatsAcc { int nm, nmm, nD, nI, tdel, tins, sD, mD, lD, sI, mI, lI; long long isum; int nlong; };

__device__ __forceinline__ void stats_op(StatsAcc &a, int ty, long long run) {
  if (ty == 0) { a.nm += (int)run; a.isum += run; }
  else if (ty == 1) { a.nmm += (int)run; a.isum -= run; }
  else {
    if (ty == 2) { a.tdel += (int)run; a.nD++; if (run <= 10) a.sD++; if (run > 10 && run < 50) a.mD++; else if (run > 50) a.lD++; }
    else { a.tins += (int)run; a.nI++; if (run <= 10) a.sI++; if (run > 10 && run < 50) a.mI++; else if (run > 50) a.lI++; }
    if (run <= 20) { a.isum -= run; if (ty == 3) a.sI++; }
    else a.nlong++;
  }
}

template <bool EMIT>
__global__ void __launch_bounds__(128) stats_warp_kernel(StatsBatch b) {
  const int lane = threadIdx.x & 31;
  const int s = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (s >= b.n_seg) return;
  const int nb = b.blk_cnt[s];
  if (nb == 0) {
    if (EMIT) { if (lane < 16) b.stats[16 * s + lane] = 0; if (lane == 0) b.value[s] = 0.0f; }
    else if (lane == 0) b.cig_off[s] = 0;
    return;
  }
  const uint32_t *B = b.blocks + 3ull * b.blk_off[s];
  uint32_t *pre = b.pre + 3ull * b.blk_off[s];
  const uint64_t qb = b.q_base[s];
  const uint32_t tb = b.t_base[s];
  // ---- pass A: exclusive offsets (columns, read advance, target advance) of every block's region
  uint32_t cC = 0, cQ = 0, cT = 0;
  for (int base = 0; base < nb; base += 32) {
    const int bi = base + lane;
    uint32_t nc = 0, qa = 0, ta = 0;
    if (bi < nb) { int len, qgp, tgp, com; stats_region(B, bi, nb, len, qgp, tgp, com); nc = (uint32_t)(len + qgp + tgp + com); qa = (uint32_t)(len + qgp + com); ta = (uint32_t)(len + tgp + com); }
    uint32_t ic = nc, iq = qa, it = ta;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, ic, o), y = __shfl_up_sync(0xffffffffu, iq, o), z = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) { ic += x; iq += y; it += z; }
    }
    if (bi < nb) { pre[3 * bi] = cC + ic - nc; pre[3 * bi + 1] = cQ + iq - qa; pre[3 * bi + 2] = cT + it - ta; }
    cC += __shfl_sync(0xffffffffu, ic, 31); cQ += __shfl_sync(0xffffffffu, iq, 31); cT += __shfl_sync(0xffffffffu, it, 31);
  }
  __syncwarp();
  const uint32_t C = cC;
  // ---- pass B: this lane's columns [lo, hi)
  const uint32_t per = (C + 31u) / 32u;
  const uint32_t lo = lane * per < C ? lane * per : C, hi = lo + per < C ? lo + per : C;
  int nruns = 0, firstTy = -1, lastTy = -1;
  uint32_t firstLen = 0, curLen = 0;
  int cur = -1;
  // emit-mode state
  uint32_t obase_l = 0, extra = 0; bool skipFirst = false;
  if (EMIT) { const uint32_t w0 = b.lane_info[64 * s + 2 * lane]; obase_l = w0 >> 1; skipFirst = (w0 & 1u) != 0; extra = b.lane_info[64 * s + 2 * lane + 1]; }
  const unsigned long long obase = EMIT ? b.cig_off[s] : 0ull;
  StatsAcc acc;
  acc.nm = acc.nmm = acc.nD = acc.nI = acc.tdel = acc.tins = acc.sD = acc.mD = acc.lD = acc.sI = acc.mI = acc.lI = 0; acc.isum = 0; acc.nlong = 0;
  auto flush = [&](bool last) {
    if (cur < 0) return;
    if (nruns == 0) { firstTy = cur; firstLen = curLen; }
    if (EMIT) {
      if (!(nruns == 0 && skipFirst)) {
        const long long run = (long long)curLen + (last ? (long long)extra : 0ll);
        const unsigned long long o = obase + obase_l + (unsigned)(nruns - (skipFirst ? 1 : 0));
        const int op = cur == 0 ? 7 : cur == 1 ? 8 : cur == 2 ? 2 : 1;
        if (o < b.cigar_cap) b.cigar[o] = ((uint32_t)run << 4) | (uint32_t)op;
        stats_op(acc, cur, run);
      }
    }
    nruns++; lastTy = cur;
  };
  auto put = [&](int ty, uint32_t n) { if (ty != cur) { flush(false); cur = ty; curLen = 0; } curLen += n; };
  if (lo < hi) {
    int bi;
    { int l2 = 0, len2 = nb;        // last block whose region starts at or before lo
      while (len2 > 0) { const int half = len2 >> 1; if (pre[3 * (l2 + half)] <= lo) { l2 += half + 1; len2 -= half + 1; } else len2 = half; }
      bi = l2 - 1; }
    uint32_t pos = lo;
    uint32_t off = lo - pre[3 * bi];
    while (pos < hi) {
      int len, qgp, tgp, com;
      stats_region(B, bi, nb, len, qgp, tgp, com);
      const uint32_t nc = (uint32_t)(len + qgp + tgp + com);
      if (off >= nc) { bi++; off = 0; continue; }
      const uint64_t q0 = qb + B[0] + pre[3 * bi + 1];
      const uint32_t t0 = tb + B[1] + pre[3 * bi + 2];
      while (off < (uint32_t)len && pos < hi) {
        put(seq_code2(b.q, q0 + off) != seq_code2(b.t, (uint64_t)(uint32_t)(t0 + off)) ? 1 : 0, 1u);
        off++; pos++;
      }
      if (off >= (uint32_t)len && off < (uint32_t)(len + qgp) && pos < hi) {
        const uint32_t n = imin((int)((uint32_t)(len + qgp) - off), (int)(hi - pos));
        put(3, n); off += n; pos += n;
      }
      if (off >= (uint32_t)(len + qgp) && off < (uint32_t)(len + qgp + tgp) && pos < hi) {
        const uint32_t n = imin((int)((uint32_t)(len + qgp + tgp) - off), (int)(hi - pos));
        put(2, n); off += n; pos += n;
      }
      while (off >= (uint32_t)(len + qgp + tgp) && off < nc && pos < hi) {
        const uint32_t j = off - (uint32_t)(len + qgp + tgp);
        put(seq_code2(b.q, q0 + (uint32_t)(len + qgp) + j) != seq_code2(b.t, (uint64_t)(uint32_t)(t0 + (uint32_t)(len + tgp) + j)) ? 1 : 0, 1u);
        off++; pos++;
      }
    }
    flush(true);
  }
  const uint32_t lastLen = curLen;
  if (!EMIT) {
    // ---- fix-up across lanes: which first runs continue the run left open by an earlier lane
    int openTy = -1, owner = -1;
    uint32_t ops = 0, myBase = 0, myExtra = 0;
    bool mySkip = false;
    for (int l = 0; l < 32; l++) {
      const int n_l = __shfl_sync(0xffffffffu, nruns, l);
      const int f_l = __shfl_sync(0xffffffffu, firstTy, l), e_l = __shfl_sync(0xffffffffu, lastTy, l);
      const uint32_t fl_l = __shfl_sync(0xffffffffu, firstLen, l);
      if (n_l == 0) continue;
      if (lane == l) myBase = ops;
      if (f_l == openTy) {
        if (lane == l) mySkip = true;
        if (lane == owner) myExtra += fl_l;
        ops += (uint32_t)(n_l - 1);
        if (n_l > 1) { owner = l; openTy = e_l; }
      } else {
        ops += (uint32_t)n_l;
        owner = l; openTy = e_l;
      }
    }
    b.lane_info[64 * s + 2 * lane] = (myBase << 1) | (mySkip ? 1u : 0u);
    b.lane_info[64 * s + 2 * lane + 1] = myExtra;
    if (lane == 0) b.cig_off[s] = ops;
    (void)lastLen;
    return;
  }
  // ---- emit: reduce the per-op statistics
  auto rsum = [&](int v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; };
  const int nm = rsum(acc.nm), nmm = rsum(acc.nmm), nD = rsum(acc.nD), nI = rsum(acc.nI), tdel = rsum(acc.tdel), tins = rsum(acc.tins);
  const int sD = rsum(acc.sD), mD = rsum(acc.mD), lD = rsum(acc.lD), sI = rsum(acc.sI), mI = rsum(acc.mI), lI = rsum(acc.lI), nlong = rsum(acc.nlong);
  long long isum = acc.isum;
  for (int o = 16; o > 0; o >>= 1) isum += __shfl_xor_sync(0xffffffffu, isum, o);
  const unsigned long long nops = b.cig_off[s + 1] - obase;
  float value;
  if (nlong == 0) value = (float)isum;
  else {
    __syncwarp();
    value = 0.0f;
    for (unsigned long long base = 0; base < nops; base += 32) {
      const unsigned long long i = base + lane;
      const uint32_t mine = (i < nops && obase + i < b.cigar_cap) ? b.cigar[obase + i] : 0u;
      const int cnt = (int)(nops - base < 32ull ? nops - base : 32ull);
      for (int j = 0; j < cnt; j++) {
        const uint32_t w = __shfl_sync(0xffffffffu, mine, j);
        const long long run = (long long)(w >> 4);
        const int op = (int)(w & 15u);
        if (op == 7) value = __fadd_rn(value, (float)(int)run);
        else if (op == 8) value = __fsub_rn(value, (float)(int)run);
        else if (run <= 20) value = __fsub_rn(value, (float)(int)run);
        else if (run <= 10001) { const int a = (int)((run - 1) / 5); value = __fadd_rn(value, __fsub_rn(__fmul_rn(-3.0f, b.lut[a]), 1.0f)); }
        else if (run <= 100001) value = __fadd_rn(value, -1000.0f);
        else value = __fadd_rn(value, -2000.0f);
      }
    }
  }
  if (lane == 0) {
    int32_t *st = b.stats + 16 * s;
    st[0] = nm; st[1] = nmm; st[2] = nD; st[3] = nI; st[4] = tdel; st[5] = tins; st[6] = sD; st[7] = mD; st[8] = lD; st[9] = sI; st[10] = mI; st[11] = lI;
    st[12] = (int32_t)(B[1] + cT); st[13] = (int32_t)B[0]; st[14] = b.read_len[s] - (int32_t)B[3 * (nb - 1)] - (int32_t)B[3 * (nb - 1) + 2]; st[15] = (int32_t)nops;
    b.value[s] = value;
  }
}

}  // namespace lra
