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

}  // namespace lra
