// a18 AffineOneGapAlign -- warp-per-job kernel for WIDE one-sided bands (doubled half-width k up to 127).
// Reference: AffineOneGapAlign.h:157-362 (prefix matrix only: the "diag + 2k >= max(len)" mode, :196-203) and the
// traceback :582-647.
//
// Layout: band cell c = i - (j - k) in [0, 2k]; lane L owns the C consecutive cells c = L*C .. L*C+C-1 in registers.
// Row j needs from row j-1 the same c (diagonal) and c+1 (target gap): one shuffle for the neighbour lane's first cell.
// The in-row dependency S[c] = max(T[c], S[c-1] + indel) is a max-plus scan: local sequential pass + 5-step warp scan of
// lane carries.  Query codes are staged once per job in shared memory (one byte per base, so a lane's C cells read C
// consecutive bytes); target codes are fetched 32 rows at a time into one register per lane and broadcast by shuffle.
// Arrows: 2 bits per cell, one 32-bit word per lane per row, written coalesced (128 B per row) into a per-warp slab.
#pragma once
#include "aog_kernels.cuh"

namespace lra {

struct AogBandScratch {
  unsigned char *base;
  unsigned long long slab_bytes;
  uint32_t max_rows, max_qlen;
};
__host__ __device__ inline unsigned long long aog_band_slab_bytes(uint32_t max_rows, uint32_t max_qlen) {
  return ((unsigned long long)max_rows + 2ull) * 128ull + ((unsigned long long)max_qlen + 2ull) * 12ull + 64ull;
}

constexpr int kBandQOff = 128;                      // smem index of query position 0
constexpr int kBandQBytes = kBandQOff + 4000 + 272; // qLen <= 4000 (class condition) + right overhang of the 32*C window

template <int C>
__global__ void __launch_bounds__(128) aog_warp_band_kernel(AogBatch b, AogPlan *plan, const uint32_t *sorted, AogBandScratch sc) {
  constexpr int cls = kAogClsBand1 + (C == 1 ? 0 : C == 2 ? 1 : C == 4 ? 2 : 3);
  __shared__ uint8_t qsm_all[4][kBandQBytes];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  uint8_t *qsm = qsm_all[wib];
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t begin = plan->bin_start[cls * kAogBuckets];
  const uint32_t end = plan->bin_start[(cls + 1) * kAogBuckets];
  unsigned char *slab = sc.base + (unsigned long long)warp_global * sc.slab_bytes;
  uint32_t *tb = (uint32_t *)slab;
  uint32_t *rblk = (uint32_t *)(slab + ((unsigned long long)sc.max_rows + 2ull) * 128ull);

  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&plan->work[cls], 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (begin + w >= end) break;
    const int job = (int)sorted[begin + w];
    const int qLen = b.q_len[job], tLen = b.t_len[job];
    const uint32_t qoff = b.q_off[job], toff = b.t_off[job];
    const int diag = imax(1, imin(qLen, tLen));
    const int k = 2 * imin(diag, b.k[job]);
    const int qB = imin(diag + k, qLen + 1), tB = imin(diag + k, tLen + 1);
    const int rows = tB - 1;
    const bool keep0 = !((qLen >= tLen && diag - k - 1 >= 0) || (qLen <= tLen && diag >= 2));
    const int m = b.m, mm = b.mm, indel = b.indel;

    __syncwarp();
    for (int p = lane; p < kBandQOff + 1; p += 32) qsm[p] = 5;
    {
      const int hi = imin(qLen + 272, kBandQBytes - kBandQOff - 1);
      for (int p = 1 + lane; p <= hi; p += 32)
        qsm[kBandQOff + p] = (p <= qLen) ? (uint8_t)seq_code(b.q, (uint64_t)qoff + (uint64_t)(p - 1)) : (uint8_t)5;
    }
    __syncwarp();

    int P[C];
#pragma unroll
    for (int x = 0; x < C; x++) {
      const int i = lane * C + x - k;
      P[x] = (i < 0 || i > k) ? kMissing : indel * i;
    }
    int tbuf = 5;
    const int d = C * indel;
    for (int j = 1; j <= rows; j++) {
      if (((j - 1) & 31) == 0) {
        const int jj = j + lane;
        tbuf = (jj <= tLen) ? seq_code(b.t, (uint64_t)toff + (uint64_t)(jj - 1)) : 5;
      }
      const int tc = __shfl_sync(0xffffffffu, tbuf, (j - 1) & 31);
      int pn = __shfl_down_sync(0xffffffffu, P[0], 1);
      if (lane == 31) pn = kMissing;
      int sM[C], sD[C], L[C];
      const uint8_t *qrow = qsm + kBandQOff + (j - k + lane * C);
#pragma unroll
      for (int x = 0; x < C; x++) {
        sM[x] = P[x] + ((int)qrow[x] == tc ? m : mm);
        sD[x] = ((x < C - 1) ? P[(x + 1) % C] : pn) + indel;
        const int t = imax(sM[x], sD[x]);
        L[x] = (x == 0) ? t : imax(t, L[(x + C - 1) % C] + indel);
      }
      int v = L[C - 1] - (lane + 1) * d;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = imax(v, u);
      }
      int excl = __shfl_up_sync(0xffffffffu, v, 1);
      if (lane == 0) excl = kNegInf;
      const int left = (j == k + 1 && keep0) ? indel * (k + 1) : kMissing;
      const int carry = lane * d + imax(left, excl);
      int prevS = carry;
      uint32_t bits = 0;
#pragma unroll
      for (int x = 0; x < C; x++) {
        int S = imax(L[x], carry + (x + 1) * indel);
        const int sI = prevS + indel;
        const int arrow = (S == sI) ? AR_LEFT : ((S == sD[x]) ? AR_DOWN : AR_DIAG);
        bits |= (uint32_t)arrow << (2 * x);
        if (lane * C + x > 2 * k) S = kMissing;  // right rail and beyond stay MISSING
        prevS = S;
        P[x] = S;
      }
      tb[(unsigned)j * 32u + lane] = bits;
    }
    // score at the traceback start cell
    const int cstar = (qB - 1) - (tB - 1) + k;
    int score = 0;
#pragma unroll
    for (int x = 0; x < C; x++) if (x == cstar % C) score = P[x];
    score = __shfl_sync(0xffffffffu, score, cstar / C);
    __syncwarp();
    // traceback: warp-uniform walk; every lane holds its word of the current row
    int i = qB - 1, j = tB - 1, run = 0, nb = 0;
    uint32_t wv = (j > 0) ? tb[(unsigned)j * 32u + lane] : 0u;
    while (i > 0 && j > 0) {
      const int c = i - j + k;
      const uint32_t word = __shfl_sync(0xffffffffu, wv, c / C);
      const int a = (int)((word >> (2 * (c % C))) & 3u);
      if (a == AR_DIAG) { run++; i--; j--; if (j > 0) wv = tb[(unsigned)j * 32u + lane]; }
      else {
        if (run) { if (lane == 0) { rblk[3 * nb] = (uint32_t)i; rblk[3 * nb + 1] = (uint32_t)j; rblk[3 * nb + 2] = (uint32_t)run; } nb++; run = 0; }
        if (a == AR_LEFT) i--; else { j--; if (j > 0) wv = tb[(unsigned)j * 32u + lane]; }
      }
    }
    if (run) { if (lane == 0) { rblk[3 * nb] = (uint32_t)i; rblk[3 * nb + 1] = (uint32_t)j; rblk[3 * nb + 2] = (uint32_t)run; } nb++; }
    unsigned long long slot = aog_reserve_blocks(b, lane == 0 ? nb : 0, lane, &plan->cls_blocks[cls]);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (lane == 0) { b.score[job] = score; b.n_blocks[job] = nb; b.block_off[job] = slot; }
    __syncwarp();
    if (slot != ~0ull) {
      uint32_t *out = b.blocks + 3ull * slot;
      for (int r = lane; r < nb; r += 32) {
        const int s2 = nb - 1 - r;
        out[3 * r] = rblk[3 * s2]; out[3 * r + 1] = rblk[3 * s2 + 1]; out[3 * r + 2] = rblk[3 * s2 + 2];
      }
    }
    __syncwarp();
  }
}

}  // namespace lra
