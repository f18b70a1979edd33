// In-worker AffineOneGapAlign for WIDE ONE-SIDED bands: the register-band form of aog_warp_band_kernel (aog_band_kernel.cuh; reference
// AffineOneGapAlign.h:157-362 prefix matrix + traceback :582-647) made callable by a mapper warp, and the dispatcher the worker calls.
// The literal form (mp_aog.cuh) keeps the reference's flat matrices in the arena and costs several global-memory round trips per row; jobs that are
// not two-sided (diag + 2 k0 >= max(qLen, tLen)), with doubled half-width 2 k + 2 <= 256 cells and qLen <= 4000, run here instead: lane L owns C
// consecutive band cells in registers, a row is one local pass + one 5-step max-plus warp scan, arrows are one coalesced 128-byte store per row.
// That covers RefineSpace's alignments (band 30) and the RefineByLinearAlignment jobs that are too long or too wide for the one-job-per-lane form.
// Query codes are staged once per job in a byte array (arena); results are the ones the literal form gives (both are pinned on the reference).
#pragma once
#include "mp_aog.cuh"
#include "aog_band_kernel.cuh"

namespace lra {
namespace mp {

#if MP_LANES == 32
template <int C>
__device__ __noinline__ int mp_aog_band(const SeqView &q, uint32_t qoff, int qLen, const SeqView &t, uint32_t toff, int tLen, int m, int mm, int indel, int k_in,
                                        Arena &ar, uint32_t *blocks_out, int cap, int *n_blocks) {
  const int lane = lane_id();
  const int diag = imax(1, imin(qLen, tLen));
  const int k = 2 * imin(diag, k_in);
  const int qB = imin(diag + k, qLen + 1), tB = imin(diag + k, tLen + 1);
  const int rows = tB - 1;
  const bool keep0 = !((qLen >= tLen && diag - k - 1 >= 0) || (qLen <= tLen && diag >= 2));
  const unsigned long long mk = ar.mark();
  uint32_t *tb = ar.alloc<uint32_t>(((unsigned long long)rows + 2ull) * 32ull);
  uint32_t *rblk = ar.alloc<uint32_t>(3ull * ((unsigned long long)diag + 4ull));
  const int qbytes = kBandQOff + qLen + 272 + 32 * C + 16;
  uint8_t *qsm = ar.alloc<uint8_t>((unsigned long long)qbytes);
  if (ar.overflow) { ar.release(mk); *n_blocks = -1; return 0; }
  for (int p = lane; p < kBandQOff + 1; p += 32) qsm[p] = 5;
  for (int p = 1 + lane; p < qbytes - kBandQOff; p += 32) qsm[kBandQOff + p] = (p <= qLen) ? (uint8_t)seq_code(q, (uint64_t)qoff + (uint64_t)(p - 1)) : (uint8_t)5;
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
      tbuf = (jj <= tLen) ? seq_code(t, (uint64_t)toff + (uint64_t)(jj - 1)) : 5;
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
      const int tt = imax(sM[x], sD[x]);
      L[x] = (x == 0) ? tt : imax(tt, L[(x + C - 1) % C] + indel);
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
  const int cstar = (qB - 1) - (tB - 1) + k;
  int score = 0;
#pragma unroll
  for (int x = 0; x < C; x++) if (x == cstar % C) score = P[x];
  score = __shfl_sync(0xffffffffu, score, cstar / C);
  __syncwarp();
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
  __syncwarp();
  if (nb > cap) { ar.release(mk); *n_blocks = -1; return score; }
  for (int r = lane; r < nb; r += 32) {
    const int s2 = nb - 1 - r;
    blocks_out[3 * r] = rblk[3 * s2]; blocks_out[3 * r + 1] = rblk[3 * s2 + 1]; blocks_out[3 * r + 2] = rblk[3 * s2 + 2];
  }
  __syncwarp();
  ar.release(mk);
  *n_blocks = nb;
  return score;
}
#endif

// AffineOneGapAlign for the calling warp: the register-band form where it applies, the literal form otherwise
__device__ __forceinline__ int mp_aog_any(const SeqView &q, uint32_t qoff, int qLen, const SeqView &t, uint32_t toff, int tLen, int m, int mm, int indel, int k_in,
                                          Arena &ar, uint32_t *blocks_out, int cap, int *n_blocks, int *err) {
#if MP_LANES == 32
  const int diag = imax(1, imin(qLen, tLen));
  const int k0 = imin(diag, k_in);
  const bool two_sided = diag + 2 * k0 < imax(qLen, tLen);
  const int k = 2 * k0, need = 2 * k + 2;
  if (!two_sided && k >= 1 && need <= 256 && qLen <= 4000 && qLen >= 1 && tLen >= 1) {
    if (need <= 32) return mp_aog_band<1>(q, qoff, qLen, t, toff, tLen, m, mm, indel, k_in, ar, blocks_out, cap, n_blocks);
    if (need <= 64) return mp_aog_band<2>(q, qoff, qLen, t, toff, tLen, m, mm, indel, k_in, ar, blocks_out, cap, n_blocks);
    if (need <= 128) return mp_aog_band<4>(q, qoff, qLen, t, toff, tLen, m, mm, indel, k_in, ar, blocks_out, cap, n_blocks);
    return mp_aog_band<8>(q, qoff, qLen, t, toff, tLen, m, mm, indel, k_in, ar, blocks_out, cap, n_blocks);
  }
#endif
  return mp_aog(q, qoff, qLen, t, toff, tLen, m, mm, indel, k_in, ar, blocks_out, cap, n_blocks, err);
}

}  // namespace mp
}  // namespace lra
