// a6: DiagonalSort / AntiDiagonalSort / CartesianSort / CartesianTargetSort (reference Sorting.h:33-224), batched over segments
// (the matches of one read strand, of one cluster, ...).
// Every comparator is a total order on the (read position, genome position) pair -- two anchors that compare equal are the same anchor,
// tuple values included -- so std::sort's instability is unobservable here and any correct sort gives the reference's result
// (unlike a3 / a12, where equal keys carry different payloads and libstdc++'s introsort is replayed).
//   mode 0  DiagonalSort        ((long) q - (long) t, q)                      Sorting.h:33-47
//   mode 1  AntiDiagonalSort    ((GenomePos)(q + t)  [32-bit wrap], q)        Sorting.h:77-90
//   mode 2  CartesianSort       (q, t)                                        Sorting.h:132-143
//   mode 3  CartesianTargetSort (t, q)                                        Sorting.h:182-193
// One CTA per segment, bitonic network over (64-bit primary key, 32-bit secondary key, 32-bit source index) records: in shared memory
// when the padded segment fits (<= 2048 records), otherwise in a global scratch slot.  The source index makes the result a permutation
// the caller can apply to any payload (tuple values, strands, lengths).
#pragma once
#include "lra_common.cuh"

namespace lra {

struct SortBatch {
  int n_seg, mode;
  const unsigned long long *seg_off;   // [n_seg + 1]
  uint32_t *q, *t;                     // sorted in place
  uint32_t *perm;                      // [total] source index (batch-wide) of every output element, or nullptr
  unsigned long long *kp;              // scratch: primary keys,   slot_off[s] .. (power-of-two sized slots, large segments only)
  uint32_t *ks, *ki;                   // scratch: secondary keys, source indices
  const unsigned long long *slot_off;  // [n_seg]
  const uint8_t *seg_mode = nullptr;   // [n_seg] per-segment mode (a15: DiagonalSort on strand 0, AntiDiagonalSort on strand 1), or nullptr: `mode` for all
};

constexpr int kSortSmem = 2048;

__device__ __forceinline__ void sort_keys(int mode, uint32_t q, uint32_t t, unsigned long long &p, uint32_t &s) {
  if (mode == 0) { p = (unsigned long long)((long long)q - (long long)t + (1ll << 32)); s = q; }
  else if (mode == 1) { p = (unsigned long long)(uint32_t)(q + t); s = q; }
  else if (mode == 2) { p = q; s = t; }
  else { p = t; s = q; }
}

__global__ void __launch_bounds__(256) sort_pairs_kernel(SortBatch b) {
  __shared__ unsigned long long sp[kSortSmem];
  __shared__ uint32_t ss[kSortSmem], si[kSortSmem];
  const int s = blockIdx.x;
  if (s >= b.n_seg) return;
  const unsigned long long o = b.seg_off[s];
  const int n = (int)(b.seg_off[s + 1] - o);
  const int mode = b.seg_mode ? (int)b.seg_mode[s] : b.mode;
  if (mode == 255) return;            // a segment its caller leaves unsorted (TrimSplitChainDiagonal: chains of one anchor)
  if (n <= 1) { if (n == 1 && b.perm && threadIdx.x == 0) b.perm[o] = (uint32_t)o; return; }
  int P = 1;
  while (P < n) P <<= 1;
  const bool in_smem = P <= kSortSmem;
  unsigned long long *kp = in_smem ? sp : b.kp + b.slot_off[s];
  uint32_t *ks = in_smem ? ss : b.ks + b.slot_off[s];
  uint32_t *ki = in_smem ? si : b.ki + b.slot_off[s];
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    if (i < n) { unsigned long long p; uint32_t s2; sort_keys(mode, b.q[o + i], b.t[o + i], p, s2); kp[i] = p; ks[i] = s2; ki[i] = (uint32_t)i; }
    else { kp[i] = ~0ull; ks[i] = 0xFFFFFFFFu; ki[i] = 0xFFFFFFFFu; }
  }
  __syncthreads();
  for (int kk = 2; kk <= P; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int x = threadIdx.x; x < (P >> 1); x += blockDim.x) {
        const int i = ((x & ~(j - 1)) << 1) | (x & (j - 1));      // the lower index of the x-th compare-exchange pair of this stage
        const int ixj = i | j;
        const unsigned long long a = kp[i], c = kp[ixj];
        const uint32_t a2 = ks[i], c2 = ks[ixj];
        const bool gt = a != c ? a > c : a2 > c2;
        const bool up = (i & kk) == 0;
        if (gt == up && (a != c || a2 != c2)) {
          kp[i] = c; kp[ixj] = a; ks[i] = c2; ks[ixj] = a2;
          const uint32_t u = ki[i]; ki[i] = ki[ixj]; ki[ixj] = u;
        }
      }
      __syncthreads();
    }
  }
  // records -> (q, t): both coordinates are recoverable from the two keys, the source index gives the permutation
  uint32_t *oq = b.q + o, *ot = b.t + o;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long p = kp[i];
    const uint32_t s2 = ks[i];
    uint32_t nq, nt;
    if (mode == 0) { nq = s2; nt = (uint32_t)((long long)s2 - ((long long)p - (1ll << 32))); }
    else if (mode == 1) { nq = s2; nt = (uint32_t)p - s2; }
    else if (mode == 2) { nq = (uint32_t)p; nt = s2; }
    else { nt = (uint32_t)p; nq = s2; }
    oq[i] = nq; ot[i] = nt;
    if (b.perm) b.perm[o + i] = (uint32_t)(o + ki[i]);
  }
}

}  // namespace lra
