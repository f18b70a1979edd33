// StoreMinimizers (MinCount.h:7-179) over a range of loop steps: the sequential scan restarted `warm` steps before the range; see gidx_kernels.cuh
// for the argument why the restarted state becomes the sequential one.  Shared by the global index builder (one thread per chunk of a contig) and
// the mapper worker (one lane per chunk of a read).
#pragma once
#include "seed_kernels.cuh"

namespace lra {

// The literal scan of seq[off + s0, off + seqLen) as if it started at s0 (ring slots, positions and window limits in contig coordinates);
// pushes made at loop steps in [p_begin, p_end) are counted (and stored when EMIT).  Returns the count; `certain` says whether the state at
// p_begin is provably the sequential one.
template <bool EMIT>
__device__ __forceinline__ uint32_t gidx_scan_range(const SeqView &seq, unsigned long long off, uint32_t seqLenFull, int k, int w, uint32_t s0, uint32_t p_begin, uint32_t p_end,
                                                    unsigned long long *ot, uint32_t *op, uint32_t gpos0, bool &certain) {
  uint32_t n_out = 0;
  certain = s0 == 0;
  const uint32_t seqLen = seqLenFull - s0;              // length of the sub-sequence
  off += s0;
  if (seqLen < (uint32_t)k) return 0;
  const int windowSpan = w + k - 1;
  if (seqLen < (uint32_t)windowSpan) return 0;
  unsigned long long m = 0;
  for (int i = 0; i < k; i++) { m <<= 2; m += 3; }
  int nextValidWindowEnd = 0, nextValidWindowStart = 0;
  bool valid = false;
  while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
    valid = true;
    for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++) {
      if (seqLen < (uint32_t)n) return 0;
      if (seq_code(seq, off + (unsigned long long)n) > 3) { nextValidWindowStart = n + 1; valid = false; }
    }
  }
  if (!valid) return 0;
  nextValidWindowEnd = nextValidWindowStart + windowSpan;
  SeqStream st;
  st.init(seq, off);
  unsigned long long cur = 0, curRC = 0;
  for (int p = 0; p <= k - 1; p++) { const int c = st.next(); cur <<= 2; cur += (unsigned long long)(c & 3) * (c != 4); }
  { unsigned long long a = cur; for (int i = 0; i < k; i++) { const unsigned long long least = ~(a & 3ull) & 3ull; a >>= 2; curRC <<= 2; curRC += least; } }
  unsigned long long ringT[kSeedMaxW];
  uint32_t ringP[kSeedMaxW];
  unsigned long long actT;
  uint32_t actP = 0;                                   // positions relative to s0 inside the scan; s0 is added when stored
  if ((cur & kForMask) < (curRC & kForMask)) actT = cur & kForMask; else actT = curRC | kRevMask;
  const uint32_t uw = (uint32_t)w;
  ringT[s0 % uw] = actT; ringP[s0 % uw] = 0;
  uint32_t p;
  for (p = 1; p < uw && p < seqLen - (uint32_t)k + 1; p++) {
    const int c = st.next();
    const unsigned long long n2 = (unsigned long long)(c & 3) * (c != 4);
    cur = ((cur << 2) & m) + n2;
    curRC >>= 2; curRC += ((~n2) & 3ull) << (2 * ((unsigned long long)k - 1));
    const unsigned long long ct = ((cur & kForMask) < (curRC & kForMask)) ? (cur & kForMask) : (curRC | kRevMask);
    if (ct < actT) { actT = ct; actP = p; }
    ringT[(p + s0) % uw] = ct; ringP[(p + s0) % uw] = p;
  }
  // the first minimizer is pushed before the main loop: it belongs to the chunk that owns step w - 1 of a scan from the contig start
  if (s0 == 0 && nextValidWindowEnd == windowSpan && p_begin <= uw - 1 && uw - 1 < p_end) {
    if (EMIT) { ot[n_out] = actT; op[n_out] = gpos0 + actP; }
    n_out++;
  }
  for (p = uw; p < seqLen - (uint32_t)k + 1; p++) {
    const uint32_t pa = p + s0;                        // loop step in contig coordinates
    if (pa >= p_end) break;
    const int c = st.next();
    if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k - 1) {
      if (c <= 3) nextValidWindowEnd++;
      else {
        nextValidWindowStart = (int)(p + (uint32_t)k);
        valid = false;
        while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
          valid = true;
          for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++)
            if (seq_code(seq, off + (unsigned long long)n) > 3) { nextValidWindowStart = n + 1; valid = false; }
        }
        if (!valid) return n_out;
        nextValidWindowEnd = nextValidWindowStart + windowSpan;
      }
    }
    const unsigned long long n2 = (unsigned long long)(c & 3) * (c != 4);
    cur = ((cur << 2) & m) + n2;
    curRC >>= 2; curRC += ((~n2) & 3ull) << (2 * ((unsigned long long)k - 1));
    const unsigned long long ct = ((cur & kForMask) < (curRC & kForMask)) ? (cur & kForMask) : (curRC | kRevMask);
    ringT[pa % uw] = ct; ringP[pa % uw] = p;
    if (pa == p_begin && !certain) return 0xffffffffu;                 // the state entering the chunk is not provably the sequential one
    const bool own = pa >= p_begin;
    bool push = false;
    if (p - uw >= actP) {
      actT = ringT[0]; actP = ringP[0];
      int ties = 1;
      for (int j = 1; j < w; j++) {
        const unsigned long long a = ringT[j] & kForMask, b = actT & kForMask;
        if (a < b) { actT = ringT[j]; actP = ringP[j]; ties = 1; } else if (a == b) ties++;
      }
      if (ties == 1 && p >= 2 * uw) certain = true;
      push = (uint32_t)nextValidWindowEnd == p + (uint32_t)k;
    } else if ((ct & kForMask) < (actT & kForMask)) {
      actT = ct; actP = p;
      if (p >= 2 * uw) certain = true;
      push = (uint32_t)nextValidWindowEnd == p + (uint32_t)k;
    }
    if (push && own) {
      if (EMIT) { ot[n_out] = actT; op[n_out] = gpos0 + s0 + actP; }
      n_out++;
    }
  }
  return n_out;
}

}  // namespace lra
