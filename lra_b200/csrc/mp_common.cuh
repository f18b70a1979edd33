// Mapper worker framework: one warp maps one read (or one chain of a read) through the glue stages of MapRead_lowacc
// (reference Map_lowacc.h:69-632), calling warp-cooperative device routines for the heavy parts.
//
// Execution model ("warp-uniform"): all lanes of the worker warp execute the scalar control flow redundantly on identical
// values (same-address loads broadcast, same-address stores collapse), and split the data-parallel loops
// `for (i = lane; i < n; i += kLanes)`.  A wsync() separates a lane-parallel region from code that reads what other lanes
// wrote.  MP_LANES == 1 (tests only, SIMT emulator) turns every collective into the identity, so the same source runs as
// plain serial code at CPU speed for bulk parity runs; MP_LANES == 32 is the product.
#pragma once
#include "lra_common.cuh"

#ifndef MP_LANES
#define MP_LANES 32
#endif

#if defined(LRA_EMU) && defined(MP_DEBUG)
#include <cstdio>
#include <cstdlib>
#define MP_CHECK(c) do { if (!(c)) { fprintf(stderr, "MP_CHECK failed: %s at %s:%d\n", #c, __FILE__, __LINE__); abort(); } } while (0)
#else
#define MP_CHECK(c) do { } while (0)
#endif

namespace lra {
namespace mp {

constexpr int kLanes = MP_LANES;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() {
#if MP_LANES == 1
  return 0;
#else
  return (int)(threadIdx.x & 31u);
#endif
}
__device__ __forceinline__ void wsync() {
#if MP_LANES > 1
  __syncwarp();
#endif
}
template <class T> __device__ __forceinline__ T bcast(T v, int src) {
#if MP_LANES == 1
  (void)src; return v;
#else
  return __shfl_sync(kFull, v, src);
#endif
}
template <class T> __device__ __forceinline__ T shfl_xor(T v, int m) {
#if MP_LANES == 1
  (void)m; return v;
#else
  return __shfl_xor_sync(kFull, v, m);
#endif
}
template <class T> __device__ __forceinline__ T shfl_up(T v, int d) {
#if MP_LANES == 1
  (void)d; return v;
#else
  return __shfl_up_sync(kFull, v, (unsigned)d);
#endif
}
__device__ __forceinline__ unsigned ballot(bool p) {
#if MP_LANES == 1
  return p ? 1u : 0u;
#else
  return __ballot_sync(kFull, p ? 1 : 0);
#endif
}
__device__ __forceinline__ bool wany(bool p) { return ballot(p) != 0u; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << lane_id()) - 1u; }

template <class T> __device__ __forceinline__ T wmax(T v) {
#if MP_LANES > 1
  for (int o = 16; o > 0; o >>= 1) { T u = __shfl_xor_sync(kFull, v, o); v = u > v ? u : v; }
#endif
  return v;
}
template <class T> __device__ __forceinline__ T wmin(T v) {
#if MP_LANES > 1
  for (int o = 16; o > 0; o >>= 1) { T u = __shfl_xor_sync(kFull, v, o); v = u < v ? u : v; }
#endif
  return v;
}
template <class T> __device__ __forceinline__ T wsum(T v) {
#if MP_LANES > 1
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
#endif
  return v;
}
// inclusive prefix sum over lanes
__device__ __forceinline__ int wscan_incl(int v) {
#if MP_LANES > 1
  const int l = lane_id();
  for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(kFull, v, (unsigned)o); if (l >= o) v += u; }
#endif
  return v;
}

// ---- per-worker bump arena with stack discipline (mark / release), replaces the function-local std::vectors of the reference
struct Arena {
  unsigned char *base;
  unsigned long long cap, top, peak, cap0;
  int overflow;
  int phase;                              // CTA phase barriers this warp has passed for its current read (mp_phase)
  unsigned phase_mask;                    // bit i: the i-th phase point of a read is a CTA barrier (uniform over the launch; all ones by default)
  unsigned long long *prof;               // optional cycle counters of this warp (LRA_B200_MAP_PROFILE)
  __device__ __forceinline__ void init(void *b, unsigned long long c) { base = (unsigned char *)b; cap = c & ~15ull; cap0 = cap; top = 0; peak = 0; overflow = 0; phase = 0; phase_mask = 0xffffffffu; prof = nullptr; }
  __device__ __forceinline__ unsigned long long tick(int slot, unsigned long long t0) {
#ifdef LRA_EMU
    (void)slot; (void)t0; return 0ull;
#else
    const unsigned long long t1 = (unsigned long long)clock64();
    if (prof && (threadIdx.x & 31u) == 0) prof[slot] += t1 - t0;
    return t1;
#endif
  }
  __device__ __forceinline__ unsigned long long now() const {
#ifdef LRA_EMU
    return 0ull;
#else
    return (unsigned long long)clock64();
#endif
  }
  template <class T> __device__ __forceinline__ T *alloc(unsigned long long n) {
    unsigned long long t = (top + 15ull) & ~15ull;
    unsigned long long e = t + n * sizeof(T);
    if (e > cap) { overflow = 1; return (T *)0; }
    top = e;
    if (e + (cap0 - cap) > peak) peak = e + (cap0 - cap);
    return (T *)(base + t);
  }
  // temporaries from the far end of the arena (released by restoring `cap`): scratch that must not sit under allocations that outlive it
  template <class T> __device__ __forceinline__ T *alloc_hi(unsigned long long n) {
    const unsigned long long bytes = (n * sizeof(T) + 15ull) & ~15ull;
    const unsigned long long t = (top + 15ull) & ~15ull;
    if (bytes > cap || t > cap - bytes) { overflow = 1; return (T *)0; }
    cap -= bytes;
    if (t + (cap0 - cap) > peak) peak = t + (cap0 - cap);
    return (T *)(base + cap);
  }
  __device__ __forceinline__ unsigned long long mark_hi() const { return cap; }
  __device__ __forceinline__ void release_hi(unsigned long long m) { cap = m; }
  __device__ __forceinline__ unsigned long long mark() const { return top; }
  __device__ __forceinline__ void release(unsigned long long m) { top = m; }
  __device__ __forceinline__ unsigned long long avail() const { return cap - ((top + 15ull) & ~15ull); }
};

// ---- phase alignment of the warps of a CTA.  The worker is instruction-fetch bound when its warps run different stages of the (800 KB) path at
// the same time (ncu: 30 of 39 stall cycles per instruction are no_instruction with 16 unaligned warps per SM), so the warps of a CTA take a group
// of reads through the stages in lock step: one CTA barrier closes every stage.  A barrier only counts arrivals, so a warp whose read leaves the
// path early (unaligned, capacity error, fewer chains) catches up with mp_phase_upto and idles at no cost to the others.
__device__ __forceinline__ void mp_phase(Arena &ar) {
#if !defined(LRA_EMU)
  if ((ar.phase_mask >> (ar.phase & 31)) & 1u) __syncthreads();
#endif
  ar.phase++;
}
__device__ __forceinline__ void mp_phase_upto(Arena &ar, int n) { while (ar.phase < n) mp_phase(ar); }
constexpr int kPhasesStage1 = 4;          // minimizers+sort | CompareLists | strand+CleanMatches+LinearExtend | first SparseDP
constexpr int kPhasesChain = 6;           // SPLITChain+Refine_splitchain | Refine_Btwnsplitchain | MergeChain+LinearExtend | second SparseDP | LocalRefineAlignment | output

__device__ __forceinline__ int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

// Warp bitonic sort of a[0..P) in (global / local) memory, P a power of two; the caller pads a[n..P) with a value that
// compares greater-or-equal to everything.  `less` must be a strict weak order; the result order of equal elements is the
// network's, so callers whose ties are observable add the source index to the key.
template <class T, class Less> __device__ inline void wsort_pow2(T *a, int P, Less less) {
  wsync();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int x = lane_id(); x < (P >> 1); x += kLanes) {
        const int i = ((x & ~(j - 1)) << 1) | (x & (j - 1));   // index with bit j clear
        const int p = i | j;
        const bool up = (i & k) == 0;
        const T u = a[i], v = a[p];
        const bool sw = up ? less(v, u) : less(u, v);
        if (sw) { a[i] = v; a[p] = u; }
      }
      wsync();
    }
  }
}

// lower_bound / upper_bound on a plain array
template <class T> __device__ __forceinline__ int lower_bound_idx(const T *a, int n, T v) {
  int first = 0, count = n;
  while (count > 0) { int step = count >> 1; if (a[first + step] < v) { first += step + 1; count -= step + 1; } else count = step; }
  return first;
}
template <class T> __device__ __forceinline__ int upper_bound_idx(const T *a, int n, T v) {
  int first = 0, count = n;
  while (count > 0) { int step = count >> 1; if (!(v < a[first + step])) { first += step + 1; count -= step + 1; } else count = step; }
  return first;
}

}  // namespace mp
}  // namespace lra
