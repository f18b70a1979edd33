// a4 in the mapper worker: CompareLists<GenomeTuple, Tuple> (CompareLists.h:8-151) of the read's sorted minimizers against the global index, split
// into three parts so that the 2 x 28 dependent probes per read minimizer of the literal form (3 Gb reference: 2 * 10^8 index entries, every
// probe a DRAM round trip) leave the sequential part:
//   A (lane-parallel)  for every read minimizer the global lower / upper bound of its masked tuple in the index, and the index tuples around them;
//   B (one lane)       the reference's two-ended walk (front / back decision by the tuple gaps, the unmasked-equality advances, the early exit
//                      that leaves the last list element unmatched) on those bounds: lower_bound over [ts, te) is the global bound clamped to
//                      the range, `q < T[ts]` is `UB[q] <= ts`, `q > T[te-1]` is `LB[q] >= te`; it emits (q range) x (t range) descriptors;
//   C (lane-parallel)  the descriptors expanded into matches in the reference's push order.
// Pinned against the literal device form (seed_kernels.cuh mm_compare) on random lists with repeated tuples and mixed strand bits
// (tests/test_emu_compare.py) and through the end-to-end SAM tests.
#pragma once
#include "mp_common.cuh"
#include "seed_kernels.cuh"

namespace lra {
namespace mp {

struct CmpPlan { int qa, qb; uint32_t ta, tb; };      // pairs (qi, ti): for ti in [ta, tb) for qi in [qa, qb]

// returns the number of descriptors (in `plan`, capacity nq + 2), or -1 when the arena is exhausted
__device__ __noinline__ int mp_compare_plan(const unsigned long long *qt, int nq, const unsigned long long *tt, long long nt, long long maxFreq, Arena &ar, CmpPlan *plan) {
  const int lane = lane_id();
  if (nq == 0 || nt == 0) return 0;
  uint32_t *LB = ar.alloc<uint32_t>(nq), *UB = ar.alloc<uint32_t>(nq);
  unsigned long long *KLB = ar.alloc<unsigned long long>(nq), *KbLB = ar.alloc<unsigned long long>(nq);
  int *np_p = ar.alloc<int>(1);
  if (ar.overflow) return -1;
  constexpr unsigned long long NONE = ~0ull;          // "no such index entry" (masked tuples never have bit 63 set)
  // ---- A
  for (int i = lane; i < nq; i += kLanes) {
    const unsigned long long key = qt[i] & kForMask;
    long long lo = 0, len = nt;
    while (len > 0) { const long long half = len >> 1, mid = lo + half; if ((tt[mid] & kForMask) < key) { lo = mid + 1; len = len - half - 1; } else len = half; }
    const long long lb = lo;
    unsigned long long klb = NONE;
    long long ub = lb;
    if (lb < nt) {
      klb = tt[lb] & kForMask;
      if (klb == key) {
        // runs are at most globalMaxFreq long in an index written by `lra index`: walk a few entries, then search
        int c = 0;
        ub = lb + 1;
        while (ub < nt && c < 8 && (tt[ub] & kForMask) == key) { ub++; c++; }
        if (ub < nt && (tt[ub] & kForMask) == key) {
          long long l2 = ub, n2 = nt - ub;
          while (n2 > 0) { const long long half = n2 >> 1, mid = l2 + half; if (!(key < (tt[mid] & kForMask))) { l2 = mid + 1; n2 = n2 - half - 1; } else n2 = half; }
          ub = l2;
        }
      }
    }
    LB[i] = (uint32_t)lb; UB[i] = (uint32_t)ub; KLB[i] = klb; KbLB[i] = lb > 0 ? (tt[lb - 1] & kForMask) : NONE;
  }
  wsync();
  // ---- B
  if (lane == 0) {
#define QK(i) (qt[i] & kForMask)
    int np = 0;
    long qs = 0, qe = nq - 1;
    long long ts = 0, te = nt;
    unsigned long long tsk = tt[0] & kForMask, tek = tt[nt - 1] & kForMask;     // T[ts], T[te - 1] (masked), kept current
    bool tsk_ok = true, tek_ok = true;
    do {
      while (qs <= qe && (long long)UB[qs] <= ts) qs++;
      if (qs >= qe) break;
      if (!tsk_ok) { tsk = tt[ts] & kForMask; tsk_ok = true; }
      const unsigned long long startGap = QK(qs) - tsk;
      while (qe > qs && te > ts && (long long)LB[qe] >= te) qe--;
      if (!tek_ok) { tek = tt[te - 1] & kForMask; tek_ok = true; }
      const unsigned long long endGap = tek - QK(qe);
      if (startGap == 0 || ((startGap & kForMask) > (endGap & kForMask))) {
        const long long tsOrig = ts; const long qsOrig = qs;
        { long long lb = (long long)LB[qs]; if (lb < ts) lb = ts; if (lb > te) lb = te;
          if (lb != ts) { if (lb == (long long)LB[qs] && KLB[qs] != NONE) { tsk = KLB[qs]; tsk_ok = true; } else tsk_ok = false; }
          ts = lb; }
        if (ts >= (long long)LB[qs] && ts < (long long)UB[qs]) {
          const long long tsStart = ts;
          long long tsi = ts;
          if (ts < te) { tsi = (long long)UB[qs] < te ? (long long)UB[qs] : te; }
          const long qsStart = qs;
          while (qs < qe && QK(qs + 1) == QK(qs)) qs++;
          if (qs - qsStart < maxFreq && tsi > tsStart) { plan[np].qa = (int)qsStart; plan[np].qb = (int)qs; plan[np].ta = (uint32_t)tsStart; plan[np].tb = (uint32_t)tsi; np++; }
        }
        if (ts == tsOrig) {        // (ts > tsOrig: T[ts] and T[tsOrig] differ already in the masked tuple)
          const unsigned long long v = tt[tsOrig];
          while (ts < te && tt[ts] == v) ts++;
          tsk_ok = false;
        }
        while (qs < qe && qt[qs] == qt[qsOrig]) qs++;
      } else {
        const bool pass = te != nt && (long long)LB[qe] < te && te <= (long long)UB[qe];
        if (!pass) {
          long long ub = (long long)UB[qe]; if (ub < ts) ub = ts; if (ub > te) ub = te;
          if (ub != te) { if (ub == (long long)UB[qe]) { tek = UB[qe] > LB[qe] ? QK(qe) : KbLB[qe]; tek_ok = tek != NONE; } else tek_ok = false; }
          te = ub;
        }
        const long long teStart = te;
        long long tei = te;
        if (te > ts && te <= (long long)UB[qe] && te > (long long)LB[qe]) tei = (long long)LB[qe] > ts ? (long long)LB[qe] : ts;
        if (tei < teStart && teStart > 0) {
          const long qeStart = qe;
          while (qe > qs && QK(qe) == QK(qe - 1)) qe--;
          if (qeStart - qe < maxFreq) { plan[np].qa = (int)qe; plan[np].qb = (int)qeStart; plan[np].ta = (uint32_t)tei; plan[np].tb = (uint32_t)teStart; np++; }
        }
        if (tei != te) { if (tei == (long long)LB[qe] && KbLB[qe] != NONE) { tek = KbLB[qe]; tek_ok = true; } else tek_ok = false; }
        te = tei;
      }
    } while (qs < qe && ts < te);
    np_p[0] = np;
#undef QK
  }
  wsync();
  return np_p[0];
}

// ---- C: the pairs of the descriptors, in push order, through `emit(slot, qi, ti)`; returns their number
template <class Emit>
__device__ __forceinline__ long long mp_compare_expand(const CmpPlan *plan, int np, Emit emit) {
  const int lane = lane_id();
  long long total = 0;
  for (int b = 0; b < np; b += kLanes) {
    const int e = b + lane;
    long long cnt = 0;
    CmpPlan p; p.qa = p.qb = 0; p.ta = p.tb = 0;
    if (e < np) { p = plan[e]; cnt = (long long)(p.tb - p.ta) * (long long)(p.qb - p.qa + 1); }
    // exclusive prefix over the lanes (64-bit)
    long long incl = cnt;
#if MP_LANES > 1
    for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(kFull, incl, (unsigned)o); if (lane >= o) incl += u; }
#endif
    long long at = total + incl - cnt;
    if (e < np)
      for (uint32_t ti = p.ta; ti < p.tb; ti++)
        for (int qi = p.qa; qi <= p.qb; qi++) emit(at++, qi, ti);
    total += bcast(incl, kLanes - 1);
  }
  wsync();
  return total;
}

}  // namespace mp
}  // namespace lra
