// a10: the SparseDP family (reference SparseDP.h:1766-2440, SparseDP_Forward.h:312-490; machinery SubRountine.h:103-458,
// DivideSubBy{Row,Col}{1,2}.h, SparseDP.h:140-310 PassValueToD1/2, :1016-1350 ProcessPoint, :1351-1576 TraceBack,
// :1658-1765 DecidePrimaryChains), one problem per warp.
//
// The reference is the Eppstein-Galil-Giancarlo-Italiano sparse DP with a piece-wise-linear gap cost that is *not* concave
// across its ceiling steps, so the answer is defined by the algorithm's own candidate-list state, not by the recurrence: the
// divide-and-conquer (numbering of the sub-problems, the order in which a start point visits them, the lazy Maximization
// state now/last/S_1/Block, strict `<` updates) is reproduced literally.  What is re-designed is the mapping:
//   * sub-problem storage is flat (one arena allocation per array, no per-fragment copies of the SS lists: a point uses
//     the lists of its own row / column);
//   * the D&C scan, the diagonal sort + unique, Db/Eb and the SS-list pushes are lane-parallel;
//   * in ProcessPoint the sub-problems a start point visits (<= 2 log2 N, all distinct) are evaluated one per lane, then
//     an ordered warp reduction applies the reference's "first strictly greater wins" rule; an end point updates its
//     sub-problems one per lane.
// All float arithmetic uses explicitly rounded binary32 operations in the reference's association order.
#pragma once
#include "mp_common.cuh"
#include "introsort.cuh"

namespace lra {
namespace mp {

// host-built InitPWL tables (SubRountine.h:43-101), uploaded once
struct Pwl {
  long long stops[25];
  float slope[25], inter[25];
  int ceil1, ceil2;
};

struct SdpPt { uint32_t q, t, frag; int32_t cl; uint32_t fl; uint32_t src; };   // fl: bit0 ind (start), bit1 inv (forward family), bit2 orient
struct SdpSub {
  int m, n, now, last, nB, nS, capB, capS;
  long long *Di, *Ei;
  float *Dv, *Ev;
  uint32_t *Dp, *Ep;
  int *Db, *Eb;
  int2 *Bk, *S;
};
struct SdpFam {
  SdpSub *sub; int nsub, cap;
  uint32_t *ssA, *ssB; int *nA, *nB; int stride;    // per row / column SS lists
  int inv, cols, desc, swp;
};
// growth heap for the Block / S_1 lists: the part of the worker arena above the set-up allocations, bumped atomically because
// lanes grow the lists of different sub-problems concurrently.  (The reference re-runs Maximization over entries it has
// already seen whenever `now` moves backwards, so the lists have no bound in terms of m.)
struct SdpDyn { unsigned long long *top; unsigned char *base; unsigned long long cap; int *err; unsigned long long base_off; };
__device__ __noinline__ bool sdp_grow(const SdpDyn &D, int2 *&arr, int &cap, int used) {
  const int ncap = cap * 2 + 16;
  const unsigned long long bytes = ((unsigned long long)ncap * sizeof(int2) + 15ull) & ~15ull;
  const unsigned long long off = atomicAdd(D.top, bytes);
  if (off + bytes > D.cap) { atomicExch(D.err, 1); return false; }
  int2 *na = (int2 *)(D.base + off);
  for (int i = 0; i < used; i++) na[i] = arr[i];
  arr = na; cap = ncap;
  return true;
}
struct SdpVal { float val; int cl; int prev_sub, prev_ind; int prev, inv, orient; };

struct SdpWork {
  SdpPt *H1; uint32_t *H2; int N;
  int *rowS, *rowE, *colS, *colE; int R, C;     // pstart / pend per row (H1 positions) and column (H2 positions)
  int *rowOf, *colOfPos, *colOfPt;              // row of H1 position; column of H2 position; column of H1 position
  SdpFam fam[4];                                // R1, C1, R2, C2
  SdpVal *val; int nfrag;
  long long *tmp; uint8_t *flag;                // scan scratch
  SdpDyn dyn;
};

__device__ __forceinline__ float sdp_w(long long i, long long j, const Pwl &P) {
  long long d = j - i; if (d < 0) d = -d;
  const long long x = d + 1;
  if (x == 1) return 0.0f;
  long long penalty;
  if (x <= 2) penalty = 0;
  else {
    int bound = 0;   // upper_bound over STOPS[0..24)
    { int first = 0, count = 24; while (count > 0) { int step = count >> 1; if (!(x < P.stops[first + step])) { first += step + 1; count -= step + 1; } else count = step; } bound = first; }
    const float f = __fadd_rn(__fmul_rn(P.slope[bound - 1], __ll2float_rn(x)), P.inter[bound - 1]);
    penalty = (long long)f;
    if (penalty >= P.ceil1 && penalty < P.ceil2) penalty = P.ceil1;
    else if (penalty > P.ceil2) penalty = P.ceil2;
  }
  return -__ll2float_rn(penalty);
}

__device__ __forceinline__ long long sdp_diag(const SdpPt &p, int inv) {
  return inv ? (long long)p.t - (long long)p.q : (long long)p.t + (long long)p.q;
}

// Lower_Bound (Sorting.h:303) over an ascending array (forward iterators) or a descending one (reverse iterators);
// returns the element index the iterator points at, or -1 for end / rend.
__device__ __forceinline__ int sdp_lb(const long long *a, int n, long long v, int desc) {
  int first = 0, count = n;
  if (!desc) {
    while (count > 0) { int step = count >> 1; if (a[first + step] < v) { first += step + 1; count -= step + 1; } else count = step; }
    return first < n ? first : -1;
  }
  while (count > 0) { int step = count >> 1; if (a[n - 1 - (first + step)] < v) { first += step + 1; count -= step + 1; } else count = step; }
  return first < n ? n - 1 - first : -1;
}

// FindBoundary (SubRountine.h:334-355)
__device__ __forceinline__ int sdp_find_boundary(int first, int last, int a, int b, const SdpSub &s, const Pwl &P) {
  if (b != -1) {
    unsigned count = (unsigned)(last - first);
    while (count > 0) {
      unsigned step = count / 2; int it = first + (int)step;
      if (__fadd_rn(s.Dv[a], sdp_w(s.Di[a], s.Ei[it], P)) > __fadd_rn(s.Dv[b], sdp_w(s.Di[b], s.Ei[it], P))) { first = it + 1; count -= step + 1; }
      else count = step;
    }
    return first;
  }
  return s.n;
}

// Maximization (SubRountine.h:357-458): advances the candidate list of one sub-problem from `last` to `now`
#define SDP_PUSH_B(v) do { if (nB >= s.capB && !sdp_grow(D, s.Bk, s.capB, nB)) { s.nS = nS; s.nB = nB; return false; } s.Bk[nB++] = (v); } while (0)
#define SDP_PUSH_S(v) do { if (nS >= s.capS && !sdp_grow(D, s.S, s.capS, nS)) { s.nS = nS; s.nB = nB; return false; } s.S[nS++] = (v); } while (0)
__device__ __noinline__ bool sdp_maximization(SdpSub &s, const Pwl &P, const SdpDyn &D) {
  const int n = s.n, m = s.m;
  int nS = s.nS, nB = s.nB;
  for (unsigned i = (unsigned)(s.last + 1); i <= (unsigned)s.now; ++i) {
    MP_CHECK((int)i < m);
    const int dbi = s.Db[i];
    if (dbi == -1) break;
    MP_CHECK(dbi >= 0 && dbi < n && nS >= 1);
    if (s.S[nS - 1].y == n + 1) { SDP_PUSH_B(make_int2(-1, dbi)); SDP_PUSH_S(make_int2((int)i, n)); }
    while (dbi >= s.S[nS - 1].y) { SDP_PUSH_B(s.S[nS - 1]); nS--; }
    MP_CHECK(nS >= 1);
    const int l = s.S[nS - 1].x;
    MP_CHECK(l >= 0 && l < m);
    const long long e = s.Ei[dbi];
    if (__fadd_rn(s.Dv[i], sdp_w(s.Di[i], e, P)) > __fadd_rn(s.Dv[l], sdp_w(s.Di[l], e, P))) {
      if (dbi < s.S[nS - 1].y && nB > 0 && dbi > s.Bk[nB - 1].y) SDP_PUSH_B(make_int2(s.S[nS - 1].x, dbi));
      int2 cur = s.S[nS - 1], prev = cur;
      while (nS > 0 && __fadd_rn(s.Dv[i], sdp_w(s.Di[i], s.Ei[cur.y - 1], P)) > __fadd_rn(s.Dv[cur.x], sdp_w(s.Di[cur.x], s.Ei[cur.y - 1], P))) {
        nS--; prev = cur; MP_CHECK(nS >= 1); cur = s.S[nS - 1];
        if (cur.y == n + 1) break;
        MP_CHECK(cur.x >= 0 && cur.x < m && cur.y >= 1 && cur.y <= n);
      }
      const int h = sdp_find_boundary(prev.y, cur.y, (int)i, cur.x, s, P);
      SDP_PUSH_S(make_int2((int)i, h));
    }
  }
  if (s.now == m - 1) { while (s.S[nS - 1].y != n + 1) { SDP_PUSH_B(s.S[nS - 1]); nS--; } }
  else { while (s.Db[s.now + 1] >= s.S[nS - 1].y) { SDP_PUSH_B(s.S[nS - 1]); nS--; } }
  s.nS = nS; s.nB = nB;
  s.last = s.now;
  return true;
}
#undef SDP_PUSH_B
#undef SDP_PUSH_S

// one start point against one sub-problem (the body of the k loops of ProcessPoint, SparseDP.h:1029-1063); returns false
// when the sub-problem is skipped, else Ev and the Ei index
__device__ __noinline__ bool sdp_eval_start(SdpSub &s, long long diag, int desc, float bonus, const Pwl &P, const SdpDyn &D, float &ev, int &i1out) {
  if (s.m == 0 || s.n == 0) return false;      // Di.empty(); a kept one-sided node with Ei empty is never in a B list
  const int t = sdp_lb(s.Ei, s.n, diag, desc);
  if (t < 0) return false;
  if (s.Eb[t] == -1) return false;
  s.now = s.Eb[t];
  if (!sdp_maximization(s, P, D)) return false;
  s.last = s.Eb[t];
  const int i1 = t; int i2;
  // FindValueInBlock (SubRountine.h:317-330)
  if (s.nB > 0 && i1 >= s.Bk[s.nB - 1].y && i1 < s.S[s.nS - 1].y) i2 = s.S[s.nS - 1].x;
  else {
    int first = 0, count = s.nB;
    while (count > 0) { int step = count >> 1; if (i1 >= s.Bk[first + step].y) { first += step + 1; count -= step + 1; } else count = step; }
    MP_CHECK(first < s.nB);
    i2 = s.Bk[first].x;
  }
  MP_CHECK(i2 >= 0 && i2 < s.m);
  ev = __fadd_rn(__fadd_rn(s.Dv[i2], sdp_w(s.Di[i2], s.Ei[i1], P)), bonus);
  s.Ev[i1] = ev; s.Ep[i1] = (uint32_t)i2;
  i1out = i1;
  return true;
}

// ---- divide and conquer ------------------------------------------------------------------------------------------
__device__ __forceinline__ const SdpPt &sdp_pt(const SdpWork &W, const SdpFam &F, int pos) { return F.cols ? W.H1[W.H2[pos]] : W.H1[pos]; }

// collect the diagonals of the points of rows/cols [s,e) with ind == DE of this family; flag rows that have one
__device__ __noinline__ int sdp_scan(SdpWork &W, const SdpFam &F, int s, int e, int DE, long long *out) {
  const int *vs = F.cols ? W.colS : W.rowS, *ve = F.cols ? W.colE : W.rowE;
  const int *rof = F.cols ? W.colOfPos : W.rowOf;
  const int p0 = vs[s], p1 = ve[e - 1];
  int cnt = 0;
  for (int b = p0; b < p1; b += kLanes) {
    const int p = b + lane_id();
    bool take = false; long long d = 0;
    if (p < p1) {
      const SdpPt &pt = sdp_pt(W, F, p);
      take = ((int)(pt.fl & 1u) == DE) && ((int)((pt.fl >> 1) & 1u) == F.inv);
      if (take) { d = sdp_diag(pt, F.inv); W.flag[rof[p]] = 1; }
    }
    const unsigned mk = ballot(take);
    if (take) out[cnt + __popc(mk & lanemask_lt())] = d;
    cnt += __popc(mk);
  }
  wsync();
  return cnt;
}
// sort + unique `a[0..cnt)` (scratch, capacity next_pow2(cnt)), write the distinct values to dst in ascending / descending order
__device__ __noinline__ int sdp_sort_unique(long long *a, int cnt, long long *dst, int desc) {
  if (cnt == 0) return 0;
  const int P = next_pow2(cnt);
  for (int i = cnt + lane_id(); i < P; i += kLanes) a[i] = 0x7fffffffffffffffll;
  wsort_pow2(a, P, [](long long x, long long y) { return x < y; });
  int u = 0;
  for (int b = 0; b < cnt; b += kLanes) {
    const int i = b + lane_id();
    const bool keep = i < cnt && (i == 0 || a[i] != a[i - 1]);
    const unsigned mk = ballot(keep);
    if (keep) dst[u + __popc(mk & lanemask_lt())] = a[i];
    u += __popc(mk);
  }
  wsync();
  if (desc) {
    for (int i = lane_id(); i < u / 2; i += kLanes) { const long long x = dst[i]; dst[i] = dst[u - 1 - i]; dst[u - 1 - i] = x; }
    wsync();
  }
  return u;
}
// push node number n into the A or B list of every flagged row of [s,e); clears the flags
__device__ __noinline__ void sdp_push_ss(SdpWork &W, SdpFam &F, int s, int e, int n, bool toA, bool toB) {
  for (int r = s + lane_id(); r < e; r += kLanes) {
    if (W.flag[r]) {
      if (toA) { F.ssA[(long long)r * F.stride + F.nA[r]] = (uint32_t)n; F.nA[r]++; }
      if (toB) { F.ssB[(long long)r * F.stride + F.nB[r]] = (uint32_t)n; F.nB[r]++; }
      W.flag[r] = 0;
    }
  }
  wsync();
}
__device__ inline void sdp_clear_flags(SdpWork &W, int s, int e) {
  for (int r = s + lane_id(); r < e; r += kLanes) W.flag[r] = 0;
  wsync();
}

// allocate and initialise the arrays of a kept two-sided node (Decide_Eb_Db_*, DivideSubBy*.h)
__device__ __noinline__ bool sdp_setup_sub(SdpSub &s, int desc, Arena &ar) {
  const int m = s.m, n = s.n;
  s.Dv = ar.alloc<float>(m); s.Dp = ar.alloc<uint32_t>(m); s.Db = ar.alloc<int>(m);
  s.Ev = ar.alloc<float>(n); s.Ep = ar.alloc<uint32_t>(n); s.Eb = ar.alloc<int>(n);
  s.capB = 2 * m + 4; s.capS = m + 3;
  s.Bk = ar.alloc<int2>((unsigned long long)s.capB); s.S = ar.alloc<int2>((unsigned long long)s.capS);
  if (ar.overflow) return false;
  for (int i = lane_id(); i < m; i += kLanes) { s.Dv[i] = 0.0f; s.Dp[i] = 0; }
  for (int i = lane_id(); i < n; i += kLanes) { s.Ev[i] = 0.0f; s.Ep[i] = 0; s.Eb[i] = -1; }
  wsync();
  for (int i = lane_id(); i < m; i += kLanes) {
    int db = -1;
    if (!desc) { const int t = lower_bound_idx(s.Ei, n, s.Di[i]); if (t < n) db = t; }
    else {
      // reverse-iterator lower bound, then one step back: the first (descending) element strictly below Di[i]
      int first = 0, count = n;
      while (count > 0) { int step = count >> 1; if (s.Ei[n - 1 - (first + step)] < s.Di[i]) { first += step + 1; count -= step + 1; } else count = step; }
      if (first != 0) db = n - first;
    }
    s.Db[i] = db;
    if (db >= 0) atomicMax(&s.Eb[db], i);     // Eb[*t] = s for increasing s: the last writer is the largest
  }
  wsync();
  // forward fill (values are non-decreasing along the index, so the fill is a running maximum)
  int carry = -1;
  for (int b = 0; b < n; b += kLanes) {
    const int i = b + lane_id();
    int v = i < n ? s.Eb[i] : -1;
#if MP_LANES > 1
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFull, v, (unsigned)o); if (lane_id() >= o && u > v) v = u; }
#endif
    if (carry > v) v = carry;
    if (i < n) s.Eb[i] = v;
    carry = bcast(v, kLanes - 1);
  }
  wsync();
  s.now = 0; s.last = -1; s.nB = 0; s.nS = 1;
  if (lane_id() == 0) s.S[0] = make_int2(-1, n + 1);
  wsync();
  return true;
}

// DivideSubProbBy{Row1,Col1,Row2,Col2}: pre-order numbering with dropped nodes, explicit stack.
// The reference collects and sorts the diagonals of a node's points at every node (O(N log^2 N) per level with a sorting network).  Here the D and
// the E points of the family are sorted by diagonal ONCE; a node owns the segment of each list that holds the points of its rows, still sorted, and
// hands its two halves to its children with a stable partition by row (ping-pong between two buffers by depth).  A node's Di / Ei are the distinct
// diagonals of the segment of its D half / E half.  Numbering, kept / dropped nodes, the SS lists and the sub-problem set-up are as before.
struct SdpDE { long long d; int row; int pad; };
// distinct diagonals of the sorted segment a[0..cnt) -> dst (ascending, or descending when desc); returns their number
__device__ __noinline__ int sdp_unique_sorted(const SdpDE *a, int cnt, long long *dst, int desc) {
  int u = 0;
  for (int b = 0; b < cnt; b += kLanes) {
    const int i = b + lane_id();
    const bool keep = i < cnt && (i == 0 || a[i].d != a[i - 1].d);
    const unsigned mk = ballot(keep);
    if (keep) dst[u + __popc(mk & lanemask_lt())] = a[i].d;
    u += __popc(mk);
  }
  wsync();
  if (desc) {
    for (int i = lane_id(); i < u / 2; i += kLanes) { const long long x = dst[i]; dst[i] = dst[u - 1 - i]; dst[u - 1 - i] = x; }
    wsync();
  }
  return u;
}
// stable partition of src[lo, hi) by row < med into dst[lo, hi); returns the size of the left part
__device__ __noinline__ int sdp_partition(const SdpDE *src, SdpDE *dst, int lo, int hi, int med) {
  int nl = 0;
  for (int b = lo; b < hi; b += kLanes) { const int i = b + lane_id(); nl += __popc(ballot(i < hi && src[i].row < med)); }
  int cl = 0, cr = 0;
  for (int b = lo; b < hi; b += kLanes) {
    const int i = b + lane_id();
    const bool in = i < hi;
    SdpDE x; x.d = 0; x.row = 0; x.pad = 0;
    if (in) x = src[i];
    const bool left = in && x.row < med, right = in && !left;
    const unsigned ml = ballot(left), mr = ballot(right);
    if (left) dst[lo + cl + __popc(ml & lanemask_lt())] = x;
    if (right) dst[lo + nl + cr + __popc(mr & lanemask_lt())] = x;
    cl += __popc(ml); cr += __popc(mr);
  }
  wsync();
  return nl;
}
// push node number n into the A or B list of every row of [s,e) that has a D (A) / an E (B) point of the family
__device__ __noinline__ void sdp_push_ss2(SdpFam &F, const uint8_t *has, int s, int e, int n, bool toA, bool toB) {
  for (int r = s + lane_id(); r < e; r += kLanes) {
    if (toA && (has[r] & 1)) { F.ssA[(long long)r * F.stride + F.nA[r]] = (uint32_t)n; F.nA[r]++; }
    if (toB && (has[r] & 2)) { F.ssB[(long long)r * F.stride + F.nB[r]] = (uint32_t)n; F.nB[r]++; }
  }
  wsync();
}

// ---- small subtrees of the divide-and-conquer tree, one per LANE ------------------------------------------------------------------------------
// Most nodes of the tree are small (a tree over V rows has V leaves), and a warp-wide pass over a node of a few points costs the same few hundred
// instructions as over a node of a thousand.  Subtrees of at most kSdpSmallRows rows are therefore taken out of the warp's traversal: their number of
// kept nodes depends only on the point counts of their rows (computed for all of them, one per lane, from prefix counts), so the warp's traversal
// can skip over them and keep numbering; afterwards every lane builds one small subtree serially (same partition / unique / set-up, plain loops),
// in its own slice of the two list buffers and of a private heap.
constexpr int kSdpSmallRows = 16;
struct SdpLaneHeap {
  unsigned char *p, *end; int ovf;
  template <class T> __device__ __forceinline__ T *alloc(int n) {
    unsigned char *q = (unsigned char *)(((unsigned long long)p + 15ull) & ~15ull);
    unsigned char *e = q + (unsigned long long)(n > 0 ? n : 0) * sizeof(T);
    if (e > end) { ovf = 1; return (T *)0; }
    p = e;
    return (T *)q;
  }
};
struct SdpSmallJob { int start, end, dlo, dhi, elo, ehi, dep, base; unsigned long long heap_off, heap_len; };

// number of kept nodes of the subtree over rows [s, e) (the subtree root is assumed to be visited)
__device__ __noinline__ int sdp_subtree_size(const int *PD, const int *PE, int s0, int e0, int swp) {
  int stS[24], stE[24], sp = 0, total = 0;
  stS[0] = s0; stE[0] = e0; sp = 1;
  while (sp > 0) {
    --sp;
    const int s = stS[sp], e = stE[sp];
    if (e == s + 1) { if (PD[e] - PD[s] != 0 && PE[e] - PE[s] != 0) total++; continue; }
    const int med = (s + e) / 2;
    const int dS = swp ? med : s, dE = swp ? e : med, eS = swp ? s : med, eE = swp ? med : e;
    const int cd = PD[dE] - PD[dS], ce = PE[eE] - PE[eS];
    if (cd == 0 && ce == 0) continue;
    total++;
    if (ce) { stS[sp] = eS; stE[sp] = eE; sp++; }
    if (cd) { stS[sp] = dS; stE[sp] = dE; sp++; }
  }
  return total;
}

__device__ __forceinline__ int sdp_unique_serial(const SdpDE *a, int cnt, long long *dst, int desc) {
  int u = 0;
  for (int i = 0; i < cnt; i++) if (i == 0 || a[i].d != a[i - 1].d) dst[u++] = a[i].d;
  if (desc) for (int i = 0; i < u / 2; i++) { const long long x = dst[i]; dst[i] = dst[u - 1 - i]; dst[u - 1 - i] = x; }
  return u;
}
__device__ __forceinline__ int sdp_partition_serial(const SdpDE *src, SdpDE *dst, int lo, int hi, int med) {
  int nl = 0;
  for (int i = lo; i < hi; i++) nl += src[i].row < med ? 1 : 0;
  int cl = lo, cr = lo + nl;
  for (int i = lo; i < hi; i++) { if (src[i].row < med) dst[cl++] = src[i]; else dst[cr++] = src[i]; }
  return nl;
}
__device__ __forceinline__ void sdp_push_ss_serial(SdpFam &F, const uint8_t *has, int s, int e, int n, bool toA, bool toB) {
  for (int r = s; r < e; r++) {
    if (toA && (has[r] & 1)) { F.ssA[(long long)r * F.stride + F.nA[r]] = (uint32_t)n; F.nA[r]++; }
    if (toB && (has[r] & 2)) { F.ssB[(long long)r * F.stride + F.nB[r]] = (uint32_t)n; F.nB[r]++; }
  }
}
__device__ __forceinline__ bool sdp_setup_sub_serial(SdpSub &s, int desc, SdpLaneHeap &H) {
  const int m = s.m, n = s.n;
  s.Dv = H.alloc<float>(m); s.Dp = H.alloc<uint32_t>(m); s.Db = H.alloc<int>(m);
  s.Ev = H.alloc<float>(n); s.Ep = H.alloc<uint32_t>(n); s.Eb = H.alloc<int>(n);
  s.capB = 2 * m + 4; s.capS = m + 3;
  s.Bk = H.alloc<int2>(s.capB); s.S = H.alloc<int2>(s.capS);
  if (H.ovf) return false;
  for (int i = 0; i < m; i++) { s.Dv[i] = 0.0f; s.Dp[i] = 0; }
  for (int i = 0; i < n; i++) { s.Ev[i] = 0.0f; s.Ep[i] = 0; s.Eb[i] = -1; }
  for (int i = 0; i < m; i++) {
    int db = -1;
    if (!desc) { const int t = lower_bound_idx(s.Ei, n, s.Di[i]); if (t < n) db = t; }
    else {
      int first = 0, count = n;
      while (count > 0) { int step = count >> 1; if (s.Ei[n - 1 - (first + step)] < s.Di[i]) { first += step + 1; count -= step + 1; } else count = step; }
      if (first != 0) db = n - first;
    }
    s.Db[i] = db;
    if (db >= 0 && s.Eb[db] < i) s.Eb[db] = i;
  }
  int carry = -1;
  for (int i = 0; i < n; i++) { if (s.Eb[i] < carry) s.Eb[i] = carry; else carry = s.Eb[i]; }
  s.now = 0; s.last = -1; s.nB = 0; s.nS = 1;
  s.S[0] = make_int2(-1, n + 1);
  return true;
}
// one small subtree, serially (called by one lane per subtree)
__device__ __noinline__ bool sdp_subtree_lane(SdpFam &F, const uint8_t *has, SdpDE *const *bufD, SdpDE *const *bufE, const SdpSmallJob &J, SdpLaneHeap &H) {
  int stS[24], stE[24], stDl[24], stDh[24], stEl[24], stEh[24], stDep[24]; int sp = 0;
  stS[0] = J.start; stE[0] = J.end; stDl[0] = J.dlo; stDh[0] = J.dhi; stEl[0] = J.elo; stEh[0] = J.ehi; stDep[0] = J.dep; sp = 1;
  int num = J.base;
  while (sp > 0) {
    --sp;
    const int start = stS[sp], end = stE[sp], dlo = stDl[sp], dhi = stDh[sp], elo = stEl[sp], ehi = stEh[sp], dep = stDep[sp];
    SdpSub s; s.m = s.n = 0; s.now = 0; s.last = -1; s.nB = 0; s.nS = 0; s.capB = s.capS = 0;
    s.Di = s.Ei = 0; s.Dv = s.Ev = 0; s.Dp = s.Ep = 0; s.Db = s.Eb = 0; s.Bk = s.S = 0;
    const SdpDE *curD = bufD[dep & 1], *curE = bufE[dep & 1];
    if (end == start + 1) {
      const int cE = ehi - elo, cD = dhi - dlo;
      if (cE != 0 && cD != 0) {
        s.Ei = H.alloc<long long>(cE); s.Di = H.alloc<long long>(cD);
        if (H.ovf) return false;
        s.n = sdp_unique_serial(curE + elo, cE, s.Ei, F.desc);
        s.m = sdp_unique_serial(curD + dlo, cD, s.Di, F.desc);
        sdp_push_ss_serial(F, has, start, end, num, true, true);
        if (!sdp_setup_sub_serial(s, F.desc, H)) return false;
        F.sub[num] = s;
        num++;
      }
      continue;
    }
    const int med = (start + end) / 2;
    const int dS = F.swp ? med : start, dE = F.swp ? end : med;
    const int eS = F.swp ? start : med, eE = F.swp ? med : end;
    SdpDE *nxtD = bufD[(dep + 1) & 1], *nxtE = bufE[(dep + 1) & 1];
    const int dl = sdp_partition_serial(curD, nxtD, dlo, dhi, med);
    const int el = sdp_partition_serial(curE, nxtE, elo, ehi, med);
    const int ldl = dlo, ldh = dlo + dl, rdl = dlo + dl, rdh = dhi;
    const int lel = elo, leh = elo + el, rel = elo + el, reh = ehi;
    const int Ddl = F.swp ? rdl : ldl, Ddh = F.swp ? rdh : ldh;
    const int Eel = F.swp ? lel : rel, Eeh = F.swp ? leh : reh;
    const int cD = Ddh - Ddl, cE = Eeh - Eel;
    if (cD) { s.Di = H.alloc<long long>(cD); if (H.ovf) return false; s.m = sdp_unique_serial(nxtD + Ddl, cD, s.Di, F.desc); }
    if (cE) { s.Ei = H.alloc<long long>(cE); if (H.ovf) return false; s.n = sdp_unique_serial(nxtE + Eel, cE, s.Ei, F.desc); }
    if (s.n == 0 && s.m == 0) continue;
    sdp_push_ss_serial(F, has, dS, dE, num, true, false);
    sdp_push_ss_serial(F, has, eS, eE, num, false, true);
    if (s.n != 0 && s.m != 0) { if (!sdp_setup_sub_serial(s, F.desc, H)) return false; }
    F.sub[num] = s;
    num++;
    const bool goD = s.m != 0, goE = s.n != 0;
    const int Dhalf_el = F.swp ? rel : lel, Dhalf_eh = F.swp ? reh : leh;
    const int Ehalf_dl = F.swp ? ldl : rdl, Ehalf_dh = F.swp ? ldh : rdh;
    if (goE) { stS[sp] = eS; stE[sp] = eE; stDl[sp] = Ehalf_dl; stDh[sp] = Ehalf_dh; stEl[sp] = Eel; stEh[sp] = Eeh; stDep[sp] = dep + 1; sp++; }
    if (goD) { stS[sp] = dS; stE[sp] = dE; stDl[sp] = Ddl; stDh[sp] = Ddh; stEl[sp] = Dhalf_el; stEh[sp] = Dhalf_eh; stDep[sp] = dep + 1; sp++; }
  }
  return true;
}

__device__ __noinline__ bool sdp_divide_impl(SdpWork &W, SdpFam &F, Arena &ar) {
  const int V = F.cols ? W.C : W.R;
  F.nsub = 0;
  if (V == 0) return true;
  // ---- the family's D and E points (diagonal, row), sorted by diagonal; has[r]: bit 0 row r has a D point, bit 1 an E point
  const int *rof = F.cols ? W.colOfPos : W.rowOf;
  const int N = W.N, P2 = next_pow2(N > 0 ? N : 1);
  SdpDE *bufD[2], *bufE[2];
  bufD[0] = ar.alloc_hi<SdpDE>(P2); bufD[1] = ar.alloc_hi<SdpDE>(N > 0 ? N : 1); bufE[0] = ar.alloc_hi<SdpDE>(P2); bufE[1] = ar.alloc_hi<SdpDE>(N > 0 ? N : 1);
  uint8_t *has = ar.alloc_hi<uint8_t>(V);
  if (ar.overflow) return false;
  for (int r = lane_id(); r < V; r += kLanes) has[r] = 0;
  wsync();
  int nD = 0, nE = 0;
  for (int b = 0; b < N; b += kLanes) {
    const int p = b + lane_id();
    bool tD = false, tE = false; SdpDE x; x.d = 0; x.row = 0; x.pad = 0;
    if (p < N) {
      const SdpPt &pt = sdp_pt(W, F, p);
      if ((int)((pt.fl >> 1) & 1u) == F.inv) { tD = (pt.fl & 1u) == 0u; tE = !tD; x.d = sdp_diag(pt, F.inv); x.row = rof[p]; }
    }
    const unsigned mD = ballot(tD), mE = ballot(tE);
    if (tD) bufD[0][nD + __popc(mD & lanemask_lt())] = x;
    if (tE) bufE[0][nE + __popc(mE & lanemask_lt())] = x;
    nD += __popc(mD); nE += __popc(mE);
  }
  wsync();
  // (rows are contiguous in family order, so the flag writes of one pass never collide on a byte with different values: a row's D and E bits are
  //  set by separate passes)
  for (int i = lane_id(); i < nD; i += kLanes) has[bufD[0][i].row] |= 1;
  wsync();
  for (int i = lane_id(); i < nE; i += kLanes) has[bufE[0][i].row] |= 2;
  wsync();
  if (nD == 0 && nE == 0) return true;                 // the root is dropped: no sub-problems
  { const int PD = next_pow2(nD > 0 ? nD : 1), PE = next_pow2(nE > 0 ? nE : 1);
    for (int i = nD + lane_id(); i < PD; i += kLanes) { bufD[0][i].d = 0x7fffffffffffffffll; bufD[0][i].row = 0x7fffffff; }
    for (int i = nE + lane_id(); i < PE; i += kLanes) { bufE[0][i].d = 0x7fffffffffffffffll; bufE[0][i].row = 0x7fffffff; }
    wsync();
    auto less = [](const SdpDE &x, const SdpDE &y) { return x.d < y.d; };
    if (nD > 1) wsort_pow2(bufD[0], PD, less);
    if (nE > 1) wsort_pow2(bufE[0], PE, less); }
  // ---- prefix counts of the D / E points over the rows, the small subtrees of the (fixed, midpoint-split) tree and their numbers of kept nodes
  int *PD = ar.alloc_hi<int>(V + 1), *PE = ar.alloc_hi<int>(V + 1);
  const int max_slots = 2 * (V / kSdpSmallRows) + 4;
  int *slotS = ar.alloc_hi<int>(max_slots), *slotE = ar.alloc_hi<int>(max_slots), *slotSize = ar.alloc_hi<int>(max_slots), *slotOf = ar.alloc_hi<int>(V + 1);
  SdpSmallJob *jobs = ar.alloc_hi<SdpSmallJob>(max_slots);
  int *nslot_p = ar.alloc_hi<int>(4);
  if (ar.overflow) return false;
  for (int r = lane_id(); r <= V; r += kLanes) { PD[r] = 0; PE[r] = 0; }
  wsync();
  for (int i = lane_id(); i < nD; i += kLanes) atomicAdd(&PD[bufD[0][i].row + 1], 1);
  for (int i = lane_id(); i < nE; i += kLanes) atomicAdd(&PE[bufE[0][i].row + 1], 1);
  wsync();
  { int cd = 0, ce = 0;          // inclusive scans in place: P[r] = points in rows < r
    for (int b = 0; b <= V; b += kLanes) {
      const int r = b + lane_id();
      const int vd = r <= V ? PD[r] : 0, ve = r <= V ? PE[r] : 0;
      const int id = wscan_incl(vd), ie = wscan_incl(ve);
      if (r <= V) { PD[r] = cd + id; PE[r] = ce + ie; }
      cd += bcast(id, kLanes - 1); ce += bcast(ie, kLanes - 1);
    } }
  wsync();
  if (lane_id() == 0) {
    int ss[72], se[72], sp0 = 0, ns = 0;
    ss[0] = 0; se[0] = V; sp0 = 1;
    while (sp0 > 0) {
      --sp0;
      const int a0 = ss[sp0], e0 = se[sp0];
      if (e0 - a0 <= kSdpSmallRows) { slotS[ns] = a0; slotE[ns] = e0; slotOf[a0] = ns; ns++; continue; }
      const int med = (a0 + e0) / 2;
      ss[sp0] = a0; se[sp0] = med; sp0++;
      ss[sp0] = med; se[sp0] = e0; sp0++;
    }
    nslot_p[0] = ns;
  }
  wsync();
  const int nslots = nslot_p[0];
  for (int k = lane_id(); k < nslots; k += kLanes) slotSize[k] = sdp_subtree_size(PD, PE, slotS[k], slotE[k], F.swp);
  wsync();
  int njobs = 0;
  unsigned long long heap_total = 0;
  int stS[72], stE[72], stDl[72], stDh[72], stEl[72], stEh[72], stDep[72]; int sp = 0;
  stS[0] = 0; stE[0] = V; stDl[0] = 0; stDh[0] = nD; stEl[0] = 0; stEh[0] = nE; stDep[0] = 0; sp = 1;
  while (sp > 0) {
    --sp;
    const int start = stS[sp], end = stE[sp], dlo = stDl[sp], dhi = stDh[sp], elo = stEl[sp], ehi = stEh[sp], dep = stDep[sp];
    if (end - start <= kSdpSmallRows) {
      const int k = slotOf[start], sz = slotSize[k];
      if (F.nsub + sz > F.cap) { ar.overflow = 1; return false; }
      if (sz > 0) {
        const unsigned long long pD = (unsigned long long)(dhi - dlo), pE = (unsigned long long)(ehi - elo);
        int lv = 1; while ((1 << (lv - 1)) < end - start) lv++;             // levels of the subtree
        const unsigned long long need = (unsigned long long)lv * (44ull * pD + 20ull * pE) + 416ull * (unsigned long long)(end - start) + 256ull;
        if (lane_id() == 0) {
          SdpSmallJob J; J.start = start; J.end = end; J.dlo = dlo; J.dhi = dhi; J.elo = elo; J.ehi = ehi; J.dep = dep; J.base = F.nsub; J.heap_off = heap_total; J.heap_len = need;
          jobs[njobs] = J;
        }
        njobs++; heap_total += need;
        F.nsub += sz;
      }
      continue;
    }
    if (F.nsub >= F.cap) { ar.overflow = 1; return false; }
    SdpSub s; s.m = s.n = 0; s.now = 0; s.last = -1; s.nB = 0; s.nS = 0; s.capB = s.capS = 0;
    s.Di = s.Ei = 0; s.Dv = s.Ev = 0; s.Dp = s.Ep = 0; s.Db = s.Eb = 0; s.Bk = s.S = 0;
    const int n = F.nsub;
    const SdpDE *curD = bufD[dep & 1], *curE = bufE[dep & 1];
    if (end == start + 1) {
      // leaf: starts and ends of one row
      const int cE = ehi - elo, cD = dhi - dlo;
      if (cE != 0 && cD != 0) {
        s.Ei = ar.alloc<long long>(cE); s.Di = ar.alloc<long long>(cD);
        if (ar.overflow) return false;
        s.n = sdp_unique_sorted(curE + elo, cE, s.Ei, F.desc);
        s.m = sdp_unique_sorted(curD + dlo, cD, s.Di, F.desc);
        sdp_push_ss2(F, has, start, end, n, true, true);
        if (!sdp_setup_sub(s, F.desc, ar)) return false;
        if (lane_id() == 0) F.sub[n] = s;
        wsync();
        F.nsub++;
      }
      continue;
    }
    const int med = (start + end) / 2;
    const int dS = F.swp ? med : start, dE = F.swp ? end : med;     // the half the D (end) points come from
    const int eS = F.swp ? start : med, eE = F.swp ? med : end;     // the half the E (start) points come from
    SdpDE *nxtD = bufD[(dep + 1) & 1], *nxtE = bufE[(dep + 1) & 1];
    const int dl = sdp_partition(curD, nxtD, dlo, dhi, med);        // rows < med in [dlo, dlo + dl), the rest behind
    const int el = sdp_partition(curE, nxtE, elo, ehi, med);
    // segments of the halves in the next buffer
    const int ldl = dlo, ldh = dlo + dl, rdl = dlo + dl, rdh = dhi;   // D points of the left / right half of the rows
    const int lel = elo, leh = elo + el, rel = elo + el, reh = ehi;   // E points of the left / right half
    const int Ddl = F.swp ? rdl : ldl, Ddh = F.swp ? rdh : ldh;       // D points of the D half
    const int Eel = F.swp ? lel : rel, Eeh = F.swp ? leh : reh;       // E points of the E half
    const int cD = Ddh - Ddl, cE = Eeh - Eel;
    if (!F.swp) {
      if (cD) { s.Di = ar.alloc<long long>(cD); if (ar.overflow) return false; s.m = sdp_unique_sorted(nxtD + Ddl, cD, s.Di, F.desc); }
      if (cE) { s.Ei = ar.alloc<long long>(cE); if (ar.overflow) return false; s.n = sdp_unique_sorted(nxtE + Eel, cE, s.Ei, F.desc); }
    } else {
      if (cE) { s.Ei = ar.alloc<long long>(cE); if (ar.overflow) return false; s.n = sdp_unique_sorted(nxtE + Eel, cE, s.Ei, F.desc); }
      if (cD) { s.Di = ar.alloc<long long>(cD); if (ar.overflow) return false; s.m = sdp_unique_sorted(nxtD + Ddl, cD, s.Di, F.desc); }
    }
    if (s.n == 0 && s.m == 0) continue;
    sdp_push_ss2(F, has, dS, dE, n, true, false);
    sdp_push_ss2(F, has, eS, eE, n, false, true);
    if (s.n != 0 && s.m != 0) { if (!sdp_setup_sub(s, F.desc, ar)) return false; }
    if (lane_id() == 0) F.sub[n] = s;
    wsync();
    F.nsub++;
    // children: the first visited (the D half) is pushed last; a half keeps BOTH its D and its E points
    const bool goD = s.m != 0, goE = s.n != 0;
    const int Dhalf_dl = Ddl, Dhalf_dh = Ddh, Dhalf_el = F.swp ? rel : lel, Dhalf_eh = F.swp ? reh : leh;
    const int Ehalf_dl = F.swp ? ldl : rdl, Ehalf_dh = F.swp ? ldh : rdh, Ehalf_el = Eel, Ehalf_eh = Eeh;
    if (goE) { stS[sp] = eS; stE[sp] = eE; stDl[sp] = Ehalf_dl; stDh[sp] = Ehalf_dh; stEl[sp] = Ehalf_el; stEh[sp] = Ehalf_eh; stDep[sp] = dep + 1; sp++; }
    if (goD) { stS[sp] = dS; stE[sp] = dE; stDl[sp] = Dhalf_dl; stDh[sp] = Dhalf_dh; stEl[sp] = Dhalf_el; stEh[sp] = Dhalf_eh; stDep[sp] = dep + 1; sp++; }
  }
  // ---- the small subtrees, one per lane
  wsync();
  if (njobs > 0) {
    unsigned char *heap = ar.alloc<unsigned char>(heap_total + 16);
    if (ar.overflow) return false;
    bool bad = false;
    for (int j = lane_id(); j < njobs; j += kLanes) {
      const SdpSmallJob J = jobs[j];
      SdpLaneHeap H; H.p = heap + J.heap_off; H.end = H.p + J.heap_len; H.ovf = 0;
      if (!sdp_subtree_lane(F, has, bufD, bufE, J, H)) bad = true;
    }
    wsync();
    if (wany(bad)) { ar.overflow = 1; return false; }
  }
  return true;
}

__device__ __forceinline__ bool sdp_divide(SdpWork &W, SdpFam &F, Arena &ar) {
  const unsigned long long hi = ar.mark_hi();          // the list buffers, prefix counts and subtree queue live at the far end of the arena
  const bool ok = sdp_divide_impl(W, F, ar);
  ar.release_hi(hi);
  return ok;
}

// ---- problem set-up ----------------------------------------------------------------------------------------------
struct SdpAnchors {          // concatenated anchors of the clusters of one problem
  const uint32_t *q, *t; const int32_t *len; int nfrag;
  const int *cl_off; const uint8_t *cl_strand; int ncl;      // mode 0 / 1 / 3
  // mode 4 (split clusters as fragments, SparseDP.h:1956): boxes (q, t = starts; qe, te = ends), per-fragment strand, value (Cluster::Val) and NumofAnchors0
  const uint32_t *qe, *te; const uint8_t *fstrand; const float *fval; const int32_t *fn0;
};

__device__ __forceinline__ void sdp_put_pair(SdpPt *H, int at, uint32_t frag, uint32_t qs, uint32_t ts, int len, int cl, int pair, int strand) {
  SdpPt s, e;
  s.frag = e.frag = frag; s.cl = e.cl = cl; s.src = (uint32_t)at; e.src = (uint32_t)at + 1;
  if (pair == 0) {
    s.q = qs; s.t = ts; s.fl = 1u | 2u | ((uint32_t)strand << 2);
    e.q = qs + (uint32_t)len; e.t = ts + (uint32_t)len; e.fl = 0u | 2u | ((uint32_t)strand << 2);
  } else {
    s.q = qs; s.t = ts + (uint32_t)len; s.fl = 1u | ((uint32_t)strand << 2);
    e.q = qs + (uint32_t)len; e.t = ts; e.fl = 0u | ((uint32_t)strand << 2);
  }
  H[at] = s; H[at + 1] = e;
}

// mode 0: pure matches of all clusters (SparseDP.h:2139); 1: one cluster `only_cl` (:2287); 2: forward only (SparseDP_Forward.h:312);
// 3: the Cluster_SameDiag anchors of the clusters of a split chain (:1766, high-accuracy pipeline): like 0 without the boundary pairs;
// 4: split clusters as fragments with four points each (:1956, the first SparseDP of the high-accuracy pipeline)
__device__ __noinline__ bool sdp_build(SdpWork &W, const SdpAnchors &A, int mode, int only_cl, float rate, int irate, Arena &ar) {
  unsigned long long tk_ = ar.now();
  // count points
  int N = 0, f0 = 0, f1 = A.nfrag;
  if (mode == 0) { N = 2 * A.nfrag; for (int c = 0; c < A.ncl; c++) { const int sz = A.cl_off[c + 1] - A.cl_off[c]; if (sz == 1) N += 2; else if (sz > 1) N += 4; } }
  else if (mode == 1) { f0 = A.cl_off[only_cl]; f1 = A.cl_off[only_cl + 1]; N = 2 * (f1 - f0); }
  else if (mode == 4) N = 4 * A.nfrag;
  else N = 2 * A.nfrag;
  W.N = N; W.nfrag = mode == 1 ? f1 - f0 : A.nfrag;
  const int P = next_pow2(N > 0 ? N : 1);
  W.H1 = ar.alloc<SdpPt>(P); W.H2 = ar.alloc<uint32_t>(P);
  W.val = ar.alloc<SdpVal>(W.nfrag > 0 ? W.nfrag : 1);
  if (ar.overflow) return false;
  // points (serial over clusters for the extra pairs, lane-parallel over anchors)
  if (mode == 0) {
    int at = 0;
    for (int c = 0; c < A.ncl; c++) {
      const int o = A.cl_off[c], sz = A.cl_off[c + 1] - o, st = A.cl_strand[c];
      // anchor i of the cluster occupies 2 points, +2 for i == 0 and (again) for i == sz-1 when sz > 1
      for (int i = lane_id(); i < sz; i += kLanes) {
        const int extra_before = (i > 0 ? 2 : 0);
        const int base = at + 2 * i + extra_before;
        const uint32_t g = (uint32_t)(o + i);
        if (st == 0) {
          sdp_put_pair(W.H1, base, g, A.q[g], A.t[g], A.len[g], c, 0, 1);
          if (i == 0 || i == sz - 1) sdp_put_pair(W.H1, base + 2, g, A.q[g], A.t[g], A.len[g], c, 1, 1);
        } else {
          sdp_put_pair(W.H1, base, g, A.q[g], A.t[g], A.len[g], c, 1, 0);
          if (i == 0 || i == sz - 1) sdp_put_pair(W.H1, base + 2, g, A.q[g], A.t[g], A.len[g], c, 0, 0);
        }
      }
      at += 2 * sz + (sz == 1 ? 2 : (sz > 1 ? 4 : 0));
    }
  } else if (mode == 1) {
    const int st = A.cl_strand[only_cl];
    for (int i = lane_id(); i < f1 - f0; i += kLanes) {
      const int g = f0 + i;
      if (st == 0) sdp_put_pair(W.H1, 2 * i, (uint32_t)i, A.q[g], A.t[g], A.len[g], only_cl, 0, 1);
      else sdp_put_pair(W.H1, 2 * i, (uint32_t)i, A.q[g], A.t[g], A.len[g], only_cl, 1, 0);
    }
  } else if (mode == 4) {
    // s1 (qS+1, tS+1), e1 (qE-1, tE-1) forward; s2 (qS+1, tE-1), e2 (qE-1, tS+1) backward; orient = strand 0 (SparseDP.h:1960-2012)
    for (int i = lane_id(); i < A.nfrag; i += kLanes) {
      const uint32_t or_ = A.fstrand[i] == 0 ? 1u : 0u;
      SdpPt x; x.frag = (uint32_t)i; x.cl = 0;
      x.q = A.q[i] + 1u; x.t = A.t[i] + 1u; x.fl = 1u | 2u | (or_ << 2); x.src = (uint32_t)(4 * i); W.H1[4 * i] = x;
      x.q = A.qe[i] - 1u; x.t = A.te[i] - 1u; x.fl = 0u | 2u | (or_ << 2); x.src = (uint32_t)(4 * i + 1); W.H1[4 * i + 1] = x;
      x.q = A.q[i] + 1u; x.t = A.te[i] - 1u; x.fl = 1u | 0u | (or_ << 2); x.src = (uint32_t)(4 * i + 2); W.H1[4 * i + 2] = x;
      x.q = A.qe[i] - 1u; x.t = A.t[i] + 1u; x.fl = 0u | 0u | (or_ << 2); x.src = (uint32_t)(4 * i + 3); W.H1[4 * i + 3] = x;
    }
  } else if (mode == 3) {
    for (int c = 0; c < A.ncl; c++) {
      const int o = A.cl_off[c], sz = A.cl_off[c + 1] - o, st = A.cl_strand[c];
      for (int i = lane_id(); i < sz; i += kLanes) {
        const uint32_t g = (uint32_t)(o + i);
        if (st == 0) sdp_put_pair(W.H1, 2 * (int)g, g, A.q[g], A.t[g], A.len[g], c, 0, 1);
        else sdp_put_pair(W.H1, 2 * (int)g, g, A.q[g], A.t[g], A.len[g], c, 1, 0);
      }
    }
  } else {
    for (int i = lane_id(); i < A.nfrag; i += kLanes) sdp_put_pair(W.H1, 2 * i, (uint32_t)i, A.q[i], A.t[i], A.len[i], 0, 0, 1);
  }
  // Value
  for (int i = lane_id(); i < W.nfrag; i += kLanes) {
    SdpVal v; v.prev_sub = -1; v.prev_ind = -1; v.prev = 1; v.inv = 1; v.orient = 1; v.cl = 0;
    const int g = f0 + i;
    if (mode == 2) v.val = (float)(A.len[g] * irate);
    else if (mode == 4) v.val = __fmul_rn(A.fval[g], rate);
    else v.val = __fmul_rn((float)A.len[g], rate);
    W.val[i] = v;
  }
  for (int i = N + lane_id(); i < P; i += kLanes) { SdpPt x; x.q = 0xffffffffu; x.t = 0xffffffffu; x.frag = 0; x.cl = 0; x.fl = 3u; x.src = 0xffffffffu; W.H1[i] = x; }
  wsync();
  if (N == 0) { W.R = W.C = 0; return true; }
  // sort(H1, SortByRowOp) -- ties by source order
  wsort_pow2(W.H1, P, [](const SdpPt &a, const SdpPt &b) {
    if (a.q != b.q) return a.q < b.q;
    if (a.t != b.t) return a.t < b.t;
    if ((a.fl & 1u) != (b.fl & 1u)) return (a.fl & 1u) < (b.fl & 1u);
    return a.src < b.src;
  });
  for (int i = lane_id(); i < P; i += kLanes) W.H2[i] = i < N ? (uint32_t)i : 0xffffffffu;
  wsync();
  { const SdpPt *H = W.H1;
    wsort_pow2(W.H2, P, [H](uint32_t a, uint32_t b) {
      if (a == 0xffffffffu || b == 0xffffffffu) return a < b;
      const SdpPt &x = H[a], &y = H[b];
      if (x.t != y.t) return x.t < y.t;
      if (x.q != y.q) return x.q < y.q;
      if ((x.fl & 1u) != (y.fl & 1u)) return (x.fl & 1u) < (y.fl & 1u);
      return a < b;
    }); }
  // Value[ii].cl / orient from the fragment's start points (SparseDP.h:2205-2267)
  for (int i = lane_id(); i < N; i += kLanes) { const SdpPt &p = W.H1[i]; if (p.fl & 1u) { W.val[p.frag].cl = p.cl; W.val[p.frag].orient = (int)((p.fl >> 2) & 1u); } }
  // rows / columns (GetRowInfo, GetColInfo)
  W.rowOf = ar.alloc<int>(N); W.colOfPos = ar.alloc<int>(N); W.colOfPt = ar.alloc<int>(N);
  W.rowS = ar.alloc<int>(N + 1); W.rowE = ar.alloc<int>(N + 1); W.colS = ar.alloc<int>(N + 1); W.colE = ar.alloc<int>(N + 1);
  W.tmp = ar.alloc<long long>(2ull * P + 64); W.flag = ar.alloc<uint8_t>(N + 1);
  W.dyn.top = 0; W.dyn.err = 0; W.dyn.base = 0; W.dyn.cap = 0;
  if (ar.overflow) return false;
  int R = 0, C = 0;
  for (int b = 0; b < N; b += kLanes) {
    const int i = b + lane_id();
    const bool hr = i < N && (i == 0 || W.H1[i].q != W.H1[i - 1].q);
    const bool hc = i < N && (i == 0 || W.H1[W.H2[i]].t != W.H1[W.H2[i - 1]].t);
    const unsigned mr = ballot(hr), mc = ballot(hc);
    if (hr) W.rowS[R + __popc(mr & lanemask_lt())] = i;
    if (hc) W.colS[C + __popc(mc & lanemask_lt())] = i;
    R += __popc(mr); C += __popc(mc);
  }
  wsync();
  for (int r = lane_id(); r < R; r += kLanes) W.rowE[r] = r + 1 < R ? W.rowS[r + 1] : N;
  for (int c = lane_id(); c < C; c += kLanes) W.colE[c] = c + 1 < C ? W.colS[c + 1] : N;
  wsync();
  for (int r = lane_id(); r < R; r += kLanes) for (int p = W.rowS[r]; p < W.rowE[r]; p++) W.rowOf[p] = r;
  for (int c = lane_id(); c < C; c += kLanes) for (int p = W.colS[c]; p < W.colE[c]; p++) { W.colOfPos[p] = c; W.colOfPt[W.H2[p]] = c; }
  for (int i = lane_id(); i <= N; i += kLanes) W.flag[i] = 0;
  wsync();
  W.R = R; W.C = C;
  tk_ = ar.tick(16, tk_);
  // families
  const int nf = mode == 2 ? 2 : 4;
  for (int f = 0; f < 4; f++) {
    SdpFam &F = W.fam[f];
    F.inv = f < 2 ? 1 : 0; F.cols = f & 1; F.desc = (f == 1 || f == 2) ? 1 : 0; F.swp = f == 3 ? 1 : 0;
    F.nsub = 0; F.cap = 0; F.sub = 0; F.ssA = F.ssB = 0; F.nA = F.nB = 0; F.stride = 0;
    if (f >= nf) continue;
    const int V = F.cols ? C : R;
    int depth = 2; while ((1 << (depth - 2)) < V) depth++;
    F.stride = depth; F.cap = 2 * V + 2;
    F.sub = ar.alloc<SdpSub>(F.cap);
    F.ssA = ar.alloc<uint32_t>((unsigned long long)V * depth); F.ssB = ar.alloc<uint32_t>((unsigned long long)V * depth);
    F.nA = ar.alloc<int>(V); F.nB = ar.alloc<int>(V);
    if (ar.overflow) return false;
    for (int i = lane_id(); i < V; i += kLanes) { F.nA[i] = 0; F.nB[i] = 0; }
    wsync();
    if (!sdp_divide(W, F, ar)) return false;
  }
  tk_ = ar.tick(17, tk_);
  return true;
}
// hand the rest of the arena to the growth heap (call after every uniform allocation the problem still needs)
__device__ inline bool sdp_open_dyn(SdpWork &W, Arena &ar) {
  unsigned long long *cell = ar.alloc<unsigned long long>(2);
  if (ar.overflow) return false;
  if (lane_id() == 0) { cell[0] = 0ull; cell[1] = 0ull; }
  wsync();
  const unsigned long long t = (ar.top + 15ull) & ~15ull;
  W.dyn.top = cell; W.dyn.err = (int *)(cell + 1); W.dyn.base = ar.base + t; W.dyn.cap = ar.cap > t ? ar.cap - t : 0ull; W.dyn.base_off = t;
  return true;
}

// ---- ProcessPoint (SparseDP.h:1016-1350; forward-only: SparseDP_Forward.h:37-130) ------------------------------------
__device__ __noinline__ void sdp_process(SdpWork &W, const SdpAnchors &A, int f0, int mode, float rate, int irate, const Pwl &P) {
  const int lane = lane_id();
  for (int i = 0; i < W.N; i++) {
    const SdpPt pt = W.H1[i];
    const int ind = (int)(pt.fl & 1u), inv = (int)((pt.fl >> 1) & 1u);
    const int fr = inv ? 0 : 2, fc = fr + 1;
    SdpFam &FR = W.fam[fr], &FC = W.fam[fc];
    const long long diag = sdp_diag(pt, inv);
    const int row = W.rowOf[i], col = W.colOfPt[i];
    const uint32_t ii = pt.frag;
    if (ind == 1) {
      const int nR = FR.nB[row], nC = FC.nB[col], total = nR + nC;
      float bonus;
      if (mode == 2) bonus = (float)(A.len[f0 + ii] * irate);
      else if (mode == 4) bonus = __fmul_rn(A.fval[f0 + ii], rate);
      else bonus = __fmul_rn(rate, (float)A.len[f0 + ii]);
      SdpVal v = W.val[ii];
      for (int b = 0; b < total; b += kLanes) {
        const int k = b + lane;
        bool ok = false; float ev = 0.0f; int i1 = 0, j = 0, isRow = 0;
        if (k < total) {
          if (k < nR) { j = (int)FR.ssB[(long long)row * FR.stride + (nR - 1 - k)]; isRow = 1; ok = sdp_eval_start(FR.sub[j], diag, FR.desc, bonus, P, W.dyn, ev, i1); }
          else { const int kk = k - nR; j = (int)FC.ssB[(long long)col * FC.stride + (nC - 1 - kk)]; ok = sdp_eval_start(FC.sub[j], diag, FC.desc, bonus, P, W.dyn, ev, i1); }
        }
        // ordered reduction: the first k (lowest lane) whose Ev is the maximum, if it beats the running value
#if MP_LANES == 1
        if (ok && v.val < ev) { v.val = ev; v.prev_sub = j; v.prev_ind = i1; v.prev = isRow; v.inv = inv; }
#else
        float best = ok ? ev : -3.0e38f;
        best = wmax(best);
        const unsigned mk = ballot(ok && ev == best);
        if (mk != 0u && v.val < best) {
          const int src = __ffs((int)mk) - 1;
          v.val = best; v.prev_sub = bcast(j, src); v.prev_ind = bcast(i1, src); v.prev = bcast(isRow, src); v.inv = inv;
        }
#endif
      }
      if (lane == 0) W.val[ii] = v;
    } else {
      const int nR = FR.nA[row], nC = FC.nA[col], total = nR + nC;
      const float val = W.val[ii].val;
      for (int k = lane; k < total; k += kLanes) {
        SdpFam &F = k < nR ? FR : FC;
        const int j = k < nR ? (int)FR.ssA[(long long)row * FR.stride + (nR - 1 - k)] : (int)FC.ssA[(long long)col * FC.stride + (nC - 1 - (k - nR))];
        SdpSub &s = F.sub[j];
        if (s.n == 0 || s.m == 0) continue;
        const int it = sdp_lb(s.Di, s.m, diag, F.desc);
        if (it >= 0 && s.Dv[it] < val) { s.Dv[it] = val; s.Dp[it] = ii; }
      }
    }
    wsync();
  }
}

// ---- TraceBack (SparseDP.h:1520-1575; with `used`: :1351-1518) --------------------------------------------------------
__device__ __forceinline__ uint32_t sdp_prev_frag(const SdpWork &W, const SdpVal &v) {
  const SdpFam &F = W.fam[(v.inv ? 0 : 2) + (v.prev ? 0 : 1)];
  const SdpSub &s = F.sub[v.prev_sub];
  return s.Dp[s.Ep[v.prev_ind]];
}
// plain traceback; returns the chain length (chain and link need room for nfrag entries)
__device__ __noinline__ int sdp_traceback(const SdpWork &W, uint32_t i, uint32_t *chain, uint8_t *link) {
  int n = 0;
  chain[n++] = i;
  while (W.val[i].prev_sub != -1 && W.val[i].prev_ind != -1) {
    if (n > W.nfrag) break;                     // defensive: a cycle cannot happen in a consistent state
    const SdpVal &v = W.val[i];
    link[n - 1] = v.inv ? 0 : 1;
    i = sdp_prev_frag(W, v);
    chain[n++] = i;
  }
  return n;
}
// traceback that refuses anchors already used by an earlier chain (chain is dropped as a whole)
__device__ __noinline__ int sdp_traceback_used(const SdpWork &W, uint32_t i, uint32_t *chain, uint8_t *link, uint8_t *used) {
  int n = 0;
  if (used[i]) return 0;
  chain[n++] = i; used[i] = 1;
  while (W.val[i].prev_sub != -1 && W.val[i].prev_ind != -1) {
    const SdpVal &v = W.val[i];
    const uint32_t nx = sdp_prev_frag(W, v);
    if (used[nx]) { for (int u = 0; u < n; u++) used[chain[u]] = 0; return 0; }
    link[n - 1] = v.inv ? 0 : 1;
    i = nx;
    chain[n++] = i; used[i] = 1;
  }
  return n;
}

}  // namespace mp
}  // namespace lra
