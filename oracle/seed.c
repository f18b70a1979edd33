/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the seeding prefix of MapRead (reference MapRead.h:169-203):
 *   a1  CreateRC                          SeqUtils.h:112-158
 *   a2  StoreMinimizers<GenomeTuple,Tuple>(canonical, Global)   MinCount.h:7-179  (+ TupleOps.h:104-138)
 *   a3  std::sort on GenomeTuple::operator< (masked key)        MapRead.h:185, TupleOps.h:76-78
 *         -- std::sort is libstdc++'s introsort (GCC 13.3, bits/stl_algo.h: __introsort_loop, threshold 16, median of
 *            three to first, unguarded partition, final insertion sort; heapsort when the depth limit 2*floor(log2 n)
 *            is hit).  It is unstable, and the order it leaves equal keys in is observable downstream
 *            (CompareLists emits duplicates for the last element of an equal-key run whose strand bits differ), so the
 *            algorithm is restated here and pinned against the real std::sort through oracle/_ref/libref_lra.so.
 *   a4  CompareLists<GenomeTuple,Tuple>(..., Global = true)     CompareLists.h:8-151
 *   a5  SeparateMatchesByStrand (the strncmp test)              MapRead.h:109-150
 * Pinned by tests/test_oracle_seed.py against the unmodified reference headers (oracle/ref_wrap.cpp).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FOR_MASK 0x7FFFFFFFFFFFFFFFull
#define REV_MASK 0x8000000000000000ull

static inline int map2(unsigned char c) { /* seqMap: non-ACGT -> 0 (SeqUtils.h:7-40) */
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3;
               default: return c < 8 ? (c & 3) : 0; }
}
static inline int mapN(unsigned char c) { /* seqMapN: non-ACGT -> 4 */
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3;
               default: return c < 8 ? (c & 3) : 4; }
}

/* a1 */
void lra_oracle_create_rc(const char *seq, long l, char *dest) {
  for (long i = 0; i < l; i++) {
    char c = seq[i], r;
    switch (c) { case 'A': r = 'T'; break; case 'C': r = 'G'; break; case 'G': r = 'C'; break; case 'T': r = 'A'; break;
                 case 'a': r = 't'; break; case 'c': r = 'g'; break; case 'g': r = 'c'; break; case 't': r = 'a'; break;
                 case 'n': r = 'n'; break; default: r = 'N'; }
    dest[l - i - 1] = r;
  }
}

typedef struct { uint64_t t; uint32_t pos; } mtup;

/* a2.  Returns the number of minimizers (all are counted; only the first `cap` are stored). */
static long store_minimizers_impl(const char *seq, uint32_t seqLen, int k, int w, uint64_t *t_out, uint32_t *pos_out, long cap, int canonical) {
  long n_out = 0;
#define EMIT(M) do { if (n_out < cap) { t_out[n_out] = (M).t; pos_out[n_out] = (M).pos; } n_out++; } while (0)
  if (seqLen < (uint32_t)k) return 0;
  const int windowSpan = w + k - 1;
  uint64_t m = 0;
  for (int i = 0; i < k; i++) { m <<= 2; m += 3; }
  int nextValidWindowEnd = 0, nextValidWindowStart = 0, valid = 0;
  if (seqLen < (uint32_t)windowSpan) return 0;
  while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
    valid = 1;
    for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++) {
      if (seqLen < (uint32_t)n) return 0;
      if (mapN((unsigned char)seq[n]) > 3) { nextValidWindowStart = n + 1; valid = 0; }
    }
  }
  if (!valid) return 0;
  nextValidWindowEnd = nextValidWindowStart + windowSpan;
  uint64_t cur = 0, curRC = 0;
  for (int p = 0; p <= k - 1; p++) { cur <<= 2; cur += (uint64_t)map2((unsigned char)seq[p]); }
  { uint64_t a = cur; curRC = 0; for (int i = 0; i < k; i++) { uint64_t least = ~(a & 3) & 3; a >>= 2; curRC <<= 2; curRC += least; } }
  mtup active, curM;
  if (!canonical) active.t = cur;   /* StoreMinimizers_noncanonical (MinCount.h:181-337): forward tuples only, can.t = cur.t */
  else if ((cur & FOR_MASK) < (curRC & FOR_MASK)) active.t = cur & FOR_MASK; else active.t = curRC | REV_MASK;
  active.pos = 0;
  mtup *ring = (mtup *)calloc((size_t)w, sizeof(mtup));
  ring[0] = active;
  uint32_t p;
  for (p = 1; p < (uint32_t)w && p < seqLen - (uint32_t)k + 1; p++) {
    cur = ((cur << 2) & m) + (uint64_t)map2((unsigned char)seq[p + k - 1]);
    curRC >>= 2; curRC += ((~(uint64_t)map2((unsigned char)seq[p + k - 1])) & 3ull) << (2 * ((uint64_t)k - 1));
    curM.pos = p;
    if (!canonical) curM.t = cur & FOR_MASK;
    else if ((cur & FOR_MASK) < (curRC & FOR_MASK)) curM.t = cur & FOR_MASK; else curM.t = curRC | REV_MASK;
    if (curM.t < active.t) { active.t = curM.t; active.pos = p; }   /* first window: UNMASKED comparison (MinCount.h:91) */
    ring[p % (uint32_t)w] = curM;
  }
  if (nextValidWindowEnd == windowSpan) EMIT(active);
  for (p = (uint32_t)w; p < seqLen - (uint32_t)k + 1; p++) {
    if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k - 1) {
      if (mapN((unsigned char)seq[p + k - 1]) <= 3) nextValidWindowEnd++;
      else {
        nextValidWindowStart = (int)(p + (uint32_t)k);
        valid = 0;
        while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
          valid = 1;
          for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++)
            if (mapN((unsigned char)seq[n]) > 3) { nextValidWindowStart = n + 1; valid = 0; }
        }
        if (!valid) { free(ring); return n_out; }
        nextValidWindowEnd = nextValidWindowStart + windowSpan;
      }
    }
    cur = ((cur << 2) & m) + (uint64_t)map2((unsigned char)seq[p + k - 1]);
    curRC >>= 2; curRC += ((~(uint64_t)map2((unsigned char)seq[p + k - 1])) & 3ull) << (2 * ((uint64_t)k - 1));
    if (!canonical) curM.t = cur & FOR_MASK;
    else if ((cur & FOR_MASK) < (curRC & FOR_MASK)) curM.t = cur & FOR_MASK; else curM.t = curRC | REV_MASK;
    curM.pos = p;
    ring[p % (uint32_t)w] = curM;
    if (p - (uint32_t)w >= active.pos) {
      active = ring[0];
      for (int j = 1; j < w; j++) if ((ring[j].t & FOR_MASK) < (active.t & FOR_MASK)) active = ring[j];
      if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k) EMIT(active);
    } else if ((curM.t & FOR_MASK) < (active.t & FOR_MASK)) {
      active = curM;
      if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k) EMIT(active);
    }
  }
  free(ring);
  return n_out;
#undef EMIT
}

long lra_oracle_store_minimizers(const char *seq, uint32_t seqLen, int k, int w, uint64_t *t_out, uint32_t *pos_out, long cap) {
  return store_minimizers_impl(seq, seqLen, k, w, t_out, pos_out, cap, 1);
}
/* StoreMinimizers_noncanonical<GenomeTuple, Tuple>(seq, len, k, w, out, Global = false)   MinCount.h:181-337 (RefineSpace, ClusterRefine.h:297-300) */
long lra_oracle_store_minimizers_nc(const char *seq, uint32_t seqLen, int k, int w, uint64_t *t_out, uint32_t *pos_out, long cap) {
  return store_minimizers_impl(seq, seqLen, k, w, t_out, pos_out, cap, 0);
}

/* a3: libstdc++ introsort on (t & FOR_MASK) */
#define LESS(a, b) (((a).t & FOR_MASK) < ((b).t & FOR_MASK))
static inline void mswap(mtup *a, mtup *b) { mtup x = *a; *a = *b; *b = x; }
static void unguarded_linear_insert(mtup *last) {
  mtup val = *last; mtup *next = last - 1;
  while (LESS(val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void insertion_sort(mtup *first, mtup *last) {
  if (first == last) return;
  for (mtup *i = first + 1; i != last; ++i) {
    if (LESS(*i, *first)) { mtup val = *i; memmove(first + 1, first, (size_t)(i - first) * sizeof(mtup)); *first = val; }
    else unguarded_linear_insert(i);
  }
}
static void push_heap_(mtup *first, long holeIndex, long topIndex, mtup value) {
  long parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && LESS(first[parent], value)) { first[holeIndex] = first[parent]; holeIndex = parent; parent = (holeIndex - 1) / 2; }
  first[holeIndex] = value;
}
static void adjust_heap(mtup *first, long holeIndex, long len, mtup value) {
  const long topIndex = holeIndex;
  long secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (LESS(first[secondChild], first[secondChild - 1])) secondChild--;
    first[holeIndex] = first[secondChild]; holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) { secondChild = 2 * (secondChild + 1); first[holeIndex] = first[secondChild - 1]; holeIndex = secondChild - 1; }
  push_heap_(first, holeIndex, topIndex, value);
}
static void heap_sort_range(mtup *first, mtup *last) { /* __partial_sort(first, last, last) = make_heap + sort_heap */
  long len = last - first;
  if (len >= 2) for (long parent = (len - 2) / 2;; parent--) { mtup v = first[parent]; adjust_heap(first, parent, len, v); if (parent == 0) break; }
  while (last - first > 1) { --last; mtup v = *last; *last = *first; adjust_heap(first, 0, last - first, v); }
}
static void introsort_loop(mtup *first, mtup *last, long depth_limit) {
  while (last - first > 16) {
    if (depth_limit == 0) { heap_sort_range(first, last); return; }
    --depth_limit;
    mtup *mid = first + (last - first) / 2;
    mtup *a = first + 1, *b = mid, *c = last - 1;
    if (LESS(*a, *b)) { if (LESS(*b, *c)) mswap(first, b); else if (LESS(*a, *c)) mswap(first, c); else mswap(first, a); }
    else if (LESS(*a, *c)) mswap(first, a);
    else if (LESS(*b, *c)) mswap(first, c);
    else mswap(first, b);
    mtup *lo = first + 1, *hi = last;
    for (;;) {
      while (LESS(*lo, *first)) ++lo;
      --hi;
      while (LESS(*first, *hi)) --hi;
      if (!(lo < hi)) break;
      mswap(lo, hi);
      ++lo;
    }
    introsort_loop(lo, last, depth_limit);
    last = lo;
  }
}
void lra_oracle_sort_minimizers(uint64_t *t, uint32_t *pos, long n) {
  if (n <= 0) return;
  mtup *v = (mtup *)malloc((size_t)n * sizeof(mtup));
  for (long i = 0; i < n; i++) { v[i].t = t[i]; v[i].pos = pos[i]; }
  long lg = 0; { unsigned long x = (unsigned long)n; while (x > 1) { x >>= 1; lg++; } }
  introsort_loop(v, v + n, lg * 2);
  if (n > 16) { insertion_sort(v, v + 16); for (mtup *i = v + 16; i != v + n; ++i) unguarded_linear_insert(i); }
  else insertion_sort(v, v + n);
  for (long i = 0; i < n; i++) { t[i] = v[i].t; pos[i] = v[i].pos; }
  free(v);
}

/* a4.  Returns the number of pairs (all counted; first `cap` stored). */
static long compare_lists_impl(const uint64_t *qt, const uint32_t *qpos, long nq, const uint64_t *tt, const uint32_t *tpos, long nt, long maxFreq,
                               long maxDiagNum, long minDiagNum, uint64_t *r_qt, uint32_t *r_qpos, uint64_t *r_tt, uint32_t *r_tpos, long cap) {
  long n_out = 0;
#define QK(i) (qt[i] & FOR_MASK)
#define TK(i) (tt[i] & FOR_MASK)
/* the diagonal band applies only when both bounds are non-zero (CompareLists.h:86-95) */
#define PUSH(qi, ti) do { int ok_ = 1; if (maxDiagNum != 0 && minDiagNum != 0) { const long D_ = (long)tpos[ti] - (long)qpos[qi]; ok_ = D_ <= maxDiagNum && D_ >= minDiagNum; } \
    if (ok_) { if (n_out < cap) { r_qt[n_out] = qt[qi]; r_qpos[n_out] = qpos[qi]; r_tt[n_out] = tt[ti]; r_tpos[n_out] = tpos[ti]; } n_out++; } } while (0)
  if (nq == 0 || nt == 0) return 0;
  long qs = 0, qe = nq - 1, ts = 0, te = nt;
  do {
    uint64_t startGap = 0, endGap = 0;   /* (uninitialised in the reference when not assigned; see below) */
    while (qs <= qe && QK(qs) < TK(ts)) qs++;
    if (qs >= qe) return n_out;
    if (qs < qe) startGap = QK(qs) - TK(ts);
    if (qs == qe) endGap = startGap;
    else {
      while (qe > qs && te > ts && QK(qe) > TK(te - 1)) qe--;
      endGap = TK(te - 1) - QK(qe);
    }
    if (startGap == 0 || ((startGap & FOR_MASK) > (endGap & FOR_MASK))) {
      long tsOrig = ts, qsOrig = qs;
      { /* lower_bound(tBegin+ts, tBegin+te, qBegin[qs]) on the masked key */
        long lo = ts, len = te - ts;
        while (len > 0) { long half = len >> 1, mid = lo + half; if (TK(mid) < QK(qs)) { lo = mid + 1; len = len - half - 1; } else len = half; }
        ts = lo;
      }
      if (ts < nt && TK(ts) == QK(qs)) {   /* the reference also reads tBegin[ts] when ts == te; ts == nt would be past the vector (UB), treated as no match */
        long tsStart = ts, tsi = ts;
        while (tsi != te && QK(qs) == TK(tsi)) tsi++;
        long qsStart = qs;
        while (qs < qe && QK(qs + 1) == QK(qs)) qs++;
        for (long ti = tsStart; ti != tsi; ti++)
          if (qs - qsStart < maxFreq)
            for (long qi = qsStart; qi <= qs; qi++) PUSH(qi, ti);
      }
      while (ts < te && tt[ts] == tt[tsOrig]) ts++;        /* UNMASKED comparisons, against the ORIGINAL positions */
      while (qs < qe && qt[qs] == qt[qsOrig]) qs++;
    } else {
      if (te != nt && TK(te - 1) == QK(qe)) { /* pass */ }
      else { /* upper_bound(tBegin+ts, tBegin+te, qBegin[qe]) */
        long lo = ts, len = te - ts;
        while (len > 0) { long half = len >> 1, mid = lo + half; if (QK(qe) < TK(mid)) len = half; else { lo = mid + 1; len = len - half - 1; } }
        te = lo;
      }
      long teStart = te, tei = te;
      while (tei > ts && TK(tei - 1) == QK(qe)) tei--;
      if (tei < teStart && teStart > 0) {
        long qeStart = qe;
        while (qe > qs && QK(qe) == QK(qe - 1)) qe--;
        for (long ti = tei; ti < teStart; ti++)
          if (qeStart - qe < maxFreq)
            for (long qi = qe; qi <= qeStart; qi++) PUSH(qi, ti);
      }
      te = tei;
    }
  } while (qs < qe && ts < te);
  return n_out;
#undef QK
#undef TK
#undef PUSH
}

long lra_oracle_compare_lists(const uint64_t *qt, const uint32_t *qpos, long nq, const uint64_t *tt, const uint32_t *tpos, long nt, long maxFreq,
                              uint64_t *r_qt, uint32_t *r_qpos, uint64_t *r_tt, uint32_t *r_tpos, long cap) {
  return compare_lists_impl(qt, qpos, nq, tt, tpos, nt, maxFreq, 0, 0, r_qt, r_qpos, r_tt, r_tpos, cap);
}
/* CompareLists<GenomeTuple, Tuple>(.., Global = false, maxDiagNum, minDiagNum, canonical = false): maxFreq = opts.localMaxFreq, banded */
long lra_oracle_compare_lists_band(const uint64_t *qt, const uint32_t *qpos, long nq, const uint64_t *tt, const uint32_t *tpos, long nt, long maxFreq,
                                   long maxDiagNum, long minDiagNum, uint64_t *r_qt, uint32_t *r_qpos, uint64_t *r_tt, uint32_t *r_tpos, long cap) {
  return compare_lists_impl(qt, qpos, nq, tt, tpos, nt, maxFreq, maxDiagNum, minDiagNum, r_qt, r_qpos, r_tt, r_tpos, cap);
}

/* a5: strand of every match: 0 if the k read bases equal the k genome bases (strncmp == 0), else 1 */
void lra_oracle_match_strands(const char *read, const char *genome_concat, const uint32_t *qpos, const uint32_t *tpos, long n, int k, uint8_t *strand) {
  for (long i = 0; i < n; i++) strand[i] = strncmp(read + qpos[i], genome_concat + tpos[i], (size_t)k) == 0 ? 0 : 1;
}

/* a2..a5 for one read; returns the number of matches */
long lra_oracle_seed_read(const char *read, uint32_t len, const char *genome_concat, const uint64_t *tt, const uint32_t *tpos, long nt, int k, int w,
                          long maxFreq, uint64_t *r_qt, uint32_t *r_qpos, uint64_t *r_tt, uint32_t *r_tpos, uint8_t *strand, long cap) {
  long mcap = (long)len + 16;
  uint64_t *mt = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)mcap);
  uint32_t *mp = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)mcap);
  long nm = lra_oracle_store_minimizers(read, len, k, w, mt, mp, mcap);
  lra_oracle_sort_minimizers(mt, mp, nm);
  long n = lra_oracle_compare_lists(mt, mp, nm, tt, tpos, nt, maxFreq, r_qt, r_qpos, r_tt, r_tpos, cap);
  lra_oracle_match_strands(read, genome_concat, r_qpos, r_tpos, n < cap ? n : cap, k, strand);
  free(mt); free(mp);
  return n;
}
