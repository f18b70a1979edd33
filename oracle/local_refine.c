/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the local-index refinement stage of MapRead:
 *   a12  LocalIndex::IndexSeq                                         /root/reference/MMIndex.h:200-245
 *          StoreMinimizers_noncanonical<LocalTuple,SmallTuple>(.., Global = false)   MinCount.h:181-338
 *          std::sort on LocalTuple::operator< (20-bit tuple only)     TupleOps.h:19-45   (libstdc++ introsort, GCC 13.3: unstable,
 *                                                                     and the order of equal tuples decides the order of the anchors)
 *          RemoveFrequent                                             MMIndex.h:69-85
 *        LocalIndex::LookupIndex                                      MMIndex.h:175-190
 *   a13  REFINEclusters                                               ClusterRefine.h:50-240
 *          SwapStrand                                                 ClusterRefine.h:24-31
 *          Cluster::CHROMIndex / Header::Find / GetNextOffset         Clustering.h:326-336, Genome.h:19-47
 *          CartesianTargetSort / LowerBound / UpperBound              Sorting.h:182-224
 *          CompareLists<LocalTuple,SmallTuple>(.., Global = false, 0, 0, canonical = false)   CompareLists.h:8-146
 *          AppendValues (band + box filter)                           TupleOps.h:159-195
 *          Cluster::SetClusterBoundariesFromMatches                   Clustering.h:308-322
 * A LocalTuple is one uint32: tuple in bits 0..19, window-relative position in bits 20..31 (the layout g++ gives the two
 * bit-fields, which is also the layout of <ref>.gli, MMIndex.h:138-151).
 * Pinned by tests/test_oracle_local_refine.py against the unmodified reference (oracle/ref_wrap.cpp: ref_index_seq,
 * ref_refine_clusters). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LT(v) ((v) & 0xFFFFFu)
#define LP(v) ((v) >> 20)

static inline int l_map2(unsigned char c) {
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3;
               default: return c < 8 ? (c & 3) : 0; }
}
static inline int l_mapN(unsigned char c) {
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3;
               default: return c < 8 ? (c & 3) : 4; }
}

/* StoreMinimizers_noncanonical on one window (seqLen <= 2048, so positions fit the 12-bit field).  Returns the count; the
 * first `cap` are stored. */
long lra_oracle_local_minimizers(const char *seq, uint32_t seqLen, int k, int w, uint32_t *out, long cap) {
  long n_out = 0;
#define EMIT(T, P) do { if (n_out < cap) out[n_out] = ((T) & 0xFFFFFu) | (((uint32_t)(P) & 0xFFFu) << 20); n_out++; } while (0)
  if (seqLen < (uint32_t)k) return 0;
  const int windowSpan = w + k - 1;
  uint32_t m = 0;
  for (int i = 0; i < k; i++) { m <<= 2; m += 3; }
  m &= 0xFFFFFu;
  int nextValidWindowEnd = 0, nextValidWindowStart = 0, valid = 0;
  if (seqLen < (uint32_t)windowSpan) return 0;
  while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
    valid = 1;
    for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++)
      if (l_mapN((unsigned char)seq[n]) > 3) { nextValidWindowStart = n + 1; valid = 0; }
  }
  if (!valid) return 0;
  nextValidWindowEnd = nextValidWindowStart + windowSpan;
  uint32_t cur = 0;
  for (int p = 0; p <= k - 1; p++) { cur = (cur << 2) & 0xFFFFFu; cur = (cur + (uint32_t)l_map2((unsigned char)seq[p])) & 0xFFFFFu; }
  uint32_t actT = cur, actP = 0;
  uint32_t *ringT = (uint32_t *)calloc((size_t)w, sizeof(uint32_t)), *ringP = (uint32_t *)calloc((size_t)w, sizeof(uint32_t));
  ringT[0] = actT; ringP[0] = 0;
  uint32_t p;
  for (p = 1; p < (uint32_t)w && p < seqLen - (uint32_t)k + 1; p++) {
    cur = ((cur << 2) & m); cur = (cur + (uint32_t)l_map2((unsigned char)seq[p + k - 1])) & 0xFFFFFu;
    if (cur < actT) { actT = cur; actP = p; }
    ringT[p % (uint32_t)w] = cur; ringP[p % (uint32_t)w] = p;
  }
  if (nextValidWindowEnd == windowSpan) EMIT(actT, actP);
  for (p = (uint32_t)w; p < seqLen - (uint32_t)k + 1; p++) {
    cur = ((cur << 2) & m); cur = (cur + (uint32_t)l_map2((unsigned char)seq[p + k - 1])) & 0xFFFFFu;
    if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k - 1) {
      if (l_mapN((unsigned char)seq[p + k - 1]) <= 3) nextValidWindowEnd++;
      else {
        nextValidWindowStart = (int)(p + (uint32_t)k);
        valid = 0;
        while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
          valid = 1;
          for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++)
            if (l_mapN((unsigned char)seq[n]) > 3) { nextValidWindowStart = n + 1; valid = 0; }
        }
        if (!valid) { free(ringT); free(ringP); return n_out; }
        nextValidWindowEnd = nextValidWindowStart + windowSpan;
      }
    }
    ringT[p % (uint32_t)w] = cur; ringP[p % (uint32_t)w] = p & 0xFFFu;
    if (p - (uint32_t)w >= actP) {
      actT = ringT[0]; actP = ringP[0];
      for (int j = 1; j < w; j++) if (ringT[j] < actT) { actT = ringT[j]; actP = ringP[j]; }
      if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k) EMIT(actT, actP);
    } else if (cur < actT) {
      actT = cur; actP = p & 0xFFFu;
      if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k) EMIT(actT, actP);
    }
  }
  free(ringT); free(ringP);
  return n_out;
#undef EMIT
}

/* libstdc++ std::sort (introsort) on the 20-bit tuple */
#define LESS(a, b) (LT(a) < LT(b))
static inline void lswap(uint32_t *a, uint32_t *b) { uint32_t x = *a; *a = *b; *b = x; }
static void l_unguarded_linear_insert(uint32_t *last) {
  uint32_t val = *last; uint32_t *next = last - 1;
  while (LESS(val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void l_insertion_sort(uint32_t *first, uint32_t *last) {
  if (first == last) return;
  for (uint32_t *i = first + 1; i != last; ++i) {
    if (LESS(*i, *first)) { uint32_t val = *i; memmove(first + 1, first, (size_t)(i - first) * sizeof(uint32_t)); *first = val; }
    else l_unguarded_linear_insert(i);
  }
}
static void l_push_heap(uint32_t *first, long holeIndex, long topIndex, uint32_t value) {
  long parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && LESS(first[parent], value)) { first[holeIndex] = first[parent]; holeIndex = parent; parent = (holeIndex - 1) / 2; }
  first[holeIndex] = value;
}
static void l_adjust_heap(uint32_t *first, long holeIndex, long len, uint32_t value) {
  const long topIndex = holeIndex;
  long secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (LESS(first[secondChild], first[secondChild - 1])) secondChild--;
    first[holeIndex] = first[secondChild]; holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) { secondChild = 2 * (secondChild + 1); first[holeIndex] = first[secondChild - 1]; holeIndex = secondChild - 1; }
  l_push_heap(first, holeIndex, topIndex, value);
}
static void l_heap_sort_range(uint32_t *first, uint32_t *last) {
  long len = last - first;
  if (len >= 2) for (long parent = (len - 2) / 2;; parent--) { uint32_t v = first[parent]; l_adjust_heap(first, parent, len, v); if (parent == 0) break; }
  while (last - first > 1) { --last; uint32_t v = *last; *last = *first; l_adjust_heap(first, 0, last - first, v); }
}
static void l_introsort_loop(uint32_t *first, uint32_t *last, long depth_limit) {
  while (last - first > 16) {
    if (depth_limit == 0) { l_heap_sort_range(first, last); return; }
    --depth_limit;
    uint32_t *mid = first + (last - first) / 2;
    uint32_t *a = first + 1, *b = mid, *c = last - 1;
    if (LESS(*a, *b)) { if (LESS(*b, *c)) lswap(first, b); else if (LESS(*a, *c)) lswap(first, c); else lswap(first, a); }
    else if (LESS(*a, *c)) lswap(first, a);
    else if (LESS(*b, *c)) lswap(first, c);
    else lswap(first, b);
    uint32_t *lo = first + 1, *hi = last;
    for (;;) {
      while (LESS(*lo, *first)) ++lo;
      --hi;
      while (LESS(*first, *hi)) --hi;
      if (!(lo < hi)) break;
      lswap(lo, hi);
      ++lo;
    }
    l_introsort_loop(lo, last, depth_limit);
    last = lo;
  }
}
void lra_oracle_sort_local(uint32_t *v, long n) {
  if (n <= 0) return;
  long lg = 0; { unsigned long x = (unsigned long)n; while (x > 1) { x >>= 1; lg++; } }
  l_introsort_loop(v, v + n, lg * 2);
  if (n > 16) { l_insertion_sort(v, v + 16); for (uint32_t *i = v + 16; i != v + n; ++i) l_unguarded_linear_insert(i); }
  else l_insertion_sort(v, v + n);
}
#undef LESS

long lra_oracle_remove_frequent(uint32_t *v, long n, int maxFreq) {
  long c = 0, i = 0;
  while (i < n) {
    long ne = i;
    while (ne < n && LT(v[ne]) == LT(v[i])) ne++;
    if (ne - i < maxFreq) for (long j = i; j < ne; j++, c++) v[c] = v[j];
    i = ne;
  }
  return c;
}

/* IndexSeq on a fresh LocalIndex (offset = 0).  seqOffsets / tupleBoundaries receive nIndex + 1 entries (the leading 0 of the
 * constructor included); returns the number of minimizers, or -1 if `cap` is too small. *n_index_out = nIndex. */
long lra_oracle_index_seq(const char *seq, int seqLen, int k, int w, int window, int maxFreq, uint64_t *seqOffsets, uint64_t *tupleBoundaries,
                          uint32_t *minimizers, long cap, int *n_index_out) {
  int nIndex = seqLen / window;
  if (seqLen % window != 0) nIndex += 1;
  uint32_t seqPos = 0;
  long total = 0;
  seqOffsets[0] = 0; tupleBoundaries[0] = 0;
  uint32_t *loc = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(window + 16));
  for (int i = 0; i < nIndex; i++) {
    uint32_t end = (uint32_t)seqLen < seqPos + (uint32_t)window ? (uint32_t)seqLen : seqPos + (uint32_t)window;
    long n = lra_oracle_local_minimizers(seq + seqPos, end - seqPos, k, w, loc, window + 16);
    lra_oracle_sort_local(loc, n);
    n = lra_oracle_remove_frequent(loc, n, maxFreq);
    int adv = window < (int)((uint32_t)seqLen - seqPos) ? window : (int)((uint32_t)seqLen - seqPos);
    seqPos += (uint32_t)adv;
    seqOffsets[i + 1] = seqPos;
    if (total + n > cap) { free(loc); return -1; }
    memcpy(minimizers + total, loc, sizeof(uint32_t) * (size_t)n);
    total += n;
    tupleBoundaries[i + 1] = (uint64_t)total;
  }
  free(loc);
  *n_index_out = nIndex;
  return total;
}

long lra_oracle_lookup_index(const uint64_t *seqOffsets, long nOff, uint64_t query) {
  if (nOff == 0) return 0;
  long lo = 0, len = nOff;
  while (len > 0) { long half = len >> 1; if (seqOffsets[lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
  if (lo >= nOff || seqOffsets[lo] != query) return lo - 1;      /* (lo == nOff: the reference reads one past the vector) */
  return lo;
}

/* CompareLists<LocalTuple,SmallTuple>(.., Global = false, maxDiagNum = 0, minDiagNum = 0, canonical = false).  startGap / endGap are
 * LocalTuples: the differences are truncated to the 20-bit field.  Pairs are stored as (query LocalTuple, target LocalTuple). */
long lra_oracle_compare_lists_local(const uint32_t *q, long nq, const uint32_t *t, long nt, long maxFreq, uint32_t *rq, uint32_t *rt, long cap) {
  long n_out = 0;
#define QK(i) LT(q[i])
#define TK(i) LT(t[i])
#define PUSH(qi, ti) do { if (n_out < cap) { rq[n_out] = q[qi]; rt[n_out] = t[ti]; } n_out++; } while (0)
  if (nq == 0 || nt == 0) return 0;
  long qs = 0, qe = nq - 1, ts = 0, te = nt;
  do {
    uint32_t startGap = 0, endGap = 0;
    while (qs <= qe && QK(qs) < TK(ts)) qs++;
    if (qs >= qe) return n_out;
    startGap = (QK(qs) - TK(ts)) & 0xFFFFFu;
    while (qe > qs && te > ts && QK(qe) > TK(te - 1)) qe--;
    endGap = (TK(te - 1) - QK(qe)) & 0xFFFFFu;
    if (startGap == 0 || startGap > endGap) {
      long tsOrig = ts, qsOrig = qs;
      { long lo = ts, len = te - ts;
        while (len > 0) { long half = len >> 1, mid = lo + half; if (TK(mid) < QK(qs)) { lo = mid + 1; len = len - half - 1; } else len = half; }
        ts = lo; }
      if (ts < nt && TK(ts) == QK(qs)) {   /* (ts == nt would read one past the list in the reference) */
        long tsStart = ts, tsi = ts;
        while (tsi != te && QK(qs) == TK(tsi)) tsi++;
        long qsStart = qs;
        while (qs < qe && QK(qs + 1) == QK(qs)) qs++;
        for (long ti = tsStart; ti != tsi; ti++)
          if (qs - qsStart < maxFreq)
            for (long qi = qsStart; qi <= qs; qi++) PUSH(qi, ti);
      }
      while (ts < te && TK(ts) == TK(tsOrig)) ts++;
      while (qs < qe && QK(qs) == QK(qsOrig)) qs++;
    } else {
      if (te != nt && TK(te - 1) == QK(qe)) { /* pass */ }
      else { long lo = ts, len = te - ts;
        while (len > 0) { long half = len >> 1, mid = lo + half; if (QK(qe) < TK(mid)) len = half; else { lo = mid + 1; len = len - half - 1; } }
        te = lo; }
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

static int hdr_find(const uint64_t *pos, int n, uint64_t query) {   /* Header::Find, Genome.h:19-31 */
  if (n > 0 && query == pos[0]) return 0;
  int lo = 0, len = n;
  while (len > 0) { int half = len >> 1; if (pos[lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
  if (lo < n && query == pos[lo]) return lo;
  return lo - 1;
}

typedef struct { uint32_t t, q; } tq_pair;
static int cmp_tq(const void *a, const void *b) {
  const tq_pair *x = (const tq_pair *)a, *y = (const tq_pair *)b;
  if (x->t != y->t) return x->t < y->t ? -1 : 1;
  if (x->q != y->q) return x->q < y->q ? -1 : 1;
  return 0;
}

/* One iteration (one cluster) of the loop of REFINEclusters.
 *   mq / mt [nm]   the cluster's anchors (read position on the cluster's strand, GLOBAL genome position); on return they hold what the
 *                  reference leaves in clusters[ph].matches (chromosome-relative, forward read coordinates, sorted by (t, q))
 *   box[4]         qStart, qEnd, tStart, tEnd of the cluster (tStart/tEnd global); on return the cluster's box after SwapStrand
 *   hdr_pos[n_hdr] genome.header.pos (n contigs + 1)
 *   gl_* / rd_*    the genome's LocalIndex and the read's LocalIndex of the cluster's strand (nOff = number of seqOffsets)
 *   out            refined anchors: rq (read position on the cluster's strand after the final SwapStrand), rt (chromosome-relative),
 *                  rtup (the shared 20-bit tuple)
 *   info[8]        0: status (0 refined, 1 skipped: no anchors, 2 dropped: spans two contigs), 1: chromIndex, 2..5: refined cluster
 *                  qStart,qEnd,tStart,tEnd, 6: n refined; diag[2]: minDiagNum, maxDiagNum; *eff: refineEffiency
 * Returns the number of refined anchors (all counted, the first `cap` stored). */
long lra_oracle_refine_cluster(uint32_t *mq, uint32_t *mt, long nm, uint32_t *box, int strand, uint32_t readLen, const uint64_t *hdr_pos, int n_hdr,
                               const uint64_t *gl_off, long gl_noff, const uint64_t *gl_bnd, const uint32_t *gl_min,
                               const uint64_t *rd_off, long rd_noff, const uint64_t *rd_bnd, const uint32_t *rd_min,
                               int globalK, int smallK, int window, long localMaxFreq,
                               uint32_t *rq, uint32_t *rt, uint32_t *rtup, long cap, int32_t *info, int64_t *diag, float *eff) {
  for (int i = 0; i < 8; i++) info[i] = 0;
  diag[0] = diag[1] = 0; *eff = 0;
  if (nm == 0) { info[0] = 1; return 0; }
  /* CHROMIndex */
  int first = hdr_find(hdr_pos, n_hdr, (uint64_t)box[2] + 1), last = hdr_find(hdr_pos, n_hdr, (uint64_t)box[3]);
  if (first != last) { info[0] = 2; return 0; }
  const int chrom = first;
  info[1] = chrom;
  const uint32_t chromOffset = (uint32_t)hdr_pos[chrom];
  for (long m = 0; m < nm; m++) mt[m] -= chromOffset;
  const uint32_t chromEndOffset = (uint32_t)hdr_pos[hdr_find(hdr_pos, n_hdr, (uint64_t)box[3]) + 1];
  if (strand == 1) {
    for (long m = 0; m < nm; m++) mq[m] = readLen - (mq[m] + (uint32_t)globalK);
    uint32_t r = box[0]; box[0] = readLen - box[1]; box[1] = readLen - r;
  }
  int64_t maxDN = (int64_t)mt[0] - (int64_t)mq[0], minDN = maxDN;
  for (long m = 0; m < nm; m++) { int64_t d = (int64_t)mt[m] - (int64_t)mq[m]; if (d > maxDN) maxDN = d; if (d < minDN) minDN = d; }
  maxDN += 100; minDN -= 100;
  diag[0] = minDN; diag[1] = maxDN;
  { tq_pair *v = (tq_pair *)malloc(sizeof(tq_pair) * (size_t)nm);   /* CartesianTargetSort: a total order on (t, q) */
    for (long m = 0; m < nm; m++) { v[m].t = mt[m]; v[m].q = mq[m]; }
    qsort(v, (size_t)nm, sizeof(tq_pair), cmp_tq);
    for (long m = 0; m < nm; m++) { mt[m] = v[m].t; mq[m] = v[m].q; }
    free(v); }
  const uint32_t segStart = box[2], segEnd = box[3];
  uint32_t wts, wte;
  if (chromOffset + (uint32_t)window > segStart) wts = chromOffset; else wts = segStart - (uint32_t)window;
  if (segEnd + (uint32_t)window > chromEndOffset) wte = chromEndOffset - 1; else wte = segEnd + (uint32_t)window;
  const long ls = lra_oracle_lookup_index(gl_off, gl_noff, wts), le = lra_oracle_lookup_index(gl_off, gl_noff, wte);
  long n_out = 0;
  uint32_t *sq = (uint32_t *)malloc(sizeof(uint32_t) * 65536), *st = (uint32_t *)malloc(sizeof(uint32_t) * 65536);
  long scap = 65536;
  for (long lsi = ls; lsi <= le; lsi++) {
    if (gl_off[lsi] < chromOffset || gl_off[lsi + 1] < chromOffset) continue;
    const uint32_t gStart = (uint32_t)(gl_off[lsi] - chromOffset), gEnd = (uint32_t)(gl_off[lsi + 1] - 1 - chromOffset);
    if (gStart >= gEnd) continue;
    long matchStart, matchEnd;
    { long lo = 0, len = nm;      /* first anchor with (t, q) >= (gStart, 0) */
      while (len > 0) { long half = len >> 1; if (mt[lo + half] < gStart) { lo += half + 1; len -= half + 1; } else len = half; }
      matchStart = lo; }
    { long lo = matchStart, len = nm - matchStart;   /* first anchor with (gEnd, 0) < (t, q) */
      while (len > 0) { long half = len >> 1, mid = lo + half; const int less = (gEnd != mt[mid]) ? (gEnd < mt[mid]) : (0 < mq[mid]);
                        if (less) len = half; else { lo = mid + 1; len -= half + 1; } }
      matchEnd = lo; }
    if (matchEnd == nm) matchEnd--;
    if (matchStart >= nm) continue;
    const uint32_t prev_readEnd = 0;           /* re-declared in every iteration of the reference's loop before its only read */
    uint32_t readStart = mq[matchStart], readEnd = mq[matchEnd];
    if (readStart == readEnd) { if (lsi > ls && readStart > prev_readEnd) readStart = prev_readEnd; }
    if (lsi == ls) { if (readStart < (uint32_t)window) readStart = 0; else readStart -= (uint32_t)window; }
    if (lsi == le) { if (readEnd + (uint32_t)window > readLen) readEnd = readLen; else readEnd += (uint32_t)window; }
    if (readStart > readEnd) continue;
    const long qis = lra_oracle_lookup_index(rd_off, rd_noff, readStart);
    const long qie = lra_oracle_lookup_index(rd_off, rd_noff, readEnd < readLen - 1 ? readEnd : readLen - 1);
    for (long qi = qis; qi <= qie; ++qi) {
      const uint32_t qb0 = (uint32_t)rd_bnd[qi], qb1 = (uint32_t)rd_bnd[qi + 1];
      const uint32_t readSegmentStart = (uint32_t)rd_off[qi];
      long n = lra_oracle_compare_lists_local(rd_min + qb0, (long)qb1 - (long)qb0, gl_min + gl_bnd[lsi], (long)(gl_bnd[lsi + 1] - gl_bnd[lsi]), localMaxFreq, sq, st, scap);
      if (n > scap) {
        scap = n; sq = (uint32_t *)realloc(sq, sizeof(uint32_t) * (size_t)scap); st = (uint32_t *)realloc(st, sizeof(uint32_t) * (size_t)scap);
        lra_oracle_compare_lists_local(rd_min + qb0, (long)qb1 - (long)qb0, gl_min + gl_bnd[lsi], (long)(gl_bnd[lsi + 1] - gl_bnd[lsi]), localMaxFreq, sq, st, scap);
      }
      for (long i = 0; i < n; i++) {      /* AppendValues with band and box */
        const uint32_t qp = LP(sq[i]) + readSegmentStart, tp = LP(st[i]) + gStart;
        const int64_t d = (int64_t)tp - (int64_t)qp;
        if (d >= minDN && d <= maxDN && qp >= box[0] && qp < box[1] && tp >= box[2] - chromOffset && tp < box[3] - chromOffset) {
          if (n_out < cap) { rq[n_out] = qp; rt[n_out] = tp; rtup[n_out] = LT(sq[i]); }
          n_out++;
        }
      }
    }
  }
  free(sq); free(st);
  info[6] = (int32_t)n_out;
  if (n_out == 0) return 0;
  const long ns = n_out < cap ? n_out : cap;
  if (strand == 1) for (long i = 0; i < ns; i++) rq[i] = readLen - (rq[i] + (uint32_t)smallK);
  if (n_out <= cap) {          /* SetClusterBoundariesFromMatches(smallOpts) */
    uint32_t qS = rq[0], qE = qS + (uint32_t)smallK, tS = rt[0], tE = tS + (uint32_t)smallK;
    for (long i = 1; i < n_out; i++) {
      if (rt[i] + (uint32_t)smallK > tE) tE = rt[i] + (uint32_t)smallK;
      if (rt[i] < tS) tS = rt[i];
      if (rq[i] + (uint32_t)smallK > qE) qE = rq[i] + (uint32_t)smallK;
      if (rq[i] < qS) qS = rq[i];
    }
    info[2] = (int32_t)qS; info[3] = (int32_t)qE; info[4] = (int32_t)tS; info[5] = (int32_t)tE;
    const uint32_t den = (qE - qS) < (tE - tS) ? (qE - qS) : (tE - tS);
    *eff = ((float)n_out) / den;       /* float / uint32 -> float division (ClusterRefine.h:236) */
  }
  return n_out;
}

/* ---- a13 (low-accuracy pipeline): one iteration (one split chain) of the loop of Refine_splitchain, ChainRefine.h:383-576.
 *   mq / mt / mlen / mstrand [n]   the chain's anchors in chain order: read position and GLOBAL genome position as stored in the clusters,
 *                                  anchor length (matchesLengths) and the strand of the cluster the anchor comes from.  The reference flips
 *                                  the clusters in place for the duration of the call (t -= chromOffset; q = L - (q + opts.globalK) on
 *                                  strand-1 clusters, ChainRefine.h:399-411) and restores them before returning (:554-565).
 *   box[4]                         QStart, QEnd, TStart, TEnd of the split chain (T global), chrom = its chromIndex, strand = its Strand
 *   limitrefine                    opts.limitrefine (default true): the band of every genome window is [min diag of the anchors in the
 *                                  window - 100, +inf): the upper bound `miniMaxDiag` is read uninitialised in the reference (:493) and holds
 *                                  a stack address in the stock build, i.e. it never filters
 *   info / diag / eff as in lra_oracle_refine_cluster (status 1: empty chain).  A window index past the genome's last window (le can be one
 *   past it when TEnd + window reaches the end of the last contig; the reference then reads past seqOffsets) is skipped. */
long lra_oracle_refine_splitchain(const uint32_t *mq_in, const uint32_t *mt_in, const uint32_t *mlen, const uint8_t *mstrand, long n, const uint32_t *box,
                                  int chrom, int strand, uint32_t readLen, const uint64_t *hdr_pos, int n_hdr,
                                  const uint64_t *gl_off, long gl_noff, const uint64_t *gl_bnd, const uint32_t *gl_min,
                                  const uint64_t *rd_off, long rd_noff, const uint64_t *rd_bnd, const uint32_t *rd_min,
                                  int globalK, int smallK, int window, long localMaxFreq, int limitrefine,
                                  uint32_t *rq, uint32_t *rt, uint32_t *rtup, long cap, int32_t *info, int64_t *diag, float *eff) {
  for (int i = 0; i < 8; i++) info[i] = 0;
  diag[0] = diag[1] = 0; *eff = 0;
  if (n == 0) { info[0] = 1; return 0; }
  info[1] = chrom;
  const uint32_t chromOffset = (uint32_t)hdr_pos[chrom];
  uint32_t *mq = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n), *mt = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
  for (long i = 0; i < n; i++) { mt[i] = mt_in[i] - chromOffset; mq[i] = mstrand[i] ? readLen - (mq_in[i] + (uint32_t)globalK) : mq_in[i]; }
  const uint32_t QStart = box[0], QEnd = box[1], TStart = box[2], TEnd = box[3];
  const uint32_t chromEndOffset = (uint32_t)hdr_pos[hdr_find(hdr_pos, n_hdr, (uint64_t)TEnd) + 1];
  int64_t maxDN = (int64_t)mt[0] - (int64_t)mq[0], minDN = maxDN;
  for (long i = 0; i < n; i++) { int64_t d = (int64_t)mt[i] - (int64_t)mq[i]; if (d > maxDN) maxDN = d; if (d < minDN) minDN = d; }
  maxDN += 50; minDN -= 50;
  diag[0] = minDN; diag[1] = maxDN;
  const uint32_t wts = (TStart >= chromOffset + (uint32_t)window) ? TStart - (uint32_t)window : chromOffset;
  const uint32_t wte = (TEnd + (uint32_t)window < chromEndOffset) ? TEnd + (uint32_t)window : chromEndOffset;
  const long ls = lra_oracle_lookup_index(gl_off, gl_noff, wts), le = lra_oracle_lookup_index(gl_off, gl_noff, wte);
  long n_out = 0, scap = 65536;
  uint32_t *sq = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)scap), *st = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)scap);
  long matchStart = 0, matchEnd = 0;
  uint32_t bqs, bqe;
  if (strand == 0) { bqs = QStart; bqe = QEnd; } else { bqs = readLen - QEnd; bqe = readLen - QStart; }
  const uint32_t bts = TStart - chromOffset, bte = TEnd - chromOffset;
  for (long lsi = ls; lsi <= le; lsi++) {
    if (lsi + 1 >= gl_noff) continue;
    if (gl_off[lsi] < chromOffset || gl_off[lsi + 1] < chromOffset) continue;
    const uint32_t gStart = (uint32_t)(gl_off[lsi] - chromOffset), gEnd = (uint32_t)(gl_off[lsi + 1] - 1 - chromOffset);
    if (gStart >= gEnd) continue;
    while (matchStart < n && mt[matchStart] <= gStart) matchStart++;
    matchEnd = matchStart;
    while (matchEnd < n && mt[matchEnd] < gEnd) matchEnd++;
    if (matchStart >= n) continue;
    if (matchEnd == matchStart) continue;
    uint32_t readStart = mq[matchStart], readEnd = mq[matchEnd - 1];
    for (long mi = matchStart; mi < matchEnd; mi++) {
      if (mq[mi] < readStart) readStart = mq[mi];
      if (mq[mi] + mlen[mi] > readEnd) readEnd = mq[mi] + mlen[mi];
    }
    if (readStart == readEnd) { if (lsi > ls && readStart > 0u) readStart = 0u; }
    int64_t bandMin = minDN, bandMax = maxDN;
    if (limitrefine) {
      int64_t mn = (int64_t)mt[matchStart] - (int64_t)mq[matchStart];
      for (long mi = matchStart; mi < matchEnd; mi++) { int64_t d = (int64_t)mt[mi] - (int64_t)mq[mi]; if (d < mn) mn = d; }
      bandMin = mn - 100; bandMax = INT64_MAX;
    }
    const uint32_t sow = 500;
    if (lsi == ls) readStart = (readStart < sow) ? 0 : readStart - sow;
    if (lsi == le) readEnd = (readEnd + sow > readLen) ? readLen : readEnd + sow;
    if (readStart > readEnd) continue;
    const long qis = lra_oracle_lookup_index(rd_off, rd_noff, readStart);
    const long qie = lra_oracle_lookup_index(rd_off, rd_noff, readEnd < readLen - 1 ? readEnd : readLen - 1);
    for (long qi = qis; qi <= qie; ++qi) {
      const uint32_t qb0 = (uint32_t)rd_bnd[qi], qb1 = (uint32_t)rd_bnd[qi + 1];
      const uint32_t readSegmentStart = (uint32_t)rd_off[qi];
      long m = lra_oracle_compare_lists_local(rd_min + qb0, (long)qb1 - (long)qb0, gl_min + gl_bnd[lsi], (long)(gl_bnd[lsi + 1] - gl_bnd[lsi]), localMaxFreq, sq, st, scap);
      if (m > scap) {
        scap = m; sq = (uint32_t *)realloc(sq, sizeof(uint32_t) * (size_t)scap); st = (uint32_t *)realloc(st, sizeof(uint32_t) * (size_t)scap);
        lra_oracle_compare_lists_local(rd_min + qb0, (long)qb1 - (long)qb0, gl_min + gl_bnd[lsi], (long)(gl_bnd[lsi + 1] - gl_bnd[lsi]), localMaxFreq, sq, st, scap);
      }
      for (long i = 0; i < m; i++) {
        const uint32_t qp = LP(sq[i]) + readSegmentStart, tp = LP(st[i]) + gStart;
        const int64_t d = (int64_t)tp - (int64_t)qp;
        if (d >= bandMin && d <= bandMax && qp >= bqs && qp < bqe && tp >= bts && tp < bte) {
          if (n_out < cap) { rq[n_out] = qp; rt[n_out] = tp; rtup[n_out] = LT(sq[i]); }
          n_out++;
        }
      }
    }
  }
  free(sq); free(st); free(mq); free(mt);
  info[6] = (int32_t)n_out;
  if (n_out == 0) return 0;
  const long ns = n_out < cap ? n_out : cap;
  if (strand == 1) for (long i = 0; i < ns; i++) rq[i] = readLen - (rq[i] + (uint32_t)smallK);
  if (n_out <= cap) {
    uint32_t qS = rq[0], qE = qS + (uint32_t)smallK, tS = rt[0], tE = tS + (uint32_t)smallK;
    for (long i = 1; i < n_out; i++) {
      if (rt[i] + (uint32_t)smallK > tE) tE = rt[i] + (uint32_t)smallK;
      if (rt[i] < tS) tS = rt[i];
      if (rq[i] + (uint32_t)smallK > qE) qE = rq[i] + (uint32_t)smallK;
      if (rq[i] < qS) qS = rq[i];
    }
    info[2] = (int32_t)qS; info[3] = (int32_t)qE; info[4] = (int32_t)tS; info[5] = (int32_t)tE;
    const uint32_t den = (qE - qS) < (tE - tS) ? (qE - qS) : (tE - tS);
    *eff = ((float)n_out) / den;
  }
  return n_out;
}
