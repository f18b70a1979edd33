/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the per-read tail of MapRead after the statistics:
 *   SegAlignmentGroup::SetFromSegAlignment   /root/reference/Alignment.h:944-983   aggregate the segments of an alignment, flags
 *   AlignmentsOrder::Update / operator() / Sort   Alignment.h:1024-1062             rank the alignments added since the last Update by (value,
 *                                                                                   NumOfAnchors0) descending with std::sort, primary / secondary
 *   SimpleMapQV                              Mapping_ultility.h:497-589             MAPQ of every segment (binary32, host logf)
 * Groups g = 0 .. n_groups-1 own segments seg_off[g] .. seg_off[g+1].  update_at[u] = number of groups that exist when the u-th Update is
 * called (Map_lowacc.h:609: once, at the end; Map_highacc.h:737: after every primary chain).  An Update that finds no new group reads one past
 * its index vector in the reference (undefined); it is a no-op here.
 * Per segment in/out: flag, typeofaln, issec (ISsecondary), supp (Supplymentary); out: mapq.  Per group out: g_issec, g_value, g_n0, g_n1,
 * g_nm[4] (nm, nmm, ndel, nins), order[] (AlignmentsOrder::index).
 * Pinned by tests/test_mapq.py against the unmodified reference (oracle/ref_wrap.cpp: ref_mapq). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float value; int n0; } ord_key;
static int ord_less(const ord_key *K, int i, int j) {       /* AlignmentsOrder::operator() */
  if (K[i].value != K[j].value) return K[i].value > K[j].value;
  return K[i].n0 > K[j].n0;
}
#define LESS(a, b) ord_less(K, (a), (b))
static void o_unguarded_linear_insert(const ord_key *K, int *last) { int val = *last; int *next = last - 1; while (LESS(val, *next)) { *last = *next; last = next; --next; } *last = val; }
static void o_insertion_sort(const ord_key *K, int *first, int *last) {
  if (first == last) return;
  for (int *i = first + 1; i != last; ++i) {
    if (LESS(*i, *first)) { int val = *i; memmove(first + 1, first, (size_t)(i - first) * sizeof(int)); *first = val; }
    else o_unguarded_linear_insert(K, i);
  }
}
static void o_adjust_heap(const ord_key *K, int *first, long holeIndex, long len, int value) {
  const long topIndex = holeIndex; long secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) { secondChild = 2 * (secondChild + 1); if (LESS(first[secondChild], first[secondChild - 1])) secondChild--; first[holeIndex] = first[secondChild]; holeIndex = secondChild; }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) { secondChild = 2 * (secondChild + 1); first[holeIndex] = first[secondChild - 1]; holeIndex = secondChild - 1; }
  long parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && LESS(first[parent], value)) { first[holeIndex] = first[parent]; holeIndex = parent; parent = (holeIndex - 1) / 2; }
  first[holeIndex] = value;
}
static void o_heap_sort(const ord_key *K, int *first, int *last) {
  long len = last - first;
  if (len >= 2) for (long parent = (len - 2) / 2;; parent--) { int v = first[parent]; o_adjust_heap(K, first, parent, len, v); if (parent == 0) break; }
  while (last - first > 1) { --last; int v = *last; *last = *first; o_adjust_heap(K, first, 0, last - first, v); }
}
static void o_introsort_loop(const ord_key *K, int *first, int *last, long depth_limit) {
  while (last - first > 16) {
    if (depth_limit == 0) { o_heap_sort(K, first, last); return; }
    --depth_limit;
    int *mid = first + (last - first) / 2, *a = first + 1, *b = mid, *c = last - 1, t;
#define SWP(x, y) do { t = *(x); *(x) = *(y); *(y) = t; } while (0)
    if (LESS(*a, *b)) { if (LESS(*b, *c)) SWP(first, b); else if (LESS(*a, *c)) SWP(first, c); else SWP(first, a); }
    else if (LESS(*a, *c)) SWP(first, a);
    else if (LESS(*b, *c)) SWP(first, c);
    else SWP(first, b);
    int *lo = first + 1, *hi = last;
    for (;;) { while (LESS(*lo, *first)) ++lo; --hi; while (LESS(*first, *hi)) --hi; if (!(lo < hi)) break; SWP(lo, hi); ++lo; }
    o_introsort_loop(K, lo, last, depth_limit);
    last = lo;
  }
}
static void o_sort(const ord_key *K, int *v, long n) {
  if (n <= 1) return;
  long lg = 0; { unsigned long x = (unsigned long)n; while (x > 1) { x >>= 1; lg++; } }
  o_introsort_loop(K, v, v + n, lg * 2);
  if (n > 16) { o_insertion_sort(K, v, v + 16); for (int *i = v + 16; i != v + n; ++i) o_unguarded_linear_insert(K, i); }
  else o_insertion_sort(K, v, v + n);
}

/* seg arrays: value, n0 (NumOfAnchors0), n1, nm, nmm, ndel, nins (the Alignment members), strand.  bypass = opts.bypassClustering,
 * read_type: 0 ont, 1 clr, 2 ccs, 3 contig (Options::AlignType), K = opts.globalK (smallOpts at the call sites). */
void lra_oracle_mapq(int n_groups, const int32_t *seg_off, const float *value, const int32_t *n0, const int32_t *n1, const int32_t *nm, const int32_t *nmm, const int32_t *ndel,
                     const int32_t *nins, const uint8_t *strand, const int32_t *update_at, int n_updates, int bypass, int read_type, int K,
                     int32_t *flag, int32_t *typeofaln, uint8_t *issec, uint8_t *supp, int32_t *mapq,
                     uint8_t *g_issec, float *g_value, int32_t *g_n0, int32_t *g_n1, int32_t *g_nm, int32_t *order) {
  ord_key *keys = (ord_key *)calloc((size_t)(n_groups + 1), sizeof(ord_key));
  int oldend = 0, u = 0;
  const int S = n_groups ? seg_off[n_groups] : 0;
  for (int s = 0; s < S; s++) mapq[s] = 0;
  for (int g = 0; g <= n_groups; g++) {
    /* Updates that happen when g groups exist */
    while (u < n_updates && update_at[u] == g) {
      if (g > oldend) {
        for (int i = oldend; i < g; i++) order[i] = i;
        o_sort(keys, order + oldend, g - oldend);
        g_issec[order[oldend]] = 0;
        for (int i = oldend + 1; i < g; i++) g_issec[order[i]] = 1;
        oldend = g;
        for (int i = 0; i < g; i++)
          if (g_issec[i] == 1) for (int z = seg_off[i]; z < seg_off[i + 1]; z++) { flag[z] |= 0x100; if (typeofaln[z] != 3) typeofaln[z] = 2; }
      }
      u++;
    }
    if (g == n_groups) break;
    /* SetFromSegAlignment */
    const int a = seg_off[g], b = seg_off[g + 1];
    g_issec[g] = 0; g_value[g] = 0; g_n0[g] = 0; g_n1[g] = 0; g_nm[4 * g] = g_nm[4 * g + 1] = g_nm[4 * g + 2] = g_nm[4 * g + 3] = 0;
    if (b > a) {
      g_issec[g] = issec[a]; g_n0[g] = n0[a];
      float v = 0;
      for (int s = a; s < b; s++) { g_n1[g] += n1[s]; g_nm[4 * g] += nm[s]; g_nm[4 * g + 1] += nmm[s]; g_nm[4 * g + 2] += ndel[s]; g_nm[4 * g + 3] += nins[s]; v += value[s]; }
      g_value[g] = v;
      int pry = 0;
      for (int s = a; s < b; s++) if (supp[s] == 0) pry++;
      if (pry == 0) supp[a] = 0;
      for (int s = a; s < b; s++) {
        if (s > a) issec[s] = g_issec[g];
        if (strand[s] == 1) flag[s] |= 0x10;
        if (supp[s] == 1) flag[s] |= 0x800;
      }
    }
    keys[g].value = g_value[g]; keys[g].n0 = g_n0[g];
  }
  /* SimpleMapQV over the ordered alignments (index has `oldend` entries) */
  const int len = oldend;
  float q_coef;
  if (bypass && read_type == 1) q_coef = 4.0f; else if (bypass && read_type == 0) q_coef = 30.0f; else q_coef = 1.0f;
  for (int r = 0; r < len; r++) {
    const int g = order[r];
    if (r == 0) {
      float x = 0, y = 1.0f;
      if (len > 1) x = g_value[order[1]] / g_value[g];
      for (int s = seg_off[g + 1] - 1; s >= seg_off[g]; s--) {
        float pen;
        if (!bypass) { pen = (n0[s] > 20 ? 1.0f : 0.05f) * n0[s]; pen = (n0[s] >= 5 ? 1.0f : 0.1f) * pen; }
        else { if (len > 1) y = ((float)g_n0[g]) / ((float)g_n0[order[1]]); pen = (n0[s] > 10 ? 1.0f : 0.05f) * n0[s]; pen = (n0[s] >= 5 ? 1.0f : 0.02f) * pen; }
        float identity;
        if (nmm[s] + ndel[s] + nins[s] == 0) identity = 1.0f; else identity = ((float)nm[s]) / (nmm[s] + ndel[s] + nins[s]);
        identity = identity < 1 ? identity : 1;
        const float l = value[s] > 3 ? logf(value[s] / K) : 0;
        long m;
        if (len == 1) { if (!bypass) m = (int)(pen * q_coef * l * identity); else m = (int)(pen * q_coef * identity); }
        else {
          if (x >= 0.990f) m = (int)(pen * (1.0f - x) * y * identity);
          else if (!bypass) m = (int)(pen * q_coef * (1.0f - x) * l * y * identity);
          else m = (int)(pen * q_coef * (1.0f - x) * y * identity);
          m -= (int)(4.343f * logf(len) + .499f);
        }
        m = m > 0 ? m : 0;
        mapq[s] = (int32_t)(uint8_t)(m < 60 ? m : 60);
        if (len == 2 && mapq[s] == 0) mapq[s] = 1;
      }
    } else for (int s = seg_off[g]; s < seg_off[g + 1]; s++) mapq[s] = 0;
  }
  free(keys);
}
