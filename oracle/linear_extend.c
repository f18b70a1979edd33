/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the low-accuracy pipeline's linear extension:
 *   LinearExtend (GenomePairs overload)      /root/reference/LinearExtend.h:658-716   merge co-diagonal overlapping K-mers, extend a K-mer by exact
 *                                                                                     base comparison up to the next anchor of the diagonal
 *   Checkbp                                  LinearExtend.h:50-84                     the base-by-base extension
 *   DecideCoordinates                        LinearExtend.h:104-128                   bounding box of the extended cluster
 *   TrimOverlappedAnchors (vector<Cluster>)  LinearExtend.h:573-647, LongAnchors :11-47   trim anchors >= 40 that overlap the next one by <= 30
 * as the two call sites use them: Map_lowacc.h:132-136 (every cluster, skipsorting = 1, no trimming) and Map_lowacc.h:460-474 (the refined
 * clusters merged into one extended cluster per split chain, DiagonalSort first, then TrimOverlappedAnchors over all of them).
 * One read.  Group g (one extended cluster) owns parts g_off[g] .. g_off[g+1]; part p (one input cluster) owns anchors p_off[p] .. p_off[p+1]
 * of (q, t) (t relative to its contig), lies on the contig at chrom_off[p] / chrom_len[p] of `genome`, strand p_strand[p].  The outputs of the
 * parts of a group are appended in order; strand and contig of the group are those of its last part (the reference's loop variables).
 * The reference compares genome.seqs[chrom][curT] with read.seq[curQ] directly on BOTH strands (no complement on strand 1) -- restated as is.
 * All position arithmetic is GenomePos (uint32) arithmetic as in the reference; lengths are int.
 * trim: 0 none, 1 the vector<Cluster> overload of TrimOverlappedAnchors, 2 its GenomePairs overload (LocalRefineAlignment.h:358-373).
 * lra_oracle_linear_extend_chain restates the high-accuracy pipeline's overload LinearExtend(vector<Cluster*>, ..., chain, ...)
 * (LinearExtend.h:134-350, with CheckOverlap :87-101) as LinearExtend_chain (:782-792) and Map_highacc.h:571-582 call it, followed by
 * MergeMatchesSameDiag (:794-823, Map_highacc.h:642).
 * Pinned by tests/test_linear_extend.py against the unmodified reference (oracle/ref_wrap.cpp: ref_linear_extend, ref_linear_extend_chain). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { const uint32_t *q, *t; const int32_t *len; int strand; } la_ctx;
static int la_less(const la_ctx *c, int i, int j) {       /* LongAnchors::operator() */
  if (c->strand == 0) {
    if (c->q[i] != c->q[j]) return c->q[i] < c->q[j];
    return c->t[i] < c->t[j];
  }
  const uint32_t ei = c->q[i] + (uint32_t)c->len[i], ej = c->q[j] + (uint32_t)c->len[j];
  if (ei != ej) return ei > ej;
  return c->t[i] < c->t[j];
}
/* libstdc++ std::sort (GCC 13.3 bits/stl_algo.h) on the index vector: two long anchors can tie under the comparator, and which one is
 * `prev` decides which is trimmed */
#define LESS(a, b) la_less(K, (a), (b))
static void la_unguarded_linear_insert(const la_ctx *K, int *last) { int val = *last; int *next = last - 1; while (LESS(val, *next)) { *last = *next; last = next; --next; } *last = val; }
static void la_insertion_sort(const la_ctx *K, int *first, int *last) {
  if (first == last) return;
  for (int *i = first + 1; i != last; ++i) {
    if (LESS(*i, *first)) { int val = *i; memmove(first + 1, first, (size_t)(i - first) * sizeof(int)); *first = val; }
    else la_unguarded_linear_insert(K, i);
  }
}
static void la_adjust_heap(const la_ctx *K, int *first, long holeIndex, long len, int value) {
  const long topIndex = holeIndex; long secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) { secondChild = 2 * (secondChild + 1); if (LESS(first[secondChild], first[secondChild - 1])) secondChild--; first[holeIndex] = first[secondChild]; holeIndex = secondChild; }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) { secondChild = 2 * (secondChild + 1); first[holeIndex] = first[secondChild - 1]; holeIndex = secondChild - 1; }
  long parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && LESS(first[parent], value)) { first[holeIndex] = first[parent]; holeIndex = parent; parent = (holeIndex - 1) / 2; }
  first[holeIndex] = value;
}
static void la_heap_sort(const la_ctx *K, int *first, int *last) {
  long len = last - first;
  if (len >= 2) for (long parent = (len - 2) / 2;; parent--) { int v = first[parent]; la_adjust_heap(K, first, parent, len, v); if (parent == 0) break; }
  while (last - first > 1) { --last; int v = *last; *last = *first; la_adjust_heap(K, first, 0, last - first, v); }
}
static void la_introsort_loop(const la_ctx *K, int *first, int *last, long depth_limit) {
  while (last - first > 16) {
    if (depth_limit == 0) { la_heap_sort(K, first, last); return; }
    --depth_limit;
    int *mid = first + (last - first) / 2, *a = first + 1, *b = mid, *c = last - 1, t;
#define SWP(x, y) do { t = *(x); *(x) = *(y); *(y) = t; } while (0)
    if (LESS(*a, *b)) { if (LESS(*b, *c)) SWP(first, b); else if (LESS(*a, *c)) SWP(first, c); else SWP(first, a); }
    else if (LESS(*a, *c)) SWP(first, a);
    else if (LESS(*b, *c)) SWP(first, c);
    else SWP(first, b);
    int *lo = first + 1, *hi = last;
    for (;;) { while (LESS(*lo, *first)) ++lo; --hi; while (LESS(*first, *hi)) --hi; if (!(lo < hi)) break; SWP(lo, hi); ++lo; }
    la_introsort_loop(K, lo, last, depth_limit);
    last = lo;
  }
}
static void la_sort(const la_ctx *K, int *v, long n) {
  if (n < 2) return;
  long lg = 0; for (long m = n; m > 1; m >>= 1) lg++;
  la_introsort_loop(K, v, v + n, lg * 2);
  if (n > 16) { la_insertion_sort(K, v, v + 16); for (int *i = v + 16; i != v + n; ++i) la_unguarded_linear_insert(K, i); }
  else la_insertion_sort(K, v, v + n);
}
#undef LESS

typedef struct { uint32_t q, t; } le_pair;
static int diag_cmp(const void *a, const void *b) {        /* DiagonalSortOp (Sorting.h:35-46): a total order on (q, t) */
  const le_pair *x = (const le_pair *)a, *y = (const le_pair *)b;
  const long dx = (long)x->q - (long)x->t, dy = (long)y->q - (long)y->t;
  if (dx != dy) return dx < dy ? -1 : 1;
  return x->q < y->q ? -1 : (x->q > y->q ? 1 : 0);
}

/* Checkbp (LinearExtend.h:50-84) */
static void checkbp(uint32_t cq, uint32_t ct, uint32_t nq, uint32_t nt, const uint8_t *contig, int clen, const uint8_t *read, int read_len, int strand, int K,
                    uint32_t *qe, uint32_t *te) {
  uint32_t curQ, curT, nextQ, nextT;
  const uint32_t L = (uint32_t)clen;
  if (strand == 0) {
    curQ = cq + (uint32_t)K; curT = ct + (uint32_t)K < L ? ct + (uint32_t)K : L;
    nextQ = nq; nextT = nt < L ? nt : L;
    while (curQ < (uint32_t)read_len && curT < L && nextQ > curQ && nextT > curT && contig[curT] == read[curQ]) { curQ++; curT++; }
  } else {
    curQ = cq + (uint32_t)K; curT = ct - 1u < L - 1u ? ct - 1u : L - 1u;
    nextQ = nq; nextT = nt + (uint32_t)K - 1u < L - 1u ? nt + (uint32_t)K - 1u : L - 1u;
    while (curQ < (uint32_t)read_len && nextQ > curQ && nextT < curT && contig[curT] == read[curQ]) { curQ++; curT--; }
  }
  *qe = curQ; *te = curT;
}

/* TrimOverlappedAnchors: the vector<Cluster> overload (LinearExtend.h:573-647; thr 40, strand-aware) and the GenomePairs overload (:724-777; thr 50,
 * forward only = strand 0 here) */
static void trim_group(uint32_t *Q, uint32_t *T, int32_t *L, long cnt, int st, int thr) {
  int *idx = (int *)malloc((size_t)(cnt > 0 ? cnt : 1) * sizeof(int));
  long nl = 0;
  for (long i = 0; i < cnt; i++) if (L[i] >= thr) idx[nl++] = (int)i;
  la_ctx c = {Q, T, L, st};
  la_sort(&c, idx, nl);
  for (long ln = 1; ln < nl; ln++) {
    const int prev = idx[ln - 1], cur = idx[ln];
    int overlap_r = 0, overlap_g = 0;
    if (st == 0) {
      if (Q[cur] < Q[prev] + (uint32_t)L[prev] && Q[cur] >= Q[prev] + (uint32_t)L[prev] - 30u) overlap_r = (int)(Q[prev] + (uint32_t)L[prev] - Q[cur]);
    } else {
      if (Q[cur] + (uint32_t)L[cur] > Q[prev] && Q[cur] + (uint32_t)L[cur] <= Q[prev] + 30u) overlap_r = (int)(Q[cur] + (uint32_t)L[cur] - Q[prev]);
    }
    if (T[cur] < T[prev] + (uint32_t)L[prev] && T[cur] >= T[prev] + (uint32_t)L[prev] - 30u) overlap_g = (int)(T[prev] + (uint32_t)L[prev] - T[cur]);
    if (overlap_r > 0 || overlap_g > 0) {
      const int overlap = overlap_r > overlap_g ? overlap_r : overlap_g;
      if (st == 1) Q[prev] += (uint32_t)(overlap + 1);
      L[prev] -= overlap + 1;
    }
  }
  free(idx);
}

/* Returns the number of extended anchors (== e_off[n_groups]).  q, t are sorted in place per part when skipsorting == 0.
 * box[4g..] = qStart, qEnd, tStart, tEnd (DecideCoordinates, before trimming; zeros for a group without anchors). */
long lra_oracle_linear_extend(const uint8_t *read, int read_len, const uint8_t *genome, const uint64_t *chrom_off, const int32_t *chrom_len, int n_groups,
                              const int32_t *g_off, const int32_t *p_off, const uint8_t *p_strand, uint32_t *q, uint32_t *t, int K, int skipsorting, int trim,
                              int32_t *e_off, uint32_t *eq, uint32_t *et, int32_t *elen, uint32_t *box) {
  long no = 0;
  for (int g = 0; g < n_groups; g++) {
    e_off[g] = (int32_t)no;
    int st = 0;
    for (int p = g_off[g]; p < g_off[g + 1]; p++) {
      const long a = p_off[p], size = p_off[p + 1] - a;
      uint32_t *pq = q + a, *pt = t + a;
      const uint8_t *contig = genome + chrom_off[p];
      const int strand = p_strand[p];
      st = strand;
      if (!skipsorting && size > 1) {
        le_pair *v = (le_pair *)malloc((size_t)size * sizeof(le_pair));
        for (long i = 0; i < size; i++) { v[i].q = pq[i]; v[i].t = pt[i]; }
        qsort(v, (size_t)size, sizeof(le_pair), diag_cmp);
        for (long i = 0; i < size; i++) { pq[i] = v[i].q; pt[i] = v[i].t; }
        free(v);
      }
      long n = 1, m = 0;
      while (n < size) {
        long curDiag, nextDiag;
        if (strand == 0) { curDiag = (long)pq[n - 1] - (long)pt[n - 1]; nextDiag = (long)pq[n] - (long)pt[n]; }
        else { curDiag = (long)pq[n - 1] + (long)pt[n - 1]; nextDiag = (long)pq[n] + (long)pt[n]; }
        if (curDiag == nextDiag) {
          if (pq[n] < pq[n - 1] + (uint32_t)K) n++;
          else {
            uint32_t qe, te;
            checkbp(pq[n - 1], pt[n - 1], pq[n], pt[n], contig, chrom_len[p], read, read_len, strand, K, &qe, &te);
            if (strand == 0 && qe == pq[n] && te == pt[n]) n++;
            else if (strand == 1 && qe == pq[n] && te == pt[n] + (uint32_t)K - 1u) n++;
            else {
              eq[no] = pq[m]; et[no] = strand == 0 ? pt[m] : te + 1u; elen[no] = (int32_t)(qe - pq[m]); no++;
              m = n; n++;
            }
          }
        } else {
          eq[no] = pq[m]; et[no] = strand == 0 ? pt[m] : pt[n - 1]; elen[no] = (int32_t)(pq[n - 1] + (uint32_t)K - pq[m]); no++;
          m = n; n++;
        }
      }
      if (n == size) {
        eq[no] = pq[m]; et[no] = strand == 0 ? pt[m] : pt[n - 1]; elen[no] = (int32_t)(pq[n - 1] + (uint32_t)K - pq[m]); no++;
      }
    }
    /* DecideCoordinates */
    const long b = e_off[g], cnt = no - b;
    box[4 * g] = box[4 * g + 1] = box[4 * g + 2] = box[4 * g + 3] = 0;
    if (cnt > 0) {
      uint32_t qs = eq[b], qe2 = qs + (uint32_t)elen[b], ts = et[b], te2 = ts + (uint32_t)elen[b];
      for (long i = b + 1; i < no; i++) {
        if (eq[i] < qs) qs = eq[i];
        if (eq[i] + (uint32_t)elen[i] > qe2) qe2 = eq[i] + (uint32_t)elen[i];
        if (et[i] < ts) ts = et[i];
        if (et[i] + (uint32_t)elen[i] > te2) te2 = et[i] + (uint32_t)elen[i];
      }
      box[4 * g] = qs; box[4 * g + 1] = qe2; box[4 * g + 2] = ts; box[4 * g + 3] = te2;
    }
    if (trim && cnt > 0) trim_group(eq + b, et + b, elen + b, cnt, trim == 2 ? 0 : st, trim == 2 ? 50 : 40);
  }
  e_off[n_groups] = (int32_t)no;
  return no;
}

static int adiag_cmp(const void *a, const void *b) {       /* AntiDiagonalSortOp (Sorting.h:77-90): (GenomePos)(q + t), then q */
  const le_pair *x = (const le_pair *)a, *y = (const le_pair *)b;
  const uint32_t dx = x->q + x->t, dy = y->q + y->t;
  if (dx != dy) return dx < dy ? -1 : 1;
  return x->q < y->q ? -1 : (x->q > y->q ? 1 : 0);
}

static int check_overlap(uint32_t q, uint32_t t, int K, const uint32_t *set_pos, const uint8_t *set_flag, int ns) {   /* CheckOverlap */
  for (int i = 0; i < ns; i++) {
    if (set_flag[i] == 0 && set_pos[i] >= q && set_pos[i] < q + (uint32_t)K) return 1;
    if (set_flag[i] == 1 && set_pos[i] >= t && set_pos[i] < t + (uint32_t)K) return 1;
  }
  return 0;
}

/* One chain of one read.  Clusters: anchors cl_off[c] .. cl_off[c+1] of (cq, ct) (sorted in place: DiagonalSort on strand 0, AntiDiagonalSort on
 * strand 1), cl_box[4c..] = qStart, qEnd, tStart, tEnd, strand, contig (arena position, length), anchorfreq.  chain[0..n_chain) = cluster indices.
 * Out per chain entry e: extended anchors e_off[e] .. e_off[e+1] (q, t, len, overlap flag), box, and the same-diagonal runs md_off[e] ..
 * md_off[e+1] as (md_start, md_end) -- one run (0, 1) for an entry without anchors, where the reference reads matches[0] of an empty vector.
 * *overlap_count = the reference's `overlap` counter.  Returns the number of extended anchors. */
long lra_oracle_linear_extend_chain(const uint8_t *read, int read_len, const uint8_t *genome, int n_cl, const int32_t *cl_off, uint32_t *cq, uint32_t *ct,
                                    const uint32_t *cl_box, const uint8_t *cl_strand, const uint64_t *cl_chrom_off, const int32_t *cl_chrom_len, const float *cl_freq,
                                    int n_chain, const int32_t *chain, int K, int skiprepetitive, int trim, int merge_dist,
                                    int32_t *e_off, uint32_t *eq, uint32_t *et, int32_t *elen, uint8_t *eovp, uint32_t *box, int32_t *overlap_count,
                                    int32_t *md_off, int32_t *md_start, int32_t *md_end) {
  (void)n_cl;
  long no = 0, nmd = 0;
  int overlap = 0;
  for (int c = 0; c < n_chain; c++) {
    e_off[c] = (int32_t)no;
    md_off[c] = (int32_t)nmd;
    const int cm = chain[c];
    const long a = cl_off[cm], size = cl_off[cm + 1] - a;
    box[4 * c] = box[4 * c + 1] = box[4 * c + 2] = box[4 * c + 3] = 0;
    const int strand = cl_strand[cm];
    if (size > 0) {
      uint32_t set_pos[8]; uint8_t set_flag[8]; int ns = 0;
      const uint32_t qsb = cl_box[4 * cm], qeb = cl_box[4 * cm + 1], tsb = cl_box[4 * cm + 2], teb = cl_box[4 * cm + 3];
      if (skiprepetitive && cl_freq[cm] <= 1.1f) {
        for (int side = 0; side < 2; side++) {
          const int o = side == 0 ? c - 1 : c + 1;
          if (o < 0 || o >= n_chain) continue;
          const uint32_t *ob = cl_box + 4 * chain[o];
          if (ob[0] > qsb && ob[0] < qeb) { set_pos[ns] = ob[0]; set_flag[ns++] = 0; }
          if (ob[1] > qsb && ob[1] < qeb) { set_pos[ns] = ob[1]; set_flag[ns++] = 0; }
          if (ob[2] > tsb && ob[2] < teb) { set_pos[ns] = ob[2]; set_flag[ns++] = 1; }
          if (ob[3] > tsb && ob[3] < teb) { set_pos[ns] = ob[3]; set_flag[ns++] = 1; }
        }
      }
      uint32_t *pq = cq + a, *pt = ct + a;
      if (size > 1) {
        le_pair *v = (le_pair *)malloc((size_t)size * sizeof(le_pair));
        for (long i = 0; i < size; i++) { v[i].q = pq[i]; v[i].t = pt[i]; }
        qsort(v, (size_t)size, sizeof(le_pair), strand == 0 ? diag_cmp : adiag_cmp);
        for (long i = 0; i < size; i++) { pq[i] = v[i].q; pt[i] = v[i].t; }
        free(v);
      }
      const uint8_t *contig = genome + cl_chrom_off[cm];
      long n = 1, m = 0;
      int chm = 1;
#define PUSH(Q_, T_, L_, O_) do { eq[no] = (Q_); et[no] = (T_); elen[no] = (int32_t)(L_); eovp[no] = (O_); no++; } while (0)
      while (n < size) {
        if (chm == 1) {
          if (check_overlap(pq[m], pt[m], K, set_pos, set_flag, ns)) {
            PUSH(pq[m], pt[m], K, 1); overlap++;
            m = n; n++; chm = 1;
            continue;
          } else chm = 0;
        }
        if (check_overlap(pq[n], pt[n], K, set_pos, set_flag, ns)) {
          PUSH(pq[m], strand == 0 ? pt[m] : pt[n - 1], pq[n - 1] + (uint32_t)K - pq[m], 0);
          PUSH(pq[n], pt[n], K, 1); overlap++;
          m = n + 1; n = m + 1; chm = 1;
          continue;
        }
        long curDiag, nextDiag;
        if (strand == 0) { curDiag = (long)pq[n - 1] - (long)pt[n - 1]; nextDiag = (long)pq[n] - (long)pt[n]; }
        else { curDiag = (long)pq[n - 1] + (long)pt[n - 1]; nextDiag = (long)pq[n] + (long)pt[n]; }
        if (curDiag == nextDiag) {
          if (pq[n] < pq[n - 1] + (uint32_t)K) n++;
          else {
            uint32_t qe, te;
            checkbp(pq[n - 1], pt[n - 1], pq[n], pt[n], contig, cl_chrom_len[cm], read, read_len, strand, K, &qe, &te);
            if (strand == 0 && qe == pq[n] && te == pt[n]) n++;
            else if (strand == 1 && qe == pq[n] && te == pt[n] + (uint32_t)K - 1u) n++;
            else { PUSH(pq[m], strand == 0 ? pt[m] : te + 1u, qe - pq[m], 0); m = n; n++; }
          }
        } else { PUSH(pq[m], strand == 0 ? pt[m] : pt[n - 1], pq[n - 1] + (uint32_t)K - pq[m], 0); m = n; n++; }
        chm = 0;
      }
      if (n == size) PUSH(pq[m], strand == 0 ? pt[m] : pt[n - 1], pq[n - 1] + (uint32_t)K - pq[m], 0);
#undef PUSH
      const long b = e_off[c], cnt = no - b;
      if (cnt > 0) {
        uint32_t qs = eq[b], qe2 = qs + (uint32_t)elen[b], ts = et[b], te2 = ts + (uint32_t)elen[b];
        for (long i = b + 1; i < no; i++) {
          if (eq[i] < qs) qs = eq[i];
          if (eq[i] + (uint32_t)elen[i] > qe2) qe2 = eq[i] + (uint32_t)elen[i];
          if (et[i] < ts) ts = et[i];
          if (et[i] + (uint32_t)elen[i] > te2) te2 = et[i] + (uint32_t)elen[i];
        }
        box[4 * c] = qs; box[4 * c + 1] = qe2; box[4 * c + 2] = ts; box[4 * c + 3] = te2;
        if (trim) trim_group(eq + b, et + b, elen + b, cnt, strand, 40);
      }
    }
    /* MergeMatchesSameDiag */
    const long b = e_off[c], cnt = no - b;
    md_start[nmd] = 0; md_end[nmd] = 1; nmd++;
    if (cnt > 0) {
      const uint32_t *Q = eq + b, *T = et + b; const int32_t *L = elen + b; const uint8_t *O = eovp + b;
#define GDIAG(i) (strand == 0 ? (long)T[i] - (long)Q[i] : (long)Q[i] + (long)T[i] + (long)L[i])
      long prev_diag = GDIAG(0);
      uint32_t prev_qEnd = Q[0] + (uint32_t)L[0];
      for (long qi = 1; qi < cnt; qi++) {
        const long cur_diag = GDIAG(qi);
        long gd = (long)Q[qi] - ((long)Q[qi - 1] + (long)L[qi - 1]); if (gd < 0) gd = -gd;
        if (O[qi - 1] == 0 && O[qi] == 0 && prev_diag == cur_diag && prev_qEnd < Q[qi] && gd <= merge_dist) md_end[nmd - 1] = (int32_t)(qi + 1);
        else { md_start[nmd] = (int32_t)qi; md_end[nmd] = (int32_t)(qi + 1); nmd++; }
        prev_qEnd = Q[qi] + (uint32_t)L[qi];
        prev_diag = cur_diag;
      }
#undef GDIAG
    }
  }
  e_off[n_chain] = (int32_t)no;
  md_off[n_chain] = (int32_t)nmd;
  *overlap_count = overlap;
  return no;
}
