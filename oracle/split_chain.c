/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the low-accuracy pipeline's chain splitting:
 *   SPLITChain (UltimateChain overload)  /root/reference/Mapping_ultility.h:380-437   cut the anchor chain at unchained gaps (>= 1000 on both axes along
 *                                                                                     one diagonal), translocations (opts.splitdist) and strand switches
 *   push_new                             Mapping_ultility.h:355-378                   build one SplitChain (ClusterIndex run-length list, box), drop it
 *                                                                                     when it spans two contigs (SplitChain::CHROMIndex, Chain.h:388-396)
 *   MergeSplitchainINS                   Mapping_ultility.h:163-264                   re-join the two sides of an insertion-like 'T' piece
 *   RemoveSpuriousSplitChain             Map_lowacc.h:38-66                           drop tiny pieces
 * as Map_lowacc.h:261-262 calls them for every UltimateChain.  The chain enters as its anchors in chain order: q, t (global genome
 * coordinate), len, strand and ClusterNum of each, link[i] between anchors i and i+1.
 * Out: the split chains in order; piece s owns entries sp_off[s] .. sp_off[s+1] of sptc (anchor indices, already reversed for forward pieces) with
 * sp_lk[sp_off[s] + j] = its link j (size - 1 of them), ClusterIndex entries ci_off[s] .. ci_off[s+1] of ci, sp_box[4s..] = QStart, QEnd, TStart,
 * TEnd, sp_chrom, sp_type ('N' 'T' 'I'), sp_strand; and spchain_link as sp_link[0 .. *n_link).
 * vector<bool> sizes are tracked as the reference leaves them (splitchains_link can be one longer than pieces - 1 when the last piece was
 * dropped; MergeSplitchainINS only resizes it).  GenomePos arithmetic is uint32 arithmetic.
 * Pinned by tests/test_split_chain.py against the unmodified reference (oracle/ref_wrap.cpp: ref_split_chain). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int *sptc; uint8_t *link; int size;          /* link has size - 1 entries */
  int *ci; int nci;
  uint32_t QS, QE, TS, TE;
  int chrom; char type; int strand;
} spc;

static int hdr_find(const uint64_t *pos, int n, uint64_t query) {   /* Header::Find, Genome.h:19-31 */
  if (n > 0 && query == pos[0]) return 0;
  int lo = 0, len = n;
  while (len > 0) { int half = len >> 1; if (pos[lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
  if (lo < n && query == pos[lo]) return lo;
  return lo - 1;
}

long lra_oracle_split_chain(const uint32_t *q, const uint32_t *t, const int32_t *len, const uint8_t *strand, const int32_t *cnum, const uint8_t *link, int n,
                            const uint64_t *hdr_pos, int n_hdr, int splitdist, int bypass,
                            int32_t *sp_off, int32_t *sptc, uint8_t *sp_lk, int32_t *ci_off, int32_t *ci, uint32_t *sp_box, int32_t *sp_chrom, uint8_t *sp_type,
                            uint8_t *sp_strand, uint8_t *sp_link, int32_t *n_link) {
  spc *S = (spc *)calloc((size_t)n + 1, sizeof(spc));
  uint8_t *SL = (uint8_t *)calloc((size_t)n + 2, 1);    /* splitchains_link */
  int ns = 0, nsl = 0;
  int *onec = (int *)malloc(((size_t)n + 1) * sizeof(int)); uint8_t *lk = (uint8_t *)malloc((size_t)n + 1);
  int no = 0, nlk = 0;
#define QSTART(i) (q[i])
#define TSTART(i) (t[i])
#define QEND(i) (q[i] + (uint32_t)len[i])
#define TEND(i) (t[i] + (uint32_t)len[i])
#define DIAG(i) (strand[i] == 1 ? (long)QEND(i) + (long)TSTART(i) : (long)TSTART(i) - (long)QSTART(i))
  /* push_new: returns 1 if the piece was kept */
#define PUSH_NEW(cur_, kept_) do { \
    spc *x = &S[ns]; memset(x, 0, sizeof *x); \
    x->size = no; x->sptc = (int *)malloc(((size_t)no + 1) * sizeof(int)); x->link = (uint8_t *)malloc((size_t)no + 1); \
    memcpy(x->sptc, onec, (size_t)no * sizeof(int)); memcpy(x->link, lk, (size_t)nlk); \
    x->strand = strand[onec[0]]; x->type = 'N'; \
    x->ci = (int *)malloc(((size_t)no + 1) * sizeof(int)); x->nci = 0; x->ci[x->nci++] = cnum[onec[0]]; \
    for (int c_ = 1; c_ < no; c_++) if (cnum[onec[c_]] != x->ci[x->nci - 1]) x->ci[x->nci++] = cnum[onec[c_]]; \
    x->QS = QSTART(onec[no - 1]); x->QE = QEND(onec[0]); \
    if (strand[onec[0]] == 0) { x->TS = TSTART(onec[no - 1]); x->TE = TEND(onec[0]); } else { x->TS = TSTART(onec[0]); x->TE = TEND(onec[no - 1]); } \
    const int f_ = hdr_find(hdr_pos, n_hdr, (uint64_t)(uint32_t)(x->TS + 1u)), l_ = hdr_find(hdr_pos, n_hdr, (uint64_t)x->TE); \
    if (f_ != l_) { free(x->sptc); free(x->link); free(x->ci); (kept_) = 0; } else { x->chrom = f_; ns++; (kept_) = 1; } \
    no = 0; nlk = 0; onec[no++] = (cur_); \
  } while (0)
  if (n > 0) {
    onec[no++] = 0;
    int im = 0, cur = 0, prev = 0;
    while (im < n - 1) {
      cur = im + 1; prev = im;
      const int qdist = (int)(QSTART(prev) - QEND(cur));
      const int tdist = (TSTART(prev) > TEND(cur)) ? (int)(TSTART(prev) - TEND(cur)) : (int)(TEND(cur) - TSTART(prev));
      const int dist = qdist < tdist ? qdist : tdist;
      int kept;
      if (strand[cur] == strand[prev] && dist >= 1000 && (double)labs(DIAG(cur) - DIAG(prev)) <= ceil(0.15 * dist)) {
        PUSH_NEW(cur, kept);
        if (kept) { SL[nsl++] = 0; S[ns - 1].type = 'N'; }
      } else if (TSTART(cur) > TEND(prev) + (uint32_t)splitdist || TEND(cur) + (uint32_t)splitdist < TSTART(prev)) {
        PUSH_NEW(cur, kept);
        if (kept) { SL[nsl++] = 0; S[ns - 1].type = 'T'; }
      } else if (strand[cur] != strand[prev]) {
        PUSH_NEW(cur, kept);
        if (kept) { S[ns - 1].type = 'I'; SL[nsl++] = 1; }
      } else { onec[no++] = cur; lk[nlk++] = link[im]; }
      im++;
    }
    if (no > 0) { int kept; PUSH_NEW(cur, kept); (void)kept; }
  }
  /* MergeSplitchainINS */
  if (ns >= 3) {
    int *cur_ind = (int *)malloc((size_t)ns * sizeof(int)); uint8_t *keep = (uint8_t *)malloc((size_t)ns);
    for (int i = 0; i < ns; i++) { cur_ind[i] = i; keep[i] = 1; }
    int change = 0, im = 0;
    while (im <= ns - 3) {
      const int c = cur_ind[im];
      if (S[c].type != 'T') { im++; continue; }
      int nn = cur_ind[im + 2];
      while (nn < ns) {
        const long tdist = (S[c].TS > S[nn].TE) ? ((long)S[c].TS - (long)S[nn].TE) : ((long)S[nn].TE - (long)S[c].TS);
        if (tdist > 1500) { nn++; continue; }
        if (S[c].strand != S[nn].strand) { nn++; continue; }
        if (S[c].chrom != S[nn].chrom) { nn++; continue; }
        change = 1;
        const int t1 = S[c].size, tt = t1 + S[nn].size;
        S[c].sptc = (int *)realloc(S[c].sptc, ((size_t)tt + 1) * sizeof(int)); S[c].link = (uint8_t *)realloc(S[c].link, (size_t)tt + 1);
        for (int s = t1; s < tt; s++) {
          S[c].sptc[s] = S[nn].sptc[s - t1];
          if (s == t1) S[c].link[s - 1] = 0; else S[c].link[s - 1] = S[nn].link[s - t1 - 1];
          if (S[nn].QS < S[c].QS) S[c].QS = S[nn].QS;
          if (S[nn].TS < S[c].TS) S[c].TS = S[nn].TS;
          if (S[nn].QE > S[c].QE) S[c].QE = S[nn].QE;
          if (S[nn].TE > S[c].TE) S[c].TE = S[nn].TE;
          S[c].type = S[nn].type;
        }
        S[c].size = tt;
        if (bypass) {
          S[c].ci = (int *)realloc(S[c].ci, ((size_t)S[c].nci + (size_t)S[nn].nci + 1) * sizeof(int));
          int pv = S[c].ci[S[c].nci - 1];
          for (int s = 0; s < S[nn].nci; s++) { const int cu = S[nn].ci[s]; if (pv != cu) { S[c].ci[S[c].nci++] = cu; pv = cu; } }
        }
        cur_ind[nn] = cur_ind[c];
        keep[nn] = 0;
        break;
      }
      im = nn;
    }
    if (change) {
      int r = 0;
      for (int s = 0; s < ns; s++) {
        if (keep[s]) { if (r != s) { S[r] = S[s]; memset(&S[s], 0, sizeof(spc)); } r++; }
        else { free(S[s].sptc); free(S[s].link); free(S[s].ci); memset(&S[s], 0, sizeof(spc)); }
      }
      ns = r;
      for (int i = nsl; i < r - 1; i++) SL[i] = 0;      /* vector<bool>::resize grows with false */
      nsl = r - 1;
      if (bypass) for (int i = 1; i < ns; i++) SL[i - 1] = S[i].type == 'I' ? 1 : 0;
    }
    free(cur_ind); free(keep);
  }
  /* the pieces of forward strand are reversed for refining */
  for (int s = 0; s < ns; s++) if (S[s].strand == 0) {
    for (int i = 0, j = S[s].size - 1; i < j; i++, j--) { int x = S[s].sptc[i]; S[s].sptc[i] = S[s].sptc[j]; S[s].sptc[j] = x; }
    for (int i = 0, j = S[s].size - 2; i < j; i++, j--) { uint8_t x = S[s].link[i]; S[s].link[i] = S[s].link[j]; S[s].link[j] = x; }
  }
  /* RemoveSpuriousSplitChain */
  {
    int total = 0;
    for (int i = 0; i < ns; i++) total += S[i].size;
    int filter = (int)floorf(0.02f * (float)total); if (filter < 2) filter = 2;
    int filter2 = (int)floorf(0.03f * (float)total); if (filter2 < 2) filter2 = 2;
    const int f1 = filter < 2 ? filter : 2, f2 = filter2 < 4 ? filter2 : 4;
    uint8_t *rem = (uint8_t *)calloc((size_t)ns + 1, 1);
    for (int i = 0; i < ns; i++) {
      if (S[i].size < f1) rem[i] = 1;
      if (i > 0 && SL[i - 1] == 1 && S[i].size < f2) rem[i] = 1;
    }
    int c = 0;
    for (int i = 0; i < ns; i++) {
      if (!rem[i]) {
        if (c != i) { free(S[c].sptc); free(S[c].link); free(S[c].ci); S[c] = S[i]; memset(&S[i], 0, sizeof(spc)); }
        if (c > 1) SL[c - 1] = SL[i - 1];
        c++;
      }
    }
    for (int i = c; i < ns; i++) { free(S[i].sptc); free(S[i].link); free(S[i].ci); memset(&S[i], 0, sizeof(spc)); }
    ns = c;
    if (c > 1) { for (int i = nsl; i < c - 1; i++) SL[i] = 0; nsl = c - 1; } else nsl = 0;
    free(rem);
  }
  int o = 0, oc = 0;
  for (int s = 0; s < ns; s++) {
    sp_off[s] = o; ci_off[s] = oc;
    for (int i = 0; i < S[s].size; i++) { sptc[o + i] = S[s].sptc[i]; sp_lk[o + i] = i + 1 < S[s].size ? S[s].link[i] : 0; }
    o += S[s].size;
    for (int i = 0; i < S[s].nci; i++) ci[oc++] = S[s].ci[i];
    sp_box[4 * s] = S[s].QS; sp_box[4 * s + 1] = S[s].QE; sp_box[4 * s + 2] = S[s].TS; sp_box[4 * s + 3] = S[s].TE;
    sp_chrom[s] = S[s].chrom; sp_type[s] = (uint8_t)S[s].type; sp_strand[s] = (uint8_t)S[s].strand;
    free(S[s].sptc); free(S[s].link); free(S[s].ci);
  }
  sp_off[ns] = o; ci_off[ns] = oc;
  for (int i = 0; i < nsl; i++) sp_link[i] = SL[i];
  *n_link = nsl;
  free(S); free(SL); free(onec); free(lk);
  return ns;
}
