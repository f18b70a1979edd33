/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of SplitRoughClustersWithGaps (first half):
 *   SplitRoughClustersWithGaps   /root/reference/Clustering.h:1358-1430   cut a rough cluster where consecutive anchors (Cartesian order) are more than
 *                                                                          RoughClustermaxGap apart, keep pieces of >= minClusterSize anchors, re-join a
 *                                                                          piece to the previous one when it is close on one axis and on the diagonal
 *   CloseToPreviousCluster / MergeTwoClusters / UpdateCluster   Clustering.h:1332-1356
 * for all rough clusters of one anchor list, in order, into ONE vector of split clusters, as MatchesToFineClusters calls it (Clustering.h:1578-1581,
 * :1632-1635): the vector's size and its last element's chromIndex carry over from one rough cluster to the next.
 * The anchors of every rough cluster must already be in Cartesian order (CartesianSort, Clustering.h:1579).
 * Out: split clusters (start, end, box, strand, coarse = index of the rough cluster, anchorfreq, chromIndex -- -1 where the reference leaves the
 * constructor's value) and their splitmatchindex as pieces (cluster, first, last+1), in order.
 * Pinned by tests/test_split_rough.py against the unmodified reference (oracle/ref_wrap.cpp: ref_split_rough). */
#include <stdint.h>
#include <stdlib.h>

typedef struct { int start, end; uint32_t qs, qe, ts, te; int strand, coarse, chrom; float freq; } srcl;

static long labs_l(long x) { return x < 0 ? -x : x; }

long lra_oracle_split_rough(const uint32_t *q, const uint32_t *t, int n_rough, const int32_t *r_start, const int32_t *r_end, const uint32_t *r_box,
                            const uint8_t *r_strand, const float *r_freq, const int32_t *r_chrom, int globalK, int maxGap, int minClusterSize, int maxDiag,
                            int32_t *s_start, int32_t *s_end, uint32_t *s_box, uint8_t *s_strand, int32_t *s_coarse, float *s_freq, int32_t *s_chrom,
                            int32_t *p_cluster, int32_t *p_start, int32_t *p_end, int32_t *n_piece) {
  long total = 0;
  for (int c = 0; c < n_rough; c++) total += r_end[c] - r_start[c];
  srcl *S = (srcl *)malloc(((size_t)total + (size_t)n_rough + 1) * sizeof(srcl));
  long ns = 0, np = 0;
  const uint32_t K = (uint32_t)globalK;
#define PIECE(cl_, a_, b_) do { p_cluster[np] = (int32_t)(cl_); p_start[np] = (a_); p_end[np] = (b_); np++; } while (0)
#define CLOSE(a_, qS_, tS_, tE_, res_) do { \
    const long aDiff = labs_l((long)(qS_) - (long)(a_)->qe); \
    const long bDiff = (a_)->strand == 0 ? labs_l((long)(tS_) - (long)(a_)->te) : labs_l((long)(a_)->ts - (long)(tE_)); \
    long aDiag, bDiag; \
    if ((a_)->strand == 0) { aDiag = (long)(a_)->te - (long)(a_)->qe; bDiag = (long)(tS_) - (long)(qS_); } \
    else { aDiag = (long)(a_)->qe + (long)(a_)->ts; bDiag = (long)(qS_) + (long)(tE_); } \
    (res_) = ((aDiff < bDiff ? aDiff : bDiff) <= maxGap && labs_l(aDiag - bDiag) < maxDiag); } while (0)
#define MERGE(a_, qS_, qE_, tS_, tE_, st_, en_) do { \
    if ((qS_) < (a_)->qs) (a_)->qs = (qS_); if ((qE_) > (a_)->qe) (a_)->qe = (qE_); if ((tS_) < (a_)->ts) (a_)->ts = (tS_); if ((tE_) > (a_)->te) (a_)->te = (tE_); \
    PIECE(ns - 1, (st_), (en_)); (a_)->end = (en_); } while (0)
#define NEWCL(st_, en_, qS_, qE_, tS_, tE_, strand_, coarse_, freq_, chrom_) do { \
    srcl *x = &S[ns]; x->start = (st_); x->end = (en_); x->qs = (qS_); x->qe = (qE_); x->ts = (tS_); x->te = (tE_); x->strand = (strand_); x->coarse = (coarse_); \
    x->freq = (freq_); x->chrom = (chrom_); ns++; PIECE(ns - 1, (st_), (en_)); } while (0)
  for (int c = 0; c < n_rough; c++) {
    const int os = r_start[c], oe = r_end[c];
    if (oe - os == 0) continue;
    if (r_freq[c] >= 10.0f) { NEWCL(os, oe, r_box[4 * c], r_box[4 * c + 1], r_box[4 * c + 2], r_box[4 * c + 3], r_strand[c], c, r_freq[c], -1); continue; }
    const long cur_s = ns;
    int split_cs = os;
    uint32_t sqS = q[os], stS = t[os], sqE = sqS + K, stE = stS + K;
    for (int m = os + 1; m < oe; m++) {
      long ad = labs_l((long)q[m - 1] - (long)q[m]), bd = labs_l((long)t[m - 1] - (long)t[m]);      /* minGapDifference(matches[m], matches[m-1]) */
      const int gap = (int)(ad < bd ? ad : bd);
      if (gap > maxGap || (ns > 1 && S[ns - 1].chrom != r_chrom[c])) {
        if (m - split_cs >= minClusterSize) {
          int close = 0;
          if (ns > cur_s && S[ns - 1].chrom == r_chrom[c]) CLOSE(&S[ns - 1], sqS, stS, stE, close);
          if (close) MERGE(&S[ns - 1], sqS, sqE, stS, stE, split_cs, m);
          else NEWCL(split_cs, m, sqS, sqE, stS, stE, r_strand[c], c, r_freq[c], r_chrom[c]);
        }
        sqS = q[m]; stS = t[m]; sqE = sqS + K; stE = stS + K; split_cs = m;
      } else {
        if (q[m] < sqS) sqS = q[m]; if (t[m] < stS) stS = t[m];
        if (q[m] + K > sqE) sqE = q[m] + K; if (t[m] + K > stE) stE = t[m] + K;
      }
    }
    if (oe - split_cs >= minClusterSize) {
      int close = 0;
      if (ns > cur_s) CLOSE(&S[ns - 1], sqS, stS, stE, close);
      if (close) MERGE(&S[ns - 1], sqS, sqE, stS, stE, split_cs, oe);
      else NEWCL(split_cs, oe, sqS, sqE, stS, stE, r_strand[c], c, r_freq[c], -1);
    }
  }
  for (long i = 0; i < ns; i++) {
    s_start[i] = S[i].start; s_end[i] = S[i].end; s_box[4 * i] = S[i].qs; s_box[4 * i + 1] = S[i].qe; s_box[4 * i + 2] = S[i].ts; s_box[4 * i + 3] = S[i].te;
    s_strand[i] = (uint8_t)S[i].strand; s_coarse[i] = S[i].coarse; s_freq[i] = S[i].freq; s_chrom[i] = S[i].chrom;
  }
  *n_piece = (int32_t)np;
  free(S);
  return ns;
}
