/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of StoreDiagonalClusters (/root/reference/Clustering.h:1442-1487, with
 * RemoveSuperRepetitiveClusters :1432-1439 and DiagonalDifference :502-514): the clusters CleanMatches builds from the cleaned, diagonal-sorted
 * anchors of one strand when opts.ExtractDiagonalFromClean is off (Clustering.h:1861-1864, :1891-1894) -- runs of anchors whose consecutive diagonal
 * difference stays below opts.maxDiag, kept when they have minClusterSize anchors, span minClusterLength on both axes and are not one repeated
 * read k-mer.  anchorfreq = (sum of matches_freq since the last KEPT cluster) / size, accumulated in binary32 in anchor order.
 * Pinned by tests/test_store_diagonal.py against the unmodified reference (oracle/ref_wrap.cpp: ref_store_diagonal). */
#include <stdint.h>
#include <stdlib.h>

static int hdr_find(const uint64_t *pos, int n, uint64_t query) {   /* Header::Find, Genome.h:19-31 */
  if (n > 0 && query == pos[0]) return 0;
  int lo = 0, len = n;
  while (len > 0) { int half = len >> 1; if (pos[lo + half] < query) { lo += half + 1; len -= half + 1; } else len = half; }
  if (lo < n && query == pos[lo]) return lo;
  return lo - 1;
}

long lra_oracle_store_diagonal(const uint32_t *q, const uint32_t *t, const uint64_t *qt, const float *freq, int n, int strand, const uint64_t *hdr_pos, int n_hdr,
                               int globalK, int maxDiag, int minClusterSize, int minClusterLength, int bypass,
                               int32_t *c_start, int32_t *c_end, uint32_t *c_box, float *c_freq, int32_t *c_chrom) {
  long nc = 0;
  int cs = 0, ce;
  float totalfreq = 0.0f;
  const uint32_t K = (uint32_t)globalK;
  while (cs < n) {
    ce = cs + 1;
    uint32_t qS = q[cs], qE = q[cs] + K, tS = t[cs], tE = t[cs] + K;
    totalfreq += freq[cs];
    const int cI = hdr_find(hdr_pos, n_hdr, (uint64_t)tS);
    while (ce < n) {
      long d;
      if (strand == 0) d = ((long)t[ce] - (long)q[ce]) - ((long)t[ce - 1] - (long)q[ce - 1]);
      else d = (long)(uint32_t)(q[ce] + t[ce]) - (long)(uint32_t)(q[ce - 1] + t[ce - 1]);
      if (!((d < 0 ? -d : d) < maxDiag)) break;
      if (q[ce] < qS) qS = q[ce]; if (q[ce] + K > qE) qE = q[ce] + K; if (t[ce] < tS) tS = t[ce]; if (t[ce] + K > tE) tE = t[ce] + K;
      totalfreq += freq[ce];
      ce++;
    }
    int rep = 1;
    for (int m = cs + 1; m < ce; m++) if (qt[m] != qt[cs]) { rep = 0; break; }
    if (ce - cs >= minClusterSize && qE - qS >= (uint32_t)minClusterLength && tE - tS >= (uint32_t)minClusterLength && !rep) {
      c_start[nc] = cs; c_end[nc] = ce; c_box[4 * nc] = qS; c_box[4 * nc + 1] = qE; c_box[4 * nc + 2] = tS; c_box[4 * nc + 3] = tE;
      c_freq[nc] = totalfreq / (float)(ce - cs);
      c_chrom[nc] = bypass ? hdr_find(hdr_pos, n_hdr, (uint64_t)tS) : cI;
      totalfreq = 0.0f;
      nc++;
    }
    cs = ce;
  }
  return nc;
}
