/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of RefineBreakpoint, /root/reference/RefineBreakpoint.h:
 *   RSdp             :150-197   full DP (match 2, mismatch -2, indel -4) of the unaligned read span against the flanking genome, scores
 *                               and arrows kept for every cell; arrow priority diag, left, down
 *   FindMax          :199-210   first maximum in row-major order
 *   StoreQScoreVect  :120-147   score and matrix index of the path cell of every read column
 *   TraceBack        :92-118, PathToBlocks :50-83, PrependBlocks / AppendBlocks :6-47
 *   RefineBreakpoint :212-462   left / right alignment, forward / reverse strand: which side of which alignment is extended, reversed
 *                               strings for backward extensions, the split that maximises the summed scores when the two extensions overlap
 * An alignment enters through its first and last block only (GetQStart/GetQEnd/GetTStart/GetTEnd, Alignment.h:130-156, and the block
 * Prepend/Append may merge into).  Output per side: mode 0 (nothing), 1 (append), 2 (prepend); the blocks to splice in; the boundary block
 * of the alignment after a possible merge.
 * Pinned by tests/test_refine_breakpoint.py against the unmodified reference (oracle/ref_wrap.cpp: ref_refine_breakpoint). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void rsdp(const char *q, int qs, const char *t, int ts, int *path, int *score, int mat, int mis, int indel) {
  const int row = qs + 1;
  for (long i = 0; i < (long)(qs + 1) * (ts + 1); i++) { path[i] = -1; score[i] = 0; }
  for (int i = 1; i < qs + 1; i++) { path[i] = 1; score[i] = score[i - 1] + indel; }
  for (int i = 1; i < ts + 1; i++) { path[row * i] = 2; score[i * row] = score[(i - 1) * row] + indel; }
  for (int i = 0; i < ts; i++)
    for (int j = 0; j < qs; j++) {
      int diagScore = score[i * row + j] + (q[j] == t[i] ? mat : mis);
      int leftScore = score[(i + 1) * row + j] + indel;
      int downScore = score[i * row + (j + 1)] + indel;
      int m = diagScore > leftScore ? diagScore : leftScore; if (downScore > m) m = downScore;
      score[(i + 1) * row + (j + 1)] = m;
      path[(i + 1) * row + (j + 1)] = m == diagScore ? 3 : (m == leftScore ? 1 : 2);
    }
}
static int find_max(const int *score, long n, int row, int *q, int *t) {
  if (n == 0) { *q = *t = 0; return 0; }
  long best = 0;
  for (long i = 1; i < n; i++) if (score[i] > score[best]) best = i;
  *t = (int)(best / row) - 1; *q = (int)(best % row) - 1;
  return score[best];
}
static void store_qscore(const int *score, const int *path, int q, int t, int r, int *qv, int *index) {
  for (int i = 0; i < r - 1; i++) { qv[i] = 0; index[i] = 0; }
  long i = (long)(t + 1) * r + q + 1;
  q++; t++;
  while (i > 0) {
    if (path[i] == 3 || path[i] == 1) { qv[q - 1] = score[i]; index[q - 1] = (int)i; }
    if (path[i] == 3) { q--; t--; }
    if (path[i] == 1) q--;
    if (path[i] == 2) t--;
    i = (long)t * r + q;
  }
}
static int trace_back(const int *path, int q, int t, int r, int *tb) {
  int n = 0;
  q++; t++;
  long i = (long)t * r + q;
  while (q > 0 || t > 0) {
    if (path[i] == 3) { q--; t--; tb[n++] = 3; }
    if (path[i] == 1) { q--; tb[n++] = 1; }
    if (path[i] == 2) { t--; tb[n++] = 2; }
    i = (long)t * r + q;
  }
  for (int a = 0, b = n - 1; a < b; a++, b--) { int x = tb[a]; tb[a] = tb[b]; tb[b] = x; }
  return n;
}
static int path_to_blocks(const int *path, int n, int32_t *blocks) {     /* blocks: (q, t, len) */
  int i = 0, q = 0, t = 0, nb = 0;
  while (i < n && path[i] != 3 && (path[i] == 1 || path[i] == 2)) { if (path[i] == 1) q++; if (path[i] == 2) t++; i++; }
  while (i < n) {
    int qs = q, ts = t;
    while (i < n && path[i] == 3) { q++; t++; i++; }
    while (i < n && (path[i] == 1 || path[i] == 2)) { if (path[i] == 1) q++; if (path[i] == 2) t++; i++; }
    int match = (q - qs) < (t - ts) ? (q - qs) : (t - ts);
    if (match > 0) { blocks[3 * nb] = qs; blocks[3 * nb + 1] = ts; blocks[3 * nb + 2] = match; nb++; }
  }
  return nb;
}
static void rev(char *s, int n) { for (int a = 0, b = n - 1; a < b; a++, b--) { char x = s[a]; s[a] = s[b]; s[b] = x; } }

/* splice `src` (n blocks, already offset) onto an alignment whose boundary block is `bound`: mode 1 append, 2 prepend */
static int splice(int mode, int32_t *src, int n, uint32_t *bound, uint32_t *out) {
  if (n == 0) return 0;
  int s0 = 0, s1 = n;
  if (mode == 1) {
    if (bound[1] + bound[2] == (uint32_t)src[1] && bound[0] + bound[2] == (uint32_t)src[0]) { bound[2] += (uint32_t)src[2]; s0 = 1; }
  } else {
    const int last = n - 1;
    if ((uint32_t)src[3 * last + 1] + (uint32_t)src[3 * last + 2] == bound[1] && (uint32_t)src[3 * last] + (uint32_t)src[3 * last + 2] == bound[0]) {
      bound[1] -= (uint32_t)src[3 * last + 2]; bound[0] -= (uint32_t)src[3 * last + 2]; bound[2] += (uint32_t)src[3 * last + 2]; s1 = last;
    }
  }
  int k = 0;
  for (int i = s0; i < s1; i++, k++) { out[3 * k] = (uint32_t)src[3 * i]; out[3 * k + 1] = (uint32_t)src[3 * i + 1]; out[3 * k + 2] = (uint32_t)src[3 * i + 2]; }
  return k;
}

/* lf / ll: first and last block (q, t, len) of the left alignment, rf / rl of the right one; lread / rread: the read strand each alignment is
 * on (Alignment::read); lchrom / rchrom: their contigs.  mode[2], n_out[2], bound[2][3] (the boundary block after the splice: the last block
 * for an append, the first for a prepend), out[2][cap][3].  Returns 1 if the breakpoint was refined, 0 if the two alignments are not within
 * MAX_GAP of each other. */
int lra_oracle_refine_breakpoint(const char *lread, const char *rread, int readLen, const char *lchrom, int lchromLen, const char *rchrom, int rchromLen,
                                 const uint32_t *lf, const uint32_t *ll, int lstrand, const uint32_t *rf, const uint32_t *rl, int rstrand,
                                 int32_t *mode, int32_t *n_out, uint32_t *bound, uint32_t *out, int cap) {
  mode[0] = mode[1] = 0; n_out[0] = n_out[1] = 0;
  const int lqs = (int)lf[0], lqe = (int)(ll[0] + ll[2]), lts = (int)lf[1], lte = (int)(ll[1] + ll[2]);
  const int rqs = (int)rf[0], rqe = (int)(rl[0] + rl[2]), rts = (int)rf[1], rte = (int)(rl[1] + rl[2]);
  int flqe, frqs;
  if (lstrand == 0) flqe = lqe; else flqe = readLen - lqs;
  if (rstrand == 0) frqs = rqs; else frqs = readLen - rqe;
  const int MAX_GAP = 500;
  if (!(frqs > flqe && frqs - flqe < MAX_GAP)) return 0;
  const int span = frqs - flqe;
  char *lq = (char *)malloc((size_t)span + 1), *lt = (char *)malloc((size_t)span + 1), *rq = (char *)malloc((size_t)span + 1), *rt = (char *)malloc((size_t)span + 1);
  int ltLen, rtLen, lPrefix = 0, rPrefix = 0;
  if (lstrand == 0) {
    memcpy(lq, lread + lqe, (size_t)span);
    int tSpan = lchromLen - lte < span ? lchromLen - lte : span;
    ltLen = tSpan; memcpy(lt, lchrom + lte, (size_t)tSpan);
  } else {
    memcpy(lq, lread + (lqs - span), (size_t)span);
    int ltExtEnd = lts, ltExtStart = ltExtEnd - span > 0 ? ltExtEnd - span : 0;
    ltLen = ltExtEnd - ltExtStart; memcpy(lt, lchrom + ltExtStart, (size_t)ltLen);
    lPrefix = 1; rev(lq, span); rev(lt, ltLen);
  }
  const long lcells = (long)(span + 1) * (ltLen + 1);
  int *lPath = (int *)malloc(sizeof(int) * (size_t)lcells), *lScore = (int *)malloc(sizeof(int) * (size_t)lcells);
  rsdp(lq, span, lt, ltLen, lPath, lScore, 2, -2, -4);
  if (rstrand == 0) {
    memcpy(rq, rread + (rqs - span), (size_t)span);
    int rtSpan = rts < span ? rts : span;
    rtLen = rtSpan; memcpy(rt, rchrom + (rts - rtSpan), (size_t)rtSpan);
    rev(rq, span); rev(rt, rtLen); rPrefix = 1;
  } else {
    memcpy(rq, rread + rqe, (size_t)span);
    int tSpan = span;
    if (rte + span >= rchromLen) tSpan = rchromLen - rte;
    rtLen = tSpan; memcpy(rt, rchrom + rte, (size_t)tSpan);
  }
  const long rcells = (long)(span + 1) * (rtLen + 1);
  int *rPath = (int *)malloc(sizeof(int) * (size_t)rcells), *rScore = (int *)malloc(sizeof(int) * (size_t)rcells);
  rsdp(rq, span, rt, rtLen, rPath, rScore, 2, -2, -4);
  int mlq, mlt, mrq, mrt;
  find_max(lScore, lcells, span + 1, &mlq, &mlt);
  find_max(rScore, rcells, span + 1, &mrq, &mrt);
  if (!(mlq < span - mrq)) {
    int *lqS = (int *)malloc(sizeof(int) * (size_t)span * 4), *rqS = lqS + span, *lqI = rqS + span, *rqI = lqI + span;
    store_qscore(lScore, lPath, mlq, mlt, span + 1, lqS, lqI);
    store_qscore(rScore, rPath, mrq, mrt, span + 1, rqS, rqI);
    int maxScore = 0, maxL = 0, maxR = 0;
    for (int i = 0; i < span; i++)
      if (lqS[i] + rqS[span - i - 1] > maxScore) { maxScore = lqS[i] + rqS[span - i - 1]; maxL = i; maxR = span - i - 1; }
    mlq = maxL; mlt = lqI[maxL] / (span + 1) - 1; mrq = maxR; mrt = rqI[maxR] / (span + 1) - 1;
    free(lqS);
  }
  int *ltb = (int *)malloc(sizeof(int) * (size_t)(2 * span + 4)), *rtb = (int *)malloc(sizeof(int) * (size_t)(2 * span + 4));
  int nl = trace_back(lPath, mlq, mlt, span + 1, ltb), nr = trace_back(rPath, mrq, mrt, span + 1, rtb);
  int32_t *lB = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)(span + 2)), *rB = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)(span + 2));
  int lqBlockStart, ltBlockStart, rqBlockStart, rtBlockStart;
  if (lPrefix) { for (int a = 0, b = nl - 1; a < b; a++, b--) { int x = ltb[a]; ltb[a] = ltb[b]; ltb[b] = x; } lqBlockStart = lqs - mlq - 1; ltBlockStart = lts - mlt - 1; }
  else { lqBlockStart = lqe; ltBlockStart = lte; }
  int nlb = path_to_blocks(ltb, nl, lB);
  for (int i = 0; i < nlb; i++) { lB[3 * i] += lqBlockStart; lB[3 * i + 1] += ltBlockStart; }
  if (rPrefix) { for (int a = 0, b = nr - 1; a < b; a++, b--) { int x = rtb[a]; rtb[a] = rtb[b]; rtb[b] = x; } rqBlockStart = rqs - mrq - 1; rtBlockStart = rts - mrt - 1; }
  else { rqBlockStart = rqe; rtBlockStart = rte; }
  int nrb = path_to_blocks(rtb, nr, rB);
  for (int i = 0; i < nrb; i++) { rB[3 * i] += rqBlockStart; rB[3 * i + 1] += rtBlockStart; }
  mode[0] = lPrefix ? 2 : 1; mode[1] = rPrefix ? 2 : 1;
  memcpy(bound, lPrefix ? lf : ll, 12); memcpy(bound + 3, rPrefix ? rf : rl, 12);
  if (nlb > cap || nrb > cap) { n_out[0] = nlb; n_out[1] = nrb; mode[0] = mode[1] = -1; }
  else { n_out[0] = splice(mode[0], lB, nlb, bound, out); n_out[1] = splice(mode[1], rB, nrb, bound + 3, out + 3 * (size_t)cap); }
  free(lq); free(lt); free(rq); free(rt); free(lPath); free(lScore); free(rPath); free(rScore); free(ltb); free(rtb); free(lB); free(rB);
  return 1;
}
