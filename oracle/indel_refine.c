/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's banded affine-gap refinement of a segment's
 * block list.  Follows  IndelRefineAlignment()  /root/reference/IndelRefine.h:53-784.
 * Pinned by tests/test_oracle_indel_refine.py against calls captured from the unmodified reference
 * (oracle/lra_capture.cpp, tests/golden/ir_*.bin).
 *
 * Steps kept (all observable in the output blocks):
 *   1. optional end padding (endAlign, :89-130)
 *   2. grouping of consecutive blocks whose gaps are < k-1 on both axes (:132-162), trimming of long first / last
 *      blocks to k-1 bases inside the window (:178-211)
 *   3. per group: ragged band qS[t]..qE[t] = envelope of the old path +-k, made monotone (:220-333)
 *   4. tiny windows -> AffineOneGapAlign (:344-357); otherwise a 3-state DP (match / ins / del) with
 *      gap = indel, gapOpen = 2*indel+1, gapExtend = 0 over the band (:383-622), traceback (:626-674), path -> blocks
 *      (:702-745).  Characters are compared raw (the reference upper-cases reads and genome at input).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int lra_oracle_aog_ws(void *wsp, const char *q, int qLen, const char *t, int tLen, int m, int mm, int indel, int k,
                      uint32_t *blocks, int cap, int *n_blocks, int *status);
void *lra_oracle_aog_ws_new(void);
void lra_oracle_aog_ws_free(void *p);

typedef struct { uint32_t q, t, len; } blk_t;
typedef struct { blk_t *v; int n, cap; } blkvec;
static void bv_push(blkvec *b, uint32_t q, uint32_t t, uint32_t len) {
  if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 64; b->v = (blk_t *)realloc(b->v, (size_t)b->cap * sizeof(blk_t)); }
  b->v[b->n].q = q; b->v[b->n].t = t; b->v[b->n].len = len; b->n++;
}
static inline long lmin(long a, long b) { return a < b ? a : b; }
static inline long lmax(long a, long b) { return a > b ? a : b; }
static inline int imax2(int a, int b) { return a > b ? a : b; }

enum { P_DIAG = 0, P_LEFT = 1, P_DOWN = 2, P_BOUND = 3, P_DELOPEN = 4, P_DELEXT = 5, P_DELCLOSE = 6, P_INSOPEN = 7,
       P_INSEXT = 8, P_INSCLOSE = 9, P_DONE = 20 };
#define IR_BAD (-999999999)

/* One banded group.  qS/qE are absolute read coordinates per target row (tLen rows).  Emits blocks through bv_push.
 * Exposed separately (lra_oracle_indel_dp) so that the DP kernel can be tested on its own inputs. */
static int indel_dp(const char *qSeq, const char *tSeqBase, long tWinOff, long qStart, long tStart, long tLen, const int *qS,
                    const int *qE, long matSize, int gap, int match, int mismatch, blkvec *out, long qSeqLen, long tSeqLen) {
  const int gapOpen = gap * 2 + 1, gapExtend = 0;
  int st = 0;
  int *score = (int *)malloc(sizeof(int) * (size_t)matSize * 9);
  int *path = score + matSize, *index = path + matSize;
  int *dScore = index + matSize, *dPath = dScore + matSize, *dIndex = dPath + matSize;
  int *iScore = dIndex + matSize, *iPath = iScore + matSize, *iIndex = iPath + matSize;
  for (long x = 0; x < matSize; x++) {
    score[x] = 0; path[x] = P_BOUND; index[x] = -1;
    dScore[x] = IR_BAD; dPath[x] = P_BOUND; dIndex[x] = -1;
    iScore[x] = IR_BAD; iPath[x] = P_BOUND; iIndex[x] = -1;
  }
  index[0] = 0; path[0] = P_DONE;
  {
    int rowStart = 0, rowEnd;
    for (long ti = 0; ti < tLen; ti++) {
      int rowLen = qE[ti] - qS[ti] + 1;
      rowEnd = rowStart + rowLen - 1;
      if (rowStart > 0) { score[rowStart] = IR_BAD; path[rowStart] = P_BOUND; iPath[rowStart] = P_BOUND; }
      else for (int qi = 1; qi < rowEnd; qi++) { score[qi] = score[qi - 1] + gap; path[qi] = P_LEFT; index[qi] = qi - 1; }
      if (ti < tLen - 1) { score[rowEnd] = IR_BAD; path[rowEnd] = P_BOUND; }
      rowStart += rowLen;
    }
  }
  int curRowStart = qE[0] - qS[0] + 1, prevRowStart = 0, prevRowLen = qE[0] - qS[0] + 1;
  for (long ti = 1; ti < tLen; ti++) {
    int curRowLen = qE[ti] - qS[ti] + 1;
    int curRowOffset = qS[ti] - qS[ti - 1];
    int curRowPos = curRowStart + 1;
    int prevRowPos = prevRowStart + curRowOffset + 1;
    int rowEnd = (ti == tLen - 1) ? curRowLen : curRowLen - 1;
    char tChar = tSeqBase[ti + tStart - tWinOff];
    int qEPrev = qE[ti - 1], qSCur = qS[ti];
    for (int qi = 1; qi < rowEnd; qi++, curRowPos++, prevRowPos++) {
      int upOk = (qEPrev >= qi + qSCur) && (path[prevRowPos] != P_BOUND);
      int delOpenScore = upOk ? score[prevRowPos] + gapOpen : IR_BAD;
      int delExtendScore = upOk ? dScore[prevRowPos] + gapExtend : IR_BAD;
      int mx = delOpenScore > delExtendScore ? delOpenScore : delExtendScore;
      dPath[curRowPos] = (mx == delOpenScore) ? P_DELOPEN : P_DELEXT;
      dIndex[curRowPos] = prevRowPos;
      dScore[curRowPos] = mx;
      int insOpenScore = score[curRowPos - 1] + gapOpen;
      int insExtendScore = iScore[curRowPos - 1] + gapExtend;
      mx = insOpenScore > insExtendScore ? insOpenScore : insExtendScore;
      iPath[curRowPos] = (mx == insOpenScore) ? P_INSOPEN : P_INSEXT;
      iIndex[curRowPos] = curRowPos - 1;
      iScore[curRowPos] = mx;
      int matchScore;
      if ((qEPrev >= qi + qSCur) && path[prevRowPos - 1] != P_BOUND)
        matchScore = score[prevRowPos - 1] + (tChar == qSeq[qi + qSCur] ? match : mismatch);
      else matchScore = IR_BAD;
      int insScore = score[curRowPos - 1] + gap;
      int delScore = upOk ? score[prevRowPos] + gap : IR_BAD;
      int delCloseScore = dScore[curRowPos], insCloseScore = iScore[curRowPos];
      mx = matchScore;
      if (insScore > mx) mx = insScore;
      if (delScore > mx) mx = delScore;
      if (delCloseScore > mx) mx = delCloseScore;
      if (insCloseScore > mx) mx = insCloseScore;
      score[curRowPos] = mx;
      if (mx == matchScore) { path[curRowPos] = P_DIAG; index[curRowPos] = prevRowPos - 1; }
      else if (mx == insScore) { path[curRowPos] = P_LEFT; index[curRowPos] = curRowPos - 1; }
      else if (mx == delScore) { path[curRowPos] = P_DOWN; index[curRowPos] = prevRowPos; }
      else if (mx == delCloseScore) { path[curRowPos] = P_DELCLOSE; index[curRowPos] = curRowPos; }
      else if (mx == insCloseScore) { path[curRowPos] = P_INSCLOSE; index[curRowPos] = curRowPos; }
    }
    prevRowStart += prevRowLen; curRowStart += curRowLen; prevRowLen = curRowLen;
  }
  /* traceback (:626-674) */
  long pcap = 1024, pn = 0;
  int *pth = (int *)malloc(sizeof(int) * pcap);
#define PPUSH(v) do { if (pn == pcap) { pcap *= 2; pth = (int *)realloc(pth, sizeof(int) * pcap); } pth[pn++] = (v); } while (0)
  int curMat = 0;
  long pos = matSize - 1, guard = 0;
  while (pos > 0) {
    if (++guard > 4 * matSize + 16) { st = 1; break; }
    if (curMat == 0) {
      if (path[pos] == P_DELCLOSE) curMat = 1;
      else if (path[pos] == P_INSCLOSE) curMat = 2;
      else PPUSH(path[pos]);
      pos = index[pos];
    } else if (curMat == 1) {
      PPUSH(P_DOWN);
      curMat = (dPath[pos] == P_DELOPEN) ? 0 : 1;
      pos = dIndex[pos];
    } else {
      PPUSH(P_LEFT);
      curMat = (iPath[pos] == P_INSOPEN) ? 0 : 2;
      pos = iIndex[pos];
    }
    if (pos < 0) { st = 1; break; }
  }
  PPUSH(P_DIAG);
  /* reversed path -> blocks (:702-745) */
  long qPath = qStart, tPath = tStart;
  long pi = pn - 1;
  while (pi >= 0) {
    long blockLen = 0, tg = 0, qg = 0;
    while (pi >= 0 && pth[pi] == P_DIAG) { blockLen++; pi--; }
    if (pi >= 0) {
      if (pth[pi] == P_LEFT) while (pi >= 0 && pth[pi] == P_LEFT) { qg++; pi--; }
      else if (pth[pi] == P_DOWN) while (pi >= 0 && pth[pi] == P_DOWN) { tg++; pi--; }
      else { st = 1; break; } /* reference would loop forever on an unexpected arrow */
    }
    bv_push(out, (uint32_t)qPath, (uint32_t)tPath, (uint32_t)blockLen);
    qPath += blockLen + qg; tPath += blockLen + tg;
  }
  if (qPath != qSeqLen + qStart || tPath != tSeqLen + tStart) st |= 4;
  free(pth); free(score);
  return st;
}

/* Band construction for blocks [startBlock, endBlock] (:220-333).  Returns matSize; qS/qE must hold tLen ints. */
static long build_band(const blk_t *b, int startBlock, int endBlock, int k, long qStart, long qEnd, long tLen, int *qS, int *qE) {
  for (long x = 0; x < tLen; x++) { qS[x] = -1; qE[x] = -1; }
  long t = b[startBlock].t, q = b[startBlock].q;
  long tOff = 0;
  (void)t;
  for (int bb = startBlock; bb <= endBlock; bb++) {
    int qGap = 0, tGap = 0;
    int blockLength = (int)b[bb].len;
    if (bb < endBlock) {
      qGap = (int)(b[bb + 1].q - (b[bb].q + blockLength));
      tGap = (int)(b[bb + 1].t - (b[bb].t + blockLength));
      if (qGap > 0 && tGap > 0) { int c = qGap < tGap ? qGap : tGap; qGap -= c; tGap -= c; blockLength += c; }
    }
    for (int bi = 0; bi < blockLength; tOff++, bi++, q++) {
      if (qS[tOff] == -1) qS[tOff] = (int)lmax(q - k, qStart);
      else qS[tOff] = (int)lmin((long)qS[tOff], lmax(q - k, qStart));
      if (qE[tOff] == -1 || qE[tOff] < q + k) qE[tOff] = (int)lmin(qEnd - 1, q + k);
      for (int ki = 0; ki < k; ki++) {
        if (tOff - ki >= 0) { if (qE[tOff - ki] < q) qE[tOff - ki] = (int)q; }
        if (tOff + ki < tLen) { if (qS[tOff + ki] == -1 || qS[tOff + ki] > q) qS[tOff + ki] = (int)q; }
      }
    }
    if (qGap > tGap) {
      for (int qi = 0; qi < qGap; qi++, q++)
        for (int ki = 0; ki < k; ki++) {
          if (tOff - ki >= 0 && tOff - ki < tLen) { if (qE[tOff - ki] < q) qE[tOff - ki] = (int)q; }
          if (tOff + ki < tLen) { if (qS[tOff + ki] == 0 || qS[tOff + ki] > q) qS[tOff + ki] = (int)q; }
        }
    }
    if (tGap > qGap) {
      for (int ti = 0; ti < tGap; tOff++, ti++) { qS[tOff] = (int)lmax(q - k, qStart); qE[tOff] = (int)lmin(qEnd - 1, q + k); }
    }
  }
  for (long qi = tLen; qi > 1; qi--) if (qS[qi - 1] < qS[qi - 2]) qS[qi - 2] = qS[qi - 1];
  for (long qi = 0; qi < tLen - 1; qi++) if (qE[qi] > qE[qi + 1]) qE[qi + 1] = qE[qi];
  long matSize = 0;
  for (long qi = 0; qi < tLen; qi++) matSize += qE[qi] - qS[qi] + 1;
  return matSize;
}

/* Whole function.  tSeq[i] (contig coordinate i) is twin[i - tWinOff].  Returns 0; *status bit0 = traceback problem,
 * bit1 = output inconsistent ("ERROR with alignment consistency" in the reference), bit2 = path end mismatch. */
/* Optional dump of the banded (DP) groups of a segment, for testing the DP kernel on its own inputs. */
typedef struct {
  int32_t *meta;      /* per group 8 ints: qStart, tStart, tLen, qSeqLen, tSeqLen, bandOff, firstOutBlock, nOutBlocks */
  int max_groups, n_groups;
  int32_t *band;      /* qS rows then qE rows of each group, concatenated: band[bandOff .. +tLen) = qS, [+tLen .. +2tLen) = qE */
  long band_cap, band_used;
} ir_dump;
static ir_dump *g_dump = NULL;

int lra_oracle_indel_refine(const char *qSeq, int readLen, const char *twin, long tWinOff, long contigLen,
                            const uint32_t *blocks_in, int n_in, int k, int match, int mismatch, int indel, int endAlign,
                            uint32_t *blocks_out, int cap_out, int *n_out, int *status, long *cells_out) {
  int st = 0;
  long cells = 0;
  blkvec refined = {0, 0, 0};
  const int maxGap = k - 1;
  if (n_in == 0 || n_in == 1) {
    for (int i = 0; i < n_in && i < cap_out; i++) { blocks_out[3 * i] = blocks_in[3 * i]; blocks_out[3 * i + 1] = blocks_in[3 * i + 1]; blocks_out[3 * i + 2] = blocks_in[3 * i + 2]; }
    *n_out = n_in; if (status) *status = 0; if (cells_out) *cells_out = 0;
    return 0;
  }
  int nb = n_in;
  blk_t *b = (blk_t *)malloc(sizeof(blk_t) * (size_t)(n_in + 2));
  {
    int addStart = 0, addEnd = 0, startMatch = 0, endMatch = 0;
    long qS0 = blocks_in[0], tS0 = blocks_in[1];
    long qAlnEnd = (long)blocks_in[3 * (n_in - 1)] + blocks_in[3 * (n_in - 1) + 2];
    long tAlnEnd = (long)blocks_in[3 * (n_in - 1) + 1] + blocks_in[3 * (n_in - 1) + 2];
    if (endAlign) {
      int minStart = (int)lmin(qS0, tS0);
      if (minStart < 40) { tS0 -= minStart; qS0 -= minStart; startMatch = minStart; addStart = 1; }
      int minEnd = (int)lmin((long)readLen - qAlnEnd, contigLen - tAlnEnd);
      if (minEnd < 40) { endMatch = minEnd; addEnd = 1; }
    }
    int o = 0;
    if (addStart) { b[o].q = (uint32_t)qS0; b[o].t = (uint32_t)tS0; b[o].len = (uint32_t)startMatch; o++; }
    for (int i = 0; i < n_in; i++, o++) { b[o].q = blocks_in[3 * i]; b[o].t = blocks_in[3 * i + 1]; b[o].len = blocks_in[3 * i + 2]; }
    if (addEnd) { b[o].q = (uint32_t)qAlnEnd; b[o].t = (uint32_t)tAlnEnd; b[o].len = (uint32_t)endMatch; o++; }
    nb = o;
  }
  void *aogws = lra_oracle_aog_ws_new();
  int *qS = NULL, *qE = NULL; long bandCap = 0;
  int startBlock = 0, endBlock = 0;
  while (endBlock < nb) {
    long qStart = b[startBlock].q, tStart = b[startBlock].t;
    int blockLen = (int)b[startBlock].len;
    long qPos = (long)b[startBlock].q + blockLen, tPos = (long)b[startBlock].t + blockLen;
    int tGap = 0, qGap = 0;
    if (endBlock < nb - 1) { tGap = (int)(b[endBlock + 1].t - tPos); qGap = (int)(b[endBlock + 1].q - qPos); }
    while (endBlock < nb - 1 && qGap < maxGap && tGap < maxGap && (startBlock == endBlock || b[endBlock].len < 100)) {
      endBlock++;
      int bl = (int)b[endBlock].len;
      qPos = (long)b[endBlock].q + bl; tPos = (long)b[endBlock].t + bl;
      if (endBlock + 1 < nb - 1) { tGap = (int)(b[endBlock + 1].t - tPos); qGap = (int)(b[endBlock + 1].q - qPos); }
    }
    blk_t altEnd; int usedAlt = 0;
    if (endBlock == startBlock) {
      bv_push(&refined, b[startBlock].q, b[startBlock].t, b[startBlock].len);
    } else {
      if ((long)b[startBlock].len > maxGap) {
        int advanced = (int)b[startBlock].len - maxGap;
        b[startBlock].len -= maxGap;
        bv_push(&refined, b[startBlock].q, b[startBlock].t, b[startBlock].len);
        b[startBlock].q += advanced; b[startBlock].t += advanced; b[startBlock].len = maxGap;
        qStart += advanced; tStart += advanced;
      }
      if ((long)b[endBlock].len > maxGap) {
        usedAlt = 1; altEnd = b[endBlock];
        altEnd.q += maxGap; altEnd.t += maxGap; altEnd.len -= maxGap;
        b[endBlock].len = maxGap;
        qPos = (long)b[endBlock].q + maxGap; tPos = (long)b[endBlock].t + maxGap;
      }
      long qEnd = (long)b[endBlock].q + b[endBlock].len, tEnd = (long)b[endBlock].t + b[endBlock].len;
      long tLen = tPos - tStart;
      if (tLen > bandCap) { bandCap = tLen * 2 + 64; free(qS); free(qE); qS = (int *)malloc(sizeof(int) * bandCap); qE = (int *)malloc(sizeof(int) * bandCap); }
      long matSize = build_band(b, startBlock, endBlock, k, qStart, qEnd, tLen, qS, qE);
      long tSeqLen = tEnd - tStart, qSeqLen = qEnd - qStart;
      if (tSeqLen < k || qSeqLen < k) {
        int cap = (int)(lmin(qSeqLen, tSeqLen) + 2), nbk = 0, s2 = 0;
        uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (size_t)cap);
        lra_oracle_aog_ws(aogws, qSeq + qStart, (int)qSeqLen, twin + (tStart - tWinOff), (int)tSeqLen, match, mismatch, indel, k, tmp, cap, &nbk, &s2);
        if (s2) st |= 1;
        for (int i = 0; i < nbk; i++) bv_push(&refined, tmp[3 * i] + (uint32_t)qStart, tmp[3 * i + 1] + (uint32_t)tStart, tmp[3 * i + 2]);
        free(tmp);
      } else {
        cells += matSize;
        int first_out = refined.n;
        if (g_dump && g_dump->n_groups < g_dump->max_groups && g_dump->band_used + 2 * tLen <= g_dump->band_cap) {
          int32_t *m8 = g_dump->meta + 8 * g_dump->n_groups;
          m8[0] = (int32_t)qStart; m8[1] = (int32_t)tStart; m8[2] = (int32_t)tLen; m8[3] = (int32_t)qSeqLen; m8[4] = (int32_t)tSeqLen;
          m8[5] = (int32_t)g_dump->band_used; m8[6] = first_out; m8[7] = -1;
          memcpy(g_dump->band + g_dump->band_used, qS, sizeof(int) * tLen);
          memcpy(g_dump->band + g_dump->band_used + tLen, qE, sizeof(int) * tLen);
          g_dump->band_used += 2 * tLen;
        }
        st |= indel_dp(qSeq, twin, tWinOff, qStart, tStart, tLen, qS, qE, matSize, indel, match, mismatch, &refined, qSeqLen, tSeqLen);
        if (g_dump && g_dump->n_groups < g_dump->max_groups && g_dump->meta[8 * g_dump->n_groups + 7] == -1 &&
            g_dump->meta[8 * g_dump->n_groups + 6] == first_out) {
          g_dump->meta[8 * g_dump->n_groups + 7] = refined.n - first_out;
          g_dump->n_groups++;
        }
      }
    }
    if (!usedAlt) endBlock++; else b[endBlock] = altEnd;
    startBlock = endBlock;
  }
  for (int i = 0; i + 1 < refined.n; i++)
    if ((long)refined.v[i].q + refined.v[i].len > refined.v[i + 1].q || (long)refined.v[i].t + refined.v[i].len > refined.v[i + 1].t) st |= 2;
  for (int i = 0; i < refined.n && i < cap_out; i++) { blocks_out[3 * i] = refined.v[i].q; blocks_out[3 * i + 1] = refined.v[i].t; blocks_out[3 * i + 2] = refined.v[i].len; }
  *n_out = refined.n;
  if (status) *status = st;
  if (cells_out) *cells_out = cells;
  free(refined.v); free(b); free(qS); free(qE); lra_oracle_aog_ws_free(aogws);
  return 0;
}

/* Same as lra_oracle_indel_refine, additionally dumping the DP groups (see ir_dump). Returns the number of groups. */
int lra_oracle_indel_refine_groups(const char *qSeq, int readLen, const char *twin, long tWinOff, long contigLen,
                                   const uint32_t *blocks_in, int n_in, int k, int match, int mismatch, int indel, int endAlign,
                                   uint32_t *blocks_out, int cap_out, int *n_out, int *status, int32_t *meta, int max_groups,
                                   int32_t *band, long band_cap) {
  ir_dump d = {meta, max_groups, 0, band, band_cap, 0};
  for (int i = 0; i < max_groups; i++) meta[8 * i + 7] = -2;
  g_dump = &d;
  long cells;
  lra_oracle_indel_refine(qSeq, readLen, twin, tWinOff, contigLen, blocks_in, n_in, k, match, mismatch, indel, endAlign, blocks_out,
                          cap_out, n_out, status, &cells);
  g_dump = NULL;
  return d.n_groups;
}
