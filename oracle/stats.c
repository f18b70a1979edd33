/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of Alignment::CalculateStatistics for one segment, showmm == true:
 *   CreateAlignmentStrings   /root/reference/Alignment.h:247-333
 *   AlignStringsToCigar      /root/reference/Alignment.h:414-504   (value uses the host-built logf table, LogLookUpTable.h:9-15)
 *   CalculateStatistics      /root/reference/Alignment.h:513-531
 * Bases are compared through seqMap (non-ACGT -> 0, so N equals A); value is accumulated in float in CIGAR order.
 * Outputs are the FUNCTION-LOCAL quantities of one call: n_D / n_I are the numbers of 'D' / 'I' runs.  (The reference stores
 * them swapped into Alignment::nins / ndel, Alignment.h:414 vs :516, and never resets tdel..nLargeIns between calls --
 * SURVEY.md Appendix D-1; both are host-side bookkeeping on top of these per-call numbers.)
 * Pinned by tests/test_oracle_stats.py against the unmodified reference (oracle/ref_wrap.cpp: ref_calc_stats). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline int map2s(unsigned char c) {
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3;
               default: return c < 8 ? (c & 3) : 0; }
}
enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_EQ = 7, OP_X = 8 };

/* stats[16]: nm, nmm, n_D, n_I, tdel, tins, nSmallDel, nMedDel, nLargeDel, nSmallIns, nMedIns, nLargeIns, refLen, preClip, sufClip, n_cigar
 * cigar: BAM-style (len << 4 | op); returns the number of ops (all counted, first cap stored); *value_out = NV */
long lra_oracle_calc_stats(const char *read, int readLen, const char *text, long tWinOff, const uint32_t *blocks, int nb, const float *lut,
                           uint32_t *cigar, long cap, int32_t *stats, float *value_out) {
  long nops = 0;
  int nm = 0, nmm = 0, nD = 0, nI = 0, tdel = 0, tins = 0, sD = 0, mD = 0, lD = 0, sI = 0, mI = 0, lI = 0;
  float value = 0;
  const float coefficient = 3.0f;
  for (int i = 0; i < 16; i++) stats[i] = 0;
  *value_out = 0;
  if (nb == 0) return 0;
  /* column stream: 0 '=' , 1 'X', 2 'D' (query gap char), 3 'I' (target gap char); RLE with the reference's run rules */
  int cur = -1; long run = 0;
  uint32_t q = blocks[0], t = blocks[1];
#define FLUSH() do { if (run > 0) { \
    int op = cur == 0 ? OP_EQ : cur == 1 ? OP_X : cur == 2 ? OP_D : OP_I; \
    if (nops < cap) cigar[nops] = ((uint32_t)run << 4) | (uint32_t)op; nops++; \
    if (cur == 0) { nm += (int)run; value += (float)run; } \
    else if (cur == 1) { nmm += (int)run; value -= (float)run; } \
    else { \
      if (cur == 2) { tdel += (int)run; nD++; if (run <= 10) sD++; if (run > 10 && run < 50) mD++; else if (run > 50) lD++; } \
      else { tins += (int)run; nI++; if (run <= 10) sI++; if (run > 10 && run < 50) mI++; else if (run > 50) lI++; } \
      if (run <= 20) { value -= (float)run; if (cur == 3) sI++; } \
      else if (run <= 10001) { int a = (int)floor((double)((run - 1) / 5)); value += -coefficient * lut[a] - 1; } \
      else if (run <= 100001) value += -1000; else value += -2000; } \
    run = 0; } } while (0)
#define COL(c) do { if ((c) != cur) { FLUSH(); cur = (c); } run++; } while (0)
  for (int b = 0; b < nb; b++) {
    uint32_t len = blocks[3 * b + 2];
    for (uint32_t bl = 0; bl < len; bl++, q++, t++) COL(map2s((unsigned char)read[q]) != map2s((unsigned char)text[(long)t - tWinOff]) ? 1 : 0);
    if (b == nb - 1) continue;
    int qg = (int)(blocks[3 * (b + 1)] - blocks[3 * b] - len);
    int tg = (int)(blocks[3 * (b + 1) + 1] - blocks[3 * b + 1] - len);
    if (qg > 0 || tg > 0) {
      int common = qg > tg ? tg : qg;
      tg -= common; qg -= common;
      for (int g = 0; g < qg; g++, q++) COL(3);
      for (int g = 0; g < tg; g++, t++) COL(2);
      for (int g = 0; g < common; g++, q++, t++) COL(map2s((unsigned char)read[q]) != map2s((unsigned char)text[(long)t - tWinOff]) ? 1 : 0);
    }
  }
  FLUSH();
  stats[0] = nm; stats[1] = nmm; stats[2] = nD; stats[3] = nI; stats[4] = tdel; stats[5] = tins; stats[6] = sD; stats[7] = mD; stats[8] = lD;
  stats[9] = sI; stats[10] = mI; stats[11] = lI; stats[12] = (int32_t)t;  /* refLen = t - refStart(0) */
  stats[13] = (int32_t)blocks[0];
  stats[14] = readLen - (int32_t)blocks[3 * (nb - 1)] - (int32_t)blocks[3 * (nb - 1) + 2];
  stats[15] = (int32_t)nops;
  *value_out = value;
  return nops;
}
