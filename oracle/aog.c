/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's banded one-free-gap global
 * aligner.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product library never does.
 *
 * Follows  AffineOneGapAlign()  /root/reference/AffineOneGapAlign.h:157-649  (index helpers :12-27).
 * Pinned against the real reference (oracle/_ref/libref_lra.so, built from the unmodified header) by
 * tests/test_oracle_aog.py and against tests/golden/aog_*.bin (captured from reference runs).
 *
 * Model (all of it observable in the block list, so all of it is kept):
 *   - sequences are compared through the 5-letter code  A,C,G,T -> 0..3, anything else -> 4  (:173-182)
 *   - diag = max(1,min(qLen,tLen)), k = min(diag,k); if diag+2k >= max(qLen,tLen) the band half-width
 *     is doubled and only the prefix matrix is used, traced back from (qB-1,tB-1)        (:194-203,:582-586)
 *   - otherwise a prefix band from (0,0) and a suffix band ending in (qLen,tLen) are joined by one free
 *     gap through per-row / per-column running maxima                                      (:347-360,:474-518)
 *   - matrices are flat with a row stride R=2k+3; every cell that is not written holds MISSING
 *   - tie order on equal scores: query-gap ("left") > target-gap ("down") > diagonal > close-gaps
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AOG_MISSING ((int64_t)INT_MIN)
enum { AR_DONE = 0, AR_LEFT = 1, AR_DOWN = 2, AR_DIAG = 3, AR_BORDER = 4, AR_GAPLEFT = 5, AR_GAPDOWN = 6 };

static inline int base_code(unsigned char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return c < 8 ? (c & 3) : 4; /* SeqUtils.h:42-75: raw 0..7 map to 0..3 twice */
  }
}

typedef struct {
  int64_t *ps, *ss;
  int32_t *pp, *sp;
  size_t mat_cap;
  int32_t *umax, *uidx, *lmax, *lidx;
  size_t diag_cap;
  int32_t *ops, *lens;
  size_t ops_cap;
  int32_t *qc, *tc;
  size_t q_cap, t_cap;
} aog_ws;

void *lra_oracle_aog_ws_new(void) { return calloc(1, sizeof(aog_ws)); }
void lra_oracle_aog_ws_free(void *p) {
  aog_ws *w = (aog_ws *)p;
  if (!w) return;
  free(w->ps); free(w->ss); free(w->pp); free(w->sp); free(w->umax); free(w->uidx); free(w->lmax);
  free(w->lidx); free(w->ops); free(w->lens); free(w->qc); free(w->tc); free(w);
}
#define GROW(ptr, cap, need, T) \
  do { if ((cap) < (size_t)(need)) { (cap) = (size_t)(need) * 2 + 64; free(ptr); (ptr) = (T *)malloc((cap) * sizeof(T)); } } while (0)

static inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* Returns the alignment score; writes up to `cap` (qPos,tPos,len) triples and the true block count.
 * *status: 0 ok, 1 = a matrix index left the allocated area (input outside the reference's own
 * defined behaviour; callers treat as "not in domain"). */
int lra_oracle_aog_ws(void *wsp, const char *q, int qLen, const char *t, int tLen, int m, int mm, int indel,
                      int k, uint32_t *blocks, int cap, int *n_blocks, int *status) {
  aog_ws *w = (aog_ws *)wsp;
  int st = 0;
  *n_blocks = 0;
  int diag = imax(1, imin(qLen, tLen));
  GROW(w->qc, w->q_cap, qLen + 1, int32_t);
  GROW(w->tc, w->t_cap, tLen + 1, int32_t);
  w->qc[0] = w->tc[0] = 0;
  for (int s = 0; s < qLen; s++) w->qc[s + 1] = base_code((unsigned char)q[s]);
  for (int s = 0; s < tLen; s++) w->tc[s + 1] = base_code((unsigned char)t[s]);
  if (w->diag_cap < (size_t)diag + 1) {
    size_t nc = (size_t)(diag + 1) * 2 + 64;
    free(w->umax); free(w->uidx); free(w->lmax); free(w->lidx);
    w->umax = malloc(nc * 4); w->uidx = malloc(nc * 4); w->lmax = malloc(nc * 4); w->lidx = malloc(nc * 4);
    w->diag_cap = nc;
  }
  for (int i = 0; i <= diag; i++) { w->umax[i] = INT_MIN; w->lmax[i] = INT_MIN; w->uidx[i] = 0; w->lidx[i] = 0; }

  k = imin(diag, k);
  int two_sided = 1;
  if (diag + 2 * k >= imax(qLen, tLen)) { k = 2 * k; two_sided = 0; }
  const int R = 2 * k + 3;
  const long matSize = (long)(3 + k + diag) * R;
  if (w->mat_cap < (size_t)matSize) {
    size_t nc = (size_t)matSize * 2 + 64;
    free(w->ps); free(w->ss); free(w->pp); free(w->sp);
    w->ps = malloc(nc * 8); w->ss = malloc(nc * 8); w->pp = malloc(nc * 4); w->sp = malloc(nc * 4);
    w->mat_cap = nc;
  }
  int64_t *ps = w->ps, *ss = w->ss;
  int32_t *pp = w->pp, *sp = w->sp;
  for (long x = 0; x < matSize; x++) { ps[x] = AOG_MISSING; ss[x] = AOG_MISSING; pp[x] = -1; sp[x] = -1; }
  if (w->ops_cap < (size_t)qLen + tLen + 16) {
    w->ops_cap = ((size_t)qLen + tLen + 16) * 2;
    free(w->ops); free(w->lens);
    w->ops = malloc(w->ops_cap * 4); w->lens = malloc(w->ops_cap * 4);
  }
  int32_t *ops = w->ops, *lens = w->lens;
  int nops = 0;

#define PIDX(i, j) ((long)(j) * R + ((i) - (j)) + k + 1)
#define CHK(ix) (((ix) < 0 || (ix) >= matSize) ? (st = 1, 0L) : (ix))
#define PS(i, j) ps[CHK(PIDX(i, j))]
#define PP(i, j) pp[CHK(PIDX(i, j))]

  /* ---- prefix matrix: borders (:229-241) */
  for (int i = 1; i < k + 1; i++) { PS(i, 0) = (int64_t)indel * i; PP(i, 0) = AR_LEFT; }
  for (int j = 1; j <= k + 1; j++) { PS(0, j) = (int64_t)indel * j; PP(0, j) = AR_DOWN; }
  PS(0, 0) = 0; PP(0, 0) = AR_DONE;
  /* rails (:248-306); note these overwrite (0,k+1) in most cases */
  if (qLen >= tLen) {
    for (int i = 0; i <= diag - k - 1; i++) { PS(i, i + k + 1) = AOG_MISSING; PP(i, i + k + 1) = AR_BORDER; }
    for (int i = 1; i < diag + k - 1; i++) { PS(i + k + 1, i) = AOG_MISSING; PP(i + k + 1, i) = AR_BORDER; }
    w->lmax[0] = 0; w->lidx[0] = 0;
  }
  if (qLen <= tLen) {
    for (int j = 0; j < diag - 1; j++) { PS(j + k + 1, j) = AOG_MISSING; PP(j + k + 1, j) = AR_BORDER; }
    for (int j = 1; j < diag + k; j++) { PS(j - k - 1, j) = AOG_MISSING; PP(j - k - 1, j) = AR_BORDER; }
    w->umax[0] = 0; w->uidx[0] = 0;
  }
  const int qB = imin(diag + k, qLen + 1), tB = imin(diag + k, tLen + 1);
  /* ---- prefix fill (:313-362) */
  for (int j = 1; j < tB; j++) {
    int ihi = imin(qB, j + k + 1);
    for (int i = imax(1, j - k); i < ihi; i++) {
      int64_t sIns = PS(i - 1, j) + indel;
      int64_t sDel = PS(i, j - 1) + indel;
      int64_t sMat = PS(i - 1, j - 1) + (w->qc[i] == w->tc[j] ? m : mm);
      int64_t best = max64(sIns, max64(sDel, sMat));
      PS(i, j) = best;
      PP(i, j) = best == sIns ? AR_LEFT : (best == sDel ? AR_DOWN : AR_DIAG);
      if (i < qLen - k && best >= (int64_t)w->lmax[j]) { w->lmax[j] = (int32_t)best; w->lidx[j] = i; }
      if (j < tLen && i < diag + 1 && best > (int64_t)w->umax[i]) { w->umax[i] = (int32_t)best; w->uidx[i] = j; }
    }
  }

  int i, j, score = -1;
#define PUSH(op)                                                              \
  do { if (nops == 0 || ops[nops - 1] != (op)) { if ((size_t)nops + 1 >= w->ops_cap) { st = 1; goto done; } \
         ops[nops] = (op); lens[nops] = 1; nops++; } else lens[nops - 1]++; } while (0)
  if (two_sided) {
    /* ---- suffix matrix (:409-518) */
    const int qStart = imax(0, qLen - diag), qEnd = qLen + 1;
    const int tStart = imax(0, tLen - diag), tEnd = tLen + 1;
    const int tLow = imax(0, tLen - diag - k - 1 - 1), qLow = imax(0, qLen - diag - k - 1);
#define SIDX(ii, jj) ((long)((jj) - tLow) * R + (((ii) - qLow) - ((jj) - tLow)) + k + 1)
#define SS(ii, jj) ss[CHK(SIDX(ii, jj))]
#define SP(ii, jj) sp[CHK(SIDX(ii, jj))]
    if (qLen >= tLen) {
      for (i = qLow, j = 0; i < qStart + k + 1; i++) { SS(i, j) = w->lmax[j]; SP(i, j) = AR_GAPLEFT; }
      for (i = qLow, j = 1; i < qLow + diag; i++, j++) { SS(i, j) = w->lmax[j]; SP(i, j) = AR_GAPLEFT; }
      for (j = tStart + 1, i = qStart; j < tEnd - k; i++, j++) { SS(i + k + 1, j) = AOG_MISSING; SP(i + k + 1, j) = AR_BORDER; }
    }
    if (qLen <= tLen) {
      for (j = tLow, i = qStart; j < tStart + k + 2; j++) { SS(i, j) = w->umax[0]; SP(i, j) = AR_GAPDOWN; }
      for (j = tStart + 1, i = qStart + 1; j < tEnd; i++, j++) { SS(i, j - k - 1) = w->umax[i]; SP(i, j - k - 1) = AR_GAPDOWN; }
      for (j = tStart, i = qStart; j < tEnd - k - 1; i++, j++) { SS(i, j + k + 1) = AOG_MISSING; SP(i, j + k + 1) = AR_BORDER; }
    }
    for (j = tLow + 1; j < tEnd; j++) {
      int doff = diag + 1 - (tEnd - j);
      int ihi = imin(qEnd, qStart + doff + k + 1);
      for (i = imax(qLow + 1, qStart + doff - k); i < ihi; i++) {
        int64_t delClose = AOG_MISSING, insClose = AOG_MISSING;
        if (qLen >= tLen) delClose = w->lmax[j];
        if (tLen > qLen) insClose = w->umax[i];
        int64_t sIns = SS(i - 1, j) + indel;
        int64_t sDel = SS(i, j - 1) + indel;
        int64_t sMat = SS(i - 1, j - 1) + (w->qc[i] == w->tc[j] ? m : mm);
        int64_t best = max64(delClose, max64(insClose, max64(sIns, max64(sDel, sMat))));
        SS(i, j) = best;
        if (best == sIns) SP(i, j) = AR_LEFT;
        else if (best == sDel) SP(i, j) = AR_DOWN;
        else if (best == sMat) SP(i, j) = AR_DIAG;
        else if (best == delClose) SP(i, j) = AR_GAPLEFT;
        else if (best == insClose) SP(i, j) = AR_GAPDOWN;
      }
    }
    /* ---- suffix traceback (:523-580) */
    i = qLen; j = tLen;
    int arrow = SP(i, j);
    score = (int)SS(i, j);
    long guard = 0;
    while (arrow != AR_DONE && arrow != AR_GAPDOWN && arrow != AR_GAPLEFT && i >= 0 && j >= 0) {
      PUSH(arrow);
      if (arrow == AR_DIAG) { i--; j--; }
      else if (arrow == AR_LEFT) i--;
      else if (arrow == AR_DOWN) j--;
      else if (++guard > 4) { st = 1; goto done; } /* reference would spin forever on -1/border */
      if (i >= 0 && j >= 0) arrow = SP(i, j);
    }
    if (i < 0 || j < 0) { st = 1; goto done; }
    if (arrow == AR_GAPDOWN) {
      if (i > diag) { st = 1; goto done; }
      ops[nops] = arrow; lens[nops] = j - w->uidx[i]; nops++;
      j = w->uidx[i];
    }
    if (arrow == AR_GAPLEFT) {
      if (j > diag) { st = 1; goto done; }
      ops[nops] = arrow; lens[nops] = i - w->lidx[j]; nops++;
      i = w->lidx[j];
    }
  } else {
    i = qB - 1; j = tB - 1;
    score = (int)PS(i, j);
  }
  /* ---- prefix traceback (:589-629) */
  {
    if (i < 0 || j < 0) { st = 1; goto done; }
    int arrow = PP(i, j);
    long guard = 0;
    while (arrow != AR_BORDER && arrow != AR_DONE && i >= 0 && j >= 0) {
      PUSH(arrow);
      if (arrow == AR_DIAG) { i--; j--; }
      else if (arrow == AR_LEFT) i--;
      else if (arrow == AR_DOWN) j--;
      else if (arrow == AR_GAPLEFT || arrow == AR_GAPDOWN) break;
      else if (++guard > 4) { st = 1; goto done; }
      if (i < 0 || j < 0) break; /* reference: harmless read at a negative coordinate, then loop exit */
      arrow = PP(i, j);
    }
  }
  /* ---- ops (reverse order) -> blocks (:630-647) */
  {
    uint32_t qPos = 0, tPos = 0;
    int nb = 0;
    for (int x = nops; x > 0; x--) {
      int op = ops[x - 1], len = lens[x - 1];
      if (op == AR_LEFT || op == AR_GAPLEFT) qPos += len;
      else if (op == AR_DOWN || op == AR_GAPDOWN) tPos += len;
      else if (op == AR_DIAG) {
        if (nb < cap) { blocks[3 * nb] = qPos; blocks[3 * nb + 1] = tPos; blocks[3 * nb + 2] = (uint32_t)len; }
        nb++; qPos += len; tPos += len;
      }
    }
    *n_blocks = nb;
  }
done:
  if (status) *status = st;
  return score;
}

int lra_oracle_aog(const char *q, int qLen, const char *t, int tLen, int m, int mm, int indel, int k,
                   uint32_t *blocks, int cap, int *n_blocks, int *status) {
  void *w = lra_oracle_aog_ws_new();
  int s = lra_oracle_aog_ws(w, q, qLen, t, tLen, m, mm, indel, k, blocks, cap, n_blocks, status);
  lra_oracle_aog_ws_free(w);
  return s;
}

/* Batch form over SoA jobs (single thread; the "port" CPU baseline).  blocks laid out at block_off[j]. */
int lra_oracle_aog_batch(const char *q_arena, const char *t_arena, const uint32_t *q_off, const uint32_t *t_off,
                         const int32_t *q_len, const int32_t *t_len, const int32_t *k, int n_jobs, int m, int mm,
                         int indel, int32_t *score, int32_t *n_blocks, const int64_t *block_off,
                         const int32_t *block_cap, uint32_t *blocks_out, int32_t *status_out) {
  void *w = lra_oracle_aog_ws_new();
  for (int j = 0; j < n_jobs; j++) {
    int nb = 0, st = 0;
    score[j] = lra_oracle_aog_ws(w, q_arena + q_off[j], q_len[j], t_arena + t_off[j], t_len[j], m, mm, indel, k[j],
                                 blocks_out ? blocks_out + 3 * block_off[j] : NULL,
                                 blocks_out ? block_cap[j] : 0, &nb, &st);
    n_blocks[j] = nb;
    if (status_out) status_out[j] = st;
  }
  lra_oracle_aog_ws_free(w);
  return 0;
}
