// TEST INFRASTRUCTURE ONLY.  The reference `lra` binary with every AffineOneGapAlign call logged
// (inputs and outputs) to the file named by $LRA_CAPTURE_AOG, for golden vectors and for the job-shape
// tables that bench.py's synthetic workload is drawn from.  Built by oracle/Makefile from the
// unmodified reference sources under /root/reference; run with -t 1.
//
// Mechanism: the reference function is defined first (header guard), then the name is macro-renamed
// so that every later call site in the reference headers goes through the logger.
#include "htslib/hts.h"
#include "htslib/kseq.h"
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "AffineOneGapAlign.h"

static FILE *lra_cap_aog_fp() {
  static FILE *fp = NULL; static bool init = false;
  if (!init) { init = true; const char *p = getenv("LRA_CAPTURE_AOG"); if (p) fp = fopen(p, "wb"); }
  return fp;
}
// record: int32 {qLen,tLen,m,mm,indel,k,score,nBlocksAdded} ; q bytes ; t bytes ; nBlocksAdded x 3 uint32
static inline int AffineOneGapAlign_logged(string &qSeq, int qLen, string &tSeq, int tLen, int m, int mm,
                                           int indel, int k, Alignment &aln, AffineAlignBuffers &b) {
  size_t n0 = aln.blocks.size();
  int score = AffineOneGapAlign(qSeq, qLen, tSeq, tLen, m, mm, indel, k, aln, b);
  FILE *fp = lra_cap_aog_fp();
  if (fp) {
    int32_t h[8] = {qLen, tLen, m, mm, indel, k, score, (int32_t)(aln.blocks.size() - n0)};
    fwrite(h, 4, 8, fp);
    fwrite(qSeq.data(), 1, qLen, fp);
    fwrite(tSeq.data(), 1, tLen, fp);
    for (size_t i = n0; i < aln.blocks.size(); i++) {
      uint32_t v[3] = {aln.blocks[i].qPos, aln.blocks[i].tPos, aln.blocks[i].length};
      fwrite(v, 4, 3, fp);
    }
  }
  return score;
}
#define AffineOneGapAlign AffineOneGapAlign_logged
#include "lra.cpp"
