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

// ---- IndelRefineAlignment (IndelRefine.h:53): log the segment before and after.  $LRA_CAPTURE_IR
// record: int32 {readLen, contigLen, k, match, mismatch, indel, endAlign, nIn, nOut, tWinOff, tWinLen, strandFlag}
//         ; read strand bytes [readLen] ; contig window bytes [tWinLen] ; nIn x 3 uint32 ; nOut x 3 uint32
#include <zlib.h>
#include <string>
#include <vector>
#include <iomanip>
#include "htslib/sam.h"
#include "Input.h"       // declares KSEQ_INIT(gzFile, gzread), which Genome.h (pulled in by IndelRefine.h) relies on
#include "IndelRefine.h"
static FILE *lra_cap_ir_fp() {
  static FILE *fp = NULL; static bool init = false;
  if (!init) { init = true; const char *p = getenv("LRA_CAPTURE_IR"); if (p) fp = fopen(p, "wb"); }
  return fp;
}
static inline void IndelRefineAlignment_logged(Read &read, Genome &genome, Alignment &alignment, const Options &opts,
                                               IndelRefineBuffers &buffers, bool endAlign = false) {
  std::vector<Block> before = alignment.blocks;
  IndelRefineAlignment(read, genome, alignment, opts, buffers, endAlign);
  FILE *fp = lra_cap_ir_fp();
  if (fp && before.size() > 0) {
    long contigLen = genome.lengths[alignment.chromIndex];
    long lo = before[0].tPos, hi = before.back().tPos + before.back().length;
    for (size_t i = 0; i < alignment.blocks.size(); i++) {
      if ((long)alignment.blocks[i].tPos < lo) lo = alignment.blocks[i].tPos;
      long e = (long)alignment.blocks[i].tPos + alignment.blocks[i].length; if (e > hi) hi = e;
    }
    lo = lo - 64 < 0 ? 0 : lo - 64; hi = hi + 64 > contigLen ? contigLen : hi + 64;
    int32_t h[12] = {(int32_t)read.length, (int32_t)contigLen, opts.refineBand, opts.localMatch, opts.localMismatch, opts.localIndel,
                     endAlign ? 1 : 0, (int32_t)before.size(), (int32_t)alignment.blocks.size(), (int32_t)lo, (int32_t)(hi - lo), alignment.strand};
    fwrite(h, 4, 12, fp);
    fwrite(alignment.read, 1, read.length, fp);
    fwrite(genome.seqs[alignment.chromIndex] + lo, 1, hi - lo, fp);
    for (size_t i = 0; i < before.size(); i++) { uint32_t v[3] = {before[i].qPos, before[i].tPos, before[i].length}; fwrite(v, 4, 3, fp); }
    for (size_t i = 0; i < alignment.blocks.size(); i++) { uint32_t v[3] = {alignment.blocks[i].qPos, alignment.blocks[i].tPos, alignment.blocks[i].length}; fwrite(v, 4, 3, fp); }
  }
}
#define IndelRefineAlignment IndelRefineAlignment_logged
#include "lra.cpp"
