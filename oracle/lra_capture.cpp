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

// ---- SparseDP (SparseDP.h:2139 pure matches, :2287 one cluster) and SparseDP_ForwardOnly (SparseDP_Forward.h:312): inputs and outputs of every
// call of the low-accuracy pipeline.  $LRA_CAPTURE_SDP.  Records (all 32-bit words, little endian):
//   kind 0: {0, n_cl, nfrag, rate(f32), alnthres(f32), NumAln, read_len} cl_off[n_cl+1] cl_strand[n_cl] q[nfrag] t[nfrag] len[nfrag]
//           {n_chains} then per chain {n, value(f32), QStart, QEnd, TStart, TEnd} chain[n] (global fragment index) link[n-1]
//   kind 1: {1, strand, nfrag, rate(f32)} q t len {n, value(f32)} chain[n] link[n-1]
//   kind 2: {2, nfrag, rate} q t len {n, value(f32)} chain[n]
//   kind 3: see below
#include "SparseDP.h"
#include "SparseDP_Forward.h"
static FILE *lra_cap_sdp_fp() {
  static FILE *fp = NULL; static bool init = false;
  if (!init) { init = true; const char *p = getenv("LRA_CAPTURE_SDP"); if (p) fp = fopen(p, "wb"); }
  return fp;
}
static inline void cap_w32(FILE *fp, uint32_t v) { fwrite(&v, 4, 1, fp); }
static inline void cap_wf(FILE *fp, float v) { fwrite(&v, 4, 1, fp); }
static inline int SparseDP_logged(vector<Cluster> &FragInput, vector<UltimateChain> &chains, const Options &opts, const vector<float> &LookUpTable, Read &read, float rate) {
  FILE *fp = lra_cap_sdp_fp();
  size_t c0 = chains.size();
  int un = read.unaligned;
  int r = SparseDP(FragInput, chains, opts, LookUpTable, read, rate);
  if (fp && !un) {
    int nfrag = 0; for (size_t c = 0; c < FragInput.size(); c++) nfrag += FragInput[c].matches.size();
    cap_w32(fp, 0); cap_w32(fp, FragInput.size()); cap_w32(fp, nfrag); cap_wf(fp, rate); cap_wf(fp, opts.alnthres); cap_w32(fp, opts.NumAln); cap_w32(fp, read.length);
    std::vector<int> off(FragInput.size() + 1, 0);
    for (size_t c = 0; c < FragInput.size(); c++) off[c + 1] = off[c] + FragInput[c].matches.size();
    for (size_t c = 0; c <= FragInput.size(); c++) cap_w32(fp, off[c]);
    for (size_t c = 0; c < FragInput.size(); c++) cap_w32(fp, FragInput[c].strand);
    for (size_t c = 0; c < FragInput.size(); c++) for (size_t i = 0; i < FragInput[c].matches.size(); i++) cap_w32(fp, FragInput[c].matches[i].first.pos);
    for (size_t c = 0; c < FragInput.size(); c++) for (size_t i = 0; i < FragInput[c].matches.size(); i++) cap_w32(fp, FragInput[c].matches[i].second.pos);
    for (size_t c = 0; c < FragInput.size(); c++) for (size_t i = 0; i < FragInput[c].matches.size(); i++) cap_w32(fp, FragInput[c].matchesLengths[i]);
    cap_w32(fp, chains.size() - c0);
    for (size_t c = c0; c < chains.size(); c++) {
      cap_w32(fp, chains[c].chain.size()); cap_wf(fp, chains[c].FirstSDPValue);
      cap_w32(fp, chains[c].QStart); cap_w32(fp, chains[c].QEnd); cap_w32(fp, chains[c].TStart); cap_w32(fp, chains[c].TEnd);
      for (size_t s = 0; s < chains[c].chain.size(); s++) cap_w32(fp, chains[c].chain[s] + off[chains[c].ClusterIndex[s]]);
      for (size_t s = 0; s + 1 < chains[c].chain.size(); s++) cap_w32(fp, s < chains[c].link.size() ? (uint32_t)chains[c].link[s] : 255u);
    }
  }
  return r;
}
static inline int SparseDP_logged(int ClusterIndex, vector<Cluster> &FragInput, UltimateChain &ultimatechain, const Options &opts, const vector<float> &LookUpTable, Read &read) {
  FILE *fp = lra_cap_sdp_fp();
  int un = read.unaligned;
  int r = SparseDP(ClusterIndex, FragInput, ultimatechain, opts, LookUpTable, read);
  if (fp && !un && FragInput[ClusterIndex].matches.size() > 0) {
    Cluster &C = FragInput[ClusterIndex];
    cap_w32(fp, 1); cap_w32(fp, C.strand); cap_w32(fp, C.matches.size()); cap_wf(fp, opts.second_anchorbonus);
    for (size_t i = 0; i < C.matches.size(); i++) cap_w32(fp, C.matches[i].first.pos);
    for (size_t i = 0; i < C.matches.size(); i++) cap_w32(fp, C.matches[i].second.pos);
    for (size_t i = 0; i < C.matches.size(); i++) cap_w32(fp, C.matchesLengths[i]);
    cap_w32(fp, ultimatechain.chain.size()); cap_wf(fp, ultimatechain.FirstSDPValue);
    for (size_t s = 0; s < ultimatechain.chain.size(); s++) cap_w32(fp, ultimatechain.chain[s]);
    for (size_t s = 0; s + 1 < ultimatechain.chain.size(); s++) cap_w32(fp, s < ultimatechain.link.size() ? (uint32_t)ultimatechain.link[s] : 255u);
  }
  return r;
}
// kind 3 (SparseDP.h:1766, the high-accuracy pipeline's second SparseDP over the Cluster_SameDiag anchors of a split chain):
//   {3, n_cl, nfrag, rate(f32)} cl_off[n_cl+1] cl_strand[n_cl] q[nfrag] t[nfrag] len[nfrag] {n, value(f32)} chain[n] (index into the concatenation)
static inline int SparseDP_logged(SplitChain &inputChain, vector<Cluster_SameDiag *> &FragInput, FinalChain &finalchain, const Options &opts, const vector<float> &LookUpTable, Read &read) {
  FILE *fp = lra_cap_sdp_fp();
  int un = read.unaligned;
  int r = SparseDP(inputChain, FragInput, finalchain, opts, LookUpTable, read);
  if (fp && !un && inputChain.size() > 0) {
    const int ncl = inputChain.size();
    std::vector<int> off(ncl + 1, 0);
    for (int c = 0; c < ncl; c++) off[c + 1] = off[c] + FragInput[inputChain[c]]->size();
    cap_w32(fp, 3); cap_w32(fp, ncl); cap_w32(fp, off[ncl]); cap_wf(fp, opts.second_anchorbonus);
    for (int c = 0; c <= ncl; c++) cap_w32(fp, off[c]);
    for (int c = 0; c < ncl; c++) cap_w32(fp, FragInput[inputChain[c]]->strand);
    for (int c = 0; c < ncl; c++) for (int i = 0; i < FragInput[inputChain[c]]->size(); i++) cap_w32(fp, FragInput[inputChain[c]]->GetqStart(i));
    for (int c = 0; c < ncl; c++) for (int i = 0; i < FragInput[inputChain[c]]->size(); i++) cap_w32(fp, FragInput[inputChain[c]]->GettStart(i));
    for (int c = 0; c < ncl; c++) for (int i = 0; i < FragInput[inputChain[c]]->size(); i++) cap_w32(fp, FragInput[inputChain[c]]->length(i));
    cap_w32(fp, finalchain.chain.size()); cap_wf(fp, finalchain.SecondSDPValue);
    for (size_t s = 0; s < finalchain.chain.size(); s++) cap_w32(fp, finalchain.chain[s] + FragInput[finalchain.ClusterIndex[s]]->matchStart);
  }
  return r;
}
// kind 4 (SparseDP.h:1956, the first SparseDP of the high-accuracy pipeline over the split clusters):
//   {4, nfrag, rate(f32), alnthres(f32), globalK, NumAln, read_len} qS[n] qE[n] tS[n] tE[n] strand[n] Val(f32)[n] NumofAnchors0[n]
//   {n_chains} then per chain of Primary_chains[0] {n, value(f32), qStart, qEnd, tStart, tEnd, NumOfAnchors0} ch[n] link[n-1]
static inline int SparseDP_logged(vector<Cluster> &FragInput, vector<Primary_chain> &Primary_chains, const Options &opts, const vector<float> &LookUpTable, Read &read, float &rate) {
  FILE *fp = lra_cap_sdp_fp();
  size_t p0 = Primary_chains.size();
  int r = SparseDP(FragInput, Primary_chains, opts, LookUpTable, read, rate);
  if (fp && FragInput.size() > 0 && p0 == 0) {
    const size_t n = FragInput.size();
    cap_w32(fp, 4); cap_w32(fp, n); cap_wf(fp, rate); cap_wf(fp, opts.alnthres); cap_w32(fp, opts.globalK); cap_w32(fp, opts.NumAln); cap_w32(fp, read.length);
    for (size_t i = 0; i < n; i++) cap_w32(fp, FragInput[i].qStart);
    for (size_t i = 0; i < n; i++) cap_w32(fp, FragInput[i].qEnd);
    for (size_t i = 0; i < n; i++) cap_w32(fp, FragInput[i].tStart);
    for (size_t i = 0; i < n; i++) cap_w32(fp, FragInput[i].tEnd);
    for (size_t i = 0; i < n; i++) cap_w32(fp, FragInput[i].strand);
    for (size_t i = 0; i < n; i++) cap_wf(fp, FragInput[i].Val);
    for (size_t i = 0; i < n; i++) cap_w32(fp, FragInput[i].NumofAnchors0);
    const size_t nc = Primary_chains.size() ? Primary_chains[0].chains.size() : 0;
    cap_w32(fp, nc);
    for (size_t c = 0; c < nc; c++) {
      CHain &ch = Primary_chains[0].chains[c];
      cap_w32(fp, ch.ch.size()); cap_wf(fp, ch.value); cap_w32(fp, ch.qStart); cap_w32(fp, ch.qEnd); cap_w32(fp, ch.tStart); cap_w32(fp, ch.tEnd); cap_w32(fp, ch.NumOfAnchors0);
      for (size_t x = 0; x < ch.ch.size(); x++) cap_w32(fp, ch.ch[x]);
      for (size_t x = 0; x + 1 < ch.ch.size(); x++) cap_w32(fp, x < ch.link.size() ? (uint32_t)ch.link[x] : 255u);
    }
  }
  return r;
}
static inline int SparseDP_ForwardOnly_logged(const GenomePairs &FragInput, const vector<int> &MatchLengths, std::vector<unsigned int> &chain, const Options &opts,
                                              const std::vector<float> &LookUpTable, float &inv_value, int &inv_NumOfAnchors, int rate = 5) {
  FILE *fp = lra_cap_sdp_fp();
  int r = SparseDP_ForwardOnly(FragInput, MatchLengths, chain, opts, LookUpTable, inv_value, inv_NumOfAnchors, rate);
  if (fp && FragInput.size() > 0) {
    cap_w32(fp, 2); cap_w32(fp, FragInput.size()); cap_w32(fp, rate);
    for (size_t i = 0; i < FragInput.size(); i++) cap_w32(fp, FragInput[i].first.pos);
    for (size_t i = 0; i < FragInput.size(); i++) cap_w32(fp, FragInput[i].second.pos);
    for (size_t i = 0; i < FragInput.size(); i++) cap_w32(fp, MatchLengths[i]);
    cap_w32(fp, chain.size()); cap_wf(fp, inv_value);
    for (size_t s = 0; s < chain.size(); s++) cap_w32(fp, chain[s]);
  }
  return r;
}
#define SparseDP(...) SparseDP_logged(__VA_ARGS__)
#define SparseDP_ForwardOnly(...) SparseDP_ForwardOnly_logged(__VA_ARGS__)
#include "lra.cpp"
