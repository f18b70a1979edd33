// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or shipped with liblra_b200.so.
//
// Thin extern "C" entry points around the UNMODIFIED reference headers, compiled from
// /root/reference where they lie (see oracle/Makefile) into oracle/_ref/libref_lra.so.
// Used (a) to pin the C restatement in oracle/*.c, (b) to generate tests/golden/*, and
// (c) as the "reference" CPU arm of bench.py (cpu_baseline.kind == "reference").
//
// Reference entry points wrapped here:
//   AffineOneGapAlign            AffineOneGapAlign.h:157-649
//   IndelRefineAlignment         IndelRefine.h:53-784
//   StoreMinimizers / std::sort / CompareLists / SeparateMatchesByStrand semantics   MinCount.h:7-179, MapRead.h:185,
//                                CompareLists.h:8-151, MapRead.h:109-150   (the seeding prefix of MapRead, MapRead.h:169-203)
//   LocalIndex::IndexSeq         MMIndex.h:200-245
//   REFINEclusters               ClusterRefine.h:50-240
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <zlib.h>
#include <iomanip>
#include "htslib/kseq.h"
#include "htslib/sam.h"
// the reference's headers only compile in the include order of its own translation unit (lra.cpp:18-29)
#include "Input.h"         // declares KSEQ_INIT(gzFile, gzread) for Genome.h
#include "MMIndex.h"
#include "TupleOps.h"
#include "MinCount.h"
#include "MapRead.h"
#include "SeqUtils.h"
#include "Options.h"
#include "Alignment.h"
#include "LogLookUpTable.h"
#include "GlobalChain.h"
#include "Fragment.h"

// A chain type for the unmodified filter templates of Chain.h (they only use these members)
struct WrapChain {
  std::vector<uint32_t> q, t, len; std::vector<uint8_t> st;
  std::vector<unsigned int> chain; std::vector<int> ClusterIndex; std::vector<bool> link;
  int size() { return (int)chain.size(); }
  bool strand(int i) { return st[chain[i]] != 0; }
  GenomePos &qStart(int i) { return q[chain[i]]; }
  GenomePos &tStart(int i) { return t[chain[i]]; }
  GenomePos qEnd(int i) { return q[chain[i]] + len[chain[i]]; }
  GenomePos tEnd(int i) { return t[chain[i]] + len[chain[i]]; }
  int length(int i) { return (int)len[chain[i]]; }
};

extern "C" {

// One call. blocks_out receives up to cap (qPos,tPos,length) triples; *n_blocks the true count.
int ref_aog(const char *q, int qLen, const char *t, int tLen, int m, int mm, int indel, int k,
            uint32_t *blocks_out, int cap, int *n_blocks) {
  std::string qs(q, qLen), ts(t, tLen);
  Alignment aln;
  AffineAlignBuffers buf;
  int score = AffineOneGapAlign(qs, qLen, ts, tLen, m, mm, indel, k, aln, buf);
  int n = (int)aln.blocks.size();
  *n_blocks = n;
  for (int i = 0; i < n && i < cap; i++) {
    blocks_out[3 * i] = aln.blocks[i].qPos;
    blocks_out[3 * i + 1] = aln.blocks[i].tPos;
    blocks_out[3 * i + 2] = aln.blocks[i].length;
  }
  return score;
}

// Batch over SoA job arrays on `nthreads` host threads (each with its own, re-used
// AffineAlignBuffers, the favourable case for the reference: LocalRefineAlignment.h:114 passes a
// long-lived buffer).  block_off[j] = index of job j's first triple in blocks_out, laid out with the
// caller-provided per-job capacity prefix (block_off is an input).  Returns 0.
int ref_aog_batch(const char *q_arena, const char *t_arena, const uint32_t *q_off,
                  const uint32_t *t_off, const int32_t *q_len, const int32_t *t_len,
                  const int32_t *k, int n_jobs, int m, int mm, int indel, int32_t *score,
                  int32_t *n_blocks, const int64_t *block_off, uint32_t *blocks_out, int nthreads) {
  std::atomic<int> next(0);
  auto work = [&]() {
    AffineAlignBuffers buf;
    Alignment aln;
    const int CH = 64;
    for (;;) {
      int s = next.fetch_add(CH);
      if (s >= n_jobs) break;
      int e = s + CH < n_jobs ? s + CH : n_jobs;
      for (int j = s; j < e; j++) {
        std::string qs(q_arena + q_off[j], q_len[j]), ts(t_arena + t_off[j], t_len[j]);
        aln.blocks.clear();
        score[j] = AffineOneGapAlign(qs, q_len[j], ts, t_len[j], m, mm, indel, k[j], aln, buf);
        n_blocks[j] = (int)aln.blocks.size();
        if (blocks_out) {
          uint32_t *o = blocks_out + 3 * block_off[j];
          for (size_t i = 0; i < aln.blocks.size(); i++) {
            o[3 * i] = aln.blocks[i].qPos; o[3 * i + 1] = aln.blocks[i].tPos; o[3 * i + 2] = aln.blocks[i].length;
          }
        }
      }
    }
  };
  if (nthreads <= 1) { work(); return 0; }
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; i++) th.emplace_back(work);
  for (auto &x : th) x.join();
  return 0;
}

// IndelRefineAlignment over a batch of segments on `nthreads` host threads.  Segment s: blocks_in[blk_off[s] .. +blk_cnt[s]),
// read strand = q_arena + q_base[s] (read_len[s] bases), contig = t_arena + t_base[s] (contig_len[s] bases; block tPos
// are contig-relative).  Output blocks at out_blocks[3*out_off[s]] with capacity read_len[s]+contig_len[s] each (out_off is an
// input).  Returns 0.
int ref_indel_refine_batch(const char *q_arena, const char *t_arena, const uint32_t *blocks_in, const uint64_t *blk_off,
                           const int32_t *blk_cnt, const uint32_t *q_base, const uint32_t *t_base, const int32_t *read_len,
                           const int32_t *contig_len, int n_seg, int refineBand, int match, int mismatch, int indel, int endAlign,
                           int32_t *out_n, const uint64_t *out_off, uint32_t *out_blocks, int nthreads) {
  std::atomic<int> next(0);
  auto work = [&]() {
    Options opts;
    opts.refineBand = refineBand; opts.localMatch = match; opts.localMismatch = mismatch; opts.localIndel = indel;
    IndelRefineBuffers buffers;
    for (;;) {
      int s = next.fetch_add(1);
      if (s >= n_seg) break;
      Read read;
      read.length = read_len[s];
      read.name = "r";
      Genome genome;
      genome.seqs.push_back((char *)t_arena + t_base[s]);
      genome.lengths.push_back(contig_len[s]);
      Alignment aln;
      aln.chromIndex = 0;
      aln.read = (char *)q_arena + q_base[s];
      aln.blocks.resize(blk_cnt[s]);
      for (int i = 0; i < blk_cnt[s]; i++) {
        const uint32_t *b = blocks_in + 3 * (blk_off[s] + i);
        aln.blocks[i] = Block(b[0], b[1], b[2]);
      }
      IndelRefineAlignment(read, genome, aln, opts, buffers, endAlign != 0);
      out_n[s] = (int32_t)aln.blocks.size();
      if (out_blocks) {
        uint32_t *o = out_blocks + 3 * out_off[s];
        for (size_t i = 0; i < aln.blocks.size(); i++) { o[3 * i] = aln.blocks[i].qPos; o[3 * i + 1] = aln.blocks[i].tPos; o[3 * i + 2] = aln.blocks[i].length; }
      }
      genome.seqs.clear();   // the arena is not ours: keep ~Genome from delete[]-ing it
      read.seq = NULL; read.qual = NULL;
    }
  };
  if (nthreads <= 1) { work(); return 0; }
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; i++) th.emplace_back(work);
  for (auto &x : th) x.join();
  return 0;
}

// Alignment::CalculateStatistics for one segment (Alignment.h:513-531).  stats as in oracle/stats.c, but taken from the
// MEMBERS after one call on a fresh Alignment: stats[2] (n_D) = member nins, stats[3] (n_I) = member ndel (the swap).
int ref_calc_stats(const char *read, int readLen, const char *text, int textLen, const uint32_t *blocks, int nb, int32_t *stats, float *value_out,
                   char *cigar_out, int cigar_cap) {
  Alignment a;
  a.read = (char *)read; a.genome = (char *)text; a.readLen = readLen; a.genomeLen = textLen;
  a.blocks.resize(nb);
  for (int i = 0; i < nb; i++) a.blocks[i] = Block(blocks[3 * i], blocks[3 * i + 1], blocks[3 * i + 2]);
  Options opts;
  std::vector<float> lut;
  for (int i = 1; i <= 10001; i = i + 5) lut.push_back(logf(i));     // LogLookUpTable.h:9-15
  a.CalculateStatistics(opts, NULL, lut);
  stats[0] = a.nm; stats[1] = a.nmm; stats[2] = a.nins; stats[3] = a.ndel; stats[4] = a.tdel; stats[5] = a.tins;
  stats[6] = a.nSmallDel; stats[7] = a.nMedDel; stats[8] = a.nLargeDel; stats[9] = a.nSmallIns; stats[10] = a.nMedIns; stats[11] = a.nLargeIns;
  stats[12] = a.refLen; stats[13] = a.preClip; stats[14] = a.sufClip; stats[15] = 0;
  *value_out = a.value;
  snprintf(cigar_out, cigar_cap, "%s", a.cigar.c_str());
  return 0;
}

static void ref_init_static() {   // lra.cpp:1008-1018 (InitStatic)
  Tuple mask = 1;
  GenomeTuple::for_mask_s = ~(mask << (sizeof(mask) * 8 - 1));
  GenomeTuple::rev_mask_s = (mask << (sizeof(mask) * 8 - 1));
  LocalTuple::for_mask_s = 1;
  for (int i = 1; i < 32 - LOCAL_POS_BITS; i++) { LocalTuple::for_mask_s = LocalTuple::for_mask_s << 1; LocalTuple::for_mask_s += 1; }
  LocalTuple::rev_mask_s = 0;
}

// ---- a12: LocalIndex.  A handle owns one LocalIndex; IndexSeq may be called once per contig (IndexFile, MMIndex.h:246-253).
void *ref_lidx_new(int k, int w, int window, int maxFreq) {
  ref_init_static();
  LocalIndex *li = new LocalIndex(window);
  li->k = k; li->w = w; li->maxFreq = maxFreq;
  return li;
}
void ref_lidx_index_seq(void *h, const char *seq, int len) { ((LocalIndex *)h)->IndexSeq((char *)seq, len); }
void ref_lidx_sizes(void *h, long *n_off, long *n_bnd, long *n_min) {
  LocalIndex *li = (LocalIndex *)h;
  *n_off = (long)li->seqOffsets.size(); *n_bnd = (long)li->tupleBoundaries.size(); *n_min = (long)li->minimizers.size();
}
void ref_lidx_copy(void *h, uint64_t *off, uint64_t *bnd, uint32_t *mins) {
  LocalIndex *li = (LocalIndex *)h;
  memcpy(off, li->seqOffsets.data(), sizeof(uint64_t) * li->seqOffsets.size());
  memcpy(bnd, li->tupleBoundaries.data(), sizeof(uint64_t) * li->tupleBoundaries.size());
  static_assert(sizeof(LocalTuple) == 4, "LocalTuple is one 32-bit word");
  memcpy(mins, li->minimizers.data(), sizeof(uint32_t) * li->minimizers.size());
}
void ref_lidx_free(void *h) { delete (LocalIndex *)h; }

// ---- a13: REFINEclusters on ONE cluster (ClusterRefine.h:50-240).  Arguments as oracle/local_refine.c: lra_oracle_refine_cluster;
// gl / rd_fwd / rd_rev are LocalIndex handles.  mq/mt/box are updated to what the reference leaves in clusters[0].
long ref_refine_cluster(uint32_t *mq, uint32_t *mt, long nm, uint32_t *box, int strand, uint32_t readLen, const uint64_t *hdr_pos, int n_hdr,
                        void *gl, void *rd_fwd, void *rd_rev, int globalK, int smallK, int window, long localMaxFreq,
                        uint32_t *rq, uint32_t *rt, uint32_t *rtup, long cap, int32_t *info, int64_t *diag, float *eff) {
  ref_init_static();
  Options opts; opts.globalK = globalK;
  Options smallOpts = opts; smallOpts.globalK = smallK; smallOpts.window = window; smallOpts.localMaxFreq = (int)localMaxFreq;
  Genome genome;
  genome.header.pos.assign(hdr_pos, hdr_pos + n_hdr);
  Read read; read.length = (int)readLen; read.unaligned = 0;
  std::vector<Cluster> clusters(1), refined(1);
  Cluster &c = clusters[0];
  c.matches.resize(nm);
  for (long i = 0; i < nm; i++) { c.matches[i].first.pos = mq[i]; c.matches[i].second.pos = mt[i]; }
  c.qStart = box[0]; c.qEnd = box[1]; c.tStart = box[2]; c.tEnd = box[3]; c.strand = strand; c.refined = 0;
  LocalIndex *lis[2] = {(LocalIndex *)rd_fwd, (LocalIndex *)rd_rev};
  for (int i = 0; i < 8; i++) info[i] = 0;
  diag[0] = diag[1] = 0; *eff = 0;
  REFINEclusters(clusters, refined, genome, read, *(LocalIndex *)gl, lis, smallOpts, opts);
  if (nm == 0) { info[0] = 1; return 0; }
  if (c.matches.size() == 0) { info[0] = 2; return 0; }
  for (long i = 0; i < nm; i++) { mq[i] = c.matches[i].first.pos; mt[i] = c.matches[i].second.pos; }
  box[0] = c.qStart; box[1] = c.qEnd; box[2] = c.tStart; box[3] = c.tEnd;
  Cluster &r = refined[0];
  info[1] = r.chromIndex; info[6] = (int32_t)r.matches.size();
  diag[0] = c.minDiagNum; diag[1] = c.maxDiagNum;
  long n = (long)r.matches.size();
  for (long i = 0; i < n && i < cap; i++) { rq[i] = r.matches[i].first.pos; rt[i] = r.matches[i].second.pos; rtup[i] = (uint32_t)r.matches[i].first.t; }
  if (n > 0) { info[2] = r.qStart; info[3] = r.qEnd; info[4] = r.tStart; info[5] = r.tEnd; *eff = r.refineEffiency; }
  return n;
}

// canonical (w,k) minimizers of one sequence, in the reference's emission order (MinCount.h:7-179)
long ref_store_minimizers(const char *seq, uint32_t len, int k, int w, uint64_t *t_out, uint32_t *pos_out, long cap) {
  ref_init_static();
  std::vector<GenomeTuple> mm;
  StoreMinimizers<GenomeTuple, Tuple>((char *)seq, len, k, w, mm, true);
  for (size_t i = 0; i < mm.size() && (long)i < cap; i++) { t_out[i] = mm[i].t; pos_out[i] = mm[i].pos; }
  return (long)mm.size();
}

// std::sort with GenomeTuple::operator< (masked key), exactly as MapRead.h:185
void ref_sort_minimizers(uint64_t *t, uint32_t *pos, long n) {
  ref_init_static();
  std::vector<GenomeTuple> mm(n);
  for (long i = 0; i < n; i++) { mm[i].t = t[i]; mm[i].pos = pos[i]; }
  std::sort(mm.begin(), mm.end());
  for (long i = 0; i < n; i++) { t[i] = mm[i].t; pos[i] = mm[i].pos; }
}

// CompareLists<GenomeTuple,Tuple>(query, target, result, opts, Global=true)  (CompareLists.h:148-151): pairs as 4 arrays
long ref_compare_lists(const uint64_t *qt, const uint32_t *qpos, long nq, const uint64_t *tt, const uint32_t *tpos, long nt, int maxFreq,
                       uint64_t *r_qt, uint32_t *r_qpos, uint64_t *r_tt, uint32_t *r_tpos, long cap) {
  ref_init_static();
  std::vector<GenomeTuple> q(nq), t(nt);
  for (long i = 0; i < nq; i++) { q[i].t = qt[i]; q[i].pos = qpos[i]; }
  for (long i = 0; i < nt; i++) { t[i].t = tt[i]; t[i].pos = tpos[i]; }
  Options opts; opts.globalMaxFreq = maxFreq;
  std::vector<std::pair<GenomeTuple, GenomeTuple> > res;
  CompareLists<GenomeTuple, Tuple>(q, t, res, opts, true);
  for (size_t i = 0; i < res.size() && (long)i < cap; i++) {
    r_qt[i] = res[i].first.t; r_qpos[i] = res[i].first.pos; r_tt[i] = res[i].second.t; r_tpos[i] = res[i].second.pos;
  }
  return (long)res.size();
}

// The seeding prefix of MapRead for one read (MapRead.h:169-203): minimizers, sort, CompareLists against the global index,
// strand split by strncmp against the genome (contigs concatenated; the reference resolves the contig through
// Genome::GlobalIndexToSeq, which addresses the same bytes).  strand_out[i] = 0 forward, 1 reverse, in allMatches order.
long ref_seed_read(const char *read, uint32_t len, const char *genome_concat, const uint64_t *tt, const uint32_t *tpos, long nt,
                   int k, int w, int maxFreq, uint64_t *r_qt, uint32_t *r_qpos, uint64_t *r_tt, uint32_t *r_tpos, uint8_t *strand_out, long cap) {
  ref_init_static();
  std::vector<GenomeTuple> mm, t(nt);
  StoreMinimizers<GenomeTuple, Tuple>((char *)read, len, k, w, mm, true);
  std::sort(mm.begin(), mm.end());
  for (long i = 0; i < nt; i++) { t[i].t = tt[i]; t[i].pos = tpos[i]; }
  Options opts; opts.globalMaxFreq = maxFreq;
  std::vector<std::pair<GenomeTuple, GenomeTuple> > res;
  CompareLists<GenomeTuple, Tuple>(mm, t, res, opts, true);
  for (size_t i = 0; i < res.size() && (long)i < cap; i++) {
    r_qt[i] = res[i].first.t; r_qpos[i] = res[i].first.pos; r_tt[i] = res[i].second.t; r_tpos[i] = res[i].second.pos;
    strand_out[i] = strncmp(read + res[i].first.pos, genome_concat + res[i].second.pos, k) == 0 ? 0 : 1;   // MapRead.h:125
  }
  return (long)res.size();
}

// ---- a13, low-accuracy pipeline: Refine_splitchain on ONE split chain (ChainRefine.h:383-576).  The chain's anchors live in clusters
// (cluster_of[i] = index of the cluster anchor i belongs to, n_clusters of them, strand per cluster); arguments otherwise as
// oracle/local_refine.c: lra_oracle_refine_splitchain.
long ref_refine_splitchain(const uint32_t *mq, const uint32_t *mt, const uint32_t *mlen, const int32_t *cluster_of, long n, const uint8_t *cluster_strand,
                           int n_clusters, const uint32_t *box, int chrom, int strand, uint32_t readLen, const uint64_t *hdr_pos, int n_hdr,
                           void *gl, void *rd_fwd, void *rd_rev, int globalK, int smallK, int window, long localMaxFreq, int limitrefine,
                           uint32_t *rq, uint32_t *rt, uint32_t *rtup, long cap, int32_t *info, int64_t *diag, float *eff) {
  ref_init_static();
  Options opts; opts.globalK = globalK; opts.limitrefine = limitrefine != 0;
  Options smallOpts = opts; smallOpts.globalK = smallK; smallOpts.window = window; smallOpts.localMaxFreq = (int)localMaxFreq;
  Genome genome;
  genome.header.pos.assign(hdr_pos, hdr_pos + n_hdr);
  Read read; read.length = (int)readLen; read.unaligned = 0;
  std::vector<Cluster> clusters(n_clusters), refined(1);
  for (int c = 0; c < n_clusters; c++) { clusters[c].strand = cluster_strand[c]; clusters[c].flip = 0; }
  UltimateChain chain(&clusters);
  std::vector<int> sptc;
  std::vector<bool> link;
  for (long i = 0; i < n; i++) {
    Cluster &c = clusters[cluster_of[i]];
    GenomePair gp; gp.first.pos = mq[i]; gp.second.pos = mt[i];
    chain.chain.push_back((unsigned)c.matches.size());
    chain.ClusterIndex.push_back(cluster_of[i]);
    c.matches.push_back(gp); c.matchesLengths.push_back((int)mlen[i]);
    sptc.push_back((int)i);
  }
  std::vector<SplitChain> sps;
  sps.push_back(SplitChain(sptc, link, &chain, strand != 0));
  SplitChain &sp = sps[0];
  sp.QStart = box[0]; sp.QEnd = box[1]; sp.TStart = box[2]; sp.TEnd = box[3]; sp.chromIndex = chrom;
  for (int c = 0; c < n_clusters; c++) sp.ClusterIndex.push_back(c);
  LocalIndex *lis[2] = {(LocalIndex *)rd_fwd, (LocalIndex *)rd_rev};
  for (int i = 0; i < 8; i++) info[i] = 0;
  diag[0] = diag[1] = 0; *eff = 0;
  Refine_splitchain(sps, chain, refined, clusters, genome, read, *(LocalIndex *)gl, lis, smallOpts, opts);
  if (n == 0) { info[0] = 1; return 0; }
  Cluster &r = refined[0];
  info[1] = r.chromIndex; info[6] = (int32_t)r.matches.size();
  diag[0] = r.minDiagNum; diag[1] = r.maxDiagNum;
  long m = (long)r.matches.size();
  for (long i = 0; i < m && i < cap; i++) { rq[i] = r.matches[i].first.pos; rt[i] = r.matches[i].second.pos; rtup[i] = (uint32_t)r.matches[i].first.t; }
  if (m > 0) { info[2] = r.qStart; info[3] = r.qEnd; info[4] = r.tStart; info[5] = r.tEnd; *eff = r.refineEffiency; }
  // the clusters must be back in their original coordinates (ChainRefine.h:554-565)
  for (long i = 0; i < n; i++) {
    const GenomePair &gp = clusters[cluster_of[i]].matches[chain.chain[i]];
    if (gp.first.pos != mq[i] || gp.second.pos != mt[i]) return -1;
  }
  return m;
}

// ---- a6: the anchor sorts of Sorting.h on a vector of GenomePairs (tuple values derived from the positions, as in real anchors)
void ref_sort_matches(int mode, uint32_t *q, uint32_t *t, long n) {
  ref_init_static();
  GenomePairs v(n);
  for (long i = 0; i < n; i++) {
    v[i].first.pos = q[i]; v[i].second.pos = t[i];
    v[i].first.t = (Tuple)q[i] * 0x9E3779B97F4A7C15ull; v[i].second.t = v[i].first.t;
  }
  if (mode == 0) DiagonalSort<GenomeTuple>(v);
  else if (mode == 1) AntiDiagonalSort<GenomeTuple>(v);
  else if (mode == 2) CartesianSort<GenomeTuple>(v);
  else CartesianTargetSort<GenomeTuple>(v.begin(), v.end());
  for (long i = 0; i < n; i++) { q[i] = v[i].first.pos; t[i] = v[i].second.pos; }
}

// ---- a24: GlobalChain<Fragment,Endpoint> (GlobalChain.h:88-189) as the reference's driver calls it (TestGlobalChain.cpp:9-27)
long ref_global_chain(const int32_t *frag, int32_t *score, int32_t *prev, long n, int32_t *chain_out) {
  std::vector<Endpoint> endpoints;
  std::vector<Fragment> fragments;
  std::vector<int> opt;
  for (long i = 0; i < n; i++) fragments.push_back(Fragment(frag[4 * i], frag[4 * i + 1], frag[4 * i + 2], frag[4 * i + 3], score[i], 0));
  GlobalChain(fragments, opt, endpoints);
  for (long i = 0; i < n; i++) { score[i] = fragments[i].score; prev[i] = fragments[i].prev; }
  for (size_t i = 0; i < opt.size(); i++) chain_out[i] = opt[i];
  return (long)opt.size();
}

// ---- a20: RefineBreakpoint (RefineBreakpoint.h:212-462) on two alignments given by their block lists.  The block lists are updated in
// place (capacity cap blocks each); returns 0.
int ref_refine_breakpoint(const char *lread, const char *rread, int readLen, const char *lchrom, int lchromLen, const char *rchrom, int rchromLen,
                          uint32_t *lblocks, int *ln, int lstrand, uint32_t *rblocks, int *rn, int rstrand, int cap) {
  Read read; read.length = readLen;
  Genome genome;
  genome.seqs.push_back((char *)lchrom); genome.lengths.push_back(lchromLen);
  genome.seqs.push_back((char *)rchrom); genome.lengths.push_back(rchromLen);
  Alignment L, R;
  L.read = (char *)lread; R.read = (char *)rread; L.strand = lstrand; R.strand = rstrand; L.chromIndex = 0; R.chromIndex = 1;
  for (int i = 0; i < *ln; i++) L.blocks.push_back(Block(lblocks[3 * i], lblocks[3 * i + 1], lblocks[3 * i + 2]));
  for (int i = 0; i < *rn; i++) R.blocks.push_back(Block(rblocks[3 * i], rblocks[3 * i + 1], rblocks[3 * i + 2]));
  Options opts;
  RefineBreakpoint(read, genome, L, R, opts);
  genome.seqs.clear();
  read.seq = NULL; read.qual = NULL;
  if ((int)L.blocks.size() > cap || (int)R.blocks.size() > cap) return -1;
  *ln = (int)L.blocks.size(); *rn = (int)R.blocks.size();
  for (int i = 0; i < *ln; i++) { lblocks[3 * i] = L.blocks[i].qPos; lblocks[3 * i + 1] = L.blocks[i].tPos; lblocks[3 * i + 2] = L.blocks[i].length; }
  for (int i = 0; i < *rn; i++) { rblocks[3 * i] = R.blocks[i].qPos; rblocks[3 * i + 1] = R.blocks[i].tPos; rblocks[3 * i + 2] = R.blocks[i].length; }
  return 0;
}

// ---- a16: the chain filters of Chain.h:546-960 (modes as in oracle/chain_filters.c).  keep[i] = 1 if anchor i survives.
void ref_chain_filter(int mode, const uint32_t *q, const uint32_t *t, const uint32_t *len, const uint8_t *strand, long n, uint8_t *keep) {
  for (long i = 0; i < n; i++) keep[i] = 0;
  if (mode == 3) {
    GenomePairs matches(n);
    std::vector<unsigned int> chain(n);
    std::vector<int> lengths(n);
    for (long i = 0; i < n; i++) { matches[i].first.pos = q[i]; matches[i].second.pos = t[i]; chain[i] = (unsigned)i; lengths[i] = (int)len[i]; }
    RemovePairedIndels(matches, chain, lengths);
    for (size_t i = 0; i < chain.size(); i++) keep[chain[i]] = 1;
    return;
  }
  WrapChain c;
  c.q.assign(q, q + n); c.t.assign(t, t + n); c.len.assign(len, len + n); c.st.assign(strand, strand + n);
  c.chain.resize(n); c.ClusterIndex.assign(n, 0);
  for (long i = 0; i < n; i++) c.chain[i] = (unsigned)i;
  if (mode == 0) RemoveSmallPairedIndels<WrapChain>(c);
  else if (mode == 1) RemovePairedIndels<WrapChain>(c, true);
  else if (mode == 2) RemovePairedIndels<WrapChain>(c, false);
  else if (mode == 4) RemoveSpuriousAnchors<WrapChain>(c);
  else RemoveSpuriousJump<WrapChain>(c);
  for (size_t i = 0; i < c.chain.size(); i++) keep[c.chain[i]] = 1;
}

// ---- a7: CleanOffDiagonal (Clustering.h:565-800) on the anchors of one read strand.  opt[10] as cod_opts in oracle/clean_off_diagonal.c.
// Outputs the surviving anchors (q, t, freq) and the clusters (cl[7k..] = start, end, qStart, qEnd, tStart, tEnd, chromIndex; cl_freq).
long ref_clean_off_diagonal(uint32_t *q, uint32_t *t, const uint64_t *qt, long n, int strand, const int32_t *opt, const uint64_t *hdr_pos, int n_hdr,
                            float *freq, long *n_kept, int32_t *cl, float *cl_freq) {
  ref_init_static();
  Options opts;
  opts.cleanMaxDiag = opt[0]; opts.minDiagCluster = opt[1]; opts.bypassClustering = opt[2] != 0; opts.cleanClustersize = opt[3];
  opts.SecondCleanMinDiagCluster = opt[4]; opts.punish_anchorfreq = opt[5]; opts.anchorPerlength = opt[6]; opts.SecondCleanMaxDiag = opt[7];
  opts.ExtractDiagonalFromClean = opt[8] != 0; opts.globalK = opt[9];
  opts.readname = "\x01not-a-read-name"; opts.dotPlot = false;
  Genome genome;
  genome.header.pos.assign(hdr_pos, hdr_pos + n_hdr);
  Read read; read.name = "r";
  std::vector<Cluster> clusters;
  GenomePairs matches(n);
  for (long i = 0; i < n; i++) { matches[i].first.pos = q[i]; matches[i].second.pos = t[i]; matches[i].first.t = qt[i]; matches[i].second.t = qt[i]; }
  std::vector<float> mf(n, 0.0f);
  CleanOffDiagonal(genome, clusters, matches, mf, opts, read, strand);
  *n_kept = (long)matches.size();
  for (size_t i = 0; i < matches.size(); i++) { q[i] = matches[i].first.pos; t[i] = matches[i].second.pos; freq[i] = mf[i]; }
  for (size_t k = 0; k < clusters.size(); k++) {
    cl[7 * k] = clusters[k].start; cl[7 * k + 1] = clusters[k].end; cl[7 * k + 2] = (int32_t)clusters[k].qStart; cl[7 * k + 3] = (int32_t)clusters[k].qEnd;
    cl[7 * k + 4] = (int32_t)clusters[k].tStart; cl[7 * k + 5] = (int32_t)clusters[k].tEnd; cl[7 * k + 6] = opts.bypassClustering ? clusters[k].chromIndex : 0;
    cl_freq[k] = clusters[k].anchorfreq;
  }
  read.seq = NULL; read.qual = NULL;
  return (long)clusters.size();
}

// ---- a9: SplitClusters + DecideSplitClustersValue (SplitClusters.h:63-248) on the clusters of one read.  Arguments as
// oracle/split_clusters.c: lra_oracle_split_clusters (mt: genome positions of the anchors, only carried along).
long ref_split_clusters(const uint32_t *box, const uint8_t *strand, const float *freq, long n, int contig, const uint32_t *mq, const uint64_t *m_off, int globalK,
                        uint8_t *split, int32_t *val_cluster, uint32_t *sp, int32_t *sp_val, int32_t *sp_n0, long cap) {
  ref_init_static();
  Options opts; opts.globalK = globalK; opts.readType = contig ? Options::contig : Options::ccs;
  Read read; read.unaligned = 0;
  std::vector<Cluster> clusters(n), splitclusters;
  for (long m = 0; m < n; m++) {
    Cluster &c = clusters[m];
    c.qStart = box[4 * m]; c.qEnd = box[4 * m + 1]; c.tStart = box[4 * m + 2]; c.tEnd = box[4 * m + 3]; c.strand = strand[m]; c.anchorfreq = freq[m]; c.Val = 0;
    for (uint64_t i = m_off[m]; i < m_off[m + 1]; i++) { GenomePair gp; gp.first.pos = mq[i]; gp.second.pos = 0; c.matches.push_back(gp); }
  }
  SplitClusters(clusters, splitclusters, read, opts);
  DecideSplitClustersValue(clusters, splitclusters, opts, read);
  for (long m = 0; m < n; m++) { split[m] = clusters[m].split; val_cluster[m] = clusters[m].Val; }
  const long ns = (long)splitclusters.size();
  for (long k = 0; k < ns && k < cap; k++) {
    const Cluster &c = splitclusters[k];
    sp[6 * k] = c.qStart; sp[6 * k + 1] = c.qEnd; sp[6 * k + 2] = c.tStart; sp[6 * k + 3] = c.tEnd; sp[6 * k + 4] = c.strand; sp[6 * k + 5] = (uint32_t)c.coarse;
    sp_val[k] = c.Val; sp_n0[k] = c.NumofAnchors0;
  }
  read.seq = NULL; read.qual = NULL;
  return ns;
}

// ---- a22: SetFromSegAlignment / AlignmentsOrder::Update / SimpleMapQV (Alignment.h:944-1062, Mapping_ultility.h:497-589) for one read.
// Arguments as oracle/mapq.c: lra_oracle_mapq.
void ref_mapq(int n_groups, const int32_t *seg_off, const float *value, const int32_t *n0, const int32_t *n1, const int32_t *nm, const int32_t *nmm, const int32_t *ndel,
              const int32_t *nins, const uint8_t *strand, const int32_t *update_at, int n_updates, int bypass, int read_type, int K,
              int32_t *flag, int32_t *typeofaln, uint8_t *issec, uint8_t *supp, int32_t *mapq,
              uint8_t *g_issec, float *g_value, int32_t *g_n0, int32_t *g_n1, int32_t *g_nm, int32_t *order) {
  Options opts; opts.bypassClustering = bypass != 0; opts.globalK = K;
  opts.readType = read_type == 0 ? Options::ont : read_type == 1 ? Options::clr : read_type == 2 ? Options::ccs : Options::contig;
  Read read; read.unaligned = 0;
  std::vector<SegAlignmentGroup> alignments;
  alignments.reserve(n_groups + 1);
  AlignmentsOrder ao(&alignments);
  std::vector<Alignment *> all;
  int u = 0;
  for (int g = 0; g <= n_groups; g++) {
    while (u < n_updates && update_at[u] == g) { if (g > ao.Oldend) ao.Update(&alignments); u++; }
    if (g == n_groups) break;
    alignments.resize(alignments.size() + 1);
    for (int s = seg_off[g]; s < seg_off[g + 1]; s++) {
      Alignment *a = new Alignment();
      a->value = value[s]; a->NumOfAnchors0 = n0[s]; a->NumOfAnchors1 = n1[s]; a->nm = nm[s]; a->nmm = nmm[s]; a->ndel = ndel[s]; a->nins = nins[s];
      a->strand = strand[s]; a->flag = flag[s]; a->typeofaln = typeofaln[s]; a->ISsecondary = issec[s]; a->Supplymentary = supp[s];
      alignments.back().SegAlignment.push_back(a); all.push_back(a);
    }
    alignments.back().SetFromSegAlignment(opts);
  }
  SimpleMapQV(ao, read, opts);
  for (size_t s = 0; s < all.size(); s++) {
    flag[s] = all[s]->flag; typeofaln[s] = all[s]->typeofaln; issec[s] = all[s]->ISsecondary; supp[s] = all[s]->Supplymentary; mapq[s] = all[s]->mapqv;
  }
  for (int g = 0; g < n_groups; g++) {
    g_issec[g] = alignments[g].ISsecondary; g_value[g] = alignments[g].value; g_n0[g] = alignments[g].NumOfAnchors0; g_n1[g] = alignments[g].NumOfAnchors1;
    g_nm[4 * g] = alignments[g].nm; g_nm[4 * g + 1] = alignments[g].nmm; g_nm[4 * g + 2] = alignments[g].ndel; g_nm[4 * g + 3] = alignments[g].nins;
  }
  for (size_t i = 0; i < ao.index.size(); i++) order[i] = ao.index[i];
  for (Alignment *a : all) delete a;
  read.seq = NULL; read.qual = NULL;
}

// ---- a15: LinearExtend (GenomePairs overload) + DecideCoordinates + TrimOverlappedAnchors(vector<Cluster>&) for one read, as Map_lowacc.h:132-136
// (skipsorting = 1, trim = 0) and Map_lowacc.h:460-474 (skipsorting = 0, trim = 1) call them.  Arguments as oracle/linear_extend.c.
long ref_linear_extend(const uint8_t *readseq, int read_len, const uint8_t *genome_arena, const uint64_t *chrom_off, const int32_t *chrom_len, int n_groups,
                       const int32_t *g_off, const int32_t *p_off, const uint8_t *p_strand, uint32_t *q, uint32_t *t, int K, int skipsorting, int trim,
                       int32_t *e_off, uint32_t *eq, uint32_t *et, int32_t *elen, uint32_t *box) {
  ref_init_static();
  Options opts; opts.globalK = K;
  Read read; read.seq = (char *)readseq; read.length = read_len; read.unaligned = 0;
  const int n_parts = g_off[n_groups];
  Genome genome;
  for (int p = 0; p < n_parts; p++) { genome.seqs.push_back((char *)genome_arena + chrom_off[p]); genome.lengths.push_back(chrom_len[p]); }
  std::vector<Cluster> ext(n_groups);
  for (int g = 0; g < n_groups; g++) {
    bool st = 0; int chromIndex = 0;
    for (int p = g_off[g]; p < g_off[g + 1]; p++) {
      GenomePairs pairs(p_off[p + 1] - p_off[p]);
      for (size_t i = 0; i < pairs.size(); i++) { pairs[i].first.pos = q[p_off[p] + i]; pairs[i].second.pos = t[p_off[p] + i]; }
      st = p_strand[p]; chromIndex = p;       // every part carries its own contig entry
      LinearExtend(&pairs, ext[g].matches, ext[g].matchesLengths, opts, genome, read, chromIndex, st, skipsorting != 0, K);
      for (size_t i = 0; i < pairs.size(); i++) { q[p_off[p] + i] = pairs[i].first.pos; t[p_off[p] + i] = pairs[i].second.pos; }
    }
    ext[g].qStart = ext[g].qEnd = ext[g].tStart = ext[g].tEnd = 0;
    DecideCoordinates(ext[g], st, chromIndex, 0.0f);
    box[4 * g] = ext[g].qStart; box[4 * g + 1] = ext[g].qEnd; box[4 * g + 2] = ext[g].tStart; box[4 * g + 3] = ext[g].tEnd;
    if (ext[g].matches.size() == 0) ext[g].strand = st;
  }
  if (trim) TrimOverlappedAnchors(ext, 0);
  long no = 0;
  for (int g = 0; g < n_groups; g++) {
    e_off[g] = (int32_t)no;
    for (size_t i = 0; i < ext[g].matches.size(); i++, no++) { eq[no] = ext[g].matches[i].first.pos; et[no] = ext[g].matches[i].second.pos; elen[no] = ext[g].matchesLengths[i]; }
  }
  e_off[n_groups] = (int32_t)no;
  genome.seqs.clear();
  read.seq = NULL; read.qual = NULL;
  return no;
}

// ---- a15 (high-accuracy overload): LinearExtend_chain (LinearExtend.h:782-792 = LinearExtend(vector<Cluster*>, ..., chain, ...) + TrimOverlappedAnchors
// from `start`) for ONE chain, as Map_highacc.h:571-582 calls it with skiprepetitive = 1, then MergeMatchesSameDiag (Map_highacc.h:642).
// Arguments as oracle/linear_extend.c: lra_oracle_linear_extend_chain.  trim = 0 calls LinearExtend alone.  Entries whose cluster has no anchors are
// left out of MergeMatchesSameDiag (the reference would index an empty vector there) and reported as one run (0, 1).
long ref_linear_extend_chain(const uint8_t *readseq, int read_len, const uint8_t *genome_arena, int n_cl, const int32_t *cl_off, uint32_t *cq, uint32_t *ct,
                             const uint32_t *cl_box, const uint8_t *cl_strand, const uint64_t *cl_chrom_off, const int32_t *cl_chrom_len, const float *cl_freq,
                             int n_chain, const int32_t *chain, int K, int skiprepetitive, int trim, int merge_dist,
                             int32_t *e_off, uint32_t *eq, uint32_t *et, int32_t *elen, uint8_t *eovp, uint32_t *box, int32_t *overlap_count,
                             int32_t *md_off, int32_t *md_start, int32_t *md_end) {
  ref_init_static();
  Options opts; opts.globalK = K; opts.merge_dist = merge_dist;
  Read read; read.seq = (char *)readseq; read.length = read_len; read.unaligned = 0;
  Genome genome;
  std::vector<Cluster> cl(n_cl);
  std::vector<Cluster *> clp(n_cl);
  for (int c = 0; c < n_cl; c++) {
    genome.seqs.push_back((char *)genome_arena + cl_chrom_off[c]); genome.lengths.push_back(cl_chrom_len[c]);
    cl[c].qStart = cl_box[4 * c]; cl[c].qEnd = cl_box[4 * c + 1]; cl[c].tStart = cl_box[4 * c + 2]; cl[c].tEnd = cl_box[4 * c + 3];
    cl[c].strand = cl_strand[c]; cl[c].chromIndex = c; cl[c].anchorfreq = cl_freq[c];
    cl[c].matches.resize(cl_off[c + 1] - cl_off[c]);
    for (size_t i = 0; i < cl[c].matches.size(); i++) { cl[c].matches[i].first.pos = cq[cl_off[c] + i]; cl[c].matches[i].second.pos = ct[cl_off[c] + i]; }
    clp[c] = &cl[c];
  }
  std::vector<unsigned int> ch(chain, chain + n_chain);
  std::vector<Cluster> ext(n_chain);
  for (int e = 0; e < n_chain; e++) { ext[e].qStart = ext[e].qEnd = ext[e].tStart = ext[e].tEnd = 0; }
  int overlap = 0;
  if (trim) LinearExtend_chain<unsigned int>(ch, ext, clp, opts, genome, read, 0, overlap, skiprepetitive != 0, K);
  else LinearExtend<unsigned int>(clp, ext, ch, opts, genome, read, 0, overlap, skiprepetitive != 0, K);
  *overlap_count = overlap;
  for (int c = 0; c < n_cl; c++)
    for (size_t i = 0; i < cl[c].matches.size(); i++) { cq[cl_off[c] + i] = cl[c].matches[i].first.pos; ct[cl_off[c] + i] = cl[c].matches[i].second.pos; }
  long no = 0, nmd = 0;
  for (int e = 0; e < n_chain; e++) {
    e_off[e] = (int32_t)no; md_off[e] = (int32_t)nmd;
    for (size_t i = 0; i < ext[e].matches.size(); i++, no++) {
      eq[no] = ext[e].matches[i].first.pos; et[no] = ext[e].matches[i].second.pos; elen[no] = ext[e].matchesLengths[i]; eovp[no] = ext[e].overlap[i];
    }
    box[4 * e] = ext[e].qStart; box[4 * e + 1] = ext[e].qEnd; box[4 * e + 2] = ext[e].tStart; box[4 * e + 3] = ext[e].tEnd;
    if (ext[e].matches.size() == 0) { md_start[nmd] = 0; md_end[nmd] = 1; nmd++; continue; }
    std::vector<Cluster> one(1); one[0] = ext[e];
    std::vector<Cluster_SameDiag> merged;
    MergeMatchesSameDiag(one, merged, opts);
    for (size_t i = 0; i < merged[0].start.size(); i++, nmd++) { md_start[nmd] = merged[0].start[i]; md_end[nmd] = merged[0].end[i]; }
  }
  e_off[n_chain] = (int32_t)no; md_off[n_chain] = (int32_t)nmd;
  genome.seqs.clear();
  read.seq = NULL; read.qual = NULL;
  return no;
}

// ---- a11 (low-accuracy pipeline): SPLITChain(genome, read, UltimateChain&, splitchains, splitchains_link, opts) (Mapping_ultility.h:380-437, with push_new
// and MergeSplitchainINS) followed by RemoveSpuriousSplitChain (Map_lowacc.h:38-66), as Map_lowacc.h:261-262.  Arguments as oracle/split_chain.c; anchor i
// lives in cluster cnum[i] (all anchors of a cluster share its strand).
long ref_split_chain(const uint32_t *q, const uint32_t *t, const int32_t *len, const uint8_t *strand, const int32_t *cnum, const uint8_t *link, int n,
                     const uint64_t *hdr_pos, int n_hdr, int splitdist, int bypass,
                     int32_t *sp_off, int32_t *sptc, uint8_t *sp_lk, int32_t *ci_off, int32_t *ci, uint32_t *sp_box, int32_t *sp_chrom, uint8_t *sp_type,
                     uint8_t *sp_strand, uint8_t *sp_link, int32_t *n_link) {
  ref_init_static();
  Options opts; opts.splitdist = splitdist; opts.bypassClustering = bypass != 0;
  Genome genome; genome.header.pos.assign(hdr_pos, hdr_pos + n_hdr);
  Read read; read.unaligned = 0;
  int ncl = 0;
  for (int i = 0; i < n; i++) if (cnum[i] + 1 > ncl) ncl = cnum[i] + 1;
  std::vector<Cluster> clusters(ncl);
  UltimateChain chain(&clusters);
  chain.chain.resize(n); chain.ClusterIndex.resize(n); chain.link.resize(n > 0 ? n - 1 : 0);
  for (int i = 0; i < n; i++) {
    Cluster &c = clusters[cnum[i]];
    c.strand = strand[i];
    GenomePair gp; gp.first.pos = q[i]; gp.second.pos = t[i];
    chain.chain[i] = (unsigned)c.matches.size(); chain.ClusterIndex[i] = cnum[i];
    c.matches.push_back(gp); c.matchesLengths.push_back(len[i]);
    if (i + 1 < n) chain.link[i] = link[i] != 0;
  }
  std::vector<SplitChain> sp; std::vector<bool> spl;
  SPLITChain(genome, read, chain, sp, spl, opts);
  RemoveSpuriousSplitChain(sp, spl);
  int o = 0, oc = 0;
  for (size_t s = 0; s < sp.size(); s++) {
    sp_off[s] = o; ci_off[s] = oc;
    for (int i = 0; i < sp[s].size(); i++) { sptc[o + i] = sp[s].sptc[i]; sp_lk[o + i] = i + 1 < sp[s].size() ? (uint8_t)sp[s].link[i] : 0; }
    o += sp[s].size();
    for (size_t i = 0; i < sp[s].ClusterIndex.size(); i++) ci[oc++] = sp[s].ClusterIndex[i];
    sp_box[4 * s] = sp[s].QStart; sp_box[4 * s + 1] = sp[s].QEnd; sp_box[4 * s + 2] = sp[s].TStart; sp_box[4 * s + 3] = sp[s].TEnd;
    sp_chrom[s] = sp[s].chromIndex; sp_type[s] = (uint8_t)sp[s].type; sp_strand[s] = sp[s].Strand;
  }
  sp_off[sp.size()] = o; ci_off[sp.size()] = oc;
  for (size_t i = 0; i < spl.size(); i++) sp_link[i] = spl[i];
  *n_link = (int32_t)spl.size();
  read.seq = NULL; read.qual = NULL;
  return (long)sp.size();
}

// ---- a11: MergeChain (ChainRefine.h:767-802) for one split chain; arguments as oracle/chain_glue.c: lra_oracle_merge_chain.
long ref_merge_chain(const int32_t *sp, int n, const int32_t *chrom, const uint8_t *strand, const uint32_t *box, uint8_t *head) {
  ref_init_static();
  int ncl = 0;
  for (int i = 0; i < n; i++) if (sp[i] + 1 > ncl) ncl = sp[i] + 1;
  std::vector<Cluster> cl(ncl); std::vector<Cluster *> clp(ncl);
  for (int c = 0; c < ncl; c++) {
    cl[c].chromIndex = chrom[c]; cl[c].strand = strand[c]; cl[c].qStart = box[4 * c]; cl[c].qEnd = box[4 * c + 1]; cl[c].tStart = box[4 * c + 2]; cl[c].tEnd = box[4 * c + 3];
    clp[c] = &cl[c];
  }
  SplitChain spc, merged;
  spc.sptc.assign(sp, sp + n);
  std::vector<Merge_SplitChain> mergeinfo;
  MergeChain(clp, mergeinfo, merged, spc);
  int t = 0;
  for (size_t g = 0; g < mergeinfo.size(); g++)
    for (size_t i = 0; i < mergeinfo[g].merged_clusterIndex.size(); i++, t++) head[t] = i == 0;
  return (long)mergeinfo.size();
}

// ---- a11: switchindex (Mapping_ultility.h:39-161) for one chain; arguments as oracle/chain_glue.c: lra_oracle_switchindex (n_cl clusters, n_sc split clusters).
long ref_switchindex(int32_t *ch, int n, uint8_t *link, int n_link, const int32_t *coarse, int n_sc, const uint32_t *cq, int n_cl, int32_t *n_link_out) {
  ref_init_static();
  std::vector<Cluster> splitclusters(n_sc), clusters(n_cl);
  for (int i = 0; i < n_sc; i++) splitclusters[i].coarse = coarse[i];
  for (int i = 0; i < n_cl; i++) { clusters[i].qStart = cq[2 * i]; clusters[i].qEnd = cq[2 * i + 1]; }
  std::vector<Primary_chain> pcs(1);
  pcs[0].chains.resize(1);
  CHain &c = pcs[0].chains[0];
  c.ch.assign(ch, ch + n);
  c.link.resize(n_link);
  for (int i = 0; i < n_link; i++) c.link[i] = link[i] != 0;
  Genome genome; Read read; read.unaligned = 0;
  switchindex(splitclusters, pcs, clusters, genome, read);
  CHain &r = pcs[0].chains[0];
  for (size_t i = 0; i < r.ch.size(); i++) ch[i] = (int32_t)r.ch[i];
  for (size_t i = 0; i < r.link.size(); i++) link[i] = r.link[i];
  *n_link_out = (int32_t)r.link.size();
  read.seq = NULL; read.qual = NULL;
  return (long)r.ch.size();
}

// ---- a17 (leaf): RefineByLinearAlignment (LocalRefineAlignment.h:144-185) for one gap of a read (the strand's sequence) against its contig.
long ref_refine_linear(const uint8_t *readseq, int read_len, const uint8_t *contig, int contig_len, uint32_t cur_read_end, uint32_t next_read_start,
                       uint32_t cur_genome_end, uint32_t next_genome_start, int m, int mm, int indel, int local_band, uint32_t *blocks_out, long cap) {
  ref_init_static();
  Options opts; opts.localMatch = m; opts.localMismatch = mm; opts.localIndel = indel; opts.localBand = local_band;
  opts.refineLevel = REF_LOC | REF_DYN | REF_DP; opts.dotPlot = false;
  Read read; read.seq = (char *)readseq; read.length = read_len; read.unaligned = 0;
  Genome genome; genome.seqs.push_back((char *)contig); genome.lengths.push_back(contig_len);
  char *strands[2] = {(char *)readseq, (char *)readseq};
  Alignment aln;
  std::vector<int> scoreMat; std::vector<Arrow> pathMat;
  AffineAlignBuffers buff;
  GenomePos a = cur_read_end, b = cur_genome_end, c = next_read_start, d = next_genome_start;
  RefineByLinearAlignment(a, b, c, d, 0, 0, &aln, read, genome, strands, scoreMat, pathMat, opts, buff);
  const long n = (long)aln.blocks.size();
  for (long i = 0; i < n && i < cap; i++) { blocks_out[3 * i] = aln.blocks[i].qPos; blocks_out[3 * i + 1] = aln.blocks[i].tPos; blocks_out[3 * i + 2] = aln.blocks[i].length; }
  genome.seqs.clear();
  read.seq = NULL; read.qual = NULL;
  return n;
}

// ---- a14 (core): RefineSpace (ClusterRefine.h:242-327) for one space of a read (the sequence of the strand the space is refined on) against its contig.
// Returns the number of EndPairs; *identity = the function's return value.
long ref_refine_space(const uint8_t *strandseq, int read_len, const uint8_t *contig, int contig_len, int K, int W, int refineSpaceDiag, int consider_str,
                      uint32_t qe, uint32_t qs, uint32_t te, uint32_t ts, int st, uint32_t lrts, uint32_t lrlength, int m, int mm, int indel,
                      uint32_t *pq, uint32_t *pt, long cap, float *identity, int localMaxFreq) {
  ref_init_static();
  Options opts; opts.localMaxFreq = localMaxFreq; opts.localMatch = m; opts.localMismatch = mm; opts.localIndel = indel; opts.dotPlot = false;
  Read read; read.seq = (char *)strandseq; read.length = read_len; read.unaligned = 0;
  Genome genome; genome.seqs.push_back((char *)contig); genome.lengths.push_back(contig_len);
  char *strands[2] = {(char *)strandseq, (char *)strandseq};
  GenomePairs EndPairs;
  int chromIndex = 0;
  *identity = RefineSpace(K, W, refineSpaceDiag, consider_str != 0, EndPairs, opts, genome, read, strands, chromIndex, qe, qs, te, ts, st != 0, lrts, lrlength);
  const long n = (long)EndPairs.size();
  for (long i = 0; i < n && i < cap; i++) { pq[i] = EndPairs[i].first.pos; pt[i] = EndPairs[i].second.pos; }
  genome.seqs.clear();
  read.seq = NULL; read.qual = NULL;
  return n;
}

// ---- a17: SwitchToOriginalAnchors (LocalRefineAlignment.h:187-198) for one FinalChain: entry i = (cluster cnum[i], run k[i]); clusters enter as their
// run lists (start / end of cluster c at run_off[c] ..) and coarse.  Returns the length of the UltimateChain.
long ref_switch_to_original(const int32_t *cnum, const int32_t *k, int n, const int32_t *run_off, const int32_t *start, const int32_t *end, const int32_t *coarse,
                            int n_cl, uint32_t *chain_out, int32_t *ci_out) {
  ref_init_static();
  std::vector<Cluster> ext(n_cl);
  std::vector<Cluster_SameDiag> sd(n_cl);
  std::vector<Cluster_SameDiag *> sdp(n_cl);
  for (int c = 0; c < n_cl; c++) {
    sd[c].cluster = &ext[c]; sd[c].coarse = coarse[c];
    sd[c].start.assign(start + run_off[c], start + run_off[c + 1]); sd[c].end.assign(end + run_off[c], end + run_off[c + 1]);
    sdp[c] = &sd[c];
  }
  FinalChain prev(&sdp);
  prev.chain.assign(k, k + n); prev.ClusterIndex.assign(cnum, cnum + n); prev.SecondSDPValue = 0;
  UltimateChain cur;
  SwitchToOriginalAnchors(prev, cur, sdp, ext);
  for (size_t i = 0; i < cur.chain.size(); i++) { chain_out[i] = cur.chain[i]; ci_out[i] = cur.ClusterIndex[i]; }
  return (long)cur.chain.size();
}

// ---- a8 (first half): SplitRoughClustersWithGaps (Clustering.h:1358-1430) over all rough clusters of one anchor list into one vector, as
// MatchesToFineClusters does (Clustering.h:1578-1581).  Arguments as oracle/split_rough.c.
long ref_split_rough(const uint32_t *q, const uint32_t *t, int n_rough, const int32_t *r_start, const int32_t *r_end, const uint32_t *r_box,
                     const uint8_t *r_strand, const float *r_freq, const int32_t *r_chrom, int globalK, int maxGap, int minClusterSize, int maxDiag,
                     int32_t *s_start, int32_t *s_end, uint32_t *s_box, uint8_t *s_strand, int32_t *s_coarse, float *s_freq, int32_t *s_chrom,
                     int32_t *p_cluster, int32_t *p_start, int32_t *p_end, int32_t *n_piece) {
  ref_init_static();
  Options opts; opts.globalK = globalK; opts.RoughClustermaxGap = maxGap; opts.minClusterSize = minClusterSize; opts.maxDiag = maxDiag;
  opts.debug = false; opts.dotPlot = false;
  Read read; read.unaligned = 0;
  int N = 0;
  for (int c = 0; c < n_rough; c++) if (r_end[c] > N) N = r_end[c];
  GenomePairs matches(N);
  for (int i = 0; i < N; i++) { matches[i].first.pos = q[i]; matches[i].second.pos = t[i]; }
  std::vector<Cluster> split;
  for (int c = 0; c < n_rough; c++) {
    Cluster rc(r_start[c], r_end[c], r_box[4 * c], r_box[4 * c + 1], r_box[4 * c + 2], r_box[4 * c + 3], r_strand[c]);
    rc.anchorfreq = r_freq[c]; rc.chromIndex = r_chrom[c];
    int outIter = c;
    SplitRoughClustersWithGaps(matches, rc, split, opts, outIter, read, r_strand[c]);
  }
  long np = 0;
  for (size_t i = 0; i < split.size(); i++) {
    const Cluster &x = split[i];
    s_start[i] = x.start; s_end[i] = x.end; s_box[4 * i] = x.qStart; s_box[4 * i + 1] = x.qEnd; s_box[4 * i + 2] = x.tStart; s_box[4 * i + 3] = x.tEnd;
    s_strand[i] = x.strand; s_coarse[i] = x.coarse; s_freq[i] = x.anchorfreq; s_chrom[i] = x.chromIndex;
    // splitmatchindex as maximal runs of consecutive indices: every piece the reference appends is one run (pieces of one cluster never touch: a
    // boundary exists only where a gap separated them... two appended pieces CAN be adjacent, so runs are cut at the recorded piece ends instead)
    size_t k = 0;
    while (k < x.splitmatchindex.size()) {
      size_t e = k + 1;
      while (e < x.splitmatchindex.size() && x.splitmatchindex[e] == x.splitmatchindex[e - 1] + 1) e++;
      p_cluster[np] = (int32_t)i; p_start[np] = x.splitmatchindex[k]; p_end[np] = x.splitmatchindex[e - 1] + 1; np++;
      k = e;
    }
  }
  *n_piece = (int32_t)np;
  read.seq = NULL; read.qual = NULL;
  return (long)split.size();
}

// ---- StoreDiagonalClusters (Clustering.h:1442-1487) on the cleaned anchors of one strand; arguments as oracle/store_diagonal.c.
long ref_store_diagonal(const uint32_t *q, const uint32_t *t, const uint64_t *qt, const float *freq, int n, int strand, const uint64_t *hdr_pos, int n_hdr,
                        int globalK, int maxDiag, int minClusterSize, int minClusterLength, int bypass,
                        int32_t *c_start, int32_t *c_end, uint32_t *c_box, float *c_freq, int32_t *c_chrom) {
  ref_init_static();
  Options opts; opts.globalK = globalK; opts.maxDiag = maxDiag; opts.minClusterSize = minClusterSize; opts.minClusterLength = minClusterLength;
  opts.bypassClustering = bypass != 0;
  Genome genome; genome.header.pos.assign(hdr_pos, hdr_pos + n_hdr);
  GenomePairs matches(n);
  std::vector<float> mf(freq, freq + n);
  for (int i = 0; i < n; i++) { matches[i].first.pos = q[i]; matches[i].second.pos = t[i]; matches[i].first.t = qt[i]; matches[i].second.t = qt[i]; }
  std::vector<Cluster> clusters;
  StoreDiagonalClusters(genome, mf, matches, clusters, opts, 0, n, strand);
  for (size_t k = 0; k < clusters.size(); k++) {
    c_start[k] = clusters[k].start; c_end[k] = clusters[k].end; c_box[4 * k] = clusters[k].qStart; c_box[4 * k + 1] = clusters[k].qEnd;
    c_box[4 * k + 2] = clusters[k].tStart; c_box[4 * k + 3] = clusters[k].tEnd; c_freq[k] = clusters[k].anchorfreq; c_chrom[k] = clusters[k].chromIndex;
  }
  return (long)clusters.size();
}

// ---- TrimSplitChainDiagonal (ChainRefine.h:189-331) for one split chain and its refined cluster; arguments as oracle/trim_splitchain.c.  Returns the number
// of anchors kept; q/t receive them (the refined cluster's matches after the call), *removed the function's return value.
long ref_trim_splitchain(const uint32_t *cq, const uint32_t *ct, int n_chain, int strand, uint32_t *q, uint32_t *t, int n, long *removed) {
  ref_init_static();
  std::vector<Cluster> cl(1);
  cl[0].strand = strand;
  UltimateChain chain(&cl);
  chain.chain.resize(n_chain); chain.ClusterIndex.assign(n_chain, 0);
  for (int i = 0; i < n_chain; i++) {
    GenomePair gp; gp.first.pos = cq[i]; gp.second.pos = ct[i];
    chain.chain[i] = (unsigned)i; cl[0].matches.push_back(gp); cl[0].matchesLengths.push_back(17);
  }
  std::vector<int> sp(n_chain); std::vector<bool> lk(n_chain > 0 ? n_chain - 1 : 0, false);
  for (int i = 0; i < n_chain; i++) sp[i] = i;
  std::vector<SplitChain> spchain; spchain.push_back(SplitChain(sp, lk, &chain, strand != 0));
  std::vector<Cluster> refined(1);
  refined[0].matches.resize(n);
  for (int i = 0; i < n; i++) { refined[0].matches[i].first.pos = q[i]; refined[0].matches[i].second.pos = t[i]; }
  *removed = TrimSplitChainDiagonal(spchain, refined);
  for (size_t i = 0; i < refined[0].matches.size(); i++) { q[i] = refined[0].matches[i].first.pos; t[i] = refined[0].matches[i].second.pos; }
  return (long)refined[0].matches.size();
}

// ---- a10: the SparseDP drivers of the low-accuracy pipeline (SparseDP.h:2139, :2287; SparseDP_Forward.h:312) -----------------------------
// InitPWL (SubRountine.h:43-101) with the given preset values; the tables it builds are returned for upload.
void ref_init_pwl(float gapopen, float gapextend, float gaproot, int ceil1, int ceil2, int64_t *stops, float *slope, float *inter) {
  InitPWL(gapopen, gapextend, gaproot, ceil1, ceil2);
  for (int i = 0; i < NUMPWL; i++) { stops[i] = STOPS[i]; slope[i] = SLOPE[i]; inter[i] = INTER[i]; }
}
float ref_pwl_w(long x) { return PWL_w(x); }

static void ref_make_clusters(std::vector<Cluster> &cl, int n_cl, const int32_t *cl_off, const uint8_t *cl_strand, const uint32_t *q, const uint32_t *t, const int32_t *len) {
  cl.resize(n_cl);
  for (int c = 0; c < n_cl; c++) {
    cl[c].strand = cl_strand[c]; cl[c].chromIndex = 0; cl[c].matchStart = -1;
    for (int i = cl_off[c]; i < cl_off[c + 1]; i++) {
      GenomePair gp; gp.first.pos = q[i]; gp.second.pos = t[i]; gp.first.t = 0; gp.second.t = 0;
      cl[c].matches.push_back(gp); cl[c].matchesLengths.push_back(len[i]);
    }
  }
}
// Pure matches + DecidePrimaryChains.  Outputs per chain c (< max_aln): chain_len[c], value[c], bounds[4c..] = QStart,QEnd,TStart,TEnd, and at
// chain[c * nfrag ..] the anchors as global fragment indices (cluster offset added back), link[c * nfrag ..] the link bits.  Returns #chains.
int ref_sdp_pure(int n_cl, const int32_t *cl_off, const uint8_t *cl_strand, const uint32_t *q, const uint32_t *t, const int32_t *len, float rate,
                 float alnthres, int NumAln, int read_len, int max_aln, int32_t *chain_len, float *value, uint32_t *bounds, uint32_t *chain, uint8_t *link) {
  ref_init_static();
  std::vector<Cluster> cl; ref_make_clusters(cl, n_cl, cl_off, cl_strand, q, t, len);
  const int nfrag = cl_off[n_cl];
  Options opts; opts.alnthres = alnthres; opts.NumAln = NumAln; opts.readname = "";
  Read read; read.length = read_len; read.name = "r"; read.unaligned = 0;
  std::vector<float> lut;
  std::vector<UltimateChain> chains;
  SparseDP(cl, chains, opts, lut, read, rate);
  int n = 0;
  for (size_t c = 0; c < chains.size() && (int)c < max_aln; c++, n++) {
    chain_len[c] = (int)chains[c].chain.size(); value[c] = chains[c].FirstSDPValue;
    bounds[4 * c] = chains[c].QStart; bounds[4 * c + 1] = chains[c].QEnd; bounds[4 * c + 2] = chains[c].TStart; bounds[4 * c + 3] = chains[c].TEnd;
    for (size_t s = 0; s < chains[c].chain.size(); s++) {
      chain[c * nfrag + s] = chains[c].chain[s] + cl_off[chains[c].ClusterIndex[s]];
      if (s + 1 < chains[c].chain.size()) link[c * nfrag + s] = s < chains[c].link.size() ? (uint8_t)chains[c].link[s] : 255;
    }
  }
  read.seq = NULL; read.qual = NULL; read.passthrough = NULL;
  return n;
}
// One cluster (second SDP of the low-accuracy pipeline).  Returns the chain length; chain holds cluster-local indices.
int ref_sdp_cluster(int n_cl, const int32_t *cl_off, const uint8_t *cl_strand, const uint32_t *q, const uint32_t *t, const int32_t *len, int cluster,
                    float second_anchorbonus, float *value, uint32_t *chain, uint8_t *link) {
  ref_init_static();
  std::vector<Cluster> cl; ref_make_clusters(cl, n_cl, cl_off, cl_strand, q, t, len);
  Options opts; opts.second_anchorbonus = second_anchorbonus;
  Read read; read.length = 1000; read.name = "r"; read.unaligned = 0;
  std::vector<float> lut;
  UltimateChain uc(&cl);
  uc.FirstSDPValue = 0;
  SparseDP(cluster, cl, uc, opts, lut, read);
  *value = uc.FirstSDPValue;
  for (size_t s = 0; s < uc.chain.size(); s++) { chain[s] = uc.chain[s]; if (s < uc.link.size()) link[s] = (uint8_t)uc.link[s]; }
  read.seq = NULL; read.qual = NULL; read.passthrough = NULL;
  return (int)uc.chain.size();
}
// Forward only (third SDP).  Returns the chain length.
int ref_sdp_forward(int n, const uint32_t *q, const uint32_t *t, const int32_t *len, int rate, float *value, uint32_t *chain) {
  ref_init_static();
  GenomePairs gp(n); std::vector<int> ml(n);
  for (int i = 0; i < n; i++) { gp[i].first.pos = q[i]; gp[i].second.pos = t[i]; gp[i].first.t = 0; gp[i].second.t = 0; ml[i] = len[i]; }
  Options opts; std::vector<float> lut; std::vector<unsigned int> ch; float v = 0; int na = 0;
  SparseDP_ForwardOnly(gp, ml, ch, opts, lut, v, na, rate);
  *value = v;
  for (size_t s = 0; s < ch.size(); s++) chain[s] = ch[s];
  return (int)ch.size();
}

}  // extern "C"
