/* lra_b200 -- C ABI of the B200-native MapRead hot path (drop-in boundary).
 *
 * Plain C, no STL / torch types across the boundary.  One context per GPU; calls on one context are serialised on the
 * context's stream (thread-compatible, like one reference worker thread, lra.cpp:103-172).  All functions return
 * LRA_B200_OK (0) or a non-zero error code; the message is available from lra_b200_last_error().  Nothing here ever
 * calls exit() (the reference does, lra.cpp:42-45), and there is no CPU fallback: without a CUDA device every compute
 * entry point fails with LRA_B200_ECUDA.
 *
 * Each entry point names the reference interface it replaces (file:line under the reference tree).
 */
#ifndef LRA_B200_H_
#define LRA_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRA_B200_OK 0
#define LRA_B200_EINVAL 1     /* bad argument / job outside the domain (negative length, window outside its arena) */
#define LRA_B200_ECUDA 2      /* CUDA runtime error or no device */
#define LRA_B200_EOVERFLOW 3  /* caller-provided output capacity too small; required size is reported */
#define LRA_B200_EINTERNAL 4  /* kernel self-check failed (never expected) */

typedef struct lra_b200_ctx lra_b200_ctx;
typedef struct lra_b200_seq lra_b200_seq;

int lra_b200_version(void);

/* Context = one GPU + one stream + grow-only scratch.  Replaces the per-thread state the reference threads through
 * MapRead (IndelRefineBuffers / AffineAlignBuffers / Timing: lra.cpp:106-108, AffineOneGapAlign.h:141-154). */
int lra_b200_create(lra_b200_ctx **ctx, int device);
void lra_b200_destroy(lra_b200_ctx *ctx);
/* Last error message of this context (or of the failed lra_b200_create when ctx == NULL). */
const char *lra_b200_last_error(const lra_b200_ctx *ctx);
/* Run this context's work on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the own stream. */
int lra_b200_set_stream(lra_b200_ctx *ctx, void *cuda_stream);
int lra_b200_synchronize(lra_b200_ctx *ctx);

/* ---- sequences -------------------------------------------------------------------------------------------------
 * A device-resident packed sequence arena: 2 bits per base + a 1-bit mask of non-ACGT symbols, i.e. exactly the
 * comparison alphabet of the reference's seqMapN table (SeqUtils.h:42-75).  Replaces the `char*` read / readRC buffers
 * (Read.h:6-74, MapRead.h:169) and the per-contig genome strings (Genome.h:94-138) on the device side. */
int lra_b200_seq_upload(lra_b200_ctx *ctx, const char *ascii_host, uint64_t n, lra_b200_seq **out);
int lra_b200_seq_from_device(lra_b200_ctx *ctx, const void *ascii_dev, uint64_t n, lra_b200_seq **out);
/* Re-use an existing arena's storage when it is large enough (per-batch read arenas). */
int lra_b200_seq_reupload(lra_b200_ctx *ctx, lra_b200_seq *seq, const char *ascii_host, uint64_t n);
void lra_b200_seq_free(lra_b200_ctx *ctx, lra_b200_seq *seq);
uint64_t lra_b200_seq_length(const lra_b200_seq *seq);
/* Test aid: copy the packed words back (b2: ceil(n/16) words, nmask: ceil(n/32) words). */
int lra_b200_seq_download(lra_b200_ctx *ctx, const lra_b200_seq *seq, uint32_t *b2, uint32_t *nmask);

/* ---- a18  AffineOneGapAlign, batched ---------------------------------------------------------------------------
 * Replaces   int AffineOneGapAlign(string &qSeq, int qLen, string &tSeq, int tLen, int m, int mm, int indel, int k,
 *                                  Alignment &aln, AffineAlignBuffers &b)            AffineOneGapAlign.h:157-649
 * for a whole batch of calls.  Job j aligns q[q_off[j] .. +q_len[j]) against t[t_off[j] .. +t_len[j]) with band
 * parameter k[j]; (match, mismatch, indel) are the reference's (m, mm, indel) = opts.localMatch / localMismatch /
 * localIndel.  Results per job: the return value (score) and the blocks appended to aln.blocks, as (qPos,tPos,length)
 * triples starting at blocks[3*block_off[j]], n_blocks[j] of them, in the reference's order.  Bit-exact.
 * Domain: q_len >= 0, t_len >= 0 (the reference does call it with empty windows), windows inside their arenas (else
 * LRA_B200_EINVAL). */
typedef struct lra_b200_aog_jobs {
  const uint32_t *q_off;
  const uint32_t *t_off;
  const int32_t *q_len;
  const int32_t *t_len;
  const int32_t *k;
  int32_t n_jobs;
  int32_t match, mismatch, indel;
} lra_b200_aog_jobs;

typedef struct lra_b200_aog_result {
  int32_t *score;          /* [n_jobs] */
  int32_t *n_blocks;       /* [n_jobs] */
  uint64_t *block_off;     /* [n_jobs] */
  uint32_t *blocks;        /* [block_cap * 3] */
  uint64_t block_cap;      /* in triples */
  uint64_t n_blocks_total; /* out: triples produced (== required capacity on LRA_B200_EOVERFLOW) */
  uint64_t cells;          /* out: reference-equivalent DP cells of the batch */
} lra_b200_aog_result;

/* Host buffers in, host buffers out (the copies are part of the call). */
int lra_b200_aog_batch(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_aog_jobs *jobs,
                       lra_b200_aog_result *res);
/* Same with every array already resident in device memory (pointers are device pointers). */
int lra_b200_aog_batch_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t,
                              const lra_b200_aog_jobs *jobs_dev, lra_b200_aog_result *res_dev);

/* ---- a19  IndelRefineAlignment: the banded 3-state DP, batched --------------------------------------------------
 * Replaces the per-group body of  void IndelRefineAlignment(Read&, Genome&, Alignment&, const Options&,
 * IndelRefineBuffers&, bool endAlign)   IndelRefine.h:359-745  (matrices :383-431, recurrence :438-622, traceback
 * :626-674, path -> blocks :702-745) for a batch of groups.  A group is one banded window of a segment, as the
 * reference builds it in IndelRefine.h:132-333: rows t = 0..t_len-1 are the target bases tSeq[t_start+t], row t spans
 * the read positions qS[t]..qE[t] (band[band_off .. +t_len) = qS, band[band_off+t_len .. +2 t_len) = qE; absolute read
 * coordinates).  qSeq[i] = q arena [q_base + i], tSeq[i] = t arena [t_base + i] (32-bit wrap-around arithmetic, so a
 * contig-relative t with a window-relative arena works).  Scoring: gap = indel, gapOpen = 2*indel+1, gapExtend = 0.
 * Output per group: the blocks the reference pushes to `refined` for it (including its zero-length blocks), in order.
 * Characters are compared through the packed alphabet: identical to the reference's raw comparison for A,C,G,T,N. */
typedef struct lra_b200_ir_groups {
  const uint32_t *q_base;
  const uint32_t *t_base;
  const int32_t *q_start;
  const int32_t *t_start;
  const int32_t *t_len;
  const int32_t *q_seq_len;
  const int32_t *t_seq_len;
  const uint32_t *band_off;
  const int32_t *band;
  uint64_t band_len;      /* number of int32 in band */
  int32_t n_groups;
  int32_t match, mismatch, indel;
} lra_b200_ir_groups;

typedef struct lra_b200_ir_result {
  int32_t *n_blocks;       /* [n_groups] */
  uint64_t *block_off;     /* [n_groups] */
  uint32_t *blocks;        /* [block_cap * 3] */
  uint64_t block_cap;
  uint64_t n_blocks_total; /* out */
  uint64_t cells;          /* out: DP cells (matSize summed over groups) */
} lra_b200_ir_result;

int lra_b200_indel_dp_batch(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_ir_groups *groups,
                            lra_b200_ir_result *res);
int lra_b200_indel_dp_batch_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t,
                                   const lra_b200_ir_groups *groups_dev, lra_b200_ir_result *res_dev);

/* The whole function, batched over segments:
 *   void IndelRefineAlignment(Read &read, Genome &genome, Alignment &alignment, const Options &opts,
 *                             IndelRefineBuffers &buffers, bool endAlign)                    IndelRefine.h:53-784
 * Segment s has blk_cnt[s] blocks (alignment.blocks, (qPos,tPos,length) triples) starting at triple blk_off[s] of
 * blocks_in (blk_off = exclusive prefix sum of blk_cnt); qSeq = alignment.read = q arena + q_base[s] (the strand the
 * blocks refer to), tSeq = genome.seqs[alignment.chromIndex] = t arena + t_base[s]; read_len = read.length,
 * contig_len = genome.lengths[chromIndex]; refine_band = opts.refineBand, (match, mismatch, indel) = opts.localMatch /
 * localMismatch / localIndel.  Output: the segment's new alignment.blocks.  Small windows go through the batched
 * AffineOneGapAlign exactly as the reference's fallback (IndelRefine.h:344-357). */
typedef struct lra_b200_ir_segments {
  const uint32_t *blocks_in;
  const uint64_t *blk_off;
  const int32_t *blk_cnt;
  const uint32_t *q_base;
  const uint32_t *t_base;
  const int32_t *read_len;
  const int32_t *contig_len;
  uint64_t n_blocks_in;   /* total triples in blocks_in */
  int32_t n_segments;
  int32_t refine_band, match, mismatch, indel, end_align;
} lra_b200_ir_segments;

typedef struct lra_b200_ir_seg_result {
  int32_t *n_blocks;       /* [n_segments] */
  uint64_t *block_off;     /* [n_segments] */
  uint32_t *blocks;        /* [block_cap * 3] */
  uint64_t block_cap;
  uint64_t n_blocks_total; /* out */
  uint64_t cells;          /* out: banded DP cells */
  uint64_t n_dp_groups;    /* out */
  uint64_t n_aog_jobs;     /* out */
} lra_b200_ir_seg_result;

int lra_b200_indel_refine_batch(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_ir_segments *segs,
                                lra_b200_ir_seg_result *res);
int lra_b200_indel_refine_batch_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t,
                                       const lra_b200_ir_segments *segs_dev, lra_b200_ir_seg_result *res_dev);

/* ---- a1-a5  the seeding prefix of MapRead, batched over reads --------------------------------------------------
 * Replaces, for a batch of reads, MapRead.h:169-203:
 *   CreateRC (SeqUtils.h:151)  ->  lra_b200_seq_revcomp
 *   StoreMinimizers<GenomeTuple,Tuple>(read.seq, read.length, opts.globalK, opts.globalW, readmm, true)   MinCount.h:7
 *   sort(readmm.begin(), readmm.end())                                                                   MapRead.h:185
 *   CompareLists<GenomeTuple,Tuple>(readmm, genomemm, allMatches, opts, true)                            CompareLists.h:148
 *   SeparateMatchesByStrand(read, genome, opts.globalK, allMatches, forMatches, revMatches, baseName)    MapRead.h:109
 * The global index image is the array of GenomeTuple{t,pos} of a `<ref>.mms` file (MMIndex.h:402-424), uploaded once.
 * Output: for read r the matches match_off[r] .. match_off[r+1] in the reference's allMatches order, each with the read
 * tuple (q_t, q_pos), the index tuple (t_t, t_pos) and strand = 0 (goes to forMatches) or 1 (revMatches). */
typedef struct lra_b200_index lra_b200_index;
int lra_b200_index_upload(lra_b200_ctx *ctx, const uint64_t *t, const uint32_t *pos, uint64_t n, lra_b200_index **out);
void lra_b200_index_free(lra_b200_ctx *ctx, lra_b200_index *idx);
/* `lra index -MODE` / `lra global` on the device (StoreIndex, MMIndex.h:286-399): canonical (k, w) minimizers of every contig (StoreMinimizers,
 * MinCount.h:7-179) with the contig offset added to pos, sorted by the masked tuple, tuples occurring more than max_freq times removed, at most
 * per_window minimizers kept per win_size window of the genome (lowest multiplicity first, CountSort order), RemoveFrequent.  The index presets
 * (lra.cpp:884-911): -ONT / -CCS 17,10,150,15,1; -CLR 15,10,250,12,1; -CONTIG 19,10,30,20,1.  The result is the image lra_b200_index_upload makes
 * of <ref>.mms; inside a run of equal tuples the entries are ordered by position (std::sort leaves them in introsort's order). */
int lra_b200_gindex_build(lra_b200_ctx *ctx, const lra_b200_seq *genome, const uint64_t *contig_start, const uint32_t *contig_len, int32_t n_contigs, int32_t k, int32_t w,
                          int32_t max_freq, int32_t win_size, int32_t per_window, lra_b200_index **out);
uint64_t lra_b200_index_size(const lra_b200_index *idx);
int lra_b200_index_download(lra_b200_ctx *ctx, const lra_b200_index *idx, uint64_t *t, uint32_t *pos);

/* reads: a packed arena holding the forward strands; read r = [read_off[r], read_off[r] + read_len[r]).  *out_rc must be NULL or an
 * arena of an earlier call, which is then re-used (no allocation in the steady state of a batch loop). */
int lra_b200_seq_revcomp(lra_b200_ctx *ctx, const lra_b200_seq *reads, const uint64_t *read_off, const uint32_t *read_len,
                         int32_t n_reads, lra_b200_seq **out_rc);

typedef struct lra_b200_seed_reads {
  const uint64_t *read_off;
  const uint32_t *read_len;
  int32_t n_reads;
  int32_t k, w;          /* opts.globalK (from the index file), opts.globalW (align preset) */
  int64_t max_freq;      /* opts.globalMaxFreq */
} lra_b200_seed_reads;

typedef struct lra_b200_seed_result {
  uint64_t *match_off;   /* [n_reads + 1] */
  uint64_t *q_t;         /* [match_cap] */
  uint32_t *q_pos;
  uint64_t *t_t;
  uint32_t *t_pos;
  uint8_t *strand;
  uint64_t match_cap;
  uint64_t n_matches;    /* out (== required capacity on LRA_B200_EOVERFLOW) */
  uint32_t *n_minimizers; /* [n_reads], may be NULL */
} lra_b200_seed_result;

int lra_b200_seed_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_index *index,
                        const lra_b200_seed_reads *in, lra_b200_seed_result *res);

/* ---- a21  Alignment::CalculateStatistics, batched over segments ------------------------------------------------
 * Replaces  void Alignment::CalculateStatistics(const Options&, ostream*, const std::vector<float> &LookUpTable)
 * (Alignment.h:513-531, with CreateAlignmentStrings :247-333 and AlignStringsToCigar :414-504; opts.showmm == true, the
 * default) for a batch of segments given as in lra_b200_ir_segments (blocks, blk_off, blk_cnt, q_base, t_base,
 * read_len).  log_lut = the caller's LookUpTable (2001 floats built by CreateLookUpTable with the host logf,
 * LogLookUpTable.h:9-15).  Outputs per segment: CIGAR ops (BAM encoding len<<4|op with '='=7 'X'=8 'I'=1 'D'=2),
 * `value` (the NV tag, bit-identical float) and stats[16] = nm, nmm, n_D, n_I, tdel, tins, nSmallDel, nMedDel, nLargeDel,
 * nSmallIns, nMedIns, nLargeIns, refLen, preClip, sufClip, n_cigar -- the quantities of ONE call; the reference stores
 * n_D into Alignment::nins and n_I into Alignment::ndel (Alignment.h:414 vs :516) and keeps adding tdel..nLargeIns across
 * calls; that bookkeeping stays with the caller. */
typedef struct lra_b200_stats_result {
  int32_t *stats;        /* [n_segments * 16] */
  float *value;          /* [n_segments] */
  uint64_t *cigar_off;   /* [n_segments + 1] */
  uint32_t *cigar;       /* [cigar_cap] */
  uint64_t cigar_cap;
  uint64_t n_cigar_total; /* out */
} lra_b200_stats_result;

int lra_b200_calc_stats_batch(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_ir_segments *segs,
                              const float *log_lut, lra_b200_stats_result *res);
/* the same with the array pointers of segs / res in device memory (log_lut stays a host pointer: 8 KB, uploaded per call);
 * blk_cnt / blk_off may be the n_blocks / block_off outputs of lra_b200_indel_refine_batch_device */
int lra_b200_calc_stats_batch_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_ir_segments *segs_dev,
                                     const float *log_lut, lra_b200_stats_result *res_dev);

/* ---- a6  DiagonalSort / AntiDiagonalSort / CartesianSort / CartesianTargetSort, batched over segments ------------------
 * Replaces the anchor sorts of Sorting.h (:49-75, :111-139, :155-169, :196-209) for any number of anchor lists at once: segment s =
 * anchors seg_off[s] .. seg_off[s+1] of (q, t) = (first.pos, second.pos).  mode 0 DiagonalSort ((long) q - (long) t, then q),
 * 1 AntiDiagonalSort ((GenomePos)(q + t), then q), 2 CartesianSort (q, then t), 3 CartesianTargetSort (t, then q).  q and t are sorted in
 * place; perm (may be NULL) receives, for every output position, the batch-wide index of the anchor that moved there, so that the
 * caller can permute the rest of the GenomePair (tuple values) or any side array (strands, lengths).  Every comparator is a total order
 * on the position pair, so the result is the reference's for any input order. */
int lra_b200_sort_matches_batch(lra_b200_ctx *ctx, int32_t mode, uint32_t *q, uint32_t *t, const uint64_t *seg_off, int32_t n_seg, uint32_t *perm);

/* ---- a7  CleanOffDiagonal, batched over anchor lists ------------------------------------------------------------------------
 * Replaces  void CleanOffDiagonal(Genome&, vector<Cluster> &clusters, vector<pair<Tup,Tup>> &matches, vector<float> &matches_freq,
 *                                 const Options&, Read&, int strand = 0, int diagOrigin = -1, int diagDrift = -1)
 * (Clustering.h:565-800, with SecondRoundCleanOffDiagonal :801-868 and AVGfreq :549-563; every call site leaves diagOrigin / diagDrift at
 * -1) for the forward and reverse anchor lists of a whole batch of reads.  List s = anchors list_off[s] .. list_off[s+1], already sorted by
 * DiagonalSort (strand 0) / AntiDiagonalSort (strand 1): q = first.pos, t = second.pos, qt = first.t (the read tuple AVGfreq counts).
 * Per INPUT anchor: keep (the reference erases the others), freq = matches_freq, cnt = the run counter.  With ExtractDiagonalFromClean the
 * clusters of list s, over its COMPACTED anchors, are cl[7 * (list_off[s] + k) ..] = start, end, qStart, qEnd, tStart, tEnd, chromIndex
 * (chromIndex only with bypassClustering) and cl_freq[list_off[s] + k] = anchorfreq, k < n_cl[s]; with bypassClustering the reference also
 * copies the anchors start .. end into the cluster. */
typedef struct lra_b200_anchor_lists {
  int32_t n_lists;
  const uint32_t *q, *t;
  const uint64_t *qt;
  const uint64_t *list_off;     /* [n_lists + 1] */
  const uint8_t *strand;        /* [n_lists] */
  const uint64_t *hdr_pos;      /* genome.header.pos (needed with bypassClustering) */
  int32_t n_hdr;
} lra_b200_anchor_lists;

typedef struct lra_b200_clean_opts {      /* Options fields read by CleanOffDiagonal */
  int32_t cleanMaxDiag, minDiagCluster, bypassClustering, cleanClustersize, SecondCleanMinDiagCluster, punish_anchorfreq, anchorPerlength,
          SecondCleanMaxDiag, ExtractDiagonalFromClean, globalK;
} lra_b200_clean_opts;

typedef struct lra_b200_clean_result {
  uint8_t *keep;                /* [total anchors] */
  float *freq;
  int32_t *cnt;
  int32_t *cl;                  /* [total anchors * 7] or NULL */
  float *cl_freq;               /* [total anchors] or NULL */
  int32_t *n_cl;                /* [n_lists] */
} lra_b200_clean_result;

int lra_b200_clean_off_diagonal_batch(lra_b200_ctx *ctx, const lra_b200_anchor_lists *al, const lra_b200_clean_opts *opts, lra_b200_clean_result *res);

/* ---- a9  SplitClusters + DecideSplitClustersValue, batched over reads ----------------------------------------------------
 * Replaces  void SplitClusters(vector<Cluster> &clusters, vector<Cluster> &splitclusters, Read&, const Options&)   (SplitClusters.h:63-171)
 * and       void DecideSplitClustersValue(vector<Cluster> &clusters, vector<Cluster> &splitclusters, const Options&, Read&)   (:174-248)
 * for the clusters of a whole batch of reads (Map_highacc.h:154-155).  Read r owns clusters cl_off[r] .. cl_off[r+1]; cluster c: box[4c..] =
 * qStart, qEnd, tStart, tEnd, strand[c], freq[c] = anchorfreq, and the read positions of its anchors m_q[m_off[c] .. m_off[c+1]) in
 * CartesianSort order.  contig = (opts.readType == Options::contig).  Results: split[c] = Cluster::split, val_cluster[c] = Cluster::Val; the
 * pieces of read r are sp_off[r] .. sp_off[r+1] in the reference's order: sp[6k..] = qStart, qEnd, tStart, tEnd, strand, coarse (index of the
 * cluster WITHIN its read), sp_val[k] = Val, sp_n0[k] = NumofAnchors0. */
typedef struct lra_b200_read_clusters {
  int32_t n_reads;
  const uint64_t *cl_off;       /* [n_reads + 1] */
  const uint32_t *box;          /* [clusters * 4] */
  const uint8_t *strand;        /* [clusters] */
  const float *freq;            /* [clusters] */
  const uint64_t *m_off;        /* [clusters + 1] */
  const uint32_t *m_q;
  int32_t contig, global_k;
} lra_b200_read_clusters;

typedef struct lra_b200_split_result {
  uint8_t *split;               /* [clusters] */
  int32_t *val_cluster;         /* [clusters] */
  uint64_t *sp_off;             /* [n_reads + 1] */
  uint32_t *sp;                 /* [piece_cap * 6] */
  int32_t *sp_val, *sp_n0;      /* [piece_cap] */
  uint64_t piece_cap;
  uint64_t n_pieces;            /* out (== required capacity on LRA_B200_EOVERFLOW) */
} lra_b200_split_result;

int lra_b200_split_clusters_batch(lra_b200_ctx *ctx, const lra_b200_read_clusters *rc, lra_b200_split_result *res);

/* ---- a12  LocalIndex::IndexSeq, batched over sequences ---------------------------------------------------------
 * Replaces  void LocalIndex::IndexSeq(char *seq, int seqLen)  (MMIndex.h:200-245; StoreMinimizers_noncanonical
 * MinCount.h:181-338, std::sort on LocalTuple::operator< TupleOps.h:30-32, RemoveFrequent MMIndex.h:69-85) for any number of
 * sequences of one packed arena: the forward strands of a read batch (Map_highacc.h:401, Map_lowacc.h:249), their reverse
 * complements (:402 / :250; the arena made by lra_b200_seq_revcomp), or the contigs of the genome (IndexFile, MMIndex.h:246-253:
 * the image of <ref>.gli).  Sequence s = [seq_start[s], seq_start[s] + seq_len[s]) is cut into windows of `window` bases
 * (<= 2048); the image holds, like the reference's LocalIndex, the window offsets (arena-relative), the tuple boundaries and
 * the LocalTuples (uint32: tuple in bits 0..19, window-relative position in bits 20..31 -- the <ref>.gli layout, MMIndex.h:138-151). */
/* *out must be NULL or an image of an earlier call, which is then rebuilt in place (its device buffers are re-used). */
typedef struct lra_b200_lindex lra_b200_lindex;
int lra_b200_lindex_build(lra_b200_ctx *ctx, const lra_b200_seq *seq, const uint64_t *seq_start, const uint32_t *seq_len, int32_t n_seqs,
                          int32_t k, int32_t w, int32_t window, int32_t max_freq, lra_b200_lindex **out);
/* an image from host arrays (a <ref>.gli read by the caller): win_off[n_win + 1], bnd[n_win + 1], mins[bnd[n_win]] */
int lra_b200_lindex_upload(lra_b200_ctx *ctx, const uint64_t *seq_start, const uint32_t *seq_len, int32_t n_seqs, int32_t window,
                           const uint64_t *win_off, const uint64_t *bnd, const uint32_t *mins, uint64_t n_win, lra_b200_lindex **out);
void lra_b200_lindex_sizes(const lra_b200_lindex *li, uint64_t *n_win, uint64_t *n_mins);
/* copies the image to host arrays: win_off[n_win + 1], bnd[n_win + 1], mins[n_mins] (any may be NULL) */
int lra_b200_lindex_download(lra_b200_ctx *ctx, const lra_b200_lindex *li, uint64_t *win_off, uint64_t *bnd, uint32_t *mins);
void lra_b200_lindex_free(lra_b200_ctx *ctx, lra_b200_lindex *li);

/* ---- a13  REFINEclusters, batched over clusters ------------------------------------------------------------------
 * Replaces  int REFINEclusters(vector<Cluster> &clusters, vector<Cluster> &refinedclusters, Genome&, Read&, LocalIndex &glIndex,
 *                              LocalIndex *localIndexes[2], const Options &smallOpts, const Options &opts)
 * (ClusterRefine.h:50-240) for the clusters of a whole batch of reads.  Cluster c: anchors m_q/m_t[m_off[c] .. m_off[c+1])
 * (read position on the cluster's strand, GLOBAL genome position), box[4c..] = qStart, qEnd, tStart, tEnd, strand[c], and
 * read_id[c] = the sequence index of its read in reads_fwd / reads_rc.  hdr_pos = genome.header.pos (n_hdr = contigs + 1).
 * Outputs per cluster: status (0 refined, 1 skipped: no anchors, 2 dropped by CHROMIndex: spans two contigs), chrom, the
 * diagonal band diag[2c..] = minDiagNum, maxDiagNum, the refined anchors r_off[c] .. r_off[c+1] in the reference's order
 * (r_q on the cluster's strand, r_t chromosome-relative, r_tup the 20-bit tuple), the refined cluster's boundaries rbox and
 * refineEffiency (bit-identical float); m_q_out / m_t_out / box_out (may be NULL) receive what the reference leaves in clusters[ph]
 * (chromosome-relative, forward coordinates, sorted by (t, q)). */
typedef struct lra_b200_clusters {
  int32_t n_clusters;
  const uint32_t *m_q, *m_t;
  const uint64_t *m_off;        /* [n_clusters + 1] */
  const uint32_t *box;          /* [n_clusters * 4] */
  const uint8_t *strand;        /* [n_clusters] */
  const uint32_t *read_id;      /* [n_clusters] */
  const uint64_t *hdr_pos;
  int32_t n_hdr;
  int32_t global_k;             /* opts.globalK */
  int32_t small_k;              /* smallOpts.globalK (= glIndex.k) */
  int32_t window;               /* smallOpts.window */
  int64_t local_max_freq;       /* smallOpts.localMaxFreq */
} lra_b200_clusters;

typedef struct lra_b200_refined {
  int32_t *status, *chrom;      /* [n_clusters] */
  int64_t *diag;                /* [n_clusters * 2] */
  uint64_t *r_off;              /* [n_clusters + 1] */
  uint32_t *r_q, *r_t, *r_tup;  /* [anchor_cap] */
  uint64_t anchor_cap;
  uint64_t n_anchors;           /* out (== required capacity on LRA_B200_EOVERFLOW) */
  uint32_t *rbox;               /* [n_clusters * 4] */
  float *eff;                   /* [n_clusters] */
  uint32_t *m_q_out, *m_t_out;  /* [m_off[n_clusters]] or NULL */
  uint32_t *box_out;            /* [n_clusters * 4] or NULL */
  uint64_t n_units, n_tasks;    /* out: (cluster, genome window) pairs and (cluster, genome window, read window) triples */
} lra_b200_refined;

int lra_b200_refine_clusters_batch(lra_b200_ctx *ctx, const lra_b200_lindex *genome_li, const lra_b200_lindex *reads_fwd,
                                   const lra_b200_lindex *reads_rc, const lra_b200_clusters *cl, lra_b200_refined *res);
/* the same with every array pointer of cl / res (hdr_pos included) in device memory; n_anchors_in = m_off[n_clusters] */
int lra_b200_refine_clusters_batch_device(lra_b200_ctx *ctx, const lra_b200_lindex *genome_li, const lra_b200_lindex *reads_fwd,
                                          const lra_b200_lindex *reads_rc, const lra_b200_clusters *cl_dev, uint64_t n_anchors_in,
                                          lra_b200_refined *res_dev);

/* ---- a13 (low-accuracy pipeline)  Refine_splitchain, batched over split chains ------------------------------------
 * Replaces  int Refine_splitchain(vector<SplitChain> &splitchains, UltimateChain &chain, vector<Cluster> &refinedclusters,
 *                                 vector<Cluster> &clusters, Genome&, Read&, LocalIndex &glIndex, LocalIndex *localIndexes[2],
 *                                 const Options &smallOpts, const Options &opts)
 * (ChainRefine.h:383-576) for the split chains of a whole batch of reads.  Chain c: anchors m_off[c] .. m_off[c+1] IN CHAIN ORDER, each
 * with the read / GLOBAL genome position stored in its cluster (m_q, m_t), its length (matchesLengths, m_len) and the strand of the
 * cluster it comes from (m_strand); box[4c..] = QStart, QEnd, TStart, TEnd; strand[c] = Strand; chrom[c] = chromIndex.  limitrefine =
 * opts.limitrefine (default 1): per genome window the band is [min diagonal of the window's anchors - 100, +inf) -- the reference's
 * upper bound is an uninitialised variable that never filters in the stock build (ChainRefine.h:491-501).  Results as for
 * lra_b200_refine_clusters_batch (status 1: empty chain; diag = the refined cluster's minDiagNum / maxDiagNum; m_q_out / m_t_out /
 * box_out are not written: the reference restores the clusters before it returns). */
typedef struct lra_b200_splitchains {
  int32_t n_chains;
  const uint32_t *m_q, *m_t, *m_len;
  const uint8_t *m_strand;
  const uint64_t *m_off;        /* [n_chains + 1] */
  const uint32_t *box;          /* [n_chains * 4] */
  const uint8_t *strand;        /* [n_chains] */
  const int32_t *chrom;         /* [n_chains] */
  const uint32_t *read_id;      /* [n_chains] */
  const uint64_t *hdr_pos;
  int32_t n_hdr;
  int32_t global_k, small_k, window;
  int64_t local_max_freq;
  int32_t limitrefine;
} lra_b200_splitchains;

int lra_b200_refine_splitchains_batch(lra_b200_ctx *ctx, const lra_b200_lindex *genome_li, const lra_b200_lindex *reads_fwd,
                                      const lra_b200_lindex *reads_rc, const lra_b200_splitchains *sc, lra_b200_refined *res);
int lra_b200_refine_splitchains_batch_device(lra_b200_ctx *ctx, const lra_b200_lindex *genome_li, const lra_b200_lindex *reads_fwd,
                                             const lra_b200_lindex *reads_rc, const lra_b200_splitchains *sc_dev, uint64_t n_anchors_in,
                                             lra_b200_refined *res_dev);

/* ---- a16  chain filters, batched over chains --------------------------------------------------------------------------
 * Replaces, for any number of chains at once (chain c = anchors chain_off[c] .. chain_off[c+1] in chain order: q = qStart, t = tStart,
 * len = length, strand), the remove-mask computation of
 *   mode 0  RemoveSmallPairedIndels<Tup>(chain)                 Chain.h:546-606      mode 3  RemovePairedIndels(matches, chain, lengths)  :754-822
 *   mode 1  RemovePairedIndels<Tup>(chain, refineEnds = true)   Chain.h:611-748      mode 4  RemoveSpuriousAnchors<Tup>(chain)             :828-890
 *   mode 2  RemovePairedIndels<Tup>(chain, refineEnds = false)                       mode 5  RemoveSpuriousJump<Tup>(chain)                :896-960
 * keep[i] = 0 for the anchors the reference removes.  The caller compacts chain.chain / ClusterIndex with it, and link as the reference
 * does: link[m-1] = link[i-1] for every kept anchor i that has m >= 1 kept anchors before it.  strand may be NULL for mode 3. */
int lra_b200_chain_filter_batch(lra_b200_ctx *ctx, int32_t mode, const uint32_t *q, const uint32_t *t, const uint32_t *len, const uint8_t *strand,
                                const uint64_t *chain_off, int32_t n_chains, uint8_t *keep);

/* ---- a15  LinearExtend (low-accuracy pipeline), batched over reads --------------------------------------------------
 * Replaces  void LinearExtend(GenomePairs *pairs, GenomePairs &Extendpairs, vector<int> &ExtendpairsMatchesLength, const Options&, Genome&, Read&,
 *                             int chromIndex, bool strand, bool skipsorting, int K)            (LinearExtend.h:658-716, with Checkbp :50-84)
 * followed by DecideCoordinates (:104-128) and, when trim != 0, TrimOverlappedAnchors(vector<Cluster>&, 0) (:573-647), as the two call sites
 * of the low-accuracy pipeline use them: Map_lowacc.h:132-136 (every cluster, skipsorting = 1, trim = 0) and Map_lowacc.h:460-474 (the
 * refined clusters of one split chain appended into one extended cluster: several parts per group, skipsorting = 0, trim = 1).
 * Group g (one extended cluster) owns parts g_off[g] .. g_off[g+1]; part p (one input cluster) owns anchors p_off[p] .. p_off[p+1] of
 * (q, t) -- t relative to the part's contig, which lies at chrom_off[p] (length chrom_len[p]) of the packed genome -- on strand
 * p_strand[p] of the read at read_off[p] (length read_len[p]) of the read arena.  Strand and contig of a group are those of its last part.
 * Results: extended anchors e_off[g] .. e_off[g+1] as (q, t, len) and box[4g..] = qStart, qEnd, tStart, tEnd (computed before trimming,
 * as in the reference; zeros for a group without anchors).  The result arrays must hold one entry per input anchor (cap >= p_off[n_parts]);
 * n_total = the number produced.  The input anchors are not modified (with skipsorting = 0 the reference leaves them diagonal-sorted). */
typedef struct lra_b200_extend_parts {
  int32_t n_groups;
  const uint64_t *g_off;        /* [n_groups + 1] */
  const uint64_t *p_off;        /* [n_parts + 1] */
  const uint8_t *p_strand;      /* [n_parts] */
  const uint64_t *chrom_off;    /* [n_parts] */
  const uint32_t *chrom_len;
  const uint64_t *read_off;     /* [n_parts] */
  const uint32_t *read_len;
  const uint32_t *q, *t;        /* [p_off[n_parts]] */
  int32_t K;                    /* opts.globalK */
  int32_t skipsorting;          /* 0: DiagonalSort every part first */
  int32_t trim;                 /* 1: TrimOverlappedAnchors(vector<Cluster>&) over every group; 2: its GenomePairs overload (anchors >= 50, forward;
                                   LinearExtend.h:724-777, as LocalRefineAlignment.h:358-373 uses it on the anchors of one gap) */
} lra_b200_extend_parts;

typedef struct lra_b200_extended {
  uint64_t *e_off;              /* [n_groups + 1] */
  uint32_t *q, *t;              /* [cap] */
  int32_t *len;                 /* [cap] */
  uint64_t cap;
  uint64_t n_total;             /* out */
  uint32_t *box;                /* [n_groups * 4] */
} lra_b200_extended;

int lra_b200_linear_extend_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_extend_parts *in,
                                 lra_b200_extended *res);

/* ---- a15  LinearExtend (high-accuracy pipeline: the vector<Cluster*> + chain overload), batched over chains --------------
 * Replaces  LinearExtend_chain(chain, ExtendClusters, RefinedClusters, smallOpts, genome, read, start, overlap, skiprepetitive, K)
 * (LinearExtend.h:782-792 = LinearExtend(vector<Cluster*>, vector<Cluster>&, vector<Tup> &chain, ...) :134-350 with CheckOverlap :87-101 and
 * DecideCoordinates, then TrimOverlappedAnchors(ExtendClusters, start) :573-647; trim = 0 leaves the trimming out) for every chain of a batch
 * (Map_highacc.h:571-582), and MergeMatchesSameDiag (:794-823, Map_highacc.h:642) over the result.
 * Chain k owns entries ch_off[k] .. ch_off[k+1] of ch[] = cluster indices; cluster c owns anchors cl_off[c] .. cl_off[c+1] of (q, t) (t relative
 * to its contig), cl_box[4c..] = qStart, qEnd, tStart, tEnd, strand, anchorfreq, contig and read as in lra_b200_extend_parts.  The anchors need
 * not be sorted (DiagonalSort / AntiDiagonalSort by strand is part of the call; the input arrays are not modified).
 * Results, one extended cluster per chain entry u (= ExtendClusters[start + c]): anchors e_off[u] .. e_off[u+1] as (q, t, len), ovp =
 * Cluster::overlap, md_head = 1 where MergeMatchesSameDiag starts a new run (start[] = positions of the ones, end[] = the next one or the
 * cluster's size), box, overlap[u] = what the entry adds to the reference's `overlap` counter.  An entry whose cluster has no anchors yields an
 * empty extended cluster with a zero box.  cap must be >= the sum over entries of their clusters' sizes (n_total = that sum on
 * LRA_B200_EOVERFLOW, else the number of anchors produced). */
typedef struct lra_b200_extend_chains {
  int32_t n_chains;
  const uint64_t *ch_off;       /* [n_chains + 1] */
  const uint32_t *ch;           /* [ch_off[n_chains]] */
  int32_t n_clusters;
  const uint64_t *cl_off;       /* [n_clusters + 1] */
  const uint32_t *q, *t;        /* [cl_off[n_clusters]] */
  const uint32_t *cl_box;       /* [n_clusters * 4] */
  const uint8_t *cl_strand;     /* [n_clusters] */
  const float *cl_freq;         /* [n_clusters] anchorfreq */
  const uint64_t *chrom_off;    /* [n_clusters] */
  const uint32_t *chrom_len;
  const uint64_t *read_off;     /* [n_clusters] */
  const uint32_t *read_len;
  int32_t K;                    /* the K LinearExtend_chain is called with */
  int32_t skiprepetitive;
  int32_t trim;
  int32_t merge_dist;           /* opts.merge_dist */
} lra_b200_extend_chains;

typedef struct lra_b200_extended_chains {
  uint64_t *e_off;              /* [n_entries + 1] */
  uint32_t *q, *t;              /* [cap] */
  int32_t *len;                 /* [cap] */
  uint8_t *ovp, *md_head;       /* [cap] */
  uint64_t cap;
  uint64_t n_total;             /* out */
  uint32_t *box;                /* [n_entries * 4] */
  int32_t *overlap;             /* [n_entries] */
} lra_b200_extended_chains;

int lra_b200_linear_extend_chains_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_extend_chains *in,
                                        lra_b200_extended_chains *res);

/* ---- a11  SPLITChain on an UltimateChain (low-accuracy pipeline), batched over chains -------------------------------------
 * Replaces  SPLITChain(genome, read, chains[p], spchain, spchain_link, opts); RemoveSpuriousSplitChain(spchain, spchain_link);
 * (Map_lowacc.h:261-262; Mapping_ultility.h:380-437 with push_new :355-378, SplitChain::CHROMIndex Chain.h:388-396, MergeSplitchainINS
 * Mapping_ultility.h:163-264; Map_lowacc.h:38-66).  Chain k owns anchors c_off[k] .. c_off[k+1] in chain order: qStart, tStart (global
 * genome coordinate), length, the strand of the anchor's cluster, UltimateChain::ClusterNum, and link[i] = chain.link between anchors i and
 * i+1.  hdr_pos = genome.header.pos.
 * Results in slot layout (nothing is compacted across chains): chain k produced n_sp[k] split chains and n_link[k] spchain_link bits.  Its
 * piece s has anchors  sptc[c_off[k] + o0 .. c_off[k] + o1)  with o0 = sp_off[c_off[k] + k + s], o1 = sp_off[c_off[k] + k + s + 1]
 * (indices into the chain, already reversed for forward pieces as the reference does for refining), sp_lk at the same places = SplitChain::link
 * (one fewer valid entry than anchors), ClusterIndex = ci[c_off[k] + ci_off[c_off[k] + k + s] ..), and at c_off[k] + s: sp_box[4*..] = QStart, QEnd,
 * TStart, TEnd, sp_chrom, sp_type ('N', 'T' or 'I'), sp_strand; spchain_link = sp_link[c_off[k] ..]. */
typedef struct lra_b200_anchor_chains {
  int32_t n_chains;
  const uint64_t *c_off;        /* [n_chains + 1] */
  const uint32_t *q, *t;        /* [c_off[n_chains]] */
  const int32_t *len;
  const uint8_t *strand;
  const int32_t *cnum;
  const uint8_t *link;
  const uint64_t *hdr_pos;
  int32_t n_hdr;
  int32_t splitdist;            /* opts.splitdist */
  int32_t bypass_clustering;    /* opts.bypassClustering */
} lra_b200_anchor_chains;

typedef struct lra_b200_split_chain_result {
  int32_t *n_sp, *n_link;       /* [n_chains] */
  int32_t *sp_off, *ci_off;     /* [N + n_chains] */
  int32_t *sptc, *ci;           /* [N] */
  uint8_t *sp_lk;               /* [N] */
  uint32_t *sp_box;             /* [N * 4] */
  int32_t *sp_chrom;            /* [N] */
  uint8_t *sp_type, *sp_strand, *sp_link;   /* [N] */
} lra_b200_split_chain_result;

int lra_b200_split_chains_batch(lra_b200_ctx *ctx, const lra_b200_anchor_chains *in, lra_b200_split_chain_result *res);

/* ---- a11  MergeChain, batched over split chains ------------------------------------------------------------------------------
 * Replaces  MergeChain(Refined_Clusters, mergeinfo, merge_spcluster, spcluster)  (ChainRefine.h:767-802; Map_lowacc.h:440).  Split chain k owns
 * entries sc_off[k] .. sc_off[k+1] of sp[] = cluster indices; clusters enter as chromIndex, strand and box[4c..] = qStart, qEnd, tStart, tEnd.
 * head[e] = 1 iff entry e starts a new Merge_SplitChain: mergeinfo[r].merged_clusterIndex = the entries from the r-th 1 up to the next one
 * (these are the groups lra_b200_linear_extend_batch takes), merge_sp.sptc = 0 .. #groups-1. */
int lra_b200_merge_chain_batch(lra_b200_ctx *ctx, const int32_t *sp, const uint64_t *sc_off, int32_t n_chains, const int32_t *chrom, const uint8_t *strand,
                               const uint32_t *box, int32_t n_clusters, uint8_t *head);

/* ---- a11  switchindex, batched over chains -------------------------------------------------------------------------------------
 * Replaces  switchindex(splitclusters, Primary_chains, clusters, genome, read)  (Mapping_ultility.h:39-161; Map_highacc.h:215) for every chain
 * Primary_chains[p].chains[h]: chain k owns entries c_off[k] .. c_off[k+1] of ch[] (split-cluster indices) and link[c_off[k] + i] = its link
 * between entries i and i+1; coarse[] = splitclusters[].coarse, cq[2c], cq[2c+1] = clusters[c].qStart, qEnd.  ch and link are rewritten in
 * place (slot layout): chain k keeps n_out[k] entries (cluster indices) and nl_out[k] links at the start of its slot. */
int lra_b200_switchindex_batch(lra_b200_ctx *ctx, int32_t *ch, uint8_t *link, const uint64_t *c_off, int32_t n_chains, const int32_t *coarse,
                               int32_t n_splitclusters, const uint32_t *cq, int32_t n_clusters, int32_t *n_out, int32_t *nl_out);

/* ---- a17 (leaf)  RefineByLinearAlignment, batched over the gaps between consecutive anchors --------------------------------
 * Replaces  RefineByLinearAlignment(btc_curReadEnd, btc_curGenomeEnd, btc_nextReadStart, btc_nextGenomeStart, str, chromIndex, alignment, read,
 *                                   genome, strands, scoreMat, pathMat, opts, buff)      (LocalRefineAlignment.h:144-185, with SetMatchAndGaps /
 * Matched :89-99, RefineSubstrings :131-142, AlignSubstrings :101-129; opts.refineLevel & REF_DP set, the default) for a batch of gaps:
 * gap g aligns read[cur_read_end .. next_read_start) (the read starts at read_off[g] of `reads`, the arena of its strand) to
 * contig[cur_genome_end .. next_genome_start) (the contig starts at chrom_off[g] of `genome`) with AffineOneGapAlign(localMatch, localMismatch,
 * localIndel, min(2 |qLen - tLen| + 1, localBand)) when Matched() > 0, and returns the blocks the reference appends to alignment->blocks (read /
 * contig coordinates), n_blocks[g] of them from blocks[3 * block_off[g]]; a gap with Matched() <= 0 yields none.  score[g] = the value
 * AlignSubstrings returns (meaningful for aligned gaps only).  Capacity protocol as lra_b200_aog_batch. */
typedef struct lra_b200_linear_gaps {
  int32_t n_gaps;
  const uint32_t *cur_read_end, *next_read_start, *cur_genome_end, *next_genome_start;   /* [n_gaps] */
  const uint32_t *read_off, *chrom_off;                                                    /* [n_gaps] */
  int32_t match, mismatch, indel, local_band;
} lra_b200_linear_gaps;

int lra_b200_refine_linear_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_linear_gaps *in,
                                 lra_b200_aog_result *res);

/* ---- a14 (core)  RefineSpace, batched over spaces -----------------------------------------------------------------------------
 * Replaces  float RefineSpace(K, W, refineSpaceDiag, consider_str, EndPairs, opts, genome, read, strands, ChromIndex, qe, qs, te, ts, st, lrts, lrlength)
 * (ClusterRefine.h:242-327), both branches: a space with qe - qs < 1000 and te - ts + lrlength < 1000 is aligned with AffineOneGapAlign(localMatch,
 * localMismatch, localIndel, 30) and yields the exact K-mers at multiples of K inside blocks longer than K (:262-294) and the identity; a larger space
 * yields the pairs of CompareLists over the sorted non-canonical (K, W) minimizers of its two windows inside the diagonal band
 * [min(0, d) - refineSpaceDiag, max(0, d) + refineSpaceDiag], d = (te - (ts - lrts)) - (qe - qs), with opts.localMaxFreq (:296-305), identity = -1.
 * Space g: read[qs, qe) on the strand whose arena holds the read at read_off[g] (length read_len[g]), contig[ts - lrts, te + lrlength - lrts) with
 * the contig at chrom_off[g]; flip[g] = consider_str && st == 1 (read positions are reported as read_len - pos - K); diag[g] = refineSpaceDiag
 * (read for the larger spaces only; the array may be NULL when there are none).  Results in slot layout: EndPairs of space g =
 * (pq, pt)[pair_off[g] .. pair_off[g] + n_pairs[g]), identity[g] = the return value.
 * pair_cap must cover the slots: min(qLen, tLen) / K + 1 per aligned space plus the pairs of the larger spaces, which are only known after their
 * count pass -- n_pairs_total reports the requirement, also with LRA_B200_EOVERFLOW (call again with at least that). */
typedef struct lra_b200_spaces {
  int32_t n_spaces;
  const uint32_t *qs, *qe, *ts, *te, *lrts, *lrlength;      /* [n_spaces] */
  const uint32_t *read_off, *read_len, *chrom_off;          /* [n_spaces] */
  const uint8_t *flip;                                      /* [n_spaces] */
  int32_t K;                                                /* the K RefineSpace is called with (opts.globalK) */
  int32_t match, mismatch, indel;                           /* opts.localMatch, localMismatch, localIndel */
  int32_t W;                                                /* the W RefineSpace is called with (opts.globalW) */
  int32_t local_max_freq;                                   /* opts.localMaxFreq */
  const int32_t *diag;                                      /* [n_spaces] refineSpaceDiag */
} lra_b200_spaces;

typedef struct lra_b200_space_result {
  uint64_t *pair_off;       /* [n_spaces + 1] */
  int32_t *n_pairs;         /* [n_spaces] */
  float *identity;          /* [n_spaces] */
  uint32_t *pq, *pt;        /* [pair_cap] */
  uint64_t pair_cap;
  uint64_t n_pairs_total;   /* out: slots used */
} lra_b200_space_result;

int lra_b200_refine_space_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_spaces *in,
                                lra_b200_space_result *res);

/* ---- a17  SwitchToOriginalAnchors, batched over FinalChain entries ----------------------------------------------------------
 * Replaces  SwitchToOriginalAnchors(finalchain, ultimatechain, ExtendClusters, extend_clusters)  (LocalRefineAlignment.h:187-198, :576): entry i of the
 * chains of a batch (concatenated) names the same-diagonal run [run_start[i], run_end[i]) = ExtendClusters[c]->start[k], ->end[k] of its extended cluster
 * (the runs lra_b200_linear_extend_chains_batch flags with md_head) and that cluster's `coarse`.  Results: the anchors of entry i are
 * chain[off[i] .. off[i+1]) = run_end-1 .. run_start (descending) with cluster_index[] = coarse[i]; off has n_entries + 1 elements, *n_total = off[n_entries]
 * (the capacity needed on LRA_B200_EOVERFLOW). */
int lra_b200_switch_to_original_batch(lra_b200_ctx *ctx, const int32_t *run_start, const int32_t *run_end, const int32_t *coarse, uint64_t n_entries,
                                      uint64_t *off, uint32_t *chain, int32_t *cluster_index, uint64_t cap, uint64_t *n_total);

/* ---- a8 (first half)  SplitRoughClustersWithGaps, batched over anchor lists ---------------------------------------------------
 * Replaces the loop  for (c ...) SplitRoughClustersWithGaps(Matches, roughClusters[c], split_roughClusters, opts, c, read, strand)
 * (Clustering.h:1578-1581 forward, :1632-1635 reverse; the function :1358-1430 with CloseToPreviousCluster / MergeTwoClusters :1332-1356) for every list
 * of a batch.  List l owns anchors l_off[l] .. l_off[l+1] (Cartesian-sorted inside every rough cluster: lra_b200_sort_matches_batch mode 2 with the
 * rough clusters as segments = the CartesianSort of :1579) and rough clusters lr_off[l] .. lr_off[l+1]: start / end (relative to the list), box,
 * strand, anchorfreq, chromIndex (-1 where CleanOffDiagonal leaves the constructor's value).
 * Results in slot layout, base(l) = l_off[l] + lr_off[l]: split cluster i of list l at base(l) + i = start, end, box, strand, coarse (index of its
 * rough cluster in the list), anchorfreq, chromIndex; its splitmatchindex = the anchor ranges [p_start, p_end) of the pieces j at base(l) + j with
 * p_cluster == i, in order. */
typedef struct lra_b200_rough_lists {
  int32_t n_lists;
  const uint64_t *l_off, *lr_off;       /* [n_lists + 1] */
  const uint32_t *q, *t;
  const int32_t *r_start, *r_end;       /* [lr_off[n_lists]] */
  const uint32_t *r_box;
  const uint8_t *r_strand;
  const float *r_freq;
  const int32_t *r_chrom;
  int32_t globalK, rough_cluster_max_gap, min_cluster_size, max_diag;   /* opts.globalK, RoughClustermaxGap, minClusterSize, maxDiag */
} lra_b200_rough_lists;

typedef struct lra_b200_split_rough_result {
  int32_t *n_split, *n_piece;           /* [n_lists] */
  int32_t *s_start, *s_end, *s_coarse, *s_chrom;   /* [N + n_rough] */
  uint32_t *s_box;                      /* [(N + n_rough) * 4] */
  uint8_t *s_strand;
  float *s_freq;
  int32_t *p_cluster, *p_start, *p_end; /* [N + n_rough] */
} lra_b200_split_rough_result;

int lra_b200_split_rough_batch(lra_b200_ctx *ctx, const lra_b200_rough_lists *in, lra_b200_split_rough_result *res);

/* ---- StoreDiagonalClusters (the cluster builder of CleanMatches when opts.ExtractDiagonalFromClean is off), batched over anchor lists ----------
 * Replaces  StoreDiagonalClusters(genome, Matches_freq, Matches, clusters, opts, 0, Matches.size(), strand)  (Clustering.h:1442-1487; called
 * :1863, :1893 after CleanOffDiagonal, and :1569, :1624) for every list of a batch: list l owns anchors l_off[l] .. l_off[l+1] (cleaned, in diagonal
 * order) with first.pos, second.pos, first.t and matches_freq (lra_b200_clean_off_diagonal_batch's `freq`), on strand[l].
 * Results in slot layout: cluster i of list l at l_off[l] + i = start, end (anchor range inside the list), box (qStart, qEnd, tStart, tEnd),
 * anchorfreq (bit-identical binary32), chromIndex; n_cl[l] clusters. */
typedef struct lra_b200_cleaned_lists {
  int32_t n_lists;
  const uint64_t *l_off;        /* [n_lists + 1] */
  const uint32_t *q, *t;
  const uint64_t *qt;
  const float *freq;
  const uint8_t *strand;        /* [n_lists] */
  const uint64_t *hdr_pos;
  int32_t n_hdr;
  int32_t globalK, max_diag, min_cluster_size, min_cluster_length, bypass_clustering;   /* opts.globalK, maxDiag, minClusterSize, minClusterLength, bypassClustering */
} lra_b200_cleaned_lists;

typedef struct lra_b200_diag_clusters {
  int32_t *n_cl;                /* [n_lists] */
  int32_t *c_start, *c_end, *c_chrom;   /* [N] */
  uint32_t *c_box;              /* [N * 4] */
  float *c_freq;                /* [N] */
} lra_b200_diag_clusters;

int lra_b200_store_diagonal_batch(lra_b200_ctx *ctx, const lra_b200_cleaned_lists *in, lra_b200_diag_clusters *res);

/* ---- TrimSplitChainDiagonal, batched over split chains ------------------------------------------------------------------------
 * Replaces  TrimSplitChainDiagonal(spchain, refined_clusters)  (ChainRefine.h:189-331; Map_lowacc.h:331, right after Refine_splitchain): split chain c owns
 * chain anchors c_off[c] .. c_off[c+1] (cq / ct = SplitChain::qStart(i) / tStart(i) in sptc order) with SplitChain::Strand strand[c], and the refined anchors
 * m_off[c] .. m_off[c+1] of (q, t) = refined_clusters[c].matches.  q and t come back in the order the reference leaves them (CartesianSort; untouched for a
 * chain of one anchor), keep[i] = 0 marks the anchors it erases, removed[c] = its contribution to the return value.  A split chain WITHOUT chain anchors is accepted
 * on the forward strand (its refined anchors are only sorted, as the reference does) and refused with LRA_B200_EINVAL on the reverse strand, where the reference reads
 * splitchain[size - 1] of an empty chain. */
int lra_b200_trim_splitchains_batch(lra_b200_ctx *ctx, const uint32_t *cq, const uint32_t *ct, const uint64_t *c_off, const uint8_t *strand, int32_t n_chains,
                                    uint32_t *q, uint32_t *t, const uint64_t *m_off, uint8_t *keep, int32_t *removed);

/* ---- a20  RefineBreakpoint, batched over pairs of adjacent segments --------------------------------------------------
 * Replaces  void RefineBreakpoint(Read &read, Genome &genome, Alignment &leftAln, Alignment &rightAln, const Options &opts)
 * (RefineBreakpoint.h:212-462; called for consecutive segments of a split read, Map_highacc.h:725, Map_lowacc.h:592).  Pair p: the left /
 * right alignment enter through their first and last block (lf, ll, rf, rl: q, t, len -- GetQStart/GetQEnd/GetTStart/GetTEnd and the
 * block Prepend/AppendBlocks may merge into), their strand, and the contig they lie on (arena position and length in the packed genome);
 * the read is at read_off[p] in both strand arenas (Alignment::read is one or the other).  Results per side s in {0 left, 1 right}:
 * mode[2p+s] = 0 nothing (the two segments are not within MAX_GAP = 500 of each other), 1 AppendBlocks, 2 PrependBlocks;
 * out[(2p+s) * 512 * 3 ..] the n_out[2p+s] blocks to splice in, bound[(2p+s) * 3 ..] the alignment's boundary block (the last block for an
 * append, the first for a prepend) after a possible gapless merge; refined[p] = 1 if the pair was refined. */
typedef struct lra_b200_breakpoints {
  int32_t n_pairs;
  const uint32_t *lf, *ll, *rf, *rl;           /* [n_pairs * 3] */
  const uint8_t *lstrand, *rstrand;            /* [n_pairs] */
  const uint64_t *read_off;                    /* [n_pairs] */
  const uint32_t *read_len;
  const uint64_t *lchrom_off, *rchrom_off;     /* [n_pairs] */
  const uint32_t *lchrom_len, *rchrom_len;
} lra_b200_breakpoints;

typedef struct lra_b200_breakpoint_result {
  int32_t *mode, *n_out;        /* [n_pairs * 2] */
  uint32_t *bound;              /* [n_pairs * 2 * 3] */
  uint32_t *out;                /* [n_pairs * 2 * 512 * 3] */
  int32_t *refined;             /* [n_pairs] */
} lra_b200_breakpoint_result;

int lra_b200_refine_breakpoint_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads_fwd, const lra_b200_seq *reads_rc, const lra_b200_seq *genome,
                                     const lra_b200_breakpoints *bp, lra_b200_breakpoint_result *res);

/* ---- a22  SetFromSegAlignment + AlignmentsOrder::Update + SimpleMapQV, batched over reads ------------------------------------
 * Replaces, for every read of a batch, the tail of MapRead after the statistics: SegAlignmentGroup::SetFromSegAlignment (Alignment.h:944-983)
 * for each alignment, AlignmentsOrder::Update (Alignment.h:1024-1062) at the points the pipeline calls it, and SimpleMapQV
 * (Mapping_ultility.h:497-589).  Read r owns alignments grp_off[r] .. grp_off[r+1]; alignment g owns segments seg_off[g] .. seg_off[g+1] (the
 * Alignment members value, NumOfAnchors0 / 1, nm, nmm, ndel, nins, strand).  update_at[upd_off[r] .. upd_off[r+1]) = how many of the read's
 * alignments exist at each Update call, ascending (Map_lowacc.h:609: one entry, the total; Map_highacc.h:737,743: one per primary chain, then
 * the total again); an Update that finds nothing new is a no-op (the reference reads past its index vector there).  bypass_clustering /
 * read_type (0 ont, 1 clr, 2 ccs, 3 contig) / global_k are the Options fields SimpleMapQV reads (smallOpts at the call sites).
 * In / out per segment: flag, typeofaln, issec (ISsecondary), supp (Supplymentary); out: mapq (mapqv).  Out per alignment: g_issec, g_value,
 * g_n0, g_n1, g_nm[4] (nm, nmm, ndel, nins) and order (AlignmentsOrder::index, positions within the read).  The two logf terms are evaluated
 * on the host with its libm, as in the reference. */
typedef struct lra_b200_alignment_groups {
  int32_t n_reads;
  const int32_t *grp_off;       /* [n_reads + 1] */
  const int32_t *seg_off;       /* [alignments + 1] */
  const int32_t *upd_off;       /* [n_reads + 1] */
  const int32_t *update_at;
  const float *value;           /* per segment */
  const int32_t *n0, *n1, *nm, *nmm, *ndel, *nins;
  const uint8_t *strand;
  int32_t bypass_clustering, read_type, global_k;
} lra_b200_alignment_groups;

typedef struct lra_b200_mapq_result {
  int32_t *flag, *typeofaln;    /* per segment, in / out */
  uint8_t *issec, *supp;        /* per segment, in / out */
  int32_t *mapq;                /* per segment */
  uint8_t *g_issec;             /* per alignment */
  float *g_value;
  int32_t *g_n0, *g_n1, *g_nm /* [alignments * 4] */, *order;
} lra_b200_mapq_result;

int lra_b200_mapq_batch(lra_b200_ctx *ctx, const lra_b200_alignment_groups *ag, lra_b200_mapq_result *res);

/* ---- a24  GlobalChain over a PrioritySearchTree, batched over independent problems --------------------------------------
 * Replaces  int GlobalChain(vector<T_Fragment> &fragments, vector<int> &optFragmentChainIndices, vector<T_Endpoint> &endpoints)
 * (GlobalChain.h:88-189, with PrioritySearchTree.h:47-291) as the reference's driver TestGlobalChain.cpp:9-27 calls it.  Problem p:
 * fragments frag_off[p] .. frag_off[p+1], frag[4i..] = xl, yl, xh, yh.  score: in = each fragment's own score, out = the score of the
 * best chain ending in it (Fragment::score); prev: Fragment::prev (problem-local index, -1 = none); chain[frag_off[p] ..] = the optimal
 * chain of problem p, first fragment first, chain_len[p] fragments.  (`lra` itself never calls GlobalChain.) */
int lra_b200_global_chain_batch(lra_b200_ctx *ctx, const int32_t *frag, const uint64_t *frag_off, int32_t n_prob, int32_t *score, int32_t *prev,
                                int32_t *chain, int32_t *chain_len);

/* ---- a10  the SparseDP family, batched (one problem per warp) ------------------------------------------------------
 * Replaces, for a batch of calls,
 *   mode 0:  int SparseDP(vector<Cluster> &FragInput, vector<UltimateChain> &chains, const Options&, const vector<float>&, Read&, float rate)
 *            SparseDP.h:2139-2282 (first SDP of MapRead_lowacc, Map_lowacc.h:188) incl. DecidePrimaryChains :1658-1765
 *   mode 1:  int SparseDP(int ClusterIndex, vector<Cluster> &FragInput, UltimateChain&, const Options&, const vector<float>&, Read&)
 *            SparseDP.h:2287-2440 (second SDP, Map_lowacc.h:535)
 *   mode 2:  int SparseDP_ForwardOnly(const GenomePairs&, const vector<int> &MatchLengths, vector<unsigned int> &chain, ..., int rate)
 *            SparseDP_Forward.h:312-490 (third SDP, LocalRefineAlignment.h:377)
 *   mode 3:  int SparseDP(SplitChain &inputChain, vector<Cluster_SameDiag*> &FragInput, FinalChain&, ...)   SparseDP.h:1766-1952 (second SDP of
 *            MapRead_highacc, LocalRefineAlignment.h:563): the same-diagonal anchors (GetqStart, GettStart, length) of the clusters of the split chain,
 *            in split-chain order, as anchors grouped by cl_off / cl_strand; rate = opts.second_anchorbonus; chain = indices into the concatenation
 *   mode 4:  int SparseDP(vector<Cluster> &FragInput, vector<Primary_chain>&, ..., float &rate)   SparseDP.h:1956-2135 (first SDP of MapRead_highacc on
 *            the split clusters, Map_highacc.h:229) incl. DecidePrimaryChains :1586-1652: every fragment is a box with four points; chains of
 *            Primary_chains[0] with value, box and NumOfAnchors0
 * Problem p owns the anchors frag_off[p] .. frag_off[p+1] (q = first.pos, t = second.pos, len = matchesLengths), grouped into clusters by
 * cl_off[cl_off_off[p] ..] (ncl + 1 problem-relative offsets) with strands cl_strand[cl_off_off[p] ..].  rate = the anchor bonus (mode 0:
 * the `rate` argument, mode 1: opts.second_anchorbonus), irate = the integer rate of mode 2.  The gap cost is the reference's PWL_w over
 * the tables InitPWL builds (SubRountine.h:43-126): build them on the host with lra_b200_init_pwl and pass them in.
 * Output: chain c (< max_aln) of problem p has chain_len[p*max_aln+c] anchors at chain[max_aln*frag_off[p] + c*nfrag_p ..] (mode 0: problem-
 * relative fragment index = MatchStart[cluster] + index in cluster; modes 1, 2: index in the cluster / list), link bits beside them,
 * chain_val = FirstSDPValue / max_value / inv_value, bounds = QStart,QEnd,TStart,TEnd (mode 0).  Bit-exact incl. the float value. */
typedef struct lra_b200_sdp_problems {
  int32_t n_prob, max_aln;
  const int32_t *mode;
  const uint64_t *frag_off;
  const uint32_t *q, *t;
  const int32_t *len;
  const uint64_t *cl_off_off;
  const int32_t *cl_off;
  const uint8_t *cl_strand;
  const int32_t *only_cl;
  const float *rate;
  const int32_t *irate;
  const int32_t *read_len;
  float alnthres;          /* opts.alnthres */
  int32_t num_aln;         /* opts.NumAln */
  const int64_t *pwl_stops; const float *pwl_slope; const float *pwl_inter;   /* 25 entries each */
  int32_t ceil1, ceil2;    /* opts.gapCeiling1 / 2 */
  /* mode 4 only (may be NULL otherwise), one entry per fragment at frag_off: q / t above are the box starts (qStart, tStart) */
  const uint32_t *q_end, *t_end;   /* Cluster::qEnd, tEnd */
  const uint8_t *frag_strand;      /* Cluster::strand */
  const float *frag_val;           /* Cluster::Val */
  const int32_t *frag_n0;          /* Cluster::NumofAnchors0 */
  int32_t global_k;                /* opts.globalK (the value threshold of DecidePrimaryChains, SparseDP.h:1593) */
} lra_b200_sdp_problems;
typedef struct lra_b200_sdp_result {
  int32_t *n_chains;       /* [n_prob] */
  int32_t *chain_len;      /* [n_prob * max_aln] */
  float *chain_val;        /* [n_prob * max_aln] */
  uint32_t *bounds;        /* [n_prob * max_aln * 4] */
  uint32_t *chain;         /* [max_aln * total anchors] */
  uint8_t *link;           /* [max_aln * total anchors] */
  int32_t *cl_of_frag;     /* [total anchors] (mode 0: cluster of every anchor) */
  uint64_t arena_peak;     /* out: largest per-problem scratch use in bytes */
  int32_t *num_anchors0;   /* [n_prob * max_aln] mode 4: CHain::NumOfAnchors0 (may be NULL) */
} lra_b200_sdp_result;
int lra_b200_sdp_batch(lra_b200_ctx *ctx, const lra_b200_sdp_problems *problems, lra_b200_sdp_result *res);
/* InitPWL (SubRountine.h:43-101): the piece-wise-linear gap cost tables for (opts.gapopen, opts.gapextend, opts.gaproot), 25 entries each. */
int lra_b200_init_pwl(float intercept, float scalar, float root, int32_t ceil1, int32_t ceil2, int64_t *stops, float *slope, float *inter);

/* ---- the MapRead seam: reads in, alignment records out -------------------------------------------------
 * Replaces the body of the worker loop  numAligned += MapRead(LookUpTable, read, genome, genomemm, glIndex, opts, &strm, ...)
 * (lra.cpp:115-121; MapRead.h:153-263 -> MapRead_lowacc, Map_lowacc.h:69-632) for a batch of reads.  The low-accuracy presets
 * (-ONT, -CLR) are implemented; the record is the per-segment POD the printers consume (Alignment.h:591-808), so the host formats
 * SAM text (lra_b200_format_sam). */
typedef struct lra_b200_map_opts {      /* the Options members the path reads (Options.h:8-241) after the align preset (lra.cpp:268-431) */
  int32_t globalK, globalW, globalMaxFreq;
  int32_t localW, localMaxFreq;
  int32_t smallK, smallW;                /* k / w of the LocalIndex (<ref>.gli): smallOpts.globalK / globalW, Map_lowacc.h:232-233 */
  int32_t cleanMaxDiag, minDiagCluster, cleanClustersize, SecondCleanMinDiagCluster, SecondCleanMaxDiag, punish_anchorfreq, anchorPerlength;
  int32_t NumAln, PrintNumAln, splitdist, readType;      /* readType: 0 ont, 1 clr, 2 ccs, 3 contig */
  float initial_anchorbonus, second_anchorbonus, alnthres, anchorstoosparse;
  int32_t refineSpaceDist, window, limitrefine, RefineBySDP;
  int32_t localMatch, localMismatch, localIndel, localBand, refineBand;
  int32_t hardClip, bypassClustering;
  /* host-side only (not read by the kernels) */
  float gapopen, gapextend, gaproot;
  int32_t gapCeiling1, gapCeiling2;
  int32_t localIndexWindow, localIndexMaxFreq;
  /* read by MapRead_highacc only (-CCS, -CONTIG; Map_highacc.h:37-798) */
  int32_t HighlyAccurate, maxDiag, maxGap, RoughClustermaxGap, minClusterSize, minUniqueStretchNum, minUniqueStretchDist, merge_dist;
} lra_b200_map_opts;
/* the align preset of `lra align -ONT | -CLR | -CCS | -CONTIG` (lra.cpp:268-431) on top of the defaults (Options.h:123-240); globalK is overwritten
 * by the value stored in <ref>.mms (MMIndex.h:409), smallK / smallW by the <ref>.gli header */
int lra_b200_map_opts_preset(const char *mode, lra_b200_map_opts *opts);

typedef struct lra_b200_record {         /* one segment (an Alignment) */
  int32_t read, chain, seg, n_seg;       /* read of the batch, alignment slot (chain p), index s in SegAlignment, segments of that alignment */
  uint32_t flag;                         /* SAM flag (READ_REVERSE | READ_SECONDARY | READ_SUPPLEMENTARY) */
  int32_t chrom, strand, mapq, order, typeofaln, supplementary;
  uint32_t tStart, tEnd, qStart, qEnd;
  int32_t preClip, sufClip;
  int32_t nm, nmm, nins, ndel, tins, tdel, nSmallDel, nMedDel, nLargeDel, nSmallIns, nMedIns, nLargeIns;   /* members of Alignment (nins / ndel swapped as in the reference) */
  float value;
  int32_t NumOfAnchors0, NumOfAnchors1;
  int32_t n_blocks, n_cigar;
  uint64_t cigar_off;                    /* into the cigar array: BAM-encoded ops (len << 4 | op; '=' 7, 'X' 8, 'I' 1, 'D' 2) */
} lra_b200_record;

typedef struct lra_b200_map_result {     /* caller-owned host buffers; *_cap in elements */
  int32_t *status;                       /* [n_reads] 0 mapped, 1 unaligned, >= 2 internal capacity error (the read is reported unaligned) */
  int32_t *n_aln;                        /* [n_reads] alignments.size() */
  int32_t *aln_nseg, *aln_seg0, *aln_rank;   /* [n_reads * 4] per alignment slot: segments, first record, and the slot printed a-th (AlignmentsOrder) */
  lra_b200_record *records; uint64_t record_cap; uint64_t n_records;
  uint32_t *cigar; uint64_t cigar_cap; uint64_t n_cigar;
  uint64_t aligned_bases;                /* out: sum of qEnd - qStart over the non-supplementary, non-secondary segments */
} lra_b200_map_result;

typedef struct lra_b200_mapper lra_b200_mapper;
/* genome: contigs concatenated (ASCII, upper-cased as Genome::Read does); hdr_pos[n_contigs + 1] cumulative offsets; the <ref>.mms tuples (t, pos);
 * the <ref>.gli arrays (seq_offsets / tuple_boundaries have n_regions entries, MMIndex.h:138-151). */
int lra_b200_mapper_create(lra_b200_ctx *ctx, const lra_b200_map_opts *opts, const char *genome_ascii, uint64_t genome_len, const uint64_t *hdr_pos, int32_t n_contigs,
                           const uint64_t *mms_t, const uint32_t *mms_pos, uint64_t n_mms, const uint64_t *gli_seq_offsets, const uint64_t *gli_tuple_boundaries,
                           int32_t gli_n_regions, const uint32_t *gli_minimizers, uint64_t gli_n_min, lra_b200_mapper **out);
void lra_b200_mapper_destroy(lra_b200_ctx *ctx, lra_b200_mapper *m);
/* reads: ASCII bases of all reads concatenated (upper case), read r = [read_off[r], +read_len[r]).  Records of read r come back at
 * aln_seg0 .. in SegAlignment order. */
int lra_b200_map_batch(lra_b200_ctx *ctx, lra_b200_mapper *m, const char *reads_ascii, uint64_t reads_len, const uint64_t *read_off, const uint32_t *read_len,
                       int32_t n_reads, lra_b200_map_result *res);
/* The same in three steps, for callers that keep batches resident (double buffering, multi-GPU shards; bench.py's `value` leg):
 * lra_b200_readset_upload  host ASCII reads -> packed forward strands + reverse complements + descriptors in HBM (H2D inside);
 * lra_b200_map_resident    every kernel of the path over a resident batch, records left in the mapper's device buffers;
 * lra_b200_map_download    records of the last lra_b200_map_resident call -> caller-owned host buffers (D2H inside).
 * lra_b200_map_batch == upload + resident + download.  `reads_ascii` and the arrays of lra_b200_map_result may be host or device pointers (the copies
 * use cudaMemcpyDefault), so a shard received over NCCL is mapped, and its records sent back, without a host bounce; read_off / read_len are host arrays. */
typedef struct lra_b200_readset lra_b200_readset;
int lra_b200_readset_upload(lra_b200_ctx *ctx, const char *reads_ascii, uint64_t reads_len, const uint64_t *read_off, const uint32_t *read_len, int32_t n_reads,
                            lra_b200_readset **out);
void lra_b200_readset_free(lra_b200_ctx *ctx, lra_b200_readset *rs);
int lra_b200_map_resident(lra_b200_ctx *ctx, lra_b200_mapper *m, const lra_b200_readset *rs);
int lra_b200_map_download(lra_b200_ctx *ctx, lra_b200_mapper *m, lra_b200_map_result *res);
/* SAM text of the batch in input order (Alignment::PrintSAM, Alignment.h:658-808; unaligned reads: SimplePrintSAM :811-832; order of the
 * segments: OUTPUT, Mapping_ultility.h:465-494).  names: n_reads NUL-terminated strings back to back; contig_names likewise.  Returns the
 * number of bytes written, or -(required size) when cap is too small.  Pure host code. */
int64_t lra_b200_format_sam(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                            const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, int32_t n_contigs, int32_t runtime, char *out, int64_t cap);

/* The other printers of OUTPUT (Mapping_ultility.h:465-494): fmt 's' SAM, 'p' PAF (Alignment::PrintPAF, Alignment.h:600-656; `-p p`, the reference's
 * default), 'c' PAF with CG:z: (`-p pc`), 'b' BED (PrintBed :591-598).  contig_len[n_contigs] is the `genomeLen` column of PAF. */
int64_t lra_b200_format_records(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, const uint64_t *contig_len, int32_t n_contigs,
                                int32_t fmt, int32_t runtime, char *out, int64_t cap);

/* The same with the QUAL column of FASTQ input (quals_ascii: one quality string per read at the offsets of reads_ascii, or NULL for '*').  As in
 * Alignment::PrintSAM (Alignment.h:719-732) the string is printed as given (not reversed with the strand) and hard-clipped like SEQ. */
int64_t lra_b200_format_records_qual(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                     const char *quals_ascii, const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, const uint64_t *contig_len,
                                     int32_t n_contigs, int32_t fmt, int32_t runtime, char *out, int64_t cap);

/* The printers that read the reference bases: fmt 'a' the pairwise view of `-p a` (Alignment::PrintPairwise, Alignment.h:564-589) and, with print_md != 0 and fmt 's', the
 * MD:Z: tag of `--printMD` (AlignmentStringsToMD, Alignment.h:204-245, :763-767).  genome_ascii: contigs back to back as given to lra_b200_mapper_create, contig c at
 * contig_off[c].  Other arguments and the return value as lra_b200_format_records_qual (which this call equals for the other formats). */
int64_t lra_b200_format_records_ref(const lra_b200_map_opts *opts, const lra_b200_map_result *res, int32_t n_reads, const char *names, const char *reads_ascii,
                                    const char *quals_ascii, const uint64_t *read_off, const uint32_t *read_len, const char *contig_names, const uint64_t *contig_len,
                                    int32_t n_contigs, const char *genome_ascii, const uint64_t *contig_off, int32_t fmt, int32_t print_md, int32_t runtime, char *out,
                                    int64_t cap);

/* ---- per-kernel timing of the last batch call (CUDA events on the context's stream) ---------------------------- */
typedef struct lra_b200_kernel_stat {
  char name[48];
  float ms;             /* device time of the launch */
  uint64_t jobs;        /* units processed */
  uint64_t cells;       /* DP cells (0 if not a DP kernel) */
  uint64_t algo_bytes;  /* algorithmic HBM bytes */
} lra_b200_kernel_stat;
/* Returns the number of records available; fills up to cap. */
int lra_b200_last_kernel_stats(lra_b200_ctx *ctx, lra_b200_kernel_stat *out, int cap);
/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t lra_b200_launch_count(const lra_b200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
