// Data model of the mapper worker (a23): device forms of Options, the index views, clusters, chains and the per-segment record.
#pragma once
#include "../../include/lra_b200.h"
#include "mp_common.cuh"
#include "mp_sdp.cuh"
#include "lref_kernels.cuh"   // LidxView, lref_lookup, lref_hdr_find

namespace lra {
namespace mp {

// The subset of Options (Options.h:8-241) the low-accuracy pipeline reads, after the align preset (lra.cpp:268-431): the public POD.  The derived
// option sets of Map_lowacc.h:228-242 are folded in: smallOpts.globalK/W = the LocalIndex k/w (smallK/smallW), tinyOpts.globalW = localW.
typedef lra_b200_map_opts MpOpts;

struct MpIndex {
  SeqView genome;                         // contigs concatenated in Header::pos order
  const unsigned long long *hdr_pos;      // [n_hdr + 1] cumulative contig offsets (Genome.h:59-84)
  int n_hdr;
  const unsigned long long *idx_t;        // global minimizer index (<ref>.mms), sorted by masked tuple
  const uint32_t *idx_pos;
  long long n_idx;
  LidxView gl;                            // <ref>.gli image
};

struct MpReads {
  SeqView fwd, rc;                        // packed arenas, same per-read offsets
  const unsigned long long *read_off;
  const uint32_t *read_len;
  int n_reads;
  LidxView rd[2];                         // LocalIndex::IndexSeq of every read, forward / reverse complement (Map_lowacc.h:246-250)
  const int *lidx_slot;                   // nullptr: rd[] hold every read of the batch (sequence r = read r); else the sequence of read r in rd[], -1 = not indexed
                                          // (high-accuracy presets index only the reads that reach REFINEclusters, in a second pass)
};

// anchors of a set of clusters, SoA; cluster c owns [off[c], off[c+1])
struct ClusterSet {
  uint32_t *q, *t; int *len;              // len == nullptr: raw K-mers
  int *off;                               // [ncl + 1]
  uint32_t *qS, *qE, *tS, *tE;            // cluster boundaries
  int *strand, *chrom;                    // strand: -1 for a default-constructed (empty) cluster
  float *freq;                            // anchorfreq
  int ncl, cap_cl, cap_a;
};

// UltimateChain (Chain.h:172-258): anchors as (cluster, index in cluster) over a ClusterSet
struct UChain {
  uint32_t *idx; int *cl; uint8_t *link;  // idx: index inside the cluster; link has n - 1 (or more: RemoveSpuriousAnchors leaves it) entries
  int n, nlink;
  float FirstSDPValue;
  int NumOfAnchors0, NumOfAnchors1;
  uint32_t QStart, QEnd, TStart, TEnd;
};

// per-segment record: what IndelRefineAlignment / CalculateStatistics / the printers consume (Alignment.h)
struct SegRec {
  int read, chain, order_in_chain;        // chain p of the read, index s in alignments[p].SegAlignment
  int strand, chrom;
  int NumOfAnchors0, NumOfAnchors1;
  int Supplymentary, ISsecondary;
  float FirstSDPValue;
  unsigned long long blk_off; int blk_cnt; // blocks of the segment in the block arena (triples)
};

// status of a read after the map kernel
enum { MP_OK = 0, MP_UNALIGNED = 1, MP_ERR_ARENA = 2, MP_ERR_CAP = 3, MP_ERR_UNSUPPORTED = 4, MP_NEED_LIDX = 5 };

}  // namespace mp
}  // namespace lra
