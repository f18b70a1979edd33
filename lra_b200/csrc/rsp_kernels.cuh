// a14 (core, small spaces): RefineSpace (reference ClusterRefine.h:242-327) for spaces shorter than 1000 bases on both axes -- the branch that aligns
// the space with AffineOneGapAlign (band 30) and harvests exact K-mers from its blocks (:262-294), the `identity` it returns, and the coordinate
// shift of its tail (:314-325).  rsp_jobs_kernel writes the a18 job arrays, the a18 kernels align, rsp_harvest_kernel walks the blocks of a space
// (one thread per space: <= 1000 bases) and writes the pairs into the space's slot.  The branch for larger spaces (on-the-fly minimizers +
// CompareLists) is not built: the host side rejects such a space.
#pragma once
#include "lra_common.cuh"

namespace lra {

struct RspBatch {
  int n;
  int K;
  SeqView reads, genome;
  const uint32_t *qs, *qe, *ts, *te, *lrts, *lrlength;     // RefineSpace's arguments (qs / qe on the strand the space is refined on)
  const uint32_t *read_off, *read_len, *chrom_off;         // the read on that strand's arena, its length, the contig in the packed genome
  const uint8_t *flip;                                      // consider_str && st == 1: report read positions on the other strand
  uint32_t *q_off, *t_off;                                  // a18 job arrays (out)
  int32_t *q_len, *t_len, *k;
  const int32_t *n_blocks;                                  // a18 results
  const unsigned long long *block_off;
  const uint32_t *blocks;
  const unsigned long long *pair_off;                       // [n] slot of the space's pairs
  uint32_t *pq, *pt;                                        // out
  int32_t *n_pairs;                                         // [n] out
  float *identity;                                          // [n] out
};

__global__ void __launch_bounds__(256) rsp_jobs_kernel(RspBatch b) {
  const int g = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (g >= b.n) return;
  b.q_off[g] = b.read_off[g] + b.qs[g]; b.q_len[g] = (int)(b.qe[g] - b.qs[g]);
  b.t_off[g] = b.chrom_off[g] + (b.ts[g] - b.lrts[g]); b.t_len[g] = (int)(b.te[g] - b.ts[g] + b.lrlength[g]);
  b.k[g] = 30;
}

__global__ void __launch_bounds__(128) rsp_harvest_kernel(RspBatch b) {
  const int g = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (g >= b.n) return;
  const uint32_t K = (uint32_t)b.K;
  const unsigned long long q0 = (unsigned long long)b.q_off[g], t0 = (unsigned long long)b.t_off[g];
  const uint32_t *bl = b.blocks + 3ull * b.block_off[g];
  const unsigned long long po = b.pair_off[g];
  const uint32_t qs = b.qs[g], tshift = b.ts[g] - b.lrts[g];
  int nMatch = 0, np = 0;
  for (int i = 0; i < b.n_blocks[g]; i++) {
    const uint32_t bq = bl[3 * i], bt = bl[3 * i + 1], len = bl[3 * i + 2];
    for (uint32_t x = 0; x < len; x++) nMatch += seq_code(b.reads, q0 + bq + x) == seq_code(b.genome, t0 + bt + x) ? 1 : 0;
    if (len > K) {
      for (uint32_t bp = 0; bp + K < len; bp += K) {
        bool mis = false;
        for (uint32_t x = 0; x < K; x++) if (seq_code(b.genome, t0 + bt + bp + x) != seq_code(b.reads, q0 + bq + bp + x)) { mis = true; break; }
        if (!mis) {
          uint32_t fq = bq + bp + qs;
          if (b.flip[g]) fq = b.read_len[g] - fq - K;
          b.pq[po + np] = fq; b.pt[po + np] = bt + bp + tshift;
          np++;
        }
      }
    }
  }
  b.n_pairs[g] = np;
  const uint32_t ql = (uint32_t)b.q_len[g], tl = (uint32_t)b.t_len[g];
  const uint32_t mn = ql < tl ? ql : tl;
  // nMatch / (float) min(querySeq.size(), refSeq.size()); 0 / 0 is the x86 default NaN (sign bit set)
  b.identity[g] = (mn == 0 && nMatch == 0) ? __uint_as_float(0xFFC00000u) : __fdiv_rn((float)nMatch, (float)mn);
}

}  // namespace lra
