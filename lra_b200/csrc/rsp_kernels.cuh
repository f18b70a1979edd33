// a14 (core, small spaces): RefineSpace (reference ClusterRefine.h:242-327) for spaces shorter than 1000 bases on both axes -- the branch that aligns
// the space with AffineOneGapAlign (band 30) and harvests exact K-mers from its blocks (:262-294), the `identity` it returns, and the coordinate
// shift of its tail (:314-325).  rsp_jobs_kernel writes the a18 job arrays, the a18 kernels align, rsp_harvest_kernel walks the blocks of a space
// (one thread per space: <= 1000 bases) and writes the pairs into the space's slot.
// Larger spaces take the other branch (:296-305): StoreMinimizers_noncanonical of both windows, std::sort, CompareLists(Global = false) inside the
// diagonal band -- the literal device routines of the seeding stage (seed_kernels.cuh: mm_scan<false>, mm_sort, mm_compare), one thread per window for
// the minimizers and the sort, one thread per space for the comparison (count, then emit once the slots are known).
#pragma once
#include "lra_common.cuh"
#include "seed_kernels.cuh"

namespace lra {

struct RspBatch {
  int n;
  int K;
  SeqView reads, genome;
  const uint32_t *qs, *qe, *ts, *te, *lrts, *lrlength;     // RefineSpace's arguments (qs / qe on the strand the space is refined on)
  const uint32_t *read_off, *read_len, *chrom_off;         // the read on that strand's arena, its length, the contig in the packed genome
  const uint8_t *flip;                                      // consider_str && st == 1: report read positions on the other strand
  const uint8_t *large;                                     // 1 = the space takes the minimizer branch (no alignment job, nothing to harvest)
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
  const bool lg = b.large[g] != 0;
  b.q_off[g] = b.read_off[g] + b.qs[g]; b.q_len[g] = lg ? 0 : (int)(b.qe[g] - b.qs[g]);
  b.t_off[g] = b.chrom_off[g] + (b.ts[g] - b.lrts[g]); b.t_len[g] = lg ? 0 : (int)(b.te[g] - b.ts[g] + b.lrlength[g]);
  b.k[g] = 30;
}

__global__ void __launch_bounds__(128) rsp_harvest_kernel(RspBatch b) {
  const int g = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (g >= b.n || b.large[g]) return;
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

struct RsplBatch {
  int n_large;
  int K, W;
  long long max_freq;                                       // opts.localMaxFreq
  SeqView reads, genome;
  const uint32_t *idx;                                      // [n_large] the space's index in the batch
  const uint32_t *qs, *qe, *ts, *te, *lrts, *lrlength, *read_off, *read_len, *chrom_off;   // batch arrays
  const uint8_t *flip;
  const int32_t *diag;                                      // refineSpaceDiag per space (batch array)
  const unsigned long long *mq_off, *mt_off;                // [n_large] minimizer slots
  unsigned long long *mq_t, *mt_t;
  uint32_t *mq_p, *mt_p;
  uint32_t *mq_n, *mt_n;                                    // [n_large]
  unsigned long long *cnt;                                  // [n_large] pairs found by the count pass
  const unsigned long long *pair_off;                       // batch array
  uint32_t *pq, *pt;
  int32_t *n_pairs;
  float *identity;
};

// one thread per (large space, window): minimizers + std::sort of the read window (even threads) and of the genome window (odd threads)
__global__ void __launch_bounds__(64) rspl_mins_kernel(RsplBatch b) {
  const int id = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  const int j = id >> 1;
  if (j >= b.n_large) return;
  const uint32_t g = b.idx[j];
  if ((id & 1) == 0) {
    unsigned long long *ot = b.mq_t + b.mq_off[j]; uint32_t *op = b.mq_p + b.mq_off[j];
    const uint32_t n = mm_scan<false>(b.reads, (unsigned long long)b.read_off[g] + b.qs[g], b.qe[g] - b.qs[g], b.K, b.W, ot, op);
    mm_sort(MmRef{ot, op}, (long)n);
    b.mq_n[j] = n;
  } else {
    unsigned long long *ot = b.mt_t + b.mt_off[j]; uint32_t *op = b.mt_p + b.mt_off[j];
    const uint32_t n = mm_scan<false>(b.genome, (unsigned long long)b.chrom_off[g] + (b.ts[g] - b.lrts[g]), b.te[g] - b.ts[g] + b.lrlength[g], b.K, b.W, ot, op);
    mm_sort(MmRef{ot, op}, (long)n);
    b.mt_n[j] = n;
  }
}

template <bool EMIT>
__global__ void __launch_bounds__(64) rspl_compare_kernel(RsplBatch b) {
  const int j = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (j >= b.n_large) return;
  const uint32_t g = b.idx[j];
  const unsigned long long *qt = b.mq_t + b.mq_off[j], *tt = b.mt_t + b.mt_off[j];
  const uint32_t *qpos = b.mq_p + b.mq_off[j], *tpos = b.mt_p + b.mt_off[j];
  const long nq = (long)b.mq_n[j], nt = (long)b.mt_n[j];
  // the band of the space in local coordinates (ClusterRefine.h:248-255)
  const long long diag2 = (long long)(uint32_t)(b.te[g] - (b.ts[g] - b.lrts[g])) - (long long)(uint32_t)(b.qe[g] - b.qs[g]);
  const long long minDiagNum = (diag2 < 0 ? diag2 : 0) - (long long)b.diag[g], maxDiagNum = (diag2 > 0 ? diag2 : 0) + (long long)b.diag[g];
  const unsigned long long po = EMIT ? b.pair_off[g] : 0ull;
  const uint32_t qs = b.qs[g], tshift = b.ts[g] - b.lrts[g], K = (uint32_t)b.K;
  const bool flip = b.flip[g] != 0;
  const uint32_t rlen = b.read_len[g];
  unsigned long long n_out = 0;
  if (nq != 0 && nt != 0)
    mm_compare(qt, nq, tt, nt, b.max_freq, [&](long qi, long ti) {
      if (maxDiagNum != 0 && minDiagNum != 0) {
        const long long D = (long long)tpos[ti] - (long long)qpos[qi];
        if (!(D <= maxDiagNum && D >= minDiagNum)) return;
      }
      if (EMIT) {
        uint32_t fq = qpos[qi] + qs;
        if (flip) fq = rlen - fq - K;
        b.pq[po + n_out] = fq; b.pt[po + n_out] = tpos[ti] + tshift;
      }
      n_out++;
    });
  if (!EMIT) b.cnt[j] = n_out;
  else { b.n_pairs[g] = (int32_t)n_out; b.identity[g] = -1.0f; }
}

}  // namespace lra
