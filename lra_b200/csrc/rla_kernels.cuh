// a17 (leaf): RefineByLinearAlignment (reference LocalRefineAlignment.h:144-185) = SetMatchAndGaps / Matched (:89-99), RefineSubstrings (:131-142)
// and AlignSubstrings (:101-129) around a18 AffineOneGapAlign, batched over the gaps between consecutive anchors.
// The job of a gap is pure arithmetic on its four coordinates: m = min(qe - qs + 1, te - ts + 1) in GenomePos arithmetic read as int; if m > 0 the
// read window [qs, qe) is aligned to the contig window [ts, te) with band min(2 |qLen - tLen| + 1, opts.localBand), and the blocks are shifted
// back by (qs, ts).  rla_jobs_kernel writes the a18 job arrays, the a18 kernels run them, rla_shift_kernel shifts the blocks in place.
#pragma once
#include "lra_common.cuh"

namespace lra {

struct RlaBatch {
  int n_gaps;
  int local_band;
  const uint32_t *cur_read_end, *next_read_start, *cur_genome_end, *next_genome_start;   // relative to the read / the contig
  const uint32_t *read_off, *chrom_off;                                                    // arena positions of the read (on its strand) and the contig
  uint32_t *q_off, *t_off;                                                                 // a18 job arrays (out)
  int32_t *q_len, *t_len, *k;
  const int32_t *n_blocks;                                                                 // a18 results (shift pass)
  const unsigned long long *block_off;
  uint32_t *blocks;
};

__global__ void __launch_bounds__(256) rla_jobs_kernel(RlaBatch b) {
  const int g = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (g >= b.n_gaps) return;
  const uint32_t qs = b.cur_read_end[g], qe = b.next_read_start[g], ts = b.cur_genome_end[g], te = b.next_genome_start[g];
  const uint32_t a = qe - qs + 1u, c = te - ts + 1u;
  const int m = (int)(a < c ? a : c);                       // Matched(): min of two GenomePos, returned as int
  int ql = 0, tl = 0, k = 1;
  if (m > 0) {
    ql = (int)(qe - qs); tl = (int)(te - ts);
    int drift = ql - tl; if (drift < 0) drift = -drift;
    k = drift * 2 + 1 < b.local_band ? drift * 2 + 1 : b.local_band;
  }
  b.q_off[g] = b.read_off[g] + qs; b.t_off[g] = b.chrom_off[g] + ts;
  b.q_len[g] = ql; b.t_len[g] = tl; b.k[g] = k;
}

__global__ void __launch_bounds__(256) rla_shift_kernel(RlaBatch b) {
  const int g = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (g >= b.n_gaps) return;
  const uint32_t qs = b.cur_read_end[g], ts = b.cur_genome_end[g];
  uint32_t *bl = b.blocks + 3ull * b.block_off[g];
  for (int i = 0; i < b.n_blocks[g]; i++) { bl[3 * i] += qs; bl[3 * i + 1] += ts; }
}

}  // namespace lra
