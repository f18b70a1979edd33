// f2 (global half): the global minimizer index of a genome (`lra index` / `lra global`: StoreIndex, MMIndex.h:286-399) built on the device.
//
//   StoreMinimizers (MinCount.h:7-179) is a sequential scan whose state is (active minimizer, the last w tuples, the N-free window tracker).
//   Here every thread scans one chunk of loop steps of one contig, starting `warm` steps earlier with a fresh state; the fresh state becomes
//   identical to the sequential one at the first step at which the window minimum is unique (rescan with a single minimum, or a strictly
//   smaller new tuple), and stays identical from there on.  A chunk whose warm-up never reaches such a step (low-complexity sequence) is
//   re-scanned with a longer warm-up and finally from the contig start, so the emitted list is exactly the sequential one.
//   Sorting (std::sort on the masked tuple, MMIndex.h:316) is an LSD radix sort; the frequency filter (:328-352), the per-window thinning
//   (CountSort + winCount, :361-378) and RemoveFrequent (:88-98) are flat passes.  The one difference from the reference's file: the order of
//   entries INSIDE a run of equal tuples is by position here, where std::sort leaves introsort's (CompareLists emits the same pairs either way).
#pragma once
#include "mm_range.cuh"

namespace lra {

struct GidxScan {
  SeqView genome;
  const unsigned long long *contig_start;   // [n_contigs] offset of every contig in the genome arena
  const uint32_t *contig_len;               // [n_contigs]
  const unsigned long long *chunk_first;    // [n_contigs + 1] first chunk of every contig
  int n_contigs, k, w, chunk;
  unsigned long long n_chunks;
  const unsigned long long *todo;           // optional list of chunks to (re)scan; n_todo entries
  unsigned long long n_todo;
  uint32_t *warm;                           // [n_chunks] warm-up steps of every chunk (0xffffffff: from the contig start)
  uint32_t *cnt;                            // [n_chunks] minimizers pushed at the loop steps of the chunk
  uint8_t *uncertain;                       // [n_chunks]
  const unsigned long long *off;            // emit pass: exclusive scan of cnt
  unsigned long long *ot; uint32_t *op;     // emit pass: tuple (strand in bit 63) and GLOBAL position (contig offset added, MMIndex.h:305-307)
};

template <bool EMIT>
__global__ void __launch_bounds__(128) gidx_scan_kernel(GidxScan b) {
  const unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= (b.todo ? b.n_todo : b.n_chunks)) return;
  const unsigned long long ch = b.todo ? b.todo[x] : x;
  int lo = 0, hi = b.n_contigs;                        // contig of the chunk: last c with chunk_first[c] <= ch
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (b.chunk_first[mid] <= ch) lo = mid; else hi = mid; }
  const int c = lo;
  const uint32_t L = b.contig_len[c];
  const uint32_t p_begin = (uint32_t)(ch - b.chunk_first[c]) * (uint32_t)b.chunk;
  uint32_t p_end = p_begin + (uint32_t)b.chunk;
  if (ch + 1 == b.chunk_first[c + 1]) p_end = 0xffffffffu;            // the last chunk of a contig runs to its end
  const uint32_t warm = b.warm[ch];
  uint32_t s0 = (warm == 0xffffffffu || p_begin <= warm) ? 0u : p_begin - warm;
  bool certain;
  unsigned long long *ot = EMIT ? b.ot + b.off[ch] : nullptr;
  uint32_t *op = EMIT ? b.op + b.off[ch] : nullptr;
  const uint32_t n = gidx_scan_range<EMIT>(b.genome, b.contig_start[c], L, b.k, b.w, s0, p_begin, p_end, ot, op, (uint32_t)b.contig_start[c], certain);
  if (!EMIT) {
    if (n == 0xffffffffu) { b.uncertain[ch] = 1; b.cnt[ch] = 0; }
    else { b.uncertain[ch] = 0; b.cnt[ch] = n; }
  }
}

// ---- after the radix sort (keys = masked tuples, vals = emission index) ---------------------------------------------------------------------
__global__ void gidx_run_start_kernel(const unsigned long long *key, unsigned long long n, uint32_t *run_start_flagged) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  run_start_flagged[i] = (i == 0 || key[i] != key[i - 1]) ? (uint32_t)i : 0u;      // inclusive max-scan -> first index of the run of i
}
__global__ void gidx_run_count_kernel(const uint32_t *run_start, unsigned long long n, uint32_t *run_cnt) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(&run_cnt[run_start[i]], 1u);
}
// emission order <- sorted order: multiplicity of the tuple and the sorted index
__global__ void gidx_scatter_freq_kernel(const uint32_t *run_start, const uint32_t *run_cnt, const uint32_t *val, unsigned long long n, uint32_t *freq_e, uint32_t *sidx_e) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = val[i];
  freq_e[e] = run_cnt[run_start[i]]; sidx_e[e] = (uint32_t)i;
}
// keep flag of every minimizer (emission order): multiplicity <= maxFreq (MMIndex.h:336), and among the unremoved minimizers of its
// globalWinsize window one of the first `per_window` in CountSort order -- multiplicity ascending, sorted index descending (:263-283, 366-378)
__global__ void gidx_thin_kernel(const uint32_t *pos_e, const uint32_t *freq_e, const uint32_t *sidx_e, unsigned long long n, uint32_t max_freq, uint32_t win, uint32_t per_window,
                                 uint8_t *keep_e) {
  const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t f = freq_e[e];
  if (f > max_freq) { keep_e[e] = 0; return; }
  const uint32_t id = pos_e[e] / win, si = sidx_e[e];
  uint32_t rank = 0;
  for (long long x = (long long)e - 1; x >= 0 && pos_e[x] / win == id; x--) {
    const uint32_t fx = freq_e[x];
    if (fx <= max_freq && (fx < f || (fx == f && sidx_e[x] > si))) rank++;
  }
  for (unsigned long long x = e + 1; x < n && pos_e[x] / win == id; x++) {
    const uint32_t fx = freq_e[x];
    if (fx <= max_freq && (fx < f || (fx == f && sidx_e[x] > si))) rank++;
  }
  keep_e[e] = rank < per_window ? 1 : 0;
}
__global__ void gidx_gather_kernel(const uint32_t *val, const unsigned long long *t_e, const uint32_t *pos_e, const uint8_t *keep_e, unsigned long long n, unsigned long long *t_s,
                                   uint32_t *pos_s, uint8_t *keep_s) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = val[i];
  t_s[i] = t_e[e]; pos_s[i] = pos_e[e]; keep_s[i] = keep_e[e];
}
__global__ void gidx_iota_mask_kernel(const unsigned long long *t_e, unsigned long long n, unsigned long long *key, uint32_t *val) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  key[i] = t_e[i] & kForMask; val[i] = (uint32_t)i;
}
__global__ void gidx_collect_uncertain_kernel(const uint8_t *uncertain, unsigned long long n, unsigned long long *list, unsigned long long *count) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !uncertain[i]) return;
  list[atomicAdd(count, 1ull)] = i;
}
__global__ void gidx_set_warm_kernel(const unsigned long long *list, unsigned long long n, uint32_t *warm, uint32_t v) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) warm[list[i]] = v;
}
__global__ void gidx_fill_u32_kernel(uint32_t *a, unsigned long long n, uint32_t v) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

}  // namespace lra
