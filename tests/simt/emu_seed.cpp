// TEST INFRASTRUCTURE ONLY: runs lra_b200/csrc/seed_kernels.cuh on the CPU through the SIMT emulator.
#include "emu_common.h"
#include "seed_kernels.cuh"

using namespace lra;
using namespace emuh;

extern "C" int emu_seq_revcomp(const uint8_t *ascii, uint64_t n, const uint64_t *read_off, const uint32_t *read_len, int n_reads,
                               uint32_t *b2_out, uint32_t *nm_out) {
  Packed p; pack(ascii, n, p);
  std::vector<uint32_t> b2(p.b2.size(), 0), nm(p.nm.size(), 0xFFFFFFFFu);
  uint32_t *pb = b2.data(), *pn = nm.data();
  emu::launch(dim3((unsigned)((n_reads + 7) / 8)), dim3(256), 0, [&] { seq_revcomp_kernel(p.view, (const unsigned long long *)read_off, read_len, n_reads, pb, pn); });
  memcpy(b2_out, b2.data(), ((n + 15) / 16) * 4);
  memcpy(nm_out, nm.data(), ((n + 31) / 32) * 4);
  return 0;
}

// returns the total number of matches; fills up to cap
extern "C" long emu_seed_batch(const uint8_t *reads_ascii, uint64_t rn, const uint64_t *read_off, const uint32_t *read_len, int n_reads,
                               const uint8_t *genome_ascii, uint64_t gn, const uint64_t *idx_t, const uint32_t *idx_pos, long n_idx, int k, int w,
                               long max_freq, uint64_t *match_off, uint64_t *q_t, uint32_t *q_pos, uint64_t *t_t, uint32_t *t_pos, uint8_t *strand,
                               uint64_t cap, uint32_t *n_mm, uint64_t *mm_t_out, uint32_t *mm_pos_out) {
  Packed reads, genome; pack(reads_ascii, rn, reads); pack(genome_ascii, gn, genome);
  std::vector<unsigned long long> mm_t(rn + 64), cnt(n_reads + 2);
  std::vector<uint32_t> mm_pos(rn + 64), mm_n(n_reads + 1);
  int err = 0;
  SeedBatch b;
  b.reads = reads.view; b.read_off = (const unsigned long long *)read_off; b.read_len = read_len; b.n_reads = n_reads; b.k = k; b.w = w;
  b.max_freq = max_freq; b.idx_t = (const unsigned long long *)idx_t; b.idx_pos = idx_pos; b.n_idx = n_idx; b.genome = genome.view;
  b.mm_t = mm_t.data(); b.mm_pos = mm_pos.data(); b.mm_n = mm_n.data(); b.match_cnt = cnt.data();
  b.m_qt = (unsigned long long *)q_t; b.m_tt = (unsigned long long *)t_t; b.m_qpos = q_pos; b.m_tpos = t_pos; b.m_strand = strand;
  b.match_cap = cap; b.err = &err;
  unsigned nb = (unsigned)((n_reads + 127) / 128);
  emu::launch(dim3(nb), dim3(128), 0, [&] { seed_minimizers_kernel(b); });
  emu::launch(dim3(nb), dim3(128), 0, [&] { seed_sort_kernel(b); });
  emu::launch(dim3(nb), dim3(128), 0, [&] { seed_compare_kernel<false>(b); });
  emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.match_cnt, n_reads, cap, b.err); });
  emu::launch(dim3(nb), dim3(128), 0, [&] { seed_compare_kernel<true>(b); });
  for (int r = 0; r <= n_reads; r++) match_off[r] = cnt[r];
  for (int r = 0; r < n_reads; r++) n_mm[r] = mm_n[r];
  if (mm_t_out) { memcpy(mm_t_out, mm_t.data(), rn * 8); memcpy(mm_pos_out, mm_pos.data(), rn * 4); }
  return (long)cnt[n_reads];
}

// ---- a21 statistics
#include "stats_kernels.cuh"
extern "C" long emu_calc_stats(const uint8_t *q_arena, uint64_t qn, const uint8_t *t_arena, uint64_t tn, const uint32_t *blocks, const uint64_t *blk_off,
                               const int32_t *blk_cnt, const uint32_t *q_base, const uint32_t *t_base, const int32_t *read_len, int S, const float *lut,
                               int32_t *stats, float *value, uint64_t *cig_off, uint32_t *cigar, uint64_t cap, int thread_kernels) {
  Packed q, t; pack(q_arena, qn, q); pack(t_arena, tn, t);
  int err = 0;
  size_t T = 0;
  for (int s = 0; s < S; s++) T = std::max(T, (size_t)(blk_off[s] + (uint64_t)blk_cnt[s]));
  std::vector<uint32_t> pre(3 * T + 16), lane_info((size_t)S * 64 + 16);
  StatsBatch b{q.view, t.view, blocks, (const unsigned long long *)blk_off, blk_cnt, q_base, t_base, read_len, S, lut, stats, value,
               (unsigned long long *)cig_off, cigar, cap, pre.data(), lane_info.data()};
  if (thread_kernels) {
    unsigned nb = (unsigned)((S + 127) / 128);
    emu::launch(dim3(nb), dim3(128), 0, [&] { stats_kernel<false>(b); });
    emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.cig_off, S, cap, &err); });
    emu::launch(dim3(nb), dim3(128), 0, [&] { stats_kernel<true>(b); });
  } else {
    unsigned nb = (unsigned)((S + 3) / 4);
    emu::launch(dim3(nb), dim3(128), 0, [&] { stats_warp_kernel<false>(b); });
    emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.cig_off, S, cap, &err); });
    emu::launch(dim3(nb), dim3(128), 0, [&] { stats_warp_kernel<true>(b); });
  }
  return (long)cig_off[S];
}
