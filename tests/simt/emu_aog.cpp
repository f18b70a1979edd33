// TEST INFRASTRUCTURE ONLY: runs the device code of lra_b200/csrc/{seq,aog}_kernels.cuh on the CPU through the
// lock-step SIMT emulator (cuda_emu.h) so that `pytest -m "not gpu"` can compare the kernel logic with the oracle.
#include "emu_common.h"

using namespace lra;
using namespace emuh;

extern "C" int emu_seq_pack(const uint8_t *ascii, uint64_t n, uint32_t *b2_out, uint32_t *nm_out) {
  Packed p; pack(ascii, n, p);
  memcpy(b2_out, p.b2.data(), ((n + 15) / 16) * 4);
  memcpy(nm_out, p.nm.data(), ((n + 31) / 32) * 4);
  return 0;
}

// use_band: 0 = only thread + literal kernels, 1 = route wide one-sided jobs to the band kernel
// force_literal: send everything to the literal kernel (to test it on the narrow shapes too)
extern "C" int emu_aog_batch(const uint8_t *q_arena, uint64_t qn, const uint8_t *t_arena, uint64_t tn, const uint32_t *q_off,
                             const uint32_t *t_off, const int32_t *q_len, const int32_t *t_len, const int32_t *k, int n_jobs,
                             int m, int mm, int indel, int32_t *score, int32_t *n_blocks, uint64_t *block_off,
                             uint32_t *blocks, uint64_t block_cap, int use_band, int force_literal, uint64_t *cells_out) {
  Packed q, t; pack(q_arena, qn, q); pack(t_arena, tn, t);
  unsigned long long cursor = 0; int err = 0;
  AogBatch b{q.view, t.view, q_off, t_off, q_len, t_len, k, n_jobs, m, mm, indel, score, n_blocks,
             (unsigned long long *)block_off, blocks, block_cap, &cursor, &err};
  uint64_t cells = run_aog(b, force_literal ? 2 : use_band);
  if (cells_out) *cells_out = cells;
  return err;
}

// lane scheduling order of the emulator between collectives (0 ascending, 1 descending, >= 2 random with that seed)
extern "C" void emu_set_lane_order(long mode) { emu::g_order_mode() = mode; }

// ---- a17 leaf RefineByLinearAlignment / a14 core RefineSpace: the job and post-processing kernels around the a18 kernels
#include "rla_kernels.cuh"
#include "rsp_kernels.cuh"
extern "C" int emu_refine_linear(const uint8_t *q_arena, uint64_t qn, const uint8_t *t_arena, uint64_t tn, int n, const uint32_t *cre, const uint32_t *nrs,
                                 const uint32_t *cge, const uint32_t *ngs, const uint32_t *read_off, const uint32_t *chrom_off, int m, int mm, int indel,
                                 int local_band, int32_t *score, int32_t *n_blocks, uint64_t *block_off, uint32_t *blocks, uint64_t block_cap) {
  Packed q, t; pack(q_arena, qn, q); pack(t_arena, tn, t);
  std::vector<uint32_t> qo(n + 1), to(n + 1); std::vector<int32_t> ql(n + 1), tl(n + 1), kk(n + 1);
  RlaBatch rb{n, local_band, cre, nrs, cge, ngs, read_off, chrom_off, qo.data(), to.data(), ql.data(), tl.data(), kk.data(), n_blocks,
              (const unsigned long long *)block_off, blocks};
  if (n) emu::launch(dim3((unsigned)((n + 255) / 256)), dim3(256), 0, [&] { rla_jobs_kernel(rb); });
  unsigned long long cursor = 0; int err = 0;
  AogBatch b{q.view, t.view, qo.data(), to.data(), ql.data(), tl.data(), kk.data(), n, m, mm, indel, score, n_blocks, (unsigned long long *)block_off, blocks, block_cap,
             &cursor, &err};
  run_aog(b, 1);
  if (n) emu::launch(dim3((unsigned)((n + 255) / 256)), dim3(256), 0, [&] { rla_shift_kernel(rb); });
  return err;
}

extern "C" int emu_refine_space(const uint8_t *q_arena, uint64_t qn, const uint8_t *t_arena, uint64_t tn, int n, int K, const uint32_t *qs, const uint32_t *qe,
                                const uint32_t *ts, const uint32_t *te, const uint32_t *lrts, const uint32_t *lrlength, const uint32_t *read_off, const uint32_t *read_len,
                                const uint32_t *chrom_off, const uint8_t *flip, int m, int mm, int indel, const uint64_t *pair_off, uint32_t *pq, uint32_t *pt,
                                int32_t *n_pairs, float *identity, uint64_t block_cap) {
  Packed q, t; pack(q_arena, qn, q); pack(t_arena, tn, t);
  std::vector<uint32_t> qo(n + 1), to(n + 1), blocks(3 * block_cap + 3); std::vector<int32_t> ql(n + 1), tl(n + 1), kk(n + 1), score(n + 1), nb(n + 1);
  std::vector<unsigned long long> boff(n + 1);
  std::vector<uint8_t> not_large(n + 1, 0);
  RspBatch rb{n, K, q.view, t.view, qs, qe, ts, te, lrts, lrlength, read_off, read_len, chrom_off, flip, not_large.data(), qo.data(), to.data(), ql.data(), tl.data(), kk.data(),
              nb.data(), boff.data(), blocks.data(), (const unsigned long long *)pair_off, pq, pt, n_pairs, identity};
  if (n) emu::launch(dim3((unsigned)((n + 255) / 256)), dim3(256), 0, [&] { rsp_jobs_kernel(rb); });
  unsigned long long cursor = 0; int err = 0;
  AogBatch b{q.view, t.view, qo.data(), to.data(), ql.data(), tl.data(), kk.data(), n, m, mm, indel, score.data(), nb.data(), boff.data(), blocks.data(), block_cap,
             &cursor, &err};
  run_aog(b, 1);
  if (n) emu::launch(dim3((unsigned)((n + 127) / 128)), dim3(128), 0, [&] { rsp_harvest_kernel(rb); });
  return err;
}

// the minimizer branch of RefineSpace: every space given here is treated as a large one
extern "C" long emu_refine_space_large(const uint8_t *q_arena, uint64_t qn, const uint8_t *t_arena, uint64_t tn, int n, int K, int W, long long max_freq, const uint32_t *qs,
                                       const uint32_t *qe, const uint32_t *ts, const uint32_t *te, const uint32_t *lrts, const uint32_t *lrlength, const uint32_t *read_off,
                                       const uint32_t *read_len, const uint32_t *chrom_off, const uint8_t *flip, const int32_t *diag, uint64_t *pair_off, uint32_t *pq,
                                       uint32_t *pt, uint64_t cap, int32_t *n_pairs, float *identity) {
  Packed q, t; pack(q_arena, qn, q); pack(t_arena, tn, t);
  std::vector<uint32_t> idx(n + 1), mqn(n + 1), mtn(n + 1);
  std::vector<unsigned long long> mqo(n + 1), mto(n + 1), cnt(n + 1);
  size_t MQ = 0, MT = 0;
  for (int g = 0; g < n; g++) { idx[g] = g; mqo[g] = MQ; mto[g] = MT; MQ += qe[g] - qs[g] + 1; MT += te[g] - ts[g] + lrlength[g] + 1; }
  std::vector<unsigned long long> mqt(MQ + 1), mtt(MT + 1); std::vector<uint32_t> mqp(MQ + 1), mtp(MT + 1);
  RsplBatch lb; memset(&lb, 0, sizeof lb);
  lb.n_large = n; lb.K = K; lb.W = W; lb.max_freq = max_freq; lb.reads = q.view; lb.genome = t.view; lb.idx = idx.data(); lb.qs = qs; lb.qe = qe; lb.ts = ts; lb.te = te;
  lb.lrts = lrts; lb.lrlength = lrlength; lb.read_off = read_off; lb.read_len = read_len; lb.chrom_off = chrom_off; lb.flip = flip; lb.diag = diag;
  lb.mq_off = mqo.data(); lb.mt_off = mto.data(); lb.mq_t = mqt.data(); lb.mt_t = mtt.data(); lb.mq_p = mqp.data(); lb.mt_p = mtp.data(); lb.mq_n = mqn.data(); lb.mt_n = mtn.data();
  lb.cnt = cnt.data(); lb.pair_off = (const unsigned long long *)pair_off; lb.pq = pq; lb.pt = pt; lb.n_pairs = n_pairs; lb.identity = identity;
  if (!n) return 0;
  emu::launch(dim3((unsigned)((2 * n + 63) / 64)), dim3(64), 0, [&] { rspl_mins_kernel(lb); });
  emu::launch(dim3((unsigned)((n + 63) / 64)), dim3(64), 0, [&] { rspl_compare_kernel<false>(lb); });
  unsigned long long P = 0;
  for (int g = 0; g < n; g++) { pair_off[g] = P; P += cnt[g]; }
  pair_off[n] = P;
  if (P > cap) return -(long)P;
  emu::launch(dim3((unsigned)((n + 63) / 64)), dim3(64), 0, [&] { rspl_compare_kernel<true>(lb); });
  return (long)P;
}
