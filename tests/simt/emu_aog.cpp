// TEST INFRASTRUCTURE ONLY: runs the device code of lra_b200/csrc/{seq,aog}_kernels.cuh on the CPU through the
// lock-step SIMT emulator (cuda_emu.h) so that `pytest -m "not gpu"` can compare the kernel logic with the oracle.
// Mirrors the launch sequence of lra_b200/csrc/aog.cu.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "aog_kernels.cuh"
#include "aog_band_kernel.cuh"
#include "seq_kernels.cuh"

using namespace lra;

struct Packed { std::vector<uint32_t> b2, nm; SeqView view; };
static void pack(const uint8_t *ascii, uint64_t n, Packed &p) {
  uint64_t groups = (n + 31) / 32 + 1;
  p.b2.assign(groups * 2 + 8, 0); p.nm.assign(groups + 8, 0);
  std::vector<uint8_t> padded(n + 64, 'N');
  memcpy(padded.data(), ascii, n);
  const uint8_t *src = padded.data();
  // the kernel wants 16-byte aligned input for its vector path
  std::vector<uint8_t> al(n + 96);
  uint8_t *a = al.data(); while (((uintptr_t)a) & 15) a++;
  memcpy(a, src, n);
  uint32_t *b2 = p.b2.data(), *nm = p.nm.data();
  emu::launch(dim3((unsigned)((groups + 63) / 64)), dim3(64), 0, [&] { seq_pack_kernel(a, n, b2, nm, groups); });
  p.view = SeqView{p.b2.data(), p.nm.data(), n};
}

extern "C" int emu_seq_pack(const uint8_t *ascii, uint64_t n, uint32_t *b2_out, uint32_t *nm_out) {
  Packed p; pack(ascii, n, p);
  memcpy(b2_out, p.b2.data(), ((n + 15) / 16) * 4);
  memcpy(nm_out, p.nm.data(), ((n + 31) / 32) * 4);
  return 0;
}

template <int K> static void run_thread_class(AogBatch &b, AogPlan *plan, const uint32_t *sorted) {
  uint32_t n = plan->bin_start[(K / 2) * kAogBuckets] - plan->bin_start[(K / 2 - 1) * kAogBuckets];
  if (!n) return;
  unsigned blocks = (n + 127) / 128; if (blocks > 3) blocks = 3;  // persistent warps: fewer CTAs than work
  emu::launch(dim3(blocks), dim3(128), 0, [&] { aog_thread_kernel<K>(b, plan, sorted); });
}
template <int C> static void run_band_class(AogBatch &b, AogPlan *plan, const uint32_t *sorted, int ci, AogBandScratch sc) {
  uint32_t n = plan->bin_start[(kAogClsBand1 + ci + 1) * kAogBuckets] - plan->bin_start[(kAogClsBand1 + ci) * kAogBuckets];
  if (!n) return;
  emu::launch(dim3(2), dim3(128), 0, [&] { aog_warp_band_kernel<C>(b, plan, sorted, sc); });
}

// use_band: 0 = only thread + literal kernels, 1 = route wide one-sided jobs to the band kernel
// force_literal: send everything to the literal kernel (to test it on the narrow shapes too)
extern "C" int emu_aog_batch(const uint8_t *q_arena, uint64_t qn, const uint8_t *t_arena, uint64_t tn, const uint32_t *q_off,
                             const uint32_t *t_off, const int32_t *q_len, const int32_t *t_len, const int32_t *k, int n_jobs,
                             int m, int mm, int indel, int32_t *score, int32_t *n_blocks, uint64_t *block_off,
                             uint32_t *blocks, uint64_t block_cap, int use_band, int force_literal, uint64_t *cells_out) {
  Packed q, t; pack(q_arena, qn, q); pack(t_arena, tn, t);
  unsigned long long cursor = 0; int err = 0;
  std::vector<int32_t> kk(k, k + n_jobs);
  AogBatch b{q.view, t.view, q_off, t_off, q_len, t_len, kk.data(), n_jobs, m, mm, indel, score, n_blocks,
             (unsigned long long *)block_off, blocks, block_cap, &cursor, &err};
  std::vector<AogPlan> planv(1); AogPlan *plan = planv.data(); memset(plan, 0, sizeof(AogPlan));
  std::vector<uint32_t> bin(n_jobs + 1), sorted(n_jobs + 1);
  unsigned nb = (unsigned)((n_jobs + 127) / 128);
  int mode = force_literal ? 2 : use_band;
  emu::launch(dim3(nb), dim3(128), 0, [&] { aog_classify_kernel(b, plan, bin.data(), mode); });
  emu::launch(dim3(1), dim3(512), 0, [&] { aog_scan_kernel(plan); });
  emu::launch(dim3(nb), dim3(128), 0, [&] { aog_scatter_kernel(n_jobs, plan, bin.data(), sorted.data()); });
  run_thread_class<2>(b, plan, sorted.data()); run_thread_class<4>(b, plan, sorted.data());
  run_thread_class<6>(b, plan, sorted.data()); run_thread_class<8>(b, plan, sorted.data());
  run_thread_class<10>(b, plan, sorted.data()); run_thread_class<12>(b, plan, sorted.data());
  run_thread_class<14>(b, plan, sorted.data());
  uint32_t nlit = plan->bin_start[(kAogClsLiteral + 1) * kAogBuckets] - plan->bin_start[kAogClsLiteral * kAogBuckets];
  if (nlit) {
    const int warps = 8;  // 2 CTAs x 4 warps
    AogLiteralScratch sc; sc.max_mat = plan->max_mat; sc.max_diag = plan->max_diag;
    sc.slab_bytes = aog_literal_slab_bytes(sc.max_mat, sc.max_diag);
    std::vector<unsigned char> slab((size_t)sc.slab_bytes * warps + 64);
    sc.base = slab.data(); while (((uintptr_t)sc.base) & 15) sc.base++;
    emu::launch(dim3(2), dim3(128), 0, [&] { aog_warp_literal_kernel(b, plan, sorted.data(), sc); });
  }
  {
    const int warps = 8;
    AogBandScratch sc; sc.max_rows = plan->max_rows_band; sc.max_qlen = plan->max_qlen_band;
    sc.slab_bytes = aog_band_slab_bytes(sc.max_rows, sc.max_qlen);
    std::vector<unsigned char> slab((size_t)sc.slab_bytes * warps + 64);
    sc.base = slab.data(); while (((uintptr_t)sc.base) & 15) sc.base++;
    run_band_class<1>(b, plan, sorted.data(), 0, sc); run_band_class<2>(b, plan, sorted.data(), 1, sc);
    run_band_class<4>(b, plan, sorted.data(), 2, sc); run_band_class<8>(b, plan, sorted.data(), 3, sc);
  }
  if (cells_out) *cells_out = plan->cells;
  return err;
}
