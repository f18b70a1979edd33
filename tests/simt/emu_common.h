// TEST INFRASTRUCTURE ONLY: shared helpers of the emulator harnesses (packing and the AffineOneGapAlign launch sequence,
// mirroring lra_b200/csrc/lra_b200.cu).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "aog_kernels.cuh"
#include "aog_band_kernel.cuh"
#include "seq_kernels.cuh"

namespace emuh {
using namespace lra;

struct Packed { std::vector<uint32_t> b2, nm; SeqView view; };
inline void pack(const uint8_t *ascii, uint64_t n, Packed &p) {
  uint64_t groups = (n + 31) / 32 + 1;
  p.b2.assign(groups * 2 + 8, 0); p.nm.assign(groups + 8, 0);
  std::vector<uint8_t> al(n + 96);
  uint8_t *a = al.data(); while (((uintptr_t)a) & 15) a++;
  memcpy(a, ascii, n);
  uint32_t *b2 = p.b2.data(), *nm = p.nm.data();
  emu::launch(dim3((unsigned)((groups + 63) / 64)), dim3(64), 0, [&] { seq_pack_kernel(a, n, b2, nm, groups); });
  p.view = SeqView{p.b2.data(), p.nm.data(), n};
}

template <int K> inline void run_thread_class(AogBatch &b, AogPlan *plan, const uint32_t *sorted) {
  uint32_t n = plan->bin_start[(K / 2) * kAogBuckets] - plan->bin_start[(K / 2 - 1) * kAogBuckets];
  if (!n) return;
  unsigned blocks = (n + 127) / 128; if (blocks > 3) blocks = 3;  // persistent warps: fewer CTAs than work
  emu::launch(dim3(blocks), dim3(128), 0, [&] { aog_thread_kernel<K>(b, plan, sorted); });
}
template <int C> inline void run_band_class(AogBatch &b, AogPlan *plan, const uint32_t *sorted, int ci, AogBandScratch sc) {
  uint32_t n = plan->bin_start[(kAogClsBand1 + ci + 1) * kAogBuckets] - plan->bin_start[(kAogClsBand1 + ci) * kAogBuckets];
  if (!n) return;
  emu::launch(dim3(2), dim3(128), 0, [&] { aog_warp_band_kernel<C>(b, plan, sorted, sc); });
}

// mode: 0 = thread + literal kernels, 1 = thread + band + literal, 2 = literal only
inline uint64_t run_aog(AogBatch &b, int mode) {
  const int n_jobs = b.n_jobs;
  std::vector<AogPlan> planv(1); AogPlan *plan = planv.data(); memset(plan, 0, sizeof(AogPlan));
  std::vector<uint32_t> bin(n_jobs + 1), sorted(n_jobs + 1);
  unsigned nb = (unsigned)((n_jobs + 127) / 128);
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
  return plan->cells;
}
}  // namespace emuh
