// TEST INFRASTRUCTURE ONLY: runs lra_b200/csrc/ir_kernels.cuh on the CPU through the SIMT emulator.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ir_kernels.cuh"
#include "seq_kernels.cuh"

using namespace lra;

struct PackedIr { std::vector<uint32_t> b2, nm; SeqView view; };
static void pack_ir(const uint8_t *ascii, uint64_t n, PackedIr &p) {
  uint64_t groups = (n + 31) / 32 + 1;
  p.b2.assign(groups * 2 + 8, 0); p.nm.assign(groups + 8, 0);
  std::vector<uint8_t> al(n + 96);
  uint8_t *a = al.data(); while (((uintptr_t)a) & 15) a++;
  memcpy(a, ascii, n);
  uint32_t *b2 = p.b2.data(), *nm = p.nm.data();
  emu::launch(dim3((unsigned)((groups + 63) / 64)), dim3(64), 0, [&] { seq_pack_kernel(a, n, b2, nm, groups); });
  p.view = SeqView{p.b2.data(), p.nm.data(), n};
}

// force_class: -1 natural, 2 = send everything to the generic kernel
extern "C" int emu_ir_dp_batch(const uint8_t *q_arena, uint64_t qn, const uint8_t *t_arena, uint64_t tn, const uint32_t *q_base,
                               const uint32_t *t_base, const int32_t *q_start, const int32_t *t_start, const int32_t *t_len,
                               const int32_t *q_seq_len, const int32_t *t_seq_len, const uint32_t *band_off, const int32_t *band,
                               int n_groups, int match, int mismatch, int indel, int32_t *n_blocks, uint64_t *block_off,
                               uint32_t *blocks, uint64_t block_cap, int force_generic, uint64_t *cells_out) {
  PackedIr q, t; pack_ir(q_arena, qn, q); pack_ir(t_arena, tn, t);
  unsigned long long cursor = 0, tb_cursor = 0, cells = 0; int err = 0;
  std::vector<unsigned long long> tb_off(n_groups + 1);
  std::vector<int32_t> maxw(n_groups + 1);
  IrBatch b{q.view, t.view, q_base, t_base, q_start, t_start, t_len, q_seq_len, t_seq_len, band_off, band, n_groups, match, mismatch,
            indel, n_blocks, (unsigned long long *)block_off, blocks, block_cap, &cursor, &err, nullptr, tb_off.data(), maxw.data()};
  std::vector<AogPlan> planv(1); AogPlan *plan = planv.data(); memset(plan, 0, sizeof(AogPlan));
  std::vector<uint32_t> bin(n_groups + 1), sorted(n_groups + 1);
  emu::launch(dim3((unsigned)((n_groups + 3) / 4)), dim3(128), 0, [&] { ir_classify_kernel(b, plan, bin.data(), &tb_cursor, &cells); });
  if (force_generic) {
    // re-bin everything into the generic class (recompute storage: generic needs more)
    memset(plan->hist, 0, sizeof plan->hist);
    tb_cursor = 0;
    for (int g = 0; g < n_groups; g++) if (bin[g] != 0xFFFFFFFFu) {
      bin[g] = kIrClsGeneric * kAogBuckets + (bin[g] % kAogBuckets);
      plan->hist[bin[g]]++;
      tb_off[g] = tb_cursor;
      tb_cursor += ((unsigned long long)t_len[g] * maxw[g] + 3) / 4 + 2ull * maxw[g] + 4;
    }
  }
  emu::launch(dim3(1), dim3(512), 0, [&] { aog_scan_kernel(plan); });
  unsigned nb = (unsigned)((n_groups + 127) / 128);
  emu::launch(dim3(nb), dim3(128), 0, [&] { aog_scatter_kernel(n_groups, plan, bin.data(), sorted.data()); });
  std::vector<uint32_t> tb(tb_cursor + 16);
  b.tb = tb.data();
  auto cnt = [&](int c) { return plan->bin_start[(c + 1) * kAogBuckets] - plan->bin_start[c * kAogBuckets]; };
  if (cnt(kIrClsW24)) emu::launch(dim3(2), dim3(64), 0, [&] { ir_dp_thread_kernel<24>(b, plan, sorted.data(), kIrClsW24); });
  if (cnt(kIrClsW64)) emu::launch(dim3(2), dim3(64), 0, [&] { ir_dp_thread_kernel<64>(b, plan, sorted.data(), kIrClsW64); });
  if (cnt(kIrClsGeneric)) emu::launch(dim3(2), dim3(64), 0, [&] { ir_dp_generic_kernel(b, plan, sorted.data()); });
  if (cells_out) *cells_out = cells;
  return err;
}
