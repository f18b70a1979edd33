// TEST INFRASTRUCTURE ONLY: runs lra_b200/csrc/ir_kernels.cuh + ir_segment_kernels.cuh on the CPU through the SIMT emulator.
#include "emu_common.h"
#include "ir_kernels.cuh"
#include "ir_segment_kernels.cuh"

using namespace lra;

typedef emuh::Packed PackedIr;
static void pack_ir(const uint8_t *ascii, uint64_t n, PackedIr &p) { emuh::pack(ascii, n, p); }

// the DP launch sequence (mirrors ir_run_device in lra_b200.cu); returns cells
static uint64_t run_ir_dp(IrBatch &b, int force_generic, std::vector<uint32_t> &tb_store) {
  const int n_groups = b.n_groups;
  unsigned long long tb_cursor = 0, cells = 0;
  std::vector<AogPlan> planv(1); AogPlan *plan = planv.data(); memset(plan, 0, sizeof(AogPlan));
  std::vector<uint32_t> bin(n_groups + 1), sorted(n_groups + 1);
  emu::launch(dim3((unsigned)((n_groups + 3) / 4)), dim3(128), 0, [&] { ir_classify_kernel(b, plan, bin.data(), &tb_cursor, &cells, force_generic == 3 ? 1 : force_generic == 4 ? 2 : 0, 1 << 30); });
  if (force_generic == 1) {
    memset(plan->hist, 0, sizeof plan->hist);
    tb_cursor = 0;
    for (int g = 0; g < n_groups; g++) if (bin[g] != 0xFFFFFFFFu) {
      bin[g] = kIrClsGeneric * kAogBuckets + (bin[g] % kAogBuckets);
      plan->hist[bin[g]]++;
      b.tb_off[g] = tb_cursor;
      tb_cursor += ((unsigned long long)b.t_len[g] * b.max_width[g] + 3) / 4 + 2ull * b.max_width[g] + 4;
    }
  }
  emu::launch(dim3(1), dim3(512), 0, [&] { aog_scan_kernel(plan); });
  unsigned nb = (unsigned)((n_groups + 127) / 128);
  emu::launch(dim3(nb), dim3(128), 0, [&] { aog_scatter_kernel(n_groups, plan, bin.data(), sorted.data()); });
  tb_store.assign(tb_cursor + 16, 0);
  b.tb = tb_store.data();
  auto cnt = [&](int c) { return plan->bin_start[(c + 1) * kAogBuckets] - plan->bin_start[c * kAogBuckets]; };
  if (cnt(kIrClsW24)) emu::launch(dim3(2), dim3(64), 0, [&] { ir_dp_thread_kernel<24>(b, plan, sorted.data(), kIrClsW24); });
  if (cnt(kIrClsW64)) emu::launch(dim3(2), dim3(64), 0, [&] { ir_dp_thread_kernel<64>(b, plan, sorted.data(), kIrClsW64); });
  if (cnt(kIrClsGeneric)) emu::launch(dim3(2), dim3(64), 0, [&] { ir_dp_generic_kernel(b, plan, sorted.data()); });
  if (cnt(kIrClsWarp32)) emu::launch(dim3(2), dim3(128), 0, [&] { ir_dp_warp_kernel(b, plan, sorted.data()); });
  if (cnt(kIrClsWarp64)) emu::launch(dim3(2), dim3(128), 0, [&] { ir_dp_warp64_kernel(b, plan, sorted.data()); });
  if (cnt(kIrClsPipe)) emu::launch(dim3(2), dim3(128), 0, [&] { if (force_generic == 5) ir_dp_pipe_kernel<2>(b, plan, sorted.data(), kIrClsPipe); else ir_dp_pipe_kernel<1>(b, plan, sorted.data(), kIrClsPipe); });
  return cells;
}

// force_generic: 1 = send everything to the generic kernel, 3 = thread kernels only, 4 = long groups through the scan-based warp kernel, 5 = pipeline kernel with two cells per step
// (default: the 8-lane row-pipeline kernel)
extern "C" int emu_ir_dp_batch(const uint8_t *q_arena, uint64_t qn, const uint8_t *t_arena, uint64_t tn, const uint32_t *q_base,
                               const uint32_t *t_base, const int32_t *q_start, const int32_t *t_start, const int32_t *t_len,
                               const int32_t *q_seq_len, const int32_t *t_seq_len, const uint32_t *band_off, const int32_t *band,
                               int n_groups, int match, int mismatch, int indel, int32_t *n_blocks, uint64_t *block_off,
                               uint32_t *blocks, uint64_t block_cap, int force_generic, uint64_t *cells_out) {
  PackedIr q, t; pack_ir(q_arena, qn, q); pack_ir(t_arena, tn, t);
  unsigned long long cursor = 0; int err = 0;
  std::vector<unsigned long long> tb_off(n_groups + 1);
  std::vector<int32_t> maxw(n_groups + 1);
  IrBatch b{q.view, t.view, q_base, t_base, q_start, t_start, t_len, q_seq_len, t_seq_len, band_off, band, n_groups, match, mismatch,
            indel, n_blocks, (unsigned long long *)block_off, blocks, block_cap, &cursor, &err, nullptr, tb_off.data(), maxw.data()};
  std::vector<uint32_t> tbs;
  uint64_t cells = run_ir_dp(b, force_generic, tbs);
  if (cells_out) *cells_out = cells;
  return err;
}


// Whole IndelRefineAlignment over segments (mirrors ir_segments_run_device in lra_b200.cu)
extern "C" int emu_ir_segments(const uint8_t *q_arena, uint64_t qn, const uint8_t *t_arena, uint64_t tn, const uint32_t *blocks_in,
                               const uint64_t *blk_off, const int32_t *blk_cnt, const uint32_t *q_base, const uint32_t *t_base,
                               const int32_t *read_len, const int32_t *contig_len, uint64_t T, int S, int k, int match, int mismatch,
                               int indel, int end_align, int32_t *out_n, uint64_t *out_off, uint32_t *out_blocks, uint64_t out_cap,
                               uint64_t *info /* n_aog, n_groups, cells */) {
  PackedIr q, t; pack_ir(q_arena, qn, q); pack_ir(t_arena, tn, t);
  const size_t nd = T + 2 * (size_t)S + 8;
  std::vector<uint32_t> work((T + 2 * S + 4) * 3), pieces((2 * T + 8 * S + 8) * 4), aq(nd), at(nd), gqb(nd), gtb(nd), gbo(nd), gfirst(nd * 3), glast(nd * 3);
  std::vector<int32_t> npieces(S + 1), aql(nd), atl(nd), ak(nd), gqs(nd), gts(nd), gtl(nd), gqsl(nd), gtsl(nd), gseg(nd), gfb(nd), glb(nd);
  unsigned long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  IrSegBatch b;
  b.blocks_in = blocks_in; b.blk_off = (const unsigned long long *)blk_off; b.blk_cnt = blk_cnt; b.q_base = q_base; b.t_base = t_base;
  b.read_len = read_len; b.contig_len = contig_len; b.n_seg = S; b.k = k; b.end_align = end_align;
  b.work = work.data(); b.pieces = pieces.data(); b.n_pieces = npieces.data();
  b.aog_q_off = aq.data(); b.aog_t_off = at.data(); b.aog_q_len = aql.data(); b.aog_t_len = atl.data(); b.aog_k = ak.data();
  b.g_q_base = gqb.data(); b.g_t_base = gtb.data(); b.g_q_start = gqs.data(); b.g_t_start = gts.data(); b.g_t_len = gtl.data();
  b.g_q_seq_len = gqsl.data(); b.g_t_seq_len = gtsl.data(); b.g_band_off = gbo.data(); b.g_seg = gseg.data();
  b.g_first_block = gfb.data(); b.g_last_block = glb.data(); b.g_first = gfirst.data(); b.g_last = glast.data(); b.counters = counters;
  emu::launch(dim3((unsigned)((S + 3) / 4)), dim3(128), 0, [&] { ir_group_kernel(b); });
  const int nA = (int)counters[0], nG = (int)counters[1];
  std::vector<int32_t> band(counters[2] + 16), band_lit(counters[2] + 16), tmpb(counters[2] + 16);
  if (nG) {
    emu::launch(dim3((unsigned)((nG + 3) / 4)), dim3(128), 0, [&] { ir_band_kernel(b, nG, band.data(), tmpb.data()); });
    // cross-check the closed form against the step-by-step kernel on every group
    emu::launch(dim3((unsigned)((nG + 3) / 4)), dim3(128), 0, [&] { ir_band_literal_kernel(b, nG, band_lit.data()); });
    for (size_t i = 0; i < (size_t)counters[2]; i++) if (band[i] != band_lit[i]) return 1 << 20;
  }
  int err = 0;
  // AffineOneGapAlign fallback
  std::vector<int32_t> a_score(nA + 1), a_nb(nA + 1); std::vector<unsigned long long> a_off(nA + 1);
  std::vector<uint32_t> a_blk(((size_t)nA * k + 16) * 3);
  unsigned long long a_cursor = 0;
  if (nA) {
    AogBatch ab{q.view, t.view, aq.data(), at.data(), aql.data(), atl.data(), ak.data(), nA, match, mismatch, indel, a_score.data(), a_nb.data(),
                a_off.data(), a_blk.data(), (unsigned long long)nA * k + 16, &a_cursor, &err};
    emuh::run_aog(ab, 1);
  }
  // banded DP
  std::vector<int32_t> d_nb(nG + 1), maxw(nG + 1); std::vector<unsigned long long> d_off(nG + 1), tb_off(nG + 1);
  std::vector<uint32_t> d_blk((2 * T + 64 * (size_t)nG + 1024 + (size_t)counters[2]) * 3);
  unsigned long long d_cursor = 0, cells = 0;
  std::vector<uint32_t> tbs;
  if (nG) {
    IrBatch ib{q.view, t.view, gqb.data(), gtb.data(), gqs.data(), gts.data(), gtl.data(), gqsl.data(), gtsl.data(), gbo.data(), band.data(), nG,
               match, mismatch, indel, d_nb.data(), d_off.data(), d_blk.data(), d_blk.size() / 3, &d_cursor, &err, nullptr, tb_off.data(), maxw.data()};
    cells = run_ir_dp(ib, 0, tbs);
  }
  unsigned long long out_cursor = 0;
  IrAssemble a{a_nb.data(), a_off.data(), a_blk.data(), d_nb.data(), d_off.data(), d_blk.data(), out_n, (unsigned long long *)out_off, out_blocks,
               out_cap, &out_cursor, &err};
  emu::launch(dim3((unsigned)((S + 3) / 4)), dim3(128), 0, [&] { ir_assemble_kernel(b, a); });
  if (info) { info[0] = nA; info[1] = nG; info[2] = cells; }
  return err;
}
