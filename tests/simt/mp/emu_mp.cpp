// TEST INFRASTRUCTURE ONLY: the mapper worker code of lra_b200/csrc/mp_*.cuh compiled for the CPU through the SIMT emulator.
// Built twice by tests/emu_mp.py: -DMP_LANES=32 (the product's lane count, lock-step fibers) and -DMP_LANES=1 (every warp
// collective is the identity; same source at CPU speed for bulk parity runs).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "mp_sdp_driver.cuh"

using namespace lra;
using namespace lra::mp;

extern "C" int emu_mp_lanes() { return MP_LANES; }

extern "C" int emu_sdp_batch_ext(int n_prob, int max_aln, const int *mode, const uint64_t *frag_off, const uint32_t *q, const uint32_t *t, const int32_t *len,
                                 const uint64_t *cl_off_off, const int *cl_off, const uint8_t *cl_strand, const int *only_cl, const float *rate, const int *irate,
                                 const int *read_len, float alnthres, int NumAln, const int64_t *stops, const float *slope, const float *inter, int ceil1, int ceil2,
                                 int *n_chains, int *chain_len, float *chain_val, uint32_t *bounds, uint32_t *chain, uint8_t *link, int *cl_of_frag,
                                 uint64_t arena_bytes, uint64_t *peak,
                                 const uint32_t *qe, const uint32_t *te, const uint8_t *fstrand, const float *fval, const int32_t *fn0, int globalK, int *out_n0) {
  Pwl pwl; for (int i = 0; i < 25; i++) { pwl.stops[i] = stops[i]; pwl.slope[i] = slope[i]; pwl.inter[i] = inter[i]; } pwl.ceil1 = ceil1; pwl.ceil2 = ceil2;
  std::vector<unsigned char> arena(arena_bytes + 64);
  int err = 0;
  SdpBatch b;
  b.n_prob = n_prob; b.max_aln = max_aln; b.mode = mode; b.frag_off = (const unsigned long long *)frag_off; b.q = q; b.t = t; b.len = len;
  b.cl_off_off = (const unsigned long long *)cl_off_off; b.cl_off = cl_off; b.cl_strand = cl_strand; b.only_cl = only_cl; b.rate = rate; b.irate = irate;
  b.read_len = read_len; b.alnthres = alnthres; b.NumAln = NumAln; b.pwl = &pwl;
  b.n_chains = n_chains; b.chain_len = chain_len; b.chain_val = chain_val; b.bounds = bounds; b.chain = chain; b.link = link; b.cl_of_frag = cl_of_frag;
  b.qe = qe; b.te = te; b.fstrand = fstrand; b.fval = fval; b.fn0 = fn0; b.globalK = globalK; b.out_n0 = out_n0;
  unsigned char *base = arena.data(); while (((uintptr_t)base) & 15) base++;
  b.arena = base; b.arena_per_warp = arena_bytes; b.err = &err; b.peak = (unsigned long long *)peak;
  emu::launch(dim3(1), dim3(MP_LANES), 0, [&] { sdp_batch_kernel(b); });
  return err;
}
extern "C" int emu_sdp_batch(int n_prob, int max_aln, const int *mode, const uint64_t *frag_off, const uint32_t *q, const uint32_t *t, const int32_t *len,
                             const uint64_t *cl_off_off, const int *cl_off, const uint8_t *cl_strand, const int *only_cl, const float *rate, const int *irate,
                             const int *read_len, float alnthres, int NumAln, const int64_t *stops, const float *slope, const float *inter, int ceil1, int ceil2,
                             int *n_chains, int *chain_len, float *chain_val, uint32_t *bounds, uint32_t *chain, uint8_t *link, int *cl_of_frag,
                             uint64_t arena_bytes, uint64_t *peak) {
  return emu_sdp_batch_ext(n_prob, max_aln, mode, frag_off, q, t, len, cl_off_off, cl_off, cl_strand, only_cl, rate, irate, read_len, alnthres, NumAln, stops, slope, inter, ceil1,
                           ceil2, n_chains, chain_len, chain_val, bounds, chain, link, cl_of_frag, arena_bytes, peak, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr);
}

// ---- the mapper worker (map_reads_kernel) and map_finalize_kernel under the emulator ------------------------------------------------
#include "mp_map.cuh"
#include "seq_kernels.cuh"

namespace {
struct Packed { std::vector<uint32_t> b2, nm; SeqView view; };
void pack_ascii(const uint8_t *ascii, uint64_t n, Packed &p) {
  // host packing with the alphabet of seq_pack_kernel: A,C,G,T (either case) -> 0..3, everything else -> N
  p.b2.assign((n + 15) / 16 + 8, 0); p.nm.assign((n + 31) / 32 + 8, 0xFFFFFFFFu);
  for (uint64_t i = 0; i < n; i++) {
    int c;
    switch (ascii[i]) { case 'A': case 'a': c = 0; break; case 'C': case 'c': c = 1; break; case 'G': case 'g': c = 2; break; case 'T': case 't': c = 3; break; default: c = 4; }
    if (c < 4) { p.b2[i >> 4] |= (uint32_t)c << (2 * (i & 15)); p.nm[i >> 5] &= ~(1u << (i & 31)); }
  }
  p.view = SeqView{p.b2.data(), p.nm.data(), n};
}
}  // namespace

extern "C" int emu_map_reads(const uint8_t *reads_fwd, const uint8_t *reads_rc, uint64_t rn, const uint64_t *read_off, const uint32_t *read_len, int n_reads,
                             const uint8_t *genome, uint64_t gn, const uint64_t *hdr_pos, int n_hdr, const uint64_t *idx_t, const uint32_t *idx_pos, int64_t n_idx,
                             const uint64_t *gl_win_off, const uint64_t *gl_bnd, const uint32_t *gl_mins, int gl_n_win,
                             const uint32_t *rf_win_first, const uint64_t *rf_win_off, const uint64_t *rf_bnd, const uint32_t *rf_mins,
                             const uint32_t *rr_win_first, const uint64_t *rr_win_off, const uint64_t *rr_bnd, const uint32_t *rr_mins,
                             const MpOpts *opts, const int64_t *stops, const float *slope, const float *inter, int ceil1, int ceil2,
                             int *status, int *n_chains, int *chain_nseg, int *chain_seg0, SegRec *seg, int seg_cap, uint32_t *blocks, uint64_t blk_cap,
                             uint64_t *counts /* n_seg, n_blk, err, peak */, uint64_t arena_bytes) {
  Packed pf, pr, pg; pack_ascii(reads_fwd, rn, pf); pack_ascii(reads_rc, rn, pr); pack_ascii(genome, gn, pg);
  Pwl pwl; for (int i = 0; i < 25; i++) { pwl.stops[i] = stops[i]; pwl.slope[i] = slope[i]; pwl.inter[i] = inter[i]; } pwl.ceil1 = ceil1; pwl.ceil2 = ceil2;
  std::vector<unsigned char> arena(arena_bytes + 64);
  unsigned char *base = arena.data(); while (((uintptr_t)base) & 15) base++;
  MapBatch b;
  b.C.o = *opts; b.C.pwl = &pwl; b.C.prof = nullptr;
  b.C.ix.genome = pg.view; b.C.ix.hdr_pos = (const unsigned long long *)hdr_pos; b.C.ix.n_hdr = n_hdr; b.C.ix.idx_t = (const unsigned long long *)idx_t; b.C.ix.idx_pos = idx_pos;
  b.C.ix.n_idx = n_idx;
  b.C.ix.gl = LidxView{(const unsigned long long *)gl_win_off, nullptr, (const unsigned long long *)gl_bnd, gl_mins, nullptr, nullptr, nullptr, gl_n_win, n_hdr};
  b.C.rd.fwd = pf.view; b.C.rd.rc = pr.view; b.C.rd.read_off = (const unsigned long long *)read_off; b.C.rd.read_len = read_len; b.C.rd.n_reads = n_reads; b.C.rd.lidx_slot = nullptr;
  b.C.rd.rd[0] = LidxView{(const unsigned long long *)rf_win_off, nullptr, (const unsigned long long *)rf_bnd, rf_mins, rf_win_first, (const unsigned long long *)read_off, read_len, 0, n_reads};
  b.C.rd.rd[1] = LidxView{(const unsigned long long *)rr_win_off, nullptr, (const unsigned long long *)rr_bnd, rr_mins, rr_win_first, (const unsigned long long *)read_off, read_len, 0, n_reads};
  unsigned long long cur[4] = {0, 0, 0, 0}; int err = 0, work = 0;
  b.out.status = status; b.out.n_chains = n_chains; b.out.chain_nseg = chain_nseg; b.out.chain_seg0 = chain_seg0;
  b.out.seg = seg; b.out.seg_cap = seg_cap; b.out.seg_cursor = &cur[0]; b.out.blocks = blocks; b.out.blk_cap = blk_cap; b.out.blk_cursor = &cur[1]; b.out.err = &err;
  b.out.peak = &cur[2];
  b.arena = base; b.arena_per_warp = arena_bytes; b.work = &work; b.order = nullptr; b.n_work = n_reads; b.phase_mask = 0xffffffffu;
  emu::launch(dim3(1), dim3(MP_LANES), 0, [&] { map_reads_kernel(b); });
  counts[0] = cur[0] >> 40; counts[1] = cur[0] & ((1ull << 40) - 1ull); counts[2] = (uint64_t)err; counts[3] = cur[2];
  return err;
}

extern "C" int emu_map_finalize(int n_reads, const MpOpts *opts, const uint64_t *read_off, const uint32_t *read_len, const int *status, const int *n_chains,
                                const int *chain_nseg, const int *chain_seg0, const SegRec *seg, const int32_t *ir_nblk, const uint64_t *ir_off, const uint32_t *ir_blocks,
                                const int32_t *stats, const float *value, const uint64_t *cigar_off, const float *logf_len, lra_b200_record *rec, int *rank,
                                uint64_t *aligned_bases, const float *seg_l, const int32_t *stats_first) {
  FinalBatch b;
  b.n_reads = n_reads; b.o = *opts; b.read_off = (const unsigned long long *)read_off; b.read_len = read_len; b.status = status; b.n_chains = n_chains;
  b.chain_nseg = chain_nseg; b.chain_seg0 = chain_seg0; b.seg = seg; b.ir_nblk = ir_nblk; b.ir_off = (const unsigned long long *)ir_off; b.ir_blocks = ir_blocks;
  b.stats = stats; b.value = value; b.cigar_off = (const unsigned long long *)cigar_off; b.logf_len = logf_len; b.seg_l = seg_l; b.stats_first = stats_first; b.rec = rec; b.rank = rank;
  b.aligned_bases = (unsigned long long *)aligned_bases;
  emu::launch(dim3((unsigned)((n_reads + 127) / 128)), dim3(128), 0, [&] { map_finalize_kernel(b); });
  return 0;
}
extern "C" int emu_sizeof_segrec() { return (int)sizeof(SegRec); }
extern "C" int emu_sizeof_mpopts() { return (int)sizeof(MpOpts); }
extern "C" int emu_sizeof_record() { return (int)sizeof(lra_b200_record); }

// ---- CompareLists: the worker's planned form (mp_compare.cuh) and the literal device form (seed_kernels.cuh) on the same lists
extern "C" long long emu_compare_plan(const uint64_t *qt, int nq, const uint64_t *tt, long long nt, long long maxFreq, int32_t *out_q, uint32_t *out_t, long long cap) {
  std::vector<unsigned char> arena((size_t)nq * 64 + (1 << 20));
  unsigned char *base = arena.data(); while (((uintptr_t)base) & 15) base++;
  long long total = 0;
  emu::launch(dim3(1), dim3(MP_LANES), 0, [&] {
    Arena ar; ar.init(base, arena.size() - 32);
    CmpPlan *plan = ar.alloc<CmpPlan>((unsigned long long)nq + 2);
    const int np = mp_compare_plan((const unsigned long long *)qt, nq, (const unsigned long long *)tt, nt, maxFreq, ar, plan);
    const long long n = mp_compare_expand(plan, np < 0 ? 0 : np, [&](long long slot, int qi, uint32_t ti) { if (slot < cap) { out_q[slot] = qi; out_t[slot] = ti; } });
    if (lane_id() == 0) total = np < 0 ? -1 : n;
  });
  return total;
}
extern "C" long long emu_compare_literal(const uint64_t *qt, int nq, const uint64_t *tt, long long nt, long long maxFreq, int32_t *out_q, uint32_t *out_t, long long cap) {
  long long n = 0;
  lra::mm_compare((const unsigned long long *)qt, (long)nq, (const unsigned long long *)tt, (long)nt, maxFreq, [&](long qi, long ti) { if (n < cap) { out_q[n] = (int32_t)qi; out_t[n] = (uint32_t)ti; } n++; });
  return n;
}
