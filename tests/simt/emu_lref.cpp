// TEST INFRASTRUCTURE ONLY: runs lra_b200/csrc/lidx_kernels.cuh and lref_kernels.cuh on the CPU through the SIMT emulator,
// with the launch sequence of lra_b200/csrc/lref_host.cuh.
#include "emu_common.h"
#include "seed_kernels.cuh"
#include "lidx_kernels.cuh"
#include "lref_kernels.cuh"

using namespace lra;
using namespace emuh;

struct EmuLidx {              // host arrays of one LocalIndex image
  const uint64_t *win_off;    // [n_win + 1]
  const uint32_t *win_len;    // [n_win]
  const uint64_t *bnd;        // [n_win + 1]
  const uint32_t *mins;
  const uint32_t *win_first;  // [n_seq + 1]
  const uint64_t *seq_start;  // [n_seq]
  const uint32_t *seq_len;    // [n_seq]
  int32_t n_win, n_seq;
};

// win_off[n_win] / win_len[n_win] are inputs (the layout of lindex_layout); bnd[n_win + 1] and mins[<= arena length] are outputs.
extern "C" long emu_lindex_build(const uint8_t *ascii, uint64_t n, const uint64_t *win_off, const uint32_t *win_len, int n_win, int k, int w,
                                 int max_freq, uint64_t *bnd, uint32_t *mins) {
  Packed p; pack(ascii, n, p);
  std::vector<uint32_t> tmp(n + 64);
  int err = 0;
  LidxBuild b;
  b.seq = p.view; b.win_off = (const unsigned long long *)win_off; b.win_len = win_len; b.n_win = n_win; b.k = k; b.w = w; b.max_freq = max_freq;
  b.tmp = tmp.data(); b.cnt = (unsigned long long *)bnd; b.mins = mins;
  if (n_win == 0) { bnd[0] = 0; return 0; }
  emu::launch(dim3((unsigned)((n_win + kLidxWarps - 1) / kLidxWarps)), dim3(32 * kLidxWarps), 0, [&] { lidx_window_kernel(b); });
  emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.cnt, n_win, ~0ull, &err); });
  emu::launch(dim3((unsigned)((n_win + 7) / 8)), dim3(256), 0, [&] { lidx_compact_kernel(b); });
  return (long)bnd[n_win];
}

static LidxView view(const EmuLidx *x) {
  return LidxView{(const unsigned long long *)x->win_off, x->win_len, (const unsigned long long *)x->bnd, x->mins, x->win_first,
                  (const unsigned long long *)x->seq_start, x->seq_len, x->n_win, x->n_seq};
}

extern "C" long emu_refine_clusters(const EmuLidx *gl, const EmuLidx *rf, const EmuLidx *rr, int n, const uint32_t *m_q, const uint32_t *m_t,
                                    const uint64_t *m_off, const uint32_t *box, const uint8_t *strand, const uint32_t *read_id,
                                    const uint64_t *hdr_pos, int n_hdr, int global_k, int small_k, int window, long local_max_freq,
                                    int32_t *status, int32_t *chrom, int64_t *diag, uint64_t *r_off, uint32_t *r_q, uint32_t *r_t, uint32_t *r_tup,
                                    uint64_t cap, uint32_t *rbox, float *eff, uint32_t *m_q_out, uint32_t *m_t_out, uint32_t *box_out, uint64_t *counts, int literal,
                                    int mode, const uint32_t *m_len, const uint8_t *m_strand, const int32_t *chrom_in, int limitrefine) {
  const size_t M = (size_t)m_off[n];
  std::vector<unsigned long long> key_off(n + 1), keys, unit_off(n + 2, 0);
  size_t kt = 0;
  for (int c = 0; c < n; c++) { size_t nm = m_off[c + 1] - m_off[c], P = 1; while (P < nm) P <<= 1; key_off[c] = kt; kt += P; }
  keys.resize(kt + 2);
  std::vector<uint32_t> chrom_off(n + 1);
  std::vector<int32_t> ls(n + 1);
  int err = 0;
  LrefBatch b;
  memset(&b, 0, sizeof b);
  b.n_clusters = n; b.in_q = m_q; b.in_t = m_t; b.m_off = (const unsigned long long *)m_off; b.in_box = box; b.strand = strand; b.read_id = read_id;
  b.hdr_pos = (const unsigned long long *)hdr_pos; b.n_hdr = n_hdr; b.gl = view(gl); b.rd[0] = view(rf); b.rd[1] = view(rr);
  b.global_k = global_k; b.small_k = small_k; b.window = window; b.local_max_freq = local_max_freq;
  b.mode = mode; b.limitrefine = limitrefine; b.in_len = m_len; b.in_mstrand = m_strand; b.in_chrom = chrom_in;
  std::vector<uint32_t> fbox(4 * (size_t)n + 4);
  b.fbox = fbox.data();
  std::vector<uint32_t> mqv(M + 4), mtv(M + 4);
  b.m_q = m_q_out ? m_q_out : mqv.data(); b.m_t = m_t_out ? m_t_out : mtv.data(); b.box = box_out; b.keys = keys.data(); b.key_off = key_off.data();
  b.status = status; b.chrom = chrom; b.diag = (long long *)diag; b.chrom_off = chrom_off.data(); b.ls = ls.data(); b.unit_off = unit_off.data();
  b.r_q = r_q; b.r_t = r_t; b.r_tup = r_tup; b.out_cap = cap; b.r_off = (unsigned long long *)r_off; b.rbox = rbox; b.eff = eff;
  emu::launch(dim3((unsigned)((n + 3) / 4)), dim3(128), 0, [&] { lref_prep_kernel(b); });
  emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.unit_off, n, ~0ull, &err); });
  const unsigned long long n_units = unit_off[n];
  std::vector<uint32_t> uc(n_units + 1), uq(n_units + 1), ug(n_units + 1);
  std::vector<unsigned long long> task_off(n_units + 2, 0);
  std::vector<long long> uband(2 * n_units + 2, 0);
  b.u_band = uband.data();
  b.u_cluster = uc.data(); b.u_qis = uq.data(); b.u_gstart = ug.data(); b.task_off = task_off.data();
  unsigned long long n_tasks = 0;
  if (n_units) {
    if (mode == 1) emu::launch(dim3((unsigned)((n + 127) / 128)), dim3(128), 0, [&] { lref_chain_unit_kernel(b); });
    else emu::launch(dim3((unsigned)((n_units + 127) / 128)), dim3(128), 0, [&] { lref_unit_kernel(b, n_units); });
    emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.task_off, (int)n_units, ~0ull, &err); });
    n_tasks = task_off[n_units];
  }
  std::vector<unsigned long long> out_off(n_tasks + 2, 0);
  b.out_off = out_off.data();
  // a small slot, so that window pairs below and above it both occur: the copy kernel and the re-run of the emit pass are both exercised
  std::vector<uint32_t> slot(literal ? n_tasks * 6 * 3 + 3 : 3);
  if (literal) { b.slot = slot.data(); b.slot_cap = 6; }
  if (n_tasks) {
    if (literal) emu::launch(dim3((unsigned)((n_tasks + 127) / 128)), dim3(128), 0, [&] { lref_task_literal_kernel<false>(b, n_units, n_tasks); });
    else emu::launch(dim3((unsigned)((n_tasks + 3) / 4)), dim3(128), 0, [&] { lref_task_kernel<false>(b, n_units, n_tasks); });
    emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.out_off, (int)n_tasks, cap, &err); });
    if (literal) emu::launch(dim3((unsigned)((n_tasks * 8 + 255) / 256)), dim3(256), 0, [&] { lref_task_copy_kernel(b, n_tasks); });
    if (literal) emu::launch(dim3((unsigned)((n_tasks + 127) / 128)), dim3(128), 0, [&] { lref_task_literal_kernel<true>(b, n_units, n_tasks); });
    else emu::launch(dim3((unsigned)((n_tasks + 3) / 4)), dim3(128), 0, [&] { lref_task_kernel<true>(b, n_units, n_tasks); });
  }
  emu::launch(dim3((unsigned)((n + 3) / 4)), dim3(128), 0, [&] { lref_finish_kernel(b, n_units, n_tasks); });
  counts[0] = n_units; counts[1] = n_tasks;
  return (long)out_off[n_tasks];
}

// ---- a6 anchor sorts
#include "sort_kernels.cuh"
extern "C" int emu_sort_matches(int mode, uint32_t *q, uint32_t *t, const uint64_t *seg_off, int n_seg, uint32_t *perm) {
  std::vector<unsigned long long> slot(n_seg + 1);
  size_t slots = 0;
  for (int s = 0; s < n_seg; s++) { size_t n = seg_off[s + 1] - seg_off[s], P = 1; while (P < n) P <<= 1; slot[s] = slots; if (P > (size_t)kSortSmem) slots += P; }
  std::vector<unsigned long long> kp(slots + 2);
  std::vector<uint32_t> ks(slots + 2), ki(slots + 2);
  SortBatch b{n_seg, mode, (const unsigned long long *)seg_off, q, t, perm, kp.data(), ks.data(), ki.data(), slot.data()};
  if (n_seg) emu::launch(dim3((unsigned)n_seg), dim3(256), 0, [&] { sort_pairs_kernel(b); });
  return 0;
}

// ---- a24 GlobalChain
#include "gchain_kernels.cuh"
extern "C" int emu_global_chain(const int32_t *frag, const uint64_t *frag_off, int n_prob, int32_t *score, int32_t *prev, int32_t *chain, int32_t *chain_len) {
  const size_t N = (size_t)frag_off[n_prob];
  std::vector<GcEndpoint> ep(2 * N + 2);
  std::vector<GcVertex> tree(4 * N + 2);
  GcBatch b{n_prob, (const unsigned long long *)frag_off, frag, score, prev, chain, chain_len, ep.data(), tree.data()};
  if (n_prob) emu::launch(dim3((unsigned)((n_prob + 63) / 64)), dim3(64), 0, [&] { gchain_kernel(b); });
  return 0;
}

// ---- a20 RefineBreakpoint
#include "rbp_kernels.cuh"
extern "C" int emu_refine_breakpoint(const uint8_t *fwd, const uint8_t *rcs, uint64_t rn, const uint8_t *genome, uint64_t gn, int n, const uint32_t *lf, const uint32_t *ll,
                                     const uint32_t *rf, const uint32_t *rl, const uint8_t *lstrand, const uint8_t *rstrand, const uint64_t *read_off,
                                     const uint32_t *read_len, const uint64_t *lchrom_off, const uint64_t *rchrom_off, const uint32_t *lchrom_len,
                                     const uint32_t *rchrom_len, int32_t *mode, int32_t *n_out, uint32_t *bound, uint32_t *out, int32_t *refined) {
  Packed pf, pr, pg; pack(fwd, rn, pf); pack(rcs, rn, pr); pack(genome, gn, pg);
  const int slabs = n < 4 ? n : 4;
  std::vector<int32_t> score((size_t)slabs * 2 * kRbpCells), walk((size_t)slabs * 8 * 512);
  std::vector<uint8_t> path((size_t)slabs * 2 * kRbpCells);
  RbpBatch b{n, pf.view, pr.view, pg.view, lf, ll, rf, rl, lstrand, rstrand, (const unsigned long long *)read_off, read_len, (const unsigned long long *)lchrom_off,
             (const unsigned long long *)rchrom_off, lchrom_len, rchrom_len, score.data(), path.data(), walk.data(), mode, n_out, bound, out, refined};
  for (int first = 0; first < n; first += slabs) {
    const int here = n - first < slabs ? n - first : slabs;
    emu::launch(dim3((unsigned)((here + 3) / 4)), dim3(128), 0, [&] { rbp_kernel(b, first, here); });
  }
  return 0;
}

// ---- a16 chain filters
#include "chainf_kernels.cuh"
extern "C" int emu_chain_filter(int mode, const uint32_t *q, const uint32_t *t, const uint32_t *len, const uint8_t *strand, const uint64_t *off, int n_chains, uint8_t *keep) {
  const size_t N = (size_t)off[n_chains];
  std::vector<int32_t> sv(N + 1), svpos(N + 1), svg(N + 1);
  ChainfBatch b{n_chains, mode, (const unsigned long long *)off, q, t, len, strand, keep, sv.data(), svpos.data(), svg.data()};
  if (n_chains) emu::launch(dim3((unsigned)((n_chains + 127) / 128)), dim3(128), 0, [&] { chainf_kernel(b); });
  return 0;
}

// ---- a7 CleanOffDiagonal
#include "cod_kernels.cuh"
extern "C" int emu_clean_off_diagonal(const uint32_t *q, const uint32_t *t, const uint64_t *qt, const uint64_t *off, const uint8_t *strand, int n_lists, const int32_t *opt,
                                      const uint64_t *hdr_pos, int n_hdr, uint8_t *keep, float *freq, int32_t *cnt, int32_t *cl, float *cl_freq, int32_t *n_cl) {
  const size_t N = (size_t)off[n_lists];
  std::vector<uint8_t> flags(3 * N + 8), hused(4 * N + 8);
  std::vector<unsigned long long> hkeys(4 * N + 8);
  CodBatch b;
  b.n_lists = n_lists; b.off = (const unsigned long long *)off; b.q = q; b.t = t; b.qt = (const unsigned long long *)qt; b.strand = strand;
  b.o = CodOpts{opt[0], opt[1], opt[2], opt[3], opt[4], opt[5], opt[6], opt[7], opt[8], opt[9]};
  b.hdr_pos = (const unsigned long long *)hdr_pos; b.n_hdr = n_hdr; b.keep = keep; b.freq = freq; b.cnt = cnt; b.cl = cl; b.cl_freq = cl_freq; b.n_cl = n_cl;
  b.flags = flags.data(); b.hkeys = hkeys.data(); b.hused = hused.data();
  if (n_lists) emu::launch(dim3((unsigned)((n_lists + 63) / 64)), dim3(64), 0, [&] { cod_kernel(b); });
  return 0;
}

// ---- a9 SplitClusters
#include "split_kernels.cuh"
extern "C" long emu_split_clusters(int R, const uint64_t *cl_off, const uint32_t *box, const uint8_t *strand, const float *freq, const uint64_t *m_off, const uint32_t *mq,
                                   int contig, int globalK, uint8_t *split, int32_t *val_cluster, uint64_t *sp_off, uint32_t *sp, int32_t *sp_val, int32_t *sp_n0, uint64_t cap) {
  const size_t Cn = (size_t)cl_off[R];
  std::vector<uint32_t> sets(4 * Cn + 8);
  std::vector<ScPoint> pts(4 * Cn + 8);
  int err = 0;
  SplitBatch b{R, contig, globalK, (const unsigned long long *)cl_off, box, strand, freq, (const unsigned long long *)m_off, mq, split, val_cluster,
               (unsigned long long *)sp_off, sp, sp_val, sp_n0, cap, sets.data(), pts.data()};
  if (R == 0) return 0;
  emu::launch(dim3((unsigned)((R + 63) / 64)), dim3(64), 0, [&] { split_kernel<false>(b); });
  emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.sp_off, R, cap, &err); });
  emu::launch(dim3((unsigned)((R + 63) / 64)), dim3(64), 0, [&] { split_kernel<true>(b); });
  return (long)sp_off[R];
}

// ---- a22 ordering + MAPQ (logv / lenpen are host-computed inputs, as in lref_host.cuh)
#include "mapq_kernels.cuh"
extern "C" int emu_mapq(int R, const int32_t *grp_off, const int32_t *seg_off, const int32_t *upd_off, const int32_t *update_at, const float *value, const int32_t *n0,
                        const int32_t *n1, const int32_t *nm, const int32_t *nmm, const int32_t *ndel, const int32_t *nins, const uint8_t *strand, const float *logv,
                        const int32_t *lenpen, int bypass, int read_type, int32_t *flag, int32_t *typeofaln, uint8_t *issec, uint8_t *supp, int32_t *mapq, uint8_t *g_issec,
                        float *g_value, int32_t *g_n0, int32_t *g_n1, int32_t *g_nm, int32_t *order) {
  MapqBatch b{R, bypass, read_type, grp_off, seg_off, upd_off, update_at, value, n0, n1, nm, nmm, ndel, nins, strand, logv, lenpen, flag, typeofaln, issec, supp, mapq,
              g_issec, g_value, g_n0, g_n1, g_nm, order};
  if (R) emu::launch(dim3((unsigned)((R + 63) / 64)), dim3(64), 0, [&] { mapq_kernel(b); });
  return 0;
}

// ---- a15 LinearExtend (pairs) + DecideCoordinates + TrimOverlappedAnchors
#include "lext_kernels.cuh"
extern "C" int emu_linear_extend(const uint8_t *reads, uint64_t rn, const uint8_t *genome, uint64_t gn, int n_groups, const uint64_t *g_off, const uint64_t *p_off,
                                 const uint8_t *p_strand, const uint64_t *chrom_off, const uint32_t *chrom_len, const uint64_t *read_off, const uint32_t *read_len,
                                 uint32_t *q, uint32_t *t, int K, int skipsorting, int trim, uint64_t *e_off, uint32_t *eq, uint32_t *et, int32_t *elen, uint32_t *box) {
  Packed pr, pg; pack(reads, rn, pr); pack(genome, gn, pg);
  const size_t P = (size_t)g_off[n_groups], N = P ? (size_t)p_off[P] : 0;
  std::vector<uint32_t> part_of(N + 1), end_q(N + 1), end_t(N + 1);
  std::vector<unsigned long long> run(N + 2, 0ull);
  std::vector<int> lidx(N + 1);
  if (!skipsorting && N) {
    std::vector<unsigned long long> slot(P), kp; std::vector<uint32_t> ks, ki;
    size_t slots = 0;
    for (size_t p = 0; p < P; p++) { size_t n = p_off[p + 1] - p_off[p], P2 = 1; while (P2 < n) P2 <<= 1; slot[p] = slots; if (P2 > (size_t)kSortSmem) slots += P2; }
    kp.resize(slots + 2); ks.resize(slots + 2); ki.resize(slots + 2);
    SortBatch sb{(int)P, 0, (const unsigned long long *)p_off, q, t, nullptr, kp.data(), ks.data(), ki.data(), slot.data()};
    emu::launch(dim3((unsigned)P), dim3(256), 0, [&] { sort_pairs_kernel(sb); });
  }
  LextBatch b;
  b.n_groups = n_groups; b.n_parts = (long long)P; b.N = N; b.K = K; b.trim = trim; b.reads = pr.view; b.genome = pg.view;
  b.g_off = (const unsigned long long *)g_off; b.p_off = (const unsigned long long *)p_off; b.p_strand = p_strand; b.chrom_off = (const unsigned long long *)chrom_off;
  b.chrom_len = chrom_len; b.read_off = (const unsigned long long *)read_off; b.read_len = read_len; b.q = q; b.t = t; b.part_of = part_of.data(); b.run = run.data();
  b.end_q = end_q.data(); b.end_t = end_t.data(); b.lidx = lidx.data(); b.e_off = (unsigned long long *)e_off; b.eq = eq; b.et = et; b.elen = elen; b.box = box;
  if (N) {
    const unsigned blocks = (unsigned)((N + 255) / 256);
    int err = 0;
    emu::launch(dim3(blocks), dim3(256), 0, [&] { lext_link_kernel(b); });
    emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.run, (int)N, ~0ull, &err); });
    emu::launch(dim3(blocks), dim3(256), 0, [&] { lext_emit_kernel(b); });
  }
  if (n_groups) emu::launch(dim3((unsigned)((n_groups + 3) / 4)), dim3(128), 0, [&] { lext_group_kernel(b); });
  return 0;
}

extern "C" int emu_linear_extend_chains(const uint8_t *reads, uint64_t rn, const uint8_t *genome, uint64_t gn, int n_chains, const uint64_t *ch_off, const uint32_t *ch,
                                        int n_cl, const uint64_t *cl_off, uint32_t *q, uint32_t *t, const uint32_t *cl_box, const uint8_t *cl_strand, const float *cl_freq,
                                        const uint64_t *chrom_off, const uint32_t *chrom_len, const uint64_t *read_off, const uint32_t *read_len, int K, int skiprepetitive,
                                        int trim, int merge_dist, uint64_t *e_off, uint32_t *eq, uint32_t *et, int32_t *elen, uint8_t *eovp, uint8_t *md_head, uint32_t *box,
                                        int32_t *overlap) {
  Packed pr, pg; pack(reads, rn, pr); pack(genome, gn, pg);
  const size_t U = (size_t)ch_off[n_chains], N = (size_t)cl_off[n_cl];
  if (U == 0) { e_off[0] = 0; return 0; }
  std::vector<unsigned long long> slot_off(U + 1), cnt(U + 2, 0ull);
  std::vector<uint8_t> edge(U, 0);
  size_t S = 0;
  for (int k = 0; k < n_chains; k++)
    for (size_t u = ch_off[k]; u < ch_off[k + 1]; u++) {
      edge[u] = (uint8_t)((u == ch_off[k] ? 1 : 0) | (u + 1 == ch_off[k + 1] ? 2 : 0));
      slot_off[u] = S; S += cl_off[ch[u] + 1] - cl_off[ch[u]];
    }
  slot_off[U] = S;
  if (N) {
    std::vector<unsigned long long> slot(n_cl), kp; std::vector<uint32_t> ks, ki;
    size_t slots = 0;
    for (int c = 0; c < n_cl; c++) { size_t n = cl_off[c + 1] - cl_off[c], P2 = 1; while (P2 < n) P2 <<= 1; slot[c] = slots; if (P2 > (size_t)kSortSmem) slots += P2; }
    kp.resize(slots + 2); ks.resize(slots + 2); ki.resize(slots + 2);
    SortBatch sb{n_cl, 0, (const unsigned long long *)cl_off, q, t, nullptr, kp.data(), ks.data(), ki.data(), slot.data(), cl_strand};
    emu::launch(dim3((unsigned)n_cl), dim3(256), 0, [&] { sort_pairs_kernel(sb); });
  }
  std::vector<uint32_t> sq(S + 1), st(S + 1); std::vector<int32_t> sl(S + 1); std::vector<uint8_t> so(S + 1); std::vector<int> lidx(S + 1);
  LextChainBatch b;
  b.n_units = (int)U; b.K = K; b.skiprepetitive = skiprepetitive; b.trim = trim; b.merge_dist = merge_dist; b.reads = pr.view; b.genome = pg.view;
  b.unit_cl = ch; b.unit_edge = edge.data(); b.slot_off = slot_off.data(); b.cl_off = (const unsigned long long *)cl_off; b.cq = q; b.ct = t; b.cl_box = cl_box;
  b.cl_strand = cl_strand; b.cl_freq = cl_freq; b.cl_chrom_off = (const unsigned long long *)chrom_off; b.cl_chrom_len = chrom_len;
  b.cl_read_off = (const unsigned long long *)read_off; b.cl_read_len = read_len; b.sq = sq.data(); b.st = st.data(); b.sl = sl.data(); b.so = so.data();
  b.cnt = cnt.data(); b.u_overlap = overlap; b.lidx = lidx.data(); b.eq = eq; b.et = et; b.elen = elen; b.eovp = eovp; b.md_head = md_head; b.box = box;
  int err = 0;
  emu::launch(dim3((unsigned)((U + 127) / 128)), dim3(128), 0, [&] { lextc_walk_kernel(b); });
  emu::launch(dim3(1), dim3(1024), 0, [&] { seed_scan_kernel(b.cnt, (int)U, ~0ull, &err); });
  emu::launch(dim3((unsigned)((U + 3) / 4)), dim3(128), 0, [&] { lextc_group_kernel(b); });
  for (size_t u = 0; u <= U; u++) e_off[u] = cnt[u];
  return 0;
}

// ---- a11 SPLITChain (UltimateChain) + MergeSplitchainINS + RemoveSpuriousSplitChain
#include "spchain_kernels.cuh"
extern "C" int emu_split_chains(int n_chains, const uint64_t *c_off, const uint32_t *q, const uint32_t *t, const int32_t *len, const uint8_t *strand, const int32_t *cnum,
                                const uint8_t *link, const uint64_t *hdr_pos, int n_hdr, int splitdist, int bypass, int32_t *n_sp, int32_t *n_link, int32_t *sp_off,
                                int32_t *ci_off, int32_t *sptc, int32_t *ci, uint8_t *sp_lk, uint32_t *sp_box, int32_t *sp_chrom, uint8_t *sp_type, uint8_t *sp_strand,
                                uint8_t *sp_link) {
  const size_t N = (size_t)c_off[n_chains] + 1;
  std::vector<int32_t> pa(N), pb(N), pnext(N), tail(N), size(N), chrom(N), cur_ind(N), ord(N);
  std::vector<uint32_t> QS(N), QE(N), TS(N), TE(N);
  std::vector<uint8_t> type(N), pstrand(N), keep(N), SL(N);
  SpChainBatch b{n_chains, splitdist, bypass, (const unsigned long long *)c_off, q, t, len, strand, cnum, link, (const unsigned long long *)hdr_pos, n_hdr,
                 pa.data(), pb.data(), pnext.data(), tail.data(), size.data(), chrom.data(), cur_ind.data(), ord.data(), QS.data(), QE.data(), TS.data(), TE.data(),
                 type.data(), pstrand.data(), keep.data(), SL.data(), n_sp, n_link, sp_off, ci_off, sptc, ci, sp_lk, sp_box, sp_chrom, sp_type, sp_strand, sp_link};
  if (n_chains) emu::launch(dim3((unsigned)((n_chains + 63) / 64)), dim3(64), 0, [&] { spchain_kernel(b); });
  return 0;
}

// ---- a11 MergeChain / switchindex
#include "cglue_kernels.cuh"
extern "C" int emu_merge_chain(uint64_t n_entries, const int32_t *sp, const uint8_t *first, const int32_t *chrom, const uint8_t *strand, const uint32_t *box, uint8_t *head) {
  MergeChainBatch b{n_entries, sp, first, chrom, strand, box, head};
  if (n_entries) emu::launch(dim3((unsigned)((n_entries + 255) / 256)), dim3(256), 0, [&] { merge_chain_kernel(b); });
  return 0;
}
extern "C" int emu_switchindex(int n_chains, const uint64_t *c_off, int32_t *ch, uint8_t *link, const int32_t *coarse, const uint32_t *cq, int32_t *n_out, int32_t *nl_out) {
  const size_t E = (size_t)c_off[n_chains] + 1;
  std::vector<int32_t> ss(E), se(E), newch(E); std::vector<uint8_t> newlink(E), flag(E);
  SwitchIndexBatch b{n_chains, (const unsigned long long *)c_off, ch, link, coarse, cq, ss.data(), se.data(), newch.data(), newlink.data(), flag.data(), n_out, nl_out};
  if (n_chains) emu::launch(dim3((unsigned)((n_chains + 63) / 64)), dim3(64), 0, [&] { switchindex_kernel(b); });
  return 0;
}

// ---- a8 (first half) SplitRoughClustersWithGaps
#include "srough_kernels.cuh"
extern "C" int emu_split_rough(int n_lists, const uint64_t *l_off, const uint64_t *lr_off, const uint32_t *q, const uint32_t *t, const int32_t *r_start, const int32_t *r_end,
                               const uint32_t *r_box, const uint8_t *r_strand, const float *r_freq, const int32_t *r_chrom, int globalK, int maxGap, int minClusterSize, int maxDiag,
                               int32_t *n_split, int32_t *n_piece, int32_t *s_start, int32_t *s_end, int32_t *s_coarse, int32_t *s_chrom, uint32_t *s_box, uint8_t *s_strand,
                               float *s_freq, int32_t *p_cluster, int32_t *p_start, int32_t *p_end) {
  SplitRoughBatch b{n_lists, globalK, maxGap, minClusterSize, maxDiag, (const unsigned long long *)l_off, (const unsigned long long *)lr_off, q, t, r_start, r_end, r_box,
                    r_strand, r_freq, r_chrom, n_split, n_piece, s_start, s_end, s_coarse, s_chrom, s_box, s_strand, s_freq, p_cluster, p_start, p_end};
  if (n_lists) emu::launch(dim3((unsigned)((n_lists + 63) / 64)), dim3(64), 0, [&] { split_rough_kernel(b); });
  return 0;
}

extern "C" int emu_store_diagonal(int n_lists, const uint64_t *l_off, const uint32_t *q, const uint32_t *t, const uint64_t *qt, const float *freq, const uint8_t *strand,
                                  const uint64_t *hdr_pos, int n_hdr, int globalK, int maxDiag, int minClusterSize, int minClusterLength, int bypass, int32_t *n_cl,
                                  int32_t *c_start, int32_t *c_end, int32_t *c_chrom, uint32_t *c_box, float *c_freq) {
  StoreDiagBatch b{n_lists, globalK, maxDiag, minClusterSize, minClusterLength, bypass, (const unsigned long long *)l_off, q, t, (const unsigned long long *)qt, freq, strand,
                   (const unsigned long long *)hdr_pos, n_hdr, n_cl, c_start, c_end, c_chrom, c_box, c_freq};
  if (n_lists) emu::launch(dim3((unsigned)((n_lists + 63) / 64)), dim3(64), 0, [&] { store_diagonal_kernel(b); });
  return 0;
}

extern "C" int emu_trim_splitchains(int n_chains, const uint64_t *c_off, const uint32_t *cq, const uint32_t *ct, const uint8_t *strand, const uint64_t *m_off, uint32_t *q,
                                    uint32_t *t, uint8_t *keep, int32_t *removed) {
  std::vector<uint8_t> mode(n_chains + 1);
  std::vector<unsigned long long> slot(n_chains + 1), kp; std::vector<uint32_t> ks, ki;
  size_t slots = 0;
  for (int c = 0; c < n_chains; c++) {
    mode[c] = (c_off[c + 1] - c_off[c]) == 1 ? 255 : 2;
    size_t n = m_off[c + 1] - m_off[c], P2 = 1; while (P2 < n) P2 <<= 1; slot[c] = slots; if (mode[c] == 2 && P2 > (size_t)kSortSmem) slots += P2;
  }
  kp.resize(slots + 2); ks.resize(slots + 2); ki.resize(slots + 2);
  if (n_chains && m_off[n_chains]) {
    SortBatch sb{n_chains, 2, (const unsigned long long *)m_off, q, t, nullptr, kp.data(), ks.data(), ki.data(), slot.data(), mode.data()};
    emu::launch(dim3((unsigned)n_chains), dim3(256), 0, [&] { sort_pairs_kernel(sb); });
  }
  TrimChainBatch b{n_chains, (const unsigned long long *)c_off, cq, ct, strand, (const unsigned long long *)m_off, q, t, keep, removed};
  if (n_chains) emu::launch(dim3((unsigned)((n_chains + 63) / 64)), dim3(64), 0, [&] { trim_splitchain_kernel(b); });
  return 0;
}
