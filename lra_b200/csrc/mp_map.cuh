// The mapper kernels: map_reads_kernel runs MapRead (MapRead.h:153-263) + MapRead_lowacc (Map_lowacc.h:69-632) up to and including
// LocalRefineAlignment for one read per warp and emits the segments (block lists) of every chain; the batched a19 / a21 kernels then refine and score
// the segments, and map_finalize_kernel applies SetFromSegAlignment, AlignmentsOrder::Update and SimpleMapQV (Alignment.h:944-1048,
// Mapping_ultility.h:497-589) per read.
#pragma once
#include "mp_align.cuh"

namespace lra {
namespace mp {

constexpr int kMaxChains = 4;            // opts.NumAln is 2 or 3 under every preset
constexpr int kMaxSegPerChain = 256;         // segments of one alignment (a contig with seeded inversions has a few per inversion)

struct MapOut {
  // per read
  int *status;                            // MP_OK / MP_UNALIGNED / error
  int *n_chains;                          // alignments.size()
  int *chain_nseg;                        // [n_reads][kMaxChains]
  int *chain_seg0;                        // [n_reads][kMaxChains] first segment id
  // segments (global, reserved atomically)
  SegRec *seg; int seg_cap; unsigned long long *seg_cursor;
  uint32_t *blocks; unsigned long long blk_cap; unsigned long long *blk_cursor;
  int *err;                               // bit0 segment capacity, bit1 block capacity
  unsigned long long *peak;               // arena high-water mark
};

__device__ __noinline__ bool mp_segbuild_alloc(SegBuild &B, Arena &ar, const UChain *uc, int ng, uint32_t L);
__device__ __noinline__ int mp_emit_segments(int r, int p, const SegBuild &B, const MapOut &out, int &nseg_out, int &seg0_out);

// one chain p of a read: everything from SPLITChain to LocalRefineAlignment.  Returns 0 ok (segments appended, possibly none),
// 1 "this chain ends the loop over chains" (Map_lowacc.h: unaligned for p == 0, break for p > 0), < 0 error
__device__ __noinline__ int mp_map_chain(const MpCtx &C, int r, Arena &ar, const ClusterSet &ext, const UChain &chain, int p, const MapOut &out, int &nseg_out, int &seg0_out) {
  const int lane = lane_id();
  const MpOpts &O = C.o;
  const uint32_t L = C.rd.read_len[r];
  nseg_out = 0; seg0_out = 0;
  unsigned long long tk = mp_clock();
  SplitSet sp;
  if (!mp_split_chain(C, ext, chain, ar, sp)) return -MP_ERR_ARENA;
  tk = mp_tick(C, PF_SPLIT, tk);
  if (sp.n == 0) return 1;
  // ---- Refine_splitchain, Refine_Btwnsplitchain
  RCluster *RC = ar.alloc<RCluster>(sp.n);
  RSeg *nodes = ar.alloc<RSeg>(sp.n);
  if (ar.overflow) return -MP_ERR_ARENA;
  for (int ph = 0; ph < sp.n; ph++) if (!mp_refine_splitchain(C, r, ext, chain, sp, ph, ar, RC[ph], nodes + ph)) return -MP_ERR_ARENA;
  tk = mp_tick(C, PF_REFINE_SPLIT, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  if (!mp_refine_btwn_splitchain(C, r, ar, sp, RC)) return -MP_ERR_ARENA;
  tk = mp_tick(C, PF_REFINE_BTWN, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  // ---- MergeChain (ChainRefine.h:767-802): groups of consecutive refined clusters
  int *grp = ar.alloc<int>(sp.n + 1);     // group id of every refined cluster
  int *ng_p = ar.alloc<int>(2);
  if (ar.overflow) return -MP_ERR_ARENA;
  if (lane == 0) {
    int g = 0; grp[0] = 0;
    for (int t = 1; t < sp.n; t++) {
      const RCluster &cur = RC[t], &prev = RC[t - 1];
      int qdist = 9999, tdist = 9999;
      if (prev.chrom == cur.chrom && prev.strand == cur.strand) {
        qdist = (prev.qS > cur.qE) ? (int)(prev.qS - cur.qE) : 0;
        if (prev.strand == 0) tdist = (prev.tS >= cur.tE) ? (int)(prev.tS - cur.tE) : 9999;
        else if (prev.strand == 1) tdist = (prev.tE <= cur.tS) ? (int)(cur.tS - prev.tE) : 9999;
      }
      if (!(qdist <= 500 && tdist <= 500)) g++;
      grp[t] = g;
    }
    int tot = 0; for (int t = 0; t < sp.n; t++) tot += RC[t].n;
    ng_p[0] = g + 1; ng_p[1] = tot;
  }
  wsync();
  const int ng = ng_p[0], total_refined = ng_p[1];
  // ---- LinearExtend of every refined cluster into its group's extended cluster (Map_lowacc.h:458-474), TrimOverlappedAnchors
  ClusterSet xs;
  if (!mp_alloc_clusterset(xs, ar, ng, total_refined, true)) return -MP_ERR_ARENA;
  if (lane == 0) xs.off[0] = 0;
  {
    int o = 0, t = 0;
    for (int g = 0; g < ng; g++) {
      int st = 0, chrom = 0; float freq = 0.0f;
      const int o_begin = o;
      while (t < sp.n && grp[t] == g) {
        RCluster &rc = RC[t];
        st = rc.strand != 0 ? 1 : 0; chrom = rc.chrom; freq = rc.freq;
        const int n = rc.n;
        if (n > 0) {
          // DiagonalSort of the cluster's matches (skipsorting == 0), then the extension
          const unsigned long long mk = ar.mark();
          MpKey *keys = ar.alloc<MpKey>((unsigned long long)next_pow2(n));
          uint32_t *gq = ar.alloc<uint32_t>(n), *gt = ar.alloc<uint32_t>(n), *sq = ar.alloc<uint32_t>(n), *stt = ar.alloc<uint32_t>(n);
          if (ar.overflow) return -MP_ERR_ARENA;
          { int k = 0; for (RSeg *s = rc.head; s; s = s->next) { const int sn = s->n; for (int i = lane; i < sn; i += kLanes) { gq[k + i] = s->q[i]; gt[k + i] = s->t[i]; } k += sn; } }
          wsync();
          for (int i = lane; i < n; i += kLanes) { keys[i].k = (unsigned long long)((long long)gq[i] - (long long)gt[i] + (1ll << 33)); keys[i].q = gq[i]; keys[i].idx = (uint32_t)i; }
          wsync();
          mp_sort_keys(keys, n);
          for (int i = lane; i < n; i += kLanes) { sq[i] = gq[keys[i].idx]; stt[i] = gt[keys[i].idx]; }
          wsync();
          o = mp_linear_extend_warp(C, C.rd.read_off[r], L, chrom, sq, stt, n, st, O.smallK, xs.q, xs.t, xs.len, o);
          ar.release(mk);
        }
        t++;
      }
      if (lane == 0) {
        xs.off[g + 1] = o;
        xs.strand[g] = -1; xs.chrom[g] = 0; xs.freq[g] = 0.0f; xs.qS[g] = 0xffffffffu; xs.qE[g] = 0; xs.tS[g] = 0xffffffffu; xs.tE[g] = 0;
        mp_decide_coordinates(xs, g, st, chrom, freq);
      }
      wsync();
      (void)o_begin;
    }
    xs.ncl = ng;
    // TrimOverlappedAnchors(extend_clusters, 0)
    wsync();
    for (int g = 0; g < ng; g++) {
      const int a0 = xs.off[g], n = xs.off[g + 1] - a0;
      if (n == 0) continue;
      if (!mp_trim_overlapped_warp(xs.q + a0, xs.t + a0, xs.len + a0, n, xs.strand[g], 40, true, ar)) return -MP_ERR_ARENA;
    }
    wsync();
  }
  tk = mp_tick(C, PF_LEXT2, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  if (total_refined == 0) return 1;
  // ---- second SparseDP per extended cluster + RemovePairedIndels + RemoveSpuriousAnchors
  UChain *uc = ar.alloc<UChain>(ng);
  uint8_t *clst = ar.alloc<uint8_t>(ng + 1);
  float *valp = ar.alloc<float>(1);
  if (ar.overflow) return -MP_ERR_ARENA;
  for (int g = lane; g < ng; g += kLanes) clst[g] = (uint8_t)(xs.strand[g] != 0);
  wsync();
  SdpAnchors A; A.q = xs.q; A.t = xs.t; A.len = xs.len; A.nfrag = xs.off[ng]; A.cl_off = xs.off; A.cl_strand = clst; A.ncl = ng;
  for (int g = 0; g < ng; g++) {
    const int nf = xs.off[g + 1] - xs.off[g];
    UChain u; u.n = 0; u.nlink = 0; u.FirstSDPValue = 0.0f; u.NumOfAnchors0 = chain.NumOfAnchors0; u.NumOfAnchors1 = 0; u.QStart = u.QEnd = u.TStart = u.TEnd = 0;
    u.idx = ar.alloc<uint32_t>(nf + 1); u.cl = ar.alloc<int>(nf + 1); u.link = ar.alloc<uint8_t>(nf + 1);
    if (ar.overflow) return -MP_ERR_ARENA;
    if (nf > 0) {
      if (lane == 0) *valp = 0.0f;
      wsync();
      const int n = sdp_one_cluster(A, g, O.second_anchorbonus, *C.pwl, ar, u.idx, u.link, valp);
      if (n < 0) return -MP_ERR_ARENA;
      wsync();
      u.n = n; u.nlink = n > 0 ? n - 1 : 0; u.FirstSDPValue = *valp; u.NumOfAnchors1 = n;
      for (int i = lane; i < n; i += kLanes) u.cl[i] = g;
      wsync();
      mp_chain_filter(1, xs, u, ar, true);      // RemovePairedIndels<UltimateChain>(chain) (refineEnds = true)
      mp_chain_filter(4, xs, u, ar, false);     // RemoveSpuriousAnchors
    }
    if (lane == 0) uc[g] = u;
    wsync();
  }
  tk = mp_tick(C, PF_SDP2, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  // LargestSplitChain
  int LSC = 0;
  for (int g = 1; g < ng; g++) if (uc[g].n > uc[LSC].n) LSC = g;
  // ---- LocalRefineAlignment -> segments
  SegBuild B;
  if (!mp_segbuild_alloc(B, ar, uc, ng, L)) return -MP_ERR_ARENA;
  if (!mp_local_refine_alignment(C, r, ar, B, xs, uc, ng, LSC, 2)) return ar.overflow ? -MP_ERR_ARENA : -MP_ERR_CAP;
  wsync();
  tk = mp_tick(C, PF_LOCAL_REFINE, tk);
  mp_phase(ar);
  tk = mp_tick(C, PF_BARRIER, tk);
  const int rc_emit = mp_emit_segments(r, p, B, out, nseg_out, seg0_out);
  tk = mp_tick(C, PF_OUTPUT, tk);
  return rc_emit;
}

// the SegBuild of one chain: room for the anchors of its ultimate chains and for the linear alignments between them
__device__ __noinline__ bool mp_segbuild_alloc(SegBuild &B, Arena &ar, const UChain *uc, int ng, uint32_t L) {
  int max_blocks = 64;
  for (int g = 0; g < ng; g++) max_blocks += 2 * uc[g].n + 8;
  max_blocks += (int)(L / 2u) + (int)(L / 8u);     // blocks of the linear alignments between the anchors (a block covers at least one base, gaps at least one more)
  B.cap = max_blocks; B.nblk = 0; B.nseg = 0; B.cap_seg = kMaxSegPerChain;
  B.blk = ar.alloc<uint32_t>(3ull * max_blocks);
  B.seg_start = ar.alloc<int>(kMaxSegPerChain + 1); B.seg_strand = ar.alloc<int>(kMaxSegPerChain); B.seg_chrom = ar.alloc<int>(kMaxSegPerChain);
  B.seg_n0 = ar.alloc<int>(kMaxSegPerChain); B.seg_n1 = ar.alloc<int>(kMaxSegPerChain); B.seg_supp = ar.alloc<int>(kMaxSegPerChain); B.seg_val = ar.alloc<float>(kMaxSegPerChain);
  B.cap_jobs = 4096; B.njobs = 0;
  B.jobs = ar.alloc<AogJob>(B.cap_jobs);
  return !ar.overflow;
}

// hand the segments of chain p of read r to the global lists.  Returns 0 (also with no segment), < 0 on a capacity error.
__device__ __noinline__ int mp_emit_segments(int r, int p, const SegBuild &B, const MapOut &out, int &nseg_out, int &seg0_out) {
  const int lane = lane_id();
  if (B.nseg == 0) { nseg_out = 0; return 0; }
  // ---- hand the segments to the global lists
  // one atomic reserves the segment ids (high 24 bits) and the block range (low 40 bits) together, so that in segment order the block offsets are
  // the exclusive prefix sum of the block counts (the layout the a19 / a21 batch kernels take)
  unsigned long long s0 = 0, b0 = 0;
  if (lane == 0) {
    const unsigned long long c = atomicAdd(out.seg_cursor, ((unsigned long long)B.nseg << 40) | (unsigned long long)B.nblk);
    s0 = c >> 40; b0 = c & ((1ull << 40) - 1ull);
  }
  s0 = bcast(s0, 0); b0 = bcast(b0, 0);
  if (s0 + (unsigned long long)B.nseg > (unsigned long long)out.seg_cap) { if (lane == 0) atomicOr(out.err, 1); return -MP_ERR_CAP; }
  if (b0 + (unsigned long long)B.nblk > out.blk_cap) { if (lane == 0) atomicOr(out.err, 2); return -MP_ERR_CAP; }
  for (int i = lane; i < 3 * B.nblk; i += kLanes) out.blocks[3ull * b0 + i] = B.blk[i];
  for (int s = lane; s < B.nseg; s += kLanes) {
    SegRec x;
    x.read = r; x.chain = p; x.order_in_chain = s; x.strand = B.seg_strand[s]; x.chrom = B.seg_chrom[s]; x.NumOfAnchors0 = B.seg_n0[s]; x.NumOfAnchors1 = B.seg_n1[s];
    x.Supplymentary = B.seg_supp[s]; x.ISsecondary = (s == 0 && p > 0) ? 1 : 0; x.FirstSDPValue = B.seg_val[s];
    const int bs = B.seg_start[s], be = s + 1 < B.nseg ? B.seg_start[s + 1] : B.nblk;
    x.blk_off = b0 + (unsigned long long)bs; x.blk_cnt = be - bs;
    out.seg[s0 + s] = x;
  }
  wsync();
  nseg_out = B.nseg; seg0_out = (int)s0;
  return 0;
}

}  // namespace mp
}  // namespace lra
#include "mp_highacc.cuh"
namespace lra {
namespace mp {

__device__ __noinline__ void mp_map_read(const MpCtx &C, int r, Arena &ar, const MapOut &out) {
  const int lane = lane_id();
  ar.top = 0; ar.overflow = 0;
  int status = MP_OK, n_al = 0;
  int nseg[kMaxChains], seg0[kMaxChains];
  for (int p = 0; p < kMaxChains; p++) { nseg[p] = 0; seg0[p] = 0; }
  ClusterSet ext; UChain *chains = 0; int nch = 0;
  ar.phase = 0;
  const int PB = (C.o.NumAln < kMaxChains ? C.o.NumAln : kMaxChains);      // chains a read can have (uniform over the batch)
  if (C.o.HighlyAccurate) {
    // MapRead_highacc (Map_highacc.h:37-798): every chain of Primary_chains[0] becomes an alignment (:690-735)
    HaState hs;
    status = mp_stage1_highacc(C, r, ar, hs);
    mp_phase_upto(ar, kPhasesStage1);
    if (status == MP_OK) {
      const unsigned long long mk = ar.mark();
      for (int h = 0; h < hs.H.nch && h < kMaxChains; h++) {
        if (hs.H.n[h] == 0) continue;
        ar.release(mk);
        wsync();
        int ns = 0, s0 = 0;
        const int rc = mp_map_chain_highacc(C, r, ar, hs, h, out, ns, s0);
        mp_phase_upto(ar, kPhasesStage1 + kPhasesChain * (h + 1));
        if (rc < 0) { status = -rc; break; }
        nseg[n_al] = ns; seg0[n_al] = s0; n_al++;
      }
      if (status == MP_OK && n_al == 0) status = MP_UNALIGNED;
    }
  } else {
  status = mp_stage1(C, r, ar, ext, chains, nch);
  mp_phase_upto(ar, kPhasesStage1);
  if (status == MP_OK) {
    const unsigned long long mk = ar.mark();
    for (int p = 0; p < nch && p < kMaxChains; p++) {
      ar.release(mk);
      wsync();
      const UChain ch = chains[p];
      int ns = 0, s0 = 0;
      const int rc = mp_map_chain(C, r, ar, ext, ch, p, out, ns, s0);
      mp_phase_upto(ar, kPhasesStage1 + kPhasesChain * (p + 1));
      if (rc < 0) { status = -rc; break; }
      if (rc == 1) { if (p == 0) status = MP_UNALIGNED; break; }
      // alignments.resize(+1) happened for this chain; p == 0 without a segment makes the read unaligned (Map_lowacc.h:577-580)
      if (p == 0 && ns == 0) { status = MP_UNALIGNED; break; }
      nseg[n_al] = ns; seg0[n_al] = s0; n_al++;
    }
  }
  }
  mp_phase_upto(ar, kPhasesStage1 + kPhasesChain * PB);
  if (status != MP_OK) n_al = 0;
  if (lane == 0) {
    out.status[r] = status; out.n_chains[r] = n_al;
    for (int p = 0; p < kMaxChains; p++) { out.chain_nseg[r * kMaxChains + p] = p < n_al ? nseg[p] : 0; out.chain_seg0[r * kMaxChains + p] = p < n_al ? seg0[p] : 0; }
    if (out.peak) atomicMax(out.peak, ar.peak);
  }
  wsync();
}

struct MapBatch {
  MpCtx C;
  MapOut out;
  unsigned char *arena; unsigned long long arena_per_warp;
  int *work;                              // dynamic read counter
  const int *order;                       // optional: reads in decreasing length (longest first), or the reads of a retry pass
  int n_work;                             // reads to map in this launch (entries of `order`)
  unsigned phase_mask;                    // which phase points are CTA barriers (mp_phase)
};

// One CTA per SM; its warps take a group of consecutive reads of the length-sorted order (similar lengths, similar stage times) through the
// stages in lock step (mp_phase).  blockDim.x = 32 * warps per CTA (host: 24 warps, 80 registers per thread: measured 576 ms per 16 k ONT reads
// against 633 ms for 16 warps with 128 registers and 692+ ms for 32 warps with 64, profiles/r02_experiments.md).
#ifndef MP_BLOCK_THREADS
#define MP_BLOCK_THREADS 768
#endif
__global__ void __launch_bounds__(MP_BLOCK_THREADS, 1) map_reads_kernel(MapBatch b) {
  const int warps_per_block = (int)blockDim.x / kLanes;
  const int wib = (int)threadIdx.x / kLanes;
  const int wid = (int)blockIdx.x * warps_per_block + wib;
  Arena ar; ar.init(b.arena + (unsigned long long)wid * b.arena_per_warp, b.arena_per_warp);
  if (b.C.prof) ar.prof = b.C.prof + (unsigned long long)wid * kProfStages;
  ar.phase_mask = b.phase_mask;
  const int PB = (b.C.o.NumAln < kMaxChains ? b.C.o.NumAln : kMaxChains);
#if !defined(LRA_EMU)
  __shared__ int s_base;
#endif
  for (;;) {
    int base;
#if !defined(LRA_EMU)
    __syncthreads();
    if (threadIdx.x == 0) s_base = atomicAdd(b.work, warps_per_block);
    __syncthreads();
    base = s_base;
#else
    base = 0;
    if (lane_id() == 0) base = atomicAdd(b.work, warps_per_block);
    base = bcast(base, 0);
#endif
    if (base >= b.n_work) break;
    const int w = base + wib;
    if (w < b.n_work) mp_map_read(b.C, b.order ? b.order[w] : w, ar, b.out);
    else { ar.phase = 0; mp_phase_upto(ar, kPhasesStage1 + kPhasesChain * PB); }
  }
}

// ---- finalize: per read, after IndelRefineAlignment and CalculateStatistics of every segment ---------------------------------------------------------
// x86-64 cvttss2si: out-of-range and NaN give INT_MIN (the reference's (int) casts of a float); the GPU conversion saturates
__device__ __forceinline__ int x86_f2i(float f) { return (f >= 2147483648.0f || f < -2147483648.0f || f != f) ? (int)0x80000000 : (int)f; }

struct FinalBatch {
  int n_reads;
  MpOpts o;
  const unsigned long long *read_off; const uint32_t *read_len;
  const int *status, *n_chains, *chain_nseg, *chain_seg0;
  const SegRec *seg;
  // a19 output blocks of every segment and a21 output (stats_kernels.cuh: 16 ints per segment, value, cigar offsets)
  const int32_t *ir_nblk; const unsigned long long *ir_off; const uint32_t *ir_blocks;
  const int32_t *stats; const float *value; const unsigned long long *cigar_off;
  const float *logf_len;                  // host-built: logf(len) for len = 0..kMaxChains
  const int32_t *stats_first;             // MapRead_highacc calls CalculateStatistics twice (Map_highacc.h:720, 729) and tdel / tins / the six size-class counters
                                          // are never reset (Alignment.h:418-505): the printed values are the sums over both calls.  nullptr: one call
  const float *seg_l;                     // host-built (glibc logf), presets without bypassClustering: value > 3 ? logf(value / globalK) : 0 per segment
  lra_b200_record *rec;                   // [n_seg]
  int *rank;                              // [n_reads][kMaxChains]
  unsigned long long *aligned_bases;
};

__global__ void __launch_bounds__(128) map_finalize_kernel(FinalBatch b) {
  const int r = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
  if (r >= b.n_reads) return;
  for (int p = 0; p < kMaxChains; p++) b.rank[r * kMaxChains + p] = p;
  if (b.status[r] != MP_OK) return;
  const int na = b.n_chains[r];
  const uint32_t L = b.read_len[r];
  // SetFromSegAlignment (Alignment.h:944-983) per alignment; the record of every segment
  float gval[kMaxChains]; int gn0[kMaxChains];
  for (int a = 0; a < na; a++) {
    const int ns = b.chain_nseg[r * kMaxChains + a], s0 = b.chain_seg0[r * kMaxChains + a];
    gval[a] = 0.0f; gn0[a] = 0;
    if (ns == 0) continue;
    gn0[a] = b.seg[s0].NumOfAnchors0;
    float v = 0.0f;
    int pry = 0;
    for (int s = 0; s < ns; s++) { v = __fadd_rn(v, b.value[s0 + s]); if (b.seg[s0 + s].Supplymentary == 0) pry++; }
    gval[a] = v;
    for (int s = 0; s < ns; s++) {
      const SegRec &sg = b.seg[s0 + s];
      const int32_t *st = b.stats + 16ll * (s0 + s);
      lra_b200_record x;
      x.read = r; x.chain = a; x.seg = s; x.n_seg = ns;
      int supp = sg.Supplymentary;
      if (s == 0 && pry == 0) supp = 0;
      x.flag = 0;
      if (sg.strand == 1) x.flag |= 0x10u;
      if (supp == 1) x.flag |= 0x800u;
      x.chrom = sg.chrom; x.strand = sg.strand; x.mapq = 0; x.order = ns - 1 - s; x.typeofaln = 0; x.supplementary = supp;
      const int nb = b.ir_nblk[s0 + s];
      const uint32_t *bl = b.ir_blocks + 3ull * b.ir_off[s0 + s];
      x.n_blocks = nb;
      x.tStart = x.tEnd = x.qStart = x.qEnd = 0; x.preClip = 0; x.sufClip = 0;
      if (nb > 0) {
        x.preClip = (int)bl[0]; x.sufClip = (int)(L - bl[3 * (nb - 1)] - bl[3 * (nb - 1) + 2]);
        x.qStart = bl[0]; x.qEnd = bl[3 * (nb - 1)] + bl[3 * (nb - 1) + 2]; x.tStart = bl[1]; x.tEnd = bl[3 * (nb - 1) + 1] + bl[3 * (nb - 1) + 2];
      }
      // members after one CalculateStatistics call: the D count lands in `nins`, the I count in `ndel` (Alignment.h:414 vs :516)
      x.nm = st[0]; x.nmm = st[1]; x.nins = st[2]; x.ndel = st[3]; x.tdel = st[4]; x.tins = st[5];
      x.nSmallDel = st[6]; x.nMedDel = st[7]; x.nLargeDel = st[8]; x.nSmallIns = st[9]; x.nMedIns = st[10]; x.nLargeIns = st[11];
      if (b.stats_first) {
        const int32_t *s1 = b.stats_first + 16ll * (s0 + s);
        x.tdel += s1[4]; x.tins += s1[5]; x.nSmallDel += s1[6]; x.nMedDel += s1[7]; x.nLargeDel += s1[8]; x.nSmallIns += s1[9]; x.nMedIns += s1[10]; x.nLargeIns += s1[11];
      }
      x.value = b.value[s0 + s];
      x.NumOfAnchors0 = sg.NumOfAnchors0; x.NumOfAnchors1 = sg.NumOfAnchors1;
      x.cigar_off = b.cigar_off[s0 + s]; x.n_cigar = (int)(b.cigar_off[s0 + s + 1] - b.cigar_off[s0 + s]);
      b.rec[s0 + s] = x;
    }
  }
  // AlignmentsOrder::Update (Alignment.h:1024-1048): std::sort of the alignment indices by (value desc, NumOfAnchors0 desc)
  int idx[kMaxChains];
  for (int a = 0; a < na; a++) idx[a] = a;
  std_sort_replay(idx, na, [&](int i, int j) { if (gval[i] != gval[j]) return gval[i] > gval[j]; return gn0[i] > gn0[j]; });
  for (int a = 0; a < na; a++) b.rank[r * kMaxChains + a] = idx[a];
  for (int k = 1; k < na; k++) {
    const int a = idx[k];
    const int ns = b.chain_nseg[r * kMaxChains + a], s0 = b.chain_seg0[r * kMaxChains + a];
    for (int s = 0; s < ns; s++) { b.rec[s0 + s].flag |= 0x100u; if (b.rec[s0 + s].typeofaln != 3) b.rec[s0 + s].typeofaln = 2; }
  }
  if (na == 0) return;
  // SimpleMapQV (Mapping_ultility.h:497-589), bypassClustering presets (the logf(value / K) factor only enters the other branch)
  const float q_coef = (b.o.bypassClustering && b.o.readType == 1) ? 4.0f : ((b.o.bypassClustering && b.o.readType == 0) ? 30.0f : 1.0f);
  const int a = idx[0];
  const int ns = b.chain_nseg[r * kMaxChains + a], s0 = b.chain_seg0[r * kMaxChains + a];
  float x = 0.0f, y = 1.0f;
  if (na > 1) x = __fdiv_rn(gval[idx[1]], gval[a]);
  unsigned long long ab = 0;
  for (int s = ns - 1; s >= 0; s--) {
    lra_b200_record &rc = b.rec[s0 + s];
    const int n0 = rc.NumOfAnchors0;
    if (na > 1 && b.o.bypassClustering) y = __fdiv_rn((float)gn0[a], (float)gn0[idx[1]]);
    float pen;
    if (!b.o.bypassClustering) { pen = __fmul_rn(n0 > 20 ? 1.0f : 0.05f, (float)n0); pen = __fmul_rn(n0 >= 5 ? 1.0f : 0.1f, pen); y = 1.0f; }
    else { pen = __fmul_rn(n0 > 10 ? 1.0f : 0.05f, (float)n0); pen = __fmul_rn(n0 >= 5 ? 1.0f : 0.02f, pen); }
    const float l = b.o.bypassClustering ? 1.0f : b.seg_l[s0 + s];
    const int den = rc.nmm + rc.ndel + rc.nins;
    float identity = den == 0 ? 1.0f : __fdiv_rn((float)rc.nm, (float)den);
    identity = identity < 1 ? identity : 1;
    long long mapq;
    if (na == 1) {
      if (!b.o.bypassClustering) mapq = (long long)x86_f2i(__fmul_rn(__fmul_rn(__fmul_rn(pen, q_coef), l), identity));
      else mapq = (long long)x86_f2i(__fmul_rn(__fmul_rn(pen, q_coef), identity));
    } else {
      if (x >= 0.990f) mapq = (long long)x86_f2i(__fmul_rn(__fmul_rn(__fmul_rn(pen, __fsub_rn(1.0f, x)), y), identity));
      else if (!b.o.bypassClustering) mapq = (long long)x86_f2i(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(pen, q_coef), __fsub_rn(1.0f, x)), l), y), identity));
      else mapq = (long long)x86_f2i(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(pen, q_coef), __fsub_rn(1.0f, x)), y), identity));
      mapq -= (long long)x86_f2i(__fadd_rn(__fmul_rn(4.343f, b.logf_len[na]), .499f));
    }
    mapq = mapq > 0 ? mapq : 0;
    int mq = (int)(mapq < 60 ? mapq : 60);
    if (na == 2 && mq == 0) mq = 1;
    rc.mapq = mq;
    if (!rc.supplementary) ab += (unsigned long long)(rc.qEnd - rc.qStart);
  }
  if (ab) atomicAdd(b.aligned_bases, ab);
}

}  // namespace mp
}  // namespace lra
