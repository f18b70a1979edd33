// Host side of the global index builder (included at the end of lra_b200.cu): lra_b200_gindex_build, lra_b200_index_size / _download.
#pragma once
#include <cub/cub.cuh>
#include "gidx_kernels.cuh"

namespace {
struct GidxTmp {
  std::vector<void *> ptrs;
  ~GidxTmp() { for (void *p : ptrs) if (p) cudaFree(p); }
  template <class T> T *get(size_t n) { void *p = nullptr; if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr; ptrs.push_back(p); return (T *)p; }
  void drop(void *p) { for (auto &q : ptrs) if (q == p) { cudaFree(q); q = nullptr; } }
};
struct GidxWiden { __host__ __device__ unsigned long long operator()(const uint32_t &v) const { return (unsigned long long)v; } };
inline unsigned gidx_blocks(unsigned long long n, int t) { return (unsigned)((n + (unsigned long long)t - 1) / (unsigned long long)t); }
}  // namespace

extern "C" uint64_t lra_b200_index_size(const lra_b200_index *ix) { return ix ? ix->n : 0; }

extern "C" int lra_b200_index_download(lra_b200_ctx *ctx, const lra_b200_index *ix, uint64_t *t, uint32_t *pos) {
  if (!ctx || !ix || (ix->n && (!t || !pos))) return fail(ctx, LRA_B200_EINVAL, "index_download: bad argument");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(t, ix->t, ix->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(pos, ix->pos, ix->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LRA_B200_OK;
}

extern "C" int lra_b200_gindex_build(lra_b200_ctx *ctx, const lra_b200_seq *genome, const uint64_t *contig_start, const uint32_t *contig_len, int32_t n_contigs, int32_t k, int32_t w,
                                     int32_t max_freq, int32_t win_size, int32_t per_window, lra_b200_index **out) {
  using namespace lra;
  if (!ctx || !genome || !contig_start || !contig_len || n_contigs < 1 || !out) return fail(ctx, LRA_B200_EINVAL, "gindex_build: NULL argument");
  if (k < 1 || k > 31 || w < 1 || w > kSeedMaxW || max_freq < 1 || win_size < 1 || per_window < 0) return fail(ctx, LRA_B200_EINVAL, "gindex_build: k in 1..31, w in 1..%d, max_freq, win_size >= 1", kSeedMaxW);
  *out = nullptr;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int CH = 512;                                   // loop steps per chunk
  std::vector<unsigned long long> cfirst(n_contigs + 1, 0), cstart(n_contigs);
  for (int c = 0; c < n_contigs; c++) {
    if (contig_start[c] + contig_len[c] > genome->n || contig_start[c] + contig_len[c] >= (1ull << 32)) return fail(ctx, LRA_B200_EINVAL, "gindex_build: contig %d ends beyond the arena (or beyond 2^32)", c);
    const unsigned long long steps = contig_len[c] >= (uint32_t)k ? (unsigned long long)contig_len[c] - k + 1 : 1ull;
    cfirst[c + 1] = cfirst[c] + (steps + CH - 1) / CH;
    cstart[c] = contig_start[c];
  }
  const unsigned long long NC = cfirst[n_contigs];
  GidxTmp T;
  unsigned long long *d_cstart = T.get<unsigned long long>(n_contigs), *d_cfirst = T.get<unsigned long long>(n_contigs + 1);
  uint32_t *d_clen = T.get<uint32_t>(n_contigs);
  uint32_t *d_warm = T.get<uint32_t>(NC), *d_cnt = T.get<uint32_t>(NC + 1);
  uint8_t *d_unc = T.get<uint8_t>(NC);
  unsigned long long *d_off = T.get<unsigned long long>(NC + 1), *d_list = T.get<unsigned long long>(NC), *d_count = T.get<unsigned long long>(1);
  if (!d_cstart || !d_cfirst || !d_clen || !d_warm || !d_cnt || !d_unc || !d_off || !d_list || !d_count) return fail(ctx, LRA_B200_ECUDA, "gindex_build: allocation failed");
  CU(cudaMemcpyAsync(d_cstart, cstart.data(), (size_t)n_contigs * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_cfirst, cfirst.data(), (size_t)(n_contigs + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_clen, contig_len, (size_t)n_contigs * 4, cudaMemcpyHostToDevice, st));
  gidx_fill_u32_kernel<<<gidx_blocks(NC, 256), 256, 0, st>>>(d_warm, NC, 64u);
  CU(cudaMemsetAsync(d_cnt, 0, (NC + 1) * 4, st));
  GidxScan b;
  b.genome = SeqView{genome->b2, genome->nm, genome->n}; b.contig_start = d_cstart; b.contig_len = d_clen; b.chunk_first = d_cfirst; b.n_contigs = n_contigs; b.k = k; b.w = w;
  b.chunk = CH; b.n_chunks = NC; b.todo = nullptr; b.n_todo = 0; b.warm = d_warm; b.cnt = d_cnt; b.uncertain = d_unc; b.off = nullptr; b.ot = nullptr; b.op = nullptr;
  // ---- count pass (+ longer warm-ups for the chunks whose entry state is not provably the sequential one)
  gidx_scan_kernel<false><<<gidx_blocks(NC, 128), 128, 0, st>>>(b); ctx->launches++;
  const uint32_t retry_warm[2] = {8192u, 0xffffffffu};
  for (int round = 0; round < 2; round++) {
    CU(cudaMemsetAsync(d_count, 0, 8, st));
    gidx_collect_uncertain_kernel<<<gidx_blocks(NC, 256), 256, 0, st>>>(d_unc, NC, d_list, d_count); ctx->launches++;
    unsigned long long nu = 0;
    CU(cudaMemcpyAsync(&nu, d_count, 8, cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st));
    if (nu == 0) break;
    gidx_set_warm_kernel<<<gidx_blocks(nu, 256), 256, 0, st>>>(d_list, nu, d_warm, retry_warm[round]);
    GidxScan b2 = b; b2.todo = d_list; b2.n_todo = nu;
    gidx_scan_kernel<false><<<gidx_blocks(nu, 128), 128, 0, st>>>(b2); ctx->launches += 2;
  }
  CU(cudaGetLastError());
  // ---- exclusive scan, emit pass
  size_t tmp_bytes = 0;
  cub::TransformInputIterator<unsigned long long, GidxWiden, const uint32_t *> cnt_in(d_cnt, GidxWiden());      // counts are 32-bit, offsets 64-bit
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt_in, d_off, (long long)(NC + 1), st);
  void *d_tmp = T.get<unsigned char>(tmp_bytes);
  if (!d_tmp) return fail(ctx, LRA_B200_ECUDA, "gindex_build: allocation failed");
  cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, cnt_in, d_off, (long long)(NC + 1), st); ctx->launches++;
  unsigned long long M = 0;
  CU(cudaMemcpyAsync(&M, d_off + NC, 8, cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st));
  if (M >= 0xfffffff0ull) return fail(ctx, LRA_B200_EOVERFLOW, "gindex_build: %llu minimizers do not fit 32-bit indices", M);
  lra_b200_index *ix = new lra_b200_index();
  if (M == 0) { ix->n = 0; CU(cudaMalloc((void **)&ix->t, 32)); CU(cudaMalloc((void **)&ix->pos, 16)); *out = ix; return LRA_B200_OK; }
  unsigned long long *t_e = T.get<unsigned long long>(M); uint32_t *pos_e = T.get<uint32_t>(M);
  if (!t_e || !pos_e) { delete ix; return fail(ctx, LRA_B200_ECUDA, "gindex_build: allocation failed (%llu minimizers)", M); }
  b.off = d_off; b.ot = t_e; b.op = pos_e;
  gidx_scan_kernel<true><<<gidx_blocks(NC, 128), 128, 0, st>>>(b); ctx->launches++;
  CU(cudaGetLastError());
  // ---- sort by the masked tuple (stable: ties stay in position order)
  unsigned long long *key_a = T.get<unsigned long long>(M), *key_b = T.get<unsigned long long>(M);
  uint32_t *val_a = T.get<uint32_t>(M), *val_b = T.get<uint32_t>(M);
  if (!key_a || !key_b || !val_a || !val_b) { delete ix; return fail(ctx, LRA_B200_ECUDA, "gindex_build: allocation failed (sort buffers)"); }
  gidx_iota_mask_kernel<<<gidx_blocks(M, 256), 256, 0, st>>>(t_e, M, key_a, val_a); ctx->launches++;
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, key_a, key_b, val_a, val_b, (long long)M, 0, 2 * k, st);
  void *d_sort = T.get<unsigned char>(sort_bytes);
  if (!d_sort) { delete ix; return fail(ctx, LRA_B200_ECUDA, "gindex_build: allocation failed (sort scratch)"); }
  cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, key_a, key_b, val_a, val_b, (long long)M, 0, 2 * k, st); ctx->launches++;
  T.drop(d_sort); T.drop(key_a); T.drop(val_a);
  // ---- multiplicities
  uint32_t *run_start = T.get<uint32_t>(M), *run_cnt = T.get<uint32_t>(M), *freq_e = T.get<uint32_t>(M), *sidx_e = T.get<uint32_t>(M);
  uint8_t *keep_e = T.get<uint8_t>(M), *keep_s = T.get<uint8_t>(M);
  if (!run_start || !run_cnt || !freq_e || !sidx_e || !keep_e || !keep_s) { delete ix; return fail(ctx, LRA_B200_ECUDA, "gindex_build: allocation failed (multiplicities)"); }
  gidx_run_start_kernel<<<gidx_blocks(M, 256), 256, 0, st>>>(key_b, M, run_start); ctx->launches++;
  size_t scan_bytes = 0;
  cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, run_start, run_start, cub::Max(), (long long)M, st);
  void *d_scan = T.get<unsigned char>(scan_bytes);
  if (!d_scan) { delete ix; return fail(ctx, LRA_B200_ECUDA, "gindex_build: allocation failed"); }
  cub::DeviceScan::InclusiveScan(d_scan, scan_bytes, run_start, run_start, cub::Max(), (long long)M, st); ctx->launches++;
  CU(cudaMemsetAsync(run_cnt, 0, M * 4, st));
  gidx_run_count_kernel<<<gidx_blocks(M, 256), 256, 0, st>>>(run_start, M, run_cnt);
  gidx_scatter_freq_kernel<<<gidx_blocks(M, 256), 256, 0, st>>>(run_start, run_cnt, val_b, M, freq_e, sidx_e);
  gidx_thin_kernel<<<gidx_blocks(M, 256), 256, 0, st>>>(pos_e, freq_e, sidx_e, M, (uint32_t)max_freq, (uint32_t)win_size, (uint32_t)per_window, keep_e);
  ctx->launches += 3;
  T.drop(run_start); T.drop(run_cnt);
  // ---- RemoveFrequent: compaction in sorted order
  unsigned long long *t_s = T.get<unsigned long long>(M); uint32_t *pos_s = T.get<uint32_t>(M);
  unsigned long long *d_nsel = T.get<unsigned long long>(1);
  if (!t_s || !pos_s || !d_nsel) { delete ix; return fail(ctx, LRA_B200_ECUDA, "gindex_build: allocation failed (compaction)"); }
  gidx_gather_kernel<<<gidx_blocks(M, 256), 256, 0, st>>>(val_b, t_e, pos_e, keep_e, M, t_s, pos_s, keep_s); ctx->launches++;
  CU(cudaMalloc((void **)&ix->t, (M + 4) * 8));
  CU(cudaMalloc((void **)&ix->pos, (M + 4) * 4));
  size_t sel_bytes = 0, sel2 = 0;
  cub::DeviceSelect::Flagged(nullptr, sel_bytes, t_s, keep_s, ix->t, d_nsel, (long long)M, st);
  cub::DeviceSelect::Flagged(nullptr, sel2, pos_s, keep_s, ix->pos, d_nsel, (long long)M, st);
  if (sel2 > sel_bytes) sel_bytes = sel2;
  void *d_sel = T.get<unsigned char>(sel_bytes);
  if (!d_sel) { lra_b200_index_free(ctx, ix); return fail(ctx, LRA_B200_ECUDA, "gindex_build: allocation failed"); }
  cub::DeviceSelect::Flagged(d_sel, sel_bytes, t_s, keep_s, ix->t, d_nsel, (long long)M, st);
  cub::DeviceSelect::Flagged(d_sel, sel_bytes, pos_s, keep_s, ix->pos, d_nsel, (long long)M, st);
  ctx->launches += 2;
  unsigned long long nsel = 0;
  CU(cudaMemcpyAsync(&nsel, d_nsel, 8, cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  ix->n = nsel;
  *out = ix;
  return LRA_B200_OK;
}
