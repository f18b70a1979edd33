// liblra_b200.so -- host side of the C ABI declared in include/lra_b200.h: context, packed-sequence arenas and the
// batched AffineOneGapAlign launcher.  CUDA only: every compute entry point fails loudly without a device.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>

#include "../../include/lra_b200.h"
#include "aog_band_kernel.cuh"
#include "aog_kernels.cuh"
#include "ir_kernels.cuh"
#include "ir_segment_kernels.cuh"
#include "seq_kernels.cuh"
#include "seed_kernels.cuh"
#include "stats_kernels.cuh"
#include "lidx_kernels.cuh"
#include "lref_kernels.cuh"
#include "sort_kernels.cuh"
#include "gchain_kernels.cuh"
#include "rbp_kernels.cuh"
#include "chainf_kernels.cuh"
#include "cod_kernels.cuh"
#include "split_kernels.cuh"
#include "mapq_kernels.cuh"
#include "lext_kernels.cuh"
#include "spchain_kernels.cuh"
#include "cglue_kernels.cuh"
#include "rla_kernels.cuh"
#include "rsp_kernels.cuh"
#include "srough_kernels.cuh"

using namespace lra;

struct lra_b200_seq {
  uint32_t *b2 = nullptr, *nm = nullptr;
  uint8_t *ascii_dev = nullptr;   // staging for uploads (kept for re-use)
  uint64_t n = 0, cap_groups = 0, ascii_cap = 0;
};

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct lra_b200_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::string err;
  int n_sm = 148;
  uint64_t launches = 0;
  // grow-only scratch
  DevBuf plan, bin_of_job, sorted, lit_slab, band_slab, misc;  // misc: block cursor (8 B) + err flag (4 B)
  DevBuf d_qoff, d_toff, d_qlen, d_tlen, d_k, d_score, d_nb, d_boff, d_blocks;
  DevBuf ir_tb, ir_tboff, ir_maxw, ir_in[9], ir_band;
  DevBuf sg[40];          // segment-level IndelRefine scratch
  DevBuf sd[12];          // seeding scratch
  DevBuf stt[12];         // statistics scratch
  DevBuf stt_x[2];        // statistics: block offsets, per-lane op indices
  DevBuf lr[32];          // local index / cluster refinement scratch
  DevBuf li_tmp;          // LocalIndex staging (one slot per arena base)
  DevBuf lr_x[8];         // more cluster refinement scratch
  DevBuf so[8];           // anchor sort scratch
  DevBuf gc[8];           // GlobalChain scratch
  DevBuf rb[20];          // RefineBreakpoint scratch
  DevBuf cf[9];           // chain filter scratch
  DevBuf cd[15];          // CleanOffDiagonal scratch
  DevBuf sc[15];          // SplitClusters scratch
  DevBuf mq[25];          // ordering / MAPQ scratch
  DevBuf le[24];          // linear extension scratch
  DevBuf lc[32];          // linear extension (chain overload) scratch
  DevBuf sp[40];          // chain splitting scratch
  DevBuf cg[16];          // MergeChain / switchindex scratch
  DevBuf rl[8];           // RefineByLinearAlignment: gap descriptors
  DevBuf rs[16];          // RefineSpace: space descriptors, pairs
  DevBuf sr[24];          // SplitRoughClustersWithGaps
  DevBuf rs2[12];         // RefineSpace, minimizer branch
  DevBuf mp[48];          // mapper / SparseDP
  bool keep_stats = false;  // sub-launchers append to stats instead of clearing
  AogPlan *h_plan = nullptr;            // pinned
  unsigned long long *h_misc = nullptr;  // pinned (2 x u64)
  std::vector<lra_b200_kernel_stat> stats;
  std::vector<cudaEvent_t> ev;
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};  // kernel classes of one batch run concurrently
  cudaEvent_t fork_ev = nullptr, join_ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

static std::string g_create_err;

static int fail(lra_b200_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_err = buf;
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(ctx, LRA_B200_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static int ensure(lra_b200_ctx *ctx, DevBuf &b, size_t bytes) {
  if (b.cap >= bytes && b.p) return 0;
  if (b.p) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
  size_t want = bytes + bytes / 4 + 256;
  CU(cudaMalloc(&b.p, want));
  b.cap = want;
  return 0;
}

extern "C" int lra_b200_version(void) { return 100; }

extern "C" int lra_b200_create(lra_b200_ctx **out, int device) {
  lra_b200_ctx *ctx = nullptr;
  if (!out) return fail(nullptr, LRA_B200_EINVAL, "ctx out pointer is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, LRA_B200_ECUDA, "no CUDA device available (%s); lra_b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, LRA_B200_EINVAL, "device %d out of range [0,%d)", device, n);
  ctx = new lra_b200_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return fail(nullptr, LRA_B200_ECUDA, "cudaSetDevice(%d) failed", device); }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->n_sm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx; return fail(nullptr, LRA_B200_ECUDA, "cudaStreamCreate failed");
  }
  ctx->stream = ctx->own_stream;
  if (cudaHostAlloc((void **)&ctx->h_plan, sizeof(AogPlan), cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc((void **)&ctx->h_misc, 64, cudaHostAllocDefault) != cudaSuccess) {
    delete ctx; return fail(nullptr, LRA_B200_ECUDA, "cudaHostAlloc failed");
  }
  ctx->ev.resize(40);
  for (auto &ev : ctx->ev) cudaEventCreate(&ev);
  for (int i = 0; i < 4; i++) {
    cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->join_ev[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
  *out = ctx;
  return LRA_B200_OK;
}

extern "C" void lra_b200_destroy(lra_b200_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  DevBuf *bufs[] = {&ctx->plan, &ctx->bin_of_job, &ctx->sorted, &ctx->lit_slab, &ctx->band_slab, &ctx->misc, &ctx->d_qoff,
                    &ctx->d_toff, &ctx->d_qlen, &ctx->d_tlen, &ctx->d_k, &ctx->d_score, &ctx->d_nb, &ctx->d_boff, &ctx->d_blocks,
                    &ctx->ir_tb, &ctx->ir_tboff, &ctx->ir_maxw, &ctx->ir_band, &ctx->ir_in[0], &ctx->ir_in[1], &ctx->ir_in[2], &ctx->ir_in[3],
                    &ctx->ir_in[4], &ctx->ir_in[5], &ctx->ir_in[6], &ctx->ir_in[7], &ctx->ir_in[8]};
  for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
  for (DevBuf &b : ctx->sg) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->sd) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->stt) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->stt_x) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->lr) if (b.p) cudaFree(b.p);
  if (ctx->li_tmp.p) cudaFree(ctx->li_tmp.p);
  for (DevBuf &b : ctx->lr_x) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->so) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->gc) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->mp) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->rb) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->cf) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->cd) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->sc) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->mq) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->le) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->lc) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->sp) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->cg) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->rl) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->rs) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->sr) if (b.p) cudaFree(b.p);
  for (DevBuf &b : ctx->rs2) if (b.p) cudaFree(b.p);
  for (auto &ev : ctx->ev) cudaEventDestroy(ev);
  for (int i = 0; i < 4; i++) { if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]); if (ctx->join_ev[i]) cudaEventDestroy(ctx->join_ev[i]); }
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  if (ctx->h_plan) cudaFreeHost(ctx->h_plan);
  if (ctx->h_misc) cudaFreeHost(ctx->h_misc);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" const char *lra_b200_last_error(const lra_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int lra_b200_set_stream(lra_b200_ctx *ctx, void *s) {
  if (!ctx) return LRA_B200_EINVAL;
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return LRA_B200_OK;
}
extern "C" int lra_b200_synchronize(lra_b200_ctx *ctx) {
  if (!ctx) return LRA_B200_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return LRA_B200_OK;
}
extern "C" uint64_t lra_b200_launch_count(const lra_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------------- sequences
static int seq_reserve(lra_b200_ctx *ctx, lra_b200_seq *s, uint64_t n) {
  uint64_t groups = (n + 31) / 32 + 1;
  if (groups > s->cap_groups) {
    if (s->b2) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(s->b2)); CU(cudaFree(s->nm)); s->b2 = s->nm = nullptr; }
    uint64_t cap = groups + groups / 4 + 8;
    CU(cudaMalloc((void **)&s->b2, (cap * 2 + 8) * 4));
    CU(cudaMalloc((void **)&s->nm, (cap + 8) * 4));
    CU(cudaMemsetAsync(s->b2, 0, (cap * 2 + 8) * 4, ctx->stream));
    CU(cudaMemsetAsync(s->nm, 0xFF, (cap + 8) * 4, ctx->stream));
    s->cap_groups = cap;
  }
  return 0;
}
static int seq_pack_launch(lra_b200_ctx *ctx, lra_b200_seq *s, const uint8_t *ascii_dev, uint64_t n) {
  int rc = seq_reserve(ctx, s, n);
  if (rc) return rc;
  uint64_t groups = (n + 31) / 32 + 1;  // one extra all-N group as padding
  unsigned blocks = (unsigned)((groups + 255) / 256);
  seq_pack_kernel<<<blocks, 256, 0, ctx->stream>>>(ascii_dev, n, s->b2, s->nm, groups);
  ctx->launches++;
  CU(cudaGetLastError());
  s->n = n;
  return 0;
}

extern "C" int lra_b200_seq_from_device(lra_b200_ctx *ctx, const void *ascii_dev, uint64_t n, lra_b200_seq **out) {
  if (!ctx || !out || (!ascii_dev && n)) return fail(ctx, LRA_B200_EINVAL, "seq_from_device: bad argument");
  CU(cudaSetDevice(ctx->device));
  if (((uintptr_t)ascii_dev) & 15) return fail(ctx, LRA_B200_EINVAL, "seq_from_device: device pointer must be 16-byte aligned");
  lra_b200_seq *s = new lra_b200_seq();
  int rc = seq_pack_launch(ctx, s, (const uint8_t *)ascii_dev, n);
  if (rc) { delete s; return rc; }
  *out = s;
  return LRA_B200_OK;
}

extern "C" int lra_b200_seq_reupload(lra_b200_ctx *ctx, lra_b200_seq *s, const char *ascii_host, uint64_t n) {
  if (!ctx || !s || (!ascii_host && n)) return fail(ctx, LRA_B200_EINVAL, "seq_reupload: bad argument");
  CU(cudaSetDevice(ctx->device));
  if (n + 64 > s->ascii_cap) {
    if (s->ascii_dev) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(s->ascii_dev)); s->ascii_dev = nullptr; }
    uint64_t cap = n + n / 4 + 256;
    CU(cudaMalloc((void **)&s->ascii_dev, cap));
    s->ascii_cap = cap;
  }
  if (n) CU(cudaMemcpyAsync(s->ascii_dev, ascii_host, n, cudaMemcpyHostToDevice, ctx->stream));
  return seq_pack_launch(ctx, s, s->ascii_dev, n);
}

extern "C" int lra_b200_seq_upload(lra_b200_ctx *ctx, const char *ascii_host, uint64_t n, lra_b200_seq **out) {
  if (!ctx || !out) return fail(ctx, LRA_B200_EINVAL, "seq_upload: bad argument");
  lra_b200_seq *s = new lra_b200_seq();
  int rc = lra_b200_seq_reupload(ctx, s, ascii_host, n);
  if (rc) { lra_b200_seq_free(ctx, s); return rc; }
  *out = s;
  return LRA_B200_OK;
}

extern "C" void lra_b200_seq_free(lra_b200_ctx *ctx, lra_b200_seq *s) {
  if (!s) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  if (s->b2) cudaFree(s->b2);
  if (s->nm) cudaFree(s->nm);
  if (s->ascii_dev) cudaFree(s->ascii_dev);
  delete s;
}
extern "C" uint64_t lra_b200_seq_length(const lra_b200_seq *s) { return s ? s->n : 0; }

extern "C" int lra_b200_seq_download(lra_b200_ctx *ctx, const lra_b200_seq *s, uint32_t *b2, uint32_t *nmask) {
  if (!ctx || !s || !b2 || !nmask) return fail(ctx, LRA_B200_EINVAL, "seq_download: bad argument");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(b2, s->b2, ((s->n + 15) / 16) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(nmask, s->nm, ((s->n + 31) / 32) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a18 launcher
static const char *kClsName[kAogNumClasses] = {"aog_thread<K=2>", "aog_thread<K=4>", "aog_thread<K=6>", "aog_thread<K=8>",
                                               "aog_thread<K=10>", "aog_thread<K=12>", "aog_thread<K=14>", "aog_warp_literal",
                                               "aog_warp_band<C=1>", "aog_warp_band<C=2>", "aog_warp_band<C=4>", "aog_warp_band<C=8>"};

template <int K>
static void launch_thread(lra_b200_ctx *ctx, cudaStream_t st, const AogBatch &b, AogPlan *plan, const uint32_t *sorted, uint32_t n) {
  unsigned blocks = (n + 127) / 128;
  unsigned cap = (unsigned)ctx->n_sm * 16u;
  if (blocks > cap) blocks = cap;
  aog_thread_kernel<K><<<blocks, 128, 0, st>>>(b, plan, sorted);
}
template <int C>
static void launch_band(cudaStream_t st, const AogBatch &b, AogPlan *plan, const uint32_t *sorted, unsigned blocks, AogBandScratch sc) {
  aog_warp_band_kernel<C><<<blocks, 128, 0, st>>>(b, plan, sorted, sc);
}

static int aog_run_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_aog_jobs *jobs,
                          lra_b200_aog_result *res) {
  const int n = jobs->n_jobs;
  if (!ctx->keep_stats) ctx->stats.clear();
  res->n_blocks_total = 0;
  res->cells = 0;
  if (n == 0) return LRA_B200_OK;
  int rc;
  if ((rc = ensure(ctx, ctx->plan, sizeof(AogPlan)))) return rc;
  if ((rc = ensure(ctx, ctx->bin_of_job, (size_t)n * 4))) return rc;
  if ((rc = ensure(ctx, ctx->sorted, (size_t)n * 4))) return rc;
  if ((rc = ensure(ctx, ctx->misc, 64))) return rc;
  AogPlan *plan = (AogPlan *)ctx->plan.p;
  unsigned long long *cursor = (unsigned long long *)ctx->misc.p;
  int *errflag = (int *)((char *)ctx->misc.p + 8);
  cudaStream_t st = ctx->stream;
  CU(cudaMemsetAsync(plan, 0, sizeof(AogPlan), st));
  CU(cudaMemsetAsync(ctx->misc.p, 0, 64, st));

  AogBatch b;
  b.q = SeqView{q->b2, q->nm, q->n};
  b.t = SeqView{t->b2, t->nm, t->n};
  b.q_off = jobs->q_off; b.t_off = jobs->t_off; b.q_len = jobs->q_len; b.t_len = jobs->t_len; b.k = jobs->k;
  b.n_jobs = n; b.m = jobs->match; b.mm = jobs->mismatch; b.indel = jobs->indel;
  b.score = res->score; b.n_blocks = res->n_blocks; b.block_off = (unsigned long long *)res->block_off;
  b.blocks = res->blocks; b.block_cap = res->block_cap; b.block_cursor = cursor; b.err = errflag;

  int evi = 0;
  auto rec = [&]() { cudaEventRecord(ctx->ev[evi++], st); };
  const unsigned nb = (unsigned)((n + 255) / 256);
  rec();
  aog_classify_kernel<<<nb, 256, 0, st>>>(b, plan, (uint32_t *)ctx->bin_of_job.p, 1);
  aog_scan_kernel<<<1, 512, 0, st>>>(plan);
  aog_scatter_kernel<<<nb, 256, 0, st>>>(n, plan, (const uint32_t *)ctx->bin_of_job.p, (uint32_t *)ctx->sorted.p);
  ctx->launches += 3;
  rec();
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ctx->h_plan, plan, sizeof(AogPlan), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const AogPlan &hp = *ctx->h_plan;
  uint32_t cnt[kAogNumClasses];
  for (int c = 0; c < kAogNumClasses; c++) cnt[c] = hp.bin_start[(c + 1) * kAogBuckets] - hp.bin_start[c * kAogBuckets];
  const uint32_t *sorted = (const uint32_t *)ctx->sorted.p;

  struct Launched { int cls; int ev0; };
  std::vector<Launched> launched;
  // The classes are independent: latency-bound warp kernels (few long jobs) overlap with the throughput-bound thread
  // kernels on side streams.  LRA_B200_SERIAL=1 keeps everything on the main stream (per-kernel timing without overlap).
  static const bool serial = getenv("LRA_B200_SERIAL") != nullptr;
  cudaStream_t S[4];
  for (int i = 0; i < 4; i++) S[i] = serial ? st : ctx->side[i];
  if (!serial) {
    CU(cudaEventRecord(ctx->fork_ev, st));
    for (int i = 0; i < 4; i++) CU(cudaStreamWaitEvent(S[i], ctx->fork_ev, 0));
  }
  cudaStream_t cur = st;
  auto begin_cls = [&](int c, cudaStream_t s2) { cur = s2; launched.push_back({c, evi}); cudaEventRecord(ctx->ev[evi++], cur); };
  auto end_cls = [&]() { cudaEventRecord(ctx->ev[evi++], cur); ctx->launches++; };

  // stream 0: literal kernel (longest jobs, launched first)
  if (cnt[kAogClsLiteral]) {
    AogLiteralScratch sc;
    sc.max_mat = hp.max_mat; sc.max_diag = hp.max_diag;
    sc.slab_bytes = (aog_literal_slab_bytes(sc.max_mat, sc.max_diag) + 127ull) & ~127ull;
    unsigned blocks = (cnt[kAogClsLiteral] + 3) / 4;
    unsigned max_blocks = (unsigned)ctx->n_sm * 4u;
    if (blocks > max_blocks) blocks = max_blocks;
    while (blocks > 1 && (unsigned long long)blocks * 4ull * sc.slab_bytes > (8ull << 30)) blocks /= 2;
    if ((rc = ensure(ctx, ctx->lit_slab, (size_t)blocks * 4 * sc.slab_bytes))) return rc;
    sc.base = (unsigned char *)ctx->lit_slab.p;
    begin_cls(kAogClsLiteral, S[0]);
    aog_warp_literal_kernel<<<blocks, 128, 0, cur>>>(b, plan, sorted, sc);
    end_cls();
  }
  // stream 1: band classes, widest first
  uint32_t nband = cnt[8] + cnt[9] + cnt[10] + cnt[11];
  if (nband) {
    AogBandScratch sc;
    sc.max_rows = hp.max_rows_band; sc.max_qlen = hp.max_qlen_band;
    sc.slab_bytes = (aog_band_slab_bytes(sc.max_rows, sc.max_qlen) + 127ull) & ~127ull;
    unsigned max_blocks = (unsigned)ctx->n_sm * 4u;
    uint32_t biggest = cnt[8];
    for (int c = 9; c < 12; c++) if (cnt[c] > biggest) biggest = cnt[c];
    unsigned blocks_cap = (biggest + 3) / 4;
    if (blocks_cap > max_blocks) blocks_cap = max_blocks;
    while (blocks_cap > 1 && (unsigned long long)blocks_cap * 4ull * sc.slab_bytes > (8ull << 30)) blocks_cap /= 2;
    // the band classes run one after another on one stream, so they can share the slabs
    if ((rc = ensure(ctx, ctx->band_slab, (size_t)blocks_cap * 4 * sc.slab_bytes))) return rc;
    sc.base = (unsigned char *)ctx->band_slab.p;
    auto blocks_for = [&](uint32_t c) { unsigned x = (c + 3) / 4; return x > blocks_cap ? blocks_cap : x; };
    if (cnt[11]) { begin_cls(11, S[1]); launch_band<8>(cur, b, plan, sorted, blocks_for(cnt[11]), sc); end_cls(); }
    if (cnt[10]) { begin_cls(10, S[1]); launch_band<4>(cur, b, plan, sorted, blocks_for(cnt[10]), sc); end_cls(); }
    if (cnt[9]) { begin_cls(9, S[1]); launch_band<2>(cur, b, plan, sorted, blocks_for(cnt[9]), sc); end_cls(); }
    if (cnt[8]) { begin_cls(8, S[1]); launch_band<1>(cur, b, plan, sorted, blocks_for(cnt[8]), sc); end_cls(); }
  }
  // streams 2,3: thread-per-job classes
  if (cnt[6]) { begin_cls(6, S[2]); launch_thread<14>(ctx, cur, b, plan, sorted, cnt[6]); end_cls(); }
  if (cnt[4]) { begin_cls(4, S[3]); launch_thread<10>(ctx, cur, b, plan, sorted, cnt[4]); end_cls(); }
  if (cnt[2]) { begin_cls(2, S[2]); launch_thread<6>(ctx, cur, b, plan, sorted, cnt[2]); end_cls(); }
  if (cnt[0]) { begin_cls(0, S[3]); launch_thread<2>(ctx, cur, b, plan, sorted, cnt[0]); end_cls(); }
  if (cnt[5]) { begin_cls(5, S[2]); launch_thread<12>(ctx, cur, b, plan, sorted, cnt[5]); end_cls(); }
  if (cnt[3]) { begin_cls(3, S[3]); launch_thread<8>(ctx, cur, b, plan, sorted, cnt[3]); end_cls(); }
  if (cnt[1]) { begin_cls(1, S[2]); launch_thread<4>(ctx, cur, b, plan, sorted, cnt[1]); end_cls(); }
  if (!serial) {
    for (int i = 0; i < 4; i++) { CU(cudaEventRecord(ctx->join_ev[i], S[i])); CU(cudaStreamWaitEvent(st, ctx->join_ev[i], 0)); }
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ctx->h_plan, plan, sizeof(AogPlan), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(ctx->h_misc, ctx->misc.p, 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  res->n_blocks_total = ctx->h_misc[0];
  res->cells = ctx->h_plan->cells;
  const int err = *(int *)((char *)ctx->h_misc + 8);
  // stats
  {
    lra_b200_kernel_stat s;
    memset(&s, 0, sizeof s);
    snprintf(s.name, sizeof s.name, "aog_plan(classify+scan+scatter)");
    cudaEventElapsedTime(&s.ms, ctx->ev[0], ctx->ev[1]);
    s.jobs = (uint64_t)n;
    s.algo_bytes = (uint64_t)n * (20 + 4 + 4);
    ctx->stats.push_back(s);
    for (auto &L : launched) {
      memset(&s, 0, sizeof s);
      snprintf(s.name, sizeof s.name, "%s", kClsName[L.cls]);
      cudaEventElapsedTime(&s.ms, ctx->ev[L.ev0], ctx->ev[L.ev0 + 1]);
      s.jobs = cnt[L.cls];
      s.cells = ctx->h_plan->cls_cells[L.cls];
      s.algo_bytes = ctx->h_plan->cls_bytes[L.cls] + 12ull * ctx->h_plan->cls_blocks[L.cls];
      ctx->stats.push_back(s);
    }
  }
  if (err & 8) return fail(ctx, LRA_B200_EINVAL, "aog_batch: at least one job has a negative length or a window outside its arena");
  if (err & 1) return fail(ctx, LRA_B200_EOVERFLOW, "aog_batch: block capacity %llu too small, %llu needed",
                           (unsigned long long)res->block_cap, (unsigned long long)res->n_blocks_total);
  if (err & 6) return fail(ctx, LRA_B200_EINTERNAL, "aog_batch: kernel self-check failed (flags 0x%x)", err);
  return LRA_B200_OK;
}

extern "C" int lra_b200_aog_batch_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t,
                                         const lra_b200_aog_jobs *jobs, lra_b200_aog_result *res) {
  if (!ctx || !q || !t || !jobs || !res) return fail(ctx, LRA_B200_EINVAL, "aog_batch_device: NULL argument");
  if (jobs->n_jobs < 0) return fail(ctx, LRA_B200_EINVAL, "aog_batch_device: negative job count");
  CU(cudaSetDevice(ctx->device));
  return aog_run_device(ctx, q, t, jobs, res);
}

extern "C" int lra_b200_aog_batch(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_aog_jobs *jobs,
                                  lra_b200_aog_result *res) {
  if (!ctx || !q || !t || !jobs || !res) return fail(ctx, LRA_B200_EINVAL, "aog_batch: NULL argument");
  const int n = jobs->n_jobs;
  if (n < 0) return fail(ctx, LRA_B200_EINVAL, "aog_batch: negative job count");
  CU(cudaSetDevice(ctx->device));
  if (n == 0) { res->n_blocks_total = 0; res->cells = 0; ctx->stats.clear(); return LRA_B200_OK; }
  int rc;
  const size_t nb4 = (size_t)n * 4;
  if ((rc = ensure(ctx, ctx->d_qoff, nb4)) || (rc = ensure(ctx, ctx->d_toff, nb4)) || (rc = ensure(ctx, ctx->d_qlen, nb4)) ||
      (rc = ensure(ctx, ctx->d_tlen, nb4)) || (rc = ensure(ctx, ctx->d_k, nb4)) || (rc = ensure(ctx, ctx->d_score, nb4)) ||
      (rc = ensure(ctx, ctx->d_nb, nb4)) || (rc = ensure(ctx, ctx->d_boff, (size_t)n * 8)) ||
      (rc = ensure(ctx, ctx->d_blocks, (size_t)(res->block_cap ? res->block_cap : 1) * 12)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(ctx->d_qoff.p, jobs->q_off, nb4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_toff.p, jobs->t_off, nb4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_qlen.p, jobs->q_len, nb4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_tlen.p, jobs->t_len, nb4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_k.p, jobs->k, nb4, cudaMemcpyHostToDevice, st));
  lra_b200_aog_jobs dj = *jobs;
  dj.q_off = (const uint32_t *)ctx->d_qoff.p; dj.t_off = (const uint32_t *)ctx->d_toff.p;
  dj.q_len = (const int32_t *)ctx->d_qlen.p; dj.t_len = (const int32_t *)ctx->d_tlen.p; dj.k = (const int32_t *)ctx->d_k.p;
  lra_b200_aog_result dr = *res;
  dr.score = (int32_t *)ctx->d_score.p; dr.n_blocks = (int32_t *)ctx->d_nb.p; dr.block_off = (uint64_t *)ctx->d_boff.p;
  dr.blocks = (uint32_t *)ctx->d_blocks.p;
  rc = aog_run_device(ctx, q, t, &dj, &dr);
  res->n_blocks_total = dr.n_blocks_total;
  res->cells = dr.cells;
  if (rc != LRA_B200_OK && rc != LRA_B200_EOVERFLOW) return rc;
  CU(cudaMemcpyAsync(res->score, dr.score, nb4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->n_blocks, dr.n_blocks, nb4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->block_off, dr.block_off, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (rc == LRA_B200_OK && dr.n_blocks_total)
    CU(cudaMemcpyAsync(res->blocks, dr.blocks, (size_t)dr.n_blocks_total * 12, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return rc;
}

extern "C" int lra_b200_last_kernel_stats(lra_b200_ctx *ctx, lra_b200_kernel_stat *out, int cap) {
  if (!ctx) return 0;
  int n = (int)ctx->stats.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = ctx->stats[i];
  return n;
}


// ---------------------------------------------------------------------------------------------------- a19 launcher
static int ir_run_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_ir_groups *gr,
                         lra_b200_ir_result *res) {
  const int n = gr->n_groups;
  if (!ctx->keep_stats) ctx->stats.clear();
  res->n_blocks_total = 0;
  res->cells = 0;
  if (n == 0) return LRA_B200_OK;
  int rc;
  if ((rc = ensure(ctx, ctx->plan, sizeof(AogPlan))) || (rc = ensure(ctx, ctx->bin_of_job, (size_t)n * 4)) ||
      (rc = ensure(ctx, ctx->sorted, (size_t)n * 4)) || (rc = ensure(ctx, ctx->misc, 64)) ||
      (rc = ensure(ctx, ctx->ir_tboff, (size_t)n * 8)) || (rc = ensure(ctx, ctx->ir_maxw, (size_t)n * 4)))
    return rc;
  AogPlan *plan = (AogPlan *)ctx->plan.p;
  unsigned long long *cursor = (unsigned long long *)ctx->misc.p;          // [0] block cursor
  int *errflag = (int *)((char *)ctx->misc.p + 8);
  unsigned long long *tb_cursor = (unsigned long long *)((char *)ctx->misc.p + 16);
  unsigned long long *cells_total = (unsigned long long *)((char *)ctx->misc.p + 24);
  cudaStream_t st = ctx->stream;
  CU(cudaMemsetAsync(plan, 0, sizeof(AogPlan), st));
  CU(cudaMemsetAsync(ctx->misc.p, 0, 64, st));
  IrBatch b;
  b.q = SeqView{q->b2, q->nm, q->n};
  b.t = SeqView{t->b2, t->nm, t->n};
  b.q_base = gr->q_base; b.t_base = gr->t_base; b.q_start = gr->q_start; b.t_start = gr->t_start; b.t_len = gr->t_len;
  b.q_seq_len = gr->q_seq_len; b.t_seq_len = gr->t_seq_len; b.band_off = gr->band_off; b.band = gr->band;
  b.n_groups = n; b.match = gr->match; b.mismatch = gr->mismatch; b.gap = gr->indel;
  b.n_blocks = res->n_blocks; b.block_off = (unsigned long long *)res->block_off; b.blocks = res->blocks;
  b.block_cap = res->block_cap; b.block_cursor = cursor; b.err = errflag;
  b.tb = nullptr; b.tb_off = (unsigned long long *)ctx->ir_tboff.p; b.max_width = (int32_t *)ctx->ir_maxw.p;
  int evi = 0;
  auto rec = [&]() { cudaEventRecord(ctx->ev[evi++], st); };
  rec();
  static const int no_warp = getenv("LRA_B200_IR_NO_WARP") ? 1 : getenv("LRA_B200_IR_SCAN_KERNEL") ? 2 : 0;
  // measured on B200 (tools/ir_sweep.sh): a small batch is bound by the row chain of its longest groups, which the scan kernel
  // walks at half the latency per row; a large batch is bound by throughput, where the row-pipeline kernel is ahead
  static const int long_rows_env = getenv("LRA_B200_IR_LONG_ROWS") ? atoi(getenv("LRA_B200_IR_LONG_ROWS")) : 0;
  const int long_rows = long_rows_env > 0 ? long_rows_env : (n < 40000 ? 6000 : 24576);
  ir_classify_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(b, plan, (uint32_t *)ctx->bin_of_job.p, tb_cursor, cells_total, no_warp, long_rows);
  aog_scan_kernel<<<1, 512, 0, st>>>(plan);
  aog_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, plan, (const uint32_t *)ctx->bin_of_job.p, (uint32_t *)ctx->sorted.p);
  ctx->launches += 3;
  rec();
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ctx->h_plan, plan, sizeof(AogPlan), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(ctx->h_misc, ctx->misc.p, 32, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (*(int *)((char *)ctx->h_misc + 8) & 8)
    return fail(ctx, LRA_B200_EINVAL, "indel_dp_batch: a group has fewer than 2 rows, a row narrower than 2 cells or a non-monotone band");
  const AogPlan &hp = *ctx->h_plan;
  uint32_t cnt[kIrNumCls];
  for (int c = 0; c < kIrNumCls; c++) cnt[c] = hp.bin_start[(c + 1) * kAogBuckets] - hp.bin_start[c * kAogBuckets];
  const unsigned long long tb_words = ctx->h_misc[2];
  if ((rc = ensure(ctx, ctx->ir_tb, (size_t)(tb_words + 64) * 4))) return rc;
  b.tb = (uint32_t *)ctx->ir_tb.p;
  const uint32_t *sorted = (const uint32_t *)ctx->sorted.p;
  struct Launched { int cls; int ev0; };
  std::vector<Launched> launched;
  auto blocks_for = [&](uint32_t c) { unsigned x = (c + 63) / 64; unsigned cap = (unsigned)ctx->n_sm * 8u; return x > cap ? cap : x; };
  // the DP classes are independent: each runs on its own side stream (LRA_B200_SERIAL=1 keeps them on the main stream)
  static const bool ir_serial = getenv("LRA_B200_SERIAL") != nullptr;
  cudaStream_t S[4];
  for (int i = 0; i < 4; i++) S[i] = ir_serial ? st : ctx->side[i];
  if (!ir_serial) {
    CU(cudaEventRecord(ctx->fork_ev, st));
    for (int i = 0; i < 4; i++) CU(cudaStreamWaitEvent(S[i], ctx->fork_ev, 0));
  }
  auto recs = [&](cudaStream_t s) { cudaEventRecord(ctx->ev[evi++], s); };
  if (cnt[kIrClsWarp32]) {   // the longest groups first: their row chains are the critical path of the batch
    unsigned wb = (cnt[kIrClsWarp32] + 3) / 4; const unsigned wcap = (unsigned)ctx->n_sm * 16u; if (wb > wcap) wb = wcap;
    launched.push_back({kIrClsWarp32, evi}); recs(S[0]); ir_dp_warp_kernel<<<wb, 128, 0, S[0]>>>(b, plan, sorted); recs(S[0]); ctx->launches++;
  }
  if (cnt[kIrClsPipe]) {
    unsigned pb = (cnt[kIrClsPipe] + 15) / 16; const unsigned pcap = (unsigned)ctx->n_sm * 5u; if (pb > pcap) pb = pcap;
    launched.push_back({kIrClsPipe, evi}); recs(S[1]); { static const bool two = getenv("LRA_B200_IR_PIPE2") != nullptr;
      if (two) ir_dp_pipe_kernel<2><<<pb, 128, 0, S[1]>>>(b, plan, sorted, kIrClsPipe); else ir_dp_pipe_kernel<1><<<pb, 128, 0, S[1]>>>(b, plan, sorted, kIrClsPipe); } recs(S[1]); ctx->launches++;
  }
  if (cnt[kIrClsWarp64]) {
    unsigned wb = (cnt[kIrClsWarp64] + 3) / 4; const unsigned wcap = (unsigned)ctx->n_sm * 16u; if (wb > wcap) wb = wcap;
    launched.push_back({kIrClsWarp64, evi}); recs(S[0]); ir_dp_warp64_kernel<<<wb, 128, 0, S[0]>>>(b, plan, sorted); recs(S[0]); ctx->launches++;
  }
  if (cnt[kIrClsGeneric]) { launched.push_back({kIrClsGeneric, evi}); recs(S[2]); ir_dp_generic_kernel<<<blocks_for(cnt[kIrClsGeneric]), 64, 0, S[2]>>>(b, plan, sorted); recs(S[2]); ctx->launches++; }
  if (cnt[kIrClsW64]) { launched.push_back({kIrClsW64, evi}); recs(S[3]); ir_dp_thread_kernel<64><<<blocks_for(cnt[kIrClsW64]), 64, 0, S[3]>>>(b, plan, sorted, kIrClsW64); recs(S[3]); ctx->launches++; }
  if (cnt[kIrClsW24]) { launched.push_back({kIrClsW24, evi}); recs(S[2]); ir_dp_thread_kernel<24><<<blocks_for(cnt[kIrClsW24]), 64, 0, S[2]>>>(b, plan, sorted, kIrClsW24); recs(S[2]); ctx->launches++; }
  if (!ir_serial)
    for (int i = 0; i < 4; i++) { CU(cudaEventRecord(ctx->join_ev[i], S[i])); CU(cudaStreamWaitEvent(st, ctx->join_ev[i], 0)); }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ctx->h_plan, plan, sizeof(AogPlan), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(ctx->h_misc, ctx->misc.p, 32, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  res->n_blocks_total = ctx->h_misc[0];
  res->cells = ctx->h_misc[3];
  const int err = *(int *)((char *)ctx->h_misc + 8);
  static const char *names[kIrNumCls] = {"ir_dp_thread<W=24>", "ir_dp_thread<W=64>", "ir_dp_generic", "ir_dp_warp<W=32>", "ir_dp_pipe<W=32>", "ir_dp_warp<W=64>"};
  {
    lra_b200_kernel_stat s;
    memset(&s, 0, sizeof s);
    snprintf(s.name, sizeof s.name, "ir_plan(classify+scan+scatter)");
    cudaEventElapsedTime(&s.ms, ctx->ev[0], ctx->ev[1]);
    s.jobs = (uint64_t)n;
    ctx->stats.push_back(s);
    for (auto &L : launched) {
      memset(&s, 0, sizeof s);
      snprintf(s.name, sizeof s.name, "%s", names[L.cls]);
      cudaEventElapsedTime(&s.ms, ctx->ev[L.ev0], ctx->ev[L.ev0 + 1]);
      s.jobs = cnt[L.cls];
      s.cells = ctx->h_plan->cls_cells[L.cls];
      s.algo_bytes = ctx->h_plan->cls_bytes[L.cls] + 12ull * ctx->h_plan->cls_blocks[L.cls];
      ctx->stats.push_back(s);
    }
  }
  if (err & 1) return fail(ctx, LRA_B200_EOVERFLOW, "indel_dp_batch: block capacity %llu too small, %llu needed",
                           (unsigned long long)res->block_cap, (unsigned long long)res->n_blocks_total);
  if (err & ~1) return fail(ctx, LRA_B200_EINTERNAL, "indel_dp_batch: kernel self-check failed (flags 0x%x)", err);
  return LRA_B200_OK;
}

extern "C" int lra_b200_indel_dp_batch_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t,
                                              const lra_b200_ir_groups *gr, lra_b200_ir_result *res) {
  if (!ctx || !q || !t || !gr || !res) return fail(ctx, LRA_B200_EINVAL, "indel_dp_batch_device: NULL argument");
  if (gr->n_groups < 0) return fail(ctx, LRA_B200_EINVAL, "indel_dp_batch_device: negative group count");
  CU(cudaSetDevice(ctx->device));
  return ir_run_device(ctx, q, t, gr, res);
}

extern "C" int lra_b200_indel_dp_batch(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_ir_groups *gr,
                                       lra_b200_ir_result *res) {
  if (!ctx || !q || !t || !gr || !res) return fail(ctx, LRA_B200_EINVAL, "indel_dp_batch: NULL argument");
  const int n = gr->n_groups;
  if (n < 0) return fail(ctx, LRA_B200_EINVAL, "indel_dp_batch: negative group count");
  CU(cudaSetDevice(ctx->device));
  if (n == 0) { res->n_blocks_total = 0; res->cells = 0; ctx->stats.clear(); return LRA_B200_OK; }
  int rc;
  const size_t nb4 = (size_t)n * 4;
  for (int i = 0; i < 8; i++) if ((rc = ensure(ctx, ctx->ir_in[i], nb4))) return rc;
  if ((rc = ensure(ctx, ctx->ir_band, (size_t)gr->band_len * 4 + 16)) || (rc = ensure(ctx, ctx->d_nb, nb4)) ||
      (rc = ensure(ctx, ctx->d_boff, (size_t)n * 8)) || (rc = ensure(ctx, ctx->d_blocks, (size_t)(res->block_cap ? res->block_cap : 1) * 12)))
    return rc;
  cudaStream_t st = ctx->stream;
  const void *src[8] = {gr->q_base, gr->t_base, gr->q_start, gr->t_start, gr->t_len, gr->q_seq_len, gr->t_seq_len, gr->band_off};
  for (int i = 0; i < 8; i++) CU(cudaMemcpyAsync(ctx->ir_in[i].p, src[i], nb4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->ir_band.p, gr->band, (size_t)gr->band_len * 4, cudaMemcpyHostToDevice, st));
  lra_b200_ir_groups dg = *gr;
  dg.q_base = (const uint32_t *)ctx->ir_in[0].p; dg.t_base = (const uint32_t *)ctx->ir_in[1].p;
  dg.q_start = (const int32_t *)ctx->ir_in[2].p; dg.t_start = (const int32_t *)ctx->ir_in[3].p;
  dg.t_len = (const int32_t *)ctx->ir_in[4].p; dg.q_seq_len = (const int32_t *)ctx->ir_in[5].p;
  dg.t_seq_len = (const int32_t *)ctx->ir_in[6].p; dg.band_off = (const uint32_t *)ctx->ir_in[7].p;
  dg.band = (const int32_t *)ctx->ir_band.p;
  lra_b200_ir_result dr = *res;
  dr.n_blocks = (int32_t *)ctx->d_nb.p; dr.block_off = (uint64_t *)ctx->d_boff.p; dr.blocks = (uint32_t *)ctx->d_blocks.p;
  rc = ir_run_device(ctx, q, t, &dg, &dr);
  res->n_blocks_total = dr.n_blocks_total;
  res->cells = dr.cells;
  if (rc != LRA_B200_OK && rc != LRA_B200_EOVERFLOW) return rc;
  CU(cudaMemcpyAsync(res->n_blocks, dr.n_blocks, nb4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->block_off, dr.block_off, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (rc == LRA_B200_OK && dr.n_blocks_total)
    CU(cudaMemcpyAsync(res->blocks, dr.blocks, (size_t)dr.n_blocks_total * 12, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return rc;
}


// ---------------------------------------------------------------------------------------------------- a19 whole function
static int ir_segments_run_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_ir_segments *sg,
                                  lra_b200_ir_seg_result *res) {
  const int S = sg->n_segments;
  const size_t T = (size_t)sg->n_blocks_in;
  ctx->stats.clear();
  res->n_blocks_total = 0; res->cells = 0; res->n_dp_groups = 0; res->n_aog_jobs = 0;
  if (S == 0) return LRA_B200_OK;
  if (sg->refine_band < 2) return fail(ctx, LRA_B200_EINVAL, "indel_refine_batch: refine_band must be >= 2");
  int rc;
  const size_t nd = T + 2 * (size_t)S + 8;   // capacity of job / group descriptor arrays
  enum { WORK, PIECES, NPIECES, AQ, AT, AQL, ATL, AK, GQB, GTB, GQS, GTS, GTL, GQSL, GTSL, GBO, GSEG, GFB, GLB, GFIRST, GLAST, CNT,
         A_SCORE, A_NB, A_BOFF, A_BLK, D_NB, D_BOFF, D_BLK, BAND, OUTCUR };
  DevBuf *B = ctx->sg;
  if ((rc = ensure(ctx, B[WORK], (T + 2 * (size_t)S + 4) * 12)) || (rc = ensure(ctx, B[PIECES], (2 * T + 8 * (size_t)S + 8) * 16)) ||
      (rc = ensure(ctx, B[NPIECES], (size_t)S * 4)) || (rc = ensure(ctx, B[CNT], 64)) || (rc = ensure(ctx, B[OUTCUR], 64)))
    return rc;
  for (int i : {AQ, AT, AQL, ATL, AK, GQB, GTB, GQS, GTS, GTL, GQSL, GTSL, GBO, GSEG, GFB, GLB}) if ((rc = ensure(ctx, B[i], nd * 4))) return rc;
  if ((rc = ensure(ctx, B[GFIRST], nd * 12)) || (rc = ensure(ctx, B[GLAST], nd * 12))) return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemsetAsync(B[CNT].p, 0, 64, st));
  CU(cudaMemsetAsync(B[OUTCUR].p, 0, 64, st));
  IrSegBatch b;
  b.blocks_in = sg->blocks_in; b.blk_off = (const unsigned long long *)sg->blk_off; b.blk_cnt = sg->blk_cnt;
  b.q_base = sg->q_base; b.t_base = sg->t_base; b.read_len = sg->read_len; b.contig_len = sg->contig_len;
  b.n_seg = S; b.k = sg->refine_band; b.end_align = sg->end_align;
  b.work = (uint32_t *)B[WORK].p; b.pieces = (uint32_t *)B[PIECES].p; b.n_pieces = (int32_t *)B[NPIECES].p;
  b.aog_q_off = (uint32_t *)B[AQ].p; b.aog_t_off = (uint32_t *)B[AT].p; b.aog_q_len = (int32_t *)B[AQL].p;
  b.aog_t_len = (int32_t *)B[ATL].p; b.aog_k = (int32_t *)B[AK].p;
  b.g_q_base = (uint32_t *)B[GQB].p; b.g_t_base = (uint32_t *)B[GTB].p; b.g_q_start = (int32_t *)B[GQS].p;
  b.g_t_start = (int32_t *)B[GTS].p; b.g_t_len = (int32_t *)B[GTL].p; b.g_q_seq_len = (int32_t *)B[GQSL].p;
  b.g_t_seq_len = (int32_t *)B[GTSL].p; b.g_band_off = (uint32_t *)B[GBO].p; b.g_seg = (int32_t *)B[GSEG].p;
  b.g_first_block = (int32_t *)B[GFB].p; b.g_last_block = (int32_t *)B[GLB].p; b.g_first = (uint32_t *)B[GFIRST].p;
  b.g_last = (uint32_t *)B[GLAST].p; b.counters = (unsigned long long *)B[CNT].p;

  cudaEvent_t e0 = ctx->ev[36], e1 = ctx->ev[37], e2 = ctx->ev[38], e3 = ctx->ev[39];
  cudaEventRecord(e0, st);
  ir_group_kernel<<<(unsigned)((S + 3) / 4), 128, 0, st>>>(b);
  ctx->launches++;
  cudaEventRecord(e1, st);
  CU(cudaGetLastError());
  unsigned long long hc[8];
  CU(cudaMemcpyAsync(ctx->h_misc, B[CNT].p, 32, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  memcpy(hc, ctx->h_misc, 32);
  const int nA = (int)hc[0], nG = (int)hc[1];
  const unsigned long long bandInts = hc[2];
  if (bandInts > 0xFFFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "indel_refine_batch: batch too large (band offsets exceed 32 bits); split it");
  res->n_dp_groups = (uint64_t)nG; res->n_aog_jobs = (uint64_t)nA;
  std::vector<lra_b200_kernel_stat> all;
  auto stat = [&](const char *name, cudaEvent_t a, cudaEvent_t c, uint64_t jobs) {
    lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "%s", name);
    cudaEventElapsedTime(&s2.ms, a, c); s2.jobs = jobs; all.push_back(s2);
  };
  stat("ir_group", e0, e1, (uint64_t)S);
  // ---- band construction for the DP groups
  if (nG) {
    if ((rc = ensure(ctx, B[BAND], (size_t)(2 * bandInts + 32) * 4))) return rc;
    static const bool band_literal = getenv("LRA_B200_IR_BAND_LITERAL") != nullptr;
    cudaEventRecord(e2, st);
    if (band_literal) ir_band_literal_kernel<<<(unsigned)((nG + 3) / 4), 128, 0, st>>>(b, nG, (int32_t *)B[BAND].p);
    else ir_band_kernel<<<(unsigned)((nG + 3) / 4), 128, 0, st>>>(b, nG, (int32_t *)B[BAND].p, (int32_t *)B[BAND].p + bandInts + 16);
    ctx->launches++;
    cudaEventRecord(e3, st);
    CU(cudaGetLastError());
  }
  // ---- small windows: AffineOneGapAlign
  lra_b200_aog_result ar; memset(&ar, 0, sizeof ar);
  if (nA) {
    const size_t acap = (size_t)nA * (size_t)sg->refine_band + 16;
    if ((rc = ensure(ctx, B[A_SCORE], (size_t)nA * 4)) || (rc = ensure(ctx, B[A_NB], (size_t)nA * 4)) ||
        (rc = ensure(ctx, B[A_BOFF], (size_t)nA * 8)) || (rc = ensure(ctx, B[A_BLK], acap * 12)))
      return rc;
    lra_b200_aog_jobs aj = {b.aog_q_off, b.aog_t_off, b.aog_q_len, b.aog_t_len, b.aog_k, nA, sg->match, sg->mismatch, sg->indel};
    ar.score = (int32_t *)B[A_SCORE].p; ar.n_blocks = (int32_t *)B[A_NB].p; ar.block_off = (uint64_t *)B[A_BOFF].p;
    ar.blocks = (uint32_t *)B[A_BLK].p; ar.block_cap = acap;
    rc = aog_run_device(ctx, q, t, &aj, &ar);
    for (auto &s2 : ctx->stats) all.push_back(s2);
    if (rc) return rc;
  }
  // ---- banded DP
  lra_b200_ir_result dr; memset(&dr, 0, sizeof dr);
  if (nG) {
    if ((rc = ensure(ctx, B[D_NB], (size_t)nG * 4)) || (rc = ensure(ctx, B[D_BOFF], (size_t)nG * 8))) return rc;
    lra_b200_ir_groups dg;
    dg.q_base = b.g_q_base; dg.t_base = b.g_t_base; dg.q_start = b.g_q_start; dg.t_start = b.g_t_start; dg.t_len = b.g_t_len;
    dg.q_seq_len = b.g_q_seq_len; dg.t_seq_len = b.g_t_seq_len; dg.band_off = b.g_band_off; dg.band = (const int32_t *)B[BAND].p;
    dg.band_len = bandInts; dg.n_groups = nG; dg.match = sg->match; dg.mismatch = sg->mismatch; dg.indel = sg->indel;
    size_t dcap = 2 * T + 64 * (size_t)nG + 1024;
    for (int attempt = 0; attempt < 2; attempt++) {
      if ((rc = ensure(ctx, B[D_BLK], dcap * 12))) return rc;
      dr.n_blocks = (int32_t *)B[D_NB].p; dr.block_off = (uint64_t *)B[D_BOFF].p; dr.blocks = (uint32_t *)B[D_BLK].p; dr.block_cap = dcap;
      rc = ir_run_device(ctx, q, t, &dg, &dr);
      if (rc != LRA_B200_EOVERFLOW) break;
      dcap = (size_t)dr.n_blocks_total + 64;
    }
    {
      lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "ir_band");
      cudaEventElapsedTime(&s2.ms, e2, e3); s2.jobs = (uint64_t)nG; s2.algo_bytes = bandInts * 4; all.push_back(s2);
    }
    for (auto &s2 : ctx->stats) all.push_back(s2);
    if (rc) return rc;
    res->cells = dr.cells;
  }
  // ---- assembly
  IrAssemble a;
  a.aog_n_blocks = ar.n_blocks; a.aog_block_off = (const unsigned long long *)ar.block_off; a.aog_blocks = ar.blocks;
  a.dp_n_blocks = dr.n_blocks; a.dp_block_off = (const unsigned long long *)dr.block_off; a.dp_blocks = dr.blocks;
  a.out_n = res->n_blocks; a.out_off = (unsigned long long *)res->block_off; a.out_blocks = res->blocks; a.out_cap = res->block_cap;
  a.out_cursor = (unsigned long long *)B[OUTCUR].p; a.err = (int *)((char *)B[OUTCUR].p + 8);
  cudaEventRecord(e0, st);
  ir_assemble_kernel<<<(unsigned)((S + 3) / 4), 128, 0, st>>>(b, a);
  ctx->launches++;
  cudaEventRecord(e1, st);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ctx->h_misc, B[OUTCUR].p, 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  stat("ir_assemble", e0, e1, (uint64_t)S);
  ctx->stats = all;
  res->n_blocks_total = ctx->h_misc[0];
  const int err = *(int *)((char *)ctx->h_misc + 8);
  if (err & 1) return fail(ctx, LRA_B200_EOVERFLOW, "indel_refine_batch: block capacity %llu too small, %llu needed",
                           (unsigned long long)res->block_cap, (unsigned long long)res->n_blocks_total);
  if (err & 64) return fail(ctx, LRA_B200_EINTERNAL, "indel_refine_batch: refined blocks overlap (the reference prints 'ERROR with alignment consistency')");
  return LRA_B200_OK;
}

extern "C" int lra_b200_indel_refine_batch_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t,
                                                  const lra_b200_ir_segments *sg, lra_b200_ir_seg_result *res) {
  if (!ctx || !q || !t || !sg || !res) return fail(ctx, LRA_B200_EINVAL, "indel_refine_batch_device: NULL argument");
  if (sg->n_segments < 0) return fail(ctx, LRA_B200_EINVAL, "indel_refine_batch_device: negative segment count");
  CU(cudaSetDevice(ctx->device));
  return ir_segments_run_device(ctx, q, t, sg, res);
}

extern "C" int lra_b200_indel_refine_batch(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t,
                                           const lra_b200_ir_segments *sg, lra_b200_ir_seg_result *res) {
  if (!ctx || !q || !t || !sg || !res) return fail(ctx, LRA_B200_EINVAL, "indel_refine_batch: NULL argument");
  const int S = sg->n_segments;
  if (S < 0) return fail(ctx, LRA_B200_EINVAL, "indel_refine_batch: negative segment count");
  CU(cudaSetDevice(ctx->device));
  if (S == 0) { res->n_blocks_total = 0; res->cells = 0; res->n_dp_groups = 0; res->n_aog_jobs = 0; ctx->stats.clear(); return LRA_B200_OK; }
  int rc;
  DevBuf *H = ctx->sg + 31;  // host-variant staging: 31..39
  const size_t T = (size_t)sg->n_blocks_in;
  if ((rc = ensure(ctx, H[0], T * 12 + 16)) || (rc = ensure(ctx, H[1], (size_t)S * 8)) || (rc = ensure(ctx, H[2], (size_t)S * 4)) ||
      (rc = ensure(ctx, H[3], (size_t)S * 4)) || (rc = ensure(ctx, H[4], (size_t)S * 4)) || (rc = ensure(ctx, H[5], (size_t)S * 4)) ||
      (rc = ensure(ctx, H[6], (size_t)S * 4)) || (rc = ensure(ctx, H[7], (size_t)S * 12)) ||
      (rc = ensure(ctx, H[8], (size_t)(res->block_cap ? res->block_cap : 1) * 12)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(H[0].p, sg->blocks_in, T * 12, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(H[1].p, sg->blk_off, (size_t)S * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(H[2].p, sg->blk_cnt, (size_t)S * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(H[3].p, sg->q_base, (size_t)S * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(H[4].p, sg->t_base, (size_t)S * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(H[5].p, sg->read_len, (size_t)S * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(H[6].p, sg->contig_len, (size_t)S * 4, cudaMemcpyHostToDevice, st));
  lra_b200_ir_segments ds = *sg;
  ds.blocks_in = (const uint32_t *)H[0].p; ds.blk_off = (const uint64_t *)H[1].p; ds.blk_cnt = (const int32_t *)H[2].p;
  ds.q_base = (const uint32_t *)H[3].p; ds.t_base = (const uint32_t *)H[4].p; ds.read_len = (const int32_t *)H[5].p;
  ds.contig_len = (const int32_t *)H[6].p;
  lra_b200_ir_seg_result dr = *res;
  dr.block_off = (uint64_t *)H[7].p; dr.n_blocks = (int32_t *)((char *)H[7].p + (size_t)S * 8); dr.blocks = (uint32_t *)H[8].p;
  rc = ir_segments_run_device(ctx, q, t, &ds, &dr);
  res->n_blocks_total = dr.n_blocks_total; res->cells = dr.cells; res->n_dp_groups = dr.n_dp_groups; res->n_aog_jobs = dr.n_aog_jobs;
  if (rc != LRA_B200_OK && rc != LRA_B200_EOVERFLOW) return rc;
  CU(cudaMemcpyAsync(res->n_blocks, dr.n_blocks, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->block_off, dr.block_off, (size_t)S * 8, cudaMemcpyDeviceToHost, st));
  if (rc == LRA_B200_OK && dr.n_blocks_total)
    CU(cudaMemcpyAsync(res->blocks, dr.blocks, (size_t)dr.n_blocks_total * 12, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return rc;
}


// ---------------------------------------------------------------------------------------------------- a1-a5 seeding
struct lra_b200_index {
  unsigned long long *t = nullptr;
  uint32_t *pos = nullptr;
  uint64_t n = 0;
};

extern "C" int lra_b200_index_upload(lra_b200_ctx *ctx, const uint64_t *t, const uint32_t *pos, uint64_t n, lra_b200_index **out) {
  if (!ctx || !out || (n && (!t || !pos))) return fail(ctx, LRA_B200_EINVAL, "index_upload: bad argument");
  CU(cudaSetDevice(ctx->device));
  lra_b200_index *ix = new lra_b200_index();
  ix->n = n;
  CU(cudaMalloc((void **)&ix->t, (n + 4) * 8));
  CU(cudaMalloc((void **)&ix->pos, (n + 4) * 4));
  CU(cudaMemcpyAsync(ix->t, t, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ix->pos, pos, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *out = ix;
  return LRA_B200_OK;
}
extern "C" void lra_b200_index_free(lra_b200_ctx *ctx, lra_b200_index *ix) {
  if (!ix) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  if (ix->t) cudaFree(ix->t);
  if (ix->pos) cudaFree(ix->pos);
  delete ix;
}

extern "C" int lra_b200_seq_revcomp(lra_b200_ctx *ctx, const lra_b200_seq *reads, const uint64_t *read_off, const uint32_t *read_len,
                                    int32_t n_reads, lra_b200_seq **out_rc) {
  if (!ctx || !reads || !out_rc || n_reads < 0 || (n_reads && (!read_off || !read_len))) return fail(ctx, LRA_B200_EINVAL, "seq_revcomp: bad argument");
  CU(cudaSetDevice(ctx->device));
  int rc;
  lra_b200_seq *o = *out_rc ? *out_rc : new lra_b200_seq();     // an arena passed in is re-used (re-sized if needed)
  const bool fresh = *out_rc == nullptr;
  *out_rc = nullptr;
  if ((rc = seq_reserve(ctx, o, reads->n))) { if (fresh) delete o; return rc; }
  o->n = reads->n;
  // positions outside every read keep the padding value (N)
  CU(cudaMemsetAsync(o->b2, 0, (o->cap_groups * 2 + 8) * 4, ctx->stream));
  CU(cudaMemsetAsync(o->nm, 0xFF, (o->cap_groups + 8) * 4, ctx->stream));
  if (n_reads) {
    if ((rc = ensure(ctx, ctx->sd[0], (size_t)n_reads * 8)) || (rc = ensure(ctx, ctx->sd[1], (size_t)n_reads * 4))) { lra_b200_seq_free(ctx, o); return rc; }
    CU(cudaMemcpyAsync(ctx->sd[0].p, read_off, (size_t)n_reads * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->sd[1].p, read_len, (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->stream));
    seq_revcomp_kernel<<<(unsigned)((n_reads + 7) / 8), 256, 0, ctx->stream>>>(SeqView{reads->b2, reads->nm, reads->n}, (const unsigned long long *)ctx->sd[0].p,
                                                                                (const uint32_t *)ctx->sd[1].p, n_reads, o->b2, o->nm);
    ctx->launches++;
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(ctx->stream));
  *out_rc = o;
  return LRA_B200_OK;
}

extern "C" int lra_b200_seed_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_index *index,
                                   const lra_b200_seed_reads *in, lra_b200_seed_result *res) {
  if (!ctx || !reads || !genome || !index || !in || !res) return fail(ctx, LRA_B200_EINVAL, "seed_batch: NULL argument");
  const int R = in->n_reads;
  if (R < 0 || in->k < 1 || in->k > 31 || in->w < 1 || in->w > kSeedMaxW) return fail(ctx, LRA_B200_EINVAL, "seed_batch: need 1 <= k <= 31, 1 <= w <= %d", kSeedMaxW);
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_matches = 0;
  if (R == 0) return LRA_B200_OK;
  if (!in->read_off || !in->read_len) return fail(ctx, LRA_B200_EINVAL, "seed_batch: NULL read descriptors");
  // the minimizer scratch is indexed by read_off: reads must lie inside the arena, in ascending order, without overlap
  for (int r = 0; r < R; r++) {
    if (in->read_off[r] + in->read_len[r] > reads->n) return fail(ctx, LRA_B200_EINVAL, "seed_batch: read %d ends beyond the arena", r);
    if (r && in->read_off[r] < in->read_off[r - 1] + in->read_len[r - 1]) return fail(ctx, LRA_B200_EINVAL, "seed_batch: reads overlap or are not in ascending order at %d", r);
  }
  int rc;
  DevBuf *B = ctx->sd;
  const size_t mmcap = (size_t)reads->n + 64;
  const size_t mcap = (size_t)(res->match_cap ? res->match_cap : 1);
  if ((rc = ensure(ctx, B[0], (size_t)R * 8)) || (rc = ensure(ctx, B[1], (size_t)R * 4)) || (rc = ensure(ctx, B[2], mmcap * 8)) ||
      (rc = ensure(ctx, B[3], mmcap * 4)) || (rc = ensure(ctx, B[4], (size_t)R * 4)) || (rc = ensure(ctx, B[5], ((size_t)R + 1) * 8)) ||
      (rc = ensure(ctx, B[6], mcap * 8)) || (rc = ensure(ctx, B[7], mcap * 8)) || (rc = ensure(ctx, B[8], mcap * 4)) ||
      (rc = ensure(ctx, B[9], mcap * 4)) || (rc = ensure(ctx, B[10], mcap)) || (rc = ensure(ctx, B[11], 64)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, in->read_off, (size_t)R * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[1].p, in->read_len, (size_t)R * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(B[11].p, 0, 64, st));
  SeedBatch b;
  b.reads = SeqView{reads->b2, reads->nm, reads->n};
  b.read_off = (const unsigned long long *)B[0].p; b.read_len = (const uint32_t *)B[1].p; b.n_reads = R;
  b.k = in->k; b.w = in->w; b.max_freq = in->max_freq;
  b.idx_t = index->t; b.idx_pos = index->pos; b.n_idx = (long long)index->n;
  b.genome = SeqView{genome->b2, genome->nm, genome->n};
  b.mm_t = (unsigned long long *)B[2].p; b.mm_pos = (uint32_t *)B[3].p; b.mm_n = (uint32_t *)B[4].p;
  b.match_cnt = (unsigned long long *)B[5].p;
  b.m_qt = (unsigned long long *)B[6].p; b.m_tt = (unsigned long long *)B[7].p; b.m_qpos = (uint32_t *)B[8].p; b.m_tpos = (uint32_t *)B[9].p;
  b.m_strand = (uint8_t *)B[10].p; b.match_cap = res->match_cap; b.err = (int *)B[11].p;
  const unsigned nb = (unsigned)((R + 127) / 128);
  int evi = 0;
  auto rec = [&]() { cudaEventRecord(ctx->ev[evi++], st); };
  rec(); seed_minimizers_kernel<<<nb, 128, 0, st>>>(b);
  rec(); seed_sort_kernel<<<nb, 128, 0, st>>>(b);
  rec(); seed_compare_kernel<false><<<nb, 128, 0, st>>>(b);
  rec(); seed_scan_kernel<<<1, 1024, 0, st>>>(b.match_cnt, R, b.match_cap, b.err);
  rec(); seed_compare_kernel<true><<<nb, 128, 0, st>>>(b);
  rec();
  ctx->launches += 5;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->match_off, b.match_cnt, ((size_t)R + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (res->n_minimizers) CU(cudaMemcpyAsync(res->n_minimizers, b.mm_n, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const uint64_t total = res->match_off[R];
  res->n_matches = total;
  static const char *names[5] = {"seed_minimizers", "seed_sort", "seed_compare<count>", "seed_scan", "seed_compare<emit>"};
  for (int i = 0; i < 5; i++) {
    lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "%s", names[i]);
    cudaEventElapsedTime(&s2.ms, ctx->ev[i], ctx->ev[i + 1]); s2.jobs = (uint64_t)R;
    ctx->stats.push_back(s2);
  }
  if (total > res->match_cap) return fail(ctx, LRA_B200_EOVERFLOW, "seed_batch: match capacity %llu too small, %llu needed",
                                         (unsigned long long)res->match_cap, (unsigned long long)total);
  if (total) {
    CU(cudaMemcpyAsync(res->q_t, b.m_qt, total * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->t_t, b.m_tt, total * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->q_pos, b.m_qpos, total * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->t_pos, b.m_tpos, total * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->strand, b.m_strand, total, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return LRA_B200_OK;
}


// count -> scan -> emit of a21.  One warp per segment (stats_warp_kernel); LRA_B200_STATS_THREAD=1 selects the statement-by-statement
// one-segment-per-thread kernels, which are kept as the cross-check.
static int stats_launch(lra_b200_ctx *ctx, StatsBatch &b, int S, size_t n_blocks_in, int *errflag) {
  int rc;
  if ((rc = ensure(ctx, ctx->stt_x[0], n_blocks_in * 12 + 64)) || (rc = ensure(ctx, ctx->stt_x[1], (size_t)S * 256))) return rc;
  b.pre = (uint32_t *)ctx->stt_x[0].p; b.lane_info = (uint32_t *)ctx->stt_x[1].p;
  cudaStream_t st = ctx->stream;
  static const bool thread_kernels = getenv("LRA_B200_STATS_THREAD") != nullptr;
  cudaEventRecord(ctx->ev[0], st);
  if (thread_kernels) {
    const unsigned nb = (unsigned)((S + 127) / 128);
    stats_kernel<false><<<nb, 128, 0, st>>>(b);
    seed_scan_kernel<<<1, 1024, 0, st>>>(b.cig_off, S, b.cigar_cap, errflag);
    stats_kernel<true><<<nb, 128, 0, st>>>(b);
  } else {
    const unsigned nb = (unsigned)((S + 3) / 4);
    stats_warp_kernel<false><<<nb, 128, 0, st>>>(b);
    seed_scan_kernel<<<1, 1024, 0, st>>>(b.cig_off, S, b.cigar_cap, errflag);
    stats_warp_kernel<true><<<nb, 128, 0, st>>>(b);
  }
  cudaEventRecord(ctx->ev[1], st);
  return 0;
}

// ---------------------------------------------------------------------------------------------------- a21 statistics
extern "C" int lra_b200_calc_stats_batch(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_ir_segments *sg,
                                         const float *log_lut, lra_b200_stats_result *res) {
  if (!ctx || !q || !t || !sg || !log_lut || !res) return fail(ctx, LRA_B200_EINVAL, "calc_stats_batch: NULL argument");
  const int S = sg->n_segments;
  if (S < 0) return fail(ctx, LRA_B200_EINVAL, "calc_stats_batch: negative segment count");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_cigar_total = 0;
  if (S == 0) return LRA_B200_OK;
  int rc;
  DevBuf *B = ctx->stt;
  const size_t T = (size_t)sg->n_blocks_in;
  const size_t ccap = (size_t)(res->cigar_cap ? res->cigar_cap : 1);
  if ((rc = ensure(ctx, B[0], T * 12 + 16)) || (rc = ensure(ctx, B[1], (size_t)S * 8)) || (rc = ensure(ctx, B[2], (size_t)S * 4)) ||
      (rc = ensure(ctx, B[3], (size_t)S * 4)) || (rc = ensure(ctx, B[4], (size_t)S * 4)) || (rc = ensure(ctx, B[5], (size_t)S * 4)) ||
      (rc = ensure(ctx, B[6], 2001 * 4)) || (rc = ensure(ctx, B[7], (size_t)S * 64)) || (rc = ensure(ctx, B[8], (size_t)S * 4)) ||
      (rc = ensure(ctx, B[9], ((size_t)S + 1) * 8)) || (rc = ensure(ctx, B[10], ccap * 4)) || (rc = ensure(ctx, B[11], 64)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, sg->blocks_in, T * 12, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[1].p, sg->blk_off, (size_t)S * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[2].p, sg->blk_cnt, (size_t)S * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[3].p, sg->q_base, (size_t)S * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[4].p, sg->t_base, (size_t)S * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[5].p, sg->read_len, (size_t)S * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[6].p, log_lut, 2001 * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(B[11].p, 0, 64, st));
  StatsBatch b;
  b.q = SeqView{q->b2, q->nm, q->n}; b.t = SeqView{t->b2, t->nm, t->n};
  b.blocks = (const uint32_t *)B[0].p; b.blk_off = (const unsigned long long *)B[1].p; b.blk_cnt = (const int32_t *)B[2].p;
  b.q_base = (const uint32_t *)B[3].p; b.t_base = (const uint32_t *)B[4].p; b.read_len = (const int32_t *)B[5].p; b.n_seg = S;
  b.lut = (const float *)B[6].p; b.stats = (int32_t *)B[7].p; b.value = (float *)B[8].p; b.cig_off = (unsigned long long *)B[9].p;
  b.cigar = (uint32_t *)B[10].p; b.cigar_cap = res->cigar_cap;
  if ((rc = stats_launch(ctx, b, S, (size_t)sg->n_blocks_in, (int *)B[11].p))) return rc;
  ctx->launches += 3;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->cigar_off, b.cig_off, ((size_t)S + 1) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->stats, b.stats, (size_t)S * 64, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->value, b.value, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  res->n_cigar_total = res->cigar_off[S];
  {
    lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "stats(count+scan+emit)");
    cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)S; ctx->stats.push_back(s2);
  }
  if (res->n_cigar_total > res->cigar_cap) return fail(ctx, LRA_B200_EOVERFLOW, "calc_stats_batch: cigar capacity %llu too small, %llu needed",
                                                      (unsigned long long)res->cigar_cap, (unsigned long long)res->n_cigar_total);
  if (res->n_cigar_total) {
    CU(cudaMemcpyAsync(res->cigar, b.cigar, (size_t)res->n_cigar_total * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return LRA_B200_OK;
}

extern "C" int lra_b200_calc_stats_batch_device(lra_b200_ctx *ctx, const lra_b200_seq *q, const lra_b200_seq *t, const lra_b200_ir_segments *sg,
                                                const float *log_lut, lra_b200_stats_result *res) {
  if (!ctx || !q || !t || !sg || !log_lut || !res) return fail(ctx, LRA_B200_EINVAL, "calc_stats_batch_device: NULL argument");
  const int S = sg->n_segments;
  if (S < 0) return fail(ctx, LRA_B200_EINVAL, "calc_stats_batch_device: negative segment count");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_cigar_total = 0;
  if (S == 0) return LRA_B200_OK;
  int rc;
  DevBuf *B = ctx->stt;
  if ((rc = ensure(ctx, B[6], 2001 * 4)) || (rc = ensure(ctx, B[11], 64))) return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[6].p, log_lut, 2001 * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(B[11].p, 0, 64, st));
  StatsBatch b;
  b.q = SeqView{q->b2, q->nm, q->n}; b.t = SeqView{t->b2, t->nm, t->n};
  b.blocks = sg->blocks_in; b.blk_off = (const unsigned long long *)sg->blk_off; b.blk_cnt = sg->blk_cnt;
  b.q_base = sg->q_base; b.t_base = sg->t_base; b.read_len = sg->read_len; b.n_seg = S;
  b.lut = (const float *)B[6].p; b.stats = res->stats; b.value = res->value; b.cig_off = (unsigned long long *)res->cigar_off;
  b.cigar = res->cigar; b.cigar_cap = res->cigar_cap;
  if ((rc = stats_launch(ctx, b, S, (size_t)sg->n_blocks_in, (int *)B[11].p))) return rc;
  ctx->launches += 3;
  CU(cudaGetLastError());
  unsigned long long total = 0;
  CU(cudaMemcpyAsync(&total, b.cig_off + S, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  res->n_cigar_total = total;
  {
    lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "stats(count+scan+emit)");
    cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)S;
    s2.algo_bytes = 12ull * sg->n_blocks_in + 4ull * total + 100ull * (uint64_t)S;     // blocks in, CIGAR out, descriptors + counters (bases: see DESIGN.md)
    ctx->stats.push_back(s2);
  }
  if (total > res->cigar_cap) return fail(ctx, LRA_B200_EOVERFLOW, "calc_stats_batch_device: cigar capacity %llu too small, %llu needed",
                                         (unsigned long long)res->cigar_cap, total);
  return LRA_B200_OK;
}

#include "lref_host.cuh"
#include "mp_host.cuh"
#include "mp_host_map.cuh"
#include "gidx_host.cuh"
