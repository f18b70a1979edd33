// Host side of the mapper entry points (included at the end of lra_b200.cu): lra_b200_sdp_batch (a10 stand-alone) and the
// lra_b200_mapper_* family (MapRead_lowacc composed on the device, reference MapRead.h:153-263 + Map_lowacc.h:69-632).
#pragma once
#include "mp_sdp_driver.cuh"

static void mp_fill_pwl(lra::mp::Pwl &P, const int64_t *stops, const float *slope, const float *inter, int c1, int c2) {
  for (int i = 0; i < 25; i++) { P.stops[i] = stops[i]; P.slope[i] = slope[i]; P.inter[i] = inter[i]; }
  P.ceil1 = c1; P.ceil2 = c2;
}

// InitPWL (SubRountine.h:43-101) on the host with the host libm, exactly the reference's expression order; uploaded, never
// recomputed on the device.
extern "C" int lra_b200_init_pwl(float intercept, float scalar, float root, int32_t ceil1, int32_t ceil2, int64_t *stops, float *slope, float *inter) {
  (void)ceil1; (void)ceil2;
  static const long S[25] = {0, 5, 10, 20, 40, 80, 100, 200, 300, 500, 1000, 2000, 3000, 4000, 5000, 6000, 7000, 8000, 9000, 15000, 20000, 30000, 40000, 50000, 100000};
  float vals[25];
  vals[0] = 0;
  for (int i = 0; i < 25; i++) stops[i] = S[i];
  for (int i = 1; i < 25; i++) {
    if (i <= 2) intercept = 0;
    vals[i] = intercept + scalar * std::pow((float)S[i], 1 / root);
  }
  for (int i = 0; i < 25; i++) { slope[i] = 0; inter[i] = 0; }
  for (int i = 0; i < 24; i++) {
    float sl = (vals[i + 1] - vals[i]) / (S[i + 1] - S[i]);
    if (S[i] <= 10) { slope[i] = 0; inter[i] = 0; }
    else { slope[i] = sl; inter[i] = vals[i] - S[i] * sl + intercept; }
  }
  return LRA_B200_OK;
}

extern "C" int lra_b200_sdp_batch(lra_b200_ctx *ctx, const lra_b200_sdp_problems *pr, lra_b200_sdp_result *res) {
  using namespace lra::mp;
  if (!ctx || !pr || !res) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: NULL argument");
  const int n = pr->n_prob;
  if (n < 0 || pr->max_aln < 1 || pr->max_aln > 8) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: bad problem count / max_aln");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (n == 0) return LRA_B200_OK;
  if (!pr->mode || !pr->frag_off || !pr->cl_off_off || !pr->cl_off || !pr->cl_strand || !pr->only_cl || !pr->rate || !pr->irate || !pr->read_len || !pr->pwl_stops ||
      !pr->pwl_slope || !pr->pwl_inter || !res->n_chains || !res->chain_len || !res->chain_val || !res->bounds)
    return fail(ctx, LRA_B200_EINVAL, "sdp_batch: NULL array");
  const size_t NF = (size_t)pr->frag_off[n], NC = (size_t)pr->cl_off_off[n];
  size_t maxf = 0;
  for (int p = 0; p < n; p++) {
    if (pr->frag_off[p + 1] < pr->frag_off[p] || pr->cl_off_off[p + 1] < pr->cl_off_off[p] + 1) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: offsets of problem %d", p);
    const size_t nf = (size_t)(pr->frag_off[p + 1] - pr->frag_off[p]);
    if (nf > (1u << 22)) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: problem %d has more than 2^22 anchors", p);
    if (pr->mode[p] < 0 || pr->mode[p] > 4) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: mode of problem %d", p);
    if (pr->mode[p] == 4 && (!pr->q_end || !pr->t_end || !pr->frag_strand || !pr->frag_val || !pr->frag_n0)) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: mode 4 needs q_end / t_end / frag_strand / frag_val / frag_n0");
    const int ncl = (int)(pr->cl_off_off[p + 1] - pr->cl_off_off[p]) - 1;
    if (pr->mode[p] == 1 && (pr->only_cl[p] < 0 || pr->only_cl[p] >= ncl)) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: cluster index of problem %d", p);
    if (pr->mode[p] != 2 && pr->mode[p] != 4) {
      const int32_t *co = pr->cl_off + pr->cl_off_off[p];
      if (co[0] != 0 || (size_t)co[ncl] != nf) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: cluster offsets of problem %d do not cover its anchors", p);
      for (int c = 0; c < ncl; c++) if (co[c + 1] < co[c]) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: cluster offsets of problem %d not ascending", p);
    }
    maxf = nf > maxf ? nf : maxf;
  }
  if (NF && (!pr->q || !pr->t || !pr->len || !res->chain || !res->link || !res->cl_of_frag)) return fail(ctx, LRA_B200_EINVAL, "sdp_batch: NULL anchor arrays");
  const int MA = pr->max_aln;
  int warps = ctx->n_sm * 8; if (warps > n) warps = n;
  const size_t per = (size_t)maxf * 12288 + (4u << 20);      // measured ~3.7 KB per anchor on captured ONT problems; 3x margin
  int rc;
  DevBuf *B = ctx->mp;
  size_t sz[24] = {(size_t)n * 4, ((size_t)n + 1) * 8, NF * 4 + 16, NF * 4 + 16, NF * 4 + 16, ((size_t)n + 1) * 8, NC * 4 + 16, NC + 16, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4,
                   (size_t)n * 4, sizeof(Pwl), (size_t)n * 4, (size_t)n * MA * 4, (size_t)n * MA * 4, (size_t)n * MA * 16, NF * MA * 4 + 16, NF * MA + 16, NF * 4 + 16,
                   (size_t)warps * per, 64, 0, 0};
  for (int i = 0; i < 22; i++) if ((rc = ensure(ctx, B[i], sz[i]))) return rc;
  cudaStream_t st = ctx->stream;
  Pwl hp; mp_fill_pwl(hp, pr->pwl_stops, pr->pwl_slope, pr->pwl_inter, pr->ceil1, pr->ceil2);
  const void *src[14] = {pr->mode, pr->frag_off, pr->q, pr->t, pr->len, pr->cl_off_off, pr->cl_off, pr->cl_strand, pr->only_cl, pr->rate, pr->irate, pr->read_len, &hp, nullptr};
  for (int i = 0; i < 13; i++) if (sz[i] && src[i]) {
    size_t bytes = sz[i];
    if (i >= 2 && i <= 4) bytes = NF * 4; if (i == 6) bytes = NC * 4; if (i == 7) bytes = NC;
    if (bytes) CU(cudaMemcpyAsync(B[i].p, src[i], bytes, cudaMemcpyHostToDevice, st));
  }
  CU(cudaMemsetAsync(B[21].p, 0, 64, st));
  SdpBatch b;
  b.n_prob = n; b.max_aln = MA; b.mode = (const int *)B[0].p; b.frag_off = (const unsigned long long *)B[1].p; b.q = (const uint32_t *)B[2].p; b.t = (const uint32_t *)B[3].p;
  b.len = (const int32_t *)B[4].p; b.cl_off_off = (const unsigned long long *)B[5].p; b.cl_off = (const int *)B[6].p; b.cl_strand = (const uint8_t *)B[7].p;
  b.only_cl = (const int *)B[8].p; b.rate = (const float *)B[9].p; b.irate = (const int *)B[10].p; b.read_len = (const int *)B[11].p;
  b.alnthres = pr->alnthres; b.NumAln = pr->num_aln; b.pwl = (const Pwl *)B[12].p;
  b.n_chains = (int *)B[13].p; b.chain_len = (int *)B[14].p; b.chain_val = (float *)B[15].p; b.bounds = (uint32_t *)B[16].p; b.chain = (uint32_t *)B[17].p;
  b.link = (uint8_t *)B[18].p; b.cl_of_frag = (int *)B[19].p; b.arena = (unsigned char *)B[20].p; b.arena_per_warp = per; b.err = (int *)B[21].p;
  b.peak = (unsigned long long *)((char *)B[21].p + 8);
  b.qe = b.te = nullptr; b.fstrand = nullptr; b.fval = nullptr; b.fn0 = nullptr; b.out_n0 = nullptr; b.globalK = pr->global_k;
  if (pr->q_end && pr->t_end && pr->frag_strand && pr->frag_val && pr->frag_n0 && NF) {
    DevBuf *X = ctx->mp + 24;
    const size_t xs[6] = {NF * 4 + 16, NF * 4 + 16, NF + 16, NF * 4 + 16, NF * 4 + 16, (size_t)n * MA * 4 + 16};
    for (int i = 0; i < 6; i++) if ((rc = ensure(ctx, X[i], xs[i]))) return rc;
    CU(cudaMemcpyAsync(X[0].p, pr->q_end, NF * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(X[1].p, pr->t_end, NF * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(X[2].p, pr->frag_strand, NF, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(X[3].p, pr->frag_val, NF * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(X[4].p, pr->frag_n0, NF * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemsetAsync(X[5].p, 0, xs[5], st));
    b.qe = (const uint32_t *)X[0].p; b.te = (const uint32_t *)X[1].p; b.fstrand = (const uint8_t *)X[2].p; b.fval = (const float *)X[3].p; b.fn0 = (const int32_t *)X[4].p;
    b.out_n0 = (int *)X[5].p;
  }
  CU(cudaMemsetAsync(B[14].p, 0, sz[14], st)); CU(cudaMemsetAsync(B[15].p, 0, sz[15], st)); CU(cudaMemsetAsync(B[16].p, 0, sz[16], st));
  cudaEventRecord(ctx->ev[0], st);
  sdp_batch_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->n_chains, b.n_chains, sz[13], cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->chain_len, b.chain_len, sz[14], cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->chain_val, b.chain_val, sz[15], cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->bounds, b.bounds, sz[16], cudaMemcpyDeviceToHost, st));
  if (NF) {
    CU(cudaMemcpyAsync(res->chain, b.chain, NF * MA * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->link, b.link, NF * MA, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->cl_of_frag, b.cl_of_frag, NF * 4, cudaMemcpyDeviceToHost, st));
  }
  if (b.out_n0 && res->num_anchors0) CU(cudaMemcpyAsync(res->num_anchors0, b.out_n0, (size_t)n * MA * 4, cudaMemcpyDeviceToHost, st));
  unsigned long long h[2] = {0, 0};
  CU(cudaMemcpyAsync(h, B[21].p, 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  res->arena_peak = h[1];
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "sdp_batch");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)n; s2.algo_bytes = 12ull * NF;
  ctx->stats.push_back(s2);
  if ((int)h[0]) return fail(ctx, LRA_B200_EOVERFLOW, "sdp_batch: a problem exceeded its worker arena (%zu bytes per warp)", per);
  return LRA_B200_OK;
}
