// Host side of the C ABI for a12 (lra_b200_lindex_*) and a13 (lra_b200_refine_clusters_batch); included by lra_b200.cu.
#pragma once

struct lra_b200_lindex {
  unsigned long long *win_off = nullptr, *bnd = nullptr, *seq_start = nullptr;
  uint32_t *win_len = nullptr, *mins = nullptr, *win_first = nullptr, *seq_len = nullptr;
  uint64_t n_win = 0, n_mins = 0;
  int n_seq = 0, window = 0;
  size_t cap_win = 0, cap_seq = 0, cap_mins = 0;     // allocated entries (an image handed back to lindex_build is re-used)
};

static void lindex_release(lra_b200_lindex *li) {
  void *p[] = {li->win_off, li->bnd, li->seq_start, li->win_len, li->mins, li->win_first, li->seq_len};
  for (void *x : p) if (x) cudaFree(x);
  delete li;
}

extern "C" void lra_b200_lindex_free(lra_b200_ctx *ctx, lra_b200_lindex *li) {
  if (!li) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  lindex_release(li);
}

extern "C" void lra_b200_lindex_sizes(const lra_b200_lindex *li, uint64_t *n_win, uint64_t *n_mins) {
  if (n_win) *n_win = li ? li->n_win : 0;
  if (n_mins) *n_mins = li ? li->n_mins : 0;
}

// window table of n_seqs sequences cut into `window`-base pieces (IndexSeq, MMIndex.h:200-206,227-232)
static int lindex_layout(lra_b200_ctx *ctx, lra_b200_lindex *li, const uint64_t *seq_start, const uint32_t *seq_len, int n_seqs, int window,
                         std::vector<unsigned long long> &win_off, std::vector<uint32_t> &win_len) {
  std::vector<uint32_t> win_first((size_t)n_seqs + 1);
  for (int s = 0; s < n_seqs; s++) {
    win_first[s] = (uint32_t)win_off.size();
    for (uint32_t p = 0; p < seq_len[s]; p += (uint32_t)window) {
      win_off.push_back(seq_start[s] + p);
      win_len.push_back(seq_len[s] - p < (uint32_t)window ? seq_len[s] - p : (uint32_t)window);
    }
  }
  win_first[n_seqs] = (uint32_t)win_off.size();
  li->n_win = win_off.size();
  li->n_seq = n_seqs; li->window = window;
  win_off.push_back(n_seqs ? seq_start[n_seqs - 1] + seq_len[n_seqs - 1] : 0ull);     // the closing offset of the last sequence
  const size_t W = li->n_win;
  if (W + 2 > li->cap_win) {
    if (li->win_off) { CU(cudaStreamSynchronize(ctx->stream)); cudaFree(li->win_off); cudaFree(li->win_len); cudaFree(li->bnd); }
    li->cap_win = W + W / 4 + 16;
    CU(cudaMalloc(&li->win_off, li->cap_win * 8)); CU(cudaMalloc(&li->win_len, li->cap_win * 4)); CU(cudaMalloc(&li->bnd, li->cap_win * 8));
  }
  if ((size_t)n_seqs + 1 > li->cap_seq) {
    if (li->win_first) { CU(cudaStreamSynchronize(ctx->stream)); cudaFree(li->win_first); cudaFree(li->seq_start); cudaFree(li->seq_len); }
    li->cap_seq = (size_t)n_seqs + (size_t)n_seqs / 4 + 16;
    CU(cudaMalloc(&li->win_first, li->cap_seq * 4)); CU(cudaMalloc(&li->seq_start, li->cap_seq * 8)); CU(cudaMalloc(&li->seq_len, li->cap_seq * 4));
  }
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(li->win_off, win_off.data(), (W + 1) * 8, cudaMemcpyHostToDevice, st));
  if (W) CU(cudaMemcpyAsync(li->win_len, win_len.data(), W * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(li->win_first, win_first.data(), ((size_t)n_seqs + 1) * 4, cudaMemcpyHostToDevice, st));
  if (n_seqs) {
    CU(cudaMemcpyAsync(li->seq_start, seq_start, (size_t)n_seqs * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(li->seq_len, seq_len, (size_t)n_seqs * 4, cudaMemcpyHostToDevice, st));
  }
  CU(cudaStreamSynchronize(st));     // the host vectors go out of scope with the caller
  return LRA_B200_OK;
}

extern "C" int lra_b200_lindex_build(lra_b200_ctx *ctx, const lra_b200_seq *seq, const uint64_t *seq_start, const uint32_t *seq_len, int32_t n_seqs,
                                     int32_t k, int32_t w, int32_t window, int32_t max_freq, lra_b200_lindex **out) {
  if (!ctx || !seq || !out || n_seqs < 0 || (n_seqs && (!seq_start || !seq_len))) return fail(ctx, LRA_B200_EINVAL, "lindex_build: bad argument");
  if (k < 1 || k > 10 || w < 1 || w + k - 1 > 60 || window < 1 || window > kLidxMaxWindow || max_freq < 1)
    return fail(ctx, LRA_B200_EINVAL, "lindex_build: need 1 <= k <= 10 (20-bit LocalTuple), w >= 1, w + k <= 61, 1 <= window <= %d, max_freq >= 1", kLidxMaxWindow);
  for (int s = 0; s < n_seqs; s++)
    if (seq_start[s] + seq_len[s] > seq->n) return fail(ctx, LRA_B200_EINVAL, "lindex_build: sequence %d ends beyond the arena", s);
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  lra_b200_lindex *li = *out ? *out : new lra_b200_lindex();     // an image passed in is rebuilt in place (its buffers are re-used)
  *out = nullptr;
  std::vector<unsigned long long> win_off; std::vector<uint32_t> win_len;
  int rc = lindex_layout(ctx, li, seq_start, seq_len, n_seqs, window, win_off, win_len);
  if (rc) { lindex_release(li); return rc; }
  const size_t W = li->n_win;
  cudaStream_t st = ctx->stream;
  if (W == 0) { CU(cudaMemsetAsync(li->bnd, 0, 16, st)); CU(cudaStreamSynchronize(st)); *out = li; return LRA_B200_OK; }
  // staging: one slot per arena base (transient; 4 B/base -- 12 GB for a 3 Gb genome, freed below)
  // (read batches re-use a grow-only context buffer; anything above 4 GB is allocated for this call only)
  uint32_t *tmp = nullptr;
  const size_t tmp_bytes = ((size_t)seq->n + 64) * 4;
  const bool own_tmp = tmp_bytes > (4ull << 30);
  if (own_tmp) {
    if (cudaMalloc(&tmp, tmp_bytes) != cudaSuccess) { lindex_release(li); return fail(ctx, LRA_B200_ECUDA, "lindex_build: cannot allocate %zu bytes of staging", tmp_bytes); }
  } else {
    if ((rc = ensure(ctx, ctx->li_tmp, tmp_bytes))) { lindex_release(li); return rc; }
    tmp = (uint32_t *)ctx->li_tmp.p;
  }
  LidxBuild b;
  b.seq = SeqView{seq->b2, seq->nm, seq->n};
  b.win_off = li->win_off; b.win_len = li->win_len; b.n_win = (int)W; b.k = k; b.w = w; b.max_freq = max_freq;
  b.tmp = tmp; b.cnt = li->bnd; b.mins = nullptr;
  if ((rc = ensure(ctx, ctx->lr[0], 64))) { if (own_tmp) cudaFree(tmp); lindex_release(li); return rc; }
  cudaMemsetAsync(ctx->lr[0].p, 0, 64, st);
  cudaEventRecord(ctx->ev[0], st);
  lidx_window_kernel<<<(unsigned)((W + kLidxWarps - 1) / kLidxWarps), 32 * kLidxWarps, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  seed_scan_kernel<<<1, 1024, 0, st>>>(li->bnd, (int)W, ~0ull, (int *)ctx->lr[0].p);
  ctx->launches += 2;
  unsigned long long total = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&total, li->bnd + W, 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess && (size_t)total + 16 > li->cap_mins) {
    if (li->mins) cudaFree(li->mins);
    li->mins = nullptr;
    li->cap_mins = (size_t)total + (size_t)total / 8 + 16;
    e = cudaMalloc(&li->mins, li->cap_mins * 4);
  }
  if (e != cudaSuccess) { if (own_tmp) cudaFree(tmp); lindex_release(li); return fail(ctx, LRA_B200_ECUDA, "lindex_build: %s", cudaGetErrorString(e)); }
  li->n_mins = total;
  b.mins = li->mins;
  cudaEventRecord(ctx->ev[2], st);
  lidx_compact_kernel<<<(unsigned)((W + 7) / 8), 256, 0, st>>>(b);
  cudaEventRecord(ctx->ev[3], st);
  ctx->launches++;
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (own_tmp) cudaFree(tmp);
  if (e != cudaSuccess) { lindex_release(li); return fail(ctx, LRA_B200_ECUDA, "lindex_build: %s", cudaGetErrorString(e)); }
  {
    lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "lidx_window");
    cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = W;
    // algorithmic bytes: 2-bit bases + mask read once, the kept tuples written once
    s2.algo_bytes = (seq->n * 3) / 8 + 12 * W + 4ull * total;
    ctx->stats.push_back(s2);
    memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "lidx_scan+compact");
    cudaEventElapsedTime(&s2.ms, ctx->ev[1], ctx->ev[3]); s2.jobs = W; s2.algo_bytes = 8ull * total + 16 * W;
    ctx->stats.push_back(s2);
  }
  *out = li;
  return LRA_B200_OK;
}

extern "C" int lra_b200_lindex_upload(lra_b200_ctx *ctx, const uint64_t *seq_start, const uint32_t *seq_len, int32_t n_seqs, int32_t window,
                                      const uint64_t *win_off, const uint64_t *bnd, const uint32_t *mins, uint64_t n_win, lra_b200_lindex **out) {
  if (!ctx || !out || n_seqs < 0 || !seq_start || !seq_len || !win_off || !bnd || (bnd[n_win] && !mins) || window < 1 || window > kLidxMaxWindow)
    return fail(ctx, LRA_B200_EINVAL, "lindex_upload: bad argument");
  CU(cudaSetDevice(ctx->device));
  *out = nullptr;
  lra_b200_lindex *li = new lra_b200_lindex();
  std::vector<unsigned long long> wo; std::vector<uint32_t> wl;
  int rc = lindex_layout(ctx, li, seq_start, seq_len, n_seqs, window, wo, wl);
  if (rc) { lindex_release(li); return rc; }
  if (li->n_win != n_win) { lindex_release(li); return fail(ctx, LRA_B200_EINVAL, "lindex_upload: %llu windows given, the sequences have %llu", (unsigned long long)n_win, (unsigned long long)wo.size() - 1); }
  for (uint64_t i = 0; i <= n_win; i++)
    if (win_off[i] != wo[i]) { lindex_release(li); return fail(ctx, LRA_B200_EINVAL, "lindex_upload: window offset %llu differs from the sequence layout", (unsigned long long)i); }
  li->n_mins = bnd[n_win];
  cudaStream_t st = ctx->stream;
  li->cap_mins = (size_t)li->n_mins + 16;
  cudaError_t e = cudaMalloc(&li->mins, li->cap_mins * 4);
  if (e == cudaSuccess) e = cudaMemcpyAsync(li->bnd, bnd, (n_win + 1) * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && li->n_mins) e = cudaMemcpyAsync(li->mins, mins, (size_t)li->n_mins * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { lindex_release(li); return fail(ctx, LRA_B200_ECUDA, "lindex_upload: %s", cudaGetErrorString(e)); }
  *out = li;
  return LRA_B200_OK;
}

extern "C" int lra_b200_lindex_download(lra_b200_ctx *ctx, const lra_b200_lindex *li, uint64_t *win_off, uint64_t *bnd, uint32_t *mins) {
  if (!ctx || !li) return fail(ctx, LRA_B200_EINVAL, "lindex_download: NULL argument");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (win_off) CU(cudaMemcpyAsync(win_off, li->win_off, (li->n_win + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (bnd) CU(cudaMemcpyAsync(bnd, li->bnd, (li->n_win + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (mins && li->n_mins) CU(cudaMemcpyAsync(mins, li->mins, (size_t)li->n_mins * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return LRA_B200_OK;
}

static LidxView lidx_view(const lra_b200_lindex *li) {
  return LidxView{li->win_off, li->win_len, li->bnd, li->mins, li->win_first, li->seq_start, li->seq_len, (int)li->n_win, li->n_seq};
}

// pow2ceil(anchors of cluster c) -> key_off[c] (scanned afterwards): slots of the bitonic sort scratch
__global__ void lref_keyslots_kernel(const unsigned long long *m_off, int n, unsigned long long *key_off) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const unsigned long long nm = m_off[c + 1] - m_off[c];
  unsigned long long P = 1;
  while (P < nm) P <<= 1;
  key_off[c] = P;
}

// The launch sequence of a13 on device-resident clusters (cl: device pointers) into device-resident results (res: device pointers;
// m_q_out / m_t_out / box_out may be NULL).  M = total number of input anchors.
struct LrefChainExtra { const uint32_t *m_len; const uint8_t *m_strand; const int32_t *chrom; int limitrefine; };   // device pointers (mode 1)

static int lref_run(lra_b200_ctx *ctx, const lra_b200_lindex *gl, const lra_b200_lindex *rf, const lra_b200_lindex *rr, const lra_b200_clusters *cl,
                    size_t M, lra_b200_refined *res, const LrefChainExtra *chain = nullptr) {
  const int n = cl->n_clusters;
  int rc;
  DevBuf *B = ctx->lr;
  if ((rc = ensure(ctx, B[8], M * 4 + 16)) || (rc = ensure(ctx, B[9], M * 4 + 16)) || (rc = ensure(ctx, B[10], (size_t)n * 16)) ||
      (rc = ensure(ctx, B[11], (2 * M + (size_t)n + 2) * 8)) || (rc = ensure(ctx, B[12], ((size_t)n + 1) * 8)) ||
      (rc = ensure(ctx, B[16], (size_t)n * 4)) || (rc = ensure(ctx, B[17], (size_t)n * 4)) || (rc = ensure(ctx, B[18], ((size_t)n + 1) * 8)) ||
      (rc = ensure(ctx, B[0], 64)) || (rc = ensure(ctx, ctx->lr_x[0], (size_t)n * 16)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemsetAsync(B[0].p, 0, 64, st));
  LrefBatch b;
  memset(&b, 0, sizeof b);
  b.n_clusters = n;
  b.fbox = (uint32_t *)ctx->lr_x[0].p;
  if (chain) { b.mode = 1; b.limitrefine = chain->limitrefine; b.in_len = chain->m_len; b.in_mstrand = chain->m_strand; b.in_chrom = chain->chrom; }
  b.in_q = cl->m_q; b.in_t = cl->m_t; b.m_off = (const unsigned long long *)cl->m_off; b.in_box = cl->box;
  b.strand = cl->strand; b.read_id = cl->read_id; b.hdr_pos = (const unsigned long long *)cl->hdr_pos; b.n_hdr = cl->n_hdr;
  b.gl = lidx_view(gl); b.rd[0] = lidx_view(rf); b.rd[1] = lidx_view(rr);
  b.global_k = cl->global_k; b.small_k = cl->small_k; b.window = cl->window; b.local_max_freq = cl->local_max_freq;
  b.m_q = res->m_q_out ? res->m_q_out : (uint32_t *)B[8].p; b.m_t = res->m_t_out ? res->m_t_out : (uint32_t *)B[9].p;
  b.box = res->box_out ? res->box_out : (uint32_t *)B[10].p;
  b.keys = (unsigned long long *)B[11].p; b.key_off = (const unsigned long long *)B[12].p;
  b.status = res->status; b.chrom = res->chrom; b.diag = (long long *)res->diag;
  b.chrom_off = (uint32_t *)B[16].p; b.ls = (int32_t *)B[17].p; b.unit_off = (unsigned long long *)B[18].p;
  b.r_q = res->r_q; b.r_t = res->r_t; b.r_tup = res->r_tup; b.out_cap = res->anchor_cap;
  b.r_off = (unsigned long long *)res->r_off; b.rbox = res->rbox; b.eff = res->eff;
  int *errflag = (int *)B[0].p;
  int evi = 0;
  auto rec = [&]() { cudaEventRecord(ctx->ev[evi++], st); };
  rec();
  if (!chain) {
    lref_keyslots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b.m_off, n, (unsigned long long *)B[12].p);
    seed_scan_kernel<<<1, 1024, 0, st>>>((unsigned long long *)B[12].p, n, ~0ull, errflag);
  }
  lref_prep_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(b);
  seed_scan_kernel<<<1, 1024, 0, st>>>(b.unit_off, n, ~0ull, errflag);
  ctx->launches += 4;
  rec();
  CU(cudaGetLastError());
  unsigned long long n_units = 0, n_tasks = 0, n_out = 0;
  CU(cudaMemcpyAsync(&n_units, b.unit_off + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if ((rc = ensure(ctx, B[19], (size_t)(n_units + 1) * 4)) || (rc = ensure(ctx, B[20], (size_t)(n_units + 1) * 4)) ||
      (rc = ensure(ctx, B[21], (size_t)(n_units + 1) * 4)) || (rc = ensure(ctx, B[22], (size_t)(n_units + 2) * 8)) ||
      (rc = ensure(ctx, ctx->lr_x[1], (size_t)(n_units + 1) * 16)))
    return rc;
  b.u_cluster = (uint32_t *)B[19].p; b.u_qis = (uint32_t *)B[20].p; b.u_gstart = (uint32_t *)B[21].p; b.task_off = (unsigned long long *)B[22].p;
  b.u_band = (long long *)ctx->lr_x[1].p;
  CU(cudaMemsetAsync(b.task_off, 0, (size_t)(n_units + 2) * 8, st));
  if (n_units) {
    if (n_units > 0x7FFFFFFFull) return fail(ctx, LRA_B200_EINVAL, "refine_clusters_batch: %llu (cluster, window) units in one batch", n_units);
    if (chain) lref_chain_unit_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b);
    else lref_unit_kernel<<<(unsigned)((n_units + 127) / 128), 128, 0, st>>>(b, n_units);
    seed_scan_kernel<<<1, 1024, 0, st>>>(b.task_off, (int)n_units, ~0ull, errflag);
    ctx->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&n_tasks, b.task_off + n_units, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  rec();
  if ((rc = ensure(ctx, B[23], (size_t)(n_tasks + 2) * 8))) return rc;
  b.out_off = (unsigned long long *)B[23].p;
  CU(cudaMemsetAsync(b.out_off, 0, (size_t)(n_tasks + 2) * 8, st));
  if (n_tasks) {
    if (n_tasks > 0x7FFFFFFFull) return fail(ctx, LRA_B200_EINVAL, "refine_clusters_batch: %llu window pairs in one batch", n_tasks);
    // measured on B200 (tools/lref_timing.py): one task per THREAD (3.7 ms per pass for 74 k window pairs) beats one task per warp
    // (16 ms): the stage is bound by instruction issue, not by latency, and the lock-step warp form issues 8x more instructions
    static const bool literal = getenv("LRA_B200_LREF_WARP") == nullptr;
    // anchors per window pair parked by the count pass (0 = off: the emit pass re-runs every non-empty pair)
    static const int slot_cap = getenv("LRA_B200_LREF_SLOT") ? atoi(getenv("LRA_B200_LREF_SLOT")) : 256;   // measured: 48 -> 131.4, 128 -> 128.1, 256 -> 124.8 ms/step
    b.slot = nullptr; b.slot_cap = 0;
    if (literal && slot_cap > 0 && n_tasks * (unsigned long long)slot_cap * 12ull <= (4ull << 30)) {
      if ((rc = ensure(ctx, ctx->lr_x[5], (size_t)n_tasks * (size_t)slot_cap * 12))) return rc;
      b.slot = (uint32_t *)ctx->lr_x[5].p; b.slot_cap = slot_cap;
    }
    if (literal) lref_task_literal_kernel<false><<<(unsigned)((n_tasks + 127) / 128), 128, 0, st>>>(b, n_units, n_tasks);
    else lref_task_kernel<false><<<(unsigned)((n_tasks + 3) / 4), 128, 0, st>>>(b, n_units, n_tasks);
    seed_scan_kernel<<<1, 1024, 0, st>>>(b.out_off, (int)n_tasks, res->anchor_cap, errflag);
    rec();
    if (b.slot) { lref_task_copy_kernel<<<(unsigned)((n_tasks * 8 + 255) / 256), 256, 0, st>>>(b, n_tasks); ctx->launches++; }
    if (literal) lref_task_literal_kernel<true><<<(unsigned)((n_tasks + 127) / 128), 128, 0, st>>>(b, n_units, n_tasks);
    else lref_task_kernel<true><<<(unsigned)((n_tasks + 3) / 4), 128, 0, st>>>(b, n_units, n_tasks);
    ctx->launches += 3;
  } else rec();
  rec();
  lref_finish_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(b, n_units, n_tasks);
  ctx->launches++;
  rec();
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(&n_out, b.out_off + n_tasks, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  res->n_anchors = n_out; res->n_units = n_units; res->n_tasks = n_tasks;
  const char *names[5] = {chain ? "lref_prep(chain)" : "lref_prep(+sort)", chain ? "lref_chain_unit" : "lref_unit", "lref_task<count>", "lref_task<emit>", "lref_finish"};
  const uint64_t jobs[5] = {(uint64_t)n, n_units, n_tasks, n_tasks, (uint64_t)n};
  const uint64_t nmins_q = rf->n_mins + rr->n_mins;
  for (int i = 0; i < 5; i++) {
    lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "%s", names[i]);
    cudaEventElapsedTime(&s2.ms, ctx->ev[i], ctx->ev[i + 1]); s2.jobs = jobs[i];
    // algorithmic bytes of a task pass: both tuple lists of every window pair once (~2 x 700 x 4 B), 24 B of descriptors, anchors out
    if (i == 2 || i == 3) s2.algo_bytes = n_tasks * 24 + (n_tasks ? (uint64_t)((double)n_tasks * 4.0 * ((double)nmins_q / (double)(rf->n_win + rr->n_win + 1) + (double)gl->n_mins / (double)(gl->n_win + 1))) : 0)
                                          + (i == 3 ? 12 * n_out : 0);
    ctx->stats.push_back(s2);
  }
  if (n_out > res->anchor_cap) return fail(ctx, LRA_B200_EOVERFLOW, "refine_clusters_batch: anchor capacity %llu too small, %llu needed",
                                           (unsigned long long)res->anchor_cap, n_out);
  return LRA_B200_OK;
}

extern "C" int lra_b200_refine_clusters_batch_device(lra_b200_ctx *ctx, const lra_b200_lindex *gl, const lra_b200_lindex *rf, const lra_b200_lindex *rr,
                                                     const lra_b200_clusters *cl, uint64_t n_anchors_in, lra_b200_refined *res) {
  if (!ctx || !gl || !rf || !rr || !cl || !res) return fail(ctx, LRA_B200_EINVAL, "refine_clusters_batch_device: NULL argument");
  if (cl->n_clusters < 0 || cl->n_hdr < 2 || !cl->hdr_pos) return fail(ctx, LRA_B200_EINVAL, "refine_clusters_batch_device: bad cluster batch");
  if (rf->n_seq != rr->n_seq) return fail(ctx, LRA_B200_EINVAL, "refine_clusters_batch_device: the two read images differ in their number of reads");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_anchors = 0; res->n_units = 0; res->n_tasks = 0;
  if (cl->n_clusters == 0) return LRA_B200_OK;
  return lref_run(ctx, gl, rf, rr, cl, (size_t)n_anchors_in, res);
}

// host-buffer front end shared by REFINEclusters (hchain == NULL) and Refine_splitchain (hchain: HOST pointers)
static int lref_host_batch(lra_b200_ctx *ctx, const lra_b200_lindex *gl, const lra_b200_lindex *rf, const lra_b200_lindex *rr,
                           const lra_b200_clusters *cl, lra_b200_refined *res, const LrefChainExtra *hchain) {
  if (!ctx || !gl || !rf || !rr || !cl || !res) return fail(ctx, LRA_B200_EINVAL, "refine_clusters_batch: NULL argument");
  const int n = cl->n_clusters;
  if (n < 0 || cl->n_hdr < 2 || !cl->hdr_pos) return fail(ctx, LRA_B200_EINVAL, "refine_clusters_batch: bad cluster batch");
  if (rf->n_seq != rr->n_seq) return fail(ctx, LRA_B200_EINVAL, "refine_clusters_batch: the two read images differ in their number of reads");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_anchors = 0; res->n_units = 0; res->n_tasks = 0;
  if (n == 0) { if (res->r_off) res->r_off[0] = 0; return LRA_B200_OK; }
  for (int c = 0; c < n; c++)
    if (cl->read_id[c] >= (uint32_t)rf->n_seq) return fail(ctx, LRA_B200_EINVAL, "refine_clusters_batch: cluster %d names read %u of %d", c, cl->read_id[c], rf->n_seq);
  const size_t M = (size_t)cl->m_off[n];
  int rc;
  DevBuf *B = ctx->lr;
  const size_t acap = (size_t)(res->anchor_cap ? res->anchor_cap : 1);
  if ((rc = ensure(ctx, B[1], M * 4 + 16)) || (rc = ensure(ctx, B[2], M * 4 + 16)) || (rc = ensure(ctx, B[3], ((size_t)n + 1) * 8)) ||
      (rc = ensure(ctx, B[4], (size_t)n * 16)) || (rc = ensure(ctx, B[5], (size_t)n)) || (rc = ensure(ctx, B[6], (size_t)n * 4)) ||
      (rc = ensure(ctx, B[7], (size_t)cl->n_hdr * 8)) || (rc = ensure(ctx, B[13], (size_t)n * 4)) || (rc = ensure(ctx, B[14], (size_t)n * 4)) ||
      (rc = ensure(ctx, B[15], (size_t)n * 16)) || (rc = ensure(ctx, B[24], acap * 4)) || (rc = ensure(ctx, B[25], acap * 4)) ||
      (rc = ensure(ctx, B[26], acap * 4)) || (rc = ensure(ctx, B[27], ((size_t)n + 1) * 8)) || (rc = ensure(ctx, B[28], (size_t)n * 16)) ||
      (rc = ensure(ctx, B[29], (size_t)n * 4)) || (rc = ensure(ctx, B[30], M * 4 + 16)) || (rc = ensure(ctx, B[31], M * 4 + 16)) ||
      (rc = ensure(ctx, B[10], (size_t)n * 16)))
    return rc;
  cudaStream_t st = ctx->stream;
  LrefChainExtra dchain;
  if (hchain) {
    if ((rc = ensure(ctx, ctx->lr_x[2], M * 4 + 16)) || (rc = ensure(ctx, ctx->lr_x[3], M + 16)) || (rc = ensure(ctx, ctx->lr_x[4], (size_t)n * 4))) return rc;
    for (int c = 0; c < n; c++)
      if (cl->m_off[c + 1] > cl->m_off[c] && (hchain->chrom[c] < 0 || hchain->chrom[c] + 1 >= cl->n_hdr))
        return fail(ctx, LRA_B200_EINVAL, "refine_splitchains_batch: chain %d names contig %d of %d", c, hchain->chrom[c], cl->n_hdr - 1);
    if (M) { CU(cudaMemcpyAsync(ctx->lr_x[2].p, hchain->m_len, M * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(ctx->lr_x[3].p, hchain->m_strand, M, cudaMemcpyHostToDevice, st)); }
    CU(cudaMemcpyAsync(ctx->lr_x[4].p, hchain->chrom, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    dchain = LrefChainExtra{(const uint32_t *)ctx->lr_x[2].p, (const uint8_t *)ctx->lr_x[3].p, (const int32_t *)ctx->lr_x[4].p, hchain->limitrefine};
  }
  if (M) { CU(cudaMemcpyAsync(B[1].p, cl->m_q, M * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[2].p, cl->m_t, M * 4, cudaMemcpyHostToDevice, st)); }
  CU(cudaMemcpyAsync(B[3].p, cl->m_off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[4].p, cl->box, (size_t)n * 16, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[5].p, cl->strand, (size_t)n, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[6].p, cl->read_id, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[7].p, cl->hdr_pos, (size_t)cl->n_hdr * 8, cudaMemcpyHostToDevice, st));
  lra_b200_clusters dcl = *cl;
  dcl.m_q = (const uint32_t *)B[1].p; dcl.m_t = (const uint32_t *)B[2].p; dcl.m_off = (const uint64_t *)B[3].p; dcl.box = (const uint32_t *)B[4].p;
  dcl.strand = (const uint8_t *)B[5].p; dcl.read_id = (const uint32_t *)B[6].p; dcl.hdr_pos = (const uint64_t *)B[7].p;
  lra_b200_refined dres;
  memset(&dres, 0, sizeof dres);
  dres.status = (int32_t *)B[13].p; dres.chrom = (int32_t *)B[14].p; dres.diag = (int64_t *)B[15].p; dres.r_off = (uint64_t *)B[27].p;
  dres.r_q = (uint32_t *)B[24].p; dres.r_t = (uint32_t *)B[25].p; dres.r_tup = (uint32_t *)B[26].p; dres.anchor_cap = res->anchor_cap;
  dres.rbox = (uint32_t *)B[28].p; dres.eff = (float *)B[29].p; dres.m_q_out = (uint32_t *)B[30].p; dres.m_t_out = (uint32_t *)B[31].p;
  dres.box_out = (uint32_t *)B[10].p;
  rc = lref_run(ctx, gl, rf, rr, &dcl, M, &dres, hchain ? &dchain : nullptr);
  res->n_anchors = dres.n_anchors; res->n_units = dres.n_units; res->n_tasks = dres.n_tasks;
  if (rc != LRA_B200_OK && rc != LRA_B200_EOVERFLOW) return rc;
  CU(cudaMemcpyAsync(res->status, dres.status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->chrom, dres.chrom, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->diag, dres.diag, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->r_off, dres.r_off, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->rbox, dres.rbox, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->eff, dres.eff, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (res->box_out) CU(cudaMemcpyAsync(res->box_out, dres.box_out, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  if (M && res->m_q_out) CU(cudaMemcpyAsync(res->m_q_out, dres.m_q_out, M * 4, cudaMemcpyDeviceToHost, st));
  if (M && res->m_t_out) CU(cudaMemcpyAsync(res->m_t_out, dres.m_t_out, M * 4, cudaMemcpyDeviceToHost, st));
  const uint64_t n_out = dres.n_anchors;
  if (rc == LRA_B200_OK && n_out) {
    CU(cudaMemcpyAsync(res->r_q, dres.r_q, (size_t)n_out * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->r_t, dres.r_t, (size_t)n_out * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->r_tup, dres.r_tup, (size_t)n_out * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  return rc;
}

extern "C" int lra_b200_refine_clusters_batch(lra_b200_ctx *ctx, const lra_b200_lindex *gl, const lra_b200_lindex *rf, const lra_b200_lindex *rr,
                                              const lra_b200_clusters *cl, lra_b200_refined *res) {
  return lref_host_batch(ctx, gl, rf, rr, cl, res, nullptr);
}

static lra_b200_clusters chains_as_clusters(const lra_b200_splitchains *sc) {
  lra_b200_clusters cl;
  memset(&cl, 0, sizeof cl);
  cl.n_clusters = sc->n_chains; cl.m_q = sc->m_q; cl.m_t = sc->m_t; cl.m_off = sc->m_off; cl.box = sc->box; cl.strand = sc->strand; cl.read_id = sc->read_id;
  cl.hdr_pos = sc->hdr_pos; cl.n_hdr = sc->n_hdr; cl.global_k = sc->global_k; cl.small_k = sc->small_k; cl.window = sc->window; cl.local_max_freq = sc->local_max_freq;
  return cl;
}

extern "C" int lra_b200_refine_splitchains_batch(lra_b200_ctx *ctx, const lra_b200_lindex *gl, const lra_b200_lindex *rf, const lra_b200_lindex *rr,
                                                 const lra_b200_splitchains *sc, lra_b200_refined *res) {
  if (!sc || (sc->n_chains > 0 && (!sc->m_off || !sc->chrom || (sc->m_off[sc->n_chains] && (!sc->m_len || !sc->m_strand)))))
    return fail(ctx, LRA_B200_EINVAL, "refine_splitchains_batch: NULL argument");
  const lra_b200_clusters cl = chains_as_clusters(sc);
  const LrefChainExtra ex{sc->m_len, sc->m_strand, sc->chrom, sc->limitrefine};
  return lref_host_batch(ctx, gl, rf, rr, &cl, res, &ex);
}

extern "C" int lra_b200_refine_splitchains_batch_device(lra_b200_ctx *ctx, const lra_b200_lindex *gl, const lra_b200_lindex *rf, const lra_b200_lindex *rr,
                                                        const lra_b200_splitchains *sc, uint64_t n_anchors_in, lra_b200_refined *res) {
  if (!ctx || !gl || !rf || !rr || !sc || !res) return fail(ctx, LRA_B200_EINVAL, "refine_splitchains_batch_device: NULL argument");
  if (sc->n_chains < 0 || sc->n_hdr < 2 || !sc->hdr_pos) return fail(ctx, LRA_B200_EINVAL, "refine_splitchains_batch_device: bad chain batch");
  if (rf->n_seq != rr->n_seq) return fail(ctx, LRA_B200_EINVAL, "refine_splitchains_batch_device: the two read images differ in their number of reads");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_anchors = 0; res->n_units = 0; res->n_tasks = 0;
  if (sc->n_chains == 0) return LRA_B200_OK;
  const lra_b200_clusters cl = chains_as_clusters(sc);
  const LrefChainExtra ex{sc->m_len, sc->m_strand, sc->chrom, sc->limitrefine};
  return lref_run(ctx, gl, rf, rr, &cl, (size_t)n_anchors_in, res, &ex);
}

// ---------------------------------------------------------------------------------------------------- a6 anchor sorts
extern "C" int lra_b200_sort_matches_batch(lra_b200_ctx *ctx, int32_t mode, uint32_t *q, uint32_t *t, const uint64_t *seg_off, int32_t n_seg, uint32_t *perm) {
  if (!ctx || !seg_off || n_seg < 0 || mode < 0 || mode > 3) return fail(ctx, LRA_B200_EINVAL, "sort_matches_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (n_seg == 0) return LRA_B200_OK;
  const size_t N = (size_t)seg_off[n_seg];
  if (N == 0) return LRA_B200_OK;
  if (!q || !t) return fail(ctx, LRA_B200_EINVAL, "sort_matches_batch: NULL anchors");
  if (N > 0xFFFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "sort_matches_batch: more than 2^32 anchors in one batch");
  std::vector<unsigned long long> slot((size_t)n_seg);
  size_t slots = 0;
  for (int s = 0; s < n_seg; s++) {
    if (seg_off[s + 1] < seg_off[s]) return fail(ctx, LRA_B200_EINVAL, "sort_matches_batch: segment offsets not ascending");
    if (seg_off[s + 1] - seg_off[s] > (1ull << 30)) return fail(ctx, LRA_B200_EINVAL, "sort_matches_batch: segment %d has more than 2^30 anchors (the kernel pads a segment to a power of two in int)", s);
    const size_t n = (size_t)(seg_off[s + 1] - seg_off[s]);
    size_t P = 1; while (P < n) P <<= 1;
    slot[s] = slots;
    if (P > (size_t)kSortSmem) slots += P;
  }
  int rc;
  DevBuf *B = ctx->so;
  if ((rc = ensure(ctx, B[0], N * 4)) || (rc = ensure(ctx, B[1], N * 4)) || (rc = ensure(ctx, B[2], ((size_t)n_seg + 1) * 8)) || (rc = ensure(ctx, B[3], N * 4)) ||
      (rc = ensure(ctx, B[4], slots * 8 + 16)) || (rc = ensure(ctx, B[5], slots * 4 + 16)) || (rc = ensure(ctx, B[6], slots * 4 + 16)) || (rc = ensure(ctx, B[7], (size_t)n_seg * 8)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, q, N * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[1].p, t, N * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[2].p, seg_off, ((size_t)n_seg + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[7].p, slot.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
  SortBatch b;
  b.n_seg = n_seg; b.mode = mode; b.seg_off = (const unsigned long long *)B[2].p; b.q = (uint32_t *)B[0].p; b.t = (uint32_t *)B[1].p;
  b.perm = perm ? (uint32_t *)B[3].p : nullptr; b.kp = (unsigned long long *)B[4].p; b.ks = (uint32_t *)B[5].p; b.ki = (uint32_t *)B[6].p;
  b.slot_off = (const unsigned long long *)B[7].p;
  cudaEventRecord(ctx->ev[0], st);
  sort_pairs_kernel<<<(unsigned)n_seg, 256, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(q, b.q, N * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(t, b.t, N * 4, cudaMemcpyDeviceToHost, st));
  if (perm) CU(cudaMemcpyAsync(perm, b.perm, N * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "sort_pairs<mode=%d>", mode);
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)n_seg; s2.algo_bytes = 16ull * N + (perm ? 4ull * N : 0ull) + 8ull * (uint64_t)n_seg;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a24 GlobalChain
extern "C" int lra_b200_global_chain_batch(lra_b200_ctx *ctx, const int32_t *frag, const uint64_t *frag_off, int32_t n_prob, int32_t *score, int32_t *prev,
                                           int32_t *chain, int32_t *chain_len) {
  if (!ctx || !frag_off || n_prob < 0 || !chain_len) return fail(ctx, LRA_B200_EINVAL, "global_chain_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (n_prob == 0) return LRA_B200_OK;
  const size_t N = (size_t)frag_off[n_prob];
  for (int p = 0; p < n_prob; p++) {
    if (frag_off[p + 1] < frag_off[p]) return fail(ctx, LRA_B200_EINVAL, "global_chain_batch: problem offsets not ascending");
    if (frag_off[p + 1] - frag_off[p] > (1ull << 28)) return fail(ctx, LRA_B200_EINVAL, "global_chain_batch: problem %d has more than 2^28 fragments", p);
  }
  if (N && (!frag || !score || !prev || !chain)) return fail(ctx, LRA_B200_EINVAL, "global_chain_batch: NULL fragment arrays");
  int rc;
  DevBuf *B = ctx->gc;
  if ((rc = ensure(ctx, B[0], N * 16 + 16)) || (rc = ensure(ctx, B[1], ((size_t)n_prob + 1) * 8)) || (rc = ensure(ctx, B[2], N * 4 + 16)) || (rc = ensure(ctx, B[3], N * 4 + 16)) ||
      (rc = ensure(ctx, B[4], N * 4 + 16)) || (rc = ensure(ctx, B[5], (size_t)n_prob * 4)) || (rc = ensure(ctx, B[6], 2 * N * sizeof(GcEndpoint) + 64)) ||
      (rc = ensure(ctx, B[7], 4 * N * sizeof(GcVertex) + 64)))
    return rc;
  cudaStream_t st = ctx->stream;
  if (N) { CU(cudaMemcpyAsync(B[0].p, frag, N * 16, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[2].p, score, N * 4, cudaMemcpyHostToDevice, st)); }
  CU(cudaMemcpyAsync(B[1].p, frag_off, ((size_t)n_prob + 1) * 8, cudaMemcpyHostToDevice, st));
  GcBatch b;
  b.n_prob = n_prob; b.frag_off = (const unsigned long long *)B[1].p; b.frag = (const int32_t *)B[0].p; b.score = (int32_t *)B[2].p; b.prev = (int32_t *)B[3].p;
  b.chain = (int32_t *)B[4].p; b.chain_len = (int32_t *)B[5].p; b.ep = (GcEndpoint *)B[6].p; b.tree = (GcVertex *)B[7].p;
  cudaEventRecord(ctx->ev[0], st);
  gchain_kernel<<<(unsigned)((n_prob + 63) / 64), 64, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  if (N) {
    CU(cudaMemcpyAsync(score, b.score, N * 4, cudaMemcpyDeviceToHost, st)); CU(cudaMemcpyAsync(prev, b.prev, N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(chain, b.chain, N * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaMemcpyAsync(chain_len, b.chain_len, (size_t)n_prob * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "gchain");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)n_prob; s2.algo_bytes = 32ull * N;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a20 RefineBreakpoint
extern "C" int lra_b200_refine_breakpoint_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads_fwd, const lra_b200_seq *reads_rc, const lra_b200_seq *genome,
                                                const lra_b200_breakpoints *bp, lra_b200_breakpoint_result *res) {
  if (!ctx || !reads_fwd || !reads_rc || !genome || !bp || !res) return fail(ctx, LRA_B200_EINVAL, "refine_breakpoint_batch: NULL argument");
  const int n = bp->n_pairs;
  if (n < 0) return fail(ctx, LRA_B200_EINVAL, "refine_breakpoint_batch: negative pair count");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (n == 0) return LRA_B200_OK;
  for (int p = 0; p < n; p++) {
    if (bp->read_off[p] + bp->read_len[p] > reads_fwd->n || bp->read_off[p] + bp->read_len[p] > reads_rc->n)
      return fail(ctx, LRA_B200_EINVAL, "refine_breakpoint_batch: read of pair %d ends beyond the arena", p);
    if (bp->lchrom_off[p] + bp->lchrom_len[p] > genome->n || bp->rchrom_off[p] + bp->rchrom_len[p] > genome->n)
      return fail(ctx, LRA_B200_EINVAL, "refine_breakpoint_batch: contig of pair %d ends beyond the genome arena", p);
  }
  const int slabs = n < 512 ? n : 512;
  int rc;
  DevBuf *B = ctx->rb;
  const size_t N = (size_t)n;
  if ((rc = ensure(ctx, B[0], N * 12)) || (rc = ensure(ctx, B[1], N * 12)) || (rc = ensure(ctx, B[2], N * 12)) || (rc = ensure(ctx, B[3], N * 12)) ||
      (rc = ensure(ctx, B[4], N)) || (rc = ensure(ctx, B[5], N)) || (rc = ensure(ctx, B[6], N * 8)) || (rc = ensure(ctx, B[7], N * 4)) ||
      (rc = ensure(ctx, B[8], N * 8)) || (rc = ensure(ctx, B[9], N * 8)) || (rc = ensure(ctx, B[10], N * 4)) || (rc = ensure(ctx, B[11], N * 4)) ||
      (rc = ensure(ctx, B[12], (size_t)slabs * 2 * kRbpCells * 4)) || (rc = ensure(ctx, B[13], (size_t)slabs * 2 * kRbpCells)) ||
      (rc = ensure(ctx, B[14], (size_t)slabs * 8 * 512 * 4)) || (rc = ensure(ctx, B[15], N * 8)) || (rc = ensure(ctx, B[16], N * 8)) ||
      (rc = ensure(ctx, B[17], N * 24)) || (rc = ensure(ctx, B[18], N * 2 * kRbpCap * 12)) || (rc = ensure(ctx, B[19], N * 4)))
    return rc;
  cudaStream_t st = ctx->stream;
  const void *src[12] = {bp->lf, bp->ll, bp->rf, bp->rl, bp->lstrand, bp->rstrand, bp->read_off, bp->read_len, bp->lchrom_off, bp->rchrom_off, bp->lchrom_len, bp->rchrom_len};
  const size_t sz[12] = {N * 12, N * 12, N * 12, N * 12, N, N, N * 8, N * 4, N * 8, N * 8, N * 4, N * 4};
  for (int i = 0; i < 12; i++) CU(cudaMemcpyAsync(B[i].p, src[i], sz[i], cudaMemcpyHostToDevice, st));
  RbpBatch b;
  b.n_pairs = n;
  b.reads_fwd = SeqView{reads_fwd->b2, reads_fwd->nm, reads_fwd->n}; b.reads_rc = SeqView{reads_rc->b2, reads_rc->nm, reads_rc->n};
  b.genome = SeqView{genome->b2, genome->nm, genome->n};
  b.lf = (const uint32_t *)B[0].p; b.ll = (const uint32_t *)B[1].p; b.rf = (const uint32_t *)B[2].p; b.rl = (const uint32_t *)B[3].p;
  b.lstrand = (const uint8_t *)B[4].p; b.rstrand = (const uint8_t *)B[5].p; b.read_off = (const unsigned long long *)B[6].p; b.read_len = (const uint32_t *)B[7].p;
  b.lchrom_off = (const unsigned long long *)B[8].p; b.rchrom_off = (const unsigned long long *)B[9].p; b.lchrom_len = (const uint32_t *)B[10].p; b.rchrom_len = (const uint32_t *)B[11].p;
  b.score = (int32_t *)B[12].p; b.path = (uint8_t *)B[13].p; b.walk = (int32_t *)B[14].p;
  b.mode = (int32_t *)B[15].p; b.n_out = (int32_t *)B[16].p; b.bound = (uint32_t *)B[17].p; b.out = (uint32_t *)B[18].p; b.refined = (int32_t *)B[19].p;
  cudaEventRecord(ctx->ev[0], st);
  for (int first = 0; first < n; first += slabs) {
    const int here = n - first < slabs ? n - first : slabs;
    rbp_kernel<<<(unsigned)((here + 3) / 4), 128, 0, st>>>(b, first, here);
    ctx->launches++;
  }
  cudaEventRecord(ctx->ev[1], st);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->mode, b.mode, N * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->n_out, b.n_out, N * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->bound, b.bound, N * 24, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->out, b.out, N * 2 * kRbpCap * 12, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->refined, b.refined, N * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "rbp");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)n; s2.algo_bytes = 128ull * N;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a16 chain filters
extern "C" int lra_b200_chain_filter_batch(lra_b200_ctx *ctx, int32_t mode, const uint32_t *q, const uint32_t *t, const uint32_t *len, const uint8_t *strand,
                                           const uint64_t *chain_off, int32_t n_chains, uint8_t *keep) {
  if (!ctx || !chain_off || n_chains < 0 || mode < 0 || mode > 5) return fail(ctx, LRA_B200_EINVAL, "chain_filter_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (n_chains == 0) return LRA_B200_OK;
  const size_t N = (size_t)chain_off[n_chains];
  if (N == 0) return LRA_B200_OK;
  if (!q || !t || !len || !keep || (mode != 3 && !strand)) return fail(ctx, LRA_B200_EINVAL, "chain_filter_batch: NULL anchor arrays");
  for (int c = 0; c < n_chains; c++)
    if (chain_off[c + 1] < chain_off[c] || chain_off[c + 1] - chain_off[c] > 0x7FFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "chain_filter_batch: bad chain offsets");
  int rc;
  DevBuf *B = ctx->cf;
  if ((rc = ensure(ctx, B[0], N * 4)) || (rc = ensure(ctx, B[1], N * 4)) || (rc = ensure(ctx, B[2], N * 4)) || (rc = ensure(ctx, B[3], N + 16)) ||
      (rc = ensure(ctx, B[4], ((size_t)n_chains + 1) * 8)) || (rc = ensure(ctx, B[5], N + 16)) || (rc = ensure(ctx, B[6], N * 4)) || (rc = ensure(ctx, B[7], N * 4)) ||
      (rc = ensure(ctx, B[8], N * 4)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, q, N * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[1].p, t, N * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[2].p, len, N * 4, cudaMemcpyHostToDevice, st));
  if (strand) CU(cudaMemcpyAsync(B[3].p, strand, N, cudaMemcpyHostToDevice, st)); else CU(cudaMemsetAsync(B[3].p, 0, N, st));
  CU(cudaMemcpyAsync(B[4].p, chain_off, ((size_t)n_chains + 1) * 8, cudaMemcpyHostToDevice, st));
  ChainfBatch b{n_chains, mode, (const unsigned long long *)B[4].p, (const uint32_t *)B[0].p, (const uint32_t *)B[1].p, (const uint32_t *)B[2].p, (const uint8_t *)B[3].p,
                (uint8_t *)B[5].p, (int32_t *)B[6].p, (int32_t *)B[7].p, (int32_t *)B[8].p};
  cudaEventRecord(ctx->ev[0], st);
  chainf_kernel<<<(unsigned)((n_chains + 127) / 128), 128, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(keep, b.keep, N, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "chainf<mode=%d>", mode);
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)n_chains; s2.algo_bytes = 14ull * N;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a7 CleanOffDiagonal
extern "C" int lra_b200_clean_off_diagonal_batch(lra_b200_ctx *ctx, const lra_b200_anchor_lists *al, const lra_b200_clean_opts *opts, lra_b200_clean_result *res) {
  if (!ctx || !al || !opts || !res || al->n_lists < 0 || !al->list_off) return fail(ctx, LRA_B200_EINVAL, "clean_off_diagonal_batch: bad argument");
  const int n = al->n_lists;
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (n == 0) return LRA_B200_OK;
  const size_t N = (size_t)al->list_off[n];
  for (int s = 0; s < n; s++)
    if (al->list_off[s + 1] < al->list_off[s] || al->list_off[s + 1] - al->list_off[s] > 0x3FFFFFFFull) return fail(ctx, LRA_B200_EINVAL, "clean_off_diagonal_batch: bad list offsets");
  if (N && (!al->q || !al->t || !al->qt)) return fail(ctx, LRA_B200_EINVAL, "clean_off_diagonal_batch: NULL anchors");
  if (opts->bypassClustering && opts->ExtractDiagonalFromClean && (!al->hdr_pos || al->n_hdr < 1)) return fail(ctx, LRA_B200_EINVAL, "clean_off_diagonal_batch: genome header needed");
  int rc;
  DevBuf *B = ctx->cd;
  const size_t H = al->n_hdr > 0 ? (size_t)al->n_hdr : 1;
  if ((rc = ensure(ctx, B[0], N * 4 + 16)) || (rc = ensure(ctx, B[1], N * 4 + 16)) || (rc = ensure(ctx, B[2], N * 8 + 16)) || (rc = ensure(ctx, B[3], (size_t)n)) ||
      (rc = ensure(ctx, B[4], ((size_t)n + 1) * 8)) || (rc = ensure(ctx, B[5], H * 8)) || (rc = ensure(ctx, B[6], N + 16)) || (rc = ensure(ctx, B[7], N * 4 + 16)) ||
      (rc = ensure(ctx, B[8], N * 4 + 16)) || (rc = ensure(ctx, B[9], N * 28 + 64)) || (rc = ensure(ctx, B[10], N * 4 + 16)) || (rc = ensure(ctx, B[11], (size_t)n * 4)) ||
      (rc = ensure(ctx, B[12], N * 3 + 64)) || (rc = ensure(ctx, B[13], N * 32 + 64)) || (rc = ensure(ctx, B[14], N * 4 + 64)))
    return rc;
  cudaStream_t st = ctx->stream;
  if (N) {
    CU(cudaMemcpyAsync(B[0].p, al->q, N * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[1].p, al->t, N * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[2].p, al->qt, N * 8, cudaMemcpyHostToDevice, st));
  }
  CU(cudaMemcpyAsync(B[3].p, al->strand, (size_t)n, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[4].p, al->list_off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
  if (al->n_hdr > 0) CU(cudaMemcpyAsync(B[5].p, al->hdr_pos, (size_t)al->n_hdr * 8, cudaMemcpyHostToDevice, st));
  CodBatch b;
  b.n_lists = n; b.off = (const unsigned long long *)B[4].p; b.q = (const uint32_t *)B[0].p; b.t = (const uint32_t *)B[1].p; b.qt = (const unsigned long long *)B[2].p;
  b.strand = (const uint8_t *)B[3].p;
  b.o = CodOpts{opts->cleanMaxDiag, opts->minDiagCluster, opts->bypassClustering, opts->cleanClustersize, opts->SecondCleanMinDiagCluster, opts->punish_anchorfreq,
                opts->anchorPerlength, opts->SecondCleanMaxDiag, opts->ExtractDiagonalFromClean, opts->globalK};
  b.hdr_pos = (const unsigned long long *)B[5].p; b.n_hdr = al->n_hdr;
  b.keep = (uint8_t *)B[6].p; b.freq = (float *)B[7].p; b.cnt = (int32_t *)B[8].p; b.cl = (int32_t *)B[9].p; b.cl_freq = (float *)B[10].p; b.n_cl = (int32_t *)B[11].p;
  b.flags = (uint8_t *)B[12].p; b.hkeys = (unsigned long long *)B[13].p; b.hused = (uint8_t *)B[14].p;
  cudaEventRecord(ctx->ev[0], st);
  cod_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  if (N) {
    CU(cudaMemcpyAsync(res->keep, b.keep, N, cudaMemcpyDeviceToHost, st)); CU(cudaMemcpyAsync(res->freq, b.freq, N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->cnt, b.cnt, N * 4, cudaMemcpyDeviceToHost, st));
    if (res->cl) CU(cudaMemcpyAsync(res->cl, b.cl, N * 28, cudaMemcpyDeviceToHost, st));
    if (res->cl_freq) CU(cudaMemcpyAsync(res->cl_freq, b.cl_freq, N * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaMemcpyAsync(res->n_cl, b.n_cl, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "cod");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)n; s2.algo_bytes = 25ull * N;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a9 SplitClusters
extern "C" int lra_b200_split_clusters_batch(lra_b200_ctx *ctx, const lra_b200_read_clusters *rc_, lra_b200_split_result *res) {
  if (!ctx || !rc_ || !res || rc_->n_reads < 0 || !rc_->cl_off) return fail(ctx, LRA_B200_EINVAL, "split_clusters_batch: bad argument");
  const int R = rc_->n_reads;
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_pieces = 0;
  if (R == 0) return LRA_B200_OK;
  const size_t Cn = (size_t)rc_->cl_off[R];
  if (Cn == 0) { for (int r = 0; r <= R; r++) res->sp_off[r] = 0; return LRA_B200_OK; }
  if (!rc_->box || !rc_->strand || !rc_->freq || !rc_->m_off) return fail(ctx, LRA_B200_EINVAL, "split_clusters_batch: NULL cluster arrays");
  const size_t M = (size_t)rc_->m_off[Cn];
  if (M && !rc_->m_q) return fail(ctx, LRA_B200_EINVAL, "split_clusters_batch: NULL anchors");
  int rc;
  DevBuf *B = ctx->sc;
  const size_t cap = (size_t)(res->piece_cap ? res->piece_cap : 1);
  if ((rc = ensure(ctx, B[0], ((size_t)R + 1) * 8)) || (rc = ensure(ctx, B[1], Cn * 16)) || (rc = ensure(ctx, B[2], Cn)) || (rc = ensure(ctx, B[3], Cn * 4)) ||
      (rc = ensure(ctx, B[4], (Cn + 1) * 8)) || (rc = ensure(ctx, B[5], M * 4 + 16)) || (rc = ensure(ctx, B[6], Cn)) || (rc = ensure(ctx, B[7], Cn * 4)) ||
      (rc = ensure(ctx, B[8], ((size_t)R + 2) * 8)) || (rc = ensure(ctx, B[9], cap * 24)) || (rc = ensure(ctx, B[10], cap * 4)) || (rc = ensure(ctx, B[11], cap * 4)) ||
      (rc = ensure(ctx, B[12], Cn * 16 + 64)) || (rc = ensure(ctx, B[13], Cn * 4 * sizeof(ScPoint) + 64)) || (rc = ensure(ctx, B[14], 64)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, rc_->cl_off, ((size_t)R + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[1].p, rc_->box, Cn * 16, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[2].p, rc_->strand, Cn, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[3].p, rc_->freq, Cn * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[4].p, rc_->m_off, (Cn + 1) * 8, cudaMemcpyHostToDevice, st));
  if (M) CU(cudaMemcpyAsync(B[5].p, rc_->m_q, M * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(B[14].p, 0, 64, st));
  SplitBatch b;
  b.n_reads = R; b.contig = rc_->contig; b.globalK = rc_->global_k;
  b.cl_off = (const unsigned long long *)B[0].p; b.box = (const uint32_t *)B[1].p; b.strand = (const uint8_t *)B[2].p; b.freq = (const float *)B[3].p;
  b.m_off = (const unsigned long long *)B[4].p; b.mq = (const uint32_t *)B[5].p; b.split = (uint8_t *)B[6].p; b.val_cluster = (int32_t *)B[7].p;
  b.sp_off = (unsigned long long *)B[8].p; b.sp = (uint32_t *)B[9].p; b.sp_val = (int32_t *)B[10].p; b.sp_n0 = (int32_t *)B[11].p; b.sp_cap = res->piece_cap;
  b.sets = (uint32_t *)B[12].p; b.pts = (ScPoint *)B[13].p;
  cudaEventRecord(ctx->ev[0], st);
  split_kernel<false><<<(unsigned)((R + 63) / 64), 64, 0, st>>>(b);
  seed_scan_kernel<<<1, 1024, 0, st>>>(b.sp_off, R, res->piece_cap, (int *)B[14].p);
  split_kernel<true><<<(unsigned)((R + 63) / 64), 64, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches += 3;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->sp_off, b.sp_off, ((size_t)R + 1) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->split, b.split, Cn, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->val_cluster, b.val_cluster, Cn * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const uint64_t total = res->sp_off[R];
  res->n_pieces = total;
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "split(count+scan+emit)");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)R; s2.algo_bytes = 30ull * Cn + 4ull * M + 32ull * total;
  ctx->stats.push_back(s2);
  if (total > res->piece_cap) return fail(ctx, LRA_B200_EOVERFLOW, "split_clusters_batch: piece capacity %llu too small, %llu needed", (unsigned long long)res->piece_cap, (unsigned long long)total);
  if (total) {
    CU(cudaMemcpyAsync(res->sp, b.sp, total * 24, cudaMemcpyDeviceToHost, st)); CU(cudaMemcpyAsync(res->sp_val, b.sp_val, total * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->sp_n0, b.sp_n0, total * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a22 ordering + MAPQ
extern "C" int lra_b200_mapq_batch(lra_b200_ctx *ctx, const lra_b200_alignment_groups *ag, lra_b200_mapq_result *res) {
  if (!ctx || !ag || !res || ag->n_reads < 0 || !ag->grp_off) return fail(ctx, LRA_B200_EINVAL, "mapq_batch: bad argument");
  const int R = ag->n_reads;
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (R == 0) return LRA_B200_OK;
  const size_t G = (size_t)ag->grp_off[R];
  if (G == 0) return LRA_B200_OK;
  if (!ag->seg_off || !ag->upd_off) return fail(ctx, LRA_B200_EINVAL, "mapq_batch: NULL offsets");
  const size_t S = (size_t)ag->seg_off[G], U = (size_t)ag->upd_off[R];
  // the two logf terms of SimpleMapQV, with the host libm (Mapping_ultility.h:527,563,576)
  std::vector<float> logv(S + 1);
  for (size_t s = 0; s < S; s++) logv[s] = ag->value[s] > 3 ? logf(ag->value[s] / ag->global_k) : 0;
  std::vector<int32_t> lenpen((size_t)R);
  for (int r = 0; r < R; r++) {
    int len = 0, old = 0;
    for (int u = ag->upd_off[r]; u < ag->upd_off[r + 1]; u++) {
      const int c = ag->update_at[u];
      if (c < old || c > ag->grp_off[r + 1] - ag->grp_off[r]) return fail(ctx, LRA_B200_EINVAL, "mapq_batch: update schedule of read %d not ascending / beyond its alignments", r);
      old = c; len = c;
    }
    lenpen[r] = len > 0 ? (int)(4.343f * logf(len) + .499f) : 0;
  }
  int rc;
  DevBuf *B = ctx->mq;
  const size_t sz4 = S * 4 + 16;
  if ((rc = ensure(ctx, B[0], ((size_t)R + 1) * 4)) || (rc = ensure(ctx, B[1], (G + 1) * 4)) || (rc = ensure(ctx, B[2], ((size_t)R + 1) * 4)) || (rc = ensure(ctx, B[3], U * 4 + 16)))
    return rc;
  for (int i = 4; i <= 12; i++) if ((rc = ensure(ctx, B[i], sz4))) return rc;       // value n0 n1 nm nmm ndel nins logv flag
  if ((rc = ensure(ctx, B[13], sz4)) || (rc = ensure(ctx, B[14], S + 16)) || (rc = ensure(ctx, B[15], S + 16)) || (rc = ensure(ctx, B[16], S + 16)) || (rc = ensure(ctx, B[17], sz4)) ||
      (rc = ensure(ctx, B[18], (size_t)R * 4)) || (rc = ensure(ctx, B[19], G + 16)) || (rc = ensure(ctx, B[20], G * 4 + 16)) || (rc = ensure(ctx, B[21], G * 4 + 16)) ||
      (rc = ensure(ctx, B[22], G * 4 + 16)) || (rc = ensure(ctx, B[23], G * 16 + 16)) || (rc = ensure(ctx, B[24], G * 4 + 16)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, ag->grp_off, ((size_t)R + 1) * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[1].p, ag->seg_off, (G + 1) * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[2].p, ag->upd_off, ((size_t)R + 1) * 4, cudaMemcpyHostToDevice, st));
  if (U) CU(cudaMemcpyAsync(B[3].p, ag->update_at, U * 4, cudaMemcpyHostToDevice, st));
  const void *src4[9] = {ag->value, ag->n0, ag->n1, ag->nm, ag->nmm, ag->ndel, ag->nins, logv.data(), res->flag};
  if (S) {
    for (int i = 0; i < 9; i++) CU(cudaMemcpyAsync(B[4 + i].p, src4[i], S * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[13].p, res->typeofaln, S * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[14].p, ag->strand, S, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[15].p, res->issec, S, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[16].p, res->supp, S, cudaMemcpyHostToDevice, st));
  }
  CU(cudaMemcpyAsync(B[18].p, lenpen.data(), (size_t)R * 4, cudaMemcpyHostToDevice, st));
  MapqBatch b;
  b.n_reads = R; b.bypass = ag->bypass_clustering; b.read_type = ag->read_type;
  b.grp_off = (const int32_t *)B[0].p; b.seg_off = (const int32_t *)B[1].p; b.upd_off = (const int32_t *)B[2].p; b.update_at = (const int32_t *)B[3].p;
  b.value = (const float *)B[4].p; b.n0 = (const int32_t *)B[5].p; b.n1 = (const int32_t *)B[6].p; b.nm = (const int32_t *)B[7].p; b.nmm = (const int32_t *)B[8].p;
  b.ndel = (const int32_t *)B[9].p; b.nins = (const int32_t *)B[10].p; b.logv = (const float *)B[11].p; b.flag = (int32_t *)B[12].p; b.typeofaln = (int32_t *)B[13].p;
  b.strand = (const uint8_t *)B[14].p; b.issec = (uint8_t *)B[15].p; b.supp = (uint8_t *)B[16].p; b.mapq = (int32_t *)B[17].p; b.lenpen = (const int32_t *)B[18].p;
  b.g_issec = (uint8_t *)B[19].p; b.g_value = (float *)B[20].p; b.g_n0 = (int32_t *)B[21].p; b.g_n1 = (int32_t *)B[22].p; b.g_nm = (int32_t *)B[23].p; b.order = (int32_t *)B[24].p;
  cudaEventRecord(ctx->ev[0], st);
  mapq_kernel<<<(unsigned)((R + 63) / 64), 64, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  if (S) {
    CU(cudaMemcpyAsync(res->flag, b.flag, S * 4, cudaMemcpyDeviceToHost, st)); CU(cudaMemcpyAsync(res->typeofaln, b.typeofaln, S * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->issec, b.issec, S, cudaMemcpyDeviceToHost, st)); CU(cudaMemcpyAsync(res->supp, b.supp, S, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->mapq, b.mapq, S * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaMemcpyAsync(res->g_issec, b.g_issec, G, cudaMemcpyDeviceToHost, st)); CU(cudaMemcpyAsync(res->g_value, b.g_value, G * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->g_n0, b.g_n0, G * 4, cudaMemcpyDeviceToHost, st)); CU(cudaMemcpyAsync(res->g_n1, b.g_n1, G * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->g_nm, b.g_nm, G * 16, cudaMemcpyDeviceToHost, st)); CU(cudaMemcpyAsync(res->order, b.order, G * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "mapq");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)R; s2.algo_bytes = 60ull * S + 40ull * G;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a15 LinearExtend (low-accuracy pipeline)
extern "C" int lra_b200_linear_extend_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_extend_parts *in,
                                            lra_b200_extended *res) {
  if (!ctx || !reads || !genome || !in || !res) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: NULL argument");
  const int G = in->n_groups;
  if (G < 0 || !in->g_off || in->K <= 0) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_total = 0;
  if (G == 0) { if (res->e_off) res->e_off[0] = 0; return LRA_B200_OK; }
  const size_t P = (size_t)in->g_off[G];
  if (P > 0x7FFFFFFFull) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: more than 2^31 parts");
  if (P && (!in->p_off || !in->p_strand || !in->chrom_off || !in->chrom_len || !in->read_off || !in->read_len))
    return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: NULL part array");
  const size_t N = P ? (size_t)in->p_off[P] : 0;
  if (N > 0x7FFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: more than 2^31 anchors in one batch");
  if (N && (!in->q || !in->t)) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: NULL anchors");
  if (res->cap < N) { res->n_total = N; return fail(ctx, LRA_B200_EOVERFLOW, "linear_extend_batch: result arrays hold %llu anchors, %llu may be needed",
                                                     (unsigned long long)res->cap, (unsigned long long)N); }
  for (int g = 0; g < G; g++) if (in->g_off[g + 1] < in->g_off[g]) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: group offsets not ascending");
  std::vector<unsigned long long> slot(P ? P : 1);
  size_t slots = 0;
  for (size_t p = 0; p < P; p++) {
    if (in->p_off[p + 1] < in->p_off[p]) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: part offsets not ascending");
    if (in->chrom_off[p] + in->chrom_len[p] > genome->n) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: contig of part %zu ends beyond the genome arena", p);
    if (in->read_off[p] + in->read_len[p] > reads->n) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: read of part %zu ends beyond the read arena", p);
    if (in->chrom_len[p] == 0) return fail(ctx, LRA_B200_EINVAL, "linear_extend_batch: part %zu lies on an empty contig", p);
    const size_t n = (size_t)(in->p_off[p + 1] - in->p_off[p]);
    size_t P2 = 1; while (P2 < n) P2 <<= 1;
    slot[p] = slots;
    if (!in->skipsorting && P2 > (size_t)kSortSmem) slots += P2;
  }
  int rc;
  DevBuf *B = ctx->le;
  const size_t Np = N ? N : 1, Pp = P ? P : 1;
  if ((rc = ensure(ctx, B[0], ((size_t)G + 1) * 8)) || (rc = ensure(ctx, B[1], (Pp + 1) * 8)) || (rc = ensure(ctx, B[2], Pp)) || (rc = ensure(ctx, B[3], Pp * 8)) ||
      (rc = ensure(ctx, B[4], Pp * 4)) || (rc = ensure(ctx, B[5], Pp * 8)) || (rc = ensure(ctx, B[6], Pp * 4)) || (rc = ensure(ctx, B[7], Np * 4)) ||
      (rc = ensure(ctx, B[8], Np * 4)) || (rc = ensure(ctx, B[9], Np * 4)) || (rc = ensure(ctx, B[10], (Np + 2) * 8)) || (rc = ensure(ctx, B[11], Np * 4)) ||
      (rc = ensure(ctx, B[12], Np * 4)) || (rc = ensure(ctx, B[13], Np * 4)) || (rc = ensure(ctx, B[14], ((size_t)G + 1) * 8)) || (rc = ensure(ctx, B[15], Np * 4)) ||
      (rc = ensure(ctx, B[16], Np * 4)) || (rc = ensure(ctx, B[17], Np * 4)) || (rc = ensure(ctx, B[18], (size_t)G * 16)) || (rc = ensure(ctx, B[19], slots * 8 + 16)) ||
      (rc = ensure(ctx, B[20], slots * 4 + 16)) || (rc = ensure(ctx, B[21], slots * 4 + 16)) || (rc = ensure(ctx, B[22], Pp * 8)) || (rc = ensure(ctx, B[23], 16)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, in->g_off, ((size_t)G + 1) * 8, cudaMemcpyHostToDevice, st));
  if (P) {
    CU(cudaMemcpyAsync(B[1].p, in->p_off, (P + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[2].p, in->p_strand, P, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[3].p, in->chrom_off, P * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[4].p, in->chrom_len, P * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[5].p, in->read_off, P * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[6].p, in->read_len, P * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[22].p, slot.data(), P * 8, cudaMemcpyHostToDevice, st));
  } else {
    CU(cudaMemsetAsync(B[1].p, 0, 8, st));
  }
  if (N) {
    CU(cudaMemcpyAsync(B[7].p, in->q, N * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[8].p, in->t, N * 4, cudaMemcpyHostToDevice, st));
  }
  LextBatch b;
  b.n_groups = G; b.n_parts = (long long)P; b.N = N; b.K = in->K; b.trim = in->trim;
  b.reads = SeqView{reads->b2, reads->nm, reads->n}; b.genome = SeqView{genome->b2, genome->nm, genome->n};
  b.g_off = (const unsigned long long *)B[0].p; b.p_off = (const unsigned long long *)B[1].p; b.p_strand = (const uint8_t *)B[2].p;
  b.chrom_off = (const unsigned long long *)B[3].p; b.chrom_len = (const uint32_t *)B[4].p; b.read_off = (const unsigned long long *)B[5].p;
  b.read_len = (const uint32_t *)B[6].p; b.q = (const uint32_t *)B[7].p; b.t = (const uint32_t *)B[8].p; b.part_of = (uint32_t *)B[9].p;
  b.run = (unsigned long long *)B[10].p; b.end_q = (uint32_t *)B[11].p; b.end_t = (uint32_t *)B[12].p; b.lidx = (int *)B[13].p;
  b.e_off = (unsigned long long *)B[14].p; b.eq = (uint32_t *)B[15].p; b.et = (uint32_t *)B[16].p; b.elen = (int32_t *)B[17].p; b.box = (uint32_t *)B[18].p;
  auto rec = [&](int i) { cudaEventRecord(ctx->ev[i], st); };
  rec(0);
  if (!in->skipsorting && N) {                   // DiagonalSort<GenomeTuple> of every part (LinearExtend.h:663; Sorting.h:33-60)
    SortBatch sb;
    sb.n_seg = (int)P; sb.mode = 0; sb.seg_off = b.p_off; sb.q = (uint32_t *)B[7].p; sb.t = (uint32_t *)B[8].p; sb.perm = nullptr;
    sb.kp = (unsigned long long *)B[19].p; sb.ks = (uint32_t *)B[20].p; sb.ki = (uint32_t *)B[21].p; sb.slot_off = (const unsigned long long *)B[22].p;
    sort_pairs_kernel<<<(unsigned)P, 256, 0, st>>>(sb);
    ctx->launches++;
  }
  rec(1);
  if (N) {
    const unsigned blocks = (unsigned)((N + 255) / 256);
    lext_link_kernel<<<blocks, 256, 0, st>>>(b);
    CU(cudaMemsetAsync(b.run + N, 0, 16, st));
    seed_scan_kernel<<<1, 1024, 0, st>>>(b.run, (int)N, ~0ull, (int *)B[23].p);
    lext_emit_kernel<<<blocks, 256, 0, st>>>(b);
    ctx->launches += 3;
  } else {
    CU(cudaMemsetAsync(b.run, 0, 16, st));
  }
  rec(2);
  lext_group_kernel<<<(unsigned)((G + 3) / 4), 128, 0, st>>>(b);
  ctx->launches++;
  rec(3);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->e_off, b.e_off, ((size_t)G + 1) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->box, b.box, (size_t)G * 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const size_t n_out = (size_t)res->e_off[G];
  res->n_total = n_out;
  if (n_out) {
    CU(cudaMemcpyAsync(res->q, b.eq, n_out * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->t, b.et, n_out * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->len, b.elen, n_out * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  const char *names[3] = {"lext_sort", "lext_link+scan+emit", "lext_group"};
  for (int i = 0; i < 3; i++) {
    if (i == 0 && (in->skipsorting || !N)) continue;
    lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "%s", names[i]);
    cudaEventElapsedTime(&s2.ms, ctx->ev[i], ctx->ev[i + 1]); s2.jobs = i == 2 ? (uint64_t)G : (uint64_t)N;
    s2.algo_bytes = i == 0 ? 16ull * N : i == 1 ? 8ull * N + 12ull * n_out : 12ull * n_out + 16ull * (uint64_t)G;
    ctx->stats.push_back(s2);
  }
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a15 LinearExtend (high-accuracy overload)
extern "C" int lra_b200_linear_extend_chains_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_extend_chains *in,
                                                   lra_b200_extended_chains *res) {
  if (!ctx || !reads || !genome || !in || !res) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: NULL argument");
  const int NC = in->n_chains, CL = in->n_clusters;
  if (NC < 0 || CL < 0 || !in->ch_off || in->K <= 0) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_total = 0;
  const size_t U = NC ? (size_t)in->ch_off[NC] : 0;
  if (U == 0) { if (res->e_off) res->e_off[0] = 0; return LRA_B200_OK; }
  if (U > 0x7FFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: more than 2^31 chain entries");
  if (!in->ch || !in->cl_off || !in->cl_box || !in->cl_strand || !in->cl_freq || !in->chrom_off || !in->chrom_len || !in->read_off || !in->read_len || CL == 0)
    return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: NULL cluster array");
  const size_t N = (size_t)in->cl_off[CL];
  if (N > 0x7FFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: more than 2^31 anchors in one batch");
  if (N && (!in->q || !in->t)) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: NULL anchors");
  std::vector<unsigned long long> slot((size_t)CL), slot_off(U + 1);
  std::vector<uint8_t> edge(U, 0);
  size_t slots = 0;
  for (int c = 0; c < CL; c++) {
    if (in->cl_off[c + 1] < in->cl_off[c]) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: cluster offsets not ascending");
    if (in->chrom_off[c] + in->chrom_len[c] > genome->n) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: contig of cluster %d ends beyond the genome arena", c);
    if (in->read_off[c] + in->read_len[c] > reads->n) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: read of cluster %d ends beyond the read arena", c);
    if (in->chrom_len[c] == 0) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: cluster %d lies on an empty contig", c);
    if (in->cl_strand[c] > 1) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: strand of cluster %d is not 0 / 1", c);
    const size_t n = (size_t)(in->cl_off[c + 1] - in->cl_off[c]);
    size_t P2 = 1; while (P2 < n) P2 <<= 1;
    slot[c] = slots;
    if (P2 > (size_t)kSortSmem) slots += P2;
  }
  size_t S = 0;
  for (int k = 0; k < NC; k++) {
    if (in->ch_off[k + 1] < in->ch_off[k]) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: chain offsets not ascending");
    for (size_t u = (size_t)in->ch_off[k]; u < (size_t)in->ch_off[k + 1]; u++) {
      if (in->ch[u] >= (uint32_t)CL) return fail(ctx, LRA_B200_EINVAL, "linear_extend_chains_batch: chain entry %zu names cluster %u of %d", u, in->ch[u], CL);
      edge[u] = (uint8_t)((u == (size_t)in->ch_off[k] ? 1 : 0) | (u + 1 == (size_t)in->ch_off[k + 1] ? 2 : 0));
      slot_off[u] = S;
      S += (size_t)(in->cl_off[in->ch[u] + 1] - in->cl_off[in->ch[u]]);
    }
  }
  slot_off[U] = S;
  if (res->cap < S) { res->n_total = S; return fail(ctx, LRA_B200_EOVERFLOW, "linear_extend_chains_batch: result arrays hold %llu anchors, %llu may be needed",
                                                     (unsigned long long)res->cap, (unsigned long long)S); }
  int rc;
  DevBuf *B = ctx->lc;
  const size_t Np = N ? N : 1, Sp = S ? S : 1, C1 = (size_t)CL;
  if ((rc = ensure(ctx, B[0], U * 4)) || (rc = ensure(ctx, B[1], U)) || (rc = ensure(ctx, B[2], (U + 1) * 8)) || (rc = ensure(ctx, B[3], (C1 + 1) * 8)) ||
      (rc = ensure(ctx, B[4], Np * 4)) || (rc = ensure(ctx, B[5], Np * 4)) || (rc = ensure(ctx, B[6], C1 * 16)) || (rc = ensure(ctx, B[7], C1)) ||
      (rc = ensure(ctx, B[8], C1 * 4)) || (rc = ensure(ctx, B[9], C1 * 8)) || (rc = ensure(ctx, B[10], C1 * 4)) || (rc = ensure(ctx, B[11], C1 * 8)) ||
      (rc = ensure(ctx, B[12], C1 * 4)) || (rc = ensure(ctx, B[13], Sp * 4)) || (rc = ensure(ctx, B[14], Sp * 4)) || (rc = ensure(ctx, B[15], Sp * 4)) ||
      (rc = ensure(ctx, B[16], Sp)) || (rc = ensure(ctx, B[17], (U + 2) * 8)) || (rc = ensure(ctx, B[18], U * 4)) || (rc = ensure(ctx, B[19], Sp * 4)) ||
      (rc = ensure(ctx, B[20], Sp * 4)) || (rc = ensure(ctx, B[21], Sp * 4)) || (rc = ensure(ctx, B[22], Sp * 4)) || (rc = ensure(ctx, B[23], Sp)) ||
      (rc = ensure(ctx, B[24], Sp)) || (rc = ensure(ctx, B[25], U * 16)) || (rc = ensure(ctx, B[26], slots * 8 + 16)) || (rc = ensure(ctx, B[27], slots * 4 + 16)) ||
      (rc = ensure(ctx, B[28], slots * 4 + 16)) || (rc = ensure(ctx, B[29], C1 * 8)) || (rc = ensure(ctx, B[30], 16)))
    return rc;
  cudaStream_t st = ctx->stream;
  const void *src[13] = {in->ch, edge.data(), slot_off.data(), in->cl_off, in->q, in->t, in->cl_box, in->cl_strand, in->cl_freq, in->chrom_off, in->chrom_len,
                         in->read_off, in->read_len};
  const size_t sz[13] = {U * 4, U, (U + 1) * 8, (C1 + 1) * 8, N * 4, N * 4, C1 * 16, C1, C1 * 4, C1 * 8, C1 * 4, C1 * 8, C1 * 4};
  for (int i = 0; i < 13; i++) if (sz[i]) CU(cudaMemcpyAsync(B[i].p, src[i], sz[i], cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[29].p, slot.data(), C1 * 8, cudaMemcpyHostToDevice, st));
  LextChainBatch b;
  b.n_units = (int)U; b.K = in->K; b.skiprepetitive = in->skiprepetitive; b.trim = in->trim; b.merge_dist = in->merge_dist;
  b.reads = SeqView{reads->b2, reads->nm, reads->n}; b.genome = SeqView{genome->b2, genome->nm, genome->n};
  b.unit_cl = (const uint32_t *)B[0].p; b.unit_edge = (const uint8_t *)B[1].p; b.slot_off = (const unsigned long long *)B[2].p;
  b.cl_off = (const unsigned long long *)B[3].p; b.cq = (const uint32_t *)B[4].p; b.ct = (const uint32_t *)B[5].p; b.cl_box = (const uint32_t *)B[6].p;
  b.cl_strand = (const uint8_t *)B[7].p; b.cl_freq = (const float *)B[8].p; b.cl_chrom_off = (const unsigned long long *)B[9].p;
  b.cl_chrom_len = (const uint32_t *)B[10].p; b.cl_read_off = (const unsigned long long *)B[11].p; b.cl_read_len = (const uint32_t *)B[12].p;
  b.sq = (uint32_t *)B[13].p; b.st = (uint32_t *)B[14].p; b.sl = (int32_t *)B[15].p; b.so = (uint8_t *)B[16].p; b.cnt = (unsigned long long *)B[17].p;
  b.u_overlap = (int32_t *)B[18].p; b.lidx = (int *)B[19].p; b.eq = (uint32_t *)B[20].p; b.et = (uint32_t *)B[21].p; b.elen = (int32_t *)B[22].p;
  b.eovp = (uint8_t *)B[23].p; b.md_head = (uint8_t *)B[24].p; b.box = (uint32_t *)B[25].p;
  auto rec = [&](int i) { cudaEventRecord(ctx->ev[i], st); };
  rec(0);
  if (N) {   // DiagonalSort<GenomeTuple>(matches, 500) on strand 0, AntiDiagonalSort<GenomeTuple>(matches, 500) on strand 1 (LinearExtend.h:196-205)
    SortBatch sb;
    sb.n_seg = CL; sb.mode = 0; sb.seg_off = b.cl_off; sb.q = (uint32_t *)B[4].p; sb.t = (uint32_t *)B[5].p; sb.perm = nullptr;
    sb.kp = (unsigned long long *)B[26].p; sb.ks = (uint32_t *)B[27].p; sb.ki = (uint32_t *)B[28].p; sb.slot_off = (const unsigned long long *)B[29].p;
    sb.seg_mode = b.cl_strand;
    sort_pairs_kernel<<<(unsigned)CL, 256, 0, st>>>(sb);
    ctx->launches++;
  }
  rec(1);
  lextc_walk_kernel<<<(unsigned)((U + 127) / 128), 128, 0, st>>>(b);
  CU(cudaMemsetAsync(b.cnt + U, 0, 16, st));
  seed_scan_kernel<<<1, 1024, 0, st>>>(b.cnt, (int)U, ~0ull, (int *)B[30].p);
  rec(2);
  lextc_group_kernel<<<(unsigned)((U + 3) / 4), 128, 0, st>>>(b);
  ctx->launches += 3;
  rec(3);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->e_off, b.cnt, (U + 1) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->box, b.box, U * 16, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->overlap, b.u_overlap, U * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const size_t n_out = (size_t)res->e_off[U];
  res->n_total = n_out;
  if (n_out) {
    CU(cudaMemcpyAsync(res->q, b.eq, n_out * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->t, b.et, n_out * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->len, b.elen, n_out * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->ovp, b.eovp, n_out, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->md_head, b.md_head, n_out, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  const char *names[3] = {"lextc_sort", "lextc_walk+scan", "lextc_group"};
  for (int i = 0; i < 3; i++) {
    if (i == 0 && !N) continue;
    lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "%s", names[i]);
    cudaEventElapsedTime(&s2.ms, ctx->ev[i], ctx->ev[i + 1]); s2.jobs = i == 0 ? (uint64_t)CL : (uint64_t)U;
    s2.algo_bytes = i == 0 ? 16ull * N : i == 1 ? 8ull * S + 13ull * n_out : 27ull * n_out + 16ull * U;
    ctx->stats.push_back(s2);
  }
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a11 SPLITChain (low-accuracy pipeline)
extern "C" int lra_b200_split_chains_batch(lra_b200_ctx *ctx, const lra_b200_anchor_chains *in, lra_b200_split_chain_result *res) {
  if (!ctx || !in || !res) return fail(ctx, LRA_B200_EINVAL, "split_chains_batch: NULL argument");
  const int NC = in->n_chains;
  if (NC < 0 || !in->c_off || !in->hdr_pos || in->n_hdr < 1) return fail(ctx, LRA_B200_EINVAL, "split_chains_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (NC == 0) return LRA_B200_OK;
  const size_t N = (size_t)in->c_off[NC];
  if (N > 0x7FFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "split_chains_batch: more than 2^31 anchors in one batch");
  for (int k = 0; k < NC; k++) if (in->c_off[k + 1] < in->c_off[k]) return fail(ctx, LRA_B200_EINVAL, "split_chains_batch: chain offsets not ascending");
  if (N && (!in->q || !in->t || !in->len || !in->strand || !in->cnum || !in->link)) return fail(ctx, LRA_B200_EINVAL, "split_chains_batch: NULL anchor array");
  int rc;
  DevBuf *B = ctx->sp;
  const size_t Np = N ? N : 1, C1 = (size_t)NC;
  // 0 c_off, 1..6 inputs, 7 hdr, 8..15 int scratch, 16..19 uint scratch, 20..23 byte scratch, 24.. outputs
  const size_t need[37] = {(C1 + 1) * 8, Np * 4, Np * 4, Np * 4, Np, Np * 4, Np, (size_t)in->n_hdr * 8,
                           Np * 4, Np * 4, Np * 4, Np * 4, Np * 4, Np * 4, Np * 4, Np * 4, Np * 4, Np * 4, Np * 4, Np * 4, Np, Np, Np, Np,
                           C1 * 4, C1 * 4, (Np + C1 + 1) * 4, (Np + C1 + 1) * 4, Np * 4, Np * 4, Np, Np * 16, Np * 4, Np, Np, Np, 16};
  for (int i = 0; i < 37; i++) if ((rc = ensure(ctx, B[i], need[i]))) return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, in->c_off, (C1 + 1) * 8, cudaMemcpyHostToDevice, st));
  if (N) {
    const void *src[6] = {in->q, in->t, in->len, in->strand, in->cnum, in->link};
    const size_t sz[6] = {N * 4, N * 4, N * 4, N, N * 4, N};
    for (int i = 0; i < 6; i++) CU(cudaMemcpyAsync(B[1 + i].p, src[i], sz[i], cudaMemcpyHostToDevice, st));
  }
  CU(cudaMemcpyAsync(B[7].p, in->hdr_pos, (size_t)in->n_hdr * 8, cudaMemcpyHostToDevice, st));
  SpChainBatch b;
  b.n_chains = NC; b.splitdist = in->splitdist; b.bypass = in->bypass_clustering;
  b.c_off = (const unsigned long long *)B[0].p; b.q = (const uint32_t *)B[1].p; b.t = (const uint32_t *)B[2].p; b.len = (const int32_t *)B[3].p;
  b.strand = (const uint8_t *)B[4].p; b.cnum = (const int32_t *)B[5].p; b.link = (const uint8_t *)B[6].p; b.hdr_pos = (const unsigned long long *)B[7].p; b.n_hdr = in->n_hdr;
  b.pa = (int32_t *)B[8].p; b.pb = (int32_t *)B[9].p; b.pnext = (int32_t *)B[10].p; b.tail = (int32_t *)B[11].p; b.size = (int32_t *)B[12].p; b.chrom = (int32_t *)B[13].p;
  b.cur_ind = (int32_t *)B[14].p; b.ord = (int32_t *)B[15].p; b.QS = (uint32_t *)B[16].p; b.QE = (uint32_t *)B[17].p; b.TS = (uint32_t *)B[18].p; b.TE = (uint32_t *)B[19].p;
  b.type = (uint8_t *)B[20].p; b.pstrand = (uint8_t *)B[21].p; b.keep = (uint8_t *)B[22].p; b.SL = (uint8_t *)B[23].p;
  b.n_sp = (int32_t *)B[24].p; b.n_link = (int32_t *)B[25].p; b.sp_off = (int32_t *)B[26].p; b.ci_off = (int32_t *)B[27].p; b.sptc = (int32_t *)B[28].p; b.ci = (int32_t *)B[29].p;
  b.sp_lk = (uint8_t *)B[30].p; b.sp_box = (uint32_t *)B[31].p; b.sp_chrom = (int32_t *)B[32].p; b.sp_type = (uint8_t *)B[33].p; b.sp_strand = (uint8_t *)B[34].p;
  b.sp_link = (uint8_t *)B[35].p;
  cudaEventRecord(ctx->ev[0], st);
  spchain_kernel<<<(unsigned)((NC + 63) / 64), 64, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->n_sp, b.n_sp, C1 * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->n_link, b.n_link, C1 * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->sp_off, b.sp_off, (N + C1) * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->ci_off, b.ci_off, (N + C1) * 4, cudaMemcpyDeviceToHost, st));
  if (N) {
    CU(cudaMemcpyAsync(res->sptc, b.sptc, N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->ci, b.ci, N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->sp_lk, b.sp_lk, N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->sp_box, b.sp_box, N * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->sp_chrom, b.sp_chrom, N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->sp_type, b.sp_type, N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->sp_strand, b.sp_strand, N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->sp_link, b.sp_link, N, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "spchain");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)NC; s2.algo_bytes = 18ull * N + 9ull * N;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a11 MergeChain / switchindex
extern "C" int lra_b200_merge_chain_batch(lra_b200_ctx *ctx, const int32_t *sp, const uint64_t *sc_off, int32_t n_chains, const int32_t *chrom, const uint8_t *strand,
                                          const uint32_t *box, int32_t n_clusters, uint8_t *head) {
  if (!ctx || !sc_off || n_chains < 0 || n_clusters < 0) return fail(ctx, LRA_B200_EINVAL, "merge_chain_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  const size_t E = n_chains ? (size_t)sc_off[n_chains] : 0;
  if (E == 0) return LRA_B200_OK;
  if (!sp || !chrom || !strand || !box || !head || n_clusters == 0) return fail(ctx, LRA_B200_EINVAL, "merge_chain_batch: NULL array");
  std::vector<uint8_t> first(E, 0);
  for (int k = 0; k < n_chains; k++) {
    if (sc_off[k + 1] < sc_off[k]) return fail(ctx, LRA_B200_EINVAL, "merge_chain_batch: chain offsets not ascending");
    if (sc_off[k + 1] > sc_off[k]) first[(size_t)sc_off[k]] = 1;
  }
  for (size_t e = 0; e < E; e++) if (sp[e] < 0 || sp[e] >= n_clusters) return fail(ctx, LRA_B200_EINVAL, "merge_chain_batch: entry %zu names cluster %d of %d", e, sp[e], n_clusters);
  int rc;
  DevBuf *B = ctx->cg;
  const size_t C1 = (size_t)n_clusters;
  if ((rc = ensure(ctx, B[0], E * 4)) || (rc = ensure(ctx, B[1], E)) || (rc = ensure(ctx, B[2], C1 * 4)) || (rc = ensure(ctx, B[3], C1)) || (rc = ensure(ctx, B[4], C1 * 16)) ||
      (rc = ensure(ctx, B[5], E)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, sp, E * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[1].p, first.data(), E, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[2].p, chrom, C1 * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[3].p, strand, C1, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[4].p, box, C1 * 16, cudaMemcpyHostToDevice, st));
  MergeChainBatch b{E, (const int32_t *)B[0].p, (const uint8_t *)B[1].p, (const int32_t *)B[2].p, (const uint8_t *)B[3].p, (const uint32_t *)B[4].p, (uint8_t *)B[5].p};
  cudaEventRecord(ctx->ev[0], st);
  merge_chain_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(head, b.head, E, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "merge_chain");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)E; s2.algo_bytes = 6ull * E + 21ull * C1;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

extern "C" int lra_b200_switchindex_batch(lra_b200_ctx *ctx, int32_t *ch, uint8_t *link, const uint64_t *c_off, int32_t n_chains, const int32_t *coarse,
                                          int32_t n_splitclusters, const uint32_t *cq, int32_t n_clusters, int32_t *n_out, int32_t *nl_out) {
  if (!ctx || !c_off || n_chains < 0 || !n_out || !nl_out) return fail(ctx, LRA_B200_EINVAL, "switchindex_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (n_chains == 0) return LRA_B200_OK;
  const size_t E = (size_t)c_off[n_chains];
  for (int k = 0; k < n_chains; k++) if (c_off[k + 1] < c_off[k]) return fail(ctx, LRA_B200_EINVAL, "switchindex_batch: chain offsets not ascending");
  if (E && (!ch || !link || !coarse || !cq)) return fail(ctx, LRA_B200_EINVAL, "switchindex_batch: NULL array");
  for (size_t e = 0; e < E; e++) if (ch[e] < 0 || ch[e] >= n_splitclusters) return fail(ctx, LRA_B200_EINVAL, "switchindex_batch: entry %zu names split cluster %d of %d", e, ch[e], n_splitclusters);
  for (int i = 0; i < n_splitclusters; i++) if (coarse[i] < 0 || coarse[i] >= n_clusters) return fail(ctx, LRA_B200_EINVAL, "switchindex_batch: split cluster %d maps to cluster %d of %d", i, coarse[i], n_clusters);
  int rc;
  DevBuf *B = ctx->cg;
  const size_t Ep = E ? E : 1, C1 = (size_t)n_chains;
  if ((rc = ensure(ctx, B[6], (C1 + 1) * 8)) || (rc = ensure(ctx, B[7], Ep * 4)) || (rc = ensure(ctx, B[8], Ep)) || (rc = ensure(ctx, B[9], (size_t)(n_splitclusters ? n_splitclusters : 1) * 4)) ||
      (rc = ensure(ctx, B[10], (size_t)(n_clusters ? n_clusters : 1) * 8)) || (rc = ensure(ctx, B[11], Ep * 12)) || (rc = ensure(ctx, B[12], Ep * 2)) || (rc = ensure(ctx, B[13], C1 * 8)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[6].p, c_off, (C1 + 1) * 8, cudaMemcpyHostToDevice, st));
  if (E) {
    CU(cudaMemcpyAsync(B[7].p, ch, E * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[8].p, link, E, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[9].p, coarse, (size_t)n_splitclusters * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B[10].p, cq, (size_t)n_clusters * 8, cudaMemcpyHostToDevice, st));
  }
  SwitchIndexBatch b;
  b.n_chains = n_chains; b.c_off = (const unsigned long long *)B[6].p; b.ch = (int32_t *)B[7].p; b.link = (uint8_t *)B[8].p; b.coarse = (const int32_t *)B[9].p;
  b.cq = (const uint32_t *)B[10].p; b.ss = (int32_t *)B[11].p; b.se = b.ss + Ep; b.newch = b.se + Ep; b.newlink = (uint8_t *)B[12].p; b.flag = b.newlink + Ep;
  b.n_out = (int32_t *)B[13].p; b.nl_out = b.n_out + C1;
  cudaEventRecord(ctx->ev[0], st);
  switchindex_kernel<<<(unsigned)((n_chains + 63) / 64), 64, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  if (E) {
    CU(cudaMemcpyAsync(ch, b.ch, E * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(link, b.link, E, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaMemcpyAsync(n_out, b.n_out, C1 * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(nl_out, b.nl_out, C1 * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "switchindex");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)n_chains; s2.algo_bytes = 10ull * E;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a17 (leaf) RefineByLinearAlignment
extern "C" int lra_b200_refine_linear_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_linear_gaps *in,
                                            lra_b200_aog_result *res) {
  if (!ctx || !reads || !genome || !in || !res) return fail(ctx, LRA_B200_EINVAL, "refine_linear_batch: NULL argument");
  const int n = in->n_gaps;
  if (n < 0) return fail(ctx, LRA_B200_EINVAL, "refine_linear_batch: negative gap count");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_blocks_total = 0; res->cells = 0;
  if (n == 0) return LRA_B200_OK;
  if (!in->cur_read_end || !in->next_read_start || !in->cur_genome_end || !in->next_genome_start || !in->read_off || !in->chrom_off)
    return fail(ctx, LRA_B200_EINVAL, "refine_linear_batch: NULL gap array");
  for (int g = 0; g < n; g++) {       // the windows a gap with m > 0 reads must lie inside the arenas
    const uint32_t a = in->next_read_start[g] - in->cur_read_end[g] + 1u, c = in->next_genome_start[g] - in->cur_genome_end[g] + 1u;
    if ((int)(a < c ? a : c) <= 0) continue;
    const int ql = (int)(in->next_read_start[g] - in->cur_read_end[g]), tl = (int)(in->next_genome_start[g] - in->cur_genome_end[g]);
    if (ql < 0 || tl < 0 || (uint64_t)in->read_off[g] + in->next_read_start[g] > reads->n || (uint64_t)in->chrom_off[g] + in->next_genome_start[g] > genome->n)
      return fail(ctx, LRA_B200_EINVAL, "refine_linear_batch: gap %d reaches beyond its arena", g);
  }
  int rc;
  DevBuf *B = ctx->rl;
  const size_t nb4 = (size_t)n * 4;
  for (int i = 0; i < 6; i++) if ((rc = ensure(ctx, B[i], nb4))) return rc;
  if ((rc = ensure(ctx, ctx->d_qoff, nb4)) || (rc = ensure(ctx, ctx->d_toff, nb4)) || (rc = ensure(ctx, ctx->d_qlen, nb4)) || (rc = ensure(ctx, ctx->d_tlen, nb4)) ||
      (rc = ensure(ctx, ctx->d_k, nb4)) || (rc = ensure(ctx, ctx->d_score, nb4)) || (rc = ensure(ctx, ctx->d_nb, nb4)) || (rc = ensure(ctx, ctx->d_boff, (size_t)n * 8)) ||
      (rc = ensure(ctx, ctx->d_blocks, (size_t)(res->block_cap ? res->block_cap : 1) * 12)))
    return rc;
  cudaStream_t st = ctx->stream;
  const void *src[6] = {in->cur_read_end, in->next_read_start, in->cur_genome_end, in->next_genome_start, in->read_off, in->chrom_off};
  for (int i = 0; i < 6; i++) CU(cudaMemcpyAsync(B[i].p, src[i], nb4, cudaMemcpyHostToDevice, st));
  RlaBatch b;
  b.n_gaps = n; b.local_band = in->local_band;
  b.cur_read_end = (const uint32_t *)B[0].p; b.next_read_start = (const uint32_t *)B[1].p; b.cur_genome_end = (const uint32_t *)B[2].p;
  b.next_genome_start = (const uint32_t *)B[3].p; b.read_off = (const uint32_t *)B[4].p; b.chrom_off = (const uint32_t *)B[5].p;
  b.q_off = (uint32_t *)ctx->d_qoff.p; b.t_off = (uint32_t *)ctx->d_toff.p; b.q_len = (int32_t *)ctx->d_qlen.p; b.t_len = (int32_t *)ctx->d_tlen.p; b.k = (int32_t *)ctx->d_k.p;
  b.n_blocks = (const int32_t *)ctx->d_nb.p; b.block_off = (const unsigned long long *)ctx->d_boff.p; b.blocks = (uint32_t *)ctx->d_blocks.p;
  rla_jobs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b);
  ctx->launches++;
  CU(cudaGetLastError());
  lra_b200_aog_jobs dj;
  dj.q_off = b.q_off; dj.t_off = b.t_off; dj.q_len = b.q_len; dj.t_len = b.t_len; dj.k = b.k; dj.n_jobs = n;
  dj.match = in->match; dj.mismatch = in->mismatch; dj.indel = in->indel;
  lra_b200_aog_result dr = *res;
  dr.score = (int32_t *)ctx->d_score.p; dr.n_blocks = (int32_t *)ctx->d_nb.p; dr.block_off = (uint64_t *)ctx->d_boff.p; dr.blocks = (uint32_t *)ctx->d_blocks.p;
  rc = aog_run_device(ctx, reads, genome, &dj, &dr);
  res->n_blocks_total = dr.n_blocks_total; res->cells = dr.cells;
  if (rc != LRA_B200_OK && rc != LRA_B200_EOVERFLOW) return rc;
  if (rc == LRA_B200_OK) {
    rla_shift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b);
    ctx->launches++;
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(res->score, dr.score, nb4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->n_blocks, dr.n_blocks, nb4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->block_off, dr.block_off, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (rc == LRA_B200_OK && dr.n_blocks_total) CU(cudaMemcpyAsync(res->blocks, dr.blocks, (size_t)dr.n_blocks_total * 12, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return rc;
}

// ---------------------------------------------------------------------------------------------------- a14 (core) RefineSpace
extern "C" int lra_b200_refine_space_batch(lra_b200_ctx *ctx, const lra_b200_seq *reads, const lra_b200_seq *genome, const lra_b200_spaces *in,
                                           lra_b200_space_result *res) {
  if (!ctx || !reads || !genome || !in || !res) return fail(ctx, LRA_B200_EINVAL, "refine_space_batch: NULL argument");
  const int n = in->n_spaces;
  if (n < 0 || in->K <= 0) return fail(ctx, LRA_B200_EINVAL, "refine_space_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  res->n_pairs_total = 0;
  if (n == 0) return LRA_B200_OK;
  if (!in->qs || !in->qe || !in->ts || !in->te || !in->lrts || !in->lrlength || !in->read_off || !in->read_len || !in->chrom_off || !in->flip)
    return fail(ctx, LRA_B200_EINVAL, "refine_space_batch: NULL space array");
  std::vector<unsigned long long> pair_off((size_t)n + 1), mq_off, mt_off;
  std::vector<uint8_t> large((size_t)n, 0);
  std::vector<uint32_t> lidx;
  size_t blk = 0, MQ = 0, MT = 0;
  for (int g = 0; g < n; g++) {
    const long long ql = (long long)in->qe[g] - (long long)in->qs[g], tl = (long long)in->te[g] - (long long)in->ts[g] + (long long)in->lrlength[g];
    if (ql < 0 || tl < 0 || in->lrts[g] > in->ts[g]) return fail(ctx, LRA_B200_EINVAL, "refine_space_batch: space %d has a negative extent", g);
    if ((uint64_t)in->read_off[g] + in->qe[g] > reads->n || (uint64_t)in->chrom_off[g] + in->te[g] + in->lrlength[g] > genome->n)
      return fail(ctx, LRA_B200_EINVAL, "refine_space_batch: space %d reaches beyond its arena", g);
    if (in->qe[g] > in->read_len[g]) return fail(ctx, LRA_B200_EINVAL, "refine_space_batch: space %d ends at read position %u of a read of %u bases", g, in->qe[g], in->read_len[g]);
    if (ql >= 1000 || tl >= 1000) {          // the minimizer branch (ClusterRefine.h:296-305)
      if (!in->diag || in->W <= 0 || in->W > kSeedMaxW || in->K > 31) return fail(ctx, LRA_B200_EINVAL, "refine_space_batch: space %d needs W, localMaxFreq and refineSpaceDiag", g);
      large[g] = 1; lidx.push_back((uint32_t)g); mq_off.push_back(MQ); mt_off.push_back(MT);
      MQ += (size_t)ql + 1; MT += (size_t)tl + 1;
    } else {
      const long long mn = ql < tl ? ql : tl;
      blk += (size_t)mn + 1;
    }
  }
  const int nl = (int)lidx.size();
  int rc;
  DevBuf *B = ctx->rs;
  const size_t nb4 = (size_t)n * 4;
  for (int i = 0; i < 9; i++) if ((rc = ensure(ctx, B[i], nb4))) return rc;
  if ((rc = ensure(ctx, B[9], (size_t)n)) || (rc = ensure(ctx, B[10], ((size_t)n + 1) * 8)) || (rc = ensure(ctx, B[13], nb4)) || (rc = ensure(ctx, B[14], nb4)) ||
      (rc = ensure(ctx, B[15], (size_t)n)))
    return rc;
  if ((rc = ensure(ctx, ctx->d_qoff, nb4)) || (rc = ensure(ctx, ctx->d_toff, nb4)) || (rc = ensure(ctx, ctx->d_qlen, nb4)) || (rc = ensure(ctx, ctx->d_tlen, nb4)) ||
      (rc = ensure(ctx, ctx->d_k, nb4)) || (rc = ensure(ctx, ctx->d_score, nb4)) || (rc = ensure(ctx, ctx->d_nb, nb4)) || (rc = ensure(ctx, ctx->d_boff, (size_t)n * 8)) ||
      (rc = ensure(ctx, ctx->d_blocks, (blk ? blk : 1) * 12)))
    return rc;
  cudaStream_t st = ctx->stream;
  const void *src[9] = {in->qs, in->qe, in->ts, in->te, in->lrts, in->lrlength, in->read_off, in->read_len, in->chrom_off};
  for (int i = 0; i < 9; i++) CU(cudaMemcpyAsync(B[i].p, src[i], nb4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[9].p, in->flip, (size_t)n, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[15].p, large.data(), (size_t)n, cudaMemcpyHostToDevice, st));
  RspBatch b;
  b.n = n; b.K = in->K;
  b.reads = SeqView{reads->b2, reads->nm, reads->n}; b.genome = SeqView{genome->b2, genome->nm, genome->n};
  b.qs = (const uint32_t *)B[0].p; b.qe = (const uint32_t *)B[1].p; b.ts = (const uint32_t *)B[2].p; b.te = (const uint32_t *)B[3].p; b.lrts = (const uint32_t *)B[4].p;
  b.lrlength = (const uint32_t *)B[5].p; b.read_off = (const uint32_t *)B[6].p; b.read_len = (const uint32_t *)B[7].p; b.chrom_off = (const uint32_t *)B[8].p;
  b.flip = (const uint8_t *)B[9].p; b.large = (const uint8_t *)B[15].p; b.pair_off = (const unsigned long long *)B[10].p;
  b.n_pairs = (int32_t *)B[13].p; b.identity = (float *)B[14].p;
  b.q_off = (uint32_t *)ctx->d_qoff.p; b.t_off = (uint32_t *)ctx->d_toff.p; b.q_len = (int32_t *)ctx->d_qlen.p; b.t_len = (int32_t *)ctx->d_tlen.p; b.k = (int32_t *)ctx->d_k.p;
  b.n_blocks = (const int32_t *)ctx->d_nb.p; b.block_off = (const unsigned long long *)ctx->d_boff.p; b.blocks = (const uint32_t *)ctx->d_blocks.p;
  // ---- large spaces: minimizers of both windows, sort, count the pairs inside the band
  RsplBatch lb; memset(&lb, 0, sizeof lb);
  std::vector<unsigned long long> cnt((size_t)nl + 1, 0ull);
  if (nl) {
    DevBuf *L = ctx->rs2;
    if ((rc = ensure(ctx, L[0], (size_t)nl * 4)) || (rc = ensure(ctx, L[1], (size_t)nl * 8)) || (rc = ensure(ctx, L[2], (size_t)nl * 8)) || (rc = ensure(ctx, L[3], MQ * 8)) ||
        (rc = ensure(ctx, L[4], MT * 8)) || (rc = ensure(ctx, L[5], MQ * 4)) || (rc = ensure(ctx, L[6], MT * 4)) || (rc = ensure(ctx, L[7], (size_t)nl * 4)) ||
        (rc = ensure(ctx, L[8], (size_t)nl * 4)) || (rc = ensure(ctx, L[9], (size_t)nl * 8)) || (rc = ensure(ctx, L[10], nb4)))
      return rc;
    CU(cudaMemcpyAsync(L[0].p, lidx.data(), (size_t)nl * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(L[1].p, mq_off.data(), (size_t)nl * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(L[2].p, mt_off.data(), (size_t)nl * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(L[10].p, in->diag, nb4, cudaMemcpyHostToDevice, st));
    lb.n_large = nl; lb.K = in->K; lb.W = in->W; lb.max_freq = in->local_max_freq; lb.reads = b.reads; lb.genome = b.genome;
    lb.idx = (const uint32_t *)L[0].p; lb.qs = b.qs; lb.qe = b.qe; lb.ts = b.ts; lb.te = b.te; lb.lrts = b.lrts; lb.lrlength = b.lrlength; lb.read_off = b.read_off;
    lb.read_len = b.read_len; lb.chrom_off = b.chrom_off; lb.flip = b.flip; lb.diag = (const int32_t *)L[10].p;
    lb.mq_off = (const unsigned long long *)L[1].p; lb.mt_off = (const unsigned long long *)L[2].p; lb.mq_t = (unsigned long long *)L[3].p; lb.mt_t = (unsigned long long *)L[4].p;
    lb.mq_p = (uint32_t *)L[5].p; lb.mt_p = (uint32_t *)L[6].p; lb.mq_n = (uint32_t *)L[7].p; lb.mt_n = (uint32_t *)L[8].p; lb.cnt = (unsigned long long *)L[9].p;
    lb.pair_off = b.pair_off; lb.n_pairs = b.n_pairs; lb.identity = b.identity;
    rspl_mins_kernel<<<(unsigned)((2 * nl + 63) / 64), 64, 0, st>>>(lb);
    rspl_compare_kernel<false><<<(unsigned)((nl + 63) / 64), 64, 0, st>>>(lb);
    ctx->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(cnt.data(), lb.cnt, (size_t)nl * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  // ---- slots: min(qLen, tLen) / K + 1 for an aligned space, the counted pairs for a minimizer space
  size_t P = 0;
  { int j = 0;
    for (int g = 0; g < n; g++) {
      pair_off[g] = P;
      if (large[g]) P += (size_t)cnt[j++];
      else {
        const long long ql = (long long)in->qe[g] - (long long)in->qs[g], tl = (long long)in->te[g] - (long long)in->ts[g] + (long long)in->lrlength[g];
        P += (size_t)((ql < tl ? ql : tl) / in->K + 1);
      }
    }
    pair_off[n] = P; }
  res->n_pairs_total = P;
  if (res->pair_cap < P) return fail(ctx, LRA_B200_EOVERFLOW, "refine_space_batch: pair arrays hold %llu entries, %llu slots are needed",
                                     (unsigned long long)res->pair_cap, (unsigned long long)P);
  if ((rc = ensure(ctx, B[11], (P ? P : 1) * 4)) || (rc = ensure(ctx, B[12], (P ? P : 1) * 4))) return rc;
  b.pq = (uint32_t *)B[11].p; b.pt = (uint32_t *)B[12].p; lb.pq = b.pq; lb.pt = b.pt;
  CU(cudaMemcpyAsync(B[10].p, pair_off.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
  rsp_jobs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b);
  ctx->launches++;
  CU(cudaGetLastError());
  lra_b200_aog_jobs dj;
  dj.q_off = b.q_off; dj.t_off = b.t_off; dj.q_len = b.q_len; dj.t_len = b.t_len; dj.k = b.k; dj.n_jobs = n;
  dj.match = in->match; dj.mismatch = in->mismatch; dj.indel = in->indel;
  lra_b200_aog_result dr; memset(&dr, 0, sizeof dr);
  dr.score = (int32_t *)ctx->d_score.p; dr.n_blocks = (int32_t *)ctx->d_nb.p; dr.block_off = (uint64_t *)ctx->d_boff.p; dr.blocks = (uint32_t *)ctx->d_blocks.p;
  dr.block_cap = blk ? blk : 1;
  if ((rc = aog_run_device(ctx, reads, genome, &dj, &dr))) return rc;
  rsp_harvest_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b);
  ctx->launches++;
  if (nl) { rspl_compare_kernel<true><<<(unsigned)((nl + 63) / 64), 64, 0, st>>>(lb); ctx->launches++; }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->pair_off, b.pair_off, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->n_pairs, b.n_pairs, nb4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->identity, b.identity, nb4, cudaMemcpyDeviceToHost, st));
  if (P) {
    CU(cudaMemcpyAsync(res->pq, b.pq, P * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->pt, b.pt, P * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a17 SwitchToOriginalAnchors
extern "C" int lra_b200_switch_to_original_batch(lra_b200_ctx *ctx, const int32_t *run_start, const int32_t *run_end, const int32_t *coarse, uint64_t n_entries,
                                                 uint64_t *off, uint32_t *chain, int32_t *cluster_index, uint64_t cap, uint64_t *n_total) {
  if (!ctx || !n_total) return fail(ctx, LRA_B200_EINVAL, "switch_to_original_batch: NULL argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  *n_total = 0;
  if (n_entries == 0) { if (off) off[0] = 0; return LRA_B200_OK; }
  if (!run_start || !run_end || !coarse || !off) return fail(ctx, LRA_B200_EINVAL, "switch_to_original_batch: NULL array");
  if (n_entries > 0x7FFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "switch_to_original_batch: more than 2^31 entries");
  int rc;
  DevBuf *B = ctx->cg;
  const size_t E = (size_t)n_entries;
  if ((rc = ensure(ctx, B[0], E * 4)) || (rc = ensure(ctx, B[1], E * 4)) || (rc = ensure(ctx, B[2], E * 4)) || (rc = ensure(ctx, B[3], (E + 2) * 8)) || (rc = ensure(ctx, B[14], 16)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, run_start, E * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[1].p, run_end, E * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[2].p, coarse, E * 4, cudaMemcpyHostToDevice, st));
  SwitchOrigBatch b{n_entries, (const int32_t *)B[0].p, (const int32_t *)B[1].p, (const int32_t *)B[2].p, (unsigned long long *)B[3].p, nullptr, nullptr};
  cudaEventRecord(ctx->ev[0], st);
  switch_orig_count_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(b);
  seed_scan_kernel<<<1, 1024, 0, st>>>(b.off, (int)E, ~0ull, (int *)B[14].p);
  CU(cudaMemcpyAsync(off, b.off, (E + 1) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const size_t T = (size_t)off[E];
  *n_total = T;
  if (T > cap) return fail(ctx, LRA_B200_EOVERFLOW, "switch_to_original_batch: result arrays hold %llu anchors, %llu needed", (unsigned long long)cap, (unsigned long long)T);
  if (T) {
    if (!chain || !cluster_index) return fail(ctx, LRA_B200_EINVAL, "switch_to_original_batch: NULL result array");
    if ((rc = ensure(ctx, B[4], T * 4)) || (rc = ensure(ctx, B[5], T * 4))) return rc;
    b.chain = (uint32_t *)B[4].p; b.cluster_index = (int32_t *)B[5].p;
    switch_orig_emit_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(b);
    cudaEventRecord(ctx->ev[1], st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(chain, b.chain, T * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cluster_index, b.cluster_index, T * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  } else cudaEventRecord(ctx->ev[1], st);
  ctx->launches += 3;
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- a8 (first half) SplitRoughClustersWithGaps
extern "C" int lra_b200_split_rough_batch(lra_b200_ctx *ctx, const lra_b200_rough_lists *in, lra_b200_split_rough_result *res) {
  if (!ctx || !in || !res) return fail(ctx, LRA_B200_EINVAL, "split_rough_batch: NULL argument");
  const int NL = in->n_lists;
  if (NL < 0 || !in->l_off || !in->lr_off) return fail(ctx, LRA_B200_EINVAL, "split_rough_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (NL == 0) return LRA_B200_OK;
  const size_t N = (size_t)in->l_off[NL], RC = (size_t)in->lr_off[NL];
  if (N + RC > 0x7FFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "split_rough_batch: batch too large");
  if (RC && (!in->r_start || !in->r_end || !in->r_box || !in->r_strand || !in->r_freq || !in->r_chrom)) return fail(ctx, LRA_B200_EINVAL, "split_rough_batch: NULL cluster array");
  if (N && (!in->q || !in->t)) return fail(ctx, LRA_B200_EINVAL, "split_rough_batch: NULL anchors");
  for (int l = 0; l < NL; l++) {
    if (in->l_off[l + 1] < in->l_off[l] || in->lr_off[l + 1] < in->lr_off[l]) return fail(ctx, LRA_B200_EINVAL, "split_rough_batch: offsets not ascending");
    const long long n = (long long)(in->l_off[l + 1] - in->l_off[l]);
    for (size_t c = (size_t)in->lr_off[l]; c < (size_t)in->lr_off[l + 1]; c++)
      if (in->r_start[c] < 0 || in->r_end[c] < in->r_start[c] || in->r_end[c] > n) return fail(ctx, LRA_B200_EINVAL, "split_rough_batch: rough cluster %zu lies outside its list", c);
  }
  int rc;
  DevBuf *B = ctx->sr;
  const size_t Np = N ? N : 1, Rp = RC ? RC : 1, S = N + RC + 1, L1 = (size_t)NL;
  const size_t need[22] = {(L1 + 1) * 8, (L1 + 1) * 8, Np * 4, Np * 4, Rp * 4, Rp * 4, Rp * 16, Rp, Rp * 4, Rp * 4,
                           L1 * 4, L1 * 4, S * 4, S * 4, S * 4, S * 4, S * 16, S, S * 4, S * 4, S * 4, S * 4};
  for (int i = 0; i < 22; i++) if ((rc = ensure(ctx, B[i], need[i]))) return rc;
  cudaStream_t st = ctx->stream;
  const void *src[10] = {in->l_off, in->lr_off, in->q, in->t, in->r_start, in->r_end, in->r_box, in->r_strand, in->r_freq, in->r_chrom};
  const size_t sz[10] = {(L1 + 1) * 8, (L1 + 1) * 8, N * 4, N * 4, RC * 4, RC * 4, RC * 16, RC, RC * 4, RC * 4};
  for (int i = 0; i < 10; i++) if (sz[i]) CU(cudaMemcpyAsync(B[i].p, src[i], sz[i], cudaMemcpyHostToDevice, st));
  SplitRoughBatch b;
  b.n_lists = NL; b.globalK = in->globalK; b.maxGap = in->rough_cluster_max_gap; b.minClusterSize = in->min_cluster_size; b.maxDiag = in->max_diag;
  b.l_off = (const unsigned long long *)B[0].p; b.lr_off = (const unsigned long long *)B[1].p; b.q = (const uint32_t *)B[2].p; b.t = (const uint32_t *)B[3].p;
  b.r_start = (const int32_t *)B[4].p; b.r_end = (const int32_t *)B[5].p; b.r_box = (const uint32_t *)B[6].p; b.r_strand = (const uint8_t *)B[7].p;
  b.r_freq = (const float *)B[8].p; b.r_chrom = (const int32_t *)B[9].p;
  b.n_split = (int32_t *)B[10].p; b.n_piece = (int32_t *)B[11].p; b.s_start = (int32_t *)B[12].p; b.s_end = (int32_t *)B[13].p; b.s_coarse = (int32_t *)B[14].p;
  b.s_chrom = (int32_t *)B[15].p; b.s_box = (uint32_t *)B[16].p; b.s_strand = (uint8_t *)B[17].p; b.s_freq = (float *)B[18].p; b.p_cluster = (int32_t *)B[19].p;
  b.p_start = (int32_t *)B[20].p; b.p_end = (int32_t *)B[21].p;
  cudaEventRecord(ctx->ev[0], st);
  split_rough_kernel<<<(unsigned)((NL + 63) / 64), 64, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  const size_t T = N + RC;
  CU(cudaMemcpyAsync(res->n_split, b.n_split, L1 * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res->n_piece, b.n_piece, L1 * 4, cudaMemcpyDeviceToHost, st));
  if (T) {
    void *dst[10] = {res->s_start, res->s_end, res->s_coarse, res->s_chrom, res->s_box, res->s_strand, res->s_freq, res->p_cluster, res->p_start, res->p_end};
    const void *dsrc[10] = {b.s_start, b.s_end, b.s_coarse, b.s_chrom, b.s_box, b.s_strand, b.s_freq, b.p_cluster, b.p_start, b.p_end};
    const size_t dsz[10] = {T * 4, T * 4, T * 4, T * 4, T * 16, T, T * 4, T * 4, T * 4, T * 4};
    for (int i = 0; i < 10; i++) CU(cudaMemcpyAsync(dst[i], dsrc[i], dsz[i], cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "split_rough");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)NL; s2.algo_bytes = 8ull * N + 33ull * RC;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- StoreDiagonalClusters
extern "C" int lra_b200_store_diagonal_batch(lra_b200_ctx *ctx, const lra_b200_cleaned_lists *in, lra_b200_diag_clusters *res) {
  if (!ctx || !in || !res) return fail(ctx, LRA_B200_EINVAL, "store_diagonal_batch: NULL argument");
  const int NL = in->n_lists;
  if (NL < 0 || !in->l_off || !in->hdr_pos || in->n_hdr < 1) return fail(ctx, LRA_B200_EINVAL, "store_diagonal_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (NL == 0) return LRA_B200_OK;
  const size_t N = (size_t)in->l_off[NL];
  if (N > 0x7FFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "store_diagonal_batch: more than 2^31 anchors in one batch");
  for (int l = 0; l < NL; l++) if (in->l_off[l + 1] < in->l_off[l]) return fail(ctx, LRA_B200_EINVAL, "store_diagonal_batch: list offsets not ascending");
  if (!in->strand || (N && (!in->q || !in->t || !in->qt || !in->freq))) return fail(ctx, LRA_B200_EINVAL, "store_diagonal_batch: NULL array");
  int rc;
  DevBuf *B = ctx->sr;
  const size_t Np = N ? N : 1, L1 = (size_t)NL;
  const size_t need[13] = {(L1 + 1) * 8, Np * 4, Np * 4, Np * 8, Np * 4, L1, (size_t)in->n_hdr * 8, L1 * 4, Np * 4, Np * 4, Np * 4, Np * 16, Np * 4};
  for (int i = 0; i < 13; i++) if ((rc = ensure(ctx, B[i], need[i]))) return rc;
  cudaStream_t st = ctx->stream;
  const void *src[7] = {in->l_off, in->q, in->t, in->qt, in->freq, in->strand, in->hdr_pos};
  const size_t sz[7] = {(L1 + 1) * 8, N * 4, N * 4, N * 8, N * 4, L1, (size_t)in->n_hdr * 8};
  for (int i = 0; i < 7; i++) if (sz[i]) CU(cudaMemcpyAsync(B[i].p, src[i], sz[i], cudaMemcpyHostToDevice, st));
  StoreDiagBatch b;
  b.n_lists = NL; b.globalK = in->globalK; b.maxDiag = in->max_diag; b.minClusterSize = in->min_cluster_size; b.minClusterLength = in->min_cluster_length;
  b.bypass = in->bypass_clustering;
  b.l_off = (const unsigned long long *)B[0].p; b.q = (const uint32_t *)B[1].p; b.t = (const uint32_t *)B[2].p; b.qt = (const unsigned long long *)B[3].p;
  b.freq = (const float *)B[4].p; b.strand = (const uint8_t *)B[5].p; b.hdr_pos = (const unsigned long long *)B[6].p; b.n_hdr = in->n_hdr;
  b.n_cl = (int32_t *)B[7].p; b.c_start = (int32_t *)B[8].p; b.c_end = (int32_t *)B[9].p; b.c_chrom = (int32_t *)B[10].p; b.c_box = (uint32_t *)B[11].p; b.c_freq = (float *)B[12].p;
  cudaEventRecord(ctx->ev[0], st);
  store_diagonal_kernel<<<(unsigned)((NL + 63) / 64), 64, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res->n_cl, b.n_cl, L1 * 4, cudaMemcpyDeviceToHost, st));
  if (N) {
    CU(cudaMemcpyAsync(res->c_start, b.c_start, N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->c_end, b.c_end, N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->c_chrom, b.c_chrom, N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->c_box, b.c_box, N * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res->c_freq, b.c_freq, N * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "store_diagonal");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)NL; s2.algo_bytes = 20ull * N;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}

// ---------------------------------------------------------------------------------------------------- TrimSplitChainDiagonal
extern "C" int lra_b200_trim_splitchains_batch(lra_b200_ctx *ctx, const uint32_t *cq, const uint32_t *ct, const uint64_t *c_off, const uint8_t *strand, int32_t n_chains,
                                               uint32_t *q, uint32_t *t, const uint64_t *m_off, uint8_t *keep, int32_t *removed) {
  if (!ctx || n_chains < 0) return fail(ctx, LRA_B200_EINVAL, "trim_splitchains_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  ctx->stats.clear();
  if (n_chains == 0) return LRA_B200_OK;          // an empty batch needs no arrays
  if (!c_off || !m_off || !strand || !removed) return fail(ctx, LRA_B200_EINVAL, "trim_splitchains_batch: bad argument");
  const size_t A = (size_t)c_off[n_chains], M = (size_t)m_off[n_chains], C1 = (size_t)n_chains;
  if (M > 0x7FFFFFF0ull) return fail(ctx, LRA_B200_EINVAL, "trim_splitchains_batch: more than 2^31 anchors in one batch");
  if ((A && (!cq || !ct)) || (M && (!q || !t || !keep))) return fail(ctx, LRA_B200_EINVAL, "trim_splitchains_batch: NULL array");
  std::vector<uint8_t> mode(C1);
  std::vector<unsigned long long> slot(C1);
  size_t slots = 0;
  for (int c = 0; c < n_chains; c++) {
    if (c_off[c + 1] < c_off[c] || m_off[c + 1] < m_off[c]) return fail(ctx, LRA_B200_EINVAL, "trim_splitchains_batch: offsets not ascending");
    const size_t nch = (size_t)(c_off[c + 1] - c_off[c]), n = (size_t)(m_off[c + 1] - m_off[c]);
    if (strand[c] != 0 && nch == 0) return fail(ctx, LRA_B200_EINVAL, "trim_splitchains_batch: chain %d is empty", c);
    mode[c] = nch == 1 ? 255 : 2;                  // chains of one anchor are left alone, the others get CartesianSort
    size_t P2 = 1; while (P2 < n) P2 <<= 1;
    slot[c] = slots;
    if (mode[c] == 2 && P2 > (size_t)kSortSmem) slots += P2;
  }
  int rc;
  DevBuf *B = ctx->cg;
  const size_t Ap = A ? A : 1, Mp = M ? M : 1;
  if ((rc = ensure(ctx, B[0], (C1 + 1) * 8)) || (rc = ensure(ctx, B[1], Ap * 4)) || (rc = ensure(ctx, B[2], Ap * 4)) || (rc = ensure(ctx, B[3], C1)) || (rc = ensure(ctx, B[4], (C1 + 1) * 8)) ||
      (rc = ensure(ctx, B[5], Mp * 4)) || (rc = ensure(ctx, B[6], Mp * 4)) || (rc = ensure(ctx, B[7], Mp)) || (rc = ensure(ctx, B[8], C1 * 4)) || (rc = ensure(ctx, B[9], C1)) ||
      (rc = ensure(ctx, B[10], C1 * 8)) || (rc = ensure(ctx, B[11], slots * 8 + 16)) || (rc = ensure(ctx, B[12], slots * 4 + 16)) || (rc = ensure(ctx, B[13], slots * 4 + 16)))
    return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(B[0].p, c_off, (C1 + 1) * 8, cudaMemcpyHostToDevice, st));
  if (A) { CU(cudaMemcpyAsync(B[1].p, cq, A * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[2].p, ct, A * 4, cudaMemcpyHostToDevice, st)); }
  CU(cudaMemcpyAsync(B[3].p, strand, C1, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[4].p, m_off, (C1 + 1) * 8, cudaMemcpyHostToDevice, st));
  if (M) { CU(cudaMemcpyAsync(B[5].p, q, M * 4, cudaMemcpyHostToDevice, st)); CU(cudaMemcpyAsync(B[6].p, t, M * 4, cudaMemcpyHostToDevice, st)); }
  CU(cudaMemcpyAsync(B[9].p, mode.data(), C1, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(B[10].p, slot.data(), C1 * 8, cudaMemcpyHostToDevice, st));
  cudaEventRecord(ctx->ev[0], st);
  if (M) {
    SortBatch sb;
    sb.n_seg = n_chains; sb.mode = 2; sb.seg_off = (const unsigned long long *)B[4].p; sb.q = (uint32_t *)B[5].p; sb.t = (uint32_t *)B[6].p; sb.perm = nullptr;
    sb.kp = (unsigned long long *)B[11].p; sb.ks = (uint32_t *)B[12].p; sb.ki = (uint32_t *)B[13].p; sb.slot_off = (const unsigned long long *)B[10].p;
    sb.seg_mode = (const uint8_t *)B[9].p;
    sort_pairs_kernel<<<(unsigned)n_chains, 256, 0, st>>>(sb);
    ctx->launches++;
  }
  TrimChainBatch b{n_chains, (const unsigned long long *)B[0].p, (const uint32_t *)B[1].p, (const uint32_t *)B[2].p, (const uint8_t *)B[3].p, (const unsigned long long *)B[4].p,
                   (const uint32_t *)B[5].p, (const uint32_t *)B[6].p, (uint8_t *)B[7].p, (int32_t *)B[8].p};
  trim_splitchain_kernel<<<(unsigned)((n_chains + 63) / 64), 64, 0, st>>>(b);
  cudaEventRecord(ctx->ev[1], st);
  ctx->launches++;
  CU(cudaGetLastError());
  if (M) {
    CU(cudaMemcpyAsync(q, B[5].p, M * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(t, B[6].p, M * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(keep, B[7].p, M, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaMemcpyAsync(removed, B[8].p, C1 * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  lra_b200_kernel_stat s2; memset(&s2, 0, sizeof s2); snprintf(s2.name, sizeof s2.name, "trim_splitchain(sort+trim)");
  cudaEventElapsedTime(&s2.ms, ctx->ev[0], ctx->ev[1]); s2.jobs = (uint64_t)n_chains; s2.algo_bytes = 17ull * M + 8ull * A;
  ctx->stats.push_back(s2);
  return LRA_B200_OK;
}
