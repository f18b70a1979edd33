// a18  AffineOneGapAlign  -- batched device implementation (reference: AffineOneGapAlign.h:157-649).
//
// Job model: thousands..millions of independent (q window, t window, k) alignments per launch, described by SoA
// arrays that point into two packed sequence arenas (lra_common.cuh).  Scoring (m, mm, indel) is per batch.
//
// Kernels
//   aog_classify_kernel / aog_scan_kernel / aog_scatter_kernel
//        classify every job (mode, effective band, rows) and counting-sort job ids by (class, rows descending) so
//        that the 32 jobs of a warp are alike and long jobs start first.
//   aog_thread_kernel<K>   one job per THREAD.  One-sided mode ("diag+2k >= max(len)" hack, AffineOneGapAlign.h:196-203),
//        exact doubled half-width K in {2,4,..,14}: the whole band row lives in registers, the query window is a
//        packed shift register, the match mask of a row is computed bit-parallel, traceback arrows are 2 bits/cell in
//        one local-memory word per row.  This is the many-tiny-jobs regime.
//   aog_warp_literal_kernel   one job per WARP, any band, both modes.  It mirrors the reference's flat matrices
//        (row stride R=2k+3, rails, boundary loops, their overwrites and aliasing) in a per-warp HBM/L2 scratch slab, so
//        every quirk that is observable in the block list is reproduced by construction; a row is filled in parallel
//        with a max-plus warp scan (S[x] = max(T[x], S[x-1]+indel)).
//   aog_warp_band_kernel<C>  (aog_band_kernel.cuh) fast register-resident warp kernel for wide one-sided bands.
#pragma once
#include "lra_common.cuh"

namespace lra {

// ---------------------------------------------------------------------------------------------------- batch + plan
struct AogBatch {
  SeqView q, t;
  const uint32_t *q_off, *t_off;
  const int32_t *q_len, *t_len, *k;
  int n_jobs;
  int m, mm, indel;
  int32_t *score;                       // [n_jobs]
  int32_t *n_blocks;                    // [n_jobs]
  unsigned long long *block_off;        // [n_jobs] index of the job's first triple in `blocks`
  uint32_t *blocks;                     // [block_cap][3]  (qPos, tPos, length)
  unsigned long long block_cap;
  unsigned long long *block_cursor;     // device counter
  int *err;                             // device flag: bit0 block overflow, bit1 scratch index error, bit2 bad traceback
};

constexpr int kAogThreadClasses = 7;            // K = 2,4,...,14
constexpr int kAogClsLiteral = 7;               // warp-per-job literal kernel
constexpr int kAogClsBand1 = 8;                 // warp band kernel, C=1,2,4,8 -> classes 8..11
constexpr int kAogNumClasses = 12;
constexpr int kAogBuckets = 128;
constexpr int kAogBins = kAogNumClasses * kAogBuckets;
constexpr int kAogThreadMaxRows = 512;

struct AogPlan {
  uint32_t hist[kAogBins];
  uint32_t bin_start[kAogBins + 1];
  uint32_t cursor[kAogBins];
  uint32_t work[kAogNumClasses];
  uint32_t max_mat;      // literal class: max (3+k+diag)*R
  uint32_t max_diag;     // literal class
  uint32_t max_rows_band;  // band classes: max rows
  uint32_t max_qlen_band;
  unsigned long long cells;  // reference-equivalent DP cells of the batch (GCUPS numerator)
  unsigned long long cls_cells[kAogNumClasses];   // per kernel class: cells
  unsigned long long cls_bytes[kAogNumClasses];   // per kernel class: algorithmic input+descriptor+score bytes
  unsigned long long cls_blocks[kAogNumClasses];  // per kernel class: block triples written
};

struct AogShape {
  int diag, k, two_sided, qB, tB, rows, cls, bucket;
  long long cells;
};

// mode: 0 = thread + literal kernels, 1 = thread + band + literal, 2 = literal only (tests)
__device__ __forceinline__ AogShape aog_shape(int qLen, int tLen, int k_in, int mode) {
  AogShape s;
  s.diag = imax(1, imin(qLen, tLen));
  int k0 = imin(s.diag, k_in);
  s.two_sided = (s.diag + 2 * k0 < imax(qLen, tLen)) ? 1 : 0;
  s.k = s.two_sided ? k0 : 2 * k0;
  s.qB = imin(s.diag + s.k, qLen + 1);
  s.tB = imin(s.diag + s.k, tLen + 1);
  s.rows = s.tB - 1;
  // reference-equivalent cells: prefix rows x band (+ suffix rows x band when two-sided)
  long long w = imin(2 * s.k + 1, s.qB);
  s.cells = (long long)s.rows * w;
  if (s.two_sided) {
    int tLow = imax(0, tLen - s.diag - s.k - 2);
    s.cells += (long long)(tLen + 1 - tLow) * (2 * s.k + 1);
  }
  if (mode != 2 && !s.two_sided && s.k >= 2 && s.k <= 14 && (s.k & 1) == 0 && s.rows <= kAogThreadMaxRows) {
    s.cls = s.k / 2 - 1;
    s.bucket = kAogBuckets - 1 - imin(s.rows >> 2, kAogBuckets - 1);
  } else if (mode == 1 && !s.two_sided && s.k >= 1 && 2 * s.k + 2 <= 256 && qLen <= 4000) {
    int need = 2 * s.k + 2;
    s.cls = kAogClsBand1 + (need <= 32 ? 0 : need <= 64 ? 1 : need <= 128 ? 2 : 3);
    s.bucket = kAogBuckets - 1 - imin(s.rows >> 4, kAogBuckets - 1);
  } else {
    s.cls = kAogClsLiteral;
    s.bucket = kAogBuckets - 1 - imin(s.rows >> 4, kAogBuckets - 1);
  }
  return s;
}

// Warp-aggregated shared-memory add: lanes with the same key elect a leader that adds the group's count.
__device__ __forceinline__ uint32_t aog_group_rank(uint32_t key, bool valid, uint32_t *leader_lane, uint32_t *group_size) {
#ifdef LRA_EMU
  // emulator: ballot-based equivalent of __match_any_sync
  uint32_t peers = 0;
  for (int l = 0; l < 32; l++) {
    uint32_t k2 = __shfl_sync(0xffffffffu, key, l);
    int v2 = __shfl_sync(0xffffffffu, valid ? 1 : 0, l);
    if (v2 && valid && k2 == key) peers |= 1u << l;
  }
#else
  uint32_t peers = __match_any_sync(0xffffffffu, valid ? key : 0xFFFFFFFFu);
  if (!valid) peers = 0;
#endif
  const uint32_t lane = threadIdx.x & 31;
  *leader_lane = peers ? (uint32_t)(__ffs((int)peers) - 1) : lane;
  *group_size = (uint32_t)__popc(peers);
  return (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

// classify: per-block histogram in shared memory (warp-aggregated), flushed once per block
__global__ void __launch_bounds__(256) aog_classify_kernel(AogBatch b, AogPlan *plan, uint32_t *bin_of_job, int use_band) {
  __shared__ uint32_t s_hist[kAogBins];
  __shared__ uint32_t s_cells[kAogNumClasses], s_bytes[kAogNumClasses];
  __shared__ uint32_t s_maxmat, s_maxdiag, s_maxrows, s_maxq;
  for (int i = threadIdx.x; i < kAogBins; i += blockDim.x) s_hist[i] = 0;
  if (threadIdx.x < kAogNumClasses) { s_cells[threadIdx.x] = 0; s_bytes[threadIdx.x] = 0; }
  if (threadIdx.x == 0) { s_maxmat = 0; s_maxdiag = 0; s_maxrows = 0; s_maxq = 0; }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = false;
  uint32_t bin = 0;
  AogShape s;
  s.cls = 0; s.cells = 0;
  int qLen = 0, tLen = 0;
  if (j < b.n_jobs) {
    qLen = b.q_len[j]; tLen = b.t_len[j];
    if (qLen < 0 || tLen < 0 || (uint64_t)b.q_off[j] + (uint64_t)qLen > b.q.n || (uint64_t)b.t_off[j] + (uint64_t)tLen > b.t.n) {
      // outside the domain: reported as LRA_B200_EINVAL by the host, job is skipped
      atomicOr(b.err, 8);
      bin_of_job[j] = 0xFFFFFFFFu;
      b.score[j] = 0; b.n_blocks[j] = 0; b.block_off[j] = 0;
    } else {
      valid = true;
      s = aog_shape(qLen, tLen, b.k[j], use_band);
      bin = (uint32_t)(s.cls * kAogBuckets + s.bucket);
      bin_of_job[j] = bin;
    }
  }
  uint32_t leader, gsz;
  const uint32_t rank = aog_group_rank(bin, valid, &leader, &gsz);
  if (valid && rank == 0) atomicAdd(&s_hist[bin], gsz);
  if (valid) {
    // per-class work counters (fit 32 bits per block: <= 256 jobs x <= 2^22 cells)
    atomicAdd(&s_cells[s.cls], (uint32_t)s.cells);
    // algorithmic bytes per job: 2-bit windows + 20 B SoA descriptor + 4 B score (+12 B per block, counted at output)
    atomicAdd(&s_bytes[s.cls], (uint32_t)((qLen + 3) / 4 + (tLen + 3) / 4 + 20 + 4));
    if (s.cls == kAogClsLiteral) {
      atomicMax(&s_maxmat, (uint32_t)((3 + s.k + s.diag) * (2 * s.k + 3)));
      atomicMax(&s_maxdiag, (uint32_t)s.diag);
    } else if (s.cls >= kAogClsBand1) {
      atomicMax(&s_maxrows, (uint32_t)s.rows);
      atomicMax(&s_maxq, (uint32_t)qLen);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kAogBins; i += blockDim.x) {
    const uint32_t v = s_hist[i];
    if (v) atomicAdd(&plan->hist[i], v);
  }
  if (threadIdx.x < kAogNumClasses) {
    const uint32_t c = s_cells[threadIdx.x];
    if (c) { atomicAdd(&plan->cls_cells[threadIdx.x], (unsigned long long)c); atomicAdd(&plan->cells, (unsigned long long)c); }
    if (s_bytes[threadIdx.x]) atomicAdd(&plan->cls_bytes[threadIdx.x], (unsigned long long)s_bytes[threadIdx.x]);
  }
  if (threadIdx.x == 0) {
    if (s_maxmat) atomicMax(&plan->max_mat, s_maxmat);
    if (s_maxdiag) atomicMax(&plan->max_diag, s_maxdiag);
    if (s_maxrows) atomicMax(&plan->max_rows_band, s_maxrows);
    if (s_maxq) atomicMax(&plan->max_qlen_band, s_maxq);
  }
}

// one block of kAogBins/ITEMS threads: exclusive scan of the histogram
__global__ void aog_scan_kernel(AogPlan *plan) {
  __shared__ uint32_t part[32];
  const int tid = threadIdx.x;              // blockDim.x == 512
  constexpr int ITEMS = (kAogBins + 511) / 512;
  uint32_t v[ITEMS];
  uint32_t sum = 0;
  for (int x = 0; x < ITEMS; x++) {
    int i = tid * ITEMS + x;
    v[x] = i < kAogBins ? plan->hist[i] : 0u;
    sum += v[x];
  }
  uint32_t inc = sum;
  const int lane = tid & 31, wid = tid >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) part[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t p = lane < 16 ? part[lane] : 0u;
    uint32_t pi = p;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xffffffffu, pi, o);
      if (lane >= o) pi += n;
    }
    part[lane] = pi - p;  // exclusive over warps
  }
  __syncthreads();
  uint32_t run = part[wid] + inc - sum;
  for (int x = 0; x < ITEMS; x++) {
    int i = tid * ITEMS + x;
    if (i < kAogBins) { plan->bin_start[i] = run; plan->cursor[i] = run; }
    run += v[x];
  }
  if (tid == 511) plan->bin_start[kAogBins] = run;
}

// scatter: each block counts its jobs per bin in shared memory, reserves one range per non-empty bin with a single
// global atomic, then places its jobs (order inside a bin is arbitrary; results are written back by job id).
__global__ void __launch_bounds__(256) aog_scatter_kernel(int n_jobs, AogPlan *plan, const uint32_t *bin_of_job, uint32_t *sorted) {
  __shared__ uint32_t s_cnt[kAogBins];
  __shared__ uint32_t s_base[kAogBins];
  for (int i = threadIdx.x; i < kAogBins; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t bin = 0xFFFFFFFFu;
  if (j < n_jobs) bin = bin_of_job[j];
  const bool valid = bin != 0xFFFFFFFFu;
  uint32_t leader, gsz;
  const uint32_t rank = aog_group_rank(bin, valid, &leader, &gsz);
  uint32_t first = 0;
  if (valid && rank == 0) first = atomicAdd(&s_cnt[bin], gsz);
  first = __shfl_sync(0xffffffffu, first, (int)leader);
  __syncthreads();
  for (int i = threadIdx.x; i < kAogBins; i += blockDim.x) {
    const uint32_t c = s_cnt[i];
    if (c) s_base[i] = atomicAdd(&plan->cursor[i], c);
  }
  __syncthreads();
  if (valid) sorted[s_base[bin] + first + rank] = (uint32_t)j;
}

// ---------------------------------------------------------------------------------------------------- output helper
// Warp-aggregated reservation of `n` block slots per lane.  Returns the lane's first slot (or ~0ull on overflow).
__device__ __forceinline__ unsigned long long aog_reserve_blocks(const AogBatch &b, int n, int lane,
                                                                  unsigned long long *cls_counter) {
  int inc = n;
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  int total = __shfl_sync(0xffffffffu, inc, 31);
  unsigned long long base = 0;
  if (lane == 0 && total > 0) {
    base = atomicAdd(b.block_cursor, (unsigned long long)total);
    atomicAdd(cls_counter, (unsigned long long)total);
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  bool over = base + (unsigned long long)total > b.block_cap;
  if (over) {
    if (lane == 0) atomicOr(b.err, 1);
    return ~0ull;
  }
  return base + (unsigned long long)(inc - n);
}

// ---------------------------------------------------------------------------------------------------- thread per job
template <int K> struct AogBits { typedef uint64_t type; };
template <> struct AogBits<2> { typedef uint32_t type; };
template <> struct AogBits<4> { typedef uint32_t type; };
template <> struct AogBits<6> { typedef uint32_t type; };

template <int K>
__global__ void __launch_bounds__(128) aog_thread_kernel(AogBatch b, AogPlan *plan, const uint32_t *sorted) {
  typedef typename AogBits<K>::type BT;
  constexpr int W = 2 * K + 1;
  constexpr int cls = K / 2 - 1;
  constexpr BT kOdd = (BT)0x5555555555555555ull;
  const int lane = threadIdx.x & 31;
  const uint32_t begin = plan->bin_start[cls * kAogBuckets];
  const uint32_t end = plan->bin_start[(cls + 1) * kAogBuckets];
  BT tb[kAogThreadMaxRows + 1];

  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&plan->work[cls], 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (begin + base >= end) break;
    const uint32_t idx = begin + base + lane;
    const bool active = idx < end;
    int job = 0, qLen = 1, tLen = 1, score = 0, nb = 0, qB = 1, tB = 1;
    if (active) {
      job = (int)sorted[idx];
      qLen = b.q_len[job]; tLen = b.t_len[job];
      const int diag = imax(1, imin(qLen, tLen));
      qB = imin(diag + K, qLen + 1);
      tB = imin(diag + K, tLen + 1);
      // (0,K+1) keeps its boundary value only if no rail loop overwrote it (AffineOneGapAlign.h:248-306)
      const bool keep0 = !((qLen >= tLen && diag - K - 1 >= 0) || (qLen <= tLen && diag >= 2));
      const int m = b.m, mm = b.mm, indel = b.indel;

      int prev[W + 1];
#pragma unroll
      for (int c = 0; c <= W; c++) {
        int i = c - K;
        prev[c] = (i < 0 || i > K) ? kMissing : indel * i;
      }
      SeqStream qs, ts;
      qs.init(b.q, b.q_off[job]);
      ts.init(b.t, b.t_off[job]);
      BT qw = 0, qn = 0;
      int qnext = 1;  // next 1-based query index to be pulled from the stream
#pragma unroll
      for (int c = K; c <= 2 * K; c++) {
        int code = 0;
        if (qnext <= qLen) { code = qs.next(); }
        qnext++;
        qw |= (BT)(code & 3) << (2 * c);
        qn |= (BT)(code == 4 ? 1 : 0) << (2 * c);
      }
      const int rows = tB - 1;
      for (int j = 1; j <= rows; j++) {
        const int tc = ts.next();
        BT e;
        if (tc == 4) e = qn;
        else {
          BT x = qw ^ ((BT)tc * kOdd);
          e = ~(x | (x >> 1)) & kOdd & ~qn;
        }
        int run = (j == K + 1 && keep0) ? indel * (K + 1) : kMissing;
        BT bits = 0;
#pragma unroll
        for (int c = 0; c < W; c++) {
          const int sM = prev[c] + (((e >> (2 * c)) & 1) ? m : mm);
          const int sD = prev[c + 1] + indel;
          const int sI = run + indel;
          const int best = imax(sI, imax(sD, sM));
          const int arrow = (best == sI) ? AR_LEFT : ((best == sD) ? AR_DOWN : AR_DIAG);
          bits |= (BT)arrow << (2 * c);
          prev[c] = best;
          run = best;
        }
        tb[j] = bits;
        qw >>= 2; qn >>= 2;
        int code = 0;
        if (qnext <= qLen) { code = qs.next(); }
        qnext++;
        qw |= (BT)(code & 3) << (4 * K);
        qn |= (BT)(code == 4 ? 1 : 0) << (4 * K);
      }
      const int cstar = (qB - 1) - (tB - 1) + K;
#pragma unroll
      for (int c = 0; c < W; c++) if (c == cstar) score = prev[c];
      // pass 1: count blocks
      {
        int i = qB - 1, j = tB - 1, run = 0;
        while (i > 0 && j > 0) {
          const int a = (int)((tb[j] >> (2 * (i - j + K))) & 3);
          if (a == AR_DIAG) { run++; i--; j--; }
          else { if (run) { nb++; run = 0; } if (a == AR_LEFT) i--; else j--; }
        }
        if (run) nb++;
      }
    }
    const unsigned long long slot = aog_reserve_blocks(b, nb, lane, &plan->cls_blocks[cls]);
    if (active) {
      b.score[job] = score;
      b.n_blocks[job] = nb;
      b.block_off[job] = slot;
      if (slot != ~0ull && nb > 0) {
        // pass 2: write blocks; traceback meets them last-to-first
        uint32_t *out = b.blocks + 3ull * slot;
        int i = qB - 1, j = tB - 1, run = 0, r = nb - 1;
        while (i > 0 && j > 0) {
          const int a = (int)((tb[j] >> (2 * (i - j + K))) & 3);
          if (a == AR_DIAG) { run++; i--; j--; }
          else {
            if (run) { out[3 * r] = (uint32_t)i; out[3 * r + 1] = (uint32_t)j; out[3 * r + 2] = (uint32_t)run; r--; run = 0; }
            if (a == AR_LEFT) i--; else j--;
          }
        }
        if (run) { out[3 * r] = (uint32_t)i; out[3 * r + 1] = (uint32_t)j; out[3 * r + 2] = (uint32_t)run; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------- warp literal
struct AogLiteralScratch {
  unsigned char *base;       // slabs, one per resident warp
  unsigned long long slab_bytes;
  uint32_t max_mat, max_diag;
};

__host__ __device__ inline unsigned long long aog_literal_slab_bytes(uint32_t max_mat, uint32_t max_diag) {
  unsigned long long mat = ((unsigned long long)max_mat + 15ull) & ~15ull;
  unsigned long long dg = ((unsigned long long)max_diag + 1ull + 15ull) & ~15ull;
  // ps, ss (int32) ; pp, sp (int8) ; lmax,lidx,umax,uidx (int32) ; reversed blocks (3 x uint32 per diag)
  return mat * 4 * 2 + mat * 2 + dg * 4 * 4 + dg * 12 + 64;
}

// Fill the cells i in [ilo, ihi) of one matrix row in parallel.  `mat`/`path` are the flat arrays; idx_of(i,j) the
// reference's index formula.  Returns nothing; all lanes must call.  closeRow / closeCol are the free-gap candidates
// (kNegInf when absent).  For the prefix matrix the per-row (lower) or per-column (upper) maxima are maintained.
struct AogRowCtx {
  int32_t *mat; int8_t *path; long matSize;
  int R, k, qShift, tShift;   // index = (j - tShift)*R + ((i - qShift) - (j - tShift)) + k + 1
  int m, mm, indel;
  int *err;
};
__device__ __forceinline__ long aog_idx(const AogRowCtx &c, int i, int j) {
  long jj = j - c.tShift, ii = i - c.qShift;
  long x = jj * c.R + (ii - jj) + c.k + 1;
  if (x < 0 || x >= c.matSize) { atomicOr(c.err, 2); x = 0; }
  return x;
}

constexpr int kAogLitC = 8;  // up to 256 computed cells per row

template <bool SUFFIX>
__device__ __forceinline__ void aog_literal_row(const AogRowCtx &cx, const SeqView &q, uint32_t qoff, int tcode, int j, int ilo,
                                                int ihi, int lane, bool useDelClose, int delCloseVal, bool useInsClose,
                                                const int32_t *umax_r,
                                                // prefix-only bookkeeping
                                                bool doLower, int lowerLimit, int32_t *lmax, int32_t *lidx,
                                                bool doUpper, int tLen, int diag, int32_t *umax, int32_t *uidx) {
  const int n = ihi - ilo;
  if (n <= 0) return;
  const int C = (n + 31) >> 5;
  const int indel = cx.indel;
  int sM[kAogLitC], sD[kAogLitC], sC1[kAogLitC], sC2[kAogLitC], L[kAogLitC];
  const int x0 = lane * C;
  int a_last = kNegInf;
#pragma unroll
  for (int x = 0; x < kAogLitC; x++) {
    sM[x] = sD[x] = L[x] = kNegInf;
    sC1[x] = sC2[x] = kMissing;  // an absent close-gap candidate still takes part as MISSING (AffineOneGapAlign.h:477-478)
    if (x < C) {
      const int i = ilo + x0 + x;
      if (i < ihi) {
        const int qc = seq_code(q, (uint64_t)qoff + (uint64_t)(i - 1));
        sM[x] = cx.mat[aog_idx(cx, i - 1, j - 1)] + (qc == tcode ? cx.m : cx.mm);
        sD[x] = cx.mat[aog_idx(cx, i, j - 1)] + indel;
        if (SUFFIX) {
          if (useDelClose) sC1[x] = delCloseVal;
          if (useInsClose) sC2[x] = umax_r[i];
        }
        int t = imax(imax(sM[x], sD[x]), imax(sC1[x], sC2[x]));
        L[x] = (x == 0) ? t : imax(t, L[x - 1] + indel);
        a_last = L[x];
      } else if (x > 0) {
        // keep the chain going so that a_last is the value at the lane's last REAL cell only
      }
    }
  }
  // lanes whose cells are all beyond ihi contribute nothing
  const int d = C * indel;
  const int nreal = imin(imax(n - x0, 0), C);
  // value entering lane l is max(left + l*d, max_{l'<l}(A(l') + (l-1-l')*d)); A of an incomplete lane is never consumed
  int v = (nreal == C) ? a_last - (lane + 1) * d : kNegInf;
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = imax(v, u);
  }
  int excl = __shfl_up_sync(0xffffffffu, v, 1);
  if (lane == 0) excl = kNegInf;
  const int left = cx.mat[aog_idx(cx, ilo - 1, j)];
  const int carry = lane * d + imax(left, excl);
  int prevS = carry;
  int lbest = kNegInf, lbi = 0;
#pragma unroll
  for (int x = 0; x < kAogLitC; x++) {
    if (x < C) {
      const int i = ilo + x0 + x;
      if (i < ihi) {
        const int S = imax(L[x], carry + (x + 1) * indel);
        const int sI = prevS + indel;
        int arrow;
        if (S == sI) arrow = AR_LEFT;
        else if (S == sD[x]) arrow = AR_DOWN;
        else if (S == sM[x]) arrow = AR_DIAG;
        else if (SUFFIX && S == sC1[x]) arrow = AR_GAPLEFT;
        else arrow = AR_GAPDOWN;
        const long ix = aog_idx(cx, i, j);
        cx.mat[ix] = S;
        cx.path[ix] = (int8_t)arrow;
        prevS = S;
        if (!SUFFIX) {
          if (doLower && i < lowerLimit && S >= lbest) { lbest = S; lbi = i; }
          if (doUpper && j < tLen && i < diag + 1 && S > umax[i]) { umax[i] = S; uidx[i] = j; }
        }
      }
    }
  }
  if (!SUFFIX && doLower) {
    // arg-max over the row with "last i wins" on ties (>= in increasing i, AffineOneGapAlign.h:347-352)
    for (int o = 16; o > 0; o >>= 1) {
      int ov = __shfl_down_sync(0xffffffffu, lbest, o);
      int oi = __shfl_down_sync(0xffffffffu, lbi, o);
      if (ov > lbest || (ov == lbest && oi > lbi)) { lbest = ov; lbi = oi; }
    }
    if (lane == 0 && lbest != kNegInf && lbest >= lmax[j]) { lmax[j] = lbest; lidx[j] = lbi; }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(128) aog_warp_literal_kernel(AogBatch b, AogPlan *plan, const uint32_t *sorted,
                                                              AogLiteralScratch sc) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t begin = plan->bin_start[kAogClsLiteral * kAogBuckets];
  const uint32_t end = plan->bin_start[(kAogClsLiteral + 1) * kAogBuckets];
  unsigned char *slab = sc.base + (unsigned long long)warp_global * sc.slab_bytes;
  const unsigned long long matA = ((unsigned long long)sc.max_mat + 15ull) & ~15ull;
  const unsigned long long dgA = ((unsigned long long)sc.max_diag + 1ull + 15ull) & ~15ull;
  int32_t *ps = (int32_t *)slab;
  int32_t *ss = ps + matA;
  int8_t *pp = (int8_t *)(ss + matA);
  int8_t *sp = pp + matA;
  int32_t *lmax = (int32_t *)(sp + matA);
  int32_t *lidx = lmax + dgA, *umax = lidx + dgA, *uidx = umax + dgA;
  uint32_t *rblk = (uint32_t *)(uidx + dgA);

  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&plan->work[kAogClsLiteral], 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (begin + w >= end) break;
    const int job = (int)sorted[begin + w];
    const int qLen = b.q_len[job], tLen = b.t_len[job];
    const uint32_t qoff = b.q_off[job], toff = b.t_off[job];
    const AogShape s = aog_shape(qLen, tLen, b.k[job], 2);
    const int diag = s.diag, k = s.k, R = 2 * k + 3;
    const long matSize = (long)(3 + k + diag) * R;
    const int m = b.m, mm = b.mm, indel = b.indel;
    const bool two = s.two_sided != 0;

    for (long x = lane; x < matSize; x += 32) { ps[x] = kMissing; pp[x] = -1; if (two) { ss[x] = kMissing; sp[x] = -1; } }
    for (int x = lane; x <= diag; x += 32) { lmax[x] = kMissing; lidx[x] = 0; umax[x] = kMissing; uidx[x] = 0; }
    __syncwarp();
    AogRowCtx P{ps, pp, matSize, R, k, 0, 0, m, mm, indel, b.err};
    // ---- prefix borders and rails, in the reference's order (later loops overwrite earlier cells)
    for (int i = 1 + lane; i < k + 1; i += 32) { long x = aog_idx(P, i, 0); ps[x] = indel * i; pp[x] = AR_LEFT; }
    __syncwarp();
    for (int j = 1 + lane; j <= k + 1; j += 32) { long x = aog_idx(P, 0, j); ps[x] = indel * j; pp[x] = AR_DOWN; }
    __syncwarp();
    if (lane == 0) { long x = aog_idx(P, 0, 0); ps[x] = 0; pp[x] = AR_DONE; }
    __syncwarp();
    if (qLen >= tLen) {
      for (int i = lane; i <= diag - k - 1; i += 32) { long x = aog_idx(P, i, i + k + 1); ps[x] = kMissing; pp[x] = AR_BORDER; }
      __syncwarp();
      for (int i = 1 + lane; i < diag + k - 1; i += 32) { long x = aog_idx(P, i + k + 1, i); ps[x] = kMissing; pp[x] = AR_BORDER; }
      if (lane == 0) { lmax[0] = 0; lidx[0] = 0; }
      __syncwarp();
    }
    if (qLen <= tLen) {
      for (int j = lane; j < diag - 1; j += 32) { long x = aog_idx(P, j + k + 1, j); ps[x] = kMissing; pp[x] = AR_BORDER; }
      __syncwarp();
      for (int j = 1 + lane; j < diag + k; j += 32) { long x = aog_idx(P, j - k - 1, j); ps[x] = kMissing; pp[x] = AR_BORDER; }
      if (lane == 0) { umax[0] = 0; uidx[0] = 0; }
      __syncwarp();
    }
    const int qB = s.qB, tB = s.tB;
    const bool doLower = two && qLen > tLen, doUpper = two && tLen > qLen;
    for (int j = 1; j < tB; j++) {
      const int tcode = seq_code(b.t, (uint64_t)toff + (uint64_t)(j - 1));
      aog_literal_row<false>(P, b.q, qoff, tcode, j, imax(1, j - k), imin(qB, j + k + 1), lane, false, 0, false, nullptr,
                             doLower, qLen - k, lmax, lidx, doUpper, tLen, diag, umax, uidx);
    }
    // ---- traceback state (lane 0 walks; results broadcast afterwards)
    int score = -1, nb = 0, bad = 0;
    int ti = 0, tj = 0;
    if (two) {
      const int qStart = imax(0, qLen - diag), qEnd = qLen + 1;
      const int tStart = imax(0, tLen - diag), tEnd = tLen + 1;
      const int tLow = imax(0, tLen - diag - k - 2), qLow = imax(0, qLen - diag - k - 1);
      AogRowCtx S{ss, sp, matSize, R, k, qLow, tLow, m, mm, indel, b.err};
      if (qLen >= tLen) {
        for (int i = qLow + lane; i < qStart + k + 1; i += 32) { long x = aog_idx(S, i, 0); ss[x] = lmax[0]; sp[x] = AR_GAPLEFT; }
        __syncwarp();
        for (int u = lane; u < diag; u += 32) { int i = qLow + u, j = 1 + u; long x = aog_idx(S, i, j); ss[x] = lmax[j]; sp[x] = AR_GAPLEFT; }
        __syncwarp();
        for (int j = tStart + 1 + lane; j < tEnd - k; j += 32) { int i = qStart + (j - tStart - 1); long x = aog_idx(S, i + k + 1, j); ss[x] = kMissing; sp[x] = AR_BORDER; }
        __syncwarp();
      }
      if (qLen <= tLen) {
        for (int j = tLow + lane; j < tStart + k + 2; j += 32) { long x = aog_idx(S, qStart, j); ss[x] = umax[0]; sp[x] = AR_GAPDOWN; }
        __syncwarp();
        for (int j = tStart + 1 + lane; j < tEnd; j += 32) { int i = qStart + 1 + (j - tStart - 1); long x = aog_idx(S, i, j - k - 1); ss[x] = umax[i]; sp[x] = AR_GAPDOWN; }
        __syncwarp();
        for (int j = tStart + lane; j < tEnd - k - 1; j += 32) { int i = qStart + (j - tStart); long x = aog_idx(S, i, j + k + 1); ss[x] = kMissing; sp[x] = AR_BORDER; }
        __syncwarp();
      }
      for (int j = tLow + 1; j < tEnd; j++) {
        const int doff = diag + 1 - (tEnd - j);
        const int tcode = seq_code(b.t, (uint64_t)toff + (uint64_t)(j - 1));
        const bool useDel = qLen >= tLen, useIns = tLen > qLen;
        aog_literal_row<true>(S, b.q, qoff, tcode, j, imax(qLow + 1, qStart + doff - k), imin(qEnd, qStart + doff + k + 1), lane,
                              useDel, useDel ? lmax[imin(j, diag)] : 0, useIns, umax,
                              false, 0, nullptr, nullptr, false, 0, 0, nullptr, nullptr);
      }
      if (lane == 0) {
        int i = qLen, j = tLen;
        int arrow = sp[aog_idx(S, i, j)];
        score = ss[aog_idx(S, i, j)];
        int run = 0, guard = 0;
        while (arrow != AR_DONE && arrow != AR_GAPDOWN && arrow != AR_GAPLEFT && i >= 0 && j >= 0) {
          if (arrow == AR_DIAG) { run++; i--; j--; }
          else {
            if (run) { rblk[3 * nb] = (uint32_t)i; rblk[3 * nb + 1] = (uint32_t)j; rblk[3 * nb + 2] = (uint32_t)run; nb++; run = 0; }
            if (arrow == AR_LEFT) i--; else if (arrow == AR_DOWN) j--; else if (++guard > 2) { bad = 1; break; }
          }
          if (i >= 0 && j >= 0) arrow = sp[aog_idx(S, i, j)];
        }
        if (run) { rblk[3 * nb] = (uint32_t)i; rblk[3 * nb + 1] = (uint32_t)j; rblk[3 * nb + 2] = (uint32_t)run; nb++; run = 0; }
        if (i < 0 || j < 0) bad = 1;
        if (!bad) {
          if (arrow == AR_GAPDOWN) { if (i > diag) bad = 1; else j = uidx[i]; }
          if (arrow == AR_GAPLEFT) { if (j > diag) bad = 1; else i = lidx[j]; }
        }
        ti = i; tj = j;
      }
    } else if (lane == 0) {
      ti = qB - 1; tj = tB - 1;
      score = ps[aog_idx(P, ti, tj)];
    }
    if (lane == 0 && !bad) {
      int i = ti, j = tj;
      int arrow = pp[aog_idx(P, i, j)];
      int run = 0, guard = 0;
      // a diagonal run may continue across the suffix/prefix seam only if no gap op separates them; the reference
      // pushes a gap op at the seam whenever the suffix ended in a close-gap arrow, so runs never merge there.
      while (arrow != AR_BORDER && arrow != AR_DONE && i >= 0 && j >= 0) {
        if (arrow == AR_DIAG) { run++; i--; j--; }
        else {
          if (run) { rblk[3 * nb] = (uint32_t)i; rblk[3 * nb + 1] = (uint32_t)j; rblk[3 * nb + 2] = (uint32_t)run; nb++; run = 0; }
          if (arrow == AR_LEFT) i--; else if (arrow == AR_DOWN) j--; else if (arrow == AR_GAPLEFT || arrow == AR_GAPDOWN) break;
          else if (++guard > 2) { bad = 1; break; }
        }
        if (i < 0 || j < 0) break;
        arrow = pp[aog_idx(P, i, j)];
      }
      if (run) { rblk[3 * nb] = (uint32_t)i; rblk[3 * nb + 1] = (uint32_t)j; rblk[3 * nb + 2] = (uint32_t)run; nb++; }
      // the reference replays the op list from wherever the walk stopped (normally the origin): make positions relative
      if (i != 0 || j != 0)
        for (int r = 0; r < nb; r++) { rblk[3 * r] -= (uint32_t)i; rblk[3 * r + 1] -= (uint32_t)j; }
    }
    nb = __shfl_sync(0xffffffffu, nb, 0);
    score = __shfl_sync(0xffffffffu, score, 0);
    bad = __shfl_sync(0xffffffffu, bad, 0);
    if (bad) { if (lane == 0) atomicOr(b.err, 4); nb = 0; }
    unsigned long long slot = aog_reserve_blocks(b, lane == 0 ? nb : 0, lane, &plan->cls_blocks[kAogClsLiteral]);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (lane == 0) { b.score[job] = score; b.n_blocks[job] = nb; b.block_off[job] = slot; }
    __syncwarp();
    if (slot != ~0ull) {
      uint32_t *out = b.blocks + 3ull * slot;
      for (int r = lane; r < nb; r += 32) {
        const int s2 = nb - 1 - r;
        out[3 * r] = rblk[3 * s2]; out[3 * r + 1] = rblk[3 * s2 + 1]; out[3 * r + 2] = rblk[3 * s2 + 2];
      }
    }
    __syncwarp();
  }
}

}  // namespace lra
