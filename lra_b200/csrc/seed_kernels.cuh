// a1-a5: the seeding prefix of MapRead (reference MapRead.h:169-203), batched over reads.
//   a1 CreateRC (SeqUtils.h:151-158)                    -> seq_revcomp_kernel (packed arena, per read)
//   a2 StoreMinimizers<GenomeTuple,Tuple> (MinCount.h:7-179)  -> seed_minimizers_kernel
//   a3 std::sort(readmm) (MapRead.h:185)                -> seed_sort_kernel (libstdc++ introsort, restated: its tie order is observable)
//   a4 CompareLists<GenomeTuple,Tuple> (CompareLists.h:8-151) -> seed_compare_kernel<EMIT> (count pass, scan, emit pass)
//   a5 SeparateMatchesByStrand (MapRead.h:109-150)      -> fused into the emit pass (strand flag per match)
//
// Mapping: ONE READ PER THREAD.  Every one of these reference routines is a sequential scan whose quirks (unmasked
// comparison in the first window, circular-buffer tie order, unstable sort, front/back galloping with unmasked skips and the
// resulting duplicate emissions) decide which anchors exist, so each thread replays its read literally; the parallelism is
// the 10^4..10^5 reads of a batch.  The only heavy memory traffic is a4: ~27 dependent probes per search into the global
// index image (3.2 GB for a 3 Gb genome) -- random 32-byte sectors out of HBM, the top levels of every search shared in L2.
#pragma once
#include "lra_common.cuh"

namespace lra {

constexpr unsigned long long kForMask = 0x7FFFFFFFFFFFFFFFull;
constexpr unsigned long long kRevMask = 0x8000000000000000ull;

// ---- a1: reverse complement of every read of a packed arena into a second arena with the same per-read offsets
__global__ void __launch_bounds__(256) seq_revcomp_kernel(SeqView in, const unsigned long long *read_off, const uint32_t *read_len, int n_reads,
                                                          uint32_t *out_b2, uint32_t *out_nm) {
  // one warp per read, one lane per 16 output bases: 16-base words of a read are disjoint between lanes only when the read
  // offset is 16-aligned; to stay general a lane builds whole OUTPUT words of the arena and only touches positions of its read
  const int lane = threadIdx.x & 31;
  const int r = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (r >= n_reads) return;
  const unsigned long long off = read_off[r];
  const unsigned long long len = read_len[r];
  const unsigned long long w0 = off >> 4, w1 = (off + len + 15) >> 4;
  for (unsigned long long wi = w0 + lane; wi < w1; wi += 32) {
    uint32_t word = 0, keep = 0;
    for (int j = 0; j < 16; j++) {
      const unsigned long long p = wi * 16 + j;
      if (p >= off && p < off + len) {
        const unsigned long long src = off + (len - 1 - (p - off));
        const int c = seq_code(in, src);
        word |= (uint32_t)(c == 4 ? 0 : (3 - c)) << (2 * j);
        keep |= 3u << (2 * j);
      }
    }
    // words at the read's ends may be shared with neighbouring reads: merge atomically
    if (keep == 0xFFFFFFFFu) out_b2[wi] = word;
    else { atomicAnd(&out_b2[wi], ~keep); atomicOr(&out_b2[wi], word); }
  }
  const unsigned long long m0 = off >> 5, m1 = (off + len + 31) >> 5;
  for (unsigned long long wi = m0 + lane; wi < m1; wi += 32) {
    uint32_t word = 0, keep = 0;
    for (int j = 0; j < 32; j++) {
      const unsigned long long p = wi * 32 + j;
      if (p >= off && p < off + len) {
        const unsigned long long src = off + (len - 1 - (p - off));
        word |= (uint32_t)(seq_code(in, src) == 4 ? 1 : 0) << j;
        keep |= 1u << j;
      }
    }
    if (keep == 0xFFFFFFFFu) out_nm[wi] = word;
    else { atomicAnd(&out_nm[wi], ~keep); atomicOr(&out_nm[wi], word); }
  }
}

struct SeedBatch {
  SeqView reads;                       // forward strands, packed
  const unsigned long long *read_off;  // [n_reads] arena offset of each read
  const uint32_t *read_len;            // [n_reads]
  int n_reads;
  int k, w;
  long long max_freq;
  // global index image (sorted by masked key), SoA
  const unsigned long long *idx_t;
  const uint32_t *idx_pos;
  long long n_idx;
  SeqView genome;                      // contigs concatenated in Header::pos order
  // minimizer scratch: region of read r starts at read_off[r] (at most one minimizer per base)
  unsigned long long *mm_t;
  uint32_t *mm_pos;
  uint32_t *mm_n;                      // [n_reads]
  // matches
  unsigned long long *match_cnt;       // [n_reads + 1] counts, then exclusive offsets after the scan
  unsigned long long *m_qt, *m_tt;
  uint32_t *m_qpos, *m_tpos;
  uint8_t *m_strand;
  unsigned long long match_cap;
  int *err;                            // bit0: match capacity exceeded
};

constexpr int kSeedMaxW = 64;

// ---- a2 (literal): the minimizers of seq[off, off + seqLen) into (ot, op); returns their number.  CANON: StoreMinimizers (MinCount.h:7-179, the smaller of
// the tuple and its reverse complement, strand in the top bit); !CANON: StoreMinimizers_noncanonical (MinCount.h:181-337, forward tuples only)
template <bool CANON>
__device__ __noinline__ uint32_t mm_scan(const SeqView &seq, unsigned long long off, uint32_t seqLen, int k, int w, unsigned long long *ot, uint32_t *op) {
  uint32_t n_out = 0;
  if (seqLen < (uint32_t)k) return 0;
  const int windowSpan = w + k - 1;
  if (seqLen < (uint32_t)windowSpan) return 0;
  unsigned long long m = 0;
  for (int i = 0; i < k; i++) { m <<= 2; m += 3; }
  int nextValidWindowEnd = 0, nextValidWindowStart = 0;
  bool valid = false;
  while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
    valid = true;
    for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++) {
      if (seqLen < (uint32_t)n) return 0;
      if (seq_code(seq, off + (unsigned long long)n) > 3) { nextValidWindowStart = n + 1; valid = false; }
    }
  }
  if (!valid) return 0;
  nextValidWindowEnd = nextValidWindowStart + windowSpan;
  SeqStream st;
  st.init(seq, off);
  unsigned long long cur = 0, curRC = 0;
  for (int p = 0; p <= k - 1; p++) { const int c = st.next(); cur <<= 2; cur += (unsigned long long)(c & 3) * (c != 4); }
  { unsigned long long a = cur; for (int i = 0; i < k; i++) { const unsigned long long least = ~(a & 3ull) & 3ull; a >>= 2; curRC <<= 2; curRC += least; } }
  unsigned long long ringT[kSeedMaxW];
  uint32_t ringP[kSeedMaxW];
  unsigned long long actT;
  uint32_t actP = 0;
  if (!CANON) actT = cur; else if ((cur & kForMask) < (curRC & kForMask)) actT = cur & kForMask; else actT = curRC | kRevMask;
  ringT[0] = actT; ringP[0] = 0;
  uint32_t p;
  for (p = 1; p < (uint32_t)w && p < seqLen - (uint32_t)k + 1; p++) {
    const int c = st.next();
    const unsigned long long n2 = (unsigned long long)(c & 3) * (c != 4);
    cur = ((cur << 2) & m) + n2;
    curRC >>= 2; curRC += ((~n2) & 3ull) << (2 * ((unsigned long long)k - 1));
    const unsigned long long ct = !CANON ? (cur & kForMask) : ((cur & kForMask) < (curRC & kForMask)) ? (cur & kForMask) : (curRC | kRevMask);
    if (ct < actT) { actT = ct; actP = p; }        // first window: unmasked comparison (MinCount.h:91)
    ringT[p % (uint32_t)w] = ct; ringP[p % (uint32_t)w] = p;
  }
  if (nextValidWindowEnd == windowSpan) { ot[n_out] = actT; op[n_out] = actP; n_out++; }
  for (p = (uint32_t)w; p < seqLen - (uint32_t)k + 1; p++) {
    const int c = st.next();   // code of seq[p+k-1]
    if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k - 1) {
      if (c <= 3) nextValidWindowEnd++;
      else {
        nextValidWindowStart = (int)(p + (uint32_t)k);
        valid = false;
        while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
          valid = true;
          for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++)
            if (seq_code(seq, off + (unsigned long long)n) > 3) { nextValidWindowStart = n + 1; valid = false; }
        }
        if (!valid) return n_out;
        nextValidWindowEnd = nextValidWindowStart + windowSpan;
      }
    }
    const unsigned long long n2 = (unsigned long long)(c & 3) * (c != 4);
    cur = ((cur << 2) & m) + n2;
    curRC >>= 2; curRC += ((~n2) & 3ull) << (2 * ((unsigned long long)k - 1));
    const unsigned long long ct = !CANON ? (cur & kForMask) : ((cur & kForMask) < (curRC & kForMask)) ? (cur & kForMask) : (curRC | kRevMask);
    ringT[p % (uint32_t)w] = ct; ringP[p % (uint32_t)w] = p;
    if (p - (uint32_t)w >= actP) {
      actT = ringT[0]; actP = ringP[0];
      for (int j = 1; j < w; j++) if ((ringT[j] & kForMask) < (actT & kForMask)) { actT = ringT[j]; actP = ringP[j]; }
      if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k) { ot[n_out] = actT; op[n_out] = actP; n_out++; }
    } else if ((ct & kForMask) < (actT & kForMask)) {
      actT = ct; actP = p;
      if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k) { ot[n_out] = actT; op[n_out] = actP; n_out++; }
    }
  }
  return n_out;
}

__global__ void __launch_bounds__(128) seed_minimizers_kernel(SeedBatch b) {   // one read per thread
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= b.n_reads) return;
  const unsigned long long off = b.read_off[r];
  b.mm_n[r] = mm_scan<true>(b.reads, off, b.read_len[r], b.k, b.w, b.mm_t + off, b.mm_pos + off);
}

// ---- a3: libstdc++ std::sort (introsort) on the masked key, in place, one read per thread
struct MmRef {
  unsigned long long *t; uint32_t *p;
  __device__ __forceinline__ unsigned long long key(long i) const { return t[i] & kForMask; }
  __device__ __forceinline__ void swap(long i, long j) const { const unsigned long long a = t[i]; t[i] = t[j]; t[j] = a; const uint32_t c = p[i]; p[i] = p[j]; p[j] = c; }
};
__device__ __forceinline__ void mm_unguarded_linear_insert(const MmRef &v, long last) {
  const unsigned long long vt = v.t[last]; const uint32_t vp = v.p[last];
  const unsigned long long vk = vt & kForMask;
  long next = last - 1;
  while (vk < v.key(next)) { v.t[last] = v.t[next]; v.p[last] = v.p[next]; last = next; --next; }
  v.t[last] = vt; v.p[last] = vp;
}
__device__ __forceinline__ void mm_insertion_sort(const MmRef &v, long first, long last) {
  if (first == last) return;
  for (long i = first + 1; i != last; ++i) {
    if (v.key(i) < v.key(first)) {
      const unsigned long long vt = v.t[i]; const uint32_t vp = v.p[i];
      for (long j = i; j > first; j--) { v.t[j] = v.t[j - 1]; v.p[j] = v.p[j - 1]; }
      v.t[first] = vt; v.p[first] = vp;
    } else mm_unguarded_linear_insert(v, i);
  }
}
__device__ __forceinline__ void mm_push_heap(const MmRef &v, long first, long holeIndex, long topIndex, unsigned long long vt, uint32_t vp) {
  long parent = (holeIndex - 1) / 2;
  const unsigned long long vk = vt & kForMask;
  while (holeIndex > topIndex && v.key(first + parent) < vk) {
    v.t[first + holeIndex] = v.t[first + parent]; v.p[first + holeIndex] = v.p[first + parent];
    holeIndex = parent; parent = (holeIndex - 1) / 2;
  }
  v.t[first + holeIndex] = vt; v.p[first + holeIndex] = vp;
}
__device__ __forceinline__ void mm_adjust_heap(const MmRef &v, long first, long holeIndex, long len, unsigned long long vt, uint32_t vp) {
  const long topIndex = holeIndex;
  long secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (v.key(first + secondChild) < v.key(first + secondChild - 1)) secondChild--;
    v.t[first + holeIndex] = v.t[first + secondChild]; v.p[first + holeIndex] = v.p[first + secondChild];
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    v.t[first + holeIndex] = v.t[first + secondChild - 1]; v.p[first + holeIndex] = v.p[first + secondChild - 1];
    holeIndex = secondChild - 1;
  }
  mm_push_heap(v, first, holeIndex, topIndex, vt, vp);
}
__device__ __forceinline__ void mm_heap_sort(const MmRef &v, long first, long last) {
  const long len = last - first;
  if (len >= 2)
    for (long parent = (len - 2) / 2;; parent--) { mm_adjust_heap(v, first, parent, len, v.t[first + parent], v.p[first + parent]); if (parent == 0) break; }
  while (last - first > 1) {
    --last;
    const unsigned long long vt = v.t[last]; const uint32_t vp = v.p[last];
    v.t[last] = v.t[first]; v.p[last] = v.p[first];
    mm_adjust_heap(v, first, 0, last - first, vt, vp);
  }
}

__device__ __noinline__ void mm_sort(const MmRef &v, long n) {
  if (n <= 1) return;
  long lg = 0;
  { unsigned long x = (unsigned long)n; while (x > 1) { x >>= 1; lg++; } }
  // __introsort_loop with an explicit stack for the recursive (right-hand) calls
  long stF[96], stL[96], stD[96];
  int sp = 0;
  stF[0] = 0; stL[0] = n; stD[0] = lg * 2; sp = 1;
  while (sp > 0) {
    --sp;
    long first = stF[sp], last = stL[sp], depth = stD[sp];
    while (last - first > 16) {
      if (depth == 0) { mm_heap_sort(v, first, last); break; }
      --depth;
      const long mid = first + (last - first) / 2;
      const long a = first + 1, bb = mid, c = last - 1;
      if (v.key(a) < v.key(bb)) {
        if (v.key(bb) < v.key(c)) v.swap(first, bb); else if (v.key(a) < v.key(c)) v.swap(first, c); else v.swap(first, a);
      } else if (v.key(a) < v.key(c)) v.swap(first, a);
      else if (v.key(bb) < v.key(c)) v.swap(first, c);
      else v.swap(first, bb);
      long lo = first + 1, hi = last;
      const unsigned long long pk = v.key(first);
      for (;;) {
        while (v.key(lo) < pk) ++lo;
        --hi;
        while (pk < v.key(hi)) --hi;
        if (!(lo < hi)) break;
        v.swap(lo, hi);
        ++lo;
      }
      // recurse on [lo, last) first (the reference's recursive call runs to completion before the loop continues), then
      // continue with [first, lo): push the left part, process the right part now
      if (sp < 95) { stF[sp] = first; stL[sp] = lo; stD[sp] = depth; sp++; }
      first = lo;
    }
  }
  if (n > 16) { mm_insertion_sort(v, 0, 16); for (long i = 16; i != n; ++i) mm_unguarded_linear_insert(v, i); }
  else mm_insertion_sort(v, 0, n);
}

__global__ void __launch_bounds__(128) seed_sort_kernel(SeedBatch b) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= b.n_reads) return;
  mm_sort(MmRef{b.mm_t + b.read_off[r], b.mm_pos + b.read_off[r]}, (long)b.mm_n[r]);
}

// ---- a4: literal CompareLists<GenomeTuple, Tuple> (CompareLists.h:8-146) of two sorted minimizer lists; push(qi, ti) receives every pair in the reference's
// order (the caller applies its own filter: none for the global index, the diagonal band for RefineSpace)
template <class Push>
__device__ void mm_compare(const unsigned long long *qt, long nq, const unsigned long long *tt, long nt, long long maxFreq, Push push) {
#define QK(i) (qt[i] & kForMask)
#define TK(i) (tt[i] & kForMask)
  if (nq != 0 && nt != 0) {
    long qs = 0, qe = nq - 1, ts = 0, te = nt;
    do {
      unsigned long long startGap = 0, endGap = 0;
      while (qs <= qe && QK(qs) < TK(ts)) qs++;
      if (qs >= qe) break;
      startGap = QK(qs) - TK(ts);
      if (qs == qe) endGap = startGap;
      else {
        while (qe > qs && te > ts && QK(qe) > TK(te - 1)) qe--;
        endGap = TK(te - 1) - QK(qe);
      }
      if (startGap == 0 || ((startGap & kForMask) > (endGap & kForMask))) {
        const long tsOrig = ts, qsOrig = qs;
        {
          long lo = ts, len = te - ts;
          const unsigned long long key = QK(qs);
          while (len > 0) { const long half = len >> 1, mid = lo + half; if (TK(mid) < key) { lo = mid + 1; len = len - half - 1; } else len = half; }
          ts = lo;
        }
        if (ts < nt && TK(ts) == QK(qs)) {
          const long tsStart = ts;
          long tsi = ts;
          while (tsi != te && QK(qs) == TK(tsi)) tsi++;
          const long qsStart = qs;
          while (qs < qe && QK(qs + 1) == QK(qs)) qs++;
          for (long ti = tsStart; ti != tsi; ti++)
            if (qs - qsStart < maxFreq)
              for (long qi = qsStart; qi <= qs; qi++) push(qi, ti);
        }
        while (ts < te && tt[ts] == tt[tsOrig]) ts++;     // unmasked, against the original positions (CompareLists.h:101-102)
        while (qs < qe && qt[qs] == qt[qsOrig]) qs++;
      } else {
        if (te != nt && TK(te - 1) == QK(qe)) { /* pass */ }
        else {
          long lo = ts, len = te - ts;
          const unsigned long long key = QK(qe);
          while (len > 0) { const long half = len >> 1, mid = lo + half; if (key < TK(mid)) len = half; else { lo = mid + 1; len = len - half - 1; } }
          te = lo;
        }
        const long teStart = te;
        long tei = te;
        while (tei > ts && TK(tei - 1) == QK(qe)) tei--;
        if (tei < teStart && teStart > 0) {
          const long qeStart = qe;
          while (qe > qs && QK(qe) == QK(qe - 1)) qe--;
          for (long ti = tei; ti < teStart; ti++)
            if (qeStart - qe < maxFreq)
              for (long qi = qe; qi <= qeStart; qi++) push(qi, ti);
        }
        te = tei;
      }
    } while (qs < qe && ts < te);
  }
#undef QK
#undef TK
}

// ---- a4 + a5 against the global index: EMIT == false counts, EMIT == true writes at match offsets
template <bool EMIT>
__global__ void __launch_bounds__(128) seed_compare_kernel(SeedBatch b) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= b.n_reads) return;
  const long nq = (long)b.mm_n[r];
  const long nt = (long)b.n_idx;
  const unsigned long long *qt = b.mm_t + b.read_off[r];
  const uint32_t *qpos = b.mm_pos + b.read_off[r];
  const unsigned long long *tt = b.idx_t;
  const uint32_t *tpos = b.idx_pos;
  unsigned long long n_out = 0;
  const unsigned long long obase = EMIT ? b.match_cnt[r] : 0ull;
  const unsigned long long roff = b.read_off[r];
  const int k = b.k;
  mm_compare(qt, nq, tt, nt, b.max_freq, [&](long qi, long ti) {
    if (EMIT) {
      const unsigned long long o = obase + n_out;
      if (o < b.match_cap) {
        b.m_qt[o] = qt[qi]; b.m_qpos[o] = qpos[qi]; b.m_tt[o] = tt[ti]; b.m_tpos[o] = tpos[ti];
        // a5: strncmp(read + q, genome + t, k) == 0 -> forward (0) else reverse (1)
        int differ = 0;
        for (int j = 0; j < k && !differ; j++)
          differ = seq_code(b.reads, roff + (unsigned long long)qpos[qi] + j) != seq_code(b.genome, (unsigned long long)tpos[ti] + j);
        b.m_strand[o] = (uint8_t)differ;
      }
    }
    n_out++;
  });
  if (!EMIT) b.match_cnt[r] = n_out;
}

// exclusive scan of match_cnt[0..n) in place, total in match_cnt[n]; single block (n is the number of reads of a batch)
__global__ void __launch_bounds__(1024) seed_scan_kernel(unsigned long long *cnt, int n, unsigned long long cap, int *err) {
  __shared__ unsigned long long part[1024];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int lo = tid * per, hi = (lo + per < n) ? lo + per : n;
  unsigned long long s = 0;
  for (int i = lo; i < hi; i++) s += cnt[i];
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; i++) { const unsigned long long v = part[i]; part[i] = run; run += v; }
    cnt[n] = run;
    if (run > cap) atomicOr(err, 1);
  }
  __syncthreads();
  unsigned long long run = part[tid];
  for (int i = lo; i < hi; i++) { const unsigned long long v = cnt[i]; cnt[i] = run; run += v; }
}

}  // namespace lra
