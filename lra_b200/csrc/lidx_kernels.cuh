// a12: LocalIndex::IndexSeq (reference MMIndex.h:200-245), batched over 2048-base windows of any number of sequences
// (the two strands of every read of a batch, or the contigs of the genome for <ref>.gli).
//   StoreMinimizers_noncanonical<LocalTuple,SmallTuple> (MinCount.h:181-338)   per window: (k,w) minimizers of the forward strand
//   std::sort on the 20-bit tuple (TupleOps.h:30-32)                            libstdc++ introsort: unstable, tie order observable
//   RemoveFrequent (MMIndex.h:69-85)                                            tuples occurring >= maxFreq times in the window are dropped
// A LocalTuple is one uint32: tuple in bits 0..19, window-relative position in bits 20..31 (the <ref>.gli layout).
//
// Mapping: ONE WARP PER WINDOW, the window's tuples in 8 KB of shared memory.
//   1. all lanes: the k-mer code of every position, straight from the packed 2-bit arena (two word loads per position);
//   2. lane 0: the reference's stateful scan (active minimizer, slot-ordered rescans, N handling) over the codes, in place
//      (emission n never overtakes the codes still needed: n <= p - w + 1);
//   3. all lanes: bitonic sort on (tuple, position).  When all tuples of the window are distinct (about 4 windows in 5 on
//      random sequence) ANY correct sort equals std::sort; otherwise lane 0 replays libstdc++'s introsort on the saved
//      unsorted list, because the order it leaves equal tuples in decides the order of the refined anchors downstream;
//   4. all lanes: frequency filter by bounded neighbour counts + ballot compaction.
// Windows are written to a sparse staging area (one slot per base) and compacted after a scan of the counts.
#pragma once
#include "lra_common.cuh"

namespace lra {

constexpr int kLidxMaxWindow = 2048;     // 1 << (LOCAL_POS_BITS - 1), TupleOps.h:18, MMIndex.h:110-117
constexpr int kLidxWarps = 4;

struct LidxBuild {
  SeqView seq;
  const unsigned long long *win_off;   // [n_win] arena position of the first base of each window
  const uint32_t *win_len;             // [n_win] bases in the window (<= 2048)
  int n_win;
  int k, w, max_freq;
  uint32_t *tmp;                       // staging: window wi owns tmp[win_off[wi] .. win_off[wi] + win_len[wi])
  unsigned long long *cnt;             // [n_win + 1] tuples kept per window; exclusive offsets after the scan
  uint32_t *mins;                      // dense result (compaction kernel)
};

__device__ __forceinline__ uint32_t lt_t(uint32_t v) { return v & 0xFFFFFu; }

// ---- libstdc++ std::sort on the 20-bit tuple, sequential (one lane), on a shared-memory array
__device__ __forceinline__ void lt_unguarded_linear_insert(uint32_t *v, int last) {
  const uint32_t val = v[last];
  int next = last - 1;
  while (lt_t(val) < lt_t(v[next])) { v[last] = v[next]; last = next; --next; }
  v[last] = val;
}
__device__ __forceinline__ void lt_insertion_sort(uint32_t *v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (lt_t(v[i]) < lt_t(v[first])) {
      const uint32_t val = v[i];
      for (int j = i; j > first; j--) v[j] = v[j - 1];
      v[first] = val;
    } else lt_unguarded_linear_insert(v, i);
  }
}
__device__ __forceinline__ void lt_adjust_heap(uint32_t *v, int first, int holeIndex, int len, uint32_t value) {
  const int topIndex = holeIndex;
  int secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (lt_t(v[first + secondChild]) < lt_t(v[first + secondChild - 1])) secondChild--;
    v[first + holeIndex] = v[first + secondChild];
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    v[first + holeIndex] = v[first + secondChild - 1];
    holeIndex = secondChild - 1;
  }
  int parent = (holeIndex - 1) / 2;       // __push_heap
  while (holeIndex > topIndex && lt_t(v[first + parent]) < lt_t(value)) { v[first + holeIndex] = v[first + parent]; holeIndex = parent; parent = (holeIndex - 1) / 2; }
  v[first + holeIndex] = value;
}
__device__ __forceinline__ void lt_heap_sort(uint32_t *v, int first, int last) {
  const int len = last - first;
  if (len >= 2)
    for (int parent = (len - 2) / 2;; parent--) { lt_adjust_heap(v, first, parent, len, v[first + parent]); if (parent == 0) break; }
  while (last - first > 1) {
    --last;
    const uint32_t val = v[last];
    v[last] = v[first];
    lt_adjust_heap(v, first, 0, last - first, val);
  }
}
__device__ __noinline__ void lt_introsort(uint32_t *v, int n) {
  if (n <= 1) return;
  int lg = 0;
  { unsigned x = (unsigned)n; while (x > 1) { x >>= 1; lg++; } }
  // __introsort_loop; the recursive call on the right part runs first, the left part is continued afterwards
  int stF[48], stL[48], stD[48];
  int sp = 0;
  stF[0] = 0; stL[0] = n; stD[0] = lg * 2; sp = 1;
  while (sp > 0) {
    --sp;
    int first = stF[sp], last = stL[sp], depth = stD[sp];
    while (last - first > 16) {
      if (depth == 0) { lt_heap_sort(v, first, last); break; }
      --depth;
      const int mid = first + (last - first) / 2;
      const int a = first + 1, bb = mid, c = last - 1;
      auto swp = [&](int i, int j) { const uint32_t x = v[i]; v[i] = v[j]; v[j] = x; };
      if (lt_t(v[a]) < lt_t(v[bb])) {
        if (lt_t(v[bb]) < lt_t(v[c])) swp(first, bb); else if (lt_t(v[a]) < lt_t(v[c])) swp(first, c); else swp(first, a);
      } else if (lt_t(v[a]) < lt_t(v[c])) swp(first, a);
      else if (lt_t(v[bb]) < lt_t(v[c])) swp(first, c);
      else swp(first, bb);
      int lo = first + 1, hi = last;
      const uint32_t pk = lt_t(v[first]);
      for (;;) {
        while (lt_t(v[lo]) < pk) ++lo;
        --hi;
        while (pk < lt_t(v[hi])) --hi;
        if (!(lo < hi)) break;
        swp(lo, hi);
        ++lo;
      }
      if (sp < 47) { stF[sp] = first; stL[sp] = lo; stD[sp] = depth; sp++; }
      first = lo;
    }
  }
  if (n > 16) { lt_insertion_sort(v, 0, 16); for (int i = 16; i != n; ++i) lt_unguarded_linear_insert(v, i); }
  else lt_insertion_sort(v, 0, n);
}

// lane 0: StoreMinimizers_noncanonical over the k-mer codes in v[0 .. len-k]; emissions overwrite v[0 .. n).  Returns n.
__device__ __noinline__ int lt_scan_window(uint32_t *v, const SeqView &seq, unsigned long long off, uint32_t seqLen, int k, int w) {
  if (seqLen < (uint32_t)k) return 0;
  const int windowSpan = w + k - 1;
  if (seqLen < (uint32_t)windowSpan) return 0;
  int nextValidWindowEnd = 0, nextValidWindowStart = 0;
  bool valid = false;
  while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
    valid = true;
    for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++)
      if (seq_code(seq, off + (unsigned long long)n) > 3) { nextValidWindowStart = n + 1; valid = false; }
  }
  if (!valid) return 0;
  nextValidWindowEnd = nextValidWindowStart + windowSpan;
  int n_out = 0;
  uint32_t actT = v[0], actP = 0;
  uint32_t p;
  for (p = 1; p < (uint32_t)w && p < seqLen - (uint32_t)k + 1; p++) {
    const uint32_t cur = v[p];
    if (cur < actT) { actT = cur; actP = p; }
  }
  const bool firstEmit = nextValidWindowEnd == windowSpan;
  // the first emission is stored when the main loop starts: by then v[0] has left every window that can still be rescanned
  uint32_t pendT = actT, pendP = actP;
  bool pend = firstEmit;
  for (p = (uint32_t)w; p < seqLen - (uint32_t)k + 1; p++) {
    const uint32_t cur = v[p];
    if ((uint32_t)nextValidWindowEnd == p + (uint32_t)k - 1) {
      if (seq_code(seq, off + p + (uint32_t)k - 1) <= 3) nextValidWindowEnd++;
      else {
        nextValidWindowStart = (int)(p + (uint32_t)k);
        valid = false;
        while ((uint32_t)nextValidWindowStart < seqLen - (uint32_t)windowSpan && !valid) {
          valid = true;
          for (int n = nextValidWindowStart; valid && n < nextValidWindowStart + windowSpan; n++)
            if (seq_code(seq, off + (unsigned long long)n) > 3) { nextValidWindowStart = n + 1; valid = false; }
        }
        if (!valid) break;
        nextValidWindowEnd = nextValidWindowStart + windowSpan;
      }
    }
    bool emit = false;
    if (p - (uint32_t)w >= actP) {
      // slot j of the reference's circular buffer holds the position p' in (p - w, p] with p' % w == j; first minimum in slot order
      const uint32_t base = p - (uint32_t)w + 1;
      uint32_t bestT = 0xFFFFFFFFu, bestP = 0;
      for (int j = 0; j < w; j++) {
        uint32_t pj = base + (((uint32_t)j + (uint32_t)w - base % (uint32_t)w) % (uint32_t)w);
        const uint32_t tj = v[pj];
        if (tj < bestT) { bestT = tj; bestP = pj; }
      }
      actT = bestT; actP = bestP;
      emit = (uint32_t)nextValidWindowEnd == p + (uint32_t)k;
    } else if (cur < actT) {
      actT = cur; actP = p;
      emit = (uint32_t)nextValidWindowEnd == p + (uint32_t)k;
    }
    // stores go to indices <= p - w + 1, which no later step reads as a code
    if (pend) { v[n_out++] = pendT | (pendP << 20); pend = false; }
    if (emit) v[n_out++] = actT | ((actP & 0xFFFu) << 20);      // at most p - w + 2 emissions so far: index <= p - w + 1
  }
  if (pend) v[n_out++] = pendT | (pendP << 20);
  return n_out;
}

__global__ void __launch_bounds__(32 * kLidxWarps) lidx_window_kernel(LidxBuild b) {
  __shared__ uint32_t sm[kLidxWarps][kLidxMaxWindow];
  __shared__ uint32_t nm_sm[kLidxWarps][68];
  // the window's packed bases (2048 / 16 words + the 16-byte alignment slack on both sides) and N mask, staged by the copy engine
  alignas(16) __shared__ uint32_t b2_sm[kLidxWarps][136];
  alignas(16) __shared__ uint32_t nmraw_sm[kLidxWarps][72];
  alignas(8) __shared__ unsigned long long mbar[kLidxWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wi = (int)(blockIdx.x * kLidxWarps + wib);
  if (wi >= b.n_win) return;
  uint32_t *v = sm[wib];
  const unsigned long long off = b.win_off[wi];
  const uint32_t len = b.win_len[wi];
  const int k = b.k;
  const unsigned long long wlo = (off >> 4) & ~3ull, mlo = (off >> 5) & ~3ull;      // first staged word of each plane (16-byte aligned)
  if (lane == 0) bulk_mbar_init(&mbar[wib], 1);
  __syncwarp();
  if (lane == 0) {
    bulk_mbar_expect(&mbar[wib], 136u * 4u + 72u * 4u);
    bulk_g2s(b2_sm[wib], b.seq.b2 + wlo, 136u * 4u, &mbar[wib]);
    bulk_g2s(nmraw_sm[wib], b.seq.nm + mlo, 72u * 4u, &mbar[wib]);
  }
  bulk_mbar_wait(&mbar[wib], 0u);
  __syncwarp();
  const uint32_t *b2w = b2_sm[wib], *nmr = nmraw_sm[wib];
  // 1. k-mer code of every position p <= len - k: bases p .. p+k-1, first base in the most significant digit (StoreTuple)
  const uint32_t kmask = (k >= 10) ? 0xFFFFFu : ((1u << (2 * k)) - 1u);
  if (len >= (uint32_t)k) {
    for (uint32_t p = lane; p + (uint32_t)k <= len; p += 32) {
      const unsigned long long a = off + p;
      const uint32_t w0 = b2w[(a >> 4) - wlo], w1 = b2w[(a >> 4) + 1 - wlo];
      const unsigned long long two = ((unsigned long long)w1 << 32) | w0;
      uint32_t bits = (uint32_t)(two >> ((uint32_t)(a & 15) * 2));     // base a in bits 0..1, a+1 in bits 2..3, ...
      uint32_t code = 0;
#pragma unroll
      for (int j = 0; j < 10; j++) if (j < k) code = (code << 2) | ((bits >> (2 * j)) & 3u);
      v[p] = code & kmask;
    }
  }
  __syncwarp();
  // 2. the minimizers of the window.
  // Lane-parallel form: where the minimum of every w-window is attained once, the reference's active minimizer at p is simply the
  // argmin of [p-w+1, p], and it emits exactly where the argmin changes (a rescan always moves it, a strictly smaller newcomer moves
  // it, nothing else does).  An emission is allowed iff the bases [p-w+1, p+k-1] hold no N and the N-free run they lie in starts
  // before seqLen - windowSpan (the reference only looks for a new run while nextValidWindowStart < seqLen - windowSpan).  A tied
  // minimum anywhere (low-complexity sequence) makes the active position depend on history: lane 0 then replays the reference's scan.
  uint32_t *stage = b.tmp + off;
  int n = 0;
  const int w = b.w;
  const int windowSpan = w + k - 1;
  bool tie = false;
  if (len > (uint32_t)windowSpan) {
    const int npos = (int)len - k + 1;
    // N mask of the window, bit p = base off+p is not A/C/G/T (65 words cover 2048 + 32 bases)
    uint32_t *nmw = nm_sm[wib];
    for (int j = lane; j < 65; j += 32) {
      const unsigned long long a = off + 32ull * j;
      const uint32_t w0 = nmr[(a >> 5) - mlo], w1 = nmr[(a >> 5) + 1 - mlo];
      const unsigned long long two = ((unsigned long long)w1 << 32) | w0;
      uint32_t bits = (uint32_t)(two >> (uint32_t)(a & 31));
      const int rem = (int)len - 32 * j;               // bases beyond the window count as N-free here (never inside a span)
      if (rem < 32) bits &= rem <= 0 ? 0u : ((1u << rem) - 1u);
      nmw[j] = bits;
    }
    __syncwarp();
    auto has_n = [&](int lo, int hi) {                 // any N in [lo, hi], 0 <= lo <= hi < len, hi - lo < 64
      const unsigned long long two = ((unsigned long long)nmw[(lo >> 5) + 1] << 32) | nmw[lo >> 5];
      const unsigned long long m = (two >> (lo & 31)) & ((hi - lo + 1) >= 64 ? ~0ull : ((1ull << (hi - lo + 1)) - 1ull));
      return m != 0ull;
    };
    auto argmin = [&](int p, bool &tied) {             // leftmost minimum of v[p-w+1 .. p]; tied: attained more than once
      uint32_t bt = v[p - w + 1]; int bp = p - w + 1; bool td = false;
      for (int j = p - w + 2; j <= p; j++) { const uint32_t x = v[j]; if (x < bt) { bt = x; bp = j; td = false; } else if (x == bt) td = true; }
      tied = td;
      return bp;
    };
    for (int base = w - 1; base < npos; base += 32) {
      const int p = base + lane;
      bool emit = false;
      uint32_t val = 0;
      if (p < npos) {
        bool t1 = false, t0 = false;
        const int a = argmin(p, t1);
        if (p == w - 1) { emit = true; }                                  // the first window: leftmost minimum, no history
        else { const int a0 = argmin(p - 1, t0); emit = a != a0; if (t1) tie = true; }
        if (emit) {
          const int lo = p - w + 1, hi = p + k - 1;
          bool ok = !has_n(lo, hi);
          if (ok && (uint32_t)lo == len - (uint32_t)windowSpan) ok = lo > 0 && !has_n(lo - 1, lo - 1);   // run must start before seqLen - windowSpan
          emit = ok;
          val = v[a] | (((uint32_t)a & 0xFFFu) << 20);
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, emit);
      if (emit) stage[n + __popc(m & ((1u << lane) - 1u))] = val;
      n += __popc(m);
    }
    tie = __any_sync(0xffffffffu, tie);
  }
  __syncwarp();
  if (tie) {
    if (lane == 0) n = lt_scan_window(v, b.seq, off, len, k, w);
    n = __shfl_sync(0xffffffffu, n, 0);
    __syncwarp();
    for (int i = lane; i < n; i += 32) stage[i] = v[i];          // the unsorted list, for the introsort replay
  } else {
    __syncwarp();
    for (int i = lane; i < n; i += 32) v[i] = stage[i];
  }
  __syncwarp();
  // 3. bitonic sort on (tuple << 12 | position)
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = lane; i < P; i += 32) { const uint32_t x = i < n ? v[i] : 0xFFFFFFFFu; v[i] = i < n ? ((x << 12) | (x >> 20)) : 0xFFFFFFFFu; }
  __syncwarp();
  for (int kk = 2; kk <= P; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int x = lane; x < (P >> 1); x += 32) {
        const int i = ((x & ~(j - 1)) << 1) | (x & (j - 1));      // lower index of the x-th compare-exchange pair of this stage
        const int ixj = i | j;
        const uint32_t a = v[i], c = v[ixj];
        const bool up = (i & kk) == 0;
        if ((a > c) == up) { v[i] = c; v[ixj] = a; }
      }
      __syncwarp();
    }
  }
  int dup = 0;
  for (int i = lane; i + 1 < n; i += 32) dup |= ((v[i] >> 12) == (v[i + 1] >> 12)) ? 1 : 0;
  dup = __any_sync(0xffffffffu, dup);
  __syncwarp();
  if (dup) {
    for (int i = lane; i < n; i += 32) v[i] = stage[i];
    __syncwarp();
    if (lane == 0) lt_introsort(v, n);
    __syncwarp();
  } else {
    for (int i = lane; i < n; i += 32) { const uint32_t x = v[i]; v[i] = (x >> 12) | (x << 20); }
    __syncwarp();
  }
  // 4. RemoveFrequent: an element stays iff its run of equal tuples is shorter than maxFreq
  const int mf = b.max_freq;
  int kept = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    bool keep = false;
    uint32_t x = 0;
    if (i < n) {
      x = v[i];
      const uint32_t t = lt_t(x);
      int run = 1;
      for (int j = i - 1; j >= 0 && run < mf && lt_t(v[j]) == t; j--) run++;
      for (int j = i + 1; j < n && run < mf && lt_t(v[j]) == t; j++) run++;
      keep = run < mf;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) stage[kept + __popc(m & ((1u << lane) - 1u))] = x;
    kept += __popc(m);
  }
  if (lane == 0) b.cnt[wi] = (unsigned long long)kept;
}

// after the scan of cnt: copy every window's tuples from the staging area to their dense position
__global__ void __launch_bounds__(256) lidx_compact_kernel(LidxBuild b) {
  const int lane = threadIdx.x & 31;
  const int wi = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (wi >= b.n_win) return;
  const unsigned long long o = b.cnt[wi];
  const int n = (int)(b.cnt[wi + 1] - o);
  const uint32_t *stage = b.tmp + b.win_off[wi];
  for (int i = lane; i < n; i += 32) b.mins[o + i] = stage[i];
}

}  // namespace lra
