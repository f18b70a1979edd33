// TEST INFRASTRUCTURE ONLY: a small lock-step SIMT emulator so that the kernel logic in lra_b200/csrc/*.cuh
// (warp shuffles, ballots, block barriers, atomics, persistent work queues) can be executed by the CPU test
// suite in a container that has nvcc but no GPU.  Nothing in the product library includes this file.
//
// Model: every CUDA thread of a block is a ucontext fiber; fibers of one block are scheduled round-robin on the
// calling OS thread and yield inside collective operations until all participating lanes have arrived.  Blocks of a
// grid run one after another.  Divergent exits while other lanes wait in a full-mask collective deadlock here
// (reported), as they would be undefined behaviour on the device.
#pragma once
#include <ucontext.h>
#include <math.h>
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __device__
#define __host__
#define __global__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ inline __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }

namespace emu {

struct WarpState {
  uint32_t arrived = 0, left = 0;
  uint64_t val[32];
  uint32_t pred_bits = 0;
};

struct Thread {
  ucontext_t ctx;
  uint3 tid;
  int lane = 0, warp = 0;
  bool done = false;
  char *stack = nullptr;
};

struct BlockState {
  std::vector<Thread> th;
  std::vector<WarpState> warps;
  ucontext_t sched;
  int cur = -1;
  int bar_count = 0;
  uint64_t bar_gen = 0;
  int live = 0;
  std::vector<int> order;
  std::function<void()> body;
};

inline BlockState *&B() { static BlockState *b = nullptr; return b; }
inline long &g_order_mode() { static long m = getenv("LRA_EMU_ORDER") ? atol(getenv("LRA_EMU_ORDER")) : 0; return m; }
inline uint3 &g_blockIdx() { static uint3 v{0, 0, 0}; return v; }
inline dim3 &g_blockDim() { static dim3 v; return v; }
inline dim3 &g_gridDim() { static dim3 v; return v; }
inline unsigned char *&g_dyn_smem() { static unsigned char *p = nullptr; return p; }

inline Thread &self() { return B()->th[B()->cur]; }
inline void yield() { BlockState *b = B(); swapcontext(&b->th[b->cur].ctx, &b->sched); }

inline void trampoline() {
  BlockState *b = B();
  b->body();
  b->th[b->cur].done = true;
  b->live--;
  swapcontext(&b->th[b->cur].ctx, &b->sched);
}

static const size_t kStack = 512 * 1024;

template <class F>
void launch(dim3 grid, dim3 block, size_t dyn_smem, F &&body) {
  BlockState bs;
  int nth = (int)(block.x * block.y * block.z);
  bs.th.resize(nth);
  bs.warps.resize((nth + 31) / 32);
  std::vector<char> stacks((size_t)nth * kStack);
  std::vector<unsigned char> smem(dyn_smem + 16);
  g_dyn_smem() = smem.data();
  g_blockDim() = block;
  g_gridDim() = grid;
  bs.body = body;
  BlockState *saved = B();
  B() = &bs;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        g_blockIdx() = uint3{bx, by, bz};
        for (auto &w : bs.warps) w = WarpState();
        bs.bar_count = 0;
        bs.live = nth;
        for (int t = 0; t < nth; t++) {
          Thread &T = bs.th[t];
          T.done = false;
          T.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
          T.lane = t & 31;
          T.warp = t >> 5;
          getcontext(&T.ctx);
          T.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * kStack;
          T.ctx.uc_stack.ss_size = kStack;
          T.ctx.uc_link = &bs.sched;
          makecontext(&T.ctx, (void (*)())trampoline, 0);
        }
        long idle_rounds = 0;
        while (bs.live > 0) {
          int before = bs.live;
          // LRA_EMU_ORDER: 0 ascending (default), 1 descending, >= 2 a fresh pseudo-random permutation every round (seed) --
          // the order lanes run between two collectives is unspecified on hardware, tests replay kernels under several orders
          const long order_mode = g_order_mode();
          static unsigned long long rng = 0x9E3779B97F4A7C15ull;
          rng ^= (unsigned long long)order_mode;
          std::vector<int> &ord = bs.order;
          if ((int)ord.size() != nth) { ord.resize(nth); for (int t = 0; t < nth; t++) ord[t] = t; }
          if (order_mode >= 2)
            for (int t = nth - 1; t > 0; t--) { rng = rng * 6364136223846793005ull + 1442695040888963407ull; std::swap(ord[t], ord[(int)((rng >> 33) % (unsigned)(t + 1))]); }
          for (int k = 0; k < nth; k++) {
            const int t = order_mode == 1 ? nth - 1 - k : ord[k];
            if (bs.th[t].done) continue;
            bs.cur = t;
            swapcontext(&bs.sched, &bs.th[t].ctx);
          }
          if (bs.live == before) {
            if (++idle_rounds > 50000000L) { fprintf(stderr, "cuda_emu: deadlock suspected in block %u\n", bx); abort(); }
          } else idle_rounds = 0;
        }
      }
  B() = saved;
}

// ---- warp collectives: every lane in `mask` deposits a value, waits for the others, reads, leaves.
inline void warp_arrive(uint32_t mask, uint64_t v, int pred, uint64_t *out_vals, uint32_t *out_pred) {
  Thread &T = self();
  WarpState &W = B()->warps[T.warp];
  uint32_t bit = 1u << T.lane;
  // lanes of a partial last warp that do not exist never arrive: drop them from the mask
  int nth = (int)B()->th.size();
  int base = T.warp * 32;
  uint32_t exist = (base + 32 <= nth) ? 0xffffffffu : ((1u << (nth - base)) - 1u);
  mask &= exist;
  assert(mask & bit);
  while (W.left & bit) yield();  // previous collective on this lane not fully drained
  W.val[T.lane] = v;
  if (pred) W.pred_bits |= bit; else W.pred_bits &= ~bit;
  W.arrived |= bit;
  while ((W.arrived & mask) != mask) yield();
  if (out_vals) for (int l = 0; l < 32; l++) out_vals[l] = W.val[l];
  if (out_pred) *out_pred = W.pred_bits & mask;
  W.left |= bit;
  if ((W.left & mask) == mask) { W.arrived &= ~mask; W.left &= ~mask; }
  else while (W.left & bit) yield();
}

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace emu

#define threadIdx (emu::self().tid)
#define blockIdx (emu::g_blockIdx())
#define blockDim (emu::g_blockDim())
#define gridDim (emu::g_gridDim())
#define warpSize 32

template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  uint64_t vals[32];
  emu::warp_arrive(mask, emu::to_bits(v), 0, vals, nullptr);
  int lane = emu::self().lane;
  int s = (lane & ~(width - 1)) | (src & (width - 1));
  return emu::from_bits<T>(vals[s]);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  uint64_t vals[32];
  emu::warp_arrive(mask, emu::to_bits(v), 0, vals, nullptr);
  int lane = emu::self().lane;
  int s = lane - (int)delta;
  if (s < (lane & ~(width - 1))) s = lane;
  return emu::from_bits<T>(vals[s]);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  uint64_t vals[32];
  emu::warp_arrive(mask, emu::to_bits(v), 0, vals, nullptr);
  int lane = emu::self().lane;
  int s = lane + (int)delta;
  if (s > (lane | (width - 1))) s = lane;
  return emu::from_bits<T>(vals[s]);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  uint64_t vals[32];
  emu::warp_arrive(mask, emu::to_bits(v), 0, vals, nullptr);
  int lane = emu::self().lane;
  int s = lane ^ lanemask;
  if (s > (lane | (width - 1))) s = lane;
  return emu::from_bits<T>(vals[s]);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  uint32_t p;
  emu::warp_arrive(mask, 0, pred, nullptr, &p);
  return p;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) {
  uint32_t p;
  emu::Thread &T = emu::self();
  int nth = (int)emu::B()->th.size(), base = T.warp * 32;
  uint32_t exist = (base + 32 <= nth) ? 0xffffffffu : ((1u << (nth - base)) - 1u);
  emu::warp_arrive(mask, 0, pred, nullptr, &p);
  return p == (mask & exist);
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_arrive(mask, 0, 0, nullptr, nullptr); }
inline unsigned __activemask() { return 0xffffffffu; }
inline void __syncthreads() {
  emu::BlockState *b = emu::B();
  uint64_t gen = b->bar_gen;
  b->bar_count++;
  // threads that already returned do not take part (matches CUDA's behaviour for exited threads)
  while (b->bar_gen == gen) {
    if (b->bar_count >= b->live) { b->bar_count = 0; b->bar_gen++; break; }
    emu::yield();
  }
}
inline int __reduce_max_sync(unsigned mask, int v) {
  uint64_t vals[32];
  emu::warp_arrive(mask, emu::to_bits(v), 0, vals, nullptr);
  int m = v;
  int nth = (int)emu::B()->th.size(), base = emu::self().warp * 32;
  for (int l = 0; l < 32; l++) if (((mask >> l) & 1u) && base + l < nth) { int o = emu::from_bits<int>(vals[l]); if (o > m) m = o; }
  return m;
}
inline void __threadfence() {}
inline void __threadfence_block() {}

template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (unsigned)(v >> (s & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (unsigned)((v << (s & 31)) >> 32);
}
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __uint_as_float(unsigned int x) { float f; memcpy(&f, &x, 4); return f; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline float __fsqrt_rn(float a) { volatile float r = __builtin_sqrtf(a); return r; }
inline float __ll2float_rn(long long a) { volatile float r = (float)a; return r; }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
