// Common device-side definitions for the lra_b200 kernels (sm_100a).
//
// Everything under lra_b200/csrc/*.cuh is written so that it can be compiled twice:
//   * by nvcc for sm_100a (the product), and
//   * by g++ with -DLRA_EMU against tests/simt/cuda_emu.h, a lock-step SIMT emulator used ONLY by the CPU test
//     suite to exercise the kernel logic (warp shuffles, atomics, barriers) where no GPU exists.
// The emulator is test infrastructure; the product library contains no host implementation of any kernel.
#pragma once
#include <stdint.h>

#ifdef LRA_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

namespace lra {

// Sentinels.  The reference keeps scores in `long` with MISSING = INT_MIN (AffineOneGapAlign.h:29); the kernels use
// int32 with a sentinel far below any reachable score.  Every MISSING-derived value is the sentinel plus the same
// small offsets the reference adds, so all orderings and equalities between candidates are preserved.
constexpr int kMissing = -(1 << 29);
constexpr int kNegInf = -(1 << 30);  // identity for max() in scans; below every MISSING-derived value

enum Arrow : int { AR_DONE = 0, AR_LEFT = 1, AR_DOWN = 2, AR_DIAG = 3, AR_BORDER = 4, AR_GAPLEFT = 5, AR_GAPDOWN = 6 };

// Packed sequence in HBM: 2 bits per base (A,C,G,T = 0..3; base p in bits 2*(p&15) of word p>>4) plus a 1-bit-per-base
// mask of non-ACGT symbols (bit p&31 of word p>>5).  code = mask ? 4 : 2-bit value, which is exactly the reference's
// seqMapN comparison alphabet (SeqUtils.h:42-75) for ASCII input.  Both arrays are padded by >= 4 words.
struct SeqView {
  const uint32_t *b2;
  const uint32_t *nm;
  uint64_t n;
};

__device__ __forceinline__ int seq_code(const SeqView &s, uint64_t p) {
  uint32_t w = s.b2[p >> 4];
  uint32_t c = (w >> ((uint32_t)(p & 15) * 2)) & 3u;
  uint32_t nbit = (s.nm[p >> 5] >> (uint32_t)(p & 31)) & 1u;
  return nbit ? 4 : (int)c;
}

// Forward sequential reader over a packed sequence (one 32-bit load per 16 bases + one per 32 bases).
struct SeqStream {
  const uint32_t *b2;
  const uint32_t *nm;
  uint64_t pos;
  uint32_t w, nw;
  __device__ __forceinline__ void init(const SeqView &s, uint64_t p) {
    b2 = s.b2; nm = s.nm; pos = p;
    w = b2[p >> 4] >> ((uint32_t)(p & 15) * 2);
    nw = nm[p >> 5] >> (uint32_t)(p & 31);
  }
  __device__ __forceinline__ int next() {
    int code = (nw & 1u) ? 4 : (int)(w & 3u);
    pos++;
    if ((pos & 15) == 0) w = b2[pos >> 4]; else w >>= 2;
    if ((pos & 31) == 0) nw = nm[pos >> 5]; else nw >>= 1;
    return code;
  }
};

// ---- 1-D bulk copy global -> shared through the copy engine (cp.async.bulk, SASS UBLKCP) with an mbarrier for completion: no tensor map needed.
// dst, src and bytes must be multiples of 16.  One elected lane arms the barrier with the byte count and issues the copies; every lane that will
// read the tile waits on the barrier's phase.  (The emulator copies at issue time.)
__device__ __forceinline__ void bulk_mbar_init(unsigned long long *mbar, int arrivals) {
#ifndef LRA_EMU
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(arrivals));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
#else
  (void)mbar; (void)arrivals;
#endif
}
__device__ __forceinline__ void bulk_mbar_expect(unsigned long long *mbar, uint32_t bytes) {
#ifndef LRA_EMU
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
#else
  (void)mbar; (void)bytes;
#endif
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *mbar) {
#ifndef LRA_EMU
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst), a = (uint32_t)__cvta_generic_to_shared(mbar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d), "l"(gsrc), "r"(bytes), "r"(a) : "memory");
#else
  (void)mbar;
  for (uint32_t i = 0; i < bytes; i++) ((unsigned char *)smem_dst)[i] = ((const unsigned char *)gsrc)[i];
#endif
}
__device__ __forceinline__ void bulk_mbar_wait(unsigned long long *mbar, uint32_t phase) {
#ifndef LRA_EMU
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(a), "r"(phase) : "memory");
#else
  (void)mbar; (void)phase;
#endif
}

// 4-byte asynchronous global -> shared copy (LDGSTS): the prefetches of the row-pipeline kernel never pass through registers,
// so no scoreboard wait can end up in the dependent chain.  Completion: cp_async_wait_all() by the issuing thread, then a
// warp barrier before other lanes read.  (The emulator copies at issue time.)
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
#ifdef LRA_EMU
  *(uint32_t *)smem_dst = *(const uint32_t *)gsrc;
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef LRA_EMU
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

}  // namespace lra
