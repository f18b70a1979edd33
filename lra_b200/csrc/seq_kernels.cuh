// Sequence packing kernels (a1 / a23 data model): ASCII -> 2-bit + non-ACGT mask (lra_common.cuh layout).
// Reference semantics: the comparison alphabet of seqMapN (SeqUtils.h:42-75): A/a,C/c,G/g,T/t -> 0..3, raw bytes
// 0..7 -> (b & 3), every other byte -> 4 ("N", equal only to another N).
#pragma once
#include "lra_common.cuh"

namespace lra {

__device__ __forceinline__ uint32_t ascii_code(uint32_t b) {
  const uint32_t u = b & 0xDFu;  // upper-case letters
  const bool acgt = (u == 0x41u) | (u == 0x43u) | (u == 0x47u) | (u == 0x54u);
  if (acgt) return ((u >> 1) ^ (u >> 2)) & 3u;
  if (b < 8u) return b & 3u;
  return 4u;
}

// One thread packs 32 bases: two 2-bit words and one mask word.  Input bytes are read with two 128-bit loads when the
// whole 32-byte group is inside the buffer (buffers are cudaMalloc'ed, so 32*t is 16-byte aligned).
__global__ void seq_pack_kernel(const uint8_t *__restrict__ ascii, uint64_t n, uint32_t *__restrict__ b2,
                                uint32_t *__restrict__ nm, uint64_t n_groups) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const uint64_t p0 = g * 32;
  uint32_t w[8];
  if (p0 + 32 <= n) {
    const uint4 *src = reinterpret_cast<const uint4 *>(ascii + p0);
    uint4 a = src[0], b = src[1];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t v = 0;
      for (int bb = 0; bb < 4; bb++) {
        uint64_t p = p0 + (uint64_t)i * 4 + bb;
        uint32_t c = p < n ? ascii[p] : (uint32_t)'N';
        v |= c << (8 * bb);
      }
      w[i] = v;
    }
  }
  uint32_t lo = 0, hi = 0, mask = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const uint32_t c = ascii_code((w[i >> 2] >> (8 * (i & 3))) & 0xFFu);
    const uint32_t two = c & 3u;
    if (i < 16) lo |= (c == 4u ? 0u : two) << (2 * i);
    else hi |= (c == 4u ? 0u : two) << (2 * (i - 16));
    mask |= (c == 4u ? 1u : 0u) << i;
  }
  b2[2 * g] = lo;
  b2[2 * g + 1] = hi;
  nm[g] = mask;
}

}  // namespace lra
