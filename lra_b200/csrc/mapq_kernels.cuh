// a22: the per-read tail of MapRead after the statistics -- SegAlignmentGroup::SetFromSegAlignment (reference Alignment.h:944-983),
// AlignmentsOrder::Update (Alignment.h:1024-1062: std::sort of the new alignments by (value, NumOfAnchors0) descending -- replayed, ties are
// observable -- and the primary / secondary flags) and SimpleMapQV (Mapping_ultility.h:497-589), batched over reads.
// SimpleMapQV calls logf on data-dependent arguments; anything that reaches the output must come from the HOST libm, so
// the host side of the call computes logf(value / globalK) per segment and the (int)(4.343f * logf(len) + .499f) penalty per read and the
// kernel does the rest: binary32 products in the reference's order, never fused, and the x86-64 float -> int conversion (cvttss2si).
// One read per thread: a read has a handful of alignments.
#pragma once
#include "lra_common.cuh"
#include "introsort.cuh"

namespace lra {

struct MapqBatch {
  int n_reads, bypass, read_type;
  const int32_t *grp_off;       // [n_reads + 1] groups (alignments) of every read
  const int32_t *seg_off;       // [groups + 1] segments of every group
  const int32_t *upd_off;       // [n_reads + 1]
  const int32_t *update_at;     // number of the read's groups that exist at each Update call, ascending
  const float *value; const int32_t *n0, *n1, *nm, *nmm, *ndel, *nins; const uint8_t *strand;
  const float *logv;            // per segment: value > 3 ? logf(value / globalK) : 0   (host libm)
  const int32_t *lenpen;        // per read: (int)(4.343f * logf(len) + .499f), len = alignments ranked at the end   (host libm)
  int32_t *flag, *typeofaln; uint8_t *issec, *supp;   // per segment, in / out
  int32_t *mapq;
  uint8_t *g_issec; float *g_value; int32_t *g_n0, *g_n1, *g_nm, *order;   // per group
};

struct MapqLess {
  const float *v; const int32_t *n0;
  __device__ __forceinline__ bool operator()(const int32_t &i, const int32_t &j) const { return v[i] != v[j] ? v[i] > v[j] : n0[i] > n0[j]; }
};

__device__ __forceinline__ int mapq_f2i(float v) {          // (int) of a float as x86-64 compiles it
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return (int)0x80000000;
  return (int)v;
}

__global__ void __launch_bounds__(64) mapq_kernel(MapqBatch b) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= b.n_reads) return;
  const int g0 = b.grp_off[r], G = b.grp_off[r + 1] - g0;
  const int32_t *so = b.seg_off + g0;
  float *gv = b.g_value + g0; int32_t *gn0 = b.g_n0 + g0, *gn1 = b.g_n1 + g0, *gnm = b.g_nm + 4 * g0, *order = b.order + g0;
  uint8_t *gsec = b.g_issec + g0;
  const int32_t *ua = b.update_at + b.upd_off[r];
  const int nu = b.upd_off[r + 1] - b.upd_off[r];
  int oldend = 0, u = 0;
  for (int s = so[0]; s < so[G]; s++) b.mapq[s] = 0;
  for (int g = 0; g < G; g++) order[g] = -1;
  for (int g = 0; g <= G; g++) {
    while (u < nu && ua[u] == g) {
      if (g > oldend) {
        for (int i = oldend; i < g; i++) order[i] = i;
        std_sort_replay(order + oldend, g - oldend, MapqLess{gv, gn0});
        gsec[order[oldend]] = 0;
        for (int i = oldend + 1; i < g; i++) gsec[order[i]] = 1;
        oldend = g;
        for (int i = 0; i < g; i++)
          if (gsec[i] == 1) for (int z = so[i]; z < so[i + 1]; z++) { b.flag[z] |= 0x100; if (b.typeofaln[z] != 3) b.typeofaln[z] = 2; }
      }
      u++;
    }
    if (g == G) break;
    const int a = so[g], e = so[g + 1];
    gsec[g] = 0; gv[g] = 0.0f; gn0[g] = 0; gn1[g] = 0; gnm[4 * g] = gnm[4 * g + 1] = gnm[4 * g + 2] = gnm[4 * g + 3] = 0;
    if (e > a) {
      gsec[g] = b.issec[a]; gn0[g] = b.n0[a];
      float v = 0.0f; int pry = 0;
      for (int s = a; s < e; s++) {
        gn1[g] += b.n1[s]; gnm[4 * g] += b.nm[s]; gnm[4 * g + 1] += b.nmm[s]; gnm[4 * g + 2] += b.ndel[s]; gnm[4 * g + 3] += b.nins[s];
        v = __fadd_rn(v, b.value[s]);
        if (b.supp[s] == 0) pry++;
      }
      gv[g] = v;
      if (pry == 0) b.supp[a] = 0;
      for (int s = a; s < e; s++) {
        if (s > a) b.issec[s] = gsec[g];
        if (b.strand[s] == 1) b.flag[s] |= 0x10;
        if (b.supp[s] == 1) b.flag[s] |= 0x800;
      }
    }
  }
  const int len = oldend;
  const float q_coef = (b.bypass && b.read_type == 1) ? 4.0f : ((b.bypass && b.read_type == 0) ? 30.0f : 1.0f);
  if (len >= 1) {
    const int g = order[0];
    float x = 0.0f, y = 1.0f;
    if (len > 1) {
      x = __fdiv_rn(gv[order[1]], gv[g]);
      if (b.bypass) y = __fdiv_rn((float)gn0[g], (float)gn0[order[1]]);
    }
    const float omx = __fsub_rn(1.0f, x);
    for (int s = so[g]; s < so[g + 1]; s++) {
      const int N0 = b.n0[s];
      float pen;
      if (!b.bypass) { pen = __fmul_rn(N0 > 20 ? 1.0f : 0.05f, (float)N0); pen = __fmul_rn(N0 >= 5 ? 1.0f : 0.1f, pen); }
      else { pen = __fmul_rn(N0 > 10 ? 1.0f : 0.05f, (float)N0); pen = __fmul_rn(N0 >= 5 ? 1.0f : 0.02f, pen); }
      const int den = b.nmm[s] + b.ndel[s] + b.nins[s];
      float identity = den == 0 ? 1.0f : __fdiv_rn((float)b.nm[s], (float)den);
      identity = identity < 1.0f ? identity : 1.0f;
      const float l = b.logv[s];
      long long m;
      if (len == 1) {
        if (!b.bypass) m = mapq_f2i(__fmul_rn(__fmul_rn(__fmul_rn(pen, q_coef), l), identity));
        else m = mapq_f2i(__fmul_rn(__fmul_rn(pen, q_coef), identity));
      } else {
        if (x >= 0.990f) m = mapq_f2i(__fmul_rn(__fmul_rn(__fmul_rn(pen, omx), y), identity));
        else if (!b.bypass) m = mapq_f2i(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(pen, q_coef), omx), l), y), identity));
        else m = mapq_f2i(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(pen, q_coef), omx), y), identity));
        m -= (long long)b.lenpen[r];
      }
      m = m > 0 ? m : 0;
      int q = (int)(m < 60 ? m : 60);
      if (len == 2 && q == 0) q = 1;
      b.mapq[s] = q;
    }
  }
}

}  // namespace lra
