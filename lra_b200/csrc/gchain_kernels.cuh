// a24: GlobalChain<Fragment,Endpoint> over a PrioritySearchTree<Endpoint> (reference GlobalChain.h:88-189, PrioritySearchTree.h:47-291),
// batched over independent chaining problems.  `lra` never calls it; it is kept as the stand-alone primitive the
// reference ships (driver: TestGlobalChain.cpp).
//
// Mapping: ONE PROBLEM PER THREAD, a literal replay.  The sweep is a chain of dependent tree updates (every start point queries the tree
// that all earlier end points have modified), and the tree is not a search tree on its key: it is built over the end points in (x, y)
// order but keyed by y as an unsigned int, so queries return what this exact code returns.  std::sort's tie order between a start point
// and an end point at the same (x, y) decides whether two touching fragments chain (introsort.cuh).  Parallelism = the problems of a batch.
#pragma once
#include "lra_common.cuh"
#include "introsort.cuh"

namespace lra {

struct GcEndpoint { int32_t x, y, frag_side, score; };          // frag_side = fragment << 1 | (1 if End)
struct GcVertex { uint32_t left, right, leaf, medianKey, maxKey; int32_t pointIndex, maxScoreNode; };

struct GcBatch {
  int n_prob;
  const unsigned long long *frag_off;   // [n_prob + 1]
  const int32_t *frag;                  // [total][4] xl, yl, xh, yh
  int32_t *score;                       // [total] in: own score; out: score of the best chain ending in the fragment
  int32_t *prev;                        // [total] predecessor (problem-local index) or -1
  int32_t *chain;                       // [total] the optimal chain of problem p at chain[frag_off[p] ..], first to last
  int32_t *chain_len;                   // [n_prob]
  GcEndpoint *ep;                       // scratch [2 * total]
  GcVertex *tree;                       // scratch [4 * total]
};

struct GcLess { __device__ __forceinline__ bool operator()(const GcEndpoint &a, const GcEndpoint &b) const { return a.x != b.x ? a.x < b.x : a.y < b.y; } };

__global__ void __launch_bounds__(64) gchain_kernel(GcBatch b) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= b.n_prob) return;
  const unsigned long long o = b.frag_off[p];
  const int n = (int)(b.frag_off[p + 1] - o);
  b.chain_len[p] = 0;
  if (n == 0) return;
  const int32_t *F = b.frag + 4 * o;
  int32_t *score = b.score + o, *prev = b.prev + o, *chain = b.chain + o;
  GcEndpoint *ep = b.ep + 2 * o;
  GcVertex *tree = b.tree + 4 * o;
  const int m = 2 * n;
  for (int i = 0; i < n; i++) {      // FragmentSetToEndpoints
    ep[2 * i] = GcEndpoint{F[4 * i], F[4 * i + 1], i << 1, 0};
    ep[2 * i + 1] = GcEndpoint{F[4 * i + 2], F[4 * i + 3], (i << 1) | 1, 0};
    prev[i] = -1;
  }
  std_sort_replay(ep, m, GcLess());
  for (int i = 0; i < 2 * m - 1; i++) tree[i] = GcVertex{0u, 0u, 0u, 0u, 0u, -1, -1};
  // CreateTree (PrioritySearchTree.h:69-128): the recursion as an explicit stack; `ret` carries the returned key
  {
    int sCur[48], sStart[48], sEnd[48], sState[48];
    int sp = 0;
    uint32_t it = 0, ret = 0;
    sCur[0] = 0; sStart[0] = 0; sEnd[0] = m; sState[0] = 0; sp = 1;
    while (sp > 0) {
      const int cur = sCur[sp - 1], start = sStart[sp - 1], end = sEnd[sp - 1], median = (end + start) / 2;
      if (sState[sp - 1] == 0) {
        tree[cur].medianKey = (uint32_t)ep[median].y;
        if (end == start) { tree[cur].pointIndex = start; ret = tree[cur].medianKey; sp--; continue; }
        if (end - start == 1) { tree[cur].leaf = 1; tree[cur].medianKey = (uint32_t)ep[start].y; tree[cur].pointIndex = start; ret = tree[cur].medianKey; sp--; continue; }
        tree[cur].leaf = 0;
        tree[cur].left = ++it;
        sState[sp - 1] = 1;
        sCur[sp] = (int)it; sStart[sp] = start; sEnd[sp] = median; sState[sp] = 0; sp++;
      } else if (sState[sp - 1] == 1) {
        tree[cur].medianKey = ret;
        tree[cur].right = ++it;
        sState[sp - 1] = 2;
        sCur[sp] = (int)it; sStart[sp] = median; sEnd[sp] = end; sState[sp] = 0; sp++;
      } else { tree[cur].maxKey = ret; sp--; }
    }
  }
  int maxEp = 0;
  bool found = false;
  for (int q = 0; q < m; q++) {
    const int f = ep[q].frag_side >> 1;
    if ((ep[q].frag_side & 1) == 0) {
      // FindIndexOfMaxPoint (:130-199, :269-281): depth-first, left subtree before right; a strictly larger score replaces the candidate
      int maxIndex = 0, maxValue = -1;
      bool ok = false;
      if (tree[0].maxScoreNode != -1) {
        const uint32_t maxKey = (uint32_t)ep[q].y;
        uint32_t stack[64];
        int sp = 0;
        stack[sp++] = 0u;
        while (sp > 0) {
          const uint32_t cur = stack[--sp];
          const int msn = tree[cur].maxScoreNode;
          if (msn == -1) continue;
          if ((uint32_t)ep[msn].y < maxKey) {
            if (ep[msn].score > maxValue) { maxValue = ep[msn].score; maxIndex = msn; ok = true; }
            continue;
          }
          if (!tree[cur].leaf) {
            if (maxKey <= tree[cur].medianKey) stack[sp++] = tree[cur].left;
            else { stack[sp++] = tree[cur].right; stack[sp++] = tree[cur].left; }
          }
        }
      }
      if (ok) { const int fPrev = ep[maxIndex].frag_side >> 1; prev[f] = fPrev; score[f] = score[fPrev] + score[f]; }
      else prev[f] = -1;
    } else {
      ep[q].score = score[f];
      {   // Activate (:236-267)
        int pointIndex = q;
        const int pointScore = ep[q].score;
        const uint32_t pointKey = (uint32_t)ep[q].y;
        uint32_t cur = 0;
        while (pointIndex != -1 && tree[cur].leaf == 0) {
          const int msn = tree[cur].maxScoreNode;
          if (msn == -1 || ep[msn].score <= pointScore) { tree[cur].maxScoreNode = pointIndex; pointIndex = msn; }
          cur = (pointKey <= tree[cur].medianKey) ? tree[cur].left : tree[cur].right;
        }
      }
      if (!found || score[ep[maxEp].frag_side >> 1] < score[f]) { maxEp = q; found = true; }
    }
  }
  int len = 0;
  if (found) {
    int f = ep[maxEp].frag_side >> 1;
    while (f != -1 && len < n) { chain[len++] = f; f = prev[f]; }
    for (int i = 0; i < len / 2; i++) { const int32_t t = chain[i]; chain[i] = chain[len - 1 - i]; chain[len - 1 - i] = t; }
  }
  b.chain_len[p] = len;
}

}  // namespace lra
