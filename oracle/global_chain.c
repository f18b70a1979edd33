/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the stand-alone chaining primitive named by the north star:
 *   GlobalChain<Fragment,Endpoint>          /root/reference/GlobalChain.h:88-189   (FragmentSetToEndpoints :65-86)
 *   PrioritySearchTree<Endpoint>            /root/reference/PrioritySearchTree.h:47-291  (CreateTree :69-128, FindIndexOfMaxPoint :130-199,
 *                                                                                          Activate :236-267)
 *   std::sort(endpoints, Endpoint::LessThan) -- libstdc++ introsort on (x, y); a start point and an end point may share (x, y), and the
 *   order introsort leaves them in decides whether the two fragments can chain, so the algorithm is restated (GCC 13.3 bits/stl_algo.h).
 * The tree is built over the endpoints in (x, y) order but keyed by y (GetKey, GlobalChain.h:56-58), with unsigned keys (KeyType):
 * it is not a search tree on y, and the queries return whatever this code returns -- which is what is reproduced.
 * `lra` itself never calls GlobalChain; the known answer of the reference's own driver TestGlobalChain.cpp:29-38
 * is the golden vector (tests/test_global_chain.py). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t x, y, frag, side; int32_t score; } gc_ep;      /* side: 0 Start, 1 End */
typedef struct { uint32_t left, right, leaf, medianKey, maxKey; int32_t pointIndex, maxScoreNode; } gc_vx;

#define EPLESS(a, b) ((a).x != (b).x ? (a).x < (b).x : (a).y < (b).y)
static inline void ep_swap(gc_ep *a, gc_ep *b) { gc_ep t = *a; *a = *b; *b = t; }
static void ep_unguarded_linear_insert(gc_ep *last) {
  gc_ep val = *last; gc_ep *next = last - 1;
  while (EPLESS(val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void ep_insertion_sort(gc_ep *first, gc_ep *last) {
  if (first == last) return;
  for (gc_ep *i = first + 1; i != last; ++i) {
    if (EPLESS(*i, *first)) { gc_ep val = *i; memmove(first + 1, first, (size_t)(i - first) * sizeof(gc_ep)); *first = val; }
    else ep_unguarded_linear_insert(i);
  }
}
static void ep_adjust_heap(gc_ep *first, long holeIndex, long len, gc_ep value) {
  const long topIndex = holeIndex;
  long secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (EPLESS(first[secondChild], first[secondChild - 1])) secondChild--;
    first[holeIndex] = first[secondChild]; holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) { secondChild = 2 * (secondChild + 1); first[holeIndex] = first[secondChild - 1]; holeIndex = secondChild - 1; }
  long parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && EPLESS(first[parent], value)) { first[holeIndex] = first[parent]; holeIndex = parent; parent = (holeIndex - 1) / 2; }
  first[holeIndex] = value;
}
static void ep_heap_sort(gc_ep *first, gc_ep *last) {
  long len = last - first;
  if (len >= 2) for (long parent = (len - 2) / 2;; parent--) { gc_ep v = first[parent]; ep_adjust_heap(first, parent, len, v); if (parent == 0) break; }
  while (last - first > 1) { --last; gc_ep v = *last; *last = *first; ep_adjust_heap(first, 0, last - first, v); }
}
static void ep_introsort_loop(gc_ep *first, gc_ep *last, long depth_limit) {
  while (last - first > 16) {
    if (depth_limit == 0) { ep_heap_sort(first, last); return; }
    --depth_limit;
    gc_ep *mid = first + (last - first) / 2;
    gc_ep *a = first + 1, *b = mid, *c = last - 1;
    if (EPLESS(*a, *b)) { if (EPLESS(*b, *c)) ep_swap(first, b); else if (EPLESS(*a, *c)) ep_swap(first, c); else ep_swap(first, a); }
    else if (EPLESS(*a, *c)) ep_swap(first, a);
    else if (EPLESS(*b, *c)) ep_swap(first, c);
    else ep_swap(first, b);
    gc_ep *lo = first + 1, *hi = last;
    for (;;) {
      while (EPLESS(*lo, *first)) ++lo;
      --hi;
      while (EPLESS(*first, *hi)) --hi;
      if (!(lo < hi)) break;
      ep_swap(lo, hi);
      ++lo;
    }
    ep_introsort_loop(lo, last, depth_limit);
    last = lo;
  }
}
static void ep_sort(gc_ep *v, long n) {
  if (n <= 1) return;
  long lg = 0; { unsigned long x = (unsigned long)n; while (x > 1) { x >>= 1; lg++; } }
  ep_introsort_loop(v, v + n, lg * 2);
  if (n > 16) { ep_insertion_sort(v, v + 16); for (gc_ep *i = v + 16; i != v + n; ++i) ep_unguarded_linear_insert(i); }
  else ep_insertion_sort(v, v + n);
}

static uint32_t pst_create(gc_vx *tree, const gc_ep *pts, int start, int end, uint32_t *it) {
  const int median = (end + start) / 2;
  const uint32_t cur = *it;
  tree[cur].medianKey = (uint32_t)pts[median].y;
  if (end == start) { tree[cur].pointIndex = start; return tree[cur].medianKey; }
  if (end - start == 1) { tree[cur].leaf = 1; tree[cur].medianKey = (uint32_t)pts[start].y; tree[cur].pointIndex = start; return tree[cur].medianKey; }
  tree[cur].leaf = 0;
  tree[cur].left = ++(*it);
  const uint32_t leftKey = pst_create(tree, pts, start, median, it);
  tree[cur].medianKey = leftKey;
  tree[cur].right = ++(*it);
  const uint32_t rightKey = pst_create(tree, pts, median, end, it);
  tree[cur].maxKey = rightKey;
  return rightKey;
}
static int pst_find(const gc_vx *tree, uint32_t cur, const gc_ep *pts, uint32_t maxKey, int *maxValue, int *maxIndex) {
  if (tree[cur].maxScoreNode == -1) return 0;
  if ((uint32_t)pts[tree[cur].maxScoreNode].y < maxKey) {
    if (pts[tree[cur].maxScoreNode].score > *maxValue) { *maxValue = pts[tree[cur].maxScoreNode].score; *maxIndex = tree[cur].maxScoreNode; return 1; }
    return 0;
  }
  if (!tree[cur].leaf) {
    if (maxKey <= tree[cur].medianKey) return pst_find(tree, tree[cur].left, pts, maxKey, maxValue, maxIndex);
    const int l = pst_find(tree, tree[cur].left, pts, maxKey, maxValue, maxIndex);
    const int r = pst_find(tree, tree[cur].right, pts, maxKey, maxValue, maxIndex);
    return l || r;
  }
  return 0;
}
static void pst_activate(gc_vx *tree, const gc_ep *pts, int pointIndex) {
  const int pointScore = pts[pointIndex].score;
  uint32_t cur = 0;
  const uint32_t pointKey = (uint32_t)pts[pointIndex].y;
  while (pointIndex != -1 && tree[cur].leaf == 0) {
    if (tree[cur].maxScoreNode == -1 || pts[tree[cur].maxScoreNode].score <= pointScore) {
      const int tmp = tree[cur].maxScoreNode;
      tree[cur].maxScoreNode = pointIndex;
      pointIndex = tmp;
    }
    cur = (pointKey <= tree[cur].medianKey) ? tree[cur].left : tree[cur].right;
  }
}

/* frag[4i..] = xl, yl, xh, yh; score[i] in: the fragment's own score, out: the score of the best chain ending in it; prev[i]: its
 * predecessor or -1; chain_out: the optimal chain (fragment indices, first to last).  Returns the chain length. */
long lra_oracle_global_chain(const int32_t *frag, int32_t *score, int32_t *prev, long n, int32_t *chain_out) {
  if (n == 0) return 0;
  const long m = 2 * n;
  gc_ep *ep = (gc_ep *)calloc((size_t)m, sizeof(gc_ep));
  for (long i = 0; i < n; i++) {
    ep[2 * i].x = frag[4 * i]; ep[2 * i].y = frag[4 * i + 1]; ep[2 * i].side = 0; ep[2 * i].frag = (int32_t)i;
    ep[2 * i + 1].x = frag[4 * i + 2]; ep[2 * i + 1].y = frag[4 * i + 3]; ep[2 * i + 1].side = 1; ep[2 * i + 1].frag = (int32_t)i;
    prev[i] = -1;
  }
  ep_sort(ep, m);
  gc_vx *tree = (gc_vx *)calloc((size_t)(2 * m - 1), sizeof(gc_vx));
  for (long i = 0; i < 2 * m - 1; i++) { tree[i].maxScoreNode = -1; tree[i].pointIndex = -1; }
  uint32_t it = 0;
  pst_create(tree, ep, 0, (int)m, &it);
  long maxEp = 0; int found = 0;
  for (long p = 0; p < m; p++) {
    if (ep[p].side == 0) {
      int maxIndex = 0, ok = 0;
      if (tree[0].maxScoreNode != -1) { int maxValue = -1; ok = pst_find(tree, 0, ep, (uint32_t)ep[p].y, &maxValue, &maxIndex); }
      if (ok) { const int fPrev = ep[maxIndex].frag; prev[ep[p].frag] = fPrev; score[ep[p].frag] = score[fPrev] + score[ep[p].frag]; }
      else prev[ep[p].frag] = -1;
    } else {
      ep[p].score = score[ep[p].frag];
      pst_activate(tree, ep, (int)p);
      if (!found || score[ep[maxEp].frag] < score[ep[p].frag]) { maxEp = p; found = 1; }
    }
  }
  long len = 0;
  if (found) {
    int f = ep[maxEp].frag;
    while (f != -1 && len < n) { chain_out[len++] = f; f = prev[f]; }
    for (long i = 0; i < len / 2; i++) { int32_t t = chain_out[i]; chain_out[i] = chain_out[len - 1 - i]; chain_out[len - 1 - i] = t; }
  }
  free(ep); free(tree);
  return len;
}
