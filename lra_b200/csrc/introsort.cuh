// libstdc++'s std::sort (GCC 13.3, bits/stl_algo.h: __introsort_loop with threshold 16, median of three moved to the front, unguarded
// partition, heapsort at depth 2*floor(log2 n), then one final insertion sort), restated as a sequential device routine for one thread.
// std::sort is unstable, and wherever the reference sorts records whose keys can tie while their payloads differ, the order it leaves
// them in is observable downstream -- so that order is reproduced rather than replaced by a parallel sort.
#pragma once
#include "lra_common.cuh"

namespace lra {

template <class T, class Less>
__device__ __forceinline__ void is_unguarded_linear_insert(T *v, int last, Less less) {
  const T val = v[last];
  int next = last - 1;
  while (less(val, v[next])) { v[last] = v[next]; last = next; --next; }
  v[last] = val;
}
template <class T, class Less>
__device__ __forceinline__ void is_insertion_sort(T *v, int first, int last, Less less) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (less(v[i], v[first])) {
      const T val = v[i];
      for (int j = i; j > first; j--) v[j] = v[j - 1];
      v[first] = val;
    } else is_unguarded_linear_insert(v, i, less);
  }
}
template <class T, class Less>
__device__ __forceinline__ void is_adjust_heap(T *v, int first, int holeIndex, int len, T value, Less less) {
  const int topIndex = holeIndex;
  int secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (less(v[first + secondChild], v[first + secondChild - 1])) secondChild--;
    v[first + holeIndex] = v[first + secondChild];
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    v[first + holeIndex] = v[first + secondChild - 1];
    holeIndex = secondChild - 1;
  }
  int parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && less(v[first + parent], value)) { v[first + holeIndex] = v[first + parent]; holeIndex = parent; parent = (holeIndex - 1) / 2; }
  v[first + holeIndex] = value;
}
template <class T, class Less>
__device__ __forceinline__ void is_heap_sort(T *v, int first, int last, Less less) {
  const int len = last - first;
  if (len >= 2)
    for (int parent = (len - 2) / 2;; parent--) { is_adjust_heap(v, first, parent, len, v[first + parent], less); if (parent == 0) break; }
  while (last - first > 1) {
    --last;
    const T val = v[last];
    v[last] = v[first];
    is_adjust_heap(v, first, 0, last - first, val, less);
  }
}
template <class T, class Less>
__device__ void std_sort_replay(T *v, int n, Less less) {
  if (n <= 1) return;
  int lg = 0;
  { unsigned x = (unsigned)n; while (x > 1) { x >>= 1; lg++; } }
  int stF[64], stL[64], stD[64];      // the recursive call takes the right part first; the left part is continued afterwards
  int sp = 0;
  stF[0] = 0; stL[0] = n; stD[0] = lg * 2; sp = 1;
  while (sp > 0) {
    --sp;
    int first = stF[sp], last = stL[sp], depth = stD[sp];
    while (last - first > 16) {
      if (depth == 0) { is_heap_sort(v, first, last, less); break; }
      --depth;
      const int mid = first + (last - first) / 2;
      const int a = first + 1, b = mid, c = last - 1;
      auto swp = [&](int i, int j) { const T x = v[i]; v[i] = v[j]; v[j] = x; };
      if (less(v[a], v[b])) {
        if (less(v[b], v[c])) swp(first, b); else if (less(v[a], v[c])) swp(first, c); else swp(first, a);
      } else if (less(v[a], v[c])) swp(first, a);
      else if (less(v[b], v[c])) swp(first, c);
      else swp(first, b);
      int lo = first + 1, hi = last;
      for (;;) {
        while (less(v[lo], v[first])) ++lo;
        --hi;
        while (less(v[first], v[hi])) --hi;
        if (!(lo < hi)) break;
        swp(lo, hi);
        ++lo;
      }
      if (sp < 63) { stF[sp] = first; stL[sp] = lo; stD[sp] = depth; sp++; }
      first = lo;
    }
  }
  if (n > 16) { is_insertion_sort(v, 0, 16, less); for (int i = 16; i != n; ++i) is_unguarded_linear_insert(v, i, less); }
  else is_insertion_sort(v, 0, n, less);
}

}  // namespace lra
