// urf_stdsort.cuh — the exact element order `std::sort` of libstdc++ (GCC 13, bits/stl_algo.h + bits/stl_heap.h) leaves
// when the comparator sees only part of the element.
//
// Why: starShapedSearch sorts the points of a sector with `std::sort(..., ptcmpr)`, `ptcmpr(a, b) = a.r < b.r`
// (star_shaped_search.cpp:22-25,109). std::sort is not stable: the final order of points with EQUAL planar radius is
// whatever introsort's swaps leave, and the edge search that follows depends on it (the slope between two points of equal
// radius is +-inf or NaN). Real sensors quantise ranges, so equal radii inside a sector are not rare (a ring on flat
// ground returns the same range in neighbouring columns). The reference's answer is therefore "what GCC's introsort does
// to the sector's points in push_back order" — libstdc++ is a third-party dependency of the reference (not in its tree,
// version unpinned; the oracle and the fixtures use this container's GCC 13.3), so its published algorithm is restated
// here, operation for operation: introsort loop (median-of-three to the front, unguarded Hoare partition, threshold 16,
// depth limit 2 * floor(log2 n), heapsort when the limit is hit) followed by the final insertion sort.
// tests/kat/stdsort_check.cpp checks the restatement against the real std::sort on tie-heavy and adversarial arrays.
//
// Elements are 64-bit words: high 32 bits = the float bits of r (r >= +0, so unsigned order = float order), low 32 bits =
// the payload (input index) that the comparator does not look at. One thread sorts one sector with this (only sectors
// that hold equal radii take this path).
#pragma once
#include <stdint.h>

#include "urf_math.cuh"

namespace urfsort {

typedef unsigned long long El;
URF_HD bool lt(El a, El b) { return (unsigned)(a >> 32) < (unsigned)(b >> 32); }      // ptcmpr
URF_HD void swp(El* a, El* b) { const El t = *a; *a = *b; *b = t; }

URF_HD void move_median_to_first(El* result, El* a, El* b, El* c) {
  if (lt(*a, *b)) {
    if (lt(*b, *c)) swp(result, b);
    else if (lt(*a, *c)) swp(result, c);
    else swp(result, a);
  } else if (lt(*a, *c)) swp(result, a);
  else if (lt(*b, *c)) swp(result, c);
  else swp(result, b);
}

URF_HD El* unguarded_partition(El* first, El* last, El* pivot) {
  while (true) {
    while (lt(*first, *pivot)) ++first;
    --last;
    while (lt(*pivot, *last)) --last;
    if (!(first < last)) return first;
    swp(first, last);
    ++first;
  }
}

URF_HD void push_heap(El* first, long hole, long top, El value) {
  long parent = (hole - 1) / 2;
  while (hole > top && lt(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

URF_HD void adjust_heap(El* first, long hole, long len, El value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (lt(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  push_heap(first, hole, top, value);
}

URF_HD void heap_sort_all(El* first, El* last) {           // __partial_sort(first, last, last): make_heap + sort_heap
  const long len = last - first;
  if (len >= 2) {
    long parent = (len - 2) / 2;
    while (true) {
      const El value = first[parent];
      adjust_heap(first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {                               // __sort_heap: __pop_heap(first, last, last)
    --last;
    const El value = *last;
    *last = *first;
    adjust_heap(first, 0, last - first, value);
  }
}

URF_HD void unguarded_linear_insert(El* last) {
  const El val = *last;
  El* next = last - 1;
  while (lt(val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}

URF_HD void insertion_sort(El* first, El* last) {
  if (first == last) return;
  for (El* i = first + 1; i != last; ++i) {
    if (lt(*i, *first)) {
      const El val = *i;
      for (El* p = i; p != first; --p) *p = *(p - 1);      // move_backward(first, i, i + 1)
      *first = val;
    } else unguarded_linear_insert(i);
  }
}

// std::sort(first, first + n, ptcmpr). The recursion of __introsort_loop on the right part is an explicit stack (the
// parts are disjoint, so the order they are finished in does not matter); its depth never exceeds the depth limit.
URF_HD void std_sort(El* first, long n) {
  if (n <= 0) return;
  El* last = first + n;
  int lg = 0;
  while ((n >> (lg + 1)) != 0) lg++;                       // std::__lg(n)
  struct Frame { int first, last, depth; };                // offsets: n < 2^31, so at most 2 * 30 + 1 frames are ever pending
  Frame stack[64];
  int sp = 0;
  stack[sp++] = Frame{0, (int)n, 2 * lg};
  while (sp > 0) {
    Frame f = stack[--sp];
    while (f.last - f.first > 16) {                        // _S_threshold
      if (f.depth == 0) {
#ifdef URF_STDSORT_COUNT_HEAP
        URF_STDSORT_COUNT_HEAP++;
#endif
        heap_sort_all(first + f.first, first + f.last);
        break;
      }
      --f.depth;
      El* lo = first + f.first;
      El* hi = first + f.last;
      El* mid = lo + (hi - lo) / 2;
      move_median_to_first(lo, lo + 1, mid, hi - 1);
      const int cut = (int)(unguarded_partition(lo + 1, hi, lo) - first);
      stack[sp++] = Frame{cut, f.last, f.depth};           // __introsort_loop(cut, last, depth_limit)
      f.last = cut;
    }
  }
  if (n > 16) {                                            // __final_insertion_sort
    insertion_sort(first, first + 16);
    for (El* i = first + 16; i != last; ++i) unguarded_linear_insert(i);
  } else insertion_sort(first, last);
}

}  // namespace urfsort
