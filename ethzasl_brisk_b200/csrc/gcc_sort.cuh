// std::sort of libstdc++ (GCC 13, bits/stl_algo.h:1848-1950, bits/stl_heap.h), element for element,
// as host/device templates.  The reference sorts with std::sort in two places where equal keys are
// the norm -- Harris maxima by score (score-calculator.h:82-84) and radius matches by
// distance (brute-force-matcher.cc:210) -- and introsort is not stable, so the exact permutation is
// part of the result.  The recursion on disjoint sub-ranges is replaced by an explicit stack (order
// independent).
#pragma once
#include "brisk_common.cuh"

namespace briskb200 {

template <class T>
BRISK_HD void gs_swap(T& a, T& b) { const T t = a; a = b; b = t; }

template <class T, class Less>
BRISK_HD void gs_push_heap(const Less& less, T* first, long hole, long top, T value) {
  long parent = (hole - 1) / 2;
  while (hole > top && less(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

template <class T, class Less>
BRISK_HD void gs_adjust_heap(const Less& less, T* first, long hole, long len, T value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (less(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  gs_push_heap(less, first, hole, top, value);
}

template <class T, class Less>
BRISK_HD void gs_heapsort(const Less& less, T* first, long len) {  // __partial_sort(first, last, last)
  if (len >= 2) {
    long parent = (len - 2) / 2;
    for (;;) {
      const T v = first[parent];
      gs_adjust_heap(less, first, parent, len, v);
      if (parent == 0) break;
      parent--;
    }
  }
  long last = len;
  while (last > 1) {
    --last;
    const T v = first[last];
    first[last] = first[0];
    gs_adjust_heap(less, first, 0, last, v);
  }
}

template <class T, class Less>
BRISK_HD void gs_unguarded_linear_insert(const Less& less, T* last) {
  const T val = *last;
  T* next = last - 1;
  while (less(val, *next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

template <class T, class Less>
BRISK_HD void gs_insertion_sort(const Less& less, T* first, T* last) {
  if (first == last) return;
  for (T* i = first + 1; i != last; ++i) {
    if (less(*i, *first)) {
      const T val = *i;
      for (T* j = i; j != first; --j) *j = *(j - 1);
      *first = val;
    } else {
      gs_unguarded_linear_insert(less, i);
    }
  }
}

// One pass of std::__introsort_loop's body on [first, last): __unguarded_partition_pivot (median of
// first+1, mid, last-1 moved to first, then the unguarded Hoare scan).  Returns the cut.
// __move_median_to_first(first, first + 1, mid, last - 1) of __unguarded_partition_pivot.
template <class T, class Less>
BRISK_HD void gs_median_to_first(const Less& less, T* a, int first, int last) {
  const int mid = first + (last - first) / 2;
  T& ra = a[first + 1]; T& rb = a[mid]; T& rc = a[last - 1];
  if (less(ra, rb)) {
    if (less(rb, rc)) gs_swap(a[first], rb);
    else if (less(ra, rc)) gs_swap(a[first], rc);
    else gs_swap(a[first], ra);
  } else if (less(ra, rc)) gs_swap(a[first], ra);
  else if (less(rb, rc)) gs_swap(a[first], rc);
  else gs_swap(a[first], rb);
}

template <class T, class Less>
BRISK_HD int gs_partition(const Less& less, T* a, int first, int last) {
  gs_median_to_first(less, a, first, last);
  int lo = first + 1, hi = last;
  for (;;) {
    while (less(a[lo], a[first])) ++lo;
    --hi;
    while (less(a[first], a[hi])) --hi;
    if (!(lo < hi)) break;
    gs_swap(a[lo], a[hi]);
    ++lo;
  }
  return lo;
}

// std::__introsort_loop on [first, last) with the given depth limit, followed by the part of
// __final_insertion_sort that concerns this range.  The final insertion sort never moves an element
// across a partition cut (everything left of a cut compares >= everything right of it), so it is
// the same as a stable insertion sort of every leaf range of at most 16 elements.
template <class T, class Less>
BRISK_HD void gs_sort_range(const Less& less, T* a, int first0, int last0, int depth0) {
  int st_first[64], st_last[64], st_depth[64];
  int sp = 0;
  st_first[0] = first0; st_last[0] = last0; st_depth[0] = depth0; sp = 1;
  while (sp > 0) {
    --sp;
    int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
    bool heap = false;
    while (last - first > 16) {
      if (depth == 0) { gs_heapsort(less, a + first, last - first); heap = true; break; }
      --depth;
      const int cut = gs_partition(less, a, first, last);
      st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; ++sp;   // recurse on [cut, last), go on with [first, cut)
      last = cut;
    }
    if (!heap) gs_insertion_sort(less, a + first, a + last);
  }
}

BRISK_HD int gs_depth_limit(int n) {
  int lg = 0;
  for (unsigned v = (unsigned)n; v > 1; v >>= 1) ++lg;  // std::__lg
  return 2 * lg;
}

// std::sort(a, a + n) with hp_less, element for element (libstdc++ of GCC 13).
template <class T, class Less>
BRISK_HD void gs_sort(const Less& less, T* a, int n) {
  if (n <= 0) return;
  gs_sort_range(less, a, 0, n, gs_depth_limit(n));
}

}  // namespace briskb200
