// tgq_select.cuh -- the statistics of ShapeDiameter::compute (Algorithms/SurfaceFeatures/Local/ShapeDiameter.hpp:222-257).
//
// The reference finds the median with std::nth_element and then sums weights[i] * values[i] over the PERMUTED values
// but the UNPERMUTED weights, so the result depends on the exact permutation nth_element leaves
// behind.  sdf_select() therefore restates the selection algorithm of the C++ runtime the reference is built with
// (GNU libstdc++ introselect: median-of-3 pivot to front, unguarded Hoare partition, depth limit 2 * floor(log2 n) with
// a heap-select fallback, final insertion sort of <= 3 elements) for plain `double` keys.  tests/test_select.py checks
// it against std::nth_element itself on the host, permutation for permutation.
#pragma once
#include "tgq_math.cuh"

namespace tgq {

TGQ_HD void sel_swap(double * a, int i, int j) { double t = a[i]; a[i] = a[j]; a[j] = t; }

TGQ_HD void sel_adjust_heap(double * a, int hole, int len, double value)
{
  int top = hole, child = hole;
  while (child < (len - 1) / 2)
  {
    child = 2 * (child + 1);
    if (a[child] < a[child - 1]) child--;
    a[hole] = a[child]; hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2)
  {
    child = 2 * (child + 1);
    a[hole] = a[child - 1]; hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && a[parent] < value) { a[hole] = a[parent]; hole = parent; parent = (hole - 1) / 2; }
  a[hole] = value;
}

TGQ_HD void sel_heap_select(double * a, int first, int middle, int last)
{
  double * h = a + first; int len = middle - first;
  if (len >= 2)
  {
    int parent = (len - 2) / 2;
    for (;;) { double v = h[parent]; sel_adjust_heap(h, parent, len, v); if (parent == 0) break; parent--; }
  }
  for (int i = middle; i < last; ++i)
    if (a[i] < a[first]) { double v = a[i]; a[i] = a[first]; sel_adjust_heap(h, 0, len, v); }
}

// Rearranges a[0..n) exactly as std::nth_element(a, a + nth, a + n) of libstdc++ does.
TGQ_HD void sdf_select(double * a, int n, int nth)
{
  if (n <= 0 || nth >= n) return;
  int first = 0, last = n;
  int depth = 0; { int m = n; while (m > 1) { m >>= 1; ++depth; } depth *= 2; }
  while (last - first > 3)
  {
    if (depth == 0)
    {
      sel_heap_select(a, first, nth + 1, last);
      sel_swap(a, first, nth);
      return;
    }
    --depth;
    int mid = first + (last - first) / 2;
    int ia = first + 1, ib = mid, ic = last - 1;            // median of (a, b, c) moved to `first`
    if (a[ia] < a[ib])
    {
      if (a[ib] < a[ic]) sel_swap(a, first, ib);
      else if (a[ia] < a[ic]) sel_swap(a, first, ic);
      else sel_swap(a, first, ia);
    }
    else if (a[ia] < a[ic]) sel_swap(a, first, ia);
    else if (a[ib] < a[ic]) sel_swap(a, first, ic);
    else sel_swap(a, first, ib);
    int lo = first + 1, hi = last;
    for (;;)
    {
      while (a[lo] < a[first]) ++lo;
      --hi;
      while (a[first] < a[hi]) --hi;
      if (!(lo < hi)) break;
      sel_swap(a, lo, hi);
      ++lo;
    }
    if (lo <= nth) first = lo; else last = lo;
  }
  for (int i = first + 1; i < last; ++i)                     // insertion sort of the remaining <= 3 elements
  {
    double v = a[i];
    if (v < a[first]) { for (int j = i; j > first; --j) a[j] = a[j - 1]; a[first] = v; }
    else { int j = i; while (v < a[j - 1]) { a[j] = a[j - 1]; --j; } a[j] = v; }
  }
}

// values: hit times of the rays that count (in ray order), weights: cos of their cone angle, nv <= 30.
// `values` is permuted in place.  Returns the normalised SDF, or -1 when nv == 0.
TGQ_HD double sdf_statistics(double * values, double const * weights, int nv, float scale)
{
  if (nv <= 0) return -1.0;
  int mid = nv / 2;
  sdf_select(values, nv, mid);
  double sv = 0, ss = 0;
  for (int i = 0; i < nv; ++i) { sv += values[i]; ss += values[i] * values[i]; }
  double avg = sv / nv, var = ss / nv - avg * avg;
  sv = 0; double sw = 0;
  for (int i = 0; i < nv; ++i)
  {
    double dv = values[i] - values[mid];
    if (dv * dv <= var) { sv += weights[i] * values[i]; sw += weights[i]; }
  }
  double r = sw > 0 ? sv / (sw * (double)scale) : values[mid] / (double)scale;
  return r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);
}

}  // namespace tgq
