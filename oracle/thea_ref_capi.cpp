// oracle/thea_ref_capi.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN, UNMODIFIED sources
// (/root/reference/Code/Source, included where they lie; nothing is copied into this repository), compiled against the Eigen
// stand-in of oracle/eigen_shim/ (the container has no Eigen: SURVEY.md finding 9).  The resulting oracle/_ref/libthea_ref.so is
// git-ignored and is used ONLY to validate the restatement in oracle/thea_oracle.hpp (tests/test_ref_pin.py) and, optionally, as
// the `cpu_baseline` of bench.py (kind "reference").  The product never loads it.
//
// What runs here is Thea's code: Algorithms::BvhN (kd-tree build BvhN.hpp:704-794,1365-1525; closestPair :1081-1114,1866-1893,
// 2036-2089; kClosestPairs :1137-1146; rangeQueryIndices :1208-1264,2118-2171; rayStructureIntersection :1291-1316,2240-2291;
// BVH-to-BVH closestPair :1900-2009), Triangle3 (Triangle3.hpp:143-155,294-304,536-552,1310-1524), AxisAlignedBoxN, BallN,
// MetricL2, IntersectionTester, RayIntersectionTester, Math::fastSqrt (Math.hpp:194-223), Random (Random.cpp:48-158).
// PointSet3::findSampleNeighbors and the NeighborFunctor (PointSet3.hpp:586-681) are private members; the keyword is redefined for this
// one translation unit so that the reference's code can be CALLED as is (the source is not modified).
#include <sstream>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <memory>
#include <functional>
#include <thread>
#include <atomic>
#include <mutex>
#include <algorithm>
#include <iostream>
#include <fstream>
#include <list>
#include <queue>
#include <stack>
#include <deque>
#include <array>
#include <tuple>
#include <complex>
#include <random>
#include <chrono>
#include <iomanip>
#include <typeinfo>
#include <typeindex>
#include <bitset>
#include <numeric>
#include <iterator>
#include <locale>
#include <regex>
#include <any>
#include <optional>
#include <variant>
#define private public
#define protected public
#include "Algorithms/PointSet3.hpp"
#include "Algorithms/Icp3.hpp"
#undef private
#undef protected
#include "Algorithms/BvhN.hpp"
#include "Algorithms/SurfaceFeatures/Local/ShapeDiameter.hpp"
#include "Algorithms/SurfaceFeatures/Local/SpinImage.hpp"
#include "Algorithms/SurfaceFeatures/Local/AverageDistance.hpp"
#include "Algorithms/SurfaceFeatures/Local/LocalDistanceHistogram.hpp"
#include "Algorithms/SurfaceFeatures/Local/RandomWalks.hpp"
#include "Algorithms/Histogram.hpp"
#include "Algorithms/ShortestPaths.hpp"
#include "Graphics/GeneralMesh.hpp"
#include "Algorithms/IntersectionTester.hpp"
#include "Algorithms/MetricL2.hpp"
#include "Algorithms/PointTraitsN.hpp"
#include "Algorithms/RayIntersectionTester.hpp"
#include "AxisAlignedBox3.hpp"
#include "Ball3.hpp"
#include "BoundedSortedArray.hpp"
#include "LineSegment3.hpp"
#include "Math.hpp"
#include "Random.hpp"
#include "Triangle3.hpp"

#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

using namespace Thea;
using namespace Thea::Algorithms;

namespace {

typedef BvhN<Triangle3, 3> TriBvh;
typedef BvhN<Vector3, 3> PointBvh;

struct RefHandle
{
  TriBvh * tri = nullptr;
  PointBvh * pts = nullptr;
  ~RefHandle() { delete tri; delete pts; }
};

void parallel_for(int64_t n, int nthreads, std::function<void(int64_t, int64_t)> const & fn)
{
  if (nthreads <= 1 || n < 64) { fn(0, n); return; }
  std::vector<std::thread> th;
  int64_t per = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t)
  {
    int64_t b = t * per, e = std::min<int64_t>(n, b + per);
    if (b >= e) break;
    th.emplace_back([=, &fn]() { fn(b, e); });
  }
  for (auto & x : th) x.join();
}

inline Vector3 v3(float const * p) { return Vector3(p[0], p[1], p[2]); }

template <typename BvhT>
void closest_impl(BvhT const & bvh, float const * q3, int64_t Q, double dist_bound, int nthreads, int64_t * idx, float * d2, float * cp3)
{
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      // BvhN::closestElement (BvhN.hpp:1049-1063) is a thin wrapper around closestPair; the pair also carries the monotone
      // (squared) distance, which closestElement only returns as a square root
      typename BvhT::NeighborPair pair = bvh.template closestPair<MetricL2>(v3(q3 + 3 * i), dist_bound, UniversalCompatibility(), true);
      idx[i] = pair.isValid() ? (int64_t)pair.getTargetIndex() : -1;
      if (d2) d2[i] = pair.isValid() ? (float)pair.getMonotoneApproxDistance() : -1.0f;
      if (cp3 && pair.isValid()) { Vector3 const & c = pair.getTargetPoint(); cp3[3 * i] = c[0]; cp3[3 * i + 1] = c[1]; cp3[3 * i + 2] = c[2]; }
      else if (cp3) { cp3[3 * i] = cp3[3 * i + 1] = cp3[3 * i + 2] = 0.0f; }
    }
  });
}

template <typename BvhT>
void knn_impl(BvhT const & bvh, float const * q3, int64_t Q, int k, double dist_bound, int nthreads, int64_t * idx, float * d2)
{
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      BoundedSortedArray<typename BvhT::NeighborPair> nbrs((size_t)k);
      bvh.template kClosestPairs<MetricL2>(v3(q3 + 3 * i), nbrs, dist_bound, UniversalCompatibility(), false);
      for (int j = 0; j < k; ++j)
      {
        bool v = (size_t)j < nbrs.size();
        idx[i * k + j] = v ? (int64_t)nbrs[(size_t)j].getTargetIndex() : -1;
        d2[i * k + j] = v ? (float)nbrs[(size_t)j].getMonotoneApproxDistance() : -1.0f;
      }
    }
  });
}

}  // namespace

extern "C" {

void * ref_create_tris(float const * v9, int64_t T)
{
  std::vector<Triangle3> tris; tris.reserve((size_t)T);
  for (int64_t i = 0; i < T; ++i) tris.push_back(Triangle3(v3(v9 + 9 * i), v3(v9 + 9 * i + 3), v3(v9 + 9 * i + 6)));
  RefHandle * h = new RefHandle;
  h->tri = new TriBvh(tris.begin(), tris.end());          // default arguments of BvhN::init: kd-tree, leaf 5 (BvhN.hpp:704-712)
  return h;
}

void * ref_create_points(float const * p3, int64_t P)
{
  std::vector<Vector3> pts; pts.reserve((size_t)P);
  for (int64_t i = 0; i < P; ++i) pts.push_back(v3(p3 + 3 * i));
  RefHandle * h = new RefHandle;
  h->pts = new PointBvh(pts.begin(), pts.end());
  return h;
}

void ref_destroy(void * h) { delete (RefHandle *)h; }

int64_t ref_num_nodes(void * hv) { RefHandle * h = (RefHandle *)hv; return h->tri ? (int64_t)h->tri->numNodes() : (int64_t)h->pts->numNodes(); }

void ref_closest(void * hv, float const * q3, int64_t Q, double dist_bound, int nthreads, int64_t * idx, float * d2, float * cp3)
{
  RefHandle * h = (RefHandle *)hv;
  if (h->tri) closest_impl(*h->tri, q3, Q, dist_bound, nthreads, idx, d2, cp3);
  else closest_impl(*h->pts, q3, Q, dist_bound, nthreads, idx, d2, cp3);
}

// closestElement itself, with its sqrt'ed distance (BvhN.hpp:1049-1063)
void ref_closest_element(void * hv, float const * q3, int64_t Q, double dist_bound, int64_t * idx, double * dist, float * cp3)
{
  RefHandle * h = (RefHandle *)hv;
  for (int64_t i = 0; i < Q; ++i)
  {
    Vector3 cp = Vector3::Zero(); double d = -1;
    idx[i] = h->tri ? (int64_t)h->tri->closestElement<MetricL2>(v3(q3 + 3 * i), dist_bound, UniversalCompatibility(), &d, &cp)
                    : (int64_t)h->pts->closestElement<MetricL2>(v3(q3 + 3 * i), dist_bound, UniversalCompatibility(), &d, &cp);
    if (dist) dist[i] = d;
    if (cp3) { cp3[3 * i] = cp[0]; cp3[3 * i + 1] = cp[1]; cp3[3 * i + 2] = cp[2]; }
  }
}

void ref_knn(void * hv, float const * q3, int64_t Q, int k, double dist_bound, int nthreads, int64_t * idx, float * d2)
{
  RefHandle * h = (RefHandle *)hv;
  if (h->tri) knn_impl(*h->tri, q3, Q, k, dist_bound, nthreads, idx, d2);
  else knn_impl(*h->pts, q3, Q, k, dist_bound, nthreads, idx, d2);
}

// rayStructureIntersection (BvhN.hpp:1291-1316): time, element index and the hit triangle's cached unit normal
void ref_ray_first(void * hv, float const * o3, float const * d3, float const * tmax, int64_t R, int nthreads, int64_t * idx, float * t,
                   float * n3)
{
  RefHandle * h = (RefHandle *)hv;
  parallel_for(R, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      Ray3 ray(v3(o3 + 3 * i), v3(d3 + 3 * i));
      RayStructureIntersection3 isec = h->tri->rayStructureIntersection<RayIntersectionTester>(ray, tmax ? tmax[i] : (Real)-1);
      bool hit = isec.isValid();
      idx[i] = hit ? (int64_t)isec.getElementIndex() : -1;
      t[i] = hit ? (float)isec.getTime() : -1.0f;
      if (n3) { Vector3 n = hit ? isec.getNormal() : Vector3::Zero(); n3[3 * i] = n[0]; n3[3 * i + 1] = n[1]; n3[3 * i + 2] = n[2]; }
    }
  });
}

// rayIntersectionTime / rayIntersects (BvhN.hpp:1266-1289)
void ref_ray_time(void * hv, float const * o3, float const * d3, float const * tmax, int64_t R, float * t, uint8_t * hit)
{
  RefHandle * h = (RefHandle *)hv;
  for (int64_t i = 0; i < R; ++i)
  {
    Ray3 ray(v3(o3 + 3 * i), v3(d3 + 3 * i));
    if (t) t[i] = (float)h->tri->rayIntersectionTime<RayIntersectionTester>(ray, tmax ? tmax[i] : (Real)-1);
    if (hit) hit[i] = h->tri->rayIntersects<RayIntersectionTester>(ray, tmax ? tmax[i] : (Real)-1) ? 1 : 0;
  }
}

// rangeQueryIndices<IntersectionTester>(Ball3) on a point structure (BvhN.hpp:1208-1220); returns the number found, writes up to
// `cap` indices in the reference's traversal order
int64_t ref_ball_range(void * hv, float const * c3, float radius, int64_t cap, int64_t * idx)
{
  RefHandle * h = (RefHandle *)hv;
  Array<intx> out;
  h->pts->rangeQueryIndices<IntersectionTester>(Ball3(v3(c3), radius), out);
  for (size_t j = 0; j < out.size() && (int64_t)j < cap; ++j) idx[j] = (int64_t)out[j];
  return (int64_t)out.size();
}

int64_t ref_box_range(void * hv, float const * lo3, float const * hi3, int64_t cap, int64_t * idx)
{
  RefHandle * h = (RefHandle *)hv;
  Array<intx> out;
  AxisAlignedBox3 box(v3(lo3), v3(hi3));
  if (h->pts) h->pts->rangeQueryIndices<IntersectionTester>(box, out);
  else h->tri->rangeQueryIndices<IntersectionTester>(box, out);
  for (size_t j = 0; j < out.size() && (int64_t)j < cap; ++j) idx[j] = (int64_t)out[j];
  return (int64_t)out.size();
}

// BVH-to-BVH kClosestPairs between two point structures (BvhN.hpp:1137-1146 with a structure as the query, :1900-2009): returns the
// number of pairs found; rows ascending as the reference's bounded set orders them.
int ref_k_closest_pairs_bvh(void * hv, void * qv, int k, double dist_bound, int64_t * target_idx, int64_t * query_idx, float * d2)
{
  RefHandle * h = (RefHandle *)hv; RefHandle * q = (RefHandle *)qv;
  if (!h->pts || !q->pts) return -1;
  BoundedSortedArray<PointBvh::NeighborPair> set((size_t)k);
  h->pts->kClosestPairs<MetricL2>(*q->pts, set, dist_bound, UniversalCompatibility(), false);
  for (size_t j = 0; j < set.size(); ++j)
  { target_idx[j] = (int64_t)set[j].getTargetIndex(); query_idx[j] = (int64_t)set[j].getQueryIndex(); d2[j] = (float)set[j].getMonotoneApproxDistance(); }
  return (int)set.size();
}

// BVH-to-BVH closestPair (BvhN.hpp:1081-1114 with a structure as the query, :1900-2009).  Returns 1 when a pair was found.
int ref_closest_pair_bvh(void * hv, void * qv, double dist_bound, int64_t * target_idx, int64_t * query_idx, float * d2, float * cp_query3,
                         float * cp_target3)
{
  RefHandle * h = (RefHandle *)hv; RefHandle * q = (RefHandle *)qv;
  bool ok = false; intx ti = -1, qi = -1; double m = -1; Vector3 a = Vector3::Zero(), b = Vector3::Zero();
#define TGQ_REF_PAIR(HS, QS)                                                                                          \
  { auto pair = (HS)->closestPair<MetricL2>(*(QS), dist_bound, UniversalCompatibility(), true);                         \
    ok = pair.isValid(); if (ok) { ti = pair.getTargetIndex(); qi = pair.getQueryIndex(); m = pair.getMonotoneApproxDistance(); \
                                   a = pair.getQueryPoint(); b = pair.getTargetPoint(); } }
  if (h->tri && q->tri) TGQ_REF_PAIR(h->tri, q->tri)
  else if (h->tri && q->pts) TGQ_REF_PAIR(h->tri, q->pts)
  else if (h->pts && q->pts) TGQ_REF_PAIR(h->pts, q->pts)
  else TGQ_REF_PAIR(h->pts, q->tri)
#undef TGQ_REF_PAIR
  *target_idx = ok ? (int64_t)ti : -1; *query_idx = ok ? (int64_t)qi : -1; *d2 = ok ? (float)m : -1.0f;
  for (int c = 0; c < 3; ++c) { cp_query3[c] = a[c]; cp_target3[c] = b[c]; }
  return ok ? 1 : 0;
}

// closest element to a line segment / triangle / ball centre query object (MetricL2 specialisations of the reference)
int64_t ref_closest_segment(void * hv, float const * a3, float const * b3, double dist_bound, float * d2, float * cp_query3, float * cp_target3)
{
  RefHandle * h = (RefHandle *)hv;
  LineSegment3 seg(v3(a3), v3(b3));
  bool ok; intx ti = -1; double m = -1; Vector3 a = Vector3::Zero(), b = Vector3::Zero();
  if (h->tri) { auto pair = h->tri->closestPair<MetricL2>(seg, dist_bound, UniversalCompatibility(), true); ok = pair.isValid();
                if (ok) { ti = pair.getTargetIndex(); m = pair.getMonotoneApproxDistance(); a = pair.getQueryPoint(); b = pair.getTargetPoint(); } }
  else        { auto pair = h->pts->closestPair<MetricL2>(seg, dist_bound, UniversalCompatibility(), true); ok = pair.isValid();
                if (ok) { ti = pair.getTargetIndex(); m = pair.getMonotoneApproxDistance(); a = pair.getQueryPoint(); b = pair.getTargetPoint(); } }
  *d2 = ok ? (float)m : -1.0f;
  for (int c = 0; c < 3; ++c) { cp_query3[c] = a[c]; cp_target3[c] = b[c]; }
  return ok ? (int64_t)ti : -1;
}

// The same query as a batch, under a pushed filter that rejects up to n_excl elements per query by INDEX -- what MeshFix's VertexFilter
// (Tools/MeshFix/MeshFix.cpp:678-697) does by vertex identity.  BvhN hands a filter the element itself (elementPassesFilters,
// BvhN.hpp:1674-1683, called with elems[index], :2044-2047), so the index is recovered from the element's address in getElements().
// pushFilter mutates the structure (:1005-1032): one thread, as the reference's callers do.
extern "C++" {
namespace {
template <typename ElemT>
struct IndexFilter : public Filter<ElemT>
{
  ElemT const * base; int64_t const * excl; int n;
  bool allows(ElemT const & e) const
  {
    int64_t i = (int64_t)(&e - base);
    for (int j = 0; j < n; ++j) if (excl[j] == i) return false;
    return true;
  }
};
template <typename BvhT, typename ElemT>
void closest_segment_impl(BvhT & bvh, float const * a3, float const * b3, int64_t Q, double dist_bound, int64_t const * excl, int n_excl,
                          int64_t * idx, float * d2, float * ep3, float * sp3)
{
  for (int64_t i = 0; i < Q; ++i)
  {
    LineSegment3 seg(v3(a3 + 3 * i), v3(b3 + 3 * i));
    IndexFilter<ElemT> f; f.base = bvh.numElements() > 0 ? &bvh.getElements()[0] : nullptr; f.excl = excl ? excl + i * n_excl : nullptr; f.n = excl ? n_excl : 0;
    if (f.n > 0) bvh.pushFilter(&f);
    auto pair = bvh.template closestPair<MetricL2>(seg, dist_bound, UniversalCompatibility(), true);
    if (f.n > 0) bvh.popFilter();
    bool ok = pair.isValid();
    idx[i] = ok ? (int64_t)pair.getTargetIndex() : -1;
    if (d2) d2[i] = ok ? (float)pair.getMonotoneApproxDistance() : -1.0f;
    Vector3 q = ok ? pair.getQueryPoint() : Vector3::Zero(), t = ok ? pair.getTargetPoint() : Vector3::Zero();
    for (int c = 0; c < 3; ++c) { if (ep3) ep3[3 * i + c] = t[c]; if (sp3) sp3[3 * i + c] = q[c]; }
  }
}
}  // namespace
}  // extern "C++"
void ref_closest_segment_batch(void * hv, float const * a3, float const * b3, int64_t Q, double dist_bound, int64_t const * excl, int n_excl,
                               int64_t * idx, float * d2, float * ep3, float * sp3)
{
  RefHandle * h = (RefHandle *)hv;
  if (h->tri) closest_segment_impl<TriBvh, Triangle3>(*h->tri, a3, b3, Q, dist_bound, excl, n_excl, idx, d2, ep3, sp3);
  else closest_segment_impl<PointBvh, Vector3>(*h->pts, a3, b3, Q, dist_bound, excl, n_excl, idx, d2, ep3, sp3);
}
// Per-pair metrics: MetricL2::closestPoints<3, Real>(element, LineSegment3) exactly as processNeighborsLeaf calls it (BvhN.hpp:2063)
void ref_seg_point_batch(float const * a3, float const * b3, float const * p3, int64_t n, float * d2, float * sp3)
{
  for (int64_t i = 0; i < n; ++i)
  {
    LineSegment3 seg(v3(a3 + 3 * i), v3(b3 + 3 * i)); Vector3 p = v3(p3 + 3 * i), tp, qp;
    d2[i] = MetricL2::closestPoints<3, Real>(p, seg, tp, qp);
    for (int c = 0; c < 3; ++c) sp3[3 * i + c] = qp[c];
  }
}
void ref_tri_seg_batch(float const * v9, float const * a3, float const * b3, int64_t n, float * d2, float * tp3, float * sp3)
{
  for (int64_t i = 0; i < n; ++i)
  {
    Triangle3 tri(v3(v9 + 9 * i), v3(v9 + 9 * i + 3), v3(v9 + 9 * i + 6));
    LineSegment3 seg(v3(a3 + 3 * i), v3(b3 + 3 * i)); Vector3 tp = Vector3::Zero(), qp = Vector3::Zero();
    d2[i] = MetricL2::closestPoints<3, Real>(tri, seg, tp, qp);
    for (int c = 0; c < 3; ++c) { tp3[3 * i + c] = tp[c]; sp3[3 * i + c] = qp[c]; }
  }
}

// MetricL2::closestPoints<3, Real>(Triangle3 a, Triangle3 b) -> a.squaredDistance(b, &cpa, &cpb) (Triangle3.hpp:314-392), per pair
void ref_tri_tri_batch(float const * va9, float const * vb9, int64_t n, float * d2, float * pa3, float * pb3)
{
  for (int64_t i = 0; i < n; ++i)
  {
    Triangle3 a(v3(va9 + 9 * i), v3(va9 + 9 * i + 3), v3(va9 + 9 * i + 6)), b(v3(vb9 + 9 * i), v3(vb9 + 9 * i + 3), v3(vb9 + 9 * i + 6));
    Vector3 pa = Vector3::Zero(), pb = Vector3::Zero();
    d2[i] = MetricL2::closestPoints<3, Real>(a, b, pa, pb);
    for (int c = 0; c < 3; ++c) { pa3[3 * i + c] = pa[c]; pb3[3 * i + c] = pb[c]; }
  }
}
// BvhN::closestPair<MetricL2>(Triangle3 query object, dist_bound, compat, true), one query triangle at a time
void ref_closest_triangle_batch(void * hv, float const * q9, int64_t Q, double dist_bound, int64_t * idx, float * d2, float * ep3, float * qp3)
{
  RefHandle * h = (RefHandle *)hv;
  for (int64_t i = 0; i < Q; ++i)
  {
    Triangle3 q(v3(q9 + 9 * i), v3(q9 + 9 * i + 3), v3(q9 + 9 * i + 6));
    bool ok; intx ti = -1; double m = -1; Vector3 a = Vector3::Zero(), b = Vector3::Zero();
    if (h->tri) { auto pair = h->tri->closestPair<MetricL2>(q, dist_bound, UniversalCompatibility(), true); ok = pair.isValid();
                  if (ok) { ti = pair.getTargetIndex(); m = pair.getMonotoneApproxDistance(); a = pair.getQueryPoint(); b = pair.getTargetPoint(); } }
    else        { auto pair = h->pts->closestPair<MetricL2>(q, dist_bound, UniversalCompatibility(), true); ok = pair.isValid();
                  if (ok) { ti = pair.getTargetIndex(); m = pair.getMonotoneApproxDistance(); a = pair.getQueryPoint(); b = pair.getTargetPoint(); } }
    idx[i] = ok ? (int64_t)ti : -1; if (d2) d2[i] = ok ? (float)m : -1.0f;
    for (int c = 0; c < 3; ++c) { if (ep3) ep3[3 * i + c] = b[c]; if (qp3) qp3[3 * i + c] = a[c]; }
  }
}

// ---- free functions ----------------------------------------------------------------------------------------------------------
// Triangle cache (Triangle3.hpp:143-155): unit normal, plane offset, primary axis, area, centroid
void ref_tri_cache(float const * v9, float * n3, float * d, int32_t * axis, float * area, float * centroid3)
{
  Triangle3 tri(v3(v9), v3(v9 + 3), v3(v9 + 6));
  Vector3 n = tri.getNormal(), c = tri.getCentroid();
  for (int k = 0; k < 3; ++k) { n3[k] = n[k]; centroid3[k] = c[k]; }
  *d = tri.getPlane().signedDistance(Vector3::Zero()) * -1.0f;      // not used for parity (signed zero); see ref_tri_point
  *axis = (int32_t)tri.getPrimaryAxis();
  *area = tri.getArea();
}

// MetricL2::closestPoints(triangle, point) (MetricL2.hpp:196-216): squared distance + closest point
float ref_tri_point(float const * v9, float const * p3, float * cp3)
{
  Triangle3 tri(v3(v9), v3(v9 + 3), v3(v9 + 6));
  Vector3 p = v3(p3), cpa, cpb;
  float m = (float)MetricL2::closestPoints<3, Real>(tri, p, cpa, cpb);
  cp3[0] = cpa[0]; cp3[1] = cpa[1]; cp3[2] = cpa[2];
  return m;
}

void ref_tri_point_batch(float const * v9, float const * p3, int64_t n, float * d2, float * cp3)
{ for (int64_t i = 0; i < n; ++i) d2[i] = ref_tri_point(v9 + 9 * i, p3 + 3 * i, cp3 + 3 * i); }

// Triangle3::rayIntersectionTime (Triangle3.hpp:536-540 over :1464-1524)
float ref_ray_tri(float const * v9, float const * o3, float const * d3, float max_time)
{
  Triangle3 tri(v3(v9), v3(v9 + 3), v3(v9 + 6));
  return (float)tri.rayIntersectionTime(Ray3(v3(o3), v3(d3)), max_time);
}

void ref_ray_tri_batch(float const * v9, float const * o3, float const * d3, int64_t n, float * t)
{ for (int64_t i = 0; i < n; ++i) t[i] = ref_ray_tri(v9 + 9 * i, o3 + 3 * i, d3 + 3 * i, -1.0f); }

float ref_fast_sqrt(float x) { return Math::fastSqrt(x); }

// AxisAlignedBoxN::squaredDistance(point) and rayIntersectionTime (AxisAlignedBoxN.hpp:429-462,576-658)
float ref_box_point_sqdist(float const * lo3, float const * hi3, float const * p3) { return AxisAlignedBox3(v3(lo3), v3(hi3)).squaredDistance(v3(p3)); }
float ref_box_ray_time(float const * lo3, float const * hi3, float const * o3, float const * d3, float max_time)
{ return AxisAlignedBox3(v3(lo3), v3(hi3)).rayIntersectionTime(Ray3(v3(o3), v3(d3)), max_time); }

// Thea::Random(seed, threadsafe = false).uniform01() (Random.hpp:126-130, Random.cpp:48-125)
void ref_random_uniform01(uint32_t seed, int64_t n, float * out)
{
  Random rng(seed, false);
  for (int64_t i = 0; i < n; ++i) out[i] = rng.uniform01();
}

// Icp3<Real>::align(n, from, to_structure, &error) (Icp3.hpp:110-126,193-275) against the triangle structure of the handle.
// tr12 = row-major [L | t].  The 3 x 3 SVD inside alignOneStep comes from the Eigen stand-in (not Eigen's own JacobiSVD).
void ref_icp(void * hv, float const * from3, int64_t n, int min_iters, int max_iters, double frac_thresh, float * tr12, float * error)
{
  RefHandle * h = (RefHandle *)hv;
  std::vector<Vector3> from((size_t)n);
  for (int64_t i = 0; i < n; ++i) from[(size_t)i] = v3(from3 + 3 * i);
  Icp3<Real> icp(frac_thresh, min_iters, max_iters, false);
  Real err = 0;
  AffineTransform3 tr = icp.align((intx)n, from.data(), *h->tri, &err);
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tr12[4 * r + c] = tr.getLinear()(r, c); tr12[4 * r + 3] = tr.getTranslation()[r]; }
  if (error) *error = err;
}

// Icp3<Real>::measureError(tr, n, from, nullptr, to) (Icp3.hpp:418-464) and alignOneStep(n, from, nullptr, nullptr, to, nullptr)
// (:298-384) with given correspondences (private members, called as they are).
float ref_icp_error(float const * tr12, int64_t n, float const * from3, float const * to3)
{
  Matrix3 L; Vector3 t;
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) L(r, c) = tr12[4 * r + c]; t[r] = tr12[4 * r + 3]; }
  std::vector<Vector3> f((size_t)n), q((size_t)n);
  for (int64_t i = 0; i < n; ++i) { f[(size_t)i] = v3(from3 + 3 * i); q[(size_t)i] = v3(to3 + 3 * i); }
  return Icp3<Real>::measureError(AffineTransform3(L, t), (intx)n, f.data(), (Icp3<Real>::DefaultWeightFunc<Vector3> *)nullptr, q.data());
}
void ref_icp_step(int64_t n, float const * from3, float const * to3, float * tr12)
{
  std::vector<Vector3> f((size_t)n), q((size_t)n);
  for (int64_t i = 0; i < n; ++i) { f[(size_t)i] = v3(from3 + 3 * i); q[(size_t)i] = v3(to3 + 3 * i); }
  Icp3<Real> icp(1e-4, 3, 10, false);
  AffineTransform3 tr = icp.alignOneStep((intx)n, f.data(), (Icp3<Real>::DefaultWeightFunc<Vector3> *)nullptr, (Plane3 const *)nullptr, q.data(),
                                         (Plane3 const *)nullptr);
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tr12[4 * r + c] = tr.getLinear()(r, c); tr12[4 * r + 3] = tr.getTranslation()[r]; }
}

// ShapeDiameter<Mesh, ExternalBvh>::compute(position, normal, only_hit_interior_surfaces) (ShapeDiameter.hpp:168-258) on the handle's
// triangle structure as the external BVH; compute(position) (:141-158) when n3 is null.
void ref_sdf(void * hv, float const * p3, float const * n3, int64_t S, float scale, int only_interior, int nthreads, double * sdf)
{
  RefHandle * h = (RefHandle *)hv;
  typedef SurfaceFeatures::Local::ShapeDiameter<Graphics::GeneralMesh<>, TriBvh> Sdf;
  Sdf f(h->tri, scale);
  parallel_for(S, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
      sdf[i] = n3 ? f.compute(v3(p3 + 3 * i), v3(n3 + 3 * i), only_interior != 0) : f.compute(v3(p3 + 3 * i), only_interior != 0);
  });
}

// PointSet3::findSampleNeighbors (PointSet3.hpp:660-681) for every sample, with the given initial radius (the reference derives it
// from 100 nearest-neighbour distances of samples drawn from the time-seeded Random::common(), PointSet3.hpp:363-376, which cannot
// be replayed; the callers pass sqrt(max_degree) * avg_sep of fixed seed indices instead, computed by ref_avg_separation below).
// surface: a handle with a triangle structure for the reachability test, or null.  idx / sep: P x max_degree, -1 padded.
void ref_graph(float const * p3, float const * n3, int64_t P, int max_degree, float init_radius, void * surface, int64_t * idx, float * sep, int32_t * count)
{
  PointSet3::Options opts = PointSet3::Options::defaults(); opts.setMaxDegree(max_degree);
  PointSet3 ps(opts);
  std::vector<Vector3> pos((size_t)P), nrm((size_t)P);
  for (int64_t i = 0; i < P; ++i) { pos[(size_t)i] = v3(p3 + 3 * i); if (n3) nrm[(size_t)i] = v3(n3 + 3 * i); }
  ps.addSamples((intx)P, pos.data(), n3 ? nrm.data() : nullptr);
  PointSet3::SampleBvh & bvh = const_cast<PointSet3::SampleBvh &>(ps.getBvh());
  auto elems = bvh.getElements();
  RefHandle * sh = (RefHandle *)surface;
  for (intx i = 0; i < bvh.numElements(); ++i)
  {
    if (sh) ps.findSampleNeighbors<TriBvh>(elems[i], bvh, init_radius, sh->tri);
    else ps.findSampleNeighbors<PointSet3::DummyRayQueryStructure3>(elems[i], bvh, init_radius, nullptr);
  }
  for (int64_t i = 0; i < P; ++i)
  {
    PointSet3::Sample const & s = ps.getSample((intx)i);
    if (count) count[i] = (int32_t)s.numNeighbors();
    for (int j = 0; j < max_degree; ++j)
    {
      bool v = j < s.numNeighbors();
      idx[i * max_degree + j] = v ? (int64_t)s.getNeighbor(j).getSample()->getIndex() : -1;
      sep[i * max_degree + j] = v ? (float)s.getNeighbor(j).getSeparation() : -1.0f;
    }
  }
}

// The average-separation estimate of PointSet3::updateGraph (PointSet3.hpp:363-376) for GIVEN seed indices: RMS of the distance
// from each seed sample to its nearest other sample (closestElement under a FilterSelf), accumulated in Real in seed order.
float ref_avg_separation(float const * p3, int64_t P, int64_t const * seeds, int n)
{
  PointSet3 ps;
  std::vector<Vector3> pos((size_t)P);
  for (int64_t i = 0; i < P; ++i) pos[(size_t)i] = v3(p3 + 3 * i);
  ps.addSamples((intx)P, pos.data(), nullptr);
  PointSet3::SampleBvh & bvh = const_cast<PointSet3::SampleBvh &>(ps.getBvh());
  Real avg = 0;
  for (int j = 0; j < n; ++j)
  {
    PointSet3::Sample * self = const_cast<PointSet3::Sample *>(&ps.getSample((intx)seeds[j]));
    PointSet3::FilterSelf filter(self);
    bvh.pushFilter(&filter);
      intx nn = bvh.closestElement<MetricL2>(self->getPosition());
      avg += (bvh.getElements()[nn]->getPosition() - self->getPosition()).squaredNorm();
    bvh.popFilter();
  }
  return std::sqrt(avg / n);
}

// ---- sample-graph consumers.  The reference derives the graph radius from the time-seeded Random::common()
// (PointSet3.hpp:363-376), so a graph cannot be replayed; instead a GIVEN graph (idx / sep: P x max_degree, -1 padded) is injected into
// a PointSet3 -- neighbour sets filled through SamplePoint3::addNeighbor, valid_graph set -- and the reference's own code runs on it.
namespace {
void fill_samples(PointSet3 & ps, float const * p3, float const * n3, int64_t n_orig, int64_t n_dense)
{
  std::vector<Vector3> pos((size_t)(n_orig + n_dense)), nrm((size_t)(n_orig + n_dense));
  for (int64_t i = 0; i < n_orig + n_dense; ++i) { pos[(size_t)i] = v3(p3 + 3 * i); if (n3) nrm[(size_t)i] = v3(n3 + 3 * i); }
  ps.addSamples((intx)n_orig, pos.data(), n3 ? nrm.data() : nullptr);
  if (n_dense > 0) ps.addOversampling((intx)n_dense, pos.data() + n_orig, n3 ? nrm.data() + n_orig : nullptr);
}
PointSet3::Sample * sample_at(PointSet3 & ps, int64_t i)
{
  int64_t n = (int64_t)ps.samples.size();
  return i < n ? &ps.samples[(size_t)i] : &ps.dense_samples[(size_t)(i - n)];
}
void inject_graph(PointSet3 & ps, int64_t P, int max_degree, int64_t const * idx, float const * sep)
{
  for (int64_t i = 0; i < P; ++i)
  {
    PointSet3::Sample * s = sample_at(ps, i); s->setIndex((intx)i); s->clearNeighbors();
    for (int j = 0; j < max_degree; ++j)
      if (idx[i * max_degree + j] >= 0) s->addNeighbor(PointSet3::Sample::Neighbor(sample_at(ps, idx[i * max_degree + j]), sep[i * max_degree + j]));
  }
  ps.valid_graph = true;
}
}  // namespace

// PointSet3::extractOriginalAdjacencies (PointSet3.cpp:158-178) on a given graph over n_orig original + n_dense dense samples.
void ref_graph_contract(float const * p3, int64_t n_orig, int64_t n_dense, int max_degree, int64_t const * idx, float const * sep,
                        int64_t * out_idx, float * out_sep, int32_t * out_count)
{
  PointSet3::Options opts = PointSet3::Options::defaults(); opts.setMaxDegree(max_degree);
  PointSet3 ps(opts);
  fill_samples(ps, p3, nullptr, n_orig, n_dense);
  int64_t P = n_orig + n_dense;
  inject_graph(ps, P, max_degree, idx, sep);
  std::vector<PointSet3::Sample *> ptrs((size_t)P);
  for (int64_t i = 0; i < P; ++i) ptrs[(size_t)i] = sample_at(ps, i);
  ps.extractOriginalAdjacencies((intx)P, ptrs.data());
  for (int64_t i = 0; i < n_orig; ++i)
  {
    PointSet3::Sample const & s = ps.samples[(size_t)i];
    if (out_count) out_count[i] = (int32_t)s.numNeighbors();
    for (int j = 0; j < max_degree; ++j)
    {
      bool v = j < s.numNeighbors();
      out_idx[i * max_degree + j] = v ? (int64_t)s.getNeighbor(j).getSample()->getIndex() : -1;
      out_sep[i * max_degree + j] = v ? (float)s.getNeighbor(j).getSeparation() : -1.0f;
    }
  }
}

// AverageDistance::compute(position, GEODESIC, max_distance) and LocalDistanceHistogram::compute(position, histogram, GEODESIC,
// max_distance, ratio >= 1) (SurfaceFeatures/Local/AverageDistance.cpp:84-118, LocalDistanceHistogram.cpp:90-137) on a given graph.
void ref_geodesic_features(float const * p3, int64_t P, int max_degree, int64_t const * idx, float const * sep, float scale, float const * q3, int64_t Q,
                           float max_distance, int nbins, double * avg, double * hist)
{
  PointSet3::Options opts = PointSet3::Options::defaults(); opts.setMaxDegree(max_degree);
  PointSet3 ps(opts);
  fill_samples(ps, p3, nullptr, P, 0);
  inject_graph(ps, P, max_degree, idx, sep);
  ps.setScale(scale);
  SurfaceFeatures::Local::AverageDistance ad(&ps);
  SurfaceFeatures::Local::LocalDistanceHistogram ldh(&ps);
  for (int64_t i = 0; i < Q; ++i)
  {
    if (avg) avg[i] = ad.compute(v3(q3 + 3 * i), DistanceType::GEODESIC, max_distance);
    if (hist)
    {
      Histogram h(nbins, hist + (int64_t)nbins * i);
      ldh.compute(v3(q3 + 3 * i), h, DistanceType::GEODESIC, max_distance, 2.0f);
    }
  }
}

// FurthestPointSampling::subsample (FurthestPointSampling.cpp:48-125) on a given graph: the selection loop of :72-100 -- first sample 0,
// then repeatedly the LAST vertex the reference's own ShortestPaths::dijkstraWithCallback reports from the selected set
// (include_unreachable = true) -- around the reference's Dijkstra.
int64_t ref_fps(float const * p3, int64_t P, int max_degree, int64_t const * idx, float const * sep, int64_t n_desired, int64_t * selected)
{
  PointSet3::Options opts = PointSet3::Options::defaults(); opts.setMaxDegree(max_degree);
  PointSet3 surf(opts);
  fill_samples(surf, p3, nullptr, P, 0);
  inject_graph(surf, P, max_degree, idx, sep);
  typedef PointSet3::SampleGraph::VertexHandle VH;
  struct Callback { VH furthest = nullptr; bool operator()(VH vertex, double, bool, VH) { furthest = vertex; return false; } };
  ShortestPaths<PointSet3::SampleGraph> shortest_paths;
  UnorderedMap<VH, double> src_region;
  for (int64_t i = 0; i < n_desired; ++i)
  {
    VH furthest_sample = nullptr;
    if (src_region.empty()) furthest_sample = const_cast<VH>(&surf.getSample(0));
    else
    {
      Callback callback;
      shortest_paths.dijkstraWithCallback(const_cast<PointSet3::SampleGraph &>(surf.getGraph()), nullptr, std::ref(callback), -1, &src_region, true);
      if (!callback.furthest || src_region.find(callback.furthest) != src_region.end()) return i;
      furthest_sample = callback.furthest;
    }
    selected[i] = (int64_t)furthest_sample->getIndex();
    src_region[furthest_sample] = 0;
  }
  return n_desired;
}

// SpinImage::compute(position, normal, num_radial_bins, num_height_bins, spin_image) (SurfaceFeatures/Local/SpinImage.cpp:47-75);
// PointSet3::setScale fixes the scale the reference would derive from the samples' bounding sphere.  image: [Q x nr x nh].
void ref_spin_image(float const * p3, float const * n3, int64_t P, float scale, float const * q3, float const * qn3, int64_t Q, int nr, int nh, double * image)
{
  PointSet3 ps;
  fill_samples(ps, p3, n3, P, 0);
  ps.setScale(scale);
  SurfaceFeatures::Local::SpinImage si(&ps);
  MatrixX<double> m;
  for (int64_t i = 0; i < Q; ++i)
  {
    si.compute(v3(q3 + 3 * i), v3(qn3 + 3 * i), nr, nh, m);
    for (int r = 0; r < nr; ++r) for (int c = 0; c < nh; ++c) image[((int64_t)i * nr + r) * nh + c] = m(r, c);
  }
}

// RandomWalks::compute (SurfaceFeatures/Local/RandomWalks.cpp:70-117) on a given graph.  The reference draws from Random::common(), which it
// seeds from the clock; here the shared generator is put into the state of Random(seed) first (its members are reachable in this
// translation unit only), so that the reference's own code runs on a known stream.
void ref_random_walks(float const * p3, int64_t P, int max_degree, int64_t const * idx, float const * sep, float const * q3, int64_t Q, int64_t num_steps,
                      int64_t num_walks, uint32_t seed, double * features)
{
  PointSet3::Options opts = PointSet3::Options::defaults(); opts.setMaxDegree(max_degree);
  PointSet3 ps(opts);
  fill_samples(ps, p3, nullptr, P, 0);
  inject_graph(ps, P, max_degree, idx, sep);
  Random fresh(seed);
  Random & shared = Random::common();
  for (int i = 0; i < 624; ++i) shared.state[i] = fresh.state[i];
  shared.index = fresh.index;
  SurfaceFeatures::Local::RandomWalks rw(&ps);
  for (int64_t i = 0; i < Q; ++i) rw.compute(v3(q3 + 3 * i), (intx)num_steps, features + 3 * num_steps * i, (intx)num_walks);
}

int ref_tree_depth(int64_t n, int leaf) { return (int)Math::treeDepth((intx)n, leaf); }

}  // extern "C"
