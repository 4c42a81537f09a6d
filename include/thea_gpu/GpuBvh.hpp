// GpuBvh.hpp -- header-only C++ adaptor that re-exposes libtgq.so (include/thea_gpu_query.h) under the member-function
// names of Thea's Algorithms::BvhN / MeshBvh query concept, so that code written against
//   ProximityQueryStructureN   (Algorithms/ProximityQueryStructureN.hpp:29-35,150-276)
//   RayQueryStructureN         (Algorithms/RayQueryStructureN.hpp:55-88)
//   BvhN                       (Algorithms/BvhN.hpp:254-266,704-794,952-995,1036-1146,1266-1316)
//   MeshBvh                    (Algorithms/MeshBvh.hpp:55-141)
// can take a TheaGpu::GpuMeshBvh / GpuPointBvh as its structure template argument.  The reference interface is a
// compile-time concept ("None of the functions are virtual, this just defines a concept"), so being a drop-in means
// having the same member templates, argument meaning, return conventions (-1 index / negative time = not found) and
// error behaviour (Thea::Error == std::runtime_error, Error.hpp:27).  File:line citations are relative to
// /root/reference/Code/Source/.
//
// The header depends on nothing but the C ABI and the standard library.  It is templated on the 3-vector type, which
// only needs `operator[](int)` returning a float-convertible lvalue and a default constructor: Thea's own
// Vector3 (Eigen::Matrix<Real,3,1,DontAlign>, MatVec.hpp:56) qualifies, as does the bundled TheaGpu::Vec3.
//
// Scalar calls (one query per call, the reference's shape) are routed to a batch of one.  The `*Batch` members are the
// new surface: one C-ABI call for the whole loop that the reference's callers run serially.
//
// What is NOT supported, and fails loudly instead of falling back to a CPU path:
//   - pushFilter with arbitrary host callables and compatibility functors other than UniversalCompatibility (they cannot
//     run on the device): pushFilter throws, except for the built-in NNFilter (same label, compatible normal) on point
//     structures; the other device predicates are `exclude_self` and the ball-graph normal test.
//   - metrics other than L2; range queries on triangle structures (they need the reference's triangle-box / triangle-ball tests).
#ifndef THEA_GPU_GPUBVH_HPP
#define THEA_GPU_GPUBVH_HPP

#include "../thea_gpu_query.h"

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace TheaGpu {

typedef std::ptrdiff_t intx;   // Platform.hpp:161
typedef float Real;            // NumericType.hpp:40

/** Thea::Error (Error.hpp:27). */
typedef std::runtime_error Error;

/** Minimal 3-vector for callers that do not bring Thea's Vector3. */
struct Vec3
{
  float v[3];
  Vec3() { v[0] = v[1] = v[2] = 0; }
  Vec3(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
  float & operator[](int i) { return v[i]; }
  float const & operator[](int i) const { return v[i]; }
};

/** Tag with the one member of Algorithms::MetricL2 the structure itself uses (MetricL2.hpp:97). */
struct MetricL2
{
  static double invertMonotoneApprox(double d2) { return std::sqrt(d2); }
};
/** Tags standing in for Algorithms::RayIntersectionTester / IntersectionTester / UniversalCompatibility. */
struct RayIntersectionTester {};
struct IntersectionTester {};
struct UniversalCompatibility { template <typename U, typename V> bool operator()(U const &, V const &) const { return true; } };

/** ProximityQueryStructureN::NeighborPair (ProximityQueryStructureN.hpp:48-144), including its operator< quirk. */
template <typename VectorT>
class NeighborPairT
{
  public:
    NeighborPairT(intx q = -1, intx t = -1, double mad = -1) : query_index(q), target_index(t), mon_approx_dist(mad) {}
    NeighborPairT(intx q, intx t, double mad, VectorT const & qp, VectorT const & tp)
    : query_index(q), target_index(t), mon_approx_dist(mad), query_point(qp), target_point(tp) {}
    bool isValid() const { return query_index >= 0 && target_index >= 0; }
    intx getQueryIndex() const { return query_index; }
    void setQueryIndex(intx i) { query_index = i; }
    intx getTargetIndex() const { return target_index; }
    void setTargetIndex(intx i) { target_index = i; }
    double getMonotoneApproxDistance() const { return mon_approx_dist; }
    void setMonotoneApproxDistance(double d) { mon_approx_dist = d; }
    template <typename MetricT> double getDistance() const { return MetricT::invertMonotoneApprox(mon_approx_dist); }
    VectorT const & getQueryPoint() const { return query_point; }
    VectorT const & getTargetPoint() const { return target_point; }
    void setQueryPoint(VectorT const & p) { query_point = p; }
    void setTargetPoint(VectorT const & p) { target_point = p; }
    bool operator<(NeighborPairT const & o) const
    {
      return mon_approx_dist < o.mon_approx_dist
          || (mon_approx_dist == o.mon_approx_dist
           && (query_index < o.query_index || (query_index == o.query_index && target_index == o.target_index)));
    }
    bool operator==(NeighborPairT const & o) const { return query_index == o.query_index && target_index == o.target_index; }

  private:
    intx query_index, target_index;
    double mon_approx_dist;
    VectorT query_point, target_point;
};

/** RayStructureIntersectionN<3> (RayQueryStructureN.hpp:28-53; base RayIntersectionN, RayIntersectableN.hpp:27-65). */
template <typename VectorT>
class RayStructureIntersectionT
{
  public:
    RayStructureIntersectionT(Real time_ = -1, VectorT const * normal_ = nullptr, intx element_index_ = -1)
    : time(time_), has_normal(normal_ != nullptr), element_index(element_index_)
    { if (normal_) normal = *normal_; }
    bool isValid() const { return time >= 0; }
    Real getTime() const { return time; }
    void setTime(Real t) { time = t; }
    bool hasNormal() const { return has_normal; }
    VectorT const & getNormal() const { return normal; }
    intx getElementIndex() const { return element_index; }
    void setElementIndex(intx i) { element_index = i; }

    /** Converts to any type with the reference's constructor (time, normal *, element index) -- i.e. to Thea's own
     *  RayStructureIntersectionN (RayQueryStructureN.hpp:37-39), so that reference code such as
     *  `RayStructureIntersection3 isec = external ? external->rayStructureIntersection<...>(ray) : own->rayStructureIntersection<...>(ray);`
     *  (ShapeDiameter.hpp:215-217) compiles with this structure as the external one. */
    template <typename U, typename std::enable_if<std::is_constructible<U, Real, VectorT const *, intx>::value
                                               && !std::is_same<U, RayStructureIntersectionT>::value, int>::type = 0>
    operator U() const { return U(time, has_normal ? &normal : nullptr, element_index); }

  private:
    Real time; bool has_normal; VectorT normal; intx element_index;
};

/** Ball3 / AxisAlignedBox3 as range-query arguments (BallN.hpp:28-66, AxisAlignedBoxN.hpp:40-120): any type with
 *  getCenter() / getRadius(), respectively getLow() / getHigh(), works -- Thea's own classes included. */
template <typename VectorT>
struct Ball3T
{
  VectorT center; Real radius;
  Ball3T(VectorT const & c, Real r) : center(c), radius(r) {}
  VectorT const & getCenter() const { return center; }
  Real getRadius() const { return radius; }
};
template <typename VectorT>
struct AxisAlignedBox3T
{
  VectorT lo, hi;
  AxisAlignedBox3T() {}
  AxisAlignedBox3T(VectorT const & l, VectorT const & h_) : lo(l), hi(h_) {}
  VectorT const & getLow() const { return lo; }
  VectorT const & getHigh() const { return hi; }
  VectorT getExtent() const { VectorT e; for (int a = 0; a < 3; ++a) e[a] = hi[a] - lo[a]; return e; }                 // AxisAlignedBoxN.hpp:288
  VectorT getCenter() const { VectorT c; for (int a = 0; a < 3; ++a) c[a] = 0.5f * (hi[a] + lo[a]); return c; }        // :274
};

/** LineSegment3 as a query object (LineSegmentN.hpp:38-92): stored as (point, direction = end1 - end0), like the reference's class.  Any
 *  type with getEndpoint(0) and getDirection() works as a segment query -- Thea's own LineSegment3 included. */
template <typename VectorT>
struct LineSegment3T
{
  VectorT point, direction;
  LineSegment3T() {}
  LineSegment3T(VectorT const & end0, VectorT const & end1) : point(end0) { for (int a = 0; a < 3; ++a) direction[a] = end1[a] - end0[a]; }
  VectorT getEndpoint(int i) const { if (i == 0) return point; VectorT e; for (int a = 0; a < 3; ++a) e[a] = point[a] + direction[a]; return e; }
  VectorT const & getDirection() const { return direction; }
};

/** What getElements()[i] of a triangle structure offers the reference's callers (MeshBvh elements are Triangle3 objects;
 *  ShapeDiameter.hpp:152-153 reads getNormal(), Triangle3.hpp:143-155,206-219): the vertices as given and the cached unit normal as
 *  the DEVICE computed it with the reference's arithmetic (bit-equal to Triangle3's own cache). */
template <typename VectorT>
class TriangleElementT
{
  public:
    TriangleElementT() {}
    TriangleElementT(VectorT const & a, VectorT const & b, VectorT const & c, VectorT const & n) : normal(n) { v[0] = a; v[1] = b; v[2] = c; }
    VectorT const & getVertex(int i) const { return v[i]; }
    VectorT const & getNormal() const { return normal; }
  private:
    VectorT v[3], normal;
};

/** RayN<3> (RayN.hpp:27-139): origin + direction, direction NOT normalised by the structure. */
template <typename VectorT>
struct RayT
{
  VectorT origin, direction;
  RayT() {}
  RayT(VectorT const & o, VectorT const & d) : origin(o), direction(d) {}
  VectorT const & getOrigin() const { return origin; }
  VectorT const & getDirection() const { return direction; }
};

/** BoundedSortedArray<T> (BoundedSortedArray.hpp:28-308): the subset kClosestPairs and its callers use. */
template <typename T>
class BoundedSortedArray
{
  public:
    explicit BoundedSortedArray(size_t capacity_ = 0) : capacity(capacity_) { values.reserve(capacity_); }
    size_t getCapacity() const { return capacity; }
    size_t size() const { return values.size(); }
    bool empty() const { return values.empty(); }
    T const & operator[](size_t i) const { return values[i]; }
    T const & first() const { return values.front(); }
    T const & last() const { return values.back(); }
    void clear() { values.clear(); }
    bool isInsertable(T const & t) const { return capacity > 0 && (values.size() < capacity || t < values.back()); }
    /** Insert keeping ascending order, after any equal elements (upper bound, BoundedSortedArray.hpp:243-265). */
    size_t insert(T const & t)
    {
      if (!isInsertable(t)) return capacity;
      size_t lo = 0, hi = values.size();
      while (lo < hi) { size_t mid = (lo + hi) / 2; if (t < values[mid]) hi = mid; else lo = mid + 1; }
      if (values.size() == capacity) values.pop_back();
      values.insert(values.begin() + (std::ptrdiff_t)lo, t);
      return lo;
    }

  private:
    size_t capacity;
    std::vector<T> values;
};

namespace detail {

inline void check(tgq_handle * h, int rc, char const * what)
{
  if (rc != TGQ_OK)
    throw Error(std::string("TheaGpu: ") + what + " failed (" + std::to_string(rc) + "): " + tgq_last_error(h));
}

/** Query-object kinds by their members: a line segment has getEndpoint(i) and getDirection() (LineSegmentN.hpp:72-78), a triangle has
 *  getVertex(i) and getNormal() (Triangle3.hpp:130-141). */
template <typename T, typename = void> struct is_segment : std::false_type {};
template <typename T> struct is_segment<T, decltype((void)std::declval<T const &>().getDirection(), (void)std::declval<T const &>().getEndpoint(0))> : std::true_type {};
template <typename T, typename = void> struct is_triangle : std::false_type {};
template <typename T> struct is_triangle<T, decltype((void)std::declval<T const &>().getVertex(0), (void)std::declval<T const &>().getNormal())> : std::true_type {};

template <typename VectorT> inline void put3(float * dst, VectorT const & v) { dst[0] = (float)v[0]; dst[1] = (float)v[1]; dst[2] = (float)v[2]; }
template <typename VectorT> inline VectorT get3(float const * src) { VectorT v; v[0] = src[0]; v[1] = src[1]; v[2] = src[2]; return v; }

}  // namespace detail

/**
 * The one element filter the device evaluates: same label and compatible normal, the NNFilter of
 * Tools/Register/Register.cpp:92-101 (`sample->label == label && n.dot(sample->n) >= 0`).  Push it around a query on a
 * GpuPointBvh exactly as Register pushes its filter on a BvhN (Register.cpp:218-223,274-277).
 */
template <typename VectorT>
struct NNFilterT
{
  VectorT n; int32_t label; bool use_normals, use_labels; float min_dot;
  NNFilterT(VectorT const & n_, int32_t label_, bool use_normals_ = true, bool use_labels_ = true, float min_dot_ = 0.0f)
  : n(n_), label(label_), use_normals(use_normals_), use_labels(use_labels_), min_dot(min_dot_) {}
};

/**
 * Common part of the two structures: owns the engine handle.  Mirrors the non-query members of BvhN
 * (numElements / isEmpty / getBounds / clear, BvhN.hpp:796-813,952-995; enable/disableNearestNeighborAcceleration,
 * :853-883; pushFilter / popFilter, :1005-1032).
 */
template <typename VectorT>
class GpuBvhBase
{
  public:
    typedef NeighborPairT<VectorT> NeighborPair;
    typedef RayStructureIntersectionT<VectorT> RayStructureIntersection3;
    typedef RayT<VectorT> Ray3;

    explicit GpuBvhBase(int device = 0) : h(nullptr), dev(device)
    {
      int rc = tgq_create(device, &h);
      if (rc != TGQ_OK) throw Error(std::string("TheaGpu: no engine (") + std::to_string(rc) + "): " + tgq_last_error(nullptr));
    }
    ~GpuBvhBase() { if (h) tgq_destroy(h); }
    GpuBvhBase(GpuBvhBase const &) = delete;
    GpuBvhBase & operator=(GpuBvhBase const &) = delete;

    tgq_handle * handle() const { return h; }
    int device() const { return dev; }
    void setStream(void * cuda_stream) { detail::check(h, tgq_set_stream(h, cuda_stream), "tgq_set_stream"); }
    void setOption(int option, int64_t value) { detail::check(h, tgq_set_option(h, option, value), "tgq_set_option"); }

    intx numElements() const { int64_t n = 0; detail::check(h, tgq_info(h, nullptr, &n, nullptr, nullptr), "tgq_info"); return (intx)n; }
    bool isEmpty() const { return numElements() <= 0; }
    bool empty() const { return isEmpty(); }
    /** BvhN::getBounds (BvhN.hpp:990-995,1740-1756): the reference returns its ROOT NODE's box, i.e. the union of the element
     *  boxes scaled by 1.05 about its centre (createTree pads every node box, BvhN.hpp:1437-1439; AxisAlignedBoxN::scaleCentered,
     *  AxisAlignedBoxN.hpp:316-324) -- reproduced here in the same fp32 operations, because callers derive scales from it
     *  (ShapeDiameter.hpp:116, Visibility.hpp:73). */
    AxisAlignedBox3T<VectorT> getBounds() const
    {
      float l[3], u[3]; detail::check(h, tgq_info(h, nullptr, nullptr, l, u), "tgq_info");
      for (int a = 0; a < 3; ++a)
      {
        float c = 0.5f * (u[a] + l[a]);
        float lo = 1.05f * (l[a] - c); lo = lo + c;
        float hi = 1.05f * (u[a] - c); hi = hi + c;
        l[a] = lo; u[a] = hi;
      }
      return AxisAlignedBox3T<VectorT>(detail::get3<VectorT>(l), detail::get3<VectorT>(u));
    }
    void getBounds(VectorT & lo, VectorT & hi) const { AxisAlignedBox3T<VectorT> b = getBounds(); lo = b.getLow(); hi = b.getHigh(); }
    /** The exact (unpadded) bounds of the elements. */
    void getElementBounds(VectorT & lo, VectorT & hi) const
    {
      float l[3], u[3]; detail::check(h, tgq_info(h, nullptr, nullptr, l, u), "tgq_info");
      lo = detail::get3<VectorT>(l); hi = detail::get3<VectorT>(u);
    }

    /** BvhN::setTransform / clearTransform (BvhN.hpp:903-925): queries see every element transformed by the affine map
     *  (row-major 3 x 4 [L | t]); the reference evaluates its metrics on the transformed element (MetricL2_Transform.hpp). */
    void setTransform(float const * tr12) { detail::check(h, tgq_set_transform(h, tr12), "tgq_set_transform"); }
    void clearTransform() { detail::check(h, tgq_set_transform(h, nullptr), "tgq_set_transform"); }

    /** Result-neutral in the reference (it only seeds the pruning bound, BvhN.hpp:2365-2377): accepted and ignored. */
    void enableNearestNeighborAcceleration(intx = -1) {}
    void disableNearestNeighborAcceleration(bool = true) {}
    bool hasNearestNeighborAcceleration() const { return false; }

    /** BvhN::closestPair<MetricL2>(query_structure, dist_bound, compatibility, get_closest_points) (BvhN.hpp:1081-1114,1900-2009;
     *  Test/TestBvh.cpp:251-258,380-385): the closest pair of elements between this structure and another one (points or triangles) on the
     *  same device; pairs are evaluated as MetricL2::closestPoints(target element, query element), see tgq_closest_pair. */
    template <typename MetricT, typename QueryVectorT, typename CompatibilityFunctorT = UniversalCompatibility>
    NeighborPair closestPair(GpuBvhBase<QueryVectorT> const & query, double dist_bound = -1, CompatibilityFunctorT = CompatibilityFunctorT(),
                             bool get_closest_points = false) const
    {
      int64_t ti = -1, qi = -1; float d2 = -1, qp[3] = {0, 0, 0}, tp[3] = {0, 0, 0};
      detail::check(h, tgq_closest_pair(h, query.handle(), dist_bound, &ti, &qi, &d2, get_closest_points ? qp : nullptr, get_closest_points ? tp : nullptr),
                    "tgq_closest_pair");
      if (ti < 0) return NeighborPair();
      return NeighborPair((intx)qi, (intx)ti, (double)d2, detail::get3<VectorT>(qp), detail::get3<VectorT>(tp));
    }

    /** BvhN::kClosestPairs<MetricL2>(query_structure, k_closest_pairs, dist_bound, ...) (BvhN.hpp:1137-1146 with a structure as the query):
     *  the k closest pairs between this POINT structure and the point structure `query`, offered to the caller's bounded set in ascending
     *  order through its own isInsertable / insert (the set is not cleared, as in the reference).  Returns its size. */
    template <typename MetricT, typename QueryVectorT, typename BoundedNeighborPairSetT, typename CompatibilityFunctorT = UniversalCompatibility>
    intx kClosestPairs(GpuBvhBase<QueryVectorT> const & query, BoundedNeighborPairSetT & k_closest_pairs, double dist_bound = -1,
                       CompatibilityFunctorT = CompatibilityFunctorT(), bool = false) const
    {
      int32_t k = (int32_t)k_closest_pairs.getCapacity(), found = 0;
      if (k <= 0) return (intx)k_closest_pairs.size();
      if (k > TGQ_MAX_K) throw Error("TheaGpu: kClosestPairs capacity exceeds TGQ_MAX_K");
      std::vector<int64_t> ti((size_t)k), qi((size_t)k); std::vector<float> d2((size_t)k);
      detail::check(h, tgq_k_closest_pairs(h, query.handle(), k, dist_bound, ti.data(), qi.data(), d2.data(), &found), "tgq_k_closest_pairs");
      for (int32_t j = 0; j < found; ++j)
      {
        NeighborPair pair((intx)qi[(size_t)j], (intx)ti[(size_t)j], (double)d2[(size_t)j]);
        if (k_closest_pairs.isInsertable(pair)) k_closest_pairs.insert(pair);
      }
      return (intx)k_closest_pairs.size();
    }

    /** BvhN::closestPair<MetricL2>(LineSegment3, dist_bound, compatibility, get_closest_points) (BvhN.hpp:1081-1114 with a bounded,
     *  non-structure query; per-element metric LineSegmentN.hpp:106-145 on point structures, Triangle3.hpp:495-509,559-577 on triangle
     *  structures): query index 0, query point on the segment, target point on the element. */
    template <typename MetricT, typename SegmentT, typename CompatibilityFunctorT = UniversalCompatibility,
              typename std::enable_if<detail::is_segment<SegmentT>::value, int>::type = 0>
    NeighborPair closestPair(SegmentT const & seg, double dist_bound = -1, CompatibilityFunctorT = CompatibilityFunctorT(),
                             bool get_closest_points = false) const
    {
      float a[3], d[3], d2 = -1, ep[3] = {0, 0, 0}, sp[3] = {0, 0, 0}; int32_t idx = -1;
      detail::put3(a, seg.getEndpoint(0)); detail::put3(d, seg.getDirection());
      detail::check(h, tgq_closest_segment(h, a, d, 1, 1, dist_bound, nullptr, 0, &idx, &d2, get_closest_points ? ep : nullptr, get_closest_points ? sp : nullptr),
                    "tgq_closest_segment");
      if (idx < 0) return NeighborPair();
      return NeighborPair(0, idx, (double)d2, detail::get3<VectorT>(sp), detail::get3<VectorT>(ep));
    }
    /** BvhN::closestElement<MetricL2>(LineSegment3, dist_bound, compatibility, &dist, &closest_point) (BvhN.hpp:1049-1063;
     *  Tools/MeshFix/MeshFix.cpp:762): *closest_point is the point on the ELEMENT (NeighborPair::getTargetPoint). */
    template <typename MetricT, typename SegmentT, typename CompatibilityFunctorT = UniversalCompatibility,
              typename std::enable_if<detail::is_segment<SegmentT>::value, int>::type = 0>
    intx closestElement(SegmentT const & seg, double dist_bound = -1, CompatibilityFunctorT c = CompatibilityFunctorT(), double * dist = nullptr,
                        VectorT * closest_point = nullptr) const
    {
      NeighborPair pair = closestPair<MetricT>(seg, dist_bound, c, closest_point != nullptr);
      if (!pair.isValid()) return -1;
      if (dist) *dist = MetricT::invertMonotoneApprox(pair.getMonotoneApproxDistance());
      if (closest_point) *closest_point = pair.getTargetPoint();
      return pair.getTargetIndex();
    }
    /** BvhN::closestPair<MetricL2>(Triangle3, dist_bound, compatibility, get_closest_points) (BvhN.hpp:1081-1114 with a triangle as the
     *  bounded query object; per-element metric Triangle3.hpp:314-392 on triangle structures, :294-304 on point structures): query index
     *  0, query point on the query triangle, target point on the element.  Any type with getVertex(i) and getNormal() is a triangle. */
    template <typename MetricT, typename TriangleT, typename CompatibilityFunctorT = UniversalCompatibility,
              typename std::enable_if<detail::is_triangle<TriangleT>::value, long>::type = 0>
    NeighborPair closestPair(TriangleT const & tri, double dist_bound = -1, CompatibilityFunctorT = CompatibilityFunctorT(),
                             bool get_closest_points = false) const
    {
      float q[9], d2 = -1, ep[3] = {0, 0, 0}, qp[3] = {0, 0, 0}; int32_t idx = -1;
      detail::put3(q, tri.getVertex(0)); detail::put3(q + 3, tri.getVertex(1)); detail::put3(q + 6, tri.getVertex(2));
      detail::check(h, tgq_closest_triangle(h, q, 1, dist_bound, &idx, &d2, get_closest_points ? ep : nullptr, get_closest_points ? qp : nullptr),
                    "tgq_closest_triangle");
      if (idx < 0) return NeighborPair();
      return NeighborPair(0, idx, (double)d2, detail::get3<VectorT>(qp), detail::get3<VectorT>(ep));
    }
    template <typename MetricT, typename TriangleT, typename CompatibilityFunctorT = UniversalCompatibility,
              typename std::enable_if<detail::is_triangle<TriangleT>::value, long>::type = 0>
    intx closestElement(TriangleT const & tri, double dist_bound = -1, CompatibilityFunctorT c = CompatibilityFunctorT(), double * dist = nullptr,
                        VectorT * closest_point = nullptr) const
    {
      NeighborPair pair = closestPair<MetricT>(tri, dist_bound, c, closest_point != nullptr);
      if (!pair.isValid()) return -1;
      if (dist) *dist = MetricT::invertMonotoneApprox(pair.getMonotoneApproxDistance());
      if (closest_point) *closest_point = pair.getTargetPoint();
      return pair.getTargetIndex();
    }
    /** Q triangle queries in one launch (q9: Q x 9 floats; outputs may be null; host or device pointers). */
    void closestElementTriangleBatch(float const * q9, intx Q, int32_t * idx, float * d2, float * element_point3, float * query_point3, double dist_bound = -1) const
    { detail::check(h, tgq_closest_triangle(h, q9, Q, dist_bound, idx, d2, element_point3, query_point3), "tgq_closest_triangle"); }

    /** BvhN::distance<MetricL2>(query object, dist_bound, compatibility) (BvhN.hpp:1036-1047) for a segment or a triangle: -1 when nothing
     *  lies within the bound. */
    template <typename MetricT, typename QueryT, typename CompatibilityFunctorT = UniversalCompatibility,
              typename std::enable_if<detail::is_segment<QueryT>::value || detail::is_triangle<QueryT>::value, short>::type = 0>
    double distance(QueryT const & query, double dist_bound = -1, CompatibilityFunctorT c = CompatibilityFunctorT()) const
    {
      double result = -1;
      return closestElement<MetricT>(query, dist_bound, c, &result) >= 0 ? result : -1;
    }

    /** Q segment queries in one launch: the loop of Tools/MeshFix/MeshFix.cpp:756-763 (one closestElement per boundary edge, each under
     *  a VertexFilter) as a batch.  a3 / b3: Q x 3 endpoints (b_is_direction: b3 holds the stored directions instead); exclude: Q x
     *  n_exclude element indices each query's filter rejects (-1 = unused, n_exclude <= 4) or null; outputs may be null; host or device
     *  pointers. */
    void closestElementSegmentBatch(float const * a3, float const * b3, intx Q, int32_t * idx, float * d2, float * element_point3, float * segment_point3,
                                    double dist_bound = -1, int32_t const * exclude = nullptr, int32_t n_exclude = 0, bool b_is_direction = false) const
    {
      detail::check(h, tgq_closest_segment(h, a3, b3, b_is_direction ? 1 : 0, Q, dist_bound, exclude, n_exclude, idx, d2, element_point3, segment_point3),
                    "tgq_closest_segment");
    }

    /** Host-side filters cannot run on the device and there is no CPU fallback. */
    template <typename FilterT> void pushFilter(FilterT *) { throw Error("TheaGpu: element filters are not supported by the GPU structure"); }
    void popFilter() {}

  protected:
    tgq_handle * h; int dev;
};

/**
 * Triangle structure: Algorithms::MeshBvh (MeshBvh.hpp:36-141) restricted to the query surface.  Element index = position
 * of the triangle in the order it was added (MeshTriangles::addTriangles push order, MeshTriangles.hpp:411-418).
 */
template <typename VectorT = Vec3>
class GpuMeshBvhT : public GpuBvhBase<VectorT>
{
    typedef GpuBvhBase<VectorT> Base;
    using Base::h;

  public:
    typedef typename Base::NeighborPair NeighborPair;
    typedef typename Base::RayStructureIntersection3 RayStructureIntersection3;
    typedef typename Base::Ray3 Ray3;

    explicit GpuMeshBvhT(int device = 0) : Base(device), built(false) {}
    using Base::closestPair;         // the structure-to-structure and line-segment forms
    using Base::closestElement;      // the line-segment and triangle forms
    using Base::distance;

    /** MeshBvh::addTriangles analogue: appends one triangle to the pending soup (MeshTriangles.hpp:361-370). */
    void addTriangle(VectorT const & v0, VectorT const & v1, VectorT const & v2)
    {
      size_t o = soup.size(); soup.resize(o + 9);
      detail::put3(&soup[o], v0); detail::put3(&soup[o + 3], v1); detail::put3(&soup[o + 6], v2);
    }
    /** Appends T triangles given as T x 9 floats. */
    void addTriangles(float const * v9, intx T) { soup.insert(soup.end(), v9, v9 + 9 * T); }
    /** Appends an indexed mesh; faces must be triangles (quads / n-gons are triangulated by the caller as
     *  MeshTriangles.hpp:173-339 does). */
    void addMesh(float const * verts3, intx V, int32_t const * tri3, intx T)
    {
      for (intx t = 0; t < T; ++t)
        for (int c = 0; c < 3; ++c)
        {
          int32_t v = tri3[3 * t + c];
          if (v < 0 || v >= V) throw Error("TheaGpu: vertex index out of range");
          soup.insert(soup.end(), verts3 + 3 * v, verts3 + 3 * v + 3);
        }
    }

    /** MeshBvh::init (MeshBvh.hpp:104-111): builds over the pending triangles and, like the reference, clears the
     *  triangle cache.  The BvhN tuning arguments (method, max_depth, max_elems_per_leaf, save_memory) are accepted for
     *  source compatibility; only max_elems_per_leaf is meaningful to the LBVH. */
    void init(int /* method */ = 0, intx /* max_depth */ = -1, intx max_elems_per_leaf = -1, bool /* save_memory */ = false)
    {
      if (max_elems_per_leaf > 0) this->setOption(TGQ_OPT_LEAF_SIZE, max_elems_per_leaf > 15 ? 15 : max_elems_per_leaf);
      detail::check(h, tgq_build_triangles(h, soup.data(), (int64_t)(soup.size() / 9)), "tgq_build_triangles");
      kept.swap(soup); std::vector<float>().swap(soup); elements.clear(); built = true;      // the host copy backs getElements()
    }
    /** Build straight from caller-owned (host or device) memory (no host copy is kept: getElements() is then unavailable). */
    void init(float const * v9, intx T)
    { detail::check(h, tgq_build_triangles(h, v9, T), "tgq_build_triangles"); kept.clear(); elements.clear(); built = true; }

    /** MeshBvh::add(mesh) (MeshBvh.hpp:55-67) for any mesh-like object that can enumerate its triangles: `visit` is called with a
     *  callable taking three vertex positions, once per triangle, in the order MeshTriangles would push them
     *  (MeshTriangles.hpp:173-339,411-418) -- with Thea: `bvh.add([&](auto emit) { MeshTriangles<Mesh> t; t.add(mesh); for (auto const &
     *  x : t.getTriangles()) emit(x.getVertex(0), x.getVertex(1), x.getVertex(2)); });` */
    template <typename VisitorT> void add(VisitorT visit)
    { visit([this](VectorT const & a, VectorT const & b, VectorT const & c) { this->addTriangle(a, b, c); }); }

    void clear() { soup.clear(); kept.clear(); elements.clear(); detail::check(h, tgq_build_triangles(h, nullptr, 0), "tgq_build_triangles"); built = false; }

    /** BvhN::getElements() (BvhN.hpp:952-965): the elements in index order.  Backed by the host copy of the triangles given to
     *  add* / init() and the device-computed triangle cache; fetched once per build. */
    std::vector< TriangleElementT<VectorT> > const & getElements() const
    {
      size_t T = kept.size() / 9;
      if (elements.size() != T)
      {
        if ((intx)T != this->numElements()) throw Error("TheaGpu: getElements() needs the structure to be built from host triangles (add* + init())");
        std::vector<float> n3; getElementNormals(n3);
        elements.resize(T);
        for (size_t i = 0; i < T; ++i)
          elements[i] = TriangleElementT<VectorT>(detail::get3<VectorT>(&kept[9 * i]), detail::get3<VectorT>(&kept[9 * i + 3]),
                                                  detail::get3<VectorT>(&kept[9 * i + 6]), detail::get3<VectorT>(&n3[3 * i]));
      }
      return elements;
    }

    /** MeshBvh::updateElements + BvhN::updateNodeBounds (MeshBvh.hpp:123-134, BvhN.hpp:816-834): same triangles, moved
     *  vertices. */
    void updateElements(float const * v9, intx T)
    {
      detail::check(h, tgq_refit_triangles(h, v9, T), "tgq_refit_triangles");
      if (!kept.empty()) kept.assign(v9, v9 + 9 * T);        // (host pointers only, like the build that filled it)
      elements.clear();
    }

    /** getElements()[i].getNormal() for all i (used by ShapeDiameter.hpp:152-153). */
    void getElementNormals(std::vector<float> & n3) const
    {
      n3.resize(3 * (size_t)this->numElements());
      detail::check(h, tgq_triangle_cache(h, n3.data(), nullptr, nullptr, nullptr), "tgq_triangle_cache");
    }

    // ---- ProximityQueryStructureN ---------------------------------------------------------------------------------------

    /** BvhN::closestElement (BvhN.hpp:1049-1063). */
    template <typename MetricT, typename CompatibilityFunctorT = UniversalCompatibility>
    intx closestElement(VectorT const & query, double dist_bound = -1, CompatibilityFunctorT = CompatibilityFunctorT(),
                        double * dist = nullptr, VectorT * closest_point = nullptr) const
    {
      float q[3], d2 = -1, cp[3]; int32_t idx = -1;
      detail::put3(q, query);
      detail::check(h, tgq_closest(h, q, 1, dist_bound, 0, &idx, &d2, closest_point ? cp : nullptr), "tgq_closest");
      if (idx >= 0)
      {
        if (dist) *dist = MetricT::invertMonotoneApprox((double)d2);
        if (closest_point) *closest_point = detail::get3<VectorT>(cp);
      }
      return idx;
    }

    /** BvhN::distance (BvhN.hpp:1036-1047). */
    template <typename MetricT, typename CompatibilityFunctorT = UniversalCompatibility>
    double distance(VectorT const & query, double dist_bound = -1, CompatibilityFunctorT c = CompatibilityFunctorT()) const
    {
      double result = -1;
      return closestElement<MetricT>(query, dist_bound, c, &result) >= 0 ? result : -1;
    }

    /** BvhN::closestPair (BvhN.hpp:1081-1114) for a point query: query index 0, as the reference sets it for
     *  non-structure queries. */
    template <typename MetricT, typename CompatibilityFunctorT = UniversalCompatibility>
    NeighborPair closestPair(VectorT const & query, double dist_bound = -1, CompatibilityFunctorT = CompatibilityFunctorT(),
                             bool get_closest_points = false) const
    {
      float q[3], d2 = -1, cp[3] = {0, 0, 0}; int32_t idx = -1;
      detail::put3(q, query);
      detail::check(h, tgq_closest(h, q, 1, dist_bound, 0, &idx, &d2, get_closest_points ? cp : nullptr), "tgq_closest");
      if (idx < 0) return NeighborPair();
      return NeighborPair(0, idx, (double)d2, query, detail::get3<VectorT>(cp));
    }

    /** Q calls of closestElement in one launch.  q3: Q x 3 floats; outputs may be null; host or device pointers. */
    void closestElementBatch(float const * q3, intx Q, int32_t * idx, float * d2, float * cp3, double dist_bound = -1) const
    { detail::check(h, tgq_closest(h, q3, Q, dist_bound, 0, idx, d2, cp3), "tgq_closest"); }

    // ---- RayQueryStructureN ---------------------------------------------------------------------------------------------

    /** BvhN::rayIntersects (BvhN.hpp:1266-1269). */
    template <typename RayIntersectionTesterT, typename RayLikeT = Ray3>
    bool rayIntersects(RayLikeT const & ray, Real max_time = -1) const
    {
      float o[3], d[3]; uint8_t hit = 0; detail::put3(o, ray.getOrigin()); detail::put3(d, ray.getDirection());
      detail::check(h, tgq_ray_any(h, o, d, &max_time, 1, &hit), "tgq_ray_any");
      return hit != 0;
    }
    /** BvhN::rayIntersectionTime (BvhN.hpp:1271-1289). */
    template <typename RayIntersectionTesterT, typename RayLikeT = Ray3>
    Real rayIntersectionTime(RayLikeT const & ray, Real max_time = -1) const
    {
      float o[3], d[3], t = -1; detail::put3(o, ray.getOrigin()); detail::put3(d, ray.getDirection());
      detail::check(h, tgq_ray_first(h, o, d, &max_time, 1, &t, nullptr, nullptr), "tgq_ray_first");
      return t;
    }
    /** BvhN::rayStructureIntersection (BvhN.hpp:1291-1316). */
    template <typename RayIntersectionTesterT, typename RayLikeT = Ray3>
    RayStructureIntersection3 rayStructureIntersection(RayLikeT const & ray, Real max_time = -1) const
    {
      float o[3], d[3], t = -1, n[3]; int32_t idx = -1; detail::put3(o, ray.getOrigin()); detail::put3(d, ray.getDirection());
      detail::check(h, tgq_ray_first(h, o, d, &max_time, 1, &t, &idx, n), "tgq_ray_first");
      if (idx < 0) return RayStructureIntersection3();
      VectorT nv = detail::get3<VectorT>(n);
      return RayStructureIntersection3(t, &nv, idx);
    }
    /** R first-hit queries in one launch (tmax may be null = unbounded). */
    void rayStructureIntersectionBatch(float const * o3, float const * d3, float const * tmax, intx R, float * t, int32_t * idx,
                                       float * normal3 = nullptr) const
    { detail::check(h, tgq_ray_first(h, o3, d3, tmax, R, t, idx, normal3), "tgq_ray_first"); }
    void rayIntersectsBatch(float const * o3, float const * d3, float const * tmax, intx R, uint8_t * hit) const
    { detail::check(h, tgq_ray_any(h, o3, d3, tmax, R, hit), "tgq_ray_any"); }

    // ---- callers' building blocks ---------------------------------------------------------------------------------------

    /** The loop of ShapeDiameter::compute over S samples (ShapeDiameter.hpp:168-258). */
    void shapeDiameterBatch(float const * p3, float const * n3, intx S, float scale, bool only_hit_interior_surfaces, double * sdf,
                            float * ray_t = nullptr, int32_t * ray_idx = nullptr) const
    { detail::check(h, tgq_sdf(h, p3, n3, S, scale, only_hit_interior_surfaces ? 1 : 0, sdf, ray_t, ray_idx), "tgq_sdf"); }

    /** The loop of Visibility::compute over S points (SurfaceFeatures/Local/Visibility.hpp:104-125) with caller-supplied ray
     *  directions (num_dirs x 3 shared, or S x num_dirs x 3 when per_sample_dirs). */
    void visibilityBatch(float const * p3, intx S, float const * dirs3, int32_t num_dirs, bool per_sample_dirs, float scale, double * visibility,
                         int32_t * num_escaped = nullptr) const
    { detail::check(h, tgq_visibility(h, p3, S, dirs3, num_dirs, per_sample_dirs ? 1 : 0, scale, visibility, num_escaped), "tgq_visibility"); }

    /** MeshSampler::sampleEvenlyByArea, EXACT mode (MeshSampler.hpp:147-185), with the engine's counter-based generator. */
    void sampleEvenlyByArea(intx S, uint64_t seed, float * p3, float * n3 = nullptr, int32_t * tri = nullptr) const
    { detail::check(h, tgq_area_sample(h, S, seed, p3, n3, tri), "tgq_area_sample"); }

  private:
    std::vector<float> soup, kept;
    mutable std::vector< TriangleElementT<VectorT> > elements;
    bool built;
};

/**
 * Point structure: BvhN<Sample*, 3> of PointSet3 (PointSet3.hpp:329-358,709-723) and BvhN<Vector3, 3> of the k-NN
 * callers (Tools/Register/Register.cpp:222).  Element index = position in the build array.
 */
template <typename VectorT = Vec3>
class GpuPointBvhT : public GpuBvhBase<VectorT>
{
    typedef GpuBvhBase<VectorT> Base;
    using Base::h;

  public:
    typedef typename Base::NeighborPair NeighborPair;

    typedef NNFilterT<VectorT> NNFilter;

    explicit GpuPointBvhT(int device = 0) : Base(device), filter(nullptr) {}
    using Base::closestPair;         // the structure-to-structure and line-segment forms
    using Base::closestElement;      // the line-segment and triangle forms
    using Base::distance;
    using Base::kClosestPairs;       // likewise

    /** Per-element labels (P int32, element order; null clears): the `label` member of Register's samples. */
    void setLabels(int32_t const * labels) { detail::check(h, tgq_set_point_labels(h, labels), "tgq_set_point_labels"); }
    /** BvhN::pushFilter / popFilter (BvhN.hpp:1005-1032) for the device-evaluated NNFilter; one level deep, like every use
     *  in the reference.  Any other filter type falls to the base template and throws. */
    using Base::pushFilter;
    void pushFilter(NNFilter const * f) { if (filter) throw Error("TheaGpu: only one NNFilter can be active"); filter = f; }
    void pushFilter(NNFilter * f) { pushFilter(static_cast<NNFilter const *>(f)); }   // (a non-const pointer would pick the throwing template)
    void popFilter() { filter = nullptr; }

    /** BvhN::init over P points with optional unit normals (P x 3 each; host or device pointers). */
    void init(float const * p3, intx P, float const * n3 = nullptr) { detail::check(h, tgq_build_points(h, p3, P, n3), "tgq_build_points"); elems.clear(); }
    /** BvhN::init(begin, end) (BvhN.hpp:704-794) over a host range of points: the copy the reference makes (:722-724) backs getElements(). */
    template <typename IteratorT> void init(IteratorT begin, IteratorT end)
    {
      elems.assign(begin, end);
      std::vector<float> p3(3 * elems.size());
      for (size_t i = 0; i < elems.size(); ++i) detail::put3(&p3[3 * i], elems[i]);
      detail::check(h, tgq_build_points(h, p3.data(), (int64_t)elems.size(), nullptr), "tgq_build_points");
    }
    void clear() { elems.clear(); detail::check(h, tgq_build_points(h, nullptr, 0, nullptr), "tgq_build_points"); }
    /** BvhN::getElements() (BvhN.hpp:952-965); available when the structure was built from a host range (init(begin, end)). */
    std::vector<VectorT> const & getElements() const
    {
      if ((intx)elems.size() != this->numElements()) throw Error("TheaGpu: getElements() needs the structure to be built with init(begin, end)");
      return elems;
    }

    // ---- RangeQueryStructure (Algorithms/RangeQueryStructure.hpp:27-95; BvhN.hpp:1208-1264) -------------------------------------
    // RangeT: a ball (getCenter / getRadius) or an axis-aligned box (getLow / getHigh) -- Thea's Ball3 / AxisAlignedBox3 or the
    // bundled Ball3T / AxisAlignedBox3T.  The elements in range are found on the device in one call and reported in ASCENDING
    // element index (the reference reports them in its tree's traversal order).

    /** BvhN::rangeQueryIndices (BvhN.hpp:1215-1220). */
    template <typename IntersectionTesterT, typename RangeT, typename IndexArrayT>
    void rangeQueryIndices(RangeT const & range, IndexArrayT & result, bool discard_prior_results = true) const
    {
      if (discard_prior_results) result.clear();
      std::vector<int32_t> idx; rangeIndices(range, idx);
      for (size_t j = 0; j < idx.size(); ++j) result.push_back((intx)idx[j]);
    }
    /** BvhN::rangeQuery (BvhN.hpp:1208-1213): the elements themselves. */
    template <typename IntersectionTesterT, typename RangeT, typename ElementArrayT>
    void rangeQuery(RangeT const & range, ElementArrayT & result, bool discard_prior_results = true) const
    {
      if (discard_prior_results) result.clear();
      std::vector<int32_t> idx; rangeIndices(range, idx);
      std::vector<VectorT> const & e = getElements();
      for (size_t j = 0; j < idx.size(); ++j) result.push_back(e[(size_t)idx[j]]);
    }
    /** BvhN::processRange (BvhN.hpp:1238-1264): functor(index, element) for every element in range until it returns true; returns
     *  that element's index, else a negative value.  The functor runs on the host over the device's result. */
    template <typename IntersectionTesterT, typename RangeT, typename FunctorT>
    intx processRange(RangeT const & range, FunctorT functor) const
    {
      std::vector<int32_t> idx; rangeIndices(range, idx);
      std::vector<VectorT> const & e = getElements();
      for (size_t j = 0; j < idx.size(); ++j)
        if (functor((intx)idx[j], e[(size_t)idx[j]])) return (intx)idx[j];
      return -1;
    }
    /** Q ranges in one call, CSR output: kind 0 = balls (a3 centres, b radii), 1 = boxes (a3 low, b high corners). */
    void rangeQueryBatch(int kind, float const * a3, float const * b, intx Q, std::vector<int64_t> & offsets, std::vector<int32_t> & idx) const
    {
      offsets.assign((size_t)Q + 1, 0); int64_t total = 0;
      detail::check(h, tgq_range_query(h, kind, a3, b, Q, offsets.data(), nullptr, 0, &total), "tgq_range_query");
      idx.resize((size_t)total);
      if (total > 0) detail::check(h, tgq_range_query(h, kind, a3, b, Q, offsets.data(), idx.data(), total, &total), "tgq_range_query");
    }

    template <typename MetricT, typename CompatibilityFunctorT = UniversalCompatibility>
    intx closestElement(VectorT const & query, double dist_bound = -1, CompatibilityFunctorT = CompatibilityFunctorT(),
                        double * dist = nullptr, VectorT * closest_point = nullptr) const
    {
      float q[3], d2 = -1, cp[3]; int32_t idx = -1;
      detail::put3(q, query);
      if (filter)
      {
        float fn[3]; detail::put3(fn, filter->n);
        detail::check(h, tgq_closest_filtered(h, q, filter->use_normals ? fn : nullptr, filter->use_labels ? &filter->label : nullptr, 1,
                                              filter->min_dot, dist_bound, &idx, &d2, closest_point ? cp : nullptr), "tgq_closest_filtered");
      }
      else detail::check(h, tgq_closest(h, q, 1, dist_bound, 0, &idx, &d2, closest_point ? cp : nullptr), "tgq_closest");
      if (idx >= 0)
      {
        if (dist) *dist = MetricT::invertMonotoneApprox((double)d2);
        if (closest_point) *closest_point = detail::get3<VectorT>(cp);
      }
      return idx;
    }
    template <typename MetricT, typename CompatibilityFunctorT = UniversalCompatibility>
    double distance(VectorT const & query, double dist_bound = -1, CompatibilityFunctorT c = CompatibilityFunctorT()) const
    {
      double result = -1;
      return closestElement<MetricT>(query, dist_bound, c, &result) >= 0 ? result : -1;
    }
    /** Q filtered closestElement calls (Register.cpp:262-287 updateOffsets): qn3 / qlabel per query, either may be null. */
    void closestElementFilteredBatch(float const * q3, float const * qn3, int32_t const * qlabel, intx Q, int32_t * idx, float * d2, float * cp3,
                                     float min_dot = 0.0f, double dist_bound = -1) const
    { detail::check(h, tgq_closest_filtered(h, q3, qn3, qlabel, Q, min_dot, dist_bound, idx, d2, cp3), "tgq_closest_filtered"); }
    /** Q filtered kClosestPairs calls (Register.cpp:210-236 findNeighbors). */
    void kClosestPairsFilteredBatch(float const * q3, float const * qn3, int32_t const * qlabel, intx Q, int32_t k, int32_t * idx, float * d2,
                                    float min_dot = 0.0f, double dist_bound = -1) const
    { detail::check(h, tgq_knn_filtered(h, q3, qn3, qlabel, Q, k, min_dot, dist_bound, idx, d2), "tgq_knn_filtered"); }
    void closestElementBatch(float const * q3, intx Q, int32_t * idx, float * d2, float * cp3, double dist_bound = -1,
                             bool exclude_self = false) const
    { detail::check(h, tgq_closest(h, q3, Q, dist_bound, exclude_self ? 1 : 0, idx, d2, cp3), "tgq_closest"); }

    /** BvhN::kClosestPairs (BvhN.hpp:1137-1146).  Like the reference the set is NOT cleared first: the k nearest are
     *  offered to it in ascending order through its own isInsertable / insert. */
    template <typename MetricT, typename BoundedNeighborPairSetT, typename CompatibilityFunctorT = UniversalCompatibility>
    intx kClosestPairs(VectorT const & query, BoundedNeighborPairSetT & k_closest_pairs, double dist_bound = -1,
                       CompatibilityFunctorT = CompatibilityFunctorT(), bool get_closest_points = false) const
    {
      int32_t k = (int32_t)k_closest_pairs.getCapacity();
      if (k > TGQ_MAX_K) throw Error("TheaGpu: kClosestPairs supports at most TGQ_MAX_K neighbours");
      if (k > 0)
      {
        float q[3]; detail::put3(q, query);
        std::vector<int32_t> idx((size_t)k); std::vector<float> d2((size_t)k);
        if (filter)
        {
          float fn[3]; detail::put3(fn, filter->n);
          detail::check(h, tgq_knn_filtered(h, q, filter->use_normals ? fn : nullptr, filter->use_labels ? &filter->label : nullptr, 1, k,
                                            filter->min_dot, dist_bound, idx.data(), d2.data()), "tgq_knn_filtered");
        }
        else detail::check(h, tgq_knn(h, q, 1, k, dist_bound, 0, idx.data(), d2.data()), "tgq_knn");
        for (int32_t j = 0; j < k && idx[(size_t)j] >= 0; ++j)
        {
          NeighborPair pair(0, idx[(size_t)j], (double)d2[(size_t)j]);
          if (get_closest_points) pair.setQueryPoint(query);
          if (k_closest_pairs.isInsertable(pair)) k_closest_pairs.insert(pair);
        }
      }
      return (intx)k_closest_pairs.size();
    }
    /** Q k-NN queries in one launch: idx / d2 are Q x k, ascending by (d2, idx), unused slots -1. */
    void kClosestPairsBatch(float const * q3, intx Q, int32_t k, int32_t * idx, float * d2, double dist_bound = -1,
                            bool exclude_self = false) const
    { detail::check(h, tgq_knn(h, q3, Q, k, dist_bound, exclude_self ? 1 : 0, idx, d2), "tgq_knn"); }

    /** Q ball range queries, BvhN::rangeQueryIndices<IntersectionTester>(Ball3(c_i, r_i), ...) (BvhN.hpp:1208-1220): count[Q] and
     *  the max_out (<= TGQ_MAX_K) lowest element indices in range per query, ascending, -1 padded. */
    void rangeQueryIndicesBatch(float const * c3, float const * radius, intx Q, int32_t max_out, int32_t * idx, int32_t * count,
                                bool exclude_self = false) const
    { detail::check(h, tgq_ball_range(h, c3, radius, Q, max_out, exclude_self ? 1 : 0, idx, count), "tgq_ball_range"); }
    /** The reference's scalar form for a ball: appends the indices of all elements in range to `result` (ascending here; the
     *  reference reports traversal order).  More than TGQ_MAX_K elements in range is an error of this adaptor, not a truncation. */
    template <typename IntersectionTesterT>
    void rangeQueryIndicesBall(VectorT const & center, Real radius, std::vector<intx> & result, bool discard_prior_results = true) const
    {
      if (discard_prior_results) result.clear();
      float c[3]; detail::put3(c, center);
      int32_t idx[TGQ_MAX_K]; int32_t count = 0;
      detail::check(h, tgq_ball_range(h, c, &radius, 1, TGQ_MAX_K, 0, idx, &count), "tgq_ball_range");
      if (count > TGQ_MAX_K) throw Error("TheaGpu: rangeQueryIndices found more than TGQ_MAX_K elements; use rangeQueryIndicesBatch for counts");
      for (int32_t j = 0; j < count; ++j) result.push_back(idx[j]);
    }

    /** Q processRange<IntersectionTester>(Ball3(c_i, r_i), functor) calls with the accumulating functors of AverageDistance /
     *  Curvature / LocalPca in one pass: out16 is Q x 16 doubles (layout: tgq_ball_stats in thea_gpu_query.h).  qn3 may be null. */
    void ballStatsBatch(float const * c3, float const * qn3, float const * radius, intx Q, double * out16) const
    { detail::check(h, tgq_ball_stats(h, c3, qn3, radius, Q, out16), "tgq_ball_stats"); }

    /** Q LocalDistanceHistogram::compute(c_i, histogram, EUCLIDEAN, r_i) calls: hist is Q x num_bins raw counts. */
    void ballHistogramBatch(float const * c3, float const * radius, intx Q, int32_t num_bins, uint32_t * hist) const
    { detail::check(h, tgq_ball_histogram(h, c3, radius, Q, num_bins, hist), "tgq_ball_histogram"); }

    /** Q SpinImage::compute(position_i, normal_i, num_radial_bins, num_height_bins, image) calls (SurfaceFeatures/Local/SpinImage.cpp:
     *  47-75): image is Q x num_radial_bins x num_height_bins raw counts over all samples; max_extent = 0.75f * scale there. */
    void spinImageBatch(float const * q3, float const * qn3, intx Q, int32_t num_radial_bins, int32_t num_height_bins, float max_extent,
                        uint32_t * image) const
    { detail::check(h, tgq_spin_image(h, q3, qn3, Q, num_radial_bins, num_height_bins, max_extent, image), "tgq_spin_image"); }

    /** Shortest paths on a sample graph given as sampleGraphBatch rows (the reference: ShortestPaths::dijkstraWithCallback on
     *  PointSet3's graph).  The structure only lends its device and stream. */
    void graphContractBatch(int32_t const * idx, float const * sep, intx P, int32_t max_degree, intx n_orig, int32_t * out_idx, float * out_sep,
                            int32_t * out_count = nullptr) const
    { detail::check(h, tgq_graph_contract(h, idx, sep, P, max_degree, n_orig, out_idx, out_sep, out_count), "tgq_graph_contract"); }
    void graphGeodesicBatch(int32_t const * idx, float const * sep, intx P, int32_t max_degree, int32_t const * seeds, intx Q, float max_distance,
                            float hist_range, int32_t num_bins, double * sum, int32_t * count, uint32_t * hist) const
    { detail::check(h, tgq_graph_geodesic(h, idx, sep, P, max_degree, seeds, Q, max_distance, hist_range, num_bins, sum, count, hist), "tgq_graph_geodesic"); }
    intx graphFurthestPointsBatch(int32_t const * idx, float const * sep, intx P, int32_t max_degree, intx n_desired, int32_t * selected) const
    { int64_t n = 0; detail::check(h, tgq_graph_fps(h, idx, sep, P, max_degree, n_desired, selected, &n), "tgq_graph_fps"); return (intx)n; }

    /** PointSet3::updateGraph's neighbour search for every sample (PointSet3.hpp:324-391,586-681). */
    void sampleGraphBatch(int32_t max_degree, float init_radius, int32_t * idx, float * sep, int32_t * count = nullptr, int32_t rounds = 3,
                          float min_dot = -0.5f, int32_t min_degree = -1) const
    {
      if (min_degree < 0) { min_degree = (int32_t)(0.25 * max_degree); if (min_degree < 4) min_degree = 4; if (min_degree > max_degree) min_degree = max_degree; }
      detail::check(h, tgq_ball_graph(h, max_degree, init_radius, rounds, min_dot, min_degree, idx, sep, count), "tgq_ball_graph");
    }
    /** PointSet3::updateGraph(surface) (PointSet3.hpp:324-391,639-648): the same search with the reachability ray test against a
     *  triangle structure on the same device. */
    void sampleGraphBatch(GpuMeshBvhT<VectorT> const & surface, int32_t max_degree, float init_radius, int32_t * idx, float * sep,
                          int32_t * count = nullptr, int32_t rounds = 3, float min_dot = -0.5f, int32_t min_degree = -1) const
    {
      if (min_degree < 0) { min_degree = (int32_t)(0.25 * max_degree); if (min_degree < 4) min_degree = 4; if (min_degree > max_degree) min_degree = max_degree; }
      detail::check(h, tgq_ball_graph_surface(h, surface.handle(), max_degree, init_radius, rounds, min_dot, min_degree, idx, sep, count),
                    "tgq_ball_graph_surface");
    }
    /** PointSet3.hpp:363-376 with caller-fixed seed samples. */
    float averageSeparation(int32_t const * seeds, int32_t n) const
    { float out = 0; detail::check(h, tgq_avg_separation(h, seeds, n, &out), "tgq_avg_separation"); return out; }

  private:
    template <typename RangeT> auto rangeIndices(RangeT const & range, std::vector<int32_t> & idx) const -> decltype(range.getRadius(), void())
    {
      float c[3], r = (float)range.getRadius(); detail::put3(c, range.getCenter());
      std::vector<int64_t> off; rangeQueryBatch(0, c, &r, 1, off, idx);
    }
    template <typename RangeT> auto rangeIndices(RangeT const & range, std::vector<int32_t> & idx) const -> decltype(range.getLow(), void())
    {
      float lo[3], hi[3]; detail::put3(lo, range.getLow()); detail::put3(hi, range.getHigh());
      std::vector<int64_t> off; rangeQueryBatch(1, lo, hi, 1, off, idx);
    }

    NNFilter const * filter;
    std::vector<VectorT> elems;
};

typedef GpuMeshBvhT<Vec3> GpuMeshBvh;
typedef GpuPointBvhT<Vec3> GpuPointBvh;

}  // namespace TheaGpu

#endif
