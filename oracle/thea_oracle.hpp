// thea_oracle.hpp -- CPU restatement of the Thea BvhN spatial-query path.
//
// TEST INFRASTRUCTURE ONLY.  This is the *checker* for the CUDA engine in thea_b200/csrc and the
// timed CPU baseline of bench.py.  Nothing in the product path may include, link or call it.
//
// PARITY PINNED TO THE REFERENCE'S CODE, NOT TO EIGEN: the reference (sidch/Thea @ 30b86b95) has no golden vectors or
// asserting tests on this path (Test/TestBvh.cpp:26-40 only prints) and its hard dependency Eigen >= 3.4 is absent
// (MatVec.hpp:27-28).  oracle/_ref/libthea_ref.so is the reference's own, unmodified BvhN / Triangle3 / AxisAlignedBoxN /
// MetricL2 / Icp3 / ShapeDiameter / PointSet3 sources compiled against the Eigen stand-in of oracle/eigen_shim/ (recipe:
// oracle/Makefile, target `ref`), and tests/test_ref_pin.py demands bit equality between it and this restatement on 10^7
// point-triangle pairs, 10^7 ray-triangle pairs, and every structure query.  What remains ASSUMED is only what the stand-in
// assumes about Eigen (header of oracle/eigen_shim/Eigen/Core): Eigen 3.4's fixed-size-3 evaluation order (sum of three
// terms = t0 + (t1 + t2); normalized() = v / sqrt(|v|^2) by true division; maxCoeff(&i) returns the first maximum), and
// JacobiSVD for ICP (tolerance 1e-5 there).  The restatement follows the reference source line by line (citations on
// every function, paths relative to Code/Source/).  Build with
//   g++ -std=c++17 -O2 -ffp-contract=off      (the reference: -O2, no -march, no fast-math,
//                                               Code/Build/Common/BuildSetup.cmake:95-111)
// so that every fp32 operation is a single IEEE-754 round-to-nearest op (no FMA contraction).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace oracle {

// ------------------------------------------------------------------------------------------------
// Vec3 with Eigen's evaluation order (MatVec.hpp:56: Matrix<T,3,1,DontAlign> => scalar, unrolled).
// ------------------------------------------------------------------------------------------------
struct V3
{
  float x, y, z;
  float & operator[](int i) { return (&x)[i]; }
  float const & operator[](int i) const { return (&x)[i]; }
};

inline V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
inline V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline V3 vmin(V3 a, V3 b) { return mk(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); }
inline V3 vmax(V3 a, V3 b) { return mk(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); }

// Eigen redux_novec_unroller<0,3>: func(x0, func(x1, x2)).
inline float sum3(float a, float b, float c) { return a + (b + c); }
inline float dot(V3 a, V3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float sqnorm(V3 a) { return dot(a, a); }
inline float norm(V3 a) { return std::sqrt(sqnorm(a)); }
inline V3 cross(V3 l, V3 r)
{ return mk(l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x); }
// Eigen MatrixBase::normalized(): z = squaredNorm; z > 0 ? n / sqrt(z) : n.
inline V3 normalized(V3 a) { float z = sqnorm(a); return z > 0.0f ? a / std::sqrt(z) : a; }
// Eigen 3.4 MatrixBase::stableNormalized() (Dot.h): w = cwiseAbs().maxCoeff(); z = (v / w).squaredNorm(); z > 0 ? v / (sqrt(z) * w) : v.
inline V3 stable_normalized(V3 a)
{
  float w = std::max(std::max(std::fabs(a.x), std::fabs(a.y)), std::fabs(a.z));
  float z = sqnorm(a / w);
  return z > 0.0f ? a / (std::sqrt(z) * w) : a;
}
// MatrixUtil.hpp:108-115 maxAbsAxis / :84-91 maxAxis -- first maximum wins.
inline int max_axis(V3 v) { int m = 0; if (v[1] > v[m]) m = 1; if (v[2] > v[m]) m = 2; return m; }
inline int max_abs_axis(V3 v) { return max_axis(mk(std::fabs(v.x), std::fabs(v.y), std::fabs(v.z))); }

// ------------------------------------------------------------------------------------------------
// Triangle cache: Triangle3.hpp:143-155 (update), Hyperplane3.hpp:34-41 (fromThreePoints),
// Triangle3.hpp:215-219 (getBounds).
// ------------------------------------------------------------------------------------------------
struct Tri
{
  V3 v0, v1, v2;
  V3 n;        // unit plane normal (zero vector for zero-area triangles)
  float d;     // n . v0
  int axis;    // primary axis
  V3 centroid;
  V3 e01, e02;
  float area;
  V3 lo, hi;   // element AABB
};

inline Tri make_tri(V3 v0, V3 v1, V3 v2)
{
  Tri t; t.v0 = v0; t.v1 = v1; t.v2 = v2;
  t.n = normalized(cross(v1 - v0, v2 - v0));
  t.d = dot(t.n, v0);
  t.axis = max_abs_axis(t.n);
  t.centroid = ((v0 + v1) + v2) / 3.0f;
  t.e01 = v1 - v0;
  t.e02 = v2 - v0;
  t.area = 0.5f * std::fabs(dot(t.n, cross(t.e01, t.e02)));
  t.lo = vmin(v0, vmin(v1, v2));
  t.hi = vmax(v0, vmax(v1, v2));
  return t;
}

// Triangle3.hpp:1414-1462 isPointInsideTriangle.
inline bool point_inside_tri(V3 const & v0, V3 const & v1, V3 const & v2, int axis, V3 const & p)
{
  int i, j;
  switch (axis) { case 1: i = 2; j = 0; break; case 2: i = 0; j = 1; break; default: i = 1; j = 2; }
#define ORACLE_A2(d, e, f) (((e)[i] - (d)[i]) * ((f)[j] - (d)[j]) - ((f)[i] - (d)[i]) * ((e)[j] - (d)[j]))
  float area = ORACLE_A2(v0, v1, v2);
  if (area == 0) return (v0 == p);
  float inv = 1.0f / area;
  float b0 = ORACLE_A2(p, v1, v2) * inv;
  if (b0 < 0 || b0 > 1) return false;
  float b1 = ORACLE_A2(v0, p, v2) * inv;
  if (b1 < 0 || b1 > 1) return false;
  float b2 = 1.0f - b0 - b1;
#undef ORACLE_A2
  return b2 >= 0 && b2 <= 1;
}

// Triangle3.hpp:1310-1338 closestPointOnLineSegment.
inline V3 closest_on_segment(V3 a, V3 b, V3 dir, float len, V3 p)
{
  V3 c = p - a;
  float t = dot(dir, c);
  if (t <= 0) return a;
  else if (t >= len) return b;
  else return a + dir * t;
}

// Triangle3.hpp:1392-1412 + 1340-1390 closestPointOnTrianglePerimeter.
inline V3 closest_on_perimeter(V3 v0, V3 v1, V3 v2, V3 p)
{
  V3 v[3] = {v0, v1, v2};
  V3 dir[3] = {v1 - v0, v2 - v1, v0 - v2};
  float len[3];
  for (int k = 0; k < 3; ++k) { len[k] = norm(dir[k]); dir[k] = dir[k] / len[k]; }
  static int const nxt[3] = {1, 2, 0};
  V3 r[3]; float d[3];
  for (int k = 0; k < 3; ++k)
  {
    r[k] = closest_on_segment(v[k], v[nxt[k]], dir[k], len[k], p);
    d[k] = sqnorm(r[k] - p);
  }
  int e;
  if (d[0] < d[1]) e = (d[0] < d[2]) ? 0 : 2;
  else             e = (d[1] < d[2]) ? 1 : 2;
  return r[e];
}

// MetricL2.hpp:210-215 closestPoints(tri, point) -> Triangle3.hpp:294-304 closestPoint ->
// HyperplaneN.hpp:145-160.  Returns fp32 squared distance; cp = closest point on the triangle.
inline float tri_point_sqdist(Tri const & t, V3 p, V3 & cp)
{
  float sd = dot(p, t.n) - t.d;
  V3 proj = p - sd * t.n;
  if (point_inside_tri(t.v0, t.v1, t.v2, t.axis, proj)) cp = proj;
  else cp = closest_on_perimeter(t.v0, t.v1, t.v2, p);
  return sqnorm(cp - p);
}

// MetricL2.hpp:187-192 closestPoints(point a = element, point b = query).
inline float point_point_sqdist(V3 elem, V3 q) { return sqnorm(elem - q); }

// Triangle3.hpp:1464-1524 rayTriangleIntersectionTime (Eberly / Wild Magic 5.3).
inline float ray_tri_time(V3 o, V3 D, V3 v0, V3 e01, V3 e02)
{
  float const EPS = std::numeric_limits<float>::epsilon();
  V3 diff = o - v0;
  V3 N = cross(e01, e02);
  float DdN = dot(D, N);
  int sign;
  if (DdN > EPS) sign = 1;
  else if (DdN < -EPS) { sign = -1; DdN = -DdN; }
  else return -1;
  float DdQxE2 = (float)sign * dot(D, cross(diff, e02));
  if (DdQxE2 >= 0)
  {
    float DdE1xQ = (float)sign * dot(D, cross(e01, diff));
    if (DdE1xQ >= 0)
    {
      if (DdQxE2 + DdE1xQ <= DdN)
      {
        float QdN = (float)(-sign) * dot(diff, N);
        if (QdN >= 0) return QdN / DdN;
      }
    }
  }
  return -1;
}

// BvhN.hpp:178-181 distanceLessThan.
template <typename U, typename W> inline bool dist_less(U lhs, W rhs) { return lhs >= 0 && (rhs < 0 || lhs < rhs); }

// Math.hpp:194-223 fastRsq / fastSqrt (float32).
inline float fast_rsq(float x)
{
  float x2 = x * 0.5f;
  int32_t i; std::memcpy(&i, &x, 4);
  i = 0x5f3759df - (i >> 1);
  float y; std::memcpy(&y, &i, 4);
  y = y * (1.5f - (x2 * y * y));
  return y;
}
inline float fast_sqrt(float x) { return 1.0f / fast_rsq(x); }

// Math.hpp:124-154 eps(a,b), fuzzyEq / fuzzyNe.
inline bool fuzzy_ne0(float a)
{
  float larger = std::max(std::fabs(a), 0.0f);
  float tol = larger > 1 ? larger * std::numeric_limits<float>::epsilon() : std::numeric_limits<float>::epsilon();
  return !((a == 0.0f) || std::fabs(a - 0.0f) <= tol);
}

inline bool fuzzy_eq0(float a) { return !fuzzy_ne0(a) && !std::isinf(a); }      // eps(a, b) keeps the plain epsilon for infinite operands (Math.hpp:137-138)

// ------------------------------------------------------------------------------------------------
// Line-segment query objects.  LineSegmentN stores (point, direction = end1 - end0),
// LineSegmentN.hpp:56-58; its far end is point + direction.
// ------------------------------------------------------------------------------------------------
struct Seg { V3 point, direction; };
inline Seg make_seg(V3 end0, V3 end1) { Seg s; s.point = end0; s.direction = end1 - end0; return s; }
inline float clamp01(float x) { return x < 0 ? 0.0f : (x > 1 ? 1.0f : x); }       // Math::clamp, Math.hpp:247-252

// LineSegmentN::closestPoint, LineSegmentN.hpp:106-145 (from G3D::LineSegment).
inline V3 seg_closest_point(Seg const & g, V3 p)
{
  float d2 = sqnorm(g.direction);
  if (fuzzy_eq0(d2)) return g.point;
  V3 v = p - g.point;
  float t = dot(g.direction, v);
  if (t >= 0 && t <= d2) return g.point + (t / d2) * g.direction;
  float d0_squared = sqnorm(v);
  float d1_squared = sqnorm(v - g.direction);
  if (d0_squared < d1_squared) return g.point;
  return g.point + g.direction;
}
// MetricL2Impl<LineSegmentN, point>::closestPoints, MetricL2.hpp:210-215: cpb = p; cpa = a.closestPoint(cpb); |cpa - cpb|^2.
inline float seg_point_sqdist(Seg const & g, V3 p, V3 & on_seg) { on_seg = seg_closest_point(g, p); return sqnorm(on_seg - p); }

// Internal::closestPtSegmentSegment, LineSegmentN.hpp:262-357.
inline float seg_seg_sqdist(V3 p1, V3 q1, bool is_line1, V3 p2, V3 q2, bool is_line2, float & s, float & t, V3 & c1, V3 & c2)
{
  V3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  float a = sqnorm(d1), e = sqnorm(d2), f = dot(d2, r);
  if (fuzzy_eq0(a))
  {
    if (fuzzy_eq0(e)) { s = t = 0; c1 = p1; c2 = p2; return dot(c1 - c2, c1 - c2); }
    s = 0; t = f / e;
    if (!is_line2) t = clamp01(t);
  }
  else
  {
    float c = dot(d1, r);
    if (fuzzy_eq0(e))
    {
      t = 0; s = -c / a;
      if (!is_line1) s = clamp01(s);
    }
    else
    {
      float b = dot(d1, d2);
      float denom = a * e - b * b;
      if (denom != 0) { s = (b * f - c * e) / denom; if (!is_line1) s = clamp01(s); }
      else s = 0;
      t = (b * s + f) / e;
      if (!is_line2)
      {
        if (t < 0) { t = 0; s = -c / a; if (!is_line1) s = clamp01(s); }
        else if (t > 1) { t = 1; s = (b - c) / a; if (!is_line1) s = clamp01(s); }
      }
    }
  }
  c1 = p1 + s * d1; c2 = p2 + t * d2;
  return sqnorm(c1 - c2);
}

// TriangleN<3>::squaredDistance(LineSegmentN, this_pt, seg_pt), Triangle3.hpp:495-509: the segment as a ray over t in [0, 1]
// (rayIntersectionTime on the cached edges, :536-540), else squaredDistanceToPerimeter (:559-577): each edge as
// LineSegmentN(v_i, v_j) against the query via LineSegmentN::squaredDistance(LineSegmentN) (LineSegmentN.hpp:153-164).
inline float tri_seg_sqdist(Tri const & tri, Seg const & g, V3 & on_tri, V3 & on_seg)
{
  float t = ray_tri_time(g.point, g.direction, tri.v0, tri.e01, tri.e02);
  if (t >= 0 && t <= 1) { V3 c = g.point + t * g.direction; on_tri = c; on_seg = c; return 0; }
  V3 const vv[3] = {tri.v0, tri.v1, tri.v2};
  float d2 = -1;
  for (int i = 0; i < 3; ++i)
  {
    Seg edge = make_seg(vv[i], vv[(i + 1) % 3]);
    V3 c1, c2; float s, u;
    float edge_d2 = seg_seg_sqdist(edge.point, edge.point + edge.direction, false, g.point, g.point + g.direction, false, s, u, c1, c2);
    if (d2 < 0 || edge_d2 < d2) { d2 = edge_d2; on_tri = c1; on_seg = c2; }
  }
  return d2;
}

// ------------------------------------------------------------------------------------------------
// Triangle query objects: TriangleN<3>::squaredDistance(TriangleN<3>), Triangle3.hpp:314-392, on top of
// Moeller's tri_tri_intersect_with_isectline as the reference carries it (Triangle3.hpp:597-1300).  The macros there are plain C on
// float[3]: THEA_TRI3_DOT sums left to right, (x + y) + z -- not Eigen's order.
// ------------------------------------------------------------------------------------------------
namespace moeller {
typedef float const * P3;
inline float DOT(P3 a, P3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void SUB(float * d, P3 a, P3 b) { d[0] = a[0] - b[0]; d[1] = a[1] - b[1]; d[2] = a[2] - b[2]; }
inline void CROSS(float * d, P3 a, P3 b) { d[0] = a[1] * b[2] - a[2] * b[1]; d[1] = a[2] * b[0] - a[0] * b[2]; d[2] = a[0] * b[1] - a[1] * b[0]; }

// THEA_TRI3_EDGE_EDGE_TEST (:649-669), THEA_TRI3_EDGE_AGAINST_TRI_EDGES (:671-683)
inline bool edge_against_tri_edges(P3 V0, P3 V1, P3 U0, P3 U1, P3 U2, int i0, int i1)
{
  float Ax = V1[i0] - V0[i0], Ay = V1[i1] - V0[i1];
  P3 ua[3] = {U0, U1, U2}, ub[3] = {U1, U2, U0};
  for (int k = 0; k < 3; ++k)
  {
    float Bx = ua[k][i0] - ub[k][i0], By = ua[k][i1] - ub[k][i1];
    float Cx = V0[i0] - ua[k][i0], Cy = V0[i1] - ua[k][i1];
    float f = Ay * Bx - Ax * By;
    float d = By * Cx - Bx * Cy;
    if ((f > 0 && d >= 0 && d <= f) || (f < 0 && d <= 0 && d >= f))
    {
      float e = Ax * Cy - Ay * Cx;
      if (f > 0) { if (e >= 0 && e <= f) return true; }
      else       { if (e <= 0 && e >= f) return true; }
    }
  }
  return false;
}
// THEA_TRI3_POINT_IN_TRI (:685-709)
inline bool point_in_tri(P3 V0, P3 U0, P3 U1, P3 U2, int i0, int i1)
{
  float a, b, c, d0, d1, d2;
  a = U1[i1] - U0[i1]; b = -(U1[i0] - U0[i0]); c = -a * U0[i0] - b * U0[i1]; d0 = a * V0[i0] + b * V0[i1] + c;
  a = U2[i1] - U1[i1]; b = -(U2[i0] - U1[i0]); c = -a * U1[i0] - b * U1[i1]; d1 = a * V0[i0] + b * V0[i1] + c;
  a = U0[i1] - U2[i1]; b = -(U0[i0] - U2[i0]); c = -a * U2[i0] - b * U2[i1]; d2 = a * V0[i0] + b * V0[i1] + c;
  if (d0 * d1 > 0) { if (d0 * d2 > 0) return true; }
  return false;
}
// coplanar_tri_tri (:711-760)
inline int coplanar_tri_tri(P3 N, P3 V0, P3 V1, P3 V2, P3 U0, P3 U1, P3 U2)
{
  float A[3] = {std::fabs(N[0]), std::fabs(N[1]), std::fabs(N[2])};
  int i0, i1;
  if (A[0] > A[1]) { if (A[0] > A[2]) { i0 = 1; i1 = 2; } else { i0 = 0; i1 = 1; } }
  else             { if (A[2] > A[1]) { i0 = 0; i1 = 1; } else { i0 = 0; i1 = 2; } }
  if (edge_against_tri_edges(V0, V1, U0, U1, U2, i0, i1)) return 1;
  if (edge_against_tri_edges(V1, V2, U0, U1, U2, i0, i1)) return 1;
  if (edge_against_tri_edges(V2, V0, U0, U1, U2, i0, i1)) return 1;
  if (point_in_tri(V0, U0, U1, U2, i0, i1)) return 1;
  if (point_in_tri(U0, V0, V1, V2, i0, i1)) return 1;
  return 0;
}
// isect2 (:1024-1039)
inline void isect2(P3 VTX0, P3 VTX1, P3 VTX2, float VV0, float VV1, float VV2, float D0, float D1, float D2, float * isect0, float * isect1,
                   float * isectpoint0, float * isectpoint1)
{
  float tmp = D0 / (D0 - D1), diff[3];
  *isect0 = VV0 + (VV1 - VV0) * tmp;
  SUB(diff, VTX1, VTX0);
  for (int k = 0; k < 3; ++k) diff[k] = tmp * diff[k];
  for (int k = 0; k < 3; ++k) isectpoint0[k] = diff[k] + VTX0[k];
  tmp = D0 / (D0 - D2);
  *isect1 = VV0 + (VV2 - VV0) * tmp;
  SUB(diff, VTX2, VTX0);
  for (int k = 0; k < 3; ++k) diff[k] = tmp * diff[k];
  for (int k = 0; k < 3; ++k) isectpoint1[k] = VTX0[k] + diff[k];
}
// compute_intervals_isectline (:1041-1077)
inline int compute_intervals_isectline(P3 VERT0, P3 VERT1, P3 VERT2, float VV0, float VV1, float VV2, float D0, float D1, float D2, float D0D1, float D0D2,
                                       float * isect0, float * isect1, float * isectpoint0, float * isectpoint1)
{
  if (D0D1 > 0)                 isect2(VERT2, VERT0, VERT1, VV2, VV0, VV1, D2, D0, D1, isect0, isect1, isectpoint0, isectpoint1);
  else if (D0D2 > 0)            isect2(VERT1, VERT0, VERT2, VV1, VV0, VV2, D1, D0, D2, isect0, isect1, isectpoint0, isectpoint1);
  else if (D1 * D2 > 0 || D0 != 0) isect2(VERT0, VERT1, VERT2, VV0, VV1, VV2, D0, D1, D2, isect0, isect1, isectpoint0, isectpoint1);
  else if (D1 != 0)             isect2(VERT1, VERT0, VERT2, VV1, VV0, VV2, D1, D0, D2, isect0, isect1, isectpoint0, isectpoint1);
  else if (D2 != 0)             isect2(VERT2, VERT0, VERT1, VV2, VV0, VV1, D2, D0, D1, isect0, isect1, isectpoint0, isectpoint1);
  else return 1;                // coplanar
  return 0;
}
// tri_tri_intersect_with_isectline (:1115-1300)
inline int tri_tri_intersect_with_isectline(P3 V0, P3 V1, P3 V2, P3 U0, P3 U1, P3 U2, int * coplanar, float * isectpt1, float * isectpt2)
{
  float const EPSILON = std::numeric_limits<float>::epsilon();
  float E1[3], E2[3], N1[3], N2[3], D[3];
  float isect1[2] = {0, 0}, isect2_[2] = {0, 0};
  float isectpointA1[3] = {0, 0, 0}, isectpointA2[3] = {0, 0, 0}, isectpointB1[3] = {0, 0, 0}, isectpointB2[3] = {0, 0, 0};
  SUB(E1, V1, V0); SUB(E2, V2, V0); CROSS(N1, E1, E2);
  float d1 = -DOT(N1, V0);
  float du0 = DOT(N1, U0) + d1, du1 = DOT(N1, U1) + d1, du2 = DOT(N1, U2) + d1;
  if (std::fabs(du0) < EPSILON) du0 = 0;
  if (std::fabs(du1) < EPSILON) du1 = 0;
  if (std::fabs(du2) < EPSILON) du2 = 0;
  float du0du1 = du0 * du1, du0du2 = du0 * du2;
  if (du0du1 > 0 && du0du2 > 0) return 0;
  SUB(E1, U1, U0); SUB(E2, U2, U0); CROSS(N2, E1, E2);
  float d2 = -DOT(N2, U0);
  float dv0 = DOT(N2, V0) + d2, dv1 = DOT(N2, V1) + d2, dv2 = DOT(N2, V2) + d2;
  if (std::fabs(dv0) < EPSILON) dv0 = 0;
  if (std::fabs(dv1) < EPSILON) dv1 = 0;
  if (std::fabs(dv2) < EPSILON) dv2 = 0;
  float dv0dv1 = dv0 * dv1, dv0dv2 = dv0 * dv2;
  if (dv0dv1 > 0 && dv0dv2 > 0) return 0;
  CROSS(D, N1, N2);
  float max = std::fabs(D[0]); int index = 0;
  float b = std::fabs(D[1]), c = std::fabs(D[2]);
  if (b > max) { max = b; index = 1; }
  if (c > max) { max = c; index = 2; }
  float vp0 = V0[index], vp1 = V1[index], vp2 = V2[index], up0 = U0[index], up1 = U1[index], up2 = U2[index];
  *coplanar = compute_intervals_isectline(V0, V1, V2, vp0, vp1, vp2, dv0, dv1, dv2, dv0dv1, dv0dv2, &isect1[0], &isect1[1], isectpointA1, isectpointA2);
  if (*coplanar) return coplanar_tri_tri(N1, V0, V1, V2, U0, U1, U2);
  compute_intervals_isectline(U0, U1, U2, up0, up1, up2, du0, du1, du2, du0du1, du0du2, &isect2_[0], &isect2_[1], isectpointB1, isectpointB2);
  int smallest1, smallest2;
  if (isect1[0] > isect1[1]) { std::swap(isect1[0], isect1[1]); smallest1 = 1; } else smallest1 = 0;       // THEA_TRI3_SORT2
  if (isect2_[0] > isect2_[1]) { std::swap(isect2_[0], isect2_[1]); smallest2 = 1; } else smallest2 = 0;
  if (isect1[1] < isect2_[0] || isect2_[1] < isect1[0]) return 0;
  float const * p1; float const * p2;
  if (isect2_[0] < isect1[0])
  {
    p1 = smallest1 == 0 ? isectpointA1 : isectpointA2;
    if (isect2_[1] < isect1[1]) p2 = smallest2 == 0 ? isectpointB2 : isectpointB1;
    else                        p2 = smallest1 == 0 ? isectpointA2 : isectpointA1;
  }
  else
  {
    p1 = smallest2 == 0 ? isectpointB1 : isectpointB2;
    if (isect2_[1] > isect1[1]) p2 = smallest1 == 0 ? isectpointA2 : isectpointA1;
    else                        p2 = smallest2 == 0 ? isectpointB2 : isectpointB1;
  }
  for (int k = 0; k < 3; ++k) { isectpt1[k] = p1[k]; isectpt2[k] = p2[k]; }
  return 1;
}
}  // namespace moeller

// TriangleN<3>::squaredDistance(other, this_pt, other_pt), Triangle3.hpp:314-392 (Ericson): `a` is the triangle the member is called on.
inline float tri_tri_sqdist(Tri const & a, Tri const & b, V3 & on_a, V3 & on_b)
{
  V3 c1 = mk(0, 0, 0), c2 = mk(0, 0, 0);
  float min_sqdist = std::numeric_limits<float>::max();
  int i_coplanar = 0; V3 isectpt1 = mk(0, 0, 0), isectpt2 = mk(0, 0, 0);
  int isec = moeller::tri_tri_intersect_with_isectline(&a.v0.x, &a.v1.x, &a.v2.x, &b.v0.x, &b.v1.x, &b.v2.x, &i_coplanar, &isectpt1.x, &isectpt2.x);
  if (isec)
  {
    if (i_coplanar) c1 = c2 = 0.5f * (a.centroid + b.centroid);
    else
    {
      Seg seg = make_seg(isectpt1, isectpt2);                                               // LineSegmentN(isectpt1, isectpt2)
      c1 = c2 = 0.5f * ((seg.point + 0.0f * seg.direction) + (seg.point + 1.0f * seg.direction));      // getPoint(0) + getPoint(1)
    }
    min_sqdist = 0;
  }
  else
  {
    V3 const av[3] = {a.v0, a.v1, a.v2}, bv[3] = {b.v0, b.v1, b.v2};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
      {
        V3 p, q; float s, t;
        float d2 = seg_seg_sqdist(av[i], av[(i + 1) % 3], false, bv[j], bv[(j + 1) % 3], false, s, t, p, q);
        if (d2 < min_sqdist) { min_sqdist = d2; c1 = p; c2 = q; }
      }
    for (int i = 0; i < 3; ++i)      // vertex of b projecting inside a: getPlane().closestPoint(q) = q - signedDistance(q) * normal
    {
      V3 q = bv[i];
      V3 p = q - (dot(q, a.n) - a.d) * a.n;
      if (point_inside_tri(a.v0, a.v1, a.v2, a.axis, p))
      {
        float d2 = sqnorm(p - q);
        if (d2 < min_sqdist) { min_sqdist = d2; c1 = p; c2 = q; }
      }
    }
    for (int i = 0; i < 3; ++i)      // vertex of a projecting inside b
    {
      V3 p = av[i];
      V3 q = p - (dot(p, b.n) - b.d) * b.n;
      if (point_inside_tri(b.v0, b.v1, b.v2, b.axis, q))
      {
        float d2 = sqnorm(p - q);
        if (d2 < min_sqdist) { min_sqdist = d2; c1 = p; c2 = q; }
      }
    }
  }
  on_a = c1; on_b = c2;
  return min_sqdist;
}

// ------------------------------------------------------------------------------------------------
// AABB: AxisAlignedBoxN.hpp:237-271 (merge), 274 (getCenter), 316-324 (scaleCentered),
// 429-456 (squaredDistance), 576-658 (Woo ray test).
// ------------------------------------------------------------------------------------------------
struct Box
{
  V3 lo, hi; bool null;
  Box() : lo(mk(0, 0, 0)), hi(mk(0, 0, 0)), null(true) {}
  void merge(V3 l, V3 h)
  { if (null) { lo = l; hi = h; null = false; } else { lo = vmin(lo, l); hi = vmax(hi, h); } }
  V3 center() const { return 0.5f * (hi + lo); }
  V3 extent() const { return hi - lo; }
  Box padded(float s) const
  {
    Box b = *this;
    if (!null) { V3 c = center(); b.lo = s * (lo - c) + c; b.hi = s * (hi - c) + c; }
    return b;
  }
  // squaredDistance(point) :448-456 ; identical arithmetic to box-box with a degenerate query box :429-442
  // except for the cwiseMax(0) which is a no-op on a point query when lo <= hi.
  float sqdist_point_as_box(V3 p) const
  {
    if (null) return 0;
    V3 vmx = vmin(p, hi), vmn = vmax(p, lo);
    V3 g = vmax(vmn - vmx, mk(0, 0, 0));
    return sqnorm(g);
  }
  // squaredDistance(box) :429-442: vmax = hi.cwiseMin(other.hi), vmin = lo.cwiseMax(other.lo), |max(vmin - vmax, 0)|^2
  float sqdist_box(Box const & o) const
  {
    if (null || o.null) return 0;
    V3 vmx = vmin(hi, o.hi), vmn = vmax(lo, o.lo);
    V3 g = vmax(vmn - vmx, mk(0, 0, 0));
    return sqnorm(g);
  }
  float sqdist_point(V3 p) const
  {
    if (null) return 0;
    V3 vmx = vmin(p, hi), vmn = vmax(p, lo);
    return sqnorm(vmn - vmx);
  }
  // :592-648 rayIntersection(...).getTime()
  float ray_time(V3 o, V3 dir, float max_time) const
  {
    if (null) return -1;
    else if (max_time >= 0)
    {
      float ray_sqlen = sqnorm(dir);
      float origin_sqdist = sqdist_point(o);
      if (origin_sqdist > max_time * max_time * ray_sqlen) return -1;
    }
    V3 max_t = mk(-1, -1, -1);
    bool inside = true;
    for (int i = 0; i < 3; ++i)
    {
      if (o[i] < lo[i]) { inside = false; if (fuzzy_ne0(dir[i])) max_t[i] = (lo[i] - o[i]) / dir[i]; }
      else if (o[i] > hi[i]) { inside = false; if (fuzzy_ne0(dir[i])) max_t[i] = (hi[i] - o[i]) / dir[i]; }
    }
    if (inside) return 0;
    int w = 0;
    for (int i = 1; i < 3; ++i) if (max_t[i] > max_t[w]) w = i;
    if (max_t[w] < -0.5f) return -1;
    for (int i = 0; i < 3; ++i)
      if (i != w)
      {
        float loc = o[i] + max_t[w] * dir[i];
        if (loc < lo[i] || loc > hi[i]) return -1;
      }
    return max_t[w];
  }
  bool ray_hits(V3 o, V3 dir, float max_time) const { return ray_time(o, dir, max_time) > -0.5f; }
  // BallN.hpp:92-100 contains(AABB): all 8 corners inside the ball.
  bool inside_ball(V3 c, float r2) const
  {
    for (int k = 0; k < 8; ++k)
    {
      V3 v = mk((k & 1) ? hi.x : lo.x, (k & 2) ? hi.y : lo.y, (k & 4) ? hi.z : lo.z);
      if (!(sqnorm(v - c) <= r2)) return false;
    }
    return true;
  }
};

// ------------------------------------------------------------------------------------------------
// Results
// ------------------------------------------------------------------------------------------------
struct Closest { int64_t idx; float d2; V3 cp; };
struct RayHit { int64_t idx; float t; };

// Candidate acceptance, canonical rule (north_star: ties broken by lowest element index).
inline bool better_canon(float d2, int64_t idx, float best, int64_t best_idx)
{ return d2 < best || (d2 == best && idx < best_idx); }

// ------------------------------------------------------------------------------------------------
// Kd-tree of BvhN (Method::KDTREE, the default): BvhN.hpp:1365-1440 (initTree), 1450-1525 (createKdTree),
// 1320-1335 (ObjectLess), Math.cpp:22-36 (treeDepth).  Elements are triangles (elem_kind 0) or points (1).
// ------------------------------------------------------------------------------------------------
// Element filter of the reference's filtered queries: BvhN::pushFilter / elementPassesFilters (BvhN.hpp:1005-1032,1674-1683)
// with the NNFilter of Tools/Register/Register.cpp:92-101 -- same label and n . n_e >= min_dot (Register uses 0).  Installed
// per worker thread around a query, exactly like the reference pushes a filter around each closestElement / kClosestPairs call.
struct ElemFilter { bool use_normal; V3 n; float min_dot; bool use_label; int32_t label; };
inline thread_local ElemFilter const * g_filter = nullptr;

struct KdNode
{
  Box bounds;             // stored padded x1.05 (BvhN.hpp:1437,1514,2458)
  int depth;
  int64_t begin, count;   // range in KdTree::order (each node keeps its whole index list, save_memory=false)
  int64_t child[2];       // -1 for leaves
};

struct KdTree
{
  std::vector<Tri> tris;          // triangle elements (copied, BvhN.hpp:722-724)
  std::vector<V3> pts;            // point elements
  std::vector<V3> nrm;            // optional per-point normals (SamplePoint3)
  std::vector<int32_t> labels;    // optional per-point labels (the `label` member of Register's samples)
  bool is_points = false;
  std::vector<int64_t> order;     // element indices; node ranges index into this
  std::vector<KdNode> nodes;      // nodes[0] = root
  int max_depth = 0;
  int max_leaf = 5;

  int64_t size() const { return is_points ? (int64_t)pts.size() : (int64_t)tris.size(); }
  // elementPassesFilters (BvhN.hpp:1674-1683): self exclusion (PointSet3::FilterSelf) and the thread's installed filter.
  bool rejects(int64_t e, int64_t skip) const
  {
    if (e == skip) return true;
    ElemFilter const * f = g_filter;
    if (!f) return false;
    if (f->use_label && !labels.empty() && labels[(size_t)e] != f->label) return true;
    if (f->use_normal && !nrm.empty() && dot(f->n, nrm[(size_t)e]) < f->min_dot) return true;
    return false;
  }
  V3 elem_center(int64_t i) const { return is_points ? pts[i] : tris[i].centroid; }
  void elem_bounds(int64_t i, V3 & l, V3 & h) const
  { if (is_points) { l = h = pts[i]; } else { l = tris[i].lo; h = tris[i].hi; } }

  static int tree_depth(int64_t n, int leaf, float mcf)
  {
    if (n <= 0) return 0;
    double lim = -std::log(mcf);   // Real max_child_fraction = 0.5f promoted to double
    int64_t d = (int64_t)std::ceil(std::log(n / (double)leaf) / lim);
    return d < 0 ? 0 : (int)d;
  }

  void build()
  {
    int64_t n = size();
    nodes.clear(); order.resize((size_t)n);
    if (n == 0) return;
    max_depth = tree_depth(n, max_leaf, 0.5f);
    nodes.reserve((size_t)(2 * n / max_leaf + 16));
    KdNode root; root.depth = 0; root.begin = 0; root.count = n; root.child[0] = root.child[1] = -1;
    for (int64_t i = 0; i < n; ++i)
    { order[(size_t)i] = i; V3 l, h; elem_bounds(i, l, h); root.bounds.merge(l, h); }
    nodes.push_back(root);
    split(0);
    nodes[0].bounds = nodes[0].bounds.padded(1.05f);
  }

  void split(int64_t ni)
  {
    if (nodes[(size_t)ni].depth >= max_depth || nodes[(size_t)ni].count <= max_leaf) return;
    int coord = max_axis(nodes[(size_t)ni].bounds.extent());
    int64_t begin = nodes[(size_t)ni].begin, cnt = nodes[(size_t)ni].count;
    int64_t mid = cnt / 2;
    KdTree const * self = this;
    std::nth_element(order.begin() + begin, order.begin() + begin + mid, order.begin() + begin + cnt,
                     [self, coord](int64_t a, int64_t b) { return self->elem_center(a)[coord] < self->elem_center(b)[coord]; });
    for (int c = 0; c < 2; ++c)
    {
      KdNode ch; ch.depth = nodes[(size_t)ni].depth + 1; ch.child[0] = ch.child[1] = -1;
      int64_t eb = (c == 0) ? 0 : cnt - mid, ee = (c == 0) ? cnt - mid : cnt;
      ch.begin = begin + eb; ch.count = ee - eb;
      for (int64_t k = ch.begin; k < ch.begin + ch.count; ++k)
      { V3 l, h; elem_bounds(order[(size_t)k], l, h); ch.bounds.merge(l, h); }
      int64_t ci = (int64_t)nodes.size();
      nodes.push_back(ch);
      nodes[(size_t)ni].child[c] = ci;
      split(ci);
      nodes[(size_t)ci].bounds = nodes[(size_t)ci].bounds.padded(1.05f);
    }
  }

  bool is_leaf(KdNode const & nd) const { return nd.child[0] < 0; }

  // ---- closest element, BvhN.hpp:1081-1114, 1866-1893, 2036-2089, functor :440-471 -------------
  // canonical == false: the reference's exact behaviour (first-found wins among equal distances).
  // canonical == true : same traversal but (d2, idx) lexicographic and non-strict node pruning.
  struct ClosestState { double best; int64_t idx; V3 cp; int64_t skip; };

  float elem_sqdist(int64_t e, V3 q, V3 & cp) const
  {
    if (is_points) { cp = pts[(size_t)e]; return point_point_sqdist(pts[(size_t)e], q); }
    return tri_point_sqdist(tris[(size_t)e], q, cp);
  }

  template <bool Canon> void closest_rec(int64_t ni, V3 q, ClosestState & s) const
  {
    KdNode const & nd = nodes[(size_t)ni];
    if (is_leaf(nd))
    {
      for (int64_t k = nd.begin; k < nd.begin + nd.count; ++k)
      {
        int64_t e = order[(size_t)k];
        if (rejects(e, s.skip)) continue;               // FilterSelf (PointSet3.hpp:576-583) / pushed filters
        V3 cp; double mad = (double)elem_sqdist(e, q, cp);
        bool take = Canon ? (s.best < 0 ? (mad == mad) : (mad < s.best || (mad == s.best && s.idx >= 0 && e < s.idx)))
                          : (s.best < 0 || mad < s.best);
        if (take) { s.best = mad; s.idx = e; s.cp = cp; }
      }
      return;
    }
    // ChildDistanceArray: sorted by lower bound, child 0 first on ties (BoundedSortedArray.hpp:243-265).
    double lb[2]; int64_t ch[2] = {nd.child[0], nd.child[1]};
    for (int c = 0; c < 2; ++c) lb[c] = (double)nodes[(size_t)ch[c]].bounds.sqdist_point_as_box(q);
    if (dist_less(lb[1], lb[0])) { std::swap(lb[0], lb[1]); std::swap(ch[0], ch[1]); }
    for (int c = 0; c < 2; ++c)
    {
      bool go = Canon ? (s.best < 0 || lb[c] <= s.best) : (lb[c] < 0 || s.best < 0 || lb[c] < s.best);
      if (go) closest_rec<Canon>(ch[c], q, s);
    }
  }

  // What BvhN::closestPair does with a non-negative distance bound (BvhN.hpp:1081-1114), found by running the reference's own code
  // (oracle/_ref): (1) the query is rejected outright when the (padded) root box is farther than the bound (:1093-1098); (2)
  // accelerationBound() hands the caller's bound straight back (:2369-2370), so closestPair takes the `accel_bound >= 0` branch and
  // WIDENS the bound by max(0.001 * |extent of the root box|, 100 eps) (:1101-1107) before ClosestPairFunctor squares it (:446-447).
  // An element up to that fudge beyond dist_bound is therefore still returned.  kClosestPairs has no such widening (:478).
  bool closest_bound(V3 q, double dist_bound, double & bound2) const
  {
    bound2 = -1;
    if (!(dist_bound >= 0)) return true;
    double lb = (double)nodes[0].bounds.sqdist_point_as_box(q);
    if (lb >= 0 && lb > dist_bound * dist_bound) return false;      // MetricL2::computeMonotoneApprox of the caller's bound
    float diag = norm(nodes[0].bounds.extent());
    double fudge = std::max((double)(0.001 * diag), (double)(100 * std::numeric_limits<float>::epsilon()));
    double eff = dist_bound + fudge;
    bound2 = eff * eff;
    return true;
  }

  template <bool Canon> Closest closest(V3 q, double dist_bound, int64_t skip = -1) const
  {
    Closest r; r.idx = -1; r.d2 = -1; r.cp = mk(0, 0, 0);
    if (nodes.empty()) return r;
    double bound2;
    if (!closest_bound(q, dist_bound, bound2)) return r;
    ClosestState s; s.best = bound2; s.idx = -1; s.cp = mk(0, 0, 0); s.skip = skip;
    closest_rec<Canon>(0, q, s);
    if (s.idx >= 0) { r.idx = s.idx; r.d2 = (float)s.best; r.cp = s.cp; }
    return r;
  }

  // ---- closest element to a LineSegment3 query object: the same closestPair / processNeighbors / processNeighborsLeaf code
  // (BvhN.hpp:1081-1114,1866-1893,2036-2089) with QueryT = LineSegmentN, a BOUNDED query: every node bound is the box-box distance
  // between the segment's bounding box (LineSegmentN::getBounds, LineSegmentN.hpp:219-224) and the node box (:1794-1803).
  // `excl`: element indices a pushed filter rejects (MeshFix's VertexFilter, Tools/MeshFix/MeshFix.cpp:678-697).
  struct SegClosest { int64_t idx; float d2; V3 on_elem, on_seg; };
  struct SegQuery { Seg g; Box qb; int64_t const * excl; int n_excl; double best; int64_t idx; V3 on_elem, on_seg; };
  static bool seg_excluded(SegQuery const & s, int64_t e) { for (int j = 0; j < s.n_excl; ++j) if (s.excl[j] == e) return true; return false; }
  float elem_seg_sqdist(int64_t e, Seg const & g, V3 & on_elem, V3 & on_seg) const
  {
    if (is_points) { on_elem = pts[(size_t)e]; return seg_point_sqdist(g, pts[(size_t)e], on_seg); }
    return tri_seg_sqdist(tris[(size_t)e], g, on_elem, on_seg);
  }
  template <bool Canon> void seg_rec(int64_t ni, SegQuery & s) const
  {
    KdNode const & nd = nodes[(size_t)ni];
    if (is_leaf(nd))
    {
      for (int64_t k = nd.begin; k < nd.begin + nd.count; ++k)
      {
        int64_t e = order[(size_t)k];
        if (seg_excluded(s, e)) continue;
        V3 oe, os; double mad = (double)elem_seg_sqdist(e, s.g, oe, os);
        bool take = Canon ? (s.best < 0 ? (mad == mad) : (mad < s.best || (mad == s.best && s.idx >= 0 && e < s.idx)))
                          : (s.best < 0 || mad < s.best);
        if (take) { s.best = mad; s.idx = e; s.on_elem = oe; s.on_seg = os; }
      }
      return;
    }
    double lb[2]; int64_t ch[2] = {nd.child[0], nd.child[1]};
    for (int c = 0; c < 2; ++c) lb[c] = (double)s.qb.sqdist_box(nodes[(size_t)ch[c]].bounds);
    if (dist_less(lb[1], lb[0])) { std::swap(lb[0], lb[1]); std::swap(ch[0], ch[1]); }
    for (int c = 0; c < 2; ++c)
    {
      bool go = Canon ? (s.best < 0 || lb[c] <= s.best) : (lb[c] < 0 || s.best < 0 || lb[c] < s.best);
      if (go) seg_rec<Canon>(ch[c], s);
    }
  }
  // The bound handling of closestPair (see closest_bound) with the query's bounding box in place of the point.
  bool seg_bound(Box const & qb, double dist_bound, double & bound2) const
  {
    bound2 = -1;
    if (!(dist_bound >= 0)) return true;
    double lb = (double)qb.sqdist_box(nodes[0].bounds);
    if (lb >= 0 && lb > dist_bound * dist_bound) return false;
    float diag = norm(nodes[0].bounds.extent());
    double fudge = std::max((double)(0.001 * diag), (double)(100 * std::numeric_limits<float>::epsilon()));
    double eff = dist_bound + fudge;
    bound2 = eff * eff;
    return true;
  }
  // mode 0: reference order (first found wins ties), 1: canonical (lowest index), 2: brute force over all elements (canonical)
  SegClosest closest_segment(int mode, V3 end0, V3 end1, double dist_bound, int64_t const * excl, int n_excl) const
  {
    SegClosest r; r.idx = -1; r.d2 = -1; r.on_elem = r.on_seg = mk(0, 0, 0);
    if (nodes.empty()) return r;
    SegQuery s; s.g = make_seg(end0, end1); s.excl = excl; s.n_excl = n_excl; s.idx = -1; s.on_elem = s.on_seg = mk(0, 0, 0);
    s.qb.merge(s.g.point, s.g.point); V3 far = s.g.point + s.g.direction; s.qb.merge(far, far);
    if (!seg_bound(s.qb, dist_bound, s.best)) return r;
    if (mode == 0) seg_rec<false>(0, s);
    else if (mode == 1) seg_rec<true>(0, s);
    else
      for (int64_t e = 0; e < size(); ++e)
      {
        if (seg_excluded(s, e)) continue;
        V3 oe, os; double mad = (double)elem_seg_sqdist(e, s.g, oe, os);
        if (s.best < 0 ? (mad == mad) : (mad < s.best || (mad == s.best && s.idx >= 0 && e < s.idx))) { s.best = mad; s.idx = e; s.on_elem = oe; s.on_seg = os; }
      }
    if (s.idx >= 0) { r.idx = s.idx; r.d2 = (float)s.best; r.on_elem = s.on_elem; r.on_seg = s.on_seg; }
    return r;
  }

  // ---- closest element to a Triangle3 query object: closestPair / processNeighbors / processNeighborsLeaf (BvhN.hpp:1081-1114,1866-1893,
  // 2036-2089) with QueryT = Triangle3 (bounded: box-box node bounds against the query's bounding box, :1794-1803).  Per element:
  // MetricL2::closestPoints(elem, query) = elem.squaredDistance(query) for triangles (tri_tri_sqdist), query.closestPoint(elem) for points.
  float elem_tri_sqdist(int64_t e, Tri const & q, V3 & on_elem, V3 & on_query) const
  {
    if (is_points) { on_elem = pts[(size_t)e]; return tri_point_sqdist(q, pts[(size_t)e], on_query); }
    return tri_tri_sqdist(tris[(size_t)e], q, on_elem, on_query);
  }
  struct TriQuery { Tri q; Box qb; double best; int64_t idx; V3 on_elem, on_query; };
  template <bool Canon> void triq_rec(int64_t ni, TriQuery & s) const
  {
    KdNode const & nd = nodes[(size_t)ni];
    if (is_leaf(nd))
    {
      for (int64_t k = nd.begin; k < nd.begin + nd.count; ++k)
      {
        int64_t e = order[(size_t)k];
        V3 oe, oq; double mad = (double)elem_tri_sqdist(e, s.q, oe, oq);
        bool take = Canon ? (s.best < 0 ? (mad == mad) : (mad < s.best || (mad == s.best && s.idx >= 0 && e < s.idx)))
                          : (s.best < 0 || mad < s.best);
        if (take) { s.best = mad; s.idx = e; s.on_elem = oe; s.on_query = oq; }
      }
      return;
    }
    double lb[2]; int64_t ch[2] = {nd.child[0], nd.child[1]};
    for (int c = 0; c < 2; ++c) lb[c] = (double)s.qb.sqdist_box(nodes[(size_t)ch[c]].bounds);
    if (dist_less(lb[1], lb[0])) { std::swap(lb[0], lb[1]); std::swap(ch[0], ch[1]); }
    for (int c = 0; c < 2; ++c)
    {
      bool go = Canon ? (s.best < 0 || lb[c] <= s.best) : (lb[c] < 0 || s.best < 0 || lb[c] < s.best);
      if (go) triq_rec<Canon>(ch[c], s);
    }
  }
  // mode 0: reference order, 1: canonical through the kd-tree (box bounds only: valid where no false intersection lies beyond a pruned
  // box, see the engine's floor), 2: brute force (the definition).  plain_bound: the structure-to-structure form (no root rejection, no
  // widening of the bound).
  SegClosest closest_triangle(int mode, V3 v0, V3 v1, V3 v2, double dist_bound, bool plain_bound = false) const
  {
    SegClosest r; r.idx = -1; r.d2 = -1; r.on_elem = r.on_seg = mk(0, 0, 0);
    if (nodes.empty()) return r;
    TriQuery s; s.q = make_tri(v0, v1, v2); s.qb.merge(s.q.lo, s.q.hi); s.idx = -1; s.on_elem = s.on_query = mk(0, 0, 0);
    if (plain_bound) s.best = dist_bound >= 0 ? dist_bound * dist_bound : -1;
    else if (!seg_bound(s.qb, dist_bound, s.best)) return r;
    if (mode == 0) triq_rec<false>(0, s);
    else if (mode == 1) triq_rec<true>(0, s);
    else
      for (int64_t e = 0; e < size(); ++e)
      {
        V3 oe, oq; double mad = (double)elem_tri_sqdist(e, s.q, oe, oq);
        if (s.best < 0 ? (mad == mad) : (mad < s.best || (mad == s.best && s.idx >= 0 && e < s.idx))) { s.best = mad; s.idx = e; s.on_elem = oe; s.on_query = oq; }
      }
    if (s.idx >= 0) { r.idx = s.idx; r.d2 = (float)s.best; r.on_elem = s.on_elem; r.on_seg = s.on_query; }
    return r;
  }

  // Conservative double-precision slab test; returns entry time clamped to >= 0, or -1 on a miss.
  static double slab_entry(Box const & b, V3 o, V3 d)
  {
    double tn = 0.0, tf = std::numeric_limits<double>::infinity();
    for (int i = 0; i < 3; ++i)
    {
      double oi = o[i], di = d[i], lo = b.lo[i], hi = b.hi[i];
      if (di == 0.0) { if (oi < lo || oi > hi) return -1; continue; }
      double t0 = (lo - oi) / di, t1 = (hi - oi) / di;
      if (t0 > t1) std::swap(t0, t1);
      tn = std::max(tn, t0); tf = std::min(tf, t1);
    }
    return (tn <= tf * (1.0 + 1e-9) + 1e-30) ? tn * (1.0 - 1e-9) : -1;
  }

  // ---- ray first hit, BvhN.hpp:1291-1316, 2240-2291 ---------------------------------------------
  template <bool Canon> void ray_rec(int64_t ni, V3 o, V3 d, float & best, int64_t & best_idx) const
  {
    KdNode const & nd = nodes[(size_t)ni];
    if (is_leaf(nd))
    {
      for (int64_t k = nd.begin; k < nd.begin + nd.count; ++k)
      {
        int64_t e = order[(size_t)k];
        Tri const & t = tris[(size_t)e];
        float tm = ray_tri_time(o, d, t.v0, t.e01, t.e02);
        // Triangle3.hpp:542-552: valid iff t >= 0 && (max_time < 0 || t <= max_time); then :2257 strict <.
        bool take = Canon ? (tm >= 0 && (best < 0 || tm < best || (tm == best && best_idx >= 0 && e < best_idx)))
                          : dist_less(tm, best);
        if (take) { best = tm; best_idx = e; }
      }
      return;
    }
    float et[2]; int64_t ch[2] = {nd.child[0], nd.child[1]};
    if (Canon)
    {
      // Canonical search: the visit test must never reject a box whose triangles could still win, so the
      // reference's fp32 Woo test (which is only an ordering/pruning heuristic there) is replaced by a
      // double-precision slab test on the padded boxes, entered on tnear <= best (non-strict, for index ties).
      double en[2];
      for (int c = 0; c < 2; ++c) en[c] = slab_entry(nodes[(size_t)ch[c]].bounds, o, d);
      if (en[1] >= 0 && (en[0] < 0 || en[1] < en[0])) { std::swap(en[0], en[1]); std::swap(ch[0], ch[1]); }
      for (int c = 0; c < 2; ++c)
        if (en[c] >= 0 && (best < 0 || en[c] <= (double)best * (1.0 + 1e-6))) ray_rec<Canon>(ch[c], o, d, best, best_idx);
      return;
    }
    for (int c = 0; c < 2; ++c) et[c] = nodes[(size_t)ch[c]].bounds.ray_time(o, d, best);
    if (dist_less(et[1], et[0])) { std::swap(et[0], et[1]); std::swap(ch[0], ch[1]); }
    for (int c = 0; c < 2; ++c)
      if (dist_less(et[c], best)) ray_rec<Canon>(ch[c], o, d, best, best_idx);
  }

  RayHit ray_first_ref(V3 o, V3 d, float max_time) const
  {
    RayHit r; r.idx = -1; r.t = -1;
    if (nodes.empty() || !nodes[0].bounds.ray_hits(o, d, max_time)) return r;
    float best = max_time; int64_t bi = -1;
    ray_rec<false>(0, o, d, best, bi);
    if (bi >= 0) { r.idx = bi; r.t = best; }
    return r;
  }


  // Canonical first hit: min over all triangles of the per-triangle fp32 time, lowest index on ties,
  // t >= 0 and (max_time < 0 or t < max_time).
  RayHit ray_first_canon(V3 o, V3 d, float max_time) const
  {
    RayHit r; r.idx = -1; r.t = -1;
    if (nodes.empty()) return r;
    float best = max_time; int64_t bi = -1;
    if (slab_entry(nodes[0].bounds, o, d) < 0) return r;
    ray_rec<true>(0, o, d, best, bi);
    if (bi >= 0) { r.idx = bi; r.t = best; }
    return r;
  }

  // Canonical k nearest: k smallest (d2, idx) pairs.
  void knn_canon_rec(int64_t ni, V3 q, double bound2, int64_t skip, int k,
                     std::vector<std::pair<float, int64_t>> & s) const
  {
    KdNode const & nd = nodes[(size_t)ni];
    if (is_leaf(nd))
    {
      for (int64_t kk = nd.begin; kk < nd.begin + nd.count; ++kk)
      {
        int64_t e = order[(size_t)kk];
        if (rejects(e, skip)) continue;
        V3 cp; float d2 = elem_sqdist(e, q, cp);
        if (!(d2 == d2)) continue;
        if (bound2 >= 0 && !((double)d2 < bound2)) continue;
        std::pair<float, int64_t> c(d2, e);
        if ((int)s.size() < k) s.insert(std::upper_bound(s.begin(), s.end(), c), c);
        else if (k > 0 && c < s.back()) { s.pop_back(); s.insert(std::upper_bound(s.begin(), s.end(), c), c); }
      }
      return;
    }
    float lb[2]; int64_t ch[2] = {nd.child[0], nd.child[1]};
    for (int c = 0; c < 2; ++c) lb[c] = nodes[(size_t)ch[c]].bounds.sqdist_point_as_box(q);
    if (lb[1] < lb[0]) { std::swap(lb[0], lb[1]); std::swap(ch[0], ch[1]); }
    for (int c = 0; c < 2; ++c)
      if ((int)s.size() < k || lb[c] <= s.back().first) knn_canon_rec(ch[c], q, bound2, skip, k, s);
  }

  // ---- k nearest, BvhN.hpp:1137-1146 + KClosestPairsFunctor :474-500 + BoundedSortedArray ---------
  // Reference-order semantics: a full set only admits strictly smaller distances; equal distances are appended
  // after existing equals (NeighborPair::operator< quirk, ProximityQueryStructureN.hpp:117-123).
  struct KnnSet { int k; std::vector<double> d; std::vector<int64_t> id; };

  void knn_rec(int64_t ni, V3 q, double bound2, int64_t skip, KnnSet & s) const
  {
    KdNode const & nd = nodes[(size_t)ni];
    auto insertable = [&](double mad) { return s.k > 0 && ((int)s.d.size() < s.k || mad < s.d.back()); };
    if (is_leaf(nd))
    {
      for (int64_t k = nd.begin; k < nd.begin + nd.count; ++k)
      {
        int64_t e = order[(size_t)k];
        if (rejects(e, skip)) continue;
        V3 cp; double mad = (double)elem_sqdist(e, q, cp);
        if ((bound2 < 0 || mad < bound2) && insertable(mad))
        {
          size_t ub = (size_t)(std::upper_bound(s.d.begin(), s.d.end(), mad) - s.d.begin());
          if ((int)s.d.size() == s.k) { s.d.pop_back(); s.id.pop_back(); }
          s.d.insert(s.d.begin() + ub, mad); s.id.insert(s.id.begin() + ub, e);
        }
      }
      return;
    }
    double lb[2]; int64_t ch[2] = {nd.child[0], nd.child[1]};
    for (int c = 0; c < 2; ++c) lb[c] = (double)nodes[(size_t)ch[c]].bounds.sqdist_point_as_box(q);
    if (dist_less(lb[1], lb[0])) { std::swap(lb[0], lb[1]); std::swap(ch[0], ch[1]); }
    // allows(NeighborPair(-1,-1,lb)): isInsertable -> (lb, -1, -1) < back  <=>  lb <= back.d (query_index -1 < 0).
    for (int c = 0; c < 2; ++c)
      if ((int)s.d.size() < s.k || lb[c] <= s.d.back()) knn_rec(ch[c], q, bound2, skip, s);
  }
};

// ------------------------------------------------------------------------------------------------
// Canonical brute-force definitions (SURVEY.md finding 4: every query result is tree independent).
// ------------------------------------------------------------------------------------------------
inline Closest brute_closest(KdTree const & T, V3 q, double dist_bound, int64_t skip = -1)
{
  Closest r; r.idx = -1; r.d2 = -1; r.cp = mk(0, 0, 0);
  double bound2 = -1;
  if (T.size() == 0 || !T.closest_bound(q, dist_bound, bound2)) return r;
  float best = 0; bool have = false;
  int64_t n = T.size();
  for (int64_t e = 0; e < n; ++e)
  {
    if (T.rejects(e, skip)) continue;
    V3 cp; float d2 = T.elem_sqdist(e, q, cp);
    if (!(d2 == d2)) continue;
    if (bound2 >= 0 && !((double)d2 < bound2)) continue;     // strict (BvhN.hpp:448,462)
    if (!have || d2 < best) { have = true; best = d2; r.idx = e; r.d2 = d2; r.cp = cp; }   // ascending e => lowest index on ties
  }
  return r;
}

inline RayHit brute_ray_first(KdTree const & T, V3 o, V3 d, float max_time)
{
  RayHit r; r.idx = -1; r.t = -1;
  int64_t n = T.size();
  for (int64_t e = 0; e < n; ++e)
  {
    Tri const & t = T.tris[(size_t)e];
    float tm = ray_tri_time(o, d, t.v0, t.e01, t.e02);
    if (!(tm >= 0)) continue;
    if (max_time >= 0 && !(tm < max_time)) continue;         // BvhN.hpp:2251,2257: strict vs the caller's bound
    if (r.idx < 0 || tm < r.t) { r.idx = e; r.t = tm; }
  }
  return r;
}

// k smallest by (d2, idx); strict dist bound; optional self exclusion.  Output ascending, padded with -1.
inline int brute_knn(KdTree const & T, V3 q, int k, double dist_bound, int64_t skip, int64_t * idx, float * d2out)
{
  double bound2 = dist_bound >= 0 ? dist_bound * dist_bound : -1;
  std::vector<std::pair<float, int64_t>> best;
  int64_t n = T.size();
  for (int64_t e = 0; e < n; ++e)
  {
    if (T.rejects(e, skip)) continue;
    V3 cp; float d2 = T.elem_sqdist(e, q, cp);
    if (!(d2 == d2)) continue;
    if (bound2 >= 0 && !((double)d2 < bound2)) continue;
    std::pair<float, int64_t> c(d2, e);
    if ((int)best.size() < k) { best.insert(std::upper_bound(best.begin(), best.end(), c), c); }
    else if (k > 0 && c < best.back()) { best.pop_back(); best.insert(std::upper_bound(best.begin(), best.end(), c), c); }
  }
  for (int i = 0; i < k; ++i)
  { idx[i] = i < (int)best.size() ? best[(size_t)i].second : -1; d2out[i] = i < (int)best.size() ? best[(size_t)i].first : -1.0f; }
  return (int)best.size();
}

// PointSet3.hpp:586-681 findSampleNeighbors + NeighborFunctor, SamplePoint3.hpp:52-55 ordering.
// Candidate j of sample i at radius r: j != i; sq(x_j - x_i) <= r*r (BallN.hpp:83); n_i . n_j >= min_dot when
// normals exist; key (fastSqrt(d2), j) ascending, keep the max_degree smallest.  Up to 3 rounds, radius doubled
// while fewer than clamp(int(0.25*max_degree), 4, max_degree) neighbours were found.
// `blocked(j, sep, diff)`: the reachability test of NeighborFunctor::isValidNeighbor (PointSet3.hpp:639-648), evaluated after the
// identity / angle tests and the separation, before insertion, for every candidate in range -- as the reference orders them.
template <typename RangeFn, typename BlockedFn>
inline int ball_graph_one(KdTree const & T, int64_t i, int max_degree, float init_radius, RangeFn range,
                          int64_t * idx, float * sep, BlockedFn blocked)
{
  int min_degree = std::min(std::max((int)(0.25 * max_degree), 4), max_degree);
  float radius = init_radius;
  std::vector<std::pair<float, int64_t>> nb;
  bool has_n = !T.nrm.empty();
  for (int it = 0; it < 3; ++it)
  {
    nb.clear();
    float r2 = radius * radius;
    V3 c = T.pts[(size_t)i];
    range(c, r2, [&](int64_t j) {
      if (j == i) return;
      if (has_n && dot(T.nrm[(size_t)i], T.nrm[(size_t)j]) < -0.5f) return;
      V3 diff = T.pts[(size_t)j] - c;
      float s = fast_sqrt(sqnorm(diff));
      if (blocked(j, s, diff)) return;
      std::pair<float, int64_t> cand(s, j);
      if ((int)nb.size() < max_degree) nb.insert(std::upper_bound(nb.begin(), nb.end(), cand), cand);
      else if (max_degree > 0 && cand < nb.back()) { nb.pop_back(); nb.insert(std::upper_bound(nb.begin(), nb.end(), cand), cand); }
    });
    if ((int)nb.size() >= min_degree) break;
    else radius *= 2.0f;
  }
  for (int k = 0; k < max_degree; ++k)
  { idx[k] = k < (int)nb.size() ? nb[(size_t)k].second : -1; sep[k] = k < (int)nb.size() ? nb[(size_t)k].first : -1.0f; }
  return (int)nb.size();
}

// BvhN.hpp:2118-2171 processRange with a Ball3 range on a point kd-tree (reference traversal).
template <typename RangeFn>
inline int ball_graph_one(KdTree const & T, int64_t i, int max_degree, float init_radius, RangeFn range, int64_t * idx, float * sep)
{ return ball_graph_one(T, i, max_degree, init_radius, range, idx, sep, [](int64_t, float, V3) { return false; }); }

template <typename Emit>
inline void kd_ball_range(KdTree const & T, int64_t ni, V3 c, float r2, Emit & emit)
{
  KdNode const & nd = T.nodes[(size_t)ni];
  if (!(nd.bounds.sqdist_point(c) <= r2)) return;                       // BallN.hpp:77-80
  if (nd.count > 0 && nd.bounds.inside_ball(c, r2))
  { for (int64_t k = nd.begin; k < nd.begin + nd.count; ++k) emit(T.order[(size_t)k]); }
  else if (T.is_leaf(nd))
  {
    for (int64_t k = nd.begin; k < nd.begin + nd.count; ++k)
    { int64_t e = T.order[(size_t)k]; if (sqnorm(T.pts[(size_t)e] - c) <= r2) emit(e); }   // BallN.hpp:83
  }
  else { kd_ball_range(T, nd.child[0], c, r2, emit); kd_ball_range(T, nd.child[1], c, r2, emit); }
}

// ------------------------------------------------------------------------------------------------
// SDF: ShapeDiameter.hpp:168-258, MatrixUtil.hpp:225-249.
// ------------------------------------------------------------------------------------------------
static float const CONE_DIRS[30][3] = {
  { 0.407497f,  0.134086f,  0.903309f}, { 0.128722f, -0.822715f,  0.553688f}, { 0.359298f,  0.686129f,  0.632560f},
  { 0.813318f, -0.153956f,  0.561081f}, {-0.707169f,  0.296987f,  0.641647f}, {-0.478783f, -0.543441f,  0.689520f},
  { 0.561083f,  0.523634f,  0.641087f}, {-0.380898f, -0.655789f,  0.651811f}, { 0.722588f,  0.472155f,  0.504912f},
  {-0.535205f,  0.421235f,  0.732200f}, { 0.577593f, -0.477896f,  0.661817f}, { 0.393652f, -0.144484f,  0.907834f},
  { 0.141508f, -0.660777f,  0.737122f}, {-0.274967f, -0.141703f,  0.950954f}, {-0.304935f,  0.474085f,  0.825989f},
  { 0.711150f,  0.301383f,  0.635164f}, {-0.696119f, -0.112341f,  0.709082f}, {-0.248327f,  0.381035f,  0.890588f},
  {-0.0772394f, 0.0462501f, 0.995939f}, { 0.438024f,  0.656463f,  0.614159f}, {-0.483384f, -0.542477f,  0.687065f},
  { 0.498569f, -0.697466f,  0.514752f}, {-0.128902f, -0.0173429f, 0.991506f}, { 0.338759f, -0.318772f,  0.885227f},
  {-0.574364f, -0.463887f,  0.674474f}, {-0.459207f,  0.260125f,  0.849390f}, {-0.628057f, -0.528928f,  0.570771f},
  { 0.232916f, -0.328273f,  0.915416f}, { 0.231667f, -0.796882f,  0.557951f}, {-0.270612f, -0.809654f,  0.520797f},
};

// Builds the 30 cone rays of one sample: origin[3], dirs[30][3].
inline void sdf_rays(V3 position, V3 normal, float scale, V3 & origin, V3 dirs[30])
{
  V3 in = -normalized(normal);
  V3 w = normalized(in);
  V3 u = (max_abs_axis(in) == 0) ? normalized(mk(in.y, -in.x, 0)) : normalized(mk(0, in.z, -in.y));
  V3 v = cross(w, u);
  V3 offset = (0.001f * scale) * in;
  origin = position + offset;
  for (int i = 0; i < 30; ++i)
  {
    V3 c = mk(CONE_DIRS[i][0], CONE_DIRS[i][1], CONE_DIRS[i][2]);
    // rot = [u v w] columns; (rot * c)_r = sum3(u_r c0, v_r c1, w_r c2)
    dirs[i] = mk(sum3(u.x * c.x, v.x * c.y, w.x * c.z), sum3(u.y * c.x, v.y * c.y, w.y * c.z),
                 sum3(u.z * c.x, v.z * c.y, w.z * c.z));
  }
}

// Statistics of ShapeDiameter.hpp:208-257 given per-ray hit times (negative = miss) and hit normals' dot sign
// already applied by the caller (valid[i]).  Reproduces quirk Q8 (nth_element permutes values, not weights).
inline double sdf_stats(float const t[30], bool const valid[30], float scale)
{
  double values[30], weights[30]; int nv = 0;
  for (int i = 0; i < 30; ++i)
    if (valid[i]) { values[nv] = t[i]; weights[nv] = CONE_DIRS[i][2]; nv++; }
  if (nv <= 0) return -1.0;
  int mid = nv / 2;
  std::nth_element(values, values + mid, values + nv);
  double sv = 0, ss = 0;
  for (int i = 0; i < nv; ++i) { sv += values[i]; ss += values[i] * values[i]; }
  double avg = sv / nv, var = ss / nv - avg * avg;
  sv = 0; double sw = 0;
  for (int i = 0; i < nv; ++i)
  { double dv = values[i] - values[mid]; if (dv * dv <= var) { sv += weights[i] * values[i]; sw += weights[i]; } }
  double r = sw > 0 ? (double)(sv / (sw * scale)) : (double)(values[mid] / scale);
  return r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);
}

}  // namespace oracle
