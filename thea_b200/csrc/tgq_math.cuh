// tgq_math.cuh -- per-pair fp32 arithmetic of the Thea BvhN query path, in the reference's exact operation order.
//
// Every function here must produce bit-identical fp32 results to the reference CPU build (x86-64 SSE2 scalar, no FMA
// contraction, Code/Build/Common/BuildSetup.cmake:95-111), so this translation unit family is compiled with
// -fmad=false and IEEE division / square root (nvcc defaults -prec-div=true -prec-sqrt=true).  Three-term sums use
// Eigen's fixed-size-3 reduction order t0 + (t1 + t2).  Citations are relative to /root/reference/Code/Source/.
#pragma once
#ifdef __CUDACC__
#include <cuda_fp16.h>
#endif
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <string.h>

#ifdef __CUDACC__
#define TGQ_HD __host__ __device__ __forceinline__
#else
#define TGQ_HD inline
#endif

namespace tgq {

#ifdef __CUDACC__
typedef float4 f4;
#else
struct alignas(16) f4 { float x, y, z, w; };
#endif

struct v3 { float x, y, z; };

TGQ_HD v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
TGQ_HD v3 add3(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
TGQ_HD v3 sub3(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
TGQ_HD v3 neg3(v3 a) { return mk3(-a.x, -a.y, -a.z); }
TGQ_HD v3 mul3(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
TGQ_HD v3 div3(v3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
TGQ_HD v3 min3(v3 a, v3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
TGQ_HD v3 max3(v3 a, v3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
TGQ_HD float sum3(float a, float b, float c) { return a + (b + c); }
TGQ_HD float dot3(v3 a, v3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
TGQ_HD float sqn3(v3 a) { return dot3(a, a); }
// Bounds-only helpers: fused multiply-add where the hardware has it.  NEVER used in reference arithmetic (the library is
// compiled with -fmad=false so that plain a * b + c is never contracted; these opt in explicitly).
TGQ_HD float ffma(float a, float b, float c)
{
#ifdef __CUDA_ARCH__
  return __fmaf_rn(a, b, c);
#else
  return a * b + c;
#endif
}
TGQ_HD float fdot3(v3 a, v3 b) { return ffma(a.x, b.x, ffma(a.y, b.y, a.z * b.z)); }
TGQ_HD v3 cross3(v3 l, v3 r) { return mk3(l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x); }
TGQ_HD v3 normalized3(v3 a) { float z = sqn3(a); return z > 0.0f ? div3(a, sqrtf(z)) : a; }
// Eigen 3.4 MatrixBase::stableNormalized (Dot.h): w = max |a_i|, z = sq(a / w); z > 0 ? a / (sqrt(z) * w) : a  (w == 0 gives
// z = NaN and returns a unchanged).  Used by Curvature's ProjectedCurvatureFunctor (SurfaceFeatures/Local/Curvature.cpp:46).
TGQ_HD v3 stable_normalized3(v3 a)
{
  float w = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fabsf(a.z));
  float z = sqn3(div3(a, w));
  return z > 0.0f ? div3(a, sqrtf(z) * w) : a;
}
// SpinImage::compute, one sample (SurfaceFeatures/Local/SpinImage.cpp:64-74): cylindrical coordinates of the sample about the axis
// through `position`, binned with the reference's fp32 expressions (Real = float).  max_radius = max_height = 0.75f * scale.
TGQ_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
TGQ_HD void spin_bins(v3 sample, v3 position, v3 axis, int nr, int nh, float max_radius, float max_height, int & radial_bin, int & height_bin)
{
  v3 offset = sub3(sample, position);
  float height = dot3(offset, axis);
  float radius = sqrtf(sqn3(sub3(offset, mul3(axis, height))));
  radial_bin = clampi((int)floorf((float)nr * (radius / max_radius)), 0, nr - 1);
  height_bin = clampi((int)floorf((float)nh * (0.5f * (height / max_height + 1.0f))), 0, nh - 1);
}
TGQ_HD float comp3(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// MatrixUtil.hpp:108-115 maxAbsAxis: first maximum of |v_i|.
TGQ_HD int max_abs_axis3(v3 v)
{
  float ax = fabsf(v.x), ay = fabsf(v.y), az = fabsf(v.z);
  int m = 0; float mv = ax;
  if (ay > mv) { m = 1; mv = ay; }
  if (az > mv) { m = 2; }
  return m;
}

TGQ_HD int f2i(float f)
{
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int i; memcpy(&i, &f, 4); return i;
#endif
}
TGQ_HD float i2f(int i)
{
#ifdef __CUDA_ARCH__
  return __int_as_float(i);
#else
  float f; memcpy(&f, &i, 4); return f;
#endif
}

// ---- Triangle cache (Triangle3.hpp:143-155, Hyperplane3.hpp:34-41) ------------------------------------------------
struct TriPlane { v3 n; float d; int axis; };

TGQ_HD TriPlane tri_plane(v3 v0, v3 v1, v3 v2)
{
  TriPlane p;
  p.n = normalized3(cross3(sub3(v1, v0), sub3(v2, v0)));
  p.d = dot3(p.n, v0);
  p.axis = max_abs_axis3(p.n);
  return p;
}
TGQ_HD v3 tri_centroid(v3 v0, v3 v1, v3 v2) { return div3(add3(add3(v0, v1), v2), 3.0f); }
TGQ_HD float tri_area(v3 n, v3 v0, v3 v1, v3 v2) { return 0.5f * fabsf(dot3(n, cross3(sub3(v1, v0), sub3(v2, v0)))); }

// ---- Point -> triangle (MetricL2.hpp:210-215, Triangle3.hpp:294-304,1310-1462, HyperplaneN.hpp:145-160) -----------
TGQ_HD float tri_area2(v3 d, v3 e, v3 f, int i, int j)
{ return (comp3(e, i) - comp3(d, i)) * (comp3(f, j) - comp3(d, j)) - (comp3(f, i) - comp3(d, i)) * (comp3(e, j) - comp3(d, j)); }

TGQ_HD bool point_inside_tri(v3 v0, v3 v1, v3 v2, int axis, v3 p)
{
  int i, j;
  if (axis == 1) { i = 2; j = 0; } else if (axis == 2) { i = 0; j = 1; } else { i = 1; j = 2; }
  float area = tri_area2(v0, v1, v2, i, j);
  if (area == 0.0f) return v0.x == p.x && v0.y == p.y && v0.z == p.z;
  float inv = 1.0f / area;
  float b0 = tri_area2(p, v1, v2, i, j) * inv;
  if (b0 < 0.0f || b0 > 1.0f) return false;
  float b1 = tri_area2(v0, p, v2, i, j) * inv;
  if (b1 < 0.0f || b1 > 1.0f) return false;
  float b2 = 1.0f - b0 - b1;
  return b2 >= 0.0f && b2 <= 1.0f;
}

TGQ_HD v3 closest_on_edge(v3 a, v3 b, v3 p, float & d2)
{
  v3 e = sub3(b, a);
  float len = sqrtf(sqn3(e));
  v3 dir = div3(e, len);
  v3 c = sub3(p, a);
  float t = dot3(dir, c);
  v3 r;
  if (t <= 0.0f) r = a;
  else if (t >= len) r = b;
  else r = add3(a, mul3(dir, t));
  d2 = sqn3(sub3(r, p));
  return r;
}

TGQ_HD float tri_point_sqdist(v3 v0, v3 v1, v3 v2, v3 n, float d, int axis, v3 p, v3 & cp)
{
  float sd = dot3(p, n) - d;
  v3 proj = sub3(p, mul3(n, sd));          // sd * n, commutative per component
  if (point_inside_tri(v0, v1, v2, axis, proj)) cp = proj;
  else
  {
    float d0, d1, d2;
    v3 r0 = closest_on_edge(v0, v1, p, d0);
    v3 r1 = closest_on_edge(v1, v2, p, d1);
    v3 r2 = closest_on_edge(v2, v0, p, d2);
    if (d0 < d1) cp = (d0 < d2) ? r0 : r2;
    else         cp = (d1 < d2) ? r1 : r2;
  }
  return sqn3(sub3(cp, p));
}

// ---- The same function with the query-independent sub-expressions taken from the triangle record --------------------
// The reference recomputes, for every (triangle, query) pair, three edge lengths (3 sqrt), three unit edge directions
// (9 divisions) and the reciprocal of the projected area (1 division) -- values that depend on the triangle alone
// (Triangle3.hpp:1319-1336,1401-1407,1435-1445).  They are computed ONCE per triangle at build time by exactly the
// operations above (same order, same rounding), so tri_point_sqdist_pre returns the bit pattern tri_point_sqdist does.
//   e0 = (dir(v0->v1), len)  e1 = (dir(v1->v2), len)  e2 = (dir(v2->v0), len)   ax = (1 / area2, area2)
TGQ_HD void tri_axes(int axis, int & i, int & j) { if (axis == 1) { i = 2; j = 0; } else if (axis == 2) { i = 0; j = 1; } else { i = 1; j = 2; } }

TGQ_HD f4 mkf4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

TGQ_HD f4 tri_edge_frame(v3 a, v3 b)
{
  v3 e = sub3(b, a); float len = sqrtf(sqn3(e)); v3 d = div3(e, len);
  return mkf4(d.x, d.y, d.z, len);
}
TGQ_HD f4 tri_area_frame(v3 v0, v3 v1, v3 v2, int axis)
{
  int i, j; tri_axes(axis, i, j);
  float area = tri_area2(v0, v1, v2, i, j);
  return mkf4(1.0f / area, area, 0.0f, 0.0f);
}

TGQ_HD v3 closest_on_edge_pre(v3 a, v3 b, f4 e, v3 p, float & d2)
{
  v3 dir = mk3(e.x, e.y, e.z);
  v3 c = sub3(p, a);
  float t = dot3(dir, c);
  v3 r;
  if (t <= 0.0f) r = a;
  else if (t >= e.w) r = b;
  else r = add3(a, mul3(dir, t));
  d2 = sqn3(sub3(r, p));
  return r;
}

TGQ_HD float tri_point_sqdist_pre(v3 v0, v3 v1, v3 v2, v3 n, float d, int axis, f4 e0, f4 e1, f4 e2, f4 ax, v3 p, v3 & cp)
{
  float sd = dot3(p, n) - d;
  v3 proj = sub3(p, mul3(n, sd));
  bool inside;
  float area = ax.y;
  if (area == 0.0f) inside = v0.x == proj.x && v0.y == proj.y && v0.z == proj.z;
  else
  {
    int i, j; tri_axes(axis, i, j);
    float inv = ax.x;
    inside = false;
    float b0 = tri_area2(proj, v1, v2, i, j) * inv;
    if (!(b0 < 0.0f || b0 > 1.0f))
    {
      float b1 = tri_area2(v0, proj, v2, i, j) * inv;
      if (!(b1 < 0.0f || b1 > 1.0f))
      {
        float b2 = 1.0f - b0 - b1;
        inside = b2 >= 0.0f && b2 <= 1.0f;
      }
    }
  }
  if (inside) cp = proj;
  else
  {
    float d0, d1, d2;
    v3 r0 = closest_on_edge_pre(v0, v1, e0, p, d0);
    v3 r1 = closest_on_edge_pre(v1, v2, e1, p, d1);
    v3 r2 = closest_on_edge_pre(v2, v0, e2, p, d2);
    if (d0 < d1) cp = (d0 < d2) ? r0 : r2;
    else         cp = (d1 < d2) ? r1 : r2;
  }
  return sqn3(sub3(cp, p));
}

// ---- Ray -> triangle (Triangle3.hpp:1464-1524); returns -1 for a miss ---------------------------------------------
TGQ_HD float ray_tri_time(v3 o, v3 D, v3 v0, v3 e01, v3 e02)
{
  v3 diff = sub3(o, v0);
  v3 N = cross3(e01, e02);
  float DdN = dot3(D, N);
  float s;
  if (DdN > FLT_EPSILON) s = 1.0f;
  else if (DdN < -FLT_EPSILON) { s = -1.0f; DdN = -DdN; }
  else return -1.0f;
  float a = s * dot3(D, cross3(diff, e02));
  if (!(a >= 0.0f)) return -1.0f;
  float b = s * dot3(D, cross3(e01, diff));
  if (!(b >= 0.0f)) return -1.0f;
  if (!(a + b <= DdN)) return -1.0f;
  float q = (-s) * dot3(diff, N);
  if (!(q >= 0.0f)) return -1.0f;
  return q / DdN;
}

// ---- Line-segment query objects -----------------------------------------------------------------
// Math::fuzzyEq(a, 0) with the default tolerance eps(a, 0) (Math.hpp:131-148): FLT_EPSILON, scaled by |a| above 1; an infinite
// operand keeps the unscaled epsilon (so inf is never "equal"), NaN compares false everywhere.
TGQ_HD bool fuzzy_eq0(float a)
{
  if (a == 0.0f) return true;
  float mag = fabsf(a);
  float tol = (!isinf(a) && mag > 1.0f) ? mag * FLT_EPSILON : FLT_EPSILON;
  return mag <= tol;
}
// Math::clamp(x, 0, 1) (Math.hpp:247-252): NaN passes through.
TGQ_HD float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }

// LineSegmentN::closestPoint (LineSegmentN.hpp:106-145).  The segment is (point, direction) with direction = endpoint 1 - endpoint 0 as
// its constructor stores it (:56-58); the far end is point + direction, which need not equal the endpoint that was passed in.
TGQ_HD v3 seg_closest_point(v3 point, v3 dir, v3 p)
{
  float d2 = sqn3(dir);
  if (fuzzy_eq0(d2)) return point;
  v3 v = sub3(p, point);
  float t = dot3(dir, v);
  if (t >= 0.0f && t <= d2) return add3(point, mul3(dir, t / d2));
  float d0 = sqn3(v), d1 = sqn3(sub3(v, dir));
  return d0 < d1 ? point : add3(point, dir);
}
// MetricL2Impl<LineSegmentN, point>::closestPoints (Algorithms/MetricL2.hpp:210-215): squared distance of a point element from the segment.
TGQ_HD float seg_point_sqdist(v3 point, v3 dir, v3 p, v3 & seg_pt)
{ seg_pt = seg_closest_point(point, dir, p); return sqn3(sub3(seg_pt, p)); }

// Internal::closestPtSegmentSegment (LineSegmentN.hpp:262-357; Ericson's routine with the reference's fuzzy degeneracy tests).
TGQ_HD float seg_seg_sqdist(v3 p1, v3 q1, bool line1, v3 p2, v3 q2, bool line2, float & s, float & t, v3 & c1, v3 & c2)
{
  v3 d1 = sub3(q1, p1), d2 = sub3(q2, p2), r = sub3(p1, p2);
  float a = sqn3(d1), e = sqn3(d2), f = dot3(d2, r);
  if (fuzzy_eq0(a))
  {
    if (fuzzy_eq0(e)) { s = t = 0.0f; c1 = p1; c2 = p2; v3 w = sub3(c1, c2); return dot3(w, w); }
    s = 0.0f; t = f / e;
    if (!line2) t = clamp01(t);
  }
  else
  {
    float c = dot3(d1, r);
    if (fuzzy_eq0(e)) { t = 0.0f; s = -c / a; if (!line1) s = clamp01(s); }
    else
    {
      float b = dot3(d1, d2);
      float denom = a * e - b * b;
      if (denom != 0.0f) { s = (b * f - c * e) / denom; if (!line1) s = clamp01(s); }
      else s = 0.0f;
      t = (b * s + f) / e;
      if (!line2)
      {
        if (t < 0.0f) { t = 0.0f; s = -c / a; if (!line1) s = clamp01(s); }
        else if (t > 1.0f) { t = 1.0f; s = (b - c) / a; if (!line1) s = clamp01(s); }
      }
    }
  }
  c1 = add3(p1, mul3(d1, s)); c2 = add3(p2, mul3(d2, t));
  return sqn3(sub3(c1, c2));
}

// TriangleN<3>::squaredDistance(LineSegmentN, &this_pt, &seg_pt) (Triangle3.hpp:495-509) with squaredDistanceToPerimeter (:559-577):
// zero when the segment, cast as a ray over t in [0, 1], hits the triangle; else the nearest of the three edges, each an edge
// LineSegment (v_i, v_j - v_i) against the query through LineSegmentN::squaredDistance(LineSegmentN) (LineSegmentN.hpp:153-164), whose
// far ends are point + direction on both sides.
TGQ_HD float tri_seg_sqdist(v3 v0, v3 v1, v3 v2, v3 sp, v3 sdir, v3 & tri_pt, v3 & seg_pt)
{
  float t = ray_tri_time(sp, sdir, v0, sub3(v1, v0), sub3(v2, v0));
  if (t >= 0.0f && t <= 1.0f) { v3 c = add3(sp, mul3(sdir, t)); tri_pt = c; seg_pt = c; return 0.0f; }
  float d2 = -1.0f;
  v3 sq = add3(sp, sdir);
  for (int i = 0; i < 3; ++i)
  {
    v3 a = i == 0 ? v0 : (i == 1 ? v1 : v2), b = i == 0 ? v1 : (i == 1 ? v2 : v0);
    v3 c1, c2; float s_, t_;
    float ed2 = seg_seg_sqdist(a, add3(a, sub3(b, a)), false, sp, sq, false, s_, t_, c1, c2);
    if (d2 < 0.0f || ed2 < d2) { d2 = ed2; tri_pt = c1; seg_pt = c2; }
  }
  return d2;
}

// ---- Triangle query objects: TriangleN<3>::squaredDistance(TriangleN<3>) -------------------------------------
// Moeller's triangle-triangle test as the reference carries it (Triangle3.hpp:597-1300).  Its macros are plain C: dot products are
// summed left to right, (x + y) + z, NOT in Eigen's order.
TGQ_HD float cdot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
TGQ_HD float tt_abs(float x) { return fabsf(x); }
// THEA_TRI3_EDGE_EDGE_TEST (Triangle3.hpp:649-669) in the projection plane (i0, i1); A = V1 - V0 of the edge being tested
TGQ_HD bool tt_edge_edge(float Ax, float Ay, v3 V0, v3 U0, v3 U1, int i0, int i1)
{
  float Bx = comp3(U0, i0) - comp3(U1, i0), By = comp3(U0, i1) - comp3(U1, i1);
  float Cx = comp3(V0, i0) - comp3(U0, i0), Cy = comp3(V0, i1) - comp3(U0, i1);
  float f = Ay * Bx - Ax * By, d = By * Cx - Bx * Cy;
  if ((f > 0.0f && d >= 0.0f && d <= f) || (f < 0.0f && d <= 0.0f && d >= f))
  {
    float e = Ax * Cy - Ay * Cx;
    if (f > 0.0f) { if (e >= 0.0f && e <= f) return true; }
    else { if (e <= 0.0f && e >= f) return true; }
  }
  return false;
}
// THEA_TRI3_EDGE_AGAINST_TRI_EDGES (:671-683)
TGQ_HD bool tt_edge_tri(v3 V0, v3 V1, v3 U0, v3 U1, v3 U2, int i0, int i1)
{
  float Ax = comp3(V1, i0) - comp3(V0, i0), Ay = comp3(V1, i1) - comp3(V0, i1);
  return tt_edge_edge(Ax, Ay, V0, U0, U1, i0, i1) || tt_edge_edge(Ax, Ay, V0, U1, U2, i0, i1) || tt_edge_edge(Ax, Ay, V0, U2, U0, i0, i1);
}
// THEA_TRI3_POINT_IN_TRI (:685-709)
TGQ_HD bool tt_point_in_tri(v3 V0, v3 U0, v3 U1, v3 U2, int i0, int i1)
{
  float a = comp3(U1, i1) - comp3(U0, i1), b = -(comp3(U1, i0) - comp3(U0, i0));
  float c = -a * comp3(U0, i0) - b * comp3(U0, i1);
  float d0 = a * comp3(V0, i0) + b * comp3(V0, i1) + c;
  a = comp3(U2, i1) - comp3(U1, i1); b = -(comp3(U2, i0) - comp3(U1, i0));
  c = -a * comp3(U1, i0) - b * comp3(U1, i1);
  float d1 = a * comp3(V0, i0) + b * comp3(V0, i1) + c;
  a = comp3(U0, i1) - comp3(U2, i1); b = -(comp3(U0, i0) - comp3(U2, i0));
  c = -a * comp3(U2, i0) - b * comp3(U2, i1);
  float d2 = a * comp3(V0, i0) + b * comp3(V0, i1) + c;
  return d0 * d1 > 0.0f && d0 * d2 > 0.0f;
}
// coplanar_tri_tri (:711-760)
TGQ_HD bool tt_coplanar(v3 N, v3 V0, v3 V1, v3 V2, v3 U0, v3 U1, v3 U2)
{
  float A0 = fabsf(N.x), A1 = fabsf(N.y), A2 = fabsf(N.z);
  int i0, i1;
  if (A0 > A1) { if (A0 > A2) { i0 = 1; i1 = 2; } else { i0 = 0; i1 = 1; } }
  else { if (A2 > A1) { i0 = 0; i1 = 1; } else { i0 = 0; i1 = 2; } }
  if (tt_edge_tri(V0, V1, U0, U1, U2, i0, i1)) return true;
  if (tt_edge_tri(V1, V2, U0, U1, U2, i0, i1)) return true;
  if (tt_edge_tri(V2, V0, U0, U1, U2, i0, i1)) return true;
  if (tt_point_in_tri(V0, U0, U1, U2, i0, i1)) return true;
  if (tt_point_in_tri(U0, V0, V1, V2, i0, i1)) return true;
  return false;
}
// isect2 (:1024-1039)
TGQ_HD void tt_isect2(v3 X0, v3 X1, v3 X2, float VV0, float VV1, float VV2, float D0, float D1, float D2, float & isect0, float & isect1, v3 & p0, v3 & p1)
{
  float tmp = D0 / (D0 - D1);
  isect0 = VV0 + (VV1 - VV0) * tmp;
  v3 diff = sub3(X1, X0); diff = mk3(tmp * diff.x, tmp * diff.y, tmp * diff.z);
  p0 = add3(diff, X0);
  tmp = D0 / (D0 - D2);
  isect1 = VV0 + (VV2 - VV0) * tmp;
  diff = sub3(X2, X0); diff = mk3(tmp * diff.x, tmp * diff.y, tmp * diff.z);
  p1 = add3(X0, diff);
}
// compute_intervals_isectline (:1041-1077): true when the triangles are coplanar
TGQ_HD bool tt_intervals(v3 X0, v3 X1, v3 X2, float VV0, float VV1, float VV2, float D0, float D1, float D2, float D0D1, float D0D2,
                         float & isect0, float & isect1, v3 & p0, v3 & p1)
{
  if (D0D1 > 0.0f) tt_isect2(X2, X0, X1, VV2, VV0, VV1, D2, D0, D1, isect0, isect1, p0, p1);
  else if (D0D2 > 0.0f) tt_isect2(X1, X0, X2, VV1, VV0, VV2, D1, D0, D2, isect0, isect1, p0, p1);
  else if (D1 * D2 > 0.0f || D0 != 0.0f) tt_isect2(X0, X1, X2, VV0, VV1, VV2, D0, D1, D2, isect0, isect1, p0, p1);
  else if (D1 != 0.0f) tt_isect2(X1, X0, X2, VV1, VV0, VV2, D1, D0, D2, isect0, isect1, p0, p1);
  else if (D2 != 0.0f) tt_isect2(X2, X0, X1, VV2, VV0, VV1, D2, D0, D1, isect0, isect1, p0, p1);
  else return true;
  return false;
}
// tri_tri_intersect_with_isectline (:1115-1300): 0 = disjoint, 1 = intersecting; coplanar and the two ends of the intersection segment
TGQ_HD int tt_intersect_line(v3 V0, v3 V1, v3 V2, v3 U0, v3 U1, v3 U2, bool & coplanar, v3 & pt1, v3 & pt2)
{
  coplanar = false;
  v3 E1 = sub3(V1, V0), E2 = sub3(V2, V0);
  v3 N1 = cross3(E1, E2);
  float d1 = -cdot3(N1, V0);
  float du0 = cdot3(N1, U0) + d1, du1 = cdot3(N1, U1) + d1, du2 = cdot3(N1, U2) + d1;
  if (fabsf(du0) < FLT_EPSILON) du0 = 0.0f;
  if (fabsf(du1) < FLT_EPSILON) du1 = 0.0f;
  if (fabsf(du2) < FLT_EPSILON) du2 = 0.0f;
  float du0du1 = du0 * du1, du0du2 = du0 * du2;
  if (du0du1 > 0.0f && du0du2 > 0.0f) return 0;
  E1 = sub3(U1, U0); E2 = sub3(U2, U0);
  v3 N2 = cross3(E1, E2);
  float d2 = -cdot3(N2, U0);
  float dv0 = cdot3(N2, V0) + d2, dv1 = cdot3(N2, V1) + d2, dv2 = cdot3(N2, V2) + d2;
  if (fabsf(dv0) < FLT_EPSILON) dv0 = 0.0f;
  if (fabsf(dv1) < FLT_EPSILON) dv1 = 0.0f;
  if (fabsf(dv2) < FLT_EPSILON) dv2 = 0.0f;
  float dv0dv1 = dv0 * dv1, dv0dv2 = dv0 * dv2;
  if (dv0dv1 > 0.0f && dv0dv2 > 0.0f) return 0;
  v3 D = cross3(N1, N2);
  float mx = fabsf(D.x); int index = 0;
  float b = fabsf(D.y), c = fabsf(D.z);
  if (b > mx) { mx = b; index = 1; }
  if (c > mx) { mx = c; index = 2; }
  float vp0 = comp3(V0, index), vp1 = comp3(V1, index), vp2 = comp3(V2, index);
  float up0 = comp3(U0, index), up1 = comp3(U1, index), up2 = comp3(U2, index);
  float i1a = 0.0f, i1b = 0.0f, i2a = 0.0f, i2b = 0.0f;
  v3 z = mk3(0.0f, 0.0f, 0.0f), A1 = z, A2 = z, B1 = z, B2 = z;
  coplanar = tt_intervals(V0, V1, V2, vp0, vp1, vp2, dv0, dv1, dv2, dv0dv1, dv0dv2, i1a, i1b, A1, A2);
  if (coplanar) return tt_coplanar(N1, V0, V1, V2, U0, U1, U2) ? 1 : 0;
  tt_intervals(U0, U1, U2, up0, up1, up2, du0, du1, du2, du0du1, du0du2, i2a, i2b, B1, B2);
  int smallest1, smallest2;
  if (i1a > i1b) { float t_ = i1a; i1a = i1b; i1b = t_; smallest1 = 1; } else smallest1 = 0;
  if (i2a > i2b) { float t_ = i2a; i2a = i2b; i2b = t_; smallest2 = 1; } else smallest2 = 0;
  if (i1b < i2a || i2b < i1a) return 0;
  if (i2a < i1a)
  {
    pt1 = smallest1 == 0 ? A1 : A2;
    if (i2b < i1b) pt2 = smallest2 == 0 ? B2 : B1;
    else pt2 = smallest1 == 0 ? A2 : A1;
  }
  else
  {
    pt1 = smallest2 == 0 ? B1 : B2;
    if (i2b > i1b) pt2 = smallest1 == 0 ? A2 : A1;
    else pt2 = smallest2 == 0 ? B2 : B1;
  }
  return 1;
}

// One triangle with the cached members the distance routines read (Triangle3.hpp:143-155): unit normal, plane offset, primary axis.
struct TriQ { v3 v0, v1, v2, n; float d; int axis; };
TGQ_HD TriQ make_triq(v3 v0, v3 v1, v3 v2) { TriPlane p = tri_plane(v0, v1, v2); TriQ t; t.v0 = v0; t.v1 = v1; t.v2 = v2; t.n = p.n; t.d = p.d; t.axis = p.axis; return t; }
TGQ_HD v3 triq_vertex(TriQ const & t, int i) { return i == 0 ? t.v0 : (i == 1 ? t.v1 : t.v2); }
// TriangleN<3>::squaredDistance(other, &this_pt, &other_pt) (Triangle3.hpp:314-392; Ericson): `a` is the triangle the member is called
// on -- the structure's element in BvhN::processNeighborsLeaf (MetricL2::closestPoints(elem, query), BvhN.hpp:2063) --, `b` the other.
TGQ_HD float tri_tri_sqdist(TriQ const & a, TriQ const & b, v3 & a_pt, v3 & b_pt)
{
  v3 c1 = mk3(0.0f, 0.0f, 0.0f), c2 = c1;
  float min_sqdist = FLT_MAX;
  bool coplanar; v3 s1, s2;
  if (tt_intersect_line(a.v0, a.v1, a.v2, b.v0, b.v1, b.v2, coplanar, s1, s2))
  {
    if (coplanar) c1 = c2 = mul3(add3(tri_centroid(a.v0, a.v1, a.v2), tri_centroid(b.v0, b.v1, b.v2)), 0.5f);
    else
    {
      v3 dir = sub3(s2, s1);                                                     // LineSegmentN(isectpt1, isectpt2)
      c1 = c2 = mul3(add3(add3(s1, mul3(dir, 0.0f)), add3(s1, mul3(dir, 1.0f))), 0.5f);     // 0.5 * (seg.getPoint(0) + seg.getPoint(1))
    }
    min_sqdist = 0.0f;
  }
  else
  {
    for (int i = 0; i < 3; ++i)
    {
      v3 ai = triq_vertex(a, i), ai2 = triq_vertex(a, (i + 1) % 3);
      for (int j = 0; j < 3; ++j)
      {
        v3 p, q; float s, t;
        float d2 = seg_seg_sqdist(ai, ai2, false, triq_vertex(b, j), triq_vertex(b, (j + 1) % 3), false, s, t, p, q);
        if (d2 < min_sqdist) { min_sqdist = d2; c1 = p; c2 = q; }
      }
    }
    for (int i = 0; i < 3; ++i)          // a vertex of `b` that projects inside `a`
    {
      v3 q = triq_vertex(b, i);
      v3 p = sub3(q, mul3(a.n, dot3(q, a.n) - a.d));
      if (point_inside_tri(a.v0, a.v1, a.v2, a.axis, p))
      {
        float d2 = sqn3(sub3(p, q));
        if (d2 < min_sqdist) { min_sqdist = d2; c1 = p; c2 = q; }
      }
    }
    for (int i = 0; i < 3; ++i)          // a vertex of `a` that projects inside `b`
    {
      v3 p = triq_vertex(a, i);
      v3 q = sub3(p, mul3(b.n, dot3(p, b.n) - b.d));
      if (point_inside_tri(b.v0, b.v1, b.v2, b.axis, q))
      {
        float d2 = sqn3(sub3(p, q));
        if (d2 < min_sqdist) { min_sqdist = d2; c1 = p; c2 = q; }
      }
    }
  }
  a_pt = c1; b_pt = c2;
  return min_sqdist;
}

// ---- Math::fastSqrt (Math.hpp:194-223) -------------------------------------------------------------------------
TGQ_HD float fast_sqrt(float x)
{
  float x2 = x * 0.5f;
  int i = f2i(x);
  i = 0x5f3759df - (i >> 1);
  float y = i2f(i);
  y = y * (1.5f - (x2 * y * y));
  return 1.0f / y;
}

// ---- Box lower bound (AxisAlignedBoxN.hpp:429-456).  Used for ordering and pruning only. ---------------------------
TGQ_HD float box_point_sqdist(v3 lo, v3 hi, v3 p)
{
  v3 g = max3(sub3(max3(lo, p), min3(hi, p)), mk3(0.0f, 0.0f, 0.0f));
  return fdot3(g, g);
}

// ---- ShapeDiameter cone rays (ShapeDiameter.hpp:168-214, MatrixUtil.hpp:225-249) ------------------------------------
struct SdfFrame { v3 origin, u, v, w; };
TGQ_HD SdfFrame sdf_frame(v3 position, v3 normal, float scale)
{
  SdfFrame f;
  v3 in = neg3(normalized3(normal));
  f.w = normalized3(in);
  f.u = (max_abs_axis3(in) == 0) ? normalized3(mk3(in.y, -in.x, 0.0f)) : normalized3(mk3(0.0f, in.z, -in.y));
  f.v = cross3(f.w, f.u);
  f.origin = add3(position, mul3(in, 0.001f * scale));
  return f;
}
TGQ_HD v3 sdf_dir(SdfFrame const & f, v3 c)
{
  return mk3(sum3(f.u.x * c.x, f.v.x * c.y, f.w.x * c.z), sum3(f.u.y * c.x, f.v.y * c.y, f.w.y * c.z),
             sum3(f.u.z * c.x, f.v.z * c.y, f.w.z * c.z));
}

}  // namespace tgq
