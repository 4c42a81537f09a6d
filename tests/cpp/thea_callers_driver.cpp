// thea_callers_driver.cpp -- TEST-ONLY program: the reference's OWN, UNMODIFIED caller templates instantiated on the GPU adaptor.
//
// Compiled (in the container that has /root/reference; recipe in tests/cpp/Makefile) against Thea's real headers -- with the Eigen
// stand-in of oracle/eigen_shim -- and include/thea_gpu/GpuBvh.hpp, linked against libtgq.so.  It runs, side by side,
//   SurfaceFeatures::Local::ShapeDiameter<Mesh, ExternalBvhT>        (Algorithms/SurfaceFeatures/Local/ShapeDiameter.hpp:41,141-258)
//   Icp3<Real>::align(n, from, to_structure, &error)                  (Algorithms/Icp3.hpp:110-126,193-464)
//   the range / nearest-neighbour call shapes of Test/TestBvh.cpp:174-236
// once with the reference's BvhN as the structure and once with TheaGpu::GpuMeshBvhT<Thea::Vector3> / GpuPointBvhT<Thea::Vector3>
// in its place, and compares what the CALLERS return.  That is the drop-in claim of INTEGRATION.md, exercised with the reference's
// code on both sides of the boundary.  Usage: thea_callers_driver <in.bin>   (input written by tests/test_adaptor.py)
#include "Algorithms/BvhN.hpp"
#include "Algorithms/Icp3.hpp"
#include "Algorithms/IntersectionTester.hpp"
#include "Algorithms/MetricL2.hpp"
#include "Algorithms/RayIntersectionTester.hpp"
#include "Algorithms/SurfaceFeatures/Local/ShapeDiameter.hpp"
#include "AxisAlignedBox3.hpp"
#include "Ball3.hpp"
#include "BoundedSortedArrayN.hpp"
#include "BoundedSortedArray.hpp"
#include "Graphics/GeneralMesh.hpp"
#include "LineSegment3.hpp"
#include "Triangle3.hpp"

#include "../../include/thea_gpu/GpuBvh.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

using namespace Thea;
using namespace Thea::Algorithms;

typedef BvhN<Triangle3, 3> RefMeshBvh;
typedef BvhN<Vector3, 3> RefPointBvh;
typedef TheaGpu::GpuMeshBvhT<Vector3> GpuMeshBvh;
typedef TheaGpu::GpuPointBvhT<Vector3> GpuPointBvh;
typedef Graphics::GeneralMesh<> Mesh;

template <typename T> static std::vector<T> rd(FILE * f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }
static Vector3 v3(float const * p) { return Vector3(p[0], p[1], p[2]); }

static int fails = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { ++fails; printf("FAIL "); printf(__VA_ARGS__); printf("\n"); } } while (0)

struct Collect
{
  std::vector<intx> * out; intx stop_at;
  bool operator()(intx index, Vector3 const &) { out->push_back(index); return index == stop_at; }
};

int main(int argc, char ** argv)
{
  if (argc < 2) { fprintf(stderr, "usage: %s in.bin\n", argv[0]); return 2; }
  FILE * fi = fopen(argv[1], "rb"); if (!fi) return 2;
  std::vector<int64_t> hdr = rd<int64_t>(fi, 4);
  int64_t T = hdr[0], S = hdr[1], NI = hdr[2], P = hdr[3];
  std::vector<float> v9 = rd<float>(fi, 9 * T), sp = rd<float>(fi, 3 * S), sn = rd<float>(fi, 3 * S), src = rd<float>(fi, 3 * NI), pts = rd<float>(fi, 3 * P);
  fclose(fi);
  try
  {
    // ---- the two triangle structures over the same triangles --------------------------------------------------------------
    std::vector<Triangle3> tris;
    for (int64_t t = 0; t < T; ++t) tris.push_back(Triangle3(v3(&v9[9 * t]), v3(&v9[9 * t + 3]), v3(&v9[9 * t + 6])));
    RefMeshBvh ref(tris.begin(), tris.end());
    GpuMeshBvh gpu;
    // MeshBvh::add(mesh) + init(): triangles arrive through a visitor, in MeshTriangles' push order
    gpu.add([&](auto emit) { for (auto const & x : tris) emit(x.getVertex(0), x.getVertex(1), x.getVertex(2)); });
    gpu.init();
    EXPECT(gpu.numElements() == (intx)T, "numElements %ld != %ld", (long)gpu.numElements(), (long)T);
    // getElements()[i].getNormal(): Triangle3's cached unit normal, bit for bit
    int bad_normals = 0;
    for (int64_t t = 0; t < T; ++t)
      if (!(gpu.getElements()[(size_t)t].getNormal() == ref.getElements()[t].getNormal())) ++bad_normals;
    EXPECT(bad_normals == 0, "%d element normals differ from Triangle3::getNormal()", bad_normals);

    // ---- ShapeDiameter with the structure as ExternalBvhT ------------------------------------------------------------------
    AxisAlignedBox3 bounds = ref.getBounds();
    Real scale = bounds.getExtent().norm();
    SurfaceFeatures::Local::ShapeDiameter<Mesh, RefMeshBvh> sdf_ref(&ref, scale);
    SurfaceFeatures::Local::ShapeDiameter<Mesh, GpuMeshBvh> sdf_gpu(&gpu, scale);
    int sdf_equal = 0, sdf_pos_equal = 0, sdf_valid = 0;
    for (int64_t i = 0; i < S; ++i)
    {
      double a = sdf_ref.compute(v3(&sp[3 * i]), v3(&sn[3 * i]), true), b = sdf_gpu.compute(v3(&sp[3 * i]), v3(&sn[3 * i]), true);
      if (a == b) ++sdf_equal;
      if (a >= 0) ++sdf_valid;
      if (i < S / 4)
      {
        // compute(position): closestElement + getElements()[nn].getNormal() + the rays (ShapeDiameter.hpp:141-158)
        double c = sdf_ref.compute(v3(&sp[3 * i]), true), d = sdf_gpu.compute(v3(&sp[3 * i]), true);
        if (c == d) ++sdf_pos_equal;
      }
    }
    EXPECT(sdf_equal == S, "ShapeDiameter::compute(p, n): %d of %ld values equal", sdf_equal, (long)S);
    // compute(position) takes the normal of the nearest triangle: where the reference's first-found nearest triangle and the
    // canonical lowest-index one differ (exact ties on shared edges) the two normals differ, so equality is required of 90 %
    EXPECT(sdf_pos_equal >= (int)(0.9 * (S / 4)), "ShapeDiameter::compute(p): %d of %ld values equal", sdf_pos_equal, (long)(S / 4));
    EXPECT(sdf_valid > S / 2, "only %d of %ld SDF values are valid", sdf_valid, (long)S);

    // ---- Icp3::align with the structure as ToProximityQueryStructureT ------------------------------------------------------
    std::vector<Vector3> from; for (int64_t i = 0; i < NI; ++i) from.push_back(v3(&src[3 * i]));
    Icp3<Real> icp(1e-4, 5, 5, false);
    Real err_ref = 0, err_gpu = 0;
    AffineTransform3 tr_ref = icp.align((intx)NI, from.data(), ref, &err_ref);
    AffineTransform3 tr_gpu = icp.align((intx)NI, from.data(), gpu, &err_gpu);
    double tdiff = 0;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tdiff = std::max(tdiff, (double)std::fabs(tr_ref.getLinear()(r, c) - tr_gpu.getLinear()(r, c)));
                                  tdiff = std::max(tdiff, (double)std::fabs(tr_ref.getTranslation()[r] - tr_gpu.getTranslation()[r])); }
    EXPECT(tdiff <= 1e-5, "Icp3::align transforms differ by %g", tdiff);
    EXPECT(std::fabs(err_ref - err_gpu) <= 1e-5 * std::max(1.0f, err_ref), "Icp3 errors %g vs %g", (double)err_ref, (double)err_gpu);

    // ---- TestBvh.cpp:143-236 call shapes on a point structure ---------------------------------------------------------------
    std::vector<Vector3> points; for (int64_t i = 0; i < P; ++i) points.push_back(v3(&pts[3 * i]));
    RefPointBvh pref(points.begin(), points.end());
    GpuPointBvh pgpu; pgpu.init(points.begin(), points.end());
    Vector3 mid = 0.5f * (pref.getBounds().getLow() + pref.getBounds().getHigh());
    Real diag = pref.getBounds().getExtent().norm();
    Ball3 ball(mid + Vector3(0.3f * diag, 0, 0), 0.12f * diag);
    AxisAlignedBox3 box(mid - Vector3(0.45f, 0.1f, 0.45f) * diag, mid + Vector3(0.45f, 0.1f, 0.2f) * diag);
    int range_ok = 0;
    {
      Array<intx> a, b;
      pref.rangeQueryIndices<IntersectionTester>(ball, a); pgpu.rangeQueryIndices<IntersectionTester>(ball, b);
      std::sort(a.begin(), a.end());
      range_ok += (a == b && !a.empty());
      EXPECT(a == b && !a.empty(), "rangeQueryIndices(ball): %zu vs %zu elements", a.size(), b.size());
      Array<intx> c, d;
      pref.rangeQueryIndices<IntersectionTester>(box, c); pgpu.rangeQueryIndices<IntersectionTester>(box, d);
      std::sort(c.begin(), c.end());
      range_ok += (c == d && !c.empty());
      EXPECT(c == d && !c.empty(), "rangeQueryIndices(box): %zu vs %zu elements", c.size(), d.size());
      Array<Vector3> ea, eb;
      pref.rangeQuery<IntersectionTester>(ball, ea); pgpu.rangeQuery<IntersectionTester>(ball, eb);
      EXPECT(ea.size() == eb.size() && ea.size() == a.size(), "rangeQuery(ball): %zu vs %zu elements", ea.size(), eb.size());
      // processRange with a functor: every element once; with early termination the returned index is the element that stopped it
      std::vector<intx> seen; Collect all = { &seen, -5 };
      intx r0 = pgpu.processRange<IntersectionTester>(box, all);
      EXPECT(r0 < 0 && seen.size() == d.size(), "processRange visited %zu of %zu elements", seen.size(), d.size());
      std::vector<intx> seen2; Collect some = { &seen2, d.empty() ? -5 : d[d.size() / 2] };
      intx r1 = pgpu.processRange<IntersectionTester>(box, some);
      EXPECT(!d.empty() && r1 == d[d.size() / 2], "processRange early termination returned %ld", (long)r1);
    }
    // nearest neighbour with a distance bound, k nearest neighbours (TestBvh.cpp:199-236)
    int nn_equal = 0, knn_equal = 0, nq = 200;
    for (int i = 0; i < nq; ++i)
    {
      Vector3 q = points[(size_t)((i * 7919) % P)] + Vector3(0.003f, -0.002f, 0.001f) * (Real)(i % 5);
      double da = -1, db = -1;
      intx ia = pref.closestElement<MetricL2>(q, 0.02 * diag, UniversalCompatibility(), &da);
      intx ib = pgpu.closestElement<MetricL2>(q, 0.02 * diag, UniversalCompatibility(), &db);
      if (ia == ib && da == db) ++nn_equal;
      BoundedSortedArrayN<3, RefPointBvh::NeighborPair> na;
      BoundedSortedArrayN<3, GpuPointBvh::NeighborPair> nb;
      intx ka = pref.kClosestPairs<MetricL2>(q, na, 0.05 * diag);
      intx kb = pgpu.kClosestPairs<MetricL2>(q, nb, 0.05 * diag);
      bool same = ka == kb;
      for (size_t j = 0; same && j < na.size(); ++j)
        same = na[j].getTargetIndex() == nb[j].getTargetIndex() && na[j].getMonotoneApproxDistance() == nb[j].getMonotoneApproxDistance();
      if (same) ++knn_equal;
    }
    EXPECT(nn_equal == nq, "closestElement with dist_bound: %d of %d equal", nn_equal, nq);
    EXPECT(knn_equal >= nq - 2, "kClosestPairs: %d of %d equal", knn_equal, nq);

    // ---- structure-to-structure closestPair (TestBvh.cpp:251-258): point structure vs point structure, triangles vs points -------
    int pair_ok = 0;
    {
      std::vector<Vector3> other; for (int64_t i = 0; i < P; i += 3) other.push_back(points[(size_t)i] + Vector3(0.013f, 0.007f, -0.011f));
      RefPointBvh oref(other.begin(), other.end());
      GpuPointBvh ogpu; ogpu.init(other.begin(), other.end());
      for (double bound : { -1.0, 0.5 * (double)diag, 1e-4 })
      {
        RefPointBvh::NeighborPair a = pref.closestPair<MetricL2>(oref, bound, UniversalCompatibility(), true);
        GpuPointBvh::NeighborPair b = pgpu.closestPair<MetricL2>(ogpu, bound, UniversalCompatibility(), true);
        bool same = a.isValid() == b.isValid() && (!a.isValid() || (a.getMonotoneApproxDistance() == b.getMonotoneApproxDistance()));
        // indices may differ only through exact ties (the reference keeps the first pair it finds)
        if (same && a.isValid() && (a.getTargetIndex() != b.getTargetIndex() || a.getQueryIndex() != b.getQueryIndex()))
          same = (points[(size_t)b.getTargetIndex()] - other[(size_t)b.getQueryIndex()]).squaredNorm() == (Real)a.getMonotoneApproxDistance();
        pair_ok += same;
        EXPECT(same, "closestPair(points, points, bound %g): %g vs %g", bound, a.getMonotoneApproxDistance(), b.getMonotoneApproxDistance());
        RefMeshBvh::NeighborPair c = ref.closestPair<MetricL2>(oref, bound, UniversalCompatibility(), true);
        GpuMeshBvh::NeighborPair d = gpu.closestPair<MetricL2>(ogpu, bound, UniversalCompatibility(), true);
        bool same2 = c.isValid() == d.isValid() && (!c.isValid() || c.getMonotoneApproxDistance() == d.getMonotoneApproxDistance());
        pair_ok += same2;
        EXPECT(same2, "closestPair(triangles, points, bound %g): %g vs %g", bound, c.getMonotoneApproxDistance(), d.getMonotoneApproxDistance());
      }
      // kClosestPairs with a structure as the query (BvhN.hpp:1137-1146): the reference's own template on both sides
      for (int k : { 1, 7, 20 })
      {
        Thea::BoundedSortedArray<RefPointBvh::NeighborPair> ra((size_t)k);
        Thea::BoundedSortedArray<GpuPointBvh::NeighborPair> ga((size_t)k);
        pref.kClosestPairs<MetricL2>(oref, ra, -1, UniversalCompatibility(), false);
        pgpu.kClosestPairs<MetricL2>(ogpu, ga, -1, UniversalCompatibility(), false);
        bool same3 = ra.size() == ga.size();
        for (size_t j = 0; same3 && j < ra.size(); ++j)
          same3 = ra[j].getMonotoneApproxDistance() == ga[j].getMonotoneApproxDistance()
                  && ((ra[j].getTargetIndex() == ga[j].getTargetIndex() && ra[j].getQueryIndex() == ga[j].getQueryIndex())
                      || (points[(size_t)ga[j].getTargetIndex()] - other[(size_t)ga[j].getQueryIndex()]).squaredNorm() == (Real)ra[j].getMonotoneApproxDistance());
        pair_ok += same3;
        EXPECT(same3, "kClosestPairs(points, points, k = %d): %zu vs %zu pairs", k, ra.size(), ga.size());
      }
    }

    // ---- line-segment query objects: Thea's own LineSegment3 handed to both structures (the call of Tools/MeshFix/MeshFix.cpp:762) ----
    int seg_ok = 0, nseg = 200;
    for (int i = 0; i < nseg; ++i)
    {
      Vector3 a = points[(size_t)((i * 104729) % P)] + Vector3(0.004f, 0.001f, -0.003f) * (Real)(i % 7);
      Vector3 b = points[(size_t)((i * 7919 + 13) % P)] * (i % 3 == 0 ? 1.0f : 0.02f) + a * (i % 3 == 0 ? 0.0f : 0.98f);
      LineSegment3 seg(a, b);
      double bound = (i % 4 == 3) ? 0.01 * (double)diag : -1.0;
      RefPointBvh::NeighborPair pa = pref.closestPair<MetricL2>(seg, bound, UniversalCompatibility(), true);
      GpuPointBvh::NeighborPair pb = pgpu.closestPair<MetricL2>(seg, bound, UniversalCompatibility(), true);
      bool same = pa.isValid() == pb.isValid()
               && (!pa.isValid() || (pa.getMonotoneApproxDistance() == pb.getMonotoneApproxDistance()
                                     && (pa.getTargetIndex() != pb.getTargetIndex()          // exact ties: the reference keeps the first it finds
                                         || (pa.getQueryPoint() == pb.getQueryPoint() && pa.getTargetPoint() == pb.getTargetPoint()))));
      RefMeshBvh::NeighborPair ta = ref.closestPair<MetricL2>(seg, bound, UniversalCompatibility(), true);
      GpuMeshBvh::NeighborPair tb = gpu.closestPair<MetricL2>(seg, bound, UniversalCompatibility(), true);
      bool same2 = ta.isValid() == tb.isValid()
                && (!ta.isValid() || (ta.getMonotoneApproxDistance() == tb.getMonotoneApproxDistance()
                                      && (ta.getTargetIndex() != tb.getTargetIndex()
                                          || (ta.getQueryPoint() == tb.getQueryPoint() && ta.getTargetPoint() == tb.getTargetPoint()))));
      double da = -1, db = -1;
      intx ia = pref.closestElement<MetricL2>(seg, bound, UniversalCompatibility(), &da);
      intx ib = pgpu.closestElement<MetricL2>(seg, bound, UniversalCompatibility(), &db);
      bool same3 = da == db && (ia == ib || ia < 0 == ib < 0) && pref.distance<MetricL2>(seg, bound) == pgpu.distance<MetricL2>(seg, bound);
      seg_ok += (same && same2 && same3);
      EXPECT(same && same2 && same3, "segment query %d: points %g vs %g, triangles %g vs %g, closestElement %ld vs %ld", i, pa.getMonotoneApproxDistance(),
             pb.getMonotoneApproxDistance(), ta.getMonotoneApproxDistance(), tb.getMonotoneApproxDistance(), (long)ia, (long)ib);
    }

    // ---- triangle query objects and mesh-to-mesh closestPair (Test/TestBvh.cpp:380-385): Thea's own Triangle3 on both sides ----------------
    int triq_ok = 0, ntriq = 120, mesh_pair_ok = 0, mesh_pair_bit_equal = 0;
    for (int i = 0; i < ntriq; ++i)
    {
      Vector3 c = points[(size_t)((i * 15485863) % P)] * (1.0f + 0.05f * (Real)(i % 5)) + Vector3(0.002f, -0.001f, 0.003f);
      Real sz = (i % 3 == 0 ? 0.15f : 0.01f) * diag;
      Triangle3 q(c, c + Vector3(sz, 0.1f * sz, 0), c + Vector3(0.2f * sz, sz, 0.3f * sz));
      double bound = (i % 4 == 3) ? 0.02 * (double)diag : -1.0;
      RefMeshBvh::NeighborPair ta = ref.closestPair<MetricL2>(q, bound, UniversalCompatibility(), true);
      GpuMeshBvh::NeighborPair tb = gpu.closestPair<MetricL2>(q, bound, UniversalCompatibility(), true);
      bool same = ta.isValid() == tb.isValid()
               && (!ta.isValid() || (ta.getMonotoneApproxDistance() == tb.getMonotoneApproxDistance()
                                     && (ta.getTargetIndex() != tb.getTargetIndex()
                                         || (ta.getQueryPoint() == tb.getQueryPoint() && ta.getTargetPoint() == tb.getTargetPoint()))));
      RefPointBvh::NeighborPair pa = pref.closestPair<MetricL2>(q, bound, UniversalCompatibility(), true);
      GpuPointBvh::NeighborPair pb = pgpu.closestPair<MetricL2>(q, bound, UniversalCompatibility(), true);
      bool same2 = pa.isValid() == pb.isValid()
                && (!pa.isValid() || (pa.getMonotoneApproxDistance() == pb.getMonotoneApproxDistance()
                                      && (pa.getTargetIndex() != pb.getTargetIndex()
                                          || (pa.getQueryPoint() == pb.getQueryPoint() && pa.getTargetPoint() == pb.getTargetPoint()))));
      triq_ok += (same && same2);
      EXPECT(same && same2, "triangle query %d: triangles %g vs %g, points %g vs %g", i, ta.getMonotoneApproxDistance(), tb.getMonotoneApproxDistance(),
             pa.getMonotoneApproxDistance(), pb.getMonotoneApproxDistance());
    }
    {
      // a second mesh: every 5th triangle, shrunk about its centroid and pushed off the surface along its normal
      for (Real off : { 0.03f, 0.004f })
      {
        std::vector<Triangle3> other;
        for (int64_t t = 0; t < T; t += 5)
        {
          Vector3 cen = tris[(size_t)t].getCentroid(), n = tris[(size_t)t].getNormal();
          Vector3 shift = (off * diag) * n + Vector3(0.0007f, 0.0003f, -0.0005f) * diag;
          other.push_back(Triangle3(cen + 0.8f * (tris[(size_t)t].getVertex(0) - cen) + shift, cen + 0.8f * (tris[(size_t)t].getVertex(1) - cen) + shift,
                                    cen + 0.8f * (tris[(size_t)t].getVertex(2) - cen) + shift));
        }
        RefMeshBvh oref(other.begin(), other.end());
        GpuMeshBvh ogpu;
        ogpu.add([&](auto emit) { for (auto const & x : other) emit(x.getVertex(0), x.getVertex(1), x.getVertex(2)); });
        ogpu.init();
        RefMeshBvh::NeighborPair a = ref.closestPair<MetricL2>(oref, -1, UniversalCompatibility(), true);
        GpuMeshBvh::NeighborPair b = gpu.closestPair<MetricL2>(ogpu, -1, UniversalCompatibility(), true);
        double da = a.getMonotoneApproxDistance(), db = b.getMonotoneApproxDistance();
        // the reference swaps the roles of some pairs depending on its trees, and triangle-triangle rounding is not symmetric: 1e-5
        bool close = a.isValid() && b.isValid() && std::fabs(da - db) <= 1e-5 * std::max(da, db);
        mesh_pair_ok += close; mesh_pair_bit_equal += (da == db);
        EXPECT(close, "closestPair(mesh, mesh, offset %g): %g vs %g", (double)off, da, db);
        // the GPU pair re-evaluated by the reference's own Triangle3::squaredDistance in the non-swapped orientation: bit equal
        Real chk = tris[(size_t)b.getTargetIndex()].squaredDistance(other[(size_t)b.getQueryIndex()]);
        EXPECT((double)chk == db, "closestPair(mesh, mesh): pair value %g re-evaluated as %g", db, (double)chk);
      }
    }

    printf("{\"triangles\": %ld, \"sdf_equal\": %d, \"sdf_samples\": %ld, \"sdf_position_equal\": %d, \"sdf_position_samples\": %ld, \"icp_max_transform_diff\": %g, "
           "\"range_sets_equal\": %d, \"nn_equal\": %d, \"knn_equal\": %d, \"queries\": %d, \"closest_pairs_equal\": %d, \"segment_queries_equal\": %d, \"segment_queries\": %d, "
           "\"triangle_queries_equal\": %d, \"triangle_queries\": %d, \"mesh_pairs_within_1e-5\": %d, \"mesh_pairs_bit_equal\": %d, \"fails\": %d}\n",
           (long)T, sdf_equal, (long)S, sdf_pos_equal, (long)(S / 4), tdiff, range_ok, nn_equal, knn_equal, nq, pair_ok, seg_ok, nseg, triq_ok, ntriq, mesh_pair_ok, mesh_pair_bit_equal, fails);
  }
  catch (std::exception const & e) { fprintf(stderr, "exception: %s\n", e.what()); return 4; }
  return fails ? 3 : 0;
}
