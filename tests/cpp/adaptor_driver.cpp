// adaptor_driver.cpp -- TEST-ONLY program: exercises include/thea_gpu/GpuBvh.hpp and Callers.hpp the way a Thea caller
// would (member templates with MetricL2 / RayIntersectionTester tags, scalar and batch forms) and dumps the results for
// tests/test_adaptor.py to compare with the oracle.  Usage: adaptor_driver <in.bin> <out.bin>
#include "../../include/thea_gpu/Callers.hpp"
#include "../../include/thea_gpu/LocalFeatures.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace TheaGpu;

template <typename T> static std::vector<T> rd(FILE * f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }
template <typename T> static void wr(FILE * f, std::vector<T> const & v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }

// compile check of the T-junction caller (exercised on the GPU through its Python twin, tests/test_gpu_segments.py)
static auto const t_junction_caller = &TheaGpu::tJunctionCandidates<TheaGpu::GpuPointBvh>;
static auto const random_walks_caller = &TheaGpu::SurfaceFeatures::Local::RandomWalks::computeBatch;

int main(int argc, char ** argv)
{
  (void)t_junction_caller; (void)random_walks_caller;
  if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  try
  {
    FILE * fi = fopen(argv[1], "rb"); if (!fi) return 2;
    std::vector<int64_t> hdr = rd<int64_t>(fi, 5);
    int64_t T = hdr[0], Q = hdr[1], R = hdr[2], S = hdr[3], NS = hdr[4];
    std::vector<float> v9 = rd<float>(fi, 9 * T), q3 = rd<float>(fi, 3 * Q), o3 = rd<float>(fi, 3 * R), d3 = rd<float>(fi, 3 * R),
                       src = rd<float>(fi, 3 * S), sp = rd<float>(fi, 3 * NS), sn = rd<float>(fi, 3 * NS);
    fclose(fi);
    FILE * fo = fopen(argv[2], "wb"); if (!fo) return 2;

    // ---- MeshBvh-style build: add triangles one by one, init() ----
    GpuMeshBvh bvh;
    for (int64_t t = 0; t < T; ++t)
      bvh.addTriangle(Vec3(v9[9 * t], v9[9 * t + 1], v9[9 * t + 2]), Vec3(v9[9 * t + 3], v9[9 * t + 4], v9[9 * t + 5]),
                      Vec3(v9[9 * t + 6], v9[9 * t + 7], v9[9 * t + 8]));
    bvh.init();
    bvh.enableNearestNeighborAcceleration();
    if (bvh.numElements() != T || bvh.isEmpty()) { fprintf(stderr, "numElements mismatch\n"); return 3; }

    // ---- closestElement: batch, then the reference's scalar form for the first few ----
    std::vector<int32_t> ci(Q); std::vector<float> cd(Q), cc(3 * Q);
    bvh.closestElementBatch(q3.data(), Q, ci.data(), cd.data(), cc.data());
    wr(fo, ci); wr(fo, cd); wr(fo, cc);
    int64_t ns = Q < 16 ? Q : 16;
    std::vector<int64_t> si(ns); std::vector<double> sd(ns), sdist(ns); std::vector<float> sc(3 * ns);
    for (int64_t i = 0; i < ns; ++i)
    {
      Vec3 q(q3[3 * i], q3[3 * i + 1], q3[3 * i + 2]), cp; double dist = -1;
      si[i] = bvh.closestElement<MetricL2>(q, -1, UniversalCompatibility(), &dist, &cp);
      sd[i] = dist; sc[3 * i] = cp[0]; sc[3 * i + 1] = cp[1]; sc[3 * i + 2] = cp[2];
      sdist[i] = bvh.distance<MetricL2>(q);
      GpuMeshBvh::NeighborPair pair = bvh.closestPair<MetricL2>(q, -1, UniversalCompatibility(), true);
      if (pair.getTargetIndex() != si[i] || pair.getDistance<MetricL2>() != dist) { fprintf(stderr, "closestPair disagrees\n"); return 3; }
    }
    wr(fo, si); wr(fo, sd); wr(fo, sdist); wr(fo, sc);
    // dist_bound smaller than the true distance: not found, in band (-1), no exception
    if (ns > 0)
    {
      Vec3 q(q3[0], q3[1], q3[2]);
      if (bvh.closestElement<MetricL2>(q, sd[0] * 0.5) != -1 || bvh.distance<MetricL2>(q, sd[0] * 0.5) != -1) { fprintf(stderr, "dist_bound not honoured\n"); return 3; }
    }

    // ---- rays ----
    std::vector<float> rt(R), rn(3 * R); std::vector<int32_t> ri(R); std::vector<uint8_t> rh(R);
    bvh.rayStructureIntersectionBatch(o3.data(), d3.data(), nullptr, R, rt.data(), ri.data(), rn.data());
    bvh.rayIntersectsBatch(o3.data(), d3.data(), nullptr, R, rh.data());
    wr(fo, rt); wr(fo, ri); wr(fo, rn); wr(fo, rh);
    int64_t nr = R < 16 ? R : 16;
    std::vector<float> st(nr), stime(nr), snrm(3 * nr); std::vector<int64_t> sidx(nr); std::vector<uint8_t> shit(nr);
    for (int64_t i = 0; i < nr; ++i)
    {
      GpuMeshBvh::Ray3 ray(Vec3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), Vec3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]));
      GpuMeshBvh::RayStructureIntersection3 isec = bvh.rayStructureIntersection<RayIntersectionTester>(ray);
      st[i] = isec.getTime(); sidx[i] = isec.getElementIndex();
      for (int c = 0; c < 3; ++c) snrm[3 * i + c] = isec.getNormal()[c];
      stime[i] = bvh.rayIntersectionTime<RayIntersectionTester>(ray);
      shit[i] = bvh.rayIntersects<RayIntersectionTester>(ray) ? 1 : 0;
    }
    wr(fo, st); wr(fo, sidx); wr(fo, snrm); wr(fo, stime); wr(fo, shit);

    // ---- ICP (Icp3 defaults: 3..10 iterations, 1e-4) ----
    Icp3Gpu icp;
    float err = -1; intx batches = 0; std::vector<int32_t> corr;
    Affine3 tr = icp.align(S, src.data(), bvh, &err, &batches, &corr);
    std::vector<float> trv(tr.m, tr.m + 12); trv.push_back(err);
    std::vector<int64_t> nb(1, batches);
    wr(fo, trv); wr(fo, nb); wr(fo, corr);

    // ---- SDF with the retry pass ----
    float lo[3], hi[3]; { Vec3 l, u; bvh.getBounds(l, u); for (int c = 0; c < 3; ++c) { lo[c] = l[c]; hi[c] = u[c]; } }
    float diag = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + ((hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2])));
    std::vector<double> sdf(NS);
    shapeDiameter(bvh, sp.data(), sn.data(), NS, diag, sdf.data());
    std::vector<float> dg(1, diag); wr(fo, dg); wr(fo, sdf);

    // ---- point structure: k-NN batch + kClosestPairs into a BoundedSortedArray, sample graph ----
    GpuPointBvh pts;
    pts.init(sp.data(), NS, sn.data());
    int const K = 8;
    std::vector<int32_t> ki(NS * K); std::vector<float> kd(NS * K);
    pts.kClosestPairsBatch(sp.data(), NS, K, ki.data(), kd.data(), -1, true);
    wr(fo, ki); wr(fo, kd);
    int64_t nk = NS < 4 ? NS : 4;
    std::vector<int64_t> ski(nk * 3, -1); std::vector<double> skd(nk * 3, -1);
    for (int64_t i = 0; i < nk; ++i)
    {
      BoundedSortedArray<GpuPointBvh::NeighborPair> set(3);
      intx found = pts.kClosestPairs<MetricL2>(Vec3(sp[3 * i], sp[3 * i + 1], sp[3 * i + 2]), set);
      for (intx j = 0; j < found; ++j) { ski[3 * i + j] = set[(size_t)j].getTargetIndex(); skd[3 * i + j] = set[(size_t)j].getMonotoneApproxDistance(); }
    }
    wr(fo, ski); wr(fo, skd);
    std::vector<int32_t> seeds; for (int j = 0; j < 100 && j < NS; ++j) seeds.push_back((int32_t)((j * 7919) % NS));
    std::vector<int32_t> gi(NS * 6), gc(NS); std::vector<float> gs(NS * 6);
    float avg = sampleGraph(pts, 6, seeds.data(), (int32_t)seeds.size(), gi.data(), gs.data(), gc.data());
    std::vector<float> av(1, avg); wr(fo, av); wr(fo, gi); wr(fo, gs); wr(fo, gc);

    // ---- range queries the way Test/TestBvh.cpp:174-197 writes them (ball and box ranges, functor / elements / indices), on a
    //      structure built from a host range so that getElements() is available; checked here against the definition ----
    {
      std::vector<Vec3> pv; for (int64_t i = 0; i < NS; ++i) pv.push_back(Vec3(sp[3 * i], sp[3 * i + 1], sp[3 * i + 2]));
      GpuPointBvh rb; rb.init(pv.begin(), pv.end());
      if ((int64_t)rb.getElements().size() != NS) { fprintf(stderr, "getElements size\n"); return 3; }
      Vec3 c(sp[0], sp[1], sp[2]);
      Ball3T<Vec3> ball(c, 0.25f);
      AxisAlignedBox3T<Vec3> box(Vec3(c[0] - 0.3f, c[1] - 0.2f, c[2] - 0.1f), Vec3(c[0] + 0.1f, c[1] + 0.2f, c[2] + 0.3f));
      std::vector<intx> want_ball, want_box;
      for (int64_t i = 0; i < NS; ++i)
      {
        float dx = pv[i][0] - c[0], dy = pv[i][1] - c[1], dz = pv[i][2] - c[2];
        if (dx * dx + (dy * dy + dz * dz) <= 0.25f * 0.25f) want_ball.push_back(i);
        bool in = true; for (int a = 0; a < 3; ++a) in = in && !(pv[i][a] < box.getLow()[a]) && !(pv[i][a] > box.getHigh()[a]);
        if (in) want_box.push_back(i);
      }
      std::vector<intx> got; rb.rangeQueryIndices<IntersectionTester>(ball, got);
      if (got != want_ball || got.empty()) { fprintf(stderr, "rangeQueryIndices(ball) %zu vs %zu\n", got.size(), want_ball.size()); return 3; }
      rb.rangeQueryIndices<IntersectionTester>(box, got);
      if (got != want_box || got.empty()) { fprintf(stderr, "rangeQueryIndices(box) %zu vs %zu\n", got.size(), want_box.size()); return 3; }
      std::vector<Vec3> elems; rb.rangeQuery<IntersectionTester>(box, elems);
      if (elems.size() != want_box.size() || elems[0][0] != pv[want_box[0]][0]) { fprintf(stderr, "rangeQuery(box)\n"); return 3; }
      size_t visited = 0; intx stop = want_box[want_box.size() / 2];
      intx r = rb.processRange<IntersectionTester>(box, [&](intx index, Vec3 const &) { ++visited; return index == stop; });
      if (r != stop || visited != want_box.size() / 2 + 1) { fprintf(stderr, "processRange early termination\n"); return 3; }
      r = rb.processRange<IntersectionTester>(ball, [&](intx, Vec3 const &) { return false; });
      if (r >= 0) { fprintf(stderr, "processRange full pass\n"); return 3; }
    }
    // getElements() of the triangle structure: vertices as given, normals as the device cached them
    {
      std::vector<float> n3; bvh.getElementNormals(n3);
      auto const & el = bvh.getElements();
      if ((int64_t)el.size() != T || el[T - 1].getVertex(2)[1] != v9[9 * (T - 1) + 7] || el[3].getNormal()[2] != n3[11]) { fprintf(stderr, "getElements (mesh)\n"); return 3; }
      auto bb = bvh.getBounds();
      Vec3 el_lo, el_hi; bvh.getElementBounds(el_lo, el_hi);
      if (!(bb.getExtent()[0] > 0) || !(bb.getLow()[0] < el_lo[0]) || !(bb.getHigh()[2] > el_hi[2])) { fprintf(stderr, "getBounds is not the padded root box\n"); return 3; }
    }

    // ---- filtered nearest neighbours the way Tools/Register/Register.cpp:210-236,262-287 writes them ----
    std::vector<int32_t> labels(NS); for (int64_t i = 0; i < NS; ++i) labels[i] = (int32_t)(i % 3);
    pts.setLabels(labels.data());
    std::vector<int32_t> fci(NS), fki(NS * 4); std::vector<float> fd(NS), fkd(NS * 4);
    pts.closestElementFilteredBatch(sp.data(), sn.data(), labels.data(), NS, fci.data(), fd.data(), nullptr);
    pts.kClosestPairsFilteredBatch(sp.data(), sn.data(), labels.data(), NS, 4, fki.data(), fkd.data());
    wr(fo, fci); wr(fo, fd); wr(fo, fki); wr(fo, fkd);
    for (int64_t i = 0; i < nk; ++i)
    {
      GpuPointBvh::NNFilter filter(Vec3(sn[3 * i], sn[3 * i + 1], sn[3 * i + 2]), labels[i]);
      pts.pushFilter(&filter);
        intx nn = pts.closestElement<MetricL2>(Vec3(sp[3 * i], sp[3 * i + 1], sp[3 * i + 2]));
        BoundedSortedArray<GpuPointBvh::NeighborPair> set(4);
        pts.kClosestPairs<MetricL2>(Vec3(sp[3 * i], sp[3 * i + 1], sp[3 * i + 2]), set);
      pts.popFilter();
      if (nn != fci[i]) { fprintf(stderr, "scalar filtered closestElement disagrees with the batch\n"); return 3; }
      for (size_t j = 0; j < set.size(); ++j)
        if (set[j].getTargetIndex() != fki[4 * i + (int64_t)j]) { fprintf(stderr, "scalar filtered kClosestPairs disagrees with the batch\n"); return 3; }
    }

    // ---- the graph again with the reachability test against the mesh (PointSet3::updateGraph(surface)) ----
    {
      std::vector<int32_t> si_(NS * 6), sc_(NS); std::vector<float> ss_(NS * 6);
      float avg2 = sampleGraph(pts, bvh, 6, seeds.data(), (int32_t)seeds.size(), si_.data(), ss_.data(), sc_.data());
      if (avg2 != avg) { fprintf(stderr, "average separation changed\n"); return 3; }
      wr(fo, si_); wr(fo, ss_); wr(fo, sc_);
    }

    // ---- graph file round trip (PointSet3::save / load format) ----
    {
      std::string gp = std::string(argv[2]) + ".graph";
      if (!saveGraph(gp, NS, 6, gi.data(), gs.data(), true)) { fprintf(stderr, "saveGraph failed\n"); return 3; }
      int32_t md = 0; std::vector<int32_t> li; std::vector<float> ls; std::string emsg;
      if (!loadGraph(gp, NS, sp.data(), md, li, ls, &emsg)) { fprintf(stderr, "loadGraph failed: %s\n", emsg.c_str()); return 3; }
      if (md != 6 || li != gi) { fprintf(stderr, "graph file round trip changed the adjacency\n"); return 3; }
      for (size_t j = 0; j < ls.size(); ++j)
        if (gi[j] >= 0 && std::fabs(ls[j] - gs[j]) > 1e-5f * std::fabs(gs[j])) { fprintf(stderr, "graph file round trip changed a separation\n"); return 3; }
      std::remove(gp.c_str());
    }

    // ---- local surface features (SurfaceFeatures::Local classes): batch, then the reference's scalar form for the first few ----
    {
      using namespace TheaGpu::SurfaceFeatures::Local;
      AverageDistance avgd(&pts, diag); Curvature curv(&pts, diag, sn.data()); LocalPca pca(&pts, diag);
      std::vector<double> fa(NS), fc(NS); std::vector<float> fe(3 * NS), fv(9 * NS);
      avgd.computeBatch(sp.data(), NS, 0.1f * diag, fa.data());
      curv.computeProjectedCurvatureBatch(sp.data(), sn.data(), NS, 0.05f, fc.data());
      pca.computeBatch(sp.data(), NS, 0.05f, fe.data(), fv.data());
      wr(fo, fa); wr(fo, fc); wr(fo, fe); wr(fo, fv);
      for (int64_t i = 0; i < nk; ++i)
      {
        Vec3 p(sp[3 * i], sp[3 * i + 1], sp[3 * i + 2]), n(sn[3 * i], sn[3 * i + 1], sn[3 * i + 2]), ev[3];
        if (avgd.compute(p, 0.1f * diag) != fa[i] || curv.computeProjectedCurvature(p, n, 0.05f) != fc[i])
        { fprintf(stderr, "scalar local feature disagrees with the batch\n"); return 3; }
        if (curv.computeProjectedCurvature(p, 0.05f) != fc[i]) { fprintf(stderr, "curvature with looked-up normal disagrees\n"); return 3; }
        Vec3 e = pca.compute(p, ev, 0.05f);
        for (int k = 0; k < 3; ++k) if (e[k] != fe[3 * i + k] || ev[k][0] != fv[9 * i + 3 * k]) { fprintf(stderr, "scalar LocalPca disagrees with the batch\n"); return 3; }
      }
      LocalDistanceHistogram ldh(&pts, diag);
      std::vector<double> hb(8 * NS);
      ldh.computeBatch(sp.data(), NS, 8, hb.data(), 0.1f * diag);
      wr(fo, hb);
      std::vector<float> dirs; for (int k = 0; k < 3 * 32 && k < (int)d3.size(); ++k) dirs.push_back(d3[k]);
      Visibility visf(&bvh, dirs.data(), (int32_t)(dirs.size() / 3));
      std::vector<double> vis(NS);
      visf.computeBatch(sp.data(), NS, vis.data());
      std::vector<float> vsc(1, visf.getScale()); wr(fo, vsc); wr(fo, vis);
      if (NS > 0 && visf.compute(Vec3(sp[0], sp[1], sp[2])) != vis[0]) { fprintf(stderr, "scalar Visibility disagrees with the batch\n"); return 3; }

      // SpinImage: batch with given normals, then the reference's first overload (normal of the nearest sample: the sample itself)
      SpinImage spin(&pts, diag, sn.data());
      std::vector<double> im(20 * NS);
      spin.computeBatch(sp.data(), sn.data(), NS, 5, 4, im.data());
      wr(fo, im);
      for (int64_t i = 0; i < nk; ++i)
      {
        double one[20]; spin.compute(Vec3(sp[3 * i], sp[3 * i + 1], sp[3 * i + 2]), 5, 4, one);
        for (int k = 0; k < 20; ++k) if (one[k] != im[20 * i + k]) { fprintf(stderr, "scalar SpinImage disagrees with the batch\n"); return 3; }
      }
      // DistanceType::GEODESIC on the sample graph built above (max_degree 6)
      std::vector<double> ga(NS), gh(8 * NS);
      avgd.computeGeodesicBatch(gi.data(), gs.data(), 6, sp.data(), NS, 0.1f * diag, ga.data());
      ldh.computeGeodesicBatch(gi.data(), gs.data(), 6, sp.data(), NS, 8, gh.data(), 0.1f * diag);
      wr(fo, ga); wr(fo, gh);
    }

    // ---- PointSet3::updateGraph with a dense oversampling (the first third of the samples play the originals) and
    //      FurthestPointSampling::subsample ----
    {
      int64_t NO = NS / 3;
      std::vector<int32_t> ci(NO * 6), cc(NO); std::vector<float> cs(NO * 6);
      float avg3 = sampleGraphOversampled(pts, NO, 6, seeds.data(), (int32_t)seeds.size(), ci.data(), cs.data(), cc.data());
      if (avg3 != avg) { fprintf(stderr, "average separation changed (oversampled graph)\n"); return 3; }
      wr(fo, ci); wr(fo, cs); wr(fo, cc);
      std::vector<int32_t> sel(NS / 10, -1);
      std::vector<int64_t> nsel(1, (int64_t)furthestPointSubsample(NS, sp.data(), NS / 10, sel.data(), seeds.data(), (int32_t)seeds.size()));
      wr(fo, nsel); wr(fo, sel);
    }

    // ---- error behaviour: filters are refused loudly, wrong structure kind raises ----
    bool threw = false;
    try { int dummy = 0; bvh.pushFilter(&dummy); } catch (Error const &) { threw = true; }
    if (!threw) { fprintf(stderr, "pushFilter did not throw\n"); return 3; }
    threw = false;
    try { BoundedSortedArray<GpuPointBvh::NeighborPair> s2(2); GpuPointBvh empty; empty.kClosestPairs<MetricL2>(Vec3(0, 0, 0), s2); } catch (Error const &) { threw = true; }
    if (!threw) { fprintf(stderr, "query on an unbuilt structure did not throw\n"); return 3; }
    fclose(fo);
    printf("adaptor driver ok\n");
    return 0;
  }
  catch (std::exception const & e) { fprintf(stderr, "exception: %s\n", e.what()); return 4; }
}
