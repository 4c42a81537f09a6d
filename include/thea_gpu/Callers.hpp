// Callers.hpp -- batched rewrites of the reference's serial query loops, on top of GpuBvh.hpp.
//
//   Icp3Gpu::align            <- Algorithms::Icp3<Real>::align(n, from, to_structure, &error)   (Algorithms/Icp3.hpp:110-126,193-275)
//   shapeDiameter             <- the per-sample loop over ShapeDiameter::compute                (Tools/MeshLocalFeatures/MeshLocalFeatures.cpp:747-764,
//                                                                                                Algorithms/SurfaceFeatures/Local/ShapeDiameter.hpp:168-258)
//   sampleGraph               <- PointSet3::updateGraph without oversampling / reachability     (Algorithms/PointSet3.hpp:324-391)
//   saveGraph / loadGraph     <- PointSet3::save / load, graph file                             (Algorithms/PointSet3.cpp:222-354)
//   tJunctionCandidates       <- the query loop of MeshFix's T-junction repair                  (Tools/MeshFix/MeshFix.cpp:699-790)
//
// Each loop body of the reference issues one structure query per iteration; here one C-ABI call answers the whole loop
// and the O(1) epilogue (3x3 polar decomposition, transform composition, convergence test) stays on the host (inside the
// library for ICP), as SURVEY.md section 2 row 7 scopes it.  File:line citations are relative to /root/reference/Code/Source/.
#ifndef THEA_GPU_CALLERS_HPP
#define THEA_GPU_CALLERS_HPP

#include "GpuBvh.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <limits>
#include <vector>

namespace TheaGpu {

/** Row-major 3 x 4 [L | t]: what AffineTransform3 holds (AffineTransformN.hpp:38-60). */
struct Affine3
{
  float m[12];
  static Affine3 identity() { Affine3 a; for (int i = 0; i < 12; ++i) a.m[i] = (i % 5 == 0) ? 1.0f : 0.0f; return a; }
  /** this * o with Eigen's row . column order t0 + (t1 + t2) (AffineTransformN.hpp:154-157). */
  Affine3 operator*(Affine3 const & o) const
  {
    Affine3 r;
    for (int i = 0; i < 3; ++i)
    {
      for (int j = 0; j < 3; ++j) r.m[4 * i + j] = m[4 * i] * o.m[j] + (m[4 * i + 1] * o.m[4 + j] + m[4 * i + 2] * o.m[8 + j]);
      r.m[4 * i + 3] = (m[4 * i] * o.m[3] + (m[4 * i + 1] * o.m[7] + m[4 * i + 2] * o.m[11])) + m[4 * i + 3];
    }
    return r;
  }
};

/** Point-to-mesh ICP with the control flow of Icp3<Real>::align (Icp3.hpp:193-275; no weights, symmetry planes or up
 *  vector), run device resident by tgq_icp_align (fixed-point sums: the same bits for any launch shape). */
class Icp3Gpu
{
  public:
    Icp3Gpu(double fractional_error_threshold_ = -1, intx min_iterations_ = -1, intx max_iterations_ = -1, bool verbose_ = false)
    : fractional_error_threshold(fractional_error_threshold_), min_iterations(min_iterations_), max_iterations(max_iterations_), verbose(verbose_)
    {
      if (fractional_error_threshold < 0) fractional_error_threshold = 0.0001;   // Icp3.hpp:61-63
      if (min_iterations < 0) min_iterations = 3;
      if (max_iterations < 0) max_iterations = 10;
    }

    /** from3: n x 3 host floats.  Optional: error, number of query batches run, per-batch correspondences
     *  (batches x n element indices appended). */
    template <typename VectorT>
    Affine3 align(intx n, float const * from3, GpuMeshBvhT<VectorT> const & to, float * error = nullptr, intx * batches = nullptr,
                  std::vector<int32_t> * correspondences = nullptr) const
    {
      if (batches) *batches = 0;
      if (n <= 0) { if (error) *error = 0; return Affine3::identity(); }
      // The whole loop runs inside the library, device resident (tgq_icp_align): per iteration one closest-point batch, two
      // reductions and one transform on the GPU, 16 sums back to the host for the 3x3 polar decomposition.
      tgq_handle * h = to.handle();
      Affine3 tr; float err = 0; int32_t nb = 0;
      std::vector<int32_t> corr;
      if (correspondences) corr.resize((size_t)(max_iterations + 1) * (size_t)n);
      detail::check(h, tgq_icp_align(h, from3, n, (int32_t)min_iterations, (int32_t)max_iterations, fractional_error_threshold, error ? 1 : 0,
                                     tr.m, &err, &nb, correspondences ? corr.data() : nullptr), "tgq_icp_align");
      if (verbose) std::fprintf(stderr, "[Icp3Gpu] %d correspondence batches, error %g\n", (int)nb, (double)err);
      if (error) *error = err;
      if (batches) *batches = nb;
      if (correspondences) correspondences->insert(correspondences->end(), corr.begin(), corr.begin() + (std::ptrdiff_t)((size_t)nb * (size_t)n));
      return tr;
    }

  private:
    double fractional_error_threshold; intx min_iterations, max_iterations; bool verbose;
};

/** SDF of S oriented samples: compute(p, n, true) and, like MeshLocalFeatures.cpp:752-756, a second pass with
 *  only_hit_interior_surfaces = false for the samples where the first gave a negative value. */
template <typename VectorT>
inline void shapeDiameter(GpuMeshBvhT<VectorT> const & bvh, float const * p3, float const * n3, intx S, float scale, double * sdf)
{
  bvh.shapeDiameterBatch(p3, n3, S, scale, true, sdf);
  std::vector<float> rp, rn; std::vector<intx> which;
  for (intx i = 0; i < S; ++i)
    if (sdf[i] < 0) { which.push_back(i); rp.insert(rp.end(), p3 + 3 * i, p3 + 3 * i + 3); rn.insert(rn.end(), n3 + 3 * i, n3 + 3 * i + 3); }
  if (which.empty()) return;
  std::vector<double> again(which.size());
  bvh.shapeDiameterBatch(rp.data(), rn.data(), (intx)which.size(), scale, false, again.data());
  for (size_t j = 0; j < which.size(); ++j) sdf[which[j]] = again[j];
}

/** Proximity graph of PointSet3::updateGraph (PointSet3.hpp:324-391) over the points of `bvh`: average separation
 *  from `seeds` (the reference draws 100 from Random::common(), PointSet3.hpp:363-376), radius sqrt(max_degree) * sep,
 *  neighbour search for every sample.  Returns the average separation used. */
template <typename VectorT>
inline float sampleGraph(GpuPointBvhT<VectorT> const & bvh, int32_t max_degree, int32_t const * seeds, int32_t num_seeds, int32_t * idx,
                         float * sep, int32_t * count = nullptr)
{
  float avg = bvh.averageSeparation(seeds, num_seeds);
  float sep_scale = std::sqrt((float)max_degree);
  bvh.sampleGraphBatch(max_degree, sep_scale * avg, idx, sep, count);
  return avg;
}

/** PointSet3::updateGraph(surface): the same with the reachability test against the mesh the samples came from. */
template <typename VectorT>
inline float sampleGraph(GpuPointBvhT<VectorT> const & bvh, GpuMeshBvhT<VectorT> const & surface, int32_t max_degree, int32_t const * seeds,
                         int32_t num_seeds, int32_t * idx, float * sep, int32_t * count = nullptr)
{
  float avg = bvh.averageSeparation(seeds, num_seeds);
  float sep_scale = std::sqrt((float)max_degree);
  bvh.sampleGraphBatch(surface, max_degree, sep_scale * avg, idx, sep, count);
  return avg;
}

/** PointSet3::updateGraph with a dense oversampling (PointSet3.hpp:329-386, PointSet3.cpp:96-178): `all` is a structure over the n_orig
 *  original samples followed by the dense ones (the reference's aggregated BVH); the proximity graph over all of them is built on the
 *  device and contracted to the originals by shortest paths (extractOriginalAdjacencies).  idx / sep / count: n_orig x max_degree
 *  rows of the contracted graph.  seeds: sample indices for the average-separation estimate (see sampleGraph).  Returns it. */
template <typename VectorT>
float sampleGraphOversampled(GpuPointBvhT<VectorT> const & all, intx n_orig, int32_t max_degree, int32_t const * seeds, int32_t num_seeds,
                             int32_t * idx, float * sep, int32_t * count = nullptr)
{
  intx P = all.numElements();
  std::vector<int32_t> gi((size_t)P * (size_t)max_degree); std::vector<float> gs((size_t)P * (size_t)max_degree);
  float avg = sampleGraph(all, max_degree, seeds, num_seeds, gi.data(), gs.data(), nullptr);
  all.graphContractBatch(gi.data(), gs.data(), P, max_degree, n_orig, idx, sep, count);
  return avg;
}

/** FurthestPointSampling::subsample(num_orig_points, orig_points, num_desired_points, selected_indices, GEODESIC)
 *  (Algorithms/FurthestPointSampling.cpp:48-125): proximity graph over the points (PointSet3's defaults: max_degree 8), then sample 0
 *  and repeatedly the sample furthest from the selected set along the graph.  seeds fix the average-separation estimate the reference
 *  draws from Random::common().  Returns the number of points selected (fewer than desired when the reference would give up). */
inline intx furthestPointSubsample(intx num_orig_points, float const * orig_points3, intx num_desired_points, int32_t * selected_indices,
                                   int32_t const * seeds, int32_t num_seeds, int32_t max_degree = 8, int device = 0)
{
  if (num_desired_points < 0) throw Error("FurthestPointSampling: Can't sample a negative number of points");
  if (num_orig_points < num_desired_points) throw Error("FurthestPointSampling: Can't subsample more points than there are");
  if (num_desired_points == 0) return 0;
  GpuPointBvh surf(device); surf.init(orig_points3, num_orig_points);
  std::vector<int32_t> gi((size_t)num_orig_points * (size_t)max_degree); std::vector<float> gs(gi.size());
  sampleGraph(surf, max_degree, seeds, num_seeds, gi.data(), gs.data(), nullptr);
  return surf.graphFurthestPointsBatch(gi.data(), gs.data(), num_orig_points, max_degree, num_desired_points, selected_indices);
}

/** MeshSampler::sampleEvenlyBySeparation (Algorithms/MeshSampler.hpp:270-329): area oversampling (default factor 3, EXACT count) on
 *  the device, furthest point subsampling of it, gather.  p3 / n3 / tri receive desired_num_samples entries (n3, tri may be null).
 *  seed: the counter-based stream of the area sampler; graph_seeds as in furthestPointSubsample. */
template <typename VectorT>
intx sampleEvenlyBySeparation(GpuMeshBvhT<VectorT> const & bvh, intx desired_num_samples, uint64_t seed, int32_t const * graph_seeds, int32_t num_graph_seeds,
                              float * p3, float * n3 = nullptr, int32_t * tri = nullptr, float oversampling_factor = -1)
{
  if (desired_num_samples <= 0) return 0;
  if (oversampling_factor < 0) oversampling_factor = 3;
  intx num_oversampling = (intx)std::ceil(oversampling_factor * desired_num_samples);
  std::vector<float> op(3 * (size_t)num_oversampling), on(3 * (size_t)num_oversampling); std::vector<int32_t> ot((size_t)num_oversampling);
  bvh.sampleEvenlyByArea(num_oversampling, seed, op.data(), on.data(), ot.data());
  std::vector<int32_t> selected((size_t)desired_num_samples);
  if (furthestPointSubsample(num_oversampling, op.data(), desired_num_samples, selected.data(), graph_seeds, num_graph_seeds, 8, bvh.device()) < desired_num_samples)
    return 0;
  for (intx i = 0; i < desired_num_samples; ++i)
  {
    size_t k = (size_t)selected[(size_t)i];
    for (int a = 0; a < 3; ++a) { p3[3 * i + a] = op[3 * k + a]; if (n3) n3[3 * i + a] = on[3 * k + a]; }
    if (tri) tri[i] = ot[k];
  }
  return desired_num_samples;
}

/** PointSet3::save, graph part (PointSet3.cpp:331-354): first line max_degree, then one line per sample,
 *  `degree nbr_0 .. nbr_{d-1} [sep_0 .. sep_{d-1}]`, floats with the default ostream precision as the reference writes them.
 *  idx / sep are P x max_degree as sampleGraph returns them (unused slots -1). */
inline bool saveGraph(std::string const & graph_path, intx P, int32_t max_degree, int32_t const * idx, float const * sep = nullptr,
                      bool write_distances = false)
{
  std::ofstream gout(graph_path.c_str(), std::ios::binary);
  if (!gout) return false;
  gout << max_degree << '\n';
  for (intx i = 0; i < P; ++i)
  {
    int32_t d = 0; while (d < max_degree && idx[i * max_degree + d] >= 0) ++d;
    gout << d;
    for (int32_t j = 0; j < d; ++j) gout << ' ' << idx[i * max_degree + j];
    if (write_distances && sep) for (int32_t j = 0; j < d; ++j) gout << ' ' << sep[i * max_degree + j];
    gout << '\n';
  }
  return (bool)gout;
}

/** PointSet3::load, graph part (PointSet3.cpp:222-321) for P samples: fills max_degree, idx and sep (P x max_degree, -1 padded),
 *  neighbours ordered by (separation, index) as BoundedSortedArray::insert leaves them.  Without stored separations they are the
 *  Euclidean distances of positions3 (PointSet3.cpp:304-311; positions3 may be null: separations stay -1).  Returns false with
 *  a message in *error on malformed input.  The reference re-inserts all max_degree slots of its scratch array, including
 *  stale ones from earlier lines when a sample has fewer neighbours (PointSet3.cpp:296-315); that is not reproduced. */
inline bool loadGraph(std::string const & graph_path, intx P, float const * positions3, int32_t & max_degree, std::vector<int32_t> & idx,
                      std::vector<float> & sep, std::string * error = nullptr)
{
  auto fail = [&](std::string const & m) { if (error) *error = "PointSet3: " + m + " '" + graph_path + "'"; return false; };
  std::ifstream gin(graph_path.c_str());
  if (!gin) return fail("Could not open graph file for reading");
  std::string line; long long md = -1;
  if (!std::getline(gin, line)) return fail("Could not read maximum degree from");
  { std::istringstream in(line); if (!(in >> md) || md < 0) return fail("Could not read valid maximum degree from"); }
  max_degree = (int32_t)md;
  idx.assign((size_t)P * (size_t)md, -1); sep.assign((size_t)P * (size_t)md, -1.0f);
  std::vector<std::pair<float, long long>> nb;
  for (intx i = 0; i < P; ++i)
  {
    if (!std::getline(gin, line)) return fail("Could not read neighbors of vertex " + std::to_string(i) + " from");
    std::istringstream in(line); long long d = -1;
    if (!(in >> d) || d < 0) return fail("Could not read valid degree of vertex " + std::to_string(i) + " from");
    nb.assign((size_t)d, std::make_pair(-1.0f, -1LL));
    for (long long j = 0; j < d; ++j)
      if (!(in >> nb[(size_t)j].second) || nb[(size_t)j].second < 0 || nb[(size_t)j].second >= P)
        return fail("Could not read valid neighbor " + std::to_string(j) + " of vertex " + std::to_string(i) + " from");
    float s;
    if (d > 0 && (in >> s))
    {
      nb[0].first = s;
      for (long long j = 1; j < d; ++j)
      {
        if (!(in >> s)) return fail("Could not read separation of neighbor " + std::to_string(j) + " of vertex " + std::to_string(i) + " from");
        nb[(size_t)j].first = s;
      }
    }
    else if (positions3)
      for (long long j = 0; j < d; ++j)
      {
        float const * a = positions3 + 3 * i; float const * b = positions3 + 3 * nb[(size_t)j].second;
        float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        nb[(size_t)j].first = std::sqrt(dx * dx + (dy * dy + dz * dz));
      }
    std::sort(nb.begin(), nb.end());
    for (long long j = 0; j < d && j < md; ++j) { idx[(size_t)(i * md + j)] = (int32_t)nb[(size_t)j].second; sep[(size_t)(i * md + j)] = nb[(size_t)j].first; }
  }
  return true;
}

/** The query half of MeshFix's T-junction repair (Tools/MeshFix/MeshFix.cpp:716-790) for E boundary edges at once.  `boundary_vertices`
 *  is the structure over the boundary vertices (tJuncts builds a BvhN<Mesh::Vertex *, 3> over them, :729-730).  Per edge e with endpoints
 *  p0_e, p1_e: the segment shrunk by `tol` at both ends (segFromEdge, :699-713: s = tol / len in double, t = 1 - s, ends
 *  (1 - (Real)s) p0 + (Real)s p1 and (1 - (Real)t) p0 + (Real)t p1; edges shorter than 2 tol are skipped), one closestElement query
 *  with dist_bound = 2 tol that ignores the `n_exclude` vertices listed for the edge (VertexFilter, :678-697: the edge's endpoints and the
 *  other vertices of its faces, as positions in the structure, -1 = unused), and the acceptance test of :766-771 on the closest point
 *  cp of the segment to the vertex found: |cp - p0|^2 > tol^2, |cp - p1|^2 > tol^2, |cp - vertex|^2 < tol^2.  split_vertex[e] = the vertex
 *  to split edge e at, or -1; the splitting itself (mesh.splitEdge, :773) and the re-queries of later iterations stay with the caller. */
template <typename PointBvhT>
void tJunctionCandidates(PointBvhT const & boundary_vertices, float const * p0, float const * p1, intx E, double tol, int32_t const * exclude, int32_t n_exclude,
                         int32_t * split_vertex)
{
  std::vector<float> a((size_t)(3 * E)), b((size_t)(3 * E)), d2((size_t)E), vp((size_t)(3 * E)), sp((size_t)(3 * E));
  std::vector<int32_t> idx((size_t)E);
  std::vector<char> skip((size_t)E, 0);
  for (intx e = 0; e < E; ++e)
  {
    float const * u = p0 + 3 * e; float const * v = p1 + 3 * e;
    float dx = v[0] - u[0], dy = v[1] - u[1], dz = v[2] - u[2];
    float len = std::sqrt(dx * dx + (dy * dy + dz * dz));                 // LineSegmentN::length = direction.norm()
    if (len < 2 * tol) { skip[(size_t)e] = 1; for (int c = 0; c < 3; ++c) a[(size_t)(3 * e + c)] = b[(size_t)(3 * e + c)] = u[c]; continue; }
    double s = tol / len, t = 1 - s;
    for (int c = 0; c < 3; ++c)
    {
      a[(size_t)(3 * e + c)] = (1 - (float)s) * u[c] + (float)s * v[c];
      b[(size_t)(3 * e + c)] = (1 - (float)t) * u[c] + (float)t * v[c];
    }
  }
  boundary_vertices.closestElementSegmentBatch(a.data(), b.data(), E, idx.data(), d2.data(), vp.data(), sp.data(), 2 * tol, exclude, n_exclude);
  double sqtol = tol * tol;
  for (intx e = 0; e < E; ++e)
  {
    split_vertex[e] = -1;
    if (skip[(size_t)e] || idx[(size_t)e] < 0) continue;
    float const * cp = &sp[(size_t)(3 * e)]; float const * vx = &vp[(size_t)(3 * e)];
    auto sq = [](float const * x, float const * y) { float ex = x[0] - y[0], ey = x[1] - y[1], ez = x[2] - y[2]; return ex * ex + (ey * ey + ez * ez); };
    if (sq(cp, p0 + 3 * e) > sqtol && sq(cp, p1 + 3 * e) > sqtol && sq(cp, vx) < sqtol) split_vertex[e] = idx[(size_t)e];
  }
}

}  // namespace TheaGpu

#endif
