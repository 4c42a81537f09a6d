// Callers.hpp -- batched rewrites of the reference's serial query loops, on top of GpuBvh.hpp.
//
//   Icp3Gpu::align            <- Algorithms::Icp3<Real>::align(n, from, to_structure, &error)   (Algorithms/Icp3.hpp:110-126,193-275)
//   shapeDiameter             <- the per-sample loop over ShapeDiameter::compute                (Tools/MeshLocalFeatures/MeshLocalFeatures.cpp:747-764,
//                                                                                                Algorithms/SurfaceFeatures/Local/ShapeDiameter.hpp:168-258)
//   sampleGraph               <- PointSet3::updateGraph without oversampling / reachability     (Algorithms/PointSet3.hpp:324-391)
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

}  // namespace TheaGpu

#endif
