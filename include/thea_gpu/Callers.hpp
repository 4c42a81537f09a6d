// Callers.hpp -- batched rewrites of the reference's serial query loops, on top of GpuBvh.hpp.
//
//   Icp3Gpu::align            <- Algorithms::Icp3<Real>::align(n, from, to_structure, &error)   (Algorithms/Icp3.hpp:110-126,193-275)
//   shapeDiameter             <- the per-sample loop over ShapeDiameter::compute                (Tools/MeshLocalFeatures/MeshLocalFeatures.cpp:747-764,
//                                                                                                Algorithms/SurfaceFeatures/Local/ShapeDiameter.hpp:168-258)
//   sampleGraph               <- PointSet3::updateGraph without oversampling / reachability     (Algorithms/PointSet3.hpp:324-391)
//
// Each loop body of the reference issues one structure query per iteration; here one C-ABI call answers the whole loop
// and the O(1) epilogue (3x3 polar decomposition, transform composition, convergence test) stays on the host, as
// SURVEY.md section 2 row 7 scopes it.  File:line citations are relative to /root/reference/Code/Source/.
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

namespace detail {

/** R = U V^T of the 3x3 matrix M (row-major) by one-sided Jacobi rotations in double: the rotation that
 *  Eigen::JacobiSVD gives Icp3::alignOneStep (Icp3.hpp:369-370), unique whenever M has rank >= 2. */
inline void polarRotation(double const M[9], double R[9])
{
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) A[i] = M[i];
  for (int sweep = 0; sweep < 64; ++sweep)
  {
    double worst = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q)
      {
        double app = 0, aqq = 0, apq = 0;
        for (int r = 0; r < 3; ++r) { app += A[3 * r + p] * A[3 * r + p]; aqq += A[3 * r + q] * A[3 * r + q]; apq += A[3 * r + p] * A[3 * r + q]; }
        double denom = std::sqrt(app * aqq);
        if (denom > 0) worst = std::max(worst, std::fabs(apq) / denom);
        if (apq == 0) continue;
        double zeta = (aqq - app) / (2.0 * apq);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int r = 0; r < 3; ++r)
        {
          double x = A[3 * r + p], y = A[3 * r + q]; A[3 * r + p] = c * x - s * y; A[3 * r + q] = s * x + c * y;
          x = V[3 * r + p]; y = V[3 * r + q]; V[3 * r + p] = c * x - s * y; V[3 * r + q] = s * x + c * y;
        }
      }
    if (worst < 1e-15) break;
  }
  double U[9], sv[3];
  for (int c = 0; c < 3; ++c)
  {
    sv[c] = std::sqrt(A[c] * A[c] + A[3 + c] * A[3 + c] + A[6 + c] * A[6 + c]);
    for (int r = 0; r < 3; ++r) U[3 * r + c] = sv[c] > 0 ? A[3 * r + c] / sv[c] : 0.0;
  }
  int mn = 0; for (int c = 1; c < 3; ++c) if (sv[c] < sv[mn]) mn = c;
  if (sv[mn] <= 1e-14 * std::max(sv[0], std::max(sv[1], sv[2])))
  {
    // rank 2: complete the missing left singular vector so that det(U) has the sign of det(V) (proper rotation U V^T)
    int a = (mn + 1) % 3, b = (mn + 2) % 3;
    U[mn] = U[3 + a] * U[6 + b] - U[6 + a] * U[3 + b];
    U[3 + mn] = U[6 + a] * U[b] - U[a] * U[6 + b];
    U[6 + mn] = U[a] * U[3 + b] - U[3 + a] * U[b];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
    { double acc = 0; for (int k = 0; k < 3; ++k) acc += U[3 * r + k] * V[3 * c + k]; R[3 * r + c] = acc; }
}

}  // namespace detail

/** Point-to-mesh ICP with the control flow of Icp3<Real>::align (Icp3.hpp:193-275; no weights, symmetry planes or up
 *  vector).  Every batch of closest-point correspondences is one tgq_closest call, every reduction one tgq_icp_sums
 *  call (fixed-point sums: the same bits for any launch shape or GPU count). */
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
      tgq_handle * h = to.handle();
      std::vector<float> from(from3, from3 + 3 * n), to_points(3 * (size_t)n);
      std::vector<int32_t> idx((size_t)n);
      double s[16];

      auto measure = [&](Affine3 const & tr) -> float    // Icp3::measureError, Icp3.hpp:387-464
      {
        to.closestElementBatch(from.data(), n, idx.data(), nullptr, to_points.data());
        for (intx i = 0; i < n; ++i)
          if (idx[(size_t)i] < 0) throw Error("Icp3: Couldn't get nearest neighbor of source point " + std::to_string(i));
        if (correspondences) correspondences->insert(correspondences->end(), idx.begin(), idx.end());
        if (batches) ++*batches;
        detail::check(h, tgq_icp_sums(h, from.data(), to_points.data(), n, tr.m, s), "tgq_icp_sums");
        return (float)s[15];
      };

      Affine3 old_tr = Affine3::identity();
      float old_error = measure(old_tr);
      if (verbose) std::fprintf(stderr, "[Icp3Gpu] Initial error: %g\n", (double)old_error);
      if (old_error <= std::numeric_limits<float>::min()) { if (error) *error = old_error; return old_tr; }

      float new_error = old_error; Affine3 new_tr = old_tr;
      for (intx i = 0; i < max_iterations; ++i)
      {
        // alignOneStep(from_points, to_points) (Icp3.hpp:228-229,298-384): the CURRENT source positions against the targets
        // found by the last measureError, which ran before the positions were updated (:237,260-261) -- reproduced as is
        detail::check(h, tgq_icp_sums(h, from.data(), to_points.data(), n, nullptr, s), "tgq_icp_sums");
        double R[9]; detail::polarRotation(&s[6], R);
        Affine3 inc;
        for (int r = 0; r < 3; ++r)
        {
          for (int c = 0; c < 3; ++c) inc.m[4 * r + c] = (float)R[3 * r + c];
          float lp = inc.m[4 * r] * (float)s[0] + (inc.m[4 * r + 1] * (float)s[1] + inc.m[4 * r + 2] * (float)s[2]);
          inc.m[4 * r + 3] = (float)s[3 + r] - lp;
        }
        new_tr = inc * old_tr;
        if (i < max_iterations - 1 || error) new_error = measure(inc);
        if (i < max_iterations - 1)
        {
          float frac_change = (old_error - new_error) / old_error;
          if (verbose) std::fprintf(stderr, "[Icp3Gpu] Iteration %ld error: %g (fractional change: %g)\n", (long)i, (double)new_error, (double)frac_change);
          if (i >= min_iterations && frac_change < fractional_error_threshold)
          {
            if (frac_change > 0) { if (error) *error = new_error; return new_tr; }
            if (error) *error = old_error;
            return old_tr;
          }
          // transform the ORIGINAL points to prevent drift (Icp3.hpp:260-261)
          detail::check(h, tgq_transform_points(h, from3, n, new_tr.m, from.data()), "tgq_transform_points");
        }
        old_error = new_error; old_tr = new_tr;
      }
      if (error) *error = new_error;
      return new_tr;
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
