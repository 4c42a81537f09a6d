// LocalFeatures.hpp -- the local surface features of Thea that are ball range queries with an accumulating functor, or
// any-hit ray fans, on top of GpuBvh.hpp.  Class and member names follow the reference:
//
//   SurfaceFeatures::Local::AverageDistance::compute(position, EUCLIDEAN, max_distance)   (Algorithms/SurfaceFeatures/Local/AverageDistance.cpp:63-81,133-143)
//   SurfaceFeatures::Local::Curvature::computeProjectedCurvature(position[, normal], r)    (…/Local/Curvature.cpp:69-98)
//   SurfaceFeatures::Local::LocalPca::compute(position, eigenvectors, r)                   (…/Local/LocalPca.cpp:32-62, Algorithms/PcaN.hpp:150-197)
//   SurfaceFeatures::Local::Visibility::compute(position, num_rays)                        (…/Local/Visibility.hpp:104-125)
//   SurfaceFeatures::Local::LocalDistanceHistogram::compute(position, histogram, EUCLIDEAN, max_distance)  (…/Local/LocalDistanceHistogram.cpp:36-88)
//
// The reference classes take a PointSet3 (samples + BvhN + scale); here they take the GpuPointBvh built over the same samples
// plus the shape scale PointSet3::getScale() would return (the diameter of its bounding sphere, PointSet3.hpp:560-572), which the
// caller supplies.  Scalar calls are batches of one; *Batch members answer a whole loop in one C-ABI call.
//   DistanceType::GEODESIC of AverageDistance / LocalDistanceHistogram (AverageDistance.cpp:84-118, LocalDistanceHistogram.cpp:90-137):
//     computeGeodesicBatch(graph rows, ...) -- nearest-sample lookup on the structure, then the bounded shortest-path search of
//     tgq_graph_geodesic on the sample graph the caller built with sampleGraph / sampleGraphBatch;
//   SurfaceFeatures::Local::SpinImage::compute(position[, normal], radial bins, height bins, image)   (…/Local/SpinImage.cpp:30-75).
//   SurfaceFeatures::Local::RandomWalks::compute(position, num_steps, features, num_walks)            (…/Local/RandomWalks.cpp:30-117): seed
//     lookup batched on the structure, the walks on the host over the caller's TheaRandom (the reference's shared Random::common()).
// Not covered: random subsampling of the histograms (Random::common()).
// File:line citations are relative to /root/reference/Code/Source/.
#ifndef THEA_GPU_LOCALFEATURES_HPP
#define THEA_GPU_LOCALFEATURES_HPP

#include "GpuBvh.hpp"

#include <algorithm>
#include <cmath>
#include <vector>

namespace TheaGpu {
namespace SurfaceFeatures {
namespace Local {

namespace detail_lf {

/** Eigen decomposition of a symmetric 3x3 matrix (cyclic Jacobi, double): eigenvalues descending, unit eigenvectors as rows.
 *  Stands in for Eigen::SelfAdjointEigenSolver (PcaN.hpp:187-195, third party). */
inline void symmetricEigen3(double const A[9], double evals[3], double evecs[9])
{
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { a[i][j] = A[3 * i + j]; v[i][j] = i == j ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 64; ++sweep)
  {
    double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    double diag = std::fabs(a[0][0]) + std::fabs(a[1][1]) + std::fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-17 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q)
      {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { double x = v[k][p], y = v[k][q]; v[k][p] = c * x - s * y; v[k][q] = s * x + c * y; }
      }
  }
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int i, int j) { return a[i][i] > a[j][j]; });
  for (int r = 0; r < 3; ++r)
  {
    int k = order[r]; evals[r] = a[k][k];
    double n = std::sqrt(v[0][k] * v[0][k] + v[1][k] * v[1][k] + v[2][k] * v[2][k]);
    for (int i = 0; i < 3; ++i) evecs[3 * r + i] = n > 0 ? v[i][k] / n : 0.0;
  }
}

}  // namespace detail_lf

/** AverageDistance (EUCLIDEAN): mean distance from the query to the samples within max_distance, over max_distance. */
template <typename VectorT = Vec3>
class AverageDistanceT
{
  public:
    AverageDistanceT(GpuPointBvhT<VectorT> const * surf_, Real scale_) : surf(surf_), scale(scale_)
    { if (!surf) throw Error("AverageDistance: Cannot construct with a null surface"); }

    /** max_distance < 0 selects the shape scale, as the reference does (AverageDistance.cpp:66-68; there it also switches to a
     *  loop over all samples, which is the same set whenever the scale covers the shape). */
    double compute(VectorT const & position, Real max_distance = -1) const
    {
      float c[3]; detail::put3(c, position); double out = 0;
      computeBatch(c, 1, max_distance, &out);
      return out;
    }
    void computeBatch(float const * p3, intx Q, Real max_distance, double * out) const
    {
      if (max_distance < 0) max_distance = scale;
      std::vector<float> rad((size_t)Q, max_distance); std::vector<double> st(16 * (size_t)Q);
      surf->ballStatsBatch(p3, nullptr, rad.data(), Q, st.data());
      for (intx i = 0; i < Q; ++i)
      {
        double n = st[16 * (size_t)i], sum = st[16 * (size_t)i + 1];
        out[i] = (n <= 0 ? 0.0 : sum / n) / max_distance;
      }
    }

    /** DistanceType::GEODESIC (AverageDistance.cpp:84-118): seed = closestElement(position), then the mean graph distance over the
     *  samples within max_distance of it, over max_distance (< 0: the whole component, over the scale).  idx / sep: the sample graph,
     *  P x max_degree rows of sampleGraphBatch over this structure's samples. */
    void computeGeodesicBatch(int32_t const * idx, float const * sep, int32_t max_degree, float const * p3, intx Q, Real max_distance, double * out) const
    {
      std::vector<int32_t> seed((size_t)Q), cnt((size_t)Q); std::vector<float> d2((size_t)Q); std::vector<double> sum((size_t)Q);
      surf->closestElementBatch(p3, Q, seed.data(), d2.data(), nullptr);
      Real range = max_distance < 0 ? scale : max_distance;
      surf->graphGeodesicBatch(idx, sep, surf->numElements(), max_degree, seed.data(), Q, max_distance, range, 0, sum.data(), cnt.data(), nullptr);
      for (intx i = 0; i < Q; ++i) out[i] = (cnt[(size_t)i] <= 0 ? 0.0 : sum[(size_t)i] / cnt[(size_t)i]) / range;
    }

  private:
    GpuPointBvhT<VectorT> const * surf; Real scale;
};

/** Curvature: projected curvature -(sum of unit offsets to the visible-side samples in the ball) . n / count. */
template <typename VectorT = Vec3>
class CurvatureT
{
  public:
    /** n3: the sample normals the structure was built with (P x 3, host), for the overload that looks the normal up. */
    CurvatureT(GpuPointBvhT<VectorT> const * surf_, Real scale_, float const * sample_normals3 = nullptr)
    : surf(surf_), scale(scale_), normals(sample_normals3)
    { if (!surf) throw Error("Curvature: Cannot construct with a null surface"); }

    /** Curvature.cpp:69-81: normal of the nearest sample; 0 when the query cannot be mapped to the surface. */
    double computeProjectedCurvature(VectorT const & position, Real nbd_radius = -1) const
    {
      if (!normals) throw Error("Curvature: sample normals were not supplied");
      intx nn = surf->template closestElement<MetricL2>(position);
      if (nn < 0) return 0.0;
      return computeProjectedCurvature(position, detail::get3<VectorT>(normals + 3 * nn), nbd_radius);
    }
    double computeProjectedCurvature(VectorT const & position, VectorT const & normal, Real nbd_radius = -1) const
    {
      float c[3], n[3]; detail::put3(c, position); detail::put3(n, normal); double out = 0;
      computeProjectedCurvatureBatch(c, n, 1, nbd_radius, &out);
      return out;
    }
    void computeProjectedCurvatureBatch(float const * p3, float const * n3, intx Q, Real nbd_radius, double * out) const
    {
      if (nbd_radius < 0) nbd_radius = 0.1f;             // Curvature.cpp:88-91
      nbd_radius *= scale;
      std::vector<float> rad((size_t)Q, nbd_radius); std::vector<double> st(16 * (size_t)Q);
      surf->ballStatsBatch(p3, n3, rad.data(), Q, st.data());
      for (intx i = 0; i < Q; ++i)
      {
        double const * s = &st[16 * (size_t)i];
        double proj = s[3] * n3[3 * i] + s[4] * n3[3 * i + 1] + s[5] * n3[3 * i + 2];
        out[i] = s[2] > 0 ? -proj / s[2] : 0.0;
      }
    }

  private:
    GpuPointBvhT<VectorT> const * surf; Real scale; float const * normals;
};

/** LocalPca: eigenvalues (descending) and eigenvectors of the covariance of the samples in the ball. */
template <typename VectorT = Vec3>
class LocalPcaT
{
  public:
    LocalPcaT(GpuPointBvhT<VectorT> const * surf_, Real scale_) : surf(surf_), scale(scale_)
    { if (!surf) throw Error("LocalPca: Cannot construct with a null surface"); }

    VectorT compute(VectorT const & position, VectorT * eigenvectors = nullptr, Real nbd_radius = -1) const
    {
      float c[3]; detail::put3(c, position); float ev[3], evec[9];
      computeBatch(c, 1, nbd_radius, ev, eigenvectors ? evec : nullptr);
      if (eigenvectors) for (int k = 0; k < 3; ++k) eigenvectors[k] = detail::get3<VectorT>(evec + 3 * k);
      return detail::get3<VectorT>(ev);
    }
    /** evals: Q x 3; evecs (may be null): Q x 3 x 3, row k = eigenvector of eigenvalue k.  n < 2 samples give zeros (PcaN.hpp:177-182). */
    void computeBatch(float const * p3, intx Q, Real nbd_radius, float * evals, float * evecs = nullptr) const
    {
      if (nbd_radius < 0) nbd_radius = 0.1f;             // LocalPca.cpp:52-55
      nbd_radius *= scale;
      std::vector<float> rad((size_t)Q, nbd_radius); std::vector<double> st(16 * (size_t)Q);
      surf->ballStatsBatch(p3, nullptr, rad.data(), Q, st.data());
      for (intx i = 0; i < Q; ++i)
      {
        double const * s = &st[16 * (size_t)i];
        double n = s[0], w[3] = {0, 0, 0}, v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (n >= 2)
        {
          double m[3] = {s[6] / n, s[7] / n, s[8] / n};
          double S2[9] = {s[9], s[10], s[11], s[10], s[12], s[13], s[11], s[13], s[14]}, cov[9];
          for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[3 * r + c] = (S2[3 * r + c] - n * m[r] * m[c]) / (n - 1);
          detail_lf::symmetricEigen3(cov, w, v);
        }
        for (int k = 0; k < 3; ++k) evals[3 * i + k] = (float)w[k];
        if (evecs) for (int k = 0; k < 9; ++k) evecs[9 * i + k] = (float)v[k];
      }
    }

  private:
    GpuPointBvhT<VectorT> const * surf; Real scale;
};

/** LocalDistanceHistogram (EUCLIDEAN, every sample accepted): histogram of the distances from the query to the samples within
 *  max_distance, over [0, max_distance].  bins receives num_bins doubles per query, like Histogram::getBins(). */
template <typename VectorT = Vec3>
class LocalDistanceHistogramT
{
  public:
    LocalDistanceHistogramT(GpuPointBvhT<VectorT> const * surf_, Real scale_) : surf(surf_), scale(scale_)
    { if (!surf) throw Error("LocalDistanceHistogram: Cannot construct with a null surface"); }

    void compute(VectorT const & position, intx num_bins, double * bins, Real max_distance = -1) const
    { float c[3]; detail::put3(c, position); computeBatch(c, 1, num_bins, bins, max_distance); }
    /** max_distance < 0 selects the shape scale (LocalDistanceHistogram.cpp:65-67). */
    void computeBatch(float const * p3, intx Q, intx num_bins, double * bins, Real max_distance = -1) const
    {
      if (max_distance < 0) max_distance = scale;
      std::vector<float> rad((size_t)Q, max_distance); std::vector<uint32_t> h((size_t)Q * (size_t)num_bins);
      surf->ballHistogramBatch(p3, rad.data(), Q, (int32_t)num_bins, h.data());
      for (size_t i = 0; i < h.size(); ++i) bins[i] = (double)h[i];
    }
    /** DistanceType::GEODESIC (LocalDistanceHistogram.cpp:90-137): graph distances from the nearest sample, range [0, max_distance]
     *  (< 0: the whole component, range [0, scale]); every sample accepted. */
    void computeGeodesicBatch(int32_t const * idx, float const * sep, int32_t max_degree, float const * p3, intx Q, intx num_bins, double * bins,
                              Real max_distance = -1) const
    {
      std::vector<int32_t> seed((size_t)Q); std::vector<float> d2((size_t)Q); std::vector<uint32_t> h((size_t)Q * (size_t)num_bins);
      surf->closestElementBatch(p3, Q, seed.data(), d2.data(), nullptr);
      Real range = max_distance < 0 ? scale : max_distance;
      surf->graphGeodesicBatch(idx, sep, surf->numElements(), max_degree, seed.data(), Q, max_distance, range, (int32_t)num_bins, nullptr, nullptr, h.data());
      for (size_t i = 0; i < h.size(); ++i) bins[i] = (double)h[i];
    }

  private:
    GpuPointBvhT<VectorT> const * surf; Real scale;
};

/** SpinImage (SpinImage.cpp:30-75): every sample of the surface binned by (radius about the axis through the query, height along it),
 *  limits 0.75f * scale.  image receives num_radial_bins x num_height_bins doubles per query, radial bin major -- the reference's
 *  spin_image(radial_bin, height_bin). */
template <typename VectorT = Vec3>
class SpinImageT
{
  public:
    /** sample_normals3 (P x 3, host, may be null): the samples' normals, needed only by the overloads that infer the query normal. */
    SpinImageT(GpuPointBvhT<VectorT> const * surf_, Real scale_, float const * sample_normals3 = nullptr) : surf(surf_), scale(scale_), sn(sample_normals3)
    { if (!surf) throw Error("SpinImage: Cannot construct with a null surface"); }

    void compute(VectorT const & position, VectorT const & normal, int num_radial_bins, int num_height_bins, double * image) const
    { float p[3], n[3]; detail::put3(p, position); detail::put3(n, normal); computeBatch(p, n, 1, num_radial_bins, num_height_bins, image); }
    /** The reference's first overload: the normal of the sample nearest to the query (SpinImage.cpp:30-45); the surface needs normals. */
    void compute(VectorT const & position, int num_radial_bins, int num_height_bins, double * image) const
    { float p[3]; detail::put3(p, position); computeBatch(p, nullptr, 1, num_radial_bins, num_height_bins, image); }
    /** qn3 == nullptr: normals of the nearest samples (a query that maps to no sample gets a zero image, as in the reference). */
    void computeBatch(float const * p3, float const * qn3, intx Q, int num_radial_bins, int num_height_bins, double * image) const
    {
      if (num_radial_bins <= 0) throw Error("SpinImage: Number of radial bins must be positive");
      if (num_height_bins <= 0) throw Error("SpinImage: Number of height bins must be positive");
      size_t bins = (size_t)num_radial_bins * (size_t)num_height_bins;
      std::vector<float> nn; std::vector<int32_t> seed;
      if (!qn3)
      {
        if (!sn) throw Error("SpinImage: Cannot infer query normal from surface lacking normals");
        seed.resize((size_t)Q); std::vector<float> d2((size_t)Q); nn.assign(3 * (size_t)Q, 0.0f);
        surf->closestElementBatch(p3, Q, seed.data(), d2.data(), nullptr);
        for (intx i = 0; i < Q; ++i) if (seed[(size_t)i] >= 0) for (int a = 0; a < 3; ++a) nn[3 * (size_t)i + a] = sn[3 * (size_t)seed[(size_t)i] + a];
        qn3 = nn.data();
      }
      std::vector<uint32_t> img((size_t)Q * bins);
      surf->spinImageBatch(p3, qn3, Q, num_radial_bins, num_height_bins, 0.75f * scale, img.data());
      for (size_t i = 0; i < img.size(); ++i) image[i] = (double)img[i];
      if (!seed.empty()) for (intx i = 0; i < Q; ++i) if (seed[(size_t)i] < 0) for (size_t b = 0; b < bins; ++b) image[(size_t)i * bins + b] = 0.0;
    }

  private:
    GpuPointBvhT<VectorT> const * surf; Real scale; float const * sn;
};

/** Visibility: fraction of rays from the query point that do not hit the shape again.  The reference draws num_rays random
 *  directions from Random::common() per call; here the caller fixes the directions at construction. */
template <typename VectorT = Vec3>
class VisibilityT
{
  public:
    VisibilityT(GpuMeshBvhT<VectorT> const * bvh_, float const * dirs3, int32_t num_dirs) : bvh(bvh_), dirs(dirs3, dirs3 + 3 * (size_t)num_dirs)
    {
      if (!bvh) throw Error("Visibility: Precomputed BVH cannot be null");
      VectorT lo, hi; bvh->getBounds(lo, hi);
      float e0 = hi[0] - lo[0], e1 = hi[1] - lo[1], e2 = hi[2] - lo[2];
      scale = std::sqrt(e0 * e0 + (e1 * e1 + e2 * e2));   // getBounds().getExtent().norm(), Visibility.hpp:73,88
    }
    Real getScale() const { return scale; }

    double compute(VectorT const & position) const
    {
      float p[3]; detail::put3(p, position); double out = 0;
      computeBatch(p, 1, &out);
      return out;
    }
    void computeBatch(float const * p3, intx S, double * visibility) const
    { bvh->visibilityBatch(p3, S, dirs.data(), (int32_t)(dirs.size() / 3), false, scale, visibility); }

  private:
    GpuMeshBvhT<VectorT> const * bvh; std::vector<float> dirs; Real scale;
};

/** Thea::Random (Random.cpp:48-158): MT19937 whose regeneration never touches state[227] (the loop bounds of Random::generate, :92-118,
 *  as they are), bits() with its tempering (:68-90), integer(lo, hi) (:147-157).  The reference's features draw from the shared
 *  Random::common(), seeded from the clock; a caller that wants reproducible features passes one of these. */
class TheaRandom
{
  public:
    explicit TheaRandom(uint32_t seed)
    {
      st[0] = seed;
      for (index = 1; index < N; ++index) st[index] = 1812433253UL * (st[index - 1] ^ (st[index - 1] >> 30)) + (uint32_t)index;
    }
    uint32_t bits()
    {
      if ((uint32_t)index >= (uint32_t)N)
      {
        uint32_t const LOWER = (1u << 31) - 1, UPPER = 0xFFFFFFFFu << 31; uint32_t const mag[2] = {0u, 0x9908B0DFu};
        for (int i = 0; i < N - M; ++i) { uint32_t x = (st[i] & UPPER) | (st[i + 1] & LOWER); st[i] = st[i + M] ^ (x >> 1) ^ mag[x & 1]; }
        for (int i = N - M + 1; i < N - 1; ++i) { uint32_t x = (st[i] & UPPER) | (st[i + 1] & LOWER); st[i] = st[i + (M - N)] ^ (x >> 1) ^ mag[x & 1]; }
        uint32_t y = (st[N - 1] & UPPER) | (st[0] & LOWER); st[N - 1] = st[M - 1] ^ (y >> 1) ^ mag[y & 1];
        index = 0;
      }
      uint32_t r = st[index++];
      r ^= r >> 11; r ^= (r << 7) & 0x9D2C5680u; r ^= (r << 15) & 0xEFC60000u; r ^= r >> 18;
      return r;
    }
    int32_t integer(int32_t low, int32_t high)
    {
      int32_t r = (int32_t)std::floor(low + (high - low + 1) * (double)bits() / 0xFFFFFFFFUL);
      return r > high ? high : r;
    }

  private:
    enum { N = 624, M = 397 };
    uint32_t st[N]; int index;
};

/** RandomWalks (RandomWalks.cpp:30-117): the average offset from the query position after each step of an n-step random walk on the
 *  sample graph.  The surface is the structure over the samples plus their host positions and the neighbour rows sampleGraph returned
 *  (graph_idx: P x max_degree, -1 padded); the generator is the caller's (see TheaRandom).  computeBatch looks up the seed samples of
 *  all positions in ONE closestElement batch; the walks are the feature's own sequential arithmetic over one random stream -- position
 *  major, walk major, step major, exactly the reference's draw order -- and stay on the host.
 *  SurfaceT: anything with closestElementBatch(q3, Q, idx, d2, cp3) -- GpuPointBvhT by default (a host stand-in lets the CPU suite run the
 *  walk arithmetic, tests/cpp/random_walks_host_check.cpp). */
template <typename VectorT = Vec3, typename SurfaceT = GpuPointBvhT<VectorT> >
class RandomWalksT
{
  public:
    RandomWalksT(SurfaceT const * surf_, float const * sample_positions3, int32_t const * graph_idx_, intx max_degree_, TheaRandom * rng_)
    : surf(surf_), pos(sample_positions3), gi(graph_idx_), md(max_degree_), rng(rng_)
    { if (!surf || !pos || !gi || !rng) throw Error("RandomWalks: Cannot construct with a null surface"); }

    void compute(VectorT const & position, intx num_steps, double * features, intx num_walks = -1) const
    { float p[3]; detail::put3(p, position); computeBatch(p, 1, num_steps, features, num_walks); }

    /** features: Q x num_steps x 3 doubles. */
    void computeBatch(float const * p3, intx Q, intx num_steps, double * features, intx num_walks = -1) const
    {
      if (num_walks < 0) num_walks = 1000;                                           // DEFAULT_NUM_WALKS, RandomWalks.cpp:73
      std::vector<int32_t> seed((size_t)Q);
      surf->closestElementBatch(p3, Q, seed.data(), nullptr, nullptr);
      std::vector<intx> counts((size_t)num_steps);
      for (intx q = 0; q < Q; ++q)
      {
        double * f = features + 3 * num_steps * q;
        for (intx i = 0; i < 3 * num_steps; ++i) f[i] = 0.0;
        if (seed[(size_t)q] < 0) throw Error("RandomWalks: Seed sample for random walks not found");
        std::fill(counts.begin(), counts.end(), (intx)0);
        for (intx w = 0; w < num_walks; ++w)
        {
          intx cur = seed[(size_t)q], steps = num_steps;
          for (intx i = 0; i < num_steps; ++i)
          {
            int32_t deg = 0; while (deg < md && gi[cur * md + deg] >= 0) ++deg;
            if (deg == 0) { steps = i; break; }
            cur = gi[cur * md + rng->integer(0, deg - 1)];
            f[3 * i] += pos[3 * cur]; f[3 * i + 1] += pos[3 * cur + 1]; f[3 * i + 2] += pos[3 * cur + 2];
          }
          for (intx j = 0; j < steps; ++j) counts[(size_t)j]++;
        }
        for (intx i = 0; i < num_steps; ++i)
        {
          intx c = counts[(size_t)i];
          f[3 * i] = f[3 * i] / c - p3[3 * q]; f[3 * i + 1] = f[3 * i + 1] / c - p3[3 * q + 1]; f[3 * i + 2] = f[3 * i + 2] / c - p3[3 * q + 2];
        }
      }
    }

  private:
    SurfaceT const * surf; float const * pos; int32_t const * gi; intx md; TheaRandom * rng;
};

typedef AverageDistanceT<Vec3> AverageDistance;
typedef CurvatureT<Vec3> Curvature;
typedef LocalPcaT<Vec3> LocalPca;
typedef VisibilityT<Vec3> Visibility;
typedef LocalDistanceHistogramT<Vec3> LocalDistanceHistogram;
typedef SpinImageT<Vec3> SpinImage;
typedef RandomWalksT<Vec3> RandomWalks;

}  // namespace Local
}  // namespace SurfaceFeatures
}  // namespace TheaGpu

#endif
