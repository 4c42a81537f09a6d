// oracle_capi.cpp -- C ABI over thea_oracle.hpp for ctypes (tests/, bench.py cpu_baseline / --impl reference).
//
// TEST INFRASTRUCTURE ONLY (see the header of thea_oracle.hpp: parity is pinned to the reference's own compiled code, oracle/_ref).
#include "thea_oracle.hpp"

#include <chrono>
#include <functional>
#include <thread>
#include <queue>

using namespace oracle;

namespace {

struct Handle { KdTree tree; };

void parallel_for(int64_t n, int nthreads, std::function<void(int64_t, int64_t)> const & fn)
{
  if (nthreads <= 1 || n < 2 * nthreads) { fn(0, n); return; }
  std::vector<std::thread> th;
  int64_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t)
  {
    int64_t b = t * chunk, e = std::min(n, b + chunk);
    if (b >= e) break;
    th.emplace_back([=, &fn] { fn(b, e); });
  }
  for (auto & t : th) t.join();
}

inline V3 ld3(float const * p, int64_t i) { return mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline void st3(float * p, int64_t i, V3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

// Deterministic generator of SURVEY.md 8(d): counter-based splitmix64, u01 = (x >> 40) * 2^-24.
inline uint64_t splitmix64_at(uint64_t seed, uint64_t i)
{
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline float u01_at(uint64_t seed, uint64_t i) { return (float)(splitmix64_at(seed, i) >> 40) * (1.0f / 16777216.0f); }

// 3x3 polar rotation R = U V^T of M via one-sided Jacobi in double (stand-in for Eigen::JacobiSVD, Icp3.hpp:369-370).
void polar_rotation(double const M[9], double R[9])
{
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) A[i] = M[i];   // row-major; columns are rotated
  for (int sweep = 0; sweep < 60; ++sweep)
  {
    double off = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q)
      {
        double a = 0, b = 0, c = 0;
        for (int r = 0; r < 3; ++r) { a += A[3 * r + p] * A[3 * r + p]; b += A[3 * r + q] * A[3 * r + q]; c += A[3 * r + p] * A[3 * r + q]; }
        off = std::max(off, std::fabs(c) / std::sqrt(std::max(a * b, 1e-300)));
        if (std::fabs(c) <= 1e-300) continue;
        double zeta = (b - a) / (2.0 * c);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
        for (int r = 0; r < 3; ++r)
        {
          double ap = A[3 * r + p], aq = A[3 * r + q];
          A[3 * r + p] = cs * ap - sn * aq; A[3 * r + q] = sn * ap + cs * aq;
          double vp = V[3 * r + p], vq = V[3 * r + q];
          V[3 * r + p] = cs * vp - sn * vq; V[3 * r + q] = sn * vp + cs * vq;
        }
      }
    if (off < 1e-15) break;
  }
  // A = U * diag(s); normalise columns (rank-deficient columns completed by cross product).
  double U[9]; double s[3];
  for (int c = 0; c < 3; ++c)
  {
    s[c] = std::sqrt(A[c] * A[c] + A[3 + c] * A[3 + c] + A[6 + c] * A[6 + c]);
    for (int r = 0; r < 3; ++r) U[3 * r + c] = s[c] > 0 ? A[3 * r + c] / s[c] : 0.0;
  }
  int mn = 0; for (int c = 1; c < 3; ++c) if (s[c] < s[mn]) mn = c;
  if (s[mn] <= 1e-14 * std::max(s[0], std::max(s[1], s[2])))
  {
    int a = (mn + 1) % 3, b = (mn + 2) % 3;
    U[mn] = U[3 + a] * U[6 + b] - U[6 + a] * U[3 + b];
    U[3 + mn] = U[6 + a] * U[b] - U[a] * U[6 + b];
    U[6 + mn] = U[a] * U[3 + b] - U[3 + a] * U[b];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
    { double acc = 0; for (int k = 0; k < 3; ++k) acc += U[3 * r + k] * V[3 * c + k]; R[3 * r + c] = acc; }
}

struct Affine { float L[9]; V3 t; };   // row-major linear part

inline V3 apply_linear(float const L[9], V3 v)
{ return mk(sum3(L[0] * v.x, L[1] * v.y, L[2] * v.z), sum3(L[3] * v.x, L[4] * v.y, L[5] * v.z), sum3(L[6] * v.x, L[7] * v.y, L[8] * v.z)); }
inline V3 apply(Affine const & a, V3 v) { return apply_linear(a.L, v) + a.t; }            // AffineTransformN.hpp:160
inline Affine compose(Affine const & a, Affine const & b)                                     // AffineTransformN.hpp:154-157
{
  Affine r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.L[3 * i + j] = sum3(a.L[3 * i] * b.L[j], a.L[3 * i + 1] * b.L[3 + j], a.L[3 * i + 2] * b.L[6 + j]);
  r.t = apply_linear(a.L, b.t) + a.t;
  return r;
}
inline Affine identity_tr() { Affine a; for (int i = 0; i < 9; ++i) a.L[i] = (i % 4 == 0) ? 1.0f : 0.0f; a.t = mk(0, 0, 0); return a; }

// Icp3.hpp:410-464 measureError(tr, n, from, nullptr, to) with ScalarT = float.
float icp_error(Affine const & tr, int64_t n, V3 const * from, V3 const * to)
{
  V3 pm = mk(0, 0, 0), qm = mk(0, 0, 0); float sw = 0;
  for (int64_t i = 0; i < n; ++i) { pm = pm + from[i]; qm = qm + to[i]; sw += 1; }
  pm = pm / sw; qm = qm / sw;
  float err = sqnorm(apply(tr, pm) - qm);
  for (int64_t i = 0; i < n; ++i)
  { V3 dp = from[i] - pm, dq = to[i] - qm; err += sqnorm(apply_linear(tr.L, dp) - dq); }
  return err;
}

// Icp3.hpp:298-384 alignOneStep with known correspondences (no weights, no symmetry, no up vector).
Affine icp_step(int64_t n, V3 const * from, V3 const * to)
{
  V3 pm = mk(0, 0, 0), qm = mk(0, 0, 0); float sw = 0;
  for (int64_t i = 0; i < n; ++i) { pm = pm + from[i]; qm = qm + to[i]; sw += 1; }
  pm = pm / sw; qm = qm / sw;
  float cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = 0; i < n; ++i)
  {
    V3 dp = from[i] - pm, dq = to[i] - qm;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[3 * r + c] += dp[r] * dq[c];
  }
  // rot = U V^T of cov = sum dp dq^T (Icp3.hpp:369-370); note that this maps q-space to p-space exactly as the
  // reference writes it.
  double M[9], R[9]; for (int i = 0; i < 9; ++i) M[i] = cov[i];
  polar_rotation(M, R);
  Affine a; for (int i = 0; i < 9; ++i) a.L[i] = (float)R[i];
  a.t = qm - apply_linear(a.L, pm);
  return a;
}

}  // namespace

// Dijkstra on a sample graph with the reference's arithmetic (see the graph section below).
namespace {
struct GraphDijkstra
{
  int64_t const * idx; float const * sep; int64_t P; int K;
  std::vector<double> dist; std::vector<char> black; std::vector<int64_t> touched;
  GraphDijkstra(int64_t const * i, float const * s, int64_t P_, int K_) : idx(i), sep(s), P(P_), K(K_), dist((size_t)P_, INFINITY), black((size_t)P_, 0) {}
  template <typename Visit> void run(std::vector<std::pair<int64_t, double>> const & sources, double limit, Visit visit)
  {
    typedef std::pair<double, int64_t> E;
    std::priority_queue<E, std::vector<E>, std::greater<E>> pq;
    for (auto const & sd : sources) { dist[(size_t)sd.first] = sd.second; touched.push_back(sd.first); pq.push(E(sd.second, sd.first)); }
    while (!pq.empty())
    {
      E top = pq.top(); pq.pop();
      int64_t u = top.second;
      if (black[(size_t)u] || top.first != dist[(size_t)u]) continue;
      for (int e = 0; e < K; ++e)
      {
        int64_t v = idx[u * K + e]; if (v < 0 || black[(size_t)v]) continue;
        double t = (double)sep[u * K + e] + dist[(size_t)u];
        if (t < dist[(size_t)v]) { if (dist[(size_t)v] == INFINITY) touched.push_back(v); dist[(size_t)v] = t; pq.push(E(t, v)); }
      }
      black[(size_t)u] = 1;
      if (limit >= 0 && dist[(size_t)u] > limit) break;
      if (visit(u, dist[(size_t)u])) break;
    }
  }
  void reset() { for (int64_t v : touched) { dist[(size_t)v] = INFINITY; black[(size_t)v] = 0; } touched.clear(); }
};
}  // namespace

extern "C" {

void * orc_create_tris(float const * v9, int64_t T)
{
  Handle * h = new Handle;
  h->tree.is_points = false;
  h->tree.tris.resize((size_t)T);
  for (int64_t i = 0; i < T; ++i) h->tree.tris[(size_t)i] = make_tri(ld3(v9, 3 * i), ld3(v9, 3 * i + 1), ld3(v9, 3 * i + 2));
  h->tree.build();
  return h;
}

void * orc_create_points(float const * p3, int64_t P, float const * n3)
{
  Handle * h = new Handle;
  h->tree.is_points = true;
  h->tree.pts.resize((size_t)P);
  for (int64_t i = 0; i < P; ++i) h->tree.pts[(size_t)i] = ld3(p3, i);
  if (n3) { h->tree.nrm.resize((size_t)P); for (int64_t i = 0; i < P; ++i) h->tree.nrm[(size_t)i] = ld3(n3, i); }
  h->tree.build();
  return h;
}

void orc_destroy(void * h) { delete (Handle *)h; }
int64_t orc_num_nodes(void * h) { return (int64_t)((Handle *)h)->tree.nodes.size(); }
int64_t orc_max_depth(void * h) { return ((Handle *)h)->tree.max_depth; }

void orc_tri_cache(void * h, float * n3, float * d, int32_t * axis, float * area)
{
  KdTree & T = ((Handle *)h)->tree;
  for (size_t i = 0; i < T.tris.size(); ++i)
  { st3(n3, (int64_t)i, T.tris[i].n); d[i] = T.tris[i].d; axis[i] = T.tris[i].axis; if (area) area[i] = T.tris[i].area; }
}

// mode 0: reference order (kd-tree DFS, first found wins); 1: canonical via kd-tree; 2: canonical brute force.
// exclude_self: query i ignores element i (PointSet3::FilterSelf).
void orc_closest(void * h, float const * q3, int64_t Q, double dist_bound, int mode, int exclude_self, int nthreads,
                 int64_t * idx, float * d2, float * cp3)
{
  KdTree const & T = ((Handle *)h)->tree;
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      V3 q = ld3(q3, i); int64_t skip = exclude_self ? i : -1;
      Closest r = mode == 0 ? T.closest<false>(q, dist_bound, skip)
                : mode == 1 ? T.closest<true>(q, dist_bound, skip) : brute_closest(T, q, dist_bound, skip);
      idx[i] = r.idx; if (d2) d2[i] = r.d2; if (cp3) st3(cp3, i, r.cp);
    }
  });
}

// Closest element to a line-segment query object (KdTree::closest_segment): excl = Q x n_excl element indices (-1 unused) or null.
void orc_closest_segment(void * h, float const * a3, float const * b3, int64_t Q, double dist_bound, int64_t const * excl, int n_excl, int mode,
                         int nthreads, int64_t * idx, float * d2, float * ep3, float * sp3)
{
  KdTree const & T = ((Handle *)h)->tree;
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      KdTree::SegClosest r = T.closest_segment(mode, ld3(a3, i), ld3(b3, i), dist_bound, excl ? excl + i * n_excl : nullptr, excl ? n_excl : 0);
      idx[i] = r.idx; if (d2) d2[i] = r.d2; if (ep3) st3(ep3, i, r.on_elem); if (sp3) st3(sp3, i, r.on_seg);
    }
  });
}

// Closest element to a triangle query object (KdTree::closest_triangle); plain_bound = the structure-to-structure bound semantics.
void orc_closest_triangle(void * h, float const * q9, int64_t Q, double dist_bound, int plain_bound, int mode, int nthreads, int64_t * idx, float * d2,
                          float * ep3, float * qp3)
{
  KdTree const & T = ((Handle *)h)->tree;
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      KdTree::SegClosest r = T.closest_triangle(mode, ld3(q9, 3 * i), ld3(q9, 3 * i + 1), ld3(q9, 3 * i + 2), dist_bound, plain_bound != 0);
      idx[i] = r.idx; if (d2) d2[i] = r.d2; if (ep3) st3(ep3, i, r.on_elem); if (qp3) st3(qp3, i, r.on_seg);
    }
  });
}

// Labels of a point structure (element order) and the per-query filter of filtered queries: query i admits element e iff
// (qlabel == null or label[e] == qlabel[i]) and (qn3 == null or qn[i] . n[e] >= min_dot)   (Register.cpp:92-101 NNFilter).
void orc_set_labels(void * h, int32_t const * labels)
{
  KdTree & T = ((Handle *)h)->tree;
  T.labels.clear();
  if (labels) T.labels.assign(labels, labels + T.size());
}

void orc_closest_filtered(void * h, float const * q3, float const * qn3, int32_t const * qlabel, int64_t Q, float min_dot, double dist_bound,
                          int mode, int nthreads, int64_t * idx, float * d2, float * cp3)
{
  KdTree const & T = ((Handle *)h)->tree;
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      ElemFilter f; f.use_normal = qn3 != nullptr; f.n = qn3 ? ld3(qn3, i) : mk(0, 0, 0); f.min_dot = min_dot;
      f.use_label = qlabel != nullptr; f.label = qlabel ? qlabel[i] : 0;
      g_filter = &f;                                              // pushFilter
      V3 q = ld3(q3, i);
      Closest r = mode == 0 ? T.closest<false>(q, dist_bound, -1) : mode == 1 ? T.closest<true>(q, dist_bound, -1) : brute_closest(T, q, dist_bound, -1);
      g_filter = nullptr;                                         // popFilter
      idx[i] = r.idx; if (d2) d2[i] = r.d2; if (cp3) st3(cp3, i, r.cp);
    }
  });
}

void orc_knn_filtered(void * h, float const * q3, float const * qn3, int32_t const * qlabel, int64_t Q, int k, float min_dot, double dist_bound,
                      int mode, int nthreads, int64_t * idx, float * d2)
{
  KdTree const & T = ((Handle *)h)->tree;
  double bound2 = dist_bound >= 0 ? dist_bound * dist_bound : -1;
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      ElemFilter f; f.use_normal = qn3 != nullptr; f.n = qn3 ? ld3(qn3, i) : mk(0, 0, 0); f.min_dot = min_dot;
      f.use_label = qlabel != nullptr; f.label = qlabel ? qlabel[i] : 0;
      g_filter = &f;
      V3 q = ld3(q3, i);
      int64_t * oi = idx + (int64_t)k * i; float * od = d2 + (int64_t)k * i;
      if (mode == 2) brute_knn(T, q, k, dist_bound, -1, oi, od);
      else if (mode == 0)
      {
        KdTree::KnnSet s; s.k = k;
        if (!T.nodes.empty()) T.knn_rec(0, q, bound2, -1, s);
        for (int j = 0; j < k; ++j) { oi[j] = j < (int)s.d.size() ? s.id[(size_t)j] : -1; od[j] = j < (int)s.d.size() ? (float)s.d[(size_t)j] : -1.0f; }
      }
      else
      {
        std::vector<std::pair<float, int64_t>> s;
        if (!T.nodes.empty()) T.knn_canon_rec(0, q, bound2, -1, k, s);
        for (int j = 0; j < k; ++j) { oi[j] = j < (int)s.size() ? s[(size_t)j].second : -1; od[j] = j < (int)s.size() ? s[(size_t)j].first : -1.0f; }
      }
      g_filter = nullptr;
    }
  });
}

// BvhN::rangeQueryIndices with Ball3(c, r) on a point structure (BvhN.hpp:1208-1220, BallN.hpp:83): brute force over all
// elements; count and the max_out lowest indices ascending.
void orc_ball_range(void * h, float const * c3, float const * radius, int64_t Q, int max_out, int exclude_self, int nthreads, int64_t * idx,
                    int32_t * count)
{
  KdTree const & T = ((Handle *)h)->tree;
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      V3 c = ld3(c3, i); float r = radius[i]; float r2 = r * r; int n = 0;
      for (int j = 0; j < max_out; ++j) idx[i * max_out + j] = -1;
      if (r >= 0)
        for (int64_t el = 0; el < T.size(); ++el)
        {
          if (exclude_self && el == i) continue;
          if (!(sqnorm(T.pts[(size_t)el] - c) <= r2)) continue;
          if (n < max_out) idx[i * max_out + n] = el;
          ++n;
        }
      count[i] = n;
    }
  });
}

// Local surface features that are processRange calls with an accumulating functor, restated with the
// reference's arithmetic in the reference's traversal order (mode 0: kd-tree processRange) or in element order (mode 2).
//   AverageDistance (EUCLIDEAN, explicit max_distance): AverageDistance.cpp:39-81 -- Real d = (x - c).norm(); double sum.
//   Curvature::computeProjectedCurvature(p, n, r):     Curvature.cpp:37-98     -- fp32 sum of stableNormalized offsets over
//       samples with n_e . n > -1e-5; -sum.dot(n) / num (fp32).
//   LocalPca::compute(p, evecs, r):                     LocalPca.cpp:32-62, PcaN.hpp:150-197, CentroidN.hpp:80-97 -- fp32
//       centroid (sum / (float)n), fp32 covariance of x - ctr, / (n - 1).  The eigen decomposition itself is Eigen's
//       SelfAdjointEigenSolver (third party, not restated); tests compare covariances and solve both sides the same way.
// radius[i] is the ABSOLUTE ball radius (the callers multiply nbd_radius by surf->getScale()).
// out[Q x 16]: [0] n, [1] sum_distances / n (0 when n == 0), [2] num_offsets, [3] curvature, [4..6] centroid,
//              [7..12] covariance xx xy xz yy yz zz (zeros when n < 2), [13..15] 0.
void orc_ball_stats(void * h, float const * c3, float const * qn3, float const * radius, int64_t Q, int mode, int nthreads, double * out)
{
  KdTree const & T = ((Handle *)h)->tree;
  int64_t P = T.size();
  bool has_n = !T.nrm.empty() && qn3;
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    std::vector<int64_t> nb;
    for (int64_t i = b; i < e; ++i)
    {
      V3 c = ld3(c3, i); float r = radius[i]; float r2 = r * r;
      double * o = out + 16 * i; for (int j = 0; j < 16; ++j) o[j] = 0;
      nb.clear();
      auto emit = [&](int64_t j) { nb.push_back(j); };
      if (r >= 0)
      {
        if (mode == 2) { for (int64_t j = 0; j < P; ++j) if (sqnorm(T.pts[(size_t)j] - c) <= r2) emit(j); }
        else if (!T.nodes.empty()) kd_ball_range(T, 0, c, r2, emit);
      }
      // AverageDistance
      double sum_d = 0; int64_t n = 0;
      for (int64_t j : nb) { float d = std::sqrt(sqnorm(T.pts[(size_t)j] - c)); sum_d += d; ++n; }
      o[0] = (double)n; o[1] = n <= 0 ? 0.0 : sum_d / (double)n;
      // Curvature
      if (has_n)
      {
        V3 nq = ld3(qn3, i), so = mk(0, 0, 0); int64_t no = 0;
        for (int64_t j : nb)
          if (dot(T.nrm[(size_t)j], nq) > -1.0e-05f)
          {
            V3 a = T.pts[(size_t)j] - c;
            float w = std::max(std::max(std::fabs(a.x), std::fabs(a.y)), std::fabs(a.z));
            float z = sqnorm(a / w);
            V3 off = z > 0 ? a / (std::sqrt(z) * w) : a;
            so = so + off; ++no;
          }
        o[2] = (double)no;
        o[3] = no > 0 ? (double)(-dot(so, nq) / (float)no) : 0.0;
      }
      // LocalPca
      if (n > 0)
      {
        V3 sp = mk(0, 0, 0);
        for (int64_t j : nb) sp = sp + T.pts[(size_t)j];
        V3 ctr = sp / (float)(double)n;
        o[4] = ctr.x; o[5] = ctr.y; o[6] = ctr.z;
        if (n >= 2)
        {
          float cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
          for (int64_t j : nb)
          {
            V3 p = T.pts[(size_t)j] * 1.0f - ctr;
            for (int rr = 0; rr < 3; ++rr) for (int cc = 0; cc < 3; ++cc) cov[3 * rr + cc] += p[rr] * p[cc];
          }
          float den = (float)(n - 1);
          for (int k = 0; k < 9; ++k) cov[k] /= den;
          o[7] = cov[0]; o[8] = cov[1]; o[9] = cov[2]; o[10] = cov[4]; o[11] = cov[5]; o[12] = cov[8];
        }
      }
    }
  });
}

// SpinImage::compute(position, normal, num_radial_bins, num_height_bins, spin_image) (SurfaceFeatures/Local/SpinImage.cpp:47-75):
// axis = normal.stableNormalized(); every sample of the set is binned by (radius about the axis, height along it) with
// max_radius = max_height = max_extent (the reference passes 0.75f * surf->getScale()).  image: [Q x nr x nh], radial bin major.
void orc_spin_image(void * h, float const * q3, float const * qn3, int64_t Q, int nr, int nh, float max_extent, int nthreads, double * image)
{
  KdTree const & T = ((Handle *)h)->tree;
  int64_t P = T.size();
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      V3 position = ld3(q3, i); V3 axis = stable_normalized(ld3(qn3, i));
      double * img = image + (int64_t)nr * nh * i;
      for (int j = 0; j < nr * nh; ++j) img[j] = 0;
      float max_radius = max_extent, max_height = max_extent;
      for (int64_t j = 0; j < P; ++j)
      {
        V3 offset = T.pts[(size_t)j] - position;
        float height = dot(offset, axis);
        float radius = std::sqrt(sqnorm(offset - axis * height));
        int radial_bin = (int)std::floor(nr * (radius / max_radius));
        radial_bin = radial_bin < 0 ? 0 : (radial_bin > nr - 1 ? nr - 1 : radial_bin);
        int height_bin = (int)std::floor(nh * (0.5f * (height / max_height + 1)));
        height_bin = height_bin < 0 ? 0 : (height_bin > nh - 1 ? nh - 1 : height_bin);
        img[radial_bin * nh + height_bin] += 1;
      }
    }
  });
}

// LocalDistanceHistogram::compute(c, histogram, EUCLIDEAN, max_distance, ratio >= 1) (LocalDistanceHistogram.cpp:36-88) with
// Histogram::setRange / insert (Histogram.hpp:110-150), in the reference's traversal order (mode 0) or element order (mode 2).
void orc_ball_hist(void * h, float const * c3, float const * radius, int64_t Q, int nbins, int mode, int nthreads, double * hist)
{
  KdTree const & T = ((Handle *)h)->tree;
  int64_t P = T.size();
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      V3 c = ld3(c3, i); float r = radius[i]; float r2 = r * r;
      double * bins = hist + (int64_t)nbins * i;
      double min_value = 0, max_value = std::max((double)r, 1.0e-30);
      double bin_size = (max_value - min_value) / nbins;
      for (int j = 0; j < nbins; ++j) bins[j] = 0;
      auto emit = [&](int64_t j) {
        float d = std::sqrt(sqnorm(T.pts[(size_t)j] - c));
        int64_t bin = (int64_t)std::floor(((double)d - min_value) / bin_size);
        bin = bin < 0 ? 0 : (bin > nbins - 1 ? nbins - 1 : bin);
        bins[bin] += 1.0;
      };
      if (r >= 0)
      {
        if (mode == 2) { for (int64_t j = 0; j < P; ++j) if (sqnorm(T.pts[(size_t)j] - c) <= r2) emit(j); }
        else if (!T.nodes.empty()) kd_ball_range(T, 0, c, r2, emit);
      }
    }
  });
}

// Visibility::compute (SurfaceFeatures/Local/Visibility.hpp:104-125) with caller-supplied directions: ray (p + 0.001f * scale * u, u),
// rayIntersects (= full first hit >= 0, BvhN.hpp:1266-1269); vis = (double)num_escaped / num_rays.
void orc_visibility(void * h, float const * p3, int64_t S, float const * dirs3, int D, int per_sample, float scale, int mode, int nthreads,
                    double * vis, int32_t * nesc)
{
  KdTree const & T = ((Handle *)h)->tree;
  parallel_for(S, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      V3 p = ld3(p3, i); int64_t esc = 0;
      for (int k = 0; k < D; ++k)
      {
        V3 u = ld3(dirs3, per_sample ? i * D + k : k);
        V3 offset = u * (0.001f * scale);
        V3 o = p + offset;
        RayHit r = mode == 0 ? T.ray_first_ref(o, u, -1.0f) : mode == 1 ? T.ray_first_canon(o, u, -1.0f) : brute_ray_first(T, o, u, -1.0f);
        if (r.idx < 0) ++esc;
      }
      if (vis) vis[i] = (double)esc / (double)D;
      if (nesc) nesc[i] = (int32_t)esc;
    }
  });
}

void orc_ray_first(void * h, float const * o3, float const * d3, float const * tmax, int64_t R, int mode, int nthreads,
                   int64_t * idx, float * t)
{
  KdTree const & T = ((Handle *)h)->tree;
  parallel_for(R, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      V3 o = ld3(o3, i), d = ld3(d3, i); float mt = tmax ? tmax[i] : -1.0f;
      RayHit r = mode == 0 ? T.ray_first_ref(o, d, mt) : mode == 1 ? T.ray_first_canon(o, d, mt) : brute_ray_first(T, o, d, mt);
      idx[i] = r.idx; t[i] = r.t;
    }
  });
}

void orc_knn(void * h, float const * q3, int64_t Q, int k, double dist_bound, int exclude_self, int mode, int nthreads,
             int64_t * idx, float * d2)
{
  KdTree const & T = ((Handle *)h)->tree;
  double bound2 = dist_bound >= 0 ? dist_bound * dist_bound : -1;
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      V3 q = ld3(q3, i); int64_t skip = exclude_self ? i : -1;
      int64_t * oi = idx + (int64_t)k * i; float * od = d2 + (int64_t)k * i;
      if (mode == 2) { brute_knn(T, q, k, dist_bound, skip, oi, od); continue; }
      if (mode == 0)
      {
        KdTree::KnnSet s; s.k = k;
        if (!T.nodes.empty()) T.knn_rec(0, q, bound2, skip, s);
        for (int j = 0; j < k; ++j) { oi[j] = j < (int)s.d.size() ? s.id[(size_t)j] : -1; od[j] = j < (int)s.d.size() ? (float)s.d[(size_t)j] : -1.0f; }
      }
      else
      {
        std::vector<std::pair<float, int64_t>> s;
        if (!T.nodes.empty()) T.knn_canon_rec(0, q, bound2, skip, k, s);
        for (int j = 0; j < k; ++j) { oi[j] = j < (int)s.size() ? s[(size_t)j].second : -1; od[j] = j < (int)s.size() ? s[(size_t)j].first : -1.0f; }
      }
    }
  });
}

// PointSet3.hpp:363-376: RMS nearest-neighbour separation over fixed seed samples (float accumulation).
float orc_avg_separation(void * h, int64_t const * seeds, int n)
{
  KdTree const & T = ((Handle *)h)->tree;
  float acc = 0;
  for (int i = 0; i < n; ++i)
  {
    int64_t s = seeds[i];
    Closest r = T.closest<true>(T.pts[(size_t)s], -1, s);
    acc += sqnorm(T.pts[(size_t)r.idx] - T.pts[(size_t)s]);
  }
  return std::sqrt(acc / (float)n);
}

// mode 0: kd-tree processRange (reference traversal); 2: brute force over all points.
void orc_ball_graph(void * h, int max_degree, float init_radius, int mode, int nthreads, int64_t * idx, float * sep,
                    int32_t * count)
{
  KdTree const & T = ((Handle *)h)->tree;
  int64_t P = T.size();
  parallel_for(P, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      auto range = [&](V3 c, float r2, std::function<void(int64_t)> emit) {
        if (mode == 2) { for (int64_t j = 0; j < P; ++j) if (sqnorm(T.pts[(size_t)j] - c) <= r2) emit(j); }
        else kd_ball_range(T, 0, c, r2, emit);
      };
      int c = ball_graph_one(T, i, max_degree, init_radius, range, idx + (int64_t)max_degree * i, sep + (int64_t)max_degree * i);
      if (count) count[i] = c;
    }
  });
}

// PointSet3::updateGraph(surface) (PointSet3.hpp:324-391,639-648): the neighbour search with the reachability test against a
// triangle structure.  Ray3(x_i + 5 * sep * normalized(n_i + n_j), x_j - x_i), rayIntersects(ray, 1).  mode as orc_ball_graph;
// the ray is answered by the reference-order traversal (mode 0) or by brute force (mode 2) of the surface oracle.
void orc_ball_graph_surface(void * h, void * hs, int max_degree, float init_radius, int mode, int nthreads, int64_t * idx, float * sep,
                            int32_t * count)
{
  KdTree const & T = ((Handle *)h)->tree;
  KdTree const & S = ((Handle *)hs)->tree;
  int64_t P = T.size();
  bool has_n = !T.nrm.empty();
  parallel_for(P, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
    {
      auto range = [&](V3 c, float r2, std::function<void(int64_t)> emit) {
        if (mode == 2) { for (int64_t j = 0; j < P; ++j) if (sqnorm(T.pts[(size_t)j] - c) <= r2) emit(j); }
        else kd_ball_range(T, 0, c, r2, emit);
      };
      auto blocked = [&](int64_t j, float s, V3 diff) {
        if (!has_n) return false;
        V3 lift_dir = normalized(T.nrm[(size_t)i] + T.nrm[(size_t)j]);
        V3 o = T.pts[(size_t)i] + (5.0f * s) * lift_dir;
        RayHit r = mode == 2 ? brute_ray_first(S, o, diff, 1.0f) : S.ray_first_ref(o, diff, 1.0f);
        return r.idx >= 0;
      };
      int c = ball_graph_one(T, i, max_degree, init_radius, range, idx + (int64_t)max_degree * i, sep + (int64_t)max_degree * i, blocked);
      if (count) count[i] = c;
    }
  });
}

// ShapeDiameter::compute(p, n, only_hit_interior) for S samples: emits the 30 rays, their first hits and the SDF.
void orc_sdf(void * h, float const * p3, float const * n3, int64_t S, float scale, int only_interior, int mode, int nthreads,
             float * ray_o3, float * ray_d3, int64_t * hit_idx, float * hit_t, double * sdf)
{
  KdTree const & T = ((Handle *)h)->tree;
  parallel_for(S, nthreads, [&](int64_t b, int64_t e) {
    for (int64_t s = b; s < e; ++s)
    {
      V3 org, dirs[30]; sdf_rays(ld3(p3, s), ld3(n3, s), scale, org, dirs);
      float t[30]; bool valid[30];
      for (int i = 0; i < 30; ++i)
      {
        RayHit r = mode == 0 ? T.ray_first_ref(org, dirs[i], -1) : mode == 1 ? T.ray_first_canon(org, dirs[i], -1)
                                                                              : brute_ray_first(T, org, dirs[i], -1);
        t[i] = r.t;
        valid[i] = r.idx >= 0 && (!only_interior || dot(T.tris[(size_t)r.idx].n, dirs[i]) >= 0);
        if (ray_o3) st3(ray_o3, 30 * s + i, org);
        if (ray_d3) st3(ray_d3, 30 * s + i, dirs[i]);
        if (hit_idx) hit_idx[30 * s + i] = r.idx;
        if (hit_t) hit_t[30 * s + i] = r.t;
      }
      sdf[s] = sdf_stats(t, valid, scale);
    }
  });
}

// Statistics only (for checking a GPU epilogue given identical per-ray hits).
double orc_sdf_stats(float const * t30, uint8_t const * valid30, float scale)
{
  bool v[30]; for (int i = 0; i < 30; ++i) v[i] = valid30[i] != 0;
  return sdf_stats(t30, v, scale);
}

// Icp3<float>::align (Icp3.hpp:193-275) against a triangle structure; records the per-batch correspondences.
// out_idx: (iters_done + 1) x n ; out_tr: 12 floats row-major [L | t]; returns the number of query batches run.
// out_from (optional): the source positions every batch queried, (iters_done + 1) x n x 3 -- lets a test feed the very same iterates
// to another implementation and demand identical correspondences batch by batch.
int orc_icp_ex(void * h, float const * from3, int64_t n, int min_iters, int max_iters, double frac_thresh, int mode, int nthreads,
               int64_t * out_idx, float * out_tr, float * out_err, float * out_from)
{
  KdTree const & T = ((Handle *)h)->tree;
  std::vector<V3> orig((size_t)n), from((size_t)n), to((size_t)n);
  for (int64_t i = 0; i < n; ++i) orig[(size_t)i] = from[(size_t)i] = ld3(from3, i);
  int batches = 0;
  auto correspond = [&]() {
    int64_t * oi = out_idx ? out_idx + (int64_t)batches * n : nullptr;
    if (out_from) for (int64_t i = 0; i < n; ++i) st3(out_from + 3 * (int64_t)batches * n, i, from[(size_t)i]);
    parallel_for(n, nthreads, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i)
      {
        Closest r = mode == 0 ? T.closest<false>(from[(size_t)i], -1) : mode == 1 ? T.closest<true>(from[(size_t)i], -1)
                                                                                  : brute_closest(T, from[(size_t)i], -1);
        to[(size_t)i] = r.cp; if (oi) oi[i] = r.idx;
      }
    });
    batches++;
  };
  auto emit = [&](Affine const & a, float err) {
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out_tr[4 * r + c] = a.L[3 * r + c]; out_tr[4 * r + 3] = a.t[r]; }
    if (out_err) *out_err = err;
  };
  Affine old_tr = identity_tr();
  correspond();
  float old_err = icp_error(old_tr, n, from.data(), to.data());
  if (old_err <= std::numeric_limits<float>::min()) { emit(old_tr, old_err); return batches; }
  float new_err = old_err; Affine new_tr = old_tr;
  for (int it = 0; it < max_iters; ++it)
  {
    Affine inc = icp_step(n, from.data(), to.data());
    new_tr = compose(inc, old_tr);
    // (error is always requested by this entry point, so the last iteration measures too: Icp3.hpp:236-237)
    correspond();
    new_err = icp_error(inc, n, from.data(), to.data());
    if (it < max_iters - 1)
    {
      float frac = (old_err - new_err) / old_err;
      if (it >= min_iters && frac < frac_thresh)
      {
        if (frac > 0) { emit(new_tr, new_err); return batches; }
        else { emit(old_tr, old_err); return batches; }
      }
      for (int64_t j = 0; j < n; ++j) from[(size_t)j] = apply(new_tr, orig[(size_t)j]);
    }
    old_err = new_err; old_tr = new_tr;
  }
  emit(new_tr, new_err);
  return batches;
}

int orc_icp(void * h, float const * from3, int64_t n, int min_iters, int max_iters, double frac_thresh, int mode, int nthreads,
            int64_t * out_idx, float * out_tr, float * out_err)
{ return orc_icp_ex(h, from3, n, min_iters, max_iters, frac_thresh, mode, nthreads, out_idx, out_tr, out_err, nullptr); }

// The two accumulation loops of Icp3 on their own, with given correspondences (pinned bit for bit to the reference's code by
// tests/test_ref_pin.py): measureError (Icp3.hpp:418-464) and alignOneStep (:298-384).  tr12 = row-major [L | t].
float orc_icp_error(float const * tr12, int64_t n, float const * from3, float const * to3)
{
  Affine a; for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) a.L[3 * r + c] = tr12[4 * r + c]; a.t[r] = tr12[4 * r + 3]; }
  std::vector<V3> f((size_t)n), t((size_t)n);
  for (int64_t i = 0; i < n; ++i) { f[(size_t)i] = ld3(from3, i); t[(size_t)i] = ld3(to3, i); }
  return icp_error(a, n, f.data(), t.data());
}
void orc_icp_step(int64_t n, float const * from3, float const * to3, float * tr12)
{
  std::vector<V3> f((size_t)n), t((size_t)n);
  for (int64_t i = 0; i < n; ++i) { f[(size_t)i] = ld3(from3, i); t[(size_t)i] = ld3(to3, i); }
  Affine a = icp_step(n, f.data(), t.data());
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tr12[4 * r + c] = a.L[3 * r + c]; tr12[4 * r + 3] = a.t[r]; }
}

// MeshSampler::sampleEvenlyByArea EXACT mode (MeshSampler.hpp:147-185, Triangle3.hpp:179-195) driven by the
// counter-based generator: sample i consumes draws 3i, 3i+1, 3i+2.
void orc_area_sample(void * h, int64_t S, uint64_t seed, float * p3, float * n3, int64_t * tri)
{
  KdTree const & T = ((Handle *)h)->tree;
  size_t n = T.tris.size();
  if (n == 0) return;
  double total = 0; for (size_t i = 0; i < n; ++i) total += T.tris[i].area;
  std::vector<double> cum(n);
  cum[0] = T.tris[0].area; for (size_t i = 1; i < n; ++i) cum[i] = cum[i - 1] + T.tris[i].area;
  for (int64_t i = 0; i < S; ++i)
  {
    double t = u01_at(seed, 3 * (uint64_t)i) * total;
    int64_t ti = std::min((int64_t)(std::lower_bound(cum.begin(), cum.end(), t) - cum.begin()), (int64_t)n - 1);
    Tri const & tr = T.tris[(size_t)ti];
    float s = u01_at(seed, 3 * (uint64_t)i + 1), u = u01_at(seed, 3 * (uint64_t)i + 2);
    if (s + u > 1) { u = 1 - u; s = 1 - s; }
    st3(p3, i, (s * tr.e01 + u * tr.e02) + tr.v0);
    if (n3) st3(n3, i, tr.n);
    if (tri) tri[i] = ti;
  }
}

// Single-primitive entry points for known-answer tests.
float orc_tri_point(float const * v9, float const * p, float * cp)
{
  Tri t = make_tri(ld3(v9, 0), ld3(v9, 1), ld3(v9, 2)); V3 c; float d2 = tri_point_sqdist(t, ld3(p, 0), c);
  if (cp) st3(cp, 0, c); return d2;
}
float orc_ray_tri(float const * v9, float const * o, float const * d)
{ Tri t = make_tri(ld3(v9, 0), ld3(v9, 1), ld3(v9, 2)); return ray_tri_time(ld3(o, 0), ld3(d, 0), t.v0, t.e01, t.e02); }
void orc_tri_point_batch(float const * v9, float const * p3, int64_t n, float * d2, float * cp3)
{ for (int64_t i = 0; i < n; ++i) d2[i] = orc_tri_point(v9 + 9 * i, p3 + 3 * i, cp3 + 3 * i); }
void orc_ray_tri_batch(float const * v9, float const * o3, float const * d3, int64_t n, float * t)
{ for (int64_t i = 0; i < n; ++i) t[i] = orc_ray_tri(v9 + 9 * i, o3 + 3 * i, d3 + 3 * i); }
// Line-segment query objects: point and triangle elements against LineSegment3(a, b) (known-answer and pin tests)
void orc_seg_point_batch(float const * a3, float const * b3, float const * p3, int64_t n, float * d2, float * sp3)
{
  for (int64_t i = 0; i < n; ++i)
  { V3 sp; d2[i] = seg_point_sqdist(make_seg(ld3(a3, i), ld3(b3, i)), ld3(p3, i), sp); if (sp3) st3(sp3, i, sp); }
}
void orc_tri_seg_batch(float const * v9, float const * a3, float const * b3, int64_t n, float * d2, float * tp3, float * sp3)
{
  for (int64_t i = 0; i < n; ++i)
  {
    Tri t = make_tri(ld3(v9, 3 * i), ld3(v9, 3 * i + 1), ld3(v9, 3 * i + 2)); V3 tp = mk(0, 0, 0), sp = mk(0, 0, 0);
    d2[i] = tri_seg_sqdist(t, make_seg(ld3(a3, i), ld3(b3, i)), tp, sp);
    if (tp3) st3(tp3, i, tp);
    if (sp3) st3(sp3, i, sp);
  }
}
void orc_tri_tri_batch(float const * va9, float const * vb9, int64_t n, float * d2, float * pa3, float * pb3)
{
  for (int64_t i = 0; i < n; ++i)
  {
    Tri a = make_tri(ld3(va9, 3 * i), ld3(va9, 3 * i + 1), ld3(va9, 3 * i + 2)), b = make_tri(ld3(vb9, 3 * i), ld3(vb9, 3 * i + 1), ld3(vb9, 3 * i + 2));
    V3 pa = mk(0, 0, 0), pb = mk(0, 0, 0);
    d2[i] = tri_tri_sqdist(a, b, pa, pb);
    if (pa3) st3(pa3, i, pa);
    if (pb3) st3(pb3, i, pb);
  }
}
float orc_fast_sqrt(float x) { return fast_sqrt(x); }
float orc_box_point_sqdist(float const * lo, float const * hi, float const * p) { Box b; b.merge(ld3(lo, 0), ld3(hi, 0)); return b.sqdist_point(ld3(p, 0)); }
float orc_box_ray_time(float const * lo, float const * hi, float const * o, float const * d, float max_time)
{ Box b; b.merge(ld3(lo, 0), ld3(hi, 0)); return b.ray_time(ld3(o, 0), ld3(d, 0), max_time); }
int orc_tree_depth(int64_t n, int leaf) { return KdTree::tree_depth(n, leaf, 0.5f); }
void orc_polar_rotation(double const * M, double * R) { polar_rotation(M, R); }

// Thea::Random (Random.cpp:48-125): MT19937 whose regeneration skips state[227]; uniform01 (Random.hpp:126-130).
void orc_thea_random_uniform01(uint32_t seed, int64_t n, float * out)
{
  enum { N = 624, M = 397 };
  uint32_t st[N]; st[0] = seed;
  for (int i = 1; i < N; ++i) st[i] = 1812433253UL * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
  int index = N;
  uint32_t const LOWER = (1u << 31) - 1, UPPER = 0xFFFFFFFFu << 31; uint32_t const mag[2] = {0u, 0x9908B0DFu};
  for (int64_t k = 0; k < n; ++k)
  {
    if (index >= N)
    {
      for (int i = 0; i < N - M; ++i) { uint32_t x = (st[i] & UPPER) | (st[i + 1] & LOWER); st[i] = st[i + M] ^ (x >> 1) ^ mag[x & 1]; }
      for (int i = N - M + 1; i < N - 1; ++i) { uint32_t x = (st[i] & UPPER) | (st[i + 1] & LOWER); st[i] = st[i + (M - N)] ^ (x >> 1) ^ mag[x & 1]; }
      uint32_t y = (st[N - 1] & UPPER) | (st[0] & LOWER); st[N - 1] = st[M - 1] ^ (y >> 1) ^ mag[y & 1];
      index = 0;
    }
    uint32_t r = st[index++];
    r ^= r >> 11; r ^= (r << 7) & 0x9D2C5680u; r ^= (r << 15) & 0xEFC60000u; r ^= r >> 18;
    out[k] = (float)r * (1.0f / (float)0xFFFFFFFFu);
  }
}


// Thea::Random as an object (Random.cpp:48-158): bits() and integer(lo, hi).
namespace {
struct TheaRandom
{
  enum { N = 624, M = 397 };
  uint32_t st[N]; int index;
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
  // Random::integer(low, high), Random.cpp:147-157
  int32_t integer(int32_t low, int32_t high)
  {
    int32_t r = (int32_t)std::floor(low + (high - low + 1) * (double)bits() / 0xFFFFFFFFUL);
    return r > high ? high : r;
  }
};
}  // namespace

// RandomWalks::compute (SurfaceFeatures/Local/RandomWalks.cpp:30-117) for Q positions in turn, all drawing from ONE generator in the
// state of Random(seed) (the reference's shared Random::common()): seed sample = the structure's closest element to the position,
// num_walks walks of up to num_steps steps over the neighbour rows idx [P x K] (-1 padded; a sample without neighbours ends the walk),
// per step the running sum of the positions reached (float coordinates added into doubles, in walk order), divided by the number of walks
// that got that far, less the query position.
void orc_random_walks(void * h, int64_t const * idx, int64_t P, int K, float const * q3, int64_t Q, int64_t num_steps, int64_t num_walks, uint32_t seed,
                      double * features)
{
  KdTree const & T = ((Handle *)h)->tree;
  TheaRandom rnd(seed);
  std::vector<int64_t> counts((size_t)num_steps);
  for (int64_t qi = 0; qi < Q; ++qi)
  {
    double * f = features + 3 * num_steps * qi;
    for (int64_t i = 0; i < 3 * num_steps; ++i) f[i] = 0.0;
    V3 pos = ld3(q3, qi);
    int64_t seed_index = T.closest<false>(pos, -1).idx;
    std::fill(counts.begin(), counts.end(), 0);
    for (int64_t w = 0; w < num_walks; ++w)
    {
      int64_t cur = seed_index, steps = num_steps;
      for (int64_t i = 0; i < num_steps; ++i)
      {
        int64_t deg = 0; while (deg < K && idx[cur * K + deg] >= 0) ++deg;
        if (deg == 0) { steps = i; break; }
        int32_t next = rnd.integer(0, (int32_t)deg - 1);
        cur = idx[cur * K + next];
        V3 p = T.pts[(size_t)cur];
        f[3 * i] += p.x; f[3 * i + 1] += p.y; f[3 * i + 2] += p.z;
      }
      for (int64_t j = 0; j < steps; ++j) counts[(size_t)j]++;
    }
    for (int64_t i = 0; i < num_steps; ++i)
    {
      int64_t c = counts[(size_t)i];
      f[3 * i] = f[3 * i] / c - pos.x; f[3 * i + 1] = f[3 * i + 1] / c - pos.y; f[3 * i + 2] = f[3 * i + 2] / c - pos.z;
    }
  }
}

// ---- shortest paths on the sample graph: ShortestPaths::dijkstraWithCallback (Algorithms/ShortestPaths.hpp:166-353) restated with a
// binary heap.  The reference's Fibonacci heap pops equal distances in an order of its own; here equal distances pop in index order
// (the canonical order the engine reproduces).  Distances are the reference's: test_dist = nbr_dist + data->dist in double (:304),
// adopted when strictly smaller (:305).  visit(vertex, distance) returns true to stop (the callback's return value, :346).

// PointSet3::extractOriginalAdjacencies (PointSet3.cpp:96-178).
void orc_graph_contract(int64_t const * idx, float const * sep, int64_t P, int K, int64_t n_orig, int nthreads, int64_t * out_idx, float * out_sep,
                        int32_t * out_count)
{
  parallel_for(n_orig, nthreads, [&](int64_t b, int64_t e) {
    GraphDijkstra D(idx, sep, P, K);
    for (int64_t s = b; s < e; ++s)
    {
      std::vector<std::pair<float, int64_t>> nb;         // the neighbour set: ordered by (Real separation, sample address) (SamplePoint3.hpp:52-55)
      D.run({{s, 0.0}}, -1.0, [&](int64_t v, double d) {
        if (v != s && v < n_orig) nb.push_back(std::make_pair((float)d, v));
        return (int)nb.size() >= K;
      });
      D.reset();
      std::sort(nb.begin(), nb.end());
      if (out_count) out_count[s] = (int32_t)nb.size();
      for (int j = 0; j < K; ++j)
      {
        bool ok = j < (int)nb.size();
        out_idx[s * K + j] = ok ? nb[(size_t)j].second : -1; out_sep[s * K + j] = ok ? nb[(size_t)j].first : -1.0f;
      }
    }
  });
}

// AverageDistance / LocalDistanceHistogram, DistanceType::GEODESIC, from given seed samples (AverageDistance.cpp:84-118,
// LocalDistanceHistogram.cpp:90-137; Histogram.hpp:110-150).  sum in the callback's order: ascending distance.
void orc_graph_geodesic(int64_t const * idx, float const * sep, int64_t P, int K, int64_t const * seeds, int64_t Q, float max_distance, float hist_range,
                        int nbins, int nthreads, double * sum, int64_t * count, double * hist)
{
  parallel_for(Q, nthreads, [&](int64_t b, int64_t e) {
    GraphDijkstra D(idx, sep, P, K);
    for (int64_t i = b; i < e; ++i)
    {
      double s = 0; int64_t n = 0;
      double hi = std::max((double)hist_range, 1.0e-30), bin_size = nbins > 0 ? (hi - 0.0) / nbins : 1.0;
      if (hist) for (int j = 0; j < nbins; ++j) hist[i * nbins + j] = 0;
      if (seeds[i] >= 0 && seeds[i] < P)
      {
        D.run({{seeds[i], 0.0}}, max_distance < 0 ? -1.0 : (double)max_distance, [&](int64_t, double d) {
          s += d; ++n;
          if (hist)
          {
            int64_t bin = (int64_t)std::floor((d - 0.0) / bin_size);
            bin = bin < 0 ? 0 : (bin > nbins - 1 ? nbins - 1 : bin);
            hist[i * nbins + bin] += 1.0;
          }
          return false;
        });
        D.reset();
      }
      if (sum) sum[i] = s;
      if (count) count[i] = n;
    }
  });
}

// FurthestPointSampling::subsample (FurthestPointSampling.cpp:48-125): sample 0, then repeatedly the vertex the multi-source search
// from the selected set reports last -- unreachable vertices are reported after the search in vertex order (ShortestPaths.hpp:350-362).
int64_t orc_graph_fps(int64_t const * idx, float const * sep, int64_t P, int K, int64_t n_desired, int64_t * selected)
{
  std::vector<std::pair<int64_t, double>> region; std::vector<char> in_region((size_t)P, 0);
  for (int64_t i = 0; i < n_desired; ++i)
  {
    int64_t furthest = -1;
    if (region.empty()) furthest = 0;
    else
    {
      GraphDijkstra D(idx, sep, P, K);
      // equal distances: the higher index is reported last (canonical; the reference: its heap's order)
      double best = -1; int64_t bi = -1;
      D.run(region, -1.0, [&](int64_t v, double d) { if (d > best || (d == best && v > bi)) { best = d; bi = v; } return false; });
      furthest = bi;
      for (int64_t v = 0; v < P; ++v) if (!D.black[(size_t)v]) furthest = v;
      if (furthest < 0 || in_region[(size_t)furthest]) return i;
    }
    selected[i] = furthest; region.push_back(std::make_pair(furthest, 0.0)); in_region[(size_t)furthest] = 1;
  }
  return n_desired;
}

}  // extern "C"
