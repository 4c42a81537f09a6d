// tgq_graph.cuh -- shortest paths on the sample graph: what the reference does with
// ShortestPaths::dijkstraWithCallback (Algorithms/ShortestPaths.hpp:166-353) on PointSet3's proximity graph:
//   * PointSet3::extractOriginalAdjacencies (Algorithms/PointSet3.cpp:158-178): per original sample, the first max_degree ORIGINAL
//     samples in Dijkstra order over the graph of original + dense samples -- the oversampling contraction;
//   * AverageDistance / LocalDistanceHistogram with DistanceType::GEODESIC (SurfaceFeatures/Local/AverageDistance.cpp:84-118,
//     LocalDistanceHistogram.cpp:90-137): statistics of all samples within a geodesic limit of a seed sample;
//   * FurthestPointSampling::subsample (Algorithms/FurthestPointSampling.cpp:48-125): repeatedly the sample furthest from the
//     selected set.
// The graph is the (P x max_degree) neighbour index / separation pair that tgq_ball_graph produces (-1 padded rows, directed).
//
// Distances are the reference's: d[v] = min over paths of the left-to-right DOUBLE sum of the float separations.  With a monotone
// rounding (fl(a + w) is non-decreasing in a) that minimum is the unique fixed point of d[v] = min_u fl(d[u] + w(u, v)), so a
// label-correcting iteration in any order converges to bit-identical values -- no priority queue is needed on the device.  One warp
// works on one source: its 32 lanes relax (node, edge) pairs of the current frontier, distances live in the warp's own row of a
// global scratch array (64-bit atomicMin on the bit pattern: non-negative doubles order like integers), touched nodes are listed so
// that the row is reset in time proportional to the neighbourhood, not to P.  Bounded searches never store a distance beyond the
// limit: every node within the limit has a shortest path all of whose prefixes are within it.
#pragma once
#include "tgq_math.cuh"

namespace tgq {

#define TGQ_GRAPH_INF 0x7ff0000000000000ull
#define TGQ_GRAPH_WARPS 8               // warps per block of the per-source kernels

struct GraphView { int const * idx; float const * sep; int P; int K; };

// Per-warp scratch rows in global memory.
struct GraphScratch
{
  unsigned long long * dist;   // [workers x P] bit patterns of doubles, TGQ_GRAPH_INF when untouched
  int * stamp;                 // [workers x P] round in which the node was last queued, -1 when untouched
  int * touched;               // [workers x P]
  int * queue;                 // [workers x 2 x qlen], qlen = max(P, 2 * TGQ_MAX_K): the second half also holds the picks of the contraction
  int qlen;
};

// Distances are updated with atomics (performed in L2): every read goes to L2 as well, never to a possibly stale L1 line.
__device__ __forceinline__ unsigned long long ld_dist(unsigned long long const * p) { return __ldcg(p); }
__device__ __forceinline__ double bits_to_double(unsigned long long b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ unsigned long long double_to_bits(double d) { return (unsigned long long)__double_as_longlong(d); }

// Relaxes every edge out of the nodes of `cur` (n_cur entries) whose distance is within the limit; returns the size of the next
// frontier written to `nxt`.  limit < 0: unbounded.  *cut is set when a relaxation was dropped because it exceeded the limit.
__device__ __forceinline__ int graph_relax_round(GraphView const & g, unsigned long long * dist, int * stamp, int * touched, int & n_touched,
                                                 int const * cur, int n_cur, int * nxt, int round, double limit, bool & cut)
{
  unsigned const FULL = 0xffffffffu;
  int const lane = threadIdx.x & 31;
  int const kk = g.K <= 1 ? 1 : (g.K <= 2 ? 2 : (g.K <= 4 ? 4 : (g.K <= 8 ? 8 : (g.K <= 16 ? 16 : 32))));    // edge slots per node per pass
  int const per = 32 / kk;                                                                                    // nodes per pass
  int n_next = 0;
  for (int e0 = 0; e0 < g.K; e0 += kk)
    for (int base = 0; base < n_cur; base += per)
    {
      int slot = base + lane / kk, e = e0 + lane % kk;
      bool have = slot < n_cur && e < g.K;
      int v = -1; double nd = 0.0;
      if (have)
      {
        int u = cur[slot];
        v = g.idx[(size_t)u * g.K + e];
        if (v >= 0)
        {
          double du = bits_to_double(ld_dist(dist + u));
          nd = du + (double)g.sep[(size_t)u * g.K + e];
          if (limit >= 0.0 && du > limit) v = -1;                       // the reference stops at the first vertex beyond the limit
          else if (limit >= 0.0 && nd > limit) { v = -1; cut = true; }
        }
      }
      bool fresh = false, enq = false;
      if (v >= 0)
      {
        unsigned long long nb = double_to_bits(nd);
        unsigned long long old = atomicMin(dist + v, nb);
        if (nb < old)
        {
          fresh = old == TGQ_GRAPH_INF;
          enq = atomicExch(stamp + v, round) != round;
        }
      }
      unsigned mf = __ballot_sync(FULL, fresh), mq = __ballot_sync(FULL, enq);
      if (fresh) touched[n_touched + __popc(mf & ((1u << lane) - 1u))] = v;
      if (enq) nxt[n_next + __popc(mq & ((1u << lane) - 1u))] = v;
      n_touched += __popc(mf); n_next += __popc(mq);
    }
  __syncwarp();
  return n_next;
}

// Runs the label-correcting iteration from the current frontier until it is empty.
__device__ __forceinline__ void graph_converge(GraphView const & g, unsigned long long * dist, int * stamp, int * touched, int & n_touched,
                                               int * q0, int * q1, int n_front, double limit, bool & cut, int & round)
{
  int * cur = q0; int * nxt = q1;
  while (n_front > 0)
  {
    ++round;
    n_front = graph_relax_round(g, dist, stamp, touched, n_touched, cur, n_front, nxt, round, limit, cut);
    int * t = cur; cur = nxt; nxt = t;
  }
}

__device__ __forceinline__ void graph_reset(unsigned long long * dist, int * stamp, int const * touched, int n_touched)
{
  int const lane = threadIdx.x & 31;
  for (int j = lane; j < n_touched; j += 32) { int v = touched[j]; dist[v] = TGQ_GRAPH_INF; stamp[v] = -1; }
  __syncwarp();
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long x)
{
#pragma unroll
  for (int m = 16; m > 0; m >>= 1)
  {
    unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)x, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(x >> 32), m);
    unsigned long long y = ((unsigned long long)hi << 32) | lo;
    x = y < x ? y : x;
  }
  return x;
}

__global__ void k_graph_init(unsigned long long * dist, int * stamp, size_t n)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { dist[i] = TGQ_GRAPH_INF; stamp[i] = -1; }
}

// ---- geodesic neighbourhood statistics ---------------------------------------------------------------------------------------
// Per query: seed sample, limit (max_distance as double, or < 0 for the whole component).  out_sum / out_count: sum of the distances
// and number of the vertices the reference's callback sees (every vertex with d <= limit, the seed with d = 0 included);
// hist (optional, num_bins per query): Histogram::insert(d) with range [0, max(max_distance, 1e-30)] (Histogram.hpp:110-150).
// The sum runs over the touched list in warp order, not in ascending distance as the reference's callback order: equal to ~1e-15
// relative (stated tolerance of the tests: 1e-12).
__global__ void __launch_bounds__(32 * TGQ_GRAPH_WARPS)
k_graph_geodesic(GraphView g, GraphScratch sc, int const * __restrict__ seeds, int Q, float max_distance, float hist_range, int num_bins, int * counter,
                 double * out_sum, int * out_count, unsigned * hist)
{
  unsigned const FULL = 0xffffffffu;
  int const lane = threadIdx.x & 31;
  size_t const worker = (size_t)blockIdx.x * TGQ_GRAPH_WARPS + (threadIdx.x >> 5);
  unsigned long long * dist = sc.dist + worker * g.P; int * stamp = sc.stamp + worker * g.P; int * touched = sc.touched + worker * g.P;
  int * q0 = sc.queue + worker * 2 * (size_t)sc.qlen; int * q1 = q0 + sc.qlen;
  double const limit = max_distance < 0.0f ? -1.0 : (double)max_distance;
  double const hi = fmax((double)hist_range, 1.0e-30), bin_size = hi / (double)(num_bins > 0 ? num_bins : 1);
  for (;;)
  {
    int q = 0;
    if (lane == 0) q = atomicAdd(counter, 1);
    q = __shfl_sync(FULL, q, 0);
    if (q >= Q) break;
    int s = seeds[q];
    int n_touched = 0, round = 0; bool cut = false;
    double sum = 0.0; int cnt = 0;
    if (s >= 0 && s < g.P)
    {
      if (lane == 0) { dist[s] = 0ull; stamp[s] = 0; touched[0] = s; q0[0] = s; }
      n_touched = 1;
      __syncwarp();
      graph_converge(g, dist, stamp, touched, n_touched, q0, q1, 1, limit, cut, round);
      for (int j = lane; j < n_touched; j += 32)
      {
        double d = bits_to_double(ld_dist(dist + touched[j]));
        if (limit < 0.0 || d <= limit)
        {
          sum += d; ++cnt;
          if (hist)
          {
            long long b = (long long)floor(d / bin_size);
            b = b < 0 ? 0 : (b > num_bins - 1 ? num_bins - 1 : b);
            atomicAdd(hist + (size_t)q * num_bins + b, 1u);
          }
        }
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) { sum += __shfl_xor_sync(FULL, sum, m); cnt += __shfl_xor_sync(FULL, cnt, m); }
      graph_reset(dist, stamp, touched, n_touched);
    }
    if (lane == 0) { if (out_sum) out_sum[q] = sum; if (out_count) out_count[q] = cnt; }
  }
}

// ---- oversampling contraction -------------------------------------------------------------------------------------------------
// Per original sample s < n_orig: the max_degree smallest (d, index) over the ORIGINAL samples v != s reachable from s, d the graph
// distance over all samples -- what DijkstraCallback (PointSet3.cpp:127-147) collects before it stops the search.  The search radius
// starts at 2.5 x the longest edge out of s and doubles until max_degree originals lie within it (or nothing was cut off).  Output
// rows as the reference's neighbour set orders them: ascending (float(d), index); -1 / -1.0f padded.
__global__ void __launch_bounds__(32 * TGQ_GRAPH_WARPS)
k_graph_contract(GraphView g, GraphScratch sc, int n_orig, int * counter, int * out_idx, float * out_sep, int * out_count)
{
  unsigned const FULL = 0xffffffffu;
  int const lane = threadIdx.x & 31, K = g.K;
  size_t const worker = (size_t)blockIdx.x * TGQ_GRAPH_WARPS + (threadIdx.x >> 5);
  unsigned long long * dist = sc.dist + worker * g.P; int * stamp = sc.stamp + worker * g.P; int * touched = sc.touched + worker * g.P;
  int * q0 = sc.queue + worker * 2 * (size_t)sc.qlen; int * q1 = q0 + sc.qlen;
  for (;;)
  {
    int s = 0;
    if (lane == 0) s = atomicAdd(counter, 1);
    s = __shfl_sync(FULL, s, 0);
    if (s >= n_orig) break;
    float longest = 0.0f;
    for (int e = lane; e < K; e += 32) if (g.idx[(size_t)s * K + e] >= 0) longest = fmaxf(longest, g.sep[(size_t)s * K + e]);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) longest = fmaxf(longest, __shfl_xor_sync(FULL, longest, m));
    double limit = longest > 0.0f ? 2.5 * (double)longest : -1.0;          // all edges of length zero: search the whole component
    if (lane == 0) { dist[s] = 0ull; stamp[s] = 0; touched[0] = s; q0[0] = s; }
    int n_touched = 1, round = 0, n_front = 1;
    __syncwarp();
    for (;;)
    {
      bool cut = false;
      graph_converge(g, dist, stamp, touched, n_touched, q0, q1, n_front, limit, cut, round);
      int c = 0;
      for (int j = lane; j < n_touched; j += 32) { int v = touched[j]; c += (v < n_orig && v != s) ? 1 : 0; }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) c += __shfl_xor_sync(FULL, c, m);
      cut = __any_sync(FULL, cut);
      if (c >= K || !cut) break;
      limit *= 2.0;
      // every touched node may now relax edges that were cut off: they all go back on the frontier
      for (int j = lane; j < n_touched; j += 32) q0[j] = touched[j];
      n_front = n_touched;
      __syncwarp();
    }
    // selection: max_degree rounds of a warp-wide minimum over (distance bits, index) of the touched originals above the last pick.
    // Two originals at the same double distance are ordered by index (the reference: by its heap's pop order).
    unsigned long long last_d = 0ull; int last_v = -1; int found = 0;
    unsigned long long my_d = 0ull; int my_v = -1;                     // lane j keeps pick j (K <= 32 per pass; K <= 64 overall: two slots)
    unsigned long long my_d2 = 0ull; int my_v2 = -1;
    for (int r = 0; r < K; ++r)
    {
      unsigned long long best_d = ~0ull; int best_v = 0x7fffffff;
      for (int j = lane; j < n_touched; j += 32)
      {
        int v = touched[j];
        if (v >= n_orig || v == s) continue;
        unsigned long long d = ld_dist(dist + v);
        bool after = last_v < 0 || d > last_d || (d == last_d && v > last_v);
        if (after && (d < best_d || (d == best_d && v < best_v))) { best_d = d; best_v = v; }
      }
      // warp minimum of (d, v): reduce d first, then v among the lanes holding that d
      unsigned long long md = warp_min_u64(best_d);
      int cand = best_d == md ? best_v : 0x7fffffff;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) cand = min(cand, __shfl_xor_sync(FULL, cand, m));
      if (md == ~0ull) break;
      last_d = md; last_v = cand; ++found;
      if (r < 32) { if (lane == r) { my_d = md; my_v = cand; } }
      else if (lane == r - 32) { my_d2 = md; my_v2 = cand; }
    }
    // the reference's neighbour set orders by (float separation, sample address): a stable pass over the picks, which are already
    // ascending in (double distance, index) -- casting to float can only merge neighbouring distances, never reorder them, and equal
    // floats then fall back to the index order only if the double order disagreed with it
    __syncwarp();
    // write picks to the scratch queue, then order by (float, index) with an insertion pass in lane 0 (K <= 64 entries)
    if (lane < min(found, 32)) { q1[2 * lane] = my_v; q1[2 * lane + 1] = __float_as_int((float)bits_to_double(my_d)); }
    if (found > 32 && lane < found - 32) { q1[2 * (lane + 32)] = my_v2; q1[2 * (lane + 32) + 1] = __float_as_int((float)bits_to_double(my_d2)); }
    __syncwarp();
    if (lane == 0)
    {
      for (int a = 1; a < found; ++a)
      {
        int v = q1[2 * a]; float f = __int_as_float(q1[2 * a + 1]); int b = a;
        while (b > 0)
        {
          int pv = q1[2 * (b - 1)]; float pf = __int_as_float(q1[2 * (b - 1) + 1]);
          if (pf < f || (pf == f && pv < v)) break;
          q1[2 * b] = pv; q1[2 * b + 1] = __float_as_int(pf); --b;
        }
        q1[2 * b] = v; q1[2 * b + 1] = __float_as_int(f);
      }
      if (out_count) out_count[s] = found;
    }
    __syncwarp();
    for (int j = lane; j < K; j += 32)
    {
      bool ok = j < found;
      out_idx[(size_t)s * K + j] = ok ? q1[2 * j] : -1;
      out_sep[(size_t)s * K + j] = ok ? __int_as_float(q1[2 * j + 1]) : -1.0f;
    }
    __syncwarp();
    graph_reset(dist, stamp, touched, n_touched);
  }
}

// ---- furthest point sampling ----------------------------------------------------------------------------------------------------
// One block.  dist[v] = distance to the selected set, kept up to date incrementally: a new source only lowers the distances of the
// nodes that are now closer to it -- the same values as the reference's full multi-source Dijkstra per selection.  The next pick is
// the vertex the reference's callback sees LAST: with include_unreachable the unreachable vertices are reported after the search, in
// vertex order, so the highest-index unreachable vertex wins; otherwise the vertex of maximal distance (equal distances: highest
// index here, heap pop order in the reference).  Stops early, like the reference, when that vertex is already selected.
#define TGQ_FPS_THREADS 1024
__global__ void __launch_bounds__(TGQ_FPS_THREADS)
k_graph_fps(GraphView g, unsigned long long * dist, int * stamp, int * q0, int * q1, int n_desired, int * selected, int * n_selected)
{
  __shared__ int s_n, s_round, s_pick;
  __shared__ unsigned long long s_key[TGQ_FPS_THREADS / 32];
  __shared__ int s_idx[TGQ_FPS_THREADS / 32];
  int const tid = threadIdx.x, lane = tid & 31, K = g.K;
  for (int v = tid; v < g.P; v += TGQ_FPS_THREADS) { dist[v] = TGQ_GRAPH_INF; stamp[v] = -1; }
  if (tid == 0) { s_round = 0; s_pick = 0; }
  __syncthreads();
  int done = 0;
  for (int i = 0; i < n_desired; ++i)
  {
    int s = s_pick;
    if (tid == 0) { selected[i] = s; dist[s] = 0ull; q0[0] = s; s_n = 0; }
    done = i + 1;
    __syncthreads();
    int * cur = q0; int * nxt = q1; int n_cur = 1;
    while (n_cur > 0)
    {
      int round = s_round + 1;
      for (long long t = tid; t < (long long)n_cur * K; t += TGQ_FPS_THREADS)
      {
        int u = cur[t / K], e = (int)(t % K);
        int v = g.idx[(size_t)u * K + e];
        if (v < 0) continue;
        double nd = bits_to_double(ld_dist(dist + u)) + (double)g.sep[(size_t)u * K + e];
        unsigned long long nb = double_to_bits(nd);
        unsigned long long old = atomicMin(dist + v, nb);
        if (nb < old && atomicExch(stamp + v, round) != round) nxt[atomicAdd(&s_n, 1)] = v;
      }
      __syncthreads();
      n_cur = s_n;
      __syncthreads();
      if (tid == 0) { s_n = 0; s_round = round; }
      int * t_ = cur; cur = nxt; nxt = t_;
      __syncthreads();
    }
    if (i + 1 >= n_desired) break;
    // argmax of (distance bits, index)
    unsigned long long bk = 0ull; int bi = -1;
    for (int v = tid; v < g.P; v += TGQ_FPS_THREADS)
    {
      unsigned long long d = ld_dist(dist + v);
      if (bi < 0 || d > bk || (d == bk && v > bi)) { bk = d; bi = v; }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
    {
      unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)bk, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(bk >> 32), m);
      unsigned long long ok = ((unsigned long long)hi << 32) | lo; int oi = __shfl_xor_sync(0xffffffffu, bi, m);
      if (oi >= 0 && (bi < 0 || ok > bk || (ok == bk && oi > bi))) { bk = ok; bi = oi; }
    }
    if (lane == 0) { s_key[tid >> 5] = bk; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0)
    {
      unsigned long long k = 0ull; int b = -1;
      for (int w = 0; w < TGQ_FPS_THREADS / 32; ++w)
        if (s_idx[w] >= 0 && (b < 0 || s_key[w] > k || (s_key[w] == k && s_idx[w] > b))) { k = s_key[w]; b = s_idx[w]; }
      s_pick = (b >= 0 && k != 0ull) ? b : -1;            // distance 0: the furthest vertex is itself selected -- nothing left to pick
    }
    __syncthreads();
    if (s_pick < 0) break;
  }
  if (tid == 0) *n_selected = done;
}

}  // namespace tgq
