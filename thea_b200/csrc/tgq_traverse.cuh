// tgq_traverse.cuh -- single-query traversals of the device LBVH (one query per thread).
//
// Each routine computes the *canonical* answer of the corresponding BvhN query (SURVEY.md findings 4-5): the optimum
// over ALL elements of the reference's per-element fp32 function, ties broken by the lowest element index.  The tree
// only prunes: a subtree is skipped when a conservative bound proves that none of its elements can win or tie.
//   closest_query  <- BvhN::closestElement / closestPair        (Algorithms/BvhN.hpp:1049-1114,1866-1893,2036-2089)
//   knn_query      <- BvhN::kClosestPairs                       (BvhN.hpp:1137-1146, functor :474-500)
//   ball_graph_query <- processRange<Ball3> + NeighborFunctor   (BvhN.hpp:2118-2171, PointSet3.hpp:586-681)
//   ray_query      <- BvhN::rayStructureIntersection / rayIntersectionTime / rayIntersects (BvhN.hpp:1266-1316,2188-2291)
#pragma once
#include "tgq_bvh.cuh"

namespace tgq {

#define TGQ_STACK 128            // > 63 Morton bits + 32 tie-break bits: ordered DFS never holds more than depth entries
#define TGQ_LEAF_BITS 4
#define TGQ_LEAF_MASK 15
#define TGQ_LB_DEFLATE 0.99999905f   // 1 - 2^-20: lower bounds are shrunk so fp32 rounding can never over-prune

// Visit counters: compiled out everywhere except the test-only host build, which uses them to study pruning quality.
#ifndef TGQ_COUNT_NODE
#define TGQ_COUNT_NODE()
#define TGQ_COUNT_PRIM()
#endif

struct BvhView
{
  NodeRec const * nodes;   // n_nodes inner nodes, root = 0
  TriRec const * tris;     // Morton-ordered triangle records (triangle structures)
  f4 const * pts;          // Morton-ordered points (point structures)
  f4 const * nrm;          // Morton-ordered normals or null
  int n_nodes;
  int n_prims;
};

TGQ_HD f4 ldq(f4 const * p)
{
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

TGQ_HD NodeRec load_node(BvhView const & b, int i)
{
  NodeRec r; f4 const * p = (f4 const *)(b.nodes + i);
  r.q0 = ldq(p); r.q1 = ldq(p + 1); r.q2 = ldq(p + 2); r.q3 = ldq(p + 3);
  return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// Closest element.  best_d2 enters as the (exclusive) squared bound or +inf, best_idx as -1.
// ---------------------------------------------------------------------------------------------------------------------
template <bool TRI>
TGQ_HD void closest_leaf(BvhView const & b, int first, int count, v3 p, int skip, float & best_d2, int & best_idx, v3 & best_cp)
{
  for (int k = first; k < first + count; ++k)
  {
    int idx; v3 cp; float d2;
    TGQ_COUNT_PRIM();
    if (TRI)
    {
      f4 const * t = (f4 const *)(b.tris + k);
      f4 a = ldq(t), bq = ldq(t + 1), c = ldq(t + 2), d = ldq(t + 3);
      idx = f2i(a.w);
      d2 = tri_point_sqdist(mk3(a.x, a.y, a.z), mk3(bq.x, bq.y, bq.z), mk3(c.x, c.y, c.z), mk3(d.x, d.y, d.z), c.w, f2i(bq.w), p, cp);
    }
    else
    {
      f4 e = ldq(b.pts + k);
      idx = f2i(e.w); cp = mk3(e.x, e.y, e.z);
      d2 = sqn3(sub3(cp, p));
    }
    if (idx == skip) continue;
    if (d2 < best_d2 || (d2 == best_d2 && idx < best_idx)) { best_d2 = d2; best_idx = idx; best_cp = cp; }
  }
}

template <bool TRI>
TGQ_HD void closest_query(BvhView const & b, v3 p, int skip, float & best_d2, int & best_idx, v3 & best_cp)
{
  if (b.n_nodes <= 0) return;
  int stk_node[TGQ_STACK]; float stk_lb[TGQ_STACK];
  int sp = 0;
  int cur = 0;
  for (;;)
  {
    while (cur >= 0)
    {
      NodeRec nd = load_node(b, cur);
      TGQ_COUNT_NODE();
      v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
      float lb0 = box_point_sqdist(lo0, hi0, p) * TGQ_LB_DEFLATE;
      float lb1 = box_point_sqdist(lo1, hi1, p) * TGQ_LB_DEFLATE;
      int c0 = f2i(nd.q3.x), c1 = f2i(nd.q3.y);
      bool go0 = lb0 <= best_d2, go1 = lb1 <= best_d2;
      if (go0 && go1)
      {
        bool swap = lb1 < lb0;
        int nearc = swap ? c1 : c0, farc = swap ? c0 : c1;
        stk_node[sp] = farc; stk_lb[sp] = swap ? lb0 : lb1; ++sp;
        cur = nearc;
      }
      else if (go0) cur = c0;
      else if (go1) cur = c1;
      else goto pop;
    }
    {
      int code = ~cur;
      closest_leaf<TRI>(b, code >> TGQ_LEAF_BITS, code & TGQ_LEAF_MASK, p, skip, best_d2, best_idx, best_cp);
    }
  pop:
    for (;;)
    {
      if (sp == 0) return;
      --sp;
      if (stk_lb[sp] <= best_d2) { cur = stk_node[sp]; break; }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Closest element with "pancake" bounds.  AABBs of a smooth surface that is oblique to the axes stick out towards an
// off-surface query by about their own size, so a far query must open every box in a wide disc around its foot point
// (hundreds of leaves on a 1M-triangle mesh).  Each child therefore also carries a slab {x : smin <= dir.x <= smax}
// around the mean normal of its triangles and the radius R of its elements about the box centre c.  For any x in the
// child, |p - x|^2 = (dir.(p - x))^2 + |(p - x)_perp|^2 >= slab_dist^2 + max(0, |(p - c)_perp| - R)^2, a bound that
// keeps the lateral offset the box bound loses.  Both bounds are valid, the traversal prunes on their maximum.
//   SlabRec (48 B): a = (dir0.xyz, smin0)  b = (smax0, dir1.xyz)  c = (smin1, smax1, R0, R1)
// ---------------------------------------------------------------------------------------------------------------------
struct SlabRec { f4 a, b, c; };

TGQ_HD float pancake_sqdist(v3 lo, v3 hi, v3 dir, float smin, float smax, float R, v3 p)
{
  float s = dot3(dir, p);
  float sd = fmaxf(fmaxf(s - smax, smin - s), 0.0f);
  v3 c = mul3(add3(lo, hi), 0.5f);
  v3 pc = sub3(p, c);
  float sc = dot3(dir, pc);
  float perp2 = fmaxf(sqn3(pc) - sc * sc, 0.0f);
  float lat = fmaxf(sqrtf(perp2) * 0.999999f - R, 0.0f);
  return sd * sd + lat * lat;
}

template <bool TRI>
TGQ_HD void closest_query_slab(BvhView const & b, SlabRec const * slabs, v3 p, int skip, float & best_d2, int & best_idx, v3 & best_cp)
{
  if (b.n_nodes <= 0) return;
  int stk_node[TGQ_STACK]; float stk_lb[TGQ_STACK];
  int sp = 0;
  int cur = 0;
  for (;;)
  {
    while (cur >= 0)
    {
      NodeRec nd = load_node(b, cur);
      TGQ_COUNT_NODE();
      v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
      float lb0 = box_point_sqdist(lo0, hi0, p) * TGQ_LB_DEFLATE;
      float lb1 = box_point_sqdist(lo1, hi1, p) * TGQ_LB_DEFLATE;
      int c0 = f2i(nd.q3.x), c1 = f2i(nd.q3.y);
      bool go0 = lb0 <= best_d2, go1 = lb1 <= best_d2;
      if (go0 || go1)
      {
        f4 const * sq = (f4 const *)(slabs + cur);
        f4 sa = ldq(sq), sb = ldq(sq + 1), sc = ldq(sq + 2);
        if (go0)
        {
          lb0 = fmaxf(lb0, pancake_sqdist(lo0, hi0, mk3(sa.x, sa.y, sa.z), sa.w, sb.x, sc.z, p) * TGQ_LB_DEFLATE);
          go0 = lb0 <= best_d2;
        }
        if (go1)
        {
          lb1 = fmaxf(lb1, pancake_sqdist(lo1, hi1, mk3(sb.y, sb.z, sb.w), sc.x, sc.y, sc.w, p) * TGQ_LB_DEFLATE);
          go1 = lb1 <= best_d2;
        }
      }
      if (go0 && go1)
      {
        bool swap = lb1 < lb0;
        int nearc = swap ? c1 : c0, farc = swap ? c0 : c1;
        stk_node[sp] = farc; stk_lb[sp] = swap ? lb0 : lb1; ++sp;
        cur = nearc;
      }
      else if (go0) cur = c0;
      else if (go1) cur = c1;
      else goto pop;
    }
    {
      int code = ~cur;
      closest_leaf<TRI>(b, code >> TGQ_LEAF_BITS, code & TGQ_LEAF_MASK, p, skip, best_d2, best_idx, best_cp);
    }
  pop:
    for (;;)
    {
      if (sp == 0) return;
      --sp;
      if (stk_lb[sp] <= best_d2) { cur = stk_node[sp]; break; }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Bounded ascending set of (key, idx) pairs: the k-NN "heap" (BoundedSortedArray.hpp:185-265 semantics with the canonical
// total order).  Storage is strided so that a thread block can keep one set per thread in shared memory without bank
// conflicts (element j of thread t lives at base[j * stride]).
// ---------------------------------------------------------------------------------------------------------------------
struct SortedSet
{
  float * key; int * idx; int stride; int cap; int n;
  TGQ_HD void init(float * k, int * i, int stride_, int cap_) { key = k; idx = i; stride = stride_; cap = cap_; n = 0; }
  TGQ_HD bool full() const { return n >= cap; }
  TGQ_HD float worst_key() const { return key[(n - 1) * stride]; }
  TGQ_HD int worst_idx() const { return idx[(n - 1) * stride]; }
  TGQ_HD bool accepts(float k, int i) const
  { return cap > 0 && (n < cap || k < worst_key() || (k == worst_key() && i < worst_idx())); }
  TGQ_HD void insert(float k, int i)   // precondition: accepts(k, i)
  {
    int j = (n < cap) ? n : cap - 1;
    while (j > 0)
    {
      float kj = key[(j - 1) * stride]; int ij = idx[(j - 1) * stride];
      if (kj < k || (kj == k && ij < i)) break;
      key[j * stride] = kj; idx[j * stride] = ij;
      --j;
    }
    key[j * stride] = k; idx[j * stride] = i;
    if (n < cap) ++n;
  }
};

// k nearest points to p: k smallest (d2, idx); d2 < bound_ub (exclusive) when bounded.
TGQ_HD void knn_query(BvhView const & b, v3 p, int skip, float bound_ub, SortedSet & set)
{
  if (b.n_nodes <= 0 || set.cap <= 0) return;
  int stk_node[TGQ_STACK]; float stk_lb[TGQ_STACK];
  int sp = 0;
  int cur = 0;
  float worst = bound_ub;     // prune radius: k-th best once the set is full, else the bound (or +inf)
  for (;;)
  {
    while (cur >= 0)
    {
      NodeRec nd = load_node(b, cur);
      v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
      float lb0 = box_point_sqdist(lo0, hi0, p) * TGQ_LB_DEFLATE;
      float lb1 = box_point_sqdist(lo1, hi1, p) * TGQ_LB_DEFLATE;
      int c0 = f2i(nd.q3.x), c1 = f2i(nd.q3.y);
      bool go0 = lb0 <= worst, go1 = lb1 <= worst;
      if (go0 && go1)
      {
        bool swap = lb1 < lb0;
        stk_node[sp] = swap ? c0 : c1; stk_lb[sp] = swap ? lb0 : lb1; ++sp;
        cur = swap ? c1 : c0;
      }
      else if (go0) cur = c0;
      else if (go1) cur = c1;
      else goto pop;
    }
    {
      int code = ~cur; int first = code >> TGQ_LEAF_BITS, count = code & TGQ_LEAF_MASK;
      for (int k = first; k < first + count; ++k)
      {
        f4 e = ldq(b.pts + k);
        int idx = f2i(e.w);
        float d2 = sqn3(sub3(mk3(e.x, e.y, e.z), p));
        if (idx == skip || !(d2 < bound_ub)) continue;
        if (set.accepts(d2, idx))
        {
          set.insert(d2, idx);
          if (set.full()) worst = set.worst_key();
        }
      }
    }
  pop:
    for (;;)
    {
      if (sp == 0) return;
      --sp;
      if (stk_lb[sp] <= worst) { cur = stk_node[sp]; break; }
    }
  }
}

// One round of PointSet3::findSampleNeighbors for sample `self` (position c, normal nc): every point j != self with
// sq(x_j - c) <= r2 and (no normals or n_c . n_j >= min_dot) is offered with key fastSqrt(d2); the set keeps the
// max_degree smallest (key, idx).  The ball/box test uses the reference arithmetic on padded boxes (a superset).
TGQ_HD void ball_round(BvhView const & b, v3 c, v3 nc, bool has_n, float min_dot, int self, float r2, SortedSet & set)
{
  set.n = 0;
  if (b.n_nodes <= 0) return;
  int stk_node[TGQ_STACK];
  int sp = 0;
  int cur = 0;
  for (;;)
  {
    while (cur >= 0)
    {
      NodeRec nd = load_node(b, cur);
      v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
      bool go0 = box_point_sqdist(lo0, hi0, c) * TGQ_LB_DEFLATE <= r2;
      bool go1 = box_point_sqdist(lo1, hi1, c) * TGQ_LB_DEFLATE <= r2;
      int c0 = f2i(nd.q3.x), c1 = f2i(nd.q3.y);
      if (go0 && go1) { stk_node[sp++] = c1; cur = c0; }
      else if (go0) cur = c0;
      else if (go1) cur = c1;
      else goto pop;
    }
    {
      int code = ~cur; int first = code >> TGQ_LEAF_BITS, count = code & TGQ_LEAF_MASK;
      for (int k = first; k < first + count; ++k)
      {
        f4 e = ldq(b.pts + k);
        int idx = f2i(e.w);
        if (idx == self) continue;
        v3 diff = sub3(mk3(e.x, e.y, e.z), c);
        float d2 = sqn3(diff);
        if (!(d2 <= r2)) continue;
        if (has_n)
        {
          f4 nj = ldq(b.nrm + k);
          if (dot3(nc, mk3(nj.x, nj.y, nj.z)) < min_dot) continue;
        }
        float sep = fast_sqrt(d2);
        if (set.accepts(sep, idx)) set.insert(sep, idx);
      }
    }
  pop:
    if (sp == 0) return;
    cur = stk_node[--sp];
  }
}

TGQ_HD void ball_graph_query(BvhView const & b, v3 c, v3 nc, bool has_n, float min_dot, int self, float init_radius,
                             int rounds, int min_degree, SortedSet & set)
{
  float radius = init_radius;
  for (int it = 0; it < rounds; ++it)
  {
    ball_round(b, c, nc, has_n, min_dot, self, radius * radius, set);
    if (set.n >= min_degree) break;
    radius *= 2.0f;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Ray first hit.  best_t enters as max_time (negative = unbounded), best_idx as -1; best_k returns the Morton position
// of the hit triangle (for its cached normal).  A hit is accepted iff t >= 0 and
// (unbounded or t < best_t) or it ties the current best with a lower index.  ANY == true stops at the first hit found
// (BvhN::rayIntersects returns the same boolean as a full first-hit search, BvhN.hpp:1266-1269).
// ---------------------------------------------------------------------------------------------------------------------
struct RaySlab { v3 o, inv; };

// Conservative slab test on a (padded) box: returns the entry time clamped to >= 0, or -1 for a miss.  NaNs from
// 0 * inf (origin on a face plane, zero direction component) are dropped by fminf / fmaxf, which errs towards "hit".
TGQ_HD float slab_entry(RaySlab const & r, v3 lo, v3 hi)
{
  float x0 = (lo.x - r.o.x) * r.inv.x, x1 = (hi.x - r.o.x) * r.inv.x;
  float y0 = (lo.y - r.o.y) * r.inv.y, y1 = (hi.y - r.o.y) * r.inv.y;
  float z0 = (lo.z - r.o.z) * r.inv.z, z1 = (hi.z - r.o.z) * r.inv.z;
  float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
  float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
  tn *= 0.9999f;                 // generous slack: the per-triangle time is what decides, this only prunes
  tf = tf * 1.0001f;
  return (tn <= tf) ? tn : -1.0f;
}

template <bool ANY>
TGQ_HD void ray_query(BvhView const & b, v3 o, v3 d, float & best_t, int & best_idx, int & best_k)
{
  if (b.n_nodes <= 0) return;
  RaySlab r; r.o = o; r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  int stk_node[TGQ_STACK]; float stk_t[TGQ_STACK];
  int sp = 0;
  int cur = 0;
  for (;;)
  {
    while (cur >= 0)
    {
      NodeRec nd = load_node(b, cur);
      v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
      float t0 = slab_entry(r, lo0, hi0), t1 = slab_entry(r, lo1, hi1);
      int c0 = f2i(nd.q3.x), c1 = f2i(nd.q3.y);
      bool go0 = t0 >= 0.0f && (best_t < 0.0f || t0 <= best_t);
      bool go1 = t1 >= 0.0f && (best_t < 0.0f || t1 <= best_t);
      if (go0 && go1)
      {
        bool swap = t1 < t0;
        stk_node[sp] = swap ? c0 : c1; stk_t[sp] = swap ? t0 : t1; ++sp;
        cur = swap ? c1 : c0;
      }
      else if (go0) cur = c0;
      else if (go1) cur = c1;
      else goto pop;
    }
    {
      int code = ~cur; int first = code >> TGQ_LEAF_BITS, count = code & TGQ_LEAF_MASK;
      for (int k = first; k < first + count; ++k)
      {
        f4 const * t = (f4 const *)(b.tris + k);
        f4 a = ldq(t), bq = ldq(t + 1), c = ldq(t + 2);
        v3 v0 = mk3(a.x, a.y, a.z);
        float tm = ray_tri_time(o, d, v0, sub3(mk3(bq.x, bq.y, bq.z), v0), sub3(mk3(c.x, c.y, c.z), v0));
        int idx = f2i(a.w);
        if (tm >= 0.0f && (best_idx < 0 ? (best_t < 0.0f || tm < best_t) : (tm < best_t || (tm == best_t && idx < best_idx))))
        { best_t = tm; best_idx = idx; best_k = k; if (ANY) return; }
      }
    }
  pop:
    for (;;)
    {
      if (sp == 0) return;
      --sp;
      if (best_t < 0.0f || stk_t[sp] <= best_t) { cur = stk_node[sp]; break; }
    }
  }
}

}  // namespace tgq
