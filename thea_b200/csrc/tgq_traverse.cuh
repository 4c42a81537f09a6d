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
// "Pancake" bounds.  AABBs of a smooth surface that is oblique to the axes stick out towards an off-surface query by
// about their own size, so a far query would have to open every box in a wide disc around its foot point (hundreds of
// leaves on a 1M-triangle mesh).  Each child of a triangle-structure node therefore also carries a slab
// {x : smin <= dir.x <= smax} around the mean normal of its triangles and the radius R of its vertices about the centre
// c of its box.  For any x in the child
//     |p - x|^2 = (dir.(p - x))^2 + |(p - x)_perp|^2 >= slab_dist(p)^2 + max(0, |(p - c)_perp| - R)^2,
// which keeps the lateral offset that the box bound loses.  Box and pancake bounds are both valid; the traversal prunes
// on their maximum.  The same bound with the triangle's own plane and the radius about its centroid screens single
// triangles before the exact (and expensive) reference arithmetic is run.
//   SlabRec (48 B): a = (dir0.xyz, smin0)  b = (smax0, dir1.xyz)  c = (smin1, smax1, R0, R1)
// All stored quantities are padded outwards at build time (scene pad); `mq` absorbs the rounding of the query-side dot
// products (it grows with the magnitude of the query coordinates), and the result is deflated like the box bound.
// ---------------------------------------------------------------------------------------------------------------------
struct SlabRec { f4 a, b, c; };

TGQ_HD float query_margin(v3 p, float scene_pad) { return scene_pad + 1e-6f * (fabsf(p.x) + fabsf(p.y) + fabsf(p.z)); }

TGQ_HD float pancake_sqdist(v3 c, v3 dir, float smin, float smax, float R, v3 p, float mq)
{
  float s = dot3(dir, p);
  float sd = fmaxf(fmaxf(s - smax, smin - s) - mq, 0.0f);
  v3 pc = sub3(p, c);
  float sc = dot3(dir, pc);
  v3 perp = sub3(pc, mul3(dir, sc));
  float lat = fmaxf(sqrtf(sqn3(perp)) - R - mq, 0.0f);
  return (sd * sd + lat * lat) * TGQ_LB_DEFLATE;
}

// Accumulates one triangle into the slab statistics of a node: extent along dir and radius about c.
TGQ_HD void slab_accumulate(v3 dir, v3 c, v3 v0, v3 v1, v3 v2, float & mn, float & mx, float & r2)
{
  float s0 = dot3(dir, v0), s1 = dot3(dir, v1), s2 = dot3(dir, v2);
  mn = fminf(mn, fminf(s0, fminf(s1, s2))); mx = fmaxf(mx, fmaxf(s0, fmaxf(s1, s2)));
  r2 = fmaxf(r2, fmaxf(sqn3(sub3(v0, c)), fmaxf(sqn3(sub3(v1, c)), sqn3(sub3(v2, c)))));
}
TGQ_HD float slab_radius(float r2, float pad) { return sqrtf(r2) * 1.000001f + 2.0f * pad; }
TGQ_HD v3 box_centre(v3 lo, v3 hi) { return mul3(add3(lo, hi), 0.5f); }
// Screening radius of one triangle about its centroid (stored in TriRec.q3.w).
TGQ_HD float tri_screen_radius(v3 v0, v3 v1, v3 v2, float pad)
{
  v3 c = mul3(add3(add3(v0, v1), v2), 0.33333334f);
  float r2 = fmaxf(sqn3(sub3(v0, c)), fmaxf(sqn3(sub3(v1, c)), sqn3(sub3(v2, c))));
  return slab_radius(r2, pad);
}

// ---------------------------------------------------------------------------------------------------------------------
// Closest element.  best_d2 enters as the (exclusive) squared bound or +inf, best_idx as -1.
// SLAB: triangle structure with pancake bounds (slabs != null) and per-triangle screening.
// ---------------------------------------------------------------------------------------------------------------------
template <bool TRI, bool SLAB>
TGQ_HD void closest_leaf(BvhView const & b, int first, int count, v3 p, int skip, float mq, float & best_d2, int & best_idx, v3 & best_cp)
{
  for (int k = first; k < first + count; ++k)
  {
    int idx; v3 cp; float d2;
    if (TRI)
    {
      f4 const * t = (f4 const *)(b.tris + k);
      f4 a = ldq(t), bq = ldq(t + 1), c = ldq(t + 2), d = ldq(t + 3);
      v3 v0 = mk3(a.x, a.y, a.z), v1 = mk3(bq.x, bq.y, bq.z), v2 = mk3(c.x, c.y, c.z), n = mk3(d.x, d.y, d.z);
      if (SLAB)
      {
        // screen: plane distance and lateral distance to the triangle's bounding circle
        v3 cen = mul3(add3(add3(v0, v1), v2), 0.33333334f);
        float sdp = dot3(p, n) - c.w;
        float sd = fmaxf(fabsf(sdp) - mq, 0.0f);
        v3 pc = sub3(p, cen);
        v3 perp = sub3(pc, mul3(n, dot3(n, pc)));
        float lat = fmaxf(sqrtf(sqn3(perp)) - d.w - mq, 0.0f);
        if ((sd * sd + lat * lat) * TGQ_LB_DEFLATE > best_d2) continue;
      }
      TGQ_COUNT_PRIM();
      idx = f2i(a.w);
      d2 = tri_point_sqdist(v0, v1, v2, n, c.w, f2i(bq.w), p, cp);
    }
    else
    {
      TGQ_COUNT_PRIM();
      f4 e = ldq(b.pts + k);
      idx = f2i(e.w); cp = mk3(e.x, e.y, e.z);
      d2 = sqn3(sub3(cp, p));
    }
    if (idx == skip) continue;
    if (d2 < best_d2 || (d2 == best_d2 && idx < best_idx)) { best_d2 = d2; best_idx = idx; best_cp = cp; }
  }
}

// Resumable traversal state: the persistent kernels advance every lane by a bounded burst of steps, retire finished
// queries and hand fresh ones to the idle lanes, so a warp stays full even though query costs are heavy tailed.
template <bool TRI, bool SLAB>
struct ClosestState
{
  v3 p; float mq; int skip;
  float best_d2; int best_idx; v3 best_cp;
  int cur, sp;
  int stk_node[TGQ_STACK]; float stk_lb[TGQ_STACK];
};

template <bool TRI, bool SLAB>
TGQ_HD void closest_begin(ClosestState<TRI, SLAB> & s, BvhView const & b, float scene_pad, v3 p, int skip, float bound_ub)
{
  s.p = p; s.skip = skip; s.mq = SLAB ? query_margin(p, scene_pad) : 0.0f;
  s.best_d2 = bound_ub; s.best_idx = -1; s.best_cp = mk3(0.0f, 0.0f, 0.0f);
  s.sp = 0; s.cur = b.n_nodes > 0 ? 0 : -1;
  if (b.n_nodes <= 0) s.sp = -1;          // empty structure: nothing to do
}

// Advances by at most `budget` node steps; returns true when the query is finished.
template <bool TRI, bool SLAB>
TGQ_HD bool closest_run(ClosestState<TRI, SLAB> & s, BvhView const & b, SlabRec const * slabs, int budget)
{
  if (s.sp < 0) return true;
  v3 p = s.p; float mq = s.mq;
  int cur = s.cur, sp = s.sp;
  for (;;)
  {
    while (cur >= 0)
    {
      if (--budget < 0) { s.cur = cur; s.sp = sp; return false; }
      NodeRec nd = load_node(b, cur);
      TGQ_COUNT_NODE();
      v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
      float lb0 = box_point_sqdist(lo0, hi0, p) * TGQ_LB_DEFLATE;
      float lb1 = box_point_sqdist(lo1, hi1, p) * TGQ_LB_DEFLATE;
      int c0 = f2i(nd.q3.x), c1 = f2i(nd.q3.y);
      bool go0 = lb0 <= s.best_d2, go1 = lb1 <= s.best_d2;
      if (SLAB && (go0 || go1))
      {
        f4 const * sq = (f4 const *)(slabs + cur);
        f4 sa = ldq(sq), sb = ldq(sq + 1), sc = ldq(sq + 2);
        if (go0)
        {
          lb0 = fmaxf(lb0, pancake_sqdist(box_centre(lo0, hi0), mk3(sa.x, sa.y, sa.z), sa.w, sb.x, sc.z, p, mq));
          go0 = lb0 <= s.best_d2;
        }
        if (go1)
        {
          lb1 = fmaxf(lb1, pancake_sqdist(box_centre(lo1, hi1), mk3(sb.y, sb.z, sb.w), sc.x, sc.y, sc.w, p, mq));
          go1 = lb1 <= s.best_d2;
        }
      }
      if (go0 && go1)
      {
        bool swap = lb1 < lb0;
        int nearc = swap ? c1 : c0, farc = swap ? c0 : c1;
        s.stk_node[sp] = farc; s.stk_lb[sp] = swap ? lb0 : lb1; ++sp;
        cur = nearc;
      }
      else if (go0) cur = c0;
      else if (go1) cur = c1;
      else goto pop;
    }
    {
      int code = ~cur;
      closest_leaf<TRI, SLAB>(b, code >> TGQ_LEAF_BITS, code & TGQ_LEAF_MASK, p, s.skip, mq, s.best_d2, s.best_idx, s.best_cp);
    }
  pop:
    for (;;)
    {
      if (sp == 0) { s.sp = -1; return true; }
      --sp;
      if (s.stk_lb[sp] <= s.best_d2) { cur = s.stk_node[sp]; break; }
    }
  }
}

template <bool TRI, bool SLAB>
TGQ_HD void closest_query(BvhView const & b, SlabRec const * slabs, float scene_pad, v3 p, int skip, float & best_d2, int & best_idx,
                          v3 & best_cp)
{
  ClosestState<TRI, SLAB> s;
  closest_begin(s, b, scene_pad, p, skip, best_d2);
  closest_run(s, b, slabs, 0x7fffffff);
  best_d2 = s.best_d2; best_idx = s.best_idx; best_cp = s.best_cp;
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
struct RayState
{
  v3 o, d; RaySlab r;
  float best_t; int best_idx, best_k;
  int cur, sp;
  int stk_node[TGQ_STACK]; float stk_t[TGQ_STACK];
};

template <bool ANY>
TGQ_HD void ray_begin(RayState<ANY> & s, BvhView const & b, v3 o, v3 d, float max_time)
{
  s.o = o; s.d = d; s.r.o = o; s.r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  s.best_t = max_time; s.best_idx = -1; s.best_k = -1;
  s.sp = b.n_nodes > 0 ? 0 : -1; s.cur = 0;
}

template <bool ANY>
TGQ_HD bool ray_run(RayState<ANY> & s, BvhView const & b, int budget)
{
  if (s.sp < 0) return true;
  int cur = s.cur, sp = s.sp;
  for (;;)
  {
    while (cur >= 0)
    {
      if (--budget < 0) { s.cur = cur; s.sp = sp; return false; }
      NodeRec nd = load_node(b, cur);
      v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
      float t0 = slab_entry(s.r, lo0, hi0), t1 = slab_entry(s.r, lo1, hi1);
      int c0 = f2i(nd.q3.x), c1 = f2i(nd.q3.y);
      bool go0 = t0 >= 0.0f && (s.best_t < 0.0f || t0 <= s.best_t);
      bool go1 = t1 >= 0.0f && (s.best_t < 0.0f || t1 <= s.best_t);
      if (go0 && go1)
      {
        bool swap = t1 < t0;
        s.stk_node[sp] = swap ? c0 : c1; s.stk_t[sp] = swap ? t0 : t1; ++sp;
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
        float tm = ray_tri_time(s.o, s.d, v0, sub3(mk3(bq.x, bq.y, bq.z), v0), sub3(mk3(c.x, c.y, c.z), v0));
        int idx = f2i(a.w);
        if (tm >= 0.0f && (s.best_idx < 0 ? (s.best_t < 0.0f || tm < s.best_t) : (tm < s.best_t || (tm == s.best_t && idx < s.best_idx))))
        { s.best_t = tm; s.best_idx = idx; s.best_k = k; if (ANY) { s.sp = -1; return true; } }
      }
    }
  pop:
    for (;;)
    {
      if (sp == 0) { s.sp = -1; return true; }
      --sp;
      if (s.best_t < 0.0f || s.stk_t[sp] <= s.best_t) { cur = s.stk_node[sp]; break; }
    }
  }
}

template <bool ANY>
TGQ_HD void ray_query(BvhView const & b, v3 o, v3 d, float & best_t, int & best_idx, int & best_k)
{
  RayState<ANY> s;
  ray_begin(s, b, o, d, best_t);
  ray_run(s, b, 0x7fffffff);
  best_t = s.best_t; best_idx = s.best_idx; best_k = s.best_k;
}

}  // namespace tgq
