#include <cstdlib>
// emu.cpp -- TEST-ONLY host build of the engine's __host__ __device__ headers (thea_b200/csrc/tgq_*.cuh).
//
// There is no GPU in the development container, so the traversal / hierarchy / selection code is written as
// host+device inline functions and this file drives the *same* functions from plain C++ loops, letting the CPU test
// suite (-m "not gpu") check the pruning logic, tie rules and the libstdc++ nth_element restatement against the oracle
// before any GPU time is spent.  It is NOT a fallback: nothing under thea_b200/ links or loads it, and libtgq.so has
// no code path that reaches it.
static thread_local long long g_cnt_node = 0, g_cnt_prim = 0, g_cnt_append = 0, g_cnt_compact = 0;
#define TGQ_COUNT_NODE() (++g_cnt_node)
#define TGQ_COUNT_PRIM() (++g_cnt_prim)
#define TGQ_COUNT_APPEND() (++g_cnt_append)
#define TGQ_COUNT_COMPACT() (++g_cnt_compact)
#include "../../thea_b200/csrc/tgq_traverse.cuh"
#include "../../thea_b200/csrc/tgq_select.cuh"

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

using namespace tgq;

namespace {

struct Emu
{
  bool tri = false, has_n = false;
  int n = 0, leaf = 4; float pad = 0; float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};     // exact element bounds (tgq_handle::lo / hi)
  std::vector<TriRec> tris; std::vector<f4> pts, nrm; std::vector<int> lab; std::vector<uint32_t> perm_keep;
  std::vector<NodeRec> nodes;
  std::vector<ObbNode> obbs;
  std::vector<WideNode> wide;   // 4-wide nodes (k_emit_wide)
  int use_wide = 0;             // the product default is the binary traversal; emu_set_wide(1) selects the 4-wide one
  BvhView view() const
  {
    BvhView v; v.nodes = nodes.data(); v.obbs = tri ? obbs.data() : nullptr; v.tris = tri ? tris.data() : nullptr; v.pts = tri ? nullptr : pts.data(); v.lab = lab.empty() ? nullptr : lab.data();
    v.nrm = has_n ? nrm.data() : nullptr; v.n_nodes = (int)nodes.size(); v.n_prims = n;
    v.wide = tri ? wide.data() : nullptr; v.n_wide = tri ? (int)wide.size() : 0; return v;
  }
};

v3 ld(float const * p, int64_t i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

void build(Emu & E, float const * src, float const * n3)
{
  int n = E.n;
  E.nodes.clear();
  if (n == 0) return;
  // scene bounds (k_scene_bounds) and pad (scene_decode)
  v3 lo = mk3(FLT_MAX, FLT_MAX, FLT_MAX), hi = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  int64_t np = E.tri ? 3 * (int64_t)n : n;
  for (int64_t i = 0; i < np; ++i) { v3 p = ld(src, i); lo = min3(lo, p); hi = max3(hi, p); }
  v3 e = sub3(hi, lo);
  v3 inv = mk3(e.x > 0 ? 1.0f / e.x : 0.0f, e.y > 0 ? 1.0f / e.y : 0.0f, e.z > 0 ? 1.0f / e.z : 0.0f);
  float m = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
  float pad = m * 1.9073486328125e-06f + FLT_MIN;
  E.pad = pad;
  E.lo[0] = lo.x; E.lo[1] = lo.y; E.lo[2] = lo.z; E.hi[0] = hi.x; E.hi[1] = hi.y; E.hi[2] = hi.z;
  // keys + stable sort (k_morton_*, radix sort)
  std::vector<uint64_t> keys((size_t)n); std::vector<uint32_t> perm((size_t)n);
  for (int i = 0; i < n; ++i)
  {
    v3 c = E.tri ? tri_centroid(ld(src, 3 * (int64_t)i), ld(src, 3 * (int64_t)i + 1), ld(src, 3 * (int64_t)i + 2)) : ld(src, i);
    keys[(size_t)i] = morton63(c, lo, inv);
  }
  std::iota(perm.begin(), perm.end(), 0u);
  std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  std::vector<uint64_t> skeys((size_t)n);
  for (int s = 0; s < n; ++s) skeys[(size_t)s] = keys[perm[(size_t)s]];
  E.perm_keep = perm;
  // records + element boxes (k_gather_*)
  std::vector<v3> plo((size_t)n), phi((size_t)n);
  if (E.tri) E.tris.resize((size_t)n); else { E.pts.resize((size_t)n + 4); if (E.has_n) E.nrm.resize((size_t)n + 4); }   // + one fetch of slack (a leaf fetch always reads 4 points)
  for (int s = 0; s < n; ++s)
  {
    int64_t i = perm[(size_t)s];
    if (E.tri)
    {
      v3 v0 = ld(src, 3 * i), v1 = ld(src, 3 * i + 1), v2 = ld(src, 3 * i + 2);
      TriPlane pl = tri_plane(v0, v1, v2);
      TriRec r;
      r.q0 = f4{v0.x, v0.y, v0.z, i2f((int)i)}; r.q1 = f4{v1.x, v1.y, v1.z, i2f(pl.axis)};
      r.q2 = f4{v2.x, v2.y, v2.z, pl.d}; r.q3 = f4{pl.n.x, pl.n.y, pl.n.z, tri_screen_radius(v0, v1, v2, pad)};
      r.e0 = tri_edge_frame(v0, v1); r.e1 = tri_edge_frame(v1, v2); r.e2 = tri_edge_frame(v2, v0); r.x = tri_area_frame(v0, v1, v2, pl.axis);
      E.tris[(size_t)s] = r;
      plo[(size_t)s] = min3(v0, min3(v1, v2)); phi[(size_t)s] = max3(v0, max3(v1, v2));
    }
    else
    {
      v3 p = ld(src, i);
      E.pts[(size_t)s] = f4{p.x, p.y, p.z, i2f((int)i)};
      if (E.has_n) { v3 q = ld(n3, i); E.nrm[(size_t)s] = f4{q.x, q.y, q.z, 0.0f}; }
      plo[(size_t)s] = phi[(size_t)s] = p;
    }
  }
  if (n == 1)
  {
    if (E.tri)
    {
      ObbNode r; TriRec const & t = E.tris[0];
      v3 v0 = mk3(t.q0.x, t.q0.y, t.q0.z), v1 = mk3(t.q1.x, t.q1.y, t.q1.z), v2 = mk3(t.q2.x, t.q2.y, t.q2.z);
      ObbAcc a; obb_acc_begin(a, cross3(sub3(v1, v0), sub3(v2, v0)));
      obb_acc_point(a, v0); obb_acc_point(a, v1); obb_acc_point(a, v2);
#ifdef TGQ_NODE64
      f4 half0[2], half1[2];
      obb_acc_store(a, pad, ~((0 << TGQ_LEAF_BITS) | 1), half0); obb_store_empty(~0, half1);
      obb_pack_node64(a, half0, half1, r.q);
#else
      obb_acc_store(a, pad, ~((0 << TGQ_LEAF_BITS) | 1), r.q); obb_store_empty(~0, r.q + TGQ_OBB_F4);
#endif
      E.obbs.assign(1, r);
      WideNode w; f4 c4[4];
      obb_acc_store_wide(a, pad, ~((0 << TGQ_LEAF_BITS) | 1), c4);
      w.q[0] = c4[0]; w.q[1] = c4[1]; w.q[8] = c4[2]; w.q[9] = c4[3];
      for (int slot = 1; slot < 4; ++slot) { wide_store_empty(c4); w.q[2 * slot] = c4[0]; w.q[2 * slot + 1] = c4[1]; w.q[8 + 2 * slot] = c4[2]; w.q[8 + 2 * slot + 1] = c4[3]; }
      E.wide.assign(1, w);
    }
    E.nodes.push_back(make_node(mk3(plo[0].x - pad, plo[0].y - pad, plo[0].z - pad), mk3(phi[0].x + pad, phi[0].y + pad, phi[0].z + pad),
                                mk3(FLT_MAX, FLT_MAX, FLT_MAX), mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX), ~((0 << TGQ_LEAF_BITS) | 1), ~0, 0, 0));
    return;
  }
  // hierarchy (k_hierarchy), refit (k_refit, here by processing nodes in decreasing range order), live + scan, emit
  int ni = n - 1;
  std::vector<int> first((size_t)ni), last((size_t)ni), split((size_t)ni);
  for (int i = 0; i < ni; ++i) lbvh_range(skeys.data(), n, i, first[(size_t)i], last[(size_t)i], split[(size_t)i]);
  std::vector<int> order((size_t)ni); std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return (last[(size_t)a] - first[(size_t)a]) < (last[(size_t)b] - first[(size_t)b]); });
  std::vector<v3> nlo((size_t)ni), nhi((size_t)ni);
  for (int i : order)
  {
    int f = first[(size_t)i], l = last[(size_t)i], sp = split[(size_t)i];
    v3 alo = f == sp ? plo[(size_t)sp] : nlo[(size_t)sp], ahi = f == sp ? phi[(size_t)sp] : nhi[(size_t)sp];
    v3 blo = sp + 1 == l ? plo[(size_t)l] : nlo[(size_t)sp + 1], bhi = sp + 1 == l ? phi[(size_t)l] : nhi[(size_t)sp + 1];
    nlo[(size_t)i] = min3(alo, blo); nhi[(size_t)i] = max3(ahi, bhi);
  }
  std::vector<int> live((size_t)ni), newidx((size_t)ni);
  int cnt = 0;
  for (int i = 0; i < ni; ++i) { live[(size_t)i] = (i == 0 || (last[(size_t)i] - first[(size_t)i] + 1) > E.leaf) ? 1 : 0; newidx[(size_t)i] = cnt; cnt += live[(size_t)i]; }
  E.nodes.resize((size_t)cnt);
  if (E.tri) E.obbs.resize((size_t)cnt);
  for (int i = 0; i < ni; ++i)
  {
    if (!live[(size_t)i]) continue;
    int f = first[(size_t)i], l = last[(size_t)i], sp = split[(size_t)i];
    int child[2]; v3 clo[2], chi[2];
    for (int c = 0; c < 2; ++c)
    {
      int cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l, k = cl - cf + 1, inner = c == 0 ? sp : sp + 1;
      v3 blo = k == 1 ? plo[(size_t)cf] : nlo[(size_t)inner], bhi = k == 1 ? phi[(size_t)cf] : nhi[(size_t)inner];
      child[c] = (k <= E.leaf) ? ~((cf << TGQ_LEAF_BITS) | k) : newidx[(size_t)inner];
      clo[c] = mk3(blo.x - pad, blo.y - pad, blo.z - pad); chi[c] = mk3(bhi.x + pad, bhi.y + pad, bhi.z + pad);
    }
    E.nodes[(size_t)newidx[(size_t)i]] = make_node(clo[0], chi[0], clo[1], chi[1], child[0], child[1], 0, 0);
    if (E.tri)
    {
      // oriented boxes through the same helpers as k_obb_small / k_obb_big / k_emit (normal sum taken in plain order here)
      ObbNode r;
#ifdef TGQ_NODE64
      f4 half[2][2];
#endif
      for (int c = 0; c < 2; ++c)
      {
        int cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l;
        v3 sum = mk3(0, 0, 0);
        // both children in the frame of the NODE's normal sum: always with 64-byte nodes, else only for visit-count studies
        // (EMU_SHARED_FRAME=1)
#ifdef TGQ_NODE64
        static bool const shared_frame = true;
#else
        static bool const shared_frame = getenv("EMU_SHARED_FRAME") != nullptr;
#endif
        for (int k = shared_frame ? f : cf; k <= (shared_frame ? l : cl); ++k)
        {
          TriRec const & t = E.tris[(size_t)k];
          v3 v0 = mk3(t.q0.x, t.q0.y, t.q0.z);
          sum = add3(sum, cross3(sub3(mk3(t.q1.x, t.q1.y, t.q1.z), v0), sub3(mk3(t.q2.x, t.q2.y, t.q2.z), v0)));
        }
        ObbAcc a; obb_acc_begin(a, sum);
        for (int k = cf; k <= cl; ++k)
        {
          TriRec const & t = E.tris[(size_t)k];
          obb_acc_point(a, mk3(t.q0.x, t.q0.y, t.q0.z)); obb_acc_point(a, mk3(t.q1.x, t.q1.y, t.q1.z)); obb_acc_point(a, mk3(t.q2.x, t.q2.y, t.q2.z));
        }
#ifdef TGQ_NODE64
        obb_acc_store(a, pad, child[c], half[c]);
        if (c == 1) obb_pack_node64(a, half[0], half[1], r.q);          // a: carries the node's frame words (same normal sum for both children)
#else
        obb_acc_store(a, pad, child[c], r.q + TGQ_OBB_F4 * c);
#endif
      }
      E.obbs[(size_t)newidx[(size_t)i]] = r;
    }
  }
  if (E.tri)
  {
    // 4-wide nodes (k_wide_flags / k_emit_wide): surviving nodes at even depth; children = binary grandchildren or leaf children.
    // Emitted depth first from the root (the device numbers them by a scan: the numbering is irrelevant to the answers).
    auto child_record = [&](int cf, int cl, int link, f4 * c4) {
      v3 sum = mk3(0, 0, 0);
      for (int k = cf; k <= cl; ++k)
      {
        TriRec const & t = E.tris[(size_t)k];
        v3 v0 = mk3(t.q0.x, t.q0.y, t.q0.z);
        sum = add3(sum, cross3(sub3(mk3(t.q1.x, t.q1.y, t.q1.z), v0), sub3(mk3(t.q2.x, t.q2.y, t.q2.z), v0)));
      }
      ObbAcc a; obb_acc_begin(a, sum);
      for (int k = cf; k <= cl; ++k)
      {
        TriRec const & t = E.tris[(size_t)k];
        obb_acc_point(a, mk3(t.q0.x, t.q0.y, t.q0.z)); obb_acc_point(a, mk3(t.q1.x, t.q1.y, t.q1.z)); obb_acc_point(a, mk3(t.q2.x, t.q2.y, t.q2.z));
      }
      obb_acc_store_wide(a, pad, link, c4);
    };
    auto put = [](WideNode & r, int slot, f4 const * c4) { r.q[2 * slot] = c4[0]; r.q[2 * slot + 1] = c4[1]; r.q[8 + 2 * slot] = c4[2]; r.q[8 + 2 * slot + 1] = c4[3]; };
    E.wide.clear(); E.wide.push_back(WideNode());
    std::vector<std::pair<int, int>> todo;     // (radix node, wide index)
    todo.push_back(std::make_pair(0, 0));
    while (!todo.empty())
    {
      int i = todo.back().first, wi = todo.back().second; todo.pop_back();
      WideNode r; int slot = 0; f4 c4[4];
      int f = first[(size_t)i], l = last[(size_t)i], sp = split[(size_t)i];
      auto add_child = [&](int cf, int cl, int inner) {
        int k = cl - cf + 1, link;
        if (k <= E.leaf) link = ~((cf << TGQ_LEAF_BITS) | k);
        else { link = (int)E.wide.size(); E.wide.push_back(WideNode()); todo.push_back(std::make_pair(inner, link)); }
        child_record(cf, cl, link, c4); put(r, slot++, c4);
      };
      for (int c = 0; c < 2; ++c)
      {
        int cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l, inner = c == 0 ? sp : sp + 1;
        if (cl - cf + 1 <= E.leaf) { add_child(cf, cl, inner); continue; }
        int gsp = split[(size_t)inner];
        add_child(cf, gsp, gsp); add_child(gsp + 1, cl, gsp + 1);
      }
      for (; slot < 4; ++slot) { wide_store_empty(c4); put(r, slot, c4); }
      E.wide[(size_t)wi] = r;
    }
  }
}

}  // namespace

extern "C" {

void * emu_build(float const * src, int64_t n, int is_tri, float const * n3, int leaf)
{
  Emu * E = new Emu; E->tri = is_tri != 0; E->has_n = (!is_tri && n3); E->n = (int)n; E->leaf = leaf;
  build(*E, src, n3);
  return E;
}
void emu_destroy(void * h) { delete (Emu *)h; }
void emu_set_wide(void * h, int on) { ((Emu *)h)->use_wide = on; }        // closest point on triangles: 4-wide nodes (default) or binary nodes
int64_t emu_num_wide(void * h) { return (int64_t)((Emu *)h)->wide.size(); }
int64_t emu_num_nodes(void * h) { return (int64_t)((Emu *)h)->nodes.size(); }

void emu_closest(void * h, float const * q3, int64_t Q, double dist_bound, int exclude_self, int32_t * idx, float * d2, float * cp3)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  float bound_ub; BoundGate gate; closest_bound_gate(E.lo, E.hi, dist_bound, bound_ub, gate);      // as tgq.cu: closest_common does
  // every query is seeded with the Morton position of the previous query's answer, as a lane of the kernel does
  // (exclude_self < 0: unseeded)
  int seed = -1; bool seeded = exclude_self >= 0; if (!seeded) exclude_self = -exclude_self - 1;
  for (int64_t i = 0; i < Q; ++i)
  {
    float best = gated_ub(gate, ld(q3, i), bound_ub); int bi = -1; v3 cp = mk3(0, 0, 0);
    if (E.tri && E.use_wide) closest_query_wide(v, E.pad, ld(q3, i), exclude_self ? (int)i : -1, best, bi, cp, seeded ? &seed : nullptr);
    else if (E.tri) closest_query<true>(v, E.pad, ld(q3, i), exclude_self ? (int)i : -1, best, bi, cp, seeded ? &seed : nullptr);
    else closest_query<false>(v, 0.0f, ld(q3, i), exclude_self ? (int)i : -1, best, bi, cp, seeded ? &seed : nullptr);
    idx[i] = bi; d2[i] = bi >= 0 ? best : -1.0f; cp3[3 * i] = cp.x; cp3[3 * i + 1] = cp.y; cp3[3 * i + 2] = cp.z;
  }
}

// Closest element to a line-segment query object, as tgq.cu: tgq_closest_segment drives it (gate on the segment's bounding box).
void emu_closest_segment(void * h, float const * a3, float const * b3, int64_t Q, double dist_bound, int32_t const * excl, int n_excl,
                         int32_t * idx, float * d2, float * ep3, float * sp3)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  float bound_ub; BoundGate gate; closest_bound_gate(E.lo, E.hi, dist_bound, bound_ub, gate);
  for (int64_t i = 0; i < Q; ++i)
  {
    v3 a = ld(a3, i), b = sub3(ld(b3, i), a), far = add3(a, b);        // b: the direction, as LineSegment3(end0, end1) stores it
    float best = gated_ub_box(gate, min3(a, far), max3(a, far), bound_ub); int bi = -1; v3 ep = mk3(0, 0, 0), sp = mk3(0, 0, 0);
    if (E.tri) seg_query<true>(v, E.pad, a, b, excl ? excl + i * n_excl : nullptr, excl ? n_excl : 0, best, bi, ep, sp);
    else seg_query<false>(v, E.pad, a, b, excl ? excl + i * n_excl : nullptr, excl ? n_excl : 0, best, bi, ep, sp);
    idx[i] = bi; d2[i] = bi >= 0 ? best : -1.0f;
    ep3[3 * i] = ep.x; ep3[3 * i + 1] = ep.y; ep3[3 * i + 2] = ep.z; sp3[3 * i] = sp.x; sp3[3 * i + 1] = sp.y; sp3[3 * i + 2] = sp.z;
  }
}

// Closest element to a triangle query object, as tgq.cu: tgq_closest_triangle drives it.
void emu_closest_triangle(void * h, float const * q9, int64_t Q, double dist_bound, int32_t * idx, float * d2, float * ep3, float * qp3)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  float bound_ub; BoundGate gate; closest_bound_gate(E.lo, E.hi, dist_bound, bound_ub, gate);
  // FLT_EPSILON over the smallest positive |N| of the elements (k_min_normal in the engine)
  float elem_tau = 0.0f;
  if (E.tri)
  {
    float mn = INFINITY;
    for (int i = 0; i < E.n; ++i)
    {
      TriRec const & t = E.tris[(size_t)i];
      v3 a = mk3(t.q0.x, t.q0.y, t.q0.z), N = cross3(sub3(mk3(t.q1.x, t.q1.y, t.q1.z), a), sub3(mk3(t.q2.x, t.q2.y, t.q2.z), a));
      float nn = sqrtf(fdot3(N, N)); if (nn > 0.0f && nn < mn) mn = nn;
    }
    elem_tau = mn < INFINITY ? FLT_EPSILON / mn : 0.0f;
  }
  for (int64_t i = 0; i < Q; ++i)
  {
    v3 v0 = ld(q9, 3 * i), v1 = ld(q9, 3 * i + 1), v2 = ld(q9, 3 * i + 2);
    float best = gated_ub_box(gate, min3(v0, min3(v1, v2)), max3(v0, max3(v1, v2)), bound_ub); int bi = -1; v3 ep = mk3(0, 0, 0), qp = mk3(0, 0, 0);
    if (E.tri) triq_query<true>(v, E.pad, elem_tau, v0, v1, v2, best, bi, ep, qp);
    else triq_query<false>(v, E.pad, 0.0f, v0, v1, v2, best, bi, ep, qp);
    idx[i] = bi; d2[i] = bi >= 0 ? best : -1.0f;
    ep3[3 * i] = ep.x; ep3[3 * i + 1] = ep.y; ep3[3 * i + 2] = ep.z; qp3[3 * i] = qp.x; qp3[3 * i + 1] = qp.y; qp3[3 * i + 2] = qp.z;
  }
}

// Per-pair triangle-triangle distance of the device headers (tgq_math.cuh: tri_tri_sqdist), for the pin against the reference.
void emu_tri_tri_batch(float const * va9, float const * vb9, int64_t n, float * d2, float * pa3, float * pb3)
{
  for (int64_t i = 0; i < n; ++i)
  {
    TriQ a = make_triq(ld(va9, 3 * i), ld(va9, 3 * i + 1), ld(va9, 3 * i + 2)), b = make_triq(ld(vb9, 3 * i), ld(vb9, 3 * i + 1), ld(vb9, 3 * i + 2));
    v3 pa, pb; d2[i] = tri_tri_sqdist(a, b, pa, pb);
    pa3[3 * i] = pa.x; pa3[3 * i + 1] = pa.y; pa3[3 * i + 2] = pa.z; pb3[3 * i] = pb.x; pb3[3 * i + 1] = pb.y; pb3[3 * i + 2] = pb.z;
  }
}

// Per-query visit counts (development aid for studying pruning quality).
void emu_closest_counts(void * h, float const * q3, int64_t Q, float const * ub, int32_t * nodes, int32_t * prims)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  for (int64_t i = 0; i < Q; ++i)
  {
    long long n0 = g_cnt_node, p0 = g_cnt_prim;
    float best = ub ? ub[i] : INFINITY; int bi = -1; v3 cp = mk3(0, 0, 0);
    if (E.tri && E.use_wide) closest_query_wide(v, E.pad, ld(q3, i), -1, best, bi, cp);
    else if (E.tri) closest_query<true>(v, E.pad, ld(q3, i), -1, best, bi, cp);
    else closest_query<false>(v, 0.0f, ld(q3, i), -1, best, bi, cp);
    nodes[i] = (int32_t)(g_cnt_node - n0); prims[i] = (int32_t)(g_cnt_prim - p0);
  }
}

// Visit counts when query i is seeded with the answer of query i - stride of the same (Morton-ordered) array: stride 1 = the
// neighbouring query, stride 32 = what a lane of a 32-lane warp sees when queries are dealt to idle lanes one at a time
// (tools/seed_stride_stats.py).
void emu_closest_counts_stride(void * h, float const * q3, int64_t Q, int stride, int32_t * nodes, int32_t * prims)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  std::vector<int> ans((size_t)Q, -1);
  for (int64_t i = 0; i < Q; ++i)
  {
    long long n0 = g_cnt_node, p0 = g_cnt_prim;
    float best = INFINITY; int bi = -1; v3 cp = mk3(0, 0, 0);
    int seed = (stride > 0 && i >= stride) ? ans[(size_t)(i - stride)] : -1;
    closest_query<true>(v, E.pad, ld(q3, i), -1, best, bi, cp, &seed);
    ans[(size_t)i] = seed;
    nodes[i] = (int32_t)(g_cnt_node - n0); prims[i] = (int32_t)(g_cnt_prim - p0);
  }
}

// k-NN / sample graph: 1 (default, as the engine) = the candidate-buffer traversals (CandBuf; serial compaction stands in for the
// warp's), 0 = the sorted-set traversals.
static int g_buffered = 1;
void emu_set_buffered(int on) { g_buffered = on; }
static void buf_out(CandBuf const & buf, int k, int32_t * idx, float * key)
{ for (int j = 0; j < k; ++j) { idx[j] = j < buf.n ? cand_idx(buf.at(j)) : -1; key[j] = j < buf.n ? cand_key(buf.at(j)) : -1.0f; } }

void emu_knn(void * h, float const * q3, int64_t Q, int k, float bound_ub, int exclude_self, int32_t * idx, float * d2)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  if (g_buffered)
  {
    std::vector<unsigned long long> row((size_t)cand_cap(k));
    for (int64_t i = 0; i < Q; ++i)
    {
      CandBuf buf; buf.init(row.data(), (int)(i % 32), k);
      knn_query_buf(v, ld(q3, i), exclude_self ? (int)i : -1, bound_ub, buf);
      buf_out(buf, k, idx + i * k, d2 + i * k);
    }
    return;
  }
  std::vector<float> keys((size_t)k); std::vector<int> ids((size_t)k);
  for (int64_t i = 0; i < Q; ++i)
  {
    SortedSet set; set.init(keys.data(), ids.data(), 1, k);
    knn_query(v, ld(q3, i), exclude_self ? (int)i : -1, bound_ub, set);
    for (int j = 0; j < k; ++j) { idx[i * k + j] = j < set.n ? ids[(size_t)j] : -1; d2[i * k + j] = j < set.n ? keys[(size_t)j] : -1.0f; }
  }
}

// Filtered queries on point structures (tgq_set_point_labels / tgq_closest_filtered / tgq_knn_filtered).
void emu_set_labels(void * h, int32_t const * labels)
{
  Emu & E = *(Emu *)h; E.lab.clear();
  if (!labels) return;
  E.lab.resize((size_t)E.n + 4);
  for (int s = 0; s < E.n; ++s) E.lab[(size_t)s] = labels[E.perm_keep[(size_t)s]];
}

static ElemFilterD make_filter(float const * qn3, int32_t const * qlabel, int64_t i, float min_dot)
{
  ElemFilterD f = no_filter(); f.mode = (qn3 ? 1 : 0) | (qlabel ? 2 : 0); f.min_dot = min_dot;
  if (qn3) f.n = ld(qn3, i);
  if (qlabel) f.label = qlabel[i];
  return f;
}

void emu_closest_filtered(void * h, float const * q3, float const * qn3, int32_t const * qlabel, int64_t Q, float min_dot, double dist_bound,
                          int32_t * idx, float * d2, float * cp3)
{
  Emu & E = *(Emu *)h; BvhView v = E.view(); int seed = -1;
  float bound_ub; BoundGate gate; closest_bound_gate(E.lo, E.hi, dist_bound, bound_ub, gate);
  for (int64_t i = 0; i < Q; ++i)
  {
    float best = gated_ub(gate, ld(q3, i), bound_ub); int bi = -1; v3 cp = mk3(0, 0, 0);
    ElemFilterD f = make_filter(qn3, qlabel, i, min_dot);
    closest_query<false>(v, 0.0f, ld(q3, i), -1, best, bi, cp, &seed, &f);
    idx[i] = bi; d2[i] = bi >= 0 ? best : -1.0f; cp3[3 * i] = cp.x; cp3[3 * i + 1] = cp.y; cp3[3 * i + 2] = cp.z;
  }
}

void emu_knn_filtered(void * h, float const * q3, float const * qn3, int32_t const * qlabel, int64_t Q, int k, float min_dot, float bound_ub,
                      int32_t * idx, float * d2)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  if (g_buffered)
  {
    std::vector<unsigned long long> row((size_t)cand_cap(k));
    for (int64_t i = 0; i < Q; ++i)
    {
      CandBuf buf; buf.init(row.data(), (int)(i % 32), k);
      ElemFilterD f = make_filter(qn3, qlabel, i, min_dot);
      knn_query_buf(v, ld(q3, i), -1, bound_ub, buf, &f);
      buf_out(buf, k, idx + i * k, d2 + i * k);
    }
    return;
  }
  std::vector<float> keys((size_t)k); std::vector<int> ids((size_t)k);
  for (int64_t i = 0; i < Q; ++i)
  {
    SortedSet set; set.init(keys.data(), ids.data(), 1, k);
    ElemFilterD f = make_filter(qn3, qlabel, i, min_dot);
    knn_query(v, ld(q3, i), -1, bound_ub, set, &f);
    for (int j = 0; j < k; ++j) { idx[i * k + j] = j < set.n ? ids[(size_t)j] : -1; d2[i * k + j] = j < set.n ? keys[(size_t)j] : -1.0f; }
  }
}

// Unbounded range query (tgq_range_query): returns the number of elements in the range, writes up to cap indices (traversal order).
int64_t emu_range_all(void * h, int box, float const * a3, float const * b3, float radius, int32_t * idx, int64_t cap)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  std::vector<int> out((size_t)E.n + 4);
  int n = box ? range_all_query<true, true>(v, ld(a3, 0), ld(b3, 0), out.data())
              : (radius >= 0.0f ? range_all_query<false, true>(v, ld(a3, 0), mk3(radius * radius, radius, 0.0f), out.data()) : 0);
  for (int j = 0; j < n && j < cap; ++j) idx[j] = out[(size_t)j];
  return n;
}

void emu_ball_range(void * h, float const * c3, float const * radius, int64_t Q, int max_out, int exclude_self, int32_t * idx, int32_t * count)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  std::vector<float> keys((size_t)max_out + 1); std::vector<int> ids((size_t)max_out + 1);
  for (int64_t i = 0; i < Q; ++i)
  {
    SortedSet set; set.init(keys.data(), ids.data(), 1, max_out); int cnt = 0;
    range_query(v, ld(c3, i), radius[i], exclude_self ? (int)i : -1, set, cnt);
    for (int j = 0; j < max_out; ++j) idx[i * max_out + j] = j < set.n ? ids[(size_t)j] : -1;
    count[i] = cnt;
  }
}

void emu_ball_graph_surface(void * h, void * hs, int k, float init_radius, int rounds, float min_dot, int min_degree, int32_t * idx, float * sep,
                            int32_t * count)
{
  Emu & E = *(Emu *)h; BvhView v = E.view(); BvhView sv = ((Emu *)hs)->view();
  std::vector<float> keys((size_t)k); std::vector<int> ids((size_t)k);
  for (int s = 0; s < E.n; ++s)
  {
    f4 e = E.pts[(size_t)s]; int64_t i = f2i(e.w);
    v3 nc = mk3(0, 0, 0); if (E.has_n) nc = mk3(E.nrm[(size_t)s].x, E.nrm[(size_t)s].y, E.nrm[(size_t)s].z);
    SortedSet set; set.init(keys.data(), ids.data(), 1, k);
    ball_graph_surface_query(v, sv, mk3(e.x, e.y, e.z), nc, E.has_n, min_dot, (int)i, init_radius, rounds, min_degree, set);
    for (int j = 0; j < k; ++j) { idx[i * k + j] = j < set.n ? ids[(size_t)j] : -1; sep[i * k + j] = j < set.n ? keys[(size_t)j] : -1.0f; }
    if (count) count[i] = set.n;
  }
}

void emu_ball_hist(void * h, float const * c3, float const * radius, int64_t Q, int nbins, uint32_t * hist)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  for (int64_t i = 0; i < Q; ++i) hist_query(v, ld(c3, i), radius[i], nbins, hist + (size_t)i * nbins);
}

#ifdef TGQ_NODE64
// The decoded frame of a 64-byte node for a given normal sum (what both the build and the query use): three axes, row major.
void emu_node_frame64(float const * nsum3, int64_t n, float * axes9)
{
  for (int64_t i = 0; i < n; ++i)
  {
    ObbAcc a = obb_node_frame64(ld(nsum3, i));
    v3 d0, d1, d2; obb_decode_frame64(a.wn, a.wr, a.wx, d0, d1, d2);          // decoding the packed words again must give the same frame
    float * o = axes9 + 9 * i;
    o[0] = d0.x; o[1] = d0.y; o[2] = d0.z; o[3] = d1.x; o[4] = d1.y; o[5] = d1.z; o[6] = d2.x; o[7] = d2.y; o[8] = d2.z;
    if (d0.x != a.f.a0.x || d1.y != a.f.a1.y || d2.z != a.f.a2.z) o[0] = NAN;
  }
}
#endif

// SpinImage (tgq_spin_image): the engine's per-pair arithmetic (spin_bins, tgq_math.cuh) over all samples of the structure.
void emu_spin_image(void * h, float const * q3, float const * qn3, int64_t Q, int nr, int nh, float max_extent, uint32_t * image)
{
  Emu & E = *(Emu *)h;
  for (int64_t i = 0; i < Q; ++i)
  {
    v3 pos = ld(q3, i), axis = stable_normalized3(ld(qn3, i));
    uint32_t * img = image + (size_t)i * nr * nh;
    for (int j = 0; j < nr * nh; ++j) img[j] = 0;
    for (int s = 0; s < E.n; ++s)
    {
      f4 e = E.pts[(size_t)s]; int rb, hb;
      spin_bins(mk3(e.x, e.y, e.z), pos, axis, nr, nh, max_extent, max_extent, rb, hb);
      img[rb * nh + hb] += 1;
    }
  }
}

void emu_ball_stats(void * h, float const * c3, float const * qn3, float const * radius, int64_t Q, double * out16)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  for (int64_t i = 0; i < Q; ++i)
    stats_query(v, ld(c3, i), radius[i], qn3 ? ld(qn3, i) : mk3(0, 0, 0), qn3 != nullptr, out16 + 16 * i);
}

void emu_ball_graph(void * h, int k, float init_radius, int rounds, float min_dot, int min_degree, int32_t * idx, float * sep, int32_t * count)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  if (g_buffered)
  {
    std::vector<unsigned long long> row((size_t)cand_cap(k));
    for (int s = 0; s < E.n; ++s)
    {
      f4 e = E.pts[(size_t)s]; int64_t i = f2i(e.w);
      v3 nc = mk3(0, 0, 0); if (E.has_n) nc = mk3(E.nrm[(size_t)s].x, E.nrm[(size_t)s].y, E.nrm[(size_t)s].z);
      CandBuf buf; buf.init(row.data(), s % 32, k);
      ball_graph_query_buf(v, mk3(e.x, e.y, e.z), nc, E.has_n, min_dot, (int)i, init_radius, rounds, min_degree, buf);
      buf_out(buf, k, idx + i * k, sep + i * k);
      if (count) count[i] = buf.n;
    }
    return;
  }
  std::vector<float> keys((size_t)k); std::vector<int> ids((size_t)k);
  for (int s = 0; s < E.n; ++s)
  {
    f4 e = E.pts[(size_t)s]; int64_t i = f2i(e.w);
    v3 nc = mk3(0, 0, 0); if (E.has_n) nc = mk3(E.nrm[(size_t)s].x, E.nrm[(size_t)s].y, E.nrm[(size_t)s].z);
    SortedSet set; set.init(keys.data(), ids.data(), 1, k);
    ball_graph_query(v, mk3(e.x, e.y, e.z), nc, E.has_n, min_dot, (int)i, init_radius, rounds, min_degree, set);
    for (int j = 0; j < k; ++j) { idx[i * k + j] = j < set.n ? ids[(size_t)j] : -1; sep[i * k + j] = j < set.n ? keys[(size_t)j] : -1.0f; }
    if (count) count[i] = set.n;
  }
}

void emu_ray(void * h, float const * o3, float const * d3, float const * tmax, int64_t R, int any, float * t, int32_t * idx)
{
  Emu & E = *(Emu *)h; BvhView v = E.view();
  for (int64_t i = 0; i < R; ++i)
  {
    float best = tmax ? tmax[i] : -1.0f; int bi = -1, bk = -1;
    if (any) ray_query<true>(v, ld(o3, i), ld(d3, i), best, bi, bk); else ray_query<false>(v, ld(o3, i), ld(d3, i), best, bi, bk);
    t[i] = bi >= 0 ? best : -1.0f; idx[i] = bi;
  }
}

void emu_counters(long long * nodes, long long * prims, int reset)
{ *nodes = g_cnt_node; *prims = g_cnt_prim; if (reset) { g_cnt_node = 0; g_cnt_prim = 0; } }
void emu_set_counters(long long * appends, long long * compactions, int reset)
{ *appends = g_cnt_append; *compactions = g_cnt_compact; if (reset) { g_cnt_append = 0; g_cnt_compact = 0; } }

// libstdc++'s own std::nth_element next to the restatement, for permutation-exact comparison.
void emu_select(double * a, int n, int nth) { sdf_select(a, n, nth); }
void emu_std_nth_element(double * a, int n, int nth) { std::nth_element(a, a + nth, a + n); }
double emu_sdf_statistics(double * values, double const * weights, int nv, float scale) { return sdf_statistics(values, weights, nv, scale); }

void emu_sdf_rays(float const * p3, float const * n3, int64_t S, float scale, float * o3, float * d3)
{
  static float const cone[30][3] = {
    { 0.407497f,  0.134086f,  0.903309f}, { 0.128722f, -0.822715f,  0.553688f}, { 0.359298f,  0.686129f,  0.632560f},
    { 0.813318f, -0.153956f,  0.561081f}, {-0.707169f,  0.296987f,  0.641647f}, {-0.478783f, -0.543441f,  0.689520f},
    { 0.561083f,  0.523634f,  0.641087f}, {-0.380898f, -0.655789f,  0.651811f}, { 0.722588f,  0.472155f,  0.504912f},
    {-0.535205f,  0.421235f,  0.732200f}, { 0.577593f, -0.477896f,  0.661817f}, { 0.393652f, -0.144484f,  0.907834f},
    { 0.141508f, -0.660777f,  0.737122f}, {-0.274967f, -0.141703f,  0.950954f}, {-0.304935f,  0.474085f,  0.825989f},
    { 0.711150f,  0.301383f,  0.635164f}, {-0.696119f, -0.112341f,  0.709082f}, {-0.248327f,  0.381035f,  0.890588f},
    {-0.0772394f, 0.0462501f, 0.995939f}, { 0.438024f,  0.656463f,  0.614159f}, {-0.483384f, -0.542477f,  0.687065f},
    { 0.498569f, -0.697466f,  0.514752f}, {-0.128902f, -0.0173429f, 0.991506f}, { 0.338759f, -0.318772f,  0.885227f},
    {-0.574364f, -0.463887f,  0.674474f}, {-0.459207f,  0.260125f,  0.849390f}, {-0.628057f, -0.528928f,  0.570771f},
    { 0.232916f, -0.328273f,  0.915416f}, { 0.231667f, -0.796882f,  0.557951f}, {-0.270612f, -0.809654f,  0.520797f},
  };
  for (int64_t s = 0; s < S; ++s)
  {
    SdfFrame f = sdf_frame(ld(p3, s), ld(n3, s), scale);
    for (int k = 0; k < 30; ++k)
    {
      v3 d = sdf_dir(f, mk3(cone[k][0], cone[k][1], cone[k][2]));
      int64_t r = 30 * s + k;
      o3[3 * r] = f.origin.x; o3[3 * r + 1] = f.origin.y; o3[3 * r + 2] = f.origin.z;
      d3[3 * r] = d.x; d3[3 * r + 1] = d.y; d3[3 * r + 2] = d.z;
    }
  }
}

}  // extern "C"
