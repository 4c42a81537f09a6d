// tgq_kernels.cuh -- __global__ kernels of the engine (sm_100a).  Build: K1 bounds+Morton, K3 hierarchy, K4 refit,
// node emission; queries: K5 closest, K6 k-NN, K7 ball graph, K8 ray first/any hit (+ SDF cone rays), K9 ICP sums,
// K11 area sampling; SURVEY.md 8f additions: ball range / statistics / histograms, visibility fans, spin images, the sample-graph searches
// (tgq_graph.cuh), and the query objects of rank 4 (k_closest_seg, k_closest_tri).  Numbering follows SURVEY.md 2.2.
#pragma once
#include "tgq_traverse.cuh"
#include "tgq_select.cuh"
#include <cuda_runtime.h>

namespace tgq {

#ifndef TGQ_BLOCK
#define TGQ_BLOCK 128
#endif

// ---- ordered-int encoding so that float min/max can use integer atomics ---------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct SceneBox { int lo[3]; int hi[3]; };   // ordered-int encoded, reduced with atomicMin / atomicMax

__global__ void k_scene_init(SceneBox * sb)
{
  for (int a = 0; a < 3; ++a) { sb->lo[a] = float_to_ordered(FLT_MAX); sb->hi[a] = float_to_ordered(-FLT_MAX); }
}

// Bounds of n points laid out as n x 3 floats (triangle soups are 3T points).
__global__ void k_scene_bounds(float const * __restrict__ p, int64_t n, SceneBox * sb)
{
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    for (int a = 0; a < 3; ++a) { float v = p[3 * i + a]; lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
  for (int a = 0; a < 3; ++a)
  {
    for (int o = 16; o > 0; o >>= 1)
    { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&sb->lo[a], float_to_ordered(lo[a])); atomicMax(&sb->hi[a], float_to_ordered(hi[a])); }
  }
}

struct SceneF { v3 lo, hi, inv_ext; float pad; };
__device__ __forceinline__ SceneF scene_decode(SceneBox const * sb)
{
  SceneF s;
  s.lo = mk3(ordered_to_float(sb->lo[0]), ordered_to_float(sb->lo[1]), ordered_to_float(sb->lo[2]));
  s.hi = mk3(ordered_to_float(sb->hi[0]), ordered_to_float(sb->hi[1]), ordered_to_float(sb->hi[2]));
  v3 e = sub3(s.hi, s.lo);
  s.inv_ext = mk3(e.x > 0 ? 1.0f / e.x : 0.0f, e.y > 0 ? 1.0f / e.y : 0.0f, e.z > 0 ? 1.0f / e.z : 0.0f);
  float m = fmaxf(fmaxf(fmaxf(fabsf(s.lo.x), fabsf(s.hi.x)), fmaxf(fabsf(s.lo.y), fabsf(s.hi.y))), fmaxf(fabsf(s.lo.z), fabsf(s.hi.z)));
  s.pad = m * 1.9073486328125e-06f + FLT_MIN;   // 2^-19 of the largest coordinate: absorbs fp32 error of computed closest points
  return s;
}

__device__ __forceinline__ v3 ld3(float const * p, int64_t i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__device__ __forceinline__ void st3(float * p, int64_t i, v3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

// K1: Morton keys of the element centres (triangle centroid as Triangle3.hpp:150, BoundedTraitsN.hpp:215; point itself).
__global__ void k_morton_tris(float const * __restrict__ v9, int n, SceneBox const * sb, uint64_t * keys, uint32_t * vals)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  SceneF s = scene_decode(sb);
  v3 c = tri_centroid(ld3(v9, 3 * (int64_t)i), ld3(v9, 3 * (int64_t)i + 1), ld3(v9, 3 * (int64_t)i + 2));
  keys[i] = morton63(c, s.lo, s.inv_ext); vals[i] = (uint32_t)i;
}
__global__ void k_morton_points(float const * __restrict__ p3, int n, SceneBox const * sb, uint64_t * keys, uint32_t * vals)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  SceneF s = scene_decode(sb);
  keys[i] = morton63(ld3(p3, i), s.lo, s.inv_ext); vals[i] = (uint32_t)i;
}

// Element records in Morton order + element boxes.
__global__ void k_gather_tris(float const * __restrict__ v9, uint32_t const * __restrict__ perm, int n, SceneBox const * sb, TriRec * tris,
                              f4 * pbox_lo, f4 * pbox_hi, int * pos_of_elem, f4 * pnrm, float * area_by_elem)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= n) return;
  int64_t i = perm[s];
  pos_of_elem[i] = s;
  float pad = scene_decode(sb).pad;
  v3 v0 = ld3(v9, 3 * i), v1 = ld3(v9, 3 * i + 1), v2 = ld3(v9, 3 * i + 2);
  TriPlane pl = tri_plane(v0, v1, v2);
  area_by_elem[i] = tri_area(pl.n, v0, v1, v2);
  v3 cr = cross3(sub3(v1, v0), sub3(v2, v0));       // area-weighted normal, summed up the tree for the slab directions
  pnrm[s] = make_float4(cr.x, cr.y, cr.z, 0.0f);
  TriRec r;
  r.q0 = make_float4(v0.x, v0.y, v0.z, __int_as_float((int)i));
  r.q1 = make_float4(v1.x, v1.y, v1.z, __int_as_float(pl.axis));
  r.q2 = make_float4(v2.x, v2.y, v2.z, pl.d);
  r.q3 = make_float4(pl.n.x, pl.n.y, pl.n.z, tri_screen_radius(v0, v1, v2, pad));
  r.e0 = tri_edge_frame(v0, v1); r.e1 = tri_edge_frame(v1, v2); r.e2 = tri_edge_frame(v2, v0);
  r.x = tri_area_frame(v0, v1, v2, pl.axis);
  tris[s] = r;
  v3 lo = min3(v0, min3(v1, v2)), hi = max3(v0, max3(v1, v2));      // Triangle3.hpp:215-219
  pbox_lo[s] = make_float4(lo.x, lo.y, lo.z, 0.0f); pbox_hi[s] = make_float4(hi.x, hi.y, hi.z, 0.0f);
}
__global__ void k_gather_points(float const * __restrict__ p3, float const * __restrict__ n3, uint32_t const * __restrict__ perm,
                                int n, f4 * pts, f4 * nrm, f4 * pbox_lo, f4 * pbox_hi, int * pos_of_elem)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= n) return;
  int64_t i = perm[s];
  pos_of_elem[i] = s;
  v3 p = ld3(p3, i);
  pts[s] = make_float4(p.x, p.y, p.z, __int_as_float((int)i));
  if (n3) { v3 q = ld3(n3, i); nrm[s] = make_float4(q.x, q.y, q.z, 0.0f); }
  pbox_lo[s] = make_float4(p.x, p.y, p.z, 0.0f); pbox_hi[s] = pbox_lo[s];
}

// Labels of a point structure in Morton order (filtered queries).
__global__ void k_gather_labels(int const * __restrict__ labels, uint32_t const * __restrict__ perm, int n, int * out)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= n) return;
  out[s] = labels[perm[s]];
}

// Source elements back out of the Morton-ordered records, in element order (tgq_set_transform keeps them as the untransformed
// originals): triangles -> 9 floats, points -> 3 floats (+ normal, + label).
__global__ void k_extract_tris(TriRec const * __restrict__ tris, int n, float * out9)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x; if (k >= n) return;
  f4 const * q = (f4 const *)(tris + k);
  f4 a = q[0], b = q[1], c = q[2];
  float * o = out9 + 9 * (size_t)__float_as_int(a.w);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = c.x; o[7] = c.y; o[8] = c.z;
}
__global__ void k_extract_points(f4 const * __restrict__ pts, f4 const * __restrict__ nrm, int const * __restrict__ lab, int n, float * out3, float * outn3,
                                 int * outlab)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x; if (k >= n) return;
  f4 a = pts[k]; size_t i = (size_t)__float_as_int(a.w);
  if (out3) { out3[3 * i] = a.x; out3[3 * i + 1] = a.y; out3[3 * i + 2] = a.z; }
  if (nrm) { f4 m = nrm[k]; outn3[3 * i] = m.x; outn3[3 * i + 1] = m.y; outn3[3 * i + 2] = m.z; }
  if (lab) outlab[i] = lab[k];
}

// K3: binary radix tree over the sorted keys.  Inner node i: range [first,last], split; parents of inner nodes at
// parent[i], of leaves (sorted position s) at parent[n - 1 + s].
__global__ void k_hierarchy(uint64_t const * __restrict__ keys, int n, int * first, int * last, int * split, int * parent)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n - 1) return;
  int f, l, s; lbvh_range(keys, n, i, f, l, s);
  first[i] = f; last[i] = l; split[i] = s;
  if (f == s) parent[n - 1 + s] = i; else parent[s] = i;
  if (s + 1 == l) parent[n - 1 + s + 1] = i; else parent[s + 1] = i;
  if (i == 0) parent[0] = -1;
}

// K4: bottom-up union of boxes (BvhN.hpp:1716-1730 updateNodeBoundsRecursive); the second thread to arrive at a node
// owns it.  nbox = unpadded box of every inner node of the binary radix tree.
__global__ void k_refit(int n, int const * __restrict__ first, int const * __restrict__ last, int const * __restrict__ split,
                        int const * __restrict__ parent, f4 const * pbox_lo, f4 const * pbox_hi, f4 * nbox_lo, f4 * nbox_hi,
                        unsigned int * arrive, f4 const * pnrm, f4 * nsum)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= n) return;
  int cur = parent[n - 1 + s];
  while (cur >= 0)
  {
    __threadfence();
    if (atomicAdd(&arrive[cur], 1u) == 0u) return;
    int f = first[cur], l = last[cur], sp = split[cur];
    volatile f4 const * vlo = nbox_lo; volatile f4 const * vhi = nbox_hi;
    volatile f4 const * vns = nsum;
    f4 alo, ahi, blo, bhi, an, bn;
    if (f == sp) { alo = pbox_lo[sp]; ahi = pbox_hi[sp]; if (pnrm) an = pnrm[sp]; }
    else
    {
      alo.x = vlo[sp].x; alo.y = vlo[sp].y; alo.z = vlo[sp].z; ahi.x = vhi[sp].x; ahi.y = vhi[sp].y; ahi.z = vhi[sp].z;
      if (pnrm) { an.x = vns[sp].x; an.y = vns[sp].y; an.z = vns[sp].z; }
    }
    if (sp + 1 == l) { blo = pbox_lo[l]; bhi = pbox_hi[l]; if (pnrm) bn = pnrm[l]; }
    else
    {
      blo.x = vlo[sp + 1].x; blo.y = vlo[sp + 1].y; blo.z = vlo[sp + 1].z; bhi.x = vhi[sp + 1].x; bhi.y = vhi[sp + 1].y; bhi.z = vhi[sp + 1].z;
      if (pnrm) { bn.x = vns[sp + 1].x; bn.y = vns[sp + 1].y; bn.z = vns[sp + 1].z; }
    }
    nbox_lo[cur] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
    nbox_hi[cur] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
    if (pnrm) nsum[cur] = make_float4(an.x + bn.x, an.y + bn.y, an.z + bn.z, 0.0f);   // left + right: order fixed by the tree
    cur = parent[cur];
  }
}

// Inner node i survives leaf collapsing iff it is the root or spans more than `leaf` elements.
__global__ void k_live(int n, int leaf, int const * __restrict__ first, int const * __restrict__ last, int * live)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n - 1) return;
  live[i] = (i == 0 || (last[i] - first[i] + 1) > leaf) ? 1 : 0;
}

// ---- oriented boxes of every radix-tree node that a surviving node references as a child --------------------------------
// obbinfo[3x .. 3x+2] = child half of an ObbNode (without the link) for inner radix node x, extents taken over all vertices of
// the node's triangles along the frame derived from the node's normal sum (obb_frame).
__device__ __forceinline__ void obb_add_tri(ObbAcc & a, TriRec const * tris, int k)
{
  f4 const * t = (f4 const *)(tris + k);
  f4 q0 = __ldg(t), q1 = __ldg(t + 1), q2 = __ldg(t + 2);
  obb_acc_point(a, mk3(q0.x, q0.y, q0.z)); obb_acc_point(a, mk3(q1.x, q1.y, q1.z)); obb_acc_point(a, mk3(q2.x, q2.y, q2.z));
}

#define TGQ_SLAB_SMALL 32
// Nodes of up to TGQ_SLAB_SMALL elements: one thread scans the node's element range; larger nodes are appended to `big`.
__global__ void k_obb_small(int n, int leaf, int const * __restrict__ first, int const * __restrict__ last, int const * __restrict__ parent,
                            f4 const * __restrict__ nsum, TriRec const * __restrict__ tris, SceneBox const * sb, f4 * obbinfo, f4 * obbinfo_w,
                            int * big, int * big_count)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x; if (x >= n - 1 || x == 0) return;
  int pa = parent[x];
  bool parent_live = pa == 0 || (last[pa] - first[pa] + 1) > leaf;
  if (!parent_live) return;                     // strictly inside a collapsed leaf: never referenced
  int f = first[x], l = last[x];
  if (l - f + 1 > TGQ_SLAB_SMALL) { big[atomicAdd(big_count, 1)] = x; return; }
  float pad = scene_decode(sb).pad;
#ifdef TGQ_NODE64
  f4 ns = nsum[pa];                             // 64-byte nodes: both children are measured in the frame of the node they belong to
#else
  f4 ns = nsum[x];
#endif
  ObbAcc a; obb_acc_begin(a, mk3(ns.x, ns.y, ns.z));
  for (int k = f; k <= l; ++k) obb_add_tri(a, tris, k);
  obb_acc_store(a, pad, 0, obbinfo + TGQ_OBB_F4 * (size_t)x);
  if (obbinfo_w) obb_acc_store_wide(a, pad, 0, obbinfo_w + 4 * (size_t)x);
}

// Larger nodes: one thread block per node, block-wide min / max reduction (exact, order independent).
__global__ void __launch_bounds__(128)
k_obb_big(int const * __restrict__ big, int const * __restrict__ big_count, int const * __restrict__ first, int const * __restrict__ last,
          int const * __restrict__ parent, f4 const * __restrict__ nsum, TriRec const * __restrict__ tris, SceneBox const * sb, f4 * obbinfo, f4 * obbinfo_w)
{
  __shared__ float s_mn[4][3], s_mx[4][3];
  for (int w = blockIdx.x; w < *big_count; w += gridDim.x)
  {
    int x = big[w];
    float pad = scene_decode(sb).pad;
#ifdef TGQ_NODE64
    f4 ns = nsum[parent[x]];
#else
    f4 ns = nsum[x];
#endif
    ObbAcc a; obb_acc_begin(a, mk3(ns.x, ns.y, ns.z));
    for (int k = first[x] + threadIdx.x; k <= last[x]; k += blockDim.x) obb_add_tri(a, tris, k);
    for (int j = 0; j < 3; ++j)
      for (int o = 16; o > 0; o >>= 1)
      { a.mn[j] = fminf(a.mn[j], __shfl_xor_sync(0xffffffffu, a.mn[j], o)); a.mx[j] = fmaxf(a.mx[j], __shfl_xor_sync(0xffffffffu, a.mx[j], o)); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) for (int j = 0; j < 3; ++j) { s_mn[threadIdx.x >> 5][j] = a.mn[j]; s_mx[threadIdx.x >> 5][j] = a.mx[j]; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      for (int j = 0; j < 3; ++j)
      {
        a.mn[j] = fminf(fminf(s_mn[0][j], s_mn[1][j]), fminf(s_mn[2][j], s_mn[3][j]));
        a.mx[j] = fmaxf(fmaxf(s_mx[0][j], s_mx[1][j]), fmaxf(s_mx[2][j], s_mx[3][j]));
      }
      obb_acc_store(a, pad, 0, obbinfo + TGQ_OBB_F4 * (size_t)x);
      if (obbinfo_w) obb_acc_store_wide(a, pad, 0, obbinfo_w + 4 * (size_t)x);
    }
  }
}

// Emits the 64-byte traversal node (and, for triangle structures, the 128-byte oriented-box node) of every surviving inner
// node at its compacted index.
__global__ void k_emit(int n, int leaf, int const * __restrict__ first, int const * __restrict__ last, int const * __restrict__ split,
                       int const * __restrict__ live, int const * __restrict__ newidx, f4 const * __restrict__ pbox_lo,
                       f4 const * __restrict__ pbox_hi, f4 const * __restrict__ nbox_lo, f4 const * __restrict__ nbox_hi,
                       SceneBox const * sb, NodeRec * nodes, TriRec const * __restrict__ tris, f4 const * __restrict__ obbinfo, ObbNode * obbs,
                       f4 const * __restrict__ nsum)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n - 1 || !live[i]) return;
  float pad = scene_decode(sb).pad;
  int f = first[i], l = last[i], sp = split[i];
  int child[2]; v3 lo[2], hi[2];
  for (int c = 0; c < 2; ++c)
  {
    int cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l;
    int cnt = cl - cf + 1;
    int inner = c == 0 ? sp : sp + 1;         // index of the child when it is an inner node of the radix tree
    f4 blo, bhi;
    if (cnt == 1) { blo = pbox_lo[cf]; bhi = pbox_hi[cf]; } else { blo = nbox_lo[inner]; bhi = nbox_hi[inner]; }
    child[c] = (cnt <= leaf) ? ~((cf << TGQ_LEAF_BITS) | cnt) : newidx[inner];
    lo[c] = mk3(blo.x - pad, blo.y - pad, blo.z - pad); hi[c] = mk3(bhi.x + pad, bhi.y + pad, bhi.z + pad);
  }
  nodes[newidx[i]] = make_node(lo[0], hi[0], lo[1], hi[1], child[0], child[1], 0, 0);
#ifdef TGQ_NODE64
  if (obbs)
  {
    // one frame per node: the children's centres and extents (measured in this frame by k_obb_small / k_obb_big, or here for a single
    // triangle) are packed behind it
    f4 ns = nsum[i]; v3 nsv = mk3(ns.x, ns.y, ns.z);
    f4 half[2][2];
    for (int c = 0; c < 2; ++c)
    {
      int cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l;
      int inner = c == 0 ? sp : sp + 1;
      if (cl == cf)
      {
        f4 const * t = (f4 const *)(tris + cf);
        f4 q0 = __ldg(t), q1 = __ldg(t + 1), q2 = __ldg(t + 2);
        ObbAcc a; obb_acc_begin(a, nsv);
        obb_acc_point(a, mk3(q0.x, q0.y, q0.z)); obb_acc_point(a, mk3(q1.x, q1.y, q1.z)); obb_acc_point(a, mk3(q2.x, q2.y, q2.z));
        obb_acc_store(a, pad, child[c], half[c]);
      }
      else
      {
        half[c][0] = obbinfo[TGQ_OBB_F4 * (size_t)inner]; half[c][1] = obbinfo[TGQ_OBB_F4 * (size_t)inner + 1];
        half[c][0].w = __int_as_float(child[c]);
      }
    }
    ObbNode r; obb_pack_node64(obb_node_frame64(nsv), half[0], half[1], r.q);
    obbs[newidx[i]] = r;
  }
#else
  if (obbs)
  {
    ObbNode r;
    for (int c = 0; c < 2; ++c)
    {
      int cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l;
      int inner = c == 0 ? sp : sp + 1;
      if (cl == cf)
      {
        // single triangle: box along its own normal
        f4 const * t = (f4 const *)(tris + cf);
        f4 q0 = __ldg(t), q1 = __ldg(t + 1), q2 = __ldg(t + 2);
        v3 v0 = mk3(q0.x, q0.y, q0.z), v1 = mk3(q1.x, q1.y, q1.z), v2 = mk3(q2.x, q2.y, q2.z);
        ObbAcc a; obb_acc_begin(a, cross3(sub3(v1, v0), sub3(v2, v0)));
        obb_acc_point(a, v0); obb_acc_point(a, v1); obb_acc_point(a, v2);
        obb_acc_store(a, pad, child[c], r.q + TGQ_OBB_F4 * c);
      }
      else
      {
        for (int j = 0; j < TGQ_OBB_F4; ++j) r.q[TGQ_OBB_F4 * c + j] = obbinfo[TGQ_OBB_F4 * (size_t)inner + j];
        r.q[TGQ_OBB_F4 * c + TGQ_OBB_LINK_F4].w = __int_as_float(child[c]);
      }
    }
    obbs[newidx[i]] = r;
  }
#endif
}

// ---- 4-wide nodes (WideNode, tgq_bvh.cuh): every second level of the surviving binary tree ---------------------------------------
// A surviving inner node is a wide node iff its depth below the root is even; its wide children are its binary grandchildren (or a
// binary child that is a leaf).  Depth by walking up the parent chain (<= ~100 steps, all in L2).
__global__ void k_wide_flags(int n, int const * __restrict__ live, int const * __restrict__ parent, int * wide)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n - 1) return;
  int w = 0;
  if (live[i])
  {
    int d = 0;
    for (int y = parent[i]; y >= 0; y = parent[y]) ++d;
    w = (d & 1) ? 0 : 1;
  }
  wide[i] = w;
}

__device__ __forceinline__ void wide_put_child(WideNode & r, int slot, f4 const * c4)
{ r.q[2 * slot] = c4[0]; r.q[2 * slot + 1] = c4[1]; r.q[8 + 2 * slot] = c4[2]; r.q[8 + 2 * slot + 1] = c4[3]; }

// Child record of the element range [cf, cl] hanging below a wide node: its oriented box from obbinfo_w (ranges of two or more
// elements: inner radix node `inner`) or from the single triangle, and the link (a leaf code, or the wide index of `inner`).
__device__ __forceinline__ void wide_child_record(int cf, int cl, int inner, int leaf, int const * __restrict__ widx, TriRec const * __restrict__ tris,
                                                  f4 const * __restrict__ obbinfo_w, float pad, f4 * c4)
{
  int cnt = cl - cf + 1;
  int link = (cnt <= leaf) ? ~((cf << TGQ_LEAF_BITS) | cnt) : widx[inner];
  if (cnt == 1)
  {
    f4 const * t = (f4 const *)(tris + cf);
    f4 q0 = __ldg(t), q1 = __ldg(t + 1), q2 = __ldg(t + 2);
    v3 v0 = mk3(q0.x, q0.y, q0.z), v1 = mk3(q1.x, q1.y, q1.z), v2 = mk3(q2.x, q2.y, q2.z);
    ObbAcc a; obb_acc_begin(a, cross3(sub3(v1, v0), sub3(v2, v0)));
    obb_acc_point(a, v0); obb_acc_point(a, v1); obb_acc_point(a, v2);
    obb_acc_store_wide(a, pad, link, c4);
  }
  else
  {
    for (int j = 0; j < 4; ++j) c4[j] = obbinfo_w[4 * (size_t)inner + j];
    c4[3].w = __int_as_float(link);
  }
}

__global__ void k_emit_wide(int n, int leaf, int const * __restrict__ first, int const * __restrict__ last, int const * __restrict__ split,
                            int const * __restrict__ wide, int const * __restrict__ widx, SceneBox const * sb, TriRec const * __restrict__ tris,
                            f4 const * __restrict__ obbinfo_w, WideNode * out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n - 1 || !wide[i]) return;
  float pad = scene_decode(sb).pad;
  WideNode r; int slot = 0; f4 c4[4];
  int f = first[i], l = last[i], sp = split[i];
  for (int c = 0; c < 2; ++c)
  {
    int cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l, inner = c == 0 ? sp : sp + 1;
    if (cl - cf + 1 <= leaf) { wide_child_record(cf, cl, inner, leaf, widx, tris, obbinfo_w, pad, c4); wide_put_child(r, slot++, c4); continue; }
    // the binary child is an inner node (odd depth): its two children become children of this wide node
    int gsp = split[inner];
    for (int g = 0; g < 2; ++g)
    {
      int gf = g == 0 ? cf : gsp + 1, gl = g == 0 ? gsp : cl, ginner = g == 0 ? gsp : gsp + 1;
      wide_child_record(gf, gl, ginner, leaf, widx, tris, obbinfo_w, pad, c4); wide_put_child(r, slot++, c4);
    }
  }
  for (; slot < 4; ++slot) { wide_store_empty(c4); wide_put_child(r, slot, c4); }
  out[widx[i]] = r;
}

__global__ void k_emit_wide_single(SceneBox const * sb, TriRec const * tris, WideNode * out)
{
  float pad = scene_decode(sb).pad;
  WideNode r; f4 c4[4];
  wide_child_record(0, 0, 0, 1, nullptr, tris, nullptr, pad, c4); wide_put_child(r, 0, c4);
  for (int slot = 1; slot < 4; ++slot) { wide_store_empty(c4); wide_put_child(r, slot, c4); }
  out[0] = r;
}

// Structure with a single element: one node whose second child is an empty leaf behind an unreachable box.
__global__ void k_emit_single(f4 const * pbox_lo, f4 const * pbox_hi, SceneBox const * sb, NodeRec * nodes, TriRec const * tris, ObbNode * obbs)
{
  float pad = scene_decode(sb).pad;
  int leaf0 = ~((0 << TGQ_LEAF_BITS) | 1);
  if (obbs)
  {
    ObbNode r;
    f4 const * t = (f4 const *)tris;
    f4 q0 = t[0], q1 = t[1], q2 = t[2];
    v3 v0 = mk3(q0.x, q0.y, q0.z), v1 = mk3(q1.x, q1.y, q1.z), v2 = mk3(q2.x, q2.y, q2.z);
    ObbAcc a; obb_acc_begin(a, cross3(sub3(v1, v0), sub3(v2, v0)));
    obb_acc_point(a, v0); obb_acc_point(a, v1); obb_acc_point(a, v2);
#ifdef TGQ_NODE64
    f4 half0[2], half1[2];
    obb_acc_store(a, pad, leaf0, half0); obb_store_empty(~0, half1);
    obb_pack_node64(a, half0, half1, r.q);
#else
    obb_acc_store(a, pad, leaf0, r.q);
    obb_store_empty(~0, r.q + TGQ_OBB_F4);
#endif
    obbs[0] = r;
  }
  f4 blo = pbox_lo[0], bhi = pbox_hi[0];
  nodes[0] = make_node(mk3(blo.x - pad, blo.y - pad, blo.z - pad), mk3(bhi.x + pad, bhi.y + pad, bhi.z + pad),
                       mk3(FLT_MAX, FLT_MAX, FLT_MAX), mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX), leaf0, ~0, 0, 0);
}

// Element-order copy of the triangle cache (for tgq_triangle_cache and hit normals).
__global__ void k_scatter_cache(TriRec const * __restrict__ tris, float const * __restrict__ area_by_elem, int n, float * n3, float * d,
                                int * axis, float * area)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= n) return;
  TriRec r = tris[s]; int64_t i = __float_as_int(r.q0.w);
  if (n3) st3(n3, i, mk3(r.q3.x, r.q3.y, r.q3.z));
  if (d) d[i] = r.q2.w;
  if (axis) axis[i] = __float_as_int(r.q1.w);
  if (area) area[i] = area_by_elem[i];
}

// ---- query ordering ---------------------------------------------------------------------------------------------------
// dir3 != null (ray batches, TGQ_OPT_RAY_ORDER): the `obits` lowest bits of the origin's code are replaced by the direction's sign
// octant (3 bits) above the remaining origin bits -- rays of one cell are grouped by the octant they head into.
__global__ void k_query_codes(float const * __restrict__ q3, int n, SceneBox const * sb, uint32_t * keys, uint32_t * vals, float const * __restrict__ dir3 = nullptr,
                              int obits = 0, int stride3 = 1)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  SceneF s = scene_decode(sb);
  // codes over the scene box grown x2 about its centre so that off-surface queries keep their locality
  v3 c = mul3(add3(s.lo, s.hi), 0.5f), e = sub3(s.hi, s.lo);
  v3 lo = sub3(c, e);
  v3 inv = mk3(e.x > 0 ? 0.5f / e.x : 0.0f, e.y > 0 ? 0.5f / e.y : 0.0f, e.z > 0 ? 0.5f / e.z : 0.0f);
  uint32_t code = morton30(ld3(q3, (int64_t)i * stride3), lo, inv);      // stride3 = 3: triangle query objects, ordered by their first vertex
  if (dir3 && obits >= 3)
  {
    v3 d = ld3(dir3, i);
    uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
    uint32_t low = code & ((1u << obits) - 1u);
    code = (code & ~((1u << obits) - 1u)) | (oct << (obits - 3)) | (low >> 3);
  }
  keys[i] = code; vals[i] = (uint32_t)i;
}

// ---- sharding a replicated batch by space (multi-GPU) --------------------------------------------------------------------------
// Every rank holds the whole batch and takes the queries whose Morton cell (the top `cell_bits` bits of the 30-bit code) it owns:
// cells are dealt round-robin, cell % world == rank, so a rank's share is spread over the whole scene in fine cells (cost balances
// statistically) while the queries inside a cell stay as dense as in the full batch -- a contiguous index range of a random batch
// would be `world` times sparser, and neighbouring queries of the Morton-ordered launch would share fewer nodes.  One pass over the
// batch: own (code, index) pairs are appended in arbitrary order (warp-aggregated atomic), then only they are sorted.
__global__ void __launch_bounds__(256)
k_query_codes_shard(float const * __restrict__ q3, int n, SceneBox const * sb, int rank, int world, int shift, uint32_t * keys, uint32_t * vals, int * count)
{
  SceneF s = scene_decode(sb);
  v3 c = mul3(add3(s.lo, s.hi), 0.5f), e = sub3(s.hi, s.lo);
  v3 lo = sub3(c, e);
  v3 inv = mk3(e.x > 0 ? 0.5f / e.x : 0.0f, e.y > 0 ? 0.5f / e.y : 0.0f, e.z > 0 ? 0.5f / e.z : 0.0f);
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int s_warp[8], s_base;
  // one global atomic per block and tile of 1024 queries (per-warp atomics on a single address serialise: 1.5 M of them cost 4 ms)
  for (int64_t tile = blockIdx.x * (int64_t)1024; tile < n; tile += (int64_t)gridDim.x * 1024)
  {
    uint32_t code[4]; unsigned m[4]; int mine = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      int64_t i = tile + k * 256 + threadIdx.x;
      bool own = false; code[k] = 0;
      if (i < n) { code[k] = morton30(ld3(q3, i), lo, inv); own = (int)((code[k] >> shift) % (uint32_t)world) == rank; }
      m[k] = __ballot_sync(0xffffffffu, own);
      mine += __popc(m[k]);
    }
    if (lane == 0) s_warp[warp] = mine;          // `mine` is warp uniform: the warp's count over its four rows
    __syncthreads();
    if (threadIdx.x == 0)
    {
      int tot = 0;
      for (int w = 0; w < 8; ++w) { int t = s_warp[w]; s_warp[w] = tot; tot += t; }
      s_base = tot ? atomicAdd(count, tot) : 0;
    }
    __syncthreads();
    int pos = s_base + s_warp[warp];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      if ((m[k] >> lane) & 1u) { int p = pos + __popc(m[k] & ((1u << lane) - 1u)); keys[p] = code[k]; vals[p] = (uint32_t)(tile + k * 256 + threadIdx.x); }
      pos += __popc(m[k]);
    }
    __syncthreads();
  }
}

// Result records of the *_records entry points: one aligned vector store per result and destination, so that a rank can write its
// results straight into every peer's copy of the gathered array over NVLink (fire-and-forget stores under the traversal).
//   closest: 32 bytes  (element index, squared distance or -1, closest point xyz, 3 words of padding)
//   ray:      8 bytes  (hit time or -1, element index or -1)
//   compact form (one destination): the record goes to position [rank of the query in this rank's Morton-ordered share] instead, and
//   carries the query index -- closest: word 5 of the 32 bytes; ray: 16 bytes { t, idx, query index, 0 } -- so that the share can be
//   moved to the other ranks as ONE contiguous block (copy engines) and scattered to [query index] there (k_scatter_records).
struct PeerOut { void * base[8]; int n; int compact; };        // n == 0: the ordinary per-field outputs are used
__device__ __forceinline__ void st_rec32(void * base, int64_t i, float4 a, float4 b)
{
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"((char *)base + 32 * i), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y),
               "f"(b.z), "f"(b.w) : "memory");
}
__device__ __forceinline__ void st_rec8(void * base, int64_t i, float a, float b)
{ asm volatile("st.global.v2.f32 [%0], {%1,%2};" :: "l"((char *)base + 8 * i), "f"(a), "f"(b) : "memory"); }

// A rank's dense share of records, split by the rank that owns each record's POSITION in the result array (query index /
// shard_len): bucket d of `out` (cap records each) receives the records for rank d, counts[d] how many (zeroed by the caller).
// One shared-memory histogram and `world` global atomics per tile of 256 records.
__global__ void __launch_bounds__(256)
k_bucket_records(void const * __restrict__ src, int64_t n, int kind, int64_t shard_len, int world, void * out, int64_t cap, unsigned long long * counts)
{
  __shared__ int s_cnt[8]; __shared__ unsigned long long s_base[8];
  for (int64_t tile = blockIdx.x * (int64_t)256; tile < n; tile += (int64_t)gridDim.x * 256)
  {
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int64_t i = tile + threadIdx.x;
    float4 a = make_float4(0, 0, 0, 0), b = a; int d = 0, local = 0;
    if (i < n)
    {
      if (kind == 0) { a = __ldg((float4 const *)src + 2 * i); b = __ldg((float4 const *)src + 2 * i + 1); d = (int)min((int64_t)__float_as_int(b.y) / shard_len, (int64_t)world - 1); }
      else { a = __ldg((float4 const *)src + i); d = (int)min((int64_t)__float_as_int(a.z) / shard_len, (int64_t)world - 1); }
      local = atomicAdd(&s_cnt[d], 1);
    }
    __syncthreads();
    if (threadIdx.x < world) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]) : 0ull;
    __syncthreads();
    if (i < n)
    {
      int64_t slot = (int64_t)s_base[d] + local;
      if (slot < cap)
      {
        if (kind == 0) { float4 * o = (float4 *)out + 2 * (d * cap + slot); o[0] = a; o[1] = b; }
        else ((float4 *)out)[d * cap + slot] = a;
      }
    }
    __syncthreads();
  }
}

// Compact records of the ranks' shares -> [query index] of the gathered array.  `slots` dense arrays of up to `cap` records each
// (slot j at src + j * cap records), counts[j] of them valid (device memory: the counts arrive with the data).  kind 0: 32-byte
// closest records (query index in word 5), kind 1: 16-byte ray records { t, idx, query index, - } -> 8-byte records.
__global__ void k_scatter_records(void const * __restrict__ src, int slots, int64_t cap, long long const * __restrict__ counts, int kind, void * dst)
{
  for (int j = 0; j < slots; ++j)
  {
    int64_t n = counts[j]; if (n > cap) n = cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    {
      if (kind == 0)
      {
        float4 const * p = (float4 const *)src + 2 * (j * cap + i);
        float4 a = __ldg(p), b = __ldg(p + 1);
        st_rec32(dst, __float_as_int(b.y), a, b);
      }
      else
      {
        float4 a = __ldg((float4 const *)src + (j * cap + i));
        st_rec8(dst, __float_as_int(a.z), a.x, a.y);
      }
    }
  }
}

// ---- persistent warp scheduler -------------------------------------------------------------------------------------------
// Query costs are heavy tailed and lanes follow different paths, so a thread-per-query kernel leaves most lanes of a warp
// idle (ncu: 1.2 - 10 active threads per instruction on the bench workload).  Instead every warp drives 32 resumable
// state machines (tgq_traverse.cuh) from a warp-uniform loop: in each iteration every lane with work pending fetches ONE
// record (an inner node or a leaf element: same fetch shape, one load phase for the warp), then the lanes at a node do
// the node arithmetic and the lanes at a leaf the element arithmetic, so the lanes reconverge after every step; lanes
// whose query finished write their result and are refilled from the warp's current chunk of the (Morton-ordered) batch,
// which the warp stages in shared memory with coalesced loads.  Chunks are handed out through one atomic counter, which
// also balances the load across SMs.
//   W: FETCH, STAGE ; int state() ; void stage(t, dst) ; void begin(t, src) ; void fetch(r) ; void node_compute(r) ;
//      void leaf_compute(r) ; void finish()  [finish leaves the lane in TGQ_ST_IDLE, or restarts it]
#ifndef TGQ_CHUNK
#define TGQ_CHUNK 64
#endif
#ifndef TGQ_POPS
#define TGQ_POPS 2
#endif
#ifndef TGQ_SSTACK
#define TGQ_SSTACK 20
#endif
// stack entries per thread kept in shared memory (deeper ones spill to local memory)
#define TGQ_WARPS (TGQ_BLOCK / 32)

// Shared memory of one block of a traversal kernel: the short stacks, then one staging buffer of TGQ_CHUNK query records
// per warp (STAGE float4 per record), then whatever the kernel itself needs (k-NN sets).
__host__ __device__ constexpr size_t trav_smem_bytes(int stage_f4, int sstack = TGQ_SSTACK) { return (size_t)sstack * TGQ_BLOCK * sizeof(StackEntry) + (size_t)TGQ_WARPS * TGQ_CHUNK * stage_f4 * sizeof(f4); }

template <typename W>
__device__ __forceinline__ void warp_run(W & w, int n, int * counter, f4 * stage)
{
  unsigned const FULL = 0xffffffffu;
  int lane = threadIdx.x & 31;
  int chunk_base = 0, chunk_next = 0, chunk_end = 0; bool exhausted = false;
  for (;;)
  {
    if (w.state() == TGQ_ST_DONE) w.finish();
    unsigned idle = __ballot_sync(FULL, w.state() == TGQ_ST_IDLE);
    if (idle != 0u && !exhausted)
    {
      if (chunk_next >= chunk_end)
      {
        int base = 0;
        if (lane == 0) base = atomicAdd(counter, TGQ_CHUNK);
        base = __shfl_sync(FULL, base, 0);
        chunk_base = chunk_next = base; chunk_end = min(base + TGQ_CHUNK, n);
        if (base >= n) exhausted = true;
        else
        {
          // all lanes fetch the chunk's query records (coalesced, one latency per TGQ_CHUNK queries instead of one per refill)
          __syncwarp();
          for (int j = lane; j < chunk_end - chunk_base; j += 32) w.stage(chunk_base + j, stage + (size_t)j * W::STAGE);
          __syncwarp();
        }
      }
      if (!exhausted)
      {
        int t = chunk_next + __popc(idle & ((1u << lane) - 1u));
        if (w.state() == TGQ_ST_IDLE && t < chunk_end) w.begin(t, stage + (size_t)(t - chunk_base) * W::STAGE);
        chunk_next = min(chunk_next + __popc(idle), chunk_end);
      }
    }
    int st = w.state();
    bool at_node = st >= 0, at_leaf = st == TGQ_ST_LEAF;
    if (!__any_sync(FULL, at_node || at_leaf || st == TGQ_ST_POP))
    {
      if (exhausted && __all_sync(FULL, w.state() != TGQ_ST_DONE)) break;
      continue;
    }
    // one load phase for the whole warp (nodes and leaf records have the same fetch shape), then the arithmetic
    f4 r[W::FETCH];
    if (at_node || at_leaf) w.fetch(r);
    if (at_node) w.node_compute(r);
    else if (at_leaf) w.leaf_compute(r);
    // pop phase: every lane that needs a deferred sibling takes (up to) TGQ_POPS off its stack, all such lanes together
#pragma unroll
    for (int j = 0; j < TGQ_POPS; ++j) if (w.state() == TGQ_ST_POP) w.pop_step();
  }
}

// Binds the calling thread's short stack and returns its warp's staging buffer; `extra` = first byte after both.
__device__ __forceinline__ f4 * trav_setup(Trav & t, TravStack & st, unsigned char * smem, int stage_f4, unsigned char ** extra, int sstack = TGQ_SSTACK)
{
  trav_bind(t, st);
  trav_bind_fast(t, (StackEntry *)smem + threadIdx.x, TGQ_BLOCK, sstack);
  f4 * stage_all = (f4 *)(smem + (size_t)sstack * TGQ_BLOCK * sizeof(StackEntry));
  if (extra) *extra = (unsigned char *)(stage_all + (size_t)TGQ_WARPS * TGQ_CHUNK * stage_f4);
  return stage_all + (size_t)(threadIdx.x >> 5) * TGQ_CHUNK * stage_f4;
}

// ---- K5 closest element -------------------------------------------------------------------------------------------------
template <bool TRI>
struct ClosestW
{
  static constexpr int STAGE = 1;       // (x, y, z, query index)
  ClosestState<TRI> s; BvhView b; float scene_pad; float const * q3; uint32_t const * perm; float ub; BoundGate gate; int exclude_self; int64_t self_offset;
  int * idx; float * d2; float * cp3; int qi; int seed;
  float const * qn3; int const * qlabel; float min_dot;      // per-query element filter (point structures), null = none
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 p = ld3(q3, i);
    *dst = make_float4(p.x, p.y, p.z, __int_as_float(i));
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 r = *src; qi = __float_as_int(r.w);
    v3 p = mk3(r.x, r.y, r.z);
    closest_begin(s, b, scene_pad, p, exclude_self ? (int)(self_offset + qi) : -1, gated_ub(gate, p, ub), seed);
    if (!TRI && (qn3 || qlabel))
    {
      s.f.mode = (qn3 ? 1 : 0) | (qlabel ? 2 : 0); s.f.min_dot = min_dot;
      if (qn3) s.f.n = ld3(qn3, qi);
      if (qlabel) s.f.label = qlabel[qi];
    }
  }
  static constexpr int FETCH = ClosestFetch<TRI>::N;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, s.best_d2); }
  __device__ __forceinline__ void fetch(f4 * r) const { closest_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { closest_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { closest_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    if (idx) idx[qi] = s.best_idx;
    if (d2) d2[qi] = s.best_idx >= 0 ? s.best_d2 : -1.0f;
    if (cp3) st3(cp3, qi, s.best_cp);
    seed = perm ? s.best_k : -1;            // only Morton-ordered batches have neighbouring queries next to each other
    s.t.cur = TGQ_ST_IDLE;
  }
};

template <bool TRI>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_closest(BvhView b, float scene_pad, float const * __restrict__ q3, uint32_t const * __restrict__ perm, int n, float bound_ub, BoundGate gate,
          int exclude_self, int64_t self_offset, int * counter, int * idx, float * d2, float * cp3, float const * __restrict__ qn3,
          int const * __restrict__ qlabel, float min_dot)
{
  extern __shared__ __align__(16) unsigned char smem[];
  ClosestW<TRI> w;
  w.b = b; w.scene_pad = scene_pad; w.q3 = q3; w.perm = perm; w.ub = bound_ub; w.gate = gate; w.exclude_self = exclude_self; w.self_offset = self_offset;
  w.idx = idx; w.d2 = d2; w.cp3 = cp3; w.qi = 0; w.seed = -1; w.qn3 = qn3; w.qlabel = qlabel; w.min_dot = min_dot; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  f4 * stage = trav_setup(w.s.t, st, smem, ClosestW<TRI>::STAGE, nullptr);
  warp_run(w, n, counter, stage);
}

// ---- K5b closest element on triangle structures, leaf tests pooled across the warp ----------------------------------------
// In k_closest<true> a lane that reaches a leaf runs the triangle screen and the reference's point-triangle arithmetic in its own
// step, so every warp iteration pays for the node arithmetic AND the leaf arithmetic while only about a third of the lanes are at a
// leaf (ncu: 11 of 32 lanes active per issued instruction in the leaf code).  Here a lane that reaches a leaf appends its triangles
// to a per-warp queue in shared memory -- entries (owner lane, Morton position) -- and goes straight back to its stack; when 32
// entries are waiting (or nobody can traverse) the whole warp runs ONE leaf iteration with every lane on a different entry: the
// owner's query is read with shuffles, the result is merged into the owner's best with a 64-bit shared-memory atomicMin on
// (squared distance bits, element index) -- exactly the canonical order: smaller distance, then lower index.  The owner's pruning
// bound is therefore a few iterations stale, which only makes pruning more conservative; the answer is the minimum over the same
// exact per-element values and does not depend on the order in which elements are tested.  A query seeded with its lane's previous
// answer holds its traversal until that one test has come back (traversing with an infinite bound would open the whole tree).
#ifndef TGQ_QCAP
#define TGQ_QCAP 256          // >= 31 waiting + 32 seeds + 32 lanes x 4 leaf triangles per iteration (leaf size <= 4)
#endif
struct PoolWarp
{
  unsigned long long best[32];
  f4 cp[32];                  // closest point + Morton position of the winning element
  int done[32];
  unsigned q[TGQ_QCAP];
  int tail; int pad_[3];
};

__device__ __forceinline__ unsigned long long pool_key(float d2, int idx)
{ return ((unsigned long long)(unsigned)__float_as_int(d2) << 32) | (unsigned long long)(unsigned)idx; }

__global__ void __launch_bounds__(TGQ_BLOCK)
k_closest_pool(BvhView b, float scene_pad, float const * __restrict__ q3, uint32_t const * __restrict__ perm, int n, float bound_ub, BoundGate gate,
               int exclude_self, int64_t self_offset, int * counter, int * idx_out, float * d2_out, float * cp3, PeerOut peers)
{
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned const FULL = 0xffffffffu;
  int const lane = threadIdx.x & 31;
  ClosestState<true> s; TravStack st; s.t.cur = TGQ_ST_IDLE;
  unsigned char * extra; f4 * stage = trav_setup(s.t, st, smem, 1, &extra);
  PoolWarp & pw = ((PoolWarp *)extra)[threadIdx.x >> 5];
  pw.best[lane] = 0ull; pw.done[lane] = 0; if (lane == 0) pw.tail = 0;
  __syncwarp();
  int qi = 0, slot = 0, seed = -1, pending = 0, head = 0; bool hold = false;
  int chunk_base = 0, chunk_next = 0, chunk_end = 0; bool exhausted = false;
  for (;;)
  {
    // ---- finished queries write their results; idle lanes take the next queries of the warp's chunk -----------------------------
    if (s.t.cur == TGQ_ST_DONE && pending == 0)
    {
#ifndef TGQ_NO_PEERS
      if (peers.n)
#else
      if (false)
#endif
      {
        float4 ra = make_float4(__int_as_float(s.best_idx), s.best_idx >= 0 ? s.best_d2 : -1.0f, s.best_cp.x, s.best_cp.y);
        float4 rb = make_float4(s.best_cp.z, __int_as_float(qi), 0.0f, 0.0f);
        if (peers.compact) st_rec32(peers.base[0], slot, ra, rb);
        else
        {
          // constant indices only: a run-time index would move the parameter array to local memory (measured: -5 % on the whole kernel)
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < peers.n) st_rec32(peers.base[j], qi, ra, rb);
        }
      }
      if (idx_out) idx_out[qi] = s.best_idx;
      if (d2_out) d2_out[qi] = s.best_idx >= 0 ? s.best_d2 : -1.0f;
      if (cp3) st3(cp3, qi, s.best_cp);
      seed = perm ? s.best_k : -1;
      s.t.cur = TGQ_ST_IDLE;
    }
    unsigned idle = __ballot_sync(FULL, s.t.cur == TGQ_ST_IDLE);
    if (idle != 0u && !exhausted)
    {
      if (chunk_next >= chunk_end)
      {
        int base = 0;
        if (lane == 0) base = atomicAdd(counter, TGQ_CHUNK);
        base = __shfl_sync(FULL, base, 0);
        chunk_base = chunk_next = base; chunk_end = min(base + TGQ_CHUNK, n);
        if (base >= n) exhausted = true;
        else
        {
          __syncwarp();
          for (int j = lane; j < chunk_end - chunk_base; j += 32)
          {
            int t = chunk_base + j; int i = perm ? (int)perm[t] : t; v3 p = ld3(q3, i);
            stage[j] = make_float4(p.x, p.y, p.z, __int_as_float(i));
          }
          __syncwarp();
        }
      }
      if (!exhausted)
      {
        int t = chunk_next + __popc(idle & ((1u << lane) - 1u));
        if (s.t.cur == TGQ_ST_IDLE && t < chunk_end)
        {
          f4 r = stage[t - chunk_base]; qi = __float_as_int(r.w); slot = t;
          v3 p = mk3(r.x, r.y, r.z);
          closest_begin(s, b, scene_pad, p, exclude_self ? (int)(self_offset + qi) : -1, gated_ub(gate, p, bound_ub), seed);
          hold = s.t.cur == TGQ_ST_LEAF;          // seeded: the seed element is tested before the traversal starts
        }
        chunk_next = min(chunk_next + __popc(idle), chunk_end);
      }
    }
    // ---- leaves reached (by the last node step, the last pops, or a seed) go to the warp's queue ---------------------------------
    if (s.t.cur == TGQ_ST_LEAF)
    {
      int cnt = s.t.kend - s.t.k;
      int pos = atomicAdd(&pw.tail, cnt);
      // straight-line (this kernel runs on leaves of at most four triangles): ncu's source view had the counted loop here at 11.5 % of
      // all warp instructions with 2.7 lanes active -- few lanes reach a leaf in the same iteration, each with its own trip count
      unsigned const entry = ((unsigned)lane << 27) | (unsigned)s.t.k;
#pragma unroll
      for (int j = 0; j < 4; ++j) if (j < cnt) pw.q[(pos + j) & (TGQ_QCAP - 1)] = entry + (unsigned)j;
      pending += cnt;
      trav_pop(s.t);
    }
    __syncwarp();
    // ---- leaf iterations: 32 queued (owner, triangle) pairs at a time, every lane on its own pair ---------------------------------
    for (;;)
    {
      int avail = pw.tail - head;
      bool runnable = (s.t.cur >= 0 || s.t.cur == TGQ_ST_POP) && !hold;
      if (!(avail >= 32 || (avail > 0 && !__any_sync(FULL, runnable)))) break;
      int m = min(avail, 32);
      pw.best[lane] = pool_key(s.best_d2, s.best_idx < 0 ? 0 : s.best_idx);      // (ub, 0) when nothing was found yet: only d2 < ub wins
      unsigned long long published = pw.best[lane];
      bool valid = lane < m;
      unsigned e = valid ? pw.q[(head + lane) & (TGQ_QCAP - 1)] : ((unsigned)lane << 27);
      int o = (int)(e >> 27), k = (int)(e & 0x7ffffffu);
      __syncwarp();
      v3 p = mk3(__shfl_sync(FULL, s.p.x, o), __shfl_sync(FULL, s.p.y, o), __shfl_sync(FULL, s.p.z, o));
      float mq = __shfl_sync(FULL, s.mq, o), bd2 = __shfl_sync(FULL, s.best_d2, o);
      int skip = __shfl_sync(FULL, s.skip, o);
      unsigned long long mykey = ~0ull; v3 cp = mk3(0.0f, 0.0f, 0.0f);
      if (valid)
      {
        f4 const * q = (f4 const *)(b.tris + k);
        f4 r0, r1, r2, r3; ldq2(q, r0, r1); ldq2(q + 2, r2, r3);
        v3 v0 = mk3(r0.x, r0.y, r0.z), v1 = mk3(r1.x, r1.y, r1.z), v2 = mk3(r2.x, r2.y, r2.z), nn = mk3(r3.x, r3.y, r3.z);
        v3 cen = mul3(add3(add3(v0, v1), v2), 0.33333334f);
        float sd = fmaxf(fabsf(fdot3(p, nn) - r2.w) - mq, 0.0f);
        v3 pc = sub3(p, cen);
        v3 perp = sub3(pc, mul3(nn, fdot3(nn, pc)));
        float lat = fmaxf(sqrtf(fdot3(perp, perp)) - r3.w - mq, 0.0f);
        int idx = f2i(r0.w);
        if ((sd * sd + lat * lat) * TGQ_LB_DEFLATE <= bd2 && idx != skip)
        {
          f4 r4, r5, r6, r7; ldq2(q + 4, r4, r5); ldq2(q + 6, r6, r7);
          float d2 = tri_point_sqdist_pre(v0, v1, v2, nn, r2.w, f2i(r1.w), r4, r5, r6, r7, p, cp);
          if (d2 <= bd2) { mykey = pool_key(d2, idx); atomicMin(&pw.best[o], mykey); }     // only candidates that can win (NaN fails, as in the reference's d2 < best)
        }
        atomicAdd(&pw.done[o], 1);
      }
      __syncwarp();
      if (mykey != ~0ull && pw.best[o] == mykey) pw.cp[o] = make_float4(cp.x, cp.y, cp.z, __int_as_float(k));
      __syncwarp();
      unsigned long long now = pw.best[lane];
      if (now != published)
      {
        f4 c = pw.cp[lane];
        s.best_d2 = __int_as_float((int)(unsigned)(now >> 32)); s.best_idx = (int)(unsigned)(now & 0xffffffffull);
        s.best_cp = mk3(c.x, c.y, c.z); s.best_k = __float_as_int(c.w);
      }
      pending -= pw.done[lane]; pw.done[lane] = 0;
      if (pending == 0) hold = false;
      head += m;
      __syncwarp();
    }
    // ---- node step for the lanes at a node, then the pop phase --------------------------------------------------------------------
    bool at_node = s.t.cur >= 0 && !hold;
    bool want_pop = s.t.cur == TGQ_ST_POP && !hold;
    if (!__any_sync(FULL, at_node || want_pop))
    {
      bool busy = s.t.cur != TGQ_ST_IDLE;        // waiting for queued tests, or about to be finished at the top of the loop
      if (exhausted && !__any_sync(FULL, busy)) break;
      continue;
    }
    if (at_node)
    {
      f4 r[2 * TGQ_OBB_F4]; f4 const * q = (f4 const *)(b.obbs + s.t.cur);
      ldq2(q, r[0], r[1]); ldq2(q + 2, r[2], r[3]);
#ifndef TGQ_NODE64
      ldq2(q + 4, r[4], r[5]);
#endif
      closest_node_compute(s, b, r);
    }
#pragma unroll
    for (int j = 0; j < TGQ_POPS; ++j) if (s.t.cur == TGQ_ST_POP && !hold) trav_pop_step(s.t, s.best_d2);
  }
}

// ---- K5c closest element on triangle structures: FOUR lanes per query on 4-wide nodes, leaf tests pooled across the warp -------
// ncu on k_closest_pool (profiles/r01_v_*): L1 LSU data-pipe wavefronts at 87 % of peak with every lane reading its own line
// (528 global wavefronts per query, 70 % of them the three 256-bit loads of a binary node), 18.5 of 32 lanes per instruction.  Here a
// warp carries EIGHT queries; the four lanes of a group hold the query state redundantly and each tests one child of the current
// WideNode, whose record is laid out so that the group's loads are two full 128-byte lines (2 wavefronts per visit of a node with up
// to four children).  The children that pass the bound are ranked by centre distance with three shuffles; the nearest is entered,
// the others go to the query's stack in shared memory (one entry written by each lane, popped by all four as a broadcast read).
// Leaves go to the per-warp queue of (owner group, triangle) pairs exactly as in k_closest_pool: 32 pairs at a time, every lane on
// its own pair, the reference's point-triangle arithmetic, 64-bit atomicMin merge in (distance, index) order.  The answer is the
// minimum over the same exact per-element values as before, so parity does not depend on any of this.
#ifndef TGQ_WSTACK
#define TGQ_WSTACK 32          // stack entries per query kept in shared memory (deeper ones spill to local memory)
#endif
#define TGQ_WOVF 96            // > 3 x (63 Morton + 32 tie-break bits) / 2 levels beyond the shared part cannot occur in ordered DFS
#ifndef TGQ_WIDE_THRESH
#define TGQ_WIDE_THRESH 32     // queued (owner, triangle) pairs that trigger a leaf iteration
#endif
struct WidePool
{
  unsigned long long best[8];
  f4 cp[8];
  int done[8];
  unsigned q[TGQ_QCAP];
  int tail; int pad_[3];
};
__host__ __device__ constexpr size_t wide_smem_bytes()
{ return (size_t)TGQ_WARPS * 8 * TGQ_WSTACK * sizeof(StackEntry) + (size_t)TGQ_WARPS * TGQ_CHUNK * sizeof(f4) + (size_t)TGQ_WARPS * sizeof(WidePool); }

__global__ void __launch_bounds__(TGQ_BLOCK)
k_closest_wide(BvhView b, float scene_pad, float const * __restrict__ q3, uint32_t const * __restrict__ perm, int n, float bound_ub, BoundGate gate,
               int exclude_self, int64_t self_offset, int * counter, int * idx_out, float * d2_out, float * cp3)
{
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned const FULL = 0xffffffffu;
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 3, gbase = lane & ~3, g = lane >> 2;
  StackEntry * stk = (StackEntry *)smem + (size_t)(warp * 8 + g) * TGQ_WSTACK;
  f4 * stage = (f4 *)(smem + (size_t)TGQ_WARPS * 8 * TGQ_WSTACK * sizeof(StackEntry)) + (size_t)warp * TGQ_CHUNK;
  WidePool & pw = ((WidePool *)(smem + (size_t)TGQ_WARPS * 8 * TGQ_WSTACK * sizeof(StackEntry) + (size_t)TGQ_WARPS * TGQ_CHUNK * sizeof(f4)))[warp];
  int ovf_node[TGQ_WOVF]; float ovf_key[TGQ_WOVF];
  if (lane < 8) { pw.best[lane] = 0ull; pw.done[lane] = 0; }
  if (lane == 0) pw.tail = 0;
  __syncwarp();
  // query state, identical in the four lanes of a group
  v3 p = mk3(0.0f, 0.0f, 0.0f), best_cp = p; float mq = 0.0f, best_d2 = 0.0f; int best_idx = -1, best_k = -1, skip = -1;
  int cur = TGQ_ST_IDLE, sp = 0, leaf_k = 0, leaf_end = 0;
  int qi = 0, seed = -1, pending = 0, head = 0; bool hold = false;
  int chunk_base = 0, chunk_next = 0, chunk_end = 0; bool exhausted = false;
#define TGQ_WIDE_ENTER(code)                                                                                              \
  { int code_ = (code);                                                                                                   \
    if (code_ >= 0) cur = code_;                                                                                          \
    else { int c_ = ~code_; leaf_k = c_ >> TGQ_LEAF_BITS; leaf_end = leaf_k + (c_ & TGQ_LEAF_MASK); cur = leaf_end > leaf_k ? TGQ_ST_LEAF : TGQ_ST_POP; } }
  for (;;)
  {
    // ---- finished queries write their results; idle groups take the next queries of the warp's chunk ----------------------------
    if (cur == TGQ_ST_DONE && pending == 0)
    {
      if (sub == 0) { if (idx_out) idx_out[qi] = best_idx; if (d2_out) d2_out[qi] = best_idx >= 0 ? best_d2 : -1.0f; }
      if (cp3 && sub < 3) cp3[3 * (size_t)qi + sub] = comp3(best_cp, sub);
      seed = perm ? best_k : -1;
      cur = TGQ_ST_IDLE;
    }
    unsigned idle = __ballot_sync(FULL, cur == TGQ_ST_IDLE);
    if (idle != 0u && !exhausted)
    {
      if (chunk_next >= chunk_end)
      {
        int base = 0;
        if (lane == 0) base = atomicAdd(counter, TGQ_CHUNK);
        base = __shfl_sync(FULL, base, 0);
        chunk_base = chunk_next = base; chunk_end = min(base + TGQ_CHUNK, n);
        if (base >= n) exhausted = true;
        else
        {
          __syncwarp();
          for (int j = lane; j < chunk_end - chunk_base; j += 32)
          {
            int t = chunk_base + j; int i = perm ? (int)perm[t] : t; v3 pp = ld3(q3, i);
            stage[j] = make_float4(pp.x, pp.y, pp.z, __int_as_float(i));
          }
          __syncwarp();
        }
      }
      if (!exhausted)
      {
        int t = chunk_next + (__popc(idle & ((1u << gbase) - 1u)) >> 2);
        if (cur == TGQ_ST_IDLE && t < chunk_end)
        {
          f4 r = stage[t - chunk_base]; qi = __float_as_int(r.w);
          p = mk3(r.x, r.y, r.z); mq = query_margin(p, scene_pad); skip = exclude_self ? (int)(self_offset + qi) : -1;
          best_d2 = gated_ub(gate, p, bound_ub); best_idx = -1; best_k = -1; best_cp = mk3(0.0f, 0.0f, 0.0f);
          sp = 0; cur = b.n_wide > 0 ? 0 : TGQ_ST_DONE; hold = false;
          if (seed >= 0 && seed < b.n_prims && cur == 0)
          {
            // the previous answer of this group is tested first; the root waits on the stack until that test is back
            if (sub == 0) { StackEntry e; e.node = 0; e.key = 0.0f; stk[0] = e; }
            sp = 1; leaf_k = seed; leaf_end = seed + 1; cur = TGQ_ST_LEAF; hold = true;
          }
        }
        chunk_next = min(chunk_next + (__popc(idle) >> 2), chunk_end);
      }
    }
    // ---- leaves reached go to the warp's queue ------------------------------------------------------------------------------------
    {
      bool at_leaf = cur == TGQ_ST_LEAF;
      int cnt = leaf_end - leaf_k, pos = 0;
      if (at_leaf && sub == 0) pos = atomicAdd(&pw.tail, cnt);
      pos = __shfl_sync(FULL, pos, gbase);
      if (at_leaf)
      {
        for (int j = sub; j < cnt; j += 4) pw.q[(pos + j) & (TGQ_QCAP - 1)] = ((unsigned)g << 27) | (unsigned)(leaf_k + j);
        pending += cnt; cur = TGQ_ST_POP;
      }
    }
    __syncwarp();
    // ---- leaf iterations: 32 queued (owner, triangle) pairs at a time, every lane on its own pair -----------------------------------
    for (;;)
    {
      int avail = *(volatile int *)&pw.tail - head;
      bool runnable = (cur >= 0 || cur == TGQ_ST_POP) && !hold;
      if (!(avail >= TGQ_WIDE_THRESH || (avail > 0 && !__any_sync(FULL, runnable)))) break;
      int m = min(avail, 32);
      unsigned long long published = pool_key(best_d2, best_idx < 0 ? 0 : best_idx);      // (ub, 0) when nothing was found yet: only d2 < ub wins
      if (sub == 0) pw.best[g] = published;
      bool valid = lane < m;
      unsigned e = valid ? pw.q[(head + lane) & (TGQ_QCAP - 1)] : ((unsigned)g << 27);
      int og = (int)(e >> 27), ol = og << 2, k = (int)(e & 0x7ffffffu);
      __syncwarp();
      v3 po = mk3(__shfl_sync(FULL, p.x, ol), __shfl_sync(FULL, p.y, ol), __shfl_sync(FULL, p.z, ol));
      float mqo = __shfl_sync(FULL, mq, ol), bd2 = __shfl_sync(FULL, best_d2, ol);
      int skipo = __shfl_sync(FULL, skip, ol);
      unsigned long long mykey = ~0ull; v3 cp = mk3(0.0f, 0.0f, 0.0f);
      if (valid)
      {
        f4 const * q = (f4 const *)(b.tris + k);
        f4 r0, r1, r2, r3; ldq2(q, r0, r1); ldq2(q + 2, r2, r3);
        v3 v0 = mk3(r0.x, r0.y, r0.z), v1 = mk3(r1.x, r1.y, r1.z), v2 = mk3(r2.x, r2.y, r2.z), nn = mk3(r3.x, r3.y, r3.z);
        v3 cen = mul3(add3(add3(v0, v1), v2), 0.33333334f);
        float sd = fmaxf(fabsf(fdot3(po, nn) - r2.w) - mqo, 0.0f);
        v3 pc = sub3(po, cen);
        v3 perp = sub3(pc, mul3(nn, fdot3(nn, pc)));
        float lat = fmaxf(sqrtf(fdot3(perp, perp)) - r3.w - mqo, 0.0f);
        int idx = f2i(r0.w);
        if ((sd * sd + lat * lat) * TGQ_LB_DEFLATE <= bd2 && idx != skipo)
        {
          f4 r4, r5, r6, r7; ldq2(q + 4, r4, r5); ldq2(q + 6, r6, r7);
          float d2 = tri_point_sqdist_pre(v0, v1, v2, nn, r2.w, f2i(r1.w), r4, r5, r6, r7, po, cp);
          if (d2 <= bd2) { mykey = pool_key(d2, idx); atomicMin(&pw.best[og], mykey); }     // only candidates that can win (NaN fails, as in the reference's d2 < best)
        }
        atomicAdd(&pw.done[og], 1);
      }
      __syncwarp();
      if (mykey != ~0ull && pw.best[og] == mykey) pw.cp[og] = make_float4(cp.x, cp.y, cp.z, __int_as_float(k));
      __syncwarp();
      unsigned long long now = pw.best[g];
      int dn = pw.done[g];
      if (now != published)
      {
        f4 c = pw.cp[g];
        best_d2 = __int_as_float((int)(unsigned)(now >> 32)); best_idx = (int)(unsigned)(now & 0xffffffffull);
        best_cp = mk3(c.x, c.y, c.z); best_k = __float_as_int(c.w);
      }
      pending -= dn;
      if (pending == 0) hold = false;
      head += m;
      __syncwarp();
      if (sub == 0) pw.done[g] = 0;
      __syncwarp();
    }
    // ---- node step for the groups at a node: lane i tests child i ---------------------------------------------------------------------
    bool at_node = cur >= 0 && !hold;
    bool want_pop = cur == TGQ_ST_POP && !hold;
    if (!__any_sync(FULL, at_node || want_pop))
    {
      bool busy = cur != TGQ_ST_IDLE;        // waiting for queued tests, or about to be finished at the top of the loop
      if (exhausted && !__any_sync(FULL, busy)) break;
      continue;
    }
    float lb = 0.0f, key = 0.0f; int link = TGQ_WIDE_EMPTY; bool pass = false;
    if (at_node)
    {
      f4 const * nd = (f4 const *)(b.wide + cur);
      f4 A0, A1, B0, B1; ldq2(nd + 2 * sub, A0, A1); ldq2(nd + 8 + 2 * sub, B0, B1);
      wide_child_bounds(A0, A1, B0, B1, p, mq, lb, key, link);
      pass = link != TGQ_WIDE_EMPTY && lb <= best_d2;
    }
    unsigned gm = (__ballot_sync(FULL, pass) >> gbase) & 15u;
    float k1 = __shfl_xor_sync(FULL, key, 1), k2 = __shfl_xor_sync(FULL, key, 2), k3 = __shfl_xor_sync(FULL, key, 3);
    int rank = 0;
    {
      int j1 = sub ^ 1, j2 = sub ^ 2, j3 = sub ^ 3;
      rank += ((gm >> j1) & 1u) && (k1 < key || (k1 == key && j1 < sub));
      rank += ((gm >> j2) & 1u) && (k2 < key || (k2 == key && j2 < sub));
      rank += ((gm >> j3) & 1u) && (k3 < key || (k3 == key && j3 < sub));
    }
    unsigned r0m = (__ballot_sync(FULL, pass && rank == 0) >> gbase) & 15u;
    int nearest = __shfl_sync(FULL, link, gbase + (r0m ? __ffs((int)r0m) - 1 : 0));
    if (at_node)
    {
      int npass = __popc(gm);
      if (npass == 0) cur = TGQ_ST_POP;
      else
      {
        if (sp + npass - 1 <= TGQ_WSTACK)
        {
          if (pass && rank > 0) { StackEntry e; e.node = link; e.key = lb; stk[sp + npass - 1 - rank] = e; }
          sp += npass - 1;
        }
        else
        {
          // rare: the shared part of the stack is full -- every lane keeps the overflow in its own local memory
          for (int r = npass - 1; r >= 1; --r)
          {
            unsigned rm = (__ballot_sync(0xfu << gbase, pass && rank == r) >> gbase) & 15u;
            int src = gbase + __ffs((int)rm) - 1;
            int ln = __shfl_sync(0xfu << gbase, link, src); float lk = __shfl_sync(0xfu << gbase, lb, src);
            if (sp < TGQ_WSTACK) { if (sub == 0) { StackEntry e; e.node = ln; e.key = lk; stk[sp] = e; } }
            else if (sp - TGQ_WSTACK < TGQ_WOVF) { ovf_node[sp - TGQ_WSTACK] = ln; ovf_key[sp - TGQ_WSTACK] = lk; }
            ++sp;
          }
        }
        TGQ_WIDE_ENTER(nearest)
      }
    }
    __syncwarp();
    // ---- pop phase ------------------------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < TGQ_POPS; ++j)
      if (cur == TGQ_ST_POP && !hold)
      {
        if (sp <= 0) cur = TGQ_ST_DONE;
        else
        {
          --sp;
          int code; float kk;
          if (sp < TGQ_WSTACK) { StackEntry e = stk[sp]; code = e.node; kk = e.key; }
          else { code = ovf_node[sp - TGQ_WSTACK]; kk = ovf_key[sp - TGQ_WSTACK]; }
          if (kk <= best_d2) TGQ_WIDE_ENTER(code)
        }
      }
  }
#undef TGQ_WIDE_ENTER
}

// ---- K6 k nearest neighbours: one bounded sorted set per thread in shared memory -------------------------------------
struct KnnW
{
  static constexpr int STAGE = 1;
  KnnState s; BvhView b; float const * q3; uint32_t const * perm; int k; float ub; int exclude_self; int64_t self_offset;
  int * idx; float * d2; int64_t qi;
  float const * qn3; int const * qlabel; float min_dot;      // per-query element filter, null = none
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 p = ld3(q3, i);
    *dst = make_float4(p.x, p.y, p.z, __int_as_float(i));
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 r = *src; qi = __float_as_int(r.w);
    knn_begin(s, b, mk3(r.x, r.y, r.z), exclude_self ? (int)(self_offset + qi) : -1, ub);
    if (qn3 || qlabel)
    {
      s.f.mode = (qn3 ? 1 : 0) | (qlabel ? 2 : 0); s.f.min_dot = min_dot;
      if (qn3) s.f.n = ld3(qn3, qi);
      if (qlabel) s.f.label = qlabel[qi];
    }
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, s.worst); }
  __device__ __forceinline__ void fetch(f4 * r) const { knn_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { knn_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { knn_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    for (int j = 0; j < k; ++j)
    {
      bool v = j < s.set.n;
      if (idx) idx[qi * k + j] = v ? s.set.idx[j * s.set.stride] : -1;
      if (d2) d2[qi * k + j] = v ? s.set.key[j * s.set.stride] : -1.0f;
    }
    s.t.cur = TGQ_ST_IDLE;
  }
};

__global__ void __launch_bounds__(TGQ_BLOCK)
k_knn(BvhView b, float const * __restrict__ q3, uint32_t const * __restrict__ perm, int n, int k, float bound_ub, int exclude_self,
      int64_t self_offset, int * counter, int * idx, float * d2, float const * __restrict__ qn3, int const * __restrict__ qlabel, float min_dot)
{
  extern __shared__ __align__(16) unsigned char smem[];
  KnnW w;
  w.b = b; w.q3 = q3; w.perm = perm; w.k = k; w.ub = bound_ub; w.exclude_self = exclude_self; w.self_offset = self_offset;
  w.idx = idx; w.d2 = d2; w.qi = 0; w.qn3 = qn3; w.qlabel = qlabel; w.min_dot = min_dot; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  unsigned char * extra; f4 * stage = trav_setup(w.s.t, st, smem, KnnW::STAGE, &extra);
  float * sets = (float *)extra;
  w.s.set.init(sets + threadIdx.x, (int *)(sets + (size_t)k * blockDim.x) + threadIdx.x, blockDim.x, k);
  warp_run(w, n, counter, stage);
}

// ---- K7 ball graph over the structure's own points (query t is the point at Morton position t) ---------------------------
template <bool SURF>
struct BallW
{
  static constexpr int STAGE = 2;       // (x, y, z, element index) (normal, -)
  BallState s; BvhView b; BvhView surf; int k; float init_radius; int rounds; float min_dot; int min_degree;
  int * idx; float * sep; int * count; int64_t qi;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    dst[0] = b.pts[t];
    dst[1] = b.nrm ? b.nrm[t] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 e = src[0], q = src[1]; qi = __float_as_int(e.w);
    ball_begin(s, b, mk3(e.x, e.y, e.z), mk3(q.x, q.y, q.z), b.nrm != nullptr, min_dot, (int)qi, init_radius, rounds, min_degree);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, s.r2); }
  __device__ __forceinline__ void fetch(f4 * r) const { ball_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { ball_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { if (SURF) ball_leaf_compute_surf(s, b, surf, r); else ball_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    if (!ball_round_done(s, b)) return;          // next round started
    for (int j = 0; j < k; ++j)
    {
      bool v = j < s.set.n;
      if (idx) idx[qi * k + j] = v ? s.set.idx[j * s.set.stride] : -1;
      if (sep) sep[qi * k + j] = v ? s.set.key[j * s.set.stride] : -1.0f;
    }
    if (count) count[qi] = s.set.n;
    s.t.cur = TGQ_ST_IDLE;
  }
};

template <bool SURF>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_ball_graph(BvhView b, BvhView surf, int n, int k, float init_radius, int rounds, float min_dot, int min_degree, int * counter, int * idx, float * sep,
             int * count)
{
  extern __shared__ __align__(16) unsigned char smem[];
  BallW<SURF> w;
  w.b = b; w.surf = surf; w.k = k; w.init_radius = init_radius; w.rounds = rounds; w.min_dot = min_dot; w.min_degree = min_degree;
  w.idx = idx; w.sep = sep; w.count = count; w.qi = 0; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  unsigned char * extra; f4 * stage = trav_setup(w.s.t, st, smem, BallW<SURF>::STAGE, &extra);
  float * sets = (float *)extra;
  w.s.set.init(sets + threadIdx.x, (int *)(sets + (size_t)k * blockDim.x) + threadIdx.x, blockDim.x, k);
  warp_run(w, n, counter, stage);
}

// ---- K6b / K7b k-NN and the sample graph on candidate buffers, compacted by the whole warp ------------------------------------------
// (CandBuf, tgq_traverse.cuh.)  The scheduler is warp_run with two warp-cooperative hooks: after every step the lanes whose buffer is
// nearly full are served one after the other, and a finished query is compacted the same way and written out by the warp (k
// consecutive outputs: coalesced).  Compacting one lane's row (n <= cap entries, keep the k smallest in ascending order):
//   cap <= 32 (k <= 16, EPL = 1): every lane takes one entry and the warp runs a 32-wide bitonic network on the packed 64-bit
//     (key, index) values -- 15 compare-exchange stages of two shuffles each, about 130 instructions at full lane utilisation,
//     against an insertion loop per candidate at 5 of 32 lanes in the sorted-set kernels;
//   larger buffers (EPL = 2, 3): rank sort -- every lane holds EPL entries in registers and counts, over broadcast reads of the
//     row, how many entries are smaller (the first version used this for every size: ncu's source view showed the rank loop at
//     34 % of all warp instructions for k = 16, profiles/r02_t_knn_buf_source_lines.txt).
// The buffer kernels are short of resident warps (shared memory: 256 bytes of buffer per lane for k = 16), so they keep only the first
// TGQ_BUF_SSTACK stack entries per lane in shared memory (deeper ones in local memory; ordered DFS over points rarely gets there).
#ifndef TGQ_BUF_SSTACK
#define TGQ_BUF_SSTACK 8
#endif
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long x, int m)
{
  unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)x, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(x >> 32), m);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long bitonic_sort32(unsigned long long x)
{
  int const lane = threadIdx.x & 31;
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1)
  {
    bool const up = (lane & size) == 0 || size == 32;
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1)
    {
      unsigned long long y = shfl_xor_u64(x, stride);
      bool const low = (lane & stride) == 0;
      bool const take_min = low == up;
      x = take_min ? (y < x ? y : x) : (y > x ? y : x);
    }
  }
  return x;
}

template <int EPL>
__device__ __forceinline__ void coop_compact(unsigned long long * rows, int cap, int k, int o, int & my_n, float & my_thr)
{
  unsigned const FULL = 0xffffffffu;
  int const lane = threadIdx.x & 31;
  int n_o = __shfl_sync(FULL, my_n, o);
  unsigned long long * row = rows + (size_t)o * cap;
  #define TGQ_PHYS(i) (((i) & ~15) | (((i) + o) & 15))
  if (EPL == 1)
  {
    unsigned long long x = lane < n_o ? row[TGQ_PHYS(lane)] : ~0ull;
    x = bitonic_sort32(x);
    int keep = min(n_o, k);
    if (lane < keep) row[TGQ_PHYS(lane)] = x;
    float kth = cand_key(__shfl_sync(FULL, x, max(k - 1, 0)));
    __syncwarp();
    if (lane == o) { my_n = keep; if (n_o >= k && k > 0) my_thr = kth; }
  }
  else
  {
    unsigned long long mine[EPL]; int rank[EPL];
#pragma unroll
    for (int t = 0; t < EPL; ++t) { int i = lane + 32 * t; mine[t] = i < n_o ? row[TGQ_PHYS(i)] : ~0ull; rank[t] = 0; }
    for (int i = 0; i < n_o; ++i)
    {
      unsigned long long x = row[TGQ_PHYS(i)];
#pragma unroll
      for (int t = 0; t < EPL; ++t) rank[t] += x < mine[t] ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < EPL; ++t) if (lane + 32 * t < n_o && rank[t] < k) row[TGQ_PHYS(rank[t])] = mine[t];
    __syncwarp();
    if (lane == o)
    {
      my_n = min(n_o, k);
      if (n_o >= k && k > 0) my_thr = cand_key(row[TGQ_PHYS(k - 1)]);
    }
  }
  #undef TGQ_PHYS
}

// Writes the (compacted) buffer of lane o: k outputs per query, entry j or (-1, -1.0f).
__device__ __forceinline__ void coop_write(unsigned long long const * rows, int cap, int k, int o, int n_o, int64_t qi, int * idx, float * key)
{
  int const lane = threadIdx.x & 31;
  unsigned long long const * row = rows + (size_t)o * cap;
  for (int j = lane; j < k; j += 32)
  {
    bool v = j < n_o;
    unsigned long long e = v ? row[((j) & ~15) | (((j) + o) & 15)] : 0ull;
    if (idx) idx[qi * k + j] = v ? cand_idx(e) : -1;
    if (key) key[qi * k + j] = v ? cand_key(e) : -1.0f;
  }
}

template <typename W>
__device__ __forceinline__ void warp_run_coop(W & w, int n, int * counter, f4 * stage)
{
  unsigned const FULL = 0xffffffffu;
  int lane = threadIdx.x & 31;
  int chunk_base = 0, chunk_next = 0, chunk_end = 0; bool exhausted = false;
  for (;;)
  {
    unsigned fin = __ballot_sync(FULL, w.state() == TGQ_ST_DONE);
    while (fin) { int o = __ffs(fin) - 1; fin &= fin - 1; w.coop_finish(o); }
    unsigned idle = __ballot_sync(FULL, w.state() == TGQ_ST_IDLE);
    if (idle != 0u && !exhausted)
    {
      if (chunk_next >= chunk_end)
      {
        int base = 0;
        if (lane == 0) base = atomicAdd(counter, TGQ_CHUNK);
        base = __shfl_sync(FULL, base, 0);
        chunk_base = chunk_next = base; chunk_end = min(base + TGQ_CHUNK, n);
        if (base >= n) exhausted = true;
        else
        {
          __syncwarp();
          for (int j = lane; j < chunk_end - chunk_base; j += 32) w.stage(chunk_base + j, stage + (size_t)j * W::STAGE);
          __syncwarp();
        }
      }
      if (!exhausted)
      {
        int t = chunk_next + __popc(idle & ((1u << lane) - 1u));
        if (w.state() == TGQ_ST_IDLE && t < chunk_end) w.begin(t, stage + (size_t)(t - chunk_base) * W::STAGE);
        chunk_next = min(chunk_next + __popc(idle), chunk_end);
      }
    }
    int st = w.state();
    bool at_node = st >= 0, at_leaf = st == TGQ_ST_LEAF;
    if (!__any_sync(FULL, at_node || at_leaf || st == TGQ_ST_POP))
    {
      if (exhausted && __all_sync(FULL, w.state() != TGQ_ST_DONE)) break;
      continue;
    }
    f4 r[W::FETCH];
    if (at_node || at_leaf) w.fetch(r);
    if (at_node) w.node_compute(r);
    else if (at_leaf) w.leaf_compute(r);
    unsigned cm = __ballot_sync(FULL, w.wants_compact());
    while (cm) { int o = __ffs(cm) - 1; cm &= cm - 1; w.coop_compact_lane(o); }
#pragma unroll
    for (int j = 0; j < TGQ_POPS; ++j) if (w.state() == TGQ_ST_POP) w.pop_step();
  }
}

template <int EPL>
struct KnnBufW
{
  static constexpr int STAGE = 1;
  KnnBufState s; BvhView b; float const * q3; uint32_t const * perm; int k; float ub; int exclude_self; int64_t self_offset;
  int * idx; float * d2; int64_t qi; unsigned long long * rows;
  float const * qn3; int const * qlabel; float min_dot;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 p = ld3(q3, i);
    *dst = make_float4(p.x, p.y, p.z, __int_as_float(i));
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 r = *src; qi = __float_as_int(r.w);
    knnb_begin(s, b, mk3(r.x, r.y, r.z), exclude_self ? (int)(self_offset + qi) : -1, ub);
    if (qn3 || qlabel)
    {
      s.f.mode = (qn3 ? 1 : 0) | (qlabel ? 2 : 0); s.f.min_dot = min_dot;
      if (qn3) s.f.n = ld3(qn3, qi);
      if (qlabel) s.f.label = qlabel[qi];
    }
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, s.worst); }
  __device__ __forceinline__ void fetch(f4 * r) const { box_fetch(s.t, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { knnb_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { knnb_leaf_compute(s, b, r); }
  __device__ __forceinline__ bool wants_compact() const { return s.buf.wants_compact(); }
  __device__ __forceinline__ void coop_compact_lane(int o) { coop_compact<EPL>(rows, s.buf.cap, k, o, s.buf.n, s.worst); }
  __device__ __forceinline__ void coop_finish(int o)
  {
    float unused = 0.0f;
    coop_compact<EPL>(rows, s.buf.cap, k, o, s.buf.n, unused);
    int n_o = __shfl_sync(0xffffffffu, s.buf.n, o);
    int64_t q_o = __shfl_sync(0xffffffffu, qi, o);
    coop_write(rows, s.buf.cap, k, o, n_o, q_o, idx, d2);
    __syncwarp();
    if ((threadIdx.x & 31) == o) s.t.cur = TGQ_ST_IDLE;
  }
};

template <int EPL>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_knn_buf(BvhView b, float const * __restrict__ q3, uint32_t const * __restrict__ perm, int n, int k, float bound_ub, int exclude_self,
          int64_t self_offset, int * counter, int * idx, float * d2, float const * __restrict__ qn3, int const * __restrict__ qlabel, float min_dot)
{
  extern __shared__ __align__(16) unsigned char smem[];
  KnnBufW<EPL> w;
  w.b = b; w.q3 = q3; w.perm = perm; w.k = k; w.ub = bound_ub; w.exclude_self = exclude_self; w.self_offset = self_offset;
  w.idx = idx; w.d2 = d2; w.qi = 0; w.qn3 = qn3; w.qlabel = qlabel; w.min_dot = min_dot; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  unsigned char * extra; f4 * stage = trav_setup(w.s.t, st, smem, KnnBufW<EPL>::STAGE, &extra, TGQ_BUF_SSTACK);
  int cap = cand_cap(k);
  w.rows = (unsigned long long *)extra + (size_t)(threadIdx.x >> 5) * 32 * cap;
  w.s.buf.init(w.rows + (size_t)(threadIdx.x & 31) * cap, threadIdx.x & 31, k);
  warp_run_coop(w, n, counter, stage);
}

template <int EPL>
struct BallBufW
{
  static constexpr int STAGE = 2;
  BallBufState s; BvhView b; int k; float init_radius; int rounds; float min_dot; int min_degree;
  int * idx; float * sep; int * count; int64_t qi; unsigned long long * rows;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    dst[0] = b.pts[t];
    dst[1] = b.nrm ? b.nrm[t] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 e = src[0], q = src[1]; qi = __float_as_int(e.w);
    ballb_begin(s, b, mk3(e.x, e.y, e.z), mk3(q.x, q.y, q.z), b.nrm != nullptr, min_dot, (int)qi, init_radius, rounds, min_degree);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, s.r2); }
  __device__ __forceinline__ void fetch(f4 * r) const { box_fetch(s.t, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { ballb_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { ballb_leaf_compute(s, b, r); }
  __device__ __forceinline__ bool wants_compact() const { return s.buf.wants_compact(); }
  __device__ __forceinline__ void coop_compact_lane(int o) { coop_compact<EPL>(rows, s.buf.cap, k, o, s.buf.n, s.worst_key); }
  __device__ __forceinline__ void coop_finish(int o)
  {
    unsigned const FULL = 0xffffffffu;
    float unused = 0.0f;
    coop_compact<EPL>(rows, s.buf.cap, k, o, s.buf.n, unused);
    int complete = 0;
    if ((threadIdx.x & 31) == o) complete = ballb_round_done(s, b) ? 1 : 0;          // otherwise the next round has started
    complete = __shfl_sync(FULL, complete, o);
    if (!complete) return;
    int n_o = __shfl_sync(FULL, s.buf.n, o);
    int64_t q_o = __shfl_sync(FULL, qi, o);
    coop_write(rows, s.buf.cap, k, o, n_o, q_o, idx, sep);
    __syncwarp();
    if ((threadIdx.x & 31) == o) { if (count) count[qi] = s.buf.n; s.t.cur = TGQ_ST_IDLE; }
  }
};

template <int EPL>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_ball_graph_buf(BvhView b, int n, int k, float init_radius, int rounds, float min_dot, int min_degree, int * counter, int * idx, float * sep, int * count)
{
  extern __shared__ __align__(16) unsigned char smem[];
  BallBufW<EPL> w;
  w.b = b; w.k = k; w.init_radius = init_radius; w.rounds = rounds; w.min_dot = min_dot; w.min_degree = min_degree;
  w.idx = idx; w.sep = sep; w.count = count; w.qi = 0; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  unsigned char * extra; f4 * stage = trav_setup(w.s.t, st, smem, BallBufW<EPL>::STAGE, &extra, TGQ_BUF_SSTACK);
  int cap = cand_cap(k);
  w.rows = (unsigned long long *)extra + (size_t)(threadIdx.x >> 5) * 32 * cap;
  w.s.buf.init(w.rows + (size_t)(threadIdx.x & 31) * cap, threadIdx.x & 31, k);
  warp_run_coop(w, n, counter, stage);
}

// ---- ball range queries on a point structure (rangeQueryIndices with a Ball3) ---------------------------------------------
struct RangeW
{
  static constexpr int STAGE = 1;       // (centre, query index)
  RangeState s; BvhView b; float const * c3; float const * rad; uint32_t const * perm; int k; int exclude_self; int64_t self_offset;
  int * idx; int * count; int64_t qi;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 p = ld3(c3, i);
    *dst = make_float4(p.x, p.y, p.z, __int_as_float(i));
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 r = *src; qi = __float_as_int(r.w);
    range_begin(s, b, mk3(r.x, r.y, r.z), rad[qi], exclude_self ? (int)(self_offset + qi) : -1);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, s.r2); }
  __device__ __forceinline__ void fetch(f4 * r) const { range_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { range_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { range_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    for (int j = 0; j < k; ++j) if (idx) idx[qi * k + j] = j < s.set.n ? s.set.idx[j * s.set.stride] : -1;
    if (count) count[qi] = s.count;
    s.t.cur = TGQ_ST_IDLE;
  }
};

__global__ void __launch_bounds__(TGQ_BLOCK)
k_ball_range(BvhView b, float const * __restrict__ c3, float const * __restrict__ rad, uint32_t const * __restrict__ perm, int n, int k,
             int exclude_self, int64_t self_offset, int * counter, int * idx, int * count)
{
  extern __shared__ __align__(16) unsigned char smem[];
  RangeW w;
  w.b = b; w.c3 = c3; w.rad = rad; w.perm = perm; w.k = k; w.exclude_self = exclude_self; w.self_offset = self_offset; w.idx = idx; w.count = count;
  w.qi = 0; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  unsigned char * extra; f4 * stage = trav_setup(w.s.t, st, smem, RangeW::STAGE, &extra);
  float * sets = (float *)extra;
  w.s.set.init(sets + threadIdx.x, (int *)(sets + (size_t)k * blockDim.x) + threadIdx.x, blockDim.x, k);
  warp_run(w, n, counter, stage);
}

// ---- unbounded range queries on a point structure (ball or axis-aligned box ranges) ------------------------------------------
template <bool BOX, bool FILL>
struct RangeAllW
{
  static constexpr int STAGE = 2;       // (a, query index) (b, -)
  RangeAllState<BOX> s; BvhView b; float const * a3; float const * bq; uint32_t const * perm; long long const * offsets; int * idx; int * count; int64_t qi;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 a = ld3(a3, i);
    dst[0] = make_float4(a.x, a.y, a.z, __int_as_float(i));
    if (BOX) { v3 h = ld3(bq, i); dst[1] = make_float4(h.x, h.y, h.z, 0.0f); }
    else { float r = bq[i]; dst[1] = make_float4(r * r, r, 0.0f, 0.0f); }       // BallN::contains: r * r in fp32
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 a = src[0], h = src[1]; qi = __float_as_int(a.w);
    bool valid = BOX ? true : h.y >= 0.0f;
    range_all_begin(s, b, mk3(a.x, a.y, a.z), mk3(h.x, h.y, h.z), FILL ? idx + offsets[qi] : nullptr, valid);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, 0.5f); }
  __device__ __forceinline__ void fetch(f4 * r) const { box_fetch(s.t, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { range_all_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { range_all_leaf_compute<BOX, FILL>(s, b, r); }
  __device__ __forceinline__ void finish() { if (!FILL) count[qi] = s.count; s.t.cur = TGQ_ST_IDLE; }
};

template <bool BOX, bool FILL>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_range_all(BvhView b, float const * __restrict__ a3, float const * __restrict__ bq, uint32_t const * __restrict__ perm, int n, int * counter,
            long long const * __restrict__ offsets, int * idx, int * count)
{
  extern __shared__ __align__(16) unsigned char smem[];
  RangeAllW<BOX, FILL> w;
  w.b = b; w.a3 = a3; w.bq = bq; w.perm = perm; w.offsets = offsets; w.idx = idx; w.count = count; w.qi = 0; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  f4 * stage = trav_setup(w.s.t, st, smem, RangeAllW<BOX, FILL>::STAGE, nullptr);
  warp_run(w, n, counter, stage);
}

// ---- ball neighbourhood statistics on a point structure (processRange with the local-feature functors) ----------------------
struct StatsW
{
  static constexpr int STAGE = 1;       // (centre, query index)
  StatsState s; BvhView b; float const * c3; float const * qn3; float const * rad; uint32_t const * perm; double * out; int64_t qi;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 p = ld3(c3, i);
    *dst = make_float4(p.x, p.y, p.z, __int_as_float(i));
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 r = *src; qi = __float_as_int(r.w);
    stats_begin(s, b, mk3(r.x, r.y, r.z), rad[qi], qn3 ? ld3(qn3, (int)qi) : mk3(0.0f, 0.0f, 0.0f), qn3 != nullptr);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, s.r2); }
  __device__ __forceinline__ void fetch(f4 * r) const { stats_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { stats_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { stats_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    double2 * o = (double2 *)(out + 16 * qi);
    for (int j = 0; j < 7; ++j) o[j] = make_double2(s.acc[2 * j], s.acc[2 * j + 1]);
    o[7] = make_double2(s.acc[14], 0.0);
    s.t.cur = TGQ_ST_IDLE;
  }
};

__global__ void __launch_bounds__(TGQ_BLOCK)
k_ball_stats(BvhView b, float const * __restrict__ c3, float const * __restrict__ qn3, float const * __restrict__ rad,
             uint32_t const * __restrict__ perm, int n, int * counter, double * out)
{
  extern __shared__ __align__(16) unsigned char smem[];
  StatsW w;
  w.b = b; w.c3 = c3; w.qn3 = qn3; w.rad = rad; w.perm = perm; w.out = out; w.qi = 0; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  f4 * stage = trav_setup(w.s.t, st, smem, StatsW::STAGE, nullptr);
  warp_run(w, n, counter, stage);
}

// Local distance histograms: one lane per ball, its bins in shared memory (bank-conflict-free strided layout, like the k-NN sets).
struct HistW
{
  static constexpr int STAGE = 1;       // (centre, query index)
  HistState s; BvhView b; float const * c3; float const * rad; uint32_t const * perm; unsigned * out; int64_t qi;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 p = ld3(c3, i);
    *dst = make_float4(p.x, p.y, p.z, __int_as_float(i));
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 r = *src; qi = __float_as_int(r.w);
    hist_begin(s, b, mk3(r.x, r.y, r.z), rad[qi]);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, s.r2); }
  __device__ __forceinline__ void fetch(f4 * r) const { hist_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { hist_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { hist_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    for (int j = 0; j < s.nbins; ++j) out[qi * s.nbins + j] = s.bins[j * s.stride];
    s.t.cur = TGQ_ST_IDLE;
  }
};

__global__ void __launch_bounds__(TGQ_BLOCK)
k_ball_hist(BvhView b, float const * __restrict__ c3, float const * __restrict__ rad, uint32_t const * __restrict__ perm, int n, int nbins,
            int * counter, unsigned * out)
{
  extern __shared__ __align__(16) unsigned char smem[];
  HistW w;
  w.b = b; w.c3 = c3; w.rad = rad; w.perm = perm; w.out = out; w.qi = 0; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  unsigned char * extra; f4 * stage = trav_setup(w.s.t, st, smem, HistW::STAGE, &extra);
  w.s.nbins = nbins; w.s.stride = blockDim.x; w.s.bins = (unsigned *)extra + threadIdx.x;
  warp_run(w, n, counter, stage);
}

// ---- SpinImage (SurfaceFeatures/Local/SpinImage.cpp:47-75): every query bins ALL samples of the point structure ------------------
// No traversal: a streaming pass, bound by the fp32 pipes (one square root and two IEEE divisions per pair), not by memory.  One WARP
// per query: its lanes take 32 samples at a time from a shared-memory tile the block's warps (= queries) share, and the warp keeps ONE
// histogram (bins x 4 bytes) in shared memory -- lanes that hit the same bin are found with a match and their leader adds the count,
// so no atomics are needed.  The first version gave every THREAD a query and a private histogram (900 bytes per thread at 15 x 15 bins):
// six resident warps per SM, latency bound at 94 G pairs/s; a histogram per warp lets the SM fill up.  grid = (query groups, sample
// slices); slices add their partial counts to the output with integer atomics, so the result does not depend on the launch shape.
#define TGQ_SPIN_TILE 1024
#define TGQ_SPIN_WARPS 8
__global__ void __launch_bounds__(32 * TGQ_SPIN_WARPS)
k_spin_image(f4 const * __restrict__ pts, int P, float const * __restrict__ q3, float const * __restrict__ qn3, int Q, int nr, int nh,
             float max_radius, float max_height, int slice_len, unsigned * out)
{
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned const FULL = 0xffffffffu;
  int const tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, bins = nr * nh;
  f4 * tile = (f4 *)smem; unsigned * hist = (unsigned *)(tile + TGQ_SPIN_TILE) + (size_t)wid * bins;
  int q = blockIdx.x * TGQ_SPIN_WARPS + wid; bool live = q < Q;
  v3 pos = mk3(0.0f, 0.0f, 0.0f), axis = mk3(0.0f, 0.0f, 1.0f);
  if (live) { pos = ld3(q3, q); axis = stable_normalized3(ld3(qn3, q)); }
  for (int b = lane; b < bins; b += 32) hist[b] = 0u;
  int s0 = blockIdx.y * slice_len, s1 = min(P, s0 + slice_len);
  for (int base = s0; base < s1; base += TGQ_SPIN_TILE)
  {
    int m = min(TGQ_SPIN_TILE, s1 - base);
    __syncthreads();
    for (int j = tid; j < m; j += 32 * TGQ_SPIN_WARPS) tile[j] = pts[base + j];
    __syncthreads();
    if (live)
      for (int j0 = 0; j0 < m; j0 += 32)
      {
        int j = j0 + lane; bool have = j < m;
        int bin = -1 - lane;                                     // lanes without a sample match nobody
        if (have)
        {
          f4 e = tile[j]; int rb, hb;
          spin_bins(mk3(e.x, e.y, e.z), pos, axis, nr, nh, max_radius, max_height, rb, hb);
          bin = rb * nh + hb;
        }
        unsigned same = __match_any_sync(FULL, bin);
        if (have && lane == __ffs(same) - 1) hist[bin] += (unsigned)__popc(same);
        __syncwarp();
      }
  }
  __syncwarp();
  if (live)
    for (int b = lane; b < bins; b += 32)
    {
      unsigned c = hist[b];
      if (c) atomicAdd(out + (size_t)q * bins + b, c);
    }
}

// Visibility rays (SurfaceFeatures/Local/Visibility.hpp:104-125): thread = (sample, direction).  Ray (p + 0.001 * scale * u, u),
// unbounded, any hit; escaped[t] = 1 when nothing was hit.  Directions: D x 3 shared by every sample, or S x D x 3 per sample.
struct VisW
{
  static constexpr int STAGE = 2;       // (origin, ray number) (direction, -)
  RayState<true> s; BvhView b; float const * p3; float const * dirs; int D; int per_sample; float scale; uint8_t * escaped; int64_t r;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int sm = t / D, k = t - sm * D;
    v3 u = ld3(dirs, per_sample ? t : k), p = ld3(p3, sm);
    v3 o = add3(p, mul3(u, 0.001f * scale));        // position + (0.001f * scale) * u
    dst[0] = make_float4(o.x, o.y, o.z, __int_as_float(t));
    dst[1] = make_float4(u.x, u.y, u.z, 0.0f);
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 o = src[0], d = src[1]; r = __float_as_int(o.w);
    ray_begin(s, b, mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z), -1.0f);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, ray_limit(s)); }
  __device__ __forceinline__ void fetch(f4 * r_) const { ray_fetch(s, b, r_); }
  __device__ __forceinline__ void node_compute(f4 const * r_) { ray_node_compute(s, b, r_); }
  __device__ __forceinline__ void leaf_compute(f4 const * r_) { ray_leaf_compute(s, b, r_); }
  __device__ __forceinline__ void finish() { escaped[r] = s.best_idx >= 0 ? 0 : 1; s.t.cur = TGQ_ST_IDLE; }
};

__global__ void __launch_bounds__(TGQ_BLOCK)
k_vis_rays(BvhView b, float const * __restrict__ p3, float const * __restrict__ dirs, int D, int per_sample, int nrays, float scale,
           int * counter, uint8_t * escaped)
{
  extern __shared__ __align__(16) unsigned char smem[];
  VisW w;
  w.b = b; w.p3 = p3; w.dirs = dirs; w.D = D; w.per_sample = per_sample; w.scale = scale; w.escaped = escaped; w.r = 0;
  w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  f4 * stage = trav_setup(w.s.t, st, smem, VisW::STAGE, nullptr);
  warp_run(w, nrays, counter, stage);
}

// (double)num_escaped / num_rays per sample (Visibility.hpp:122); one warp per sample.
__global__ void k_vis_count(uint8_t const * __restrict__ escaped, int S, int D, double * vis, int * n_escaped)
{
  int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= S) return;
  int c = 0;
  for (int k = lane; k < D; k += 32) c += escaped[(size_t)warp * D + k];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) { if (vis) vis[warp] = (double)c / (double)D; if (n_escaped) n_escaped[warp] = c; }
}

// ---- K8 ray first hit / any hit ------------------------------------------------------------------------------------------
template <bool ANY>
struct RayW
{
  static constexpr int STAGE = 2;       // (origin, query index) (direction, max time)
  RayState<ANY> s; BvhView b; float const * o3; float const * d3; float const * tmax; uint32_t const * perm;
  float * t_out; int * idx; float * normal3; uint8_t * hit; int qi; int slot; PeerOut peers;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 o = ld3(o3, i), d = ld3(d3, i);
    dst[0] = make_float4(o.x, o.y, o.z, __int_as_float(i));
    dst[1] = make_float4(d.x, d.y, d.z, tmax ? tmax[i] : -1.0f);
  }
  __device__ __forceinline__ void begin(int t, f4 const * src)
  {
    f4 o = src[0], d = src[1]; qi = __float_as_int(o.w); slot = t;
    ray_begin(s, b, mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z), d.w);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, ray_limit(s)); }
  __device__ __forceinline__ void fetch(f4 * r) const { ray_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { ray_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { ray_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    if (peers.n && peers.compact)
      ((float4 *)peers.base[0])[slot] = make_float4(s.best_idx >= 0 ? s.best_t : -1.0f, __int_as_float(s.best_idx), __int_as_float(qi), 0.0f);
    else
    {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < peers.n) st_rec8(peers.base[j], qi, s.best_idx >= 0 ? s.best_t : -1.0f, __int_as_float(s.best_idx));
    }
    if (t_out) t_out[qi] = s.best_idx >= 0 ? s.best_t : -1.0f;
    if (idx) idx[qi] = s.best_idx;
    if (hit) hit[qi] = s.best_idx >= 0 ? 1 : 0;
    if (normal3)
    {
      v3 nn = mk3(0.0f, 0.0f, 0.0f);
      if (s.best_k >= 0) { f4 q = ((f4 const *)(b.tris + s.best_k))[3]; nn = mk3(q.x, q.y, q.z); }
      st3(normal3, qi, nn);
    }
    s.t.cur = TGQ_ST_IDLE;
  }
};

template <bool ANY>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_ray(BvhView b, float const * __restrict__ o3, float const * __restrict__ d3, float const * __restrict__ tmax,
      uint32_t const * __restrict__ perm, int n, int * counter, float * t_out, int * idx, float * normal3, uint8_t * hit, PeerOut peers)
{
  extern __shared__ __align__(16) unsigned char smem[];
  RayW<ANY> w;
  w.b = b; w.o3 = o3; w.d3 = d3; w.tmax = tmax; w.perm = perm; w.t_out = t_out; w.idx = idx; w.normal3 = normal3; w.hit = hit; w.qi = 0; w.slot = 0; w.peers = peers;
  w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  f4 * stage = trav_setup(w.s.t, st, smem, RayW<ANY>::STAGE, nullptr);
  warp_run(w, n, counter, stage);
}

// ---- closest element to a line-segment query object (SegState in tgq_traverse.cuh) ------------------------
template <bool TRI>
struct SegW
{
  static constexpr int STAGE = 2;       // (endpoint 0, query index) (endpoint 1 or direction, -)
  SegState<TRI> s; BvhView b; float scene_pad; float const * a3; float const * b3; uint32_t const * perm; float ub; BoundGate gate;
  int const * excl; int n_excl; int dir_form;
  int * idx; float * d2; float * ep3; float * sp3; int qi;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 a = ld3(a3, i), e = ld3(b3, i);
    dst[0] = make_float4(a.x, a.y, a.z, __int_as_float(i));
    dst[1] = make_float4(e.x, e.y, e.z, 0.0f);
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 a = src[0], e = src[1]; qi = __float_as_int(a.w);
    v3 p0 = mk3(a.x, a.y, a.z), p1 = mk3(e.x, e.y, e.z);
    seg_begin(s, b, scene_pad, p0, dir_form ? p1 : sub3(p1, p0), ub);
    s.best_d2 = gated_ub_box(gate, s.blo, s.bhi, ub);
    if (excl)
    {
#pragma unroll
      for (int j = 0; j < TGQ_SEG_EXCL; ++j) if (j < n_excl) s.excl[j] = excl[(int64_t)qi * n_excl + j];
    }
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, s.best_d2); }
  __device__ __forceinline__ void fetch(f4 * r) const { seg_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { seg_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { seg_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    if (idx) idx[qi] = s.best_idx;
    if (d2) d2[qi] = s.best_idx >= 0 ? s.best_d2 : -1.0f;
    if (ep3) st3(ep3, qi, s.best_ep);
    if (sp3) st3(sp3, qi, s.best_sp);
    s.t.cur = TGQ_ST_IDLE;
  }
};

template <bool TRI>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_closest_seg(BvhView b, float scene_pad, float const * __restrict__ a3, float const * __restrict__ b3, uint32_t const * __restrict__ perm, int n,
              int dir_form, float bound_ub, BoundGate gate, int const * __restrict__ excl, int n_excl, int * counter, int * idx, float * d2, float * ep3,
              float * sp3)
{
  extern __shared__ __align__(16) unsigned char smem[];
  SegW<TRI> w;
  w.b = b; w.scene_pad = scene_pad; w.a3 = a3; w.b3 = b3; w.perm = perm; w.ub = bound_ub; w.gate = gate; w.excl = excl; w.n_excl = n_excl; w.dir_form = dir_form;
  w.idx = idx; w.d2 = d2; w.ep3 = ep3; w.sp3 = sp3; w.qi = 0; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  f4 * stage = trav_setup(w.s.t, st, smem, SegW<TRI>::STAGE, nullptr);
  warp_run(w, n, counter, stage);
}

// ---- closest element to a triangle query object (TriQState in tgq_traverse.cuh) -----------------------------
// Smallest positive |N| = |(v1 - v0) x (v2 - v0)| over the structure's triangles, as int bits (positive floats order like ints): the
// coplanarity tolerance of Moeller's test is FLT_EPSILON / |N| (triq_begin).
__global__ void k_min_normal(TriRec const * __restrict__ tris, int n, int * out_bits)
{
  float m = INFINITY;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
  {
    f4 const * q = (f4 const *)(tris + k);
    f4 a = q[0], b = q[1], c = q[2];
    v3 v0 = mk3(a.x, a.y, a.z);
    v3 N = cross3(sub3(mk3(b.x, b.y, b.z), v0), sub3(mk3(c.x, c.y, c.z), v0));
    float nn = sqrtf(fdot3(N, N));
    if (nn > 0.0f) m = fminf(m, nn);
  }
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMin(out_bits, __float_as_int(m));
}

template <bool TRI>
struct TriQW
{
  static constexpr int STAGE = 3;       // (v0, query index) (v1, -) (v2, -)
  TriQState<TRI> s; BvhView b; float scene_pad, elem_tau; float const * q9; uint32_t const * perm; float ub; BoundGate gate;
  int * idx; float * d2; float * ep3; float * qp3; int qi;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int i = perm ? (int)perm[t] : t; v3 v0 = ld3(q9, 3 * (int64_t)i), v1 = ld3(q9, 3 * (int64_t)i + 1), v2 = ld3(q9, 3 * (int64_t)i + 2);
    dst[0] = make_float4(v0.x, v0.y, v0.z, __int_as_float(i));
    dst[1] = make_float4(v1.x, v1.y, v1.z, 0.0f);
    dst[2] = make_float4(v2.x, v2.y, v2.z, 0.0f);
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 a = src[0], e = src[1], c = src[2]; qi = __float_as_int(a.w);
    triq_begin(s, b, scene_pad, elem_tau, mk3(a.x, a.y, a.z), mk3(e.x, e.y, e.z), mk3(c.x, c.y, c.z), ub);
    s.best_d2 = gated_ub_box(gate, s.blo, s.bhi, ub);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, triq_limit(s)); }
  __device__ __forceinline__ void fetch(f4 * r) const { triq_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { triq_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { triq_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    if (idx) idx[qi] = s.best_idx;
    if (d2) d2[qi] = s.best_idx >= 0 ? s.best_d2 : -1.0f;
    if (ep3) st3(ep3, qi, s.best_ep);
    if (qp3) st3(qp3, qi, s.best_qp);
    s.t.cur = TGQ_ST_IDLE;
  }
};

template <bool TRI>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_closest_tri(BvhView b, float scene_pad, float elem_tau, float const * __restrict__ q9, uint32_t const * __restrict__ perm, int n, float bound_ub,
              BoundGate gate, int * counter, int * idx, float * d2, float * ep3, float * qp3)
{
  extern __shared__ __align__(16) unsigned char smem[];
  TriQW<TRI> w;
  w.b = b; w.scene_pad = scene_pad; w.elem_tau = elem_tau; w.q9 = q9; w.perm = perm; w.ub = bound_ub; w.gate = gate;
  w.idx = idx; w.d2 = d2; w.ep3 = ep3; w.qp3 = qp3; w.qi = 0; w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  f4 * stage = trav_setup(w.s.t, st, smem, TriQW<TRI>::STAGE, nullptr);
  warp_run(w, n, counter, stage);
}

// SDF cone rays (ShapeDiameter.hpp:168-220): thread = (sample, ray).  valid = hit && (!interior || n_hit . dir >= 0).
__constant__ float c_cone_dirs[30][3] = {
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

struct SdfW
{
  static constexpr int STAGE = 2;       // (origin, ray number) (direction, -)
  RayState<false> s; BvhView b; float const * p3; float const * n3; float scale; int interior;
  float * t_out; int * idx; uint8_t * valid; int64_t r;
  __device__ __forceinline__ int state() const { return s.t.cur; }
  __device__ __forceinline__ void stage(int t, f4 * dst) const
  {
    int sm = t / 30, k = t - sm * 30;
    SdfFrame f = sdf_frame(ld3(p3, sm), ld3(n3, sm), scale);
    v3 d = sdf_dir(f, mk3(c_cone_dirs[k][0], c_cone_dirs[k][1], c_cone_dirs[k][2]));
    dst[0] = make_float4(f.origin.x, f.origin.y, f.origin.z, __int_as_float(t));
    dst[1] = make_float4(d.x, d.y, d.z, 0.0f);
  }
  __device__ __forceinline__ void begin(int, f4 const * src)
  {
    f4 o = src[0], d = src[1]; r = __float_as_int(o.w);
    ray_begin(s, b, mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z), -1.0f);
  }
  static constexpr int FETCH = TGQ_FETCH_BOX;
  __device__ __forceinline__ void pop_step() { trav_pop_step(s.t, ray_limit(s)); }
  __device__ __forceinline__ void fetch(f4 * r) const { ray_fetch(s, b, r); }
  __device__ __forceinline__ void node_compute(f4 const * r) { ray_node_compute(s, b, r); }
  __device__ __forceinline__ void leaf_compute(f4 const * r) { ray_leaf_compute(s, b, r); }
  __device__ __forceinline__ void finish()
  {
    bool ok = s.best_idx >= 0;
    if (ok && interior) { f4 q = ((f4 const *)(b.tris + s.best_k))[3]; ok = dot3(mk3(q.x, q.y, q.z), s.d) >= 0.0f; }
    if (t_out) t_out[r] = s.best_idx >= 0 ? s.best_t : -1.0f;
    if (idx) idx[r] = s.best_idx;
    valid[r] = ok ? 1 : 0;
    s.t.cur = TGQ_ST_IDLE;
  }
};

__global__ void __launch_bounds__(TGQ_BLOCK)
k_sdf_rays(BvhView b, float const * __restrict__ p3, float const * __restrict__ n3, int nrays, float scale, int interior, int * counter,
           float * t_out, int * idx, uint8_t * valid)
{
  extern __shared__ __align__(16) unsigned char smem[];
  SdfW w;
  w.b = b; w.p3 = p3; w.n3 = n3; w.scale = scale; w.interior = interior; w.t_out = t_out; w.idx = idx; w.valid = valid; w.r = 0;
  w.s.t.cur = TGQ_ST_IDLE; TravStack st;
  f4 * stage = trav_setup(w.s.t, st, smem, SdfW::STAGE, nullptr);
  warp_run(w, nrays, counter, stage);
}

// ---- K9 ICP sums: order-independent fixed-point accumulation (bit-identical for any launch shape / GPU count) ---------
// acc[0..2] = sum p, [3..5] = sum q, [6..14] = sum p q^T, [15] = sum |L p + t - q|^2 ; units in `unit` (pos, prod).
__global__ void k_icp_sums(float const * __restrict__ from3, float const * __restrict__ to3, int64_t n, float const * __restrict__ tr12,
                           double inv_unit_pos, double inv_unit_prod, long long * acc)
{
  long long a[16];
  for (int j = 0; j < 16; ++j) a[j] = 0;
  float L[12];
  for (int j = 0; j < 12; ++j) L[j] = tr12 ? tr12[j] : ((j % 5 == 0) ? 1.0f : 0.0f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
  {
    double p[3] = {from3[3 * i], from3[3 * i + 1], from3[3 * i + 2]}, q[3] = {to3[3 * i], to3[3 * i + 1], to3[3 * i + 2]};
    for (int r = 0; r < 3; ++r) { a[r] += __double2ll_rn(p[r] * inv_unit_pos); a[3 + r] += __double2ll_rn(q[r] * inv_unit_pos); }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[6 + 3 * r + c] += __double2ll_rn(p[r] * q[c] * inv_unit_prod);
    double e = 0;
    for (int r = 0; r < 3; ++r)
    { double v = (double)L[4 * r] * p[0] + (double)L[4 * r + 1] * p[1] + (double)L[4 * r + 2] * p[2] + (double)L[4 * r + 3] - q[r]; e += v * v; }
    a[15] += __double2ll_rn(e * inv_unit_prod);
  }
  for (int j = 0; j < 16; ++j)
  {
    long long v = a[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd((unsigned long long *)&acc[j], (unsigned long long)v);
  }
}

// ---- K9b ICP sums in the REFERENCE'S order (TGQ_OPT_ICP_REFERENCE_ORDER, the default) ---------------------------------------
// Icp3::alignOneStep / measureError (Algorithms/Icp3.hpp:306-367,418-463) are sequential fp32 loops over the correspondences:
//   pass 1   p_mean += p; q_mean += q; sum_weights += 1   (7 chains)   then   p_mean /= sum_weights; q_mean /= sum_weights
//   pass 2   cov(r, c) += (p - p_mean)[r] * (q - q_mean)[c]   (9 chains)
//            err = |tr p_mean - q_mean|^2;  err += |L (p - p_mean) - (q - q_mean)|^2   (1 chain)
// A chain is a dependent sequence of single fp32 additions, so its value is fixed by the order (and, over 10^5 terms, carries
// ~1e-4 of summation noise that an exact sum does not reproduce: after 50 ICP iterations on cfg3 the exact-sum transform differs
// from the reference's by 3e-4).  The chains are independent of each other: one thread block, whose threads stage tiles of
// TGQ_ICP_TILE points in shared memory (coalesced) and compute the per-point error terms, while lane j of warp 0 owns chain j and
// walks the tile in index order.  Bit for bit what the reference's loops give.
// out16: [0..2] p_mean, [3..5] q_mean, [6..14] cov row-major, [15] err  (all fp32 values, stored as float).
#define TGQ_ICP_TILE 1024
#define TGQ_ICP_THREADS 256
__global__ void __launch_bounds__(TGQ_ICP_THREADS)
k_icp_ref_order(float const * __restrict__ from3, float const * __restrict__ to3, int64_t n, float const * __restrict__ tr12, float * out16)
{
  int const tid = threadIdx.x, lane = tid & 31;
  bool const chain_warp = tid < 32;
  __shared__ float sp[3 * TGQ_ICP_TILE], sq[3 * TGQ_ICP_TILE], sterm[TGQ_ICP_TILE], smean[8];
  float L[12];
  for (int j = 0; j < 12; ++j) L[j] = tr12 ? tr12[j] : ((j % 5 == 0) ? 1.0f : 0.0f);
  float acc = 0.0f;
  // lane j < 3 walks p[j], 3 <= j < 6 walks q[j - 3], lane 6 counts
  float const * src1 = lane < 3 ? sp + lane : sq + (lane < 6 ? lane - 3 : 0);
  for (int64_t base = 0; base < n; base += TGQ_ICP_TILE)
  {
    int m = (int)min((int64_t)TGQ_ICP_TILE, n - base);
    for (int j = tid; j < 3 * m; j += TGQ_ICP_THREADS) { sp[j] = from3[3 * base + j]; sq[j] = to3[3 * base + j]; }
    __syncthreads();
    if (chain_warp && lane < 6)
    {
#pragma unroll 8
      for (int k = 0; k < m; ++k) acc = acc + src1[3 * k];
    }
    else if (chain_warp && lane == 6)
      for (int k = 0; k < m; ++k) acc = acc + 1.0f;
    __syncthreads();
  }
  if (chain_warp)
  {
    float sw = __shfl_sync(0xffffffffu, acc, 6);
    float mean = acc / sw;
    if (lane < 6) smean[lane] = mean;
  }
  __syncthreads();
  v3 pm = mk3(smean[0], smean[1], smean[2]), qm = mk3(smean[3], smean[4], smean[5]);
  // err starts from |tr p_mean - q_mean|^2 (AffineTransformN.hpp:160: linear * v + translation, Eigen's row . vector order)
  v3 tpm = mk3(sum3(L[0] * pm.x, L[1] * pm.y, L[2] * pm.z) + L[3], sum3(L[4] * pm.x, L[5] * pm.y, L[6] * pm.z) + L[7],
               sum3(L[8] * pm.x, L[9] * pm.y, L[10] * pm.z) + L[11]);
  float chain = (lane == 9) ? sqn3(sub3(tpm, qm)) : 0.0f;
  int const r = lane / 3, c = lane - 3 * (lane / 3);
  float const pmr = comp3(pm, r < 3 ? r : 0), qmc = comp3(qm, c);
  for (int64_t base = 0; base < n; base += TGQ_ICP_TILE)
  {
    int m = (int)min((int64_t)TGQ_ICP_TILE, n - base);
    for (int j = tid; j < 3 * m; j += TGQ_ICP_THREADS) { sp[j] = from3[3 * base + j]; sq[j] = to3[3 * base + j]; }
    __syncthreads();
    for (int k = tid; k < m; k += TGQ_ICP_THREADS)
    {
      v3 dp = sub3(mk3(sp[3 * k], sp[3 * k + 1], sp[3 * k + 2]), pm), dq = sub3(mk3(sq[3 * k], sq[3 * k + 1], sq[3 * k + 2]), qm);
      v3 ld = mk3(sum3(L[0] * dp.x, L[1] * dp.y, L[2] * dp.z), sum3(L[4] * dp.x, L[5] * dp.y, L[6] * dp.z), sum3(L[8] * dp.x, L[9] * dp.y, L[10] * dp.z));
      sterm[k] = sqn3(sub3(ld, dq));
    }
    __syncthreads();
    if (chain_warp && lane < 9)
    {
#pragma unroll 8
      for (int k = 0; k < m; ++k) { float dpr = sp[3 * k + r] - pmr, dqc = sq[3 * k + c] - qmc; chain = chain + dpr * dqc; }
    }
    else if (chain_warp && lane == 9)
    {
#pragma unroll 8
      for (int k = 0; k < m; ++k) chain = chain + sterm[k];
    }
    __syncthreads();
  }
  if (chain_warp)
  {
    if (lane < 6) out16[lane] = smean[lane];
    if (lane < 9) out16[6 + lane] = chain;
    if (lane == 9) out16[15] = chain;
  }
}

// out = L in + t, AffineTransformN.hpp:160 with Eigen's row . vector order.
__global__ void k_transform(float const * __restrict__ in3, int64_t n, float const * __restrict__ tr12, float * out3)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i >= n) return;
  v3 p = ld3(in3, i);
  st3(out3, i, mk3(sum3(tr12[0] * p.x, tr12[1] * p.y, tr12[2] * p.z) + tr12[3], sum3(tr12[4] * p.x, tr12[5] * p.y, tr12[6] * p.z) + tr12[7],
                   sum3(tr12[8] * p.x, tr12[9] * p.y, tr12[10] * p.z) + tr12[11]));
}

// ---- K11 area sampling ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64_at(uint64_t seed, uint64_t i)
{
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float u01_at(uint64_t seed, uint64_t i) { return (float)(splitmix64_at(seed, i) >> 40) * (1.0f / 16777216.0f); }

// cum[] = inclusive cumulative areas in element order (double); tri_of_elem[i] = Morton position of element i.
__global__ void k_area_sample(TriRec const * __restrict__ tris, int const * __restrict__ pos_of_elem, double const * __restrict__ cum,
                              int T, int64_t S, uint64_t seed, float * p3, float * n3, int * tri)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i >= S) return;
  double total = cum[T - 1];
  double t = (double)u01_at(seed, 3 * (uint64_t)i) * total;
  int lo = 0, hi = T;                       // std::lower_bound: first cum[j] >= t
  while (lo < hi) { int mid = (lo + hi) >> 1; if (cum[mid] < t) lo = mid + 1; else hi = mid; }
  int ti = lo < T - 1 ? lo : T - 1;
  TriRec r = tris[pos_of_elem[ti]];
  v3 v0 = mk3(r.q0.x, r.q0.y, r.q0.z), e01 = sub3(mk3(r.q1.x, r.q1.y, r.q1.z), v0), e02 = sub3(mk3(r.q2.x, r.q2.y, r.q2.z), v0);
  float s = u01_at(seed, 3 * (uint64_t)i + 1), u = u01_at(seed, 3 * (uint64_t)i + 2);
  if (s + u > 1.0f) { u = 1.0f - u; s = 1.0f - s; }
  if (p3) st3(p3, i, add3(add3(mul3(e01, s), mul3(e02, u)), v0));
  if (n3) st3(n3, i, mk3(r.q3.x, r.q3.y, r.q3.z));
  if (tri) tri[i] = ti;
}

// Sequential inclusive scan in double: cum[i] = cum[i-1] + a[i], the exact accumulation order of MeshSampler.hpp:166-168
// (a parallel scan would round differently and could move a lower_bound boundary).  One thread walks the chain; this is
// a one-off set-up step per structure (about T x 20 ns), not part of any query path.
__global__ void k_seq_scan(float const * __restrict__ a, int n, double * cum)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double acc = 0.0;
  for (int i = 0; i < n; ++i) { acc = (i == 0) ? (double)a[0] : acc + (double)a[i]; cum[i] = acc; }
}

// Indexed mesh -> triangle soup: out[3 * (3t + c) ..] = verts[tri[3t + c]] (MeshTriangles::addTriangles order).
__global__ void k_expand_indexed(float const * __restrict__ verts3, int const * __restrict__ tri3, int64_t ncorners, int64_t V, float * out9, int * bad)
{
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (c >= ncorners) return;
  int64_t v = tri3[c];
  if (v < 0 || v >= V) { *bad = 1; return; }       // reported as TGQ_EINVAL by the caller; nothing is built
  st3(out9, c, ld3(verts3, v));
}

// PointSet3.hpp:363-376: squared distance from seed sample seeds[j] to its nearest OTHER sample.
__global__ void k_self_nn(BvhView b, int const * __restrict__ pos_of_elem, int const * __restrict__ seeds, int n, float * d2_out)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x; if (j >= n) return;
  int i = seeds[j];
  f4 e = b.pts[pos_of_elem[i]];
  float best = INFINITY; int bi = -1; v3 cp = mk3(0.0f, 0.0f, 0.0f);
  closest_query<false>(b, 0.0f, mk3(e.x, e.y, e.z), i, best, bi, cp);
  d2_out[j] = bi >= 0 ? best : 0.0f;
}

// K10: SDF epilogue, one thread per sample (ShapeDiameter.hpp:208-257).
__global__ void k_sdf_stats(float const * __restrict__ ray_t, uint8_t const * __restrict__ valid, int S, float scale, double * sdf)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= S) return;
  double values[30], weights[30]; int nv = 0;
  for (int k = 0; k < 30; ++k)
    if (valid[30 * (int64_t)s + k]) { values[nv] = (double)ray_t[30 * (int64_t)s + k]; weights[nv] = (double)c_cone_dirs[k][2]; ++nv; }
  sdf[s] = sdf_statistics(values, weights, nv, scale);
}

// Structure-to-structure closestPair (BvhN.hpp:1081-1114 with a structure as the query, :1900-2009): the winner among the per-element
// answers of the query structure's elements, in the canonical order (squared distance, target index, query index).
// best[0] = (d2 bits << 32 | target index) minimum, best[1] = lowest query index attaining it.
__global__ void k_pair_min_key(int const * __restrict__ idx, float const * __restrict__ d2, int64_t n, unsigned long long * best)
{
  unsigned long long m = ~0ull;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (idx[i] >= 0) { unsigned long long k = ((unsigned long long)(unsigned)__float_as_int(d2[i]) << 32) | (unsigned)idx[i]; m = k < m ? k : m; }
  for (int o = 16; o > 0; o >>= 1) { unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o); m = x < m ? x : m; }
  if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(&best[0], m);
}
__global__ void k_pair_min_query(int const * __restrict__ idx, float const * __restrict__ d2, int64_t n, unsigned long long * best)
{
  unsigned long long want = best[0], m = ~0ull;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (idx[i] >= 0 && ((((unsigned long long)(unsigned)__float_as_int(d2[i]) << 32) | (unsigned)idx[i]) == want)) m = (unsigned long long)i < m ? (unsigned long long)i : m;
  for (int o = 16; o > 0; o >>= 1) { unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o); m = x < m ? x : m; }
  if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(&best[1], m);
}

// Structure-to-structure kClosestPairs: the k-NN rows of the query structure's points (P x k, ascending per row, -1 padded) become sort
// keys (d2 bits << 32 | target index) with the flat position as the value; a stable radix sort then leaves the k smallest pairs in the
// canonical order (squared distance, target index, query index) at the front.  k_pair_rows decodes them.
__global__ void k_pair_keys(int const * __restrict__ idx, float const * __restrict__ d2, int64_t n, unsigned long long * keys, unsigned * vals)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
  {
    keys[i] = idx[i] >= 0 ? (((unsigned long long)(unsigned)__float_as_int(d2[i]) << 32) | (unsigned)idx[i]) : ~0ull;
    vals[i] = (unsigned)i;
  }
}
__global__ void k_pair_rows(unsigned long long const * __restrict__ keys, unsigned const * __restrict__ vals, int k, int kk, long long * target, long long * query, float * d2)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x; if (j >= kk) return;
  bool ok = keys[j] != ~0ull;
  target[j] = ok ? (long long)(unsigned)(keys[j] & 0xffffffffull) : -1;
  query[j] = ok ? (long long)(vals[j] / (unsigned)k) : -1;
  d2[j] = ok ? __int_as_float((int)(unsigned)(keys[j] >> 32)) : -1.0f;
}

// max |x| over two float arrays, as int bits (non-negative floats order like ints).
__global__ void k_max_abs(float const * __restrict__ a, float const * __restrict__ b, int64_t n, int * out_bits)
{
  float m = 0.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
  { float x = fabsf(a[i]), y = fabsf(b[i]); if (x == x) m = fmaxf(m, x); if (y == y) m = fmaxf(m, y); }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_int(m));
}

}  // namespace tgq
