// tgq_kernels.cuh -- __global__ kernels of the engine (sm_100a).  Build: K1 bounds+Morton, K3 hierarchy, K4 refit,
// node emission; queries: K5 closest, K6 k-NN, K7 ball graph, K8 ray first/any hit (+ SDF cone rays), K9 ICP sums,
// K11 area sampling.  Numbering follows SURVEY.md 2.2.
#pragma once
#include "tgq_traverse.cuh"
#include "tgq_select.cuh"
#include <cuda_runtime.h>

namespace tgq {

#define TGQ_BLOCK 128

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

// ---- pancake statistics of every radix-tree node that a surviving node references as a child -------------------------------
// sinfo[2x] = (dir.xyz, smin), sinfo[2x+1] = (smax, R, 0, 0) for inner node x, about the centre of x's padded box.
struct SlabAcc { v3 dir, c; float mn, mx, r2; };

__device__ __forceinline__ SlabAcc slab_begin(f4 ns, f4 blo, f4 bhi, float pad)
{
  SlabAcc a;
  a.dir = normalized3(mk3(ns.x, ns.y, ns.z));
  if (!(sqn3(a.dir) > 0.5f)) a.dir = mk3(0.0f, 0.0f, 0.0f);          // cancelled normals (closed surface): no usable slab
  a.c = box_centre(mk3(blo.x - pad, blo.y - pad, blo.z - pad), mk3(bhi.x + pad, bhi.y + pad, bhi.z + pad));
  a.mn = FLT_MAX; a.mx = -FLT_MAX; a.r2 = 0.0f;
  return a;
}
__device__ __forceinline__ void slab_add_tri(SlabAcc & a, TriRec const * tris, int k)
{
  f4 const * t = (f4 const *)(tris + k);
  f4 q0 = __ldg(t), q1 = __ldg(t + 1), q2 = __ldg(t + 2);
  slab_accumulate(a.dir, a.c, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), a.mn, a.mx, a.r2);
}
__device__ __forceinline__ void slab_store(SlabAcc const & a, float pad, f4 * sinfo, int x)
{
  sinfo[2 * x] = make_float4(a.dir.x, a.dir.y, a.dir.z, a.mn - 2.0f * pad);
  sinfo[2 * x + 1] = make_float4(a.mx + 2.0f * pad, slab_radius(a.r2, pad), 0.0f, 0.0f);
}

#define TGQ_SLAB_SMALL 32
// Nodes of up to TGQ_SLAB_SMALL elements: one thread scans the node's element range; larger nodes are appended to `big`.
__global__ void k_slab_small(int n, int leaf, int const * __restrict__ first, int const * __restrict__ last, int const * __restrict__ parent,
                             f4 const * __restrict__ nbox_lo, f4 const * __restrict__ nbox_hi, f4 const * __restrict__ nsum,
                             TriRec const * __restrict__ tris, SceneBox const * sb, f4 * sinfo, int * big, int * big_count)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x; if (x >= n - 1 || x == 0) return;
  int pa = parent[x];
  bool parent_live = pa == 0 || (last[pa] - first[pa] + 1) > leaf;
  if (!parent_live) return;                     // strictly inside a collapsed leaf: never referenced
  int f = first[x], l = last[x];
  if (l - f + 1 > TGQ_SLAB_SMALL) { big[atomicAdd(big_count, 1)] = x; return; }
  float pad = scene_decode(sb).pad;
  SlabAcc a = slab_begin(nsum[x], nbox_lo[x], nbox_hi[x], pad);
  for (int k = f; k <= l; ++k) slab_add_tri(a, tris, k);
  slab_store(a, pad, sinfo, x);
}

// Larger nodes: one thread block per node, block-wide min / max reduction (exact, order independent).
__global__ void __launch_bounds__(128)
k_slab_big(int const * __restrict__ big, int const * __restrict__ big_count, int const * __restrict__ first, int const * __restrict__ last,
           f4 const * __restrict__ nbox_lo, f4 const * __restrict__ nbox_hi, f4 const * __restrict__ nsum, TriRec const * __restrict__ tris,
           SceneBox const * sb, f4 * sinfo)
{
  __shared__ float s_mn[4], s_mx[4], s_r2[4];
  for (int w = blockIdx.x; w < *big_count; w += gridDim.x)
  {
    int x = big[w];
    float pad = scene_decode(sb).pad;
    SlabAcc a = slab_begin(nsum[x], nbox_lo[x], nbox_hi[x], pad);
    for (int k = first[x] + threadIdx.x; k <= last[x]; k += blockDim.x) slab_add_tri(a, tris, k);
    for (int o = 16; o > 0; o >>= 1)
    {
      a.mn = fminf(a.mn, __shfl_xor_sync(0xffffffffu, a.mn, o)); a.mx = fmaxf(a.mx, __shfl_xor_sync(0xffffffffu, a.mx, o));
      a.r2 = fmaxf(a.r2, __shfl_xor_sync(0xffffffffu, a.r2, o));
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = a.mn; s_mx[threadIdx.x >> 5] = a.mx; s_r2[threadIdx.x >> 5] = a.r2; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      for (int j = 1; j < 4; ++j) { s_mn[0] = fminf(s_mn[0], s_mn[j]); s_mx[0] = fmaxf(s_mx[0], s_mx[j]); s_r2[0] = fmaxf(s_r2[0], s_r2[j]); }
      a.mn = s_mn[0]; a.mx = s_mx[0]; a.r2 = s_r2[0];
      slab_store(a, pad, sinfo, x);
    }
  }
}

// Emits the 64-byte traversal node (and, for triangle structures, the 48-byte slab record) of every surviving inner node
// at its compacted index.
__global__ void k_emit(int n, int leaf, int const * __restrict__ first, int const * __restrict__ last, int const * __restrict__ split,
                       int const * __restrict__ live, int const * __restrict__ newidx, f4 const * __restrict__ pbox_lo,
                       f4 const * __restrict__ pbox_hi, f4 const * __restrict__ nbox_lo, f4 const * __restrict__ nbox_hi,
                       SceneBox const * sb, NodeRec * nodes, TriRec const * __restrict__ tris, f4 const * __restrict__ sinfo, SlabRec * slabs)
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
  if (slabs)
  {
    f4 sa[2], sb2[2];
    for (int c = 0; c < 2; ++c)
    {
      int cf = c == 0 ? f : sp + 1, cl = c == 0 ? sp : l;
      int inner = c == 0 ? sp : sp + 1;
      if (cl == cf)
      {
        // single triangle: its own plane and its radius about the box centre
        f4 const * t = (f4 const *)(tris + cf);
        f4 q0 = __ldg(t), q1 = __ldg(t + 1), q2 = __ldg(t + 2);
        v3 v0 = mk3(q0.x, q0.y, q0.z), v1 = mk3(q1.x, q1.y, q1.z), v2 = mk3(q2.x, q2.y, q2.z);
        v3 cr = cross3(sub3(v1, v0), sub3(v2, v0));
        SlabAcc a = slab_begin(make_float4(cr.x, cr.y, cr.z, 0.0f), pbox_lo[cf], pbox_hi[cf], pad);
        slab_accumulate(a.dir, a.c, v0, v1, v2, a.mn, a.mx, a.r2);
        sa[c] = make_float4(a.dir.x, a.dir.y, a.dir.z, a.mn - 2.0f * pad);
        sb2[c] = make_float4(a.mx + 2.0f * pad, slab_radius(a.r2, pad), 0.0f, 0.0f);
      }
      else { sa[c] = sinfo[2 * inner]; sb2[c] = sinfo[2 * inner + 1]; }
    }
    SlabRec r;
    r.a = make_float4(sa[0].x, sa[0].y, sa[0].z, sa[0].w);
    r.b = make_float4(sb2[0].x, sa[1].x, sa[1].y, sa[1].z);
    r.c = make_float4(sa[1].w, sb2[1].x, sb2[0].y, sb2[1].y);
    slabs[newidx[i]] = r;
  }
}

// Structure with a single element: one node whose second child is an empty leaf behind an inverted box.
__global__ void k_emit_single(f4 const * pbox_lo, f4 const * pbox_hi, SceneBox const * sb, NodeRec * nodes, SlabRec * slabs)
{
  float pad = scene_decode(sb).pad;
  if (slabs)   // no usable direction: zero dir, unbounded slab, unbounded radius => the pancake bound is 0
  { SlabRec r; r.a = make_float4(0, 0, 0, -FLT_MAX); r.b = make_float4(FLT_MAX, 0, 0, 0); r.c = make_float4(-FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX); slabs[0] = r; }
  f4 blo = pbox_lo[0], bhi = pbox_hi[0];
  nodes[0] = make_node(mk3(blo.x - pad, blo.y - pad, blo.z - pad), mk3(bhi.x + pad, bhi.y + pad, bhi.z + pad),
                       mk3(FLT_MAX, FLT_MAX, FLT_MAX), mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX), ~((0 << TGQ_LEAF_BITS) | 1), ~0, 0, 0);
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
__global__ void k_query_codes(float const * __restrict__ q3, int n, SceneBox const * sb, uint32_t * keys, uint32_t * vals)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  SceneF s = scene_decode(sb);
  // codes over the scene box grown x2 about its centre so that off-surface queries keep their locality
  v3 c = mul3(add3(s.lo, s.hi), 0.5f), e = sub3(s.hi, s.lo);
  v3 lo = sub3(c, e);
  v3 inv = mk3(e.x > 0 ? 0.5f / e.x : 0.0f, e.y > 0 ? 0.5f / e.y : 0.0f, e.z > 0 ? 0.5f / e.z : 0.0f);
  keys[i] = morton30(ld3(q3, i), lo, inv); vals[i] = (uint32_t)i;
}

// ---- K5 closest element -------------------------------------------------------------------------------------------------
template <bool TRI, bool SLAB>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_closest(BvhView b, SlabRec const * __restrict__ slabs, float scene_pad, float const * __restrict__ q3, uint32_t const * __restrict__ perm,
          int n, float bound_ub, int exclude_self, int64_t self_offset, int * idx, float * d2, float * cp3)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
  int64_t i = perm ? perm[t] : t;
  v3 p = ld3(q3, i);
  float best = bound_ub; int bi = -1; v3 cp = mk3(0.0f, 0.0f, 0.0f);
  closest_query<TRI, SLAB>(b, slabs, scene_pad, p, exclude_self ? (int)(self_offset + i) : -1, best, bi, cp);
  if (idx) idx[i] = bi;
  if (d2) d2[i] = bi >= 0 ? best : -1.0f;
  if (cp3) st3(cp3, i, cp);
}

// ---- K6 k nearest neighbours: one bounded sorted set per thread in shared memory -------------------------------------
__global__ void __launch_bounds__(TGQ_BLOCK)
k_knn(BvhView b, float const * __restrict__ q3, uint32_t const * __restrict__ perm, int n, int k, float bound_ub, int exclude_self,
      int64_t self_offset, int * idx, float * d2)
{
  extern __shared__ float smem[];
  float * keys = smem + threadIdx.x; int * ids = (int *)(smem + (size_t)k * blockDim.x) + threadIdx.x;
  int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
  int64_t i = perm ? perm[t] : t;
  SortedSet set; set.init(keys, ids, blockDim.x, k);
  knn_query(b, ld3(q3, i), exclude_self ? (int)(self_offset + i) : -1, bound_ub, set);
  for (int j = 0; j < k; ++j)
  {
    bool v = j < set.n;
    if (idx) idx[i * k + j] = v ? ids[j * blockDim.x] : -1;
    if (d2) d2[i * k + j] = v ? keys[j * blockDim.x] : -1.0f;
  }
}

// ---- K7 ball graph over the structure's own points (thread t handles the point at Morton position t) ------------------
__global__ void __launch_bounds__(TGQ_BLOCK)
k_ball_graph(BvhView b, int n, int k, float init_radius, int rounds, float min_dot, int min_degree, int * idx, float * sep, int * count)
{
  extern __shared__ float smem[];
  float * keys = smem + threadIdx.x; int * ids = (int *)(smem + (size_t)k * blockDim.x) + threadIdx.x;
  int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= n) return;
  f4 e = b.pts[s]; int64_t i = __float_as_int(e.w);
  bool has_n = b.nrm != nullptr;
  v3 nc = mk3(0.0f, 0.0f, 0.0f);
  if (has_n) { f4 q = b.nrm[s]; nc = mk3(q.x, q.y, q.z); }
  SortedSet set; set.init(keys, ids, blockDim.x, k);
  ball_graph_query(b, mk3(e.x, e.y, e.z), nc, has_n, min_dot, (int)i, init_radius, rounds, min_degree, set);
  for (int j = 0; j < k; ++j)
  {
    bool v = j < set.n;
    if (idx) idx[i * k + j] = v ? ids[j * blockDim.x] : -1;
    if (sep) sep[i * k + j] = v ? keys[j * blockDim.x] : -1.0f;
  }
  if (count) count[i] = set.n;
}

// ---- K8 ray first hit / any hit ------------------------------------------------------------------------------------------
template <bool ANY>
__global__ void __launch_bounds__(TGQ_BLOCK)
k_ray(BvhView b, float const * __restrict__ o3, float const * __restrict__ d3, float const * __restrict__ tmax,
      uint32_t const * __restrict__ perm, int n, float * t_out, int * idx, float * normal3, uint8_t * hit)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
  int64_t i = perm ? perm[t] : t;
  float best = tmax ? tmax[i] : -1.0f; int bi = -1, bk = -1;
  ray_query<ANY>(b, ld3(o3, i), ld3(d3, i), best, bi, bk);
  if (t_out) t_out[i] = bi >= 0 ? best : -1.0f;
  if (idx) idx[i] = bi;
  if (hit) hit[i] = bi >= 0 ? 1 : 0;
  if (normal3)
  {
    v3 nn = mk3(0.0f, 0.0f, 0.0f);
    if (bk >= 0) { f4 q = ((f4 const *)(b.tris + bk))[3]; nn = mk3(q.x, q.y, q.z); }
    st3(normal3, i, nn);
  }
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

__global__ void __launch_bounds__(TGQ_BLOCK)
k_sdf_rays(BvhView b, float const * __restrict__ p3, float const * __restrict__ n3, int64_t nrays, float scale, int interior,
           float * t_out, int * idx, uint8_t * valid)
{
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (r >= nrays) return;
  int64_t s = r / 30; int k = (int)(r - s * 30);
  SdfFrame f = sdf_frame(ld3(p3, s), ld3(n3, s), scale);
  v3 dir = sdf_dir(f, mk3(c_cone_dirs[k][0], c_cone_dirs[k][1], c_cone_dirs[k][2]));
  float best = -1.0f; int bi = -1, bk = -1;
  ray_query<false>(b, f.origin, dir, best, bi, bk);
  bool ok = bi >= 0;
  if (ok && interior) { f4 q = ((f4 const *)(b.tris + bk))[3]; ok = dot3(mk3(q.x, q.y, q.z), dir) >= 0.0f; }
  if (t_out) t_out[r] = bi >= 0 ? best : -1.0f;
  if (idx) idx[r] = bi;
  valid[r] = ok ? 1 : 0;
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
__global__ void k_expand_indexed(float const * __restrict__ verts3, int const * __restrict__ tri3, int64_t ncorners, int64_t V, float * out9)
{
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (c >= ncorners) return;
  int64_t v = tri3[c]; v = v < 0 ? 0 : (v >= V ? V - 1 : v);
  st3(out9, c, ld3(verts3, v));
}

// PointSet3.hpp:363-376: squared distance from seed sample seeds[j] to its nearest OTHER sample.
__global__ void k_self_nn(BvhView b, int const * __restrict__ pos_of_elem, int const * __restrict__ seeds, int n, float * d2_out)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x; if (j >= n) return;
  int i = seeds[j];
  f4 e = b.pts[pos_of_elem[i]];
  float best = INFINITY; int bi = -1; v3 cp = mk3(0.0f, 0.0f, 0.0f);
  closest_query<false, false>(b, nullptr, 0.0f, mk3(e.x, e.y, e.z), i, best, bi, cp);
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
