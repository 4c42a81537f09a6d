// tgq_bvh.cuh -- device-resident LBVH: record layouts and the per-element build steps (Morton code, Karras hierarchy,
// bottom-up refit, node emission).  Replaces the median-split kd-tree of Algorithms/BvhN.hpp:1365-1525 (createKdTree)
// and the refit of BvhN.hpp:816-834,1694-1730 (updateNodeBounds); licensed by SURVEY.md finding 4: every query result on
// this path is tree independent, so only the leaf arithmetic (tgq_math.cuh) has to match the reference bit for bit.
//
// Layout in HBM (all records 16-byte aligned, read with 128-bit loads):
//   Node  (64 B)  : boxes of BOTH children + child links, so one node fetch decides both descents.
//        q0 = (c0.lo.x, c0.lo.y, c0.lo.z, c0.hi.x)   q1 = (c0.hi.y, c0.hi.z, c1.lo.x, c1.lo.y)
//        q2 = (c1.lo.z, c1.hi.x, c1.hi.y, c1.hi.z)   q3 = (child0, child1, count0, count1) as int bits
//        child >= 0: index of an inner node; child < 0: leaf holding elements [~child, ~child + count) of the
//        Morton-ordered element array.
//   Obb   (96 B)  : triangle structures only, same index as Node: per child (centre, h0) (axis0, h1|h2 as two round-up
//        bfloat16) (axis1, child link); axis2 = axis0 x axis1 is recomputed.  Read by closest-point queries INSTEAD of
//        Node (rays, k-NN and ball queries read Node): three 256-bit loads per node instead of four.
//   Tri   (128 B) : q0 = (v0, orig_index)  q1 = (v1, axis)  q2 = (v2, plane d)  q3 = (unit normal, screening radius about the centroid)
//        e0 e1 e2 = (unit direction, length) of the edges v0->v1, v1->v2, v2->v0   x = (1 / area2, area2, -, -)
//        (ray queries read q0..q2 = 48 B; a closest-point step reads the whole 128-byte line, the same fetch shape as an
//        oriented-box node, so that one load phase serves the lanes that are at a node and the lanes that are at a triangle)
//   Point (16 B)  : (x, y, z, orig_index); optional normals (16 B) in a parallel array.
#pragma once
#include "tgq_math.cuh"
#include <string.h>

namespace tgq {

struct NodeRec { f4 q0, q1, q2, q3; };
struct TriRec  { f4 q0, q1, q2, q3, e0, e1, e2, x; };
// ObbNode, the oriented-box node of the closest-point traversal on triangle structures: 64 bytes since round 2 -- ONE frame for both
// children (fp16 axes, re-orthonormalised by the decode), two centres + links, six half extents as round-up bfloat16; layout and
// error budget in tgq_traverse.cuh.  -DTGQ_NODE96 builds the earlier 96-byte record (a full-precision frame per child, three 256-bit
// loads per visit instead of two): measured on B200, 64 against 96 bytes: cfg5 closest point 407 against 395 M queries/s on the 50 M
// batch, 357 against 324 M/s on 8 M batches, 509 against 500 M/s on the 1 M-triangle mesh.
#if !defined(TGQ_NODE96) && !defined(TGQ_NODE64)
#define TGQ_NODE64
#endif
#ifdef TGQ_NODE64
#define TGQ_OBB_F4 2              // float4 per child half (unpacked build-time form); the packed node is 4 float4
#define TGQ_OBB_LINK_F4 0         // which float4 of a child carries the link in .w
#else
#define TGQ_OBB_F4 3
#define TGQ_OBB_LINK_F4 2
#endif
struct ObbNode { f4 q[2 * TGQ_OBB_F4]; };     // oriented boxes of both children (64 B; 96 B with TGQ_NODE96), see tgq_traverse.cuh
// 4-wide node of the closest-point traversal on triangle structures (k_closest_wide): every second level of the binary tree
// collapsed, so a node has 2..4 children.  One query is traversed by FOUR lanes, lane i testing child i: the record is laid out so
// that the four lanes' 256-bit loads are one contiguous, aligned 128-byte line each (two L1 wavefronts per node visit and query,
// against three wavefronts per BINARY node when every lane walks its own query):
//   q[2 i], q[2 i + 1]          (a0, -a0 . c) (a1, -a1 . c)      first line:  children 0..3
//   q[8 + 2 i], q[8 + 2 i + 1]  (a2, -a2 . c) (h0, h1, h2, link)  second line: children 0..3
// a0..a2: the child's own orthonormal frame (a0 = mean normal), c its box centre, h its half extents; link as in NodeRec, or
// TGQ_WIDE_EMPTY for an unused slot.  Everything is fp32 and nothing is recomputed per visit (the 96-byte binary record packs two
// extents as bfloat16 and rebuilds a2 with a cross product): with coalesced loads bytes are cheap and instructions are not.
struct WideNode { f4 q[16]; };
#define TGQ_WIDE_EMPTY (-1)     // == ~0: a leaf of no elements

// ---- Morton codes: 21 bits per axis, 63-bit keys --------------------------------------------------------------------
TGQ_HD uint64_t expand21(uint64_t v)
{
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8))  & 0x100f00f00f00f00full;
  v = (v | (v << 4))  & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2))  & 0x1249249249249249ull;
  return v;
}

TGQ_HD uint64_t morton63(v3 c, v3 lo, v3 inv_ext)
{
  float fx = fminf(fmaxf((c.x - lo.x) * inv_ext.x, 0.0f), 1.0f);
  float fy = fminf(fmaxf((c.y - lo.y) * inv_ext.y, 0.0f), 1.0f);
  float fz = fminf(fmaxf((c.z - lo.z) * inv_ext.z, 0.0f), 1.0f);
  uint64_t x = (uint64_t)fminf(fx * 2097152.0f, 2097151.0f);
  uint64_t y = (uint64_t)fminf(fy * 2097152.0f, 2097151.0f);
  uint64_t z = (uint64_t)fminf(fz * 2097152.0f, 2097151.0f);
  return (expand21(x) << 2) | (expand21(y) << 1) | expand21(z);
}

// 30-bit code for ordering query batches (10 bits per axis).
TGQ_HD uint32_t expand10(uint32_t v)
{
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8))  & 0x0300f00fu;
  v = (v | (v << 4))  & 0x030c30c3u;
  v = (v | (v << 2))  & 0x09249249u;
  return v;
}
TGQ_HD uint32_t morton30(v3 c, v3 lo, v3 inv_ext)
{
  float fx = fminf(fmaxf((c.x - lo.x) * inv_ext.x, 0.0f), 1.0f);
  float fy = fminf(fmaxf((c.y - lo.y) * inv_ext.y, 0.0f), 1.0f);
  float fz = fminf(fmaxf((c.z - lo.z) * inv_ext.z, 0.0f), 1.0f);
  uint32_t x = (uint32_t)fminf(fx * 1024.0f, 1023.0f);
  uint32_t y = (uint32_t)fminf(fy * 1024.0f, 1023.0f);
  uint32_t z = (uint32_t)fminf(fz * 1024.0f, 1023.0f);
  return (expand10(x) << 2) | (expand10(y) << 1) | expand10(z);
}

// ---- Karras 2012 hierarchy over n sorted keys (ties broken by position) ----------------------------------------------
TGQ_HD int clz64(uint64_t x)
{
#ifdef __CUDA_ARCH__
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}

// Length of the common prefix of keys i and j (augmented with the position to make all keys distinct); -1 out of range.
TGQ_HD int lbvh_delta(uint64_t const * keys, int n, int i, int j)
{
  if (j < 0 || j >= n) return -1;
  uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + clz64((uint64_t)(uint32_t)i ^ (uint64_t)(uint32_t)j) - 32;   // 64 + clz32(i ^ j)
  return clz64(a ^ b);
}

// Inner node i of the binary radix tree: range [first, last] of sorted elements and the split position: the left child
// covers [first, split], the right child [split + 1, last].
TGQ_HD void lbvh_range(uint64_t const * keys, int n, int i, int & first, int & last, int & split)
{
  int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = lbvh_delta(keys, n, i, i - d);
  int lmax = 2;
  while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = lbvh_delta(keys, n, i, j);
  int s = 0;
  int t = l;
  do
  {
    t = (t + 1) >> 1;
    if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  split = i + s * d + (d < 0 ? -1 : 0);
  first = d > 0 ? i : j;
  last = d > 0 ? j : i;
}

// ---- Node record helpers ---------------------------------------------------------------------------------------
TGQ_HD void node_boxes(NodeRec const & nd, v3 & lo0, v3 & hi0, v3 & lo1, v3 & hi1)
{
  lo0 = mk3(nd.q0.x, nd.q0.y, nd.q0.z); hi0 = mk3(nd.q0.w, nd.q1.x, nd.q1.y);
  lo1 = mk3(nd.q1.z, nd.q1.w, nd.q2.x); hi1 = mk3(nd.q2.y, nd.q2.z, nd.q2.w);
}

TGQ_HD NodeRec make_node(v3 lo0, v3 hi0, v3 lo1, v3 hi1, int c0, int c1, int n0, int n1)
{
  NodeRec r;
  r.q0.x = lo0.x; r.q0.y = lo0.y; r.q0.z = lo0.z; r.q0.w = hi0.x;
  r.q1.x = hi0.y; r.q1.y = hi0.z; r.q1.z = lo1.x; r.q1.w = lo1.y;
  r.q2.x = lo1.z; r.q2.y = hi1.x; r.q2.z = hi1.y; r.q2.w = hi1.z;
  r.q3.x = i2f(c0); r.q3.y = i2f(c1); r.q3.z = i2f(n0); r.q3.w = i2f(n1);
  return r;
}

}  // namespace tgq
