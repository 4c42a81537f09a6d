// tgq_traverse.cuh -- resumable single-query traversals of the device LBVH.
//
// Each traversal computes the *canonical* answer of the corresponding BvhN query (SURVEY.md findings 4-5): the optimum
// over ALL elements of the reference's per-element fp32 function, ties broken by the lowest element index.  The tree
// only prunes: a subtree is skipped when a conservative bound proves that none of its elements can win or tie.
//   Closest*   <- BvhN::closestElement / closestPair        (Algorithms/BvhN.hpp:1049-1114,1866-1893,2036-2089)
//   Knn*       <- BvhN::kClosestPairs                       (BvhN.hpp:1137-1146, functor :474-500)
//   Ball*      <- processRange<Ball3> + NeighborFunctor     (BvhN.hpp:2118-2171, PointSet3.hpp:586-681)
//   Ray*       <- BvhN::rayStructureIntersection / rayIntersectionTime / rayIntersects (BvhN.hpp:1266-1316,2188-2291)
//   Seg*, TriQ* <- BvhN::closestPair with a LineSegment3 / Triangle3 query object                (BvhN.hpp:1081-1114; end of this file)
//
// Every traversal is a small state machine: `cur` >= 0 means "an inner node is pending", TGQ_ST_LEAF "elements
// [k, kend) of a leaf are pending", TGQ_ST_DONE "finished".  One fetch + *_node_compute / *_leaf_compute advances the query
// by one unit of work.  The kernels (tgq_kernels.cuh) drive 32 such machines per warp from a warp-uniform loop, so the
// lanes reconverge after every step and finished lanes are refilled with fresh queries; the host emulation
// of the test suite drives the same step functions from a plain loop.
#pragma once
#include "tgq_bvh.cuh"

namespace tgq {

#define TGQ_STACK 128            // > 63 Morton bits + 32 tie-break bits: ordered DFS never holds more than depth entries
#define TGQ_LEAF_BITS 4
#define TGQ_LEAF_MASK 15
#define TGQ_LB_DEFLATE 0.99999714f   // 1 - 3 * 2^-20 (48 u): relative part of the error budget below
#ifndef TGQ_MARGIN_SCALE
#define TGQ_MARGIN_SCALE 1.0f        // test builds widen every absolute margin (libtgq_wide.so) for differential checks
#endif

#define TGQ_ST_DONE (-1)
#define TGQ_ST_LEAF (-2)
#define TGQ_ST_IDLE (-3)
#define TGQ_ST_POP (-4)            // wants the next deferred sibling from its stack

// Visit counters: compiled out everywhere except the test-only host build, which uses them to study pruning quality.
#ifndef TGQ_COUNT_NODE
#define TGQ_COUNT_NODE()
#define TGQ_COUNT_PRIM()
#endif
#ifndef TGQ_COUNT_APPEND
#define TGQ_COUNT_APPEND()
#define TGQ_COUNT_COMPACT()
#endif

struct BvhView
{
  NodeRec const * nodes;   // n_nodes inner nodes with the AABBs of both children, root = 0
  ObbNode const * obbs;    // same nodes with oriented boxes of both children (triangle structures), or null
  WideNode const * wide;   // 4-wide oriented-box nodes (triangle structures), root = 0, or null
  int n_wide;
  TriRec const * tris;     // Morton-ordered triangle records (triangle structures)
  f4 const * pts;          // Morton-ordered points (point structures)
  f4 const * nrm;          // Morton-ordered normals or null
  int const * lab;         // Morton-ordered element labels or null (filtered queries on point structures)
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

// Two consecutive float4 (32-byte aligned) in ONE 256-bit load (sm_100: LDG.E.256).  The traversal kernels are bound by
// the number of L1 wavefronts -- every lane reads a different line, so each load instruction costs one wavefront per lane
// (ncu: l1tex data-pipe wavefronts at 75 % of peak with 128-bit loads) -- and a 256-bit load halves the instruction count
// per record.
TGQ_HD void ldq2(f4 const * p, f4 & a, f4 & b)
{
#ifdef __CUDA_ARCH__
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
  a = p[0]; b = p[1];
#endif
}

TGQ_HD NodeRec load_node(BvhView const & b, int i)
{
  NodeRec r; f4 const * p = (f4 const *)(b.nodes + i);
  r.q0 = ldq(p); r.q1 = ldq(p + 1); r.q2 = ldq(p + 2); r.q3 = ldq(p + 3);
  return r;
}

// Traversal bookkeeping shared by all query kinds: pending node / leaf range and the stack of deferred siblings.
// The first `fast_cap` stack entries live in a caller-provided strided array (shared memory on the device: entry j of a
// thread at fast[j * fast_stride], conflict free for any mix of depths because the stride is a multiple of the bank
// count), deeper entries in TravStack (thread-local memory; ordered DFS rarely gets that deep).  ncu on the
// local-memory-only version showed 30 % of all stall samples on the pop loop's dependent local loads.
struct alignas(8) StackEntry { int node; float key; };
struct TravStack { int node[TGQ_STACK]; float key[TGQ_STACK]; };
struct Trav
{
  int cur, k, kend, sp;
  StackEntry * fast; int fast_stride, fast_cap;
  int * stk_node; float * stk_key;
};
TGQ_HD void trav_bind(Trav & t, TravStack & st) { t.stk_node = st.node; t.stk_key = st.key; t.fast = nullptr; t.fast_stride = 0; t.fast_cap = 0; }
TGQ_HD void trav_bind_fast(Trav & t, StackEntry * fast, int stride, int cap) { t.fast = fast; t.fast_stride = stride; t.fast_cap = cap; }

TGQ_HD void trav_enter(Trav & t, int code)
{
  // branch free: inlined at every descend / pop site inside divergent code, where a branch per site cost more than the selects
  bool leaf = code < 0; int c = ~code;
  int k = c >> TGQ_LEAF_BITS, kend = k + (c & TGQ_LEAF_MASK);
  t.k = leaf ? k : t.k; t.kend = leaf ? kend : t.kend;
  t.cur = leaf ? ((c & TGQ_LEAF_MASK) > 0 ? TGQ_ST_LEAF : TGQ_ST_POP) : code;
}
TGQ_HD void trav_begin(Trav & t, int n_nodes) { t.sp = 0; t.k = t.kend = 0; t.cur = n_nodes > 0 ? 0 : TGQ_ST_DONE; }
TGQ_HD void trav_push(Trav & t, int code, float key)
{
  if (t.sp < t.fast_cap) { StackEntry e; e.node = code; e.key = key; t.fast[t.sp * t.fast_stride] = e; }
  else { t.stk_node[t.sp - t.fast_cap] = code; t.stk_key[t.sp - t.fast_cap] = key; }
  ++t.sp;
}
// Popping is a state of its own (TGQ_ST_POP), not a loop inside the node / leaf arithmetic: ncu showed the pop loop running
// with 2.6 of 32 lanes active (a quarter of all stall samples) when every lane looped on its own stack inside divergent code.
// One call takes ONE deferred sibling: entered if its key (lower bound / entry time) is still <= limit, else dropped and the
// lane stays in TGQ_ST_POP; an empty stack finishes the query.
TGQ_HD void trav_pop_step(Trav & t, float limit)
{
  if (t.sp <= 0) { t.cur = TGQ_ST_DONE; return; }
  --t.sp;
  int code; float key;
  if (t.sp < t.fast_cap) { StackEntry e = t.fast[t.sp * t.fast_stride]; code = e.node; key = e.key; }
  else { code = t.stk_node[t.sp - t.fast_cap]; key = t.stk_key[t.sp - t.fast_cap]; }
  if (key <= limit) trav_enter(t, code);
}
TGQ_HD void trav_pop(Trav & t) { t.cur = TGQ_ST_POP; }

// ---------------------------------------------------------------------------------------------------------------------
// Oriented bounds for closest-point queries on triangle structures.  AABBs of a smooth surface that is oblique to the
// axes stick out towards an off-surface query by about their own size, so a far query would have to open every box in
// a wide disc around its foot point (hundreds of leaves on a 1M-triangle mesh).  Each child of a triangle-structure
// node therefore carries a box oriented along the mean normal of its triangles (axis a0) and two tangent directions:
//     |p - x|^2 >= sum_i max(|a_i . (p - c)| - h_i, 0)^2   for every x in the child,
// which is thin across the surface and keeps the lateral offset.  ObbNode (96 B), per child three float4:
//     (c.xyz, h0) (a0.xyz, h1|h2 as round-up bfloat16) (a1.xyz, child link)        a2 = a0 x a1 is recomputed
// Half extents are padded outwards at build time (scene pad); `mq` absorbs the rounding of the query-side products (it
// grows with the magnitude of the query coordinates), and the result is deflated.  Children are visited in the order
// of their centre distance (a better predictor of where the foot point lies than the lower bound for far queries:
// 1.5x fewer node visits on the bench workload), and pruned on the lower bound.
// The same idea with the triangle's own plane and the radius about its centroid screens single triangles before the
// exact (and expensive) reference arithmetic is run.
// ---------------------------------------------------------------------------------------------------------------------
// Error budget of the closest-point bounds (u = 2^-24, M = largest |coordinate| of the scene, pad = 32 u M = scene pad,
// D = distance from the query p to the element).  A bound may prune an element only if it is provably below the value
// f = fl(|cp - p|^2) that the REFERENCE arithmetic returns for it, where cp is the reference's computed closest point.
//   (1) cp versus the true triangle.  cp is a vertex (exact), a + dir * t on an edge (off the edge by
//       <= u (sqrt3 |a|_inf + 4 len)), or the projection p - sd n accepted by the inside test.  For the projection:
//       |d(sd)| <= 3u|p| + u D (three products, two sums, the subtraction), the subtraction p - sd n adds u(|p| + 2D), the
//       stored normal is unit to 4u (another 4u D), the plane offset d = fl(n . v0) is off by <= 3u|v0|: cp lies within
//       4u|p| + 7u D + 3 sqrt3 u M of the triangle (the barycentric test can accept a projection outside the triangle only
//       by ~4u * extent * aspect, negligible next to this for triangles of aspect < 1e3).
//   (2) f versus |cp - p|^2: three subtractions, three squares, two sums: relative 4u.
//   (3) query side of a box bound: pc = fl(p - c) (u|pc|), fused dot with a unit axis (3u|pc|): 4u(|p| + sqrt3 M); the frame
//       is orthonormal to 4u per pair, so sum_j (a_j . w)^2 <= (1 + 12u)|w|^2.
//   (4) build side of a box: vertex projections 3u sqrt3 M, centre reconstruction c = sum_j a_j c_j 4u sqrt3 M, centre
//       versus frame 4u sqrt3 M: 19u M.
// Absolute terms: (1) + (3) <= 8u|p|_2 + 12.2u M  ->  query margin mq = 10u|p|_1 + 16u M;  (4) -> box half extents grow by
// pad = 32u M;  centroid rounding of the triangle screen (3 sqrt3 u M) -> radius grows by 16u M.  Relative terms: 14u (the
// 7u D of (1), squared) + 4u (2) + 12u (3) = 30u -> TGQ_LB_DEFLATE = 1 - 48u.
// tests/test_gpu_scale.py checks, at the full bench size, that a build with every absolute margin x4 returns the same
// bits, and the CPU suite checks the same code (host emulation) against brute force.
TGQ_HD float query_margin(v3 p, float scene_pad)
{ return TGQ_MARGIN_SCALE * (0.5f * scene_pad + 5.9604645e-7f * (fabsf(p.x) + fabsf(p.y) + fabsf(p.z))); }

// Two non-negative floats as round-UP bfloat16 in one word (the lateral half extents: 0.8 % slack at most, never too small).
TGQ_HD float pack_bf16_up(float a, float b)
{
  unsigned ua = (unsigned)f2i(a), ub = (unsigned)f2i(b);
  ua = (ua + 0xffffu) & 0xffff0000u; ub = (ub + 0xffffu) >> 16;
  return i2f((int)(ua | ub));
}
TGQ_HD void unpack_bf16(float w, float & a, float & b)
{
  unsigned u = (unsigned)f2i(w);
  a = i2f((int)(u & 0xffff0000u)); b = i2f((int)(u << 16));
}

#ifdef TGQ_NODE64
// ---- 64-byte node with ONE frame for both children --------------------------------------------------------------------------------
// The two children of a node are neighbouring patches of the surface, so their mean normals differ by about the patch's own angular
// size: measuring both boxes in the frame of the NODE's mean normal makes them thicker by about their own sag again.  Host counters on
// the bench workload (tools/seed_stride_stats.py with EMU_SHARED_FRAME=1): 132 / 96 / 167 node visits per query against 135 / 95 / 161
// with a frame per child -- no difference.  With one frame the record fits 64 bytes without any expensive decode:
//     q0 = (c0.xyz, link0)   q1 = (c1.xyz, link1)
//     q2 = (a0.x|a0.y, a0.z|a1.x, a1.y|a1.z as fp16 pairs, h0|h1 of child 0 as round-up bfloat16)
//     q3 = (h2 of child 0|h0 of child 1, h1|h2 of child 1, 0, 0)
// i.e. two 256-bit loads per node visit instead of three (the closest-point kernels are bound by L1 wavefronts, one per lane and
// load), the third axis by ONE cross product per node instead of one per child, six half-to-float conversions.  The earlier 64-byte
// experiments kept a frame per child (octahedral normal + in-plane angle, or a quaternion) and lost on their decode.
// The fp16 axes are not orthonormal (components off by up to 2^-11), and a RELATIVE slack eps of the lower bound lets every box within
// d sqrt(eps) of a far query's foot point through -- a first version that simply deflated the bound by 0.4 % opened 4-6x the nodes on
// the bench workload.  So the decoded axes are re-orthonormalised (Gram-Schmidt with the reciprocal-square-root approximation: two
// MUFU operations and ~20 multiply-adds per NODE, not per child), by the same function in the build -- which measures the extents
// along the result -- and in the query.  After it |a_i|^2 = 1 +- 4 u_r and a_i . a_j = +- 4 u_r with u_r ~ 2^-22 the error of the
// approximation: lambda_max of the Gram matrix <= 1 + 1.5e-6, covered together with the 36 u of the other relative terms by
// TGQ_OBB_DEFLATE = 1 - 96 u.  (Host emulation: 1 / sqrtf; build and query agree on each side, which is all that is needed.)
#define TGQ_OBB_DEFLATE 0.99999428f
TGQ_HD float obb_rsqrt(float x)
{
#ifdef __CUDA_ARCH__
  float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
  return 1.0f / sqrtf(x);
#endif
}
// IEEE binary16 <-> binary32 in integer arithmetic (round to nearest even; inputs here are axis components in [-1, 1]): the same bits
// on the host and on the device, so that build (device), emulation (host) and query (hardware conversion, exact) agree.
TGQ_HD unsigned f32_to_f16_bits(float x)
{
  unsigned u = (unsigned)f2i(x), sign = (u >> 16) & 0x8000u; u &= 0x7fffffffu;
  if (u >= 0x47800000u) return sign | 0x7bffu;                       // >= 65536 (never for unit axes): largest finite half
  if (u < 0x38800000u)                                               // below the smallest normal half (2^-14): subnormal or zero
  {
    if (u < 0x33000000u) return sign;                                // < 2^-25: zero
    unsigned e = u >> 23, m = (u & 0x7fffffu) | 0x800000u, sh = 126u - e;      // value = m * 2^(e - 150); half subnormal unit 2^-24
    unsigned h = m >> sh, rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1u);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return sign | h;
  }
  unsigned h = ((u - 0x38000000u) >> 13), rem = u & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
  return sign | h;
}
TGQ_HD float f16_bits_to_f32(unsigned h)
{
#ifdef __CUDA_ARCH__
  return __half2float(__ushort_as_half((unsigned short)h));
#else
  unsigned sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  if (e == 0)
  {
    if (m == 0) return i2f((int)sign);
    float v = (float)m * 5.9604644775390625e-08f;                    // m * 2^-24, exact
    return sign ? -v : v;
  }
  return i2f((int)(sign | ((e + 112u) << 23) | (m << 13)));         // no infinities / NaNs are ever encoded
#endif
}
TGQ_HD unsigned pack_f16x2(float a, float b) { return f32_to_f16_bits(a) | (f32_to_f16_bits(b) << 16); }
TGQ_HD void unpack_f16x2(unsigned w, float & a, float & b) { a = f16_bits_to_f32(w & 0xffffu); b = f16_bits_to_f32(w >> 16); }
// The frame of a 64-byte node from its three axis words.
TGQ_HD void obb_decode_frame64(unsigned w0, unsigned w1, unsigned w2, v3 & a0, v3 & a1, v3 & a2)
{
  float t0, t1;
  unpack_f16x2(w0, a0.x, a0.y); unpack_f16x2(w1, t0, t1); a0.z = t0; a1.x = t1; unpack_f16x2(w2, a1.y, a1.z);
  a0 = mul3(a0, obb_rsqrt(fdot3(a0, a0)));
  float d = fdot3(a0, a1);
  a1 = mk3(ffma(-d, a0.x, a1.x), ffma(-d, a0.y, a1.y), ffma(-d, a0.z, a1.z));
  a1 = mul3(a1, obb_rsqrt(fdot3(a1, a1)));
  a2 = mk3(ffma(a0.y, a1.z, -(a0.z * a1.y)), ffma(a0.z, a1.x, -(a0.x * a1.z)), ffma(a0.x, a1.y, -(a0.y * a1.x)));
}
// Both children of one 64-byte node: lower bounds (deflated), ordering keys, links.
TGQ_HD void obb_node_bounds64(f4 const * r, v3 p, float mq, float & lb0, float & key0, int & c0, float & lb1, float & key1, int & c1)
{
  v3 a0, a1, a2; obb_decode_frame64((unsigned)f2i(r[2].x), (unsigned)f2i(r[2].y), (unsigned)f2i(r[2].z), a0, a1, a2);
  float h00, h01, h02, h10, h11, h12;
  unpack_bf16(r[2].w, h00, h01); unpack_bf16(r[3].x, h02, h10); unpack_bf16(r[3].y, h11, h12);
  {
    v3 pc = sub3(p, mk3(r[0].x, r[0].y, r[0].z));
    float x = fmaxf(fabsf(fdot3(a0, pc)) - (h00 + mq), 0.0f), y = fmaxf(fabsf(fdot3(a1, pc)) - (h01 + mq), 0.0f), z = fmaxf(fabsf(fdot3(a2, pc)) - (h02 + mq), 0.0f);
    lb0 = fdot3(mk3(x, y, z), mk3(x, y, z)) * TGQ_OBB_DEFLATE; key0 = fdot3(pc, pc); c0 = f2i(r[0].w);
  }
  {
    v3 pc = sub3(p, mk3(r[1].x, r[1].y, r[1].z));
    float x = fmaxf(fabsf(fdot3(a0, pc)) - (h10 + mq), 0.0f), y = fmaxf(fabsf(fdot3(a1, pc)) - (h11 + mq), 0.0f), z = fmaxf(fabsf(fdot3(a2, pc)) - (h12 + mq), 0.0f);
    lb1 = fdot3(mk3(x, y, z), mk3(x, y, z)) * TGQ_OBB_DEFLATE; key1 = fdot3(pc, pc); c1 = f2i(r[1].w);
  }
}
#else
TGQ_HD void obb_bounds(f4 q0, f4 q1, f4 q2, v3 p, float mq, float & lb, float & key)
{
  v3 a0 = mk3(q1.x, q1.y, q1.z), a1 = mk3(q2.x, q2.y, q2.z);
  v3 a2 = cross3(a0, a1);                       // exactly the a2 the build measured the extents along (obb_frame)
  float h1, h2; unpack_bf16(q1.w, h1, h2);
  v3 pc = sub3(p, mk3(q0.x, q0.y, q0.z));
  float x = fmaxf(fabsf(fdot3(a0, pc)) - (q0.w + mq), 0.0f);
  float y = fmaxf(fabsf(fdot3(a1, pc)) - (h1 + mq), 0.0f);
  float z = fmaxf(fabsf(fdot3(a2, pc)) - (h2 + mq), 0.0f);
  lb = fdot3(mk3(x, y, z), mk3(x, y, z)) * TGQ_LB_DEFLATE;
  key = fdot3(pc, pc);
}
#endif

TGQ_HD float slab_radius(float r2, float pad) { return sqrtf(r2) * 1.000001f + TGQ_MARGIN_SCALE * 0.5f * pad; }
TGQ_HD v3 box_centre(v3 lo, v3 hi) { return mul3(add3(lo, hi), 0.5f); }
// Screening radius of one triangle about its centroid (stored in TriRec.q3.w).
TGQ_HD float tri_screen_radius(v3 v0, v3 v1, v3 v2, float pad)
{
  v3 c = mul3(add3(add3(v0, v1), v2), 0.33333334f);
  float r2 = fmaxf(sqn3(sub3(v0, c)), fmaxf(sqn3(sub3(v1, c)), sqn3(sub3(v2, c))));
  return slab_radius(r2, pad);
}

// Frame of an oriented box from the (unnormalised) normal sum of its triangles: a0 = mean normal (or +z when the
// normals cancel, e.g. a closed surface), a1 = the world axis least aligned with a0 made orthogonal to it, a2 = a0 x a1.
struct ObbFrame { v3 a0, a1, a2; };
TGQ_HD ObbFrame obb_frame(v3 nsum)
{
  ObbFrame f;
  f.a0 = normalized3(nsum);
  if (!(sqn3(f.a0) > 0.5f)) f.a0 = mk3(0.0f, 0.0f, 1.0f);
  float ax = fabsf(f.a0.x), ay = fabsf(f.a0.y), az = fabsf(f.a0.z);
  v3 e = (ax <= ay && ax <= az) ? mk3(1.0f, 0.0f, 0.0f) : (ay <= az ? mk3(0.0f, 1.0f, 0.0f) : mk3(0.0f, 0.0f, 1.0f));
  f.a1 = normalized3(sub3(e, mul3(f.a0, dot3(f.a0, e))));
  f.a2 = cross3(f.a0, f.a1);
  return f;
}
// Extent accumulator of one oriented box.
struct ObbAcc { ObbFrame f; float mn[3], mx[3]; unsigned wn, wr, wx; };      // wn, wr, wx: the frame words of a 64-byte node
TGQ_HD void obb_acc_begin(ObbAcc & a, v3 nsum)
{
  a.f = obb_frame(nsum); a.wn = 0u; a.wr = 0u; a.wx = 0u;
#ifdef TGQ_NODE64
  {
    // the frame as the node record will carry it (fp16 axes, a2 by the cross product of the decoded ones): the extents below are then
    // measured along exactly the axes every query uses.  The caller passes the normal sum of the NODE the box is a child of.
    a.wn = pack_f16x2(a.f.a0.x, a.f.a0.y); a.wr = pack_f16x2(a.f.a0.z, a.f.a1.x); a.wx = pack_f16x2(a.f.a1.y, a.f.a1.z);
    obb_decode_frame64(a.wn, a.wr, a.wx, a.f.a0, a.f.a1, a.f.a2);
  }
#endif
  for (int j = 0; j < 3; ++j) { a.mn[j] = FLT_MAX; a.mx[j] = -FLT_MAX; }
}
TGQ_HD void obb_acc_point(ObbAcc & a, v3 v)
{
  float s0 = dot3(a.f.a0, v), s1 = dot3(a.f.a1, v), s2 = dot3(a.f.a2, v);
  a.mn[0] = fminf(a.mn[0], s0); a.mx[0] = fmaxf(a.mx[0], s0);
  a.mn[1] = fminf(a.mn[1], s1); a.mx[1] = fmaxf(a.mx[1], s1);
  a.mn[2] = fminf(a.mn[2], s2); a.mx[2] = fmaxf(a.mx[2], s2);
}
// Packs the child half of an ObbNode.  `pad` covers the rounding of the build-side projections and of the centre.
TGQ_HD void obb_acc_store(ObbAcc const & a, float pad, int child, f4 * out4)
{
  float c0 = 0.5f * (a.mn[0] + a.mx[0]), c1 = 0.5f * (a.mn[1] + a.mx[1]), c2 = 0.5f * (a.mn[2] + a.mx[2]);
  v3 c = add3(add3(mul3(a.f.a0, c0), mul3(a.f.a1, c1)), mul3(a.f.a2, c2));
  float hp = TGQ_MARGIN_SCALE * pad;
  float h0 = 0.5f * (a.mx[0] - a.mn[0]) * 1.000001f + hp, h1 = 0.5f * (a.mx[1] - a.mn[1]) * 1.000001f + hp, h2 = 0.5f * (a.mx[2] - a.mn[2]) * 1.000001f + hp;
#ifdef TGQ_NODE64
  // child half of a 64-byte node, unpacked: (centre, link) (h0, h1, h2, -).  The decoded frame is not exactly orthonormal, so the
  // centre's own projections are taken again and each half extent covers the measured range about THEM.
  float s0 = dot3(a.f.a0, c), s1 = dot3(a.f.a1, c), s2 = dot3(a.f.a2, c);
  h0 = fmaxf(a.mx[0] - s0, s0 - a.mn[0]) * 1.000001f + hp; h1 = fmaxf(a.mx[1] - s1, s1 - a.mn[1]) * 1.000001f + hp; h2 = fmaxf(a.mx[2] - s2, s2 - a.mn[2]) * 1.000001f + hp;
  out4[0].x = c.x; out4[0].y = c.y; out4[0].z = c.z; out4[0].w = i2f(child);
  out4[1].x = h0; out4[1].y = h1; out4[1].z = h2; out4[1].w = 0.0f;
#else
  out4[0].x = c.x; out4[0].y = c.y; out4[0].z = c.z; out4[0].w = h0;
  out4[1].x = a.f.a0.x; out4[1].y = a.f.a0.y; out4[1].z = a.f.a0.z; out4[1].w = pack_bf16_up(h1, h2);
  out4[2].x = a.f.a1.x; out4[2].y = a.f.a1.y; out4[2].z = a.f.a1.z; out4[2].w = i2f(child);
#endif
}
// Child record of a WideNode (four float4: A0, A1, B0, B1 as laid out in tgq_bvh.cuh).  The box is kept in its own frame
// coordinates: s_i = a_i . v was measured over the child's vertices, so the slab along a_i is [mn_i, mx_i] and the query side needs
// u_i = a_i . p - 0.5 (mn_i + mx_i) only -- no world-space centre is reconstructed.  Error budget as for obb_bounds: the build-side
// projections are off by <= 3 sqrt3 u M (inside pad = 32 u M, added to h), the query-side fused chains by <= 4u (|p|_1 + sqrt3 M)
// (inside mq), the frame is orthonormal to 4u per pair (inside TGQ_LB_DEFLATE).
TGQ_HD void obb_acc_store_wide(ObbAcc const & a, float pad, int child, f4 * out4)
{
  float c0 = 0.5f * (a.mn[0] + a.mx[0]), c1 = 0.5f * (a.mn[1] + a.mx[1]), c2 = 0.5f * (a.mn[2] + a.mx[2]);
  float hp = TGQ_MARGIN_SCALE * pad;
  float h0 = 0.5f * (a.mx[0] - a.mn[0]) * 1.000001f + hp, h1 = 0.5f * (a.mx[1] - a.mn[1]) * 1.000001f + hp, h2 = 0.5f * (a.mx[2] - a.mn[2]) * 1.000001f + hp;
  out4[0].x = a.f.a0.x; out4[0].y = a.f.a0.y; out4[0].z = a.f.a0.z; out4[0].w = -c0;
  out4[1].x = a.f.a1.x; out4[1].y = a.f.a1.y; out4[1].z = a.f.a1.z; out4[1].w = -c1;
  out4[2].x = a.f.a2.x; out4[2].y = a.f.a2.y; out4[2].z = a.f.a2.z; out4[2].w = -c2;
  out4[3].x = h0; out4[3].y = h1; out4[3].z = h2; out4[3].w = i2f(child);
}
TGQ_HD void wide_store_empty(f4 * out4)
{
  out4[0] = mkf4(1.0f, 0.0f, 0.0f, 0.0f); out4[1] = mkf4(0.0f, 1.0f, 0.0f, 0.0f); out4[2] = mkf4(0.0f, 0.0f, 1.0f, 0.0f);
  out4[3] = mkf4(-1e30f, -1e30f, -1e30f, i2f(TGQ_WIDE_EMPTY));
}
// Lower bound (deflated) of the squared distance from p to anything in the child, the ordering key (squared distance to the
// box centre) and the link.
TGQ_HD void wide_child_bounds(f4 A0, f4 A1, f4 B0, f4 B1, v3 p, float mq, float & lb, float & key, int & link)
{
  float u0 = ffma(A0.x, p.x, ffma(A0.y, p.y, ffma(A0.z, p.z, A0.w)));
  float u1 = ffma(A1.x, p.x, ffma(A1.y, p.y, ffma(A1.z, p.z, A1.w)));
  float u2 = ffma(B0.x, p.x, ffma(B0.y, p.y, ffma(B0.z, p.z, B0.w)));
  float x = fmaxf(fabsf(u0) - (B1.x + mq), 0.0f), y = fmaxf(fabsf(u1) - (B1.y + mq), 0.0f), z = fmaxf(fabsf(u2) - (B1.z + mq), 0.0f);
  lb = ffma(x, x, ffma(y, y, z * z)) * TGQ_LB_DEFLATE;
  key = ffma(u0, u0, ffma(u1, u1, u2 * u2));
  link = f2i(B1.w);
}

// An oriented box that no query can reach (the missing second child of a one-element structure).
TGQ_HD void obb_store_empty(int child, f4 * out4)
{
#ifdef TGQ_NODE64
  // a box at an unreachable distance (the packed extents cannot be negative); its link is an empty leaf anyway
  out4[0].x = 1e30f; out4[0].y = 1e30f; out4[0].z = 1e30f; out4[0].w = i2f(child);
  out4[1].x = 0.0f; out4[1].y = 0.0f; out4[1].z = 0.0f; out4[1].w = 0.0f;
  return;
#endif
  out4[0].x = 0.0f; out4[0].y = 0.0f; out4[0].z = 0.0f; out4[0].w = -1e30f;
  out4[1].x = 1.0f; out4[1].y = 0.0f; out4[1].z = 0.0f; out4[1].w = 0.0f;
  out4[2].x = 0.0f; out4[2].y = 1.0f; out4[2].z = 0.0f; out4[2].w = i2f(child);
}

#ifdef TGQ_NODE64
// The 64-byte record from the node's (quantised) frame and the two unpacked child halves written by obb_acc_store / obb_store_empty.
TGQ_HD void obb_pack_node64(ObbAcc const & fr, f4 const * half0, f4 const * half1, f4 * out4)
{
  out4[0] = half0[0]; out4[1] = half1[0];
  out4[2].x = i2f((int)fr.wn); out4[2].y = i2f((int)fr.wr); out4[2].z = i2f((int)fr.wx);
  out4[2].w = pack_bf16_up(half0[1].x, half0[1].y);
  out4[3].x = pack_bf16_up(half0[1].z, half1[1].x); out4[3].y = pack_bf16_up(half1[1].y, half1[1].z); out4[3].z = 0.0f; out4[3].w = 0.0f;
}
// The frame a node record carries for the normal sum `nsum` of the node's triangles (what obb_acc_begin measures along).
TGQ_HD ObbAcc obb_node_frame64(v3 nsum) { ObbAcc a; obb_acc_begin(a, nsum); return a; }
#endif

// ---------------------------------------------------------------------------------------------------------------------
// Step protocol.  Every traversal advances in steps of ONE record fetch followed by arithmetic on the fetched registers:
//   *_fetch(s, b, r)      loads the record the state is waiting for -- an inner node (cur >= 0) or the pending leaf
//                         element(s) (TGQ_ST_LEAF) -- into r[]; both kinds have the same fetch shape (FETCH float4), so a
//                         warp issues ONE load phase per iteration for all its lanes, whatever they are waiting for;
//   *_node_compute(s, b, r) / *_leaf_compute(s, b, r)   consume r[] and move the state on.
// ncu on the previous layout (node step and leaf step each with their own dependent load) showed the kernels latency and
// issue bound at once: 54 % issue utilisation with ~300 dependent fetches per closest-point query.
// ---------------------------------------------------------------------------------------------------------------------
#define TGQ_FETCH_TRI 8        // closest point on triangle structures: 128-byte ObbNode / 128-byte TriRec
#define TGQ_FETCH_BOX 4        // everything else: 64-byte NodeRec / up to 4 points / the 48 used bytes of a TriRec

// ---------------------------------------------------------------------------------------------------------------------
// Closest element.  best_d2 starts as the (exclusive) squared bound or +inf, best_idx as -1.
// ---------------------------------------------------------------------------------------------------------------------
// Element filter of the reference's filtered nearest-neighbour queries on point structures: BvhN::pushFilter +
// elementPassesFilters (BvhN.hpp:1005-1032,1674-1683) with the NNFilter of Tools/Register/Register.cpp:92-101: an element e is
// admissible iff (labels unused or label[e] == label) and (normals unused or n . n_e >= min_dot; Register uses 0).  The dot
// product is the reference's (Eigen order); filtered-out elements are skipped before any distance is computed, as there.
struct ElemFilterD { v3 n; float min_dot; int label; int mode; };      // mode bit 0: normals, bit 1: labels
TGQ_HD ElemFilterD no_filter() { ElemFilterD f; f.n = mk3(0.0f, 0.0f, 0.0f); f.min_dot = 0.0f; f.label = 0; f.mode = 0; return f; }
TGQ_HD bool filtered_out(BvhView const & b, int k, ElemFilterD const & f)
{
  if (f.mode == 0) return false;
  if (f.mode & 2)
  {
#ifdef __CUDA_ARCH__
    int l = __ldg(b.lab + k);
#else
    int l = b.lab[k];
#endif
    if (l != f.label) return true;
  }
  if (f.mode & 1) { f4 nj = ldq(b.nrm + k); if (dot3(f.n, mk3(nj.x, nj.y, nj.z)) < f.min_dot) return true; }
  return false;
}

// ---------------------------------------------------------------------------------------------------------------------
// What BvhN::closestPair does with a non-negative distance bound (BvhN.hpp:1081-1114; established by running the reference's own
// code against the first version of this engine, which took the bound at face value):
//   (1) the query is rejected outright when the reference's root box -- the union of all element AABBs, scaleCentered(1.05)
//       (BvhN.hpp:1437-1439) -- is farther than the bound: fp32 AxisAlignedBoxN::squaredDistance, compared in double with
//       dist_bound^2 (:1093-1098);
//   (2) otherwise accelerationBound() returns the caller's bound unchanged (:2369-2370), closestPair takes its `accel_bound >= 0`
//       branch and WIDENS the bound by max(0.001 * |extent of the root box|, 100 eps) (:1101-1107); ClosestPairFunctor squares the
//       widened bound (:446-447) and accepts d2 < that, strictly, in double.
// So an element up to the fudge beyond dist_bound is still returned, unless (1) fires.  kClosestPairs has no widening (:478).
// Both the root box and the fudge depend on the element set only, not on the tree, so the LBVH reproduces them exactly.
// ---------------------------------------------------------------------------------------------------------------------
struct BoundGate { v3 lo, hi; float reject; int on; };      // reference root box; a query with fp32 box distance > reject finds nothing
TGQ_HD BoundGate no_gate() { BoundGate g; g.lo = g.hi = mk3(0.0f, 0.0f, 0.0f); g.reject = 0.0f; g.on = 0; return g; }
// AxisAlignedBoxN::squaredDistance(point) in the reference's operation order (AxisAlignedBoxN.hpp:429-442)
TGQ_HD float ref_box_point_sqdist(v3 lo, v3 hi, v3 p) { return sqn3(max3(sub3(max3(lo, p), min3(hi, p)), mk3(0.0f, 0.0f, 0.0f))); }
// The exclusive squared bound of one query: `ub` unless the gate rejects the query (then nothing can be accepted).
TGQ_HD float gated_ub(BoundGate const & g, v3 p, float ub) { return (g.on && ref_box_point_sqdist(g.lo, g.hi, p) > g.reject) ? -1.0f : ub; }

// Host only.  Smallest float >= x (x >= 0): d2 < ub  <=>  (double)d2 < x for every float d2.
inline float float_at_least(double x)
{
  if (x > (double)FLT_MAX) return INFINITY;
  float f = (float)x;
  if ((double)f < x) f = nextafterf(f, INFINITY);
  return f;
}
// Host side of the above: from the structure's exact element bounds [lo, hi] and the caller's dist_bound, the exclusive squared
// bound `ub` the kernels test d2 against and the gate.  dist_bound < 0 (or NaN): unbounded, gate off.
inline void closest_bound_gate(float const lo[3], float const hi[3], double dist_bound, float & ub, BoundGate & g)
{
  g = no_gate(); ub = INFINITY;
  if (!(dist_bound >= 0)) return;
  // AxisAlignedBoxN::scaleCentered(1.05) (AxisAlignedBoxN.hpp:316-324): c = 0.5 * (hi + lo); lo' = s * (lo - c) + c
  float rlo[3], rhi[3];
  for (int a = 0; a < 3; ++a)
  {
    float c = 0.5f * (hi[a] + lo[a]);
    float l = 1.05f * (lo[a] - c); l = l + c;
    float u = 1.05f * (hi[a] - c); u = u + c;
    rlo[a] = l; rhi[a] = u;
  }
  float ex = rhi[0] - rlo[0], ey = rhi[1] - rlo[1], ez = rhi[2] - rlo[2];
  float xx = ex * ex, yy = ey * ey, zz = ez * ez;
  float t = yy + zz; float sq = xx + t;
  float diag = sqrtf(sq);
  double fudge = 0.001 * (double)diag;
  double floor_fudge = (double)(100 * FLT_EPSILON);
  if (fudge < floor_fudge) fudge = floor_fudge;
  double eff = dist_bound + fudge;
  ub = float_at_least(eff * eff);
  double b2 = dist_bound * dist_bound;
  float rj = b2 > (double)FLT_MAX ? FLT_MAX : (float)b2;
  if ((double)rj > b2) rj = nextafterf(rj, -INFINITY);        // largest float <= dist_bound^2: lb > b2 <=> lb > rj for a float lb
  g.lo = mk3(rlo[0], rlo[1], rlo[2]); g.hi = mk3(rhi[0], rhi[1], rhi[2]); g.reject = rj; g.on = 1;
}

template <bool TRI>
struct ClosestState
{
  v3 p; float mq; int skip;
  float best_d2; int best_idx; v3 best_cp;
  int best_k;              // Morton position of the best element (seed for the lane's next query)
  ElemFilterD f;           // point structures only
  Trav t;
};
template <bool TRI> struct ClosestFetch { static constexpr int N = TRI ? TGQ_FETCH_TRI : TGQ_FETCH_BOX; };

// seed_k >= 0: Morton position of an element to evaluate before the traversal starts (the answer of a neighbouring query
// of the Morton-ordered batch): it is an ordinary element evaluated early, so the result is unchanged, but the
// traversal starts with a near-final bound (host counters: -13 % nodes, -30 % exact tests on the bench workload).
template <bool TRI>
TGQ_HD void closest_begin(ClosestState<TRI> & s, BvhView const & b, float scene_pad, v3 p, int skip, float bound_ub, int seed_k = -1)
{
  s.p = p; s.skip = skip; s.mq = TRI ? query_margin(p, scene_pad) : 0.0f;
  s.best_d2 = bound_ub; s.best_idx = -1; s.best_cp = mk3(0.0f, 0.0f, 0.0f); s.best_k = -1; s.f = no_filter();
  trav_begin(s.t, b.n_nodes);
  if (seed_k >= 0 && seed_k < b.n_prims && s.t.cur == 0)
  {
    trav_push(s.t, 0, 0.0f);                                   // the root waits on the stack
    trav_enter(s.t, ~((seed_k << TGQ_LEAF_BITS) | 1));         // a one-element "leaf"
  }
}

template <bool TRI>
TGQ_HD void closest_fetch(ClosestState<TRI> const & s, BvhView const & b, f4 * r)
{
  f4 const * q;
  if (TRI) q = s.t.cur >= 0 ? (f4 const *)(b.obbs + s.t.cur) : (f4 const *)(b.tris + s.t.k);
  else q = s.t.cur >= 0 ? (f4 const *)(b.nodes + s.t.cur) : b.pts + s.t.k;
  if (TRI)
  {
    // 32-byte aligned records: a node is 96 bytes (three 256-bit loads), a triangle 128 (four)
#ifdef TGQ_NODE64
    ldq2(q, r[0], r[1]); ldq2(q + 2, r[2], r[3]);
    if (s.t.cur < 0) { ldq2(q + 4, r[4], r[5]); ldq2(q + 6, r[6], r[7]); }
#else
    ldq2(q, r[0], r[1]); ldq2(q + 2, r[2], r[3]); ldq2(q + 4, r[4], r[5]);
    if (s.t.cur < 0) ldq2(q + 6, r[6], r[7]);
#endif
  }
  else { for (int i = 0; i < ClosestFetch<TRI>::N; ++i) r[i] = ldq(q + i); }                        // point leaves start anywhere
}

// Descends into / defers the two children given their lower bounds, ordering keys and links.
TGQ_HD void trav_descend(Trav & t, float lb0, float lb1, float key0, float key1, int c0, int c1, float limit)
{
  bool go0 = lb0 <= limit, go1 = lb1 <= limit;
  if (go0 && go1)
  {
    bool swap = key1 < key0;
    trav_push(t, swap ? c0 : c1, swap ? lb0 : lb1);
    trav_enter(t, swap ? c1 : c0);
  }
  else if (go0 || go1) trav_enter(t, go0 ? c0 : c1);
  else trav_pop(t);
}

template <bool TRI>
TGQ_HD void closest_node_compute(ClosestState<TRI> & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  float lb0, lb1, key0, key1; int c0, c1;
  if (TRI)
  {
#ifdef TGQ_NODE64
    obb_node_bounds64(r, s.p, s.mq, lb0, key0, c0, lb1, key1, c1);
#else
    obb_bounds(r[0], r[1], r[2], s.p, s.mq, lb0, key0);
    obb_bounds(r[3], r[4], r[5], s.p, s.mq, lb1, key1);
    c0 = f2i(r[2].w); c1 = f2i(r[5].w);
#endif
  }
  else
  {
    NodeRec nd; nd.q0 = r[0]; nd.q1 = r[1]; nd.q2 = r[2]; nd.q3 = r[3];
    v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
    lb0 = box_point_sqdist(lo0, hi0, s.p) * TGQ_LB_DEFLATE; lb1 = box_point_sqdist(lo1, hi1, s.p) * TGQ_LB_DEFLATE;
    key0 = lb0; key1 = lb1;
    c0 = f2i(nd.q3.x); c1 = f2i(nd.q3.y);
  }
  trav_descend(s.t, lb0, lb1, key0, key1, c0, c1, s.best_d2);
}

// One pending triangle (triangle structures) or up to four pending points of the leaf (point structures).
template <bool TRI>
TGQ_HD void closest_leaf_compute(ClosestState<TRI> & s, BvhView const & b, f4 const * r)
{
  if (TRI)
  {
    ++s.t.k;
    f4 a = r[0], bq = r[1], c = r[2], d = r[3];
    v3 v0 = mk3(a.x, a.y, a.z), v1 = mk3(bq.x, bq.y, bq.z), v2 = mk3(c.x, c.y, c.z), n = mk3(d.x, d.y, d.z);
    // screen: plane distance and lateral distance to the triangle's bounding circle
    v3 cen = mul3(add3(add3(v0, v1), v2), 0.33333334f);
    float sd = fmaxf(fabsf(fdot3(s.p, n) - c.w) - s.mq, 0.0f);
    v3 pc = sub3(s.p, cen);
    v3 perp = sub3(pc, mul3(n, fdot3(n, pc)));
    float lat = fmaxf(sqrtf(fdot3(perp, perp)) - d.w - s.mq, 0.0f);
    int idx = f2i(a.w);
    if ((sd * sd + lat * lat) * TGQ_LB_DEFLATE <= s.best_d2 && idx != s.skip)
    {
      TGQ_COUNT_PRIM();
      v3 cp;
      float d2 = tri_point_sqdist_pre(v0, v1, v2, n, c.w, f2i(bq.w), r[4], r[5], r[6], r[7], s.p, cp);
      if (d2 < s.best_d2 || (d2 == s.best_d2 && idx < s.best_idx)) { s.best_d2 = d2; s.best_idx = idx; s.best_cp = cp; s.best_k = s.t.k - 1; }
    }
  }
  else
  {
    int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
    for (int j = 0; j < TGQ_FETCH_BOX; ++j)
    {
      if (j >= cnt) break;
      TGQ_COUNT_PRIM();
      f4 e = r[j];
      int idx = f2i(e.w); v3 cp = mk3(e.x, e.y, e.z);
      float d2 = sqn3(sub3(cp, s.p));
      if (idx == s.skip || filtered_out(b, s.t.k + j, s.f)) continue;
      if (d2 < s.best_d2 || (d2 == s.best_d2 && idx < s.best_idx)) { s.best_d2 = d2; s.best_idx = idx; s.best_cp = cp; s.best_k = s.t.k + j; }
    }
    s.t.k += cnt;
  }
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}

// Whole query in one go (host emulation; single-lane device uses).
template <bool TRI>
TGQ_HD void closest_query(BvhView const & b, float scene_pad, v3 p, int skip, float & best_d2, int & best_idx, v3 & best_cp, int * seed_k = nullptr,
                          ElemFilterD const * filter = nullptr)
{
  ClosestState<TRI> s; TravStack st; trav_bind(s.t, st);
  closest_begin(s, b, scene_pad, p, skip, best_d2, seed_k ? *seed_k : -1);
  if (filter) s.f = *filter;
  f4 r[ClosestFetch<TRI>::N];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.best_d2); continue; }
    closest_fetch(s, b, r);
    if (s.t.cur >= 0) closest_node_compute(s, b, r); else closest_leaf_compute(s, b, r);
  }
  best_d2 = s.best_d2; best_idx = s.best_idx; best_cp = s.best_cp;
  if (seed_k) *seed_k = s.best_k;
}

// The closest-element query over the 4-wide nodes, one query at a time: the sequential statement of what k_closest_wide does with
// four lanes per query (same child bounds, same ordering -- passing children by increasing centre distance, child slot as the tie
// break --, same pruning rule, same leaf arithmetic).  Drives the host emulation of the test suite.
template <int DUMMY = 0>
TGQ_HD void closest_query_wide(BvhView const & b, float scene_pad, v3 p, int skip, float & best_d2, int & best_idx, v3 & best_cp, int * seed_k = nullptr)
{
  ClosestState<true> s; TravStack st; trav_bind(s.t, st);
  closest_begin(s, b, scene_pad, p, skip, best_d2, -1);
  s.t.cur = TGQ_ST_DONE;
  f4 r[TGQ_FETCH_TRI];
  int sn[4 * TGQ_STACK]; float sk[4 * TGQ_STACK]; int sp = 0;
  // leaf helper: all elements of a leaf link
  #define TGQ_WIDE_LEAF(code) { int c_ = ~(code); s.t.k = c_ >> TGQ_LEAF_BITS; s.t.kend = s.t.k + (c_ & TGQ_LEAF_MASK); \
                                while (s.t.k < s.t.kend) { s.t.cur = TGQ_ST_LEAF; closest_fetch(s, b, r); closest_leaf_compute(s, b, r); } }
  if (b.n_wide > 0)
  {
    if (seed_k && *seed_k >= 0 && *seed_k < b.n_prims) { TGQ_WIDE_LEAF(~((*seed_k << TGQ_LEAF_BITS) | 1)) }
    sn[sp] = 0; sk[sp] = 0.0f; ++sp;
    while (sp > 0)
    {
      --sp;
      int code = sn[sp]; float klb = sk[sp];
      if (!(klb <= s.best_d2)) continue;
      if (code < 0) { TGQ_WIDE_LEAF(code) continue; }
      TGQ_COUNT_NODE();
      f4 const * nd = (f4 const *)(b.wide + code);
      float lb[4], key[4]; int link[4]; bool pass[4];
      for (int i = 0; i < 4; ++i)
      {
        wide_child_bounds(ldq(nd + 2 * i), ldq(nd + 2 * i + 1), ldq(nd + 8 + 2 * i), ldq(nd + 8 + 2 * i + 1), p, s.mq, lb[i], key[i], link[i]);
        pass[i] = link[i] != TGQ_WIDE_EMPTY && lb[i] <= s.best_d2;
      }
      // push far to near: the nearest passing child is popped next
      int order[4], m = 0;
      for (int i = 0; i < 4; ++i) if (pass[i]) order[m++] = i;
      for (int x = 1; x < m; ++x)
        for (int y = x; y > 0 && (key[order[y]] < key[order[y - 1]] || (key[order[y]] == key[order[y - 1]] && order[y] < order[y - 1])); --y)
        { int t_ = order[y]; order[y] = order[y - 1]; order[y - 1] = t_; }
      for (int x = m - 1; x >= 0; --x) { sn[sp] = link[order[x]]; sk[sp] = lb[order[x]]; ++sp; }
    }
  }
  #undef TGQ_WIDE_LEAF
  best_d2 = s.best_d2; best_idx = s.best_idx; best_cp = s.best_cp;
  if (seed_k) *seed_k = s.best_k;
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

// ---------------------------------------------------------------------------------------------------------------------
// Candidate buffer: the k-NN / graph "heap" of the warp-cooperative kernels (k_knn_buf, k_ball_graph_buf).  ncu on the sorted-set
// kernels (profiles/r02_p_knn_ball_graph_ncu.csv): 5.2 of 32 lanes active per instruction -- every accepted candidate runs a
// data-dependent insertion loop inside divergent leaf code.  Here a lane only APPENDS accepted candidates, packed as
// (key bits << 32 | element index) -- one store, no loop; the packed value orders exactly like the canonical (key, index) order because
// keys are non-negative floats.  When fewer than TGQ_FETCH_BOX free slots remain, the buffer is COMPACTED: the k smallest entries
// are kept, sorted ascending, and the k-th becomes the lane's acceptance / pruning threshold.  On the device the compaction of one
// lane's buffer is done by the whole warp (rank sort with broadcast reads, tgq_kernels.cuh); `compact_serial` is the same operation
// in one thread (host emulation, k_self_nn-style single-thread callers).  The result -- the k smallest (key, index) pairs of all
// candidates ever offered with key <= threshold at the time -- does not depend on when compactions happen: an entry is only ever
// dropped when k smaller ones are present.
// Row layout: entry i of the lane with swizzle `lane` lives at row[(i & ~15) | ((i + lane) & 15)], so that lanes appending at equal
// fill levels hit different banks and a whole-warp read of one row is conflict free.
// ---------------------------------------------------------------------------------------------------------------------
TGQ_HD unsigned long long cand_pack(float key, int idx) { return ((unsigned long long)(unsigned)f2i(key) << 32) | (unsigned long long)(unsigned)idx; }
TGQ_HD float cand_key(unsigned long long e) { return i2f((int)(unsigned)(e >> 32)); }
TGQ_HD int cand_idx(unsigned long long e) { return (int)(unsigned)(e & 0xffffffffull); }
TGQ_HD int cand_cap(int k) { int room = k < 8 ? 8 : (k > 24 ? 24 : k); return (k + room + 15) & ~15; }      // k = 16: 32 slots
struct CandBuf
{
  unsigned long long * row; int lane, cap, k, n;
  TGQ_HD void init(unsigned long long * row_, int lane_, int k_) { row = row_; lane = lane_; k = k_; cap = cand_cap(k_); n = 0; }
  TGQ_HD int phys(int i) const { return (i & ~15) | ((i + lane) & 15); }
  TGQ_HD void append(float key, int idx) { TGQ_COUNT_APPEND(); row[phys(n)] = cand_pack(key, idx); ++n; }
  TGQ_HD bool wants_compact() const { return n > cap - TGQ_FETCH_BOX; }
  TGQ_HD unsigned long long at(int i) const { return row[phys(i)]; }
  // keeps the min(n, k) smallest entries, ascending; returns true (and the k-th key) when the buffer then holds k entries
  TGQ_HD bool compact_serial(float & kth)
  {
    TGQ_COUNT_COMPACT();
    for (int i = 1; i < n; ++i)
    {
      unsigned long long x = row[phys(i)]; int j = i;
      while (j > 0 && row[phys(j - 1)] > x) { row[phys(j)] = row[phys(j - 1)]; --j; }
      row[phys(j)] = x;
    }
    if (n > k) n = k;
    if (n == k && k > 0) { kth = cand_key(row[phys(k - 1)]); return true; }
    return false;
  }
};

// Fetch of the AABB tree over points: a 64-byte node or the next (up to) four 16-byte points of the pending leaf.
TGQ_HD void box_fetch(Trav const & t, BvhView const & b, f4 * r)
{
  f4 const * q = t.cur >= 0 ? (f4 const *)(b.nodes + t.cur) : b.pts + t.k;
  for (int i = 0; i < TGQ_FETCH_BOX; ++i) r[i] = ldq(q + i);
}
TGQ_HD void box_node_bounds(f4 const * r, v3 p, float & lb0, float & lb1, int & c0, int & c1)
{
  NodeRec nd; nd.q0 = r[0]; nd.q1 = r[1]; nd.q2 = r[2]; nd.q3 = r[3];
  v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
  lb0 = box_point_sqdist(lo0, hi0, p) * TGQ_LB_DEFLATE; lb1 = box_point_sqdist(lo1, hi1, p) * TGQ_LB_DEFLATE;
  c0 = f2i(nd.q3.x); c1 = f2i(nd.q3.y);
}

// k nearest points to p: k smallest (d2, idx); d2 < bound_ub (exclusive) when bounded.
struct KnnState
{
  v3 p; int skip; float bound_ub;
  float worst;     // prune radius: k-th best once the set is full, else the bound (or +inf)
  ElemFilterD f;
  SortedSet set;
  Trav t;
};

TGQ_HD void knn_begin(KnnState & s, BvhView const & b, v3 p, int skip, float bound_ub)
{
  s.p = p; s.skip = skip; s.bound_ub = bound_ub; s.worst = bound_ub; s.set.n = 0; s.f = no_filter();
  trav_begin(s.t, s.set.cap > 0 ? b.n_nodes : 0);
}
TGQ_HD void knn_fetch(KnnState const & s, BvhView const & b, f4 * r) { box_fetch(s.t, b, r); }

TGQ_HD void knn_node_compute(KnnState & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  float lb0, lb1; int c0, c1; box_node_bounds(r, s.p, lb0, lb1, c0, c1);
  trav_descend(s.t, lb0, lb1, lb0, lb1, c0, c1, s.worst);
}

TGQ_HD void knn_leaf_compute(KnnState & s, BvhView const & b, f4 const * r)
{
  int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
  for (int j = 0; j < TGQ_FETCH_BOX; ++j)
  {
    if (j >= cnt) break;
    TGQ_COUNT_PRIM();
    f4 e = r[j];
    int idx = f2i(e.w);
    float d2 = sqn3(sub3(mk3(e.x, e.y, e.z), s.p));
    if (idx == s.skip || !(d2 < s.bound_ub) || filtered_out(b, s.t.k + j, s.f)) continue;
    if (s.set.accepts(d2, idx))
    {
      s.set.insert(d2, idx);
      if (s.set.full()) s.worst = s.set.worst_key();
    }
  }
  s.t.k += cnt;
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}

TGQ_HD void knn_query(BvhView const & b, v3 p, int skip, float bound_ub, SortedSet & set, ElemFilterD const * filter = nullptr)
{
  KnnState s; TravStack st; trav_bind(s.t, st); s.set = set;
  knn_begin(s, b, p, skip, bound_ub);
  if (filter) s.f = *filter;
  f4 r[TGQ_FETCH_BOX];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.worst); continue; }
    knn_fetch(s, b, r);
    if (s.t.cur >= 0) knn_node_compute(s, b, r); else knn_leaf_compute(s, b, r);
  }
  set.n = s.set.n;
}

// The same query on a candidate buffer (k_knn_buf): candidates with d2 <= worst are appended, `worst` tightens at compactions.
struct KnnBufState
{
  v3 p; int skip; float bound_ub; float worst; ElemFilterD f; CandBuf buf; Trav t;
};
TGQ_HD void knnb_begin(KnnBufState & s, BvhView const & b, v3 p, int skip, float bound_ub)
{
  s.p = p; s.skip = skip; s.bound_ub = bound_ub; s.worst = bound_ub; s.buf.n = 0; s.f = no_filter();
  trav_begin(s.t, s.buf.k > 0 ? b.n_nodes : 0);
}
TGQ_HD void knnb_node_compute(KnnBufState & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  float lb0, lb1; int c0, c1; box_node_bounds(r, s.p, lb0, lb1, c0, c1);
  trav_descend(s.t, lb0, lb1, lb0, lb1, c0, c1, s.worst);
}
TGQ_HD void knnb_leaf_compute(KnnBufState & s, BvhView const & b, f4 const * r)
{
  int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
  // straight-line over the four fetched slots (no early exit): the lanes of a warp that are at a leaf stay converged
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < TGQ_FETCH_BOX; ++j)
  {
    f4 e = r[j];
    int idx = f2i(e.w);
    float d2 = sqn3(sub3(mk3(e.x, e.y, e.z), s.p));
    bool ok = j < cnt && idx != s.skip && d2 < s.bound_ub && d2 <= s.worst;
    if (j < cnt) { TGQ_COUNT_PRIM(); }
    if (ok && s.f.mode != 0) ok = !filtered_out(b, s.t.k + j, s.f);
    if (ok) s.buf.append(d2, idx);
  }
  s.t.k += cnt;
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}
// One-thread driver (host emulation): out_n entries ascending in buf.
TGQ_HD void knn_query_buf(BvhView const & b, v3 p, int skip, float bound_ub, CandBuf & buf, ElemFilterD const * filter = nullptr)
{
  KnnBufState s; TravStack st; trav_bind(s.t, st); s.buf = buf;
  knnb_begin(s, b, p, skip, bound_ub);
  if (filter) s.f = *filter;
  f4 r[TGQ_FETCH_BOX];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.worst); continue; }
    box_fetch(s.t, b, r);
    if (s.t.cur >= 0) knnb_node_compute(s, b, r);
    else { knnb_leaf_compute(s, b, r); if (s.buf.wants_compact()) { float kth; if (s.buf.compact_serial(kth)) s.worst = kth; } }
  }
  float kth; s.buf.compact_serial(kth);
  buf.n = s.buf.n;
}

// PointSet3::findSampleNeighbors for sample `self` (position c, normal nc): in each round every point j != self with
// sq(x_j - c) <= r2 and (no normals or n_c . n_j >= min_dot) is offered with key fastSqrt(d2); the set keeps the
// max_degree smallest (key, idx); the radius doubles for up to `rounds` rounds while fewer than min_degree neighbours were
// found (PointSet3.hpp:660-681).  The ball/box test uses the reference arithmetic on padded boxes (a superset).
struct BallState
{
  v3 c, nc; bool has_n; float min_dot; int self; float radius, r2; int round, rounds, min_degree;
  SortedSet set;
  Trav t;
};

TGQ_HD void ball_begin(BallState & s, BvhView const & b, v3 c, v3 nc, bool has_n, float min_dot, int self, float init_radius, int rounds,
                       int min_degree)
{
  s.c = c; s.nc = nc; s.has_n = has_n; s.min_dot = min_dot; s.self = self; s.radius = init_radius; s.r2 = init_radius * init_radius;
  s.round = 0; s.rounds = rounds; s.min_degree = min_degree; s.set.n = 0;
  trav_begin(s.t, b.n_nodes);
}
TGQ_HD void ball_fetch(BallState const & s, BvhView const & b, f4 * r) { box_fetch(s.t, b, r); }

TGQ_HD void ball_node_compute(BallState & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  float lb0, lb1; int c0, c1; box_node_bounds(r, s.c, lb0, lb1, c0, c1);
  // keys 0: every deferred child that intersected the ball stays wanted (the radius does not shrink within a round)
  trav_descend(s.t, lb0, lb1, 0.0f, 0.0f, c0, c1, s.r2);
}

TGQ_HD void ball_leaf_compute(BallState & s, BvhView const & b, f4 const * r)
{
  int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
  for (int j = 0; j < TGQ_FETCH_BOX; ++j)
  {
    if (j >= cnt) break;
    TGQ_COUNT_PRIM();
    f4 e = r[j];
    int idx = f2i(e.w);
    if (idx == s.self) continue;
    v3 diff = sub3(mk3(e.x, e.y, e.z), s.c);
    float d2 = sqn3(diff);
    if (!(d2 <= s.r2)) continue;
    if (s.has_n)
    {
      f4 nj = ldq(b.nrm + s.t.k + j);
      if (dot3(s.nc, mk3(nj.x, nj.y, nj.z)) < s.min_dot) continue;
    }
    float sep = fast_sqrt(d2);
    if (s.set.accepts(sep, idx)) s.set.insert(sep, idx);
  }
  s.t.k += cnt;
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}

// Called when a round's traversal is done: returns true when the query is complete, else starts the next round.
TGQ_HD bool ball_round_done(BallState & s, BvhView const & b)
{
  ++s.round;
  if (s.set.n >= s.min_degree || s.round >= s.rounds) return true;
  s.radius *= 2.0f; s.r2 = s.radius * s.radius; s.set.n = 0;
  trav_begin(s.t, b.n_nodes);
  return s.t.cur == TGQ_ST_DONE;
}

TGQ_HD void ball_graph_query(BvhView const & b, v3 c, v3 nc, bool has_n, float min_dot, int self, float init_radius,
                             int rounds, int min_degree, SortedSet & set)
{
  BallState s; TravStack st; trav_bind(s.t, st); s.set = set;
  ball_begin(s, b, c, nc, has_n, min_dot, self, init_radius, rounds, min_degree);
  f4 r[TGQ_FETCH_BOX];
  for (;;)
  {
    while (s.t.cur != TGQ_ST_DONE)
    {
      if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.r2); continue; }
      ball_fetch(s, b, r);
      if (s.t.cur >= 0) ball_node_compute(s, b, r); else ball_leaf_compute(s, b, r);
    }
    if (ball_round_done(s, b)) break;
  }
  set.n = s.set.n;
}

// The sample graph on a candidate buffer (k_ball_graph_buf): every admissible point of the ball is appended with key fastSqrt(d2);
// once a compaction has left max_degree entries, candidates with a larger key than the last of them are not offered any more (they
// could never be among the max_degree smallest).  Rounds as above.
struct BallBufState
{
  v3 c, nc; bool has_n; float min_dot; int self; float radius, r2; int round, rounds, min_degree; float worst_key;
  CandBuf buf; Trav t;
};
TGQ_HD void ballb_begin(BallBufState & s, BvhView const & b, v3 c, v3 nc, bool has_n, float min_dot, int self, float init_radius, int rounds,
                        int min_degree)
{
  s.c = c; s.nc = nc; s.has_n = has_n; s.min_dot = min_dot; s.self = self; s.radius = init_radius; s.r2 = init_radius * init_radius;
  s.round = 0; s.rounds = rounds; s.min_degree = min_degree; s.buf.n = 0; s.worst_key = INFINITY;
  trav_begin(s.t, b.n_nodes);
}
TGQ_HD void ballb_node_compute(BallBufState & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  float lb0, lb1; int c0, c1; box_node_bounds(r, s.c, lb0, lb1, c0, c1);
  trav_descend(s.t, lb0, lb1, 0.0f, 0.0f, c0, c1, s.r2);
}
TGQ_HD void ballb_leaf_compute(BallBufState & s, BvhView const & b, f4 const * r)
{
  int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < TGQ_FETCH_BOX; ++j)
  {
    f4 e = r[j];
    int idx = f2i(e.w);
    v3 diff = sub3(mk3(e.x, e.y, e.z), s.c);
    float d2 = sqn3(diff);
    bool ok = j < cnt && idx != s.self && d2 <= s.r2;
    if (j < cnt) { TGQ_COUNT_PRIM(); }
    if (ok && s.has_n)
    {
      f4 nj = ldq(b.nrm + s.t.k + j);
      ok = !(dot3(s.nc, mk3(nj.x, nj.y, nj.z)) < s.min_dot);
    }
    float sep = fast_sqrt(d2);
    if (ok && sep <= s.worst_key) s.buf.append(sep, idx);
  }
  s.t.k += cnt;
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}
// Called (after a compaction of the finished round's buffer) when a round's traversal is done: true = the query is complete.
TGQ_HD bool ballb_round_done(BallBufState & s, BvhView const & b)
{
  ++s.round;
  if (s.buf.n >= s.min_degree || s.round >= s.rounds) return true;
  s.radius *= 2.0f; s.r2 = s.radius * s.radius; s.buf.n = 0; s.worst_key = INFINITY;
  trav_begin(s.t, b.n_nodes);
  return s.t.cur == TGQ_ST_DONE;
}
TGQ_HD void ball_graph_query_buf(BvhView const & b, v3 c, v3 nc, bool has_n, float min_dot, int self, float init_radius, int rounds, int min_degree,
                                 CandBuf & buf)
{
  BallBufState s; TravStack st; trav_bind(s.t, st); s.buf = buf;
  ballb_begin(s, b, c, nc, has_n, min_dot, self, init_radius, rounds, min_degree);
  f4 r[TGQ_FETCH_BOX];
  for (;;)
  {
    while (s.t.cur != TGQ_ST_DONE)
    {
      if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.r2); continue; }
      box_fetch(s.t, b, r);
      if (s.t.cur >= 0) ballb_node_compute(s, b, r);
      else { ballb_leaf_compute(s, b, r); if (s.buf.wants_compact()) { float kth; if (s.buf.compact_serial(kth)) s.worst_key = kth; } }
    }
    float kth; s.buf.compact_serial(kth);
    if (ballb_round_done(s, b)) break;
  }
  buf.n = s.buf.n;
}

// Ball range query on a point structure: BvhN::rangeQueryIndices / processRange<IntersectionTester>(Ball3(c, r)) (BvhN.hpp:1208-
// 1220,2118-2171; BallN.hpp:77-100): every element with sq(x_e - c) <= r * r (the reference's fp32 test).  Counts all of them and
// keeps the `cap` lowest element indices, ascending (SortedSet with a constant key orders by index alone).
struct RangeState
{
  v3 c; float r2; int self; int count;
  SortedSet set;
  Trav t;
};
TGQ_HD void range_begin(RangeState & s, BvhView const & b, v3 c, float radius, int self)
{
  s.c = c; s.r2 = radius * radius; s.self = self; s.count = 0; s.set.n = 0;
  trav_begin(s.t, (radius >= 0.0f) ? b.n_nodes : 0);
}
TGQ_HD void range_fetch(RangeState const & s, BvhView const & b, f4 * r) { box_fetch(s.t, b, r); }
TGQ_HD void range_node_compute(RangeState & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  float lb0, lb1; int c0, c1; box_node_bounds(r, s.c, lb0, lb1, c0, c1);
  trav_descend(s.t, lb0, lb1, 0.0f, 0.0f, c0, c1, s.r2);
}
TGQ_HD void range_leaf_compute(RangeState & s, BvhView const & b, f4 const * r)
{
  int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
  for (int j = 0; j < TGQ_FETCH_BOX; ++j)
  {
    if (j >= cnt) break;
    TGQ_COUNT_PRIM();
    f4 e = r[j];
    int idx = f2i(e.w);
    if (idx == s.self) continue;
    float d2 = sqn3(sub3(mk3(e.x, e.y, e.z), s.c));
    if (!(d2 <= s.r2)) continue;
    ++s.count;
    if (s.set.accepts(0.0f, idx)) s.set.insert(0.0f, idx);
  }
  s.t.k += cnt;
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}
TGQ_HD void range_query(BvhView const & b, v3 c, float radius, int self, SortedSet & set, int & count)
{
  RangeState s; TravStack st; trav_bind(s.t, st); s.set = set;
  range_begin(s, b, c, radius, self);
  f4 r[TGQ_FETCH_BOX];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.r2); continue; }
    range_fetch(s, b, r);
    if (s.t.cur >= 0) range_node_compute(s, b, r); else range_leaf_compute(s, b, r);
  }
  set.n = s.set.n; count = s.count;
}

// Unbounded range queries on a point structure: BvhN::processRange / rangeQuery / rangeQueryIndices<IntersectionTester>(range, ...)
// (BvhN.hpp:1208-1264,2118-2171) with a Ball3 or an AxisAlignedBox3 as the range.  The element tests are the reference's
// (IntersectionTester.hpp:105-125 -> BallN::contains, BallN.hpp:83: sq(x - c) <= r * r;  AxisAlignedBoxN::contains,
// AxisAlignedBoxN.hpp:401-412: lo <= x <= hi in every coordinate); the reference's shortcut for nodes wholly inside the range
// (BvhN.hpp:2129-2143) takes every element of such a node without testing it, which is the same set because a node's padded box
// contains its elements.  The result is the SET of elements in range; the reference reports it in its tree's traversal order,
// here it is canonical: ascending element index.  FILL == false counts, FILL == true writes the indices at out[n++].
template <bool BOX>
struct RangeAllState
{
  v3 a, b;          // ball: centre, (r * r, -, -);  box: lo, hi
  int count; int * out;
  Trav t;
};
template <bool BOX>
TGQ_HD void range_all_begin(RangeAllState<BOX> & s, BvhView const & bv, v3 a, v3 b, int * out, bool valid)
{ s.a = a; s.b = b; s.count = 0; s.out = out; trav_begin(s.t, valid ? bv.n_nodes : 0); }
template <bool BOX>
TGQ_HD void range_all_node_compute(RangeAllState<BOX> & s, BvhView const &, f4 const * r)
{
  TGQ_COUNT_NODE();
  NodeRec nd; nd.q0 = r[0]; nd.q1 = r[1]; nd.q2 = r[2]; nd.q3 = r[3];
  v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
  int c0 = f2i(nd.q3.x), c1 = f2i(nd.q3.y);
  float k0, k1;            // 0: the child's (padded) box meets the range, 1: it cannot
  if (BOX)
  {
    k0 = (lo0.x > s.b.x || hi0.x < s.a.x || lo0.y > s.b.y || hi0.y < s.a.y || lo0.z > s.b.z || hi0.z < s.a.z) ? 1.0f : 0.0f;
    k1 = (lo1.x > s.b.x || hi1.x < s.a.x || lo1.y > s.b.y || hi1.y < s.a.y || lo1.z > s.b.z || hi1.z < s.a.z) ? 1.0f : 0.0f;
  }
  else
  {
    k0 = box_point_sqdist(lo0, hi0, s.a) * TGQ_LB_DEFLATE <= s.b.x ? 0.0f : 1.0f;
    k1 = box_point_sqdist(lo1, hi1, s.a) * TGQ_LB_DEFLATE <= s.b.x ? 0.0f : 1.0f;
  }
  trav_descend(s.t, k0, k1, 0.0f, 0.0f, c0, c1, 0.5f);
}
template <bool BOX, bool FILL>
TGQ_HD void range_all_leaf_compute(RangeAllState<BOX> & s, BvhView const &, f4 const * r)
{
  int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
  for (int j = 0; j < TGQ_FETCH_BOX; ++j)
  {
    if (j >= cnt) break;
    TGQ_COUNT_PRIM();
    f4 e = r[j];
    bool in;
    if (BOX) in = !(e.x < s.a.x || e.x > s.b.x || e.y < s.a.y || e.y > s.b.y || e.z < s.a.z || e.z > s.b.z);
    else in = sqn3(sub3(mk3(e.x, e.y, e.z), s.a)) <= s.b.x;
    if (!in) continue;
    if (FILL) s.out[s.count] = f2i(e.w);
    ++s.count;
  }
  s.t.k += cnt;
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}
// Whole query in one go (host emulation).
template <bool BOX, bool FILL>
TGQ_HD int range_all_query(BvhView const & bv, v3 a, v3 b, int * out)
{
  RangeAllState<BOX> s; TravStack st; trav_bind(s.t, st);
  range_all_begin(s, bv, a, b, out, true);
  f4 r[TGQ_FETCH_BOX];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, 0.5f); continue; }
    box_fetch(s.t, bv, r);
    if (s.t.cur >= 0) range_all_node_compute(s, bv, r); else range_all_leaf_compute<BOX, FILL>(s, bv, r);
  }
  return s.count;
}

// Ball neighbourhood statistics on a point structure: BvhN::processRange<IntersectionTester>(Ball3(c, r), functor) (BvhN.hpp:1238-
// 1264,2118-2171) with the accumulating functors of the local surface features, all in one pass over the
// elements with sq(x_e - c) <= r * r:
//   acc[0]      n                                                  every functor's point count
//   acc[1]      sum |x_e - c|  (fp32 norm per element)             AverageDistance EuclideanCallback (AverageDistance.cpp:39-61)
//   acc[2]      #{e : n_e . qn > -1e-5}, acc[3..5] sum stableNormalized(x_e - c) over those
//                                                                  Curvature ProjectedCurvatureFunctor (Curvature.cpp:37-66)
//   acc[6..8]   sum (x_e - c), acc[9..14] sum (x_e - c)(x_e - c)^T (xx xy xz yy yz zz)
//                                                                  LocalPca: moments of the neighbourhood (LocalPca.cpp:32-47, PcaN.hpp:163-184)
// Per-element terms are the reference's fp32 expressions; the sums are taken in double, one lane per ball, in this tree's
// traversal order (the reference sums in ITS traversal order, in fp32 for the curvature offsets and the PCA moments: its own
// result is tree dependent at fp32 rounding level, which is the tolerance the tests state).
struct StatsState
{
  v3 c, qn; float r2; int has_qn;
  double acc[15];
  Trav t;
};
TGQ_HD void stats_begin(StatsState & s, BvhView const & b, v3 c, float radius, v3 qn, bool has_qn)
{
  s.c = c; s.qn = qn; s.has_qn = (has_qn && b.nrm) ? 1 : 0; s.r2 = radius * radius;
  for (int j = 0; j < 15; ++j) s.acc[j] = 0.0;
  trav_begin(s.t, (radius >= 0.0f) ? b.n_nodes : 0);
}
TGQ_HD void stats_fetch(StatsState const & s, BvhView const & b, f4 * r) { box_fetch(s.t, b, r); }
TGQ_HD void stats_node_compute(StatsState & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  float lb0, lb1; int c0, c1; box_node_bounds(r, s.c, lb0, lb1, c0, c1);
  trav_descend(s.t, lb0, lb1, 0.0f, 0.0f, c0, c1, s.r2);
}
TGQ_HD void stats_leaf_compute(StatsState & s, BvhView const & b, f4 const * r)
{
  int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
  for (int j = 0; j < TGQ_FETCH_BOX; ++j)
  {
    if (j >= cnt) break;
    TGQ_COUNT_PRIM();
    f4 e = r[j];
    v3 diff = sub3(mk3(e.x, e.y, e.z), s.c);
    float d2 = sqn3(diff);
    if (!(d2 <= s.r2)) continue;
    s.acc[0] += 1.0;
    s.acc[1] += (double)sqrtf(d2);
    if (s.has_qn)
    {
      f4 nj = ldq(b.nrm + s.t.k + j);
      if (dot3(mk3(nj.x, nj.y, nj.z), s.qn) > -1.0e-05f)
      {
        v3 o = stable_normalized3(diff);
        s.acc[2] += 1.0; s.acc[3] += (double)o.x; s.acc[4] += (double)o.y; s.acc[5] += (double)o.z;
      }
    }
    double dx = diff.x, dy = diff.y, dz = diff.z;
    s.acc[6] += dx; s.acc[7] += dy; s.acc[8] += dz;
    s.acc[9] += dx * dx; s.acc[10] += dx * dy; s.acc[11] += dx * dz; s.acc[12] += dy * dy; s.acc[13] += dy * dz; s.acc[14] += dz * dz;
  }
  s.t.k += cnt;
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}
TGQ_HD void stats_query(BvhView const & b, v3 c, float radius, v3 qn, bool has_qn, double * out16)
{
  StatsState s; TravStack st; trav_bind(s.t, st);
  stats_begin(s, b, c, radius, qn, has_qn);
  f4 r[TGQ_FETCH_BOX];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.r2); continue; }
    stats_fetch(s, b, r);
    if (s.t.cur >= 0) stats_node_compute(s, b, r); else stats_leaf_compute(s, b, r);
  }
  for (int j = 0; j < 15; ++j) out16[j] = s.acc[j];
  out16[15] = 0.0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Ray first hit.  best_t starts as max_time (negative = unbounded), best_idx as -1; best_k is the Morton position of
// the hit triangle (for its cached normal).  A hit is accepted iff t >= 0 and (unbounded or t < best_t) or it ties the
// current best with a lower index.  ANY == true stops at the first hit found (BvhN::rayIntersects returns the same
// boolean as a full first-hit search, BvhN.hpp:1266-1269).
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
  Trav t;
};

template <bool ANY>
TGQ_HD float ray_limit(RayState<ANY> const & s) { return s.best_t < 0.0f ? FLT_MAX : s.best_t; }

template <bool ANY>
TGQ_HD void ray_begin(RayState<ANY> & s, BvhView const & b, v3 o, v3 d, float max_time)
{
  s.o = o; s.d = d; s.r.o = o; s.r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  s.best_t = max_time; s.best_idx = -1; s.best_k = -1;
  trav_begin(s.t, b.n_nodes);
}

template <bool ANY>
TGQ_HD void ray_fetch(RayState<ANY> const & s, BvhView const & b, f4 * r)
{
  f4 const * q = s.t.cur >= 0 ? (f4 const *)(b.nodes + s.t.cur) : (f4 const *)(b.tris + s.t.k);
  ldq2(q, r[0], r[1]); ldq2(q + 2, r[2], r[3]);
}

template <bool ANY>
TGQ_HD void ray_node_compute(RayState<ANY> & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  NodeRec nd; nd.q0 = r[0]; nd.q1 = r[1]; nd.q2 = r[2]; nd.q3 = r[3];
  v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
  float t0 = slab_entry(s.r, lo0, hi0), t1 = slab_entry(s.r, lo1, hi1);
  int c0 = f2i(nd.q3.x), c1 = f2i(nd.q3.y);
  float lim = ray_limit(s);
#if defined(__CUDA_ARCH__) && defined(TGQ_RAY_PREFETCH)
  // EXPERIMENT (off; measured on cfg5, 50 M rays: 18.98 -> 19.41 ms, -2 %): when both children are entered, the record of the deferred
  // one is requested into L2 now (the ray kernels are bound by memory latency -- L2 hit rate 36 %, long-scoreboard stalls -- not by
  // the LSU pipe; the deferred child is usually fetched only a few steps later, too soon for the prefetch to have landed)
  if (t0 >= 0.0f && t0 <= lim && t1 >= 0.0f && t1 <= lim)
  {
    int far = t1 < t0 ? c0 : c1;
    void const * pa = far >= 0 ? (void const *)(b.nodes + far) : (void const *)(b.tris + ((~far) >> TGQ_LEAF_BITS));
    asm volatile("prefetch.global.L2 [%0];" :: "l"(pa));
  }
#endif
  // a miss (-1) is mapped above every limit (lim is at most FLT_MAX)
  trav_descend(s.t, t0 >= 0.0f ? t0 : INFINITY, t1 >= 0.0f ? t1 : INFINITY, t0, t1, c0, c1, lim);
}

// One pending triangle.
template <bool ANY>
TGQ_HD void ray_leaf_compute(RayState<ANY> & s, BvhView const & b, f4 const * r)
{
  int k = s.t.k++;
  TGQ_COUNT_PRIM();
  f4 a = r[0], bq = r[1], c = r[2];
  v3 v0 = mk3(a.x, a.y, a.z);
  float tm = ray_tri_time(s.o, s.d, v0, sub3(mk3(bq.x, bq.y, bq.z), v0), sub3(mk3(c.x, c.y, c.z), v0));
  int idx = f2i(a.w);
  if (tm >= 0.0f && (s.best_idx < 0 ? (s.best_t < 0.0f || tm < s.best_t) : (tm < s.best_t || (tm == s.best_t && idx < s.best_idx))))
  {
    s.best_t = tm; s.best_idx = idx; s.best_k = k;
    if (ANY) { s.t.cur = TGQ_ST_DONE; s.t.sp = 0; return; }
  }
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}

template <bool ANY>
TGQ_HD void ray_query(BvhView const & b, v3 o, v3 d, float & best_t, int & best_idx, int & best_k)
{
  RayState<ANY> s; TravStack st; trav_bind(s.t, st);
  ray_begin(s, b, o, d, best_t);
  f4 r[TGQ_FETCH_BOX];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, ray_limit(s)); continue; }
    ray_fetch(s, b, r);
    if (s.t.cur >= 0) ray_node_compute(s, b, r); else ray_leaf_compute(s, b, r);
  }
  best_t = s.best_t; best_idx = s.best_idx; best_k = s.best_k;
}

// ---------------------------------------------------------------------------------------------------------------------
// Local distance histogram (SurfaceFeatures/Local/LocalDistanceHistogram.cpp:36-88, EUCLIDEAN with an explicit max_distance and
// every sample accepted): processRange<IntersectionTester>(Ball3(c, max_distance), callback) with
//   Real d = (x_e - c).norm();  histogram.insert(d)
// and Histogram::insert (Algorithms/Histogram.hpp:110-150): bin = clamp((intx)floor(((double)d - min) / bin_size), 0, B - 1) with
// min = 0 and bin_size = max((double)max_distance, 1e-30) / B in double.  Counts are integers: order independent, bit exact.
struct HistState
{
  v3 c; float r2; double bin_size; int nbins; int stride; unsigned * bins;
  Trav t;
};
TGQ_HD double hist_bin_size(float max_distance, int nbins)
{
  double mx = (double)max_distance; if (mx < 1.0e-30) mx = 1.0e-30;
  return (mx - 0.0) / (double)nbins;
}
TGQ_HD void hist_begin(HistState & s, BvhView const & b, v3 c, float radius)
{
  s.c = c; s.r2 = radius * radius; s.bin_size = hist_bin_size(radius, s.nbins);
  for (int j = 0; j < s.nbins; ++j) s.bins[j * s.stride] = 0u;
  trav_begin(s.t, (radius >= 0.0f) ? b.n_nodes : 0);
}
TGQ_HD void hist_fetch(HistState const & s, BvhView const & b, f4 * r) { box_fetch(s.t, b, r); }
TGQ_HD void hist_node_compute(HistState & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  float lb0, lb1; int c0, c1; box_node_bounds(r, s.c, lb0, lb1, c0, c1);
  trav_descend(s.t, lb0, lb1, 0.0f, 0.0f, c0, c1, s.r2);
}
TGQ_HD void hist_leaf_compute(HistState & s, BvhView const & b, f4 const * r)
{
  int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
  for (int j = 0; j < TGQ_FETCH_BOX; ++j)
  {
    if (j >= cnt) break;
    TGQ_COUNT_PRIM();
    f4 e = r[j];
    float d2 = sqn3(sub3(mk3(e.x, e.y, e.z), s.c));
    if (!(d2 <= s.r2)) continue;
    double q = floor(((double)sqrtf(d2) - 0.0) / s.bin_size);
    int bin = q < 0.0 ? 0 : (q > (double)(s.nbins - 1) ? s.nbins - 1 : (int)q);
    s.bins[bin * s.stride] += 1u;
  }
  s.t.k += cnt;
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}
TGQ_HD void hist_query(BvhView const & b, v3 c, float radius, int nbins, unsigned * bins)
{
  HistState s; TravStack st; trav_bind(s.t, st); s.nbins = nbins; s.stride = 1; s.bins = bins;
  hist_begin(s, b, c, radius);
  f4 r[TGQ_FETCH_BOX];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.r2); continue; }
    hist_fetch(s, b, r);
    if (s.t.cur >= 0) hist_node_compute(s, b, r); else hist_leaf_compute(s, b, r);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Proximity graph with the reachability test of PointSet3's NeighborFunctor (PointSet3.hpp:639-648): a candidate that passed the
// identity, angle and range tests is dropped when the segment from the lifted sample position towards it is blocked by the
// surface: Ray3(x_i + 5 * sep * normalized(n_i + n_j), x_j - x_i), rayIntersects(ray, max_time = 1) on the triangle structure
// `surf`.  The ray is cast only for candidates the bounded set would accept: a blocked candidate never enters the set and the
// set only improves, so the result is the max_degree smallest valid candidates either way, as in the reference (which tests
// every candidate in range).  Needs normals (the reference skips the test without them).
TGQ_HD bool graph_edge_blocked(BvhView const & surf, v3 c, v3 nc, v3 nj, v3 diff, float sep)
{
  v3 lift_dir = normalized3(add3(nc, nj));
  v3 o = add3(c, mul3(lift_dir, 5.0f * sep));
  float best_t = 1.0f; int best_idx = -1, best_k = -1;
  ray_query<true>(surf, o, diff, best_t, best_idx, best_k);
  return best_idx >= 0;
}

TGQ_HD void ball_leaf_compute_surf(BallState & s, BvhView const & b, BvhView const & surf, f4 const * r)
{
  int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
  for (int j = 0; j < TGQ_FETCH_BOX; ++j)
  {
    if (j >= cnt) break;
    TGQ_COUNT_PRIM();
    f4 e = r[j];
    int idx = f2i(e.w);
    if (idx == s.self) continue;
    v3 diff = sub3(mk3(e.x, e.y, e.z), s.c);
    float d2 = sqn3(diff);
    if (!(d2 <= s.r2)) continue;
    v3 nj = mk3(0.0f, 0.0f, 0.0f);
    if (s.has_n)
    {
      f4 q = ldq(b.nrm + s.t.k + j); nj = mk3(q.x, q.y, q.z);
      if (dot3(s.nc, nj) < s.min_dot) continue;
    }
    float sep = fast_sqrt(d2);
    if (!s.set.accepts(sep, idx)) continue;
    if (s.has_n && graph_edge_blocked(surf, s.c, s.nc, nj, diff, sep)) continue;
    s.set.insert(sep, idx);
  }
  s.t.k += cnt;
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}

TGQ_HD void ball_graph_surface_query(BvhView const & b, BvhView const & surf, v3 c, v3 nc, bool has_n, float min_dot, int self,
                                     float init_radius, int rounds, int min_degree, SortedSet & set)
{
  BallState s; TravStack st; trav_bind(s.t, st); s.set = set;
  ball_begin(s, b, c, nc, has_n, min_dot, self, init_radius, rounds, min_degree);
  f4 r[TGQ_FETCH_BOX];
  for (;;)
  {
    while (s.t.cur != TGQ_ST_DONE)
    {
      if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.r2); continue; }
      ball_fetch(s, b, r);
      if (s.t.cur >= 0) ball_node_compute(s, b, r); else ball_leaf_compute_surf(s, b, surf, r);
    }
    if (ball_round_done(s, b)) break;
  }
  set.n = s.set.n;
}

// ---------------------------------------------------------------------------------------------------------------------
// Closest element to a LINE SEGMENT query object: BvhN::closestElement / closestPair<MetricL2>(LineSegment3, ...)
// (BvhN.hpp:1049-1114,1866-1893,2036-2089), as Tools/MeshFix/MeshFix.cpp:756-763 runs it once per boundary edge against a vertex
// structure under a VertexFilter, and as MetricL2Impl<LineSegmentN, TriangleN> (Algorithms/MetricL2.hpp:286-295) defines it on meshes.
// Per-element arithmetic: seg_point_sqdist / tri_seg_sqdist (tgq_math.cuh, reference order); the answer is canonical -- the smallest
// (d2, element index) over all admissible elements.  The reference prunes on the box-box distance between the segment's bounding box
// and the node box (monotonePruningDistance for a bounded query, BvhN.hpp:1794-1803); any conservative bound returns the same answer,
// so the larger of that and a centre-sphere bound (distance of the box centre from the segment minus the half diagonal: tight for long
// oblique segments, whose own box is mostly empty) is used on the AABB nodes of either structure kind.
// Error budget (u = 2^-24, M = largest |coordinate| of the scene, A, B the segment's ends): the reference's closest points are
// point + s * direction with s clamped to [0, 1] on BOTH objects -- whatever the conditioning of s and t -- so they lie within
// 3u(|A| + |B|) of the segment and 5u M of the element, and the returned fl(|c1 - c2|^2) is at least (D - eps)^2 (1 - 4u) with D the true
// distance and eps their sum; a "hit" of the ray test (d2 = 0) on a segment that truly misses the triangle by more than that needs the
// segment to graze the triangle's plane, the same residual the ray traversal accepts.  Box bounds are evaluated with fused
// products (4u per term).  Absolute terms -> mq = 2 pad + 64u(|A|_1 + |B|_1) (pad = 32u M), relative terms -> TGQ_LB_DEFLATE.
#define TGQ_SEG_EXCL 4            // excluded element indices per query (MeshFix's VertexFilter: the edge's ends and its faces' vertices)
template <bool TRI>
struct SegState
{
  v3 a, dir, blo, bhi; float mq;
  int excl[TGQ_SEG_EXCL];
  float best_d2; int best_idx; v3 best_ep, best_sp;       // closest point on the element / on the segment
  Trav t;
};

TGQ_HD float seg_query_margin(v3 a, v3 b, float scene_pad)
{ return TGQ_MARGIN_SCALE * (2.0f * scene_pad + 3.8146973e-6f * (fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(b.x) + fabsf(b.y) + fabsf(b.z))); }

template <bool TRI>
TGQ_HD float seg_box_lb(SegState<TRI> const & s, v3 lo, v3 hi)
{
  v3 g = max3(sub3(max3(lo, s.blo), min3(hi, s.bhi)), mk3(0.0f, 0.0f, 0.0f));
  float d_box = sqrtf(fdot3(g, g));
  v3 c = box_centre(lo, hi), h = sub3(hi, c), w = sub3(c, s.a);
  float dd = fdot3(s.dir, s.dir);
  float tt = dd > 0.0f ? fminf(fmaxf(fdot3(w, s.dir) / dd, 0.0f), 1.0f) : 0.0f;
  v3 x = sub3(w, mul3(s.dir, tt));
  float d_sph = sqrtf(fdot3(x, x)) * 0.999999f - sqrtf(fdot3(h, h)) * 1.000001f;
  float d = fmaxf(fmaxf(d_box, d_sph) - s.mq, 0.0f);
  return d * d * TGQ_LB_DEFLATE;
}

// (point, direction): the segment as LineSegmentN stores it; a caller holding the two endpoints passes direction = end1 - end0, the
// subtraction LineSegment3(end0, end1) performs (LineSegmentN.hpp:56-58).
template <bool TRI>
TGQ_HD void seg_begin(SegState<TRI> & s, BvhView const & b, float scene_pad, v3 a, v3 dir, float bound_ub)
{
  s.a = a; s.dir = dir;
  v3 far = add3(a, s.dir);
  s.blo = min3(a, far); s.bhi = max3(a, far);
  s.mq = seg_query_margin(a, far, scene_pad);
  for (int j = 0; j < TGQ_SEG_EXCL; ++j) s.excl[j] = -1;
  s.best_d2 = bound_ub; s.best_idx = -1; s.best_ep = s.best_sp = mk3(0.0f, 0.0f, 0.0f);
  trav_begin(s.t, b.n_nodes);
}

template <bool TRI>
TGQ_HD void seg_fetch(SegState<TRI> const & s, BvhView const & b, f4 * r)
{
  if (TRI)
  {
    f4 const * q = s.t.cur >= 0 ? (f4 const *)(b.nodes + s.t.cur) : (f4 const *)(b.tris + s.t.k);
    ldq2(q, r[0], r[1]); ldq2(q + 2, r[2], r[3]);
  }
  else box_fetch(s.t, b, r);
}

template <bool TRI>
TGQ_HD void seg_node_compute(SegState<TRI> & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  NodeRec nd; nd.q0 = r[0]; nd.q1 = r[1]; nd.q2 = r[2]; nd.q3 = r[3];
  v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
  float lb0 = seg_box_lb(s, lo0, hi0), lb1 = seg_box_lb(s, lo1, hi1);
  trav_descend(s.t, lb0, lb1, lb0, lb1, f2i(nd.q3.x), f2i(nd.q3.y), s.best_d2);
}

template <bool TRI>
TGQ_HD bool seg_excluded(SegState<TRI> const & s, int idx)
{
  bool out = false;
  for (int j = 0; j < TGQ_SEG_EXCL; ++j) out = out || idx == s.excl[j];
  return out;
}

// One pending triangle (triangle structures) or up to four pending points of the leaf (point structures).
template <bool TRI>
TGQ_HD void seg_leaf_compute(SegState<TRI> & s, BvhView const & b, f4 const * r)
{
  if (TRI)
  {
    ++s.t.k;
    TGQ_COUNT_PRIM();
    int idx = f2i(r[0].w);
    if (!seg_excluded(s, idx))
    {
      v3 ep, sp;
      float d2 = tri_seg_sqdist(mk3(r[0].x, r[0].y, r[0].z), mk3(r[1].x, r[1].y, r[1].z), mk3(r[2].x, r[2].y, r[2].z), s.a, s.dir, ep, sp);
      if (d2 < s.best_d2 || (d2 == s.best_d2 && idx < s.best_idx)) { s.best_d2 = d2; s.best_idx = idx; s.best_ep = ep; s.best_sp = sp; }
    }
  }
  else
  {
    int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
    for (int j = 0; j < TGQ_FETCH_BOX; ++j)
    {
      if (j >= cnt) break;
      TGQ_COUNT_PRIM();
      f4 e = r[j];
      int idx = f2i(e.w); v3 p = mk3(e.x, e.y, e.z);
      if (seg_excluded(s, idx)) continue;
      v3 sp; float d2 = seg_point_sqdist(s.a, s.dir, p, sp);
      if (d2 < s.best_d2 || (d2 == s.best_d2 && idx < s.best_idx)) { s.best_d2 = d2; s.best_idx = idx; s.best_ep = p; s.best_sp = sp; }
    }
    s.t.k += cnt;
  }
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}

// Whole query in one go (host emulation).
template <bool TRI>
TGQ_HD void seg_query(BvhView const & b, float scene_pad, v3 a, v3 dir, int const * excl, int n_excl, float & best_d2, int & best_idx, v3 & elem_pt, v3 & seg_pt)
{
  SegState<TRI> s; TravStack st; trav_bind(s.t, st);
  seg_begin(s, b, scene_pad, a, dir, best_d2);
  for (int j = 0; j < n_excl && j < TGQ_SEG_EXCL; ++j) s.excl[j] = excl[j];
  f4 r[TGQ_FETCH_BOX];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, s.best_d2); continue; }
    seg_fetch(s, b, r);
    if (s.t.cur >= 0) seg_node_compute(s, b, r); else seg_leaf_compute(s, b, r);
  }
  best_d2 = s.best_d2; best_idx = s.best_idx; elem_pt = s.best_ep; seg_pt = s.best_sp;
}

// ---------------------------------------------------------------------------------------------------------------------
// Closest element to a TRIANGLE query object: BvhN::closestPair<MetricL2>(Triangle3, ...) and, one query triangle per
// element of the query structure, the leaf work of BvhN::closestPair(BvhN) between two meshes (BvhN.hpp:1900-2009, Test/TestBvh.cpp:380-385).
// Per-element value: element.squaredDistance(query) on triangle structures (tri_tri_sqdist: Moeller's intersection test, then nine
// edge pairs and six vertex-plane projections, Triangle3.hpp:314-392), query.closestPoint(element) on point structures
// (MetricL2Impl<point, TriangleN>, Algorithms/MetricL2.hpp:219-233 -> Triangle3.hpp:294-304).  Canonical answer: smallest (d2, index).
// Bounds: box-box distance to the query's bounding box, or centre distance minus both bounding radii, less the margin mq.
// `floor2`: Moeller's test snaps a vertex within FLT_EPSILON / |N| of the other triangle's plane (N the UNNORMALISED normal, |N| = twice
// the area) onto that plane (THEA_TRI3_USE_EPSILON_TEST, Triangle3.hpp:602-603,1141-1147), so a pair up to eps / |N_element| + eps / |N_query|
// apart can be reported as intersecting (d2 = 0).  No node within that distance is ever pruned: floor = eps / (smallest positive
// |N| of the structure's elements) + eps / |N_query| + mq.  That is what a CONSISTENT perturbation of the vertices could reach; the reference
// mixes snapped distances with unsnapped coordinates (isect2, :1024-1039), and when all three query vertices are snapped it never computes the
// query's interval and compares the element's with [0, 0] (:1041-1077,1209-1216), so it can call pairs intersecting that are farther apart --
// in the last case at any lateral distance, decided by where the origin lies.  No spatial bound covers that; DESIGN.md 4.3 states this
// exception and tests/test_emu.py characterises it.
// Zero-area triangles make the reference's test degenerate to a 2-D overlap test in a coordinate plane (:711-760 with N = 0) and are outside
// the guarantee as well.
// The floor is taken four times as wide as a consistent perturbation would need: host emulation against brute force on small-scale scenes
// (tolerance 1 - 10 % of the scene) showed the reference's inconsistent verdicts at 1 - 3 times that distance (disagreements per 72,000
// queries at scene scale 0.03: 1,255 with factor 1, 63 with factor 4).
#ifndef TGQ_TRI_FLOOR_FACTOR
#define TGQ_TRI_FLOOR_FACTOR 4.0f
#endif
template <bool TRI>
struct TriQState
{
  TriQ q; v3 blo, bhi, cen; float rad, mq, floor2;
  float best_d2; int best_idx; v3 best_ep, best_qp;       // closest point on the element / on the query triangle
  Trav t;
};
template <bool TRI> TGQ_HD float triq_limit(TriQState<TRI> const & s) { return fmaxf(s.best_d2, s.floor2); }

template <bool TRI>
TGQ_HD void triq_begin(TriQState<TRI> & s, BvhView const & b, float scene_pad, float elem_tau, v3 v0, v3 v1, v3 v2, float bound_ub)
{
  s.q = make_triq(v0, v1, v2);
  s.blo = min3(v0, min3(v1, v2)); s.bhi = max3(v0, max3(v1, v2));
  s.cen = mul3(add3(add3(v0, v1), v2), 0.33333334f);
  s.rad = sqrtf(fmaxf(fdot3(sub3(v0, s.cen), sub3(v0, s.cen)), fmaxf(fdot3(sub3(v1, s.cen), sub3(v1, s.cen)), fdot3(sub3(v2, s.cen), sub3(v2, s.cen))))) * 1.000001f;
  s.mq = TGQ_MARGIN_SCALE * (2.0f * scene_pad + 3.8146973e-6f * (fabsf(v0.x) + fabsf(v0.y) + fabsf(v0.z) + fabsf(v1.x) + fabsf(v1.y) + fabsf(v1.z)
                                                               + fabsf(v2.x) + fabsf(v2.y) + fabsf(v2.z)));
  s.floor2 = 0.0f;
  if (TRI)
  {
    v3 N = cross3(sub3(v1, v0), sub3(v2, v0));
    float nn = sqrtf(fdot3(N, N));
    float tau = (nn > 0.0f ? FLT_EPSILON / nn : 0.0f) + elem_tau;
    float f = (TGQ_TRI_FLOOR_FACTOR * tau + s.mq) * 1.000001f;
    s.floor2 = f * f;
  }
  s.best_d2 = bound_ub; s.best_idx = -1; s.best_ep = s.best_qp = mk3(0.0f, 0.0f, 0.0f);
  trav_begin(s.t, b.n_nodes);
}

template <bool TRI>
TGQ_HD float triq_box_lb(TriQState<TRI> const & s, v3 lo, v3 hi)
{
  v3 g = max3(sub3(max3(lo, s.blo), min3(hi, s.bhi)), mk3(0.0f, 0.0f, 0.0f));
  float d_box = sqrtf(fdot3(g, g));
  v3 c = box_centre(lo, hi), h = sub3(hi, c), w = sub3(c, s.cen);
  float d_sph = sqrtf(fdot3(w, w)) * 0.999999f - sqrtf(fdot3(h, h)) * 1.000001f - s.rad;
  float d = fmaxf(fmaxf(d_box, d_sph) - s.mq, 0.0f);
  return d * d * TGQ_LB_DEFLATE;
}

template <bool TRI>
TGQ_HD void triq_fetch(TriQState<TRI> const & s, BvhView const & b, f4 * r)
{
  if (TRI)
  {
    f4 const * q = s.t.cur >= 0 ? (f4 const *)(b.nodes + s.t.cur) : (f4 const *)(b.tris + s.t.k);
    ldq2(q, r[0], r[1]); ldq2(q + 2, r[2], r[3]);
  }
  else box_fetch(s.t, b, r);
}

template <bool TRI>
TGQ_HD void triq_node_compute(TriQState<TRI> & s, BvhView const & b, f4 const * r)
{
  TGQ_COUNT_NODE();
  NodeRec nd; nd.q0 = r[0]; nd.q1 = r[1]; nd.q2 = r[2]; nd.q3 = r[3];
  v3 lo0, hi0, lo1, hi1; node_boxes(nd, lo0, hi0, lo1, hi1);
  float lb0 = triq_box_lb(s, lo0, hi0), lb1 = triq_box_lb(s, lo1, hi1);
  trav_descend(s.t, lb0, lb1, lb0, lb1, f2i(nd.q3.x), f2i(nd.q3.y), triq_limit(s));
}

template <bool TRI>
TGQ_HD void triq_leaf_compute(TriQState<TRI> & s, BvhView const & b, f4 const * r)
{
  if (TRI)
  {
    ++s.t.k;
    TGQ_COUNT_PRIM();
    int idx = f2i(r[0].w);
    TriQ a; a.v0 = mk3(r[0].x, r[0].y, r[0].z); a.v1 = mk3(r[1].x, r[1].y, r[1].z); a.v2 = mk3(r[2].x, r[2].y, r[2].z);
    a.n = mk3(r[3].x, r[3].y, r[3].z); a.d = r[2].w; a.axis = f2i(r[1].w);
    v3 ep, qp;
    float d2 = tri_tri_sqdist(a, s.q, ep, qp);
    if (d2 < s.best_d2 || (d2 == s.best_d2 && idx < s.best_idx)) { s.best_d2 = d2; s.best_idx = idx; s.best_ep = ep; s.best_qp = qp; }
  }
  else
  {
    int cnt = s.t.kend - s.t.k; if (cnt > TGQ_FETCH_BOX) cnt = TGQ_FETCH_BOX;
    for (int j = 0; j < TGQ_FETCH_BOX; ++j)
    {
      if (j >= cnt) break;
      TGQ_COUNT_PRIM();
      f4 e = r[j];
      int idx = f2i(e.w); v3 p = mk3(e.x, e.y, e.z);
      v3 qp; float d2 = tri_point_sqdist(s.q.v0, s.q.v1, s.q.v2, s.q.n, s.q.d, s.q.axis, p, qp);
      if (d2 < s.best_d2 || (d2 == s.best_d2 && idx < s.best_idx)) { s.best_d2 = d2; s.best_idx = idx; s.best_ep = p; s.best_qp = qp; }
    }
    s.t.k += cnt;
  }
  if (s.t.k >= s.t.kend) trav_pop(s.t);
}

// Whole query in one go (host emulation).
template <bool TRI>
TGQ_HD void triq_query(BvhView const & b, float scene_pad, float elem_tau, v3 v0, v3 v1, v3 v2, float & best_d2, int & best_idx, v3 & elem_pt, v3 & query_pt)
{
  TriQState<TRI> s; TravStack st; trav_bind(s.t, st);
  triq_begin(s, b, scene_pad, elem_tau, v0, v1, v2, best_d2);
  f4 r[TGQ_FETCH_BOX];
  while (s.t.cur != TGQ_ST_DONE)
  {
    if (s.t.cur == TGQ_ST_POP) { trav_pop_step(s.t, triq_limit(s)); continue; }
    triq_fetch(s, b, r);
    if (s.t.cur >= 0) triq_node_compute(s, b, r); else triq_leaf_compute(s, b, r);
  }
  best_d2 = s.best_d2; best_idx = s.best_idx; elem_pt = s.best_ep; query_pt = s.best_qp;
}

// BvhN::closestPair's treatment of a non-negative distance bound (see BoundGate above) for a BOUNDED query object: the root rejection
// uses the box-box distance between the query's bounding box and the reference's root box (AxisAlignedBoxN::squaredDistance(box),
// AxisAlignedBoxN.hpp:429-442: vmin - vmax clamped at zero, squared norm in Eigen's order).
TGQ_HD float ref_box_box_sqdist(v3 lo, v3 hi, v3 olo, v3 ohi) { return sqn3(max3(sub3(max3(lo, olo), min3(hi, ohi)), mk3(0.0f, 0.0f, 0.0f))); }
TGQ_HD float gated_ub_box(BoundGate const & g, v3 blo, v3 bhi, float ub) { return (g.on && ref_box_box_sqdist(blo, bhi, g.lo, g.hi) > g.reject) ? -1.0f : ub; }

}  // namespace tgq
