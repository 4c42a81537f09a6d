// tgq.cu -- host side of libtgq.so: the handle, device memory, launch sequences and the C ABI declared in
// include/thea_gpu_query.h.  No CPU query path exists here: every query entry point ends in kernels from
// tgq_kernels.cuh, and tgq_create fails when there is no CUDA device.
#include "../../include/thea_gpu_query.h"
#include "tgq_kernels.cuh"
#include "tgq_graph.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace tgq;

namespace {

std::string g_create_error;
std::mutex g_smem_mutex;
std::map<std::pair<int, void const *>, size_t> g_smem_limit;      // (device, kernel) -> dynamic shared memory limit set so far

struct DevBuf
{
  void * p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes)
  {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 16 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T * as() const { return (T *)p; }
};

#define TGQ_MAX_ARGS 8
#define TGQ_SHARD_CELL_BITS 12       // default of TGQ_OPT_SHARD_CELL_BITS (measured on cfg5, 8 ranks: 9..12 bits 17.6 ms per closest share, 18 bits 18.5, 21 bits 19.3): 2^12 Morton cells over the scene box grown x2, dealt round-robin to the ranks
#ifndef TGQ_SLOTS
// Ray kernels: resident blocks per SM (the rest of the shared-memory array becomes L1).  Measured on B200, 8 M rays on the 10 M-triangle
// mesh: 7 blocks with the maximum carve-out 1.75 G rays/s, 6 blocks 2.27 G, 5 blocks 2.11 G (1 M triangles: 2.24 / 2.73 / 2.53 G).
#ifndef TGQ_RAY_BLOCKS
#define TGQ_RAY_BLOCKS 6
#endif
#define TGQ_SLOTS 4       // pipeline depth of host-pointer calls (within ONE call, measured on cfg5: 2 -> 462, 3 -> 475, 4 -> 474 M queries/s end to end;
                          // four slots, handed out round-robin ACROSS calls, let an asynchronous ray call upload all of its four chunks under the
                          // kernels of the closest-point call before it instead of queueing behind that call's last chunks)
#endif

// Scratch of one pipeline slot.  Host-pointer calls rotate over TGQ_SLOTS slots, each with its own stream, so that the
// copies of one chunk overlap the kernels of the others.
struct Slot
{
  cudaStream_t stream = nullptr;
  DevBuf in[TGQ_MAX_ARGS], out[TGQ_MAX_ARGS], keys[2], vals[2], cubtmp, counter;
};

}  // namespace

struct tgq_handle
{
  int device = 0;
  std::string err;
  cudaStream_t own_stream = nullptr, stream = nullptr;

  int opt_sort = 2, opt_leaf = 0, opt_async = 0, opt_pool = 1, opt_icp_ref = 1, opt_wide = 0, opt_setbuf = 1; unsigned slot_rr = 0; bool wide_built = false; int64_t opt_chunk = 16777216;

  int kind = TGQ_KIND_NONE; int n = 0; int n_nodes = 0; bool has_normals = false, has_labels = false; int leaf_built = 4;
  DevBuf nodes, elems, nrm, lab, scene;                                // NodeRec[], TriRec[] | f4[], f4[], SceneBox
  DevBuf perm, pos_of_elem, first, last, split, parent, live, newidx;   // kept for refit
  DevBuf pbox_lo, pbox_hi, nbox_lo, nbox_hi, arrive, keys_a, keys_b, vals_a, cubtmp, src, src2, misc;
  DevBuf icp_cur, icp_orig, icp_to, icp_idx;                          // device-resident ICP loop
  DevBuf area_by_elem, cum;                                           // float[n] in element order; double[n] cumulative
  DevBuf obbs, pnrm, nsum, obbinfo, big, big_count;                   // oriented boxes (triangle structures)
  DevBuf wide, wflag, widx, obbinfo_w; int n_wide = 0;                 // 4-wide nodes of the closest-point traversal
  int opt_cell_bits = TGQ_SHARD_CELL_BITS, opt_ray_order = 0; bool sorting_rays = false; int sort_stride3 = 1;
  int shard_rank = 0, shard_world = 1;                                 // set by the *_records entry points for the duration of one call
  PeerOut peers = {}; int64_t last_share = 0;                          // queries this rank answered in the last *_records call
  int sm_count = 148; std::map<std::pair<void const *, size_t>, int> occupancy;   // resident blocks per SM of the persistent kernels
  bool cum_valid = false;
  DevBuf g_idx, g_sep, g_dist, g_stamp, g_touched, g_queue, g_seeds, g_out0, g_out1, g_out2;   // sample-graph searches (tgq_graph_*)
  bool orig_valid = false; DevBuf orig, orig_n, orig_lab, xf;           // untransformed source elements (tgq_set_transform)
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}; float pad = 0;

  Slot slot[TGQ_SLOTS];
  std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  tgq_stats st = {};
};

namespace {

#define CK(call)                                                                                                      \
  do {                                                                                                                \
    cudaError_t e__ = (call);                                                                                         \
    if (e__ != cudaSuccess) {                                                                                         \
      char buf__[512];                                                                                                \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);     \
      h->err = buf__;                                                                                                 \
      return TGQ_ECUDA;                                                                                               \
    }                                                                                                                 \
  } while (0)

#define CKRC(call) do { int rc__ = (call); if (rc__ != TGQ_OK) return rc__; } while (0)

int fail(tgq_handle * h, int code, char const * msg) { if (h) h->err = msg; return code; }

inline int grid_for(int64_t n, int block = TGQ_BLOCK) { return (int)((n + block - 1) / block); }

bool is_device_ptr(void const * p)
{
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

BvhView view_of(tgq_handle * h)
{
  BvhView v;
  v.nodes = h->nodes.as<NodeRec>();
  v.obbs = h->kind == TGQ_KIND_TRIANGLES ? h->obbs.as<ObbNode>() : nullptr;
  v.tris = h->kind == TGQ_KIND_TRIANGLES ? h->elems.as<TriRec>() : nullptr;
  v.pts = h->kind == TGQ_KIND_POINTS ? h->elems.as<f4>() : nullptr;
  v.nrm = (h->kind == TGQ_KIND_POINTS && h->has_normals) ? h->nrm.as<f4>() : nullptr;
  v.lab = (h->kind == TGQ_KIND_POINTS && h->has_labels) ? h->lab.as<int>() : nullptr;
  v.n_nodes = h->n_nodes; v.n_prims = h->n;
  bool wide = h->kind == TGQ_KIND_TRIANGLES && h->wide_built;
  v.wide = wide ? h->wide.as<WideNode>() : nullptr; v.n_wide = wide ? h->n_wide : 0;
  return v;
}

cudaEvent_t get_event(tgq_handle * h)
{
  if (h->ev_used == h->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
  return h->ev_pool[h->ev_used++];
}

// Exclusive squared-distance bound as a float: d2 < ub  <=>  (double)d2 < bound^2 (BvhN.hpp:448,462).
float bound_to_ub(double dist_bound)
{
  if (!(dist_bound >= 0)) return INFINITY;
  return float_at_least(dist_bound * dist_bound);   // smallest float >= bound^2 (kClosestPairs: no widening, BvhN.hpp:478)
}

// Returns a device pointer holding `bytes` of p (copying when p is a host pointer).
int to_device(tgq_handle * h, void const * p, size_t bytes, DevBuf & buf, cudaStream_t s, void const ** out)
{
  if (is_device_ptr(p)) { *out = p; return TGQ_OK; }
  CK(buf.reserve(bytes));
  CK(cudaMemcpyAsync(buf.p, p, bytes, cudaMemcpyHostToDevice, s));
  *out = buf.p;
  return TGQ_OK;
}

int from_device(tgq_handle * h, void * dst, void const * src_dev, size_t bytes, cudaStream_t s)
{
  CK(cudaMemcpyAsync(dst, src_dev, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  return TGQ_OK;
}

// ---- build --------------------------------------------------------------------------------------------------------------

int refit_and_emit(tgq_handle * h, cudaStream_t s)
{
  int n = h->n;
  if (n >= 2)
  {
    CK(cudaMemsetAsync(h->arrive.p, 0, sizeof(unsigned) * (size_t)(n - 1), s));
    bool tri = h->kind == TGQ_KIND_TRIANGLES;
    k_refit<<<grid_for(n), TGQ_BLOCK, 0, s>>>(n, h->first.as<int>(), h->last.as<int>(), h->split.as<int>(), h->parent.as<int>(),
                                              h->pbox_lo.as<f4>(), h->pbox_hi.as<f4>(), h->nbox_lo.as<f4>(), h->nbox_hi.as<f4>(),
                                              h->arrive.as<unsigned>(), tri ? h->pnrm.as<f4>() : nullptr, h->nsum.as<f4>());
    if (tri)
    {
      CK(cudaMemsetAsync(h->big_count.p, 0, 4, s));
      k_obb_small<<<grid_for(n - 1), TGQ_BLOCK, 0, s>>>(n, h->leaf_built, h->first.as<int>(), h->last.as<int>(), h->parent.as<int>(),
                                                        h->nsum.as<f4>(), h->elems.as<TriRec>(), h->scene.as<SceneBox>(), h->obbinfo.as<f4>(),
                                                        h->wide_built ? h->obbinfo_w.as<f4>() : nullptr, h->big.as<int>(), h->big_count.as<int>());
      int big_grid = std::max(1, std::min(n / TGQ_SLAB_SMALL + 1, 148 * 64));
      k_obb_big<<<big_grid, 128, 0, s>>>(h->big.as<int>(), h->big_count.as<int>(), h->first.as<int>(), h->last.as<int>(), h->parent.as<int>(), h->nsum.as<f4>(),
                                         h->elems.as<TriRec>(), h->scene.as<SceneBox>(), h->obbinfo.as<f4>(), h->wide_built ? h->obbinfo_w.as<f4>() : nullptr);
      if (h->wide_built)
      k_emit_wide<<<grid_for(n - 1), TGQ_BLOCK, 0, s>>>(n, h->leaf_built, h->first.as<int>(), h->last.as<int>(), h->split.as<int>(), h->wflag.as<int>(),
                                                        h->widx.as<int>(), h->scene.as<SceneBox>(), h->elems.as<TriRec>(), h->obbinfo_w.as<f4>(),
                                                        h->wide.as<WideNode>());
      h->st.total_launches += 3;
    }
    k_emit<<<grid_for(n - 1), TGQ_BLOCK, 0, s>>>(n, h->leaf_built, h->first.as<int>(), h->last.as<int>(), h->split.as<int>(),
                                                 h->live.as<int>(), h->newidx.as<int>(), h->pbox_lo.as<f4>(), h->pbox_hi.as<f4>(),
                                                 h->nbox_lo.as<f4>(), h->nbox_hi.as<f4>(), h->scene.as<SceneBox>(), h->nodes.as<NodeRec>(),
                                                 tri ? h->elems.as<TriRec>() : nullptr, tri ? h->obbinfo.as<f4>() : nullptr,
                                                 tri ? h->obbs.as<ObbNode>() : nullptr, tri ? h->nsum.as<f4>() : nullptr);
    h->st.total_launches += 2;
  }
  else if (n == 1)
  {
    bool tri = h->kind == TGQ_KIND_TRIANGLES;
    k_emit_single<<<1, 1, 0, s>>>(h->pbox_lo.as<f4>(), h->pbox_hi.as<f4>(), h->scene.as<SceneBox>(), h->nodes.as<NodeRec>(),
                                  tri ? h->elems.as<TriRec>() : nullptr, tri ? h->obbs.as<ObbNode>() : nullptr);
    if (tri && h->wide_built) k_emit_wide_single<<<1, 1, 0, s>>>(h->scene.as<SceneBox>(), h->elems.as<TriRec>(), h->wide.as<WideNode>());
    h->st.total_launches += tri ? 2 : 1;
  }
  CK(cudaGetLastError());
  return TGQ_OK;
}

int scene_bounds(tgq_handle * h, float const * d_pts, int64_t npts, cudaStream_t s)
{
  CK(h->scene.reserve(sizeof(SceneBox)));
  k_scene_init<<<1, 1, 0, s>>>(h->scene.as<SceneBox>());
  if (npts > 0)
  {
    int grid = (int)std::min<int64_t>((npts + 255) / 256, (int64_t)h->st.num_elements * 0 + 148 * 8);
    k_scene_bounds<<<grid, 256, 0, s>>>(d_pts, npts, h->scene.as<SceneBox>());
  }
  h->st.total_launches += 2;
  CK(cudaGetLastError());
  return TGQ_OK;
}

int fetch_scene(tgq_handle * h, cudaStream_t s)
{
  SceneBox sb;
  CK(cudaMemcpyAsync(&sb, h->scene.p, sizeof sb, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int a = 0; a < 3; ++a)
  {
    int l = sb.lo[a], u = sb.hi[a];
    l = l >= 0 ? l : l ^ 0x7fffffff; u = u >= 0 ? u : u ^ 0x7fffffff;
    memcpy(&h->lo[a], &l, 4); memcpy(&h->hi[a], &u, 4);
  }
  float m = 0;
  for (int a = 0; a < 3; ++a) m = std::max(m, std::max(std::fabs(h->lo[a]), std::fabs(h->hi[a])));
  h->pad = m * 1.9073486328125e-06f + FLT_MIN;      // same expression as scene_decode() on the device
  return TGQ_OK;
}

// d_src: device pointer to T x 9 floats (triangles) or P x 3 floats (points); d_n3: device normals or null.
int build_common(tgq_handle * h, int kind, float const * d_src, float const * d_n3, int64_t n64)
{
  cudaStream_t s = h->stream;
  if (n64 < 0) return fail(h, TGQ_EINVAL, "negative element count");
  if (n64 >= ((int64_t)1 << (31 - TGQ_LEAF_BITS))) return fail(h, TGQ_ELIMIT, "more than 2^27 - 1 elements in one structure");
  int n = (int)n64;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  h->ev_used = 0; e0 = get_event(h); e1 = get_event(h);
  CK(cudaEventRecord(e0, s));

  h->kind = kind; h->n = n; h->n_nodes = 0; h->n_wide = 0; h->wide_built = false; h->has_normals = (kind == TGQ_KIND_POINTS && d_n3 != nullptr);
  // leaf size: the caller's, or by kind -- 4 triangles (the pooled closest-point kernel's queue is sized for it; 1 / 2 / 4 measured equal),
  // 15 points (measured on 1 M samples, leaf 4 -> 15: k-NN 401 -> 464 M/s, sample graph 544 -> 592 M/s, ball statistics 40.8 -> 25.1 ms)
  h->leaf_built = h->opt_leaf > 0 ? h->opt_leaf : (kind == TGQ_KIND_POINTS ? TGQ_LEAF_MASK : 4); h->cum_valid = false; h->has_labels = false; h->orig_valid = false;
  h->st.num_elements = n;
  for (int a = 0; a < 3; ++a) { h->lo[a] = 0; h->hi[a] = 0; }
  if (n == 0) { h->st.num_nodes = 0; h->st.node_bytes = 0; h->st.element_bytes = 0; h->st.last_build_ms = 0; return TGQ_OK; }

  CKRC(scene_bounds(h, d_src, kind == TGQ_KIND_TRIANGLES ? 3 * n64 : n64, s));

  size_t nn = (size_t)n;
  CK(h->keys_a.reserve(8 * nn)); CK(h->keys_b.reserve(8 * nn)); CK(h->vals_a.reserve(4 * nn)); CK(h->perm.reserve(4 * nn));
  CK(h->pos_of_elem.reserve(4 * nn));
  CK(h->pbox_lo.reserve(16 * nn)); CK(h->pbox_hi.reserve(16 * nn));
  size_t rec = kind == TGQ_KIND_TRIANGLES ? sizeof(TriRec) : sizeof(f4);
  CK(h->elems.reserve(rec * nn));
  if (h->has_normals) CK(h->nrm.reserve(16 * nn));
  if (kind == TGQ_KIND_TRIANGLES) { CK(h->pnrm.reserve(16 * nn)); CK(h->area_by_elem.reserve(4 * nn)); }

  if (kind == TGQ_KIND_TRIANGLES) k_morton_tris<<<grid_for(n), TGQ_BLOCK, 0, s>>>(d_src, n, h->scene.as<SceneBox>(), h->keys_a.as<uint64_t>(), h->vals_a.as<uint32_t>());
  else k_morton_points<<<grid_for(n), TGQ_BLOCK, 0, s>>>(d_src, n, h->scene.as<SceneBox>(), h->keys_a.as<uint64_t>(), h->vals_a.as<uint32_t>());
  CK(cudaGetLastError());

  // K2: stable LSD radix sort of (morton63, element index)
  size_t tmp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->keys_a.as<uint64_t>(), h->keys_b.as<uint64_t>(), h->vals_a.as<uint32_t>(),
                                     h->perm.as<uint32_t>(), n, 0, 63, s));
  CK(h->cubtmp.reserve(tmp_bytes));
  CK(cub::DeviceRadixSort::SortPairs(h->cubtmp.p, tmp_bytes, h->keys_a.as<uint64_t>(), h->keys_b.as<uint64_t>(), h->vals_a.as<uint32_t>(),
                                     h->perm.as<uint32_t>(), n, 0, 63, s));

  if (kind == TGQ_KIND_TRIANGLES)
    k_gather_tris<<<grid_for(n), TGQ_BLOCK, 0, s>>>(d_src, h->perm.as<uint32_t>(), n, h->scene.as<SceneBox>(), h->elems.as<TriRec>(),
                                                    h->pbox_lo.as<f4>(), h->pbox_hi.as<f4>(), h->pos_of_elem.as<int>(), h->pnrm.as<f4>(),
                                                    h->area_by_elem.as<float>());
  else
    k_gather_points<<<grid_for(n), TGQ_BLOCK, 0, s>>>(d_src, d_n3, h->perm.as<uint32_t>(), n, h->elems.as<f4>(), h->nrm.as<f4>(),
                                                      h->pbox_lo.as<f4>(), h->pbox_hi.as<f4>(), h->pos_of_elem.as<int>());
  CK(cudaGetLastError());
  h->st.total_launches += 8;

  if (n >= 2)
  {
    size_t ni = nn - 1;
    CK(h->first.reserve(4 * ni)); CK(h->last.reserve(4 * ni)); CK(h->split.reserve(4 * ni)); CK(h->parent.reserve(4 * (2 * nn)));
    CK(h->live.reserve(4 * ni)); CK(h->newidx.reserve(4 * ni)); CK(h->arrive.reserve(4 * ni));
    CK(h->nbox_lo.reserve(16 * ni)); CK(h->nbox_hi.reserve(16 * ni)); CK(h->nsum.reserve(16 * ni));
    if (kind == TGQ_KIND_TRIANGLES)
    {
      CK(h->obbinfo.reserve(16 * TGQ_OBB_F4 * ni)); CK(h->big.reserve(4 * ni)); CK(h->big_count.reserve(4));
      if (h->opt_wide) { CK(h->obbinfo_w.reserve(64 * ni)); CK(h->wflag.reserve(4 * ni)); CK(h->widx.reserve(4 * ni)); }
    }
    k_hierarchy<<<grid_for(n - 1), TGQ_BLOCK, 0, s>>>(h->keys_b.as<uint64_t>(), n, h->first.as<int>(), h->last.as<int>(), h->split.as<int>(),
                                                      h->parent.as<int>());
    k_live<<<grid_for(n - 1), TGQ_BLOCK, 0, s>>>(n, h->leaf_built, h->first.as<int>(), h->last.as<int>(), h->live.as<int>());
    CK(cudaGetLastError());
    size_t scan_bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, h->live.as<int>(), h->newidx.as<int>(), n - 1, s));
    CK(h->cubtmp.reserve(scan_bytes));
    CK(cub::DeviceScan::ExclusiveSum(h->cubtmp.p, scan_bytes, h->live.as<int>(), h->newidx.as<int>(), n - 1, s));
    int tail[2];
    CK(cudaMemcpyAsync(&tail[0], h->newidx.as<int>() + (n - 2), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&tail[1], h->live.as<int>() + (n - 2), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    h->n_nodes = tail[0] + tail[1];
    h->st.total_launches += 4;
    h->wide_built = false;
    if (kind == TGQ_KIND_TRIANGLES && h->opt_wide)
    {
      h->wide_built = true;
      // which surviving nodes are 4-wide nodes (even depth), and their compacted indices
      k_wide_flags<<<grid_for(n - 1), TGQ_BLOCK, 0, s>>>(n, h->live.as<int>(), h->parent.as<int>(), h->wflag.as<int>());
      CK(cudaGetLastError());
      CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, h->wflag.as<int>(), h->widx.as<int>(), n - 1, s));
      CK(h->cubtmp.reserve(scan_bytes));
      CK(cub::DeviceScan::ExclusiveSum(h->cubtmp.p, scan_bytes, h->wflag.as<int>(), h->widx.as<int>(), n - 1, s));
      CK(cudaMemcpyAsync(&tail[0], h->widx.as<int>() + (n - 2), 4, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(&tail[1], h->wflag.as<int>() + (n - 2), 4, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      h->n_wide = tail[0] + tail[1];
      h->st.total_launches += 3;
    }
  }
  else { h->n_nodes = 1; h->n_wide = 1; h->wide_built = kind == TGQ_KIND_TRIANGLES && h->opt_wide; }

  CK(h->nodes.reserve(sizeof(NodeRec) * (size_t)h->n_nodes));
  if (kind == TGQ_KIND_TRIANGLES) { CK(h->obbs.reserve(sizeof(ObbNode) * (size_t)h->n_nodes)); if (h->wide_built) CK(h->wide.reserve(sizeof(WideNode) * (size_t)std::max(1, h->n_wide))); }
  CKRC(refit_and_emit(h, s));
  CK(cudaEventRecord(e1, s));
  CKRC(fetch_scene(h, s));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  h->st.last_build_ms = ms;
  h->st.num_nodes = h->n_nodes;
  h->st.node_bytes = (int64_t)(sizeof(NodeRec) + (kind == TGQ_KIND_TRIANGLES ? sizeof(ObbNode) : 0)) * h->n_nodes
                   + (h->wide_built ? (int64_t)sizeof(WideNode) * h->n_wide : 0);
  h->st.element_bytes = (int64_t)(rec * nn);
  return TGQ_OK;
}

// ---- batched query driver ---------------------------------------------------------------------------------------------------

struct Arg { void const * in; void * out; size_t stride; };   // one of in/out non-null; stride = bytes per query

struct ChunkCtx
{
  Slot * slot; cudaStream_t stream; int64_t offset; int n;
  void const * din[TGQ_MAX_ARGS]; void * dout[TGQ_MAX_ARGS];
  uint32_t const * perm;   // Morton order of the chunk, or null
};

// Orders the chunk's queries by the Morton code of their first input (positions / ray origins).
int sort_chunk(tgq_handle * h, ChunkCtx & c)
{
  bool want = h->opt_sort == 1 || (h->opt_sort == 2 && c.n >= 4096);
  c.perm = nullptr;
  Slot & sl = *c.slot; size_t nn = (size_t)c.n;
  if (h->shard_world > 1)
  {
    // this rank's share of the replicated batch (Morton cells dealt round-robin), then the Morton order of the share alone
    for (int i = 0; i < 2; ++i) { CK(sl.keys[i].reserve(4 * nn)); CK(sl.vals[i].reserve(4 * nn)); }
    int * cnt = sl.counter.as<int>() + 1;
    CK(cudaMemsetAsync(cnt, 0, 4, c.stream));
    int grid = (int)std::min<int64_t>((c.n + 1023) / 1024, (int64_t)h->sm_count * 8);
    k_query_codes_shard<<<grid, 256, 0, c.stream>>>((float const *)c.din[0], c.n, h->scene.as<SceneBox>(), h->shard_rank, h->shard_world, 30 - h->opt_cell_bits,
                                                    sl.keys[0].as<uint32_t>(), sl.vals[0].as<uint32_t>(), cnt);
    CK(cudaGetLastError());
    int own = 0;
    CK(cudaMemcpyAsync(&own, cnt, 4, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    c.n = own; h->last_share += own;
    h->st.last_launches += 1; h->st.total_launches += 1;
    if (own > 0)
    {
      size_t tmp = 0;
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, sl.keys[0].as<uint32_t>(), sl.keys[1].as<uint32_t>(), sl.vals[0].as<uint32_t>(),
                                         sl.vals[1].as<uint32_t>(), own, 0, 30, c.stream));
      CK(sl.cubtmp.reserve(tmp));
      CK(cub::DeviceRadixSort::SortPairs(sl.cubtmp.p, tmp, sl.keys[0].as<uint32_t>(), sl.keys[1].as<uint32_t>(), sl.vals[0].as<uint32_t>(),
                                         sl.vals[1].as<uint32_t>(), own, 0, 30, c.stream));
      h->st.last_launches += 3; h->st.total_launches += 3;
    }
    c.perm = sl.vals[1].as<uint32_t>();
    return TGQ_OK;
  }
  if (!want || h->n == 0) return TGQ_OK;
  for (int i = 0; i < 2; ++i) { CK(sl.keys[i].reserve(4 * nn)); CK(sl.vals[i].reserve(4 * nn)); }
  k_query_codes<<<grid_for(c.n), TGQ_BLOCK, 0, c.stream>>>((float const *)c.din[0], c.n, h->scene.as<SceneBox>(), sl.keys[0].as<uint32_t>(),
                                                           sl.vals[0].as<uint32_t>(), h->sorting_rays ? (float const *)c.din[1] : nullptr, h->opt_ray_order, h->sort_stride3);
  CK(cudaGetLastError());
  size_t tmp = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, sl.keys[0].as<uint32_t>(), sl.keys[1].as<uint32_t>(), sl.vals[0].as<uint32_t>(),
                                     sl.vals[1].as<uint32_t>(), c.n, 0, 30, c.stream));
  CK(sl.cubtmp.reserve(tmp));
  CK(cub::DeviceRadixSort::SortPairs(sl.cubtmp.p, tmp, sl.keys[0].as<uint32_t>(), sl.keys[1].as<uint32_t>(), sl.vals[0].as<uint32_t>(),
                                     sl.vals[1].as<uint32_t>(), c.n, 0, 30, c.stream));
  c.perm = sl.vals[1].as<uint32_t>();
  h->st.last_launches += 4; h->st.total_launches += 4;
  return TGQ_OK;
}

template <typename Launch>
int run_batched(tgq_handle * h, int64_t Q, Arg * args, int nargs, bool sortable, Launch launch)
{
  h->st.last_query_ms = 0; h->st.last_sort_ms = 0; h->st.last_launches = 0; h->ev_used = 0;
  if (Q == 0) return TGQ_OK;
  bool all_dev = true;
  for (int a = 0; a < nargs; ++a)
  {
    void const * p = args[a].in ? args[a].in : args[a].out;
    if (p && !is_device_ptr(p)) all_dev = false;
  }
  // host-pointer calls: at least four chunks, so that the upload of one chunk, the kernels of another and the download of a third
  // overlap even when the batch is smaller than TGQ_OPT_CHUNK (a 6 M-query shard used to be ONE chunk: copy, sort, kernel, copy in series)
  // ... but not below about 1.5 M queries per chunk (a smaller Morton-ordered batch is sparser and the traversal kernels lose more on
  // it than the exposed copies cost).  The floor is deliberately low: on ONE GPU a rank's 6.25 M + 6.25 M share of the cfg5 batches runs
  // faster through host buffers with one chunk per call (25.0 ms per step against 29.7 ms with four, tools/e2e_shard_probe.py), but with
  // eight ranks on one host the unchunked calls all upload, compute and download at the same moments and the step takes 36.1 ms
  // against 32.7 ms with four chunks per call (bench.py --gpus 8).  TGQ_DEV_MIN_CHUNK overrides the floor for experiments.
  static int64_t const min_chunk = getenv("TGQ_DEV_MIN_CHUNK") ? atoll(getenv("TGQ_DEV_MIN_CHUNK")) : 1500000;
  int64_t nchunks = std::max<int64_t>(1, std::min<int64_t>(4, (Q + min_chunk / 2) / std::max<int64_t>(1, min_chunk)));
  int64_t chunk = all_dev ? ((int64_t)1 << 26) : std::max<int64_t>(1024, std::min<int64_t>(h->opt_chunk, ((Q + nchunks - 1) / nchunks + 1023) / 1024 * 1024));
  int nslots = all_dev ? 1 : TGQ_SLOTS;
  if (!all_dev) CK(cudaStreamSynchronize(h->stream));     // the structure must be complete before other streams read it
  std::vector<cudaEvent_t> evs;
  int ci = 0;
  for (int64_t off = 0; off < Q; off += chunk, ++ci)
  {
    ChunkCtx c; c.slot = &h->slot[all_dev ? 0 : (h->slot_rr++ % nslots)]; c.stream = all_dev ? h->stream : c.slot->stream; c.offset = off;
    c.n = (int)std::min<int64_t>(chunk, Q - off); c.perm = nullptr;
    for (int a = 0; a < nargs; ++a)
    {
      c.din[a] = nullptr; c.dout[a] = nullptr;
      size_t bytes = args[a].stride * (size_t)c.n, skip = args[a].stride * (size_t)off;
      if (args[a].in)
      {
        char const * p = (char const *)args[a].in + skip;
        if (is_device_ptr(args[a].in)) c.din[a] = p;
        else { CK(c.slot->in[a].reserve(bytes)); CK(cudaMemcpyAsync(c.slot->in[a].p, p, bytes, cudaMemcpyHostToDevice, c.stream)); c.din[a] = c.slot->in[a].p; }
      }
      else if (args[a].out)
      {
        if (is_device_ptr(args[a].out)) c.dout[a] = (char *)args[a].out + skip;
        else { CK(c.slot->out[a].reserve(bytes)); c.dout[a] = c.slot->out[a].p; }
      }
    }
    cudaEvent_t e0 = get_event(h), e1 = get_event(h), e2 = get_event(h);
    CK(cudaEventRecord(e0, c.stream));
    if (sortable) CKRC(sort_chunk(h, c));
    CK(cudaEventRecord(e1, c.stream));
    if (c.n > 0) CKRC(launch(c));
    CK(cudaGetLastError());
    CK(cudaEventRecord(e2, c.stream));
    evs.push_back(e0); evs.push_back(e1); evs.push_back(e2);
    for (int a = 0; a < nargs; ++a)
      if (args[a].out && !is_device_ptr(args[a].out))
        CK(cudaMemcpyAsync((char *)args[a].out + args[a].stride * (size_t)off, c.dout[a], args[a].stride * (size_t)c.n, cudaMemcpyDeviceToHost, c.stream));
  }
  if (h->opt_async) return TGQ_OK;          // the caller orders later work on the stream / calls tgq_synchronize before reading host outputs
  if (all_dev) CK(cudaStreamSynchronize(h->stream));
  else for (int s = 0; s < TGQ_SLOTS; ++s) CK(cudaStreamSynchronize(h->slot[s].stream));
  for (size_t i = 0; i + 2 < evs.size() + 0 && i < evs.size(); i += 3)
  {
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, evs[i], evs[i + 1])); CK(cudaEventElapsedTime(&b, evs[i], evs[i + 2]));
    h->st.last_sort_ms += a; h->st.last_query_ms += b;
  }
  return TGQ_OK;
}

// Grid of a persistent kernel: enough blocks to fill the device (resident blocks per SM x SMs), fewer for small batches.
// Also zeroes the chunk counter the kernel's warps draw work from.
template <typename K>
int persistent_grid(tgq_handle * h, K kernel, size_t smem, int64_t n, int * counter, cudaStream_t s, int * grid_out, int max_blocks = 0)
{
  auto key = std::make_pair((void const *)kernel, smem);
  // cudaFuncAttributeMaxDynamicSharedMemorySize is state of the FUNCTION (per device), shared by every handle and every launch
  // configuration of the process: it is only ever raised, so that a cached larger configuration (k-NN with a larger k, another
  // handle) never launches above a limit that a later, smaller configuration lowered.
  {
    std::lock_guard<std::mutex> lock(g_smem_mutex);
    size_t & cur = g_smem_limit[std::make_pair(h->device, (void const *)kernel)];
    if (smem > cur) { CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
  }
  auto it = h->occupancy.find(key);
  if (it == h->occupancy.end())
  {
    int per_sm = 0;
    int carve = cudaSharedmemCarveoutMaxShared;
    char const * dev_blocks = getenv("TGQ_DEV_BLOCKS");     // development aid: cap resident blocks per SM and give the rest of the array to L1
    int cap = dev_blocks ? atoi(dev_blocks) : max_blocks;   // max_blocks: kernels that run faster with fewer resident blocks and a larger L1
    if (cap > 0) carve = (int)std::min<size_t>(100, ((size_t)cap * (smem + 1024) * 100 + 233471) / 233472);
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TGQ_BLOCK, smem));
    if (cap > 0) per_sm = std::min(per_sm, cap);
    it = h->occupancy.insert(std::make_pair(key, std::max(1, per_sm))).first;
  }
  int64_t want = (n + (int64_t)TGQ_CHUNK * (TGQ_BLOCK / 32) - 1) / ((int64_t)TGQ_CHUNK * (TGQ_BLOCK / 32));
  *grid_out = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)it->second * h->sm_count));
  CK(cudaMemsetAsync(counter, 0, sizeof(int), s));
  return TGQ_OK;
}

int need_built(tgq_handle * h, int kind_mask)
{
  if (!h) return TGQ_EINVAL;
  if (h->kind == TGQ_KIND_NONE) return fail(h, TGQ_ESTATE, "no structure built on this handle");
  if (!((1 << h->kind) & kind_mask)) return fail(h, TGQ_EINVAL, "query not supported by this structure kind");
  return TGQ_OK;
}

struct Guard
{
  int prev = -1; bool ok = true;
  explicit Guard(tgq_handle * h) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != h->device) ok = cudaSetDevice(h->device) == cudaSuccess; }
  ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

// =================================================================================================================================
// C ABI
// =================================================================================================================================
extern "C" {

int tgq_create(int device, tgq_handle ** out)
{
  if (!out) { g_create_error = "tgq_create: null output pointer"; return TGQ_EINVAL; }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
  {
    g_create_error = std::string("tgq_create: no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                     "); the engine has no CPU path";
    cudaGetLastError();
    return TGQ_ENODEVICE;
  }
  if (device < 0 || device >= count) { g_create_error = "tgq_create: device index out of range"; return TGQ_EINVAL; }
  tgq_handle * h = new tgq_handle;
  h->device = device;
  int prev = -1; cudaGetDevice(&prev);
  auto bail = [&](char const * what, cudaError_t err) {
    g_create_error = std::string("tgq_create: ") + what + ": " + cudaGetErrorString(err);
    if (prev >= 0) cudaSetDevice(prev);
    delete h; return TGQ_ECUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  h->stream = h->own_stream;
  for (int s = 0; s < TGQ_SLOTS; ++s)
  {
    if ((e = cudaStreamCreateWithFlags(&h->slot[s].stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = h->slot[s].counter.reserve(4 * sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
  }
  if ((e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return bail("cudaDeviceGetAttribute", e);
  if (prev >= 0) cudaSetDevice(prev);
  *out = h;
  return TGQ_OK;
}

int tgq_destroy(tgq_handle * h)
{
  if (!h) return TGQ_OK;
  Guard g(h);
  cudaDeviceSynchronize();
  DevBuf * bufs[] = {&h->nodes, &h->elems, &h->nrm, &h->lab, &h->scene, &h->perm, &h->pos_of_elem, &h->first, &h->last, &h->split, &h->parent, &h->live,
                     &h->newidx, &h->pbox_lo, &h->pbox_hi, &h->nbox_lo, &h->nbox_hi, &h->arrive, &h->keys_a, &h->keys_b, &h->vals_a,
                     &h->cubtmp, &h->src, &h->src2, &h->misc, &h->icp_cur, &h->icp_orig, &h->icp_to, &h->icp_idx, &h->area_by_elem, &h->cum, &h->obbs, &h->pnrm, &h->nsum, &h->obbinfo, &h->big, &h->big_count, &h->wide, &h->wflag, &h->widx, &h->obbinfo_w,
                     &h->orig, &h->orig_n, &h->orig_lab, &h->xf, &h->g_idx, &h->g_sep, &h->g_dist, &h->g_stamp, &h->g_touched, &h->g_queue, &h->g_seeds,
                     &h->g_out0, &h->g_out1, &h->g_out2};
  for (DevBuf * b : bufs) b->release();
  for (int s = 0; s < TGQ_SLOTS; ++s)
  {
    Slot & sl = h->slot[s];
    for (int a = 0; a < TGQ_MAX_ARGS; ++a) { sl.in[a].release(); sl.out[a].release(); }
    sl.keys[0].release(); sl.keys[1].release(); sl.vals[0].release(); sl.vals[1].release(); sl.cubtmp.release(); sl.counter.release();
    if (sl.stream) cudaStreamDestroy(sl.stream);
  }
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return TGQ_OK;
}

const char * tgq_last_error(tgq_handle * h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int tgq_set_stream(tgq_handle * h, void * cuda_stream)
{
  if (!h) return TGQ_EINVAL;
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return TGQ_OK;
}

int tgq_synchronize(tgq_handle * h)
{
  if (!h) return TGQ_EINVAL;
  Guard g(h);
  CK(cudaStreamSynchronize(h->stream));
  for (int s = 0; s < TGQ_SLOTS; ++s) CK(cudaStreamSynchronize(h->slot[s].stream));
  return TGQ_OK;
}

int tgq_set_option(tgq_handle * h, int option, int64_t value)
{
  if (!h) return TGQ_EINVAL;
  switch (option)
  {
    case TGQ_OPT_SORT_QUERIES: if (value < 0 || value > 2) return fail(h, TGQ_EINVAL, "TGQ_OPT_SORT_QUERIES takes 0, 1 or 2"); h->opt_sort = (int)value; return TGQ_OK;
    case TGQ_OPT_LEAF_SIZE: if (value < 0 || value > TGQ_LEAF_MASK) return fail(h, TGQ_EINVAL, "TGQ_OPT_LEAF_SIZE takes 0 (by kind) or 1..15"); h->opt_leaf = (int)value; return TGQ_OK;
    case TGQ_OPT_ASYNC: h->opt_async = value ? 1 : 0; return TGQ_OK;
    case TGQ_OPT_POOL: h->opt_pool = value ? 1 : 0; return TGQ_OK;
    case TGQ_OPT_ICP_REFERENCE_ORDER: h->opt_icp_ref = value ? 1 : 0; return TGQ_OK;
    case TGQ_OPT_WIDE: h->opt_wide = value ? 1 : 0; return TGQ_OK;
    case TGQ_OPT_SET_BUFFER: h->opt_setbuf = value ? 1 : 0; return TGQ_OK;
    case TGQ_OPT_RAY_ORDER: if (value != 0 && (value < 3 || value > 30)) return fail(h, TGQ_EINVAL, "TGQ_OPT_RAY_ORDER takes 0 or 3..30"); h->opt_ray_order = (int)value; return TGQ_OK;
    case TGQ_OPT_SHARD_CELL_BITS: if (value < 3 || value > 30) return fail(h, TGQ_EINVAL, "TGQ_OPT_SHARD_CELL_BITS takes 3..30"); h->opt_cell_bits = (int)value; return TGQ_OK;
    case TGQ_OPT_CHUNK: if (value < 1024) return fail(h, TGQ_EINVAL, "TGQ_OPT_CHUNK must be >= 1024"); h->opt_chunk = value; return TGQ_OK;
    default: return fail(h, TGQ_EINVAL, "unknown option");
  }
}

int tgq_build_triangles(tgq_handle * h, const float * v9, int64_t T)
{
  if (!h) return TGQ_EINVAL;
  if (T < 0 || (T > 0 && !v9)) return fail(h, TGQ_EINVAL, "tgq_build_triangles: bad arguments");
  Guard g(h);
  void const * d = nullptr;
  if (T > 0) CKRC(to_device(h, v9, sizeof(float) * 9 * (size_t)T, h->src, h->stream, &d));
  return build_common(h, TGQ_KIND_TRIANGLES, (float const *)d, nullptr, T);
}

int tgq_build_triangles_indexed(tgq_handle * h, const float * verts3, int64_t V, const int32_t * tri3, int64_t T)
{
  if (!h) return TGQ_EINVAL;
  if (T < 0 || V < 0 || (T > 0 && (!verts3 || !tri3))) return fail(h, TGQ_EINVAL, "tgq_build_triangles_indexed: bad arguments");
  Guard g(h);
  if (T == 0) return build_common(h, TGQ_KIND_TRIANGLES, nullptr, nullptr, 0);
  void const * dv = nullptr; void const * di = nullptr;
  CKRC(to_device(h, verts3, sizeof(float) * 3 * (size_t)V, h->src2, h->stream, &dv));
  CKRC(to_device(h, tri3, sizeof(int32_t) * 3 * (size_t)T, h->misc, h->stream, &di));
  CK(h->src.reserve(sizeof(float) * 9 * (size_t)T));
  CK(h->big_count.reserve(4)); CK(cudaMemsetAsync(h->big_count.p, 0, 4, h->stream));
  k_expand_indexed<<<grid_for(3 * T), TGQ_BLOCK, 0, h->stream>>>((float const *)dv, (int const *)di, 3 * T, V, h->src.as<float>(), h->big_count.as<int>());
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, h->big_count.p, 4, cudaMemcpyDeviceToHost, h->stream)); CK(cudaStreamSynchronize(h->stream));
  if (bad) return fail(h, TGQ_EINVAL, "tgq_build_triangles_indexed: vertex index out of range");
  return build_common(h, TGQ_KIND_TRIANGLES, h->src.as<float>(), nullptr, T);
}

int tgq_build_points(tgq_handle * h, const float * p3, int64_t P, const float * n3)
{
  if (!h) return TGQ_EINVAL;
  if (P < 0 || (P > 0 && !p3)) return fail(h, TGQ_EINVAL, "tgq_build_points: bad arguments");
  Guard g(h);
  void const * d = nullptr; void const * dn = nullptr;
  if (P > 0) CKRC(to_device(h, p3, sizeof(float) * 3 * (size_t)P, h->src, h->stream, &d));
  if (P > 0 && n3) CKRC(to_device(h, n3, sizeof(float) * 3 * (size_t)P, h->src2, h->stream, &dn));
  return build_common(h, TGQ_KIND_POINTS, (float const *)d, (float const *)dn, P);
}

int tgq_set_transform(tgq_handle * h, const float * tr12)
{
  CKRC(need_built(h, (1 << TGQ_KIND_TRIANGLES) | (1 << TGQ_KIND_POINTS)));
  if (h->n == 0) return TGQ_OK;
  Guard g(h);
  cudaStream_t s = h->stream;
  int n = h->n, kind = h->kind; bool tri = kind == TGQ_KIND_TRIANGLES, has_n = h->has_normals, has_l = h->has_labels;
  size_t np = (tri ? 3 : 1) * (size_t)n;
  if (!h->orig_valid)
  {
    // first transform since the build: the records still hold the caller's elements bit for bit
    CK(h->orig.reserve(12 * np));
    if (tri) k_extract_tris<<<grid_for(n), TGQ_BLOCK, 0, s>>>(h->elems.as<TriRec>(), n, h->orig.as<float>());
    else
    {
      if (has_n) CK(h->orig_n.reserve(12 * (size_t)n));
      if (has_l) CK(h->orig_lab.reserve(4 * (size_t)n));
      k_extract_points<<<grid_for(n), TGQ_BLOCK, 0, s>>>(h->elems.as<f4>(), has_n ? h->nrm.as<f4>() : nullptr, has_l ? h->lab.as<int>() : nullptr, n,
                                                         h->orig.as<float>(), has_n ? h->orig_n.as<float>() : nullptr, has_l ? h->orig_lab.as<int>() : nullptr);
    }
    CK(cudaGetLastError());
    h->st.total_launches += 1;
  }
  else if (has_l)
  {
    // labels may have been (re)set since the originals were taken: refresh them from the current Morton order
    CK(h->orig_lab.reserve(4 * (size_t)n));
    k_extract_points<<<grid_for(n), TGQ_BLOCK, 0, s>>>(h->elems.as<f4>(), nullptr, h->lab.as<int>(), n, nullptr, nullptr, h->orig_lab.as<int>());
    CK(cudaGetLastError());
    h->st.total_launches += 1;
  }
  float const * src = h->orig.as<float>();
  if (tr12)
  {
    CK(h->xf.reserve(12 * np)); CK(h->misc.reserve(64));
    CK(cudaMemcpyAsync(h->misc.p, tr12, 48, is_device_ptr(tr12) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    k_transform<<<grid_for((int64_t)np), TGQ_BLOCK, 0, s>>>(h->orig.as<float>(), (int64_t)np, h->misc.as<float>(), h->xf.as<float>());
    CK(cudaGetLastError());
    h->st.total_launches += 1;
    src = h->xf.as<float>();
  }
  int rc = build_common(h, kind, src, (!tri && has_n) ? h->orig_n.as<float>() : nullptr, n);
  if (rc != TGQ_OK) return rc;
  h->orig_valid = true;
  if (has_l)
  {
    CK(h->lab.reserve(4 * (size_t)n));
    k_gather_labels<<<grid_for(n), TGQ_BLOCK, 0, s>>>(h->orig_lab.as<int>(), h->perm.as<uint32_t>(), n, h->lab.as<int>());
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    h->st.total_launches += 1;
    h->has_labels = true;
  }
  return TGQ_OK;
}

int tgq_refit_triangles(tgq_handle * h, const float * v9, int64_t T)
{
  CKRC(need_built(h, 1 << TGQ_KIND_TRIANGLES));
  if (T != h->n) return fail(h, TGQ_ESTATE, "tgq_refit_triangles: element count differs from the built structure");
  if (T == 0) return TGQ_OK;
  if (!v9) return fail(h, TGQ_EINVAL, "tgq_refit_triangles: null vertices");
  Guard g(h);
  cudaStream_t s = h->stream;
  h->ev_used = 0; cudaEvent_t e0 = get_event(h), e1 = get_event(h);
  CK(cudaEventRecord(e0, s));
  void const * d = nullptr;
  CKRC(to_device(h, v9, sizeof(float) * 9 * (size_t)T, h->src, s, &d));
  CKRC(scene_bounds(h, (float const *)d, 3 * T, s));
  k_gather_tris<<<grid_for(h->n), TGQ_BLOCK, 0, s>>>((float const *)d, h->perm.as<uint32_t>(), h->n, h->scene.as<SceneBox>(), h->elems.as<TriRec>(),
                                                     h->pbox_lo.as<f4>(), h->pbox_hi.as<f4>(), h->pos_of_elem.as<int>(), h->pnrm.as<f4>(),
                                                     h->area_by_elem.as<float>());
  CK(cudaGetLastError());
  h->st.total_launches += 1; h->cum_valid = false; h->orig_valid = false;
  CKRC(refit_and_emit(h, s));
  CK(cudaEventRecord(e1, s));
  CKRC(fetch_scene(h, s));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); h->st.last_build_ms = ms;
  return TGQ_OK;
}

int tgq_info(tgq_handle * h, int32_t * kind, int64_t * num_elements, float * lo3, float * hi3)
{
  if (!h) return TGQ_EINVAL;
  if (kind) *kind = h->kind;
  if (num_elements) *num_elements = h->n;
  for (int a = 0; a < 3; ++a) { if (lo3) lo3[a] = h->lo[a]; if (hi3) hi3[a] = h->hi[a]; }
  return TGQ_OK;
}

int tgq_triangle_cache(tgq_handle * h, float * n3, float * d, int32_t * axis, float * area)
{
  CKRC(need_built(h, 1 << TGQ_KIND_TRIANGLES));
  if (h->n == 0) return TGQ_OK;
  Guard g(h);
  cudaStream_t s = h->stream; size_t nn = (size_t)h->n;
  // device staging for host outputs: [n3 | d | axis | area]
  CK(h->misc.reserve(4 * nn * 6));
  float * dn = h->misc.as<float>(); float * dd = dn + 3 * nn; int * da = (int *)(dd + nn); float * dar = (float *)(da + nn);
  float * on = n3 ? (is_device_ptr(n3) ? n3 : dn) : nullptr; float * od = d ? (is_device_ptr(d) ? d : dd) : nullptr;
  int * oa = axis ? (is_device_ptr(axis) ? axis : da) : nullptr; float * oar = area ? (is_device_ptr(area) ? area : dar) : nullptr;
  k_scatter_cache<<<grid_for(h->n), TGQ_BLOCK, 0, s>>>(h->elems.as<TriRec>(), h->area_by_elem.as<float>(), h->n, on, od, oa, oar);
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  if (n3 && on == dn) CKRC(from_device(h, n3, dn, 12 * nn, s));
  if (d && od == dd) CKRC(from_device(h, d, dd, 4 * nn, s));
  if (axis && oa == da) CKRC(from_device(h, axis, da, 4 * nn, s));
  if (area && oar == dar) CKRC(from_device(h, area, dar, 4 * nn, s));
  CK(cudaStreamSynchronize(s));
  return TGQ_OK;
}

static int closest_common(tgq_handle * h, const float * q3, const float * qn3, const int32_t * qlabel, float min_dot, int64_t Q, double dist_bound,
                          int32_t exclude_self, int32_t * idx, float * d2, float * cp3, bool plain_bound = false)
{
  CKRC(need_built(h, (1 << TGQ_KIND_TRIANGLES) | (1 << TGQ_KIND_POINTS)));
  if (Q < 0 || (Q > 0 && !q3)) return fail(h, TGQ_EINVAL, "tgq_closest: bad arguments");
  Guard g(h);
  float ub; BoundGate gate; closest_bound_gate(h->lo, h->hi, dist_bound, ub, gate);     // BvhN::closestPair's root rejection and bound widening
  if (plain_bound) { ub = bound_to_ub(dist_bound); gate = no_gate(); }                  // structure-to-structure queries: the bound as given
  Arg args[6] = {{q3, nullptr, 12}, {nullptr, idx, 4}, {nullptr, d2, 4}, {nullptr, cp3, 12}, {qn3, nullptr, 12}, {qlabel, nullptr, 4}};
  BvhView v = view_of(h); bool tri = h->kind == TGQ_KIND_TRIANGLES;
  size_t smem = trav_smem_bytes(1);
  return run_batched(h, Q, args, 6, true, [&](ChunkCtx & c) -> int {
    int grid = 0; int * cnt = c.slot->counter.as<int>();
    if (tri && h->opt_wide && h->wide_built && h->leaf_built <= 4)
    {
      size_t wsmem = wide_smem_bytes();
      CKRC(persistent_grid(h, k_closest_wide, wsmem, c.n, cnt, c.stream, &grid));
      k_closest_wide<<<grid, TGQ_BLOCK, wsmem, c.stream>>>(v, h->pad, (float const *)c.din[0], c.perm, c.n, ub, gate, exclude_self, c.offset, cnt,
                                                          (int *)c.dout[1], (float *)c.dout[2], (float *)c.dout[3]);
      if (h->peers.n) return fail(h, TGQ_EINVAL, "record outputs are not available with TGQ_OPT_WIDE");
    }
    else if (tri && h->opt_pool && h->leaf_built <= 4)
    {
      size_t psmem = smem + (TGQ_BLOCK / 32) * sizeof(PoolWarp);
      CKRC(persistent_grid(h, k_closest_pool, psmem, c.n, cnt, c.stream, &grid));
      k_closest_pool<<<grid, TGQ_BLOCK, psmem, c.stream>>>(v, h->pad, (float const *)c.din[0], c.perm, c.n, ub, gate, exclude_self, c.offset, cnt,
                                                          (int *)c.dout[1], (float *)c.dout[2], (float *)c.dout[3], h->peers);
    }
    else if (tri)
    {
      CKRC(persistent_grid(h, k_closest<true>, smem, c.n, cnt, c.stream, &grid));
      k_closest<true><<<grid, TGQ_BLOCK, smem, c.stream>>>(v, h->pad, (float const *)c.din[0], c.perm, c.n, ub, gate, exclude_self, c.offset, cnt,
                                                         (int *)c.dout[1], (float *)c.dout[2], (float *)c.dout[3], nullptr, nullptr, 0.0f);
    }
    else
    {
      CKRC(persistent_grid(h, k_closest<false>, smem, c.n, cnt, c.stream, &grid));
      k_closest<false><<<grid, TGQ_BLOCK, smem, c.stream>>>(v, 0.0f, (float const *)c.din[0], c.perm, c.n, ub, gate, exclude_self, c.offset, cnt,
                                                          (int *)c.dout[1], (float *)c.dout[2], (float *)c.dout[3], (float const *)c.din[4],
                                                          (int const *)c.din[5], min_dot);
    }
    h->st.last_launches += 1; h->st.total_launches += 1;
    return TGQ_OK;
  });
}

int tgq_closest(tgq_handle * h, const float * q3, int64_t Q, double dist_bound, int32_t exclude_self, int32_t * idx, float * d2, float * cp3)
{ return closest_common(h, q3, nullptr, nullptr, 0.0f, Q, dist_bound, exclude_self, idx, d2, cp3); }

// Validates the filter arguments of the *_filtered entry points against the structure.
static int check_filter(tgq_handle * h, const float * qn3, const int32_t * qlabel)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  if (qn3 && !h->has_normals) return fail(h, TGQ_ESTATE, "filtered query with normals on a structure built without normals");
  if (qlabel && !h->has_labels) return fail(h, TGQ_ESTATE, "filtered query with labels on a structure without labels (tgq_set_point_labels)");
  return TGQ_OK;
}

int tgq_closest_filtered(tgq_handle * h, const float * q3, const float * qn3, const int32_t * qlabel, int64_t Q, float min_dot, double dist_bound,
                         int32_t * idx, float * d2, float * cp3)
{
  if (!h) return TGQ_EINVAL;
  CKRC(check_filter(h, qn3, qlabel));
  return closest_common(h, q3, qn3, qlabel, min_dot, Q, dist_bound, 0, idx, d2, cp3);
}

int tgq_closest_segment(tgq_handle * h, const float * a3, const float * b3, int32_t b_is_direction, int64_t Q, double dist_bound, const int32_t * exclude, int32_t n_exclude,
                        int32_t * idx, float * d2, float * element_point3, float * segment_point3)
{
  CKRC(need_built(h, (1 << TGQ_KIND_TRIANGLES) | (1 << TGQ_KIND_POINTS)));
  if (Q < 0 || (Q > 0 && (!a3 || !b3))) return fail(h, TGQ_EINVAL, "tgq_closest_segment: bad arguments");
  if (n_exclude < 0 || n_exclude > TGQ_SEG_EXCL || (n_exclude > 0 && !exclude)) return fail(h, TGQ_ELIMIT, "tgq_closest_segment: at most 4 excluded elements per query");
  Guard g(h);
  float ub; BoundGate gate; closest_bound_gate(h->lo, h->hi, dist_bound, ub, gate);     // BvhN::closestPair's root rejection and bound widening
  if (n_exclude == 0) exclude = nullptr;
  Arg args[7] = {{a3, nullptr, 12}, {b3, nullptr, 12}, {exclude, nullptr, 4 * (size_t)std::max(1, n_exclude)}, {nullptr, idx, 4}, {nullptr, d2, 4},
                 {nullptr, element_point3, 12}, {nullptr, segment_point3, 12}};
  BvhView v = view_of(h); bool tri = h->kind == TGQ_KIND_TRIANGLES;
  size_t smem = trav_smem_bytes(2);
  return run_batched(h, Q, args, 7, true, [&](ChunkCtx & c) -> int {
    int grid = 0; int * cnt = c.slot->counter.as<int>();
    if (tri)
    {
      CKRC(persistent_grid(h, k_closest_seg<true>, smem, c.n, cnt, c.stream, &grid));
      k_closest_seg<true><<<grid, TGQ_BLOCK, smem, c.stream>>>(v, h->pad, (float const *)c.din[0], (float const *)c.din[1], c.perm, c.n, b_is_direction, ub, gate,
                                                             (int const *)c.din[2], n_exclude, cnt, (int *)c.dout[3], (float *)c.dout[4],
                                                             (float *)c.dout[5], (float *)c.dout[6]);
    }
    else
    {
      CKRC(persistent_grid(h, k_closest_seg<false>, smem, c.n, cnt, c.stream, &grid));
      k_closest_seg<false><<<grid, TGQ_BLOCK, smem, c.stream>>>(v, h->pad, (float const *)c.din[0], (float const *)c.din[1], c.perm, c.n, b_is_direction, ub, gate,
                                                              (int const *)c.din[2], n_exclude, cnt, (int *)c.dout[3], (float *)c.dout[4],
                                                              (float *)c.dout[5], (float *)c.dout[6]);
    }
    h->st.last_launches += 1; h->st.total_launches += 1;
    return TGQ_OK;
  });
}

// Q triangle query objects against the structure (plain_bound: the bound as given, without closestPair's root rejection and widening --
// the structure-to-structure form).
static int closest_triangle_common(tgq_handle * h, const float * q9, int64_t Q, double dist_bound, int32_t * idx, float * d2, float * element_point3,
                                   float * query_point3, bool plain_bound)
{
  CKRC(need_built(h, (1 << TGQ_KIND_TRIANGLES) | (1 << TGQ_KIND_POINTS)));
  if (Q < 0 || (Q > 0 && !q9)) return fail(h, TGQ_EINVAL, "tgq_closest_triangle: bad arguments");
  Guard g(h);
  float ub; BoundGate gate; closest_bound_gate(h->lo, h->hi, dist_bound, ub, gate);
  if (plain_bound) { ub = bound_to_ub(dist_bound); gate = no_gate(); }
  BvhView v = view_of(h); bool tri = h->kind == TGQ_KIND_TRIANGLES;
  float elem_tau = 0.0f;
  if (tri && h->n > 0)
  {
    // FLT_EPSILON over the smallest positive |N| of the elements: how far apart Moeller's test can still report an intersection
    CK(h->misc.reserve(64));
    int * bits = h->misc.as<int>();
    int inf_bits = 0x7f800000;
    CK(cudaMemcpyAsync(bits, &inf_bits, 4, cudaMemcpyHostToDevice, h->stream));
    k_min_normal<<<std::min(grid_for(h->n, 256), h->sm_count * 8), 256, 0, h->stream>>>(v.tris, h->n, bits);
    CK(cudaGetLastError());
    int got = 0;
    CK(cudaMemcpyAsync(&got, bits, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float mn; memcpy(&mn, &got, 4);
    elem_tau = (mn > 0.0f && mn < INFINITY) ? FLT_EPSILON / mn : 0.0f;
    h->st.total_launches += 1;
  }
  Arg args[5] = {{q9, nullptr, 36}, {nullptr, idx, 4}, {nullptr, d2, 4}, {nullptr, element_point3, 12}, {nullptr, query_point3, 12}};
  size_t smem = trav_smem_bytes(3);
  h->sort_stride3 = 3;                      // the launch order follows the Morton code of each query triangle's first vertex
  struct Restore { tgq_handle * h; ~Restore() { h->sort_stride3 = 1; } } restore{h};
  return run_batched(h, Q, args, 5, true, [&](ChunkCtx & c) -> int {
    int grid = 0; int * cnt = c.slot->counter.as<int>();
    if (tri)
    {
      CKRC(persistent_grid(h, k_closest_tri<true>, smem, c.n, cnt, c.stream, &grid));
      k_closest_tri<true><<<grid, TGQ_BLOCK, smem, c.stream>>>(v, h->pad, elem_tau, (float const *)c.din[0], c.perm, c.n, ub, gate, cnt, (int *)c.dout[1],
                                                             (float *)c.dout[2], (float *)c.dout[3], (float *)c.dout[4]);
    }
    else
    {
      CKRC(persistent_grid(h, k_closest_tri<false>, smem, c.n, cnt, c.stream, &grid));
      k_closest_tri<false><<<grid, TGQ_BLOCK, smem, c.stream>>>(v, h->pad, elem_tau, (float const *)c.din[0], c.perm, c.n, ub, gate, cnt, (int *)c.dout[1],
                                                              (float *)c.dout[2], (float *)c.dout[3], (float *)c.dout[4]);
    }
    h->st.last_launches += 1; h->st.total_launches += 1;
    return TGQ_OK;
  });
}

int tgq_closest_triangle(tgq_handle * h, const float * q9, int64_t Q, double dist_bound, int32_t * idx, float * d2, float * element_point3,
                         float * query_point3)
{ return closest_triangle_common(h, q9, Q, dist_bound, idx, d2, element_point3, query_point3, false); }

int tgq_closest_pair(tgq_handle * h, tgq_handle * query, double dist_bound, int64_t * target_index, int64_t * query_index, float * d2, float * query_point3,
                     float * target_point3)
{
  CKRC(need_built(h, (1 << TGQ_KIND_TRIANGLES) | (1 << TGQ_KIND_POINTS)));
  if (!query || !target_index || !query_index) return fail(h, TGQ_EINVAL, "tgq_closest_pair: bad arguments");
  if (query->kind != TGQ_KIND_POINTS && query->kind != TGQ_KIND_TRIANGLES) return fail(h, TGQ_ESTATE, "tgq_closest_pair: no structure built on the query handle");
  if (query->device != h->device) return fail(h, TGQ_EINVAL, "tgq_closest_pair: both structures must live on the same device");
  *target_index = -1; *query_index = -1; if (d2) *d2 = -1.0f;
  if (query->n == 0 || h->n == 0) return TGQ_OK;
  Guard g(h);
  cudaStream_t s = h->stream;
  CK(cudaStreamSynchronize(query->stream));
  bool const qtri = query->kind == TGQ_KIND_TRIANGLES;
  size_t P = (size_t)query->n, qw = qtri ? 9 : 3;
  // the query structure's elements in element order, answered one by one by this structure, then the canonical minimum
  CK(h->src2.reserve(4 * qw * P)); CK(h->slot[0].out[4].reserve(4 * P)); CK(h->slot[0].out[5].reserve(4 * P)); CK(h->slot[0].out[6].reserve(12 * P));
  CK(h->slot[0].out[7].reserve(12 * P));
  if (qtri) k_extract_tris<<<grid_for((int64_t)P), TGQ_BLOCK, 0, s>>>(query->elems.as<TriRec>(), (int)P, h->src2.as<float>());
  else k_extract_points<<<grid_for((int64_t)P), TGQ_BLOCK, 0, s>>>(query->elems.as<f4>(), nullptr, nullptr, (int)P, h->src2.as<float>(), nullptr, nullptr);
  CK(cudaGetLastError());
  int * di = h->slot[0].out[4].as<int>(); float * dd = h->slot[0].out[5].as<float>(); float * dc = h->slot[0].out[6].as<float>();
  float * dq = h->slot[0].out[7].as<float>();
  int saved_async = h->opt_async; h->opt_async = 0;
  int rc = qtri ? closest_triangle_common(h, h->src2.as<float>(), (int64_t)P, dist_bound, di, dd, dc, dq, true)
                : closest_common(h, h->src2.as<float>(), nullptr, nullptr, 0.0f, (int64_t)P, dist_bound, 0, di, dd, dc, true);
  h->opt_async = saved_async;
  if (rc != TGQ_OK) return rc;
  CK(h->misc.reserve(64));
  unsigned long long * best = h->misc.as<unsigned long long>();
  CK(cudaMemsetAsync(best, 0xff, 16, s));
  int grid = (int)std::min<int64_t>((int64_t)(P + 255) / 256, (int64_t)h->sm_count * 8);
  k_pair_min_key<<<grid, 256, 0, s>>>(di, dd, (int64_t)P, best);
  k_pair_min_query<<<grid, 256, 0, s>>>(di, dd, (int64_t)P, best);
  CK(cudaGetLastError());
  h->st.total_launches += 3;
  unsigned long long hb[2];
  CK(cudaMemcpyAsync(hb, best, 16, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (hb[0] == ~0ull || hb[1] == ~0ull) return TGQ_OK;
  *target_index = (int64_t)(unsigned)(hb[0] & 0xffffffffull); *query_index = (int64_t)hb[1];
  float hd = 0, hq[3], ht[3];
  CK(cudaMemcpyAsync(&hd, dd + hb[1], 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(hq, (qtri ? dq : h->src2.as<float>()) + 3 * hb[1], 12, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(ht, dc + 3 * hb[1], 12, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (d2) *d2 = hd;
  for (int a = 0; a < 3; ++a) { if (query_point3) query_point3[a] = hq[a]; if (target_point3) target_point3[a] = ht[a]; }
  return TGQ_OK;
}

// ---- sample-graph searches (tgq_graph.cuh) --------------------------------------------------------------------------------------
// Uploads the graph if it lives on the host, sizes the per-warp scratch rows (at most `budget` bytes) and initialises them.
static int graph_setup(tgq_handle * h, const int32_t * idx, const float * sep, int64_t P, int32_t K, int64_t sources, GraphView * g, GraphScratch * sc,
                       int * blocks)
{
  if (P < 1 || P > 0x7fffffff || K < 1 || K > TGQ_MAX_K || !idx || !sep) return fail(h, TGQ_EINVAL, "graph query: bad graph arguments");
  cudaStream_t s = h->stream;
  void const * di = nullptr; void const * ds = nullptr;
  CKRC(to_device(h, idx, 4 * (size_t)P * K, h->g_idx, s, &di));
  CKRC(to_device(h, sep, 4 * (size_t)P * K, h->g_sep, s, &ds));
  g->idx = (int const *)di; g->sep = (float const *)ds; g->P = (int)P; g->K = K;
  if (!sc) return TGQ_OK;
  size_t const row = 24 * (size_t)P;                                     // 8 (distance) + 4 (stamp) + 4 (touched) + 8 (two queues) bytes per node
  size_t const budget = (size_t)16 << 30;
  int64_t workers = std::min<int64_t>((int64_t)h->sm_count * 2 * TGQ_GRAPH_WARPS, std::max<int64_t>(TGQ_GRAPH_WARPS, (int64_t)(budget / row)));
  workers = std::min<int64_t>(workers, (sources + TGQ_GRAPH_WARPS - 1) / TGQ_GRAPH_WARPS * TGQ_GRAPH_WARPS);
  workers = std::max<int64_t>(TGQ_GRAPH_WARPS, workers / TGQ_GRAPH_WARPS * TGQ_GRAPH_WARPS);
  *blocks = (int)(workers / TGQ_GRAPH_WARPS);
  size_t const n = (size_t)workers * (size_t)P;
  sc->qlen = (int)std::max<int64_t>(P, 2 * TGQ_MAX_K);
  CK(h->g_dist.reserve(8 * n)); CK(h->g_stamp.reserve(4 * n)); CK(h->g_touched.reserve(4 * n)); CK(h->g_queue.reserve(8 * (size_t)workers * (size_t)sc->qlen));
  sc->dist = h->g_dist.as<unsigned long long>(); sc->stamp = h->g_stamp.as<int>(); sc->touched = h->g_touched.as<int>(); sc->queue = h->g_queue.as<int>();
  k_graph_init<<<h->sm_count * 8, 256, 0, s>>>(sc->dist, sc->stamp, n);
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  return TGQ_OK;
}

int tgq_graph_contract(tgq_handle * h, const int32_t * idx, const float * sep, int64_t P, int32_t max_degree, int64_t n_orig, int32_t * out_idx,
                       float * out_sep, int32_t * out_count)
{
  if (!h) return TGQ_EINVAL;
  if (n_orig < 0 || n_orig > P || (n_orig > 0 && (!out_idx || !out_sep))) return fail(h, TGQ_EINVAL, "tgq_graph_contract: bad arguments");
  if (n_orig == 0) return TGQ_OK;
  Guard g(h);
  cudaStream_t s = h->stream;
  GraphView gv; GraphScratch sc; int blocks = 0;
  CKRC(graph_setup(h, idx, sep, P, max_degree, n_orig, &gv, &sc, &blocks));
  size_t const nk = (size_t)n_orig * max_degree;
  CK(h->g_out0.reserve(4 * nk)); CK(h->g_out1.reserve(4 * nk)); CK(h->g_out2.reserve(4 * (size_t)n_orig)); CK(h->g_seeds.reserve(4));
  int * di = is_device_ptr(out_idx) ? out_idx : h->g_out0.as<int>();
  float * ds = is_device_ptr(out_sep) ? out_sep : h->g_out1.as<float>();
  int * dc = out_count ? (is_device_ptr(out_count) ? out_count : h->g_out2.as<int>()) : nullptr;
  CK(cudaMemsetAsync(h->g_seeds.p, 0, 4, s));
  k_graph_contract<<<blocks, 32 * TGQ_GRAPH_WARPS, 0, s>>>(gv, sc, (int)n_orig, h->g_seeds.as<int>(), di, ds, dc);
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  if (!is_device_ptr(out_idx)) CKRC(from_device(h, out_idx, di, 4 * nk, s));
  if (!is_device_ptr(out_sep)) CKRC(from_device(h, out_sep, ds, 4 * nk, s));
  if (out_count && !is_device_ptr(out_count)) CKRC(from_device(h, out_count, dc, 4 * (size_t)n_orig, s));
  CK(cudaStreamSynchronize(s));
  return TGQ_OK;
}

int tgq_graph_geodesic(tgq_handle * h, const int32_t * idx, const float * sep, int64_t P, int32_t max_degree, const int32_t * seeds, int64_t Q,
                       float max_distance, float hist_range, int32_t num_bins, double * sum, int32_t * count, uint32_t * hist)
{
  if (!h) return TGQ_EINVAL;
  if (Q < 0 || Q > 0x7fffffff || (Q > 0 && !seeds) || num_bins < 0 || (hist && num_bins < 1)) return fail(h, TGQ_EINVAL, "tgq_graph_geodesic: bad arguments");
  if (Q == 0) return TGQ_OK;
  Guard g(h);
  cudaStream_t s = h->stream;
  GraphView gv; GraphScratch sc; int blocks = 0;
  CKRC(graph_setup(h, idx, sep, P, max_degree, Q, &gv, &sc, &blocks));
  void const * dseeds = nullptr;
  CK(h->g_seeds.reserve(4 * (size_t)Q + 16));
  int * cnt = h->g_seeds.as<int>();                                       // [0]: work counter, [4..]: seeds when they come from the host
  if (is_device_ptr(seeds)) dseeds = seeds;
  else { CK(cudaMemcpyAsync(cnt + 4, seeds, 4 * (size_t)Q, cudaMemcpyHostToDevice, s)); dseeds = cnt + 4; }
  CK(cudaMemsetAsync(cnt, 0, 4, s));
  size_t const nb = (size_t)Q * (size_t)(hist ? num_bins : 0);
  CK(h->g_out0.reserve(8 * (size_t)Q)); CK(h->g_out1.reserve(4 * (size_t)Q)); if (hist) CK(h->g_out2.reserve(4 * nb));
  double * dsum = sum ? (is_device_ptr(sum) ? sum : h->g_out0.as<double>()) : nullptr;
  int * dcount = count ? (is_device_ptr(count) ? count : h->g_out1.as<int>()) : nullptr;
  unsigned * dhist = hist ? (is_device_ptr(hist) ? hist : h->g_out2.as<unsigned>()) : nullptr;
  if (dhist) CK(cudaMemsetAsync(dhist, 0, 4 * nb, s));
  k_graph_geodesic<<<blocks, 32 * TGQ_GRAPH_WARPS, 0, s>>>(gv, sc, (int const *)dseeds, (int)Q, max_distance, hist_range, num_bins, cnt, dsum, dcount, dhist);
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  if (sum && !is_device_ptr(sum)) CKRC(from_device(h, sum, dsum, 8 * (size_t)Q, s));
  if (count && !is_device_ptr(count)) CKRC(from_device(h, count, dcount, 4 * (size_t)Q, s));
  if (hist && !is_device_ptr(hist)) CKRC(from_device(h, hist, dhist, 4 * nb, s));
  CK(cudaStreamSynchronize(s));
  return TGQ_OK;
}

int tgq_graph_fps(tgq_handle * h, const int32_t * idx, const float * sep, int64_t P, int32_t max_degree, int64_t n_desired, int32_t * selected,
                  int64_t * n_selected)
{
  if (!h) return TGQ_EINVAL;
  if (n_desired < 0 || n_desired > P || (n_desired > 0 && !selected) || !n_selected) return fail(h, TGQ_EINVAL, "tgq_graph_fps: bad arguments");
  *n_selected = 0;
  if (n_desired == 0) return TGQ_OK;
  Guard g(h);
  cudaStream_t s = h->stream;
  GraphView gv; int blocks = 0;
  CKRC(graph_setup(h, idx, sep, P, max_degree, 1, &gv, nullptr, &blocks));
  CK(h->g_dist.reserve(8 * (size_t)P)); CK(h->g_stamp.reserve(4 * (size_t)P)); CK(h->g_queue.reserve(8 * (size_t)P));
  CK(h->g_out0.reserve(4 * (size_t)n_desired)); CK(h->g_seeds.reserve(4));
  int * dsel = is_device_ptr(selected) ? selected : h->g_out0.as<int>();
  k_graph_fps<<<1, TGQ_FPS_THREADS, 0, s>>>(gv, h->g_dist.as<unsigned long long>(), h->g_stamp.as<int>(), h->g_queue.as<int>(), h->g_queue.as<int>() + P,
                                           (int)n_desired, dsel, h->g_seeds.as<int>());
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  int n = 0;
  CK(cudaMemcpyAsync(&n, h->g_seeds.p, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (!is_device_ptr(selected) && n > 0) { CKRC(from_device(h, selected, dsel, 4 * (size_t)n, s)); CK(cudaStreamSynchronize(s)); }
  *n_selected = n;
  return TGQ_OK;
}

int tgq_set_point_labels(tgq_handle * h, const int32_t * labels)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  Guard g(h);
  if (!labels || h->n == 0) { h->has_labels = false; return TGQ_OK; }
  cudaStream_t s = h->stream;
  void const * dl = nullptr;
  CKRC(to_device(h, labels, 4 * (size_t)h->n, h->src2, s, &dl));
  CK(h->lab.reserve(4 * (size_t)h->n));
  k_gather_labels<<<grid_for(h->n), TGQ_BLOCK, 0, s>>>((int const *)dl, h->perm.as<uint32_t>(), h->n, h->lab.as<int>());
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  h->st.total_launches += 1;
  h->has_labels = true;
  return TGQ_OK;
}

static int knn_common(tgq_handle * h, const float * q3, const float * qn3, const int32_t * qlabel, float min_dot, int64_t Q, int32_t k,
                      double dist_bound, int32_t exclude_self, int32_t * idx, float * d2)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  if (Q < 0 || (Q > 0 && !q3) || k < 1) return fail(h, TGQ_EINVAL, "tgq_knn: bad arguments");
  if (k > TGQ_MAX_K) return fail(h, TGQ_ELIMIT, "tgq_knn: k exceeds TGQ_MAX_K");
  Guard g(h);
  float ub = bound_to_ub(dist_bound);
  Arg args[5] = {{q3, nullptr, 12}, {nullptr, idx, 4 * (size_t)k}, {nullptr, d2, 4 * (size_t)k}, {qn3, nullptr, 12}, {qlabel, nullptr, 4}};
  BvhView v = view_of(h);
  size_t smem = trav_smem_bytes(KnnW::STAGE) + 2 * (size_t)k * TGQ_BLOCK * 4;
  int const cap = cand_cap(k), epl = (cap + 31) / 32;
  size_t bsmem = trav_smem_bytes(1, TGQ_BUF_SSTACK) + (size_t)cap * TGQ_BLOCK * 8;
  bool const buffered = h->opt_setbuf != 0;
  return run_batched(h, Q, args, 5, true, [&](ChunkCtx & c) -> int {
    int grid = 0; int * cnt = c.slot->counter.as<int>();
#define TGQ_KNN_BUF(E) { CKRC(persistent_grid(h, k_knn_buf<E>, bsmem, c.n, cnt, c.stream, &grid)); \
      k_knn_buf<E><<<grid, TGQ_BLOCK, bsmem, c.stream>>>(v, (float const *)c.din[0], c.perm, c.n, k, ub, exclude_self, c.offset, cnt, (int *)c.dout[1], \
                                                         (float *)c.dout[2], (float const *)c.din[3], (int const *)c.din[4], min_dot); }
    if (buffered && epl == 1) TGQ_KNN_BUF(1)
    else if (buffered && epl == 2) TGQ_KNN_BUF(2)
    else if (buffered) TGQ_KNN_BUF(3)
#undef TGQ_KNN_BUF
    else
    {
    CKRC(persistent_grid(h, k_knn, smem, c.n, cnt, c.stream, &grid));
    k_knn<<<grid, TGQ_BLOCK, smem, c.stream>>>(v, (float const *)c.din[0], c.perm, c.n, k, ub, exclude_self, c.offset, cnt, (int *)c.dout[1],
                                               (float *)c.dout[2], (float const *)c.din[3], (int const *)c.din[4], min_dot);
    }
    h->st.last_launches += 1; h->st.total_launches += 1;
    return TGQ_OK;
  });
}

int tgq_knn(tgq_handle * h, const float * q3, int64_t Q, int32_t k, double dist_bound, int32_t exclude_self, int32_t * idx, float * d2)
{ return knn_common(h, q3, nullptr, nullptr, 0.0f, Q, k, dist_bound, exclude_self, idx, d2); }

int tgq_knn_filtered(tgq_handle * h, const float * q3, const float * qn3, const int32_t * qlabel, int64_t Q, int32_t k, float min_dot,
                     double dist_bound, int32_t * idx, float * d2)
{
  if (!h) return TGQ_EINVAL;
  CKRC(check_filter(h, qn3, qlabel));
  return knn_common(h, q3, qn3, qlabel, min_dot, Q, k, dist_bound, 0, idx, d2);
}

int tgq_k_closest_pairs(tgq_handle * h, tgq_handle * query, int32_t k, double dist_bound, int64_t * target_index, int64_t * query_index, float * d2,
                        int32_t * n_found)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  if (!query || k < 1 || !target_index || !query_index || !n_found) return fail(h, TGQ_EINVAL, "tgq_k_closest_pairs: bad arguments");
  if (k > TGQ_MAX_K) return fail(h, TGQ_ELIMIT, "tgq_k_closest_pairs: k exceeds TGQ_MAX_K");
  if (query->kind != TGQ_KIND_POINTS) return fail(h, TGQ_EINVAL, "tgq_k_closest_pairs: the query structure must be a point structure");
  if (query->device != h->device) return fail(h, TGQ_EINVAL, "tgq_k_closest_pairs: both structures must live on the same device");
  *n_found = 0;
  for (int j = 0; j < k; ++j) { target_index[j] = -1; query_index[j] = -1; if (d2) d2[j] = -1.0f; }
  if (query->n == 0 || h->n == 0) return TGQ_OK;
  size_t const P = (size_t)query->n, n = P * (size_t)k;
  if (n >= ((size_t)1 << 32)) return fail(h, TGQ_ELIMIT, "tgq_k_closest_pairs: query elements x k exceeds 2^32");
  Guard g(h);
  cudaStream_t s = h->stream;
  CK(cudaStreamSynchronize(query->stream));
  CK(h->src2.reserve(12 * P)); CK(h->slot[0].out[4].reserve(4 * n)); CK(h->slot[0].out[5].reserve(4 * n));
  k_extract_points<<<grid_for((int64_t)P), TGQ_BLOCK, 0, s>>>(query->elems.as<f4>(), nullptr, nullptr, (int)P, h->src2.as<float>(), nullptr, nullptr);
  CK(cudaGetLastError());
  int * di = h->slot[0].out[4].as<int>(); float * dd = h->slot[0].out[5].as<float>();
  int saved_async = h->opt_async; h->opt_async = 0;
  int rc = knn_common(h, h->src2.as<float>(), nullptr, nullptr, 0.0f, (int64_t)P, k, dist_bound, 0, di, dd);
  h->opt_async = saved_async;
  if (rc != TGQ_OK) return rc;
  Slot & sl = h->slot[0];
  CK(sl.keys[0].reserve(8 * n)); CK(sl.keys[1].reserve(8 * n)); CK(sl.vals[0].reserve(4 * n)); CK(sl.vals[1].reserve(4 * n));
  int grid = (int)std::min<int64_t>((int64_t)(n + 255) / 256, (int64_t)h->sm_count * 8);
  k_pair_keys<<<grid, 256, 0, s>>>(di, dd, (int64_t)n, sl.keys[0].as<unsigned long long>(), sl.vals[0].as<unsigned>());
  CK(cudaGetLastError());
  size_t tmp = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, sl.keys[0].as<unsigned long long>(), sl.keys[1].as<unsigned long long>(), sl.vals[0].as<unsigned>(),
                                     sl.vals[1].as<unsigned>(), (int64_t)n, 0, 64, s));
  CK(sl.cubtmp.reserve(tmp));
  CK(cub::DeviceRadixSort::SortPairs(sl.cubtmp.p, tmp, sl.keys[0].as<unsigned long long>(), sl.keys[1].as<unsigned long long>(), sl.vals[0].as<unsigned>(),
                                     sl.vals[1].as<unsigned>(), (int64_t)n, 0, 64, s));
  int const kk = (int)std::min<size_t>((size_t)k, n);
  CK(h->misc.reserve(20 * (size_t)k));
  long long * dt = h->misc.as<long long>(); long long * dq = dt + k; float * df = (float *)(dq + k);
  k_pair_rows<<<1, 64, 0, s>>>(sl.keys[1].as<unsigned long long>(), sl.vals[1].as<unsigned>(), k, kk, dt, dq, df);
  CK(cudaGetLastError());
  h->st.total_launches += 4;
  std::vector<long long> ht((size_t)kk), hq((size_t)kk); std::vector<float> hf((size_t)kk);
  CK(cudaMemcpyAsync(ht.data(), dt, 8 * (size_t)kk, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(hq.data(), dq, 8 * (size_t)kk, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(hf.data(), df, 4 * (size_t)kk, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  int found = 0;
  for (int j = 0; j < kk; ++j)
    if (ht[(size_t)j] >= 0) { target_index[found] = ht[(size_t)j]; query_index[found] = hq[(size_t)j]; if (d2) d2[found] = hf[(size_t)j]; ++found; }
  *n_found = found;
  return TGQ_OK;
}

static int ball_graph_common(tgq_handle * h, tgq_handle * surface, int32_t max_degree, float init_radius, int32_t rounds, float min_dot,
                             int32_t min_degree, int32_t * idx, float * sep, int32_t * count)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  if (max_degree < 1 || rounds < 1) return fail(h, TGQ_EINVAL, "tgq_ball_graph: bad arguments");
  if (max_degree > TGQ_MAX_K) return fail(h, TGQ_ELIMIT, "tgq_ball_graph: max_degree exceeds TGQ_MAX_K");
  BvhView sv = view_of(h);
  if (surface)
  {
    if (surface->kind != TGQ_KIND_TRIANGLES) return fail(h, TGQ_ESTATE, "tgq_ball_graph_surface: the surface handle holds no triangle structure");
    if (surface->device != h->device) return fail(h, TGQ_EINVAL, "tgq_ball_graph_surface: both structures must live on the same device");
    if (!h->has_normals) return fail(h, TGQ_ESTATE, "tgq_ball_graph_surface: the reachability test needs sample normals (PointSet3.hpp:639)");
    CK(cudaStreamSynchronize(surface->stream));      // the surface structure must be complete before this handle's stream reads it
    sv = view_of(surface);
  }
  Guard g(h);
  h->st.last_query_ms = 0; h->st.last_sort_ms = 0; h->st.last_launches = 0; h->ev_used = 0;
  if (h->n == 0) return TGQ_OK;
  cudaStream_t s = h->stream; size_t nn = (size_t)h->n, k = (size_t)max_degree;
  CK(h->slot[0].out[0].reserve(4 * nn * k)); CK(h->slot[0].out[1].reserve(4 * nn * k)); CK(h->slot[0].out[2].reserve(4 * nn));
  int * di = idx ? (is_device_ptr(idx) ? idx : h->slot[0].out[0].as<int>()) : nullptr;
  float * ds = sep ? (is_device_ptr(sep) ? sep : h->slot[0].out[1].as<float>()) : nullptr;
  int * dc = count ? (is_device_ptr(count) ? count : h->slot[0].out[2].as<int>()) : nullptr;
  cudaEvent_t e0 = get_event(h), e1 = get_event(h);
  CK(cudaEventRecord(e0, s));
  int grid = 0; int * cnt = h->slot[0].counter.as<int>();
  size_t smem = trav_smem_bytes(BallW<false>::STAGE) + 2 * k * TGQ_BLOCK * 4;
  if (surface)
  {
    CKRC(persistent_grid(h, k_ball_graph<true>, smem, h->n, cnt, s, &grid));
    k_ball_graph<true><<<grid, TGQ_BLOCK, smem, s>>>(view_of(h), sv, h->n, max_degree, init_radius, rounds, min_dot, min_degree, cnt, di, ds, dc);
  }
  else if (h->opt_setbuf)
  {
    int const cap = cand_cap(max_degree), epl = (cap + 31) / 32;
    size_t bsmem = trav_smem_bytes(2, TGQ_BUF_SSTACK) + (size_t)cap * TGQ_BLOCK * 8;
#define TGQ_BALL_BUF(E) { CKRC(persistent_grid(h, k_ball_graph_buf<E>, bsmem, h->n, cnt, s, &grid)); \
      k_ball_graph_buf<E><<<grid, TGQ_BLOCK, bsmem, s>>>(view_of(h), h->n, max_degree, init_radius, rounds, min_dot, min_degree, cnt, di, ds, dc); }
    if (epl == 1) TGQ_BALL_BUF(1) else if (epl == 2) TGQ_BALL_BUF(2) else TGQ_BALL_BUF(3)
#undef TGQ_BALL_BUF
  }
  else
  {
    CKRC(persistent_grid(h, k_ball_graph<false>, smem, h->n, cnt, s, &grid));
    k_ball_graph<false><<<grid, TGQ_BLOCK, smem, s>>>(view_of(h), sv, h->n, max_degree, init_radius, rounds, min_dot, min_degree, cnt, di, ds, dc);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(e1, s));
  h->st.last_launches += 1; h->st.total_launches += 1;
  if (idx && !is_device_ptr(idx)) CKRC(from_device(h, idx, di, 4 * nn * k, s));
  if (sep && !is_device_ptr(sep)) CKRC(from_device(h, sep, ds, 4 * nn * k, s));
  if (count && !is_device_ptr(count)) CKRC(from_device(h, count, dc, 4 * nn, s));
  CK(cudaStreamSynchronize(s));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); h->st.last_query_ms = ms;
  return TGQ_OK;
}

int tgq_ball_graph(tgq_handle * h, int32_t max_degree, float init_radius, int32_t rounds, float min_dot, int32_t min_degree, int32_t * idx,
                   float * sep, int32_t * count)
{ return ball_graph_common(h, nullptr, max_degree, init_radius, rounds, min_dot, min_degree, idx, sep, count); }

int tgq_ball_graph_surface(tgq_handle * h, tgq_handle * surface, int32_t max_degree, float init_radius, int32_t rounds, float min_dot,
                           int32_t min_degree, int32_t * idx, float * sep, int32_t * count)
{
  if (!h) return TGQ_EINVAL;
  if (!surface) return fail(h, TGQ_EINVAL, "tgq_ball_graph_surface: null surface handle");
  return ball_graph_common(h, surface, max_degree, init_radius, rounds, min_dot, min_degree, idx, sep, count);
}

int tgq_ball_range(tgq_handle * h, const float * c3, const float * radius, int64_t Q, int32_t max_out, int32_t exclude_self, int32_t * idx,
                   int32_t * count)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  if (Q < 0 || (Q > 0 && (!c3 || !radius)) || max_out < 0) return fail(h, TGQ_EINVAL, "tgq_ball_range: bad arguments");
  if (max_out > TGQ_MAX_K) return fail(h, TGQ_ELIMIT, "tgq_ball_range: max_out exceeds TGQ_MAX_K");
  Guard g(h);
  int k = max_out;
  Arg args[4] = {{c3, nullptr, 12}, {radius, nullptr, 4}, {nullptr, k > 0 ? idx : nullptr, 4 * (size_t)k}, {nullptr, count, 4}};
  BvhView v = view_of(h);
  size_t smem = trav_smem_bytes(RangeW::STAGE) + 2 * (size_t)k * TGQ_BLOCK * 4;
  return run_batched(h, Q, args, 4, true, [&](ChunkCtx & c) -> int {
    int grid = 0; int * cnt = c.slot->counter.as<int>();
    CKRC(persistent_grid(h, k_ball_range, smem, c.n, cnt, c.stream, &grid));
    k_ball_range<<<grid, TGQ_BLOCK, smem, c.stream>>>(v, (float const *)c.din[0], (float const *)c.din[1], c.perm, c.n, k, exclude_self, c.offset, cnt,
                                                      (int *)c.dout[2], (int *)c.dout[3]);
    h->st.last_launches += 1; h->st.total_launches += 1;
    return TGQ_OK;
  });
}

int tgq_range_query(tgq_handle * h, int32_t kind, const float * a3, const float * b, int64_t Q, int64_t * offsets, int32_t * idx, int64_t capacity,
                    int64_t * total)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  if (kind < 0 || kind > 1 || Q < 0 || capacity < 0 || !offsets || !total || (Q > 0 && (!a3 || !b)) || (capacity > 0 && !idx))
    return fail(h, TGQ_EINVAL, "tgq_range_query: bad arguments");
  if (Q >= ((int64_t)1 << 31)) return fail(h, TGQ_ELIMIT, "tgq_range_query: more than 2^31 - 1 ranges in one call");
  Guard g(h);
  cudaStream_t s = h->stream;
  *total = 0;
  if (Q == 0)
  {
    int64_t z = 0;
    if (!is_device_ptr(offsets)) offsets[0] = 0;
    else { CK(cudaMemcpyAsync(offsets, &z, 8, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s)); }
    return TGQ_OK;
  }
  size_t nq = (size_t)Q; bool box = kind == 1;
  void const * da = nullptr; void const * db = nullptr;
  CKRC(to_device(h, a3, 12 * nq, h->slot[0].in[0], s, &da));
  CKRC(to_device(h, b, (box ? 12 : 4) * nq, h->slot[0].in[1], s, &db));
  // pass 1: counts; exclusive scan -> offsets; pass 2: indices; segmented sort -> ascending element index per range
  CK(h->slot[0].out[0].reserve(4 * nq)); CK(h->slot[0].out[1].reserve(8 * (nq + 1)));
  int * dcount = h->slot[0].out[0].as<int>(); long long * doff = h->slot[0].out[1].as<long long>();
  BvhView v = view_of(h);
  size_t smem = trav_smem_bytes(2);
  int grid = 0; int * cnt = h->slot[0].counter.as<int>();
  if (box) { CKRC(persistent_grid(h, k_range_all<true, false>, smem, Q, cnt, s, &grid)); k_range_all<true, false><<<grid, TGQ_BLOCK, smem, s>>>(v, (float const *)da, (float const *)db, nullptr, (int)Q, cnt, nullptr, nullptr, dcount); }
  else     { CKRC(persistent_grid(h, k_range_all<false, false>, smem, Q, cnt, s, &grid)); k_range_all<false, false><<<grid, TGQ_BLOCK, smem, s>>>(v, (float const *)da, (float const *)db, nullptr, (int)Q, cnt, nullptr, nullptr, dcount); }
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(doff, 0, 8 * (nq + 1), s));
  size_t tmp = 0;
  CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, dcount, doff + 1, (int)Q, s));
  CK(h->cubtmp.reserve(tmp));
  CK(cub::DeviceScan::InclusiveSum(h->cubtmp.p, tmp, dcount, doff + 1, (int)Q, s));
  long long tot = 0;
  CK(cudaMemcpyAsync(&tot, doff + nq, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  *total = (int64_t)tot;
  h->st.total_launches += 3;
  CKRC(from_device(h, offsets, doff, 8 * (nq + 1), s));
  if (tot > capacity || tot == 0) { CK(cudaStreamSynchronize(s)); return tot > capacity && capacity > 0 ? fail(h, TGQ_ELIMIT, "tgq_range_query: index buffer too small (see *total)") : TGQ_OK; }
  CK(h->slot[0].out[2].reserve(4 * (size_t)tot)); CK(h->slot[0].out[3].reserve(4 * (size_t)tot));
  int * raw = h->slot[0].out[2].as<int>(); int * sorted = is_device_ptr(idx) ? idx : h->slot[0].out[3].as<int>();
  if (box) { CKRC(persistent_grid(h, k_range_all<true, true>, smem, Q, cnt, s, &grid)); k_range_all<true, true><<<grid, TGQ_BLOCK, smem, s>>>(v, (float const *)da, (float const *)db, nullptr, (int)Q, cnt, doff, raw, nullptr); }
  else     { CKRC(persistent_grid(h, k_range_all<false, true>, smem, Q, cnt, s, &grid)); k_range_all<false, true><<<grid, TGQ_BLOCK, smem, s>>>(v, (float const *)da, (float const *)db, nullptr, (int)Q, cnt, doff, raw, nullptr); }
  CK(cudaGetLastError());
  tmp = 0;
  CK(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, tmp, raw, sorted, (int)tot, (int)Q, doff, doff + 1, 0, 32, s));
  CK(h->cubtmp.reserve(tmp));
  CK(cub::DeviceSegmentedRadixSort::SortKeys(h->cubtmp.p, tmp, raw, sorted, (int)tot, (int)Q, doff, doff + 1, 0, 32, s));
  h->st.total_launches += 2;
  if (sorted != idx) CKRC(from_device(h, idx, sorted, 4 * (size_t)tot, s));
  CK(cudaStreamSynchronize(s));
  return TGQ_OK;
}

int tgq_ball_stats(tgq_handle * h, const float * c3, const float * qn3, const float * radius, int64_t Q, double * out16)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  if (Q < 0 || (Q > 0 && (!c3 || !radius || !out16))) return fail(h, TGQ_EINVAL, "tgq_ball_stats: bad arguments");
  if (qn3 && !h->has_normals) return fail(h, TGQ_ESTATE, "tgq_ball_stats: query normals given but the structure was built without normals");
  Guard g(h);
  Arg args[4] = {{c3, nullptr, 12}, {qn3, nullptr, 12}, {radius, nullptr, 4}, {nullptr, out16, 128}};
  BvhView v = view_of(h);
  size_t smem = trav_smem_bytes(StatsW::STAGE);
  return run_batched(h, Q, args, 4, true, [&](ChunkCtx & c) -> int {
    int grid = 0; int * cnt = c.slot->counter.as<int>();
    CKRC(persistent_grid(h, k_ball_stats, smem, c.n, cnt, c.stream, &grid));
    k_ball_stats<<<grid, TGQ_BLOCK, smem, c.stream>>>(v, (float const *)c.din[0], (float const *)c.din[1], (float const *)c.din[2], c.perm, c.n, cnt,
                                                      (double *)c.dout[3]);
    h->st.last_launches += 1; h->st.total_launches += 1;
    return TGQ_OK;
  });
}

int tgq_ball_histogram(tgq_handle * h, const float * c3, const float * radius, int64_t Q, int32_t num_bins, uint32_t * hist)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  if (Q < 0 || num_bins < 1 || (Q > 0 && (!c3 || !radius || !hist))) return fail(h, TGQ_EINVAL, "tgq_ball_histogram: bad arguments");
  if (num_bins > 256) return fail(h, TGQ_ELIMIT, "tgq_ball_histogram: more than 256 bins");
  Guard g(h);
  Arg args[3] = {{c3, nullptr, 12}, {radius, nullptr, 4}, {nullptr, hist, 4 * (size_t)num_bins}};
  BvhView v = view_of(h);
  size_t smem = trav_smem_bytes(HistW::STAGE) + (size_t)num_bins * TGQ_BLOCK * 4;
  return run_batched(h, Q, args, 3, true, [&](ChunkCtx & c) -> int {
    int grid = 0; int * cnt = c.slot->counter.as<int>();
    CKRC(persistent_grid(h, k_ball_hist, smem, c.n, cnt, c.stream, &grid));
    k_ball_hist<<<grid, TGQ_BLOCK, smem, c.stream>>>(v, (float const *)c.din[0], (float const *)c.din[1], c.perm, c.n, num_bins, cnt, (unsigned *)c.dout[2]);
    h->st.last_launches += 1; h->st.total_launches += 1;
    return TGQ_OK;
  });
}

int tgq_spin_image(tgq_handle * h, const float * q3, const float * qn3, int64_t Q, int32_t num_radial_bins, int32_t num_height_bins,
                   float max_extent, uint32_t * image)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  int64_t const bins = (int64_t)num_radial_bins * num_height_bins;
  if (Q < 0 || num_radial_bins < 1 || num_height_bins < 1 || !(max_extent > 0.0f) || (Q > 0 && (!q3 || !qn3 || !image)))
    return fail(h, TGQ_EINVAL, "tgq_spin_image: bad arguments");
  if (bins > 1024) return fail(h, TGQ_ELIMIT, "tgq_spin_image: more than 1024 bins");
  if (Q == 0) return TGQ_OK;
  Guard g(h);
  Arg args[3] = {{q3, nullptr, 12}, {qn3, nullptr, 12}, {nullptr, image, 4 * (size_t)bins}};
  f4 const * pts = h->elems.as<f4>(); int const P = h->n;
  size_t smem = (size_t)TGQ_SPIN_TILE * 16 + (size_t)bins * TGQ_SPIN_WARPS * 4;
  {
    std::lock_guard<std::mutex> lock(g_smem_mutex);
    size_t & cur = g_smem_limit[std::make_pair(h->device, (void const *)k_spin_image)];
    if (smem > cur) { CK(cudaFuncSetAttribute(k_spin_image, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
  }
  int64_t saved_chunk = h->opt_chunk;
  h->opt_chunk = std::max<int64_t>(1024, std::min<int64_t>(saved_chunk, ((int64_t)1 << 26) / bins));   // staged image bytes per chunk stay bounded
  int rc = run_batched(h, Q, args, 3, false, [&](ChunkCtx & c) -> int {
    int gx = (int)((c.n + TGQ_SPIN_WARPS - 1) / TGQ_SPIN_WARPS);
    // enough sample slices for about eight blocks per SM, each at least one tile long
    int want = std::max(1, (8 * h->sm_count + gx - 1) / gx);
    int slices = (int)std::max<int64_t>(1, std::min<int64_t>(want, ((int64_t)P + TGQ_SPIN_TILE - 1) / TGQ_SPIN_TILE));
    int slice_len = (int)(((int64_t)P + slices - 1) / slices);
    slice_len = (slice_len + TGQ_SPIN_TILE - 1) / TGQ_SPIN_TILE * TGQ_SPIN_TILE;
    slices = std::max(1, (int)(((int64_t)P + slice_len - 1) / slice_len));
    CK(cudaMemsetAsync(c.dout[2], 0, 4 * (size_t)bins * (size_t)c.n, c.stream));
    k_spin_image<<<dim3(gx, slices), 32 * TGQ_SPIN_WARPS, smem, c.stream>>>(pts, P, (float const *)c.din[0], (float const *)c.din[1], c.n, num_radial_bins, num_height_bins,
                                                          max_extent, max_extent, slice_len, (unsigned *)c.dout[2]);
    h->st.last_launches += 1; h->st.total_launches += 1;
    return TGQ_OK;
  });
  h->opt_chunk = saved_chunk;
  return rc;
}

int tgq_visibility(tgq_handle * h, const float * p3, int64_t S, const float * dirs3, int32_t num_dirs, int32_t per_sample_dirs, float scale,
                   double * visibility, int32_t * num_escaped)
{
  CKRC(need_built(h, 1 << TGQ_KIND_TRIANGLES));
  if (S < 0 || num_dirs < 1 || num_dirs > 4096 || (S > 0 && (!p3 || !dirs3))) return fail(h, TGQ_EINVAL, "tgq_visibility: bad arguments");
  if (S == 0) return TGQ_OK;
  Guard g(h);
  BvhView v = view_of(h);
  int D = num_dirs;
  void const * ddirs = nullptr;
  if (!per_sample_dirs)
  {
    CKRC(to_device(h, dirs3, 12 * (size_t)D, h->misc, h->stream, &ddirs));
    CK(cudaStreamSynchronize(h->stream));
  }
  size_t smem = trav_smem_bytes(VisW::STAGE);
  Arg args[4] = {{p3, nullptr, 12}, {per_sample_dirs ? dirs3 : nullptr, nullptr, 12 * (size_t)D}, {nullptr, visibility, 8}, {nullptr, num_escaped, 4}};
  int64_t saved_chunk = h->opt_chunk;
  h->opt_chunk = std::max<int64_t>(1024, std::min<int64_t>(saved_chunk, ((int64_t)1 << 24) / D));   // rays per chunk stay below 2^24 + slack
  int rc = run_batched(h, S, args, 4, false, [&](ChunkCtx & c) -> int {
    if ((int64_t)c.n * D > (int64_t)1 << 30) return fail(h, TGQ_ELIMIT, "tgq_visibility: more than 2^30 rays in one device-resident call; split the batch");
    Slot & sl = *c.slot; size_t nr = (size_t)D * (size_t)c.n;
    CK(sl.keys[1].reserve(nr));
    uint8_t * esc = sl.keys[1].as<uint8_t>();
    int grid = 0; int * cnt = sl.counter.as<int>();
    CKRC(persistent_grid(h, k_vis_rays, smem, (int64_t)nr, cnt, c.stream, &grid, TGQ_RAY_BLOCKS));
    k_vis_rays<<<grid, TGQ_BLOCK, smem, c.stream>>>(v, (float const *)c.din[0], per_sample_dirs ? (float const *)c.din[1] : (float const *)ddirs, D,
                                                    per_sample_dirs ? 1 : 0, (int)nr, scale, cnt, esc);
    int blocks = (int)(((int64_t)c.n * 32 + TGQ_BLOCK - 1) / TGQ_BLOCK);
    k_vis_count<<<blocks, TGQ_BLOCK, 0, c.stream>>>(esc, c.n, D, (double *)c.dout[2], (int *)c.dout[3]);
    h->st.last_launches += 2; h->st.total_launches += 2;
    return TGQ_OK;
  });
  h->opt_chunk = saved_chunk;
  return rc;
}

int tgq_avg_separation(tgq_handle * h, const int32_t * seeds, int32_t n, float * out)
{
  CKRC(need_built(h, 1 << TGQ_KIND_POINTS));
  if (n < 1 || !seeds || !out) return fail(h, TGQ_EINVAL, "tgq_avg_separation: bad arguments");
  if (h->n < 2) return fail(h, TGQ_ESTATE, "tgq_avg_separation: needs at least two points");
  Guard g(h);
  cudaStream_t s = h->stream;
  std::vector<int32_t> hs((size_t)n);
  if (is_device_ptr(seeds)) { CK(cudaMemcpyAsync(hs.data(), seeds, 4 * (size_t)n, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
  else memcpy(hs.data(), seeds, 4 * (size_t)n);
  for (int i = 0; i < n; ++i) if (hs[(size_t)i] < 0 || hs[(size_t)i] >= h->n) return fail(h, TGQ_EINVAL, "tgq_avg_separation: seed index out of range");
  CK(h->misc.reserve(8 * (size_t)n));
  int * ds = h->misc.as<int>(); float * dd = (float *)(ds + n);
  CK(cudaMemcpyAsync(ds, hs.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, s));
  k_self_nn<<<grid_for(n), TGQ_BLOCK, 0, s>>>(view_of(h), h->pos_of_elem.as<int>(), ds, n, dd);
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  std::vector<float> d2((size_t)n);
  CK(cudaMemcpyAsync(d2.data(), dd, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  float acc = 0;                                    // PointSet3.hpp:363-376: float accumulation in trial order
  for (int i = 0; i < n; ++i) acc += d2[(size_t)i];
  *out = std::sqrt(acc / (float)n);
  return TGQ_OK;
}

static int ray_common(tgq_handle * h, const float * o3, const float * d3, const float * tmax, int64_t R, float * t, int32_t * idx,
                      float * normal3, uint8_t * hit, bool any)
{
  CKRC(need_built(h, 1 << TGQ_KIND_TRIANGLES));
  if (R < 0 || (R > 0 && (!o3 || !d3))) return fail(h, TGQ_EINVAL, "ray query: bad arguments");
  Guard g(h);
  BvhView v = view_of(h);
  size_t smem = trav_smem_bytes(2);
  Arg args[7] = {{o3, nullptr, 12}, {d3, nullptr, 12}, {tmax, nullptr, 4}, {nullptr, t, 4}, {nullptr, idx, 4}, {nullptr, normal3, 12}, {nullptr, hit, 1}};
  h->sorting_rays = h->opt_ray_order >= 3;
  struct Reset { tgq_handle * h; ~Reset() { h->sorting_rays = false; } } reset = {h};
  return run_batched(h, R, args, 7, true, [&](ChunkCtx & c) -> int {
    int grid = 0; int * cnt = c.slot->counter.as<int>();
    if (any)
    {
      CKRC(persistent_grid(h, k_ray<true>, smem, c.n, cnt, c.stream, &grid, TGQ_RAY_BLOCKS));
      k_ray<true><<<grid, TGQ_BLOCK, smem, c.stream>>>(v, (float const *)c.din[0], (float const *)c.din[1], (float const *)c.din[2], c.perm, c.n, cnt,
                                                    (float *)c.dout[3], (int *)c.dout[4], (float *)c.dout[5], (uint8_t *)c.dout[6], h->peers);
    }
    else
    {
      CKRC(persistent_grid(h, k_ray<false>, smem, c.n, cnt, c.stream, &grid, TGQ_RAY_BLOCKS));
      k_ray<false><<<grid, TGQ_BLOCK, smem, c.stream>>>(v, (float const *)c.din[0], (float const *)c.din[1], (float const *)c.din[2], c.perm, c.n, cnt,
                                                     (float *)c.dout[3], (int *)c.dout[4], (float *)c.dout[5], (uint8_t *)c.dout[6], h->peers);
    }
    h->st.last_launches += 1; h->st.total_launches += 1;
    return TGQ_OK;
  });
}

// Common argument checks and handle state of the *_records entry points.
static int records_begin(tgq_handle * h, void const * in0, int32_t shard_rank, int32_t shard_world, void * const * dest, int32_t n_dest, int32_t compact)
{
  if (!h) return TGQ_EINVAL;
  if (compact && n_dest != 1) return fail(h, TGQ_EINVAL, "record query: the compact form takes one destination");
  if (shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world) return fail(h, TGQ_EINVAL, "record query: bad shard rank / world");
  if (n_dest < 1 || n_dest > 8 || !dest) return fail(h, TGQ_EINVAL, "record query: 1..8 destination arrays");
  if (in0 && !is_device_ptr(in0)) return fail(h, TGQ_EINVAL, "record query: inputs must be device pointers");
  for (int j = 0; j < n_dest; ++j) if (!dest[j] || !is_device_ptr(dest[j])) return fail(h, TGQ_EINVAL, "record query: destinations must be device (or peer-mapped) pointers");
  h->peers.n = n_dest; h->peers.compact = compact ? 1 : 0; for (int j = 0; j < 8; ++j) h->peers.base[j] = j < n_dest ? dest[j] : nullptr;
  h->shard_rank = shard_rank; h->shard_world = shard_world; h->last_share = 0;
  return TGQ_OK;
}
static void records_end(tgq_handle * h, int64_t Q, int64_t * share)
{
  if (share) *share = h->shard_world > 1 ? h->last_share : Q;
  h->peers.n = 0; h->peers.compact = 0; h->shard_rank = 0; h->shard_world = 1;
}

int tgq_closest_records(tgq_handle * h, const float * q3, int64_t Q, double dist_bound, int32_t shard_rank, int32_t shard_world, void * const * dest,
                        int32_t n_dest, int32_t compact, int64_t * share)
{
  CKRC(records_begin(h, q3, shard_rank, shard_world, dest, n_dest, compact));
  int rc = (h->kind == TGQ_KIND_TRIANGLES && h->opt_pool && !h->opt_wide && h->leaf_built <= 4)
             ? closest_common(h, q3, nullptr, nullptr, 0.0f, Q, dist_bound, 0, nullptr, nullptr, nullptr)
             : fail(h, TGQ_ESTATE, "tgq_closest_records needs a triangle structure with the default (pooled) closest-point kernel");
  records_end(h, Q, share);
  return rc;
}

int tgq_ray_first_records(tgq_handle * h, const float * o3, const float * d3, const float * tmax, int64_t R, int32_t shard_rank, int32_t shard_world,
                          void * const * dest, int32_t n_dest, int32_t compact, int64_t * share)
{
  CKRC(records_begin(h, o3, shard_rank, shard_world, dest, n_dest, compact));
  int rc = (is_device_ptr(d3) && (!tmax || is_device_ptr(tmax))) ? ray_common(h, o3, d3, tmax, R, nullptr, nullptr, nullptr, nullptr, false)
                                                                : fail(h, TGQ_EINVAL, "record query: inputs must be device pointers");
  records_end(h, R, share);
  return rc;
}

int tgq_bucket_records(tgq_handle * h, const void * compact, int64_t n, int32_t kind, int64_t shard_len, int32_t world, void * buckets, int64_t cap,
                       int64_t * counts)
{
  if (!h) return TGQ_EINVAL;
  if (n < 0 || kind < 0 || kind > 1 || shard_len < 1 || world < 1 || world > 8 || cap < 1 || !buckets || !counts || (n > 0 && !compact))
    return fail(h, TGQ_EINVAL, "tgq_bucket_records: bad arguments");
  if ((n > 0 && !is_device_ptr(compact)) || !is_device_ptr(buckets) || !is_device_ptr(counts)) return fail(h, TGQ_EINVAL, "tgq_bucket_records: device pointers only");
  Guard g(h);
  CK(cudaMemsetAsync(counts, 0, 8 * (size_t)world, h->stream));
  if (n > 0)
  {
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    k_bucket_records<<<grid, 256, 0, h->stream>>>(compact, n, kind, shard_len, world, buckets, cap, (unsigned long long *)counts);
    CK(cudaGetLastError());
    h->st.total_launches += 1;
  }
  if (!h->opt_async) CK(cudaStreamSynchronize(h->stream));
  return TGQ_OK;
}

int tgq_scatter_records(tgq_handle * h, const void * compact, int32_t slots, int64_t cap, const int64_t * counts, int32_t kind, void * dest)
{
  if (!h) return TGQ_EINVAL;
  if (slots < 1 || cap < 0 || kind < 0 || kind > 1 || !compact || !dest || !counts) return fail(h, TGQ_EINVAL, "tgq_scatter_records: bad arguments");
  if (!is_device_ptr(compact) || !is_device_ptr(dest) || !is_device_ptr(counts)) return fail(h, TGQ_EINVAL, "tgq_scatter_records: device pointers only");
  if (cap == 0) return TGQ_OK;
  Guard g(h);
  int grid = (int)std::min<int64_t>((cap + 255) / 256, (int64_t)h->sm_count * 32);
  k_scatter_records<<<grid, 256, 0, h->stream>>>(compact, slots, cap, (long long const *)counts, kind, dest);
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  if (!h->opt_async) CK(cudaStreamSynchronize(h->stream));
  return TGQ_OK;
}

int tgq_ray_first(tgq_handle * h, const float * o3, const float * d3, const float * tmax, int64_t R, float * t, int32_t * idx, float * normal3)
{ return ray_common(h, o3, d3, tmax, R, t, idx, normal3, nullptr, false); }

int tgq_ray_any(tgq_handle * h, const float * o3, const float * d3, const float * tmax, int64_t R, uint8_t * hit)
{ return ray_common(h, o3, d3, tmax, R, nullptr, nullptr, nullptr, hit, true); }

int tgq_sdf(tgq_handle * h, const float * p3, const float * n3, int64_t S, float scale, int32_t only_hit_interior, double * sdf, float * ray_t,
            int32_t * ray_idx)
{
  CKRC(need_built(h, 1 << TGQ_KIND_TRIANGLES));
  if (S < 0 || (S > 0 && (!p3 || !n3))) return fail(h, TGQ_EINVAL, "tgq_sdf: bad arguments");
  Guard g(h);
  BvhView v = view_of(h);
  // per-sample arrays; the 30 rays of a sample stay together so the epilogue is one thread per sample
  size_t smem = trav_smem_bytes(SdfW::STAGE);
  Arg args[5] = {{p3, nullptr, 12}, {n3, nullptr, 12}, {nullptr, sdf, 8}, {nullptr, ray_t, 120}, {nullptr, ray_idx, 120}};
  return run_batched(h, S, args, 5, false, [&](ChunkCtx & c) -> int {
    Slot & sl = *c.slot; size_t nr = 30 * (size_t)c.n;
    CK(sl.keys[0].reserve(4 * nr)); CK(sl.keys[1].reserve(nr));
    float * dt = c.dout[3] ? (float *)c.dout[3] : sl.keys[0].as<float>();
    uint8_t * dv = sl.keys[1].as<uint8_t>();
    int grid = 0; int * cnt = sl.counter.as<int>();
    CKRC(persistent_grid(h, k_sdf_rays, smem, (int64_t)nr, cnt, c.stream, &grid, TGQ_RAY_BLOCKS));
    k_sdf_rays<<<grid, TGQ_BLOCK, smem, c.stream>>>(v, (float const *)c.din[0], (float const *)c.din[1], (int)nr, scale, only_hit_interior, cnt, dt,
                                                 (int *)c.dout[4], dv);
    if (c.dout[2]) k_sdf_stats<<<grid_for(c.n), TGQ_BLOCK, 0, c.stream>>>(dt, dv, c.n, scale, (double *)c.dout[2]);
    h->st.last_launches += c.dout[2] ? 2 : 1; h->st.total_launches += c.dout[2] ? 2 : 1;
    return TGQ_OK;
  });
}

int tgq_icp_sums(tgq_handle * h, const float * from3, const float * to3, int64_t n, const float * tr12, double * out16)
{
  if (!h) return TGQ_EINVAL;
  if (n < 1 || !from3 || !to3 || !out16) return fail(h, TGQ_EINVAL, "tgq_icp_sums: bad arguments");
  Guard g(h);
  cudaStream_t s = h->stream;
  void const * df = nullptr; void const * dt = nullptr;
  CKRC(to_device(h, from3, 12 * (size_t)n, h->slot[0].in[0], s, &df));
  CKRC(to_device(h, to3, 12 * (size_t)n, h->slot[0].in[1], s, &dt));
  CK(h->misc.reserve(16 * 8 + 12 * 4 + 64));
  long long * acc = h->misc.as<long long>(); float * dtr = (float *)(acc + 16); int * dexp = (int *)(dtr + 12);
  CK(cudaMemsetAsync(acc, 0, 16 * 8, s));
  if (tr12) CK(cudaMemcpyAsync(dtr, tr12, 48, is_device_ptr(tr12) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  if (h->opt_icp_ref)
  {
    // the reference's own accumulation order (Icp3.hpp:306-367,418-463): one warp, one lane per chain
    float * dout = (float *)acc; float hf[16];
    k_icp_ref_order<<<1, TGQ_ICP_THREADS, 0, s>>>((float const *)df, (float const *)dt, n, tr12 ? dtr : nullptr, dout);
    CK(cudaGetLastError());
    h->st.total_launches += 1;
    CK(cudaMemcpyAsync(hf, dout, 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    double res[16]; for (int i = 0; i < 16; ++i) res[i] = (double)hf[i];
    if (is_device_ptr(out16)) { CK(cudaMemcpyAsync(out16, res, 128, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s)); }
    else memcpy(out16, res, 128);
    return TGQ_OK;
  }
  // fixed-point units from the data range: 2^-34 of the largest coordinate magnitude (2^-34 of its square for products)
  CK(cudaMemsetAsync(dexp, 0, 4, s));
  k_max_abs<<<148 * 4, 256, 0, s>>>((float const *)df, (float const *)dt, 3 * n, dexp);
  int mbits = 0;
  CK(cudaMemcpyAsync(&mbits, dexp, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  float maxabs; memcpy(&maxabs, &mbits, 4);
  int e = 0; std::frexp(maxabs > 0 ? maxabs : 1.0f, &e);          // maxabs < 2^e
  double unit_pos = std::ldexp(1.0, e - 34), unit_prod = std::ldexp(1.0, 2 * e + 2 - 34);
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  k_icp_sums<<<grid, 256, 0, s>>>((float const *)df, (float const *)dt, n, tr12 ? dtr : nullptr, 1.0 / unit_pos, 1.0 / unit_prod, acc);
  CK(cudaGetLastError());
  h->st.total_launches += 2;
  long long ha[16];
  CK(cudaMemcpyAsync(ha, acc, 16 * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  double N = (double)n, sp[3], sq[3], spq[9];
  for (int i = 0; i < 3; ++i) { sp[i] = (double)ha[i] * unit_pos; sq[i] = (double)ha[3 + i] * unit_pos; }
  for (int i = 0; i < 9; ++i) spq[i] = (double)ha[6 + i] * unit_prod;
  double res[16];
  for (int i = 0; i < 3; ++i) { res[i] = sp[i] / N; res[3 + i] = sq[i] / N; }
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) res[6 + 3 * r + c] = spq[3 * r + c] - N * res[r] * res[3 + c];
  // sum |L p + t - q|^2 = sum |(L(p-pm) - (q-qm)) + (L pm + t - qm)|^2 = centred sum + N |L pm + t - qm|^2
  double L[12]; for (int i = 0; i < 12; ++i) L[i] = tr12 ? 0.0 : ((i % 5 == 0) ? 1.0 : 0.0);
  if (tr12)
  {
    float ht[12];
    CK(cudaMemcpyAsync(ht, dtr, 48, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s));
    for (int i = 0; i < 12; ++i) L[i] = ht[i];
  }
  double rbar2 = 0;
  for (int r = 0; r < 3; ++r) { double vv = L[4 * r] * res[0] + L[4 * r + 1] * res[1] + L[4 * r + 2] * res[2] + L[4 * r + 3] - res[3 + r]; rbar2 += vv * vv; }
  double total = (double)ha[15] * unit_prod;
  res[15] = rbar2 + (total - N * rbar2);     // Icp3.hpp:441-461: |tr pm - qm|^2 + sum |L dp - dq|^2
  if (is_device_ptr(out16)) { CK(cudaMemcpyAsync(out16, res, 128, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s)); }
  else memcpy(out16, res, 128);
  return TGQ_OK;
}

int tgq_transform_points(tgq_handle * h, const float * in3, int64_t n, const float * tr12, float * out3)
{
  if (!h) return TGQ_EINVAL;
  if (n < 0 || (n > 0 && (!in3 || !out3)) || !tr12) return fail(h, TGQ_EINVAL, "tgq_transform_points: bad arguments");
  if (n == 0) return TGQ_OK;
  Guard g(h);
  cudaStream_t s = h->stream;
  void const * di = nullptr;
  CKRC(to_device(h, in3, 12 * (size_t)n, h->slot[0].in[0], s, &di));
  CK(h->misc.reserve(64));
  CK(cudaMemcpyAsync(h->misc.p, tr12, 48, is_device_ptr(tr12) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  float * dout = out3;
  if (!is_device_ptr(out3)) { CK(h->slot[0].out[0].reserve(12 * (size_t)n)); dout = h->slot[0].out[0].as<float>(); }
  k_transform<<<grid_for(n), TGQ_BLOCK, 0, s>>>((float const *)di, n, h->misc.as<float>(), dout);
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  if (dout != out3) CKRC(from_device(h, out3, dout, 12 * (size_t)n, s));
  CK(cudaStreamSynchronize(s));
  return TGQ_OK;
}

// R = U V^T of the 3x3 matrix M (row-major) by one-sided Jacobi rotations in double: the rotation Eigen::JacobiSVD gives
// Icp3::alignOneStep (Algorithms/Icp3.hpp:369-370).
static void polar_rotation(double const M[9], double R[9])
{
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) A[i] = M[i];
  for (int sweep = 0; sweep < 64; ++sweep)
  {
    double worst = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q)
      {
        double app = 0, aqq = 0, apq = 0;
        for (int r = 0; r < 3; ++r) { app += A[3 * r + p] * A[3 * r + p]; aqq += A[3 * r + q] * A[3 * r + q]; apq += A[3 * r + p] * A[3 * r + q]; }
        double denom = std::sqrt(app * aqq);
        if (denom > 0) worst = std::max(worst, std::fabs(apq) / denom);
        if (apq == 0) continue;
        double zeta = (aqq - app) / (2.0 * apq);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        for (int r = 0; r < 3; ++r)
        {
          double x = A[3 * r + p], y = A[3 * r + q]; A[3 * r + p] = c * x - sn * y; A[3 * r + q] = sn * x + c * y;
          x = V[3 * r + p]; y = V[3 * r + q]; V[3 * r + p] = c * x - sn * y; V[3 * r + q] = sn * x + c * y;
        }
      }
    if (worst < 1e-15) break;
  }
  double U[9], sv[3];
  for (int c = 0; c < 3; ++c)
  {
    sv[c] = std::sqrt(A[c] * A[c] + A[3 + c] * A[3 + c] + A[6 + c] * A[6 + c]);
    for (int r = 0; r < 3; ++r) U[3 * r + c] = sv[c] > 0 ? A[3 * r + c] / sv[c] : 0.0;
  }
  int mn = 0; for (int c = 1; c < 3; ++c) if (sv[c] < sv[mn]) mn = c;
  if (sv[mn] <= 1e-14 * std::max(sv[0], std::max(sv[1], sv[2])))
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

// a * b for row-major 3x4 [L | t] in fp32 with Eigen's t0 + (t1 + t2) order (AffineTransformN.hpp:154-157).
static void compose12(float const a[12], float const b[12], float r[12])
{
  for (int i = 0; i < 3; ++i)
  {
    for (int j = 0; j < 3; ++j) r[4 * i + j] = a[4 * i] * b[j] + (a[4 * i + 1] * b[4 + j] + a[4 * i + 2] * b[8 + j]);
    r[4 * i + 3] = (a[4 * i] * b[3] + (a[4 * i + 1] * b[7] + a[4 * i + 2] * b[11])) + a[4 * i + 3];
  }
}

int tgq_icp_align(tgq_handle * h, const float * from3, int64_t n, int32_t min_iterations, int32_t max_iterations, double fractional_error_threshold,
                  int32_t want_error, float * tr12, float * error, int32_t * batches, int32_t * correspondences)
{
  CKRC(need_built(h, 1 << TGQ_KIND_TRIANGLES));
  if (n < 0 || (n > 0 && !from3) || !tr12 || max_iterations < 0) return fail(h, TGQ_EINVAL, "tgq_icp_align: bad arguments");
  float ident[12]; for (int i = 0; i < 12; ++i) ident[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  memcpy(tr12, ident, sizeof ident);
  if (error) *error = 0; if (batches) *batches = 0;
  if (n == 0) return TGQ_OK;
  size_t bytes = 12 * (size_t)n;
  cudaStream_t s = h->stream;
  {
    Guard g(h);
    CK(h->icp_cur.reserve(bytes)); CK(h->icp_orig.reserve(bytes)); CK(h->icp_to.reserve(bytes)); CK(h->icp_idx.reserve(4 * (size_t)n));
    CK(cudaMemcpyAsync(h->icp_orig.p, from3, bytes, is_device_ptr(from3) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->icp_cur.p, h->icp_orig.p, bytes, cudaMemcpyDeviceToDevice, s));
    CK(cudaStreamSynchronize(s));
  }
  float * cur = h->icp_cur.as<float>(); float * to = h->icp_to.as<float>(); int32_t * idx = h->icp_idx.as<int32_t>();
  int nb = 0; double sums[16];
  int64_t launches0 = h->st.total_launches; double query_ms = 0;
  // Icp3::measureError (Icp3.hpp:387-464): correspondences of the CURRENT positions, then the error under `tr`
  auto measure = [&](float const * tr, float & err) -> int {
    CKRC(tgq_closest(h, cur, n, -1.0, 0, idx, nullptr, to));
    query_ms += h->st.last_query_ms;
    if (correspondences)
    {
      Guard g(h);
      CK(cudaMemcpyAsync(correspondences + (size_t)nb * (size_t)n, idx, 4 * (size_t)n, is_device_ptr(correspondences) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
    }
    ++nb;
    CKRC(tgq_icp_sums(h, cur, to, n, tr, sums));
    err = (float)sums[15];
    return TGQ_OK;
  };
  float old_tr[12], new_tr[12]; memcpy(old_tr, ident, sizeof ident); memcpy(new_tr, ident, sizeof ident);
  float old_err = 0, new_err = 0;
  CKRC(measure(old_tr, old_err));
  new_err = old_err;
  bool done = old_err <= FLT_MIN;
  for (int i = 0; i < max_iterations && !done; ++i)
  {
    // alignOneStep(from_points, to_points) (Icp3.hpp:228-229,298-384): the CURRENT source positions against the targets found by
    // the last measureError, which ran before the positions were updated (:237,260-261) -- reproduced as is
    CKRC(tgq_icp_sums(h, cur, to, n, nullptr, sums));
    double R[9]; polar_rotation(&sums[6], R);
    float inc[12];
    for (int r = 0; r < 3; ++r)
    {
      for (int c = 0; c < 3; ++c) inc[4 * r + c] = (float)R[3 * r + c];
      float lp = inc[4 * r] * (float)sums[0] + (inc[4 * r + 1] * (float)sums[1] + inc[4 * r + 2] * (float)sums[2]);
      inc[4 * r + 3] = (float)sums[3 + r] - lp;
    }
    compose12(inc, old_tr, new_tr);
    if (i < max_iterations - 1 || want_error) CKRC(measure(inc, new_err));
    if (i < max_iterations - 1)
    {
      float frac = (old_err - new_err) / old_err;
      if (i >= min_iterations && (double)frac < fractional_error_threshold)
      {
        if (frac > 0) { memcpy(old_tr, new_tr, sizeof new_tr); old_err = new_err; }      // improved slightly: return the new one
        done = true; break;                                                                // else the previous alignment was better
      }
      CKRC(tgq_transform_points(h, h->icp_orig.as<float>(), n, new_tr, cur));             // transform the ORIGINAL points (no drift)
    }
    old_err = new_err; memcpy(old_tr, new_tr, sizeof new_tr);
  }
  memcpy(tr12, old_tr, sizeof old_tr);
  if (error) *error = old_err;
  if (batches) *batches = nb;
  h->st.last_query_ms = query_ms; h->st.last_sort_ms = 0; h->st.last_launches = h->st.total_launches - launches0;
  return TGQ_OK;
}

int tgq_area_sample(tgq_handle * h, int64_t S, uint64_t seed, float * p3, float * n3, int32_t * tri)
{
  CKRC(need_built(h, 1 << TGQ_KIND_TRIANGLES));
  if (S < 0) return fail(h, TGQ_EINVAL, "tgq_area_sample: bad arguments");
  if (S == 0 || h->n == 0) return TGQ_OK;
  Guard g(h);
  cudaStream_t s = h->stream; size_t nn = (size_t)h->n;
  if (!h->cum_valid)
  {
    CK(h->cum.reserve(8 * nn));
    k_seq_scan<<<1, 32, 0, s>>>(h->area_by_elem.as<float>(), h->n, h->cum.as<double>());
    CK(cudaGetLastError());
    h->st.total_launches += 1; h->cum_valid = true;
  }
  size_t ns = (size_t)S;
  CK(h->slot[0].out[0].reserve(12 * ns)); CK(h->slot[0].out[1].reserve(12 * ns)); CK(h->slot[0].out[2].reserve(4 * ns));
  float * dp = p3 ? (is_device_ptr(p3) ? p3 : h->slot[0].out[0].as<float>()) : nullptr;
  float * dn = n3 ? (is_device_ptr(n3) ? n3 : h->slot[0].out[1].as<float>()) : nullptr;
  int * dt = tri ? (is_device_ptr(tri) ? tri : h->slot[0].out[2].as<int>()) : nullptr;
  k_area_sample<<<grid_for(S), TGQ_BLOCK, 0, s>>>(h->elems.as<TriRec>(), h->pos_of_elem.as<int>(), h->cum.as<double>(), h->n, S, seed, dp, dn, dt);
  CK(cudaGetLastError());
  h->st.total_launches += 1;
  if (p3 && dp != p3) CKRC(from_device(h, p3, dp, 12 * ns, s));
  if (n3 && dn != n3) CKRC(from_device(h, n3, dn, 12 * ns, s));
  if (tri && dt != tri) CKRC(from_device(h, tri, dt, 4 * ns, s));
  CK(cudaStreamSynchronize(s));
  return TGQ_OK;
}

int tgq_get_stats(tgq_handle * h, tgq_stats * out)
{
  if (!h || !out) return TGQ_EINVAL;
  *out = h->st;
  return TGQ_OK;
}

}  // extern "C"
