/* thea_gpu_query.h -- C ABI of the B200 batched spatial-query engine (libtgq.so).
 *
 * Drop-in boundary for the query path of Thea's Algorithms::BvhN / MeshBvh.  The reference API is a
 * compile-time C++ concept that answers ONE query per call (Algorithms/ProximityQueryStructureN.hpp:29-35,
 * Algorithms/RayQueryStructureN.hpp:55-60); each entry point below answers a whole batch of the query it cites, and the
 * header-only adaptor include/thea_gpu/GpuBvh.hpp re-exposes them under the reference's member-function names.
 * File:line citations are relative to /root/reference/Code/Source/.
 *
 * Conventions
 *   - Every function returns 0 on success, a TGQ_E* code otherwise; tgq_last_error() gives the message.  Nothing throws.
 *   - All array arguments may be HOST or DEVICE pointers (detected with cudaPointerGetAttributes); device pointers must
 *     belong to the handle's device.  Host arrays are staged through pinned buffers in chunks, copies overlapped with
 *     the kernels.  The caller owns every buffer; the library owns its device copies (BvhN copies its elements too,
 *     Algorithms/BvhN.hpp:722-724).
 *   - Element indices are int32 positions in the caller's build array, -1 = not found (the reference returns intx -1,
 *     BvhN.hpp:1086-1087); "not found" is never an error.  At most 2^27 elements per structure.
 *   - Results are the canonical answer: optimum over all elements of the reference's per-element fp32 arithmetic, ties
 *     broken by the lowest element index (BASELINE.json north_star).  There is no CPU fallback: without a CUDA device
 *     tgq_create fails.
 *   - A handle is externally synchronised: one call at a time per handle.  Work is issued on the handle's stream
 *     (tgq_set_stream) and every query call returns after its results are complete, unless TGQ_OPT_ASYNC is set and all
 *     arguments are device pointers.
 */
#ifndef THEA_GPU_QUERY_H
#define THEA_GPU_QUERY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tgq_handle tgq_handle;

enum
{
  TGQ_OK = 0,
  TGQ_EINVAL = 1,     /* bad argument (null pointer, negative count, k out of range, wrong structure kind) */
  TGQ_ECUDA = 2,      /* CUDA runtime error, message in tgq_last_error */
  TGQ_ENODEVICE = 3,  /* no usable CUDA device: the engine has no CPU path */
  TGQ_ESTATE = 4,     /* structure not built / refit with a different element count */
  TGQ_ELIMIT = 5      /* more than 2^27 elements, k > TGQ_MAX_K */
};

enum
{
  TGQ_KIND_NONE = 0,
  TGQ_KIND_TRIANGLES = 1,
  TGQ_KIND_POINTS = 2
};

enum
{
  TGQ_OPT_SORT_QUERIES = 1,   /* 0 = never, 1 = always, 2 = auto (default: batches >= 4096 are Morton-ordered first) */
  TGQ_OPT_LEAF_SIZE = 2,      /* max elements per leaf: 0 (default) = by kind, 4 for triangle and 15 for point structures; or 1..15; takes effect at
                                 the next build */
  TGQ_OPT_ASYNC = 3,          /* 1: batched query calls return without waiting: device-pointer calls are ordered on the handle's stream,
                                 host-pointer calls on the internal pipeline streams -- call tgq_synchronize before reading host outputs
                                 (lets consecutive calls overlap: the downloads of one with the uploads of the next).  Default 0 */
  TGQ_OPT_CHUNK = 4,          /* queries per staged chunk for host-pointer calls (default 16777216) */
  TGQ_OPT_POOL = 5,           /* closest point on triangle structures: 1 = leaf tests pooled across the warp (default), 0 = per lane */
  TGQ_OPT_WIDE = 7,           /* closest point on triangle structures: 1 = four lanes per query on 4-wide nodes (set BEFORE the build; measured
                                 1.5x slower than the default: profiles/r02_b_closest_wide_vs_pool_ncu.txt), 0 = binary nodes (default) */
  TGQ_OPT_SHARD_CELL_BITS = 8,   /* *_records with shard_world > 1: the batch is dealt to the ranks in 2^bits Morton cells (default 12) */
  TGQ_OPT_RAY_ORDER = 9,      /* ray batches: 0 (default) = launch order by the Morton code of the origin; b in 3..30 = the b lowest code bits give way to
                                 the direction's sign octant (rays of one cell of 2^(30-b) ... grouped by heading) */
  TGQ_OPT_SET_BUFFER = 10,    /* k-NN and the sample graph: 1 (default) = candidate buffers compacted by the whole warp (k_knn_buf / k_ball_graph_buf),
                                 0 = one sorted set per lane (k_knn / k_ball_graph); same results */
  TGQ_OPT_ICP_REFERENCE_ORDER = 6   /* 1 (default): tgq_icp_sums / tgq_icp_align accumulate in the reference's own order -- sequential fp32, the
                                       loops of Icp3.hpp:306-367,418-463 -- so iterates, per-iteration correspondences and the final transform
                                       equal the reference's bit for bit; 0: order-independent fixed-point (exact) sums, 20x faster per call, but
                                       the reference's fp32 chains carry ~1e-4 of summation noise over 10^5 points that exact sums do not
                                       reproduce (cfg3: final transform differs by 3e-4 after 50 iterations) */
};

#define TGQ_MAX_K 64

/* ---- lifetime --------------------------------------------------------------------------------------------------- */

/* Creates an engine bound to CUDA device `device` (>= 0).  Replaces the BvhN constructor (BvhN.hpp:664-676). */
int tgq_create(int device, tgq_handle ** out);
/* Releases all device memory.  Replaces BvhN::clear + destructor (BvhN.hpp:796-813). */
int tgq_destroy(tgq_handle * h);
/* Message of the last failure on this handle (or of the last failed tgq_create when h is null). */
const char * tgq_last_error(tgq_handle * h);
/* Issue all work on this cudaStream_t (passed as void*); null = the handle's own stream. */
int tgq_set_stream(tgq_handle * h, void * cuda_stream);
int tgq_set_option(tgq_handle * h, int option, int64_t value);

/* Waits for everything the handle has in flight (the stream of device-pointer calls and the pipeline streams of host-pointer
 * calls).  Needed only with TGQ_OPT_ASYNC. */
int tgq_synchronize(tgq_handle * h);

/* ---- build ------------------------------------------------------------------------------------------------------ */

/* Builds the structure over T triangles given as T x 9 floats (v0 v1 v2), element index = position.
 * Replaces MeshBvh::add + init -> BvhN::init -> createKdTree (Algorithms/MeshBvh.hpp:55-111, BvhN.hpp:704-794,
 * 1365-1525) with a Morton-code LBVH (radix sort + Karras hierarchy + bottom-up refit) built on the device, and the
 * per-triangle cache of Triangle3.hpp:143-155 (unit normal, plane offset, primary axis, area). */
int tgq_build_triangles(tgq_handle * h, const float * v9, int64_t T);
/* Same, from V x 3 vertices and T x 3 int32 vertex indices (MeshTriangles::addTriangles ordering,
 * Algorithms/MeshTriangles.hpp:173-339). */
int tgq_build_triangles_indexed(tgq_handle * h, const float * verts3, int64_t V, const int32_t * tri3, int64_t T);
/* Builds the structure over P points (P x 3 floats) with optional unit normals (P x 3, may be null).
 * Replaces BvhN<Sample*,3>::init of PointSet3 (Algorithms/PointSet3.hpp:329-358,709-723). */
int tgq_build_points(tgq_handle * h, const float * p3, int64_t P, const float * n3);
/* New vertex positions for the same T triangles, same order: recomputes the triangle cache and refits all boxes
 * bottom-up without re-sorting.  Replaces MeshBvh::updateElements + BvhN::updateNodeBounds
 * (MeshBvh.hpp:123-134, BvhN.hpp:816-834,1694-1730). */
int tgq_refit_triangles(tgq_handle * h, const float * v9, int64_t T);

/* BvhN::setTransform / clearTransform (BvhN.hpp:903-925): queries then see every element transformed by the
 * affine map tr12 (row-major 3x4 [L | t]) -- the reference evaluates its metrics on Transformer::transform(element, transform)
 * (Algorithms/MetricL2_Transform.hpp), i.e. points L p + t (AffineTransformN.hpp:160 operation order) and triangles rebuilt from
 * their transformed vertices (TriangleN::transform).  The structure is rebuilt over the transformed elements (same element indices;
 * normals and labels of point structures are the elements' own and stay as they are); closest points, distances, hit times and
 * tgq_info bounds are in the transformed space.  tr12 = null clears the transform: the structure is rebuilt from the untransformed
 * elements the library kept, bit-identical to the original build.  A new build or refit forgets the transform.
 * Used by Tools/Register/Register.cpp:644-650 (bounding-box alignment before the first correspondence search). */
int tgq_set_transform(tgq_handle * h, const float * tr12);

/* BvhN::numElements / getBounds (BvhN.hpp:952-995): kind, element count, unpadded bounds (lo[3], hi[3]; may be null). */
int tgq_info(tgq_handle * h, int32_t * kind, int64_t * num_elements, float * lo3, float * hi3);
/* Per-triangle cache in element order (any pointer may be null): unit normals T x 3 (what
 * getElements()[i].getNormal() returns, ShapeDiameter.hpp:152-153), plane offsets, primary axes, areas. */
int tgq_triangle_cache(tgq_handle * h, float * n3, float * d, int32_t * axis, float * area);

/* ---- queries ---------------------------------------------------------------------------------------------------- */

/* Q closest-element queries.  Replaces Q calls of
 *   BvhN::closestElement<MetricL2>(q, dist_bound, UniversalCompatibility(), &dist, &closest_point)   (BvhN.hpp:1049-1063)
 * dist_bound < 0: unbounded; otherwise only elements with squared distance < dist_bound^2 (strict, in double) count
 * (BvhN.hpp:448,462).  exclude_self != 0: query i ignores element i (PointSet3::FilterSelf, PointSet3.hpp:576-583).
 * Outputs (each may be null): idx[Q]; d2[Q] = fp32 squared distance (the reference's monotone approximation; its
 * *dist is sqrt((double)d2), BvhN.hpp:1058), untouched slots get -1; cp3[Q x 3] = closest point on the element. */
int tgq_closest(tgq_handle * h, const float * q3, int64_t Q, double dist_bound, int32_t exclude_self,
                int32_t * idx, float * d2, float * cp3);

/* Q k-nearest-neighbour queries on a point structure.  Replaces BvhN::kClosestPairs<MetricL2> into an empty
 * BoundedSortedArray of capacity k (BvhN.hpp:1137-1146,474-500).  idx[Q x k], d2[Q x k] ascending by (d2, idx),
 * unused slots -1. */
int tgq_knn(tgq_handle * h, const float * q3, int64_t Q, int32_t k, double dist_bound, int32_t exclude_self,
            int32_t * idx, float * d2);

/* ---- filtered nearest neighbours on point structures ----------------------------------------- */

/* Per-element labels of a point structure (P int32 in element order; null clears them).  The reference's elements carry
 * their label as a member (Tools/Register/Register.cpp:60-75); here it is a parallel array. */
int tgq_set_point_labels(tgq_handle * h, const int32_t * labels);
/* closestElement / kClosestPairs with an element filter pushed around every query.  Replaces
 *   NNFilter filter(n_i, label_i); bvh.pushFilter(&filter); bvh.closestElement<MetricL2>(p_i) | kClosestPairs<MetricL2>(p_i, set);
 *   bvh.popFilter();                      (Tools/Register/Register.cpp:92-101,218-223,274-277; BvhN.hpp:1005-1032,1674-1683)
 * Query i admits element e iff (qlabel == null or label[e] == qlabel[i]) and (qn3 == null or qn3[i] . n[e] >= min_dot);
 * Register's NNFilter is min_dot = 0.  Filtered-out elements are skipped before any distance is taken, as in the reference.
 * qn3 needs a structure built with normals, qlabel one with labels (else TGQ_ESTATE). */
int tgq_closest_filtered(tgq_handle * h, const float * q3, const float * qn3, const int32_t * qlabel, int64_t Q, float min_dot,
                         double dist_bound, int32_t * idx, float * d2, float * cp3);
int tgq_knn_filtered(tgq_handle * h, const float * q3, const float * qn3, const int32_t * qlabel, int64_t Q, int32_t k, float min_dot,
                     double dist_bound, int32_t * idx, float * d2);

/* Q ball range queries on a point structure (the primitive under the local surface features).  Replaces
 *   BvhN::rangeQueryIndices<IntersectionTester>(Ball3(c_i, r_i), result) / processRange (BvhN.hpp:1208-1264,2118-2171;
 *   BallN.hpp:77-100): every element e with sq(x_e - c_i) <= r_i * r_i in fp32, as the reference tests it.
 * count[Q] = how many elements are in range; idx[Q x max_out] = the max_out LOWEST element indices in range, ascending, unused
 * slots -1 (the reference reports them in traversal order: compare as sets; the set is complete iff count <= max_out).
 * max_out = 0 only counts (idx may be null).  A negative radius matches nothing.  exclude_self != 0: query i ignores element i. */
int tgq_ball_range(tgq_handle * h, const float * c3, const float * radius, int64_t Q, int32_t max_out, int32_t exclude_self,
                   int32_t * idx, int32_t * count);

/* Q closest-element queries whose query object is a LINE SEGMENT.  Replaces Q calls of
 *   BvhN::closestElement<MetricL2>(LineSegment3(a_i, b_i), dist_bound, ...) / closestPair<MetricL2>(segment, dist_bound, compat, true)
 * (BvhN.hpp:1049-1114; the per-element metric is MetricL2Impl<LineSegmentN, point> -> LineSegmentN::closestPoint, LineSegmentN.hpp:106-145,
 * on point structures and MetricL2Impl<LineSegmentN, TriangleN> -> TriangleN<3>::squaredDistance(LineSegmentN), Triangle3.hpp:495-509,559-577
 * with Internal::closestPtSegmentSegment, LineSegmentN.hpp:262-357, on triangle structures).  This is the query loop of the T-junction
 * repair in Tools/MeshFix/MeshFix.cpp:756-763, which pushes a VertexFilter (the edge's own endpoints and the other vertices of its faces,
 * MeshFix.cpp:678-697) around every call: `exclude` (Q x n_exclude element indices, -1 = unused slot, n_exclude <= 4, null / 0 = none)
 * is that filter as data.  dist_bound as in tgq_closest (the segment's bounding box takes the place of the point in the root rejection).
 * b_is_direction = 0: b3 holds the second endpoints, the direction is taken as b_i - a_i like the LineSegment3(a, b) constructor does
 * (LineSegmentN.hpp:56-58); != 0: b3 holds the stored direction vectors themselves (LineSegmentN::getDirection of an existing object --
 * the segment's far end is point + direction in the reference's arithmetic, which need not be the endpoint it was built from).
 * Outputs (each may be null): idx[Q] (-1: none), d2[Q] = fp32 squared distance (-1 when none), element_point3 / segment_point3
 * [Q x 3] = the closest pair of points on the element and on the segment (NeighborPair::getTargetPoint / getQueryPoint).  Ties go to the
 * lowest element index. */
int tgq_closest_segment(tgq_handle * h, const float * a3, const float * b3, int32_t b_is_direction, int64_t Q, double dist_bound,
                        const int32_t * exclude, int32_t n_exclude, int32_t * idx, float * d2, float * element_point3, float * segment_point3);

/* Q closest-element queries whose query object is a TRIANGLE.  Replaces Q calls of
 *   BvhN::closestPair<MetricL2>(Triangle3(q_i), dist_bound, compatibility, true)            (BvhN.hpp:1081-1114,1866-1893,2036-2089)
 * q9: Q x 9 floats (v0 v1 v2).  Per-element value: element.squaredDistance(query) on triangle structures -- Moeller's triangle-triangle
 * intersection test with the intersection segment, else nine edge pairs and six vertex-plane projections (Triangle3.hpp:314-392,597-1300)
 * --, query.closestPoint(element) on point structures (Algorithms/MetricL2.hpp:219-233, Triangle3.hpp:294-304).  dist_bound as in
 * tgq_closest (the query's bounding box in the root rejection).  Outputs (each may be null): idx[Q], d2[Q] (-1 when none),
 * element_point3 / query_point3 [Q x 3] = NeighborPair::getTargetPoint / getQueryPoint.  Ties go to the lowest element index.
 * Parity exception: the reference's triangle-triangle test snaps plane distances below FLT_EPSILON (in units of the
 * unnormalised normal) to zero but keeps the unsnapped coordinates, and when all three query vertices are snapped it compares against an
 * interval it never computed (Triangle3.hpp:1024-1077,1209-1216): it can call a pair intersecting that is farther apart than the
 * tolerance FLT_EPSILON / |N| allows.  Where the reference's exhaustive minimum is such a pair, this call returns the minimum over the
 * other pairs (rare unless triangles are small in absolute units).  Zero-area triangles (elements or
 * queries) make the reference's test a 2-D overlap test in a coordinate plane and are outside the guarantee as well. */
int tgq_closest_triangle(tgq_handle * h, const float * q9, int64_t Q, double dist_bound, int32_t * idx, float * d2, float * element_point3,
                         float * query_point3);

/* Closest pair of elements between two structures.  Replaces BvhN::closestPair<MetricL2>(query_bvh, dist_bound, compatibility,
 * get_closest_points) (BvhN.hpp:1081-1114 with a structure as the query; processNeighborsBvh :1900-2009; Test/TestBvh.cpp:251-258):
 * the minimum over all (target element, query element) pairs of the reference's per-pair squared distance, accepted when
 * (double)d2 < dist_bound^2 (dist_bound < 0: unbounded; a structure query gets no bound widening, BvhN.hpp:2384-2416), ties to the
 * lowest target index, then the lowest query index.  `h` is the target, `query` the other structure on the same device; either may hold
 * points or triangles (Test/TestBvh.cpp:251-258 points, :380-385 meshes).  Every pair is evaluated as the reference's leaf code does
 * without role swapping, MetricL2::closestPoints(target element, query element) (BvhN.hpp:2063): for two triangles that is
 * target.squaredDistance(query) (Triangle3.hpp:314-392), whose rounding is not symmetric in its arguments -- the reference's dual-tree
 * traversal swaps the roles of some pairs depending on the shapes of its two trees (:1960-1985), so between two MESHES its result can
 * differ from this one in the last bit; with a point structure on either side the value is symmetric and equal.  Outputs are host memory:
 * *target_index / *query_index (-1: no pair), *d2, the closest points on the query and the target element (each may be null). */
int tgq_closest_pair(tgq_handle * h, tgq_handle * query, double dist_bound, int64_t * target_index, int64_t * query_index, float * d2,
                     float * query_point3, float * target_point3);

/* BvhN::kClosestPairs with another structure as the query (Algorithms/BvhN.hpp:1137-1146 with QueryT = BvhN; the pair traversal of
 * :1900-2009): the k closest (target element, query element) pairs between this POINT structure and the point structure `query`, in
 * the canonical order (squared distance, target index, query index); dist_bound < 0: unbounded, else only pairs strictly closer.
 * target_index / query_index / d2 (may be null): k entries on the host, -1 / -1.0f padded; *n_found = pairs found (<= k).
 * Every query element's k nearest targets are found in one batch, then one device sort picks the k smallest. */
int tgq_k_closest_pairs(tgq_handle * h, tgq_handle * query, int32_t k, double dist_bound, int64_t * target_index, int64_t * query_index, float * d2,
                        int32_t * n_found);

/* Q unbounded range queries on a point structure.  Replaces BvhN::rangeQueryIndices / rangeQuery / processRange
 * <IntersectionTester>(range, ...) (BvhN.hpp:1208-1264,2118-2171) for the two range types the reference's callers and tests use:
 * kind 0: Ball3(a3[i], b[i])  -- every element with sq(x - c) <= r * r in fp32 (BallN.hpp:83);   b = Q radii
 * kind 1: AxisAlignedBox3(a3[i], b[3 i ..])  -- every element with lo <= x <= hi (AxisAlignedBoxN.hpp:401-412);   b = Q x 3 high corners
 * Output in CSR form: offsets[Q + 1] (int64) and idx[total] (int32), the elements of range i at idx[offsets[i] .. offsets[i + 1]) in
 * ASCENDING element index (the reference reports its tree's traversal order; the set is the same).  *total always receives the
 * number of (range, element) pairs; when it exceeds `capacity` (the size of idx, 0 allowed to ask for the counts only) the call
 * fills offsets only and, for capacity > 0, returns TGQ_ELIMIT. */
int tgq_range_query(tgq_handle * h, int32_t kind, const float * a3, const float * b, int64_t Q, int64_t * offsets, int32_t * idx,
                    int64_t capacity, int64_t * total);

/* Q ball neighbourhood statistics on a point structure (the local surface features are processRange calls
 * with an accumulating functor).  Replaces, for every query ball (c_i, r_i),
 *   surf->getBvh().processRange<IntersectionTester>(Ball3(c_i, r_i), functor)                    (BvhN.hpp:1238-1264,2118-2171)
 * with the functors of AverageDistance (EuclideanCallback, SurfaceFeatures/Local/AverageDistance.cpp:39-81), Curvature
 * (ProjectedCurvatureFunctor, Curvature.cpp:37-98) and LocalPca (LocalPcaFunctor, LocalPca.hpp:58-75, LocalPca.cpp:32-62,
 * PcaN.hpp:150-197), all in ONE pass over the elements e with sq(x_e - c_i) <= r_i * r_i.  out16[Q x 16] doubles:
 *   [0] n  [1] sum |x_e - c_i| (fp32 norm per element)
 *   [2] #{e : n_e . qn_i > -1e-5}  [3..5] sum stableNormalized(x_e - c_i) over those   (zero when qn3 is null)
 *   [6..8] sum (x_e - c_i)  [9..14] sum (x_e - c_i)(x_e - c_i)^T as xx xy xz yy yz zz  [15] 0
 * Per-element terms are the reference's fp32 expressions; sums are double.  The reference accumulates in its own traversal
 * order (fp32 for the curvature offsets and the covariance), so it is itself tree dependent at fp32 rounding level; the caller
 * epilogues (include/thea_gpu/LocalFeatures.hpp) turn the sums into the features.  A negative radius matches nothing.
 * qn3 needs a structure built with normals (else TGQ_ESTATE). */
int tgq_ball_stats(tgq_handle * h, const float * c3, const float * qn3, const float * radius, int64_t Q, double * out16);

/* Q local distance histograms on a point structure.  Replaces the loop of
 *   LocalDistanceHistogram::compute(c_i, histogram, EUCLIDEAN, max_distance = r_i, sample_reduction_ratio >= 1)
 * (Algorithms/SurfaceFeatures/Local/LocalDistanceHistogram.cpp:36-88; Algorithms/Histogram.hpp:110-150): every element e in
 * Ball3(c_i, r_i) falls in bin clamp((intx)floor((double)|x_e - c_i| / bin_size), 0, num_bins - 1), bin_size =
 * max((double)r_i, 1e-30) / num_bins.  hist[Q x num_bins] are the raw counts (the reference's bins hold the same integers as
 * doubles, sumValues() is their sum); num_bins <= 256.  Integer counts: bit exact, order independent.  The reference's random
 * subsampling (sample_reduction_ratio < 1, Random::common()) is not reproduced. */
int tgq_ball_histogram(tgq_handle * h, const float * c3, const float * radius, int64_t Q, int32_t num_bins, uint32_t * hist);

/* SpinImage::compute(position, normal, num_radial_bins, num_height_bins, spin_image) for Q queries on a point structure
 * (Algorithms/SurfaceFeatures/Local/SpinImage.cpp:47-75): every query bins ALL samples of the structure by (radius about the axis
 * through the query along stableNormalized(normal), height along it), radial limits [0, max_extent], height limits
 * [-max_extent, max_extent], with the reference's fp32 expressions; the reference uses max_extent = 0.75f * surf->getScale().
 * image: [Q x num_radial_bins x num_height_bins] counts, radial bin major (the reference's spin_image(radial_bin, height_bin)).
 * Exact integers.  num_radial_bins * num_height_bins <= 1024. */
int tgq_spin_image(tgq_handle * h, const float * q3, const float * qn3, int64_t Q, int32_t num_radial_bins, int32_t num_height_bins,
                   float max_extent, uint32_t * image);

/* ---- shortest paths on the sample graph (what the reference does with ShortestPaths::dijkstraWithCallback, Algorithms/ShortestPaths.hpp:
 * 166-353, on PointSet3's proximity graph).  The graph is given as the (P x max_degree) rows tgq_ball_graph produces: idx (-1 padded)
 * and sep (the float separations), directed, host or device.  Path lengths are the reference's: left-to-right double sums of the
 * float edge lengths, minimised over paths -- bit-identical to Dijkstra's values.  The handle only supplies device, stream and
 * scratch; it needs no structure. */

/* PointSet3::extractOriginalAdjacencies (Algorithms/PointSet3.cpp:96-178): the oversampling contraction of updateGraph.  Samples
 * 0 .. n_orig-1 are the original ones, n_orig .. P-1 the dense oversampling.  For every original sample s: the max_degree original
 * samples v != s with the smallest graph distance over ALL samples (equal distances: lower index; the reference: heap pop order),
 * as rows ordered like the reference's neighbour set, by (float(distance), index); out_idx / out_sep [n_orig x max_degree], -1 / -1.0f
 * padded; out_count (may be null) [n_orig]. */
int tgq_graph_contract(tgq_handle * h, const int32_t * idx, const float * sep, int64_t P, int32_t max_degree, int64_t n_orig, int32_t * out_idx,
                       float * out_sep, int32_t * out_count);

/* The geodesic neighbourhood of Q seed samples: AverageDistance::compute(p, GEODESIC, d) and LocalDistanceHistogram::compute(p, hist,
 * GEODESIC, d, ratio >= 1) (Algorithms/SurfaceFeatures/Local/AverageDistance.cpp:84-118, LocalDistanceHistogram.cpp:90-137) after their
 * nearest-sample lookup (seeds[i] = closestElement(p_i), e.g. from tgq_closest on the point structure).  Every vertex with graph
 * distance <= max_distance from its seed (max_distance < 0: the whole component) contributes: sum [Q] (double; summed in device
 * order, the reference sums in ascending distance: equal to about 1e-15 relative), count [Q], and -- when hist is given -- one count
 * in bin clamp(floor(d / (max(hist_range, 1e-30) / num_bins))) of hist [Q x num_bins] (Histogram.hpp:110-150; the reference uses
 * hist_range = max_distance, or the scale when max_distance < 0).  The average distance is sum / count / hist_range.  Any output may
 * be null. */
int tgq_graph_geodesic(tgq_handle * h, const int32_t * idx, const float * sep, int64_t P, int32_t max_degree, const int32_t * seeds, int64_t Q,
                       float max_distance, float hist_range, int32_t num_bins, double * sum, int32_t * count, uint32_t * hist);

/* FurthestPointSampling::subsample (Algorithms/FurthestPointSampling.cpp:48-125; MeshSampler::sampleEvenlyBySeparation, MeshSampler.hpp:
 * 270-329, calls it on a 3x area oversampling): sample 0 first, then repeatedly the sample the reference's multi-source search
 * reports last -- the highest-index sample not reachable from the selected set if there is one, else the sample of maximal graph
 * distance to the set (equal distances: higher index; the reference: heap pop order).  Stops early, like the reference, when nothing
 * but selected samples is left.  selected [n_desired] (host or device), *n_selected (host) = number written. */
int tgq_graph_fps(tgq_handle * h, const int32_t * idx, const float * sep, int64_t P, int32_t max_degree, int64_t n_desired, int32_t * selected,
                  int64_t * n_selected);

/* Proximity graph over the P points of a point structure.  Replaces the loop of PointSet3::updateGraph over
 * findSampleNeighbors (PointSet3.hpp:380-381,586-681): for every sample, a ball range query of radius init_radius
 * (doubled for up to `rounds` rounds while fewer than min_degree neighbours were found), candidates filtered by
 * self, n.n' >= min_dot (when the structure has normals), kept as the max_degree smallest by (Math::fastSqrt(d2), idx).
 * idx[P x max_degree], sep[P x max_degree] ascending, unused slots -1; count[P] may be null.
 * Reference defaults: rounds 3, min_dot -0.5, min_degree clamp(int(0.25 * max_degree), 4, max_degree). */
int tgq_ball_graph(tgq_handle * h, int32_t max_degree, float init_radius, int32_t rounds, float min_dot, int32_t min_degree,
                   int32_t * idx, float * sep, int32_t * count);

/* The same with PointSet3's reachability test: updateGraph(surface) (PointSet3.hpp:324-391,639-648).  A
 * candidate j of sample i is dropped when rayIntersects(Ray3(x_i + 5 * sep * normalized(n_i + n_j), x_j - x_i), max_time = 1)
 * holds on the triangle structure `surface` (same device), i.e. when the lifted segment towards it is blocked by the surface.
 * The point structure must have normals (the reference skips the test without them; here that is TGQ_ESTATE). */
int tgq_ball_graph_surface(tgq_handle * h, tgq_handle * surface, int32_t max_degree, float init_radius, int32_t rounds, float min_dot,
                           int32_t min_degree, int32_t * idx, float * sep, int32_t * count);

/* RMS nearest-neighbour separation over n seed samples (PointSet3.hpp:363-376; the reference draws the seeds from
 * Random::common(), here the caller fixes them): sqrt(sum_f32 |x_nn(s) - x_s|^2 / n). */
int tgq_avg_separation(tgq_handle * h, const int32_t * seeds, int32_t n, float * out);

/* R ray first-hit queries on a triangle structure.  Replaces BvhN::rayStructureIntersection<RayIntersectionTester>
 * (ray, max_time) (BvhN.hpp:1291-1316,2240-2291) and rayIntersectionTime (BvhN.hpp:1271-1289).
 * tmax may be null (unbounded) else tmax[R], negative = unbounded; hits need 0 <= t and t < tmax (strict, BvhN.hpp:2257).
 * Outputs (each may be null): t[R] (-1 = miss), idx[R] (-1 = miss), normal3[R x 3] = unit normal of the hit triangle
 * (RayStructureIntersectionN::getNormal, Algorithms/RayQueryStructureN.hpp:28-53). */
/* The same queries with RECORD outputs and optional sharding of a replicated batch by space -- the multi-GPU form (SURVEY.md 8e:
 * BVH replicated, queries sharded, every rank ends holding all results).  All pointers are device pointers.  Every rank calls with
 * the SAME whole batch; rank `shard_rank` of `shard_world` answers the queries whose Morton cell it owns (cells dealt round-robin,
 * so each rank's share keeps the density of the full batch) and the kernel stores each result as one aligned record
 *     closest: 32 bytes { int32 idx; float d2 (-1 = none); float cp[3]; int32 pad[3] }      ray: 8 bytes { float t (-1 = miss); int32 idx }
 * at position [query index] of EVERY destination array dest[0 .. n_dest) (n_dest <= 8): the rank's own copy and, peer-mapped, the
 * copies of the other ranks (e.g. torch symmetric memory), written straight over NVLink while the traversal runs.  shard_world = 1,
 * n_dest = 1 is the plain single-GPU call with record output.  Records of queries other ranks own are left untouched.
 * Measured on 8 x B200: small scattered peer stores are expensive on both ends (7 peers: rays -56 %, closest -14 % per rank), so
 * the bench uses the COMPACT form instead: compact != 0 (n_dest = 1) writes the rank's share as a dense array in the order the
 * share was processed -- closest: the same 32-byte record with the query index in pad[0]; ray: 16 bytes { t, idx, query index, 0 }
 * -- which moves to the other ranks as one contiguous block per peer (copy engines) and is put in place there by
 * tgq_scatter_records.  *share (may be null) receives the number of queries this rank answered. */
int tgq_closest_records(tgq_handle * h, const float * q3, int64_t Q, double dist_bound, int32_t shard_rank, int32_t shard_world,
                        void * const * dest, int32_t n_dest, int32_t compact, int64_t * share);
int tgq_ray_first_records(tgq_handle * h, const float * o3, const float * d3, const float * tmax, int64_t R, int32_t shard_rank,
                          int32_t shard_world, void * const * dest, int32_t n_dest, int32_t compact, int64_t * share);
/* Splits n compact records by the rank that owns each record's position in the result array (query index / shard_len, the last
 * rank takes the remainder): bucket d = records [d * cap, d * cap + counts[d]) of `buckets`; counts[world] is int64 DEVICE memory
 * and is (re)written by the call.  Runs on the handle's stream (asynchronously with TGQ_OPT_ASYNC). */
int tgq_bucket_records(tgq_handle * h, const void * compact, int64_t n, int32_t kind, int64_t shard_len, int32_t world, void * buckets,
                       int64_t cap, int64_t * counts);
/* Compact records -> dest[query index] in the non-compact record layout.  compact: `slots` dense arrays of up to `cap` records each
 * (slot j starts at record j * cap; kind 0: closest, 32 bytes; kind 1: ray, 16 bytes), counts[slots] (int64, DEVICE memory: they
 * travel with the data) of them valid.  Runs on the handle's stream; returns without waiting when TGQ_OPT_ASYNC is set. */
int tgq_scatter_records(tgq_handle * h, const void * compact, int32_t slots, int64_t cap, const int64_t * counts, int32_t kind, void * dest);

int tgq_ray_first(tgq_handle * h, const float * o3, const float * d3, const float * tmax, int64_t R,
                  float * t, int32_t * idx, float * normal3);
/* BvhN::rayIntersects (BvhN.hpp:1266-1269) with early exit: hit[R] = 1/0. */
int tgq_ray_any(tgq_handle * h, const float * o3, const float * d3, const float * tmax, int64_t R, uint8_t * hit);

/* Shape diameter function for S samples with outward normals.  Replaces the loop of
 * ShapeDiameter::compute(position, normal, only_hit_interior_surfaces) (Algorithms/SurfaceFeatures/Local/
 * ShapeDiameter.hpp:168-258): 30 cone rays per sample, first hits, hit-normal sign test, median +- sigma outlier
 * rejection, cos-weighted mean, divided by `scale` and clamped to [0,1]; -1 when no ray counts.
 * Outputs (may be null): sdf[S] (double), ray_t[S x 30], ray_idx[S x 30]. */
int tgq_sdf(tgq_handle * h, const float * p3, const float * n3, int64_t S, float scale, int32_t only_hit_interior,
            double * sdf, float * ray_t, int32_t * ray_idx);

/* External visibility of S points against a triangle structure.  Replaces the loop of
 * Visibility::compute(position, num_rays) (Algorithms/SurfaceFeatures/Local/Visibility.hpp:104-125): for every direction u,
 * the ray (p + 0.001f * scale * u, u) is tested with rayIntersects; visibility[i] = (double)num_escaped / num_dirs.
 * The reference draws u from Random::common().sphere() per ray; here the caller supplies the directions: dirs3 is
 * num_dirs x 3 shared by all samples (per_sample_dirs == 0) or S x num_dirs x 3.  scale = diagonal of the structure's bounds
 * (Visibility.hpp:73,88).  Outputs may be null: visibility[S], num_escaped[S]. */
int tgq_visibility(tgq_handle * h, const float * p3, int64_t S, const float * dirs3, int32_t num_dirs, int32_t per_sample_dirs, float scale,
                   double * visibility, int32_t * num_escaped);

/* ICP sums for n correspondences (from3[i] -> to3[i]).  Default (TGQ_OPT_ICP_REFERENCE_ORDER = 1): the reference's sequential
 * fp32 loops, every output an fp32 value.  With the option at 0: order-independent fixed-point accumulation, identical for any
 * launch shape on one device (sharded front ends that combine per-shard sums quantise each shard with its own unit: equal to
 * 1e-9 relative, not bit for bit, across GPU counts).  out[0..2] = mean p, out[3..5] = mean q, out[6..14] = sum (p - pm)(q - qm)^T row-major, and with tr12
 * (row-major 3x4 [L|t], may be null = identity) out[15] = |tr pm - qm|^2 + sum |L (p - pm) - (q - qm)|^2.
 * Replaces the accumulation loops of Icp3::alignOneStep / measureError (Algorithms/Icp3.hpp:298-367,410-464). */
int tgq_icp_sums(tgq_handle * h, const float * from3, const float * to3, int64_t n, const float * tr12, double * out16);
/* The whole of Icp3<Real>::align(n, from, to_structure, &error) (Algorithms/Icp3.hpp:110-126,193-275; no weights, symmetry
 * planes or up vector) against the triangle structure, device resident: every iteration is one closest-point batch, two
 * reductions and one transform on the GPU; only the 16 sums come back for the 3x3 polar decomposition on the host.
 * Reference defaults: min_iterations 3, max_iterations 10, fractional_error_threshold 1e-4.  The reference's control flow is
 * kept to the letter, including alignOneStep pairing the updated source positions with the targets found before the update.
 * Outputs: tr12 (row-major 3x4 [R | t], host), error (host, may be null; want_error != 0 measures after the last iteration as
 * the reference does when an error pointer is passed), batches = closest-point batches run (host, may be null),
 * correspondences = batches x n element indices in batch order (host or device, may be null; room for max_iterations + 1). */
int tgq_icp_align(tgq_handle * h, const float * from3, int64_t n, int32_t min_iterations, int32_t max_iterations,
                  double fractional_error_threshold, int32_t want_error, float * tr12, float * error, int32_t * batches,
                  int32_t * correspondences);
/* out3[i] = L in3[i] + t with the reference's fp32 operation order (AffineTransformN.hpp:160). */
int tgq_transform_points(tgq_handle * h, const float * in3, int64_t n, const float * tr12, float * out3);

/* Area-weighted surface samples of a triangle structure: MeshSampler::sampleEvenlyByArea in EXACT mode
 * (Algorithms/MeshSampler.hpp:147-185, Triangle3.hpp:179-195) driven by the counter-based generator of
 * thea_b200/gen.py (sample i uses draws 3i..3i+2 of splitmix64 stream `seed`).  Outputs may be null. */
int tgq_area_sample(tgq_handle * h, int64_t S, uint64_t seed, float * p3, float * n3, int32_t * tri);

/* ---- introspection ------------------------------------------------------------------------------------------------ */

typedef struct tgq_stats
{
  int64_t num_elements;
  int64_t num_nodes;        /* inner nodes of the device LBVH */
  int64_t node_bytes;       /* HBM bytes of the node array */
  int64_t element_bytes;    /* HBM bytes of the element records */
  double last_build_ms;     /* device time of the last build / refit */
  double last_query_ms;     /* device time of the kernels of the last query call (sort + traversal), CUDA events */
  double last_sort_ms;      /* share of last_query_ms spent ordering the batch */
  int64_t last_launches;    /* kernels launched by the last query call */
  int64_t total_launches;   /* kernels launched since tgq_create */
} tgq_stats;

int tgq_get_stats(tgq_handle * h, tgq_stats * out);

#ifdef __cplusplus
}
#endif
#endif
