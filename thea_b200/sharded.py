"""Multi-GPU driver: the BVH is replicated (every rank builds the same structure), the query batch is sharded
contiguously over the ranks, and the per-shard results are gathered once at the end.

One process per GPU, ``torch.distributed`` for the plumbing: NCCL over NVLink/NVSwitch when the result tensors live on
the GPU, gloo when they are host tensors (the CPU tests).  The queries themselves never cross GPUs: each one reads the
immutable structure and writes its own result slot, so there is no data-path collective; the only exchanges are the
final gather of the result arrays and, for ICP, 17 doubles per iteration.

Nothing here computes queries: `engine` is a thea_b200.Engine (or anything with the same methods).
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [begin, begin + count) of `total` items for `rank` of `world` (sizes differ by at most one)."""
    b = total * rank // world
    e = total * (rank + 1) // world
    return b, e - b


def _dist():
    import torch.distributed as dist
    return dist


def world_info(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_shards(local, total: int, group=None):
    """All-gathers the per-rank shards (first dimension sharded by shard_range) into the full array on every rank.

    `local` is a torch tensor (CUDA -> NCCL, CPU -> gloo) or a numpy array (gathered through a CPU tensor).  Shards are
    padded to the largest shard so that one all_gather_into_tensor moves everything; the result is byte-identical to
    what a single rank would have produced for the whole batch.
    """
    import torch
    dist = _dist()
    rank, world = world_info(group)
    is_np = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if is_np else local.contiguous()
    if world == 1:
        return local
    counts = [shard_range(total, r, world)[1] for r in range(world)]
    assert t.shape[0] == counts[rank], f"rank {rank}: shard has {t.shape[0]} rows, expected {counts[rank]}"
    mx = max(counts)
    tail = tuple(t.shape[1:])
    if t.shape[0] < mx:
        pad = torch.zeros((mx - t.shape[0],) + tail, dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], dim=0)
    out = torch.empty((world * mx,) + tail, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    if all(c == mx for c in counts):
        full = out
    else:
        full = torch.cat([out[r * mx:r * mx + counts[r]] for r in range(world)], dim=0)
    return full.numpy() if is_np else full


def space_shard_indices(q3, lo, hi, rank: int, world: int, bits: int = 5):
    """Indices (ascending) of the queries dealt to `rank` when the batch is sharded BY SPACE instead of by position: the box
    [lo, hi] (every rank passes the same one, e.g. the structure's bounds scaled about its centre) is cut into 2^bits cells per axis,
    the cells are numbered along the Morton curve and dealt to the ranks round-robin.  Every rank then works on spatially dense
    pieces of the batch whatever N is (a contiguous shard of a random batch gets N times sparser, which costs the
    traversal kernels 13-26 % at N = 4-8), and neighbouring cells go to different ranks, which balances the load.  Deterministic and
    identical on every rank; queries outside the box fall into the border cells.  numpy arrays or torch tensors (any device)."""
    is_np = isinstance(q3, np.ndarray)
    n = 1 << bits
    if is_np:
        lo_ = np.asarray(lo, np.float32); ext = np.maximum(np.asarray(hi, np.float32) - lo_, np.float32(1e-30))
        g = np.clip(np.floor((q3.astype(np.float32) - lo_) / ext * np.float32(n)), 0, n - 1).astype(np.int64)
        cell = np.zeros(len(q3), np.int64)
        for b in range(bits):
            cell |= ((g[:, 0] >> b) & 1) << (3 * b + 2) | ((g[:, 1] >> b) & 1) << (3 * b + 1) | ((g[:, 2] >> b) & 1) << (3 * b)
        return np.flatnonzero(cell % world == rank)
    import torch
    lo_ = torch.as_tensor(np.asarray(lo, np.float32), device=q3.device)
    ext = torch.clamp(torch.as_tensor(np.asarray(hi, np.float32), device=q3.device) - lo_, min=1e-30)
    g = torch.clamp(torch.floor((q3 - lo_) / ext * float(n)), 0, n - 1).to(torch.int64)
    cell = torch.zeros(q3.shape[0], dtype=torch.int64, device=q3.device)
    for b in range(bits):
        cell |= (((g[:, 0] >> b) & 1) << (3 * b + 2)) | (((g[:, 1] >> b) & 1) << (3 * b + 1)) | (((g[:, 2] >> b) & 1) << (3 * b))
    return torch.nonzero(cell % world == rank).reshape(-1)


def gather_scattered(local, sel, total: int, group=None):
    """Counterpart of gather_shards for shards that are index SETS: every rank contributes rows `local` for the global positions
    `sel` (the sets partition range(total)); returns the full array on every rank.  Lists are padded to the largest shard, all-gathered
    together with their positions, and scattered locally."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    is_np = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if is_np else local.contiguous()
    s = torch.from_numpy(np.ascontiguousarray(sel)).to(torch.int64) if isinstance(sel, np.ndarray) else sel.to(torch.int64)
    tail = tuple(t.shape[1:])
    full = torch.empty((total,) + tail, dtype=t.dtype, device=t.device)
    if world == 1:
        full[s] = t
        return full.numpy() if is_np else full
    cnt = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    cnts = torch.empty(world, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    counts = [int(c) for c in cnts.cpu()]
    assert sum(counts) == total, f"the index sets of the ranks hold {sum(counts)} positions, expected {total}"
    mx = max(max(counts), 1)
    pt = torch.zeros((mx,) + tail, dtype=t.dtype, device=t.device); pt[:t.shape[0]] = t
    ps = torch.zeros(mx, dtype=torch.int64, device=t.device); ps[:s.shape[0]] = s
    all_t = torch.empty((world * mx,) + tail, dtype=t.dtype, device=t.device)
    all_s = torch.empty(world * mx, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(all_t, pt, group=group)
    dist.all_gather_into_tensor(all_s, ps, group=group)
    for r in range(world):
        full[all_s[r * mx:r * mx + counts[r]]] = all_t[r * mx:r * mx + counts[r]]
    return full.numpy() if is_np else full


def combine_icp_sums(parts, tr12=None):
    """Combines per-shard tgq_icp_sums outputs [(n_s, out16_s), ...] (in rank order) into the sums of the whole set.

    out16 = (mean p, mean q, centred cross-covariance, error) as documented in include/thea_gpu_query.h; the pooled
    statistics follow from the raw sums, accumulated in float64 in rank order (deterministic for a given world size; every
    shard quantises its fixed-point sums with its own unit, so different world sizes agree to ~1e-9 relative, not bit for bit).
    """
    L = np.eye(3, 4) if tr12 is None else np.asarray(tr12, np.float64).reshape(3, 4)
    N = 0.0
    sp = np.zeros(3)
    sq = np.zeros(3)
    spq = np.zeros((3, 3))
    tot = 0.0
    for n_s, o in parts:
        o = np.asarray(o, np.float64)
        n_s = float(n_s)
        if n_s <= 0:
            continue
        pm, qm, cov, err = o[0:3], o[3:6], o[6:15].reshape(3, 3), o[15]
        rbar = L[:, :3] @ pm + L[:, 3] - qm
        N += n_s
        sp += n_s * pm
        sq += n_s * qm
        spq += cov + n_s * np.outer(pm, qm)
        tot += (err - rbar @ rbar) + n_s * (rbar @ rbar)       # raw sum |L p + t - q|^2 of the shard
    out = np.zeros(16)
    if N == 0:
        return out
    pm, qm = sp / N, sq / N
    out[0:3], out[3:6] = pm, qm
    out[6:15] = (spq - N * np.outer(pm, qm)).reshape(9)
    rbar = L[:, :3] @ pm + L[:, 3] - qm
    out[15] = rbar @ rbar + (tot - N * (rbar @ rbar))
    return out


class ShardedEngine:
    """Query front end for one rank of a replicated-BVH / sharded-queries job.

    Every rank calls the same method with the same GLOBAL query arrays (or with `local=True` and only its own shard);
    each returns the gathered global results.
    """

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.rank, self.world = world_info(group)

    def _slice(self, a, total, local):
        if local:
            return a
        b, c = shard_range(total, self.rank, self.world)
        return a[b:b + c]

    def closest(self, q3, total=None, local=False, space_box=None, **kw):
        """space_box = (lo, hi): shard the GLOBAL batch by space (space_shard_indices) instead of by position."""
        total = int(q3.shape[0]) if total is None else total
        if kw.get("exclude_self"):
            # "self" is the element with the query's GLOBAL position; a shard only knows local positions
            raise ValueError("exclude_self needs global query positions: query the shard through Engine.closest on a slice of the build array")
        if space_box is not None:
            sel = space_shard_indices(q3, space_box[0], space_box[1], self.rank, self.world)
            idx, d2, cp = self.engine.closest(q3[sel], **kw)
            return tuple(None if x is None else gather_scattered(x, sel, total, self.group) for x in (idx, d2, cp))
        idx, d2, cp = self.engine.closest(self._slice(q3, total, local), **kw)
        return tuple(None if x is None else gather_shards(x, total, self.group) for x in (idx, d2, cp))

    def ray_first(self, o3, d3, total=None, local=False, space_box=None, **kw):
        total = int(o3.shape[0]) if total is None else total
        if space_box is not None:                     # by ray origin
            sel = space_shard_indices(o3, space_box[0], space_box[1], self.rank, self.world)
            t, idx, _ = self.engine.ray_first(o3[sel], d3[sel], **kw)
            return gather_scattered(t, sel, total, self.group), gather_scattered(idx, sel, total, self.group)
        t, idx, _ = self.engine.ray_first(self._slice(o3, total, local), self._slice(d3, total, local), **kw)
        return gather_shards(t, total, self.group), gather_shards(idx, total, self.group)

    def closest_segment(self, a3, b3, dist_bound=-1.0, exclude=None, total=None, local=False, space_box=None, **kw):
        """Line-segment query objects (tgq_closest_segment) of the query shard, gathered: (idx, d2, element point, segment point) for
        the whole batch on every rank.  exclude follows the queries; space_box shards by the first endpoint."""
        total = int(a3.shape[0]) if total is None else total
        if space_box is not None:
            sel = space_shard_indices(a3, space_box[0], space_box[1], self.rank, self.world)
            out = self.engine.closest_segment(a3[sel], b3[sel], dist_bound, None if exclude is None else exclude[sel], **kw)
            return tuple(gather_scattered(x, sel, total, self.group) for x in out)
        out = self.engine.closest_segment(self._slice(a3, total, local), self._slice(b3, total, local), dist_bound,
                                          None if exclude is None else self._slice(exclude, total, local), **kw)
        return tuple(gather_shards(x, total, self.group) for x in out)

    def closest_triangle(self, q9, dist_bound=-1.0, total=None, local=False):
        """Triangle query objects (tgq_closest_triangle) of the query shard, gathered: (idx, d2, element point, query point)."""
        total = int(q9.shape[0]) if total is None else total
        out = self.engine.closest_triangle(self._slice(q9, total, local), dist_bound)
        return tuple(gather_shards(x, total, self.group) for x in out)

    def closest_pair_of_triangles(self, q9, dist_bound=-1.0):
        """BvhN::closestPair between this (replicated) structure and the mesh given by its triangles q9 [T x 9]: every rank answers its
        share of the query triangles, the canonical minimum -- smallest (d2, target index, query index) -- is taken over the gathered
        rows, identically on every rank.  Returns (target index, query index, d2, query point, target point), -1 indices when none.
        dist_bound here is a per-query-object bound (with the reference's widening); pass -1 and filter for the plain form."""
        idx, d2, ep, qp = self.closest_triangle(q9, dist_bound)
        idx, d2 = np.asarray(idx), np.asarray(d2)
        ok = np.flatnonzero(idx >= 0)
        if ok.size == 0:
            return -1, -1, -1.0, np.zeros(3, np.float32), np.zeros(3, np.float32)
        k = ok[np.lexsort((ok, idx[ok], d2[ok]))][0]
        return int(idx[k]), int(k), float(d2[k]), np.asarray(qp)[k], np.asarray(ep)[k]

    def knn(self, q3, k, total=None, local=False, **kw):
        total = int(q3.shape[0]) if total is None else total
        b, _ = shard_range(total, self.rank, self.world)
        q = self._slice(q3, total, local)
        if kw.get("exclude_self"):
            # "self" is the element with the query's GLOBAL position: shift the shard's indices
            raise ValueError("exclude_self needs global query positions: query the shard through Engine.knn on a slice of the build array")
        idx, d2 = self.engine.knn(q, k, **kw)
        return gather_shards(idx, total, self.group), gather_shards(d2, total, self.group)

    def _slice_per_query(self, a, total, local):
        """Per-query arguments (radii) follow the query shard; scalars are shared by every query."""
        if a is None or np.ndim(a) == 0:
            return a
        return self._slice(a, total, local)

    def ball_stats(self, c3, radius, qn3=None, total=None, local=False):
        """Per-ball neighbourhood sums (tgq_ball_stats) of the query shard, gathered: [total x 16] float64 on every rank."""
        total = int(c3.shape[0]) if total is None else total
        st = self.engine.ball_stats(self._slice(c3, total, local), self._slice_per_query(radius, total, local),
                                    None if qn3 is None else self._slice(qn3, total, local))
        return gather_shards(st, total, self.group)

    def ball_histogram(self, c3, radius, num_bins, total=None, local=False):
        total = int(c3.shape[0]) if total is None else total
        h = self.engine.ball_histogram(self._slice(c3, total, local), self._slice_per_query(radius, total, local), num_bins)
        return gather_shards(h, total, self.group)

    def visibility(self, p3, dirs3, scale, total=None, local=False):
        """Visibility of the sample shard ([D x 3] shared directions, or [S x D x 3] sharded with the samples)."""
        total = int(p3.shape[0]) if total is None else total
        d = dirs3 if np.ndim(dirs3) == 2 else self._slice(dirs3, total, local)
        vis, nesc = self.engine.visibility(self._slice(p3, total, local), d, scale)
        return gather_shards(vis, total, self.group), gather_shards(nesc, total, self.group)

    def icp_sums(self, from3, to3, tr12=None):
        """from3 / to3: this rank's shard of the correspondences; returns the pooled 16 sums on every rank."""
        import torch
        dist = _dist()
        n = int(from3.shape[0])
        local = np.zeros(17)
        if n > 0:
            local[0] = n
            local[1:] = self.engine.icp_sums(from3, to3, tr12)
        if self.world == 1:
            return combine_icp_sums([(local[0], local[1:])], tr12)
        dev = "cuda" if (torch.cuda.is_available() and dist.get_backend(self.group) == "nccl") else "cpu"
        t = torch.from_numpy(local).to(dev)
        out = torch.empty(self.world * 17, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, t, group=self.group)
        parts = out.cpu().numpy().reshape(self.world, 17)
        return combine_icp_sums([(p[0], p[1:]) for p in parts], tr12)
