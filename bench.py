#!/usr/bin/env python
"""bench.py -- BVH queries/sec on the BASELINE.json configs, with the HBM roofline and the CPU reference beside it.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload cfg5|cfg1] [--scale F]

Workload (default cfg5, BASELINE.json configs[4]): 10M-triangle bumpy torus (9,999,392 triangles), 100M mixed queries
= 50M closest-point (BvhN::closestElement with closest point) + 50M ray first-hit (BvhN::rayStructureIntersection),
sharded contiguously over the ranks (one C-ABI call per query kind per step); the BVH is replicated (each rank builds it).
One step = one pass over all queries of the rank.  `value` is device-resident throughput (inputs already in HBM);
`e2e` goes through the same C-ABI calls with pinned HOST buffers, H2D/D2H inside the timed region.

The CPU oracle (oracle/, the restatement of the reference's BvhN in reference traversal order) is used here ONLY for
the `cpu_baseline` object and for `--impl reference`; it is never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from thea_b200 import gen  # noqa: E402

BLOCK = 1_000_000
WORKLOADS = {
    # name: (torus grid, closest queries, rays, description)
    "cfg5": (gen.TORUS_10M, 50_000_000, 50_000_000, "cfg5: 9,999,392-tri torus, 50M closest-point + 50M ray first-hit"),
    "cfg1": (gen.TORUS_70K, 1_000_000, 0, "cfg1: 70,688-tri torus, 1M closest-point queries"),
    "cfg4r": (gen.TORUS_1M, 0, 3_000_000, "cfg4-like: 1,002,528-tri torus, 3M random first-hit rays"),
    "cfg1m": (gen.TORUS_1M, 5_000_000, 5_000_000, "1,002,528-tri torus, 5M closest-point + 5M ray first-hit"),
}


def algorithmic_bytes(T: int, kind: str) -> int:
    """in + out + D*2*24 + 5*elem with D = ceil(log2(T/5)) (one root-to-leaf path, leaf 5)."""
    D = int(np.ceil(np.log2(T / 5.0)))
    if kind == "closest":
        return 12 + 20 + D * 48 + 5 * 36
    return 28 + 8 + D * 48 + 5 * 36


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def torch_u01(seed: int, n: int, offset: int, device):
    """thea_b200.gen.u01 on a torch device (int64 two's-complement arithmetic == uint64 wraparound)."""
    import torch

    def c(x):  # uint64 constant as int64
        return x - (1 << 64) if x >= (1 << 63) else x

    def lsr(z, s):
        return (z >> s) & ((1 << (64 - s)) - 1)

    i = torch.arange(offset + 1, offset + n + 1, dtype=torch.int64, device=device)
    z = i * c(0x9E3779B97F4A7C15) + c(seed)
    z = (z ^ lsr(z, 30)) * c(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * c(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    return lsr(z, 40).to(torch.float32) * (2.0 ** -24)


def torch_closest_queries(count, bounds, seed, offset, device):
    import torch
    lo, hi = gen.box_scaled(*bounds, 1.5)
    u = torch_u01(seed, 3 * count, 3 * offset, device).reshape(count, 3)
    lo_t = torch.from_numpy(lo).to(device)
    hi_t = torch.from_numpy(hi).to(device)
    return (lo_t + u * (hi_t - lo_t)).contiguous()


def torch_random_rays(count, bounds, seed, offset, device):
    import torch
    lo, hi = gen.box_scaled(*bounds, 1.5)
    u = torch_u01(seed, 9 * count, 9 * offset, device).reshape(count, 9)
    lo_t = torch.from_numpy(lo).to(device)
    hi_t = torch.from_numpy(hi).to(device)
    o = (lo_t + u[:, 0:3] * (hi_t - lo_t)).contiguous()
    u1 = torch.clamp(u[:, 3:6].double(), min=2.0 ** -25)
    u2 = u[:, 6:9].double()
    g = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * np.pi * u2)
    nrm = torch.sqrt((g * g).sum(dim=1, keepdim=True))
    nrm[nrm == 0] = 1.0
    return o, (g / nrm).float().contiguous()


def shard(total: int, rank: int, world: int):
    b = total * rank // world
    e = total * (rank + 1) // world
    return b, e - b


def blocks(n: int):
    return [(b, min(BLOCK, n - b)) for b in range(0, n, BLOCK)]


class _CpuArm:
    """The CPU side timed by the reference arm and the cpu_baseline leg: closest(q) / ray_first(o, d) on all host cores."""

    def __init__(self, obj, closest, ray_first):
        self.obj, self.closest, self.ray_first = obj, closest, ray_first


def cpu_reference(mesh):
    """(arm, kind, cores, description): the reference's OWN BvhN code (oracle/_ref/libthea_ref.so: sidch/Thea's unmodified sources
    compiled against the Eigen stand-in of oracle/eigen_shim) when that library exists, else the oracle port in reference order."""
    from oracle import orc
    cores = orc.ncores()
    try:
        from oracle import ref
        if ref.available():
            R = ref.Reference(tris9=mesh)
            return (_CpuArm(R, lambda q: R.closest(q, threads=cores), lambda o, d: R.ray_first(o, d, threads=cores)), "reference", cores,
                    "CPU = the reference's own BvhN::closestPair / rayStructureIntersection (Thea sources compiled unmodified against an Eigen "
                    "stand-in), its single-threaded query loop spread over the host threads")
    except Exception as exc:      # noqa: BLE001
        print(f"bench.py: reference library unavailable ({type(exc).__name__}: {exc}); timing the oracle port", file=sys.stderr)
    O = orc.Oracle(tris9=mesh)
    return (_CpuArm(O, lambda q: O.closest(q, mode=orc.MODE_REFERENCE, threads=cores), lambda o, d: O.ray_first(o, d, mode=orc.MODE_REFERENCE, threads=cores)),
            "port", cores, "CPU = oracle port of the reference kd-tree (leaf 5, x1.05 boxes) in DFS order")


def run_reference(args, world, rank):
    """--impl reference: the CPU restatement of the reference's BvhN (kd-tree, DFS order) on all host cores."""
    if rank != 0:
        return
    grid, nq, nr, desc = WORKLOADS[args.workload]
    mesh = gen.torus_mesh(max(8, int(grid * args.scale)))
    T = len(mesh)
    bounds = gen.mesh_bounds(mesh)
    t0 = time.time()
    ora, kind, cores, what = cpu_reference(mesh)
    build_s = time.time() - t0
    # bounded sample per step, sized from a pilot so that steps+warmup stay within a few minutes
    sq = 2000 if nq else 0
    sr = 20000 if nr else 0
    t0 = time.time()
    if sq:
        ora.closest(gen.closest_queries(sq, bounds))
    tq = time.time() - t0
    t0 = time.time()
    if sr:
        o, d = gen.random_rays(sr, bounds)
        ora.ray_first(o, d)
    tr = time.time() - t0
    budget = 60.0 / max(1, args.steps + args.warmup)          # seconds per step
    mix_q = nq / max(1, nq + nr)
    per_pair = (tq / sq if sq else 0) * mix_q + (tr / sr if sr else 0) * (1 - mix_q)
    n_total = int(max(2000, min(2_000_000, budget / max(per_pair, 1e-9))))
    n_q = int(n_total * mix_q)
    n_r = n_total - n_q
    q = gen.closest_queries(n_q, bounds) if n_q else None
    if n_r:
        o, d = gen.random_rays(n_r, bounds)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if n_q:
            ora.closest(q)
        if n_r:
            ora.ray_first(o, d)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = (n_q + n_r) * len(times) / total
    sample = f"{n_q} closest + {n_r} rays per step (prefix of the same streams), kd-tree build {build_s:.1f}s excluded; {what}"
    line = {"impl": "reference", "metric": "bvh_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "triangles": T, "scale": args.scale},
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="development aid: scales the torus grid and query counts")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    ap.add_argument("--chunk", type=int, default=0, help="queries per staged chunk of the host-buffer leg (0: library default)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, world, rank)
        return

    import torch
    import torch.distributed as dist
    from thea_b200 import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    grid, nq_total, nr_total, desc = WORKLOADS[args.workload]
    grid = max(8, int(grid * args.scale))
    nq_total = int(nq_total * args.scale)
    nr_total = int(nr_total * args.scale)
    mesh = gen.torus_mesh(grid)
    T = len(mesh)
    bounds = gen.mesh_bounds(mesh)

    # everything below -- torch ops and the engine's kernels -- is ordered on ONE non-default stream
    main_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(main_stream)
    eng = Engine(local_rank)
    eng.set_stream(main_stream.cuda_stream)
    if os.environ.get("TGQ_BENCH_CELL_BITS"):
        eng.set_option(8, int(os.environ["TGQ_BENCH_CELL_BITS"]))
    d_mesh = torch.from_numpy(mesh).to(dev)
    eng.build_triangles(d_mesh)
    build_first_ms = eng.stats()["last_build_ms"]        # includes the first-call allocation of every build and structure buffer
    eng.build_triangles(d_mesh)                          # the same structure again: the steady-state build time
    build_ms = eng.stats()["last_build_ms"]
    del d_mesh

    # ---- device-resident leg -------------------------------------------------------------------------------------------------------
    # Every rank holds the WHOLE batch (generated on the device from the counter-based streams) and, inside the timed step, takes the
    # queries whose Morton cell it owns (tgq_*_records: cells dealt round-robin, so its share keeps the density of the full batch -- a
    # contiguous index range of a random batch is `world` times sparser after the Morton sort, which cost 22 % per rank at N = 8 in
    # round 1).  The one real exchange of the path -- every rank ends a step holding the results of the WHOLE batch, in query order
    #  -- is built from contiguous copy-engine transfers only, so that nothing but 1/N of the work is ever exposed:
    #   1. the traversal kernels write the rank's share as a dense array of records (32 B closest, 16 B ray, each carrying its query
    #      index); a bucket kernel splits it by the rank that OWNS each record's position in the result array (query index // (Q/N));
    #   2. all-to-all by peer-to-peer copies over NVLink on a side stream (one contiguous copy per peer, the counts alongside), then
    #      a device-side barrier on that stream: everybody's buckets have landed;
    #   3. a scatter kernel puts the records a rank received at [query index] -- all inside ITS contiguous 1/N of the result arrays;
    #   4. that finished 1/N is pushed into every peer's copy of the result arrays (contiguous, copy engines, under the next kernels).
    # Measured and dropped: storing each result from the traversal kernel straight into every peer's copy (small scattered peer
    # stores cost the ray kernel 56 % and the closest kernel 14 % at N = 8), and pushing whole shares to every rank and scattering
    # all 100 M records everywhere (3.6 ms per step whatever N is).
    # Fallback when symmetric memory is unavailable: contiguous shards + NCCL all-gather.
    q_off, nq = shard(nq_total, rank, world)
    r_off, nr = shard(nr_total, rank, world)
    dq = torch.empty((nq_total, 3), dtype=torch.float32, device=dev)
    for b, n in blocks(nq_total):
        dq[b:b + n] = torch_closest_queries(n, bounds, gen.SEED_CLOSEST, b, dev)
    dro = torch.empty((nr_total, 3), dtype=torch.float32, device=dev)
    drd = torch.empty((nr_total, 3), dtype=torch.float32, device=dev)
    for b, n in blocks(nr_total):
        o, d = torch_random_rays(n, bounds, gen.SEED_RAYS, b, dev)
        dro[b:b + n] = o
        drd[b:b + n] = d
    kern = {"closest_ms": 0.0, "closest_q": 0, "ray_ms": 0.0, "ray_q": 0, "sort_ms": 0.0, "launches": 0, "scatter_ms": 0.0}

    transport, sym = "none", None
    totals = {"c": nq_total, "r": nr_total}
    cwords, fwords, kinds = {"c": 8, "r": 4}, {"c": 8, "r": 2}, {"c": 0, "r": 1}
    if world > 1 and os.environ.get("TGQ_BENCH_TRANSPORT", "p2p") == "p2p":
        try:
            import torch.distributed._symmetric_memory as symm_mem
            assert nq_total % world == 0 and nr_total % world == 0, "bench workloads shard evenly over 1/2/4/8 ranks"
            sym = {}
            for name in ("c", "r"):
                tot = totals[name]
                cap_pair = int(tot / world / world * 1.1) + 4096                 # records one rank sends to one rank
                words = cap_pair * cwords[name]
                n_recv = 2 * world * words                                        # [parity][source rank][cap_pair records]
                n_cnt = 2 * world * 2                                             # int64 counts [parity][source rank]
                n_fin = max(1, tot) * fwords[name]                                # the gathered result array, query order
                numel = n_recv + n_cnt + n_fin
                t = symm_mem.empty(numel, dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(t, dist.group.WORLD)
                t.zero_()
                views = []
                for r in range(world):
                    buf = hdl.get_buffer(r, (numel,), torch.float32)
                    views.append({"recv": buf[:n_recv].view(2, world, words), "cnt": buf[n_recv:n_recv + n_cnt].view(torch.int64).view(2, world),
                                  "final": buf[n_recv + n_cnt:].view(max(1, tot), fwords[name])})
                cap_share = int(tot / world * 1.25) + 4096
                sym[name] = {"hdl": hdl, "views": views, "cap_pair": cap_pair, "cap_share": cap_share, "shard_len": tot // world,
                             "share": torch.empty(cap_share * cwords[name], dtype=torch.float32, device=dev),
                             "buckets": torch.empty((world, words), dtype=torch.float32, device=dev),
                             "bcnt": torch.zeros(world, dtype=torch.int64, device=dev)}
            crec, rrec = sym["c"]["views"][rank]["final"], sym["r"]["views"][rank]["final"]
            comm_stream = torch.cuda.Stream(device=dev)
            transport = ("dense per-rank record arrays, bucketed by owner of the result position; all-to-all and all-gather by copy-engine "
                         "P2P copies into symmetric memory (NVLink), device-side barrier, local scatter of 1/N of the records")
        except Exception as exc:      # noqa: BLE001 -- any failure here means: use NCCL
            if rank == 0:
                print(f"bench.py: symmetric memory unavailable ({type(exc).__name__}: {exc}); using contiguous shards + NCCL all-gather", file=sys.stderr)
            sym = None
    space_sharded = world == 1 or sym is not None
    if sym is None:
        crec = torch.zeros((max(1, nq_total), 8), dtype=torch.float32, device=dev)     # gathered results, query order: 32-byte records
        rrec = torch.zeros((max(1, nr_total), 2), dtype=torch.float32, device=dev)     # 8-byte records
        if world > 1:
            transport = "NCCL all-gather of contiguous shards"
    eng.set_option(3, 1)              # TGQ_OPT_ASYNC: calls are ordered on the stream; the step is timed with CUDA events
    timers = {"c": [], "r": [], "s": []}
    pending = []
    step_no = [0]

    def timed(kind, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        timers[kind].append((a, b))
        return out

    def exchange_a(name, n_own, par):
        """Steps 1b-2 for one query kind: bucket the share, all-to-all on the side stream; returns the event 'all buckets landed'."""
        S = sym[name]
        cur = torch.cuda.current_stream()
        timed("s", lambda: eng.bucket_records(S["share"], n_own, kinds[name], S["shard_len"], world, S["buckets"], S["cap_pair"], S["bcnt"]))
        comm_stream.wait_stream(cur)
        with torch.cuda.stream(comm_stream):
            for k in range(world):
                p = (rank + k) % world
                dst = S["views"][p]
                dst["recv"][par, rank].copy_(S["buckets"][p], non_blocking=True)
                dst["cnt"][par, rank].copy_(S["bcnt"][p], non_blocking=True)
            S["hdl"].barrier(channel=par)
            ev = torch.cuda.Event()
            ev.record(comm_stream)
        return ev

    def exchange_b(name, ev, par):
        """Steps 3-4: scatter what this rank received into its 1/N of the result array, push that 1/N to every peer."""
        S = sym[name]
        mine = S["views"][rank]
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        timed("s", lambda: eng.scatter_records(mine["recv"][par], world, S["cap_pair"], mine["cnt"][par], kinds[name], mine["final"]))
        comm_stream.wait_stream(cur)
        with torch.cuda.stream(comm_stream):
            lo = rank * S["shard_len"]
            for k in range(1, world):
                p = (rank + k) % world
                S["views"][p]["final"][lo:lo + S["shard_len"]].copy_(mine["final"][lo:lo + S["shard_len"]], non_blocking=True)

    pipelined = os.environ.get("TGQ_BENCH_PIPELINE", "1") == "1"

    def exchange_pipelined(name, n_own, par):
        """The whole exchange of one query kind with only the bucketing on the main stream: all-to-all, device-side barrier, local
        scatter and the pushes of the scattered slice run on the side stream, so the next traversal kernel starts at once -- the closest
        records are exchanged under the ray kernel of the same step, the ray records under the closest kernel of the NEXT step (a
        software pipeline one step deep: the timed region still ends only after every transfer has completed).  Buffer reuse: the
        main stream waits, before it buckets the next step's records of this kind, for the event 'this step's buckets have left';
        the receive buffers alternate by step parity and are released by the next step's barrier on the same (in-order) side stream."""
        S = sym[name]
        mine = S["views"][rank]
        cur = torch.cuda.current_stream()
        if S.get("ev_left") is not None:
            cur.wait_event(S["ev_left"])
        timed("s", lambda: eng.bucket_records(S["share"], n_own, kinds[name], S["shard_len"], world, S["buckets"], S["cap_pair"], S["bcnt"]))
        comm_stream.wait_stream(cur)
        with torch.cuda.stream(comm_stream):
            for k in range(world):
                p = (rank + k) % world
                dst = S["views"][p]
                dst["recv"][par, rank].copy_(S["buckets"][p], non_blocking=True)
                dst["cnt"][par, rank].copy_(S["bcnt"][p], non_blocking=True)
            S["ev_left"] = torch.cuda.Event()
            S["ev_left"].record(comm_stream)
            S["hdl"].barrier(channel=par)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(comm_stream)
            eng.set_stream(comm_stream.cuda_stream)
            eng.scatter_records(mine["recv"][par], world, S["cap_pair"], mine["cnt"][par], kinds[name], mine["final"])
            eng.set_stream(main_stream.cuda_stream)
            b.record(comm_stream)
            timers["s"].append((a, b))
            lo = rank * S["shard_len"]
            for k in range(1, world):
                p = (rank + k) % world
                S["views"][p]["final"][lo:lo + S["shard_len"]].copy_(mine["final"][lo:lo + S["shard_len"]], non_blocking=True)

    def step_device(record: bool):
        par = step_no[0] & 1
        step_no[0] += 1
        if world > 1 and space_sharded and pipelined:
            if nq_total:
                n_c = timed("c", lambda: eng.closest_records(dq, [sym["c"]["share"]], rank, world, compact=True))
                assert n_c <= sym["c"]["cap_share"], "closest share exceeds its buffer"
                exchange_pipelined("c", n_c, par)
            if nr_total:
                n_r = timed("r", lambda: eng.ray_first_records(dro, drd, [sym["r"]["share"]], rank, world, compact=True))
                assert n_r <= sym["r"]["cap_share"], "ray share exceeds its buffer"
                exchange_pipelined("r", n_r, par)
            return
        if world > 1 and space_sharded:
            evc = evr = None
            if nq_total:
                n_c = timed("c", lambda: eng.closest_records(dq, [sym["c"]["share"]], rank, world, compact=True))
                assert n_c <= sym["c"]["cap_share"], "closest share exceeds its buffer"
                evc = exchange_a("c", n_c, par)                   # the all-to-all of the closest records runs under the ray kernel
            if nr_total:
                n_r = timed("r", lambda: eng.ray_first_records(dro, drd, [sym["r"]["share"]], rank, world, compact=True))
                assert n_r <= sym["r"]["cap_share"], "ray share exceeds its buffer"
                evr = exchange_a("r", n_r, par)
            if evc is not None:
                exchange_b("c", evc, par)
            if evr is not None:
                exchange_b("r", evr, par)
            return
        # one GPU (or the NCCL fallback): records straight to [query index]
        if nq_total:
            if space_sharded:
                timed("c", lambda: eng.closest_records(dq, [crec], 0, 1))
            else:
                timed("c", lambda: eng.closest_records(dq[q_off:q_off + nq], [crec[q_off:q_off + nq]]))
                pending.append(dist.all_gather_into_tensor(crec.view(-1), crec[q_off:q_off + nq].reshape(-1), async_op=True))
        if nr_total:
            if space_sharded:
                timed("r", lambda: eng.ray_first_records(dro, drd, [rrec], 0, 1))
            else:
                timed("r", lambda: eng.ray_first_records(dro[r_off:r_off + nr], drd[r_off:r_off + nr], [rrec[r_off:r_off + nr]]))
                pending.append(dist.all_gather_into_tensor(rrec.view(-1), rrec[r_off:r_off + nr].reshape(-1), async_op=True))
        while pending:
            pending.pop().wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device(False)
    barrier()
    for k in timers:
        timers[k].clear()
    launches0 = eng.stats()["total_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device(True)
    if world > 1 and sym is not None:
        torch.cuda.current_stream().wait_stream(comm_stream)      # this rank's last pushes into the peers' result arrays are done
    e1.record()
    barrier()                                # ... and so are everybody else's into this rank's: every rank holds the whole batch's results
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dev_ms = float(ms.item())
    total_queries = nq_total + nr_total
    value = total_queries * args.steps / (dev_ms / 1e3)
    # per-call device times (select + sort + traversal kernel of the rank's share; scatter kernels)
    kern["closest_ms"] = sum(a.elapsed_time(b) for a, b in timers["c"])
    kern["ray_ms"] = sum(a.elapsed_time(b) for a, b in timers["r"])
    kern["scatter_ms"] = sum(a.elapsed_time(b) for a, b in timers["s"])
    kern["closest_q"], kern["ray_q"] = nq_total / world * len(timers["c"]), nr_total / world * len(timers["r"])
    kern["closest_launches"], kern["ray_launches"] = max(1, len(timers["c"])), max(1, len(timers["r"]))
    kern["launches"] = int(eng.stats()["total_launches"] - launches0)
    eng.set_option(3, 0)

    # checksum of the gathered results on THIS rank's copy (every rank holds all of them), so that the timed work cannot be optimised
    # away and runs at different N can be compared; all ranks must agree
    checksum = int((crec.view(torch.int32)[:, 0].long().sum() + rrec.view(torch.int32)[:, 1].long().sum()).item())
    if world > 1:
        cs = torch.tensor([checksum, -checksum], dtype=torch.int64, device=dev)
        dist.all_reduce(cs, op=dist.ReduceOp.MAX)
        assert int(cs[0].item()) == checksum and int(cs[1].item()) == -checksum, "ranks hold different gathered results"

    # ---- e2e: same calls, pinned host buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        # host-buffer leg: every rank answers a contiguous shard of the batch from / to pinned host memory
        hq = dq[q_off:q_off + nq].cpu().pin_memory().numpy()
        hro, hrd = dro[r_off:r_off + nr].cpu().pin_memory().numpy(), drd[r_off:r_off + nr].cpu().pin_memory().numpy()
        h_idx = torch.empty(nq, dtype=torch.int32).pin_memory().numpy()
        h_d2 = torch.empty(nq, dtype=torch.float32).pin_memory().numpy()
        h_cp = torch.empty((nq, 3), dtype=torch.float32).pin_memory().numpy()
        h_t = torch.empty(nr, dtype=torch.float32).pin_memory().numpy()
        h_ri = torch.empty(nr, dtype=torch.int32).pin_memory().numpy()

        if args.chunk:
            eng.set_option(4, args.chunk)

        eng.set_option(3, 1)          # TGQ_OPT_ASYNC: the two calls of a step overlap (downloads of one under the uploads of the next)

        def step_host():
            if nq:
                eng.closest(hq, out=(h_idx, h_d2, h_cp))
            if nr:
                eng.ray_first(hro, hrd, out=(h_t, h_ri, None))
            eng.synchronize()         # results are in the host buffers

        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        e2e_s = float(wall.item())
        h2d = nq * 12 + nr * 24
        d2h = nq * 20 + nr * 8
        hs = torch.tensor([int(h_idx.astype(np.int64).sum() + h_ri.astype(np.int64).sum())], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(hs)
        e2e = {"value": total_queries * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
               "matches_device_checksum": bool(int(hs.item()) == checksum)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_kind = measured_peak()
    dominant = "k_closest_pool" if kern["closest_ms"] >= kern["ray_ms"] else "k_ray<false>"
    if dominant.startswith("k_closest"):
        bytes_per_q, k_ms, k_q, k_n = algorithmic_bytes(T, "closest"), kern["closest_ms"], kern["closest_q"], kern.get("closest_launches", 1)
    else:
        bytes_per_q, k_ms, k_q, k_n = algorithmic_bytes(T, "ray"), kern["ray_ms"], kern["ray_q"], kern.get("ray_launches", 1)
    achieved = (bytes_per_q * k_q / 1e9) / (k_ms / 1e3) if k_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        rec = json.load(open(tp)).get(f"{args.workload}:{dominant}")
        if rec:   # ncu dram bytes of the bench's own launch, scaled to this run's queries per launch
            traffic = rec["dram_bytes_per_launch"] / rec["queries_per_launch"] * (k_q // max(1, k_n))
    limiter = None
    if os.path.exists(tp):
        limiter = (json.load(open(tp)).get(f"{args.workload}:{dominant}") or {}).get("limiter")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "limiter": limiter,
                "kernel": dominant, "peak_source": peak_kind, "algorithmic_bytes_per_query": bytes_per_q,
                "queries_per_launch": int(k_q // max(1, k_n)), "avg_launch_ms": k_ms / max(1, k_n)}

    # ---- CPU baseline on rank 0: the reference's own code (or the oracle port) on all host cores, bounded sample ----
    cpu = None
    if not args.no_cpu:
        from oracle import orc
        arm, kind, cores, what = cpu_reference(mesh)
        sq, sr = (1000 if nq_total else 0), (100000 if nr_total else 0)
        t0 = time.perf_counter()
        if sq:
            arm.closest(gen.closest_queries(sq, bounds))
        tq = time.perf_counter() - t0
        # scale the closest sample to ~10 s of CPU work
        if sq and tq < 8.0:
            sq = int(min(2_000_000, sq * 10.0 / max(tq, 1e-3)))
            qs = gen.closest_queries(sq, bounds)
            t0 = time.perf_counter()
            oi, od, _ = arm.closest(qs)
            tq = time.perf_counter() - t0
        else:
            qs = gen.closest_queries(sq, bounds) if sq else None
            od = arm.closest(qs)[1] if sq else None
        t0 = time.perf_counter()
        if sr:
            o, d = gen.random_rays(sr, bounds)
            arm.ray_first(o, d)
        tr = time.perf_counter() - t0
        # per-query costs combined in the workload's mix
        per_q = (tq / sq if sq else 0.0) * nq_total + (tr / sr if sr else 0.0) * nr_total
        cpu = {"value": total_queries / per_q if per_q > 0 else None, "unit": "queries/s", "cores": cores, "kind": kind,
               "sample": f"{sq} closest-point ({tq:.2f}s) + {sr} rays ({tr:.2f}s) of the same streams, combined in the workload's "
                         f"{nq_total}:{nr_total} mix; {what}"}
        # the same prefix through the GPU path must agree with the CPU side's distances (sanity, not timing)
        if sq:
            eng.set_option(3, 0)
            gi, gd, _ = eng.closest(qs[:2000])
            assert np.array_equal(gd, od[:2000]), "GPU closest-point distances differ from the CPU reference on the bench sample"

    line = {
        "metric": "bvh_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": desc, "triangles": T, "closest_queries": nq_total, "ray_queries": nr_total, "sharding": f"BVH replicated; batch replicated, each rank answers the queries of its Morton cells (1/{world})" + (f"; results gathered on every rank every step: {transport}" if world > 1 else ""),
                   "l2_policy": "inputs_exceed_l2" if (nq_total * 32 + nr_total * 32) / world > 126e6 else "structure+inputs may be L2 resident",
                   "scale": args.scale, "build_ms": build_ms, "build_first_ms": build_first_ms},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": kern["launches"], "clocks": clocks,
        "breakdown": {"closest_Mqps_kernel": kern["closest_q"] / kern["closest_ms"] / 1e3 if kern["closest_ms"] else None,
                      "ray_Mqps_kernel": kern["ray_q"] / kern["ray_ms"] / 1e3 if kern["ray_ms"] else None,
                      "per_call": "device time of one C-ABI call on this rank's share: query selection + Morton sort + traversal kernel",
                      "scatter_ms_per_step": kern["scatter_ms"] / args.steps, "checksum": checksum},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
