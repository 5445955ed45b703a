"""Row-sharded multi-GPU search (new capability; the reference is single-GPU).

One process per GPU (torchrun).  Rank g owns database rows [lo_g, hi_g), runs the replicated
hasher on them and builds its own CSR shard over all buckets; queries are replicated, every
rank scans its shard, and the per-shard top-k lists meet in ONE all-gather (NCCL over
NVLink on GPUs, gloo in the CPU tests) followed by the k-way merge kernel.  Global ids are
local ids + lo_g, and every top-k decision uses the (distance, id) order, so the merged
result equals the single-GPU result independent of the shard count.
"""
import torch
import torch.distributed as dist

from . import _native
from .indexer import Indexer


def shard_range(n_rows, rank, world_size):
    """Contiguous, balanced row range [lo, hi) of `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_topk(ids, dists, group=None):
    """[Q, k] per-rank lists -> ([G, Q, k] ids, [G, Q, k] dists), rank-major, on every rank."""
    world = dist.get_world_size(group)
    ids = ids.contiguous()
    dists = dists.contiguous()
    nq, k = ids.shape
    # flat [G*Q, k] output (concatenation along dim 0): the layout both NCCL and gloo accept
    g_ids = torch.empty((world * nq, k), dtype=ids.dtype, device=ids.device)
    g_d = torch.empty((world * nq, k), dtype=dists.dtype, device=dists.device)
    dist.all_gather_into_tensor(g_ids, ids, group=group)
    dist.all_gather_into_tensor(g_d, dists, group=group)
    return g_ids.view(world, nq, k), g_d.view(world, nq, k)


class PackedLists:
    """One rank's search result as ONE byte buffer [ids int64 [Q,k] | dists fp32 [Q,k] | n_cand
    int32 [Q]] so the exchange is a single all-gather; `gathered` holds all ranks' buffers and
    exposes them as strided [G, Q, k] views the merge kernel reads in place."""

    def __init__(self, n_queries, k, world, device):
        self.nq, self.k, self.world = n_queries, k, world
        self.off_d = n_queries * k * 8
        self.off_n = self.off_d + n_queries * k * 4
        self.nbytes = (self.off_n + n_queries * 4 + 15) // 16 * 16
        self.local = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.gathered = torch.empty(world * self.nbytes, dtype=torch.uint8, device=device)
        self.ids, self.dists, self.ncand = self._views(self.local.view(1, self.nbytes))
        self.ids, self.dists, self.ncand = self.ids[0], self.dists[0], self.ncand[0]
        self.g_ids, self.g_dists, self.g_ncand = self._views(self.gathered.view(world, self.nbytes))

    def _views(self, buf):
        g = buf.shape[0]
        ids = buf[:, :self.off_d].view(torch.int64).view(g, self.nq, self.k)
        dists = buf[:, self.off_d:self.off_n].view(torch.float32).view(g, self.nq, self.k)
        ncand = buf[:, self.off_n:self.off_n + self.nq * 4].view(torch.int32).view(g, self.nq)
        return ids, dists, ncand

    def out(self):
        return self.ids, self.dists, self.ncand

    def exchange_and_merge(self, group=None, sqrt_after=False):
        """ONE all-gather of the packed lists + the shard merge; sqrt_after: the lists hold squared L2
        distances (FLAG_SQUARED_L2_OUT) and the root is taken on the merged result."""
        dist.all_gather_into_tensor(self.gathered, self.local, group=group)
        ids, dists, ncand = _native.merge_topk(self.g_dists, self.g_ids, self.g_ncand)
        if sqrt_after:
            dists.sqrt_()
        return ids, dists, ncand


class ShardedIndexer:
    """Indexer over this rank's shard + cross-rank merge.  Same query API as Indexer."""

    def __init__(self, hashing, local_vectors_gpu, distance_func, shard_lo, metric=None, group=None):
        self.group = group
        self.shard_lo = int(shard_lo)
        self.local = Indexer(hashing, local_vectors_gpu, distance_func, metric=metric,
                             id_offset=self.shard_lo)
        self._packed = {}  # (n_queries, k, stream) -> PackedLists of the eager path
        self._squared = self.local._metric == _native.METRIC_L2  # shard lists carry squared L2 distances

    def _multi(self):
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def query_tensors(self, query_vectors, k=10, hash_times=10, probes=None):
        if not self._multi():
            return self.local.query_tensors(query_vectors, k, hash_times, probes)
        dev = query_vectors.device
        key = (query_vectors.shape[0], k, torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0)
        packed = self._packed.get(key)
        if packed is None:  # exchange buffers are kept per batch shape (and stream), not allocated per call
            packed = self._packed[key] = PackedLists(key[0], k, dist.get_world_size(self.group), dev)
        with self._shard_list_flags():
            self.local.query_tensors(query_vectors, k, hash_times, probes, out=packed.out())
        return packed.exchange_and_merge(self.group, sqrt_after=self._squared)

    def _shard_list_flags(self):
        """Context: this shard's lists are written for a cross-shard merge (squared L2 distances)."""
        local, squared = self.local, self._squared

        class _Ctx:
            def __enter__(self_inner):
                self_inner.saved = local.scan_flags
                if squared:
                    local.scan_flags |= _native.FLAG_SQUARED_L2_OUT

            def __exit__(self_inner, *exc):
                local.scan_flags = self_inner.saved
        return _Ctx()

    def capture_query(self, n_queries, k=10, hash_times=10, shard_hashing=True, group=None):
        """The whole batch as ONE CUDA graph.  One rank: hash -> probe selection -> scan + top-k.  Several ranks:
        each rank hashes its 1/N slice of the queries, the probe matrices meet in a first all-gather (Q x p int32),
        every rank scans its shard, the packed per-shard lists meet in the second all-gather, then the shard merge -
        NCCL collectives captured in the graph, so a replay costs one launch and no host work between the kernels.
        (`shard_hashing=False`: every rank hashes all queries, one all-gather.  `group`: the process group of the
        captured collectives - graphs that replay concurrently need a communicator each.)  Returns a callable
        query_vectors -> (ids, dists, n_candidates); `.kernels_per_call` counts this library's kernels per call."""
        if not self._multi():
            graphed = self.local.capture_query(n_queries, k, hash_times)

            def run(query_vectors):
                return graphed(query_vectors)

            run.kernels_per_call = graphed.kernels_per_replay
            run.graphed = graphed
            return run
        graphed = GraphedShardedQuery(self, n_queries, k, hash_times, shard_hashing, group=group)

        def run(query_vectors):
            return graphed(query_vectors)

        run.kernels_per_call = graphed.kernels_per_replay
        run.graphed = graphed
        return run

    def query(self, query_vectors, k=10, hash_times=10, probes=None):
        ids, _, ncand = self.query_tensors(query_vectors, k, hash_times, probes)
        rows = ids.cpu().tolist()
        return [[v for v in r if v >= 0] for r in rows], ncand.cpu().tolist()


class GraphedShardedQuery:
    """One batch of a row-sharded search as a single CUDA graph (see ShardedIndexer.capture_query).  The
    graph owns its scratch workspace, its padded query buffer, the probe and result exchange buffers and the
    output tensors; the NCCL all-gathers are graph nodes (captured with capture_error_mode="thread_local":
    the process group's watchdog thread may touch CUDA while this thread captures)."""

    def __init__(self, sharded, n_queries, k, hash_times, shard_hashing=True, group=None):
        local = sharded.local
        dev = local._candidate_vectors_gpu.device
        group = group if group is not None else sharded.group
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        self.n_queries = n_queries
        self.chunk = (n_queries + world - 1) // world if shard_hashing else n_queries
        rows = self.chunk * world if shard_hashing else n_queries
        self.q = torch.zeros((rows, local._dim), dtype=torch.float32, device=dev)
        self.workspace = _native.Workspace(dev, name=f"GraphedShardedQuery {id(self):#x}")
        self.packed = PackedLists(n_queries, k, world, dev)
        # first exchange: per query its probe row and, in one more int32 column, the bits of its distance bound
        # (each rank seeds the bounds of its slice from its own shard: valid for every shard)
        seeded = shard_hashing and local.uses_tensor_core_scan(n_queries, k, hash_times)
        # a rank seeds only 1/world of the queries, so from 4 ranks on it takes a sample 1.5x the library's rule
        # (on an 8-GPU shard of config 4: 192 rows against 128, scan 0.176 -> 0.154 ms, seed + 0.007 ms)
        n_rows, n_buckets = local._x_sorted.shape[0], local._offsets.shape[0] - 1
        base_rows = min(128, max(32, n_rows // max(n_buckets, 1) // 2 // 32 * 32))
        seed_rows = base_rows * 3 // 2 if world >= 4 else 0
        width = hash_times + (1 if seeded else 0)
        self.slice_out = torch.empty((self.chunk, width), dtype=torch.int32, device=dev)
        self.slices_all = torch.empty((rows, width), dtype=torch.int32, device=dev)

        def batch():
            if shard_hashing:
                mine = self.q[rank * self.chunk:(rank + 1) * self.chunk]
                probes_mine = local.hash_tensors(mine, hash_times, workspace=self.workspace)
                self.slice_out[:, :hash_times] = probes_mine
                if seeded:
                    tau_mine = local.seed_tau_tensors(mine, probes_mine, k, workspace=self.workspace,
                                                      sample_rows=seed_rows)
                    self.slice_out[:, hash_times] = tau_mine.view(torch.int32)
                dist.all_gather_into_tensor(self.slices_all, self.slice_out, group=group)
                probes = self.slices_all[:n_queries, :hash_times].contiguous()
                tau = self.slices_all[:n_queries, hash_times].contiguous().view(torch.float32) if seeded else None
            else:
                probes = local.hash_tensors(self.q, hash_times, workspace=self.workspace)
                tau = None
            with sharded._shard_list_flags():
                local.query_tensors(self.q[:n_queries], k, hash_times, probes=probes, out=self.packed.out(),
                                    workspace=self.workspace, tau_seed=tau)
            return self.packed.exchange_and_merge(group, sqrt_after=sharded._squared)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up on the capture stream: sizes the workspace, warms the communicator
            for _ in range(2):
                batch()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        launches0 = _native.kernel_launch_count()
        with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
            self.ids, self.dists, self.ncand = batch()
        self.kernels_per_replay = _native.kernel_launch_count() - launches0
        self.workspace.frozen = True

    def replay(self):
        self.graph.replay()
        return self.ids, self.dists, self.ncand

    def __call__(self, query_vectors):
        self.q[:self.n_queries].copy_(query_vectors, non_blocking=True)
        self.graph.replay()
        return self.ids, self.dists, self.ncand

    def release(self):
        """Drop the captured graph (before the process group is destroyed: a live graph holds NCCL work)."""
        self.graph = None


class PipelinedSearch:
    """Serving loop over captured query graphs (ShardedIndexer.capture_query) for a fixed batch
    shape, `depth` batches in flight, each in its own lane: a CUDA stream with its own captured
    graph, scratch workspace, input buffer and (pinned) result buffers.  A lane runs the whole
    life of a batch - input copy, hash -> probe selection -> plan -> seed -> scan -> merge,
    (all-gather + shard merge), result copy - and lanes overlap on the device, so the launch-bound
    front part of batch i+1 and the exchange of batch i run beside the scan of the other lane.
    Batches complete in submission order per lane; results are identical to the serial path.

        pipe = PipelinedSearch(index, n_queries, k=10, hash_times=8)
        t = pipe.submit(q)                     # q: pinned host or device tensor; does not block
        ids, dists, ncand = pipe.result(t)     # blocks until that batch is done; pinned host buffers
                                               # (to_host=True) or the lane's device tensors, both
                                               # reused `depth` submits later
    """

    def __init__(self, sharded_index, n_queries, k=10, hash_times=10, depth=2, to_host=True):
        local = sharded_index.local
        dev = local._candidate_vectors_gpu.device
        self.device = dev
        self.depth = int(depth)
        self.to_host = bool(to_host)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(self.depth)]
        self.runs = []
        multi = sharded_index._multi()
        # lanes replay concurrently: NCCL serialises nothing between the captured collectives of different
        # graphs, so every lane gets a communicator of its own (all ranks create them in the same order)
        self.groups = [dist.new_group() if multi and self.depth > 1 else sharded_index.group
                       for _ in range(self.depth)]
        for st, grp in zip(self.streams, self.groups):
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st):
                self.runs.append(sharded_index.capture_query(n_queries, k=k, hash_times=hash_times, group=grp))
        torch.cuda.synchronize(dev)
        self.out = [(torch.empty((n_queries, k), dtype=torch.int64).pin_memory(),
                     torch.empty((n_queries, k), dtype=torch.float32).pin_memory(),
                     torch.empty((n_queries,), dtype=torch.int32).pin_memory()) if self.to_host else None
                    for _ in range(self.depth)]
        self.dev_out = [None] * self.depth
        self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self.submitted = 0
        self.h2d_bytes = n_queries * local._dim * 4
        self.d2h_bytes = n_queries * k * 12 + n_queries * 4
        self.kernels_per_call = getattr(self.runs[0], "kernels_per_call", None)

    def submit(self, q):
        s = self.submitted % self.depth
        st = self.streams[s]
        st.wait_stream(torch.cuda.current_stream(self.device))  # q may have been produced there
        with torch.cuda.stream(st):
            ids, dists, ncand = self.runs[s](q)  # copies q into the lane's input buffer, replays
            self.dev_out[s] = (ids, dists, ncand)
            if self.to_host:
                o_ids, o_d, o_n = self.out[s]
                o_ids.copy_(ids, non_blocking=True)
                o_d.copy_(dists, non_blocking=True)
                o_n.copy_(ncand, non_blocking=True)
            self.done[s].record(st)
        self.submitted += 1
        return s

    def result(self, ticket):
        self.done[ticket].synchronize()
        return self.out[ticket] if self.to_host else self.dev_out[ticket]

    def release(self):
        """Drop the lanes' captured graphs (before the process group is destroyed)."""
        torch.cuda.synchronize(self.device)
        for run in self.runs:
            g = getattr(run, "graphed", None)
            if hasattr(g, "release"):
                g.release()
        self.runs = []

    def fence(self):
        """Orders the current stream after everything submitted so far (no host wait)."""
        cur = torch.cuda.current_stream(self.device)
        for ev in self.done[:min(self.submitted, self.depth)]:
            cur.wait_event(ev)
