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


class ShardedIndexer:
    """Indexer over this rank's shard + cross-rank merge.  Same query API as Indexer."""

    def __init__(self, hashing, local_vectors_gpu, distance_func, shard_lo, metric=None, group=None):
        self.group = group
        self.shard_lo = int(shard_lo)
        self.local = Indexer(hashing, local_vectors_gpu, distance_func, metric=metric,
                             id_offset=self.shard_lo)

    def query_tensors(self, query_vectors, k=10, hash_times=10, probes=None):
        ids, dists, ncand = self.local.query_tensors(query_vectors, k, hash_times, probes)
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return ids, dists, ncand
        g_ids, g_d = all_gather_topk(ids, dists, self.group)
        ncand = ncand.clone()
        dist.all_reduce(ncand, group=self.group)
        m_ids, m_d = _native.merge_topk(g_d, g_ids)
        return m_ids, m_d, ncand

    def capture_query(self, n_queries, k=10, hash_times=10):
        """The local search (hash -> scan + top-k) as one CUDA graph; the all-gather and the
        merge of the shard lists stay eager launches behind it.  Returns a callable
        query_vectors -> (ids, dists, n_candidates); `.kernels_per_call` counts this library's
        kernels per call."""
        graphed = self.local.capture_query(n_queries, k, hash_times)
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1

        def run(query_vectors):
            ids, dists, ncand = graphed(query_vectors)
            if not multi:
                return ids, dists, ncand
            g_ids, g_d = all_gather_topk(ids, dists, self.group)
            total = ncand.clone()
            dist.all_reduce(total, group=self.group)
            m_ids, m_d = _native.merge_topk(g_d, g_ids)
            return m_ids, m_d, total

        run.kernels_per_call = graphed.kernels_per_replay + (1 if multi else 0)
        run.graphed = graphed
        return run

    def query(self, query_vectors, k=10, hash_times=10, probes=None):
        ids, _, ncand = self.query_tensors(query_vectors, k, hash_times, probes)
        rows = ids.cpu().tolist()
        return [[v for v in r if v >= 0] for r in rows], ncand.cpu().tolist()
