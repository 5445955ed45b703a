"""Index build + query — drop-in mirror of the reference's nlsh/indexer.py.

`build_index(indexes, cuda=True)` and `Indexer(hashing, candidate_vectors_gpu,
distance_func)` with `.index2row`, `.hash(...)`, `.query(...)` keep the reference's
signatures and return types (indexer.py:6-96) so nlsh/trainers/base.py:80-115 and eval.py
run unchanged.  Underneath, the Python dict-of-lists, the per-query index_select /
distance / topk loop and the per-query `.tolist()` sync are replaced by three C-ABI calls
into libnlsh_b200.so: hasher forward -> CSR build -> batched scan + top-k.
"""
from collections.abc import Mapping
from typing import List

import numpy as np
import torch
import torch.nn.functional as F

from . import _native
from .hashings import codes_to_sets


# --------------------------------------------------------------------------------------
# metric resolution: the reference passes a Python callable (data.distance)
# --------------------------------------------------------------------------------------
def _probe_metric(distance_func):
    """Identify a distance callable by behaviour on a tiny CPU fixture: it must be either
    nlsh/data.py:201 (F.pairwise_distance, eps inside) or nlsh/data.py:109 (1 - cosine)."""
    v1 = torch.tensor([0.5, -1.0, 2.0, 0.25])
    v2 = torch.tensor([[0.5, -1.0, 2.0, 0.25], [1.0, 0.0, -1.0, 3.0], [-2.0, 0.5, 0.5, 1.0]])
    try:
        with torch.no_grad():
            got = torch.as_tensor(distance_func(v1, v2)).float().flatten()
    except Exception:  # noqa: BLE001 - any failure means "not a metric we know"
        return None
    if got.shape != (3,):
        return None
    if torch.allclose(got, F.pairwise_distance(v1, v2), rtol=1e-6, atol=1e-9):
        return "l2"
    if torch.allclose(got, 1 - F.cosine_similarity(v1, v2, dim=-1), rtol=1e-6, atol=1e-7):
        return "angular"
    return None


def resolve_metric(distance_func, metric=None) -> int:
    """Map the reference's `distance_func` argument (or an explicit metric name) to the scan
    kernel's metric id; anything that is not the L2 / angular scan distance is refused —
    there is no generic-callable fallback."""
    if metric is not None:
        name = str(metric).lower()
        if name in ("l2", "euclidean"):
            return _native.METRIC_L2
        if name in ("angular", "cosine"):
            return _native.METRIC_ANGULAR
        raise ValueError(f"unknown metric {metric!r} (expected 'l2' or 'angular')")
    if isinstance(distance_func, str):
        return resolve_metric(None, distance_func)
    tagged = getattr(distance_func, "nlsh_metric", None)
    if tagged is not None:
        return resolve_metric(None, tagged)
    if distance_func is F.pairwise_distance:
        return _native.METRIC_L2
    qual = getattr(distance_func, "__qualname__", "")
    if qual.startswith("SIFT."):
        return _native.METRIC_L2
    if qual.startswith("Glove."):
        return _native.METRIC_ANGULAR
    found = _probe_metric(distance_func) if callable(distance_func) else None
    if found is None:
        raise ValueError(
            "distance_func is neither the L2 scan distance (nlsh/data.py:201, "
            "F.pairwise_distance) nor the angular one (nlsh/data.py:109); pass metric='l2' or "
            "metric='angular'. The B200 scan kernel has no generic-callable fallback.")
    return resolve_metric(None, found)


# --------------------------------------------------------------------------------------
# build_index (stand-alone function, indexer.py:6-24)
# --------------------------------------------------------------------------------------
def build_index(indexes, cuda=True):
    """indexes: per row a set of codes (or a single int) -> {code: LongTensor of
    row ids, ascending}.  Keys appear in order of first occurrence, as the reference's dict.
    The grouping runs on the GPU (stable counting sort, nlsh_build_csr); `cuda=False` only
    moves the resulting tensors to the host, as the reference's flag does."""
    if not torch.cuda.is_available():
        raise _native.NativeLibraryError("build_index needs a CUDA device (no CPU fallback)")
    rows, codes = [], []
    for idx, index_set in enumerate(indexes):
        if isinstance(index_set, (int, np.integer)):
            rows.append(idx)
            codes.append(int(index_set))
        else:
            for index in index_set:
                rows.append(idx)
                codes.append(int(index))
    if not codes:
        return {}
    first_seen = list(dict.fromkeys(codes))
    codes_np = np.asarray(codes, dtype=np.int64)
    lo = int(codes_np.min())
    n_buckets = int(codes_np.max()) - lo + 1
    if n_buckets > (1 << 20):
        raise ValueError(f"code range {n_buckets} too wide for a dense bucket table")
    dev = torch.device("cuda")
    dense = torch.from_numpy((codes_np - lo).astype(np.int32)).to(dev)
    rows_t = torch.from_numpy(np.asarray(rows, dtype=np.int64)).to(dev)
    offsets, ids, _ = _native.build_csr(dense, n_buckets, None)
    sorted_rows = rows_t[ids.long()]  # pair order is row-major, so rows ascend inside a bucket
    off = offsets.cpu().tolist()
    if not cuda:
        sorted_rows = sorted_rows.cpu()
    return {code: sorted_rows[off[code - lo]:off[code - lo + 1]].clone() for code in first_seen}


class BucketView(Mapping):
    """Read-only dict-like view of the CSR index: code -> LongTensor of row ids (ascending).

    Satisfies what the reference's callers do with `index2row`: len(), .values() of things
    with len() (base.py:87-90), .get(key, default) (indexer.py:68), .items()/.keys()
    (eval.py:121-124).  Only non-empty buckets are keys, as in the reference's dict."""

    def __init__(self, offsets_host, ids_device, id_offset=0):
        self._off = offsets_host
        self._ids = ids_device
        self._id_offset = id_offset
        self._sizes = np.diff(offsets_host)
        self._keys = np.nonzero(self._sizes > 0)[0]

    def __len__(self):
        return int(self._keys.shape[0])

    def __iter__(self):
        return iter(self._keys.tolist())

    def __contains__(self, key):
        return isinstance(key, (int, np.integer)) and 0 <= key < self._sizes.shape[0] \
            and self._sizes[key] > 0

    def __getitem__(self, key):
        if key not in self:
            raise KeyError(key)
        rows = self._ids[int(self._off[key]):int(self._off[key + 1])].long()
        return rows + self._id_offset if self._id_offset else rows

    @property
    def sizes(self):
        """Bucket sizes of the non-empty buckets (np.int64), no device traffic."""
        return self._sizes[self._keys]


class Indexer:

    def __init__(self, hashing, candidate_vectors_gpu, distance_func, metric=None, id_offset=0):
        self._hashing = hashing
        self._candidate_vectors_gpu = candidate_vectors_gpu
        self._distance_func = distance_func
        self._metric = resolve_metric(distance_func, metric)
        self._id_offset = int(id_offset)
        # indexer.py:52 hashes the tail batch with hash_times=1; set True to
        # reproduce that in differential tests, default applies hash_times to every query.
        self.compat_tail_single_probe = False
        self.scan_flags = 0
        self._build_index()

    # ---- build ---------------------------------------------------------------------------
    def _build_index(self):
        x = self._candidate_vectors_gpu
        _native.require_cuda(x, "candidate_vectors_gpu")
        n, d = x.shape
        self._dim = d
        codes, _, _ = self._hashing.hash_tensors(x, 1)
        n_buckets = self._hashing.n_buckets
        self._offsets, self._ids, self._x_sorted, self._x_sqnorm = _native.build_csr(
            codes, n_buckets, x, want_sqnorm=True)
        off = self._offsets.cpu().numpy().astype(np.int64)  # one sync per build
        if int(off[-1]) != n:
            raise ValueError(f"index build dropped {n - int(off[-1])} rows: bucket codes outside "
                             f"[0, {n_buckets})")
        self._offsets_host = off
        sizes = np.diff(off)
        self._max_bucket_rows = int(sizes.max()) if sizes.size else 0
        self.index2row = BucketView(off, self._ids, self._id_offset)

    @property
    def bucket_sizes(self):
        return np.diff(self._offsets_host)

    # ---- hash ----------------------------------------------------------------------------
    def hash_tensors(self, query_vectors, hash_times=1, workspace=None):
        """-> probes int32 [n, hash_times] on the device (column 0 = hard code)."""
        if workspace is None:
            codes, probes, _ = self._hashing.hash_tensors(query_vectors, hash_times)
        else:  # a hashing object from elsewhere (e.g. the reference's own class) need not know the keyword
            codes, probes, _ = self._hashing.hash_tensors(query_vectors, hash_times, workspace=workspace)
        if probes is None:
            probes = codes.unsqueeze(1)
        if self.compat_tail_single_probe and hash_times > 1:
            n_full = (query_vectors.shape[0] // 4096) * 4096
            probes = probes.clone()
            probes[n_full:, 1:] = -1
        return probes

    def hash(self, query_vectors, batch_size=4096, hash_times=1):
        # indexer.py:40-54; batch_size is accepted for compatibility (the CUDA path chunks
        # internally)
        if query_vectors.shape[0] == 0:
            return []
        return codes_to_sets(self.hash_tensors(query_vectors, hash_times))

    # ---- query ---------------------------------------------------------------------------
    def seed_tau_tensors(self, query_vectors, probes, k=10, workspace=None, sample_rows=0):
        """Distance bounds fp32 [Q] of the queries from a sample of this index's rows (see
        _native.query_seed_tau): upper bounds of the k-th best distance over ANY index that holds these rows,
        so a row-sharded search seeds each query on one rank only."""
        return _native.query_seed_tau(query_vectors, probes, self._offsets, self._x_sorted, self._dim,
                                      self._metric, k, workspace=workspace, sample_rows=sample_rows)

    def uses_tensor_core_scan(self, n_queries, k, hash_times):
        return _native.scan_impl(self._dim, k, self._metric, self._x_sqnorm is not None, n_queries, hash_times,
                                 self._hashing.n_buckets) == 1 and (self.scan_flags & 2) == 0

    def query_tensors(self, query_vectors, k=10, hash_times=10, probes=None, out=None, workspace=None,
                      tau_seed=None):
        """Batched search -> (ids int64 [Q, k], dists fp32 [Q, k], n_candidates int32 [Q]),
        all on the device, no host synchronisation.  ids are -1 / dists +inf past the number
        of candidates.  `probes` (int32 [Q, p], -1 = unused) overrides the hasher's probe
        sets, e.g. with the reference's sampled sets for differential testing.  `workspace`: a
        _native.Workspace for the calls' scratch (default: the one of the current stream)."""
        _native.require_cuda(query_vectors, "query_vectors")
        if probes is None:
            probes = self.hash_tensors(query_vectors, hash_times, workspace=workspace)
        return _native.query_scan_topk(
            query_vectors, probes, self._offsets, self._ids, self._x_sorted, self._dim,
            self._max_bucket_rows, self._metric, k, id_offset=self._id_offset,
            flags=self.scan_flags, out=out, x_sqnorm=self._x_sqnorm, workspace=workspace, tau_seed=tau_seed)

    def query(self, query_vectors, k=10, hash_times=10, probes=None) -> List[List[int]]:
        # indexer.py:56-96: returns (List[List[int]] ids by ascending distance, List[int]
        # n_candidates).  Rows with fewer than k candidates return all of them, sorted
        # (the reference returns the last probed bucket unsorted there).
        if query_vectors.shape[0] == 0:
            return [], []
        ids, _, ncand = self.query_tensors(query_vectors, k, hash_times, probes)
        ids_host = ids.cpu().numpy()
        ncand_host = ncand.cpu().numpy()
        short = ncand_host < k
        result = ids_host.tolist()
        if short.any():
            for i in np.nonzero(short)[0].tolist():
                result[i] = [v for v in result[i] if v >= 0]
        return result, ncand_host.tolist()

    def capture_query(self, n_queries, k=10, hash_times=10, out=None):
        """CUDA-graph capture of query_tensors for a fixed batch shape: returns a GraphedQuery
        whose call replays hash -> probe selection -> scan + top-k (about a dozen launches) as
        one graph launch."""
        return GraphedQuery(self, n_queries, k, hash_times, out=out)

    @staticmethod
    def probes_from_sets(sets, device, width=None):
        """List[Set[int]] (e.g. the reference's hash() output) -> int32 [n, width] probes."""
        width = width or max((len(s) for s in sets), default=1) or 1
        arr = np.full((len(sets), width), -1, dtype=np.int32)
        for i, s in enumerate(sets):
            vals = list(s)
            arr[i, :len(vals)] = vals
        return torch.from_numpy(arr).to(device)


class GraphedQuery:
    """`Indexer.query_tensors` for a fixed (n_queries, k, hash_times), captured once into a CUDA
    graph and replayed per batch: no per-launch host cost, no allocations.  The graph reads the
    hasher's live parameter tensors and the index arrays in place; build a new one after the
    index is rebuilt.  Call with a CUDA (or pinned host) tensor [n_queries, d]; returns the
    graph's static output tensors (ids int64 [Q, k], dists fp32 [Q, k], n_candidates int32 [Q]),
    overwritten by the next call.

    The graph owns its scratch (a private _native.Workspace): the captured kernels hold raw
    pointers into it, so no other call can write there, and several GraphedQuery objects of one
    index can replay concurrently on different streams (nlsh.parallel.PipelinedSearch does)."""

    def __init__(self, indexer, n_queries, k=10, hash_times=10, out=None):
        dev = indexer._candidate_vectors_gpu.device
        self.indexer = indexer
        self.q = torch.zeros((n_queries, indexer._dim), dtype=torch.float32, device=dev)
        self.workspace = _native.Workspace(dev, name=f"GraphedQuery {id(self):#x}")
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up on the capture stream (sizes the workspace)
            for _ in range(2):
                indexer.query_tensors(self.q, k, hash_times, out=out, workspace=self.workspace)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        launches0 = _native.kernel_launch_count()
        with torch.cuda.graph(self.graph, stream=side):
            self.ids, self.dists, self.ncand = indexer.query_tensors(self.q, k, hash_times, out=out,
                                                                      workspace=self.workspace)
        self.kernels_per_replay = _native.kernel_launch_count() - launches0
        self.workspace.frozen = True

    def replay(self):
        """Replays on the current stream with whatever self.q holds."""
        self.graph.replay()
        return self.ids, self.dists, self.ncand

    def __call__(self, query_vectors):
        self.q.copy_(query_vectors, non_blocking=True)
        self.graph.replay()
        return self.ids, self.dists, self.ncand
