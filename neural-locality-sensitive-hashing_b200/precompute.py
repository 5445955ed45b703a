"""Brute-force ground-truth kNN — drop-in mirror of the reference's precompute.py.

`self_get_knn_pt(vectors, distance_func, k=100, batch_size=512)` keeps the reference
signature and return type (precompute.py:57-67: np int array [n, k], the first hit of the
top-(k+1) dropped as "self"), but the 512 x N distance matrix + torch.topk per batch is
replaced by one fused CUDA pass (nlsh_knn_bruteforce) that never materialises distances.
`_l2` / `_cosine_distance` are kept as the selectors the reference's DISTANCE_FUNC table
uses (precompute.py:70-76) and as plain matrix functions.  The script part of the reference
(precompute.py:79-100) reads / writes HDF5 (no h5py here); `python precompute.py <directory> <key>` does
the same job over a directory of .npy / TEXMEX files (nlsh/data_io.py, same key names).
"""
import numpy as np
import torch

from nlsh import _native

QUERY_CHUNK = 131072  # queries per kernel call (bounds the partial-list workspace)


def _cosine_distance(v1, v2):
    """Cosine distance between 2 matrices (precompute.py:22-34): v1 (n, d), v2 (m, d) -> (n, m)."""
    v1_normalized = v1 / v1.norm(dim=1)[:, None]
    v2_normalized = v2 / v2.norm(dim=1)[:, None]
    return 1 - torch.mm(v1_normalized, v2_normalized.T)


def _l2(v1, v2):
    """Squared l2 distance between 2 matrices (precompute.py:37-54): (n, d), (m, d) -> (n, m)."""
    v1_norm = v1.pow(2).sum(dim=-1, keepdim=True)
    v2_norm = v2.pow(2).sum(dim=-1, keepdim=True)
    return torch.addmm(v2_norm.transpose(-2, -1), v1, v2.transpose(-2, -1), alpha=-2).add_(v1_norm)


_l2.nlsh_metric = "l2sq"
_cosine_distance.nlsh_metric = "cosine"


def _resolve_knn_metric(distance_func):
    if isinstance(distance_func, str):
        name = distance_func.lower()
    else:
        name = getattr(distance_func, "nlsh_metric", None) or getattr(distance_func, "__name__", "")
        name = {"_l2": "l2sq", "_cosine_distance": "cosine"}.get(name, name)
    if name not in _native.METRIC_BY_NAME:
        raise ValueError(
            f"distance_func {distance_func!r} is not precompute._l2 / precompute._cosine_distance "
            "(or a metric name 'l2sq' / 'cosine' / 'l2' / 'angular'); the CUDA kNN has no "
            "generic-callable fallback")
    return _native.METRIC_BY_NAME[name]


def knn_tensors(queries, database, distance_func, k, exclude_self=False, self_offset=0,
                id_offset=0):
    """Exact kNN of CUDA `queries` [Q, d] in CUDA `database` [N, d]
    -> (ids int64 [Q, k], dists fp32 [Q, k]) on the device."""
    metric = _resolve_knn_metric(distance_func)
    ids_all, d_all = [], []
    for s in range(0, max(queries.shape[0], 1), QUERY_CHUNK):
        q = queries[s:s + QUERY_CHUNK]
        ids, dist = _native.knn_bruteforce(q, database, metric, k, exclude_self=exclude_self,
                                           self_offset=self_offset + s, id_offset=id_offset)
        ids_all.append(ids)
        d_all.append(dist)
    return torch.cat(ids_all), torch.cat(d_all)


def self_get_knn_pt(vectors, distance_func, k=100, batch_size=512, exclude="first"):
    """precompute.py:57-67.  exclude="first" reproduces the reference (top-(k+1), first hit
    dropped); exclude="self" removes row i from its own result by index instead, which is
    what the reference means and differs only when the data holds exact duplicates."""
    if not torch.cuda.is_available():
        raise _native.NativeLibraryError("self_get_knn_pt needs a CUDA device (no CPU fallback)")
    vector_pt = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32)).cuda()
    n = vectors.shape[0]
    if exclude == "first":
        ids, _ = knn_tensors(vector_pt, vector_pt, distance_func, k + 1)
        ids = ids[:, 1:]
    elif exclude == "self":
        ids, _ = knn_tensors(vector_pt, vector_pt, distance_func, k, exclude_self=True)
    else:
        raise ValueError(f"exclude={exclude!r} (expected 'first' or 'self')")
    knn = np.zeros((n, k), dtype=int)
    knn[:, :] = ids.cpu().numpy()
    return knn


DISTANCE_FUNC = {
    "glove_25": _cosine_distance,
    "glove_50": _cosine_distance,
    "glove_100": _cosine_distance,
    "glove_200": _cosine_distance,
    "sift": _l2,
}


def nearest_exclude_positive(vectors, distance_function, positive_indexes):
    """Negative mining of the triplet trainer (nlsh/trainers/triplet.py:44-74): for every row the
    index of the nearest row that is neither the row itself nor one of its `positive_indexes`
    [n, kp].  The reference materialises a 32 x n distance matrix per batch and masks it; here one
    fused kNN call asks for kp + 2 neighbours (self + kp positives + one more can never all be
    excluded) and the first admissible one is picked.  `distance_function` is a pairwise-matrix
    function of the reference (`_l2` / `_cosine_distance` / `Glove.pairwise_distance`) or a metric
    name.  Returns a LongTensor [n] on the device."""
    if not vectors.is_cuda:
        raise _native.NativeLibraryError("nearest_exclude_positive needs CUDA tensors (no CPU fallback)")
    n = vectors.shape[0]
    positive_indexes = positive_indexes.to(vectors.device).long().reshape(n, -1)
    kp = positive_indexes.shape[1]
    kk = min(kp + 2, n)
    if kk > _native.MAX_K:
        raise ValueError(f"nearest_exclude_positive: {kp} positives per row exceed the top-k capacity")
    ids, _ = knn_tensors(vectors, vectors, distance_function, kk)
    own = torch.arange(n, device=vectors.device)[:, None]
    excluded = (ids == own) | (ids[:, :, None] == positive_indexes[:, None, :]).any(-1) | (ids < 0)
    first = (~excluded).float().argmax(dim=1)  # first admissible neighbour (ids are sorted by distance)
    out = ids.gather(1, first[:, None]).squeeze(1)
    if excluded.all(dim=1).any():
        raise ValueError("nearest_exclude_positive: a row has no admissible neighbour")
    return out


if __name__ == "__main__":
    # precompute.py:79-100 over nlsh/data_io.py: <directory> holds train / test / neighbors as .npy or
    # .fvecs / .ivecs / .bvecs; <key> selects the distance as in DISTANCE_FUNC; writes train_knn.npy
    import sys

    from nlsh.data_io import load_dataset, save_processed
    if len(sys.argv) != 3 or sys.argv[2] not in DISTANCE_FUNC:
        raise SystemExit(f"usage: python precompute.py <dataset directory> <{' | '.join(DISTANCE_FUNC)}>")
    data = load_dataset(sys.argv[1])
    train_knn = self_get_knn_pt(np.asarray(data["train"], dtype=np.float32), DISTANCE_FUNC[sys.argv[2]])
    save_processed(sys.argv[1], train_knn)
    print(f"train_knn {train_knn.shape} -> {sys.argv[1]}/train_knn.npy")
