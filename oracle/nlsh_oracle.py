"""CPU oracle for the Neural-LSH hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain restatement, on the CPU, of the reference's algorithm for the index-build + query
path, function by function, each citing the reference file:line it follows.  The reference
does all of its arithmetic with torch ops, so the floating-point parts use the same torch
CPU ops (torch 2.11 as installed; the reference pins 1.3.0); the integer
parts are numpy / pure Python.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
CPU-baseline / `--impl reference` legs may import this module; nothing under
`neural-locality-sensitive-hashing_b200/` does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here against
  (a) the reference's own two golden tests for this path (nlsh/tests/test_indexer.py:6-26,
      tests/test_precompute.py:6-21), and
  (b) outputs of the reference itself (imported from /root/reference, incl. its Cython
      utils.pyx) on seeded inputs, committed as tests/golden/nlsh_golden.npz by
      tests/golden/make_golden.py.
`topp_probes` and `merge_topk` have no reference counterpart (the reference samples probes
at random and is single-GPU): they restate this repo's own specification and are marked so.
"""
import ctypes
import os
from typing import Dict, List, Sequence, Set

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))

HEAD_SIGMOID, HEAD_TANH, HEAD_SOFTMAX = 0, 1, 2


# ---------------------------------------------------------------------------------------
# nlsh/utils.pyx
# ---------------------------------------------------------------------------------------
def binarr_to_int(binarr) -> int:
    """utils.pyx:6-15: MSB-first pack in an int32 accumulator, returned as np.int16_t."""
    out = 0
    for bit in binarr:
        out = ((out << 1) | int(bit)) & 0xFFFFFFFF  # int32 accumulator (wraps)
    out &= 0xFFFF  # C conversion int32 -> int16 keeps the low 16 bits
    return out - 0x10000 if out >= 0x8000 else out


def hash_codes(codes) -> List[Set[int]]:
    """utils.pyx:18-32: per row the set of packed codes over the n_samples axis.
    Buffer protocol errors of the `int[:, :, :]` signature are restated as ValueError."""
    arr = np.asarray(codes)
    if arr.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions")
    if arr.dtype != np.intc:
        raise ValueError("Buffer dtype mismatch")
    return [set(binarr_to_int(arr[i, j, :]) for j in range(arr.shape[1]))
            for i in range(arr.shape[0])]


def hash_codes_c(codes) -> List[Set[int]]:
    """Same through the plain-C restatement oracle/pack_codes.c (built by build())."""
    lib = ctypes.CDLL(os.path.join(_HERE, "liboracle_pack.so"))
    arr = np.ascontiguousarray(codes, dtype=np.intc)
    n, s, hs = arr.shape
    out = np.empty((n, s), dtype=np.int16)
    lib.oracle_pack_codes(ctypes.c_void_p(arr.ctypes.data), ctypes.c_long(n), ctypes.c_long(s),
                          ctypes.c_long(hs), ctypes.c_void_p(out.ctypes.data))
    return [set(r) for r in out.tolist()]


# ---------------------------------------------------------------------------------------
# encoders.py + nlsh/hashings.py
# ---------------------------------------------------------------------------------------
class Layer:
    """weight [out, in], bias [out] | None (torch CPU fp32), relu flag; sin_w0 = w0 of a SIREN sine
    layer (encoders.py:58-79 through the third-party siren-torch: sin(w0 * x)), None otherwise."""

    def __init__(self, weight, bias, relu, sin_w0=None):
        self.weight = torch.as_tensor(weight, dtype=torch.float32)
        self.bias = None if bias is None else torch.as_tensor(bias, dtype=torch.float32)
        self.relu = bool(relu)
        self.sin_w0 = None if sin_w0 is None else float(sin_w0)


def mlp_logits(x, layers: Sequence[Layer]) -> torch.Tensor:
    """encoders.py:41-55 ([Linear -> ReLU] per hidden dim) followed by the output Linear of
    hashings.py:19-22; returns the pre-sigmoid logits."""
    h = torch.as_tensor(x, dtype=torch.float32)
    with torch.no_grad():
        for L in layers:
            h = F.linear(h, L.weight, L.bias)
            if L.relu:
                h = F.relu(h)
            if L.sin_w0 is not None:
                h = torch.sin(L.sin_w0 * h)
    return h


def hard_bits(logits: torch.Tensor, head: int) -> torch.Tensor:
    """hashings.py:24-26 + 67-72: probs = sigmoid(l) (or tanh(l)/2 + 0.5), bit = probs > 0.5,
    evaluated in fp32 exactly as the reference does."""
    if head == HEAD_TANH:
        probs = torch.tanh(logits) / 2. + 0.5
    else:
        probs = torch.sigmoid(logits)
    return (probs > 0.5)


def hard_codes(logits: torch.Tensor, head: int) -> np.ndarray:
    """Bucket code per row: hashings.py:72-76 + utils.pyx:6-15 (n == 1), or argmax of the
    softmax for the Categorical head (hashings.py:104-106, 131-133)."""
    if head == HEAD_SOFTMAX:
        return F.softmax(logits, dim=1).argmax(axis=1).numpy().astype(np.int32)
    bits = hard_bits(logits, head).unsqueeze(1).int().numpy()  # [n, 1, hs] as hashings.py:72
    return np.array([binarr_to_int(b[0]) for b in bits], dtype=np.int32)


def hash_sets(x, layers, head=HEAD_SIGMOID) -> List[Set[int]]:
    """MultivariateBernoulli.hash(x, n=1) (hashings.py:66-92): list of singleton sets."""
    logits = mlp_logits(x, layers)
    bits = hard_bits(logits, head).unsqueeze(1).int().numpy()
    return hash_codes(bits)


def topp_probes(logits, head: int, p: int) -> np.ndarray:
    """[NO REFERENCE COUNTERPART — specification of this repo's deterministic multi-probe,
    replacing the Bernoulli sampling of hashings.py:77-81.]
    Row i: the p cheapest codes, cost(mask) = sum over flipped bits of |logit| added in
    logit-index order in fp32, ties by smaller mask; code = hard_code ^ mask.  Softmax head:
    the p largest logits, ties by smaller index.  Padded with -1."""
    l = np.asarray(logits, dtype=np.float32)
    n, hs = l.shape
    out = np.full((n, p), -1, dtype=np.int32)
    if head == HEAD_SOFTMAX:
        for i in range(n):
            order = sorted(range(hs), key=lambda c: (-float(l[i, c]), c))[:p]
            out[i, :len(order)] = order
        return out
    base = hard_codes(torch.from_numpy(l), head)
    masks = np.arange(1 << hs, dtype=np.int64)
    for i in range(n):
        cost = np.zeros(1 << hs, dtype=np.float32)
        a = np.abs(l[i])
        for j in range(hs):
            sel = ((masks >> (hs - 1 - j)) & 1).astype(bool)
            cost = np.where(sel, (cost + a[j]).astype(np.float32), cost)
        order = np.lexsort((masks, cost))[:p]
        out[i, :len(order)] = int(base[i]) ^ masks[order]
    return out


def _philox4x32_10(ctr, key):
    """Philox-4x32-10 (Salmon et al., SC'11) on uint32 numpy arrays: ctr [4, ...], key [2, ...]."""
    c = [np.asarray(v, dtype=np.uint64) for v in ctr]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & m32, p1 & m32, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & m32, p0 & m32]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c


def sample_probes(logits, head: int, p: int, seed: int):
    """hashings.py:66-81 with n = p: column 0 the hard code, columns 1.. the packed codes of p - 1 draws of
    torch.distributions.Bernoulli(probs) - restated with the counter-based generator of this repo's kernel
    (uniform u(row, j, bit) = (Philox(seed; row, j, bit // 4)[bit % 4] >> 8) * 2^-24, bit set iff u < prob), since
    torch's own sampling stream depends on torch's launch geometry and cannot be reproduced.  Returns (probes
    int32 [n, p], margin fp32 [n, p]: min over bits of |u - prob|, so a test can skip draws that a one-ulp
    difference in exp / tanh could flip)."""
    l = torch.as_tensor(np.asarray(logits, dtype=np.float32))
    n, hs = l.shape
    probs = (torch.tanh(l) / 2. + 0.5 if head == HEAD_TANH else torch.sigmoid(l)).numpy().astype(np.float32)
    rows = np.arange(n, dtype=np.uint64)[:, None, None]
    js = np.arange(p, dtype=np.uint64)[None, :, None]
    groups = np.arange((hs + 3) // 4, dtype=np.uint64)[None, None, :]
    shape = (n, p, (hs + 3) // 4)
    words = _philox4x32_10([np.broadcast_to(rows & np.uint64(0xFFFFFFFF), shape), np.broadcast_to(rows >> np.uint64(32), shape),
                            np.broadcast_to(js, shape), np.broadcast_to(groups, shape)],
                           (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = np.stack(words, axis=-1).reshape(n, p, -1)[:, :, :hs]            # [n, p, hs] uint64 words
    u = ((u >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24))
    bits = (u < probs[:, None, :])
    bits[:, 0, :] = hard_bits(l, head).numpy()
    margin = np.abs(u - probs[:, None, :]).min(axis=2)
    margin[:, 0] = 1.0
    weights = (1 << np.arange(hs - 1, -1, -1)).astype(np.int64)
    return (bits.astype(np.int64) * weights).sum(axis=2).astype(np.int32), margin


# ---------------------------------------------------------------------------------------
# nlsh/indexer.py
# ---------------------------------------------------------------------------------------
def build_index(indexes) -> Dict[int, np.ndarray]:
    """indexer.py:6-24: dict code -> row ids in order of appearance (= ascending)."""
    index2row: Dict[int, list] = {}
    for idx, index_set in enumerate(indexes):
        for index in index_set:
            if index not in index2row:
                index2row[index] = [idx]
            else:
                index2row[index].append(idx)
    return {k: np.asarray(v, dtype=np.int64) for k, v in index2row.items()}


def scan_distance(q, cand, metric: str) -> torch.Tensor:
    """nlsh/data.py:192-201 (SIFT.distance = F.pairwise_distance: eps 1e-6 inside, sqrt) or
    nlsh/data.py:100-109 (Glove.distance = 1 - F.cosine_similarity(dim=-1))."""
    if metric == "l2":
        return F.pairwise_distance(q, cand)
    if metric == "angular":
        return 1 - F.cosine_similarity(q, cand, dim=-1)
    raise ValueError(metric)


def query(vectors, index2row, queries, probe_sets, metric: str, k: int, fallback="sorted"):
    """Indexer.query (indexer.py:56-96) for given per-query probe sets.

    Per query: concatenate the probed buckets' rows, gather the vectors (index_select,
    77-82), distance to the query (84-87), topk(k, largest=False) (90), map to row ids (91).
    Returns (ids per query, distances per query, n_candidates per query).
    fallback: what to return when there are fewer than k candidates.  "reference" = the
    last probed bucket's rows, unsorted (indexer.py:92-93); "sorted" = all candidates by
    ascending distance, which is what the CUDA path returns."""
    X = torch.as_tensor(vectors, dtype=torch.float32)
    Q = torch.as_tensor(queries, dtype=torch.float32)
    empty = torch.zeros((0,), dtype=torch.int64)
    ids_out, d_out, n_out = [], [], []
    for qi, keys in enumerate(probe_sets):
        rows_list = [torch.as_tensor(index2row.get(key, empty), dtype=torch.int64)
                     for key in list(keys)]
        rows = torch.cat(rows_list) if rows_list else empty
        n_cand = int(rows.shape[0])
        n_out.append(n_cand)
        if n_cand == 0:
            ids_out.append([])
            d_out.append([])
            continue
        cand = torch.index_select(X, 0, rows)
        dist = scan_distance(Q[qi], cand, metric)
        if n_cand >= k:
            dv, di = dist.topk(k, largest=False)
        elif fallback == "reference":
            last = rows_list[-1]
            ids_out.append(last.tolist())
            d_out.append(scan_distance(Q[qi], X[last], metric).tolist())
            continue
        else:
            dv, di = dist.topk(n_cand, largest=False)
        ids_out.append(rows[di].tolist())
        d_out.append(dv.tolist())
    return ids_out, d_out, n_out


# ---------------------------------------------------------------------------------------
# nlsh/metrics.py
# ---------------------------------------------------------------------------------------
def recall(y_true, y_pred) -> float:
    """metrics.py:4-25 with reduce_func = mean."""
    assert len(y_true) == len(y_pred)
    vals = [len(set(map(int, t)) & set(map(int, p))) / len(t) for t, p in zip(y_true, y_pred)]
    return float(np.mean(vals))


# ---------------------------------------------------------------------------------------
# precompute.py
# ---------------------------------------------------------------------------------------
def knn_distance_matrix(v1, v2, metric: str) -> torch.Tensor:
    """precompute.py:37-54 (_l2: squared, expansion form via addmm) or precompute.py:22-34
    (_cosine_distance: normalise rows, 1 - mm)."""
    if metric == "l2sq":
        v1_norm = v1.pow(2).sum(dim=-1, keepdim=True)
        v2_norm = v2.pow(2).sum(dim=-1, keepdim=True)
        return torch.addmm(v2_norm.transpose(-2, -1), v1, v2.transpose(-2, -1), alpha=-2).add_(v1_norm)
    if metric == "cosine":
        a = v1 / v1.norm(dim=1)[:, None]
        b = v2 / v2.norm(dim=1)[:, None]
        return 1 - torch.mm(a, b.T)
    raise ValueError(metric)


def self_knn(vectors, metric: str, k: int, batch_size: int = 512) -> np.ndarray:
    """self_get_knn_pt (precompute.py:57-67): batches of rows against all rows,
    topk(k+1, largest=False), first hit dropped as "self"."""
    X = torch.as_tensor(vectors, dtype=torch.float32)
    n = X.shape[0]
    knn = np.zeros((n, k), dtype=int)
    for batch_idx in range(n // batch_size + 1):
        start, end = batch_idx * batch_size, (batch_idx + 1) * batch_size
        target = X[start:end, :]
        if target.shape[0] == 0:
            continue
        knn[start:end, :] = knn_distance_matrix(target, X, metric).topk(
            k + 1, dim=1, largest=False)[1][:, 1:].numpy()
    return knn


def knn_queries(queries, vectors, metric: str, k: int, batch_size: int = 512):
    """Held-out-query variant of the above (no self to drop): (ids [Q, k], dists [Q, k]).
    metric may also be the scan metrics 'l2' / 'angular' (row-wise scan_distance)."""
    X = torch.as_tensor(vectors, dtype=torch.float32)
    Q = torch.as_tensor(queries, dtype=torch.float32)
    ids = np.zeros((Q.shape[0], k), dtype=np.int64)
    dd = np.zeros((Q.shape[0], k), dtype=np.float32)
    if metric in ("l2", "angular"):
        for i in range(Q.shape[0]):
            dv, di = scan_distance(Q[i], X, metric).topk(k, largest=False)
            ids[i], dd[i] = di.numpy(), dv.numpy()
        return ids, dd
    for s in range(0, Q.shape[0], batch_size):
        dv, di = knn_distance_matrix(Q[s:s + batch_size], X, metric).topk(k, dim=1, largest=False)
        ids[s:s + batch_size], dd[s:s + batch_size] = di.numpy(), dv.numpy()
    return ids, dd


# ---------------------------------------------------------------------------------------
# multi-GPU merge (no reference counterpart)
# ---------------------------------------------------------------------------------------
def nearest_exclude_positive(vectors, metric: str, positive_indexes, batch_size: int = 32) -> torch.Tensor:
    """nlsh/trainers/triplet.py:44-74: per batch of rows the full distance matrix, the row's positives
    and the row itself overwritten with the matrix maximum, then argmin (the `.cuda()` calls dropped)."""
    vectors = torch.as_tensor(vectors, dtype=torch.float32)
    positive_indexes = torch.as_tensor(positive_indexes, dtype=torch.int64)
    n = vectors.shape[0]
    out = []
    for start in range(0, n, batch_size):  # the reference's separate last-batch branch does the same
        batch = vectors[start:start + batch_size]
        distances = knn_distance_matrix(batch, vectors, metric).clone()
        max_value = float(distances.max())  # (the reference passes the 0-d tensor; torch 2 wants the scalar overload)
        distances.scatter_(1, positive_indexes[start:start + batch_size], max_value)
        diag = torch.arange(start, start + batch.shape[0]).reshape(-1, 1)
        distances.scatter_(1, diag, max_value)  # don't select self
        out.append(distances.argmin(dim=1))
    return torch.cat(out)


def merge_topk(dists, ids, k: int):
    """[NO REFERENCE COUNTERPART] k smallest (distance, id) pairs over G per-shard lists
    [G, Q, k]; ids < 0 are empty slots.  Returns (ids [Q, k] padded -1, dists padded +inf)."""
    dists = np.asarray(dists, dtype=np.float32)
    ids = np.asarray(ids, dtype=np.int64)
    g, nq, kk = dists.shape
    out_i = np.full((nq, k), -1, dtype=np.int64)
    out_d = np.full((nq, k), np.inf, dtype=np.float32)
    for q in range(nq):
        cand = [(float(dists[s, q, e]), int(ids[s, q, e])) for s in range(g) for e in range(kk)
                if ids[s, q, e] >= 0]
        cand.sort()
        for r, (dv, iv) in enumerate(cand[:k]):
            out_i[q, r], out_d[q, r] = iv, dv
    return out_i, out_d


# ---------------------------------------------------------------------------------------
# the reference's end-to-end CPU flow, for the timed CPU baseline (bench.py)
# ---------------------------------------------------------------------------------------
class CpuIndexer:
    """Indexer (indexer.py:27-96) restated end to end on the CPU, including the host work
    the reference spends its time in: 4096-row hash batches (40-54), Bernoulli construction
    and `.cpu().numpy()` + hash_codes per batch (hashings.py:66-92), dict build (6-24), and
    the per-query gather / distance / topk / `.tolist()` loop (62-95)."""

    def __init__(self, layers, head, vectors, metric, hash_codes_fn=None):
        self.layers, self.head, self.metric = layers, head, metric
        self.X = torch.as_tensor(vectors, dtype=torch.float32)
        self._hash_codes = hash_codes_fn or hash_codes
        self.index2row = {k: torch.from_numpy(v) for k, v in
                          build_index(self.hash(self.X)).items()}

    def _hash_batch(self, batch):
        logits = mlp_logits(batch, self.layers)
        probs = torch.tanh(logits) / 2. + 0.5 if self.head == HEAD_TANH else torch.sigmoid(logits)
        dist = torch.distributions.Bernoulli(probs)  # hashings.py:70 (argument validation cost)
        codes = (dist.probs > 0.5).unsqueeze_(1).int().cpu().numpy()
        return self._hash_codes(codes)

    def hash(self, vectors, batch_size=4096):
        keys = []
        n = vectors.shape[0]
        n_batches = n // batch_size
        for idx in range(n_batches):
            keys += self._hash_batch(vectors[idx * batch_size:(idx + 1) * batch_size, :])
        if n > n_batches * batch_size:
            keys += self._hash_batch(vectors[n_batches * batch_size:, :])
        return keys

    def query(self, queries, k=10, probe_sets=None):
        Q = torch.as_tensor(queries, dtype=torch.float32)
        query_indexes = probe_sets if probe_sets is not None else self.hash(Q)
        recall_result, n_candidates_result = [], []
        vector_buffer = torch.empty_like(self.X)  # indexer.py:60
        default_empty_rows = torch.LongTensor([])
        for idx, qi in enumerate(query_indexes):
            n_candidates, buffer_start, candidate_rows_list = 0, 0, []
            buffer_end = 0
            for key in list(qi):
                candidate_rows = self.index2row.get(key, default_empty_rows)
                candidate_rows_list.append(candidate_rows)
                n_candidates += len(candidate_rows)
                buffer_end = buffer_start + len(candidate_rows)
                torch.index_select(self.X, 0, candidate_rows,
                                   out=vector_buffer[buffer_start:buffer_end, :])
                buffer_start = buffer_end
            distance = scan_distance(Q[idx, :], vector_buffer[:buffer_end, :], self.metric)
            concat_candidate_rows = torch.cat(candidate_rows_list)
            try:
                topk_idxs = distance.topk(k, largest=False)[1]
                topk_idxs = concat_candidate_rows[topk_idxs].tolist()
            except RuntimeError:
                topk_idxs = candidate_rows.tolist()
            n_candidates_result.append(n_candidates)
            recall_result.append(topk_idxs)
        return recall_result, n_candidates_result
