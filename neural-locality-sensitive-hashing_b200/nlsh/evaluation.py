"""The callers of the hot path, wired to the CUDA Indexer.

`validate_index` is the validation block of Trainer.fit (nlsh/trainers/base.py:80-115): rebuild
the index from the hasher's live weights, log the bucket statistics, run the validation queries
and report recall / mean candidates / QPS under the reference's metric names.  `recall_sweep` is
the recall-vs-candidates loop of eval.py:148-196 with the deterministic top-p probe sets of this
library in place of the reference's Bernoulli sampling.

Both are thin: the work is Indexer(...) + Indexer.query_tensors(...) (CUDA) and
nlsh.metrics (unchanged).
"""
import time

import numpy as np
import torch

from .indexer import Indexer
from .metrics import calculate_recall, recall_at_k_tensors


def _sync(device):
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def validate_index(hashing, candidate_vectors_gpu, distance_func, validation_vectors_gpu, ground_truth,
                   k=10, hash_times=10, logger=None, global_step=0, metric=None, list_api=True):
    """nlsh/trainers/base.py:80-115.  ground_truth: [Q, >=k] ids (array / list of lists).  Returns
    {"test/n_indexes", "test/std_index_rows", "test/recall", "test/query_size", "test/qps",
    "indexer"}; every scalar is also passed to logger.log(name, value, global_step) when a logger
    is given.  list_api=True times Indexer.query (Python lists, as the reference does); False
    times the tensor API and computes recall on the device."""
    hashing.train_mode(False)
    indexer = Indexer(hashing, candidate_vectors_gpu, distance_func, metric=metric)
    sizes = indexer.bucket_sizes
    out = {"test/n_indexes": int(len(indexer.index2row)),
           "test/std_index_rows": float(np.std(sizes[sizes > 0])) if (sizes > 0).any() else 0.0}
    dev = validation_vectors_gpu.device
    _sync(dev)
    t1 = time.time()
    if list_api:
        result, n_candidates = indexer.query(validation_vectors_gpu, k=k, hash_times=hash_times)
        t2 = time.time()
        gt = [list(r[:k]) for r in (ground_truth.tolist() if hasattr(ground_truth, "tolist") else ground_truth)]
        recall = float(calculate_recall(gt, result, np.mean))
        query_size = float(np.mean(n_candidates)) if len(n_candidates) else 0.0
    else:
        ids, _, ncand = indexer.query_tensors(validation_vectors_gpu, k=k, hash_times=hash_times)
        _sync(dev)
        t2 = time.time()
        gt = torch.as_tensor(np.asarray(ground_truth)[:, :k], dtype=torch.int64, device=dev)
        recall = float(recall_at_k_tensors(gt, ids))
        query_size = float(ncand.double().mean().item()) if ncand.numel() else 0.0
    out["test/recall"] = recall
    out["test/query_size"] = query_size
    out["test/qps"] = validation_vectors_gpu.shape[0] / max(t2 - t1, 1e-12)
    if logger is not None:
        for name in ("test/n_indexes", "test/std_index_rows", "test/recall", "test/query_size", "test/qps"):
            logger.log(name, out[name], global_step)
    out["indexer"] = indexer
    return out


def recall_sweep(indexer, query_vectors_gpu, ground_truth, k=10, probe_counts=(1, 2, 4, 8, 16, 32)):
    """eval.py:148-196: one (avg_n_candidates, recall, qps) row per probe count."""
    dev = query_vectors_gpu.device
    gt = torch.as_tensor(np.asarray(ground_truth)[:, :k], dtype=torch.int64, device=dev)
    rows = []
    n_buckets = len(indexer.bucket_sizes)
    for p in probe_counts:
        if p > n_buckets:
            break
        indexer.query_tensors(query_vectors_gpu[: min(64, query_vectors_gpu.shape[0])], k=k, hash_times=p)  # warm
        _sync(dev)
        t1 = time.time()
        ids, _, ncand = indexer.query_tensors(query_vectors_gpu, k=k, hash_times=p)
        _sync(dev)
        dt = time.time() - t1
        rows.append({"probes": int(p), "avg_n_candidates": float(ncand.double().mean().item()),
                     "recall": float(recall_at_k_tensors(gt, ids)),
                     "qps": query_vectors_gpu.shape[0] / max(dt, 1e-12)})
    return rows
