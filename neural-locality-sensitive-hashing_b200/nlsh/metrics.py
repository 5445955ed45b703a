"""Recall metric — same definitions as the reference's nlsh/metrics.py:4-25, plus a device
twin (`recall_at_k_tensors`) that keeps the ids on the GPU."""
from typing import List, Union

from . import _native


def _recall(y_true: List[int], y_pred: List[int]) -> float:
    # nlsh/metrics.py:4-7
    n_true = len(y_true)
    true_positives = len(set(y_true) & set(y_pred))
    return true_positives / n_true


def calculate_recall(
        y_true: List[List[int]],
        y_pred: List[List[int]],
        reduce_func=None,
    ) -> Union[List[float], float]:
    # nlsh/metrics.py:10-25
    assert len(y_true) == len(y_pred)
    recalls = [_recall(yt, yp) for yt, yp in zip(y_true, y_pred)]
    if reduce_func is not None:
        return reduce_func(recalls)
    return recalls


def recall_at_k_tensors(gt_ids, pred_ids) -> float:
    """mean_q |set(gt[q]) & set(pred[q])| / k_gt for CUDA int64 tensors [Q, k_gt], [Q, k_pred]
    (pred may hold -1 padding).  Same value as calculate_recall(..., np.mean) when the rows
    of gt hold distinct ids."""
    hits = _native.recall_hits(gt_ids, pred_ids)
    return float(hits.double().mean().item()) / gt_ids.shape[1]


def n_recall_at_k(y_true, y_pred, n=1, k=10):
    pass
