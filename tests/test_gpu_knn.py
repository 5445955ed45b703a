"""GPU parity of the brute-force kNN (nlsh_knn_bruteforce through precompute.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rows_equal_up_to_ties(got, want, dist_matrix, rtol):
    """ids identical except where the reference's own distances tie within rtol."""
    bad = 0
    for i in range(got.shape[0]):
        if np.array_equal(got[i], want[i]):
            continue
        for a, b in zip(got[i], want[i]):
            if a != b:
                da, db = dist_matrix[i, a], dist_matrix[i, b]
                assert abs(da - db) <= rtol * max(abs(da), abs(db), 1e-30), (i, a, b, da, db)
        bad += 1
    return bad


def test_reference_test_precompute_golden_through_api():
    # tests/test_precompute.py:6-21 (the function is called self_get_knn_pt today)
    import precompute
    vectors = np.array([[1.2, 2, 3], [3, 2, 1], [1, 2, 4], [6, 4, 2.5], [2, 4, 6]], dtype=np.float32)
    result = precompute.self_get_knn_pt(vectors, precompute._cosine_distance, k=2, batch_size=2)
    assert result.dtype == np.int64 and result.shape == (5, 2)
    assert [set(r) for r in list(result)] == [{4, 2}, {3, 0}, {0, 4}, {1, 0}, {0, 2}]


def test_self_knn_matches_reference(golden, oracle):
    import precompute
    V = golden["knn_vectors"]
    Vt = torch.from_numpy(V)
    for fn, metric, key in [(precompute._l2, "l2sq", "knn_l2_k10"), (precompute._cosine_distance, "cosine", "knn_cos_k10")]:
        got = precompute.self_get_knn_pt(V, fn, k=10, batch_size=256)
        dm = oracle.knn_distance_matrix(Vt, Vt, metric).numpy()
        # the expansion form of precompute._l2 carries ~1e-6 * |x|^2 of rounding: ties at that level
        rows_equal_up_to_ties(got, golden[key], dm, rtol=1e-4 if metric == "l2sq" else 1e-5)
        got_self = precompute.self_get_knn_pt(V, fn, k=10, exclude="self")
        rows_equal_up_to_ties(got_self, golden[key], dm, rtol=1e-4 if metric == "l2sq" else 1e-5)


@pytest.mark.parametrize("metric", ["l2", "angular", "l2sq", "cosine"])
@pytest.mark.parametrize("nq,n,d,k", [(100, 5000, 128, 10), (17, 3000, 100, 100), (300, 257, 32, 5), (5, 40, 8, 64)])
def test_knn_queries_against_oracle(oracle, metric, nq, n, d, k):
    import precompute
    g = torch.Generator().manual_seed(nq + n)
    X = torch.randn(n, d, generator=g)
    Q = torch.randn(nq, d, generator=g)
    ids, dists = precompute.knn_tensors(Q.cuda(), X.cuda(), metric, k)
    ids, dists = ids.cpu().numpy(), dists.cpu().numpy()
    kk = min(k, n)
    o_ids, o_d = oracle.knn_queries(Q, X, metric, kk)
    assert (ids[:, kk:] == -1).all() and np.isinf(dists[:, kk:]).all()
    # distances: difference form here vs expansion form in the reference -> compare on the scale
    # the reference's rounding lives on (|q|^2 + |x|^2) for l2sq, 1e-5 relative otherwise
    if metric == "l2sq":
        scale = (Q.pow(2).sum(1)[:, None] + X.pow(2).sum(1).max()).numpy()
        assert (np.abs(dists[:, :kk] - o_d) <= 1e-5 * scale).all()
    else:
        np.testing.assert_allclose(dists[:, :kk], o_d, rtol=1e-5, atol=2e-6 if metric != "l2" else 0)
    mism = (ids[:, :kk] != o_ids)
    if mism.any():  # only at near-ties
        qs, pos = np.nonzero(mism)
        for q, p_ in zip(qs, pos):
            assert abs(dists[q, p_] - o_d[q, p_]) <= 1e-4 * max(abs(o_d[q, p_]), 1e-6)
        assert mism.mean() < 0.01


def test_exclude_self_and_offsets():
    import precompute
    X = torch.randn(1000, 16, generator=torch.Generator().manual_seed(4)).cuda()
    ids, _ = precompute.knn_tensors(X[200:300], X, "l2sq", 5, exclude_self=True, self_offset=200)
    assert not (ids == torch.arange(200, 300, device="cuda")[:, None]).any()
    ids0, d0 = precompute.knn_tensors(X[200:300], X, "l2sq", 5)
    assert torch.equal(ids0[:, 0], torch.arange(200, 300, device="cuda")) and (d0[:, 0] == 0).all()
    assert torch.equal(ids0[:, 1:], ids[:, :4])
    ids_off, _ = precompute.knn_tensors(X[200:300], X, "l2sq", 5, id_offset=10_000_000_000)
    assert torch.equal(ids_off, ids0 + 10_000_000_000)


def test_recall_kernel_matches_metrics():
    from nlsh.metrics import calculate_recall, recall_at_k_tensors
    g = torch.Generator().manual_seed(0)
    gt = torch.stack([torch.randperm(50, generator=g)[:10] for _ in range(200)])
    pred = torch.stack([torch.randperm(50, generator=g)[:10] for _ in range(200)])
    pred[5, 3:] = -1
    want = calculate_recall(gt.tolist(), [[v for v in r if v >= 0] for r in pred.tolist()], np.mean)
    assert recall_at_k_tensors(gt.cuda(), pred.cuda()) == pytest.approx(want)
