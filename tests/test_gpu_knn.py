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
    # expansion form (precompute._l2): d(x, x) is rounding noise on the scale of |x|^2, not exactly 0
    assert torch.equal(ids0[:, 0], torch.arange(200, 300, device="cuda"))
    assert (d0[:, 0].abs() <= 1e-5 * X[200:300].pow(2).sum(1)).all()
    assert torch.equal(ids0[:, 1:], ids[:, :4])
    ids_off, _ = precompute.knn_tensors(X[200:300], X, "l2sq", 5, id_offset=10_000_000_000)
    assert torch.equal(ids_off, ids0 + 10_000_000_000)


@pytest.mark.parametrize("metric", ["l2sq", "cosine"])
@pytest.mark.parametrize("nq,n,d,k", [(300, 70_000, 128, 10), (129, 5000, 100, 100), (1, 64, 4, 1),
                                      (260, 100, 32, 128), (64, 1_200_000, 16, 10), (1000, 20_000, 64, 101)])
def test_tensor_core_knn_matches_simt_and_oracle(oracle, metric, nq, n, d, k):
    """tc_knn.cu (tcgen05 3xTF32 GEMM + top-k epilogue) against the fp32 SIMT path and the oracle."""
    import os
    import precompute
    g = torch.Generator().manual_seed(7 * nq + n + d)
    X = torch.randn(n, d, generator=g) + 0.5
    Q = torch.randn(nq, d, generator=g) + 0.5
    Xc, Qc = X.cuda(), Q.cuda()
    os.environ["NLSH_KNN_IMPL"] = "simt"
    try:
        s_ids, s_d = precompute.knn_tensors(Qc, Xc, metric, k)
    finally:
        os.environ.pop("NLSH_KNN_IMPL")
    t_ids, t_d = precompute.knn_tensors(Qc, Xc, metric, k)
    kk = min(k, n)
    assert (t_ids[:, kk:] == -1).all() and torch.isinf(t_d[:, kk:]).all()
    t_ids, t_d, s_ids, s_d = (v.cpu().numpy()[:, :kk] for v in (t_ids, t_d, s_ids, s_d))
    assert (np.diff(t_d, axis=1) >= 0).all()
    for row in t_ids[:: max(1, nq // 50)]:
        assert len(set(row.tolist())) == kk and row.min() >= 0 and row.max() < n
    # distances: both are fp32 evaluations of the same quantity; the expansion form's rounding lives
    # on the scale |q|^2 + |x|^2 (l2sq) or 1 (cosine)
    if metric == "l2sq":
        scale = (Q.pow(2).sum(1)[:, None] + X.pow(2).sum(1).max()).numpy()
    else:
        scale = np.ones((nq, 1), dtype=np.float32)
    assert (np.abs(t_d - s_d) <= 2e-6 * scale).all(), np.abs(t_d - s_d).max()
    mism = t_ids != s_ids
    if mism.any():  # only where the two evaluations order a near-tie differently
        qs, pos = np.nonzero(mism)
        assert (np.abs(t_d[qs, pos] - s_d[qs, pos]) <= 2e-6 * scale[qs, 0]).all()
        assert mism.mean() < 0.01, mism.mean()
    if n <= 100_000 or nq <= 64:
        o_ids, o_d = oracle.knn_queries(Q, X, metric, kk)
        assert (np.abs(t_d - o_d) <= 1e-5 * scale).all()
        assert (t_ids != o_ids).mean() < 0.01


def test_tensor_core_knn_exclude_self_across_chunks():
    import precompute
    n = 1_100_000  # two database chunks of tc_knn.cu
    X = torch.randn(n, 8, generator=torch.Generator().manual_seed(11)).cuda()
    lo = 1_048_500  # queries whose own rows straddle the chunk boundary (2^20)
    ids, _ = precompute.knn_tensors(X[lo:lo + 200], X, "l2sq", 3, exclude_self=True, self_offset=lo)
    ids0, _ = precompute.knn_tensors(X[lo:lo + 200], X, "l2sq", 4)
    own = torch.arange(lo, lo + 200, device="cuda")
    assert not (ids == own[:, None]).any()
    assert torch.equal(ids0[:, 0], own) and torch.equal(ids0[:, 1:], ids)


def test_recall_kernel_matches_metrics():
    from nlsh.metrics import calculate_recall, recall_at_k_tensors
    g = torch.Generator().manual_seed(0)
    gt = torch.stack([torch.randperm(50, generator=g)[:10] for _ in range(200)])
    pred = torch.stack([torch.randperm(50, generator=g)[:10] for _ in range(200)])
    pred[5, 3:] = -1
    want = calculate_recall(gt.tolist(), [[v for v in r if v >= 0] for r in pred.tolist()], np.mean)
    assert recall_at_k_tensors(gt.cuda(), pred.cuda()) == pytest.approx(want)


@pytest.mark.parametrize("metric_fn", ["_l2", "_cosine_distance"])
def test_nearest_exclude_positive_matches_reference_loop(oracle, metric_fn):
    """nlsh/trainers/triplet.py:44-74 restated densely: mask self + positives, argmin."""
    import precompute
    fn = getattr(precompute, metric_fn)
    g = torch.Generator().manual_seed(3)
    V = torch.randn(700, 24, generator=g).cuda()
    pos = precompute.knn_tensors(V, V, fn, 6, exclude_self=True)[0][:, :5]  # 5 positives per row
    got = precompute.nearest_exclude_positive(V, fn, pos)
    want = oracle.nearest_exclude_positive(V.cpu(), "l2sq" if metric_fn == "_l2" else "cosine", pos.cpu()).cuda()
    same = (got == want)
    if not same.all():  # only where the two nearest admissible rows tie within rounding
        rows = (~same).nonzero().squeeze(1)
        d2 = fn(V, V)
        assert torch.allclose(d2[rows, got[rows]], d2[rows, want[rows]], rtol=1e-4, atol=1e-6)
    assert same.float().mean() > 0.99
    assert not (got[:, None] == pos).any() and not (got == torch.arange(700, device="cuda")).any()
