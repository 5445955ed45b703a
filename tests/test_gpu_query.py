"""GPU parity of the candidate scan + top-k (nlsh_query_scan_topk through Indexer) against
the oracle and the reference-generated golden vectors."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import (assert_topk_equal_up_to_ties, hashing_from_golden, mixture,
                     oracle_layers_from_hashing, rows_to_sets, unpad)

pytestmark = pytest.mark.gpu
ANGULAR_ATOL = 2e-6  # 1 - cos cancels near 0


@pytest.fixture(autouse=True, params=["auto", "tc"])
def scan_impl(request, monkeypatch):
    """Every test of this file runs twice: with the library's own choice of scan kernel (small
    batches go to the fp32 SIMT kernel) and with the tensor-core filtered kernel forced."""
    if request.param == "tc":
        monkeypatch.setenv("NLSH_SCAN_IMPL", "tc")
    else:
        monkeypatch.delenv("NLSH_SCAN_IMPL", raising=False)
    return request.param


def lists(ids, dists):
    ids, dists = ids.cpu().numpy(), dists.cpu().numpy()
    out_i, out_d = [], []
    for r, dd in zip(ids, dists):
        keep = r >= 0
        out_i.append(r[keep].tolist())
        out_d.append(dd[keep].tolist())
    return out_i, out_d


@pytest.mark.parametrize("tag,metric", [("l2", "l2"), ("ang", "angular")])
@pytest.mark.parametrize("flags", [0, 1])
def test_query_matches_reference(golden, oracle, tag, metric, flags):
    from nlsh.indexer import Indexer
    h = hashing_from_golden(golden, tag)
    X = torch.from_numpy(golden[f"{tag}_X"]).cuda()
    Q = torch.from_numpy(golden[f"{tag}_Q"]).cuda()
    k = int(golden[f"{tag}_k"])
    idx = Indexer(h, X, None, metric=metric)
    idx.scan_flags = flags
    atol = ANGULAR_ATOL if metric == "angular" else 0.0
    # single probe: feed the reference's own query codes so both sides scan the same buckets
    probes = torch.from_numpy(golden[f"{tag}_q_codes"].astype(np.int32))[:, None].cuda()
    ids, dists, ncand = idx.query_tensors(Q, k=k, probes=probes)
    assert ncand.cpu().tolist() == golden[f"{tag}_query_ncand"].tolist()
    want_ids = unpad(golden[f"{tag}_query_ids"])
    want_d = [row[:len(w)] for row, w in zip(golden[f"{tag}_query_dists"], want_ids)]
    got_i, got_d = lists(ids, dists)
    full = [i for i, n in enumerate(golden[f"{tag}_query_ncand"]) if n >= k]  # <k: SURVEY Q8
    assert_topk_equal_up_to_ties([got_i[i] for i in full], [got_d[i] for i in full],
                                 [want_ids[i] for i in full], [want_d[i] for i in full],
                                 rtol=1e-5, atol=atol)
    # multi-probe with the reference's sampled probe sets
    mp_sets = rows_to_sets(golden[f"{tag}_mp_sets"])
    probes = Indexer.probes_from_sets(mp_sets, "cuda")
    ids, dists, ncand = idx.query_tensors(Q, k=k, probes=probes)
    assert ncand.cpu().tolist() == golden[f"{tag}_mp_ncand"].tolist()
    index2row = {int(c): idx.index2row[c].cpu().numpy() for c in idx.index2row}
    o_ids, o_d, _ = oracle.query(X.cpu(), index2row, Q.cpu(), mp_sets, metric, k)
    got_i, got_d = lists(ids, dists)
    assert_topk_equal_up_to_ties(got_i, got_d, o_ids, o_d, rtol=1e-5, atol=atol)
    ref_ids = unpad(golden[f"{tag}_mp_ids"])
    full = [i for i, n in enumerate(golden[f"{tag}_mp_ncand"]) if n >= k]
    same = sum(got_i[i] == ref_ids[i] for i in full)
    assert same >= len(full) - 1  # identical to the reference's own output except at ties
    # list API (Indexer.query)
    l_ids, l_n = idx.query(Q, k=k, probes=probes)
    assert l_ids == got_i and l_n == golden[f"{tag}_mp_ncand"].tolist()


@pytest.mark.parametrize("n,d,hs,nq,p,k,metric", [
    (20000, 128, 4, 200, 1, 10, "l2"),        # config-1 shape, scaled
    (30000, 128, 8, 300, 4, 10, "l2"),        # config-2 shape: multi-probe
    (30000, 128, 8, 64, 16, 10, "l2"),
    (24000, 100, 10, 200, 8, 10, "angular"),  # config-3 shape: D = 100, cosine
    (6000, 960, 9, 40, 4, 100, "l2"),         # config-5 shape: wide rows, k = 100
    (5000, 30, 3, 50, 2, 7, "l2"),            # D not a multiple of 4 (padded rows, masked tail)
    (5000, 6, 2, 50, 3, 40, "angular"),
    (3000, 64, 5, 1, 5, 10, "l2"),            # one query: rows split across CTAs
    (50000, 32, 2, 9, 1, 64, "l2"),           # big buckets, few queries => many chunks
])
def test_query_against_oracle(oracle, n, d, hs, nq, p, k, metric):
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    torch.manual_seed(n + d + hs)
    X = mixture(n, d, 3 << hs, seed=n)
    Q = mixture(nq, d, 3 << hs, seed=n)  # same centres, different draws below
    Q = Q + 0.1 * torch.randn(Q.shape, generator=torch.Generator().manual_seed(7))
    h = MultivariateBernoulli(MultiLayerRelu(d, [64, 64]), hs, None)
    h.train_mode(False)
    idx = Indexer(h, X.cuda(), None, metric=metric)
    probes = idx.hash_tensors(Q.cuda(), p)
    assert probes.shape == (nq, p)
    ids, dists, ncand = idx.query_tensors(Q.cuda(), k=k, probes=probes)
    sets = [set(int(c) for c in row if c >= 0) for row in probes.cpu().numpy()]
    index2row = {int(c): idx.index2row[c].cpu().numpy() for c in idx.index2row}
    o_ids, o_d, o_n = oracle.query(X, index2row, Q, sets, metric, k)
    assert ncand.cpu().tolist() == o_n
    got_i, got_d = lists(ids, dists)
    assert_topk_equal_up_to_ties(got_i, got_d, o_ids, o_d, rtol=1e-5,
                                 atol=ANGULAR_ATOL if metric == "angular" else 0.0)
    # both staging variants give identical bits
    idx.scan_flags = 1
    ids2, dists2, _ = idx.query_tensors(Q.cuda(), k=k, probes=probes)
    assert torch.equal(ids, ids2) and torch.equal(dists, dists2)
    # results ascend by (distance, id)
    dd = dists.cpu().numpy()
    assert (np.diff(np.where(np.isinf(dd), np.float32(3e38), dd), axis=1) >= 0).all()


@pytest.mark.parametrize("n,d,hs,nq,p,k,metric,kind", [
    (200_000, 128, 6, 3000, 4, 10, "l2", "mixture"),     # ~190 queries per bucket: several query groups
    (200_000, 128, 10, 2000, 8, 10, "l2", "mixture"),    # small buckets (cfg4 / 8-GPU shard shape)
    (100_000, 100, 8, 1500, 4, 10, "angular", "mixture"),
    (60_000, 128, 5, 500, 3, 32, "l2", "offset"),        # |x|^2 >> d^2: the filter passes almost everything
    (60_000, 64, 5, 500, 3, 10, "angular", "offset"),
    (40_000, 30, 4, 300, 2, 7, "l2", "duplicates"),      # exact ties -> (distance, id) order; d % 4 != 0
    (40_000, 8, 4, 300, 2, 10, "angular", "duplicates"),
    (3000, 128, 2, 5000, 2, 10, "l2", "mixture"),        # far more queries than rows per bucket
    (30001, 128, 5, 600, 3, 10, "l2", "mixture"),        # row count not a multiple of 4: the last rows' norms
    (9999, 64, 3, 500, 2, 10, "angular", "mixture"),     # are read outside the 16-byte aligned bulk copies
    (777, 16, 1, 300, 2, 32, "l2", "mixture"),           # two buckets of a few tiles, k = 32
    (20_000, 960, 5, 300, 3, 100, "l2", "mixture"),      # config-5 shape: wide rows (queries ride with the row
    (20_000, 200, 5, 300, 3, 10, "angular", "mixture"),  # tiles' K blocks), k up to 128
    (20_000, 132, 4, 200, 2, 50, "l2", "mixture"),       # a last K block of 4 columns
    (9000, 520, 3, 150, 2, 128, "l2", "duplicates"),
])
def test_tensor_core_filter_is_exact(monkeypatch, n, d, hs, nq, p, k, metric, kind):
    """scan_tc.cu (tcgen05 tf32 GEMM as a filter + exact re-rank) must return bit-for-bit what the
    fp32 SIMT scan kernel returns: a pair the filter dropped wrongly would show up here."""
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    monkeypatch.setenv("NLSH_SCAN_IMPL", "tc")  # also where the library would pick the SIMT kernel by itself
    torch.manual_seed(n + d + hs)
    X = mixture(n, d, 3 << hs, seed=n)
    Q = mixture(nq, d, 3 << hs, seed=n) + 0.1 * torch.randn(nq, d, generator=torch.Generator().manual_seed(7))
    if kind == "offset":
        X, Q = X + 200.0, Q + 200.0
    if kind == "duplicates":
        X[n // 2:] = X[: n - n // 2]
        Q[: nq // 2] = X[: nq // 2]
    h = MultivariateBernoulli(MultiLayerRelu(d, [64, 64]), hs, None)
    h.train_mode(False)
    idx = Indexer(h, X.cuda(), None, metric=metric)
    probes = idx.hash_tensors(Q.cuda(), p)
    ids, dists, ncand = idx.query_tensors(Q.cuda(), k=k, probes=probes)
    idx.scan_flags = 2  # bit 1: no tensor-core filter
    ids2, dists2, ncand2 = idx.query_tensors(Q.cuda(), k=k, probes=probes)
    assert torch.equal(ncand, ncand2)
    assert torch.equal(ids, ids2), (ids != ids2).sum().item()
    assert torch.equal(dists, dists2)
    assert (ids[:, 0] >= 0).all()
    # the same through the filter's other regimes: candidate buffers of k entries (almost every query
    # overflows and is re-scanned exactly in the merge), no threshold ladder (seed bound only), no seed
    # (bound +inf: every row is scored), the scorer's query from global / shared memory (the latter with the
    # lazy item release), 16 SMs left free, 128 / 32 queries per item
    for flags, env in ((4, {}), (0, {"NLSH_TC_LADDER": "0"}), (0, {"NLSH_SCAN_SEED": "0"}),
                       (0, {"NLSH_TC_QGLOBAL": "1"}), (0, {"NLSH_TC_QGLOBAL": "0"}), (16 << 8, {}),
                       (0, {"NLSH_TC_NQ": "128"}), (0, {"NLSH_TC_NQ": "32"}),
                       # 32-byte row loads on / off with either query source, the accumulator ring 2 / 16 deep,
                       # the shallowest slot ring
                       (0, {"NLSH_TC_V8": "1", "NLSH_TC_QGLOBAL": "1"}), (0, {"NLSH_TC_V8": "1", "NLSH_TC_QGLOBAL": "0"}),
                       (0, {"NLSH_TC_V8": "0", "NLSH_TC_QGLOBAL": "0"}), (0, {"NLSH_TC_SETS": "2"}),
                       (0, {"NLSH_TC_SETS": "16"}), (0, {"NLSH_TC_SLOTS": "3"})):
        if env.get("NLSH_SCAN_SEED") == "0" and n * nq * p > 2e9:
            continue  # scoring every pair one thread at a time is only for the small cases
        for name, val in env.items():
            monkeypatch.setenv(name, val)
        idx.scan_flags = flags
        ids3, dists3, ncand3 = idx.query_tensors(Q.cuda(), k=k, probes=probes)
        for name in env:
            monkeypatch.delenv(name)
        assert torch.equal(ncand, ncand3), (flags, env)
        assert torch.equal(ids2, ids3), (flags, env, (ids2 != ids3).sum().item())
        assert torch.equal(dists2, dists3), (flags, env)


def test_edge_cases(oracle):
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    torch.manual_seed(0)
    X = mixture(2000, 16, 8, seed=1)
    h = MultivariateBernoulli(MultiLayerRelu(16, [32]), 6, None)
    h.train_mode(False)
    idx = Indexer(h, X.cuda(), F.pairwise_distance)
    Q = X[:8].clone()
    sizes = idx.bucket_sizes
    empty = int(np.nonzero(sizes == 0)[0][0]) if (sizes == 0).any() else None
    small = int(np.argmin(np.where(sizes > 0, sizes, 1 << 30)))
    big = int(np.argmax(sizes))
    rows = [[big, -1, -1], [small, -1, -1], [big, big, small], [-1, -1, -1],
            [small, big, -1], [empty if empty is not None else -1, -1, -1], [big, small, big], [63, 0, 1]]
    probes = torch.tensor(rows, dtype=torch.int32).cuda()
    k = 10
    ids, dists, ncand = idx.query_tensors(Q.cuda(), k=k, probes=probes)
    sets = [set(c for c in r if c >= 0) for r in rows]
    index2row = {int(c): idx.index2row[c].cpu().numpy() for c in idx.index2row}
    o_ids, o_d, o_n = oracle.query(X, index2row, Q, sets, "l2", k)  # sorted fallback for < k
    assert ncand.cpu().tolist() == o_n  # duplicates probed once, empty / unused probes ignored
    got_i, got_d = lists(ids, dists)
    assert_topk_equal_up_to_ties(got_i, got_d, o_ids, o_d, rtol=1e-5)
    assert got_i[3] == [] and np.isinf(dists[3].cpu().numpy()).all()
    # a row of the database queried against its own bucket finds itself at distance eps*sqrt(D)
    own = idx.hash_tensors(X[:64].cuda(), 1)
    ids, dists, _ = idx.query_tensors(X[:64].cuda(), k=1, probes=own)
    assert ids[:, 0].cpu().tolist() == list(range(64))
    np.testing.assert_allclose(dists[:, 0].cpu().numpy(), 1e-6 * np.sqrt(16), rtol=1e-4)
    # list API truncates rows with fewer than k candidates and empty input is fine
    l_ids, l_n = idx.query(Q.cuda(), k=k, probes=probes)
    assert l_ids == got_i and l_n == o_n
    assert idx.query(Q[:0].cuda()) == ([], [])
    with pytest.raises(ValueError):
        idx.query_tensors(Q.cuda(), k=129)
    with pytest.raises(ValueError):
        idx.query_tensors(torch.zeros(2, 17).cuda(), k=5)


def test_hash_times_semantics():
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    torch.manual_seed(1)
    X = mixture(6000, 16, 20, seed=2)
    h = MultivariateBernoulli(MultiLayerRelu(16, [32]), 5, None)
    idx = Indexer(h, X.cuda(), F.pairwise_distance)
    Q = X[:4200].cuda()
    sets = idx.hash(Q, hash_times=4)
    assert all(len(s) == 4 for s in sets)
    idx.compat_tail_single_probe = True  # indexer.py:52: the tail batch is hashed with n = 1
    sets = idx.hash(Q, hash_times=4)
    assert all(len(s) == 4 for s in sets[:4096]) and all(len(s) == 1 for s in sets[4096:])
    _, n1 = idx.query(Q[:100], k=5, hash_times=1)
    idx.compat_tail_single_probe = False
    _, n4 = idx.query(Q[:100], k=5, hash_times=4)
    assert all(b >= a for a, b in zip(n1, n4)) and sum(n4) > sum(n1)
