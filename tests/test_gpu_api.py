"""End-to-end drop-in checks on the GPU: the call sequence of nlsh/trainers/base.py:80-115
against the oracle's CPU flow, the sharded search, and BASELINE-size properties."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import assert_topk_equal_up_to_ties, mixture, oracle_layers_from_hashing

pytestmark = pytest.mark.gpu


def test_trainer_validation_block_flow(oracle):
    # what Trainer.fit does every test_every_updates steps (base.py:80-115)
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    from nlsh.metrics import calculate_recall
    import precompute
    torch.manual_seed(3)
    n, d, hs, nq, K = 30000, 64, 6, 500, 10
    X = mixture(n, d, 256, seed=11)
    Qv = mixture(nq, d, 256, seed=11) + 0.05 * torch.randn(nq, d, generator=torch.Generator().manual_seed(1))
    hashing = MultivariateBernoulli(MultiLayerRelu(d, [64, 64]), hs, F.pairwise_distance)
    hashing.train_mode(False)
    indexer = Indexer(hashing, X.cuda(), F.pairwise_distance)
    n_indexes = len(indexer.index2row)
    std_index_rows = np.std([len(idxs) for idxs in indexer.index2row.values()])
    assert 1 <= n_indexes <= 64 and std_index_rows >= 0
    assert sum(len(v) for v in indexer.index2row.values()) == n
    recalls, n_candidates = indexer.query(Qv.cuda(), k=K, hash_times=1)
    gt, _ = precompute.knn_tensors(Qv.cuda(), X.cuda(), "l2", K)
    current_recall = calculate_recall(gt.cpu().tolist(), recalls, np.mean)
    # the same flow on the CPU through the oracle, from the same weights
    layers = oracle_layers_from_hashing(hashing, oracle)
    cpu = oracle.CpuIndexer(layers, oracle.HEAD_SIGMOID, X, "l2")
    c_ids, c_ncand = cpu.query(Qv, k=K)
    agree = np.mean([a == b for a, b in zip(n_candidates, c_ncand)])
    assert agree >= 0.99  # a query lands in another bucket only if a logit sits at a rounding tie
    same = [i for i in range(nq) if n_candidates[i] == c_ncand[i] and c_ncand[i] >= K]
    assert np.mean([recalls[i] == c_ids[i] for i in same]) >= 0.99
    o_gt, _ = oracle.knn_queries(Qv, X, "l2", K)
    cpu_recall = oracle.recall(o_gt.tolist(), c_ids)
    assert abs(current_recall - cpu_recall) < 2e-3
    assert 0.0 < current_recall <= 1.0


def test_sharded_search_equals_single_index(oracle):
    # emulate G shards in one process (gloo covers the collective plumbing on the CPU, the
    # driver's N-GPU bench the NCCL path): per-shard Indexer with id_offset + merge kernel
    from encoders import MultiLayerRelu
    from nlsh import _native
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    from nlsh.parallel import shard_range
    torch.manual_seed(5)
    n, d, hs, nq, k, p = 40001, 32, 6, 300, 10, 3
    X = mixture(n, d, 200, seed=21).cuda()
    Q = mixture(nq, d, 200, seed=21).cuda()
    hashing = MultivariateBernoulli(MultiLayerRelu(d, [32]), hs, None)
    hashing.train_mode(False)
    full = Indexer(hashing, X, None, metric="l2")
    f_ids, f_d, f_n = full.query_tensors(Q, k=k, hash_times=p)
    for G in (2, 4, 8):
        parts = []
        for r in range(G):
            lo, hi = shard_range(n, r, G)
            shard = Indexer(hashing, X[lo:hi], None, metric="l2", id_offset=lo)
            parts.append(shard.query_tensors(Q, k=k, hash_times=p))
        g_ids = torch.stack([t[0] for t in parts])
        g_d = torch.stack([t[1] for t in parts])
        m_ids, m_d = _native.merge_topk(g_d, g_ids)
        assert torch.equal(m_ids, f_ids) and torch.equal(m_d, f_d)
        assert torch.equal(sum(t[2] for t in parts), f_n)
        o_ids, o_d = oracle.merge_topk(g_d.cpu().numpy(), g_ids.cpu().numpy(), k)
        assert np.array_equal(o_ids, m_ids.cpu().numpy())


@pytest.mark.parametrize("metric,d,k", [("l2", 64, 10), ("angular", 100, 10), ("l2", 200, 40)])
def test_bounds_seeded_on_one_shard_serve_every_shard(monkeypatch, metric, d, k):
    """nlsh_query_seed_tau / nlsh_query_scan_topk_seeded: a distance bound computed from a sample of ONE shard's
    rows bounds the k-th best distance over the whole database, so every shard may filter with it (a row-sharded
    search seeds each query on one rank only).  The merged result must not change - whichever shard seeds."""
    from encoders import MultiLayerRelu
    from nlsh import _native
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    from nlsh.parallel import shard_range
    monkeypatch.setenv("NLSH_SCAN_IMPL", "tc")
    torch.manual_seed(9)
    n, hs, nq, p, G = 60000, 6, 700, 4, 4
    X = mixture(n, d, 300, seed=31).cuda()
    Q = (mixture(nq, d, 300, seed=31) + 0.1 * torch.randn(nq, d, generator=torch.Generator().manual_seed(3))).cuda()
    hashing = MultivariateBernoulli(MultiLayerRelu(d, [48]), hs, None)
    hashing.train_mode(False)
    full = Indexer(hashing, X, None, metric=metric)
    probes = full.hash_tensors(Q, p)
    f_ids, f_d, f_n = full.query_tensors(Q, k=k, probes=probes)
    shards = []
    for r in range(G):
        lo, hi = shard_range(n, r, G)
        shards.append(Indexer(hashing, X[lo:hi], None, metric=metric, id_offset=lo))
    for seeder, rows in ((0, 0), (G - 1, 0), (1, 192)):  # rows: the caller's sample size (nlsh_query_seed_tau_rows)
        tau = shards[seeder].seed_tau_tensors(Q, probes, k, sample_rows=rows)
        assert tau.shape == (nq,) and tau.dtype == torch.float32
        if rows:  # a larger sample can only tighten the bound
            assert (tau <= shards[seeder].seed_tau_tensors(Q, probes, k)).all()
        # a valid bound: at least the exact k-th best distance over the whole database (squared for L2)
        kth = f_d[:, k - 1] ** 2 if metric == "l2" else f_d[:, k - 1]
        assert (tau >= kth).all()
        for s in shards:  # shard lists carry squared L2 distances, the root follows the shard merge
            s.scan_flags = _native.FLAG_SQUARED_L2_OUT if metric == "l2" else 0
        parts = [s.query_tensors(Q, k=k, probes=probes, tau_seed=tau) for s in shards]
        m_ids, m_d = _native.merge_topk(torch.stack([t[1] for t in parts]), torch.stack([t[0] for t in parts]))
        if metric == "l2":
            m_d = m_d.sqrt()
        assert torch.equal(m_ids, f_ids) and torch.equal(m_d, f_d)
        assert torch.equal(sum(t[2] for t in parts), f_n)


@pytest.mark.parametrize("k", [1, 10, 33, 100, 128])
def test_merge_kernel_against_oracle(oracle, k):
    from nlsh import _native
    g = torch.Generator().manual_seed(k)
    G, nq = 5, 77
    d = torch.rand(G, nq, k, generator=g).sort(dim=2)[0]
    d[:, :, k // 2:] = torch.round(d[:, :, k // 2:] * 8) / 8  # force cross-list ties
    d = d.sort(dim=2)[0]
    ids = torch.randint(0, 1 << 40, (G, nq, k), generator=g)
    ids[2, :, k - k // 3:] = -1  # short lists
    d[2, :, k - k // 3:] = float("inf")
    ids[4, 5] = -1
    d[4, 5] = float("inf")
    m_ids, m_d = _native.merge_topk(d.cuda(), ids.cuda())
    o_ids, o_d = oracle.merge_topk(d.numpy(), ids.numpy(), k)
    assert np.array_equal(m_ids.cpu().numpy(), o_ids)
    assert np.array_equal(m_d.cpu().numpy(), o_d)


def test_full_size_properties_config2():
    # BASELINE config 2 at full size (1M x 128, 256 buckets, p in {1, 4, 16}, k = 10), checked
    # through size-independent properties: candidates counted exactly, more probes never hurt,
    # every returned id is a true candidate with the exact distance, recall monotone in p and
    # equal to 1 when every bucket is probed.
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    from nlsh.metrics import recall_at_k_tensors
    import precompute
    torch.manual_seed(7)
    n, d, hs, nq, k = 1_000_000, 128, 8, 2000, 10
    g = torch.Generator(device="cuda").manual_seed(1002)
    centers = torch.randn(1024, d, generator=g, device="cuda") * 2
    X = centers[torch.randint(0, 1024, (n,), generator=g, device="cuda")] + torch.randn(n, d, generator=g, device="cuda")
    Q = centers[torch.randint(0, 1024, (nq,), generator=g, device="cuda")] + torch.randn(nq, d, generator=g, device="cuda")
    hashing = MultivariateBernoulli(MultiLayerRelu(d, [256, 256]), hs, F.pairwise_distance)
    hashing.train_mode(False)
    idx = Indexer(hashing, X, F.pairwise_distance)
    assert int(idx.bucket_sizes.sum()) == n
    gt, gt_d = precompute.knn_tensors(Q, X, "l2", k)
    sizes = torch.from_numpy(idx.bucket_sizes).cuda()
    prev_recall, prev_d = -1.0, None
    for p in (1, 4, 16):
        probes = idx.hash_tensors(Q, p)
        ids, dists, ncand = idx.query_tensors(Q, k=k, probes=probes)
        assert torch.equal(ncand.long(), sizes[probes.long()].sum(1))
        assert (ids >= 0).all()
        exact = F.pairwise_distance(Q[:, None, :].expand(-1, k, -1).reshape(-1, d), X[ids.reshape(-1)]).view(nq, k)
        torch.testing.assert_close(dists, exact, rtol=1e-5, atol=0)
        codes_of_ids = idx._hashing.hash_tensors(X[ids.reshape(-1)], 1)[0].view(nq, k)
        assert (codes_of_ids[:, :, None] == probes[:, None, :]).any(-1).all()
        rec = recall_at_k_tensors(gt, ids)
        assert rec >= prev_recall - 1e-9
        if prev_d is not None:
            assert (dists <= prev_d * (1 + 1e-6)).all()
        prev_recall, prev_d = rec, dists
    all_probes = torch.arange(256, dtype=torch.int32, device="cuda")[None, :].expand(64, -1).contiguous()
    ids, dists, ncand = idx.query_tensors(Q[:64], k=k, probes=all_probes)
    assert (ncand == n).all()
    torch.testing.assert_close(dists, gt_d[:64], rtol=1e-5, atol=0)
    assert recall_at_k_tensors(gt[:64], ids) >= 0.999


def test_graphed_query_equals_eager():
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    torch.manual_seed(9)
    X = mixture(50000, 64, 300, seed=31).cuda()
    hashing = MultivariateBernoulli(MultiLayerRelu(64, [64, 64]), 7, None)
    hashing.train_mode(False)
    idx = Indexer(hashing, X, None, metric="l2")
    graphed = idx.capture_query(777, k=10, hash_times=4)
    assert graphed.kernels_per_replay >= 8
    for seed in (1, 2, 3):
        Q = mixture(777, 64, 300, seed=31 + seed).cuda()
        e_ids, e_d, e_n = idx.query_tensors(Q, k=10, hash_times=4)
        g_ids, g_d, g_n = graphed(Q)
        assert torch.equal(e_ids, g_ids) and torch.equal(e_d, g_d) and torch.equal(e_n, g_n)
    # pinned host input is copied into the graph's static buffer
    Qh = mixture(777, 64, 300, seed=77).pin_memory()
    g_ids, _, _ = graphed(Qh)
    assert torch.equal(g_ids, idx.query_tensors(Qh.cuda(), k=10, hash_times=4)[0])


def test_packed_exchange_buffer_and_strided_merge(oracle):
    # PackedLists: the [ids | dists | n_cand] byte buffer of one rank and the in-place strided
    # views over the all-gathered buffers (the collective itself is covered by the gloo test
    # and the multi-GPU bench)
    from nlsh import _native
    from nlsh.parallel import PackedLists
    G, nq, k = 3, 41, 10
    g = torch.Generator().manual_seed(3)
    packs = [PackedLists(nq, k, G, torch.device("cuda")) for _ in range(G)]
    all_d = torch.rand(G, nq, k, generator=g).sort(dim=2)[0]
    all_i = torch.randint(0, 1 << 33, (G, nq, k), generator=g)
    all_n = torch.randint(0, 1000, (G, nq), generator=g, dtype=torch.int32)
    for r, pk in enumerate(packs):
        pk.ids.copy_(all_i[r]); pk.dists.copy_(all_d[r]); pk.ncand.copy_(all_n[r])
    gathered = torch.cat([pk.local for pk in packs])  # what all_gather_into_tensor produces
    packs[0].gathered.copy_(gathered)
    m_ids, m_d, m_n = _native.merge_topk(packs[0].g_dists, packs[0].g_ids, packs[0].g_ncand)
    o_ids, o_d = oracle.merge_topk(all_d.numpy(), all_i.numpy(), k)
    assert np.array_equal(m_ids.cpu().numpy(), o_ids) and np.array_equal(m_d.cpu().numpy(), o_d)
    assert torch.equal(m_n.cpu(), all_n.sum(0).int())


def test_validation_block_and_recall_sweep(oracle):
    """the Trainer.fit validation block (base.py:80-115) and the eval.py sweep,
    driven through the CUDA Indexer; recall must equal metrics.py on the oracle's CPU flow."""
    import precompute
    from encoders import MultiLayerRelu
    from helpers import mixture, oracle_layers_from_hashing
    from nlsh.evaluation import recall_sweep, validate_index
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.metrics import calculate_recall
    torch.manual_seed(3)
    X = mixture(30000, 64, 48, seed=5)
    Q = mixture(400, 64, 48, seed=5) + 0.05 * torch.randn(400, 64, generator=torch.Generator().manual_seed(1))
    h = MultivariateBernoulli(MultiLayerRelu(64, [64, 64]), 6, F.pairwise_distance)
    gt, _ = precompute.knn_tensors(Q.cuda(), X.cuda(), "l2", 10)

    class Log:
        def __init__(self):
            self.rows = {}

        def log(self, name, value, step):
            self.rows[name] = (value, step)

    log = Log()
    res = validate_index(h, X.cuda(), F.pairwise_distance, Q.cuda(), gt.cpu().numpy(), k=10, hash_times=4,
                         logger=log, global_step=300)
    assert set(log.rows) == {"test/n_indexes", "test/std_index_rows", "test/recall", "test/query_size", "test/qps"}
    assert log.rows["test/recall"][1] == 300 and res["test/qps"] > 0
    idx = res["indexer"]
    # the same numbers from the reference's flow on the CPU (oracle), fed the same probe sets
    probes = idx.hash_tensors(Q.cuda(), 4).cpu().numpy()
    sets = [set(int(c) for c in row if c >= 0) for row in probes]
    index2row = {int(c): idx.index2row[c].cpu().numpy() for c in idx.index2row}
    o_ids, _, o_n = oracle.query(X, index2row, Q, sets, "l2", 10)
    want = calculate_recall(gt.cpu().tolist(), o_ids, np.mean)
    assert res["test/recall"] == pytest.approx(want, abs=1e-9)
    assert res["test/query_size"] == pytest.approx(float(np.mean(o_n)))
    assert res["test/n_indexes"] == len(index2row)
    fast = validate_index(h, X.cuda(), F.pairwise_distance, Q.cuda(), gt.cpu().numpy(), k=10, hash_times=4,
                          list_api=False)
    assert fast["test/recall"] == pytest.approx(want, abs=1e-6)
    rows = recall_sweep(idx, Q.cuda(), gt.cpu().numpy(), k=10, probe_counts=(1, 2, 4, 8, 128))
    assert [r["probes"] for r in rows] == [1, 2, 4, 8]
    assert all(b["recall"] >= a["recall"] - 1e-9 and b["avg_n_candidates"] >= a["avg_n_candidates"]
               for a, b in zip(rows, rows[1:]))
    assert rows[2]["recall"] == pytest.approx(want, abs=1e-6)


def test_pipelined_search_matches_direct_calls():
    """nlsh.parallel.PipelinedSearch: double-buffered host-to-host loop == direct query_tensors calls."""
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.parallel import PipelinedSearch, ShardedIndexer
    torch.manual_seed(5)
    X = mixture(40000, 32, 40, seed=9)
    h = MultivariateBernoulli(MultiLayerRelu(32, [64]), 5, F.pairwise_distance)
    h.train_mode(False)
    index = ShardedIndexer(h, X.cuda(), F.pairwise_distance, shard_lo=0)
    nq, k, p = 512, 10, 4
    batches = [(mixture(nq, 32, 40, seed=9) + 0.01 * i).pin_memory() for i in range(5)]
    pipe = PipelinedSearch(index, nq, k=k, hash_times=p)
    tickets, got = [], []
    for b in batches:
        tickets.append(pipe.submit(b))
        if len(tickets) > 1:
            ids, dists, ncand = pipe.result(tickets.pop(0))
            got.append((ids.clone(), dists.clone(), ncand.clone()))
    ids, dists, ncand = pipe.result(tickets.pop(0))
    got.append((ids.clone(), dists.clone(), ncand.clone()))
    assert len(got) == len(batches)
    for b, (ids, dists, ncand) in zip(batches, got):
        w_ids, w_d, w_n = index.query_tensors(b.cuda(), k=k, hash_times=p)
        assert torch.equal(ids, w_ids.cpu()) and torch.equal(dists, w_d.cpu()) and torch.equal(ncand, w_n.cpu())


def test_full_size_config4_filter_equals_simt():
    """BASELINE config 4 at full size on one GPU (10M x 128, 4096 buckets, 10k queries, p = 8, k = 10):
    the tensor-core filtered scan and the fp32 SIMT scan must return identical bits, every id must be
    a row of a probed bucket, and the candidate counts must be the probed buckets' sizes."""
    import synth
    from encoders import MultiLayerRelu
    from nlsh import _native
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    n, d, hs, nq, k, p = 10_000_000, 128, 12, 10_000, 10, 8
    dev = torch.device("cuda")
    X = synth.make_database(n, d, hs, 1004, dev)
    Q = synth.make_queries(nq, d, hs, 1004, dev)
    torch.manual_seed(1004)
    hashing = MultivariateBernoulli(MultiLayerRelu(d, [256, 256]), hs, F.pairwise_distance)
    synth.fit_hasher(hashing, d, hs, 1004, dev, steps=60)
    idx = Indexer(hashing, X, F.pairwise_distance)
    assert _native.scan_impl(d, k, idx._metric, True, nq, p, 1 << hs) == 1
    probes = idx.hash_tensors(Q, p)
    ids, dists, ncand = idx.query_tensors(Q, k=k, probes=probes)
    idx.scan_flags = 2  # fp32 SIMT kernel
    ids2, dists2, ncand2 = idx.query_tensors(Q, k=k, probes=probes)
    assert torch.equal(ids, ids2) and torch.equal(dists, dists2) and torch.equal(ncand, ncand2)
    sizes = torch.from_numpy(idx.bucket_sizes).to(dev)
    assert torch.equal(ncand.long(), sizes[probes.long()].sum(1))
    sample = torch.arange(0, nq, 10, device=dev)
    got = ids[sample]
    exact = F.pairwise_distance(Q[sample][:, None, :].expand(-1, k, -1).reshape(-1, d),
                                X[got.reshape(-1)]).view(-1, k)
    torch.testing.assert_close(dists[sample], exact, rtol=1e-5, atol=0)
    codes_of_ids = hashing.hash_tensors(X[got.reshape(-1)], 1)[0].view(-1, k)
    assert (codes_of_ids[:, :, None] == probes[sample][:, None, :]).any(-1).all()
    assert (dists[:, 1:] >= dists[:, :-1]).all()
