"""Multi-rank host logic on the CPU: world_size-2 gloo processes shard a database by rows,
search their shard (the oracle stands in for the per-shard CUDA scan — this test covers the
sharding, global-id offsets and the all-gather plumbing of nlsh/parallel.py, not kernels),
all-gather the [Q, k] lists and merge; the result must equal the unsharded search."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (os.path.join(ROOT, "neural-locality-sensitive-hashing_b200"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from nlsh.parallel import all_gather_topk, shard_range
    from oracle import nlsh_oracle as oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        X = torch.randn(501, 12, generator=g)
        Q = torch.randn(23, 12, generator=g)
        k = 7
        lo, hi = shard_range(X.shape[0], rank, world)
        ids, dd = oracle.knn_queries(Q, X[lo:hi], "l2", k)
        ids = torch.from_numpy(ids + lo)
        g_ids, g_d = all_gather_topk(ids, torch.from_numpy(dd))
        assert g_ids.shape == (world, 23, k)
        ncand = torch.tensor([hi - lo])
        dist.all_reduce(ncand)
        assert int(ncand) == X.shape[0]
        m_ids, m_d = oracle.merge_topk(g_d.numpy(), g_ids.numpy(), k)
        full_ids, full_d = oracle.knn_queries(Q, X, "l2", k)
        assert np.array_equal(m_ids, full_ids)
        np.testing.assert_allclose(m_d, full_d, rtol=1e-6)
        # the packed exchange ShardedIndexer uses: ids | dists | n_cand of a rank in ONE byte buffer, one
        # all-gather, strided [G, Q, k] views over the gathered bytes (the merge kernel reads them in place)
        from nlsh.parallel import PackedLists
        packed = PackedLists(Q.shape[0], k, world, torch.device("cpu"))
        p_ids, p_d, p_n = packed.out()
        p_ids.copy_(ids)
        p_d.copy_(torch.from_numpy(dd))
        p_n.fill_(hi - lo)
        dist.all_gather_into_tensor(packed.gathered, packed.local)
        assert torch.equal(packed.g_ids, g_ids) and torch.equal(packed.g_dists, g_d)
        assert packed.g_ncand.sum(0).tolist() == [X.shape[0]] * Q.shape[0]
        m2_ids, _ = oracle.merge_topk(packed.g_dists.numpy(), packed.g_ids.numpy(), k)
        assert np.array_equal(m2_ids, full_ids)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_search_world_size_2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
