import sys, os
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, "tests"); sys.path.insert(0, ".")
os.environ["NLSH_SCAN_IMPL"] = "tc"
import torch
from helpers import mixture
from encoders import MultiLayerRelu
from nlsh import _native
from nlsh.hashings import MultivariateBernoulli
from nlsh.indexer import Indexer
from nlsh.parallel import shard_range
for metric, d, k in (("l2", 200, 40), ("l2", 64, 40), ("l2", 200, 10)):
    torch.manual_seed(9)
    n, hs, nq, p, G = 60000, 6, 700, 4, 4
    X = mixture(n, d, 300, seed=31).cuda()
    Q = (mixture(nq, d, 300, seed=31) + 0.1 * torch.randn(nq, d, generator=torch.Generator().manual_seed(3))).cuda()
    hashing = MultivariateBernoulli(MultiLayerRelu(d, [48]), hs, None)
    hashing.train_mode(False)
    full = Indexer(hashing, X, None, metric=metric)
    probes = full.hash_tensors(Q, p)
    f_ids, f_d, f_n = full.query_tensors(Q, k=k, probes=probes)
    full.scan_flags = 2
    s_ids, s_d, _ = full.query_tensors(Q, k=k, probes=probes)
    print(metric, d, k, "full tc==simt", torch.equal(f_ids, s_ids), torch.equal(f_d, s_d))
    shards = []
    for r in range(G):
        lo, hi = shard_range(n, r, G)
        shards.append(Indexer(hashing, X[lo:hi], None, metric=metric, id_offset=lo))
    for name, tau in (("unseeded", None), ("seed0", shards[0].seed_tau_tensors(Q, probes, k)), ("seed3", shards[3].seed_tau_tensors(Q, probes, k))):
        parts = [s.query_tensors(Q, k=k, probes=probes, tau_seed=tau) for s in shards]
        m_ids, m_d = _native.merge_topk(torch.stack([t[1] for t in parts]), torch.stack([t[0] for t in parts]))
        bad = (m_ids != f_ids).any(1).nonzero().flatten().tolist()
        print(" ", name, "ids equal", torch.equal(m_ids, f_ids), "d equal", torch.equal(m_d, f_d), "bad queries", len(bad))
        for q in bad[:3]:
            pos = (m_ids[q] != f_ids[q]).nonzero().flatten().tolist()
            print("    q", q, "pos", pos, "merged", m_ids[q, pos].tolist(), m_d[q, pos].tolist(), "full", f_ids[q, pos].tolist(), f_d[q, pos].tolist(),
                  "missing", sorted(set(f_ids[q].tolist()) - set(m_ids[q].tolist())))
