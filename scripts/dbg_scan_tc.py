import os, sys, time, torch
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, "tests")
from helpers import mixture
from encoders import MultiLayerRelu
from nlsh.hashings import MultivariateBernoulli
from nlsh.indexer import Indexer
n, d, hs, nq, p, k = [int(v) for v in sys.argv[1:7]]
torch.manual_seed(0)
X = mixture(n, d, 3 << hs, seed=n); Q = mixture(nq, d, 3 << hs, seed=n)
h = MultivariateBernoulli(MultiLayerRelu(d, [64, 64]), hs, None); h.train_mode(False)
idx = Indexer(h, X.cuda(), None, metric="l2")
probes = idx.hash_tensors(Q.cuda(), p)
torch.cuda.synchronize(); t = time.time()
try:
    ids, dists, nc = idx.query_tensors(Q.cuda(), k=k, probes=probes); torch.cuda.synchronize()
    print("tc ok", time.time() - t)
    idx.scan_flags = 2
    ids2, dists2, _ = idx.query_tensors(Q.cuda(), k=k, probes=probes); torch.cuda.synchronize()
    print("equal", torch.equal(ids, ids2), torch.equal(dists, dists2))
except Exception as e:
    print("FAILED after", time.time() - t, str(e)[:200])
