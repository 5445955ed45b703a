"""Survivor statistics of the tensor-core scan on a bench workload (NLSH_TC_STATS=1)."""
import os, sys, torch
os.environ["NLSH_TC_STATS"] = "1"
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, ".")
import bench, synth
from nlsh import _native
from nlsh.indexer import Indexer
n, d, hs, nq, k, metric, seed = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4_10m_x128_4096b"]
p = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else n
dev = torch.device("cuda")
X = synth.make_database(n, d, hs, seed, dev, sep=bench.SEP); Q = synth.make_queries(nq, d, hs, seed, dev, sep=bench.SEP)
hashing, _ = bench.make_hashing(d, hs, metric, seed, dev, 300)
idx = Indexer(hashing, X, hashing.distance, metric=metric)
for rows in (128, 256):
    os.environ["NLSH_SCAN_SEED"] = str(rows)
    ids, dd, nc = idx.query_tensors(Q, k=k, hash_times=p); torch.cuda.synchronize()
    ws = list(_native._workspaces.values())[0].buf
    B = 1 << hs
    st = ws.view(torch.uint8)[(2 * B + 2) * 4:(2 * B + 2) * 4 + 18 * 8].view(torch.int64).cpu().tolist()
    pairs = int(nc.long().sum())
    names = ["producer: wait meta_empty, wait slot empty, wait q_empty, total",
             "mma: wait acc_empty, wait slot full, -, total",
             "filter: wait meta_full, wait acc_full, wait surv_empty, total",
             "rerank: wait surv_full, batch cycles, batches, total"]
    for r in range(4):
        print("   ", names[r], st[2 + 4 * r: 6 + 4 * r])
    print(f"seed_rows={rows}: candidates(pairs)={pairs} survivors={st[0]} ({st[0]/pairs*100:.2f}%) batches={st[1]} per query={st[0]/nq:.0f}")
