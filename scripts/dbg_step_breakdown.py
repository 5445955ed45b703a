"""Warm timing (CUDA events, eager launches) of the pieces of one query step on a bench workload."""
import os, sys, torch
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, ".")
import bench, synth
from nlsh import _native
from nlsh.indexer import Indexer
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4_10m_x128_4096b"
n, d, hs, nq, k, metric, seed = bench.WORKLOADS[wl]
p = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else n
dev = torch.device("cuda")
X = synth.make_database(n, d, hs, seed, dev, sep=bench.SEP); Q = synth.make_queries(nq, d, hs, seed, dev, sep=bench.SEP)
hashing, _ = bench.make_hashing(d, hs, metric, seed, dev, 300)
idx = Indexer(hashing, X, hashing.distance, metric=metric)
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
layers, head = hashing.layer_specs(), hashing.head
logits, codes = _native.mlp_hash(Q, layers, head)
probes = _native.topp_probes(logits, head, p)
print("mlp_hash (3 layers + splits)   %.3f ms" % timed(lambda: _native.mlp_hash(Q, layers, head)))
print("topp_probes                    %.3f ms" % timed(lambda: _native.topp_probes(logits, head, p)))
print("query_scan_topk (all kernels)  %.3f ms" % timed(lambda: idx.query_tensors(Q, k=k, probes=probes)))
_native.profile_enable(True)
for _ in range(20): idx.query_tensors(Q, k=k, probes=probes)
ms = _native.profile_read(); _native.profile_enable(False)
print("  of which the scan kernel     %.3f ms" % (sum(ms) / len(ms)))
os.environ["NLSH_SCAN_SEED"] = "0"; os.environ["_x"] = "1"
t0 = timed(lambda: idx.query_tensors(Q, k=k, probes=probes))
_native.profile_enable(True)
for _ in range(20): idx.query_tensors(Q, k=k, probes=probes)
ms0 = _native.profile_read(); _native.profile_enable(False)
print("  without seeding: call %.3f ms, scan kernel %.3f ms -> seed kernel ~ %.3f ms" % (t0, sum(ms0)/len(ms0), 0))
os.environ.pop("NLSH_SCAN_SEED")
g = idx.capture_query(nq, k=k, hash_times=p)
print("whole step, CUDA graph         %.3f ms" % timed(lambda: g(Q)))
