import sys, time, torch
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, ".")
import precompute, synth
X = synth.make_database(10_000_000, 128, 12, 1004, "cuda")
Q = synth.make_queries(10_000, 128, 12, 1004, "cuda")
for metric, k in (("l2", 10), ("l2sq", 10), ("l2sq", 100)):
    precompute.knn_tensors(Q[:256], X, metric, k); torch.cuda.synchronize()
    t = time.perf_counter(); precompute.knn_tensors(Q, X, metric, k); torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"knn 10k x 10M x 128 {metric} k={k}: {dt*1e3:.1f} ms  ({2*1e4*1e7*128/dt/1e12:.1f} algorithmic TFLOP/s)")
Xs = X[:1_000_000]
t = time.perf_counter(); precompute.knn_tensors(Xs, Xs, "l2sq", 101); torch.cuda.synchronize()
dt = time.perf_counter() - t
print(f"self-knn 1M x 1M x 128 k=101: {dt:.2f} s ({2*1e6*1e6*128/dt/1e12:.1f} TFLOP/s)")
