"""Short brute-force kNN run for ncu: 10k queries x 2M rows x 128, l2sq, k from argv (default 10)."""
import sys, time, torch
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, ".")
import precompute, synth
k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
X = synth.make_database(n, 128, 12, 1004, "cuda")
Q = synth.make_queries(10_000, 128, 12, 1004, "cuda")
precompute.knn_tensors(Q, X, "l2sq", k); torch.cuda.synchronize()
t = time.perf_counter(); precompute.knn_tensors(Q, X, "l2sq", k); torch.cuda.synchronize()
dt = time.perf_counter() - t
print(f"knn 10k x {n} x 128 l2sq k={k}: {dt*1e3:.1f} ms ({2*1e4*n*128/dt/1e12:.1f} algorithmic TFLOP/s)")
