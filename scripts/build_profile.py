"""Index build (nlsh_build_csr) on a BASELINE-config-4 sized input for ncu: 10M x 128 fp32 rows, 4096
buckets, codes from a random-init hasher.  Prints the CUDA-event time of one build and its
algorithmic HBM GB/s (16 n + 2 n d 4 bytes: codes x2, ids, offsets + the row gather's read + write)."""
import sys, torch
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, ".")
from nlsh import _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d, hs = 128, 12
g = torch.Generator(device="cuda").manual_seed(1004)
x = torch.randn(n, d, device="cuda", generator=g)
codes = torch.randint(0, 1 << hs, (n,), device="cuda", generator=g, dtype=torch.int32)
_native.build_csr(codes, 1 << hs, x, want_sqnorm=True); torch.cuda.synchronize()   # warm-up (sizes the workspace)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
off, ids, xs, xn = _native.build_csr(codes, 1 << hs, x, want_sqnorm=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
c = codes[ids.long()]
same = c[1:] == c[:-1]
assert int(off[-1]) == n and bool((c[1:] >= c[:-1]).all()) and bool((ids[1:][same] > ids[:-1][same]).all())
print(f"build_csr {n} x {d}, {1 << hs} buckets: {ms:.3f} ms = {(16.0 * n + 2.0 * n * d * 4) / ms / 1e6:.0f} GB/s algorithmic")
