#!/usr/bin/env python
"""Hasher-forward + index-build throughput (not the headline metric): rows/s and algorithmic
TFLOP/s of nlsh_mlp_hash_f32 on the tensor-core path vs the fp32 SIMT path, and the CSR build."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-locality-sensitive-hashing_b200"))
sys.path.insert(0, ROOT)
from encoders import MultiLayerRelu  # noqa: E402
from nlsh import _native  # noqa: E402
from nlsh.hashings import MultivariateBernoulli  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    out = {}
    for d, hs in ((128, 12), (960, 9)):
        rows = n if d == 128 else n // 10
        torch.manual_seed(0)
        x = torch.randn(rows, d, device="cuda")
        h = MultivariateBernoulli(MultiLayerRelu(d, [256, 256]), hs, None)
        h.train_mode(False)
        flops = 2.0 * (d * 256 + 256 * 256 + 256 * hs) * rows
        res = {}
        for impl in ("tc", "simt"):
            if impl == "simt":
                os.environ["NLSH_MLP_IMPL"] = "simt"
            else:
                os.environ.pop("NLSH_MLP_IMPL", None)
            ms = timed(lambda: h.hash_tensors(x, 1))
            res[impl] = {"ms": ms, "rows_per_s": rows / ms * 1e3, "algorithmic_tflops": flops / ms / 1e9,
                         "x_read_gbs": rows * d * 4 / ms / 1e6}
        os.environ.pop("NLSH_MLP_IMPL", None)
        codes_tc = h.hash_tensors(x, 1)[0]
        os.environ["NLSH_MLP_IMPL"] = "simt"
        codes_simt = h.hash_tensors(x, 1)[0]
        os.environ.pop("NLSH_MLP_IMPL", None)
        res["bucket_agreement_tc_vs_simt"] = float((codes_tc == codes_simt).float().mean())
        ms_build = timed(lambda: _native.build_csr(codes_tc, 1 << hs, x))
        res["build_csr_ms"] = ms_build
        res["build_csr_gbs"] = (3.0 * rows * d * 4 + 16.0 * rows - rows * d * 4) / ms_build / 1e6
        out[f"{rows}x{d}_hs{hs}"] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
