"""SIREN trunk (encoders.py:58-79, what main.py:388 builds) through the CUDA hasher: logit error and
bucket agreement of both kernels (NLSH_MLP_IMPL=simt, tcgen05 3xTF32) against an fp64 evaluation of the
same network, next to the error the fp32 CPU oracle itself has against fp64.  One JSON line per case.

    python scripts/dbg_siren.py > gpurun_out/siren.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "neural-locality-sensitive-hashing_b200")
sys.path.insert(0, ".")
from encoders import Siren  # noqa: E402
from nlsh import _native  # noqa: E402
from nlsh.hashings import MultivariateBernoulli, extract_layer_specs  # noqa: E402
from oracle import nlsh_oracle as oracle  # noqa: E402

torch.set_grad_enabled(False)


def fp64_logits(x, specs):
    h = x.double()
    for w, b, act, scale in specs:
        h = h @ w.double().cpu().T
        if b is not None:
            h = h + b.double().cpu()
        if act == _native.ACT_RELU:
            h = torch.relu(h)
        elif act == _native.ACT_SIN:
            h = torch.sin(scale * h)
    return h.numpy()


def rel_err(got, ref):
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
    return float((np.abs(got - ref) / scale).max())


for dims, d, hs, xs in (([256, 256, 64], 128, 12, 0.05), ([256, 256, 64], 128, 12, 1.0),
                        ([256, 256], 128, 12, 0.05), ([256, 256, 64], 100, 10, 0.05),
                        ([256, 256, 64], 960, 9, 0.02)):
    torch.manual_seed(2)
    hashing = MultivariateBernoulli(Siren(d, dims), hs, None)
    hashing.train_mode(False)
    X = torch.randn(20000, d) * xs
    hasher = hashing._hasher
    hasher.cuda()
    specs = extract_layer_specs(hasher._encoder, hasher.output_layer)
    ref64 = fp64_logits(X, specs)
    layers = [oracle.Layer(w.cpu(), None if b is None else b.cpu(), act == _native.ACT_RELU,
                           scale if act == _native.ACT_SIN else None) for w, b, act, scale in specs]
    ref32 = oracle.mlp_logits(X, layers).numpy()
    rec = {"trunk": f"Siren({d},{dims})", "hash_size": hs, "x_scale": xs, "rows": X.shape[0],
           "oracle_fp32_vs_fp64": rel_err(ref32, ref64)}
    for impl in ("simt", "tc"):
        os.environ["NLSH_MLP_IMPL"] = impl  # "tc" forces the tensor cores, which sine trunks do not take by default
        codes, _, logits = hashing.hash_tensors(X.cuda(), 1, want_logits=True)
        got = logits.cpu().numpy()
        rec[f"{impl}_vs_fp64"] = rel_err(got, ref64)
        rec[f"{impl}_vs_oracle"] = rel_err(got, ref32)
        rec[f"{impl}_bucket_agreement_vs_oracle"] = float(
            (oracle.hard_codes(torch.from_numpy(ref32), oracle.HEAD_SIGMOID) == codes.cpu().numpy()).mean())
        rec[f"{impl}_codes_from_own_logits_exact"] = bool(np.array_equal(
            oracle.hard_codes(logits.cpu(), oracle.HEAD_SIGMOID), codes.cpu().numpy()))
    os.environ.pop("NLSH_MLP_IMPL", None)
    print(json.dumps(rec), flush=True)
