"""A/B in one process: rows per chunk of the tensor-core hasher (NLSH_MLP_CHUNK) on a 10M x 128 forward."""
import os, sys, json, torch
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, ".")
from encoders import MultiLayerRelu
from nlsh import _native
from nlsh.hashings import MultivariateBernoulli
torch.manual_seed(0)
h = MultivariateBernoulli(MultiLayerRelu(128, [256, 256]), 12, None); h.train_mode(False)
x = torch.randn(10_000_000, 128, device="cuda")
res, codes = {}, {}
for rnd in range(3):
    for chunk in ("default", "32768", "18944", "56832"):
        if chunk == "default": os.environ.pop("NLSH_MLP_CHUNK", None)
        else: os.environ["NLSH_MLP_CHUNK"] = chunk
        _native.release_workspaces()
        h.hash_tensors(x[:100000], 1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = h.hash_tensors(x, 1); e1.record(); torch.cuda.synchronize()
        res.setdefault(chunk, []).append(round(e0.elapsed_time(e1), 2)); codes[chunk] = out[0]
res["codes_equal"] = all(torch.equal(codes["default"], c) for c in codes.values())
print(json.dumps(res))
