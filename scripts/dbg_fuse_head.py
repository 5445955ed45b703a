"""A/B in one process, interleaved: hasher forward with the output layer fused into the epilogue of the layer before it
(NLSH_MLP_FUSE_HEAD unset) or as a tcgen05 layer of its own (=0), for an index-build sized input and a query batch."""
import os, sys, json, torch
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, ".")
from encoders import MultiLayerRelu
from nlsh.hashings import MultivariateBernoulli
torch.manual_seed(0)
h = MultivariateBernoulli(MultiLayerRelu(128, [256, 256]), 12, None); h.train_mode(False)
def timed(x, reps):
    h.hash_tensors(x, 1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = h.hash_tensors(x, 1, want_logits=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out
for n, reps in ((4_000_000, 5), (10_000, 200), (1_250, 200)):
    x = torch.randn(n, 128, device="cuda")
    res = {"rows": n}
    for rnd in range(3):
        for mode in ("fused", "unfused"):
            if mode == "unfused": os.environ["NLSH_MLP_FUSE_HEAD"] = "0"
            else: os.environ.pop("NLSH_MLP_FUSE_HEAD", None)
            ms, out = timed(x, reps)
            res.setdefault(mode, []).append(round(ms, 4))
            res[mode + "_codes"] = out[0]
    res["codes_equal"] = float((res.pop("fused_codes") == res.pop("unfused_codes")).float().mean())
    print(json.dumps(res), flush=True)
