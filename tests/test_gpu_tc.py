"""The tensor-core (tcgen05, 3xTF32 split) hasher path against fp64 / the fp32 SIMT path."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_mlp(x, weights, biases, head=0):
    from nlsh import _native
    layers = [_native.LayerSpec(w, b, _native.ACT_RELU if i < len(weights) - 1 else _native.ACT_IDENTITY)
              for i, (w, b) in enumerate(zip(weights, biases))]
    return _native.mlp_hash(x, layers, head)


@pytest.mark.parametrize("m,dims", [(128, [32, 12]), (1, [128, 256, 256, 12]), (300, [64, 256, 15]),
                                    (5000, [128, 256, 256, 12]), (40000, [100, 256, 256, 10]),
                                    (1000, [960, 256, 256, 9]), (777, [128, 64, 64, 8]), (129, [36, 100, 4])])
def test_tc_path_matches_fp64_and_simt(m, dims):
    g = torch.Generator().manual_seed(m + sum(dims))
    x = torch.randn(m, dims[0], generator=g)
    ws = [torch.randn(o, i, generator=g) / np.sqrt(i) for i, o in zip(dims[:-1], dims[1:])]
    bs = [torch.randn(o, generator=g) * 0.1 for o in dims[1:]]
    ref = x.double()
    for i, (w, b) in enumerate(zip(ws, bs)):
        ref = ref @ w.double().T + b.double()
        if i < len(ws) - 1:
            ref = ref.relu()
    xs, wsc, bsc = x.cuda(), [w.cuda() for w in ws], [b.cuda() for b in bs]
    os.environ["NLSH_MLP_IMPL"] = "simt"
    try:
        simt_logits, simt_codes = run_mlp(xs, wsc, bsc)
    finally:
        os.environ.pop("NLSH_MLP_IMPL")
    tc_logits, tc_codes = run_mlp(xs, wsc, bsc)
    scale = ref.abs().max(dim=1, keepdim=True)[0].clamp_min(1e-6)
    err_tc = ((tc_logits.cpu().double() - ref).abs() / scale).max().item()
    err_simt = ((simt_logits.cpu().double() - ref).abs() / scale).max().item()
    print(f"m={m} dims={dims}: max rel err tc={err_tc:.2e} simt={err_simt:.2e}")
    assert err_simt < 1e-5
    assert err_tc < 1e-5  # BASELINE.json: logits within 1e-5 relative tolerance
    agree = (tc_codes == simt_codes).float().mean().item()
    assert agree >= 0.999, agree
    # codes are exactly the threshold rule applied to the path's own logits
    from nlsh import _native
    assert torch.equal(_native.codes_from_logits(tc_logits, 0), tc_codes)
    # the default fuses the output layer into the epilogue of the layer before it (tc_linear.cu, TcArgs::head_w);
    # as a tcgen05 layer of its own it must meet the same bar
    os.environ["NLSH_MLP_FUSE_HEAD"] = "0"
    try:
        un_logits, un_codes = run_mlp(xs, wsc, bsc)
    finally:
        os.environ.pop("NLSH_MLP_FUSE_HEAD")
    err_un = ((un_logits.cpu().double() - ref).abs() / scale).max().item()
    assert err_un < 1e-5, err_un
    assert torch.equal(_native.codes_from_logits(un_logits, 0), un_codes)
    assert (un_codes == tc_codes).float().mean().item() >= 0.999
