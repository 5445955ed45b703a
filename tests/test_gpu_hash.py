"""GPU parity of the hasher forward, the bucket-code epilogue and the probe enumeration
(through the C ABI) against the oracle and the reference-generated golden vectors."""
import numpy as np
import pytest
import torch

from helpers import golden_layers, hashing_from_golden, rows_to_sets

pytestmark = pytest.mark.gpu
LOGIT_RTOL = 1e-5  # BASELINE.json: logits within 1e-5 relative tolerance


def logit_close(got, ref):
    # relative to the larger of |ref| and the row scale: a logit that cancels to ~0 cannot be
    # held to 1e-5 of itself by any summation order
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
    return np.abs(got - ref) <= LOGIT_RTOL * scale


@pytest.mark.parametrize("tag", ["l2", "ang"])
def test_logits_and_codes_match_reference(golden, oracle, tag):
    from nlsh import _native
    h = hashing_from_golden(golden, tag)
    X = torch.from_numpy(golden[f"{tag}_X"]).cuda()
    codes, _, logits = h.hash_tensors(X, 1, want_logits=True)
    ref_logits = golden[f"{tag}_logits"]
    assert logit_close(logits.cpu().numpy(), ref_logits).all()
    # bit-exact codes when computed from identical fp32 logits
    from_ref_logits = _native.codes_from_logits(torch.from_numpy(ref_logits).cuda(), h.head)
    assert from_ref_logits.cpu().tolist() == golden[f"{tag}_db_codes"].tolist()
    # and from the device's own logits, under the reference's rule
    assert np.array_equal(oracle.hard_codes(logits.cpu(), h.head), codes.cpu().numpy())
    agreement = (codes.cpu().numpy() == golden[f"{tag}_db_codes"]).mean()
    print(f"[{tag}] bucket agreement vs reference: {agreement:.6f}")
    assert agreement >= 0.999
    # set-returning API
    assert h.hash(X[:50]) == [{int(c)} for c in codes[:50].cpu().tolist()]


def test_threshold_dead_band_bit_exact(golden):
    from nlsh import _native
    l = torch.from_numpy(golden["thr_logits"]).cuda()[:, None].contiguous()
    assert np.array_equal(_native.codes_from_logits(l, _native.HEAD_SIGMOID).cpu().numpy().astype(bool),
                          golden["thr_sigmoid_bits"])
    assert np.array_equal(_native.codes_from_logits(l, _native.HEAD_TANH).cpu().numpy().astype(bool),
                          golden["thr_tanh_bits"])


def test_bit_order_msb_first():
    from nlsh import _native
    l = torch.tensor([[1., -1., 1.], [-1., 1., 1.], [1., 1., 1.], [-1., -1., -1.]]).cuda()
    assert _native.codes_from_logits(l, _native.HEAD_SIGMOID).cpu().tolist() == [5, 3, 7, 0]


def test_categorical_head(golden, oracle):
    from nlsh import _native
    h = hashing_from_golden(golden, "cat", cls="categorical")
    X = torch.from_numpy(golden["cat_X"]).cuda()
    codes, probes, logits = h.hash_tensors(X, 3, want_logits=True)
    assert logit_close(logits.cpu().numpy(), golden["cat_logits"]).all()
    assert _native.codes_from_logits(torch.from_numpy(golden["cat_logits"]).cuda(),
                                     _native.HEAD_SOFTMAX).cpu().tolist() == golden["cat_codes"].tolist()
    assert np.array_equal(probes.cpu().numpy(), oracle.topp_probes(logits.cpu().numpy(), oracle.HEAD_SOFTMAX, 3))
    assert h.hash(X[:20]) == codes[:20].cpu().tolist()


@pytest.mark.parametrize("hs,p", [(4, 1), (5, 6), (8, 16), (12, 16), (12, 40), (15, 100), (3, 16)])
def test_topp_probes_match_specification(oracle, hs, p):
    from nlsh import _native
    l = torch.randn(64, hs, generator=torch.Generator().manual_seed(hs * 100 + p))
    l[0, :] = 0.0  # all-ties row
    l[1, 0] = l[1, 1]  # equal costs
    for head in (_native.HEAD_SIGMOID, _native.HEAD_TANH):
        got = _native.topp_probes(l.cuda(), head, p).cpu().numpy()
        want = oracle.topp_probes(l.numpy(), head, p)
        assert np.array_equal(got, want)
        assert np.array_equal(got[:, 0], _native.codes_from_logits(l.cuda(), head).cpu().numpy())


@pytest.mark.parametrize("hs,p", [(4, 2), (7, 9), (12, 16), (15, 33)])
def test_sampled_probes_match_the_restated_bernoulli_draws(oracle, hs, p):
    """hashings.py:77-81: hard code + n - 1 Bernoulli(probs) draws.  The kernel's draws equal the oracle's
    (same Philox counters) except where a uniform sits within 1e-6 of the probability; the hard-code column is
    bit-exact; a different seed gives different draws; the flip frequency follows sigmoid(logit)."""
    from nlsh import _native
    l = torch.randn(300, hs, generator=torch.Generator().manual_seed(hs * 7 + p)) * 1.5
    for head in (_native.HEAD_SIGMOID, _native.HEAD_TANH):
        got = _native.sample_probes(l.cuda(), head, p, 1234).cpu().numpy()
        want, margin = oracle.sample_probes(l.numpy(), head, p, 1234)
        safe = margin > 1e-6
        assert safe.mean() > 0.99
        assert np.array_equal(got[safe], want[safe])
        assert np.array_equal(got[:, 0], _native.codes_from_logits(l.cuda(), head).cpu().numpy())
        assert np.array_equal(got, _native.sample_probes(l.cuda(), head, p, 1234).cpu().numpy())
        assert not np.array_equal(got[:, 1:], _native.sample_probes(l.cuda(), head, p, 99).cpu().numpy()[:, 1:])
    # distribution: one row of logits repeated, many draws -> per-bit frequency ~ sigmoid(logit)
    row = torch.tensor([[-2.0, -0.5, 0.0, 0.7, 3.0]])
    draws = _native.sample_probes(row.repeat(4000, 1).cuda(), _native.HEAD_SIGMOID, 9, 7).cpu().numpy()[:, 1:]
    bits = (draws[..., None] >> np.arange(4, -1, -1)) & 1
    freq = bits.reshape(-1, 5).mean(axis=0)
    assert np.abs(freq - torch.sigmoid(row[0]).numpy()).max() < 0.01, freq


def test_sampled_hash_sets_feed_the_index(oracle):
    """MultivariateBernoulli.hash(x, n, sample_seed=...) returns the reference's List[Set[int]] (every set
    holds the hard code, at most n codes) and Indexer.query accepts sampled probe rows with repeats."""
    import torch.nn.functional as F
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer
    torch.manual_seed(3)
    h = MultivariateBernoulli(MultiLayerRelu(32, [64]), 6, F.pairwise_distance)
    h.train_mode(False)
    X = torch.randn(5000, 32, generator=torch.Generator().manual_seed(1)).cuda()
    Q = torch.randn(40, 32, generator=torch.Generator().manual_seed(2)).cuda()
    sets = h.hash(Q, n=5, sample_seed=11)
    hard = h.hash(Q, n=1)
    assert all(len(s) <= 5 and next(iter(b)) in s for s, b in zip(sets, hard))
    idx = Indexer(h, X, F.pairwise_distance)
    _, probes, _ = h.hash_tensors(Q, 5, sample_seed=11)
    ids, dists, ncand = idx.query_tensors(Q, k=5, probes=probes)
    index2row = oracle.build_index([{int(c)} for c in h.hash_tensors(X, 1)[0].cpu().tolist()])
    o_ids, o_d, o_n = oracle.query(X.cpu(), index2row, Q.cpu(), sets, "l2", 5)
    assert ncand.cpu().tolist() == o_n
    for q in range(40):
        np.testing.assert_allclose(dists[q, :len(o_d[q])].cpu().numpy(), o_d[q], rtol=1e-5)


@pytest.mark.parametrize("n,d,hidden,hs", [(1, 128, [256, 256], 12), (4097, 100, [256, 256], 10),
                                           (70001, 128, [256, 256], 8), (3000, 960, [256, 256], 9),
                                           (513, 30, [17], 3), (1000, 128, [64, 64], 12)])
def test_mlp_shapes_against_oracle(oracle, n, d, hidden, hs):
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    from helpers import oracle_layers_from_hashing
    torch.manual_seed(n + d)
    h = MultivariateBernoulli(MultiLayerRelu(d, hidden), hs, None)
    h.train_mode(False)
    X = torch.randn(n, d, generator=torch.Generator().manual_seed(n))
    codes, _, logits = h.hash_tensors(X.cuda(), 1, want_logits=True)
    ref = oracle.mlp_logits(X, oracle_layers_from_hashing(h, oracle)).numpy()
    assert logit_close(logits.cpu().numpy(), ref).all()
    agree = (oracle.hard_codes(torch.from_numpy(ref), oracle.HEAD_SIGMOID) == codes.cpu().numpy()).mean()
    assert agree >= 0.995, agree


def test_empty_and_error_paths():
    from encoders import MultiLayerRelu
    from nlsh import _native
    from nlsh.hashings import MultivariateBernoulli
    h = MultivariateBernoulli(MultiLayerRelu(8, [16]), 4, None)
    assert h.hash(torch.zeros((0, 8)).cuda()) == []
    with pytest.raises(ValueError):
        h.hash(torch.zeros((2, 8)).cuda(), n=0)
    with pytest.raises(ValueError):  # hash_size beyond the int16 code range of utils.pyx
        MultivariateBernoulli(MultiLayerRelu(8, [16]), 16, None).hash(torch.zeros((2, 8)).cuda())
    with pytest.raises(ValueError):  # wrong input width
        h.hash(torch.zeros((2, 9)).cuda())
    with pytest.raises(_native.NativeLibraryError):
        h.hash(torch.zeros((2, 8)))


def _fp64_logits(x, specs, native):
    h = x.double()
    for w, b, act, scale in specs:
        h = h @ w.double().cpu().T
        if b is not None:
            h = h + b.double().cpu()
        if act == native.ACT_RELU:
            h = torch.relu(h)
        elif act == native.ACT_SIN:
            h = torch.sin(scale * h)
    return h.numpy()


def _row_scale_err(got, ref):
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
    return float((np.abs(got - ref) / scale).max())


@pytest.mark.parametrize("dims,d,hs,x_scale", [
    ([256, 256, 64], 128, 12, 0.05),   # main.py:283,388 defaults (encoder_structure 256,256 + the 64-wide output)
    ([256, 256], 128, 12, 0.05),
    ([256, 256, 64], 100, 10, 0.05),   # GloVe-100 shape (cfg3)
    ([256, 256, 64], 960, 9, 0.02),    # GIST shape (cfg5)
    ([256, 256, 64], 128, 12, 1.0),    # unit-scale inputs: sin(30 x) of arguments up to ~200
])
@pytest.mark.parametrize("impl", ["default", "simt", "tc"])
def test_siren_trunk_logits_against_oracle(oracle, dims, d, hs, x_scale, impl):
    """Hasher forward over the SIREN trunk main.py:388 builds (encoders.py:58-79).  Bars, from the B200
    measurements in profiles/r2_siren.jsonl:
    * the shipped path (sine trunks take the fp32 SIMT kernels) and NLSH_MLP_IMPL=simt hold the 1e-5
      logit bar against the fp32 oracle (measured 1.1e-6 .. 6.2e-6 of the row scale);
    * the tcgen05 3xTF32 kernels, forced with NLSH_MLP_IMPL=tc, do not: sin(30 x) amplifies the split's
      2^-21 error to 0.7e-5 .. 1.7e-5 (5e-5 allowed here) - which is why sine trunks are not routed there;
    * at unit input scale no fp32 evaluation holds 1e-5: the oracle itself is 5.2e-5 away from an fp64
      evaluation (summation order under sin(30 x)), so there the fp32 paths are held to 2x the oracle's
      own distance from fp64 and the forced tensor-core path to 10x.
    Codes must follow bit-exactly from the device's own logits in every case."""
    import os
    from encoders import Siren
    from nlsh import _native
    from nlsh.hashings import MultivariateBernoulli, extract_layer_specs
    torch.manual_seed(2)
    hashing = MultivariateBernoulli(Siren(d, dims), hs, None)
    hashing.train_mode(False)
    X = torch.randn(3000, d) * x_scale
    old = os.environ.pop("NLSH_MLP_IMPL", None)
    if impl != "default":
        os.environ["NLSH_MLP_IMPL"] = impl
    try:
        codes, _, logits = hashing.hash_tensors(X.cuda(), 1, want_logits=True)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("NLSH_MLP_IMPL", None)
        if old is not None:
            os.environ["NLSH_MLP_IMPL"] = old
    hasher = hashing._hasher
    specs = extract_layer_specs(hasher._encoder, hasher.output_layer)
    layers = [oracle.Layer(w.cpu(), None if b is None else b.cpu(), act == _native.ACT_RELU,
                           scale if act == _native.ACT_SIN else None) for w, b, act, scale in specs]
    ref = oracle.mlp_logits(X, layers).numpy()
    got = logits.cpu().numpy()
    err = _row_scale_err(got, ref)
    agree = float((oracle.hard_codes(torch.from_numpy(ref), oracle.HEAD_SIGMOID) == codes.cpu().numpy()).mean())
    print(f"siren {d}->{dims}->{hs} x{x_scale} impl={impl}: max logit err {err:.3g} of the row scale, "
          f"bucket agreement {agree:.5f}")
    if x_scale >= 1.0:
        ref64 = _fp64_logits(X, specs, _native)
        own = _row_scale_err(ref, ref64)
        assert _row_scale_err(got, ref64) <= (10.0 if impl == "tc" else 2.0) * own + 1e-5
        assert agree >= 0.999
    else:
        assert err <= (5e-5 if impl == "tc" else 1e-5), err
        assert agree >= 0.9995, agree
    assert np.array_equal(oracle.hard_codes(logits.cpu(), oracle.HEAD_SIGMOID), codes.cpu().numpy())
