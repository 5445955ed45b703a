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


@pytest.mark.skipif(__import__("os").environ.get("NLSH_TEST_SIREN") != "1",
                    reason="sine trunk (encoders.py:58-79): host side pinned on the CPU, the CUDA forward with "
                           "NLSH_ACT_SIN has not been run on a GPU yet (round 1's GPU budget ended first); "
                           "NLSH_TEST_SIREN=1 runs it")
def test_siren_trunk_logits_against_oracle(oracle):
    """Hasher forward over the SIREN trunk main.py:388 builds.  sin(30 x) amplifies the rounding of the
    first layer 30-fold, so the bar is 1e-4 of the row scale, not the 1e-5 of the ReLU trunks; codes
    must still follow bit-exactly from the device's own logits."""
    from encoders import Siren
    from nlsh import _native
    from nlsh.hashings import MultivariateBernoulli, extract_layer_specs
    torch.manual_seed(2)
    hashing = MultivariateBernoulli(Siren(128, [256, 256, 64]), 12, None)
    hashing.train_mode(False)
    X = torch.randn(3000, 128) * 0.05
    codes, _, logits = hashing.hash_tensors(X.cuda(), 1, want_logits=True)
    hasher = hashing._hasher
    layers = [oracle.Layer(w.cpu(), None if b is None else b.cpu(), act == _native.ACT_RELU,
                           scale if act == _native.ACT_SIN else None)
              for w, b, act, scale in extract_layer_specs(hasher._encoder, hasher.output_layer)]
    ref = oracle.mlp_logits(X, layers).numpy()
    got = logits.cpu().numpy()
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
    assert (np.abs(got - ref) <= 1e-4 * scale).all()
    assert np.array_equal(oracle.hard_codes(logits.cpu(), oracle.HEAD_SIGMOID), codes.cpu().numpy())
