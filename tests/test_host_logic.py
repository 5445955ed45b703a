"""Host-side logic of the drop-in package that needs no GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neural-locality-sensitive-hashing_b200")


def test_resolve_metric():
    from nlsh import _native
    from nlsh.indexer import resolve_metric

    class SIFT:
        @staticmethod
        def distance(v1, v2):
            return F.pairwise_distance(v1, v2)

    class Glove:
        @staticmethod
        def distance(v1, v2):
            return 1 - F.cosine_similarity(v1, v2, dim=-1)

    assert resolve_metric(F.pairwise_distance) == _native.METRIC_L2
    assert resolve_metric(SIFT.distance) == _native.METRIC_L2
    assert resolve_metric(Glove.distance) == _native.METRIC_ANGULAR
    assert resolve_metric(lambda a, b: 1 - F.cosine_similarity(a, b, dim=-1)) == _native.METRIC_ANGULAR
    assert resolve_metric(lambda a, b: F.pairwise_distance(a, b)) == _native.METRIC_L2
    assert resolve_metric(None, "angular") == _native.METRIC_ANGULAR
    assert resolve_metric("l2") == _native.METRIC_L2
    with pytest.raises(ValueError):  # squared L2 is not the scan metric: refused, no fallback
        resolve_metric(lambda a, b: ((a - b) ** 2).sum(-1))
    with pytest.raises(ValueError):
        resolve_metric(None, "manhattan")


def test_extract_layer_tensors_matches_module_forward(oracle):
    from encoders import MultiLayerRelu, TwoLayer256Relu
    from nlsh import _native
    from nlsh.hashings import extract_layer_tensors
    torch.manual_seed(0)
    x = torch.randn(17, 12)
    for enc in (MultiLayerRelu(12, [16, 8]), MultiLayerRelu(12, [16, 8], with_batchnorm=True),
                MultiLayerRelu(12, [16], with_bias=False), TwoLayer256Relu(12)):
        out = nn.Linear(enc.output_dim, 5)
        for m in enc.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.running_mean.normal_()
                m.running_var.uniform_(0.5, 2.0)
                m.weight.data.normal_()
                m.bias.data.normal_()
        enc.eval()
        layers = [oracle.Layer(w, b, act == _native.ACT_RELU) for w, b, act in extract_layer_tensors(enc, out)]
        with torch.no_grad():
            want = out(enc(x))
        torch.testing.assert_close(oracle.mlp_logits(x, layers), want, rtol=1e-5, atol=1e-5)


def test_siren_trunk_extraction_matches_module_forward(oracle):
    """encoders.py:58-79 (the trunk main.py:388 builds): Linear -> sin(30 x), Linear -> sin(x), Linear,
    then the hasher's output Linear; the extracted (weight, bias, act, w0) stack must reproduce the
    module's forward.  The sine trunk comes from an unpinned third-party package: parity unpinned."""
    from encoders import MultiLayerRelu, Siren  # noqa: F401 - main.py:9 imports exactly these two names
    from nlsh import _native
    from nlsh.hashings import MultivariateBernoulli, extract_layer_specs, extract_layer_tensors
    torch.manual_seed(1)
    enc = Siren(12, [16, 16, 8])
    assert enc.output_dim == 8
    hashing = MultivariateBernoulli(enc, 5, None)
    hasher = hashing._hasher.cpu()
    specs = extract_layer_specs(hasher._encoder, hasher.output_layer)
    assert [(tuple(w.shape), act, scale) for w, _, act, scale in specs] == [
        ((16, 12), _native.ACT_SIN, 30.0), ((16, 16), _native.ACT_SIN, 1.0),
        ((8, 16), _native.ACT_IDENTITY, 1.0), ((5, 8), _native.ACT_IDENTITY, 1.0)]
    layers = [oracle.Layer(w, b, act == _native.ACT_RELU, scale if act == _native.ACT_SIN else None)
              for w, b, act, scale in specs]
    x = torch.randn(17, 12) * 0.1
    with torch.no_grad():
        want = hasher.output_layer(hasher._encoder(x))
    torch.testing.assert_close(oracle.mlp_logits(x, layers), want, rtol=1e-5, atol=1e-6)
    with pytest.raises(NotImplementedError):
        extract_layer_tensors(hasher._encoder, hasher.output_layer)  # three-field form has no room for w0


def test_extract_layers_refuses_unknown_modules():
    from nlsh.hashings import extract_layer_tensors
    enc = nn.Sequential(nn.Linear(4, 4), nn.GELU())
    with pytest.raises(NotImplementedError):
        extract_layer_tensors(enc, nn.Linear(4, 2))
    bn = nn.Sequential(nn.Linear(4, 4), nn.BatchNorm1d(4))
    bn.train()
    with pytest.raises(NotImplementedError):
        extract_layer_tensors(bn, nn.Linear(4, 2))


def test_hashing_api_surface_and_loud_failure_without_cuda():
    from encoders import MultiLayerRelu
    from nlsh import _native
    from nlsh.hashings import MultivariateBernoulli, Categorical
    h = MultivariateBernoulli(MultiLayerRelu(8, [16]), 6, F.pairwise_distance)
    assert h.output_dim == 6 and h.n_buckets == 64 and h.distance is F.pairwise_distance
    assert len(list(h.parameters())) == 4
    h.train_mode(False)
    assert not h._hasher.training
    if not torch.cuda.is_available():
        assert h.predict(torch.randn(3, 8)).shape == (3, 6)
        with pytest.raises(_native.NativeLibraryError):
            h.hash(torch.randn(3, 8))  # CPU tensors: no fallback
    with pytest.raises(ValueError):
        h.hash(torch.randn(3, 8), n=0)
    c = Categorical(MultiLayerRelu(8, [16]), 10, None)
    assert c.n_buckets == 10 and c.head == _native.HEAD_SOFTMAX


def test_bucket_view():
    from nlsh.indexer import BucketView
    off = np.array([0, 2, 2, 5, 6], dtype=np.int64)
    ids = torch.tensor([3, 9, 0, 4, 7, 1], dtype=torch.int32)
    view = BucketView(off, ids)
    assert len(view) == 3 and list(view.keys()) == [0, 2, 3]
    assert [len(v) for v in view.values()] == [2, 3, 1]
    assert view[2].tolist() == [0, 4, 7] and view[2].dtype == torch.int64
    assert view.get(1, "empty") == "empty" and 1 not in view and 2 in view
    assert dict(view.items())[3].tolist() == [1]
    assert view.sizes.tolist() == [2, 3, 1]
    assert BucketView(off, ids, id_offset=100)[0].tolist() == [103, 109]


def test_probes_from_sets_and_shard_range():
    from nlsh.indexer import Indexer
    from nlsh.parallel import shard_range
    pr = Indexer.probes_from_sets([{1, 2}, {7}, set()], "cpu", width=3)
    assert pr.shape == (3, 3) and sorted(pr[0].tolist()) == [-1, 1, 2] and pr[2].tolist() == [-1, -1, -1]
    for n, w in [(10, 3), (7, 8), (1000003, 8), (16, 4)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_recall_functions_match_reference_values(golden):
    from nlsh.metrics import calculate_recall
    yt, yp = golden["recall_true"].tolist(), golden["recall_pred"].tolist()
    assert calculate_recall(yt, yp) == golden["recall_values"].tolist()
    assert calculate_recall(yt, yp, np.mean) == pytest.approx(float(golden["recall_mean"]))


def test_precompute_metric_resolution():
    import precompute
    from nlsh import _native
    assert precompute._resolve_knn_metric(precompute._l2) == _native.METRIC_L2SQ
    assert precompute._resolve_knn_metric(precompute._cosine_distance) == _native.METRIC_COSINE
    assert precompute._resolve_knn_metric("angular") == _native.METRIC_ANGULAR
    with pytest.raises(ValueError):
        precompute._resolve_knn_metric(lambda a, b: a @ b.T)
    a, b = torch.randn(5, 7), torch.randn(6, 7)
    torch.testing.assert_close(precompute._l2(a, b), torch.cdist(a, b) ** 2, rtol=1e-4, atol=1e-4)


def test_vecs_round_trip(tmp_path):
    from nlsh.data_io import load_dataset, read_vecs, save_processed, write_vecs
    rng = np.random.RandomState(0)
    f = rng.randn(37, 16).astype(np.float32)
    i = rng.randint(0, 1000, size=(37, 5)).astype(np.int32)
    b = rng.randint(0, 256, size=(9, 128)).astype(np.uint8)
    for name, arr in (("train.fvecs", f), ("neighbors.ivecs", i), ("x.bvecs", b)):
        write_vecs(str(tmp_path / name), arr)
        assert np.array_equal(read_vecs(str(tmp_path / name)), arr)
    assert read_vecs(str(tmp_path / "train.fvecs"), max_rows=10).shape == (10, 16)
    np.save(tmp_path / "test.npy", f[:5])
    ds = load_dataset(str(tmp_path))
    assert set(ds) == {"train", "test", "neighbors"} and np.array_equal(ds["test"], f[:5])
    save_processed(str(tmp_path), i)
    assert np.array_equal(load_dataset(str(tmp_path))["train_knn"], i)
    (tmp_path / "bad.fvecs").write_bytes(b"\x03\x00\x00\x00abc")
    with pytest.raises(ValueError):
        read_vecs(str(tmp_path / "bad.fvecs"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/nlsh"), reason="needs a checkout of the reference")
def test_overlay_on_the_reference_checkout():
    """NLSH_REFERENCE_PATH: the reference's training glue (nlsh.trainers, nlsh.learning) imports on top
    of this package and its Trainer picks up THIS package's Indexer / calculate_recall
    (nlsh/trainers/base.py:7-8), which is the drop-in claim of INTEGRATION.md section 1."""
    code = r"""
import os, sys, types
sys.modules["hnswlib"] = types.ModuleType("hnswlib")          # nlsh/trainers/hnsw.py:7, not installed here
import nlsh, nlsh.indexer, nlsh.metrics, nlsh.hashings, nlsh.utils
pkg = os.path.dirname(nlsh.__file__)
import nlsh.learning.distances as dist                          # the reference's own file
import nlsh.trainers.base as base                               # the reference's own file ...
assert dist.__file__.startswith(os.environ["NLSH_REFERENCE_PATH"])
assert base.__file__.startswith(os.environ["NLSH_REFERENCE_PATH"])
assert base.Indexer is nlsh.indexer.Indexer                     # ... bound to this package's hot path
assert base.calculate_recall is nlsh.metrics.calculate_recall
for m in (nlsh.indexer, nlsh.metrics, nlsh.hashings, nlsh.utils):
    assert os.path.dirname(m.__file__) == pkg, m.__file__
from encoders import MultiLayerRelu, Siren                      # main.py:9
print("overlay ok")
"""
    env = dict(os.environ, NLSH_REFERENCE_PATH="/root/reference",
               PYTHONPATH=os.pathsep.join([PKG, os.environ.get("PYTHONPATH", "")]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "overlay ok" in out.stdout, out.stderr[-2000:]


def test_load_rebuilds_a_saved_hasher(tmp_path):
    """MultivariateBernoulli.load / Categorical.load (the reference's TODO, hashings.py:58) read the
    TorchScript files save() writes (hashings.py:53-57; eval.py:113 loads the same `_gpu.pt`)."""
    from encoders import MultiLayerRelu, TwoLayer256Relu
    from nlsh.hashings import Categorical, MultivariateBernoulli
    torch.manual_seed(4)
    x = torch.randn(9, 12)
    cases = [(MultivariateBernoulli, MultiLayerRelu(12, [16, 8], with_batchnorm=True), {"tanh_output": True}),
             (MultivariateBernoulli, MultiLayerRelu(12, [16], with_bias=False), {}),
             (MultivariateBernoulli, TwoLayer256Relu(12), {}),
             (Categorical, MultiLayerRelu(12, [16, 8]), {})]
    for cls, enc, kw in cases:
        h = cls(enc, 6, F.pairwise_distance, **kw)
        h._hasher.cpu()
        for m in h._hasher.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.running_mean.normal_()
                m.running_var.uniform_(0.5, 2.0)
        h.train_mode(False)
        path = str(tmp_path / f"{cls.__name__}_{len(list(enc.parameters()))}_cpu.pt")
        torch.jit.save(torch.jit.script(h._hasher), path)  # what save() does for the "_cpu.pt" file
        g = cls.load(path, F.pairwise_distance)
        g._hasher.cpu()
        assert type(g._hasher._encoder) is type(enc) and g._hash_size == 6
        assert getattr(g, "_tanh_output", False) == kw.get("tanh_output", False)
        assert not g._hasher.training
        with torch.no_grad():
            torch.testing.assert_close(g.predict(x), h.predict(x), rtol=0, atol=0)
    other = nn.Sequential(nn.Linear(12, 8), nn.ReLU())  # not one of the reference's trunks
    other.output_dim = 8
    odd = MultivariateBernoulli(other, 4, None)
    path = str(tmp_path / "odd_cpu.pt")
    torch.jit.save(torch.jit.script(odd._hasher.cpu()), path)
    with pytest.raises(NotImplementedError):
        MultivariateBernoulli.load(path, None)
