"""Pins the oracle (oracle/nlsh_oracle.py) against the reference: the reference's own two
golden tests for this path and outputs of the reference itself (tests/golden/make_golden.py).
CPU only."""
import numpy as np
import pytest
import torch

from helpers import golden_layers, rows_to_sets, unpad, assert_topk_equal_up_to_ties


# ---- the reference's own golden tests -------------------------------------------------------
def test_reference_test_indexer_golden(oracle):
    # nlsh/tests/test_indexer.py:6-26, literal fixture
    index2row = oracle.build_index([set([1, 2]), set([2, 3, 4]), set([1, 5])])
    expected = {1: [0, 2], 2: [0, 1], 3: [1], 4: [1], 5: [2]}
    assert index2row.keys() == expected.keys()
    for k in expected:
        assert index2row[k].tolist() == expected[k]


def test_reference_test_precompute_golden(oracle):
    # tests/test_precompute.py:6-21, literal fixture (function since renamed self_get_knn_pt)
    vectors = np.array([[1.2, 2, 3], [3, 2, 1], [1, 2, 4], [6, 4, 2.5], [2, 4, 6]], dtype=np.float32)
    result = oracle.self_knn(vectors, "cosine", k=2, batch_size=2)
    assert [set(r) for r in list(result)] == [{4, 2}, {3, 0}, {0, 4}, {1, 0}, {0, 2}]


# ---- against outputs of the reference itself --------------------------------------------------
@pytest.mark.parametrize("name", ["hc", "hc_wide", "hc_strided"])
def test_hash_codes_matches_reference_cython(oracle, golden, name):
    bits = golden[f"{name}_bits"]
    want = rows_to_sets(golden[f"{name}_codes"])
    assert oracle.hash_codes(bits) == want
    assert oracle.hash_codes_c(bits) == want


def test_hash_codes_known_answers(oracle):
    # int16 wrap of the Cython return type
    assert oracle.hash_codes(np.array([[[1, 0, 1], [0, 1, 1]]], dtype=np.intc)) == [{5, 3}]
    assert oracle.hash_codes(np.ones((1, 1, 15), dtype=np.intc)) == [{32767}]
    assert oracle.hash_codes(np.ones((1, 1, 16), dtype=np.intc)) == [{-1}]
    b17 = np.zeros((1, 1, 17), dtype=np.intc)
    b17[0, 0, 0] = 1
    assert oracle.hash_codes(b17) == [{0}]
    with pytest.raises(ValueError):
        oracle.hash_codes(np.ones((1, 1, 4), dtype=np.int64))
    with pytest.raises(ValueError):
        oracle.hash_codes(np.ones((1, 4), dtype=np.intc))


def test_hash_codes_matches_compiled_reference_helper(oracle):
    # oracle/_ref holds the reference's own utils.pyx compiled here (oracle/Makefile `ref`)
    import glob
    import importlib.util
    import os
    so = glob.glob(os.path.join(os.path.dirname(oracle.__file__), "_ref", "utils*.so"))
    if not so:
        pytest.skip("oracle/_ref not built (make -C oracle ref needs /root/reference)")
    spec = importlib.util.spec_from_file_location("utils", so[0])
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    bits = np.random.RandomState(3).randint(0, 2, size=(200, 4, 17)).astype(np.intc)
    assert ref.hash_codes(bits) == oracle.hash_codes(bits) == oracle.hash_codes_c(bits)


def test_threshold_semantics(oracle, golden):
    l = torch.from_numpy(golden["thr_logits"])[:, None]
    assert np.array_equal(oracle.hard_bits(l, oracle.HEAD_SIGMOID).numpy()[:, 0], golden["thr_sigmoid_bits"])
    assert np.array_equal(oracle.hard_bits(l, oracle.HEAD_TANH).numpy()[:, 0], golden["thr_tanh_bits"])
    # closed form used by the CUDA epilogue (include/nlsh_b200.h)
    x = golden["thr_logits"]
    assert np.array_equal(x > np.float32(1.5 * 2.0 ** -24), golden["thr_sigmoid_bits"])
    assert np.array_equal(x > np.float32(2.0 ** -24), golden["thr_tanh_bits"])


@pytest.mark.parametrize("tag,metric", [("l2", "l2"), ("ang", "angular")])
def test_hash_build_query_match_reference(oracle, golden, tag, metric):
    layers = golden_layers(golden, tag, oracle)
    head = oracle.HEAD_TANH if int(golden[f"{tag}_tanh"]) else oracle.HEAD_SIGMOID
    X, Q, k = golden[f"{tag}_X"], golden[f"{tag}_Q"], int(golden[f"{tag}_k"])
    logits = oracle.mlp_logits(X, layers)
    np.testing.assert_allclose(logits.numpy(), golden[f"{tag}_logits"], rtol=1e-6, atol=1e-6)
    db_sets = oracle.hash_sets(X, layers, head)
    assert [next(iter(s)) for s in db_sets] == golden[f"{tag}_db_codes"].tolist()
    assert oracle.hard_codes(logits, head).tolist() == golden[f"{tag}_db_codes"].tolist()
    index2row = oracle.build_index(db_sets)
    keys = golden[f"{tag}_index_keys"].tolist()
    assert sorted(index2row.keys()) == keys
    for key, rows in zip(keys, unpad(golden[f"{tag}_index_rows"])):
        assert index2row[key].tolist() == rows
    q_sets = oracle.hash_sets(Q, layers, head)
    assert [next(iter(s)) for s in q_sets] == golden[f"{tag}_q_codes"].tolist()
    ids, dists, ncand = oracle.query(X, index2row, Q, q_sets, metric, k, fallback="reference")
    assert ncand == golden[f"{tag}_query_ncand"].tolist()
    assert ids == unpad(golden[f"{tag}_query_ids"])
    # the reference's sampled multi-probe sets, fed identically to both sides
    mp_sets = rows_to_sets(golden[f"{tag}_mp_sets"])
    ids, dists, ncand = oracle.query(X, index2row, Q, mp_sets, metric, k, fallback="reference")
    assert ncand == golden[f"{tag}_mp_ncand"].tolist()
    want = unpad(golden[f"{tag}_mp_ids"])
    # candidate order follows set iteration order, only visible through ties
    assert_topk_equal_up_to_ties(ids, dists, want, dists, rtol=1e-6)


def test_categorical_matches_reference(oracle, golden):
    layers = golden_layers(golden, "cat", oracle)
    logits = oracle.mlp_logits(golden["cat_X"], layers)
    np.testing.assert_allclose(logits.numpy(), golden["cat_logits"], rtol=1e-6, atol=1e-6)
    assert oracle.hard_codes(logits, oracle.HEAD_SOFTMAX).tolist() == golden["cat_codes"].tolist()


def test_knn_matches_reference(oracle, golden):
    assert np.array_equal(oracle.self_knn(golden["knn_kat_vectors"], "cosine", 2, batch_size=2),
                          golden["knn_kat_cosine_k2"])
    V = golden["knn_vectors"]
    assert np.array_equal(oracle.self_knn(V, "l2sq", 10, batch_size=256), golden["knn_l2_k10"])
    assert np.array_equal(oracle.self_knn(V, "cosine", 10, batch_size=256), golden["knn_cos_k10"])
    Vt = torch.from_numpy(V)
    np.testing.assert_array_equal(oracle.knn_distance_matrix(Vt[:64], Vt, "l2sq").numpy(), golden["knn_l2_matrix"])
    np.testing.assert_array_equal(oracle.knn_distance_matrix(Vt[:64], Vt, "cosine").numpy(), golden["knn_cos_matrix"])


def test_recall_matches_reference(oracle, golden):
    yt, yp = golden["recall_true"].tolist(), golden["recall_pred"].tolist()
    assert oracle.recall(yt, yp) == pytest.approx(float(golden["recall_mean"]))
    assert golden["recall_values"].tolist() == [0.5, 1.0, 0.0]


def test_topp_probes_spec(oracle):
    # probe 0 is the hard code; costs ascend; codes distinct; p > 2^hs pads with -1
    l = torch.randn(50, 6, generator=torch.Generator().manual_seed(2)).numpy()
    pr = oracle.topp_probes(l, oracle.HEAD_SIGMOID, 8)
    base = oracle.hard_codes(torch.from_numpy(l), oracle.HEAD_SIGMOID)
    assert np.array_equal(pr[:, 0], base)
    for i in range(50):
        assert len(set(pr[i].tolist())) == 8
        cost = [np.abs(l[i])[[j for j in range(6) if ((int(c) ^ int(base[i])) >> (5 - j)) & 1]].sum() for c in pr[i]]
        assert all(cost[j] <= cost[j + 1] + 1e-6 for j in range(7))
    assert np.array_equal(oracle.topp_probes(l[:3, :2], oracle.HEAD_SIGMOID, 6)[:, 4:], -np.ones((3, 2)))
    top = oracle.topp_probes(l, oracle.HEAD_SOFTMAX, 3)
    assert np.array_equal(top[:, 0], l.argmax(1))


def test_cpu_indexer_flow_matches_reference(oracle, golden):
    # the end-to-end CPU flow used as the timed baseline gives the reference's results
    layers = golden_layers(golden, "l2", oracle)
    idx = oracle.CpuIndexer(layers, oracle.HEAD_SIGMOID, golden["l2_X"], "l2")
    ids, ncand = idx.query(golden["l2_Q"], k=int(golden["l2_k"]))
    assert ncand == golden["l2_query_ncand"].tolist()
    assert ids == unpad(golden["l2_query_ids"])


def test_oracle_nearest_exclude_positive_small_case(oracle):
    # triplet.py:44-74 on a case small enough to check by hand: points on a line
    v = torch.tensor([[0.0], [1.0], [2.1], [3.3], [10.0]])
    pos = torch.tensor([[1], [0], [1], [2], [3]])
    got = oracle.nearest_exclude_positive(v, "l2sq", pos).tolist()
    # row 0: not itself, not 1 -> 2;  row 1: not 0 -> 2;  row 2: not 1 -> 3;  row 3: not 2 -> 1;  row 4: not 3 -> 2
    assert got == [2, 2, 3, 1, 2]


def test_philox_known_answers():
    """The oracle's Philox-4x32-10 (sample_probes) against the published Random123 known-answer vectors."""
    from oracle import nlsh_oracle as oracle
    z = [np.zeros(1)] * 4
    assert [int(w[0]) for w in oracle._philox4x32_10(z, (0, 0))] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = [np.full(1, 0xffffffff)] * 4
    assert [int(w[0]) for w in oracle._philox4x32_10(f, (0xffffffff, 0xffffffff))] == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    probes, margin = oracle.sample_probes(np.zeros((4, 5), dtype=np.float32), oracle.HEAD_SIGMOID, 6, 1)
    assert probes.shape == (4, 6) and (probes[:, 0] == 0).all() and (margin[:, 0] == 1).all()
