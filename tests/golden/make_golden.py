"""Generates tests/golden/nlsh_golden.npz by RUNNING THE REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
It imports the reference's nlsh package (pyximport compiles its utils.pyx), encoders.py and
precompute.py with the three shims of SURVEY §8c (stub `siren`, stub `h5py`, `.cuda()` as a
no-op), runs the hot path on small seeded inputs and stores inputs + outputs.  The oracle
(oracle/nlsh_oracle.py) is pinned against this file by tests/test_oracle_golden.py and the
CUDA path by tests/test_gpu_*.py; nothing reads /root/reference at test time.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("NLSH_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nlsh_golden.npz")

sys.path.insert(0, REF)
sys.modules["siren"] = types.ModuleType("siren")
sys.modules["siren"].SIREN = object  # encoders.py:5
sys.modules["h5py"] = types.ModuleType("h5py")  # precompute.py:5
torch.Tensor.cuda = lambda self, *a, **k: self  # hashings.py:37, indexer.py:20,60,61, precompute.py:58
torch.nn.Module.cuda = lambda self, *a, **k: self

from encoders import MultiLayerRelu  # noqa: E402
from nlsh.hashings import MultivariateBernoulli, Categorical  # noqa: E402
from nlsh.indexer import Indexer, build_index  # noqa: E402
from nlsh.metrics import calculate_recall  # noqa: E402
from nlsh.utils import hash_codes  # noqa: E402
import precompute  # noqa: E402

G = {}


def sets_to_array(sets, width=None):
    width = width or max(len(s) for s in sets)
    arr = np.full((len(sets), width), -32768, dtype=np.int32)  # -32768 never a valid fill here
    for i, s in enumerate(sets):
        vals = sorted(s)
        arr[i, :len(vals)] = vals
    return arr


def ragged(lists, fill=-1):
    width = max((len(r) for r in lists), default=0)
    arr = np.full((len(lists), max(width, 1)), fill, dtype=np.int64)
    for i, r in enumerate(lists):
        arr[i, :len(r)] = r
    return arr


def layer_arrays(prefix, hasher):
    mods = [m for m in hasher._encoder.children() if isinstance(m, torch.nn.Linear)]
    mods.append(hasher.output_layer)
    G[f"{prefix}_n_layers"] = np.int64(len(mods))
    for i, m in enumerate(mods):
        G[f"{prefix}_w{i}"] = m.weight.detach().numpy().copy()
        G[f"{prefix}_b{i}"] = m.bias.detach().numpy().copy()


# ---- 1. hash_codes known answers (utils.pyx), incl. the int16 wrap ------------------------
rs = np.random.RandomState(11)
bits = rs.randint(0, 2, size=(64, 3, 12)).astype(np.intc)
G["hc_bits"] = bits
G["hc_codes"] = sets_to_array(hash_codes(bits), 3)
wide = rs.randint(0, 2, size=(32, 2, 18)).astype(np.intc)
wide[0, :, :] = 1
G["hc_wide_bits"] = wide
G["hc_wide_codes"] = sets_to_array(hash_codes(wide), 2)
strided = rs.randint(0, 2, size=(10, 4, 16)).astype(np.intc)[:, ::2, ::2]  # non-contiguous view
G["hc_strided_bits"] = np.ascontiguousarray(strided)
G["hc_strided_codes"] = sets_to_array(hash_codes(strided), 2)

# ---- 2. sigmoid / tanh threshold semantics (hashings.py:24-26, 67-72) ----------------------
tiny = np.concatenate([
    np.arange(0x33000000, 0x34800000, 0x1000, dtype=np.uint32).view(np.float32),
    np.array([0.0, 5.9604645e-08, 5.960465e-08, 8.940697e-08, 8.9406974e-08, 1e-7, 1e-6, 1.0],
             dtype=np.float32)])
tiny = np.concatenate([tiny, -tiny]).astype(np.float32)
t = torch.from_numpy(tiny)
G["thr_logits"] = tiny
G["thr_sigmoid_bits"] = (torch.sigmoid(t) > 0.5).numpy()
G["thr_tanh_bits"] = ((torch.tanh(t) / 2. + 0.5) > 0.5).numpy()

# ---- 3. hasher forward + hash + build_index + query, L2 and angular -----------------------
for tag, dim, hidden, hs, n, nq, k, tanh, metric in [
        ("l2", 16, [32, 32], 5, 1500, 40, 10, False, "l2"),
        ("ang", 20, [24], 4, 1200, 30, 5, True, "angular")]:
    torch.manual_seed(1234 if tag == "l2" else 4321)
    enc = MultiLayerRelu(dim, hidden)
    dist_fn = F.pairwise_distance if metric == "l2" else (lambda a, b: 1 - F.cosine_similarity(a, b, dim=-1))
    hashing = MultivariateBernoulli(enc, hs, None, tanh_output=tanh)
    hashing.train_mode(False)
    gen = torch.Generator().manual_seed(99)
    centers = torch.randn(12, dim, generator=gen) * 2.0
    X = centers[torch.randint(0, 12, (n,), generator=gen)] + 0.5 * torch.randn(n, dim, generator=gen)
    Qv = centers[torch.randint(0, 12, (nq,), generator=gen)] + 0.5 * torch.randn(nq, dim, generator=gen)
    with torch.no_grad():
        logits = hashing._hasher.output_layer(hashing._hasher._encoder(X))
        q_logits = hashing._hasher.output_layer(hashing._hasher._encoder(Qv))
        idx = Indexer(hashing, X, dist_fn)
        db_sets = idx.hash(X, hash_times=1)
        q_sets = idx.hash(Qv, hash_times=1)
        ids, ncand = idx.query(Qv, k=k, hash_times=1)
        # multi-probe: the reference samples at random (hashings.py:77-81); fix one draw and
        # store the probe sets so both sides can be fed the same sets
        torch.manual_seed(7)
        mp_sets = hashing.hash(Qv, 6)
        # reproduce query() with those sets through the reference's own loop
        idx.hash = lambda qv, hash_times=1, batch_size=4096: mp_sets
        mp_ids, mp_ncand = idx.query(Qv, k=k, hash_times=6)
        dists = [dist_fn(Qv[i], X[torch.tensor(r, dtype=torch.long)]).numpy() if len(r) else np.zeros(0, np.float32)
                 for i, r in enumerate(ids)]
    layer_arrays(tag, hashing._hasher)
    G[f"{tag}_X"] = X.numpy()
    G[f"{tag}_Q"] = Qv.numpy()
    G[f"{tag}_hs"] = np.int64(hs)
    G[f"{tag}_k"] = np.int64(k)
    G[f"{tag}_tanh"] = np.int64(tanh)
    G[f"{tag}_logits"] = logits.numpy()
    G[f"{tag}_q_logits"] = q_logits.numpy()
    G[f"{tag}_db_codes"] = sets_to_array(db_sets, 1)[:, 0]
    G[f"{tag}_q_codes"] = sets_to_array(q_sets, 1)[:, 0]
    keys = sorted(idx.index2row.keys())
    G[f"{tag}_index_keys"] = np.asarray(keys, dtype=np.int64)
    G[f"{tag}_index_rows"] = ragged([idx.index2row[c].tolist() for c in keys])
    G[f"{tag}_query_ids"] = ragged(ids)
    G[f"{tag}_query_ncand"] = np.asarray(ncand, dtype=np.int64)
    G[f"{tag}_query_dists"] = np.stack([np.pad(d, (0, k - len(d)), constant_values=np.inf) for d in dists]).astype(np.float32)
    G[f"{tag}_mp_sets"] = sets_to_array(mp_sets, 6)
    G[f"{tag}_mp_ids"] = ragged(mp_ids)
    G[f"{tag}_mp_ncand"] = np.asarray(mp_ncand, dtype=np.int64)

# ---- 4. build_index golden of nlsh/tests/test_indexer.py:6-26 (re-run, multi-code rows) ---
ix = build_index([set([1, 2]), set([2, 3, 4]), set([1, 5])], cuda=False)
G["bi_keys"] = np.asarray(sorted(ix.keys()), dtype=np.int64)
G["bi_rows"] = ragged([ix[c].tolist() for c in sorted(ix.keys())])

# ---- 5. Categorical head (hashings.py:95-133) ---------------------------------------------
torch.manual_seed(55)
enc = MultiLayerRelu(16, [32])
cat = Categorical(enc, 24, None)
cat.train_mode(False)
Xc = torch.randn(300, 16, generator=torch.Generator().manual_seed(5))
with torch.no_grad():
    G["cat_codes"] = np.asarray(cat.hash(Xc), dtype=np.int64)
    G["cat_logits"] = cat._hasher.output_layer(cat._hasher._encoder(Xc)).numpy()
layer_arrays("cat", cat._hasher)
G["cat_X"] = Xc.numpy()

# ---- 6. brute-force kNN (precompute.py) ---------------------------------------------------
kat = np.array([[1.2, 2, 3], [3, 2, 1], [1, 2, 4], [6, 4, 2.5], [2, 4, 6]], dtype=np.float32)
G["knn_kat_vectors"] = kat  # tests/test_precompute.py:6-21
G["knn_kat_cosine_k2"] = precompute.self_get_knn_pt(kat, precompute._cosine_distance, k=2, batch_size=2)
V = torch.randn(700, 24, generator=torch.Generator().manual_seed(77)).numpy()
G["knn_vectors"] = V
G["knn_l2_k10"] = precompute.self_get_knn_pt(V, precompute._l2, k=10, batch_size=256)
G["knn_cos_k10"] = precompute.self_get_knn_pt(V, precompute._cosine_distance, k=10, batch_size=256)
Vt = torch.from_numpy(V)
G["knn_l2_matrix"] = precompute._l2(Vt[:64], Vt).numpy()
G["knn_cos_matrix"] = precompute._cosine_distance(Vt[:64], Vt).numpy()

# ---- 7. recall (metrics.py) ---------------------------------------------------------------
yt = [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]
yp = [[4, 3, 99, 98], [5, 6, 7, 8], [13, 14, 15, 16]]
G["recall_true"] = np.asarray(yt)
G["recall_pred"] = np.asarray(yp)
G["recall_values"] = np.asarray(calculate_recall(yt, yp), dtype=np.float64)
G["recall_mean"] = np.float64(calculate_recall(yt, yp, np.mean))

G["meta_torch_version"] = np.array(torch.__version__)
np.savez_compressed(OUT, **G)
print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(G), "arrays")
