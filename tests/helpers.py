"""Shared helpers for the parity tests."""
import numpy as np


def golden_layers(golden, prefix, oracle):
    """The hasher of a golden case as oracle.Layer list (ReLU on all but the last)."""
    n = int(golden[f"{prefix}_n_layers"])
    return [oracle.Layer(golden[f"{prefix}_w{i}"], golden[f"{prefix}_b{i}"], relu=(i < n - 1))
            for i in range(n)]


def rows_to_sets(arr, fill=-32768):
    return [set(int(v) for v in row if v != fill) for row in arr]


def unpad(arr, fill=-1):
    return [[int(v) for v in row if v != fill] for row in arr]


def assert_topk_equal_up_to_ties(got_ids, got_d, ref_ids, ref_d, rtol=1e-5, atol=0.0):
    """Top-k parity as BASELINE.json states it: ids identical except at distance ties,
    distances within rtol (+ atol).  A position may differ only if the two distances there
    agree within tolerance (a tie swap) and both ids' distances tie with a neighbour."""
    assert len(got_ids) == len(ref_ids)
    for q, (gi, gd, ri, rd) in enumerate(zip(got_ids, got_d, ref_ids, ref_d)):
        gi, ri = list(gi), list(ri)
        gd, rd = np.asarray(gd, dtype=np.float64), np.asarray(rd, dtype=np.float64)
        assert len(gi) == len(ri), f"query {q}: {len(gi)} ids vs {len(ri)}"
        if len(gi) == 0:
            continue
        tol = atol + rtol * np.abs(rd)
        assert np.all(np.abs(gd - rd) <= tol), f"query {q}: distances differ {gd} vs {rd}"
        if gi == ri:
            continue
        # ids may differ only inside groups of (near-)tied distances, or at the k-th boundary
        for pos, (a, b) in enumerate(zip(gi, ri)):
            if a == b:
                continue
            tied_inside = b in gi and abs(gd[gi.index(b)] - rd[pos]) <= tol[pos]
            at_boundary = abs(rd[pos] - rd[-1]) <= tol[pos]
            assert tied_inside or at_boundary, \
                f"query {q} pos {pos}: id {a} vs {b} without a distance tie ({gd} vs {rd})"


def hashing_from_golden(golden, prefix, cls="bernoulli"):
    """A nlsh.hashings object carrying the golden case's weights (on the GPU)."""
    import torch
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli, Categorical
    n = int(golden[f"{prefix}_n_layers"])
    dims = [golden[f"{prefix}_w{i}"].shape for i in range(n)]
    enc = MultiLayerRelu(dims[0][1], [s[0] for s in dims[:-1]])
    if cls == "categorical":
        h = Categorical(enc, dims[-1][0], None)
    else:
        h = MultivariateBernoulli(enc, dims[-1][0], None, tanh_output=bool(int(golden.get(f"{prefix}_tanh", 0))))
    linears = [m for m in h._hasher._encoder.children() if isinstance(m, torch.nn.Linear)]
    linears.append(h._hasher.output_layer)
    with torch.no_grad():
        for i, lin in enumerate(linears):
            lin.weight.copy_(torch.from_numpy(golden[f"{prefix}_w{i}"]))
            lin.bias.copy_(torch.from_numpy(golden[f"{prefix}_b{i}"]))
    h.train_mode(False)
    return h


def oracle_layers_from_hashing(hashing, oracle):
    from nlsh.hashings import extract_layer_tensors
    return [oracle.Layer(w.cpu(), None if b is None else b.cpu(), act == 1)
            for w, b, act in extract_layer_tensors(hashing._hasher._encoder, hashing._hasher.output_layer)]


def mixture(n, d, n_centers, seed, spread=2.0, noise=1.0):
    import torch
    g = torch.Generator().manual_seed(seed)
    centers = torch.randn(n_centers, d, generator=g) * spread
    return centers[torch.randint(0, n_centers, (n,), generator=g)] + noise * torch.randn(n, d, generator=g)
