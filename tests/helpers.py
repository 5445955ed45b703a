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
