"""Synthetic workloads of the BASELINE configs and a quick hasher fit (PyTorch training
loop — the part of the system that stays PyTorch).

Data: a seeded Gaussian mixture with 2^hash_size clusters arranged on a rotated hypercube
of side 2*sep*sigma in a random hash_size-dimensional subspace of R^d, plus isotropic noise
sigma in all d dimensions.  Neighbouring clusters overlap (sep controls how much), so the
learned partition needs multi-probe to reach recall 0.9 — the regime BASELINE.json names.
Rows are generated in fixed blocks so the database is bit-identical however many ranks it
is sharded over.  There is no network and no dataset: `data: synthetic`.
"""
import torch

BLOCK_ROWS = 1_250_000  # 10 M rows = 8 blocks: shardable over 1 / 2 / 4 / 8 ranks


def _basis(d, hs, seed, device):
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(d, d, generator=g))
    return q[:, :hs].t().contiguous().to(device)  # [hs, d] orthonormal rows


def _block(n, d, hs, sep, sigma, basis, seed, device, want_bits=False, skew=0.5):
    g = torch.Generator(device=device).manual_seed(seed)
    if skew == 0.5:
        bits = torch.randint(0, 2, (n, hs), generator=g, device=device, dtype=torch.int8)
    else:  # skewed cluster populations: bit = 1 with probability `skew` (bucket sizes spread over decades)
        bits = (torch.rand((n, hs), generator=g, device=device) < skew).to(torch.int8)
    x = torch.randn(n, d, generator=g, device=device) * sigma
    x.addmm_((bits.float() * 2 - 1) * (sep * sigma), basis)
    return (x, bits) if want_bits else x


def make_database(n, d, hs, seed, device, sep=3.0, sigma=1.0, row_lo=0, row_hi=None, skew=0.5):
    """Rows [row_lo, row_hi) of the n x d database (fp32, C-contiguous) on `device`."""
    row_hi = n if row_hi is None else row_hi
    basis = _basis(d, hs, seed, device)
    out = torch.empty((row_hi - row_lo, d), dtype=torch.float32, device=device)
    b0 = row_lo // BLOCK_ROWS
    b1 = (row_hi + BLOCK_ROWS - 1) // BLOCK_ROWS
    for b in range(b0, b1):
        lo, hi = b * BLOCK_ROWS, min((b + 1) * BLOCK_ROWS, n)
        blk = _block(hi - lo, d, hs, sep, sigma, basis, seed * 1000 + b, device, skew=skew)
        s, e = max(lo, row_lo), min(hi, row_hi)
        out[s - row_lo:e - row_lo] = blk[s - lo:e - lo]
        del blk
    return out


def make_queries(nq, d, hs, seed, device, sep=3.0, sigma=1.0, skew=0.5):
    """Held-out queries from the same mixture (the survey's `seed + 7` convention)."""
    basis = _basis(d, hs, seed, device)
    return _block(nq, d, hs, sep, sigma, basis, (seed + 7) * 1000 + 999, device, skew=skew)


def fit_hasher(hashing, d, hs, seed, device, sep=3.0, sigma=1.0, steps=300, batch=4096, lr=2e-3, skew=0.5):
    """Cheap supervised surrogate for the reference's trainers (nlsh/trainers/*): Adam on the
    BCE between `hashing.predict` and the generating cluster bits of fresh mixture samples.
    Returns the final loss.  Uses only torch (autograd); the CUDA hot path is not involved."""
    basis = _basis(d, hs, seed, device)
    opt = torch.optim.Adam(list(hashing.parameters()), lr=lr, amsgrad=True)
    hashing.train_mode(True)
    loss = None
    for step in range(steps):
        x, bits = _block(batch, d, hs, sep, sigma, basis, seed * 7919 + 17 + step, device, True, skew=skew)
        probs = hashing.predict(x)
        if getattr(hashing, "_tanh_output", False):
            probs = probs / 2. + 0.5
        loss = torch.nn.functional.binary_cross_entropy(probs.clamp(1e-6, 1 - 1e-6), bits.float())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    hashing.train_mode(False)
    return float(loss.detach().item())
