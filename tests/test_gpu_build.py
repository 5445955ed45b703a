"""GPU parity of the CSR index build against the oracle / reference goldens."""
import numpy as np
import pytest
import torch

from helpers import unpad

pytestmark = pytest.mark.gpu


def check_csr(codes_np, n_buckets, offsets, ids, oracle):
    off = offsets.cpu().numpy()
    ids = ids.cpu().numpy()
    index2row = oracle.build_index([{int(c)} for c in codes_np])
    assert off[0] == 0 and off[-1] == len(codes_np)
    for b in range(n_buckets):
        got = ids[off[b]:off[b + 1]].tolist()
        assert got == index2row.get(b, np.zeros(0, np.int64)).tolist(), f"bucket {b}"


def test_reference_test_indexer_golden_through_api(golden):
    # nlsh/tests/test_indexer.py:6-26 through the drop-in build_index (multi-code rows)
    from nlsh.indexer import build_index
    index2row = build_index([set([1, 2]), set([2, 3, 4]), set([1, 5])], cuda=False)
    expected = {1: torch.LongTensor([0, 2]), 2: torch.LongTensor([0, 1]), 3: torch.LongTensor([1]),
                4: torch.LongTensor([1]), 5: torch.LongTensor([2])}
    assert index2row.keys() == expected.keys()
    for k in index2row.keys():
        assert torch.equal(expected[k], index2row[k])
    on_gpu = build_index([set([1, 2]), set([2, 3, 4]), set([1, 5])])
    assert all(v.is_cuda and v.dtype == torch.int64 for v in on_gpu.values())
    assert list(on_gpu.keys())[0] in (1, 2)  # first-seen order, as the reference's dict
    assert build_index([]) == {}
    ints = build_index([3, 1, 3, 3], cuda=False)  # eval.py passes ints
    assert ints[3].tolist() == [0, 2, 3] and ints[1].tolist() == [1]
    neg = build_index([{-1}, {5}, {-1, 5}], cuda=False)  # int16-wrapped codes are legal keys
    assert neg[-1].tolist() == [0, 2] and neg[5].tolist() == [1, 2]


@pytest.mark.parametrize("tag", ["l2", "ang"])
def test_csr_matches_reference_index(golden, oracle, tag):
    from nlsh import _native
    codes = golden[f"{tag}_db_codes"].astype(np.int32)
    n_buckets = 1 << int(golden[f"{tag}_hs"])
    X = torch.from_numpy(golden[f"{tag}_X"]).cuda()
    offsets, ids, xs = _native.build_csr(torch.from_numpy(codes).cuda(), n_buckets, X)
    check_csr(codes, n_buckets, offsets, ids, oracle)
    keys = golden[f"{tag}_index_keys"].tolist()
    off = offsets.cpu().numpy()
    for key, rows in zip(keys, unpad(golden[f"{tag}_index_rows"])):
        assert ids[off[key]:off[key + 1]].cpu().tolist() == rows
    assert torch.equal(xs[:, :X.shape[1]], X[ids.long()])


@pytest.mark.parametrize("n,n_buckets,d", [(0, 16, 8), (1, 16, 8), (31, 1, 4), (4096, 16, 128), (8192, 256, 100),
                                           (100000, 16, 128), (99999, 4096, 30), (300001, 32768, 0)])
def test_csr_shapes(oracle, n, n_buckets, d):
    from nlsh import _native
    g = torch.Generator().manual_seed(n + n_buckets)
    # skewed bucket sizes, some empty buckets
    codes = (torch.rand(n, generator=g) ** 3 * n_buckets).long().clamp_(0, n_buckets - 1).int()
    X = torch.randn(n, d, generator=g).cuda() if d else None
    offsets, ids, xs = _native.build_csr(codes.cuda(), n_buckets, X)
    if n <= 100000:
        check_csr(codes.numpy(), n_buckets, offsets, ids, oracle)
    # size-independent properties: stable sort by code == torch's stable sort
    order = torch.sort(codes.cuda().long(), stable=True)[1]
    assert torch.equal(ids.long(), order)
    assert torch.equal(offsets.long(), torch.searchsorted(
        codes.cuda().long()[order].contiguous(), torch.arange(n_buckets + 1, device="cuda")))
    if d:
        dp = (d + 3) // 4 * 4
        assert xs.shape == (n, dp)
        assert torch.equal(xs[:, :d], X[order])
        assert (xs[:, d:] == 0).all()


def test_full_size_build_properties():
    # BASELINE config sizes: 1M x 128 / 256 buckets and 10M codes / 4096 buckets (ids only)
    from nlsh import _native
    g = torch.Generator(device="cuda").manual_seed(3)
    for n, n_buckets, d in [(1_000_000, 256, 128), (10_000_000, 4096, 0)]:
        codes = torch.randint(0, n_buckets, (n,), generator=g, device="cuda", dtype=torch.int32)
        X = torch.randn(n, d, generator=g, device="cuda") if d else None
        offsets, ids, xs = _native.build_csr(codes, n_buckets, X)
        sorted_codes = codes[ids.long()]
        assert (sorted_codes[1:] >= sorted_codes[:-1]).all()                      # bucket-major
        same = sorted_codes[1:] == sorted_codes[:-1]
        assert (ids[1:][same] > ids[:-1][same]).all()                              # ascending inside a bucket
        assert torch.equal(torch.bincount(ids.long(), minlength=n), torch.ones(n, dtype=torch.long, device="cuda"))
        assert torch.equal(offsets[1:] - offsets[:-1], torch.bincount(codes.long(), minlength=n_buckets).int())
        if d:
            assert torch.equal(xs, X[ids.long()])


def test_out_of_range_codes_are_reported():
    from encoders import MultiLayerRelu
    from nlsh import _native
    codes = torch.tensor([0, 1, 7, 2], dtype=torch.int32).cuda()
    offsets, ids, _ = _native.build_csr(codes, 4, None)
    assert int(offsets[-1]) == 3  # the stray code is dropped, the count tells the caller
    with pytest.raises(ValueError):
        _native.build_csr(codes, 4, torch.zeros(3, 4).cuda())
