"""The C-ABI library: loads, exports every symbol include/nlsh_b200.h declares, and its
host-only entry points work without a GPU.  No device compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nlsh_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nlsh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nlsh import _native
    assert os.path.exists(_native.LIB_PATH), "build the library: make -C neural-locality-sensitive-hashing_b200"
    handle = ctypes.CDLL(_native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(handle, name), f"{name} declared in nlsh_b200.h but not exported"
    # and the Python binding covers exactly the header
    assert sorted(_native.PROTOTYPES) == names


def test_version_and_error_string():
    from nlsh import _native
    assert _native.lib().nlsh_version() == 310
    assert isinstance(_native.last_error(), str)


def test_argument_validation_without_gpu():
    from nlsh import _native
    L = _native.lib()
    # invalid arguments are rejected before any CUDA call is made
    assert L.nlsh_topp_probes(None, 4, 40, _native.HEAD_SIGMOID, 2, None, None) == _native.ERR_INVALID
    assert "hash_size" in _native.last_error()
    assert L.nlsh_query_scan_topk(None, 1, 8, None, 1, None, 4, None, None, None, 0, 0, 7, 10, 0, None,
                                  None, None, None, 0, 0, None) == _native.ERR_INVALID
    assert "metric" in _native.last_error()
    assert L.nlsh_build_csr(None, -1, 4, None, 0, None, None, None, None, None, 0, None) == _native.ERR_INVALID
    # sampled multi-probe: only the bit hashers sample (hashings.py:66-92), p and hash_size bounded
    assert L.nlsh_sample_probes(None, 4, 12, _native.HEAD_SOFTMAX, 4, 1, None, None) == _native.ERR_INVALID
    assert "Bernoulli" in _native.last_error()
    assert L.nlsh_sample_probes(None, 4, 12, _native.HEAD_SIGMOID, 0, 1, None, None) == _native.ERR_INVALID
    assert L.nlsh_sample_probes(None, 4, 12, _native.HEAD_SIGMOID, 4, 1, None, None) == _native.ERR_INVALID  # NULLs
    assert L.nlsh_sample_probes(None, 0, 12, _native.HEAD_SIGMOID, 4, 1, None, None) == _native.OK  # empty batch
    # seed with a caller-chosen sample: the sample size is bounded
    assert L.nlsh_query_seed_tau_rows(None, 1, 8, None, 1, None, 4, None, 0, 0, 10, -1, None, None, 0,
                                      None) == _native.ERR_INVALID
    assert L.nlsh_build_workspace_bytes(1000, 16) > 0
    assert L.nlsh_query_workspace_bytes(100, 4, 10, 128, 256, 10000, 100) > 0
    assert L.nlsh_knn_workspace_bytes(100, 1000, 128, 10) > 0


def test_pack_codes_host_matches_reference(golden, oracle):
    from nlsh.utils import hash_codes
    for name in ("hc", "hc_wide", "hc_strided"):
        bits = golden[f"{name}_bits"]
        want = [set(int(v) for v in row if v != -32768) for row in golden[f"{name}_codes"]]
        assert hash_codes(bits) == want
    view = np.random.RandomState(0).randint(0, 2, size=(9, 6, 14)).astype(np.intc)[::2, 1::2, ::3]
    assert hash_codes(view) == oracle.hash_codes(np.ascontiguousarray(view))
    assert hash_codes(np.ones((1, 1, 16), dtype=np.intc)) == [{-1}]
    assert hash_codes(np.zeros((0, 1, 4), dtype=np.intc)) == []


def test_hash_codes_error_behaviour():
    # Cython buffer-protocol errors of utils.pyx:19 are ValueErrors
    from nlsh.utils import hash_codes
    with pytest.raises(ValueError):
        hash_codes(np.ones((2, 2, 4), dtype=np.int64))
    with pytest.raises(ValueError):
        hash_codes(np.ones((2, 4), dtype=np.intc))


def test_scan_kernel_choice_is_a_pure_host_decision(monkeypatch):
    """nlsh_query_scan_impl: tensor-core filter for L2 / angular with k <= 128, d <= 4096, row norms present and
    bucket tiles shared by >= 4 (query, probe) pairs on average; fp32 SIMT scan otherwise."""
    from nlsh import _native
    monkeypatch.delenv("NLSH_SCAN_IMPL", raising=False)
    L2, ANG = _native.METRIC_L2, _native.METRIC_ANGULAR
    assert _native.scan_impl(128, 10, L2, True, 10_000, 8, 4096) == 1      # BASELINE config 4
    assert _native.scan_impl(100, 10, ANG, True, 10_000, 2, 1024) == 1     # config 3
    assert _native.scan_impl(128, 10, L2, True, 256, 8, 4096) == 0         # about one query per bucket
    assert _native.scan_impl(128, 10, L2, False, 10_000, 8, 4096) == 0     # index without row norms
    assert _native.scan_impl(960, 100, L2, True, 1000, 128, 512) == 1      # config 5: wide rows, k = 100
    assert _native.scan_impl(128, 33, L2, True, 10_000, 8, 4096) == 1
    assert _native.scan_impl(132, 10, L2, True, 10_000, 8, 4096) == 1
    assert _native.scan_impl(4100, 10, L2, True, 10_000, 8, 4096) == 0     # wider than 128 K blocks
    monkeypatch.setenv("NLSH_SCAN_IMPL", "simt")
    assert _native.scan_impl(128, 10, L2, True, 10_000, 8, 4096) == 0
    monkeypatch.setenv("NLSH_SCAN_IMPL", "tc")
    assert _native.scan_impl(128, 10, L2, True, 256, 8, 4096) == 1
