"""Candidate helpers — mirror of the reference's Cython module nlsh/utils.pyx.

`hash_codes` keeps the reference signature and error behaviour (utils.pyx:18-32): it takes
a host int32 array [n_codes, n_samples, n_bits] of 0/1 bits and returns one Python set of
int16-wrapped codes per row.  The bit pack itself runs in libnlsh_b200.so
(nlsh_pack_codes_host); only the set construction is Python, as in the reference.
"""
from typing import List, Set

import numpy as np

from . import _native


def hash_codes(codes) -> List[Set[int]]:
    arr = np.asarray(codes)
    # Cython's `int[:, :, :]` buffer check (utils.pyx:19): wrong ndim / dtype -> ValueError
    if arr.ndim != 3:
        raise ValueError(f"Buffer has wrong number of dimensions (expected 3, got {arr.ndim})")
    if arr.dtype != np.intc:
        raise ValueError(f"Buffer dtype mismatch, expected 'int' but got '{arr.dtype}'")
    packed = _native.pack_codes_host(arr)
    return [set(row) for row in packed.tolist()]
