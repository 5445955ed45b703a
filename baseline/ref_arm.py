"""Loader for the UNMODIFIED reference under baseline/_ref (made by baseline/Makefile) with the three
shims SURVEY section 10.1 lists, for bench.py's reference arm and the caller tests.  Nothing here is
product code and nothing under neural-locality-sensitive-hashing_b200/ imports it.

The reference's package is called `nlsh`, like this repo's drop-in: a process imports ONE of them.
`load(cpu=True)` puts baseline/_ref first on sys.path and refuses to run when this repo's `nlsh` is
already imported.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "nlsh", "indexer.py"))


def _stub(name, **attrs):
    if name not in sys.modules:
        mod = types.ModuleType(name)
        for key, val in attrs.items():
            setattr(mod, key, val)
        sys.modules[name] = mod


def load(cpu=True):
    """-> namespace with the reference's own Indexer, build_index, MultivariateBernoulli, MultiLayerRelu,
    calculate_recall (nlsh/indexer.py, nlsh/hashings.py, encoders.py, nlsh/metrics.py of the copy).
    cpu=True neutralises the hard-coded `.cuda()` calls (hashings.py:37, indexer.py:20,60,61) for THIS
    process, so that the reference's code path runs on the host cores."""
    if not available():
        raise FileNotFoundError(f"{REF_ROOT} is missing: run `make -C baseline ref` where /root/reference exists")
    loaded = sys.modules.get("nlsh")
    if loaded is not None and not os.path.abspath(getattr(loaded, "__file__", "")).startswith(REF_ROOT):
        raise RuntimeError("this process has already imported the repo's own `nlsh`; the reference arm "
                           "needs a process of its own")
    import torch
    _stub("siren", SIREN=object)   # encoders.py:5 (siren-torch is not installed; MultiLayerRelu is in-repo)
    _stub("h5py")                  # precompute.py:5 / nlsh/data.py
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import encoders as ref_encoders
    from nlsh.hashings import MultivariateBernoulli
    from nlsh.indexer import Indexer, build_index
    from nlsh.metrics import calculate_recall
    import nlsh.utils as ref_utils
    assert os.path.abspath(ref_encoders.__file__).startswith(REF_ROOT)
    return types.SimpleNamespace(
        Indexer=Indexer, build_index=build_index, MultivariateBernoulli=MultivariateBernoulli,
        MultiLayerRelu=ref_encoders.MultiLayerRelu, calculate_recall=calculate_recall,
        hash_codes=ref_utils.hash_codes, utils_file=ref_utils.__file__, root=REF_ROOT)
