"""nlsh — B200-native drop-in for the index-build + query hot path of
stegben/neural-locality-sensitive-hashing (same import names as the reference's `nlsh`
package: nlsh.utils, nlsh.hashings, nlsh.indexer, nlsh.metrics).

The reference's nlsh/__init__.py:1-3 installs pyximport to compile utils.pyx at import;
here the native code is the prebuilt libnlsh_b200.so (see nlsh/_native.py), so importing
this package has no side effects and works on a CPU-only host (compute calls raise).
"""

import os as _os

__all__ = ["utils", "hashings", "indexer", "metrics"]

# Overlay on a checkout of the reference: with NLSH_REFERENCE_PATH=<reference root> the modules this
# package does not replace (nlsh.data, nlsh.loggers, nlsh.trainers, nlsh.learning - training glue,
# out of scope) resolve to the reference's own files, while nlsh.utils / hashings / indexer /
# metrics stay the ones here (this directory comes first on the package path).  That is what lets
# the reference's main.py / eval.py / Trainer.fit run unchanged on top of the CUDA hot path:
# nlsh/trainers/base.py:7-8 imports `nlsh.metrics` and `nlsh.indexer` and gets this package's.
_ref = _os.environ.get("NLSH_REFERENCE_PATH")
if _ref:
    _ref_pkg = _os.path.join(_ref, "nlsh")
    if _os.path.isdir(_ref_pkg) and _ref_pkg not in __path__:
        __path__.append(_ref_pkg)
del _ref
