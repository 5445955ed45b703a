"""nlsh — B200-native drop-in for the index-build + query hot path of
stegben/neural-locality-sensitive-hashing (same import names as the reference's `nlsh`
package: nlsh.utils, nlsh.hashings, nlsh.indexer, nlsh.metrics).

The reference's nlsh/__init__.py:1-3 installs pyximport to compile utils.pyx at import;
here the native code is the prebuilt libnlsh_b200.so (see nlsh/_native.py), so importing
this package has no side effects and works on a CPU-only host (compute calls raise).
"""

__all__ = ["utils", "hashings", "indexer", "metrics"]
