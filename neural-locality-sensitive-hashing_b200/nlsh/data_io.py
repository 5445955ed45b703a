"""Minimal on-disk formats for real data sets without h5py: the TEXMEX
.fvecs / .ivecs / .bvecs files SIFT-1M and GIST-1M ship in, and .npy.  The reference reads the
ann-benchmarks HDF5 files (nlsh/data.py:23-49: keys train / test / neighbors / train_knn) and
writes `.processed` HDF5 (precompute.py:91-99); `load_dataset` / `save_processed` keep those key
names over a directory of .npy files."""
import os

import numpy as np


def read_vecs(path, dtype=None, max_rows=None):
    """TEXMEX vector file -> [n, d] array.  Every record is int32 d followed by d components
    (float32 for .fvecs, int32 for .ivecs, uint8 for .bvecs)."""
    ext = os.path.splitext(path)[1].lower()
    comp = {".fvecs": np.float32, ".ivecs": np.int32, ".bvecs": np.uint8}.get(ext, dtype)
    if comp is None:
        raise ValueError(f"unknown vector file extension {ext!r}: pass dtype")
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size == 0:
        return np.zeros((0, 0), dtype=comp)
    d = int(raw[:4].view(np.int32)[0])
    rec = 4 + d * np.dtype(comp).itemsize
    if d <= 0 or raw.size % rec:
        raise ValueError(f"{path}: not a {ext} file (d={d}, {raw.size} bytes)")
    n = raw.size // rec
    if max_rows is not None:
        n = min(n, max_rows)
    body = raw[: n * rec].reshape(n, rec)
    if not (body[:, :4].view(np.int32)[:, 0] == d).all():
        raise ValueError(f"{path}: records of different dimension")
    return np.ascontiguousarray(body[:, 4:]).view(comp).reshape(n, d)


def write_vecs(path, arr):
    ext = os.path.splitext(path)[1].lower()
    comp = {".fvecs": np.float32, ".ivecs": np.int32, ".bvecs": np.uint8}[ext]
    arr = np.ascontiguousarray(arr, dtype=comp)
    n, d = arr.shape
    rec = np.empty((n, 4 + d * arr.itemsize), dtype=np.uint8)
    rec[:, :4] = np.full((n, 1), d, dtype=np.int32).view(np.uint8)
    rec[:, 4:] = arr.view(np.uint8).reshape(n, -1)
    rec.tofile(path)


def load_dataset(directory):
    """{train, test, neighbors[, train_knn]} from `directory` (.npy or .fvecs/.ivecs per key)."""
    out = {}
    for key, required in (("train", True), ("test", True), ("neighbors", True), ("train_knn", False)):
        for ext in (".npy", ".fvecs", ".ivecs", ".bvecs"):
            path = os.path.join(directory, key + ext)
            if os.path.exists(path):
                out[key] = np.load(path) if ext == ".npy" else read_vecs(path)
                break
        else:
            if required:
                raise FileNotFoundError(f"{directory}: no {key}.npy / .fvecs / .ivecs")
    return out


def save_processed(directory, train_knn):
    """precompute.py:91-99 writes train_knn next to the data set; here as train_knn.npy."""
    os.makedirs(directory, exist_ok=True)
    np.save(os.path.join(directory, "train_knn.npy"), np.asarray(train_knn))
