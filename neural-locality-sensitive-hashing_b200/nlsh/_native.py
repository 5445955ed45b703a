"""ctypes binding of libnlsh_b200.so (include/nlsh_b200.h).

This is the only place that touches the shared library.  There is deliberately no CPU or
PyTorch fallback: if the library is missing or CUDA is unavailable every compute wrapper
raises, so a silent slow path can never be mistaken for the product.
"""
import ctypes
import os
import threading

import numpy as np
import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# NLSH_B200_LIB=<path> loads another build of the library (A/B runs of compile-time variants)
LIB_PATH = os.environ.get("NLSH_B200_LIB") or os.path.join(_PKG_ROOT, "lib", "libnlsh_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3
ACT_IDENTITY, ACT_RELU, ACT_SIN = 0, 1, 2
HEAD_SIGMOID, HEAD_TANH, HEAD_SOFTMAX = 0, 1, 2
METRIC_L2, METRIC_ANGULAR, METRIC_L2SQ, METRIC_COSINE = 0, 1, 2, 3
MAX_K = 128
MAX_HASH_BITS = 15
FLAG_SYNC_STAGING = 1
FLAG_NO_TC_FILTER = 2
FLAG_TINY_CAND_BUFFERS = 4
FLAG_SQUARED_L2_OUT = 8

METRIC_BY_NAME = {"l2": METRIC_L2, "angular": METRIC_ANGULAR, "l2sq": METRIC_L2SQ,
                  "cosine": METRIC_COSINE}


class NativeLibraryError(RuntimeError):
    """libnlsh_b200.so is missing / unloadable, or no CUDA device is present."""


class LayerStruct(ctypes.Structure):
    _fields_ = [
        ("weight", ctypes.c_void_p),
        ("bias", ctypes.c_void_p),
        ("in_dim", ctypes.c_int32),
        ("out_dim", ctypes.c_int32),
        ("act", ctypes.c_int32),
        ("act_scale", ctypes.c_float),
    ]


_vp, _i32, _i64, _u32, _sz = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32,
                              ctypes.c_size_t)
_LP = ctypes.POINTER(LayerStruct)

# name -> (restype, argtypes): one entry per symbol declared in include/nlsh_b200.h
PROTOTYPES = {
    "nlsh_version": (ctypes.c_int, []),
    "nlsh_last_error": (ctypes.c_char_p, []),
    "nlsh_pack_codes_host": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp]),
    "nlsh_mlp_workspace_bytes": (_sz, [_i64, _LP, _i32]),
    "nlsh_mlp_hash_f32": (ctypes.c_int, [_vp, _i64, _i32, _LP, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "nlsh_codes_from_logits": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "nlsh_topp_probes": (ctypes.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "nlsh_sample_probes": (ctypes.c_int, [_vp, _i64, _i32, _i32, _i32, ctypes.c_uint64, _vp, _vp]),
    "nlsh_build_workspace_bytes": (_sz, [_i64, _i32]),
    "nlsh_build_csr": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nlsh_query_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32, _i64, _i64]),
    "nlsh_query_scan_topk": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i64,
                                            _i64, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _sz, _u32,
                                            _vp]),
    "nlsh_query_scan_topk_seeded": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i64,
                                                   _i64, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _u32,
                                                   _vp]),
    "nlsh_query_seed_workspace_bytes": (_sz, [_i64, _i32]),
    "nlsh_query_seed_tau": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _i32, _vp, _i64, _i32, _i32, _vp, _vp,
                                           _sz, _vp]),
    "nlsh_query_seed_tau_rows": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _i32, _vp, _i64, _i32, _i32, _i32,
                                                _vp, _vp, _sz, _vp]),
    "nlsh_query_scan_impl": (ctypes.c_int, [_i32, _i32, _i32, _i32, _i64, _i32, _i32]),
    "nlsh_knn_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "nlsh_knn_bruteforce": (ctypes.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i64, _i64,
                                           _vp, _vp, _vp, _sz, _vp]),
    "nlsh_merge_topk": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _vp]),
    "nlsh_recall_hits": (ctypes.c_int, [_vp, _i32, _vp, _i32, _i64, _vp, _vp]),
    "nlsh_kernel_launch_count": (ctypes.c_longlong, []),
    "nlsh_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "nlsh_profile_read": (ctypes.c_int, [_vp, ctypes.c_int]),
}

_lib = None
_lib_lock = threading.Lock()


def lib():
    """Load (once) and return the ctypes handle; raises NativeLibraryError if absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeLibraryError(
                    f"{LIB_PATH} not found: build it with `make -C {_PKG_ROOT}` "
                    "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                    "nlsh_b200 has no CPU fallback.")
            try:
                handle = ctypes.CDLL(LIB_PATH)
            except OSError as exc:  # pragma: no cover - depends on the machine
                raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
            for name, (restype, argtypes) in PROTOTYPES.items():
                fn = getattr(handle, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = handle
    return _lib


def last_error():
    return lib().nlsh_last_error().decode("utf-8", "replace")


def _check(rc, what):
    if rc == OK:
        return
    msg = f"{what}: {last_error()}"
    if rc == ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise NativeLibraryError(
            f"{name} must be a CUDA tensor: the nlsh_b200 hot path runs only on a B200 GPU "
            "(no CPU fallback).")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class Workspace:
    """Scratch memory for the library calls of ONE stream (or one captured CUDA graph).

    The C ABI's contract is caller-owned, stream-ordered scratch (the library never allocates): two
    calls may share a buffer only if they are ordered on the same stream.  A Workspace therefore binds
    to the first stream it is used on and refuses any other; it grows on demand, and the buffer it
    outgrows is handed back to the allocator only after that stream has passed it (record_stream).
    `nlsh.indexer.GraphedQuery` owns a private one: the kernels captured in its graph hold raw pointers
    into it, so nothing else may write there between replays."""

    def __init__(self, device, name="anonymous"):
        self.device = torch.device(device)
        self.name = name
        self.buf = None
        self.stream_id = None
        self.frozen = False  # set once a CUDA graph has captured pointers into buf

    def get(self, nbytes):
        stream = torch.cuda.current_stream(self.device)
        if self.stream_id is None:
            self.stream_id = stream.cuda_stream
        elif self.stream_id != stream.cuda_stream:
            raise RuntimeError(
                f"nlsh workspace {self.name!r} belongs to stream {self.stream_id:#x} but is used on "
                f"stream {stream.cuda_stream:#x}: concurrent streams need a Workspace each")
        if self.buf is None or self.buf.numel() < nbytes:
            if self.frozen:
                raise RuntimeError(
                    f"nlsh workspace {self.name!r} is captured in a CUDA graph and cannot grow "
                    f"({self.buf.numel()} -> {nbytes} bytes): capture a new graph for the new shape")
            if self.buf is not None:
                self.buf.record_stream(stream)  # kernels already enqueued may still use the old buffer
            self.buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=self.device)
        return self.buf


_workspaces = {}
_workspaces_lock = threading.Lock()


def _workspace(device, nbytes, workspace=None):
    """The scratch buffer of this call: the caller's Workspace, or the cached one of
    (device, current stream) - calls on different streams never share scratch."""
    if workspace is None:
        stream = torch.cuda.current_stream(device)
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device(),
               stream.cuda_stream)
        with _workspaces_lock:
            workspace = _workspaces.get(key)
            if workspace is None:
                workspace = _workspaces[key] = Workspace(device, name=f"stream {stream.cuda_stream:#x}")
    return workspace.get(nbytes)


def release_workspaces():
    with _workspaces_lock:
        _workspaces.clear()


def _f32c(t, name):
    require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# --------------------------------------------------------------------------------------
# host helper
# --------------------------------------------------------------------------------------
def pack_codes_host(bits):
    """int32 [n, s, hs] 0/1 bits (any strides) -> int16 [n, s]; utils.pyx:6-32 arithmetic."""
    n, s, hs = bits.shape
    out = np.empty((n, s), dtype=np.int16)
    isz = bits.itemsize
    rc = lib().nlsh_pack_codes_host(
        ctypes.c_void_p(bits.ctypes.data), n, s, hs, bits.strides[0] // isz,
        bits.strides[1] // isz, bits.strides[2] // isz, ctypes.c_void_p(out.ctypes.data))
    _check(rc, "nlsh_pack_codes_host")
    return out


# --------------------------------------------------------------------------------------
# hasher
# --------------------------------------------------------------------------------------
class LayerSpec:
    """One Linear(+activation): weight [out, in] / bias [out] CUDA fp32 tensors."""

    def __init__(self, weight, bias, act=ACT_IDENTITY, act_scale=1.0):
        self.weight = _f32c(weight, "layer weight")
        self.bias = _f32c(bias, "layer bias") if bias is not None else None
        self.act = int(act)
        self.act_scale = float(act_scale)

    @property
    def in_dim(self):
        return self.weight.shape[1]

    @property
    def out_dim(self):
        return self.weight.shape[0]


def _layer_array(layers):
    arr = (LayerStruct * len(layers))()
    for i, L in enumerate(layers):
        arr[i].weight = L.weight.data_ptr()
        arr[i].bias = L.bias.data_ptr() if L.bias is not None else None
        arr[i].in_dim = L.in_dim
        arr[i].out_dim = L.out_dim
        arr[i].act = L.act
        arr[i].act_scale = L.act_scale
    return arr


def mlp_hash(x, layers, head, want_logits=True, want_codes=True, workspace=None):
    """x [n, d] -> (logits [n, hs] fp32 | None, codes [n] int32 | None)."""
    x = _f32c(x, "x")
    n, d = x.shape
    hs = layers[-1].out_dim
    arr = _layer_array(layers)
    logits = torch.empty((n, hs), dtype=torch.float32, device=x.device) if want_logits else None
    codes = torch.empty((n,), dtype=torch.int32, device=x.device) if want_codes else None
    with torch.cuda.device(x.device):
        nbytes = lib().nlsh_mlp_workspace_bytes(n, arr, len(layers))
        ws = _workspace(x.device, nbytes, workspace)
        rc = lib().nlsh_mlp_hash_f32(_ptr(x), n, d, arr, len(layers), head, _ptr(logits),
                                     _ptr(codes), _ptr(ws), ws.numel(), _stream())
    _check(rc, "nlsh_mlp_hash_f32")
    return logits, codes


def codes_from_logits(logits, head):
    logits = _f32c(logits, "logits")
    n, hs = logits.shape
    codes = torch.empty((n,), dtype=torch.int32, device=logits.device)
    with torch.cuda.device(logits.device):
        rc = lib().nlsh_codes_from_logits(_ptr(logits), n, hs, head, _ptr(codes), _stream())
    _check(rc, "nlsh_codes_from_logits")
    return codes


def topp_probes(logits, head, p):
    logits = _f32c(logits, "logits")
    n, hs = logits.shape
    probes = torch.empty((n, p), dtype=torch.int32, device=logits.device)
    with torch.cuda.device(logits.device):
        rc = lib().nlsh_topp_probes(_ptr(logits), n, hs, head, p, _ptr(probes), _stream())
    _check(rc, "nlsh_topp_probes")
    return probes


def sample_probes(logits, head, p, seed):
    """Bernoulli-sampled probe codes int32 [n, p] (column 0 = hard code), hashings.py:77-81."""
    logits = _f32c(logits, "logits")
    n, hs = logits.shape
    probes = torch.empty((n, p), dtype=torch.int32, device=logits.device)
    with torch.cuda.device(logits.device):
        rc = lib().nlsh_sample_probes(_ptr(logits), n, hs, head, p, int(seed) & (2 ** 64 - 1), _ptr(probes),
                                      _stream())
    _check(rc, "nlsh_sample_probes")
    return probes


# --------------------------------------------------------------------------------------
# index build
# --------------------------------------------------------------------------------------
def padded_dim(d):
    return (d + 3) // 4 * 4


def build_csr(codes, n_buckets, x=None, want_sqnorm=False, workspace=None):
    """codes int32 [n] in [0, n_buckets) -> (offsets int32 [B+1], ids int32 [n], x_sorted|None)
    [, x_sqnorm fp32 [n] when want_sqnorm: |x_sorted row|^2, the tensor-core scan filter's input]."""
    require_cuda(codes, "codes")
    codes = codes.to(torch.int32).contiguous()
    n = codes.shape[0]
    dev = codes.device
    offsets = torch.empty((n_buckets + 1,), dtype=torch.int32, device=dev)
    ids = torch.empty((n,), dtype=torch.int32, device=dev)
    xs = None
    d = 0
    if x is not None:
        x = _f32c(x, "x")
        if x.shape[0] != n:
            raise ValueError(f"build_csr: {n} codes for {x.shape[0]} rows")
        d = x.shape[1]
        xs = torch.empty((n, padded_dim(d)), dtype=torch.float32, device=dev)
    xn = None
    if want_sqnorm:
        if xs is None:
            raise ValueError("build_csr: want_sqnorm needs x")
        xn = torch.empty((n,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib().nlsh_build_workspace_bytes(n, n_buckets)
        ws = _workspace(dev, nbytes, workspace)
        rc = lib().nlsh_build_csr(_ptr(codes), n, n_buckets, _ptr(x), d, _ptr(offsets), _ptr(ids),
                                  _ptr(xs), _ptr(xn), _ptr(ws), ws.numel(), _stream())
    _check(rc, "nlsh_build_csr")
    if want_sqnorm:
        return offsets, ids, xs, xn
    return offsets, ids, xs


# --------------------------------------------------------------------------------------
# query / kNN / merge
# --------------------------------------------------------------------------------------
def query_seed_tau(xq, probes, offsets, x_sorted, d, metric, k, workspace=None, sample_rows=0):
    """Distance bounds fp32 [Q] of the queries from a sample of the rows of their probed buckets (the seed of
    the tensor-core scan's filter); valid for any shard of the same database (query_scan_topk(tau_seed=...)).
    sample_rows: base sample size (0 = the library's rule)."""
    xq = _f32c(xq, "query_vectors")
    require_cuda(probes, "probes")
    probes = probes.to(torch.int32).contiguous()
    nq = xq.shape[0]
    tau = torch.empty((nq,), dtype=torch.float32, device=xq.device)
    with torch.cuda.device(xq.device):
        nbytes = lib().nlsh_query_seed_workspace_bytes(nq, d)
        ws = _workspace(xq.device, nbytes, workspace)
        rc = lib().nlsh_query_seed_tau_rows(_ptr(xq), nq, d, _ptr(probes), probes.shape[1], _ptr(offsets),
                                            offsets.shape[0] - 1, _ptr(x_sorted), x_sorted.shape[0], metric, k,
                                            int(sample_rows), _ptr(tau), _ptr(ws), ws.numel(), _stream())
    _check(rc, "nlsh_query_seed_tau_rows")
    return tau


def query_scan_topk(xq, probes, offsets, ids, x_sorted, d, max_bucket_rows, metric, k,
                    id_offset=0, flags=0, out=None, x_sqnorm=None, workspace=None, tau_seed=None):
    """-> (ids int64 [Q, k], dists fp32 [Q, k], n_cand int32 [Q]); `out` = preallocated
    contiguous (ids, dists, n_cand) tensors to write into; x_sqnorm (from build_csr) enables the
    tensor-core filtered scan; tau_seed fp32 [Q] = distance bounds from query_seed_tau (any shard)."""
    xq = _f32c(xq, "query_vectors")
    require_cuda(probes, "probes")
    probes = probes.to(torch.int32).contiguous()
    nq = xq.shape[0]
    if xq.shape[1] != d:
        raise ValueError(f"query dim {xq.shape[1]} != index dim {d}")
    p = probes.shape[1]
    n_buckets = offsets.shape[0] - 1
    n_rows = ids.shape[0]
    dev = xq.device
    if out is not None:
        out_ids, out_d, out_n = out
        assert out_ids.shape == (nq, k) and out_ids.dtype == torch.int64 and out_ids.is_contiguous()
        assert out_d.shape == (nq, k) and out_d.dtype == torch.float32 and out_d.is_contiguous()
        assert out_n.shape == (nq,) and out_n.dtype == torch.int32 and out_n.is_contiguous()
    else:
        out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_n = torch.empty((nq,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib().nlsh_query_workspace_bytes(nq, p, k, d, n_buckets, n_rows, max_bucket_rows)
        ws = _workspace(dev, nbytes, workspace)
        if tau_seed is not None:
            require_cuda(tau_seed, "tau_seed")
            assert tau_seed.dtype == torch.float32 and tau_seed.shape == (nq,) and tau_seed.is_contiguous()
        rc = lib().nlsh_query_scan_topk_seeded(_ptr(xq), nq, d, _ptr(probes), p, _ptr(offsets), n_buckets,
                                               _ptr(ids), _ptr(x_sorted), _ptr(x_sqnorm), n_rows,
                                               max_bucket_rows, metric,
                                               k, id_offset, _ptr(tau_seed), _ptr(out_ids), _ptr(out_d),
                                               _ptr(out_n), _ptr(ws), ws.numel(), flags, _stream())
    _check(rc, "nlsh_query_scan_topk")
    return out_ids, out_d, out_n


def scan_impl(d, k, metric, has_sqnorm=True, n_queries=1 << 20, p=1, n_buckets=1):
    """Which scan kernel nlsh_query_scan_topk runs for this shape: 1 = tensor-core filtered, 0 = fp32 SIMT."""
    return int(lib().nlsh_query_scan_impl(d, k, metric, 1 if has_sqnorm else 0, n_queries, p, n_buckets))


def knn_bruteforce(xq, xdb, metric, k, exclude_self=False, self_offset=0, id_offset=0, workspace=None):
    """Exact kNN of xq [Q, d] in xdb [N, d] -> (ids int64 [Q, k], dists fp32 [Q, k])."""
    xq = _f32c(xq, "xq")
    xdb = _f32c(xdb, "xdb")
    if xq.shape[1] != xdb.shape[1]:
        raise ValueError(f"knn: query dim {xq.shape[1]} != database dim {xdb.shape[1]}")
    d = xq.shape[1]
    if d % 4:
        pad = padded_dim(d) - d  # zero columns change none of the four metrics (no eps inside L2SQ;
        if metric == METRIC_L2:  # the scan metric L2 adds eps per column, so it is not paddable)
            raise ValueError("knn with the eps-L2 scan metric needs d % 4 == 0")
        xq = torch.nn.functional.pad(xq, (0, pad))
        xdb = torch.nn.functional.pad(xdb, (0, pad))
        d += pad
    nq, n = xq.shape[0], xdb.shape[0]
    dev = xq.device
    out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib().nlsh_knn_workspace_bytes(nq, n, d, k)
        ws = _workspace(dev, nbytes, workspace)
        rc = lib().nlsh_knn_bruteforce(_ptr(xq), nq, _ptr(xdb), n, d, metric, k,
                                       1 if exclude_self else 0, self_offset, id_offset,
                                       _ptr(out_ids), _ptr(out_d), _ptr(ws), ws.numel(), _stream())
    _check(rc, "nlsh_knn_bruteforce")
    return out_ids, out_d


def merge_topk(dists, ids, ncand=None):
    """dists fp32 [G, Q, k] / ids int64 [G, Q, k] per-shard lists (the G dimension may be
    strided, e.g. views into a packed all-gather buffer) -> merged ([Q, k] ids, [Q, k] dists
    [, summed ncand int32 [Q] when per-shard ncand int32 [G, Q] is given])."""
    require_cuda(dists, "dists")
    require_cuda(ids, "ids")
    g, nq, k = dists.shape
    if dists.dtype != torch.float32 or ids.dtype != torch.int64:
        dists, ids = dists.float(), ids.to(torch.int64)
    if dists.stride()[1:] != (k, 1) or ids.stride()[1:] != (k, 1):
        dists, ids = dists.contiguous(), ids.contiguous()
    dev = dists.device
    out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_n = None
    n_stride = 0
    if ncand is not None:
        require_cuda(ncand, "ncand")
        if ncand.dtype != torch.int32 or ncand.stride(1) != 1:
            ncand = ncand.to(torch.int32).contiguous()
        out_n = torch.empty((nq,), dtype=torch.int32, device=dev)
        n_stride = ncand.stride(0) if g > 1 else nq
    with torch.cuda.device(dev):
        rc = lib().nlsh_merge_topk(_ptr(dists), _ptr(ids), dists.stride(0) if g > 1 else 0,
                                   ids.stride(0) if g > 1 else 0, _ptr(ncand), n_stride, g, nq, k,
                                   _ptr(out_ids), _ptr(out_d), _ptr(out_n), _stream())
    _check(rc, "nlsh_merge_topk")
    if ncand is not None:
        return out_ids, out_d, out_n
    return out_ids, out_d


def recall_hits(gt, pred):
    """gt int64 [Q, k_gt], pred int64 [Q, k_pred] -> hits int32 [Q] (|set(gt) & set(pred)|)."""
    require_cuda(gt, "gt")
    require_cuda(pred, "pred")
    gt = gt.to(torch.int64).contiguous()
    pred = pred.to(torch.int64).contiguous()
    nq = gt.shape[0]
    hits = torch.empty((nq,), dtype=torch.int32, device=gt.device)
    with torch.cuda.device(gt.device):
        rc = lib().nlsh_recall_hits(_ptr(gt), gt.shape[1], _ptr(pred), pred.shape[1], nq, _ptr(hits),
                                    _stream())
    _check(rc, "nlsh_recall_hits")
    return hits


# --------------------------------------------------------------------------------------
# measurement hooks
# --------------------------------------------------------------------------------------
def kernel_launch_count():
    return int(lib().nlsh_kernel_launch_count())


def profile_enable(on=True):
    lib().nlsh_profile_enable(1 if on else 0)


def profile_read(capacity=256):
    """Durations (ms) of the scan kernels recorded since profile_enable / the last read."""
    buf = (ctypes.c_float * capacity)()
    n = lib().nlsh_profile_read(ctypes.cast(buf, ctypes.c_void_p), capacity)
    if n < 0:
        raise RuntimeError("nlsh_profile_read failed")
    return [float(buf[i]) for i in range(n)]
