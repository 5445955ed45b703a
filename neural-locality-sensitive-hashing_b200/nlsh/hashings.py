"""Learnable hash functions — drop-in mirror of the reference's nlsh/hashings.py.

Same classes, constructor arguments, attributes (`_hasher`, `_encoder`, `_hash_size`,
`_distance_func`) and methods (`predict`, `hash`, `parameters`, `save`, `train_mode`,
`distance`, `output_dim`) as hashings.py:11-139.  `predict` stays a PyTorch module call
(it is the training path and needs autograd); `hash` — the inference path that
Indexer._build_index / Indexer.query sit on (indexer.py:36-57) — runs the hand-written
CUDA forward + bucket-code epilogue of libnlsh_b200.so and never touches torch ops.

Differences from the reference, all deliberate:
  * multi-probe (`hash(x, n>1)`) is deterministic: the n most probable codes under
    independent bits instead of n-1 Bernoulli samples (hashings.py:77-81, Q5);
  * `hash_tensors` is a tensor-returning fast path (no Python sets, no host sync);
  * encoders that are not Linear/ReLU(/eval BatchNorm1d) stacks are refused — there is no
    PyTorch fallback for the hot path.
"""
from typing import List, Set

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native


def _maybe_cuda(module):
    # The reference calls .cuda() unconditionally (hashings.py:37); on a CPU-only host we
    # keep the module on the CPU so training-side code can still be imported and tested.
    return module.cuda() if torch.cuda.is_available() else module


def _fold_batchnorm(weight, bias, bn):
    if bn.training:
        raise NotImplementedError(
            "BatchNorm1d in training mode cannot be folded; call hashing.train_mode(False) first")
    scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps) if bn.affine else \
        1.0 / torch.sqrt(bn.running_var.detach() + bn.eps)
    shift = (bn.bias.detach() if bn.affine else 0.0) - bn.running_mean.detach() * scale
    w = weight * scale[:, None]
    b = (bias if bias is not None else torch.zeros_like(scale)) * scale + shift
    return w, b


def _leaf_modules(module):
    """Leaf modules of `module` in definition order (containers such as nn.Sequential or the SIREN
    wrapper of encoders.py:72-76 are walked, not returned)."""
    for child in module.children():
        if next(child.children(), None) is None:
            yield child
        else:
            yield from _leaf_modules(child)


def extract_layer_specs(encoder, output_layer):
    """Read the live weights of `encoder` (+ `output_layer`) as [(weight, bias, act, act_scale), ...].

    Supported trunks: any nn.Sequential-like module whose leaves are nn.Linear, nn.ReLU, a `Sine`
    (sin(w0 x), the SIREN trunk of encoders.py:58-79), nn.Identity, nn.Dropout (eval) or eval-mode
    nn.BatchNorm1d (encoders.py:24-55), and modules exposing fc1/fc2 Linear attributes applied
    with ReLU (encoders.py:8-21).  Weights are re-read at every call because they change between
    index builds while training (base.py:80-86).
    """
    pending = []  # [weight, bias, act, act_scale]

    def add_linear(lin):
        pending.append([lin.weight.detach(), lin.bias.detach() if lin.bias is not None else None,
                        _native.ACT_IDENTITY, 1.0])

    if hasattr(encoder, "fc1") and hasattr(encoder, "fc2") and isinstance(encoder.fc1, nn.Linear):
        for lin in (encoder.fc1, encoder.fc2):
            add_linear(lin)
            pending[-1][2] = _native.ACT_RELU
    else:
        leaves = list(_leaf_modules(encoder)) if isinstance(encoder, nn.Module) else None
        if not leaves:
            raise NotImplementedError(
                f"encoder {type(encoder).__name__} is not a Linear/ReLU stack; the B200 hasher "
                "kernel has no PyTorch fallback")
        for child in leaves:
            if isinstance(child, nn.Linear):
                add_linear(child)
            elif isinstance(child, nn.ReLU):
                if not pending or pending[-1][2] != _native.ACT_IDENTITY:
                    raise NotImplementedError("ReLU without a preceding Linear")
                pending[-1][2] = _native.ACT_RELU
            elif type(child).__name__ == "Sine" and hasattr(child, "w0"):
                if not pending or pending[-1][2] != _native.ACT_IDENTITY:
                    raise NotImplementedError("Sine without a preceding Linear")
                pending[-1][2] = _native.ACT_SIN
                pending[-1][3] = float(child.w0)
            elif isinstance(child, nn.BatchNorm1d):
                if not pending or pending[-1][2] != _native.ACT_IDENTITY:
                    raise NotImplementedError("BatchNorm1d must directly follow a Linear")
                pending[-1][0], pending[-1][1] = _fold_batchnorm(pending[-1][0], pending[-1][1], child)
            elif isinstance(child, nn.Identity) or (isinstance(child, nn.Dropout) and not child.training):
                continue
            else:
                raise NotImplementedError(
                    f"encoder child {type(child).__name__} is not supported by the B200 hasher "
                    "kernel (Linear / ReLU / Sine / eval BatchNorm1d only; no PyTorch fallback)")
    add_linear(output_layer)
    return [tuple(entry) for entry in pending]


def extract_layer_tensors(encoder, output_layer):
    """extract_layer_specs without the activation scale: [(weight, bias, act), ...] for the ReLU trunks
    (the callers that unpack three fields); a Sine layer needs its w0, so it is refused here."""
    specs = extract_layer_specs(encoder, output_layer)
    if any(act == _native.ACT_SIN for _, _, act, _ in specs):
        raise NotImplementedError("sine layers carry a scale: use extract_layer_specs")
    return [(w, b, act) for w, b, act, _ in specs]


def extract_layers(encoder, output_layer):
    """extract_layer_specs as CUDA LayerSpecs for the C ABI."""
    return [_native.LayerSpec(w, b, act, scale) for w, b, act, scale in extract_layer_specs(encoder, output_layer)]


def codes_to_sets(probes) -> List[Set[int]]:
    """int32 [n, p] probe matrix (-1 = unused) -> list of sets of Python ints."""
    rows = probes.cpu().tolist()
    return [set(c for c in row if c >= 0) for row in rows]


def _trunk_from_state_dict(sd):
    """Rebuild the encoder module of a saved hasher from its state dict: the `{i}_linear` /
    `{i}_batch_norm` naming of encoders.py:41-48 or the fc1 / fc2 pair of encoders.py:8-21."""
    import re

    from encoders import MultiLayerRelu, TwoLayer256Relu  # top-level module of this package's root
    if "_encoder.fc1.weight" in sd:
        return TwoLayer256Relu(sd["_encoder.fc1.weight"].shape[1], with_bias="_encoder.fc1.bias" in sd)
    idx = sorted(int(m.group(1)) for m in (re.fullmatch(r"_encoder\.(\d+)_linear\.weight", k) for k in sd) if m)
    if not idx or idx != list(range(len(idx))):
        raise NotImplementedError(
            "saved hasher is not a MultiLayerRelu / TwoLayer256Relu trunk (encoders.py:8-55); "
            "construct the hashing yourself and load_state_dict into hashing._hasher")
    weights = [sd[f"_encoder.{i}_linear.weight"] for i in idx]
    return MultiLayerRelu(weights[0].shape[1], [w.shape[0] for w in weights],
                          with_batchnorm=any("_batch_norm." in k for k in sd),
                          with_bias="_encoder.0_linear.bias" in sd)


def _load_scripted(cls, path, distance_func, kwargs):
    scripted = torch.jit.load(path, map_location="cpu")
    sd = {k: v.detach().clone() for k, v in scripted.state_dict().items()}
    hashing = cls(_trunk_from_state_dict(sd), sd["output_layer.weight"].shape[0], distance_func, **kwargs(scripted))
    hashing._hasher.load_state_dict(sd)
    hashing.train_mode(False)
    return hashing


class MultivariateBernoulli:

    class _Hasher(nn.Module):
        # hashings.py:13-27
        def __init__(self, encoder, hash_size, tanh_output=False):
            super().__init__()
            self._encoder = encoder
            self._tanh_output = tanh_output
            self.output_layer = nn.Linear(encoder.output_dim, hash_size)

        def forward(self, x):
            x = self._encoder(x)
            if self._tanh_output:
                x = torch.tanh(self.output_layer(x))
            else:
                x = torch.sigmoid(self.output_layer(x))
            return x

    def __init__(self, encoder, hash_size, distance_func, tanh_output=False):
        self._encoder = encoder
        self._hash_size = hash_size
        self._distance_func = distance_func
        self._tanh_output = tanh_output
        self._hasher = _maybe_cuda(self._Hasher(self._encoder, self._hash_size, tanh_output))

    # ---- training-side API (PyTorch, unchanged semantics) --------------------------------
    def predict(self, x):
        return self._hasher(x)

    @property
    def distance(self):
        return self._distance_func

    @property
    def output_dim(self):
        return self._hash_size

    def parameters(self):
        return self._hasher.parameters()

    def save(self, base_name):
        # hashings.py:53-57
        scripted_model_cpu = torch.jit.script(self._hasher.cpu())
        torch.jit.save(scripted_model_cpu, base_name + "_cpu.pt")
        scripted_model_gpu = torch.jit.script(self._hasher.cuda())
        torch.jit.save(scripted_model_gpu, base_name + "_gpu.pt")

    @classmethod
    def load(cls, path, distance_func):
        """The `load` classmethod hashings.py:58 leaves as a TODO: rebuild the hashing (in eval mode)
        from a TorchScript file written by `save` (base_name + "_cpu.pt" / "_gpu.pt", hashings.py:53-57;
        eval.py:113 loads the same file), so that an index can be built from a stored model."""
        return _load_scripted(cls, path, distance_func,
                              kwargs=lambda m: {"tanh_output": bool(getattr(m, "_tanh_output", False))})

    def train_mode(self, on):
        if on:
            self._hasher.train()
        else:
            self._hasher.eval()

    # ---- inference hot path (CUDA) -------------------------------------------------------
    @property
    def head(self):
        return _native.HEAD_TANH if self._tanh_output else _native.HEAD_SIGMOID

    @property
    def n_buckets(self):
        return 1 << self._hash_size

    def layer_specs(self):
        return extract_layers(self._hasher._encoder, self._hasher.output_layer)

    def hash_tensors(self, query_vectors, n=1, want_logits=False, workspace=None, sample_seed=None):
        """-> (codes int32 [B], probes int32 [B, n] | None, logits fp32 [B, hs] | None).

        codes[i] is the hard code (`probs > 0.5`, hashings.py:72); probes[i, 0] == codes[i]
        and probes[i, 1:] are the next most probable codes (-1 padded) - or, with `sample_seed`
        (an int), n - 1 draws of Bernoulli(probs) as in hashings.py:77-81 (rows may then repeat a
        code; the reference collects them into a set), reproducible for a given seed."""
        if n < 1:
            raise ValueError(f"`n` should be positive integer, but got {n}")
        need_logits = want_logits or n > 1
        logits, codes = _native.mlp_hash(query_vectors, self.layer_specs(), self.head,
                                         want_logits=need_logits, want_codes=True, workspace=workspace)
        if n == 1:
            probes = None
        elif sample_seed is None:
            probes = _native.topp_probes(logits, self.head, n)
        else:
            probes = _native.sample_probes(logits, self.head, n, sample_seed)
        return codes, probes, (logits if want_logits else None)

    def hash(self, query_vectors, n=1, sample_seed=None) -> List[Set[int]]:
        # hashings.py:66-92
        if n < 1:
            raise ValueError(f"`n` should be positive integer, but got {n}")
        if query_vectors.shape[0] == 0:
            return []
        codes, probes, _ = self.hash_tensors(query_vectors, n, sample_seed=sample_seed)
        if probes is None:
            return [{c} for c in codes.cpu().tolist()]
        return codes_to_sets(probes)


class Categorical:

    class _Hasher(nn.Module):
        # hashings.py:97-107
        def __init__(self, encoder, hash_size):
            super().__init__()
            self._encoder = encoder
            self.output_layer = nn.Linear(encoder.output_dim, hash_size)

        def forward(self, x):
            prob = self._encoder(x)
            prob = F.softmax(self.output_layer(prob), dim=1)
            return prob

    def __init__(self, encoder, hash_size, distance_func):
        self._encoder = encoder
        self._hash_size = hash_size
        self._distance_func = distance_func
        self._hasher = _maybe_cuda(self._Hasher(self._encoder, self._hash_size))

    def predict(self, x):
        return self._hasher(x)

    def distance(self, y1, y2):
        return self._distance_func(y1, y2)

    @property
    def output_dim(self):
        return self._hash_size

    def parameters(self):
        return self._hasher.parameters()

    def save(self, base_name):
        scripted_model_cpu = torch.jit.script(self._hasher.cpu())
        torch.jit.save(scripted_model_cpu, base_name + "_cpu.pt")
        scripted_model_gpu = torch.jit.script(self._hasher.cuda())
        torch.jit.save(scripted_model_gpu, base_name + "_gpu.pt")

    @classmethod
    def load(cls, path, distance_func):
        """As MultivariateBernoulli.load, for the softmax hasher (hashings.py:124-128 writes the files)."""
        return _load_scripted(cls, path, distance_func, kwargs=lambda m: {})

    def train_mode(self, on):
        if on:
            self._hasher.train()
        else:
            self._hasher.eval()

    head = _native.HEAD_SOFTMAX

    @property
    def n_buckets(self):
        return self._hash_size

    def layer_specs(self):
        return extract_layers(self._hasher._encoder, self._hasher.output_layer)

    def hash_tensors(self, query_vectors, n=1, want_logits=False, workspace=None):
        if n < 1:
            raise ValueError(f"`n` should be positive integer, but got {n}")
        need_logits = want_logits or n > 1
        logits, codes = _native.mlp_hash(query_vectors, self.layer_specs(), self.head,
                                         want_logits=need_logits, want_codes=True, workspace=workspace)
        probes = _native.topp_probes(logits, self.head, n) if n > 1 else None
        return codes, probes, (logits if want_logits else None)

    def hash(self, query_vectors, n=1):
        # hashings.py:131-133 returns List[int]; n > 1 (an extension) returns sets of the
        # top-n classes so the result can feed Indexer.query.
        if query_vectors.shape[0] == 0:
            return []
        codes, probes, _ = self.hash_tensors(query_vectors, n)
        if probes is None:
            return codes.cpu().tolist()
        return codes_to_sets(probes)


class ProductQuantization:

    def __init__(self, bits_of_each_band: List[int]):
        pass
