"""Runs the reference's OWN callers of the hot path on the CUDA drop-in (launched by
tests/test_gpu_reference_callers.py in a process of its own; prints one JSON line).

With NLSH_REFERENCE_PATH = a verbatim copy of the reference (baseline/_ref) the package overlay of
neural-locality-sensitive-hashing_b200/nlsh/__init__.py resolves nlsh.trainers / nlsh.learning / nlsh.loggers
to the reference's files while nlsh.indexer / hashings / metrics / utils stay this repo's, so
* nlsh/trainers/base.py:36-115 (Trainer.fit: training loop + the validation block) and siamese.py run unmodified,
  building `Indexer(hashing, X, data.distance)` and calling `indexer.query(...)` on the CUDA path;
* nlsh/trainers/proposed.py:84-115 calls `hashing.hash(batch, n=1)` inside its loss;
* eval.py's helpers (hash_all / hash_by_batch, eval.py:56-100) feed int codes to `build_index` (eval.py:118-120).
Every validation block's logged test/recall and test/query_size are re-computed on the CPU by the oracle
from the hasher's weights at that moment.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stub(name, **attrs):
    mod = types.ModuleType(name)
    for key, val in attrs.items():
        setattr(mod, key, val)
    sys.modules[name] = mod


# third-party modules the reference's glue imports and this image does not have (none is on the hot path)
stub("hnswlib")
stub("tensorboardX", SummaryWriter=object)
stub("comet_ml", Experiment=object)
stub("wandb")
stub("h5py")
stub("ipdb")
stub("dotenv", load_dotenv=lambda *a, **k: None)
os.environ.setdefault("NLSH_MODEL_SAVE_DIR", "/tmp")

import nlsh  # noqa: E402  (this repo's package, overlaid on the reference copy)
import nlsh.indexer  # noqa: E402
import nlsh.trainers.base as ref_base  # noqa: E402  the reference's file
from nlsh.trainers.siamese import SiameseTrainer  # noqa: E402
from nlsh.trainers.proposed import ProposedTrainer  # noqa: E402
from nlsh.learning.distances import MVBernoulliL2  # noqa: E402
from nlsh.hashings import MultivariateBernoulli, extract_layer_tensors  # noqa: E402
from nlsh.metrics import calculate_recall  # noqa: E402
from encoders import MultiLayerRelu  # noqa: E402
import precompute  # noqa: E402
from oracle import nlsh_oracle as oracle  # noqa: E402  (test infrastructure: the checker)

REF = os.environ["NLSH_REFERENCE_PATH"]
assert ref_base.__file__.startswith(REF), ref_base.__file__
assert ref_base.Indexer is nlsh.indexer.Indexer
assert not nlsh.indexer.__file__.startswith(REF)


class StopFit(Exception):
    pass


class SyntheticData:
    """What Trainer.fit reads of a nlsh.data object (base.py:37-44,84): arrays + the distance callable."""

    def __init__(self, n, nq, d, k_train, K, seed):
        g = torch.Generator().manual_seed(seed)
        centers = torch.randn(48, d, generator=g) * 2.0
        self.training = (centers[torch.randint(0, 48, (n,), generator=g)] + torch.randn(n, d, generator=g)).numpy()
        self.testing = (centers[torch.randint(0, 48, (nq,), generator=g)] + torch.randn(nq, d, generator=g)).numpy()
        self.prepared = True
        self.dim = d
        # the label producers are this repo's brute-force kNN kernels (precompute.py:57-67 / ground truth)
        self.training_self_knn = precompute.self_get_knn_pt(self.training, precompute._l2, k=k_train)
        gt, _ = precompute.knn_tensors(torch.from_numpy(self.testing).cuda(), torch.from_numpy(self.training).cuda(),
                                       "l2", K)
        self.ground_truth = gt.cpu().numpy()

    @staticmethod
    def distance(v1, v2):  # nlsh/data.py:192-201 (SIFT.distance)
        return F.pairwise_distance(v1, v2)

    def load(self):
        pass


class CheckingLogger:
    """The logger protocol of nlsh/loggers.py:6-24.  At the end of each validation block it repeats the block
    on the CPU through the oracle, from the hasher's current weights, and records both sides."""
    run_name = "reference_callers"

    def __init__(self, hashing, data, K, stop_after):
        self.hashing, self.data, self.K, self.stop_after = hashing, data, K, stop_after
        self.logged = {}
        self.blocks = []

    def log(self, name, value, step):
        self.logged.setdefault(name, []).append((step, float(value)))
        if name != "training/query_size":  # the last line of a validation block (base.py:115)
            return
        hasher = self.hashing._hasher
        layers = [oracle.Layer(w.cpu(), None if b is None else b.cpu(), act == 1)
                  for w, b, act in extract_layer_tensors(hasher._encoder, hasher.output_layer)]
        X = torch.from_numpy(self.data.training)
        Qv = torch.from_numpy(self.data.testing)
        codes = oracle.hard_codes(oracle.mlp_logits(X, layers), oracle.HEAD_SIGMOID)
        index2row = oracle.build_index([{int(c)} for c in codes.tolist()])
        probes = oracle.topp_probes(oracle.mlp_logits(Qv, layers).numpy(), oracle.HEAD_SIGMOID, 10)  # query(): hash_times=10
        sets = [set(int(c) for c in row if c >= 0) for row in probes]
        o_ids, _, o_n = oracle.query(X, index2row, Qv, sets, "l2", self.K)
        self.blocks.append({
            "step": step,
            "test/recall": self.logged["test/recall"][-1][1],
            "test/query_size": self.logged["test/query_size"][-1][1],
            "test/n_indexes": self.logged["test/n_indexes"][-1][1],
            "oracle_recall": float(calculate_recall(list(self.data.ground_truth[:, :self.K]), o_ids, np.mean)),
            "oracle_query_size": float(np.mean(o_n)),
            "oracle_n_indexes": len(index2row),
        })
        if len(self.blocks) >= self.stop_after:
            raise StopFit()


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    K, d, hs = 10, 32, 6
    data = SyntheticData(n=12000, nq=600, d=d, k_train=100, K=K, seed=5)
    out = {"base_py": ref_base.__file__}

    # ---- SiameseTrainer.fit (reference's base.py + siamese.py) on the CUDA Indexer ---------------
    hashing = MultivariateBernoulli(MultiLayerRelu(d, [64, 64]).cuda(), hs, MVBernoulliL2())
    logger = CheckingLogger(hashing, data, K, stop_after=3)
    trainer = SiameseTrainer(hashing, data, "/tmp", logger)
    try:
        trainer.fit(K=K, batch_size=512, learning_rate=1e-3, test_every_updates=40)
    except StopFit:
        pass
    out["siamese_blocks"] = logger.blocks
    out["siamese_losses"] = len(logger.logged.get("training/loss", []))
    out["siamese_qps"] = [v for _, v in logger.logged.get("test/qps", [])]

    # ---- ProposedTrainer (proposed.py): hashing.hash(batch, n=1) inside the loss ---------------------
    hashing2 = MultivariateBernoulli(MultiLayerRelu(d, [64, 64]).cuda(), hs, MVBernoulliL2())
    logger2 = CheckingLogger(hashing2, data, K, stop_after=1)
    trainer2 = ProposedTrainer(hashing2, data, "/tmp", logger2, train_k=5)
    try:
        trainer2.fit(K=K, batch_size=256, learning_rate=1e-3, test_every_updates=4)
    except StopFit:
        pass
    out["proposed_blocks"] = logger2.blocks
    out["proposed_losses"] = [v for _, v in logger2.logged.get("training/loss", [])]

    # ---- eval.py: its own hashing helpers feed int codes to build_index (eval.py:118-120) ----------------
    sys.path.insert(0, REF)
    stub("nlsh.data", Glove=object)  # eval.py:11 (HDF5 loader, out of scope)
    import eval as ref_eval
    assert ref_eval.__file__.startswith(REF)
    hashing.train_mode(False)
    hasher_cpu = torch.jit.script(hashing._hasher.cpu())  # what eval.py:113 loads: the TorchScript hasher
    cand = torch.from_numpy(data.training)
    with torch.no_grad():
        indexes = ref_eval.hash_by_batch(hasher_cpu, cand, 4096)  # List[int]
    index2row = nlsh.indexer.build_index(indexes)  # int codes, grouped on the GPU
    want = oracle.build_index([{int(c)} for c in indexes])
    out["eval_build_index_equal"] = bool(
        list(index2row.keys()) == list(want.keys()) and
        all(index2row[c].cpu().tolist() == rows.tolist() for c, rows in want.items()))
    out["eval_n_buckets"] = len(index2row)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
