"""The reference's own callers executed on the CUDA hot path: Trainer.fit of
nlsh/trainers/base.py:36-115 through SiameseTrainer and ProposedTrainer, and eval.py's hashing helpers +
build_index, all UNMODIFIED files of the reference copy under baseline/_ref (made by `make -C baseline ref`),
with this repo's nlsh.indexer / hashings / metrics overlaid (NLSH_REFERENCE_PATH).  The logged test/recall,
test/query_size and test/n_indexes of every validation block must equal the oracle's CPU flow."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-locality-sensitive-hashing_b200")
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "nlsh", "trainers", "base.py")),
                    reason="baseline/_ref (the reference copy made by `make -C baseline ref`) is absent")
def test_reference_trainers_and_eval_run_on_the_cuda_hot_path():
    env = dict(os.environ, NLSH_REFERENCE_PATH=REF,
               PYTHONPATH=os.pathsep.join([PKG, os.environ.get("PYTHONPATH", "")]))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_callers_script.py")], env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    rec = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["base_py"].startswith(REF)
    assert len(rec["siamese_blocks"]) == 3 and rec["siamese_losses"] >= 120
    assert len(rec["proposed_blocks"]) == 1 and len(rec["proposed_losses"]) >= 4
    for blk in rec["siamese_blocks"] + rec["proposed_blocks"]:
        # a logit at a rounding tie may put a row into another bucket (tensor cores vs CPU fp32)
        assert abs(blk["test/recall"] - blk["oracle_recall"]) <= 2e-3, blk
        assert abs(blk["test/query_size"] - blk["oracle_query_size"]) <= 5e-3 * blk["oracle_query_size"] + 1, blk
        assert abs(blk["test/n_indexes"] - blk["oracle_n_indexes"]) <= 1, blk
        assert 0.0 < blk["test/recall"] <= 1.0
    assert all(q > 0 for q in rec["siamese_qps"])
    assert rec["eval_build_index_equal"] and rec["eval_n_buckets"] >= 2
    print("reference callers:", json.dumps(rec["siamese_blocks"][-1]))
