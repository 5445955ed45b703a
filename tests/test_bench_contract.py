"""bench.py --impl reference (the reference's own CPU code from baseline/_ref when that copy exists, the
oracle port otherwise) on a tiny case: runs without a GPU, prints ONE JSON line with the contract's keys,
uses all host threads even under torchrun's OMP_NUM_THREADS=1, and ranks other than 0 exit 0 without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMD = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
       "cfg1_100k_x128_16b", "--rows", "5000", "--queries", "64", "--cpu-sample", "64", "--steps", "1",
       "--warmup", "1", "--fit-steps", "5"]


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")  # what torchrun exports
    out = subprocess.run(CMD, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["gpu_launches"] == 0
    assert rec["metric"] == "QPS at recall@10>=0.9" and rec["unit"] == "queries/s" and rec["higher_is_better"]
    assert rec["value"] > 0 and rec["steps"] == 1 and rec["dtype"] == "f32" and rec["data"] == "synthetic"
    assert rec["config"]["workload"] == "cfg1_100k_x128_16b" and rec["config"]["rows"] == 5000
    cpu = rec["cpu_baseline"]
    have_ref = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "nlsh", "indexer.py"))
    assert cpu["kind"] == ("reference" if have_ref else "port")
    assert cpu["cores"] == os.cpu_count() and cpu["value"] == rec["value"] and cpu["sample"]
    assert rec["e2e"] == {"value": rec["value"], "unit": "queries/s", "h2d_bytes_per_step": 0,
                          "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run(CMD, env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
