"""Throughput of the serving loop (nlsh.parallel.PipelinedSearch, queries resident in HBM) against the number of
lanes and the SMs the tensor-core scan leaves free, in one process: batches of the bench workload per second.

    python scripts/dbg_lanes.py [workload] [p] [rows]        LANE_VARIANTS="lanes:reserve:seed;..."
"""
import json
import os
import sys

import torch

sys.path.insert(0, "neural-locality-sensitive-hashing_b200")
sys.path.insert(0, ".")
import bench  # noqa: E402
import synth  # noqa: E402
from nlsh.parallel import PipelinedSearch, ShardedIndexer  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4_10m_x128_4096b"
n, d, hs, nq, k, metric, seed = bench.WORKLOADS[wl]
p = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else n
dev = torch.device("cuda")
X = synth.make_database(n, d, hs, seed, dev, sep=bench.SEP)
Q = synth.make_queries(nq, d, hs, seed, dev, sep=bench.SEP)
hashing, _ = bench.make_hashing(d, hs, metric, seed, dev, 300)
index = ShardedIndexer(hashing, X, hashing.distance, shard_lo=0, metric=metric)
ref = index.query_tensors(Q, k=k, hash_times=p)
variants = os.environ.get("LANE_VARIANTS", "1:0:-;2:0:-;2:8:-;2:16:-;2:32:-;3:16:-;3:32:-;2:16:128;2:32:128;3:32:128;1:0:128")
for v in variants.split(";"):
    lanes, reserve, seed_rows = v.split(":")
    os.environ.pop("NLSH_SCAN_SEED", None)
    if seed_rows != "-":
        os.environ["NLSH_SCAN_SEED"] = seed_rows
    index.local.scan_flags = int(reserve) << 8
    pipe = PipelinedSearch(index, nq, k=k, hash_times=p, depth=int(lanes), to_host=False)
    ids, dd, _ = pipe.result(pipe.submit(Q))
    same = bool(torch.equal(ids, ref[0]) and torch.equal(dd, ref[1]))
    nb = 40

    def run():
        for _ in range(nb):
            pipe.submit(Q)
        pipe.fence()
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nb
    print(json.dumps({"lanes": int(lanes), "sm_reserve": int(reserve), "seed_rows": seed_rows, "ms_per_batch": ms,
                      "qps": nq / ms * 1e3, "equal_to_serial": same}), flush=True)
    del pipe
