"""Row-sharded serving loop (nlsh.parallel.PipelinedSearch, queries resident in HBM) against the number of lanes
and the SMs the scan leaves free, N ranks in one run:  torchrun --nproc-per-node N scripts/dbg_lanes_multi.py
LANE_VARIANTS="lanes:reserve;..."  (one JSON line per variant from rank 0)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, "neural-locality-sensitive-hashing_b200")
sys.path.insert(0, ".")
import bench  # noqa: E402
import synth  # noqa: E402
from nlsh.parallel import PipelinedSearch, ShardedIndexer, shard_range  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wl = "cfg4_10m_x128_4096b"
n, d, hs, nq, k, metric, seed = bench.WORKLOADS[wl]
p = 8
lo, hi = shard_range(n, rank, world)
X = synth.make_database(n, d, hs, seed, dev, sep=bench.SEP, row_lo=lo, row_hi=hi)
Q = synth.make_queries(nq, d, hs, seed, dev, sep=bench.SEP)
hashing, _ = bench.make_hashing(d, hs, metric, seed, dev, 300)
index = ShardedIndexer(hashing, X, hashing.distance, shard_lo=lo, metric=metric)
ref = index.query_tensors(Q, k=k, hash_times=p)
for v in os.environ.get("LANE_VARIANTS", "1:0;2:0;2:16;3:0;3:16;4:16").split(";"):
    lanes, reserve = (int(t) for t in v.split(":"))
    index.local.scan_flags = reserve << 8
    pipe = PipelinedSearch(index, nq, k=k, hash_times=p, depth=lanes, to_host=False)
    ids, dd, _ = pipe.result(pipe.submit(Q))
    same = bool(torch.equal(ids, ref[0]) and torch.equal(dd, ref[1]))
    nb = 60

    def run():
        for _ in range(nb):
            pipe.submit(Q)
        pipe.fence()
    run()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / nb], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "lanes": lanes, "sm_reserve": reserve, "ms_per_batch": float(t),
                          "qps": nq / float(t) * 1e3, "equal_to_eager": same}), flush=True)
    pipe.release()
    del pipe
    dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
