"""Per-role wait cycles of scan_tc_kernel on a bench workload (the -DNLSH_TC_ROLE_STATS library).

    make -C neural-locality-sensitive-hashing_b200 stats
    python scripts/dbg_tc_roles.py [workload] [p] [rows]

Prints, per role, the share of its life spent in each wait (summed over all CTAs; the filter over its 8 warps)
and the scan-kernel time of the instrumented and of the shipped library is NOT comparable (clock reads)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NLSH_B200_LIB"] = os.path.join(ROOT, "neural-locality-sensitive-hashing_b200", "lib", "libnlsh_b200_stats.so")
os.environ["NLSH_TC_STATS"] = "1"
sys.path.insert(0, os.path.join(ROOT, "neural-locality-sensitive-hashing_b200")); sys.path.insert(0, ROOT)
import torch
import bench, synth
from nlsh import _native
from nlsh.indexer import Indexer

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4_10m_x128_4096b"
n, d, hs, nq, k, metric, seed = bench.WORKLOADS[wl]
p = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else n
dev = torch.device("cuda")
X = synth.make_database(n, d, hs, seed, dev, sep=bench.SEP)
Q = synth.make_queries(nq, d, hs, seed, dev, sep=bench.SEP)
hashing, _ = bench.make_hashing(d, hs, metric, seed, dev, 300)
idx = Indexer(hashing, X, hashing.distance, metric=metric)
probes = idx.hash_tensors(Q, p)
for _ in range(3):
    idx.query_tensors(Q, k=k, probes=probes)
torch.cuda.synchronize()
_native.profile_enable(True); _native.profile_read()
_, _, nc = idx.query_tensors(Q, k=k, probes=probes)
torch.cuda.synchronize()
ms = _native.profile_read(); _native.profile_enable(False)
ws = list(_native._workspaces.values())[0].buf
B = 1 << hs
st = ws.view(torch.uint8)[(2 * B + 2) * 4:(2 * B + 2) * 4 + 26 * 8].view(torch.int64).cpu().tolist()
roles = {
    "planner": (6, ["wait q_empty + qb_empty (threshold refresh rounds)", "-", "-", "-"]),
    "streamer": (21, ["wait itm_full", "wait meta_empty", "wait slot empty", "-"]),
    "mma": (11, ["wait itm_full + qb_full", "wait acc_empty", "wait slot full", "-"]),
    "filter(8 warps)": (16, ["wait itm_full", "wait meta_full", "wait acc_full", "scoring full batches"]),
}
out = {"workload": wl, "rows": n, "p": p, "scan_ms_instrumented": ms, "pairs": int(nc.long().sum()),
       "survivors": st[0], "full_batches": st[1], "end_batches": st[2], "candidates": st[3]}
for name, (b, labels) in roles.items():
    tot = max(st[b + 4], 1)
    out[name] = {"total_cycles": tot, **{l: round(st[b + i] / tot, 4) for i, l in enumerate(labels) if l != "-"}}
print(json.dumps(out), flush=True)
