"""A/B of run-time switches of the tensor-core scan on a bench workload, all in one process (the
library reads its NLSH_* switches at every launch): scan-kernel ms from the library's event ring,
whole-call ms, survivor / candidate counters (NLSH_TC_STATS), and torch.equal against the default.

    python scripts/dbg_tc_variants.py [workload] [p] [rows]

Switches of the shipped library: NLSH_SCAN_SEED (base seed sample rows, 0 = no seed), NLSH_SCAN_SEED_DIV (the
sample is at least 1/DIV of the query's first bucket), NLSH_TC_LADDER (0 = no threshold ladder), NLSH_TC_SLOTS
(row-tile slot ring depth), NLSH_TC_CAND_CAP (candidate buffer entries per query), NLSH_SCAN_IMPL=simt (the fp32
SIMT kernel).  TC_VARIANTS="A=1,B=2;C=3" sets the list.
"""
import os, sys, json, torch
sys.path.insert(0, "neural-locality-sensitive-hashing_b200"); sys.path.insert(0, ".")
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
SWITCHES = ("NLSH_TC_SETS", "NLSH_TC_NQ", "NLSH_TC_QGLOBAL", "NLSH_TC_GRID", "NLSH_TC_SLOTS", "NLSH_SCAN_SEED", "NLSH_SCAN_SEED_DIV", "NLSH_TC_LADDER", "NLSH_TC_CAND_CAP", "NLSH_SCAN_IMPL",
            "NLSH_TC_STATS", "NLSH_TC_V8", "NLSH_TC_SMEM_KB")
VARIANTS = [{}] + [dict(kv.split("=") for kv in v.split(",")) for v in os.environ.get(
    "TC_VARIANTS", "NLSH_TC_LADDER=0;NLSH_SCAN_SEED=128;NLSH_SCAN_SEED=256;NLSH_SCAN_SEED=512;NLSH_TC_SLOTS=6;"
    "NLSH_SCAN_IMPL=simt").split(";")] + [{}]


def set_env(v):
    for s in SWITCHES:
        os.environ.pop(s, None)
    os.environ.update(v)


def run():
    return idx.query_tensors(Q, k=k, probes=probes)


def stats():
    os.environ["NLSH_TC_STATS"] = "1"
    _, _, nc = run(); torch.cuda.synchronize()
    ws = list(_native._workspaces.values())[0].buf
    B = 1 << hs
    st = ws.view(torch.uint8)[(2 * B + 2) * 4:(2 * B + 2) * 4 + 6 * 8].view(torch.int64).cpu().tolist()
    os.environ.pop("NLSH_TC_STATS")
    return {"pairs": int(nc.long().sum()), "survivors": st[0], "full_batches": st[1], "item_end_batches": st[2],
            "candidates": st[3], "overflow_rescans": st[5]}


base = None
for v in VARIANTS:
    set_env(v)
    for _ in range(5):
        out = run()
    torch.cuda.synchronize()
    if base is None:
        base = [t.clone() for t in out]
    same = all(torch.equal(a, b) for a, b in zip(base, out))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _native.profile_enable(True); _native.profile_read()
    e0.record()
    for _ in range(30):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = _native.profile_read(); _native.profile_enable(False)
    rec = {"variant": v, "scan_ms_mean": sum(ms) / len(ms), "scan_ms_min": min(ms), "call_ms": e0.elapsed_time(e1) / 30,
           "equal_to_default": same}
    rec.update(stats())
    print(json.dumps(rec), flush=True)
