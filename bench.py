#!/usr/bin/env python
"""Headline benchmark: QPS at recall@10 >= 0.9 of the Neural-LSH query path on B200.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this repo (CUDA hot path)
    python bench.py --impl reference --gpus 1 --steps 3 ...   # the reference's own CPU code (baseline/_ref)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the one the metric is quoted on; it fits one GPU):
synthetic 10M x 128 fp32 L2 database, 4096 buckets (hash_size 12, MultiLayerRelu 256,256
hasher fitted on the synthetic mixture), 10k queries per step, k = 10, multi-probe p = the
smallest power of two reaching recall@10 >= 0.9 against brute-force ground truth.  With N
ranks the SAME database is row-sharded (strong scaling): every rank scans its shard and the
per-shard top-k lists are merged after one NCCL all-gather.  A step = --batches-per-step (10) batches
of 10k queries, each through hash -> probe selection -> candidate scan + top-k -> (all-gather, merge),
submitted to the package's serving loop (nlsh.parallel.PipelinedSearch, --lanes batches in flight).
Prints ONE JSON line (see the key list in DESIGN.md section 5).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "neural-locality-sensitive-hashing_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: rows, dim, hash_size, queries, k, metric, seed   
    "cfg1_100k_x128_16b": (100_000, 128, 4, 1_000, 10, "l2", 1001),
    "cfg2_1m_x128_256b": (1_000_000, 128, 8, 10_000, 10, "l2", 1002),
    "cfg3_1.2m_x100_1024b_angular": (1_200_000, 100, 10, 10_000, 10, "angular", 1003),
    "cfg4_10m_x128_4096b": (10_000_000, 128, 12, 10_000, 10, "l2", 1004),
    "cfg5_1m_x960_512b_k100": (1_000_000, 960, 9, 1_000, 100, "l2", 1005),
    # config 4 with skewed bucket sizes (SURVEY section 7 "bucket skew"): every cluster bit is 1 with
    # probability 0.7, so the 4096 clusters hold between ~5 and ~138 000 of the 10M rows
    "cfg4_10m_x128_4096b_skewed": (10_000_000, 128, 12, 10_000, 10, "l2", 1004),
}
SKEW = {"cfg4_10m_x128_4096b_skewed": 0.7}  # P(cluster bit = 1); 0.5 = balanced
DEFAULT_WORKLOAD = "cfg4_10m_x128_4096b"
RECALL_TARGET = 0.9
PROBE_LADDER = (1, 2, 4, 8, 16, 32, 64, 128)
SEP = 3.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the row count (debug)")
    ap.add_argument("--queries", type=int, default=0, help="override the query count (debug)")
    ap.add_argument("--probes", type=int, default=0, help="fix p instead of searching the ladder")
    ap.add_argument("--fit-steps", type=int, default=300)
    ap.add_argument("--batches-per-step", type=int, default=10, help="query batches per timed step")
    ap.add_argument("--lanes", type=int, default=0,
                    help="batches in flight in the serving loop (0 = 2 on one GPU, 3 on 2, 4 on 4 or 8: measured best)")
    ap.add_argument("--cpu-sample", type=int, default=4096,
                    help="queries per step of the reference arm (indexer.py:45-53 multi-probes only full 4096-row batches)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the CUDA graph")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
def metric_distance(metric):
    if metric == "l2":
        return F.pairwise_distance
    fn = lambda a, b: 1 - F.cosine_similarity(a, b, dim=-1)  # noqa: E731
    fn.nlsh_metric = "angular"
    return fn


def make_hashing(d, hs, metric, seed, device, fit_steps, src_rank_trains=True, skew=0.5):
    """MultiLayerRelu(d, [256, 256]) + Linear(256, hs) + sigmoid (main.py:279-284 defaults),
    fitted on rank 0 and broadcast so every rank hashes with identical weights."""
    import synth
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    torch.manual_seed(seed)
    hashing = MultivariateBernoulli(MultiLayerRelu(d, [256, 256]), hs, metric_distance(metric))
    hashing._hasher.to(device)
    rank = dist.get_rank() if dist.is_initialized() else 0
    loss = 0.0
    if rank == 0 or not src_rank_trains:
        loss = synth.fit_hasher(hashing, d, hs, seed, device, sep=SEP, steps=fit_steps, skew=skew)
    if dist.is_initialized() and dist.get_world_size() > 1:
        for prm in hashing.parameters():
            dist.broadcast(prm.data, src=0)
    hashing.train_mode(False)
    return hashing, loss


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU sampled every 200 ms while the bench is under load (rank 0
    only: one poller per rank stalls the ranks in turn, and the captured collectives make every rank wait)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200", "-f", self.path],
                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def count(self):
        """Samples written so far."""
        try:
            return sum(1 for line in open(self.path) if line.strip())
        except (OSError, TypeError):
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8 or not parts[0].isdigit() or int(parts[0]) != self.gpu_index:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def timed_steps(step_fn, steps, barrier, finish=None):
    """EXACTLY `steps` calls bracketed by barrier + synchronize, timed with CUDA events on the
    launching stream; returns this rank's elapsed ms.  `finish` (the serving loop's fence) orders the
    launching stream after everything the steps submitted to other streams, inside the timed region."""
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_fn()
    if finish is not None:
        finish()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1)


def max_over_ranks(value, device):
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def bucket_recall(gt_ids, probes, codes_of_gt):
    """recall@k of a bucket-probing search = fraction of the true neighbours whose bucket is
    probed (every probed bucket is scanned exhaustively, so a true neighbour that is a
    candidate is always returned)."""
    hit = (codes_of_gt[:, :, None] == probes[:, None, :]).any(-1)
    return float(hit.float().mean().item())


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def cpu_query_baseline(oracle, layers, X_host, index2row, Q_host, probe_sets, metric, k, repeats=1):
    """The reference's query path restated by the oracle (Indexer.query, indexer.py:56-96, incl. query
    hashing) on the host cores - used only where baseline/_ref is absent (cpu_baseline.kind = "port")."""
    cpu = oracle.CpuIndexer.__new__(oracle.CpuIndexer)
    cpu.layers, cpu.head, cpu.metric = layers, oracle.HEAD_SIGMOID, metric
    cpu.X = X_host
    cpu._hash_codes = oracle.hash_codes_c if os.path.exists(os.path.join(ROOT, "oracle", "liboracle_pack.so")) \
        else oracle.hash_codes
    cpu.index2row = index2row
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu.hash(Q_host)  # query hashing is inside the reference's timed region
        ids, ncand = cpu.query(Q_host, k=k, probe_sets=probe_sets)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return Q_host.shape[0] / best, ids, ncand


def index2row_from_csr(offsets_host, ids_host):
    ids_t = torch.from_numpy(ids_host.astype(np.int64))
    return {int(c): ids_t[int(offsets_host[c]):int(offsets_host[c + 1])]
            for c in range(len(offsets_host) - 1) if offsets_host[c + 1] > offsets_host[c]}


def time_on_stream(fn, reps, device):
    """Mean ms of `fn()` over `reps` calls, CUDA events on the current stream (this rank only)."""
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / reps


# ------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import synth
    import precompute
    from nlsh import _native
    from nlsh.metrics import recall_at_k_tensors
    from nlsh.parallel import PipelinedSearch, ShardedIndexer, all_gather_topk, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    n, d, hs, nq, k, metric, seed = WORKLOADS[args.workload]
    skew = SKEW.get(args.workload, 0.5)
    n = args.rows or n
    nq = args.queries or nq
    nb = max(1, args.batches_per_step)
    lo, hi = shard_range(n, rank, world)
    X = synth.make_database(n, d, hs, seed, device, sep=SEP, row_lo=lo, row_hi=hi, skew=skew)
    Q = synth.make_queries(nq, d, hs, seed, device, sep=SEP, skew=skew)
    hashing, fit_loss = make_hashing(d, hs, metric, seed, device, args.fit_steps, skew=skew)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index = ShardedIndexer(hashing, X, hashing.distance, shard_lo=lo, metric=metric)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    sizes_host = index.local.bucket_sizes

    # ---- ground truth (brute force, same sharding) and the probe count for recall >= 0.9 ----
    gt_ids, gt_d = precompute.knn_tensors(Q, X, metric, k, id_offset=lo)
    if world > 1:
        g_ids, g_d = all_gather_topk(gt_ids, gt_d)
        gt_ids, gt_d = _native.merge_topk(g_d, g_ids)
    recalls = {}
    p_used = args.probes
    ladder = (args.probes,) if args.probes else tuple(p for p in PROBE_LADDER if p <= (1 << hs))
    for p in ladder:
        ids, _, _ = index.query_tensors(Q, k=k, hash_times=p)
        recalls[p] = recall_at_k_tensors(gt_ids, ids)
        p_used = p
        if recalls[p] >= RECALL_TARGET:
            break
    recall = recalls[p_used]
    ids_serial, dists_serial, _ = index.query_tensors(Q, k=k, hash_times=p_used)

    # ---- the timed step: nb batches through the serving loop, queries resident in HBM ------------
    lanes = args.lanes if args.lanes > 0 else (2 if world == 1 else (3 if world < 4 else 4))
    if args.no_graph:
        pipe = None

        def step():
            for _ in range(nb):
                index.query_tensors(Q, k=k, hash_times=p_used)
        kernels_per_batch = None
    else:
        pipe = PipelinedSearch(index, nq, k=k, hash_times=p_used, depth=lanes, to_host=False)
        kernels_per_batch = pipe.kernels_per_call

        def step():  # the lanes stay full across steps: the fence comes once, before the closing event
            for _ in range(nb):
                pipe.submit(Q)
        # the pipelined loop returns what the serial call returns
        t = pipe.submit(Q)
        p_ids, p_d, _ = pipe.result(t)
        assert torch.equal(p_ids, ids_serial) and torch.equal(p_d, dists_serial), "pipelined != serial results"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    finish = pipe.fence if pipe is not None else None
    for _ in range(max(args.warmup, 3)):
        step()
    if finish is not None:
        finish()
    launches0 = _native.kernel_launch_count()
    ms = max_over_ranks(timed_steps(step, args.steps, barrier, finish), device)
    launches = _native.kernel_launch_count() - launches0
    if kernels_per_batch is not None:  # graph replays do not pass through the counter
        launches = kernels_per_batch * nb * args.steps

    # ---- end to end: pinned host queries in, pinned host results out, every batch ----------------
    q_pinned = Q.cpu().pin_memory()
    e2e_api = (f"nlsh.parallel.PipelinedSearch ({lanes} batches in flight, one CUDA graph + stream + workspace per "
               "lane) over ShardedIndexer.capture_query; pinned host queries in, pinned host results out, every "
               "batch's results read on the host")
    pipe_h = PipelinedSearch(index, nq, k=k, hash_times=p_used, depth=lanes, to_host=True)
    pending = []

    def e2e_step():
        for _ in range(nb):
            pending.append(pipe_h.submit(q_pinned))
            if len(pending) >= lanes:
                ids_h, _, _ = pipe_h.result(pending.pop(0))  # the caller reads every result (lanes - 1 late)
                assert ids_h.shape[0] == nq

    def e2e_drain():
        while pending:
            pipe_h.result(pending.pop(0))

    for _ in range(3):
        e2e_step()
    e2e_drain()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    e2e_drain()  # the last results are read inside the timed region
    pipe_h.fence()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1), device)
    ids_h, d_h, _ = pipe_h.result(pipe_h.submit(q_pinned))
    assert torch.equal(ids_h, ids_serial.cpu()) and torch.equal(d_h, dists_serial.cpu()), "e2e != serial results"

    # ---- where a batch's time goes (this rank, serial, nothing else on the device) ----------------
    serial_run = index.capture_query(nq, k=k, hash_times=p_used) if not args.no_graph else \
        (lambda qv: index.query_tensors(qv, k=k, hash_times=p_used))
    for _ in range(3):
        serial_run(Q)
    serial_ms = time_on_stream(lambda: serial_run(Q), 30, device)
    local_ms = serial_ms
    if world > 1:  # the same batch without the exchange: this rank's shard alone (hashes all queries itself)
        local_only = index.local.capture_query(nq, k=k, hash_times=p_used)
        for _ in range(3):
            local_only(Q)
        local_ms = time_on_stream(lambda: local_only(Q), 30, device)
        barrier()

    # ---- scan-kernel roofline: CUDA events around the kernel inside the library --------------
    _native.profile_enable(True)
    prof_steps = min(max(args.steps, 3), 200)
    torch.cuda.nvtx.range_push("nlsh_steps")  # ncu --nvtx --nvtx-include "nlsh_steps/": one step's launches
    for _ in range(prof_steps):
        _, _, ncand_local = index.local.query_tensors(Q, k=k, hash_times=p_used)
    torch.cuda.nvtx.range_pop()
    scan_ms = _native.profile_read()
    # the same call where no bucket tile is shared between queries (about one probing query
    # per two buckets): every candidate byte has to come from HBM
    n_lr = max(8, min(nq, (1 << hs) // (2 * p_used)))
    for _ in range(3):
        index.local.query_tensors(Q[:n_lr], k=k, hash_times=p_used)
    _native.profile_read()
    for _ in range(50):
        _, _, ncand_lr = index.local.query_tensors(Q[:n_lr], k=k, hash_times=p_used)
    lr_ms = _native.profile_read()
    _native.profile_enable(False)
    # nvidia-smi needs about a second to deliver its first sample; the timed region is shorter than
    # that: keep the same step running, untimed, until a few samples exist (a fixed number of steps
    # derived from the rank-maximum step time, so that every rank issues the same number of all-gathers)
    for _ in range(int(min(2000, max(0.0, 1500.0 / max(ms / args.steps, 1e-3))))):
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop()

    sizes_dev = torch.from_numpy(sizes_host).to(device)
    row_bytes = 4 * d + 4  # fp32 vector + int32 id per candidate

    def distinct_bytes(queries):
        pr = index.local.hash_tensors(queries, p_used)
        return float(sizes_dev[torch.unique(pr[pr >= 0]).long()].sum().item()) * row_bytes

    lr_pairs = float(ncand_lr.double().sum().item()) * row_bytes + n_lr * (4 * d + 8 * k)
    lr_distinct = distinct_bytes(Q[:n_lr])
    lr_avg_ms = float(np.mean(lr_ms)) if lr_ms else float("nan")
    pair_bytes = float(ncand_local.double().sum().item()) * row_bytes + nq * (4 * d + 8 * k)
    distinct = distinct_bytes(Q)
    scan_avg_ms = float(np.mean(scan_ms)) if scan_ms else float("nan")
    peaks, peak_kind = load_peaks()
    achieved = distinct / (scan_avg_ms * 1e-3) / 1e9
    mean_cand = float(ncand_local.double().mean().item())

    qps = nq * nb * args.steps / (ms * 1e-3)
    e2e_qps = nq * nb * args.steps / (e2e_ms * 1e-3)
    scan_impl = _native.scan_impl(d, k, index.local._metric, index.local._x_sqnorm is not None, nq, p_used, 1 << hs)
    lr_impl = _native.scan_impl(d, k, index.local._metric, index.local._x_sqnorm is not None, n_lr, p_used, 1 << hs)
    # DRAM bytes per launch of the scan kernel from an `ncu --set full` capture of THIS library version
    # (profiles/ncu_scan_traffic.json is regenerated by scripts/gpu_profile_scan_tc.sh); null otherwise
    traffic, traffic_src = None, None
    prof_path = os.path.join(ROOT, "profiles", "ncu_scan_traffic.json")
    if os.path.exists(prof_path):
        rec = json.load(open(prof_path)).get(
            f"{args.workload}/p{p_used}/gpus{world}/{'tc' if scan_impl == 1 else 'simt'}")
        if rec and rec.get("lib_version") == _native.lib().nlsh_version():
            traffic, traffic_src = rec["dram_bytes_per_launch"], rec.get("report")
    names = {1: "scan_tc_kernel (tcgen05 tf32 filter + thread-per-survivor exact fp32 scoring)",
             0: "scan_kernel (fp32 SIMT candidate scan + top-k)"}
    batch_ms = ms / args.steps / nb
    line = {
        "metric": "QPS at recall@10>=0.9", "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "b200",
        "config": {
            "workload": args.workload, "rows": n, "dim": d, "buckets": 1 << hs, "queries_per_batch": nq,
            "batches_per_step": nb, "queries_per_step": nq * nb, "lanes": lanes if pipe is not None else 1,
            "k": k, "metric": metric, "probes": p_used, "recall_at_k": recall,
            "recall_by_probes": {str(a): b for a, b in recalls.items()},
            "mean_candidates_per_query_per_rank": mean_cand, "parallelism": f"row-shard x{world}",
            "bucket_rows_per_rank": {"mean": float(sizes_host.mean()), "max": int(sizes_host.max()),
                                     "p50": float(np.median(sizes_host)), "nonempty": int((sizes_host > 0).sum())},
            "hasher": "MultiLayerRelu(256,256)+sigmoid, fitted %d Adam steps (loss %.4f)" % (args.fit_steps, fit_loss),
            "l2_policy": "database shard (%.0f MB) larger than the 126 MB L2" % (X.numel() * 4 / 1e6),
            "index_build_s": build_s,
        },
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": nb * nq * d * 4,
                "d2h_bytes_per_step": nb * (nq * k * 12 + nq * 4), "api": e2e_api},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "breakdown_ms_per_batch": {
            "pipelined_batch": batch_ms, "serial_batch": serial_ms, "scan_kernel": scan_avg_ms,
            "rest_of_serial_batch": serial_ms - scan_avg_ms, "local_only_batch": local_ms,
            "note": "serial_batch = one captured batch alone on the device (this rank): scan_kernel + the rest (query "
                    "hashing - of this rank's 1/N slice of the queries when N > 1 -, probe selection, planning, seeding, "
                    "candidate merge, and for N > 1 the two captured NCCL all-gathers + the shard merge); local_only_batch "
                    "= the same batch on this rank's shard without any exchange (it hashes all queries itself); "
                    "pipelined_batch = ms_per_step / batches_per_step with the lanes overlapping"},
        "roofline": {"bound": "hbm", "kernel": names[scan_impl], "achieved": achieved,
                     "peak": peaks["hbm_gbs"], "peak_kind": peak_kind, "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                     "bytes_per_launch": distinct, "kernel_ms": scan_avg_ms,
                     "kernel_share_of_serial_batch": scan_avg_ms / serial_ms,
                     "pair_bytes_per_launch": pair_bytes,
                     "achieved_algorithmic": pair_bytes / (scan_avg_ms * 1e-3) / 1e9,
                     "reuse_factor": pair_bytes / distinct if distinct else None,
                     "dram_gbs": (traffic / (scan_avg_ms * 1e-3) / 1e9) if traffic else None,
                     "note": "achieved / frac count every probed bucket ONCE: (4 D + 4) bytes per row of the distinct "
                             "buckets the batch probes = what HBM has to deliver per launch; pair_bytes counts every "
                             "(query, candidate) pair as SURVEY 8d does (a tile serves reuse_factor queries, so "
                             "achieved_algorithmic exceeds any memory peak); dram_gbs = ncu dram bytes / live kernel time",
                     "hbm_bound_case": {"queries": n_lr, "probes": p_used, "kernel_ms": lr_avg_ms,
                                        "kernel": names[lr_impl].split(" ")[0],
                                        "bytes_per_launch": lr_distinct,
                                        "achieved": lr_distinct / (lr_avg_ms * 1e-3) / 1e9,
                                        "frac": lr_distinct / (lr_avg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                        "pair_bytes_per_launch": lr_pairs,
                                        "note": "same call with so few queries that buckets are (almost) never shared"}},
    }

    for obj in (pipe, pipe_h):
        if obj is not None:
            obj.release()
    g = getattr(serial_run, "graphed", None)
    if hasattr(g, "release"):
        g.release()
    del pipe, pipe_h, serial_run
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_subprocess(args, p_used)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        barrier()
        sys.stdout.flush()
        # captured NCCL work can keep destroy_process_group() waiting: the line is printed, leave without it
        os._exit(0)


def cpu_baseline_subprocess(args, p_used):
    """The reference arm (below) in a process of its own - the reference's package is called `nlsh`
    like this repo's, one process imports one of them - on a bounded sample: 2 steps of 4096 queries."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
           "--steps", "2", "--warmup", "1", "--fit-steps", str(args.fit_steps), "--cpu-sample", str(args.cpu_sample)]
    if args.rows:
        cmd += ["--rows", str(args.rows)]
    if args.queries:
        cmd += ["--queries", str(args.queries)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        rec = json.loads(lines[-1])
        base = rec["cpu_baseline"]
        base["config"] = {k: rec["config"].get(k) for k in ("queries_per_step", "probes", "hash_times", "recall_at_k",
                                                             "mean_candidates_per_query")}
        return base
    except Exception as exc:  # noqa: BLE001 - the GPU line is still valid without it
        return {"value": None, "unit": "queries/s", "cores": None, "kind": "unavailable", "sample": repr(exc)[:300]}


# ------------------------------------------------------------------------------------------
# reference arm: the reference's OWN CPU implementation of the path (baseline/_ref), host cores
# ------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_arm
    use_ref = ref_arm.available()
    if use_ref:
        ref = ref_arm.load(cpu=True)  # baseline/_ref first on sys.path: `nlsh`, `encoders` are the reference's
    import synth  # this repo's synthetic data + hasher fit (pure torch, set-up only)

    n, d, hs, nq, k, metric, seed = WORKLOADS[args.workload]
    skew = SKEW.get(args.workload, 0.5)
    n = args.rows or n
    nq = args.queries or nq
    distance = F.pairwise_distance if metric == "l2" else (lambda a, b: 1 - F.cosine_similarity(a, b, dim=-1))
    # set-up only (data, hasher fit, bucket grouping, ground truth) may use torch on the GPU when there
    # is one; none of this repo's kernels run, and everything timed is the reference's code on the CPU
    setup_dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    X = synth.make_database(n, d, hs, seed, setup_dev, sep=SEP, skew=skew)
    ns = min(args.cpu_sample, nq)
    Q = synth.make_queries(nq, d, hs, seed, setup_dev, sep=SEP, skew=skew)[:ns]
    if use_ref:
        torch.manual_seed(seed)
        hashing = ref.MultivariateBernoulli(ref.MultiLayerRelu(d, [256, 256]), hs, distance)
    else:
        from encoders import MultiLayerRelu
        from nlsh.hashings import MultivariateBernoulli
        torch.manual_seed(seed)
        hashing = MultivariateBernoulli(MultiLayerRelu(d, [256, 256]), hs, distance)
    hashing._hasher.to(setup_dev)
    fit_loss = synth.fit_hasher(hashing, d, hs, seed, setup_dev, sep=SEP, steps=args.fit_steps, skew=skew)
    hashing.train_mode(False)
    with torch.no_grad():
        enc, out = hashing._hasher._encoder, hashing._hasher.output_layer
        codes = torch.cat([((torch.sigmoid(out(enc(X[s:s + 262144]))) > 0.5).long() *
                            (1 << torch.arange(hs - 1, -1, -1, device=setup_dev))).sum(1)
                           for s in range(0, n, 262144)])
        order = torch.sort(codes, stable=True)[1]
        counts = torch.bincount(codes, minlength=1 << hs).cpu().numpy()
        offsets = np.concatenate([[0], np.cumsum(counts)])
        gt = []  # ground truth of the sample (torch matmul), to pick the probe count as the other arm does
        for s in range(0, ns, 128):
            dmat = torch.cdist(Q[s:s + 128], X) if metric == "l2" else \
                1 - F.normalize(Q[s:s + 128], dim=1) @ F.normalize(X, dim=1).T
            gt.append(dmat.topk(k, dim=1, largest=False)[1])
        gt = torch.cat(gt).cpu()
        codes_of_gt = codes.cpu()[gt]
    # bucket content = what build_index (indexer.py:6-24) produces from these codes: rows ascending per code
    index2row = index2row_from_csr(offsets, order.cpu().numpy())
    X_host, Q_host = X.cpu(), Q.cpu()
    del X, codes, order
    hashing._hasher.cpu()

    if use_ref:
        # the reference's own Indexer (indexer.py:27-96) around the grouped rows; its constructor would hash all
        # n rows through 4096-row batches and Python sets (minutes at 10M) - the query path below is unchanged
        idx = ref.Indexer.__new__(ref.Indexer)
        idx._hashing, idx._candidate_vectors_gpu, idx._distance_func = hashing, X_host, distance
        idx.index2row = index2row
        kind = "reference"
        # hash_times: the reference samples probes (hashings.py:77-81) and multi-probes only full 4096-row
        # batches (indexer.py:45-53); smallest ladder value whose probed buckets hold >= 0.9 of the true neighbours
        h_used, recalls = args.probes, {}
        ladder = (args.probes,) if args.probes else tuple(p for p in PROBE_LADDER if p <= (1 << hs))
        with torch.no_grad():
            for h in ladder:
                sets = idx.hash(Q_host, hash_times=h)
                width = max(len(sset) for sset in sets)
                probes = torch.full((ns, width), -1, dtype=torch.long)
                for i, sset in enumerate(sets):
                    probes[i, :len(sset)] = torch.tensor(sorted(sset), dtype=torch.long)
                recalls[h] = bucket_recall(gt, probes, codes_of_gt)
                h_used = h
                if recalls[h] >= RECALL_TARGET:
                    break

        probe_mode = "the reference's own sampled multi-probe (hashings.py:77-81)"
        if recalls[h_used] < RECALL_TARGET and not args.probes:
            # Bernoulli samples of a confident hasher repeat the hard code (mean distinct probes stay near 1), so
            # the reference's sampling does not reach the recall target on this workload at any ladder value.  For
            # an equal-recall comparison the deterministic top-p probe sets of the other arm are handed
            # to the reference's query loop through Indexer.hash, which still runs its own hashing pass for the cost.
            from oracle import nlsh_oracle as oracle
            with torch.no_grad():
                q_logits = hashing._hasher.output_layer(hashing._hasher._encoder(Q_host)).numpy()
            for h in ladder:
                probes = oracle.topp_probes(q_logits, oracle.HEAD_SIGMOID, h)
                recalls[f"top-p {h}"] = bucket_recall(gt, torch.from_numpy(probes).long(), codes_of_gt)
                h_used = h
                if recalls[f"top-p {h}"] >= RECALL_TARGET:
                    break
            sets = [set(int(c) for c in row if c >= 0) for row in probes]
            ref_hash = idx.hash

            def injected_hash(query_vectors, batch_size=4096, hash_times=1):
                ref_hash(query_vectors, batch_size=batch_size, hash_times=1)  # the reference's hashing cost
                return sets
            idx.hash = injected_hash
            probe_mode = ("deterministic top-p probe sets injected through Indexer.hash (the reference's sampling "
                          "stays below the recall target); hashing pass still run")

        def one_step():
            with torch.no_grad():
                return idx.query(Q_host, k=k, hash_times=h_used)  # indexer.py:56-96, includes query hashing
        mean_probes = float(np.mean([len(sset) for sset in sets]))
    else:
        from oracle import nlsh_oracle as oracle
        from nlsh.hashings import extract_layer_tensors
        kind = "port"
        with torch.no_grad():
            q_logits = out(enc(Q_host)).numpy()
        h_used, recalls = args.probes, {}
        ladder = (args.probes,) if args.probes else tuple(p for p in PROBE_LADDER if p <= (1 << hs))
        for h in ladder:
            probes = oracle.topp_probes(q_logits, oracle.HEAD_SIGMOID, h)
            recalls[h] = bucket_recall(gt, torch.from_numpy(probes).long(), codes_of_gt)
            h_used = h
            if recalls[h] >= RECALL_TARGET:
                break
        sets = [set(int(c) for c in row if c >= 0) for row in probes]
        probe_mode = "deterministic top-p probe sets"
        layers = [oracle.Layer(w.cpu(), None if b is None else b.cpu(), act == 1)
                  for w, b, act in extract_layer_tensors(enc, out)]

        def one_step():
            _, ids, ncand = cpu_query_baseline(oracle, layers, X_host, index2row, Q_host, sets, metric, k)
            return ids, ncand
        mean_probes = float(np.mean([len(sset) for sset in sets]))

    for _ in range(min(args.warmup, 1)):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids, ncand = one_step()
    wall = time.perf_counter() - t0
    qps = ns * args.steps / wall
    got_recall = float(np.mean([len(set(g) & set(r)) / k for g, r in zip(gt.tolist(), ids)]))
    cores = torch.get_num_threads()
    sample = (f"each step = {ns} of the {nq} queries against the full {n}-row index on the host cores through "
              + ("the reference's own Indexer.query (baseline/_ref/nlsh/indexer.py:56-96 + hashings.py:66-92 + the "
                 "Cython hash_codes), unmodified, `.cuda()` neutralised" if use_ref else
                 "the oracle's restatement of Indexer.query (baseline/_ref absent)")
              + f", hash_times={h_used}, includes query hashing; set-up (data, hasher fit, bucket grouping) untimed")
    line = {
        "metric": "QPS at recall@10>=0.9", "value": qps, "unit": "queries/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": args.workload, "rows": n, "dim": d, "buckets": 1 << hs,
                   "queries_per_step": ns, "k": k, "metric": metric, "probes": h_used, "hash_times": h_used,
                   "mean_distinct_probes": mean_probes, "probe_mode": probe_mode, "recall_at_k": got_recall,
                   "bucket_recall_by_hash_times": {str(a): b for a, b in recalls.items()},
                   "mean_candidates_per_query": float(np.mean(ncand)),
                   "hasher": "MultiLayerRelu(256,256)+sigmoid, fitted %d Adam steps (loss %.4f)" % (args.fit_steps, fit_loss)},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": kind, "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    _args = parse_args()
    if _args.impl == "reference":
        run_reference(_args)
    else:
        run_b200(_args)
